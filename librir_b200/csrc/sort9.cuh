// librir_b200/csrc/sort9.cuh -- register sorting network shared by the bad-pixel kernels.
#pragma once

namespace rirb {

__device__ __forceinline__ void cswap(unsigned& a, unsigned& b)
{
    unsigned lo = min(a, b), hi = max(a, b);
    a = lo;
    b = hi;
}

// Sort 9 values ascending (25-exchange, depth-7 network; checked with the 0/1 principle) -- sentinels 0xFFFFFFFF pad short windows.
__device__ __forceinline__ void sort9(unsigned (&v)[9])
{
    cswap(v[0], v[3]); cswap(v[1], v[7]); cswap(v[2], v[5]); cswap(v[4], v[8]);
    cswap(v[0], v[7]); cswap(v[2], v[4]); cswap(v[3], v[8]); cswap(v[5], v[6]);
    cswap(v[0], v[2]); cswap(v[1], v[3]); cswap(v[4], v[5]); cswap(v[7], v[8]);
    cswap(v[1], v[4]); cswap(v[3], v[6]); cswap(v[5], v[7]);
    cswap(v[0], v[1]); cswap(v[2], v[4]); cswap(v[3], v[5]); cswap(v[6], v[8]);
    cswap(v[2], v[3]); cswap(v[4], v[5]); cswap(v[6], v[7]);
    cswap(v[1], v[2]); cswap(v[3], v[4]); cswap(v[5], v[6]);
}

// element c/2 of the c valid (smallest) entries of a sorted 9-array, c in [1,9]
__device__ __forceinline__ unsigned pick_mid(const unsigned (&v)[9], int c)
{
    int k = c >> 1;  // 0..4
    unsigned r = v[0];
    r = (k == 1) ? v[1] : r;
    r = (k == 2) ? v[2] : r;
    r = (k == 3) ? v[3] : r;
    r = (k == 4) ? v[4] : r;
    return r;
}


}  // namespace rirb
