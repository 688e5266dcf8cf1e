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


// BadPixels::correct's replacement value for pixel (x, y): the (c/2)-th smallest of the c in-bounds cells of its 3x3
// neighbourhood, read from the UNCORRECTED frame (BadPixels.cpp:41-59).
__device__ __forceinline__ unsigned median3x3_global(const u16* __restrict__ frame, int w, int h, int x, int y)
{
    unsigned v[9];
    if (x > 0 && y > 0 && x < w - 1 && y < h - 1) {  // interior: 9 unconditional loads
        const u16* p = frame + (size_t)(y - 1) * w + (x - 1);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) v[r * 3 + c] = p[(size_t)r * w + c];
        sort9(v);
        return v[4];
    }
    int c = 0;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const int xx = x + dx, yy = y + dy;
            const bool ok = xx >= 0 && yy >= 0 && xx < w && yy < h;
            v[(dy + 1) * 3 + dx + 1] = ok ? (unsigned)frame[(size_t)yy * w + xx] : 0xFFFFFFFFu;
            c += ok;
        }
    sort9(v);
    return pick_mid(v, c);
}

}  // namespace rirb
