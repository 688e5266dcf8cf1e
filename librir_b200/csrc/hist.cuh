// librir_b200/csrc/hist.cuh -- the shared-memory histogram shared by stats.cu and the fused pre-coder.
// Values below HIST_SMEM_BINS (49,152 -- every 13/14-bit IR camera) are counted in a shared-memory
// histogram of u32 (192 KB of the SM's 227 KB); the rest go straight to the global u64 histogram.
#pragma once
#include "common.cuh"

namespace rirb {

constexpr unsigned HIST_SMEM_BINS = 49152;

__device__ __forceinline__ void count_px(unsigned v, unsigned* sh, unsigned long long* hist)
{
    if (v < HIST_SMEM_BINS)
        atomicAdd(&sh[v], 1u);
    else
        atomicAdd(&hist[v], 1ull);
}

// per-CTA epilogue: packed per-halfword min / max of the thread -> one global atomic pair per warp;
// non-zero shared bins -> global histogram
__device__ __forceinline__ void hist_flush(unsigned lo, unsigned hi, const unsigned* sh, unsigned* __restrict__ minmax,
                                           unsigned long long* __restrict__ hist, bool with_hist)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xFFFFFFFFu, hi, o));
    }
    if ((threadIdx.x & 31) == 0 && minmax && lo <= hi) {  // lo > hi: this warp saw no pixel
        atomicMin(&minmax[0], lo);
        atomicMax(&minmax[1], hi);
    }
    if (with_hist) {
        __syncthreads();
        for (unsigned i = threadIdx.x; i < HIST_SMEM_BINS; i += blockDim.x) {
            const unsigned c = sh[i];
            if (c) atomicAdd(&hist[i], (unsigned long long)c);
        }
    }
}

}  // namespace rirb
