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

// N packed words (2 pixels each): ONE range test for all of them -- a value is >= 49,152 exactly when its two top bits
// are set -- then unconditional shared-memory atomics; only a vector that really holds such a value takes the per-pixel
// test.  (The per-pixel test cost 2 of the ~9.6 instructions per pixel of an ALU-bound kernel.)
template <int N> __device__ __forceinline__ void count_words(const unsigned (&w)[N], unsigned* sh, unsigned long long* hist)
{
    unsigned top = 0;
#pragma unroll
    for (int k = 0; k < N; ++k) top |= w[k] & (w[k] << 1);
    if ((top & 0x80008000u) == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            atomicAdd(&sh[w[k] & 0xFFFFu], 1u);
            atomicAdd(&sh[w[k] >> 16], 1u);
        }
    } else {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            count_px(w[k] & 0xFFFFu, sh, hist);
            count_px(w[k] >> 16, sh, hist);
        }
    }
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
