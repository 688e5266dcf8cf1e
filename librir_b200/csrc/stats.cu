// librir_b200/csrc/stats.cu -- per-movie statistics: min, max and the exact 65,536-bin histogram.
//
// These are the quantities the multi-GPU path all-reduces and from which the
// reference's per-image statistics follow exactly:
//   find_median_pixel(_mask)  Filters.cpp:56-101   (quantile of the 65,535-bin histogram)
//   get_background            h264.cpp:1955-1991   (mode of the 16,384-bin histogram of p>>2)
//   frame min / max           h264.cpp:2093-2097, masked_registration_ecc.py:157-163
//
// Kernel: persistent, one CTA per SM.  Values below HIST_SMEM_BINS (49,152 -- every 13/14-bit IR
// camera) are counted in a shared-memory histogram of u32 (192 KB of the SM's 227 KB); the rest go
// straight to the global u64 histogram.  At the end each CTA adds its non-zero bins to the global
// histogram.  Algorithmic traffic: 2 B/px read.
#include "common.cuh"
#include "hist.cuh"
#include "kernels.h"

namespace rirb {

constexpr int ST_THREADS = 1024;

__global__ void stats_init_kernel(unsigned* minmax, unsigned long long* hist)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && minmax) {
        minmax[0] = 65535u;
        minmax[1] = 0u;
    }
    if (hist && i < 65536) hist[i] = 0ull;
}

int launch_stats_init(unsigned* minmax, unsigned long long* hist, cudaStream_t st)
{
    RIRB_LAUNCH(stats_init_kernel, 65536 / 256, 256, 0, st, minmax, hist);
    return 0;
}

template <bool HIST>
__global__ void __launch_bounds__(ST_THREADS, 1)
movie_stats_kernel(const u16* __restrict__ p, size_t n, const u8* __restrict__ mask, unsigned* __restrict__ minmax,
                   unsigned long long* __restrict__ hist)
{
    extern __shared__ unsigned sh[];
    if (HIST) {
        for (unsigned i = threadIdx.x; i < HIST_SMEM_BINS; i += ST_THREADS) sh[i] = 0;
        __syncthreads();
    }
    unsigned lo = 65535u, hi = 0u;
    const size_t tid = (size_t)blockIdx.x * ST_THREADS + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * ST_THREADS;

    if (mask == nullptr && (reinterpret_cast<uintptr_t>(p) & 15u) == 0) {
        const size_t nvec = n >> 3;
        const uint4* pv = reinterpret_cast<const uint4*>(p);
        unsigned lo2 = 0xFFFFFFFFu, hi2 = 0u;  // packed per-halfword min / max
        constexpr int U = 4;                   // 128-bit loads in flight per thread
        for (size_t i = tid; i < nvec; i += U * nthreads) {
            uint4 a[U];
#pragma unroll
            for (int q = 0; q < U; ++q) {
                const size_t j = i + q * nthreads;
                a[q] = (j < nvec) ? ld_stream(pv + j) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int q = 0; q < U; ++q) {
                if (i + q * nthreads >= nvec) break;
                const unsigned w[4] = {a[q].x, a[q].y, a[q].z, a[q].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    lo2 = __vminu2(lo2, w[k]);
                    hi2 = __vmaxu2(hi2, w[k]);
                }
                if (HIST) count_words(w, sh, hist);
            }
        }
        lo = min(lo2 & 0xFFFFu, lo2 >> 16);
        hi = max(hi2 & 0xFFFFu, hi2 >> 16);
        for (size_t i = (nvec << 3) + tid; i < n; i += nthreads) {  // tail
            const unsigned v = p[i];
            lo = min(lo, v);
            hi = max(hi, v);
            if (HIST) count_px(v, sh, hist);
        }
    } else {
        for (size_t i = tid; i < n; i += nthreads) {
            if (mask && !mask[i]) continue;
            const unsigned v = p[i];
            lo = min(lo, v);
            hi = max(hi, v);
            if (HIST) count_px(v, sh, hist);
        }
    }
    hist_flush(lo, hi, sh, minmax, hist, HIST);
}

int launch_movie_stats(const u16* mov, size_t n, const u8* mask, unsigned* minmax, unsigned long long* hist, cudaStream_t st)
{
    if (n == 0) return 0;
    long long grid = sm_count();
    const long long need = ceil_div((long long)n, (long long)ST_THREADS * 16);
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    if (hist) {
        const size_t smem = HIST_SMEM_BINS * sizeof(unsigned);
        RIRB_SMEM_ATTR(movie_stats_kernel<true>, smem);
        RIRB_LAUNCH(movie_stats_kernel<true>, (unsigned)grid, ST_THREADS, smem, st, mov, n, mask, minmax, hist);
    } else {
        RIRB_LAUNCH(movie_stats_kernel<false>, (unsigned)(grid * 2), ST_THREADS, 0, st, mov, n, mask, minmax, hist);
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// small single-CTA reductions over the 65,536-bin histogram (results stay on the device)
// ------------------------------------------------------------------------------------------------
// find_median_pixel rule (Filters.cpp:63-71): s = round(count * percent) evaluated in float, answer
// = first bin i < 65535 whose cumulative count reaches s, 0 if none.  count < 0: use the
// histogram's own total (the masked variant, Filters.cpp:92).
__global__ void __launch_bounds__(1024)
hist_quantile_kernel(const unsigned long long* __restrict__ hist, long long count, float percent, int masked_rule,
                     int* __restrict__ out)
{
    __shared__ unsigned long long part[1024];
    __shared__ unsigned long long total_sh;
    __shared__ int best;
    const int t = threadIdx.x;
    unsigned long long local = 0;
    for (int i = 0; i < 64; ++i) local += hist[t * 64 + i];
    part[t] = local;
    if (t == 0) best = 0x7FFFFFFF;
    __syncthreads();
    if (t == 0) {  // exclusive scan of 1024 partials; serial is fine for a once-per-call reduction
        unsigned long long run = 0;
        for (int i = 0; i < 1024; ++i) {
            unsigned long long v = part[i];
            part[i] = run;
            run += v;
        }
        total_sh = run;
    }
    __syncthreads();
    const unsigned long long total = (count < 0) ? total_sh : (unsigned long long)count;
    const float target = roundf((float)total * percent);
    // (size_t)round(...) for the plain rule, (size_t)(int)round(...) for the masked one
    const unsigned long long s = masked_rule ? (unsigned long long)(long long)(int)target : (unsigned long long)target;
    unsigned long long run = part[t];
    for (int i = 0; i < 64; ++i) {
        const int bin = t * 64 + i;
        if (bin >= 65535) break;
        run += hist[bin];
        if (run >= s) {
            atomicMin(&best, bin);
            break;
        }
    }
    __syncthreads();
    if (t == 0) *out = (best == 0x7FFFFFFF) ? 0 : best;
}

int launch_hist_quantile(const unsigned long long* hist, long long count, float percent, int masked_rule, int* out,
                         cudaStream_t st)
{
    RIRB_LAUNCH(hist_quantile_kernel, 1, 1024, 0, st, hist, count, percent, masked_rule, out);
    return 0;
}

// get_background rule (h264.cpp:1955-1991): mode of the 16,384-bin histogram of p>>2, the first
// maximum wins, answer (bin<<2)+1.
__global__ void __launch_bounds__(1024)
hist_mode4_kernel(const unsigned long long* __restrict__ hist, unsigned* __restrict__ out)
{
    __shared__ unsigned long long bestv[1024];
    __shared__ int besti[1024];
    const int t = threadIdx.x;
    unsigned long long bv = 0;
    int bi = 0x7FFFFFFF;
    for (int i = 0; i < 16; ++i) {
        const int bin = t * 16 + i;
        const unsigned long long c = hist[4 * bin] + hist[4 * bin + 1] + hist[4 * bin + 2] + hist[4 * bin + 3];
        if (bi == 0x7FFFFFFF || c > bv) {
            bv = c;
            bi = bin;
        }
    }
    bestv[t] = bv;
    besti[t] = bi;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if (t < s) {
            // keep the larger count; on ties the lower bin (the reference scans upward with '>')
            if (bestv[t + s] > bestv[t] || (bestv[t + s] == bestv[t] && besti[t + s] < besti[t])) {
                bestv[t] = bestv[t + s];
                besti[t] = besti[t + s];
            }
        }
        __syncthreads();
    }
    if (t == 0) *out = ((unsigned)besti[0] << 2) + 1u;
}

int launch_hist_mode4(const unsigned long long* hist, unsigned* out, cudaStream_t st)
{
    RIRB_LAUNCH(hist_mode4_kernel, 1, 1024, 0, st, hist, out);
    return 0;
}

}  // namespace rirb
