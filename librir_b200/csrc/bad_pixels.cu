// librir_b200/csrc/bad_pixels.cu -- bad-pixel detection (once per movie) and median correction.
//
// Reference semantics (restated in SURVEY.md appendix A.1/A.2, checked by tests against oracle/):
//   detection  rir::badPixels<unsigned short>     Filters.h:135-193, BadPixels::init BadPixels.cpp:13-32
//   correction rir::BadPixels::correct            BadPixels.cpp:34-66 + clampMin Filters.cpp:7-50
//   loader variant IRFileLoader::removeBadPixels  IRFileLoader.cpp:722-802
//
// Layout in HBM: the bad-pixel set of a movie is a BITMAP, one bit per pixel, row-padded to
// bytes (bit x&7 of mask[y*mstride + x/8]); at 640x512 that is 40 KB, shared by every frame and
// L2-resident.  Correction is one streaming pass (4 B/px algorithmic): a CTA stages a band of
// rows in shared memory, each thread re-reads its 8 pixels, patches the (rare) flagged ones with
// the 3x3 median taken from the staged INPUT band (halo rows from global), clamps, and stores.
#include "common.cuh"
#include "kernels.h"

namespace rirb {

// ------------------------------------------------------------------------------------------------
// small sorting helpers (registers only)
// ------------------------------------------------------------------------------------------------
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

// ------------------------------------------------------------------------------------------------
// a-2  correction, banded streaming kernel
// ------------------------------------------------------------------------------------------------
// 3x3 median around (x,y) as BadPixels::correct takes it: in-bounds neighbours incl. centre, all
// read from the INPUT frame; band rows come from shared memory, the two halo rows from global.
__device__ __forceinline__ unsigned median3x3_input(const u16* __restrict__ frame, const u16* tile, int w, int h, int y0,
                                                    int y1, int x, int y)
{
    unsigned v[9];
    int c = 0;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
        int yy = y + dy;
        bool rowok = (yy >= 0) && (yy < h);
        bool in_tile = (yy >= y0) && (yy < y1);
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            int xx = x + dx;
            bool ok = rowok && (xx >= 0) && (xx < w);
            unsigned val = 0xFFFFFFFFu;
            if (ok) {
                val = in_tile ? (unsigned)tile[(yy - y0) * w + xx] : (unsigned)frame[(size_t)yy * w + xx];
                ++c;
            }
            v[(dy + 1) * 3 + (dx + 1)] = val;
        }
    }
    sort9(v);
    return pick_mid(v, c);
}

constexpr int BP_THREADS = 256;

// VEC == 8: w % 8 == 0 and 16-byte aligned frames (128-bit path).  VEC == 1: anything else.
template <int VEC>
__global__ void __launch_bounds__(BP_THREADS)
bp_correct_kernel(const u16* __restrict__ in, u16* __restrict__ out, const u8* __restrict__ mask, int w, int h, int mstride,
                  int band_rows, int bands, unsigned clamp, size_t frame_stride)
{
    extern __shared__ __align__(16) u16 tile[];
    const int band = blockIdx.x % bands;
    const size_t f = blockIdx.x / bands;
    const int y0 = band * band_rows;
    const int y1 = min(h, y0 + band_rows);
    const u16* frame = in + f * frame_stride;
    u16* oframe = out + f * frame_stride;
    const int npx = (y1 - y0) * w;

    if (VEC == 8) {
        const uint4* g = reinterpret_cast<const uint4*>(frame + (size_t)y0 * w);
        uint4* s = reinterpret_cast<uint4*>(tile);
        const int nvec = npx >> 3;
        // phase 1: band -> shared memory, loads issued in independent batches of 4
        for (int i = threadIdx.x; i < nvec; i += 4 * BP_THREADS) {
            uint4 r[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int j = i + k * BP_THREADS;
                if (j < nvec) r[k] = ld_stream(g + j);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int j = i + k * BP_THREADS;
                if (j < nvec) s[j] = r[k];
            }
        }
        __syncthreads();
        // phase 2: patch flagged pixels, clamp, store
        const unsigned c2 = clamp | (clamp << 16);
        uint4* o = reinterpret_cast<uint4*>(oframe + (size_t)y0 * w);
        const int vec_per_row = w >> 3;
        for (int j = threadIdx.x; j < nvec; j += BP_THREADS) {
            uint4 v = s[j];
            int ry = j / vec_per_row;
            int xv = j - ry * vec_per_row;
            int y = y0 + ry;
            unsigned m = mask[(size_t)y * mstride + xv];
            if (m) {
                unsigned px[8] = {v.x & 0xFFFF, v.x >> 16, v.y & 0xFFFF, v.y >> 16, v.z & 0xFFFF, v.z >> 16, v.w & 0xFFFF, v.w >> 16};
#pragma unroll
                for (int b = 0; b < 8; ++b)
                    if (m & (1u << b)) px[b] = median3x3_input(frame, tile, w, h, y0, y1, xv * 8 + b, y);
                v.x = px[0] | (px[1] << 16);
                v.y = px[2] | (px[3] << 16);
                v.z = px[4] | (px[5] << 16);
                v.w = px[6] | (px[7] << 16);
            }
            v.x = vmaxu2(v.x, c2);
            v.y = vmaxu2(v.y, c2);
            v.z = vmaxu2(v.z, c2);
            v.w = vmaxu2(v.w, c2);
            st_stream(o + j, v);
        }
    } else {
        const u16* g = frame + (size_t)y0 * w;
        for (int i = threadIdx.x; i < npx; i += BP_THREADS) tile[i] = g[i];
        __syncthreads();
        u16* o = oframe + (size_t)y0 * w;
        for (int i = threadIdx.x; i < npx; i += BP_THREADS) {
            int ry = i / w;
            int x = i - ry * w;
            int y = y0 + ry;
            unsigned v = tile[i];
            if (mask[(size_t)y * mstride + (x >> 3)] & (1u << (x & 7))) v = median3x3_input(frame, tile, w, h, y0, y1, x, y);
            o[i] = (u16)max(v, clamp);
        }
    }
}

int launch_bp_correct(const u16* in, u16* out, const u8* mask, int w, int h, int clamp_value, long long nframes,
                      size_t frame_stride, cudaStream_t st)
{
    if (nframes <= 0 || w <= 0 || h <= 0) return 0;
    const int mstride = (w + 7) / 8;
    const unsigned clamp = clamp_value > 0 ? (unsigned)clamp_value & 0xFFFFu : 0u;
    // band: <= 24 KB of shared memory and <= 2048 vectors, so >= 8 CTAs fit on an SM
    int band_rows = (int)(12288 / w);
    if (band_rows < 1) band_rows = 1;
    if (band_rows > 32) band_rows = 32;
    if (band_rows > h) band_rows = h;
    const int bands = (int)ceil_div(h, band_rows);
    const size_t smem = (size_t)band_rows * w * sizeof(u16);
    if (smem > 200 * 1024) {
        set_error("bad_pixels_correct: image width %d too large", w);
        return -1;
    }
    const long long grid = nframes * bands;
    if (grid > 0x7FFFFFFFLL) {
        set_error("bad_pixels_correct: too many frames in one call (%lld)", nframes);
        return -1;
    }
    const bool vec = (w % 8 == 0) && aligned16(in) && aligned16(out) && (frame_stride % 8 == 0);
    if (vec) {
        if (smem > 48 * 1024)
            RIRB_CUDA_OK(cudaFuncSetAttribute(bp_correct_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RIRB_LAUNCH(bp_correct_kernel<8>, (unsigned)grid, BP_THREADS, smem, st, in, out, mask, w, h, mstride, band_rows, bands,
                    clamp, frame_stride);
    } else {
        if (smem > 48 * 1024)
            RIRB_CUDA_OK(cudaFuncSetAttribute(bp_correct_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RIRB_LAUNCH(bp_correct_kernel<1>, (unsigned)grid, BP_THREADS, smem, st, in, out, mask, w, h, mstride, band_rows, bands,
                    clamp, frame_stride);
    }
    return 0;
}

// In-place correction keeps the reference's sequential meaning (BadPixels.cpp:41-59 with in == out:
// later pixels see earlier corrections).  One thread walks one frame's raster-ordered list; rare
// path (the Python API never calls it), parallel over frames only.
__global__ void bp_correct_inplace_kernel(u16* img, const int* __restrict__ xy, int count, int w, int h, long long nframes,
                                          size_t frame_stride)
{
    long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    u16* frame = img + (size_t)f * frame_stride;
    for (int i = 0; i < count; ++i) {
        int x = xy[2 * i], y = xy[2 * i + 1];
        unsigned v[9];
        int c = 0;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            int xx = x + k / 3 - 1, yy = y + k % 3 - 1;
            bool ok = xx >= 0 && yy >= 0 && xx < w && yy < h;
            v[k] = ok ? (unsigned)frame[(size_t)yy * w + xx] : 0xFFFFFFFFu;
            c += ok;
        }
        sort9(v);
        frame[(size_t)y * w + x] = (u16)pick_mid(v, c);
    }
}

__global__ void clamp_min_kernel(u16* img, size_t n, unsigned clamp)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) img[i] = (u16)max((unsigned)img[i], clamp);
}

int launch_bp_correct_inplace(u16* img, const int* xy_dev, int count, int w, int h, int clamp_value, long long nframes,
                              size_t frame_stride, cudaStream_t st)
{
    if (nframes <= 0) return 0;
    RIRB_LAUNCH(bp_correct_inplace_kernel, (unsigned)ceil_div(nframes, 64), 64, 0, st, img, xy_dev, count, w, h, nframes,
                frame_stride);
    if (clamp_value > 0) {
        // frames may be strided: clamp each frame's w*h pixels; contiguous movies take one launch
        if (frame_stride == (size_t)w * h) {
            size_t n = (size_t)nframes * frame_stride;
            RIRB_LAUNCH(clamp_min_kernel, (unsigned)min((long long)ceil_div((long long)n, 256), (long long)sm_count() * 32), 256, 0,
                        st, img, n, (unsigned)clamp_value & 0xFFFFu);
        } else {
            for (long long f = 0; f < nframes; ++f)
                RIRB_LAUNCH(clamp_min_kernel, (unsigned)ceil_div((long long)w * h, 256), 256, 0, st, img + f * frame_stride,
                            (size_t)w * h, (unsigned)clamp_value & 0xFFFFu);
        }
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// a-3  loader variant: window shifted inside the image, flagged cells skipped, in place, no clamp
// ------------------------------------------------------------------------------------------------
// Only flagged pixels are written and only unflagged pixels are read, so the in-place update is
// order-independent (the reference's sequential loop reads the same values): one thread per
// (frame, flagged pixel) off the raster-ordered list.
__global__ void loader_bp_kernel(u16* img, const int* __restrict__ xy, const u8* __restrict__ mask, int count, int w, int h,
                                 int mstride, long long nframes, size_t frame_stride)
{
    long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= nframes * count) return;
    long long f = id / count;
    int i = (int)(id - f * count);
    u16* frame = img + (size_t)f * frame_stride;
    int x = xy[2 * i], y = xy[2 * i + 1];
    int x0 = x - 1, y0 = y - 1;
    if (x == 0) x0 = 0; else if (x == w - 1) x0 = w - 3;
    if (y == 0) y0 = 0; else if (y == h - 1) y0 = h - 3;
    unsigned v[9];
    int c = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        int xx = x0 + k / 3, yy = y0 + k % 3;
        bool ok = !(mask[(size_t)yy * mstride + (xx >> 3)] & (1u << (xx & 7)));
        v[k] = ok ? (unsigned)frame[(size_t)yy * w + xx] : 0xFFFFFFFFu;
        c += ok;
    }
    if (c == 0) return;  // reference reads a stale stack slot here (undefined); leave the pixel
    sort9(v);
    frame[(size_t)y * w + x] = (u16)pick_mid(v, c);
}

// w < 3 or h < 3: the reference falls back to the sequential clipped-window loop
// (IRFileLoader.cpp:735-752), which reads earlier corrections: same walk as the in-place kernel.
int launch_loader_bp(u16* img, const int* xy_dev, const u8* mask, int count, int w, int h, long long nframes,
                     size_t frame_stride, cudaStream_t st)
{
    if (nframes <= 0 || count <= 0) return 0;
    if (w < 3 || h < 3) {
        RIRB_LAUNCH(bp_correct_inplace_kernel, (unsigned)ceil_div(nframes, 64), 64, 0, st, img, xy_dev, count, w, h, nframes,
                    frame_stride);
        return 0;
    }
    long long total = nframes * count;
    RIRB_LAUNCH(loader_bp_kernel, (unsigned)ceil_div(total, 256), 256, 0, st, img, xy_dev, mask, count, w, h, (w + 7) / 8,
                nframes, frame_stride);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// a-1  detection
// ------------------------------------------------------------------------------------------------
__global__ void hist_frame_kernel(const u16* __restrict__ img, size_t n, unsigned* __restrict__ hist)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) atomicAdd(&hist[img[i]], 1u);
}

int launch_hist_frame(const u16* img, size_t n, unsigned* hist65536, cudaStream_t st)
{
    RIRB_CUDA_OK(cudaMemsetAsync(hist65536, 0, 65536 * sizeof(unsigned), st));
    if (n == 0) return 0;
    RIRB_LAUNCH(hist_frame_kernel, (unsigned)min((long long)ceil_div((long long)n, 256), 4096LL), 256, 0, st, img, n, hist65536);
    return 0;
}

// One thread per pixel: in-bounds 5x5 window, full sort (odd-even transposition on 25 registers,
// sentinels pad short windows), median w[n/2], trimmed spread over sorted [n/5, 4n/5), then the
// three tests of Filters.h:184-188 in non-contracted fp64 (the flagged set is a discrete decision).
__global__ void __launch_bounds__(256)
bp_detect_kernel(const u16* __restrict__ img, int w, int h, double std_factor, unsigned gthr, u8* __restrict__ mask, int mstride)
{
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    bool bad = false;
    if (x < w && y < h) {
        unsigned s[25];
        int n = 0;
#pragma unroll
        for (int dy = -2; dy <= 2; ++dy)
#pragma unroll
            for (int dx = -2; dx <= 2; ++dx) {
                int xx = x + dx, yy = y + dy;
                bool ok = xx >= 0 && yy >= 0 && xx < w && yy < h;
                s[(dy + 2) * 5 + dx + 2] = ok ? (unsigned)img[(size_t)yy * w + xx] : 0xFFFFFFFFu;
                n += ok;
            }
#pragma unroll
        for (int round = 0; round < 25; ++round) {
#pragma unroll
            for (int i = (round & 1); i + 1 < 25; i += 2) cswap(s[i], s[i + 1]);
        }
        const int imed = n / 2, lo = n / 5, hi = n * 4 / 5;
        long long med = 0;
#pragma unroll
        for (int i = 0; i < 25; ++i)
            if (i == imed) med = (long long)s[i];
        long long acc = 0;
#pragma unroll
        for (int i = 0; i < 25; ++i)
            if (i >= lo && i < hi) {
                long long d = (long long)s[i] - med;
                acc += d * d;
            }
        double sum2 = __ddiv_rn((double)acc, (double)(hi - lo));
        double sd = __dsqrt_rn(sum2);
        double spread = __dmul_rn(std_factor, sd);
        double lower = __dsub_rn((double)med, spread);
        double upper = __dadd_rn((double)med, spread);
        unsigned p = img[(size_t)y * w + x];
        bad = ((double)p < lower) || ((double)p > upper) || (p < gthr);
    }
    unsigned bits = __ballot_sync(0xFFFFFFFFu, bad);
    if (threadIdx.x < 4 && y < h) {
        int byte = blockIdx.x * 4 + threadIdx.x;
        if (byte < mstride) mask[(size_t)y * mstride + byte] = (u8)(bits >> (8 * threadIdx.x));
    }
}

int launch_bp_detect(const u16* img, int w, int h, double std_factor, unsigned gthr, u8* mask, cudaStream_t st)
{
    dim3 block(32, 8);
    dim3 grid((unsigned)ceil_div(w, 32), (unsigned)ceil_div(h, 8));
    RIRB_LAUNCH(bp_detect_kernel, grid, block, 0, st, img, w, h, std_factor, gthr, mask, (w + 7) / 8);
    return 0;
}

}  // namespace rirb
