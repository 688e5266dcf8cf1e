// librir_b200/csrc/bad_pixels.cu -- bad-pixel detection (once per movie) and median correction.
//
// Reference semantics (restated in SURVEY.md appendix A.1/A.2, checked by tests against oracle/):
//   detection  rir::badPixels<unsigned short>     Filters.h:135-193, BadPixels::init BadPixels.cpp:13-32
//   correction rir::BadPixels::correct            BadPixels.cpp:34-66 + clampMin Filters.cpp:7-50
//   loader variant IRFileLoader::removeBadPixels  IRFileLoader.cpp:722-802
//
// Layout in HBM: the bad-pixel set of a movie is a BITMAP, one bit per pixel, row-padded to
// bytes (bit x&7 of mask[y*mstride + x/8]); at 640x512 that is 40 KB, shared by every frame and
// L2-resident, plus the raster-ordered (x,y) list the reference keeps.  Correction is one
// streaming pass (4 B/px algorithmic): copy + clamp at full width, then one thread per flagged
// pixel takes the 3x3 median from the INPUT frame and overwrites the pixel.
#include "common.cuh"
#include "kernels.h"
#include "sort9.cuh"

namespace rirb {

// ------------------------------------------------------------------------------------------------
// a-2  correction: streaming copy+clamp, then list-driven fix-ups
// ------------------------------------------------------------------------------------------------
// A CTA owns a flat span of BP_SPAN pixels of one frame.
//   pass 1 (every byte of the movie, once): 256-bit streaming load -> per-halfword max with the
//           clamp -> 256-bit streaming store.  ~1 instruction per 16 bytes moved.
//   pass 2 (K flagged pixels per frame, K/N ~ 1 %): the raster-ordered list of flagged pixels is
//           cut per span at create time (span_off); ONE THREAD PER FLAGGED PIXEL gathers the
//           in-bounds 3x3 neighbourhood from the INPUT frame (L2 hits: the CTA has just streamed
//           those rows), sorts 9 registers, and overwrites the pixel.  No divergence tax on the
//           clean pixels, which is what made the band-in-shared-memory version ALU-bound
//           (profiles/r1_v1_ncu_summary.md: 64 % ALU at 29 % of HBM peak).
// The two passes touch the same output address from different threads: __syncthreads orders them.
constexpr int BP_THREADS = 256;
constexpr int BP_UNROLL = BP_SPAN / (BP_THREADS * 16);  // 256-bit vectors per thread
static_assert(BP_UNROLL * BP_THREADS * 16 == BP_SPAN, "BP_SPAN must be a multiple of one CTA-wide row of vectors");

// VEC: frames 32-byte aligned (base and stride), so span starts are too.
template <bool VEC>
__global__ void __launch_bounds__(BP_THREADS)
bp_correct_kernel(const u16* __restrict__ in, u16* __restrict__ out, const int* __restrict__ xy, const int* __restrict__ span_off,
                  int w, int h, int npx, int spans, unsigned clamp, size_t frame_stride)
{
    const int span = blockIdx.x % spans;
    const size_t f = blockIdx.x / spans;
    const int s0 = span * BP_SPAN;
    const int s1 = min(npx, s0 + BP_SPAN);
    const u16* frame = in + f * frame_stride;
    u16* oframe = out + f * frame_stride;
    int done = s0;
    const int a = span_off[span], b = span_off[span + 1];
    U32x8 r[BP_UNROLL];
    const int nvec = VEC ? (s1 - s0) >> 4 : 0;
    if (VEC) {
        const U32x8* g = reinterpret_cast<const U32x8*>(frame + s0);
#pragma unroll
        for (int k = 0; k < BP_UNROLL; ++k) {
            const int j = threadIdx.x + k * BP_THREADS;
            if (j < nvec) r[k] = ld_stream256_keep(g + j);
        }
    }
    // first fix-up of this thread: its gather is issued while the stream loads are still in flight,
    // so the neighbour rows are found in L2 (behind the same misses) instead of being re-read from HBM
    const int i0 = a + (int)threadIdx.x;
    int2 p0 = make_int2(0, 0);
    unsigned med0 = 0;
    if (i0 < b) {
        p0 = reinterpret_cast<const int2*>(xy)[i0];
        med0 = median3x3_global(frame, w, h, p0.x, p0.y);
    }
    if (VEC) {
        const unsigned c2 = clamp | (clamp << 16);
        U32x8* o = reinterpret_cast<U32x8*>(oframe + s0);
#pragma unroll
        for (int k = 0; k < BP_UNROLL; ++k) {
            const int j = threadIdx.x + k * BP_THREADS;
            if (j < nvec) {
#pragma unroll
                for (int q = 0; q < 8; ++q) r[k].v[q] = vmaxu2(r[k].v[q], c2);
                st_stream256(o + j, r[k]);
            }
        }
        done = s0 + (nvec << 4);
    }
    for (int i = done + threadIdx.x; i < s1; i += BP_THREADS) oframe[i] = (u16)max((unsigned)frame[i], clamp);
    if (a == b) return;  // CTA-uniform
    __syncthreads();
    if (i0 < b) oframe[(size_t)p0.y * w + p0.x] = (u16)max(med0, clamp);
    for (int i = i0 + BP_THREADS; i < b; i += BP_THREADS) {
        const int2 p = reinterpret_cast<const int2*>(xy)[i];
        oframe[(size_t)p.y * w + p.x] = (u16)max(median3x3_global(frame, w, h, p.x, p.y), clamp);
    }
}

int launch_bp_correct(const u16* in, u16* out, const int* xy_dev, const int* span_off_dev, int w, int h, int clamp_value,
                      long long nframes, size_t frame_stride, cudaStream_t st)
{
    if (nframes <= 0 || w <= 0 || h <= 0) return 0;
    const unsigned clamp = clamp_value > 0 ? (unsigned)clamp_value & 0xFFFFu : 0u;
    const long long npx = (long long)w * h;
    const long long spans = ceil_div(npx, BP_SPAN);
    const long long grid = nframes * spans;
    if (npx > 0x7FFFFFFFLL || grid > 0x7FFFFFFFLL) {
        set_error("bad_pixels_correct: too many pixels or frames in one call (%lld frames)", nframes);
        return -1;
    }
    const bool vec = aligned32(in) && aligned32(out) && (frame_stride % 16 == 0);
    if (vec)
        RIRB_LAUNCH(bp_correct_kernel<true>, (unsigned)grid, BP_THREADS, 0, st, in, out, xy_dev, span_off_dev, w, h, (int)npx,
                    (int)spans, clamp, frame_stride);
    else
        RIRB_LAUNCH(bp_correct_kernel<false>, (unsigned)grid, BP_THREADS, 0, st, in, out, xy_dev, span_off_dev, w, h, (int)npx,
                    (int)spans, clamp, frame_stride);
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
