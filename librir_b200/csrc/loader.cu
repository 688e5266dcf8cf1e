// librir_b200/csrc/loader.cu -- the reader's post-decode chain on a run of frames.
//
// Reference: IRFileLoader::readImage, IRFileLoader.cpp:1148-1247 (calibration == 0 branch):
//   bin_read_image -> H264_Loader -> VideoGrabber::toArray   byte-plane merge      h264.cpp:3016-3051
//   pixels[i] += min_T on the first min_T_height rows                              IRFileLoader.cpp:1174-1179
//   removeBadPixels(pixels, w, h - 3)                                              IRFileLoader.cpp:722-802
//   removeMotion(pixels, w, h - 3, pos) -> removeMotionGeneric                     IRFileLoader.cpp:617-627
//
// The reference walks the frame four times (merge, add, fix, translate + copy back).  Here:
//   loader_merge_kernel   ONE streaming pass: 16 low bytes + 16 high bytes (two 128-bit loads) are
//                         interleaved with PRMT, min_T is added per halfword where the row asks for
//                         it, 256-bit store; then one thread per flagged pixel of the span takes the
//                         median of the un-flagged cells of its shifted 3x3 window straight from the
//                         two byte planes (rows the CTA has just streamed: L1 hits, because spans with
//                         flagged pixels load with default caching) and overwrites the pixel.
//                         4 B/px: 2 read + 2 written.
//   translate_u16_tma_kernel<MOTION>  (translate.cu) for the motion step, 4 B/px.
#include "common.cuh"
#include "kernels.h"
#include "sort9.cuh"

namespace rirb {

constexpr int LD_THREADS = 256;
constexpr int LD_UNROLL = BP_SPAN / (LD_THREADS * 16);
static_assert(LD_UNROLL * LD_THREADS * 16 == BP_SPAN, "a span is a whole number of CTA-wide rows of 16-pixel vectors");

__device__ __forceinline__ unsigned merged_px(const u8* __restrict__ lo, const u8* __restrict__ hi, size_t i, unsigned add)
{
    return (((unsigned)lo[i] | ((unsigned)hi[i] << 8)) + add) & 0xFFFFu;
}

// VEC: planes 16-byte aligned, output 32-byte aligned, plane/frame strides multiples of 16 pixels
// and w % 16 == 0 (so a 16-pixel vector never straddles the min_T row limit).
template <bool VEC>
__global__ void __launch_bounds__(LD_THREADS)
loader_merge_kernel(const u8* __restrict__ lo, const u8* __restrict__ hi, u16* __restrict__ out, const int* __restrict__ xy,
                    const int* __restrict__ span_off, const int* __restrict__ nbr, int handle_spans, int w, int h, int hb, int npx,
                    int spans, unsigned min_t, int t_limit_px, size_t frame_stride)
{
    const int span = blockIdx.x % spans;
    const size_t f = blockIdx.x / spans;
    const int s0 = span * BP_SPAN;
    const int s1 = min(npx, s0 + BP_SPAN);
    const u8* flo = lo + f * frame_stride;
    const u8* fhi = hi + f * frame_stride;
    u16* oframe = out + f * frame_stride;
    int a = 0, b = 0;
    if (xy != nullptr && span < handle_spans) {
        a = span_off[span];
        b = span_off[span + 1];
    }
    uint4 rl[LD_UNROLL], rh[LD_UNROLL];
    const int nvec = VEC ? (s1 - s0) >> 4 : 0;
    if (VEC) {
        const uint4* gl = reinterpret_cast<const uint4*>(flo + s0);
        const uint4* gh = reinterpret_cast<const uint4*>(fhi + s0);
#pragma unroll
        for (int k = 0; k < LD_UNROLL; ++k) {
            const int j = threadIdx.x + k * LD_THREADS;
            if (j < nvec) {
                if (a != b) {  // spans with flagged pixels: default caching, their medians re-read these rows
                    rl[k] = __ldg(gl + j);
                    rh[k] = __ldg(gh + j);
                } else {
                    rl[k] = ld_stream(gl + j);
                    rh[k] = ld_stream(gh + j);
                }
            }
        }
    }
    // first fix-up of this thread (evaluated after the streaming stores, see below)
    const int i0 = a + (int)threadIdx.x;
    int2 p0 = make_int2(0, 0);
    unsigned med0 = 0;
    bool have0 = false;
    auto fix = [&](int2 p, int flags, unsigned& med) -> bool {
        // IRFileLoader.cpp:754-795: 3x3 window shifted inside the hb-row image, flagged cells skipped
        int x0 = p.x - 1, y0 = p.y - 1;
        if (p.x == 0) x0 = 0; else if (p.x == w - 1) x0 = w - 3;
        if (p.y == 0) y0 = 0; else if (p.y == hb - 1) y0 = hb - 3;
        // which of the 9 cells are flagged was worked out once, at create time (nbr): 18 independent byte loads
        unsigned v[9];
        int c = 0;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const int xx = x0 + k / 3, yy = y0 + k % 3;
            const size_t i = (size_t)yy * w + xx;
            const unsigned val = merged_px(flo, fhi, i, (int)i < t_limit_px ? min_t : 0u);
            const bool ok = !((flags >> k) & 1);
            v[k] = ok ? val : 0xFFFFFFFFu;
            c += ok;
        }
        if (c == 0) return false;  // the reference reads a stale stack slot here (undefined): leave the pixel
        sort9(v);
        med = pick_mid(v, c);
        return true;
    };
    int done = s0;
    if (VEC) {
        const unsigned t2 = min_t | (min_t << 16);
        U32x8* o = reinterpret_cast<U32x8*>(oframe + s0);
#pragma unroll
        for (int k = 0; k < LD_UNROLL; ++k) {
            const int j = threadIdx.x + k * LD_THREADS;
            if (j < nvec) {
                const unsigned add = (s0 + 16 * j < t_limit_px) ? t2 : 0u;
                const unsigned l[4] = {rl[k].x, rl[k].y, rl[k].z, rl[k].w};
                const unsigned hh[4] = {rh[k].x, rh[k].y, rh[k].z, rh[k].w};
                U32x8 r;
#pragma unroll
                for (int q = 0; q < 4; ++q) {  // bytes (l0 h0 l1 h1) and (l2 h2 l3 h3): v = lo | hi << 8 (h264.cpp:3030,3044)
                    r.v[2 * q] = __vadd2(__byte_perm(l[q], hh[q], 0x5140), add);
                    r.v[2 * q + 1] = __vadd2(__byte_perm(l[q], hh[q], 0x7362), add);
                }
                st_stream256(o + j, r);
            }
        }
        done = s0 + (nvec << 4);
    }
    for (int i = done + threadIdx.x; i < s1; i += LD_THREADS) oframe[i] = (u16)merged_px(flo, fhi, i, i < t_limit_px ? min_t : 0u);
    if (a == b) return;  // CTA-uniform
    // The medians re-read, byte by byte, rows this CTA has just streamed.  With the streaming (L1 no-allocate) loads
    // every one of those 18 sectors per flagged pixel came from DRAM a second time (ncu: 1.73 GB read for 1.31 GB of
    // planes) and the pass ran at 0.78 of peak; spans that hold flagged pixels therefore load through L1 (above), and
    // the gather runs after the stores, when those lines have arrived: 1.03 of peak, the same as without medians.
    if (i0 < b) {
        p0 = reinterpret_cast<const int2*>(xy)[i0];
        have0 = fix(p0, nbr[i0], med0);
    }
    __syncthreads();
    if (have0) oframe[(size_t)p0.y * w + p0.x] = (u16)med0;
    for (int i = i0 + LD_THREADS; i < b; i += LD_THREADS) {
        const int2 p = reinterpret_cast<const int2*>(xy)[i];
        unsigned med;
        if (fix(p, nbr[i], med)) oframe[(size_t)p.y * w + p.x] = (u16)med;
    }
}

// xy_dev / span_off_dev / nbr_dev: the handle's list, span offsets and window flags for a w x hb image
// (nullptr: no bad-pixel correction in this pass).  frame_stride in pixels for all three buffers.
int launch_loader_merge(const u8* lo, const u8* hi, u16* out, int w, int h, int hb, long long nframes, size_t frame_stride,
                        int min_t, int min_t_height, const int* xy_dev, const int* span_off_dev, const int* nbr_dev,
                        cudaStream_t st)
{
    if (nframes <= 0 || w <= 0 || h <= 0) return 0;
    const long long npx = (long long)w * h;
    const long long spans = ceil_div(npx, BP_SPAN);
    const long long grid = nframes * spans;
    if (npx > 0x7FFFFFFFLL || grid > 0x7FFFFFFFLL) {
        set_error("loader_read: too many pixels or frames in one call (%lld frames)", nframes);
        return -1;
    }
    const int handle_spans = (int)ceil_div((long long)w * hb, BP_SPAN);
    const int rows_t = min_t_height < 0 ? 0 : (min_t_height > h ? h : min_t_height);
    const int t_limit_px = min_t != 0 ? rows_t * w : 0;
    const unsigned mt = (unsigned)min_t & 0xFFFFu;
    const bool vec = (w % 16 == 0) && aligned16(lo) && aligned16(hi) && aligned32(out) && (frame_stride % 16 == 0);
    if (vec)
        RIRB_LAUNCH(loader_merge_kernel<true>, (unsigned)grid, LD_THREADS, 0, st, lo, hi, out, xy_dev, span_off_dev, nbr_dev,
                    handle_spans, w, h, hb, (int)npx, (int)spans, mt, t_limit_px, frame_stride);
    else
        RIRB_LAUNCH(loader_merge_kernel<false>, (unsigned)grid, LD_THREADS, 0, st, lo, hi, out, xy_dev, span_off_dev, nbr_dev,
                    handle_spans, w, h, hb, (int)npx, (int)spans, mt, t_limit_px, frame_stride);
    return 0;
}

// pixels[i] += min_T on the first t_rows rows of every frame, modulo 2^16 (IRFileLoader.cpp:1174-1179), in place: for frames
// that reach the chain as uint16 already (the zstd movie file), where there is no merge pass to fold it into
__global__ void loader_add_min_kernel(u16* __restrict__ frames, int t_px, unsigned min_t, size_t frame_stride)
{
    u16* fr = frames + (size_t)blockIdx.y * frame_stride;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < t_px; i += gridDim.x * blockDim.x) fr[i] = (u16)(fr[i] + min_t);
}

int launch_loader_add_min(u16* frames, int w, int t_rows, long long nframes, size_t frame_stride, int min_t, cudaStream_t st)
{
    if (nframes <= 0 || t_rows <= 0 || min_t == 0) return 0;
    const int t_px = w * t_rows;
    for (long long f0 = 0; f0 < nframes; f0 += 65535) {
        const long long n = min(nframes - f0, 65535LL);
        const dim3 grid((unsigned)min((long long)ceil_div(t_px, 256), 64LL), (unsigned)n);
        RIRB_LAUNCH(loader_add_min_kernel, grid, 256, 0, st, frames + f0 * frame_stride, t_px, (unsigned)min_t & 0xFFFFu, frame_stride);
    }
    return 0;
}

}  // namespace rirb
