// librir_b200/csrc/translate.cu -- sub-pixel translation (the resampling step of motion correction).
//
// Reference semantics: rir::translate<T,U> Filters.h:249-326 (+ TranslateBorder :238-244,
// detail::wrap/cast :231-235), C facade signal_processing.cpp:14-73, motion-correction variant
// removeMotionGeneric IRFileLoader.cpp:617-627.  Restated in SURVEY.md appendix A.4.
//
// Arithmetic: source coordinates in float32, blend in float64 with the reference's operation
// order and NO fused multiply-add (__dmul_rn/__dadd_rn), truncating store -- so integer outputs
// are bit-identical to the reference, not merely within +-1 LSB.
//
// Two kernels:
//   translate_generic_kernel<T,U>  every dtype / strategy, one thread per destination pixel.
//   translate_u16_kernel           the uint16 hot path: 4 pixels per thread, the two source rows
//                                  staged once per thread as 5 columns, the vertical blend done
//                                  in exact 64-bit integer arithmetic (see the comment there) so
//                                  that the FP64 pipe carries 5.25 instead of 13 ops per pixel.
#include "common.cuh"
#include "kernels.h"

namespace rirb {

// ------------------------------------------------------------------------------------------------
// pixel traits: promotion to double in the blend, detail::cast<U>(double) on the way out
// ------------------------------------------------------------------------------------------------
template <typename T> struct Pix {
    __device__ static __forceinline__ double to_double(T v) { return (double)v; }
    __device__ static __forceinline__ T from_double(double v) { return (T)v; }  // cvt.rzi for integers, rn for float
};
struct BoolPix {  // storage type of numpy bool: one byte, 0/1
    u8 v;
};
template <> struct Pix<BoolPix> {
    __device__ static __forceinline__ double to_double(BoolPix b) { return b.v ? 1.0 : 0.0; }  // bool -> int -> double
    __device__ static __forceinline__ BoolPix from_double(double v) { return BoolPix{(u8)(v != 0.0)}; }
};

// (size_t)float as x86-64 evaluates it for negative inputs: truncate as signed 64-bit, reinterpret.
__device__ __forceinline__ unsigned long long f2size(float v) { return (unsigned long long)(long long)v; }
__device__ __forceinline__ unsigned long long wrap_idx(unsigned long long v, unsigned long long n) { return (v + n) % n; }

__device__ __forceinline__ double blend4(double p1, double p2, double p3, double p4, double u, double v)
{
    const double omv = __dsub_rn(1.0, v);
    const double omu = __dsub_rn(1.0, u);
    const double left = __dadd_rn(__dmul_rn(p1, omv), __dmul_rn(p2, v));
    const double right = __dadd_rn(__dmul_rn(p3, omv), __dmul_rn(p4, v));
    return __dadd_rn(__dmul_rn(left, omu), __dmul_rn(right, u));
}

// One destination pixel, any strategy.  Returns false when dst must be left untouched.
template <typename T, typename U>
__device__ __forceinline__ bool translate_pixel(const T* __restrict__ src, int w, int h, int x, int y, float dx, float dy,
                                                int strategy, U background, U& result)
{
    const float px = (float)x - dx;
    const float py = (float)y - dy;
    const float fw = (float)w, fh = (float)h;
    long long l, rt, t, b;
    double u, v;
    if (px < 0 || px >= fw || py < 0 || py >= fh) {
        if (strategy == STRAT_NOBORDER) return false;
        if (strategy == STRAT_BACKGROUND) {
            result = background;
            return true;
        }
        if (strategy == STRAT_NEAREST) {
            long long sx = px < 0 ? 0 : (px >= fw ? w - 1 : (long long)px);
            long long sy = py < 0 ? 0 : (py >= fh ? h - 1 : (long long)py);
            // plain conversion T -> U (identity for the facade; u16 -> float in the motion variant)
            result = Pix<U>::from_double(Pix<T>::to_double(src[sy * w + sx]));
            return true;
        }
        l = (long long)wrap_idx(f2size(px), (unsigned long long)w);
        rt = (long long)wrap_idx(f2size(px + 1.0f), (unsigned long long)w);
        t = (long long)wrap_idx(f2size(py), (unsigned long long)h);
        b = (long long)wrap_idx(f2size(py + 1.0f), (unsigned long long)h);
        u = (double)fabsf(px - (float)(int)px);
        v = (double)fabsf(py - (float)(int)py);
    } else {
        l = (long long)px;
        rt = (long long)(px + 1.0f);
        if (rt == w) rt = l;
        t = (long long)py;
        b = (long long)(py + 1.0f);
        if (b == h) b = t;
        u = (double)(px - (float)l);
        v = (double)((float)b - py);
    }
    const double p1 = Pix<T>::to_double(src[b * w + l]);
    const double p2 = Pix<T>::to_double(src[t * w + l]);
    const double p3 = Pix<T>::to_double(src[b * w + rt]);
    const double p4 = Pix<T>::to_double(src[t * w + rt]);
    result = Pix<U>::from_double(blend4(p1, p2, p3, p4, u, v));
    return true;
}

template <typename T>
__global__ void __launch_bounds__(256)
translate_generic_kernel(const T* __restrict__ src, T* __restrict__ dst, int w, int h, long long nframes, size_t frame_stride,
                         const float* __restrict__ dxs, const float* __restrict__ dys, float dx0, float dy0, int strategy,
                         T background)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    for (long long f = blockIdx.z; f < nframes; f += gridDim.z) {
        const float dx = dxs ? dxs[f] : dx0;
        const float dy = dys ? dys[f] : dy0;
        T r;
        if (translate_pixel<T, T>(src + f * frame_stride, w, h, x, y, dx, dy, strategy, background, r))
            dst[f * frame_stride + (size_t)y * w + x] = r;
    }
}

template <typename T>
static int launch_generic(const void* src, void* dst, int w, int h, long long nframes, const float* dxs, const float* dys,
                          float dx0, float dy0, int strategy, const void* background_host, cudaStream_t st)
{
    T bg = *reinterpret_cast<const T*>(background_host);
    dim3 block(32, 8);
    dim3 grid((unsigned)ceil_div(w, 32), (unsigned)ceil_div(h, 8), (unsigned)min(nframes, 32768LL));
    RIRB_LAUNCH(translate_generic_kernel<T>, grid, block, 0, st, (const T*)src, (T*)dst, w, h, nframes, (size_t)w * h, dxs, dys,
                dx0, dy0, strategy, bg);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// uint16 hot path
// ------------------------------------------------------------------------------------------------
// Exact-integer vertical blend.  For a destination row, py = (float)y - dy, and in the in-range
// case v = (float)b - py, 1-v are the reference's vertical weights.  Whenever v is a multiple of
// 2^-23 with |v| <= 1 (always true for |py| >= 1, where float32 spacing is >= 2^-23 ... 2^-13)
// both weights are integers/2^23 below 2^24+1, a pixel is below 2^16, so
//     p_b*(1-v) + p_t*v  =  (p_b*A + p_t*B) / 2^23,   A = (1-v)*2^23, B = v*2^23
// has at most 41 significant bits: each fp64 product and their sum in the reference are EXACT,
// and equal the integer N = p_b*A + p_t*B scaled by 2^-23.  N is formed with two 64-bit integer
// multiply-adds and turned into the double N*2^-23 by writing it into the mantissa of 2^29 and
// subtracting 2^29 (one exact DADD).  Rows where the condition fails (0 <= py < 1 with a finer
// fraction, e.g. denormal shifts) take the plain fp64 column blend.  The horizontal blend keeps
// the reference's rounded fp64 operations: c_l*(1-u) + c_r*u.
__device__ __forceinline__ double int_to_double_scaled23(long long n)
{
    // n in [0, 2^41): double with exponent of 2^29 has ulp 2^-23 => bits = bits(2^29) + n
    const long long bits = 0x41C0000000000000LL + n;
    return __dsub_rn(__longlong_as_double(bits), 536870912.0);
}

constexpr int TR_PX = 4;  // destination pixels per thread

template <bool MOTION>
__global__ void __launch_bounds__(256)
translate_u16_kernel(const u16* __restrict__ src, u16* __restrict__ dst, int w, int h, long long nframes, size_t src_stride,
                     size_t dst_stride, const float* __restrict__ dxs, const float* __restrict__ dys, float dx0, float dy0,
                     int strategy, unsigned background)
{
    const int xg = blockIdx.x * blockDim.x + threadIdx.x;  // group of TR_PX pixels
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int x0 = xg * TR_PX;
    if (x0 >= w || y >= h) return;
    const float fw = (float)w, fh = (float)h;
    for (long long f = blockIdx.z; f < nframes; f += gridDim.z) {
        const float dx = dxs ? dxs[f] : dx0;
        const float dy = dys ? dys[f] : dy0;
        const u16* frame = src + f * src_stride;
        u16* orow = dst + f * dst_stride + (size_t)y * w + x0;
        const float py = (float)y - dy;
        const bool row_in = !(py < 0 || py >= fh);
        // fast path: whole group in range, source columns consecutive, v on the 2^-23 grid
        const float px_first = (float)x0 - dx;
        const float px_last = (float)(x0 + TR_PX - 1) - dx;
        bool fast = row_in && (x0 + TR_PX <= w) && !(px_first < 0) && (px_last < fw);
        long long l0 = 0;
        int t = 0, b = 0;
        float vf = 0.f;
        if (fast) {
            l0 = (long long)px_first;
            t = (int)py;
            b = (int)(py + 1.0f);
            if (b == h) b = t;
            vf = (float)b - py;
            const float vs = vf * 8388608.0f;
            fast = (vs == truncf(vs)) && (fabsf(vf) <= 1.0f);
#pragma unroll
            for (int i = 0; i < TR_PX; ++i) {  // l_i = l0+i and rt_i = l_i+1, as the reference would compute them
                const float pxi = (float)(x0 + i) - dx;
                fast = fast && ((long long)pxi == l0 + i) && ((long long)(pxi + 1.0f) == l0 + i + 1);
            }
        }
        if (fast) {
            const int B = (int)(vf * 8388608.0f);
            const int A = 8388608 - B;
            const u16* rb = frame + (size_t)b * w + l0;
            const u16* rt_ = frame + (size_t)t * w + l0;
            const int ncol = (l0 + TR_PX < w) ? TR_PX + 1 : TR_PX;  // right edge: rt clamps to l
            double c[TR_PX + 1];
#pragma unroll
            for (int i = 0; i <= TR_PX; ++i) {
                if (i < ncol) {
                    long long n = (long long)rb[i] * A + (long long)rt_[i] * B;
                    c[i] = int_to_double_scaled23(n);
                } else {
                    c[i] = c[i - 1];
                }
            }
            unsigned outv[TR_PX];
#pragma unroll
            for (int i = 0; i < TR_PX; ++i) {
                const float px = (float)(x0 + i) - dx;
                const double u = (double)(px - (float)(l0 + i));
                const double omu = __dsub_rn(1.0, u);
                const double val = __dadd_rn(__dmul_rn(c[i], omu), __dmul_rn(c[i + 1], u));
                outv[i] = MOTION ? (unsigned)(u16)(float)val : (unsigned)(u16)val;
            }
            if ((((uintptr_t)orow) & 7u) == 0) {
                uint2 o;
                o.x = outv[0] | (outv[1] << 16);
                o.y = outv[2] | (outv[3] << 16);
                *reinterpret_cast<uint2*>(orow) = o;
            } else {
#pragma unroll
                for (int i = 0; i < TR_PX; ++i) orow[i] = (u16)outv[i];
            }
        } else {
#pragma unroll
            for (int i = 0; i < TR_PX; ++i) {
                if (x0 + i < w) {
                    if (MOTION) {
                        float r;
                        if (translate_pixel<u16, float>(frame, w, h, x0 + i, y, dx, dy, strategy, (float)background, r))
                            orow[i] = (u16)r;
                    } else {
                        u16 r;
                        if (translate_pixel<u16, u16>(frame, w, h, x0 + i, y, dx, dy, strategy, (u16)background, r)) orow[i] = r;
                    }
                }
            }
        }
    }
}

int launch_translate_u16(const u16* src, u16* dst, int w, int h, long long nframes, size_t src_stride, size_t dst_stride,
                         const float* dxs, const float* dys, float dx0, float dy0, int strategy, unsigned background, bool motion,
                         cudaStream_t st)
{
    if (nframes <= 0 || w <= 0 || h <= 0) return 0;
    dim3 block(32, 8);
    dim3 grid((unsigned)ceil_div(ceil_div(w, TR_PX), 32), (unsigned)ceil_div(h, 8), (unsigned)min(nframes, 32768LL));
    if (motion)
        RIRB_LAUNCH(translate_u16_kernel<true>, grid, block, 0, st, src, dst, w, h, nframes, src_stride, dst_stride, dxs, dys, dx0,
                    dy0, strategy, background);
    else
        RIRB_LAUNCH(translate_u16_kernel<false>, grid, block, 0, st, src, dst, w, h, nframes, src_stride, dst_stride, dxs, dys,
                    dx0, dy0, strategy, background);
    return 0;
}

int launch_translate(int type, const void* src, void* dst, int w, int h, long long nframes, const float* dxs, const float* dys,
                     float dx0, float dy0, int strategy, const void* background_host, cudaStream_t st)
{
    if (nframes <= 0 || w <= 0 || h <= 0) return 0;
    switch (type) {
    case '?': return launch_generic<BoolPix>(src, dst, w, h, nframes, dxs, dys, dx0, dy0, strategy, background_host, st);
    case 'b': return launch_generic<signed char>(src, dst, w, h, nframes, dxs, dys, dx0, dy0, strategy, background_host, st);
    case 'B': return launch_generic<unsigned char>(src, dst, w, h, nframes, dxs, dys, dx0, dy0, strategy, background_host, st);
    case 'h': return launch_generic<short>(src, dst, w, h, nframes, dxs, dys, dx0, dy0, strategy, background_host, st);
    case 'H':
        return launch_translate_u16((const u16*)src, (u16*)dst, w, h, nframes, (size_t)w * h, (size_t)w * h, dxs, dys, dx0, dy0,
                                    strategy, *(const u16*)background_host, false, st);
    case 'i': return launch_generic<int>(src, dst, w, h, nframes, dxs, dys, dx0, dy0, strategy, background_host, st);
    case 'I': return launch_generic<unsigned int>(src, dst, w, h, nframes, dxs, dys, dx0, dy0, strategy, background_host, st);
    case 'l': return launch_generic<long long>(src, dst, w, h, nframes, dxs, dys, dx0, dy0, strategy, background_host, st);
    case 'L': return launch_generic<unsigned long long>(src, dst, w, h, nframes, dxs, dys, dx0, dy0, strategy, background_host, st);
    case 'f': return launch_generic<float>(src, dst, w, h, nframes, dxs, dys, dx0, dy0, strategy, background_host, st);
    case 'd': return launch_generic<double>(src, dst, w, h, nframes, dxs, dys, dx0, dy0, strategy, background_host, st);
    default: set_error("translate: unknown dtype code %d", type); return -1;
    }
}

}  // namespace rirb
