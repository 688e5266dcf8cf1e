// librir_b200/csrc/gaussian.cu -- Gaussian filter, separable, streaming.
//
// Reference semantics: gaussian_filter / generate_kernel, signal_processing.cpp:79-148 (dense
// (2r+1)^2 float convolution, r = max(1,(int)(2*sigma)), partial-kernel renormalisation at the
// borders).  The reference's 2-D kernel is a product of 1-D kernels, and so is its border
// normaliser (the in-bounds tap set is a rectangle), so a separable evaluation agrees with it to
// float rounding (tolerance 1e-5 relative, BASELINE.json).
//
// Kernel design (r <= 4 -- sigma < 2.5, every shape in BASELINE.json):
//  gauss_tile_kernel (TMA-compatible layouts: 16-byte aligned rows): a CTA owns a 128 x 64 output
//   tile; ONE TMA box load brings the (128 + 2 halo vectors) x (64+2r) source window into shared memory, with
//   the out-of-image halo zero-filled by the hardware -- which is exactly the reference's
//   "taps outside the image do not count" once the sum is renormalised.  Each of the 4 warps walks
//   16 output rows: a lane reads its 4+2r source pixels of a row with two vector LDS, runs the
//   horizontal pass in registers and keeps a (2r+1)-row register ring for the vertical pass, which
//   is issued as packed FFMA2 (two columns per instruction, sm_100's fma.rn.f32x2).
//  gauss_sep_kernel (fallback, w % 4 == 0, no shared memory):
//   one WARP owns a 128-pixel-wide column strip of one frame and walks down its rows.  A lane
//   loads 4 pixels of the row (128-bit for f32, 64-bit for u16), takes the r pixels it needs from
//   each neighbour lane with warp shuffles (lanes 0/31 fetch the strip's halo vector, an L2 hit),
//   does the horizontal pass in registers, and pushes the result into a (2r+1)-row register ring
//   from which the vertical pass produces one output row per input row.  Every input byte is
//   read from HBM once and every output byte written once: 8 B/px (f32 -> f32), 6 B/px (u16 -> f32).
//   No shared memory, no __syncthreads; rows are loaded in batches of 4 for memory parallelism.
#include <math.h>

#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "sort9.cuh"
#include "tma.cuh"

namespace rirb {

int gaussian_taps_host(float sigma, GaussTaps* taps)
{
    int r = (int)(sigma * 2);
    if (r < 1) r = 1;
    if (r > GAUSS_MAX_RADIUS) {
        set_error("gaussian_filter: radius %d (sigma %g) exceeds the supported maximum %d", r, (double)sigma, GAUSS_MAX_RADIUS);
        return -1;
    }
    const int kw = 2 * r + 1;
    // the reference's 2-D taps, in float, same expression and accumulation order
    float* k2 = new float[(size_t)kw * kw];
    const float s = 2.0f * sigma * sigma;
    float sum = 0.0f;
    for (int x = -r; x <= r; ++x)
        for (int y = -r; y <= r; ++y) {
            float rho = (float)sqrt((double)(x * x + y * y));
            float e = expf(-(rho * rho) / s);
            float t = (float)((double)e / (3.14159265358979323846 * (double)s));
            k2[x + r + (y + r) * kw] = t;
            sum += t;
        }
    for (int i = 0; i < kw * kw; ++i) k2[i] /= sum;
    // 1-D taps = row sums (K[dx,dy] = k1[dx] * k1[dy])
    taps->radius = r;
    for (int d = 0; d < kw; ++d) {
        double acc = 0;
        for (int y = 0; y < kw; ++y) acc += (double)k2[d + y * kw];
        taps->k[d] = (float)acc;
    }
    delete[] k2;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// fast path
// ------------------------------------------------------------------------------------------------
template <typename TIN> struct RowVec;
template <> struct RowVec<float> {
    __device__ static __forceinline__ void load(const float* p, float (&a)[4])
    {
        float4 v = ld_stream(reinterpret_cast<const float4*>(p));
        a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
    }
};
template <> struct RowVec<u16> {
    __device__ static __forceinline__ void load(const u16* p, float (&a)[4])
    {
        uint2 v = ld_stream(reinterpret_cast<const uint2*>(p));
        a[0] = (float)(v.x & 0xFFFFu); a[1] = (float)(v.x >> 16);
        a[2] = (float)(v.y & 0xFFFFu); a[3] = (float)(v.y >> 16);
    }
};

constexpr int GS_ROWS = 4;     // rows loaded per batch
constexpr int GS_WARPS = 4;    // warps (independent work items) per CTA
constexpr int GS_STRIP = 128;  // pixels per warp-row

template <int R, typename TIN>
__global__ void __launch_bounds__(GS_WARPS * 32)
gauss_sep_kernel(const TIN* __restrict__ src, float* __restrict__ dst, int w, int h, long long nframes, int xstrips, int ystrips,
                 int rows_per_strip, GaussTaps taps)
{
    const int lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * GS_WARPS + (threadIdx.x >> 5);
    const long long items = nframes * ystrips * xstrips;
    if (item >= items) return;  // warp-uniform
    const int xstrip = (int)(item % xstrips);
    const int ystrip = (int)((item / xstrips) % ystrips);
    const long long f = item / ((long long)xstrips * ystrips);

    const int xs = xstrip * GS_STRIP;
    const int x = xs + 4 * lane;
    const bool xin = x < w;  // w % 4 == 0: a lane is entirely inside or outside
    const int ys = ystrip * rows_per_strip;
    const int ye = min(h, ys + rows_per_strip);
    const TIN* frame = src + (size_t)f * w * h;
    float* oframe = dst + (size_t)f * w * h;

    float k[2 * R + 1];
#pragma unroll
    for (int d = 0; d <= 2 * R; ++d) k[d] = taps.k[d];

    float kfull = 0.f;
#pragma unroll
    for (int d = 0; d <= 2 * R; ++d) kfull += k[d];

    // horizontal normaliser per owned pixel (sum of taps that fall inside the row)
    float nx[4];
    bool xborder[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float s = 0.f;
#pragma unroll
        for (int d = 0; d <= 2 * R; ++d) {
            int xx = x + j + d - R;
            if (xx >= 0 && xx < w) s += k[d];
        }
        nx[j] = s;
        xborder[j] = (x + j < R) || (x + j >= w - R);
    }

    const bool halo_l = (lane == 0) && (xs > 0);
    const bool halo_r = (lane == 31) && (xs + GS_STRIP < w);
    const int halo_x = halo_l ? x - 4 : x + 4;

    float ring[2 * R + 1][4];
#pragma unroll
    for (int i = 0; i <= 2 * R; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) ring[i][j] = 0.f;

    for (int y0 = ys - R; y0 < ye + R; y0 += GS_ROWS) {
        float a[GS_ROWS][4], hv[GS_ROWS][4];
#pragma unroll
        for (int g = 0; g < GS_ROWS; ++g) {
            const int yy = y0 + g;
            const bool rowok = (yy >= 0) && (yy < h) && (yy < ye + R);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                a[g][j] = 0.f;
                hv[g][j] = 0.f;
            }
            if (rowok && xin) RowVec<TIN>::load(frame + (size_t)yy * w + x, a[g]);
            if (rowok && (halo_l || halo_r)) RowVec<TIN>::load(frame + (size_t)yy * w + halo_x, hv[g]);
        }
#pragma unroll
        for (int g = 0; g < GS_ROWS; ++g) {
            const int yy = y0 + g;
            if (yy >= ye + R) break;  // warp-uniform
            // neighbours' pixels: left lane's last R, right lane's first R
            float win[4 + 2 * R];
#pragma unroll
            for (int i = 0; i < R; ++i) {
                float fromL = __shfl_up_sync(0xFFFFFFFFu, a[g][4 - R + i], 1);
                float fromR = __shfl_down_sync(0xFFFFFFFFu, a[g][i], 1);
                win[i] = (lane == 0) ? hv[g][4 - R + i] : fromL;
                win[R + 4 + i] = (lane == 31) ? hv[g][i] : fromR;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) win[R + j] = a[g][j];
            // shift the ring, insert the horizontal pass of this row
#pragma unroll
            for (int i = 0; i < 2 * R; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) ring[i][j] = ring[i + 1][j];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float s = 0.f;
#pragma unroll
                for (int d = 0; d <= 2 * R; ++d) s = fmaf(k[d], win[j + d], s);
                ring[2 * R][j] = s;
            }
            // vertical pass -> output row yo
            const int yo = yy - R;
            if (yo >= ys && yo < ye) {
                float o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float s = 0.f;
#pragma unroll
                    for (int d = 0; d <= 2 * R; ++d) s = fmaf(k[d], ring[d][j], s);
                    o[j] = s;
                }
                const bool yborder = (yo < R) || (yo >= h - R);
                float ny = 0.f;
                if (yborder) {
#pragma unroll
                    for (int d = 0; d <= 2 * R; ++d) {
                        int yyy = yo + d - R;
                        if (yyy >= 0 && yyy < h) ny += k[d];
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (yborder || xborder[j]) o[j] = o[j] / ((yborder ? ny : kfull) * (xborder[j] ? nx[j] : kfull));
                }
                if (xin) st_stream(reinterpret_cast<float4*>(oframe + (size_t)yo * w + x), make_float4(o[0], o[1], o[2], o[3]));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// TMA-tiled path
// ------------------------------------------------------------------------------------------------
constexpr int GT_W = 128;          // output columns per CTA (4 per lane)
constexpr int GT_WARPS = 4;        // warps per CTA
constexpr int GT_RG = 16;          // output rows per warp
constexpr int GT_H = GT_WARPS * GT_RG;
// TMA boxes start on 16-byte boundaries in x (tma.cuh), so the left halo is a whole vector: GT_HALO
// source pixels (16 bytes) on each side, of which the r nearest are used.
template <typename TIN> struct GtBox {
    static constexpr int HALO = 16 / (int)sizeof(TIN);  // 8 (u16) or 4 (f32)
    static constexpr int BW = GT_W + 2 * HALO;
};

// 4 consecutive tile elements starting at p (8-byte aligned for u16, 16-byte aligned for f32) -> float
__device__ __forceinline__ void load4(const u16* p, float (&a)[4])
{
    // uint16 -> float without I2F (16 lanes/clk/SM): PRMT builds the bits of 2^23 + p, FADD removes 2^23
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    a[0] = __uint_as_float(__byte_perm(v.x, 0x4B000000u, 0x7610)) - 8388608.0f;
    a[1] = __uint_as_float(__byte_perm(v.x, 0x4B000000u, 0x7632)) - 8388608.0f;
    a[2] = __uint_as_float(__byte_perm(v.y, 0x4B000000u, 0x7610)) - 8388608.0f;
    a[3] = __uint_as_float(__byte_perm(v.y, 0x4B000000u, 0x7632)) - 8388608.0f;
}
__device__ __forceinline__ void load4(const float* p, float (&a)[4])
{
    const float4 v = *reinterpret_cast<const float4*>(p);
    a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
}

// Bad-pixel correction fused in front of the filter (uint16 input only): the box arrives RAW, the flagged pixels that
// fall into it (halo included) are replaced in shared memory by BadPixels::correct's median -- taken from the raw frame
// in global memory, so neighbouring flagged pixels do not see each other's replacement -- every pixel is raised to the
// clamp level, the corrected interior is stored (it is what translate reads next) and the filter runs on the corrected
// tile.  The movie is then read once for correction + filter: 8 B/px (2 in, 2 + 4 out) instead of 4 + 6.
struct BpFuse {
    const u16* raw;       // the movie the tensor map describes
    const int* xy;        // flagged pixels, raster order (x, y)
    const int* row_off;   // first list entry of each image row, h + 1 entries
    u16* corrected;       // [n][h][w] output of the correction
    unsigned clamp;       // 0: none
};

template <int R, typename TIN, bool BP>
__global__ void __launch_bounds__(GT_WARPS * 32)
gauss_tile_kernel(const __grid_constant__ CUtensorMap tmap, float* __restrict__ dst, int w, int h, int tiles_x, int tiles_y,
                  GaussTaps taps, BpFuse bf)
{
    constexpr int BH = GT_H + 2 * R;
    constexpr int HALO = GtBox<TIN>::HALO;
    static_assert(R <= 4, "the window below is 4 pixels either side of the lane's own 4");
    __shared__ __align__(128) TIN tile[BH][GtBox<TIN>::BW];
    __shared__ __align__(8) unsigned long long bar;
    const int tiles = tiles_x * tiles_y;
    const long long f = blockIdx.x / tiles;
    const int tile_id = (int)(blockIdx.x - f * tiles);
    const int ty = tile_id / tiles_x, tx = tile_id - ty * tiles_x;
    const int x0t = tx * GT_W, y0t = ty * GT_H;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, (unsigned)sizeof(tile));
        tma_load_box(&tile[0][0], &tmap, &bar, x0t - HALO, y0t - R, (int)f);
    }

    // ---- per-thread constants (while the box is in flight) ---------------------------------------
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int x = x0t + 4 * lane;
    float k[2 * R + 1];
#pragma unroll
    for (int d = 0; d <= 2 * R; ++d) k[d] = taps.k[d];
    float kfull = 0.f;
#pragma unroll
    for (int d = 0; d <= 2 * R; ++d) kfull += k[d];
    // Horizontal normaliser of each owned column: the partial tap sum at the image's left / right
    // edge, the full sum elsewhere.  rx[j] is what an output of a row that is NOT a top/bottom
    // border row is multiplied by: 1 for interior columns (the reference does not normalise interior
    // pixels, and x * 1.0f is exact), 1 / (kfull * partial) for edge columns.  Multiplying instead of
    // branching keeps the warps that own the image's edge columns from taking a divergent division
    // on every row (that cost 40 % of the tiles of a 640-wide frame their speed).
    float nx[4], rx[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float s = 0.f;
#pragma unroll
        for (int d = 0; d <= 2 * R; ++d) {
            const int xx = x + j + d - R;
            if (xx >= 0 && xx < w) s += k[d];
        }
        const bool xb = (x + j < R) || (x + j >= w - R);
        nx[j] = xb ? s : kfull;
        rx[j] = xb ? 1.0f / (kfull * s) : 1.0f;
    }
    float* oframe = dst + (size_t)f * w * h;
    float2 ring[2 * R + 1][2];
#pragma unroll
    for (int i = 0; i <= 2 * R; ++i) ring[i][0] = ring[i][1] = make_float2(0.f, 0.f);

    if (BP) {
        static_assert(!BP || sizeof(TIN) == 2, "the fused correction is for uint16 frames");
        static_assert(BH <= GT_WARPS * 32, "one thread per box row");
        constexpr int BW = GtBox<TIN>::BW;
        u16* t16 = reinterpret_cast<u16*>(&tile[0][0]);
        const u16* rawf = bf.raw + (size_t)f * w * h;
        const int bx0 = x0t - HALO, by = y0t - R + (int)threadIdx.x;  // this thread's box row
        // the row's flagged pixels inside the box; the first one's median is gathered while the box is in flight
        int a = 0, b = 0, first_x = -1;
        unsigned first_v = 0;
        if ((int)threadIdx.x < BH && by >= 0 && by < h) {
            a = bf.row_off[by];
            b = bf.row_off[by + 1];
            for (; a < b; ++a) {
                const int x = bf.xy[2 * a];
                if (x >= bx0 + BW) {
                    a = b;
                    break;
                }
                if (x >= bx0) {
                    first_x = x;
                    first_v = median3x3_global(rawf, w, h, x, by);
                    ++a;
                    break;
                }
            }
        }
        mbar_wait(&bar, 0);
        if (first_x >= 0) t16[threadIdx.x * BW + (first_x - bx0)] = (u16)first_v;
        for (; a < b; ++a) {  // further flagged pixels of the row: rare
            const int x = bf.xy[2 * a];
            if (x >= bx0 + BW) break;
            t16[threadIdx.x * BW + (x - bx0)] = (u16)median3x3_global(rawf, w, h, x, by);
        }
        __syncthreads();
        // clamp the whole box, store the corrected interior: 8 pixels (16 bytes) per step
        const unsigned c2 = bf.clamp | (bf.clamp << 16);
        u16* cframe = bf.corrected + (size_t)f * w * h;
        for (int v = threadIdx.x; v < BH * (BW / 8); v += GT_WARPS * 32) {
            const int r = v / (BW / 8), c8 = v - r * (BW / 8);
            const int gy = y0t - R + r, gx = bx0 + 8 * c8;
            if (gy < 0 || gy >= h || gx < 0 || gx >= w) continue;  // outside the image the box keeps the hardware's zeros
            uint4* p = reinterpret_cast<uint4*>(t16 + r * BW + 8 * c8);  // w % 8 == 0 on this path: a vector is in or out
            uint4 q = *p;
            q.x = vmaxu2(q.x, c2);
            q.y = vmaxu2(q.y, c2);
            q.z = vmaxu2(q.z, c2);
            q.w = vmaxu2(q.w, c2);
            *p = q;
            if (r >= R && r < R + GT_H && c8 >= 1 && c8 <= GT_W / 8) st_stream(reinterpret_cast<uint4*>(cframe + (size_t)gy * w + gx), q);
        }
        __syncthreads();
    } else {
        mbar_wait(&bar, 0);
    }

#pragma unroll
    for (int rr = 0; rr < GT_RG + 2 * R; ++rr) {
        // aligned vectors left / own / right of the lane's 4 pixels; win[j] = source column x - R + j
        const TIN* srow = &tile[wi * GT_RG + rr][HALO + 4 * lane];
        float vl[4], vc[4], vr[4], win[4 + 2 * R];
        load4(srow - 4, vl);
        load4(srow, vc);
        load4(srow + 4, vr);
#pragma unroll
        for (int j = 0; j < R; ++j) {
            win[j] = vl[4 - R + j];
            win[R + 4 + j] = vr[j];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) win[R + j] = vc[j];
        // horizontal pass of this source row -> newest ring slot
#pragma unroll
        for (int i = 0; i < 2 * R; ++i) {
            ring[i][0] = ring[i + 1][0];
            ring[i][1] = ring[i + 1][1];
        }
        float hsum[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float s = k[0] * win[j];
#pragma unroll
            for (int d = 1; d <= 2 * R; ++d) s = fmaf(k[d], win[j + d], s);
            hsum[j] = s;
        }
        ring[2 * R][0] = make_float2(hsum[0], hsum[1]);
        ring[2 * R][1] = make_float2(hsum[2], hsum[3]);
        if (rr >= 2 * R) {
            const int yo = y0t + wi * GT_RG + rr - 2 * R;
            // vertical pass, two columns per FFMA2
            float2 o0 = make_float2(k[0] * ring[0][0].x, k[0] * ring[0][0].y);
            float2 o1 = make_float2(k[0] * ring[0][1].x, k[0] * ring[0][1].y);
#pragma unroll
            for (int d = 1; d <= 2 * R; ++d) {
                const float2 kk = make_float2(k[d], k[d]);
                o0 = __ffma2_rn(kk, ring[d][0], o0);
                o1 = __ffma2_rn(kk, ring[d][1], o1);
            }
            float o[4] = {o0.x, o0.y, o1.x, o1.y};
            const bool yborder = (yo < R) || (yo >= h - R);
            if (yborder) {  // warp-uniform; partial-kernel renormalisation (signal_processing.cpp:130-144)
                float ny = 0.f;
#pragma unroll
                for (int d = 0; d <= 2 * R; ++d) {
                    const int yyy = yo + d - R;
                    if (yyy >= 0 && yyy < h) ny += k[d];
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] = o[j] / (ny * nx[j]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] *= rx[j];
            }
            if (x < w && yo < h) st_stream(reinterpret_cast<float4*>(oframe + (size_t)yo * w + x), make_float4(o[0], o[1], o[2], o[3]));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// TMA-tiled path for the wide kernels (R = 3, 4: sigma 1.5 .. 2.49, a named configuration of BASELINE.json)
// ------------------------------------------------------------------------------------------------
// gauss_tile_kernel is FMA-bound there (0.80 / 0.61 of the copy bandwidth in round 1): 2R+1 scalar FFMA per pixel in the
// horizontal pass, run over 16 + 2R source rows for 16 output rows, in a fully unrolled body of 24 rows that no longer
// fits the 32 KB instruction cache.  Same tile staging, different inner loop:
//   * source rows are taken TWO AT A TIME and the horizontal pass is packed across the two rows: (win_r[i], win_r+1[i])
//     sit in one register pair because the unpack writes them there, so one FFMA2 does a tap of both rows -- half the
//     instructions of the pass, no repacking of its inputs;
//   * a warp owns 32 output rows instead of 16: the 2R halo rows cost 2R/32 instead of 2R/16 of the horizontal work;
//   * the ring of horizontally filtered rows has 2R+2 slots addressed modulo at compile time and the loop is unrolled by
//     exactly one turn of the ring, so the body is 2R+2 rows long whatever the strip height.
// 4 consecutive elements of two staged rows -> (row 0, row 1) pairs as float: for uint16 two PRMT build the bits of
// 2^23 + p in an aligned register pair and ONE packed add (sm_100 add.f32x2) removes the bias of both
__device__ __forceinline__ void load4x2(const u16* p0, const u16* p1, float2 (&a)[4])
{
    const uint2 v = *reinterpret_cast<const uint2*>(p0);
    const uint2 u = *reinterpret_cast<const uint2*>(p1);
    const float2 bias = make_float2(-8388608.0f, -8388608.0f);
    a[0] = __fadd2_rn(make_float2(__uint_as_float(__byte_perm(v.x, 0x4B000000u, 0x7610)), __uint_as_float(__byte_perm(u.x, 0x4B000000u, 0x7610))), bias);
    a[1] = __fadd2_rn(make_float2(__uint_as_float(__byte_perm(v.x, 0x4B000000u, 0x7632)), __uint_as_float(__byte_perm(u.x, 0x4B000000u, 0x7632))), bias);
    a[2] = __fadd2_rn(make_float2(__uint_as_float(__byte_perm(v.y, 0x4B000000u, 0x7610)), __uint_as_float(__byte_perm(u.y, 0x4B000000u, 0x7610))), bias);
    a[3] = __fadd2_rn(make_float2(__uint_as_float(__byte_perm(v.y, 0x4B000000u, 0x7632)), __uint_as_float(__byte_perm(u.y, 0x4B000000u, 0x7632))), bias);
}
__device__ __forceinline__ void load4x2(const float* p0, const float* p1, float2 (&a)[4])
{
    const float4 v = *reinterpret_cast<const float4*>(p0);
    const float4 u = *reinterpret_cast<const float4*>(p1);
    a[0] = make_float2(v.x, u.x);
    a[1] = make_float2(v.y, u.y);
    a[2] = make_float2(v.z, u.z);
    a[3] = make_float2(v.w, u.w);
}

constexpr int GW_RG = 32;                    // output rows per warp
constexpr int GW_H = GT_WARPS * GW_RG;       // 128 output rows per CTA

template <int R, typename TIN>
__global__ void __launch_bounds__(GT_WARPS * 32, sizeof(TIN) == 2 ? 4 : 3)
gauss_tile_wide_kernel(const __grid_constant__ CUtensorMap tmap, float* __restrict__ dst, int w, int h, int tiles_x, int tiles_y,
                       GaussTaps taps)
{
    constexpr int BH = GW_H + 2 * R;
    constexpr int BW = GtBox<TIN>::BW;
    constexpr int HALO = GtBox<TIN>::HALO;
    constexpr int RS = 2 * R + 2;        // ring slots (even: rows come in pairs)
    constexpr int S = GW_RG + 2 * R;     // source rows of a warp's strip (even)
    static_assert(R >= 1 && R <= 4 && (S % 2) == 0, "window: 4 pixels either side of the lane's own 4; rows in pairs");
    extern __shared__ __align__(128) unsigned char gw_smem[];
    TIN(*tile)[BW] = reinterpret_cast<TIN(*)[BW]>(gw_smem);
    __shared__ __align__(8) unsigned long long bar;
    const int tiles = tiles_x * tiles_y;
    const long long f = blockIdx.x / tiles;
    const int tile_id = (int)(blockIdx.x - f * tiles);
    const int ty = tile_id / tiles_x, tx = tile_id - ty * tiles_x;
    const int x0t = tx * GT_W, y0t = ty * GW_H;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, (unsigned)(BH * BW * sizeof(TIN)));
        tma_load_box(&tile[0][0], &tmap, &bar, x0t - HALO, y0t - R, (int)f);
    }
    // ---- per-thread constants (while the box is in flight) ----
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int x = x0t + 4 * lane;
    float k[2 * R + 1];
    float2 kk[2 * R + 1];
    float kfull = 0.f;
#pragma unroll
    for (int d = 0; d <= 2 * R; ++d) {
        k[d] = taps.k[d];
        kk[d] = make_float2(k[d], k[d]);
        kfull += k[d];
    }
    float nx[4], rx[4];  // see gauss_tile_kernel: horizontal normaliser of each owned column
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float s = 0.f;
#pragma unroll
        for (int d = 0; d <= 2 * R; ++d) {
            const int xx = x + j + d - R;
            if (xx >= 0 && xx < w) s += k[d];
        }
        const bool xb = (x + j < R) || (x + j >= w - R);
        nx[j] = xb ? s : kfull;
        rx[j] = xb ? 1.0f / (kfull * s) : 1.0f;
    }
    float* oframe = dst + (size_t)f * w * h;
    float2 ring[RS][2];
#pragma unroll
    for (int i = 0; i < RS; ++i) ring[i][0] = ring[i][1] = make_float2(0.f, 0.f);
    mbar_wait(&bar, 0);

#pragma unroll 1
    for (int b = 0; b < S; b += RS) {
#pragma unroll
        for (int i = 0; i < RS; i += 2) {
            const int rr = b + i;  // source rows rr and rr + 1 of the warp's strip
            if (rr >= S) break;    // warp-uniform (the last turn of the ring may be partial)
            // ---- unpack both rows, interleaved: wp[c] = (row rr, row rr + 1) at source column x - R + c ----
            const TIN* s0 = &tile[wi * GW_RG + rr][HALO + 4 * lane];
            const TIN* s1 = s0 + BW;
            float2 av[4];
            float2 wp[4 + 2 * R];
            load4x2(s0 - 4, s1 - 4, av);
#pragma unroll
            for (int j = 0; j < R; ++j) wp[j] = av[4 - R + j];
            load4x2(s0, s1, av);
#pragma unroll
            for (int j = 0; j < 4; ++j) wp[R + j] = av[j];
            load4x2(s0 + 4, s1 + 4, av);
#pragma unroll
            for (int j = 0; j < R; ++j) wp[R + 4 + j] = av[j];
            // ---- horizontal pass of both rows: one FFMA2 per tap and column ----
            float2 P[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 acc = __fmul2_rn(kk[0], wp[j]);
#pragma unroll
                for (int d = 1; d <= 2 * R; ++d) acc = __ffma2_rn(kk[d], wp[j + d], acc);
                P[j] = acc;
            }
            ring[i][0] = make_float2(P[0].x, P[1].x);
            ring[i][1] = make_float2(P[2].x, P[3].x);
            ring[i + 1][0] = make_float2(P[0].y, P[1].y);
            ring[i + 1][1] = make_float2(P[2].y, P[3].y);
            // ---- vertical pass: output rows rr - 2R and rr + 1 - 2R of the strip ----
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int ro = rr + q - 2 * R;
                if (ro < 0) continue;  // warp-uniform: the first 2R source rows only fill the ring
                const int yo = y0t + wi * GW_RG + ro;
                // taps d = 0..2R sit in slots (i + q - 2R + d) mod RS = (i + q + 2 + d) mod RS
                float2 o0 = __fmul2_rn(kk[0], ring[(i + q + 2) % RS][0]);
                float2 o1 = __fmul2_rn(kk[0], ring[(i + q + 2) % RS][1]);
#pragma unroll
                for (int d = 1; d <= 2 * R; ++d) {
                    o0 = __ffma2_rn(kk[d], ring[(i + q + 2 + d) % RS][0], o0);
                    o1 = __ffma2_rn(kk[d], ring[(i + q + 2 + d) % RS][1], o1);
                }
                const bool yborder = (yo < R) || (yo >= h - R);
                if (yborder) {  // warp-uniform; partial-kernel renormalisation (signal_processing.cpp:130-144)
                    float ny = 0.f;
#pragma unroll
                    for (int d = 0; d <= 2 * R; ++d) {
                        const int yyy = yo + d - R;
                        if (yyy >= 0 && yyy < h) ny += k[d];
                    }
                    o0 = make_float2(o0.x / (ny * nx[0]), o0.y / (ny * nx[1]));
                    o1 = make_float2(o1.x / (ny * nx[2]), o1.y / (ny * nx[3]));
                } else {
                    o0 = __fmul2_rn(o0, make_float2(rx[0], rx[1]));
                    o1 = __fmul2_rn(o1, make_float2(rx[2], rx[3]));
                }
                if (x < w && yo < h) st_stream(reinterpret_cast<float4*>(oframe + (size_t)yo * w + x), make_float4(o0.x, o0.y, o1.x, o1.y));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// generic path: any width, any radius up to GAUSS_MAX_RADIUS -- one thread per output pixel
// ------------------------------------------------------------------------------------------------
template <typename TIN>
__global__ void __launch_bounds__(256)
gauss_generic_kernel(const TIN* __restrict__ src, float* __restrict__ dst, int w, int h, long long nframes, GaussTaps taps)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const int r = taps.radius;
    const bool border = (x < r) || (x >= w - r) || (y < r) || (y >= h - r);
    for (long long f = blockIdx.z; f < nframes; f += gridDim.z) {
        const TIN* frame = src + (size_t)f * w * h;
        float acc = 0.f, nxs = 0.f, nys = 0.f;
        for (int dy = -r; dy <= r; ++dy) {  // rows outside the image contribute nothing
            int yy = y + dy;
            if (yy < 0 || yy >= h) continue;
            float ky = taps.k[dy + r];
            nys += ky;
            float rowacc = 0.f;
            for (int dx = -r; dx <= r; ++dx) {
                int xx = x + dx;
                if (xx < 0 || xx >= w) continue;
                rowacc = fmaf(taps.k[dx + r], (float)frame[(size_t)yy * w + xx], rowacc);
            }
            acc = fmaf(ky, rowacc, acc);
        }
        if (border) {
            for (int dx = -r; dx <= r; ++dx) {
                int xx = x + dx;
                if (xx >= 0 && xx < w) nxs += taps.k[dx + r];
            }
            acc = acc / (nxs * nys);
        }
        dst[(size_t)f * w * h + (size_t)y * w + x] = acc;
    }
}

// bad_pixels_correct + gaussian_filter in one pass over the raw movie (see BpFuse).  Returns 1 when the layout cannot
// take the tiled path (the caller then runs the two kernels), 0 / -1 otherwise.
int launch_gaussian_bp_u16(const u16* raw, u16* corrected, float* dst, int w, int h, long long nframes, const GaussTaps& taps,
                           const int* xy_dev, const int* row_off_dev, int clamp_value, cudaStream_t st)
{
    if (nframes <= 0 || w <= 0 || h <= 0) return 0;
    const int r = taps.radius;
    const int tiles_x = (int)ceil_div(w, GT_W), tiles_y = (int)ceil_div(h, GT_H);
    const long long tgrid = nframes * tiles_x * tiles_y;
    if (r > 4 || (w % 8) != 0 || !aligned16(dst) || !aligned16(corrected) || !option_enabled(OPT_GAUSS_TMA) ||
        !tma_compatible(raw, (size_t)w * 2, (size_t)w * h * 2) || nframes > 0x7FFFFFFFLL || tgrid > 0x7FFFFFFFLL)
        return 1;
    BpFuse bf;
    bf.raw = raw;
    bf.xy = xy_dev;
    bf.row_off = row_off_dev;
    bf.corrected = corrected;
    bf.clamp = clamp_value > 0 ? (unsigned)clamp_value & 0xFFFFu : 0u;
    CUtensorMap tmap;
#define RIRB_GTB(RR)                                                                                                              \
    do {                                                                                                                          \
        if (make_movie_tensor_map(&tmap, raw, 2, w, h, nframes, (size_t)w * 2, (size_t)w * h * 2, GtBox<u16>::BW, GT_H + 2 * RR) != 0) \
            return -1;                                                                                                            \
        RIRB_LAUNCH((gauss_tile_kernel<RR, u16, true>), (unsigned)tgrid, GT_WARPS * 32, 0, st, tmap, dst, w, h, tiles_x, tiles_y, taps, bf); \
    } while (0)
    switch (r) {
    case 1: RIRB_GTB(1); break;
    case 2: RIRB_GTB(2); break;
    case 3: RIRB_GTB(3); break;
    default: RIRB_GTB(4); break;
    }
#undef RIRB_GTB
    return 0;
}

template <typename TIN>
static int launch_gaussian(const TIN* src, float* dst, int w, int h, long long nframes, const GaussTaps& taps, cudaStream_t st,
                           size_t src_row = 0, size_t src_frame = 0)
{
    // src_row / src_frame (pixels; 0 = dense): the source is a w x h REGION of larger frames.  Only the TMA-tiled kernels
    // take that (the strides are the tensor map's); anything else answers 1 and the caller filters whole frames.
    if (nframes <= 0 || w <= 0 || h <= 0) return 0;
    const bool region = src_row != 0;
    if (!region) {
        src_row = (size_t)w;
        src_frame = (size_t)w * h;
    }
    const int r = taps.radius;
    const bool fast = (r <= 4) && (w % 4 == 0) && aligned16(dst) && aligned16(src);
    if (!fast) {
        if (region) return 1;
        dim3 block(32, 8);
        dim3 grid((unsigned)ceil_div(w, 32), (unsigned)ceil_div(h, 8), (unsigned)min(nframes, 32768LL));
        RIRB_LAUNCH(gauss_generic_kernel<TIN>, grid, block, 0, st, src, dst, w, h, nframes, taps);
        return 0;
    }
    const bool tma_enabled = option_enabled(OPT_GAUSS_TMA);  // "gauss_tma" = 0 selects the warp-strip kernel (A/B measurements)
    const size_t esz = sizeof(TIN);
    const int tiles_x = (int)ceil_div(w, GT_W), tiles_y = (int)ceil_div(h, GT_H);
    const long long tgrid = nframes * tiles_x * tiles_y;
    if (tma_enabled && tma_compatible(src, src_row * esz, src_frame * esz) && nframes <= 0x7FFFFFFFLL && tgrid <= 0x7FFFFFFFLL) {
        CUtensorMap tmap;
#define RIRB_GT(RR)                                                                                                          \
    do {                                                                                                                     \
        if (make_movie_tensor_map(&tmap, src, (int)esz, w, h, nframes, src_row * esz, src_frame * esz, GtBox<TIN>::BW,       \
                                  GT_H + 2 * RR) != 0)                                                                       \
            return -1;                                                                                                       \
        RIRB_LAUNCH((gauss_tile_kernel<RR, TIN, false>), (unsigned)tgrid, GT_WARPS * 32, 0, st, tmap, dst, w, h, tiles_x, tiles_y, taps, BpFuse{}); \
    } while (0)
        // R >= 3: the row-pair kernel with 32-row strips (gauss_tile_wide_kernel); "gauss_tma" = 2 keeps the first-generation one
        const bool wide = r >= 3 && option_value(OPT_GAUSS_TMA) != 2;
        const int wtiles_y = (int)ceil_div(h, GW_H);
        const long long wgrid = nframes * tiles_x * wtiles_y;
#define RIRB_GW(RR)                                                                                                          \
    do {                                                                                                                     \
        const size_t smem = (size_t)(GW_H + 2 * RR) * GtBox<TIN>::BW * esz;                                                  \
        RIRB_SMEM_ATTR((gauss_tile_wide_kernel<RR, TIN>), smem);                                                             \
        if (make_movie_tensor_map(&tmap, src, (int)esz, w, h, nframes, src_row * esz, src_frame * esz, GtBox<TIN>::BW,       \
                                  GW_H + 2 * RR) != 0)                                                                       \
            return -1;                                                                                                       \
        RIRB_LAUNCH((gauss_tile_wide_kernel<RR, TIN>), (unsigned)wgrid, GT_WARPS * 32, smem, st, tmap, dst, w, h, tiles_x, wtiles_y, taps); \
    } while (0)
        if (wide && wgrid <= 0x7FFFFFFFLL) {
            if (r == 3) RIRB_GW(3);
            else RIRB_GW(4);
            return 0;
        }
#undef RIRB_GW
        switch (r) {
        case 1: RIRB_GT(1); break;
        case 2: RIRB_GT(2); break;
        case 3: RIRB_GT(3); break;
        default: RIRB_GT(4); break;
        }
#undef RIRB_GT
        return 0;
    }
    if (region) return 1;
    const int xstrips = (int)ceil_div(w, GS_STRIP);
    // full-height strips when there are enough frames; otherwise cut rows to fill the GPU
    const long long target = (long long)sm_count() * 16;
    long long ystrips = ceil_div(target, nframes * xstrips);
    const long long max_ystrips = ceil_div(h, 16);
    if (ystrips > max_ystrips) ystrips = max_ystrips;
    if (ystrips < 1) ystrips = 1;
    int rows = (int)ceil_div(h, ystrips);
    rows = (int)ceil_div(rows, GS_ROWS) * GS_ROWS;
    ystrips = ceil_div(h, rows);
    const long long items = nframes * ystrips * xstrips;
    const long long grid = ceil_div(items, GS_WARPS);
    if (grid > 0x7FFFFFFFLL) {
        set_error("gaussian_filter: too many frames in one call (%lld)", nframes);
        return -1;
    }
#define RIRB_GS(RR)                                                                                                         \
    RIRB_LAUNCH((gauss_sep_kernel<RR, TIN>), (unsigned)grid, GS_WARPS * 32, 0, st, src, dst, w, h, nframes, xstrips, (int)ystrips, \
                rows, taps)
    switch (r) {
    case 1: RIRB_GS(1); break;
    case 2: RIRB_GS(2); break;
    case 3: RIRB_GS(3); break;
    default: RIRB_GS(4); break;
    }
#undef RIRB_GS
    return 0;
}

int launch_gaussian_f32(const float* src, float* dst, int w, int h, long long nframes, const GaussTaps& taps, cudaStream_t st)
{
    return launch_gaussian<float>(src, dst, w, h, nframes, taps, st);
}
int launch_gaussian_u16(const u16* src, float* dst, int w, int h, long long nframes, const GaussTaps& taps, cudaStream_t st)
{
    return launch_gaussian<u16>(src, dst, w, h, nframes, taps, st);
}
// A w x h region (top-left pixel at src) of frames whose rows are src_row pixels and whose frames are src_frame pixels apart,
// filtered as an image of its own into dense dst[n][h][w]: a border of the region that is not a border of the frame gets the
// image-border renormalisation, so the caller leaves `radius` pixels of margin there.  1: this layout is not taken.
int launch_gaussian_u16_region(const u16* src, size_t src_row, size_t src_frame, float* dst, int w, int h, long long nframes,
                               const GaussTaps& taps, cudaStream_t st)
{
    return launch_gaussian<u16>(src, dst, w, h, nframes, taps, st, src_row, src_frame);
}
int launch_gaussian_f32_region(const float* src, size_t src_row, size_t src_frame, float* dst, int w, int h, long long nframes,
                               const GaussTaps& taps, cudaStream_t st)
{
    return launch_gaussian<float>(src, dst, w, h, nframes, taps, st, src_row, src_frame);
}

}  // namespace rirb
