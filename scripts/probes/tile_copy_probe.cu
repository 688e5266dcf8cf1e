// scripts/probes/tile_copy_probe.cu -- what does a plain copy achieve with translate's traversal?  (measurement only)
//   mode 0: linear 128-bit copy (the roofline denominator's pattern)
//   mode 1: one CTA per 128-pixel tile column of a frame, 128 threads, thread = 8 pixels x 8 consecutive rows of a 64-row tile,
//           direct LDG.128 / STG.128 (no shared memory, no TMA)
//   mode 2: as 1 but each thread reads the next 16-byte chunk as well (the second, overlapping load of the blend)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tile_copy_probe.cu -o tile_copy_probe ; run: ./tile_copy_probe [w h frames]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void linear_copy(const uint4* __restrict__ a, uint4* __restrict__ b, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(a + i));
        asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(b + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
}

template <int MODE>
__global__ void __launch_bounds__(128) tile_copy(const unsigned short* __restrict__ src, unsigned short* __restrict__ dst, int w, int h)
{
    const int f = blockIdx.y, x0t = blockIdx.x * 128;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cx = lane & 15, half = lane >> 4;
    const int x0 = x0t + 8 * cx;
    if (x0 >= w) return;
    const unsigned short* fr = src + (size_t)f * w * h;
    unsigned short* out = dst + (size_t)f * w * h;
    for (int y0 = 0; y0 < h; y0 += 64) {
        const int r0 = y0 + 16 * warp + 8 * half;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int y = r0 + k;
            if (y >= h) break;
            const uint4* p = reinterpret_cast<const uint4*>(fr + (size_t)y * w + x0);
            uint4 v = __ldg(p);
            if (MODE == 2 && x0 + 8 < w) {
                const uint4 v2 = __ldg(p + 1);
                v.x ^= v2.x & 0;  // keep the load alive
                asm volatile("" ::"r"(v2.y), "r"(v2.z), "r"(v2.w));
            }
            asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(out + (size_t)y * w + x0), "r"(v.x), "r"(v.y), "r"(v.z),
                         "r"(v.w)
                         : "memory");
        }
    }
}

// mode 3: one CTA per (tile column, 64-row tile, frame): the traversal of mode 1 cut into short CTAs that are launched in
//         memory order (column fastest, then tile row, then frame)
// mode 4: one CTA per (64-row tile, frame), covering the full row width: warp k = tile column k (blockDim = 32 * columns)
template <int MODE>
__global__ void tile_copy_short(const unsigned short* __restrict__ src, unsigned short* __restrict__ dst, int w, int h)
{
    const int f = blockIdx.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cx = lane & 15, half = lane >> 4;
    int x0, r0;
    if (MODE == 3) {
        x0 = blockIdx.x * 128 + 8 * cx;
        r0 = blockIdx.y * 64 + 16 * warp + 8 * half;
    } else {  // warp = tile column; the CTA's 64 rows are walked in 4 bands of 16
        x0 = warp * 128 + 8 * cx;
        r0 = blockIdx.y * 64 + 8 * half;
    }
    if (x0 >= w) return;
    const unsigned short* fr = src + (size_t)f * w * h;
    unsigned short* out = dst + (size_t)f * w * h;
    for (int b = 0; b < (MODE == 3 ? 1 : 4); ++b) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int y = r0 + 16 * b + k;
            if (y >= h) break;
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(fr + (size_t)y * w + x0));
            asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(out + (size_t)y * w + x0), "r"(v.x), "r"(v.y), "r"(v.z),
                         "r"(v.w)
                         : "memory");
        }
    }
}

// mode 5+: as mode 3 with T = 2, 4, 8 tiles (of 64 rows) per CTA, walked downwards; launch order: column fastest, then
//          tile group, then frame
__global__ void tile_copy_multi(const unsigned short* __restrict__ src, unsigned short* __restrict__ dst, int w, int h, int T)
{
    const int f = blockIdx.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cx = lane & 15, half = lane >> 4;
    const int x0 = blockIdx.x * 128 + 8 * cx;
    if (x0 >= w) return;
    const unsigned short* fr = src + (size_t)f * w * h;
    unsigned short* out = dst + (size_t)f * w * h;
    for (int t = 0; t < T; ++t) {
        const int r0 = (blockIdx.y * T + t) * 64 + 16 * warp + 8 * half;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int y = r0 + k;
            if (y >= h) break;
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(fr + (size_t)y * w + x0));
            asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(out + (size_t)y * w + x0), "r"(v.x), "r"(v.y), "r"(v.z),
                         "r"(v.w)
                         : "memory");
        }
    }
}

int main(int argc, char** argv)
{
    const int w = argc > 1 ? atoi(argv[1]) : 640, h = argc > 2 ? atoi(argv[2]) : 512, n = argc > 3 ? atoi(argv[3]) : 4000;
    const size_t px = (size_t)w * h * n;
    unsigned short *a, *b;
    cudaMalloc(&a, px * 2);
    cudaMalloc(&b, px * 2);
    cudaMemset(a, 1, px * 2);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int mode = 0; mode < 8; ++mode) {
        float best = 1e9f;
        for (int rep = 0; rep < 6; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0)
                linear_copy<<<148 * 16, 512>>>((const uint4*)a, (uint4*)b, px / 8);
            else if (mode == 1)
                tile_copy<1><<<dim3((w + 127) / 128, n), 128>>>(a, b, w, h);
            else if (mode == 2)
                tile_copy<2><<<dim3((w + 127) / 128, n), 128>>>(a, b, w, h);
            else if (mode == 3)
                tile_copy_short<3><<<dim3((w + 127) / 128, (h + 63) / 64, n), 128>>>(a, b, w, h);
            else if (mode == 4)
                tile_copy_short<4><<<dim3(1, (h + 63) / 64, n), 32 * ((w + 127) / 128)>>>(a, b, w, h);
            else {
                const int T = 1 << (mode - 4);
                tile_copy_multi<<<dim3((w + 127) / 128, (h + 64 * T - 1) / (64 * T), n), 128>>>(a, b, w, h, T);
            }
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep >= 2 && ms < best) best = ms;
        }
        printf("{\"probe\": \"tile_copy\", \"mode\": %d, \"frame\": [%d, %d], \"frames\": %d, \"ms\": %.4f, \"gbs\": %.1f, \"err\": \"%s\"}\n", mode, w, h, n,
               best, 4.0 * px / (best * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
