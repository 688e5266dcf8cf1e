// scripts/probes/tma_probe.cu -- stand-alone check of the TMA box load used by translate/gaussian.
// usage: tma_probe <box_w> <box_h> <x> <y> <fence:0|1> <w> <h> <n>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("FAIL %s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int FENCE>
__global__ void probe(const __grid_constant__ CUtensorMap tmap, unsigned short* out, int box_w, int box_h, int x, int y, int f)
{
    extern __shared__ __align__(128) unsigned short tile[];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&bar)), "r"(1) : "memory");
        if (FENCE == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        else asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&bar)), "r"(box_w * box_h * 2) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                         smem_addr(tile)), "l"(&tmap), "r"(smem_addr(&bar)), "r"(x), "r"(y), "r"(f) : "memory");
    }
    unsigned ok = 0;
    while (!ok)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_addr(&bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < box_w * box_h; i += blockDim.x) out[i] = tile[i];
}

int main(int argc, char** argv)
{
    int box_w = argc > 1 ? atoi(argv[1]) : 136, box_h = argc > 2 ? atoi(argv[2]) : 34;
    int x = argc > 3 ? atoi(argv[3]) : -2, y = argc > 4 ? atoi(argv[4]) : 2, fence = argc > 5 ? atoi(argv[5]) : 0;
    int w = argc > 6 ? atoi(argv[6]) : 128, h = argc > 7 ? atoi(argv[7]) : 96, n = argc > 8 ? atoi(argv[8]) : 1;
    printf("box %dx%d at (%d,%d) fence=%d tensor %dx%dx%d: ", box_w, box_h, x, y, fence, w, h, n);
    std::vector<unsigned short> host((size_t)w * h * n);
    for (size_t i = 0; i < host.size(); ++i) host[i] = (unsigned short)(i * 7 + 1);
    unsigned short *d, *o;
    CK(cudaMalloc(&d, host.size() * 2));
    CK(cudaMalloc(&o, (size_t)box_w * box_h * 2));
    CK(cudaMemcpy(d, host.data(), host.size() * 2, cudaMemcpyHostToDevice));
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    CUtensorMap map;
    cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t strides[2] = {(cuuint64_t)w * 2, (cuuint64_t)w * h * 2};
    cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = ((Enc)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    size_t smem = (size_t)box_w * box_h * 2;
    if (fence == 0) probe<0><<<1, 256, smem>>>(map, o, box_w, box_h, x, y, 0);
    else probe<1><<<1, 256, smem>>>(map, o, box_w, box_h, x, y, 0);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<unsigned short> got((size_t)box_w * box_h);
    CK(cudaMemcpy(got.data(), o, got.size() * 2, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (int j = 0; j < box_h; ++j)
        for (int i = 0; i < box_w; ++i) {
            int sx = x + i, sy = y + j;
            unsigned short want = (sx >= 0 && sx < w && sy >= 0 && sy < h) ? host[(size_t)sy * w + sx] : 0;
            bad += got[(size_t)j * box_w + i] != want;
        }
    printf("%s (%zu mismatches)\n", bad ? "WRONG" : "OK", bad);
    return bad != 0;
}
