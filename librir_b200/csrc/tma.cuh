// librir_b200/csrc/tma.cuh -- Tensor Memory Accelerator plumbing for the tiled kernels (sm_100a).
//
// A movie is described to the TMA unit as a 3-D tensor (x, y, frame) of 2- or 4-byte elements.
// One elected thread asks for a box at signed element coordinates (negative / past-the-edge
// parts are zero-filled by the hardware) and the box lands densely in shared memory; completion
// is signalled on an mbarrier.  Two things on this path are exactly that access pattern:
//   * translate: the source window of a destination tile starts at (x0 + floor(-dx), y0 + floor(-dy)),
//     a position that differs per frame and can hang over any image edge;
//   * gaussian: tiles need a halo of `radius` pixels whose out-of-image taps count as zero, which
//     is the hardware's out-of-bounds fill.
// Measured constraint (scripts/probes/tma_probe.cu on B200): the box's INNERMOST start coordinate
// must fall on a 16-byte boundary (x % 8 == 0 for uint16, x % 4 == 0 for float32), otherwise the
// load traps ("illegal instruction"); negative and past-the-edge coordinates are fine in every
// dimension.  The kernels therefore round the box origin down and keep the residual in registers.
// SASS: UTMALDG (cp.async.bulk.tensor), SYNCS (mbarrier).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace rirb {

// ---- host: tensor-map encoding through the driver entry point (no link against libcuda) ---------
// elem_bytes 1 (uint8), 2 (uint16) or 4 (float32).  Strides in BYTES, multiples of 16; base 16-byte aligned;
// box_w * elem_bytes a multiple of 16, box dims <= 256.  Returns 0 / -1 (set_error).
int make_movie_tensor_map(CUtensorMap* map, const void* base, int elem_bytes, int w, int h, long long nframes,
                          size_t row_stride_bytes, size_t frame_stride_bytes, int box_w, int box_h);
// true if (base, strides) can be described to TMA at all
static inline bool tma_compatible(const void* base, size_t row_stride_bytes, size_t frame_stride_bytes)
{
    return ((reinterpret_cast<uintptr_t>(base) & 15u) == 0) && (row_stride_bytes % 16 == 0) && (frame_stride_bytes % 16 == 0) &&
           row_stride_bytes < (1ull << 40) && frame_stride_bytes < (1ull << 40);
}

#ifdef __CUDACC__
// ---- device -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrivals) : "memory");
}
// make the initialised barrier visible to the async (TMA) proxy
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity)
{
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}"
        : "=r"(ok)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}
// box at element coordinates (x, y, frame), x on a 16-byte boundary -> dense [box_h][box_w] at `dst` (128-byte aligned)
__device__ __forceinline__ void tma_load_box(void* dst, const CUtensorMap* map, unsigned long long* bar, int x, int y, int frame)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                     smem_addr(dst)),
                 "l"(map), "r"(smem_addr(bar)), "r"(x), "r"(y), "r"(frame)
                 : "memory");
}
// the same box, only as far as L2 (no shared-memory destination, no completion): covers the DRAM part of the latency for
// tiles that have no free stage yet
__device__ __forceinline__ void tma_prefetch_box(const CUtensorMap* map, int x, int y, int frame)
{
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(map), "r"(x), "r"(y), "r"(frame) : "memory");
}
#endif  // __CUDACC__

}  // namespace rirb
