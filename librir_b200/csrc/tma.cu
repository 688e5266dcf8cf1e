// librir_b200/csrc/tma.cu -- host side of tma.cuh: encode a movie as a 3-D TMA tensor map.
// cuTensorMapEncodeTiled is a driver-API function; it is resolved at run time through
// cudaGetDriverEntryPoint so that the library links against the (static) runtime only.
#include <mutex>

#include "common.cuh"
#include "tma.cuh"

namespace rirb {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn()
{
    static std::once_flag once;
    static EncodeTiledFn fn = nullptr;
    std::call_once(once, []() {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    });
    return fn;
}

int make_movie_tensor_map(CUtensorMap* map, const void* base, int elem_bytes, int w, int h, long long nframes,
                          size_t row_stride_bytes, size_t frame_stride_bytes, int box_w, int box_h)
{
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return -1;
    }
    if ((elem_bytes != 1 && elem_bytes != 2 && elem_bytes != 4) || !tma_compatible(base, row_stride_bytes, frame_stride_bytes) || box_w > 256 ||
        box_h > 256 || (box_w * elem_bytes) % 16 != 0 || nframes <= 0) {
        set_error("tensor map: unsupported layout");
        return -1;
    }
    const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)nframes};
    const cuuint64_t strides[2] = {(cuuint64_t)row_stride_bytes, (cuuint64_t)frame_stride_bytes};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    // a one-frame movie must not carry a zero / unaligned outer stride
    const CUresult r = enc(map, elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : (elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32), 3,
                           const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (w=%d h=%d n=%lld box=%dx%d)", (int)r, w, h, nframes, box_w, box_h);
        return -1;
    }
    return 0;
}

}  // namespace rirb
