// tma.cu -- host side of tma.cuh: tensor-map encoding through the runtime's driver entry point.
#include <mutex>

#include "tma.cuh"

namespace admm {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static std::once_flag once;
    static EncodeTiledFn fn = nullptr;
    std::call_once(once, [] {
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

bool tma_encode_f32(CUtensorMap* map, const void* base, int rank, const unsigned long long* dims,
                    const unsigned long long* strides_bytes, const unsigned* box) {
    EncodeTiledFn fn = encode_fn();
    if (!fn || !map || !base || rank < 2 || rank > 3) return false;
    if (reinterpret_cast<uintptr_t>(base) & 15) return false;
    cuuint64_t gd[3], gs[2];
    cuuint32_t bx[3], es[3] = {1, 1, 1};
    for (int k = 0; k < rank; ++k) {
        if (dims[k] == 0 || box[k] == 0 || box[k] > 256) return false;
        gd[k] = dims[k];
        bx[k] = box[k];
    }
    if ((box[0] * 4u) & 15u) return false;
    for (int k = 0; k + 1 < rank; ++k) {
        if (strides_bytes[k] & 15ULL) return false;
        gs[k] = strides_bytes[k];
    }
    CUtensorMap tmp;
    const CUresult r = fn(&tmp, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    *map = tmp;
    return true;
}

}  // namespace admm
