#include "tmap.cuh"

namespace dcgru {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// resolved once, thread-safely (C++11 magic static): forward runs on the training thread, backward on autograd's
// device thread, and several GPUs' threads may arrive together
static EncodeTiledFn resolve_encode_fn() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
        return reinterpret_cast<EncodeTiledFn>(p);
    return nullptr;
}
static EncodeTiledFn encode_fn() {
    static const EncodeTiledFn fn = resolve_encode_fn();
    return fn;
}

cudaError_t make_tmap_f32(CUtensorMap* out, const void* base, int rank, const unsigned long long* dims,
                          const unsigned long long* strides_bytes, const unsigned* box, CUtensorMapSwizzle swizzle) {
    return make_tmap(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box, swizzle);
}

cudaError_t make_tmap(CUtensorMap* out, CUtensorMapDataType dtype, const void* base, int rank, const unsigned long long* dims,
                      const unsigned long long* strides_bytes, const unsigned* box, CUtensorMapSwizzle swizzle) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return cudaErrorNotSupported;
    cuuint64_t gd[5], gs[5];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 1; i < rank; ++i) gs[i - 1] = strides_bytes[i];
    CUresult r = fn(out, dtype, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

}  // namespace dcgru
