// Host-side creation of TMA tensor maps.  cuTensorMapEncodeTiled is a driver-API entry point; it is resolved at
// run time through cudaGetDriverEntryPoint so the library keeps linking against the CUDA runtime only.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace dcgru {

// rank <= 5; dims[i] elements, strides[i] bytes for i >= 1 (dimension 0 is contiguous), box[i] elements
cudaError_t make_tmap_f32(CUtensorMap* out, const void* base, int rank, const unsigned long long* dims,
                          const unsigned long long* strides_bytes, const unsigned* box, CUtensorMapSwizzle swizzle);

// same for any element type (fp16 operand images of the second-generation kernels)
cudaError_t make_tmap(CUtensorMap* out, CUtensorMapDataType dtype, const void* base, int rank, const unsigned long long* dims,
                      const unsigned long long* strides_bytes, const unsigned* box, CUtensorMapSwizzle swizzle);

}  // namespace dcgru
