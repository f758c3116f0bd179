// Host-side helpers shared by the launchers: CUDA error plumbing and TMA descriptor (CUtensorMap) encoding.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "common.cuh"

namespace icb {

#define ICB_CUDA_CHECK(expr)                                                                       \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      fprintf(stderr, "[icb] CUDA error %s at %s:%d: %s\n", #expr, __FILE__, __LINE__,             \
              cudaGetErrorString(_e));                                                             \
      return IC_ERR_CUDA;                                                                          \
    }                                                                                              \
  } while (0)

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

// The driver entry point is resolved at run time so the library links against cudart only
// (there is no libcuda on the build box).
inline PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

// bf16 tensor, innermost dimension contiguous, 128-byte swizzle, zero fill out of bounds.
// dims[0] is the innermost extent (elements); strides_bytes[i] is the byte stride of dims[i+1].
inline int make_tmap_typed(CUtensorMap* out, CUtensorMapDataType dtype, const void* base, int rank,
                           const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                           int swizzle_bytes) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) {
    fprintf(stderr, "[icb] cuTensorMapEncodeTiled unavailable\n");
    return IC_ERR_CUDA;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    estr[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CUresult r = enc(out, dtype, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[icb] cuTensorMapEncodeTiled failed: %d (rank %d dims %llu %llu box %u %u base %p)\n", (int)r,
            rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0],
            rank > 1 ? box[1] : 0, base);
    return IC_ERR_CUDA;
  }
  return IC_OK;
}

inline int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes = 128) {
  return make_tmap_typed(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box, swizzle_bytes);
}

// fp32 tensor (destination of the TMA reduce-add epilogue)
inline int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes = 128) {
  return make_tmap_typed(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box, swizzle_bytes);
}

inline int num_sms() {
  static int n = 0;
  if (n) return n;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}

}  // namespace icb
