// Internal interface of the tcgen05 GEMM (see gemm_sm100.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace icb {

// Fused epilogue description.  For the accumulator tile acc[r, c] (fp32, TMEM):
//   v = acc + bias[c]            (bias_per_row: bias[r])
//   v = gelu_tanh(v)             (act == 1)
//   out_bf16[r*ld_out + c] = bf16(v)                                (if out_bf16)
//   rowss[r*rowss_ld + n_block] = sum_c bf16(v)^2 over this N tile   (if rowss; needs out_bf16)
//   out_f32[r*ld_f32 + c] = v + addend[r*ld_add + c]                 (if out_f32; addend optional)
//   resid[r*ld_res + c] += gate[c] * v                               (if resid; gate optional -> 1)
struct GemmEpilogue {
  const float* bias = nullptr;
  int bias_per_row = 0;
  int act = 0;
  __nv_bfloat16* out_bf16 = nullptr;
  int ld_out = 0;
  float* rowss = nullptr;
  int rowss_ld = 0;
  float* out_f32 = nullptr;
  int ld_f32 = 0;
  const float* addend = nullptr;
  int ld_add = 0;
  float* resid = nullptr;
  int ld_res = 0;
  const float* gate = nullptr;
};

// C[M,N] = A[M,K] * B[N,K]^T, A and B bf16 row-major (K contiguous), fp32 accumulation on tcgen05.
// Returns IC_OK or a negative error code.  N-tile width used for rowss is gemm_block_n(N).
int gemm_bf16_tn(const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb, int M, int N, int K,
                 const GemmEpilogue& ep, cudaStream_t stream);
int gemm_block_n(int N);

}  // namespace icb
