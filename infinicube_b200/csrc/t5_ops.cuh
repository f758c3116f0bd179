// Internal interface of the umT5 prompt-encoder helper kernels (see t5_ops.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace icb {

int t5_embed(const int* ids, int L, const __nv_bfloat16* table, int vocab, int D, float* x, int ldx,
             cudaStream_t stream);
int t5_rmsnorm(const float* x, int ldx, const float* w, __nv_bfloat16* out, int ldo, int rows, int D, float eps,
               int zero_from_row, cudaStream_t stream);
// q / k / v: bf16 [L, ld] with head h in columns [64h, 64h + 64); bias_by_offset fp32 [n_heads][2L - 1]
int t5_attention(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, int ld,
                 const float* bias_by_offset, const unsigned char* key_mask, __nv_bfloat16* out, int ldo, int L,
                 int n_heads, cudaStream_t stream);
int mul_bf16(const __nv_bfloat16* a, const __nv_bfloat16* b, __nv_bfloat16* out, long long n, cudaStream_t stream);

}  // namespace icb
