// Internal interface of the tcgen05 flash-attention forward (see fmha_sm100.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace icb {

// O[Sq, n_heads*128] = softmax(Q K^T * softmax_scale) V, per head, non-causal, no mask.
// K is [n_seg][seg_len, ldk] and V^T is [n_seg][n_heads*128, ldvt]; segment strides in elements
// (ignored when n_seg == 1).  seg_len, ld* and strides must be multiples of 8 elements.
int fmha_fwd(const __nv_bfloat16* Q, int ldq, const __nv_bfloat16* K, int ldk, long long k_seg_stride,
             const __nv_bfloat16* VT, int ldvt, long long vt_seg_stride, __nv_bfloat16* O, int ldo, int Sq,
             int seg_len, int n_seg, int n_heads, float softmax_scale, cudaStream_t stream);

}  // namespace icb
