// Internal interface of the tcgen05 flash-attention forward (see fmha_sm100.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace icb {

// O[Sq, n_heads*128] = softmax(Q K^T * softmax_scale) V, per head, non-causal, no mask.
// K is [n_seg][seg_len, ldk] and V^T is [n_seg][n_heads*128, ldvt]; segment strides in elements
// (ignored when n_seg == 1).  seg_len, ld* and strides must be multiples of 8 elements.
// Peer-memory exchange (optional): with seg_ready != nullptr the segments are consumed in ring order starting at
// seg_first (the local one) and segment s != seg_first is read only once seg_ready[s] (device memory, written by
// the peer that owns s) has reached `epoch`.
int fmha_fwd(const __nv_bfloat16* Q, int ldq, const __nv_bfloat16* K, int ldk, long long k_seg_stride,
             const __nv_bfloat16* VT, int ldvt, long long vt_seg_stride, __nv_bfloat16* O, int ldo, int Sq,
             int seg_len, int n_seg, int n_heads, float softmax_scale, cudaStream_t stream, int seg_first = 0,
             const unsigned* seg_ready = nullptr, unsigned epoch = 0);

}  // namespace icb
