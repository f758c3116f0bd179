// Internal interface of the implicit-GEMM causal convolution (see conv_sm100.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace icb {

struct ConvTap {
  int dt, dh, dw;  // input offset of this tap relative to the output position
};

// out[t,h,w,:] = bias + sum_taps W[tap] * in[t+dt, h+dh, w+dw, :]  (+ resid[t,h,w,:]); out-of-range input = 0.
// in: bf16 channels-last [Tin, Hin, Win, Cin]; out / resid: bf16 channels-last [T, H, W, Cout] with row pitch ld_out;
// weight: bf16 [Cout, ntaps*Cin] (K index = tap*Cin + c); bias fp32 [Cout] or null.  Cin % 32 == 0.
int conv_igemm(const __nv_bfloat16* in, int Tin, int Hin, int Win, int Cin, const __nv_bfloat16* weight, const float* bias,
               const ConvTap* taps, int ntaps, __nv_bfloat16* out, int T, int H, int W, int Cout, int ld_out,
               const __nv_bfloat16* resid, int ld_resid, cudaStream_t stream);

}  // namespace icb
