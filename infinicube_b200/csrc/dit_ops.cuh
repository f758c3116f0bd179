// Internal interface of the HBM-bound DiT helper kernels (see dit_ops.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace icb {

// (cos, sin) tables of the 3-axis rotary embedding, fp32 pairs: tab_f [n_f][22], tab_h [n_h][21],
// tab_w [n_w][21] (head_dim 128 -> 64 complex pairs split 22 | 21 | 21).
struct RopeTables {
  const float* tab_f;
  const float* tab_h;
  const float* tab_w;
  int n_f, n_h, n_w;
};

int ln_modulate(const float* x, int ldx, const float* mul, const float* add, int mul_plus_one, __nv_bfloat16* out,
                int ldo, int rows, int D, float eps, cudaStream_t stream);
int rmsnorm_rope(const __nv_bfloat16* src, int ld_src, const float* ss, int ss_ld, int ss_off, int ss_cnt,
                 const float* w, __nv_bfloat16* dst, int ld_dst, int rows, int D, float eps, const RopeTables* rope,
                 int f0, cudaStream_t stream, int group_cols = 0, long long group_stride = 0);
int patchify(const float* lat, __nv_bfloat16* out, int C, int F, int H, int W, int ld_out, int col_off,
             cudaStream_t stream);
int unpatchify_cfg_step(float* lat, const float* vpos, const float* vneg, int C, int F, int H, int W, float cfg,
                        float dsigma, float* v_out, cudaStream_t stream);
int gemv_bf16(const __nv_bfloat16* W, int ldw, const float* x, const float* b, float* y, int N, int K, int act_in,
              int act_out, cudaStream_t stream);
int add_bcast(const float* a, const float* b, float* out, long long n, int period, cudaStream_t stream);
int cast_f32_bf16(const float* in, __nv_bfloat16* out, long long n, cudaStream_t stream);

}  // namespace icb
