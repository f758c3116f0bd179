// extern "C" surface for the granular ops (declared in include/infinicube_b200.h).
#include "../../include/infinicube_b200.h"
#include "dit_ops.cuh"
#include "fmha_sm100.cuh"
#include "gemm_sm100.cuh"
#include "host_util.h"
#include "t5_ops.cuh"

using namespace icb;

extern "C" {

int ic_version(void) { return 100; }

const char* ic_error_string(int code) {
  switch (code) {
    case IC_OK: return "ok";
    case IC_ERR_INVALID: return "invalid argument";
    case IC_ERR_CUDA: return "CUDA error";
    case IC_ERR_NO_DEVICE: return "no sm_100 (B200) device: this library has no fallback path";
    case IC_ERR_UNSUPPORTED: return "unsupported configuration";
    case IC_ERR_NCCL: return "NCCL error";
    default: return "unknown error";
  }
}

int ic_device_check(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return IC_ERR_NO_DEVICE;
  }
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return IC_ERR_NO_DEVICE;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  return major == 10 ? IC_OK : IC_ERR_NO_DEVICE;
}

int ic_gemm_block_n(int N) { return gemm_block_n(N); }

int ic_gemm_bf16(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const ic_gemm_epilogue* e,
                 void* stream) {
  if (!A || !B || !e) return IC_ERR_INVALID;
  int r = ic_device_check();
  if (r != IC_OK) return r;
  GemmEpilogue ep;
  ep.bias = e->bias;
  ep.bias_per_row = e->bias_per_row;
  ep.act = e->act;
  ep.out_bf16 = static_cast<__nv_bfloat16*>(e->out_bf16);
  ep.ld_out = e->ld_out;
  ep.rowss = e->rowss;
  ep.rowss_ld = e->rowss_ld;
  ep.out_f32 = e->out_f32;
  ep.ld_f32 = e->ld_f32;
  ep.addend = e->addend;
  ep.ld_add = e->ld_add;
  ep.resid = e->resid;
  ep.ld_res = e->ld_res;
  ep.gate = e->gate;
  return gemm_bf16_tn(static_cast<const __nv_bfloat16*>(A), lda, static_cast<const __nv_bfloat16*>(B), ldb, M, N, K,
                      ep, static_cast<cudaStream_t>(stream));
}

int ic_fmha_fwd(const void* Q, int ldq, const void* K, int ldk, long long k_seg_stride, const void* VT, int ldvt,
                long long vt_seg_stride, void* O, int ldo, int Sq, int seg_len, int n_seg, int n_heads,
                float softmax_scale, void* stream) {
  if (!Q || !K || !VT || !O) return IC_ERR_INVALID;
  int r = ic_device_check();
  if (r != IC_OK) return r;
  return fmha_fwd(static_cast<const __nv_bfloat16*>(Q), ldq, static_cast<const __nv_bfloat16*>(K), ldk, k_seg_stride,
                  static_cast<const __nv_bfloat16*>(VT), ldvt, vt_seg_stride, static_cast<__nv_bfloat16*>(O), ldo, Sq,
                  seg_len, n_seg, n_heads, softmax_scale, static_cast<cudaStream_t>(stream));
}

int ic_ln_modulate(const float* x, int ldx, const float* mul, const float* add, int mul_plus_one, void* out, int ldo,
                   int rows, int D, float eps, void* stream) {
  if (!x || !mul || !add || !out) return IC_ERR_INVALID;
  return ln_modulate(x, ldx, mul, add, mul_plus_one, static_cast<__nv_bfloat16*>(out), ldo, rows, D, eps,
                     static_cast<cudaStream_t>(stream));
}

int ic_rmsnorm_rope(const void* src, int ld_src, const float* rowss, int ss_ld, int ss_off, int ss_cnt,
                    const float* weight, void* dst, int ld_dst, int rows, int D, float eps, const float* tab_f,
                    const float* tab_h, const float* tab_w, int n_f, int n_h, int n_w, int frame0, void* stream) {
  if (!src || !rowss || !weight || !dst) return IC_ERR_INVALID;
  RopeTables rt{tab_f, tab_h, tab_w, n_f, n_h, n_w};
  return rmsnorm_rope(static_cast<const __nv_bfloat16*>(src), ld_src, rowss, ss_ld, ss_off, ss_cnt, weight,
                      static_cast<__nv_bfloat16*>(dst), ld_dst, rows, D, eps, tab_f ? &rt : nullptr, frame0,
                      static_cast<cudaStream_t>(stream));
}

int ic_patchify(const float* latents, void* out, int C, int F, int H, int W, int ld_out, int col_off, void* stream) {
  if (!latents || !out) return IC_ERR_INVALID;
  return patchify(latents, static_cast<__nv_bfloat16*>(out), C, F, H, W, ld_out, col_off,
                  static_cast<cudaStream_t>(stream));
}

int ic_unpatchify_cfg_step(float* latents, const float* head_pos, const float* head_neg, int C, int F, int H, int W,
                           float cfg_scale, float dsigma, float* v_out, void* stream) {
  if (!head_pos || (!latents && !v_out)) return IC_ERR_INVALID;
  return unpatchify_cfg_step(latents, head_pos, head_neg, C, F, H, W, cfg_scale, dsigma, v_out,
                             static_cast<cudaStream_t>(stream));
}

int ic_t5_embed(const int* ids, int L, const void* table_bf16, int vocab, int D, float* x, int ldx, void* stream) {
  if (!ids || !table_bf16 || !x) return IC_ERR_INVALID;
  int r = ic_device_check();
  if (r != IC_OK) return r;
  return t5_embed(ids, L, static_cast<const __nv_bfloat16*>(table_bf16), vocab, D, x, ldx,
                  static_cast<cudaStream_t>(stream));
}

int ic_t5_rmsnorm(const float* x, int ldx, const float* weight, void* out_bf16, int ldo, int rows, int D, float eps,
                  int zero_from_row, void* stream) {
  if (!x || !weight || !out_bf16) return IC_ERR_INVALID;
  int r = ic_device_check();
  if (r != IC_OK) return r;
  return t5_rmsnorm(x, ldx, weight, static_cast<__nv_bfloat16*>(out_bf16), ldo, rows, D, eps, zero_from_row,
                    static_cast<cudaStream_t>(stream));
}

int ic_t5_attention(const void* q, const void* k, const void* v, int ld, const float* bias_by_offset,
                    const unsigned char* key_mask, void* out, int ldo, int L, int n_heads, void* stream) {
  if (!q || !k || !v || !bias_by_offset || !out) return IC_ERR_INVALID;
  int r = ic_device_check();
  if (r != IC_OK) return r;
  return t5_attention(static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(k),
                      static_cast<const __nv_bfloat16*>(v), ld, bias_by_offset, key_mask,
                      static_cast<__nv_bfloat16*>(out), ldo, L, n_heads, static_cast<cudaStream_t>(stream));
}

int ic_mul_bf16(const void* a, const void* b, void* out, long long n, void* stream) {
  if (!a || !b || !out) return IC_ERR_INVALID;
  int r = ic_device_check();
  if (r != IC_OK) return r;
  return mul_bf16(static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(b),
                  static_cast<__nv_bfloat16*>(out), n, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
