// HBM-bound helper kernels of the Wan2.1 DiT forward: LayerNorm+AdaLN modulation, full-width RMSNorm
// + 3-axis RoPE, patchify / unpatchify+CFG+Euler, time-embedding GEMVs.  Each replaces a handful of
// unfused PyTorch elementwise kernels in the reference stack.
#include "dit_ops.cuh"

#include <stdlib.h>
#include "host_util.h"

namespace icb {

namespace {

// ------------------------------------------------------------------------------------------------
// LayerNorm (no affine) + y = xhat * mul + add  ->  bf16.   mul_plus_one: mul := 1 + mul (AdaLN scale)
// one 128-thread block per row; row cached in registers (float4 x VPT)
// ------------------------------------------------------------------------------------------------
template <int VPT>
__global__ void __launch_bounds__(128)
ln_modulate_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ mul,
                   const float* __restrict__ add, int mul_plus_one, __nv_bfloat16* __restrict__ out, int ldo,
                   int D, float eps) {
  __shared__ float red[4];
  const int row = blockIdx.x;
  const int nvec = D >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * ldx);
  float4 v[VPT];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int idx = threadIdx.x + i * 128;
    v[i] = idx < nvec ? xr[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  const float mean = (red[0] + red[1] + red[2] + red[3]) / static_cast<float>(D);
  __syncthreads();
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int idx = threadIdx.x + i * 128;
    if (idx < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
  q = warp_sum(q);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
  __syncthreads();
  const float rstd = rsqrtf((red[0] + red[1] + red[2] + red[3]) / static_cast<float>(D) + eps);
  const float one = mul_plus_one ? 1.f : 0.f;
  uint2* orow = reinterpret_cast<uint2*>(out + static_cast<size_t>(row) * ldo);
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int idx = threadIdx.x + i * 128;
    if (idx < nvec) {
      const float4 m = __ldg(reinterpret_cast<const float4*>(mul) + idx);
      const float4 a = __ldg(reinterpret_cast<const float4*>(add) + idx);
      const float y0 = (v[i].x - mean) * rstd * (one + m.x) + a.x;
      const float y1 = (v[i].y - mean) * rstd * (one + m.y) + a.y;
      const float y2 = (v[i].z - mean) * rstd * (one + m.z) + a.z;
      const float y3 = (v[i].w - mean) * rstd * (one + m.w) + a.w;
      orow[idx] = make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// RMSNorm over the full model width (sum of squares arrives as per-N-tile partials from the GEMM
// epilogue) followed by the 3-axis rotary embedding on adjacent pairs.  bf16 in, bf16 out.
// ------------------------------------------------------------------------------------------------
struct RopeGeom {
  const float2* tab_f;  // [n_f][22] (cos, sin)
  const float2* tab_h;  // [n_h][21]
  const float2* tab_w;  // [n_w][21]
  int n_h, n_w;         // patch grid height / width
  int f0;               // global frame index of local token 0
};

constexpr int RR_ROWS = 4;  // token rows per block: the weight chunk is loaded once, the row loads overlap

__global__ void __launch_bounds__(192)
rmsnorm_rope_kernel(const __nv_bfloat16* __restrict__ src, int ld_src, const float* __restrict__ ss, int ss_ld,
                    int ss_off, int ss_cnt, const float* __restrict__ w, __nv_bfloat16* __restrict__ dst,
                    int ld_dst, int group_cols, long long group_stride, int D, float eps, int use_rope, RopeGeom g,
                    int rows) {
  const int row0 = blockIdx.x * RR_ROWS;
  const int lane = threadIdx.x & 31;
  // per-row 1/rms from the GEMM epilogue's partial sums of squares: lane i fetches partial i, butterfly-reduce
  float rstd[RR_ROWS];
  {
    float part[RR_ROWS];
#pragma unroll
    for (int r = 0; r < RR_ROWS; ++r) {
      part[r] = 0.f;
      const int row = min(row0 + r, rows - 1);
      for (int i = lane; i < ss_cnt; i += 32) part[r] += ss[static_cast<size_t>(row) * ss_ld + ss_off + i];
    }
#pragma unroll
    for (int r = 0; r < RR_ROWS; ++r) {
      float t = part[r];
#pragma unroll
      for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      rstd[r] = rsqrtf(t / static_cast<float>(D) + eps);
    }
  }
  const int hw = g.n_h * g.n_w;
  for (int v = threadIdx.x; v < (D >> 3); v += blockDim.x) {
    uint4 in[RR_ROWS];
#pragma unroll
    for (int r = 0; r < RR_ROWS; ++r) {
      const int row = min(row0 + r, rows - 1);
      in[r] = reinterpret_cast<const uint4*>(src + static_cast<size_t>(row) * ld_src)[v];
    }
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w) + 2 * v);
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(w) + 2 * v + 1);
    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    // destination may be split into column groups (head groups of the multi-GPU gather buffer)
    const int col = v * 8;
    const int grp = col / group_cols;
    const int pair0 = (col & 127) >> 1;
#pragma unroll
    for (int r = 0; r < RR_ROWS; ++r) {
      const int row = row0 + r;
      if (row >= rows) break;
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&in[r]);
      int pf = 0, phh = 0, pww = 0;
      if (use_rope) {
        pf = g.f0 + row / hw;
        const int rem = row % hw;
        phh = rem / g.n_w;
        pww = rem % g.n_w;
      }
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(h2[j]);
        float a = f.x * rstd[r] * wv[2 * j];
        float b = f.y * rstd[r] * wv[2 * j + 1];
        if (use_rope) {
          const int pi = pair0 + j;
          float2 cs;
          if (pi < 22)
            cs = __ldg(g.tab_f + pf * 22 + pi);
          else if (pi < 43)
            cs = __ldg(g.tab_h + phh * 21 + (pi - 22));
          else
            cs = __ldg(g.tab_w + pww * 21 + (pi - 43));
          const float ra = a * cs.x - b * cs.y;
          const float rb = a * cs.y + b * cs.x;
          a = ra;
          b = rb;
        }
        o[j] = pack_bf16x2(a, b);
      }
      *reinterpret_cast<uint4*>(dst + grp * group_stride + static_cast<size_t>(row) * ld_dst + (col - grp * group_cols)) =
          make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// v2 (the default since round 2: the full DiT / pipeline GPU suites pass with it and the step gains ~3 ms;
// ICB_RMSROPE_V2=0 selects v1 for A/B): same contract as rmsnorm_rope_kernel.  ncu shows v1 at 2.8 TB/s (83 us for 230 MB): it is instruction-bound, not
// HBM-bound - every thread redoes two integer div/mod pairs and four branchy table look-ups per row.
// Here the per-row work that does not depend on the column is done once per block: threads < rows
// compute 1/rms and the (frame, y, x) position of their row, the block stages the 64 (cos, sin) pairs of
// each row in shared memory (heads share them), and the streaming loop is LDG.128 + 2 LDS.128 +
// ~35 arithmetic instructions + STG.128 per 8 elements.
// ------------------------------------------------------------------------------------------------
constexpr int RR2_ROWS = 8;

__global__ void __launch_bounds__(192)
rmsnorm_rope_v2_kernel(const __nv_bfloat16* __restrict__ src, int ld_src, const float* __restrict__ ss, int ss_ld,
                       int ss_off, int ss_cnt, const float* __restrict__ w, __nv_bfloat16* __restrict__ dst,
                       int ld_dst, int group_cols, long long group_stride, int D, float eps, int use_rope, RopeGeom g,
                       int rows) {
  __shared__ float s_rstd[RR2_ROWS];
  __shared__ int s_pos[RR2_ROWS][3];
  __shared__ __align__(16) float2 s_cs[RR2_ROWS][64];
  const int row0 = blockIdx.x * RR2_ROWS;
  const int nrows = min(RR2_ROWS, rows - row0);
  if (threadIdx.x < nrows) {
    const int row = row0 + threadIdx.x;
    float t = 0.f;
    for (int i = 0; i < ss_cnt; ++i) t += ss[static_cast<size_t>(row) * ss_ld + ss_off + i];
    s_rstd[threadIdx.x] = rsqrtf(t / static_cast<float>(D) + eps);
    if (use_rope) {
      const int hw = g.n_h * g.n_w;
      const int rem = row % hw;
      s_pos[threadIdx.x][0] = g.f0 + row / hw;
      s_pos[threadIdx.x][1] = rem / g.n_w;
      s_pos[threadIdx.x][2] = rem % g.n_w;
    }
  }
  __syncthreads();
  if (use_rope) {
    for (int idx = threadIdx.x; idx < nrows * 64; idx += blockDim.x) {
      const int r = idx >> 6, pi = idx & 63;
      float2 cs;
      if (pi < 22)
        cs = __ldg(g.tab_f + s_pos[r][0] * 22 + pi);
      else if (pi < 43)
        cs = __ldg(g.tab_h + s_pos[r][1] * 21 + (pi - 22));
      else
        cs = __ldg(g.tab_w + s_pos[r][2] * 21 + (pi - 43));
      s_cs[r][pi] = cs;
    }
    __syncthreads();
  }
  for (int v = threadIdx.x; v < (D >> 3); v += blockDim.x) {
    uint4 in[RR2_ROWS];
#pragma unroll
    for (int r = 0; r < RR2_ROWS; ++r) {
      if (r < nrows) in[r] = reinterpret_cast<const uint4*>(src + static_cast<size_t>(row0 + r) * ld_src)[v];
    }
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w) + 2 * v);
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(w) + 2 * v + 1);
    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    const int col = v * 8;
    const int grp = col / group_cols;
    const int pair0 = (col & 127) >> 1;  // multiple of 4: the four pairs are 32 contiguous bytes of s_cs[r]
    __nv_bfloat16* dbase = dst + grp * group_stride + (col - grp * group_cols);
#pragma unroll
    for (int r = 0; r < RR2_ROWS; ++r) {
      if (r >= nrows) break;
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&in[r]);
      const float rs = s_rstd[r];
      float cs[8] = {1.f, 0.f, 1.f, 0.f, 1.f, 0.f, 1.f, 0.f};
      if (use_rope) {
        const float4 c01 = *reinterpret_cast<const float4*>(&s_cs[r][pair0]);
        const float4 c23 = *reinterpret_cast<const float4*>(&s_cs[r][pair0 + 2]);
        cs[0] = c01.x, cs[1] = c01.y, cs[2] = c01.z, cs[3] = c01.w;
        cs[4] = c23.x, cs[5] = c23.y, cs[6] = c23.z, cs[7] = c23.w;
      }
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(h2[j]);
        float a = f.x * rs * wv[2 * j];
        float b = f.y * rs * wv[2 * j + 1];
        if (use_rope) {
          const float ra = a * cs[2 * j] - b * cs[2 * j + 1];
          const float rb = a * cs[2 * j + 1] + b * cs[2 * j];
          a = ra;
          b = rb;
        }
        o[j] = pack_bf16x2(a, b);
      }
      *reinterpret_cast<uint4*>(dbase + static_cast<size_t>(row0 + r) * ld_dst) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// patchify: latent video fp32 [C, F, H, W] -> bf16 tokens [F*(H/2)*(W/2), C*4], column = c*4 + py*2 + px
// (the im2col of Conv3d(C, D, kernel = stride = (1,2,2)))
// ------------------------------------------------------------------------------------------------
__global__ void patchify_kernel(const float* __restrict__ lat, __nv_bfloat16* __restrict__ out, int C, int F, int H,
                                int W, int ld_out, int col_off) {
  const int hp = H >> 1, wp = W >> 1;
  const long long ntok = static_cast<long long>(F) * hp * wp;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= ntok * C) return;
  const int c = static_cast<int>(idx / ntok);
  const long long tok = idx - static_cast<long long>(c) * ntok;
  const int f = static_cast<int>(tok / (hp * wp));
  const int rem = static_cast<int>(tok - static_cast<long long>(f) * hp * wp);
  const int hh = rem / wp, ww = rem - hh * wp;
  const float* base = lat + ((static_cast<size_t>(c) * F + f) * H + 2 * hh) * W + 2 * ww;
  const float2 r0 = *reinterpret_cast<const float2*>(base);
  const float2 r1 = *reinterpret_cast<const float2*>(base + W);
  uint2 pk = make_uint2(pack_bf16x2(r0.x, r0.y), pack_bf16x2(r1.x, r1.y));
  *reinterpret_cast<uint2*>(out + tok * ld_out + col_off + c * 4) = pk;
}

// ------------------------------------------------------------------------------------------------
// unpatchify + classifier-free guidance + flow-match Euler step, in place on the latent video:
//   v = v_neg + cfg * (v_pos - v_neg);   lat += v * dsigma
// head outputs are [tokens, 4*C] fp32 with column = (py*2 + px)*C + c
// ------------------------------------------------------------------------------------------------
__global__ void unpatchify_cfg_step_kernel(float* __restrict__ lat, const float* __restrict__ vpos,
                                           const float* __restrict__ vneg, int C, int F, int H, int W, float cfg,
                                           float dsigma, float* __restrict__ v_out) {
  const long long n = static_cast<long long>(C) * F * H * W;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int x = static_cast<int>(idx % W);
  long long r = idx / W;
  const int y = static_cast<int>(r % H);
  r /= H;
  const int f = static_cast<int>(r % F);
  const int c = static_cast<int>(r / F);
  const int hp = H >> 1, wp = W >> 1;
  const long long tok = (static_cast<long long>(f) * hp + (y >> 1)) * wp + (x >> 1);
  const int col = (((y & 1) << 1) | (x & 1)) * C + c;
  const float vp = vpos[tok * (4 * C) + col];
  float v = vp;
  if (vneg) {
    const float vn = vneg[tok * (4 * C) + col];
    v = vn + cfg * (vp - vn);
  }
  if (v_out) v_out[idx] = v;
  if (lat) lat[idx] += v * dsigma;
}

// ------------------------------------------------------------------------------------------------
// y[n] = act_out( sum_k W[n,k] * act_in(x[k]) + b[n] ),  W bf16 [N,K]; one warp per output row.
// act codes: 0 none, 1 SiLU.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float silu(float x) { return x / (1.f + __expf(-x)); }

__global__ void __launch_bounds__(256)
gemv_kernel(const __nv_bfloat16* __restrict__ W, int ldw, const float* __restrict__ x,
            const float* __restrict__ b, float* __restrict__ y, int N, int K, int act_in, int act_out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= N) return;
  const uint4* wr = reinterpret_cast<const uint4*>(W + static_cast<size_t>(warp) * ldw);
  float acc = 0.f;
  for (int v = lane; v < (K >> 3); v += 32) {
    const uint4 pk = __ldg(wr + v);
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 wf = __bfloat1622float2(h2[j]);
      float x0 = x[v * 8 + 2 * j], x1 = x[v * 8 + 2 * j + 1];
      if (act_in == 1) {
        x0 = silu(x0);
        x1 = silu(x1);
      }
      acc += wf.x * x0 + wf.y * x1;
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    float r = acc + (b ? b[warp] : 0.f);
    if (act_out == 1) r = silu(r);
    y[warp] = r;
  }
}

// out[i] = a[i % period] + b[i]   (per-layer modulation table + time projection)
__global__ void add_bcast_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                                 long long n, int period) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i % period];
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(in + i);
    *reinterpret_cast<uint2*>(out + i) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  } else {
    for (long long k = i; k < n; ++k) out[k] = __float2bfloat16(in[k]);
  }
}

}  // namespace

int ln_modulate(const float* x, int ldx, const float* mul, const float* add, int mul_plus_one, __nv_bfloat16* out,
                int ldo, int rows, int D, float eps, cudaStream_t stream) {
  if (rows <= 0 || D <= 0 || (D & 3) || (ldx & 3) || (ldo & 3)) return IC_ERR_INVALID;
  const int nvec = D >> 2;
  const int vpt = (nvec + 127) / 128;
  if (vpt <= 3)
    ln_modulate_kernel<3><<<rows, 128, 0, stream>>>(x, ldx, mul, add, mul_plus_one, out, ldo, D, eps);
  else if (vpt <= 10)
    ln_modulate_kernel<10><<<rows, 128, 0, stream>>>(x, ldx, mul, add, mul_plus_one, out, ldo, D, eps);
  else
    return IC_ERR_UNSUPPORTED;
  ICB_CUDA_CHECK(cudaGetLastError());
  return IC_OK;
}

int rmsnorm_rope(const __nv_bfloat16* src, int ld_src, const float* ss, int ss_ld, int ss_off, int ss_cnt,
                 const float* w, __nv_bfloat16* dst, int ld_dst, int rows, int D, float eps, const RopeTables* rope,
                 int f0, cudaStream_t stream, int group_cols, long long group_stride) {
  if (rows <= 0 || (D & 127) || (ld_src & 7) || (ld_dst & 7)) return IC_ERR_INVALID;
  if (group_cols <= 0) group_cols = D;
  if ((group_cols & 127) || (group_stride & 7)) return IC_ERR_INVALID;
  RopeGeom g{};
  if (rope) {
    g.tab_f = reinterpret_cast<const float2*>(rope->tab_f);
    g.tab_h = reinterpret_cast<const float2*>(rope->tab_h);
    g.tab_w = reinterpret_cast<const float2*>(rope->tab_w);
    g.n_h = rope->n_h;
    g.n_w = rope->n_w;
    g.f0 = f0;
  }
  const char* v2env = getenv("ICB_RMSROPE_V2");  // read per call (a few hundred per step) so one process can A/B
  const int v2 = v2env ? atoi(v2env) : 1;
  if (v2)
    rmsnorm_rope_v2_kernel<<<(rows + RR2_ROWS - 1) / RR2_ROWS, 192, 0, stream>>>(
        src, ld_src, ss, ss_ld, ss_off, ss_cnt, w, dst, ld_dst, group_cols, group_stride, D, eps, rope ? 1 : 0, g, rows);
  else
    rmsnorm_rope_kernel<<<(rows + RR_ROWS - 1) / RR_ROWS, 192, 0, stream>>>(src, ld_src, ss, ss_ld, ss_off, ss_cnt, w, dst,
                                                                            ld_dst, group_cols, group_stride, D, eps,
                                                                            rope ? 1 : 0, g, rows);
  ICB_CUDA_CHECK(cudaGetLastError());
  return IC_OK;
}

int patchify(const float* lat, __nv_bfloat16* out, int C, int F, int H, int W, int ld_out, int col_off,
             cudaStream_t stream) {
  if ((H & 1) || (W & 1) || (ld_out & 3) || (col_off & 3)) return IC_ERR_INVALID;
  const long long n = static_cast<long long>(C) * F * (H / 2) * (W / 2);
  patchify_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(lat, out, C, F, H, W, ld_out, col_off);
  ICB_CUDA_CHECK(cudaGetLastError());
  return IC_OK;
}

int unpatchify_cfg_step(float* lat, const float* vpos, const float* vneg, int C, int F, int H, int W, float cfg,
                        float dsigma, float* v_out, cudaStream_t stream) {
  const long long n = static_cast<long long>(C) * F * H * W;
  unpatchify_cfg_step_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(lat, vpos, vneg, C, F, H, W,
                                                                                       cfg, dsigma, v_out);
  ICB_CUDA_CHECK(cudaGetLastError());
  return IC_OK;
}

int gemv_bf16(const __nv_bfloat16* W, int ldw, const float* x, const float* b, float* y, int N, int K, int act_in,
              int act_out, cudaStream_t stream) {
  if ((K & 7) || (ldw & 7)) return IC_ERR_INVALID;
  const int warps_per_block = 8;
  gemv_kernel<<<(N + warps_per_block - 1) / warps_per_block, 256, 0, stream>>>(W, ldw, x, b, y, N, K, act_in,
                                                                              act_out);
  ICB_CUDA_CHECK(cudaGetLastError());
  return IC_OK;
}

int add_bcast(const float* a, const float* b, float* out, long long n, int period, cudaStream_t stream) {
  add_bcast_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(a, b, out, n, period);
  ICB_CUDA_CHECK(cudaGetLastError());
  return IC_OK;
}

int cast_f32_bf16(const float* in, __nv_bfloat16* out, long long n, cudaStream_t stream) {
  const long long nthreads = (n + 3) / 4;
  cast_f32_bf16_kernel<<<static_cast<unsigned>((nthreads + 255) / 256), 256, 0, stream>>>(in, out, n);
  ICB_CUDA_CHECK(cudaGetLastError());
  return IC_OK;
}

}  // namespace icb
