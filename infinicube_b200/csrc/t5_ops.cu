// Helper kernels of the umT5 prompt encoder: token-embedding gather, T5 RMS
// layer-norm, self-attention with the per-block relative-position bias and key padding mask, and the
// gated-GELU product.  The seven nn.Linear of a T5 block run on the tcgen05 GEMM (gemm_sm100.cu); these
// kernels are the glue between them.  The encoder runs twice per generate() call on 512 tokens, so this is
// latency-class work: 4.3 GFLOP of attention per block on the FP32 pipe, one launch per block.
#include "t5_ops.cuh"

#include <float.h>

#include "host_util.h"

namespace icb {

namespace {

// ------------------------------------------------------------------------------------------------
// x[row, :] = float(table[ids[row], :])        one 128-thread block per token, 16-byte loads
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
t5_embed_kernel(const int* __restrict__ ids, const __nv_bfloat16* __restrict__ table, int vocab, int D,
                float* __restrict__ x, int ldx) {
  const int row = blockIdx.x;
  int id = ids[row];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);  // the host wrapper rejects out-of-range ids before the upload
  const uint4* src = reinterpret_cast<const uint4*>(table + static_cast<size_t>(id) * D);
  float4* dst = reinterpret_cast<float4*>(x + static_cast<size_t>(row) * ldx);
  for (int i = threadIdx.x; i < (D >> 3); i += 128) {
    const uint4 v = __ldg(src + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    float f[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      f[2 * k] = __uint_as_float(w[k] << 16);
      f[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
    }
    dst[2 * i] = make_float4(f[0], f[1], f[2], f[3]);
    dst[2 * i + 1] = make_float4(f[4], f[5], f[6], f[7]);
  }
}

// ------------------------------------------------------------------------------------------------
// T5LayerNorm: out = bf16(weight * x * rsqrt(mean(x^2) + eps)); rows >= zero_from_row are written as 0
// (the prompter's "zero the padding rows").  One 256-thread block per row; the second pass re-reads the
// row from L1/L2 (16 KB at D = 4096).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
t5_rmsnorm_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                  int ldo, int D, float eps, int zero_from_row) {
  __shared__ float red[8];
  const int row = blockIdx.x;
  const int nvec = D >> 2;
  uint2* orow = reinterpret_cast<uint2*>(out + static_cast<size_t>(row) * ldo);
  if (row >= zero_from_row) {
    for (int i = threadIdx.x; i < nvec; i += 256) orow[i] = make_uint2(0u, 0u);
    return;
  }
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * ldx);
  float s = 0.f;
  for (int i = threadIdx.x; i < nvec; i += 256) {
    const float4 v = xr[i];
    s += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  const float r = rsqrtf(tot / static_cast<float>(D) + eps);
  for (int i = threadIdx.x; i < nvec; i += 256) {
    const float4 v = xr[i];
    const float4 g = __ldg(reinterpret_cast<const float4*>(w) + i);
    orow[i] = make_uint2(pack_bf16x2(g.x * (v.x * r), g.y * (v.y * r)), pack_bf16x2(g.z * (v.z * r), g.w * (v.w * r)));
  }
}

// ------------------------------------------------------------------------------------------------
// Self-attention of one T5 block, head_dim 64, no 1/sqrt(d) scaling:
//   P = softmax_j(q_i . k_j + bias[h][j - i + L - 1] + (mask[j] ? 0 : -FLT_MAX)),  out_i = sum_j P_ij v_j
// CTA = one head x 128 queries, one thread per query (q and the output row live in registers);
// K / V stream through shared memory in 32-key chunks, converted to fp32 ONCE while staging (every lane then
// reads the same key row: broadcast LDS.128, no conversions in the FMA loop); the chunk is consumed in
// 8-key groups by a rolled loop so the unrolled body (8 x (16 LDS + 64 FMA) x 2) stays inside the
// instruction cache (the fully unrolled 32-key body stalled 64 % of issue slots on instruction fetch);
// online softmax per group in fp32, the 64-wide rescale only when the running maximum moved.
// bias_by_offset is the block's relative-position table expanded to one value per offset j - i
// (2L - 1 floats per head), staged in shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int T5_DK = 64;
constexpr int T5_QT = 128;
constexpr int T5_KC = 32;
constexpr int T5_KG = 8;

__device__ __forceinline__ void bf16x8_to_f32(const uint4& v, float* f) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    f[2 * k] = __uint_as_float(w[k] << 16);
    f[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
  }
}

__global__ void __launch_bounds__(T5_QT)
t5_attention_kernel(const __nv_bfloat16* __restrict__ Q, const __nv_bfloat16* __restrict__ K,
                    const __nv_bfloat16* __restrict__ V, int ld, const float* __restrict__ bias_by_offset,
                    const unsigned char* __restrict__ key_mask, __nv_bfloat16* __restrict__ O, int ldo, int L) {
  extern __shared__ __align__(16) unsigned char t5_smem[];
  float4* ks = reinterpret_cast<float4*>(t5_smem);                     // [T5_KC][16] float4 = 32 keys x 64 fp32
  float4* vs = ks + T5_KC * (T5_DK / 4);                               // same for V
  float* sbias = reinterpret_cast<float*>(vs + T5_KC * (T5_DK / 4));   // [2L - 1]
  float* smask = sbias + (2 * L - 1);                                  // [L] 0 = attend, 1 = padding key

  const int h = blockIdx.y;
  const int tid = threadIdx.x;
  const int i = blockIdx.x * T5_QT + tid;        // this thread's query
  const int ic = i < L ? i : L - 1;              // clamped copy for address arithmetic (tail threads still help load)
  const size_t col = static_cast<size_t>(h) * T5_DK;

  for (int t = tid; t < 2 * L - 1; t += T5_QT) sbias[t] = bias_by_offset[static_cast<size_t>(h) * (2 * L - 1) + t];
  for (int t = tid; t < L; t += T5_QT) smask[t] = (key_mask == nullptr || key_mask[t]) ? 0.f : 1.f;

  float q[T5_DK];
  {
    const uint4* qr = reinterpret_cast<const uint4*>(Q + static_cast<size_t>(ic) * ld + col);
#pragma unroll
    for (int c = 0; c < T5_DK / 8; ++c) bf16x8_to_f32(qr[c], q + 8 * c);
  }
  float o[T5_DK];
#pragma unroll
  for (int c = 0; c < T5_DK; ++c) o[c] = 0.f;
  float m = -INFINITY, l = 0.f;
  const float* brow = sbias + (L - 1 - ic);      // brow[j] = bias of key j for this query

  for (int k0 = 0; k0 < L; k0 += T5_KC) {
    __syncthreads();  // previous chunk fully consumed (also orders the sbias / smask fill before first use)
    // 32 keys x 8 segments of 8 bf16 per matrix = 256 segments: two per thread and matrix
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int e = tid + r * T5_QT;
      const int kr = e >> 3, kc = e & 7;
      const int j = k0 + kr;
      uint4 kv = make_uint4(0u, 0u, 0u, 0u), vv = kv;
      if (j < L) {
        kv = *(reinterpret_cast<const uint4*>(K + static_cast<size_t>(j) * ld + col) + kc);
        vv = *(reinterpret_cast<const uint4*>(V + static_cast<size_t>(j) * ld + col) + kc);
      }
      float f[8];
      bf16x8_to_f32(kv, f);
      ks[kr * 16 + kc * 2] = make_float4(f[0], f[1], f[2], f[3]);
      ks[kr * 16 + kc * 2 + 1] = make_float4(f[4], f[5], f[6], f[7]);
      bf16x8_to_f32(vv, f);
      vs[kr * 16 + kc * 2] = make_float4(f[0], f[1], f[2], f[3]);
      vs[kr * 16 + kc * 2 + 1] = make_float4(f[4], f[5], f[6], f[7]);
    }
    __syncthreads();

#pragma unroll 1
    for (int g0 = 0; g0 < T5_KC; g0 += T5_KG) {
      if (k0 + g0 >= L) break;  // uniform over the CTA
      float s[T5_KG];
      float mc = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < T5_KG; ++jj) {
        const int j = k0 + g0 + jj;
        const float4* kr4 = ks + (g0 + jj) * 16;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;  // four chains: the dot product is latency-, not issue-bound
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const float4 kk = kr4[c];
          a0 = fmaf(q[4 * c], kk.x, a0);
          a1 = fmaf(q[4 * c + 1], kk.y, a1);
          a2 = fmaf(q[4 * c + 2], kk.z, a2);
          a3 = fmaf(q[4 * c + 3], kk.w, a3);
        }
        float sv = -INFINITY;  // keys past the end of a ragged sequence do not exist
        if (j < L) {
          sv = ((a0 + a1) + (a2 + a3)) + brow[j];
          if (smask[j] != 0.f) sv = -FLT_MAX;  // masked_fill_(mask == 0, finfo.min)
        }
        s[jj] = sv;
        mc = fmaxf(mc, sv);
      }
      const float m_new = fmaxf(m, mc);  // finite: every group that runs holds at least one existing key
      if (m_new != m) {
        const float corr = expf(m - m_new);
        l *= corr;
#pragma unroll
        for (int c = 0; c < T5_DK; ++c) o[c] *= corr;
        m = m_new;
      }
#pragma unroll
      for (int jj = 0; jj < T5_KG; ++jj) {
        const float p = expf(s[jj] - m);
        l += p;
        const float4* vr4 = vs + (g0 + jj) * 16;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const float4 vv = vr4[c];
          o[4 * c] = fmaf(p, vv.x, o[4 * c]);
          o[4 * c + 1] = fmaf(p, vv.y, o[4 * c + 1]);
          o[4 * c + 2] = fmaf(p, vv.z, o[4 * c + 2]);
          o[4 * c + 3] = fmaf(p, vv.w, o[4 * c + 3]);
        }
      }
    }
  }

  if (i < L) {
    const float inv = 1.f / l;
    uint4* orow = reinterpret_cast<uint4*>(O + static_cast<size_t>(i) * ldo + col);
#pragma unroll
    for (int c = 0; c < T5_DK / 8; ++c) {
      orow[c] = make_uint4(pack_bf16x2(o[8 * c] * inv, o[8 * c + 1] * inv), pack_bf16x2(o[8 * c + 2] * inv, o[8 * c + 3] * inv),
                           pack_bf16x2(o[8 * c + 4] * inv, o[8 * c + 5] * inv),
                           pack_bf16x2(o[8 * c + 6] * inv, o[8 * c + 7] * inv));
    }
  }
}

// out = bf16(float(a) * float(b)), 8 elements per thread
__global__ void __launch_bounds__(256)
mul_bf16_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ out, long long nvec) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= nvec) return;
  float fa[8], fb[8];
  bf16x8_to_f32(a[i], fa);
  bf16x8_to_f32(b[i], fb);
  out[i] = make_uint4(pack_bf16x2(fa[0] * fb[0], fa[1] * fb[1]), pack_bf16x2(fa[2] * fb[2], fa[3] * fb[3]),
                      pack_bf16x2(fa[4] * fb[4], fa[5] * fb[5]), pack_bf16x2(fa[6] * fb[6], fa[7] * fb[7]));
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

int t5_embed(const int* ids, int L, const __nv_bfloat16* table, int vocab, int D, float* x, int ldx,
             cudaStream_t stream) {
  if (L <= 0 || vocab <= 0 || D <= 0 || (D & 7) || (ldx & 3) || ldx < D) return IC_ERR_INVALID;
  if (!aligned16(table) || !aligned16(x)) return IC_ERR_INVALID;
  t5_embed_kernel<<<L, 128, 0, stream>>>(ids, table, vocab, D, x, ldx);
  ICB_CUDA_CHECK(cudaGetLastError());
  return IC_OK;
}

int t5_rmsnorm(const float* x, int ldx, const float* w, __nv_bfloat16* out, int ldo, int rows, int D, float eps,
               int zero_from_row, cudaStream_t stream) {
  if (rows <= 0 || D <= 0 || (D & 3) || (ldx & 3) || (ldo & 3) || ldx < D || ldo < D) return IC_ERR_INVALID;
  if (!aligned16(x) || !aligned16(w) || (reinterpret_cast<uintptr_t>(out) & 7)) return IC_ERR_INVALID;
  if (zero_from_row < 0) zero_from_row = rows;
  t5_rmsnorm_kernel<<<rows, 256, 0, stream>>>(x, ldx, w, out, ldo, D, eps, zero_from_row);
  ICB_CUDA_CHECK(cudaGetLastError());
  return IC_OK;
}

int t5_attention(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, int ld,
                 const float* bias_by_offset, const unsigned char* key_mask, __nv_bfloat16* out, int ldo, int L,
                 int n_heads, cudaStream_t stream) {
  if (L <= 0 || n_heads <= 0 || (ld & 7) || (ldo & 7) || ldo < n_heads * T5_DK) return IC_ERR_INVALID;
  if (!aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(out)) return IC_ERR_INVALID;
  const size_t smem = 2 * T5_KC * T5_DK * sizeof(float) + (static_cast<size_t>(2 * L - 1) + L) * sizeof(float);
  if (smem > 200 * 1024) return IC_ERR_UNSUPPORTED;
  if (smem > 48 * 1024) {
    ICB_CUDA_CHECK(cudaFuncSetAttribute(t5_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
  }
  const dim3 grid((L + T5_QT - 1) / T5_QT, n_heads);
  t5_attention_kernel<<<grid, T5_QT, smem, stream>>>(q, k, v, ld, bias_by_offset, key_mask, out, ldo, L);
  ICB_CUDA_CHECK(cudaGetLastError());
  return IC_OK;
}

int mul_bf16(const __nv_bfloat16* a, const __nv_bfloat16* b, __nv_bfloat16* out, long long n, cudaStream_t stream) {
  if (n <= 0 || (n & 7)) return IC_ERR_INVALID;
  if (!aligned16(a) || !aligned16(b) || !aligned16(out)) return IC_ERR_INVALID;
  const long long nvec = n >> 3;
  mul_bf16_kernel<<<static_cast<unsigned>((nvec + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const uint4*>(a), reinterpret_cast<const uint4*>(b), reinterpret_cast<uint4*>(out), nvec);
  ICB_CUDA_CHECK(cudaGetLastError());
  return IC_OK;
}

}  // namespace icb
