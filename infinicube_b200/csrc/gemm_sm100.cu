// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma
// (accumulators in TMEM, double buffered) -> fused epilogue straight from TMEM.
//
// Replaces the cuBLASLt calls the reference stack issues for every nn.Linear of the Wan2.1 DiT
// (reference call site infinicube/videogen/inference.py:216-226,
// arithmetic in the un-vendored diffsynth WanModel).
//
// Roles (192 threads): warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane) + TMEM owner,
// warps 2..5 = epilogue (warp w owns TMEM lanes 32*(w%4) .. +31, one accumulator row per thread).
#include <stdlib.h>

#include <algorithm>

#include "gemm_sm100.cuh"
#include "host_util.h"

namespace icb {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int GEMM_THREADS = 192;
constexpr int GROUP_M = 16;  // m-blocks per L2 rasterisation group

template <int BN>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int TMEM_COLS = 2 * BN;  // two accumulator stages
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct GemmParams {
  int M, N, K;
  int num_m_blocks, num_n_blocks, num_k_blocks, num_tiles;
  int use_red;  // pair kernel: residual epilogue leaves the SM as TMA reduce-adds
  GemmEpilogue ep;
};

__device__ __forceinline__ void tile_coords(int tile, int num_m_blocks, int num_n_blocks, int& m_blk, int& n_blk) {
  // groups of GROUP_M m-blocks sweep all n-blocks, m fastest: the A panel of a group and the whole
  // weight matrix stay L2-resident while the group is in flight.
  const int group_tiles = GROUP_M * num_n_blocks;
  const int g = tile / group_tiles;
  const int first_m = g * GROUP_M;
  const int gm = min(GROUP_M, num_m_blocks - first_m);
  const int idx = tile - g * group_tiles;
  m_blk = first_m + idx % gm;
  n_blk = idx / gm;
}

// ---------------------------------------------------------------------------------------------------------
// Residual epilogue  x[row, :] += gate * (acc + bias)  on an fp32 stream that does not fit L2: a read-modify-write
// whose DRAM latency (not the math) paces the tile.  The x values of a chunk are fetched RES_PF chunks ahead of
// the accumulator columns they meet, and the first RES_PF chunks of a tile before the wait on its accumulator.
// ---------------------------------------------------------------------------------------------------------
constexpr int RES_PF = 3;
struct ResidRegs {
  float4 x[RES_PF][8];
};

__device__ __forceinline__ bool resid_fast_path(const GemmEpilogue& ep, const GemmParams& p) {
  return ep.resid && !ep.out_bf16 && !ep.out_f32 && !ep.rowss && !ep.bias_per_row && ep.act == 0 && (p.N & 31) == 0;
}

__device__ __forceinline__ void resid_load_chunk(const GemmEpilogue& ep, const GemmParams& p, int row, bool row_ok,
                                                 int col0, float4 (&x)[8]) {
  if (row_ok && col0 < p.N) {
    const float4* src = reinterpret_cast<const float4*>(ep.resid + static_cast<size_t>(row) * ep.ld_res + col0);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = src[i];
  }
}

template <int BN>
__device__ __forceinline__ void resid_prefetch(const GemmEpilogue& ep, const GemmParams& p, int row, bool row_ok, int n0,
                                               ResidRegs& r) {
#pragma unroll
  for (int c = 0; c < RES_PF && c < BN / 32; ++c) resid_load_chunk(ep, p, row, row_ok, n0 + c * 32, r.x[c]);
}

template <int BN>
__device__ __forceinline__ void epilogue_row_resid(const GemmEpilogue& ep, const GemmParams& p, int row, bool row_ok,
                                                   int n0, uint32_t taddr, ResidRegs& r) {
  constexpr int NCH = BN / 32;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int col0 = n0 + c * 32;
    if (col0 < p.N) {  // warp-uniform
      uint32_t raw[32];
      tmem_ld_x32(taddr + c * 32, raw);
      tmem_wait_ld();
      float4(&x)[8] = r.x[c % RES_PF];
      if (row_ok) {
        float* dst = ep.resid + static_cast<size_t>(row) * ep.ld_res + col0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 v = make_float4(__uint_as_float(raw[4 * i]), __uint_as_float(raw[4 * i + 1]),
                                 __uint_as_float(raw[4 * i + 2]), __uint_as_float(raw[4 * i + 3]));
          if (ep.bias) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col0) + i);
            v.x += b.x;
            v.y += b.y;
            v.z += b.z;
            v.w += b.w;
          }
          float4 g = make_float4(1.f, 1.f, 1.f, 1.f);
          if (ep.gate) g = __ldg(reinterpret_cast<const float4*>(ep.gate + col0) + i);
          float4 o = x[i];
          o.x += g.x * v.x;
          o.y += g.y * v.y;
          o.z += g.z * v.z;
          o.w += g.w * v.w;
          reinterpret_cast<float4*>(dst)[i] = o;
        }
      }
      if (c + RES_PF < NCH) resid_load_chunk(ep, p, row, row_ok, n0 + (c + RES_PF) * 32, x);
    }
  }
}

// Epilogue of one accumulator tile: this thread owns accumulator row `row` (TMEM lane), columns [n0, n0 + BN).
template <int BN>
__device__ __forceinline__ float epilogue_row(const GemmEpilogue& ep, const GemmParams& p, int row, bool row_ok, int n0,
                                              uint32_t taddr) {
  float ss = 0.f;
  const float rbias = (ep.bias && ep.bias_per_row && row_ok) ? ep.bias[row] : 0.f;
#pragma unroll 1
  for (int c = 0; c < BN / 32; ++c) {
    const int col0 = n0 + c * 32;
    if (col0 >= p.N) break;  // warp-uniform
    uint32_t raw[32];
    tmem_ld_x32(taddr + c * 32, raw);
    tmem_wait_ld();
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
    if (ep.bias) {
      if (ep.bias_per_row) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += rbias;
      } else {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          if (col0 + i < p.N) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + i));
            v[i] += b.x;
            v[i + 1] += b.y;
            v[i + 2] += b.z;
            v[i + 3] += b.w;
          }
        }
      }
    }
    if (ep.act == 1) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = gelu_tanh(v[i]);
    }
    if (row_ok) {
      if (ep.out_bf16) {
        __nv_bfloat16* dst = ep.out_bf16 + static_cast<size_t>(row) * ep.ld_out + col0;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          if (col0 + i < p.N) {
            uint4 pk;
            pk.x = pack_bf16x2(v[i], v[i + 1]);
            pk.y = pack_bf16x2(v[i + 2], v[i + 3]);
            pk.z = pack_bf16x2(v[i + 4], v[i + 5]);
            pk.w = pack_bf16x2(v[i + 6], v[i + 7]);
            *reinterpret_cast<uint4*>(dst + i) = pk;
            if (ep.rowss) {
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = __bfloat1622float2(h[j]);
                ss += f.x * f.x + f.y * f.y;
              }
            }
          }
        }
      }
      if (ep.out_f32) {
        float* dst = ep.out_f32 + static_cast<size_t>(row) * ep.ld_f32 + col0;
        const float* add = ep.addend ? ep.addend + static_cast<size_t>(row) * ep.ld_add + col0 : nullptr;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          if (col0 + i < p.N) {
            float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            if (add) {
              const float4 a = *reinterpret_cast<const float4*>(add + i);
              o.x += a.x;
              o.y += a.y;
              o.z += a.z;
              o.w += a.w;
            }
            *reinterpret_cast<float4*>(dst + i) = o;
          }
        }
      }
      if (ep.resid) {
        float* dst = ep.resid + static_cast<size_t>(row) * ep.ld_res + col0;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          if (col0 + i < p.N) {
            float4 g = make_float4(1.f, 1.f, 1.f, 1.f);
            if (ep.gate) g = __ldg(reinterpret_cast<const float4*>(ep.gate + col0 + i));
            float4 x = *reinterpret_cast<const float4*>(dst + i);
            x.x += g.x * v[i];
            x.y += g.y * v[i + 1];
            x.z += g.z * v[i + 2];
            x.w += g.w * v[i + 3];
            *reinterpret_cast<float4*>(dst + i) = x;
          }
        }
      }
    }
  }
  return ss;
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::STAGES * C::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full = bars;                       // [STAGES]  TMA -> MMA
  uint64_t* empty = bars + C::STAGES;          // [STAGES]  MMA -> TMA
  uint64_t* tmem_full = bars + 2 * C::STAGES;  // [2]       MMA -> epilogue
  uint64_t* tmem_empty = tmem_full + 2;        // [2]       epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    // warp-uniform control flow (operands stay in uniform registers); one elected lane issues
    {
      const bool leader = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int m_blk, n_blk;
        tile_coords(tile, p.num_m_blocks, p.num_n_blocks, m_blk, n_blk);
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (leader) {
            mbar_arrive_expect_tx(&full[stage], C::STAGE_BYTES);
            tma_load_2d(smem_a + stage * C::A_BYTES, &tmA, &full[stage], kb * BK, m_blk * BM, kEvictNormal);
            tma_load_2d(smem_b + stage * C::B_BYTES, &tmB, &full[stage], kb * BK, n_blk * BN, kEvictLast);
          }
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------
    {
      const bool leader = elect_one();
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128_kmajor(smem_u32(smem_a + stage * C::A_BYTES));
          const uint64_t bdesc = umma_desc_sw128_kmajor(smem_u32(smem_b + stage * C::B_BYTES));
          if (leader) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // advancing 16 bf16 (32 B) along K inside the swizzle row: +2 in the (addr >> 4) field
              umma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            }
            umma_commit(&empty[stage]);
          }
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (leader) umma_commit(&tmem_full[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------- epilogue -----------------------------------
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const GemmEpilogue& ep = p.ep;
    const bool fast_resid = resid_fast_path(ep, p);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      int m_blk, n_blk;
      tile_coords(tile, p.num_m_blocks, p.num_n_blocks, m_blk, n_blk);
      const int row = m_blk * BM + quad * 32 + lane;
      const bool row_ok = row < p.M;
      const int n0 = n_blk * BN;
      ResidRegs rr;
      if (fast_resid) resid_prefetch<BN>(ep, p, row, row_ok, n0, rr);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN;
      float ss = 0.f;
      if (fast_resid)
        epilogue_row_resid<BN>(ep, p, row, row_ok, n0, taddr, rr);
      else
        ss = epilogue_row<BN>(ep, p, row, row_ok, n0, taddr);
      // all TMEM reads of this accumulator stage are complete (tmem_wait_ld above): hand it back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (ep.rowss && row_ok) ep.rowss[static_cast<size_t>(row) * ep.rowss_ld + n_blk] = ss;
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): a cluster of two SMs owns one 256 x 256 tile.  Each CTA stages its own 128
// rows of A and its own 128 rows of B per k-block (32 KB / stage instead of 48 KB, six stages), the leader CTA
// issues one 256 x 256 x 16 MMA for the pair, and each CTA drains its own 128 accumulator rows.  Halves the
// B-operand traffic from L2 per SM and deepens the TMA ring.
// ---------------------------------------------------------------------------------------------------------
constexpr int PAIR_BN = 256;
constexpr int PAIR_STAGES = 6;
constexpr int PAIR_A_BYTES = 128 * BK * 2;
constexpr int PAIR_B_BYTES = 128 * BK * 2;
constexpr int PAIR_STAGE_BYTES = PAIR_A_BYTES + PAIR_B_BYTES;
constexpr int PAIR_RED_BYTES = 4 * 2 * 4096;  // per epilogue warp: two 32 x 32 fp32 boxes (reduce-add staging)
constexpr int PAIR_SMEM_BYTES = PAIR_STAGES * PAIR_STAGE_BYTES + PAIR_RED_BYTES + 1024 + 256;
constexpr int PAIR_GROUP_M = 8;  // 256-row blocks per L2 rasterisation group

__device__ __forceinline__ void pair_tile_coords(int tile, int num_m_blocks, int num_n_blocks, int& m_blk, int& n_blk) {
  const int group_tiles = PAIR_GROUP_M * num_n_blocks;
  const int g = tile / group_tiles;
  const int first_m = g * PAIR_GROUP_M;
  const int gm = min(PAIR_GROUP_M, num_m_blocks - first_m);
  const int idx = tile - g * group_tiles;
  m_blk = first_m + idx % gm;
  n_blk = idx / gm;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tn_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmR, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + PAIR_STAGES * PAIR_A_BYTES;
  uint8_t* smem_red = smem + PAIR_STAGES * PAIR_STAGE_BYTES;  // 1024-aligned: 128-byte swizzle atoms line up
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_red + PAIR_RED_BYTES);
  uint64_t* full = bars;                          // [STAGES] used in the leader CTA only
  uint64_t* empty = bars + PAIR_STAGES;           // [STAGES] one per CTA, signalled by the multicast commit
  uint64_t* tmem_full = bars + 2 * PAIR_STAGES;   // [2] one per CTA (multicast commit)
  uint64_t* tmem_empty = tmem_full + 2;           // [2] leader only: 8 arrivals (4 epilogue warps x 2 CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < PAIR_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(tmem_slot, 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // TMA producer (both CTAs): own halves of A and B; completion bytes go to the leader's full barrier
    const bool leader_lane = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = cluster_id; tile < p.num_tiles; tile += num_clusters) {
      int m_blk, n_blk;
      pair_tile_coords(tile, p.num_m_blocks, p.num_n_blocks, m_blk, n_blk);
      for (int kb = 0; kb < p.num_k_blocks; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (leader_lane) {
          if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * PAIR_STAGE_BYTES);
          tma_load_2d_2cta(smem_a + stage * PAIR_A_BYTES, &tmA, &full[stage], kb * BK, m_blk * 256 + rank * 128,
                           kEvictNormal);
          tma_load_2d_2cta(smem_b + stage * PAIR_B_BYTES, &tmB, &full[stage], kb * BK, n_blk * PAIR_BN + rank * 128,
                           kEvictLast);
        }
        if (++stage == PAIR_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // MMA issuer: leader CTA only
    if (rank == 0) {
      const bool leader_lane = elect_one();
      constexpr uint32_t idesc = umma_idesc_bf16(256, PAIR_BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < p.num_tiles; tile += num_clusters) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * PAIR_BN;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128_kmajor(smem_u32(smem_a + stage * PAIR_A_BYTES));
          const uint64_t bdesc = umma_desc_sw128_kmajor(smem_u32(smem_b + stage * PAIR_B_BYTES));
          if (leader_lane) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma_ss_2cta(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            umma_commit_2cta(&empty[stage]);
          }
          if (++stage == PAIR_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (leader_lane) umma_commit_2cta(&tmem_full[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // epilogue (both CTAs): each drains its own 128 accumulator rows
    const int quad = warp & 3;
    const GemmEpilogue& ep = p.ep;
    const bool fast_resid = resid_fast_path(ep, p);
    uint32_t red_chunk = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = cluster_id; tile < p.num_tiles; tile += num_clusters) {
      int m_blk, n_blk;
      pair_tile_coords(tile, p.num_m_blocks, p.num_n_blocks, m_blk, n_blk);
      const int row = m_blk * 256 + rank * 128 + quad * 32 + lane;
      const bool row_ok = row < p.M;
      const int n0 = n_blk * PAIR_BN;
      ResidRegs rr;
      if (fast_resid && !p.use_red) resid_prefetch<PAIR_BN>(ep, p, row, row_ok, n0, rr);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * PAIR_BN;
      float ss = 0.f;
      if (p.use_red) {
        // x += gate * (acc + bias) as TMA reduce-adds: this warp's 32 rows x 32 columns per box, two boxes in
        // flight (N is a multiple of 32 on this path).  The tensor map's row extent is M, so the box that straddles
        // row M is clipped by the TMA unit itself (rows >= M hold bias-only values computed from zero-filled A rows and
        // never leave shared memory); boxes entirely past M are not issued.
        const int row_base = m_blk * 256 + rank * 128 + quad * 32;
        const bool any_row = row_base < p.M;  // warp-uniform
#pragma unroll 1
        for (int c = 0; c < PAIR_BN / 32 && any_row; ++c) {
          const int col0 = n0 + c * 32;
          if (col0 >= p.N) break;  // warp-uniform
          uint32_t raw[32];
          tmem_ld_x32(taddr + c * 32, raw);
          tmem_wait_ld();
          uint8_t* buf = smem_red + quad * 8192 + (red_chunk & 1) * 4096;
          ++red_chunk;
          if (lane == 0) bulk_wait_group_read<1>();  // the box issued two chunks ago has left this buffer
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 v = make_float4(__uint_as_float(raw[4 * j]), __uint_as_float(raw[4 * j + 1]),
                                   __uint_as_float(raw[4 * j + 2]), __uint_as_float(raw[4 * j + 3]));
            if (ep.bias) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col0) + j);
              v.x += b.x;
              v.y += b.y;
              v.z += b.z;
              v.w += b.w;
            }
            if (ep.gate) {
              const float4 g = __ldg(reinterpret_cast<const float4*>(ep.gate + col0) + j);
              v.x *= g.x;
              v.y *= g.y;
              v.z *= g.z;
              v.w *= g.w;
            }
            *reinterpret_cast<float4*>(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) = v;  // SWIZZLE_128B
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_reduce_add_2d(&tmR, buf, col0, row_base);
            bulk_commit_group();
          }
        }
      } else if (fast_resid) {
        epilogue_row_resid<PAIR_BN>(ep, p, row, row_ok, n0, taddr, rr);
      } else {
        ss = epilogue_row<PAIR_BN>(ep, p, row, row_ok, n0, taddr);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], 0);
      if (ep.rowss && row_ok) ep.rowss[static_cast<size_t>(row) * ep.rowss_ld + n_blk] = ss;
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (p.use_red && lane == 0) bulk_wait_group<0>();  // staging memory and x are quiescent before the CTA retires
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2cta(tmem_base, 512);
}

template <int BN>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  using C = Cfg<BN>;
  static bool configured = false;
  if (!configured) {
    ICB_CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        C::SMEM_BYTES));
    configured = true;
  }
  const int grid = min(p.num_tiles, num_sms());
  gemm_bf16_tn_kernel<BN><<<grid, GEMM_THREADS, C::SMEM_BYTES, stream>>>(tmA, tmB, p);
  ICB_CUDA_CHECK(cudaGetLastError());
  return IC_OK;
}

}  // namespace

static int launch_pair(const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb, int M, int N, int K,
                       const GemmEpilogue& ep, cudaStream_t stream) {
  GemmParams p;
  p.M = M;
  p.N = N;
  p.K = K;
  p.num_m_blocks = (M + 255) / 256;
  p.num_n_blocks = (N + PAIR_BN - 1) / PAIR_BN;
  p.num_k_blocks = (K + BK - 1) / BK;
  p.num_tiles = p.num_m_blocks * p.num_n_blocks;
  p.ep = ep;
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)lda * 2};
    const uint32_t box[2] = {BK, 128};
    int r = make_tmap_bf16(&tmA, A, 2, dims, strides, box);
    if (r) return r;
  }
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    const uint64_t strides[1] = {(uint64_t)ldb * 2};
    const uint32_t box[2] = {BK, 128};
    int r = make_tmap_bf16(&tmB, B, 2, dims, strides, box);
    if (r) return r;
  }
  static bool configured = false;
  if (!configured) {
    ICB_CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_tn_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        PAIR_SMEM_BYTES));
    configured = true;
  }
  const int clusters = static_cast<int>(std::min<long long>(p.num_tiles, num_sms() / 2));
  CUtensorMap tmR = tmA;
  static int red_mode = -1;
  if (red_mode < 0) {
    const char* e = getenv("ICB_GEMM_RED");
    red_mode = e ? atoi(e) : 1;
  }
  p.use_red = 0;
  if (red_mode && ep.resid && !ep.out_bf16 && !ep.out_f32 && !ep.rowss && !ep.bias_per_row && ep.act == 0 &&
      (N % 32) == 0 && (ep.ld_res % 4) == 0 && (reinterpret_cast<uintptr_t>(ep.resid) & 15) == 0) {
    const uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)ep.ld_res * 4};
    const uint32_t box[2] = {32, 32};
    int r = make_tmap_f32(&tmR, ep.resid, 2, dims, strides, box);
    if (r) return r;
    p.use_red = 1;
  }
  gemm_bf16_tn_pair_kernel<<<2 * clusters, GEMM_THREADS, PAIR_SMEM_BYTES, stream>>>(tmA, tmB, tmR, p);
  ICB_CUDA_CHECK(cudaGetLastError());
  return IC_OK;
}

int gemm_block_n(int N) { return N >= 256 ? 256 : (N >= 128 ? 128 : 64); }

// Tile width actually launched: 256 unless 128-wide tiles fill the persistent grid's waves clearly better
// (small-M shards on multi-GPU runs: e.g. 37 x 6 tiles on 148 SMs = 1.5 waves, 37 x 12 = 3.0 waves).
static int pick_block_n(int M, int N, bool needs_fixed) {
  const int base = gemm_block_n(N);
  if (base != 256 || needs_fixed) return base;
  const int sms = num_sms();
  const long long mb = (M + BM - 1) / BM;
  auto waves_cost = [&](int bn) {
    const long long tiles = mb * ((N + bn - 1) / bn);
    const long long waves = (tiles + sms - 1) / sms;
    return static_cast<double>(waves) * bn;  // time ~ waves x tile width
  };
  return waves_cost(128) < 0.9 * waves_cost(256) ? 128 : 256;
}

int gemm_bf16_tn(const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb, int M, int N, int K,
                 const GemmEpilogue& ep, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return IC_ERR_INVALID;
  if ((K % 8) || (lda % 8) || (ldb % 8) || (N % 8)) return IC_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) return IC_ERR_INVALID;
  if (ep.rowss && !ep.out_bf16) return IC_ERR_INVALID;
  // CTA-pair kernel (256 x 256 tiles on two SMs) whenever it fills its waves well
  static int pair_mode = -1;
  if (pair_mode < 0) {
    const char* e = getenv("ICB_GEMM_PAIR");
    pair_mode = e ? atoi(e) : 1;
  }
  if (pair_mode && M >= 256 && N >= 256) {
    const long long tiles = static_cast<long long>((M + 255) / 256) * ((N + 255) / 256);
    const int clusters = num_sms() / 2;
    const long long waves = (tiles + clusters - 1) / clusters;
    if (static_cast<double>(tiles) / (waves * clusters) >= 0.85) return launch_pair(A, lda, B, ldb, M, N, K, ep, stream);
  }
  // rowss consumers index partial sums by gemm_block_n(N)-wide tiles, so those launches keep the nominal width
  const int bn = pick_block_n(M, N, ep.rowss != nullptr);

  GemmParams p;
  p.M = M;
  p.N = N;
  p.K = K;
  p.num_m_blocks = (M + BM - 1) / BM;
  p.num_n_blocks = (N + bn - 1) / bn;
  p.num_k_blocks = (K + BK - 1) / BK;
  p.num_tiles = p.num_m_blocks * p.num_n_blocks;
  p.use_red = 0;
  p.ep = ep;

  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)lda * 2};
    const uint32_t box[2] = {BK, BM};
    int r = make_tmap_bf16(&tmA, A, 2, dims, strides, box);
    if (r) return r;
  }
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    const uint64_t strides[1] = {(uint64_t)ldb * 2};
    const uint32_t box[2] = {BK, (uint32_t)bn};
    int r = make_tmap_bf16(&tmB, B, 2, dims, strides, box);
    if (r) return r;
  }
  switch (bn) {
    case 256:
      return launch<256>(tmA, tmB, p, stream);
    case 128:
      return launch<128>(tmA, tmB, p, stream);
    default:
      return launch<64>(tmA, tmB, p, stream);
  }
}

}  // namespace icb
