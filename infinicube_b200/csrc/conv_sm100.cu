// Implicit-GEMM causal convolution for the Wan 3-D VAE on sm_100a: no im2col buffer — each (tap, channel-chunk)
// K-step is one 4-D TMA box load {channels, 16 px in W, 8 px in H, 1 frame} from the channels-last activation,
// shifted by the tap offset; TMA's out-of-bounds zero fill IS the spatial zero padding and the causal temporal
// padding.  MMA, TMEM double buffering and the epilogue follow gemm_sm100.cu.
//
// Replaces the cuDNN conv3d / conv2d calls of diffsynth's WanVideoVAE (CausalConv3d, Resample) used by the
// reference at infinicube/videogen/inference.py:216-226.
#include <stdlib.h>

#include <type_traits>

#include "conv_sm100.cuh"
#include "host_util.h"

namespace icb {

namespace {

constexpr int CONV_THREADS = 192;
constexpr int TILE_W = 16, TILE_H = 8;  // 128 output pixels per M tile
constexpr int MAX_TAPS = 27;
constexpr int kMaxCout = 1024;  // bias vector staged in shared memory

struct ConvParams {
  int T, H, W, Cout;
  int ntaps, k_chunks;  // channel chunks per tap
  int Cin;
  int tiles_w, tiles_h, num_m_tiles, num_n_tiles, num_tiles;
  int tap[MAX_TAPS][3];
  const float* bias;
  const __nv_bfloat16* resid;
  int ld_resid;
  __nv_bfloat16* out;
  int ld_out;
};

// KSUB: swizzle-wide channel chunks staged per pipeline stage (3 for C = 96: one whole tap per barrier round trip)
template <int BN, int BKC, int KSUB>
struct CCfg {
  static constexpr int A_SUB = 128 * BKC * 2;
  static constexpr int B_SUB = BN * BKC * 2;
  static constexpr int A_BYTES = A_SUB * KSUB;
  static constexpr int B_BYTES = B_SUB * KSUB;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 8 ? 8 : (200 * 1024) / STAGE_BYTES;
  static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

template <int BN, int BKC, int KSUB>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmW, const ConvParams p) {
  using C = CCfg<BN, BKC, KSUB>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::STAGES * C::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::STAGES;
  uint64_t* tmem_full = bars + 2 * C::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  __shared__ float sbias[kMaxCout];  // the epilogue reads the bias from here (an L1 round trip per use paced its drain)
  for (int i = threadIdx.x; i < kMaxCout; i += CONV_THREADS) sbias[i] = (p.bias && i < p.Cout) ? p.bias[i] : 0.f;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmIn);
    tma_prefetch_desc(&tmW);
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
  const int k_steps = p.ntaps * p.k_chunks;

  if (warp == 0) {
    {
      const bool leader = elect_one();  // warp-uniform control flow, one lane issues
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int n_blk = tile % p.num_n_tiles;  // n fastest: the activation tile is reused from L2
        const int m_tile = tile / p.num_n_tiles;
        const int wb = m_tile % p.tiles_w;
        const int hb = (m_tile / p.tiles_w) % p.tiles_h;
        const int t = m_tile / (p.tiles_w * p.tiles_h);
        for (int ks = 0; ks < k_steps; ++ks) {
          const int tap = ks / p.k_chunks;
          const int cc = ks - tap * p.k_chunks;
          mbar_wait(&empty[stage], phase ^ 1);
          if (leader) {
            mbar_arrive_expect_tx(&full[stage], C::STAGE_BYTES);
#pragma unroll
            for (int sub = 0; sub < KSUB; ++sub) {
              const int c0 = (cc * KSUB + sub) * BKC;
              tma_load_4d(smem_a + stage * C::A_BYTES + sub * C::A_SUB, &tmIn, &full[stage], c0,
                          wb * TILE_W + p.tap[tap][2], hb * TILE_H + p.tap[tap][1], t + p.tap[tap][0], kEvictNormal);
              tma_load_2d(smem_b + stage * C::B_BYTES + sub * C::B_SUB, &tmW, &full[stage], tap * p.Cin + c0, n_blk * BN,
                          kEvictLast);
            }
          }
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      const bool leader = elect_one();
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int ks = 0; ks < k_steps; ++ks) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * C::A_BYTES);
          const uint32_t b_addr = smem_u32(smem_b + stage * C::B_BYTES);
          const uint64_t adesc = BKC == 64 ? umma_desc_sw128_kmajor(a_addr) : umma_desc_sw64_kmajor(a_addr);
          const uint64_t bdesc = BKC == 64 ? umma_desc_sw128_kmajor(b_addr) : umma_desc_sw64_kmajor(b_addr);
          if (leader) {
#pragma unroll
            for (int sub = 0; sub < KSUB; ++sub) {
#pragma unroll
              for (int k = 0; k < BKC / 16; ++k)
                umma_ss(d_tmem, adesc + ((sub * C::A_SUB) >> 4) + 2 * k, bdesc + ((sub * C::B_SUB) >> 4) + 2 * k, idesc,
                        (ks | sub | k) != 0);
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
    const int quad = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int n_blk = tile % p.num_n_tiles;
      const int m_tile = tile / p.num_n_tiles;
      const int wb = m_tile % p.tiles_w;
      const int hb = (m_tile / p.tiles_w) % p.tiles_h;
      const int t = m_tile / (p.tiles_w * p.tiles_h);
      const int r = quad * 32 + lane;
      const int h = hb * TILE_H + (r >> 4);
      const int w = wb * TILE_W + (r & 15);
      const bool ok = h < p.H && w < p.W;
      const size_t pix = (static_cast<size_t>(t) * p.H + h) * p.W + w;
      const int n0 = n_blk * BN;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int col0 = n0 + c * 32;
        if (col0 >= p.Cout) break;
        uint32_t raw[32];
        tmem_ld_x32(taddr + c * 32, raw);
        // the residual row of this chunk is requested before the TMEM round trip is waited for
        uint4 rv[4] = {};
        if (ok && p.resid) {
          const __nv_bfloat16* res = p.resid + pix * p.ld_resid + col0;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (col0 + 8 * i < p.Cout) rv[i] = *reinterpret_cast<const uint4*>(res + 8 * i);
        }
        tmem_wait_ld();
        if (ok) {
          __nv_bfloat16* dst = p.out + pix * p.ld_out + col0;
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            if (col0 + i < p.Cout) {  // Cout is a multiple of 8 (validated on the host)
              float v[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(raw[i + j]) + sbias[col0 + i + j];
              if (p.resid) {
                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&rv[i >> 3]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 f = __bfloat1622float2(h2[j]);
                  v[2 * j] += f.x;
                  v[2 * j + 1] += f.y;
                }
              }
              uint4 pk;
              pk.x = pack_bf16x2(v[0], v[1]);
              pk.y = pack_bf16x2(v[2], v[3]);
              pk.z = pack_bf16x2(v[4], v[5]);
              pk.w = pack_bf16x2(v[6], v[7]);
              *reinterpret_cast<uint4*>(dst + i) = pk;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
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

// ------------------------------------------------------------------------------------------------------------------
// Causal 3x3x3 convolution, Cin = 96, with the temporal taps shared in shared memory ("multi-frame" tiles).
//
// The generic kernel above fetches one 128-pixel x 96-channel box per tap: 27 boxes (648 KB) through L2 per output tile,
// and ncu shows it waiting on exactly that (profiles/r2_conv96.ncu-rep: tensor pipe 35 % active, TMA delivering
// 47 B/clk/SM where the full tensor rate needs 146).  Here one CTA owns a 16 x 8 pixel tile of kMfFrames = 4 CONSECUTIVE
// output frames (4 x BN fp32 accumulator columns in TMEM).  For each of the 9 spatial taps it loads the 3 temporal
// weight slices once (W stage) and the 6 input frames t0-2 .. t0+3 once each (A stages); input frame f feeds the
// accumulators of output frames f, f+1, f+2 through the weight slices kt = 2, 1, 0 - so an A box is used by up to
// three MMA groups instead of one: 144 KB of A + 54 KB of W per spatial tap for 12 tap-MMA groups, 16.5 KB per group
// instead of 42 KB.  Out-of-range frames (before the clip, causal padding; past its end in the last group) are TMA
// zero fill exactly as the spatial padding is.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kMfFrames = 4;
constexpr int kMfAStages = 4;
constexpr int kMfWStages = 2;

// Shared-memory layout: the 96 channels are split 64 + 32.  The 64-channel part uses 128-byte rows / SWIZZLE_128B, the
// 32-channel part 64-byte rows / SWIZZLE_64B: ncu on an all-64-byte version (profiles/r2_conv96_mf.ncu-rep) showed the
// producer waiting for free stages and UTCHMMA throttled by the MIO queue at 42 % tensor-pipe activity - MMAs fed from
// 64-byte-swizzled operands run at less than half the rate of 128-byte-swizzled ones - so two thirds of the K-steps
// are moved to the fast layout.
template <int BN>
struct MfCfg {
  static constexpr int A64 = 128 * 64 * 2;              // 128 pixels x 64 channels (128-byte rows)
  static constexpr int A32 = 128 * 32 * 2;              // 128 pixels x 32 channels (64-byte rows)
  static constexpr int A_BYTES = A64 + A32;
  static constexpr int B64 = BN * 64 * 2;               // one temporal slice, channels 0-63
  static constexpr int B32 = BN * 32 * 2;               // one temporal slice, channels 64-95
  static constexpr int W_BYTES = 3 * (B64 + B32);       // [3 slices x B64 | 3 slices x B32], slices in reverse kt order
  static constexpr int TMEM_COLS = kMfFrames * BN <= 256 ? 256 : 512;
  static constexpr int SMEM_BYTES = kMfAStages * A_BYTES + kMfWStages * W_BYTES + 1024 + 256;
};

template <int BN>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv3x3x3_mf_kernel(const __grid_constant__ CUtensorMap tmIn64, const __grid_constant__ CUtensorMap tmIn32,
                    const __grid_constant__ CUtensorMap tmW64, const __grid_constant__ CUtensorMap tmW32, const ConvParams p) {
  using C = MfCfg<BN>;
  static_assert(kMfFrames * BN <= 512, "accumulators must fit TMEM");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_w = smem + kMfAStages * C::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_w + kMfWStages * C::W_BYTES);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + kMfAStages;
  uint64_t* w_full = a_empty + kMfAStages;
  uint64_t* w_empty = w_full + kMfWStages;
  uint64_t* acc_full = w_empty + kMfWStages;    // [kMfFrames] accumulator o holds its finished tile
  uint64_t* acc_empty = acc_full + kMfFrames;   // [kMfFrames] accumulator o has been drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kMfFrames);

  __shared__ float sbias[BN];  // the epilogue reads the bias from here (an L1 round trip per use paced its drain)
  for (int i = threadIdx.x; i < BN; i += CONV_THREADS) sbias[i] = (p.bias && i < p.Cout) ? p.bias[i] : 0.f;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmIn64);
    tma_prefetch_desc(&tmIn32);
    tma_prefetch_desc(&tmW64);
    tma_prefetch_desc(&tmW32);
    for (int s = 0; s < kMfAStages; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < kMfWStages; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    for (int o = 0; o < kMfFrames; ++o) {
      mbar_init(&acc_full[o], 1);
      mbar_init(&acc_empty[o], 4);
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
  const int groups = (p.T + kMfFrames - 1) / kMfFrames;
  const int tiles_hw = p.tiles_w * p.tiles_h;
  const int num_tiles = tiles_hw * groups;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    const bool leader = elect_one();
    int as = 0, ws = 0;
    uint32_t aph = 0, wph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int sp = tile % tiles_hw;  // spatial tile fastest: neighbours of one frame group share their halos in L2
      const int wb = sp % p.tiles_w, hb = sp / p.tiles_w;
      const int t0 = (tile / tiles_hw) * kMfFrames;
      for (int s = 0; s < 9; ++s) {
        const int dh = s / 3 - 1, dw = s % 3 - 1;
        mbar_wait(&w_empty[ws], wph ^ 1);
        if (leader) {
          mbar_arrive_expect_tx(&w_full[ws], C::W_BYTES);
#pragma unroll
          for (int kt = 0; kt < 3; ++kt) {
            // per channel part: [temporal slice in REVERSE order][BN rows] - the slices that share one A operand
            // (kt = 2, 1, 0 -> ascending output frames) are adjacent row blocks, so one MMA can span them
            uint8_t* w0 = smem_w + ws * C::W_BYTES;
            tma_load_2d(w0 + (2 - kt) * C::B64, &tmW64, &w_full[ws], (kt * 9 + s) * 96, 0, kEvictLast);
            tma_load_2d(w0 + 3 * C::B64 + (2 - kt) * C::B32, &tmW32, &w_full[ws], (kt * 9 + s) * 96 + 64, 0, kEvictLast);
          }
        }
        if (++ws == kMfWStages) {
          ws = 0;
          wph ^= 1;
        }
        for (int fi = 0; fi < kMfFrames + 2; ++fi) {
          mbar_wait(&a_empty[as], aph ^ 1);
          if (leader) {
            mbar_arrive_expect_tx(&a_full[as], C::A_BYTES);
            tma_load_4d(smem_a + as * C::A_BYTES, &tmIn64, &a_full[as], 0, wb * TILE_W + dw, hb * TILE_H + dh, t0 - 2 + fi,
                        kEvictNormal);
            tma_load_4d(smem_a + as * C::A_BYTES + C::A64, &tmIn32, &a_full[as], 64, wb * TILE_W + dw, hb * TILE_H + dh,
                        t0 - 2 + fi, kEvictNormal);
          }
          if (++as == kMfAStages) {
            as = 0;
            aph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------
    // Input frame t0-2+fi reaches output frames o = fi-kt (kt = 0..2, dt = kt-2) through temporal slice kt.  The valid
    // o form a run [o_lo, o_hi]; their slices are adjacent row blocks of the W stage (position 2 - kt) and their
    // accumulators adjacent TMEM columns, so a run is issued as MMAs of up to kMaxRun x BN columns sharing one read of
    // A - except at the first spatial tap, where the run mixes an accumulator that starts (o = fi, accumulate = 0)
    // with accumulators that continue.  Everything about a frame slot fi is a compile-time constant (the first
    // version computed runs, positions and descriptors in a run-time loop inside the elected lane: ~160 dependent
    // scalar instructions per A stage paced the tensor pipe at 42 %).
    const bool leader = elect_one();
    constexpr int kMaxRun = 256 / BN >= 3 ? 3 : 256 / BN;
    int as = 0, ws = 0;
    uint32_t aph = 0, wph = 0, tph = 0;
    auto issue = [&](auto fic, const bool first_tap, const bool last_tap, const uint32_t a_addr, const uint32_t w_addr) {
      constexpr int FI = decltype(fic)::value;
      constexpr int o_lo = FI - 2 > 0 ? FI - 2 : 0;
      constexpr int o_hi = FI < kMfFrames - 1 ? FI : kMfFrames - 1;
      const uint64_t adesc64 = umma_desc_sw128_kmajor(a_addr);
      const uint64_t adesc32 = umma_desc_sw64_kmajor(a_addr + C::A64);
      if (first_tap) {
#pragma unroll
        for (int o = o_lo; o <= o_hi; ++o) {
          const int pos = 2 - (FI - o);
          const uint64_t bdesc64 = umma_desc_sw128_kmajor(w_addr + pos * C::B64);
          const uint64_t bdesc32 = umma_desc_sw64_kmajor(w_addr + 3 * C::B64 + pos * C::B32);
          constexpr uint32_t id = umma_idesc_bf16(128, BN);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tmem_base + o * BN, adesc64 + 2 * k, bdesc64 + 2 * k, id, !(o == FI && k == 0));
#pragma unroll
          for (int k = 0; k < 2; ++k) umma_ss(tmem_base + o * BN, adesc32 + 2 * k, bdesc32 + 2 * k, id, 1);
        }
      } else {
#pragma unroll
        for (int o = o_lo; o <= o_hi; o += kMaxRun) {
          const int run = o_hi - o + 1 < kMaxRun ? o_hi - o + 1 : kMaxRun;
          const int pos = 2 - (FI - o);
          const uint64_t bdesc64 = umma_desc_sw128_kmajor(w_addr + pos * C::B64);
          const uint64_t bdesc32 = umma_desc_sw64_kmajor(w_addr + 3 * C::B64 + pos * C::B32);
          const uint32_t id = umma_idesc_bf16(128, run * BN);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tmem_base + o * BN, adesc64 + 2 * k, bdesc64 + 2 * k, id, 1);
#pragma unroll
          for (int k = 0; k < 2; ++k) umma_ss(tmem_base + o * BN, adesc32 + 2 * k, bdesc32 + 2 * k, id, 1);
        }
      }
      if (FI >= 2 && last_tap) umma_commit(&acc_full[FI >= 2 ? FI - 2 : 0]);  // the last MMA into accumulator FI - 2
    };
    auto slot = [&](auto fic, const bool first_tap, const bool last_tap, const uint32_t w_addr) {
      constexpr int FI = decltype(fic)::value;
      mbar_wait(&a_full[as], aph);
      // Accumulators are handed over one by one: output frame o gets its first MMA of a tile at the first spatial tap
      // from input slot o (it must have been drained) and its last one at the last tap from slot o + 2 (committed to the
      // epilogue right there), so the drain of a tile overlaps the tail of its main loop and the head of the next.
      if (FI < kMfFrames && first_tap) mbar_wait(&acc_empty[FI < kMfFrames ? FI : 0], tph ^ 1);
      tc_fence_after();
      if (leader) {
        issue(fic, first_tap, last_tap, smem_u32(smem_a + as * C::A_BYTES), w_addr);
        umma_commit(&a_empty[as]);
      }
      if (++as == kMfAStages) {
        as = 0;
        aph ^= 1;
      }
    };
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int s = 0; s < 9; ++s) {
        mbar_wait(&w_full[ws], wph);
        const uint32_t w_addr = smem_u32(smem_w + ws * C::W_BYTES);
        const bool first_tap = s == 0, last_tap = s == 8;
        slot(std::integral_constant<int, 0>{}, first_tap, last_tap, w_addr);
        slot(std::integral_constant<int, 1>{}, first_tap, last_tap, w_addr);
        slot(std::integral_constant<int, 2>{}, first_tap, last_tap, w_addr);
        slot(std::integral_constant<int, 3>{}, first_tap, last_tap, w_addr);
        slot(std::integral_constant<int, 4>{}, first_tap, last_tap, w_addr);
        slot(std::integral_constant<int, 5>{}, first_tap, last_tap, w_addr);
        static_assert(kMfFrames + 2 == 6, "six input slots per spatial tap");
        if (leader) umma_commit(&w_empty[ws]);
        if (++ws == kMfWStages) {
          ws = 0;
          wph ^= 1;
        }
      }
      tph ^= 1;
    }
  } else {
    // ------------------------------- epilogue -----------------------------------
    const int quad = warp & 3;
    uint32_t tph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int sp = tile % tiles_hw;
      const int wb = sp % p.tiles_w, hb = sp / p.tiles_w;
      const int t0 = (tile / tiles_hw) * kMfFrames;
      const int r = quad * 32 + lane;
      const int h = hb * TILE_H + (r >> 4);
      const int w = wb * TILE_W + (r & 15);
      const bool ok = h < p.H && w < p.W;
#pragma unroll 1
      for (int o = 0; o < kMfFrames; ++o) {
        mbar_wait(&acc_full[o], tph);
        tc_fence_after();
        if (t0 + o >= p.T) {  // a frame past the end of the clip (last group): nothing to store, hand it straight back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[o]);
          continue;
        }
        const size_t pix = (static_cast<size_t>(t0 + o) * p.H + h) * p.W + w;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + o * BN;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          const int col0 = c * 32;
          if (col0 >= p.Cout) break;
          uint32_t raw[32];
          tmem_ld_x32(taddr + c * 32, raw);
          // the residual row of this chunk is requested before the TMEM round trip is waited for
          uint4 rv[4] = {};
          if (ok && p.resid) {
            const __nv_bfloat16* res = p.resid + pix * p.ld_resid + col0;
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (col0 + 8 * i < p.Cout) rv[i] = *reinterpret_cast<const uint4*>(res + 8 * i);
          }
          tmem_wait_ld();
          if (ok) {
            __nv_bfloat16* dst = p.out + pix * p.ld_out + col0;
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              if (col0 + i < p.Cout) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(raw[i + j]) + sbias[col0 + i + j];
                if (p.resid) {
                  const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&rv[i >> 3]);
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float2 f = __bfloat1622float2(h2[j]);
                    v[2 * j] += f.x;
                    v[2 * j + 1] += f.y;
                  }
                }
                uint4 pk;
                pk.x = pack_bf16x2(v[0], v[1]);
                pk.y = pack_bf16x2(v[2], v[3]);
                pk.z = pack_bf16x2(v[4], v[5]);
                pk.w = pack_bf16x2(v[6], v[7]);
                *reinterpret_cast<uint4*>(dst + i) = pk;
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[o]);
      }
      tph ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

template <int BN>
int launch_conv_mf(const __nv_bfloat16* in, int T, int H, int W, const __nv_bfloat16* weight, int Cout, const ConvParams& p,
                   cudaStream_t stream) {
  using C = MfCfg<BN>;
  CUtensorMap tmIn64, tmIn32, tmW64, tmW32;
  {
    const uint64_t dims[4] = {96, (uint64_t)W, (uint64_t)H, (uint64_t)T};
    const uint64_t strides[3] = {96 * 2, (uint64_t)W * 96 * 2, (uint64_t)H * W * 96 * 2};
    const uint32_t box64[4] = {64, TILE_W, TILE_H, 1};
    const uint32_t box32[4] = {32, TILE_W, TILE_H, 1};
    int r = make_tmap_bf16(&tmIn64, in, 4, dims, strides, box64, 128);
    if (r) return r;
    r = make_tmap_bf16(&tmIn32, in, 4, dims, strides, box32, 64);
    if (r) return r;
  }
  {
    const uint64_t K = 27 * 96;
    const uint64_t dims[2] = {K, (uint64_t)Cout};
    const uint64_t strides[1] = {K * 2};
    const uint32_t box64[2] = {64, (uint32_t)BN};
    const uint32_t box32[2] = {32, (uint32_t)BN};
    int r = make_tmap_bf16(&tmW64, weight, 2, dims, strides, box64, 128);
    if (r) return r;
    r = make_tmap_bf16(&tmW32, weight, 2, dims, strides, box32, 64);
    if (r) return r;
  }
  static bool configured = false;
  if (!configured) {
    ICB_CUDA_CHECK(cudaFuncSetAttribute(conv3x3x3_mf_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    configured = true;
  }
  const int num_tiles = p.tiles_w * p.tiles_h * ((p.T + kMfFrames - 1) / kMfFrames);
  const int grid = min(num_tiles, num_sms());
  conv3x3x3_mf_kernel<BN><<<grid, CONV_THREADS, C::SMEM_BYTES, stream>>>(tmIn64, tmIn32, tmW64, tmW32, p);
  ICB_CUDA_CHECK(cudaGetLastError());
  return IC_OK;
}

template <int BN, int BKC, int KSUB>
int launch_conv(const CUtensorMap& tmIn, const CUtensorMap& tmW, const ConvParams& p, cudaStream_t stream) {
  using C = CCfg<BN, BKC, KSUB>;
  static bool configured = false;
  if (!configured) {
    ICB_CUDA_CHECK(cudaFuncSetAttribute(conv_igemm_kernel<BN, BKC, KSUB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        C::SMEM_BYTES));
    configured = true;
  }
  const int grid = min(p.num_tiles, num_sms());
  conv_igemm_kernel<BN, BKC, KSUB><<<grid, CONV_THREADS, C::SMEM_BYTES, stream>>>(tmIn, tmW, p);
  ICB_CUDA_CHECK(cudaGetLastError());
  return IC_OK;
}

}  // namespace

int conv_igemm(const __nv_bfloat16* in, int Tin, int Hin, int Win, int Cin, const __nv_bfloat16* weight, const float* bias,
               const ConvTap* taps, int ntaps, __nv_bfloat16* out, int T, int H, int W, int Cout, int ld_out,
               const __nv_bfloat16* resid, int ld_resid, cudaStream_t stream) {
  if (!in || !weight || !out || !taps || ntaps < 1 || ntaps > MAX_TAPS) return IC_ERR_INVALID;
  if (Cin % 32 || Cout % 8 || ld_out % 8 || (resid && ld_resid % 8)) return IC_ERR_INVALID;
  if (Cout > kMaxCout) return IC_ERR_UNSUPPORTED;
  if (T <= 0 || H <= 0 || W <= 0 || Tin <= 0) return IC_ERR_INVALID;
  // Channel chunk per K-step.  Cin = 96 (the full-resolution VAE stage) has two options: three 32-channel boxes with
  // 64-byte swizzle, or (ICB_CONV_PAD64=1) two 64-channel boxes with 128-byte swizzle where the second box reaches 32
  // channels past the tensor - TMA zero-fills them (and they meet weights of the next tap, which the zeros cancel),
  // so a third of that MMA work is wasted but each pixel costs 2 requests (128 + 64 bytes) instead of 3.  ncu
  // (profiles/r2_conv96.ncu-rep) shows the 32-channel form at 35 % tensor-pipe activity with TMA delivering
  // 47 B/clk/SM of the 146 B/clk/SM the full tensor rate needs, L2 at 62 %.  Measured: the padded form is 4 % SLOWER
  // on the whole decode (1.085 vs 1.041 s), so the limit is bytes through L2, not requests - what would help is
  // fetching each input pixel once per output tile instead of once per tap.  Default stays 3 x 32.
  static int pad64 = -1;
  if (pad64 < 0) {
    const char* e = getenv("ICB_CONV_PAD64");
    pad64 = e ? atoi(e) : 0;
  }
  const bool padded = pad64 && Cin % 64 != 0 && Cin > 64;
  const int bkc = (Cin % 64 == 0 || padded) ? 64 : 32;
  const int bn = Cout >= 192 ? 192 : (Cout > 64 ? 96 : 64);

  ConvParams p;
  p.T = T;
  p.H = H;
  p.W = W;
  p.Cout = Cout;
  p.Cin = Cin;
  p.ntaps = ntaps;
  const int ksub = (bkc == 32 && Cin % 96 == 0) ? 3 : 1;
  p.k_chunks = (Cin + bkc * ksub - 1) / (bkc * ksub);
  for (int i = 0; i < ntaps; ++i) {
    p.tap[i][0] = taps[i].dt;
    p.tap[i][1] = taps[i].dh;
    p.tap[i][2] = taps[i].dw;
  }
  p.tiles_w = (W + TILE_W - 1) / TILE_W;
  p.tiles_h = (H + TILE_H - 1) / TILE_H;
  p.num_m_tiles = p.tiles_w * p.tiles_h * T;
  p.num_n_tiles = (Cout + bn - 1) / bn;
  p.num_tiles = p.num_m_tiles * p.num_n_tiles;
  p.bias = bias;
  p.resid = resid;
  p.ld_resid = ld_resid;
  p.out = out;
  p.ld_out = ld_out;

  CUtensorMap tmIn, tmW;
  {
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)Win, (uint64_t)Hin, (uint64_t)Tin};
    const uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)Win * Cin * 2, (uint64_t)Hin * Win * Cin * 2};
    const uint32_t box[4] = {(uint32_t)bkc, TILE_W, TILE_H, 1};
    int r = make_tmap_bf16(&tmIn, in, 4, dims, strides, box, bkc == 64 ? 128 : 64);
    if (r) return r;
  }
  {
    const uint64_t K = static_cast<uint64_t>(ntaps) * Cin;
    const uint64_t dims[2] = {K, (uint64_t)Cout};
    const uint64_t strides[1] = {K * 2};
    const uint32_t box[2] = {(uint32_t)bkc, (uint32_t)bn};
    int r = make_tmap_bf16(&tmW, weight, 2, dims, strides, box, bkc == 64 ? 128 : 64);
    if (r) return r;
  }
  // multi-frame path: the canonical causal 3x3x3 tap set over 96 input channels, one n-tile, same clip length in and out
  static int mf = -1;
  if (mf < 0) {
    const char* e = getenv("ICB_CONV_MF");
    mf = e ? atoi(e) : 1;
  }
  if (mf && !padded && ntaps == 27 && Cin == 96 && Tin == T && Hin == H && Win == W && Cout <= 96) {
    bool canonical = true;
    for (int i = 0; i < 27 && canonical; ++i)
      canonical = taps[i].dt == i / 9 - 2 && taps[i].dh == (i / 3) % 3 - 1 && taps[i].dw == i % 3 - 1;
    if (canonical) {
      if (bn == 96) return launch_conv_mf<96>(in, T, H, W, weight, Cout, p, stream);
      if (bn == 64) return launch_conv_mf<64>(in, T, H, W, weight, Cout, p, stream);
    }
  }
  if (bkc == 64) {
    if (bn == 192) return launch_conv<192, 64, 1>(tmIn, tmW, p, stream);
    if (bn == 96) return launch_conv<96, 64, 1>(tmIn, tmW, p, stream);
    return launch_conv<64, 64, 1>(tmIn, tmW, p, stream);
  }
  if (ksub == 3) {
    if (bn == 192) return launch_conv<192, 32, 3>(tmIn, tmW, p, stream);
    if (bn == 96) return launch_conv<96, 32, 3>(tmIn, tmW, p, stream);
    return launch_conv<64, 32, 3>(tmIn, tmW, p, stream);
  }
  if (bn == 192) return launch_conv<192, 32, 1>(tmIn, tmW, p, stream);
  if (bn == 96) return launch_conv<96, 32, 1>(tmIn, tmW, p, stream);
  return launch_conv<64, 32, 1>(tmIn, tmW, p, stream);
}

}  // namespace icb
