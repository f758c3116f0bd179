// Non-causal flash attention forward for sm_100a, head_dim 128, bf16 in / bf16 out, fp32 softmax.
//
// Replaces diffsynth's flash_attention() dispatch (FA3 -> FA2 -> SDPA) used by Wan2.1 self- and
// cross-attention (72 % of the DiT FLOPs at S = 37 440).
//
// One CTA = one head x 256 query rows (two 128-row tiles that ping-pong on the tensor core):
//   warp 0      TMA producer: Q once, then a 2-deep ring of K tiles and V^T tiles (128 keys each)
//   warp 1      tcgen05.mma issuer: S_t = Q_t K^T (SS), O_t += P_t V (TS: P read straight from TMEM)
//   warp 2      TMEM allocator (512 columns: S0 S1 O0 O1, 128 fp32 columns each)
//   warps 4-7   softmax for tile 0, warps 8-11 softmax for tile 1 (one query row per thread)
// P (bf16) overwrites the first 64 columns of its S tile, so the second GEMM needs no shared memory.
// O is rescaled lazily: only when a row maximum grows by more than 2^8 (exact after the final 1/l).
//
// Operand layouts (all bf16, token-major, head h = columns [128h, 128h+128)):
//   Q  [Sq, ldq]            K  [n_seg][seg_len, ldk]         V^T [n_seg][D, ldvt] (keys contiguous)
// K and V^T are segmented so that a multi-GPU all-gather buffer (one segment per rank) is consumed
// in place; a key tile never straddles two segments.
#include <stdlib.h>

#include <type_traits>

#include "fmha_sm100.cuh"
#include "host_util.h"

namespace icb {

namespace {

constexpr int FMHA_THREADS = 384;
constexpr int TILE = 128;             // rows per Q tile, keys per KV tile, head_dim
constexpr int HALF_BYTES = 128 * 128; // 128 rows x 64 bf16 (one swizzle-128B box)
constexpr int TILE_BYTES = 2 * HALF_BYTES;
constexpr int KV_STAGES = 2;
#ifndef ICB_FMHA_HEAD_CHUNKS
#define ICB_FMHA_HEAD_CHUNKS 3
#endif
constexpr int kHeadChunks = ICB_FMHA_HEAD_CHUNKS;  // 32-key chunks of P handed to the tensor core before the row is finished
static_assert(kHeadChunks == 2 || kHeadChunks == 3, "P is handed over after 64 or 96 keys");
constexpr int kHeadSteps = kHeadChunks * 2;  // = MMA k-steps (16 keys each) covered by those chunks
constexpr int FMHA_SMEM = 2 * TILE_BYTES + 2 * KV_STAGES * TILE_BYTES + 1024 + 256;

struct FmhaParams {
  int Sq;
  int seg_len, n_seg, tiles_per_seg, n_tiles;
  float scale_log2;  // softmax scale * log2(e)
  __nv_bfloat16* O;
  int ldo;
  // peer-memory K / V^T exchange (kSegFlags instantiation only; appended so the default kernel's parameter
  // offsets do not move): segments are consumed in ring order starting at the local one, and a remote
  // segment is touched only after its owner has raised seg_ready[seg] to this launch's epoch
  int seg_first;
  const unsigned* seg_ready;
  unsigned epoch;
  // persistent variant (kPersist): a CTA walks work items (query block, head) = item % n_qblocks, item / n_qblocks
  int n_qblocks, n_items;
};

// Spin until *flag has reached `epoch` (wrap-safe), then order the generic-proxy acquire before the
// async-proxy (TMA) reads that follow.  Bounded: a lost signal must surface as a launch failure, not a hung GPU.
__device__ __forceinline__ void wait_segment_epoch(const unsigned* flag, unsigned epoch) {
  unsigned spins = 0;
  for (;;) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (static_cast<int>(v - epoch) >= 0) break;
    __nanosleep(200);
    if (++spins > (1u << 26)) __trap();
  }
  asm volatile("fence.proxy.async;" ::: "memory");
}

// kEmuEighths: of every 8 exponential pairs, this many run on the FMA pipe instead of the MUFU.
// kSegFlags: peer-memory exchange variant (see FmhaParams); false compiles to the plain kernel.
// kWhatIf (measurement only, WRONG results; ICB_FMHA_WHATIF): bit 0 = no max tree / rescale after the first tile,
//   bit 1 = 2 of 8 exponential pairs replaced by a move, bit 2 = all exponentials replaced - how fast the kernel
//   would run if that part of the softmax were free, i.e. which part paces the tensor pipe.
// kLazy: no per-tile maximum at all.  The reference point only has to keep P inside the fp32 / bf16 exponent range, so
//   tile j > 0 is exponentiated against the reference it inherits and the ROW SUM it produces anyway is the detector:
//   a partial sum >= 2^100 (or inf / NaN) means an element would leave the range - that tile (its head before p_full
//   is raised, its tail before p_tail) is redone exactly against its true maximum; a tile sum > 2^8 merely raises the
//   reference by a whole power of two before the next tile.  Exact (softmax is shift-invariant and every rescale is
//   applied to O and l), and the 64-instruction FMNMX3 tree plus the max -> exp dependency leave the S -> P -> PV
//   critical path of every tile (what-if bound: +4 %).
// kPersist: one CTA per SM walks several (query block, head) items, keeping TMEM, barriers and the pipelines alive
//   across them - for short key sequences (cross-attention: 4 KV tiles) where set-up, pipeline fill and drain of a
//   CTA cost as much as its attention.  Barrier parities run on a tile counter that continues across items; the next
//   item's Q is loaded as soon as the last S MMA of the current one has consumed the old one (q_empty), and the first
//   PV of an item waits until the softmax warps have read the previous item's O out of TMEM (o_free).
template <int kEmuEighths, bool kSegFlags, int kWhatIf = 0, bool kLazy = false, bool kPersist = false>
__global__ void __launch_bounds__(FMHA_THREADS, 1)
fmha_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmVT, const FmhaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;                                  // [2 tiles][2 halves][128 x 128 B]
  uint8_t* smem_k = smem + 2 * TILE_BYTES;                 // [KV_STAGES][2 halves (d)][128 keys x 128 B]
  uint8_t* smem_v = smem_k + KV_STAGES * TILE_BYTES;       // [KV_STAGES][2 halves (keys)][128 d x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_v + KV_STAGES * TILE_BYTES);
  uint64_t* q_full = bars;          // 1
  uint64_t* k_full = bars + 1;      // [2]
  uint64_t* k_empty = bars + 3;     // [2]
  uint64_t* v_full = bars + 5;      // [2]
  uint64_t* v_empty = bars + 7;     // [2]
  uint64_t* s_full = bars + 9;      // [2] per Q tile: S ready in TMEM
  uint64_t* p_full = bars + 11;     // [2] per Q tile: P written (and O rescaled)
  uint64_t* pv_done = bars + 13;    // [2] per Q tile: O += P V retired
  uint64_t* p_tail = bars + 15;     // [2] per Q tile: last quarter of P written
  uint64_t* pv_head = bars + 17;    // [2] per Q tile: the head part of O += P V retired (kLazy's tail redo waits on it)
  uint64_t* q_empty = bars + 19;    // kPersist: the S MMAs of an item have consumed Q
  uint64_t* o_free = bars + 20;     // [2] kPersist, per Q tile: the softmax warps have read the item's O out of TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);
  static_assert(!(kPersist && (kLazy || kSegFlags || kWhatIf != 0)), "the persistent variant is the plain kernel only");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_items = kPersist ? p.n_items : 1;        // items of this launch; a non-persistent CTA has exactly one
  const int item0 = kPersist ? static_cast<int>(blockIdx.x) : 0;
  const int item_step = kPersist ? static_cast<int>(gridDim.x) : 1;
  auto item_head = [&](int item) { return kPersist ? item / p.n_qblocks : static_cast<int>(blockIdx.y); };
  auto item_q0 = [&](int item) { return (kPersist ? item % p.n_qblocks : static_cast<int>(blockIdx.x)) * (2 * TILE); };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmVT);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 4);
      mbar_init(&p_tail[s], 4);
      mbar_init(&pv_done[s], 1);
      mbar_init(&pv_head[s], 1);
      mbar_init(&o_free[s], 4);
    }
    mbar_init(q_empty, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
  // register budget: 384 threads x 168 at launch = 64512 = 128 x 88 + 256 x 208
  reg_dealloc<88>();  // warpgroup 0 (TMA / MMA / allocator) hands its registers to the softmax groups
  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    {
      const bool leader = elect_one();  // warp-uniform control flow, one lane issues
      int g0 = 0, it = 0;                // tiles issued by earlier items, items done
      for (int item = item0; item < n_items; item += item_step, g0 += p.n_tiles, ++it) {
      const int head = item_head(item), q0 = item_q0(item);
      if (kPersist && it > 0) mbar_wait(q_empty, (it - 1) & 1);
      if (leader) {
        mbar_arrive_expect_tx(q_full, 2 * TILE_BYTES);
        for (int t = 0; t < 2; ++t)
          for (int hf = 0; hf < 2; ++hf)
            tma_load_2d(smem_q + t * TILE_BYTES + hf * HALF_BYTES, &tmQ, q_full, head * TILE + hf * 64,
                        q0 + t * TILE, kEvictFirst);
      }
      for (int j = 0; j < p.n_tiles; ++j) {
        int seg = j / p.tiles_per_seg;
        const int key0 = (j - seg * p.tiles_per_seg) * TILE;
        if constexpr (kSegFlags) {
          seg += p.seg_first;  // ring order: local segment first, then the peers in the order their pushes are issued
          if (seg >= p.n_seg) seg -= p.n_seg;
          if (key0 == 0 && seg != p.seg_first) {
            if (leader) wait_segment_epoch(p.seg_ready + seg, p.epoch);
            __syncwarp();
          }
        }
        const int st = (g0 + j) & 1;
        const uint32_t ph = ((g0 + j) >> 1) & 1;
        mbar_wait(&k_empty[st], ph ^ 1);
        if (leader) {
          mbar_arrive_expect_tx(&k_full[st], TILE_BYTES);
          for (int hf = 0; hf < 2; ++hf)
            tma_load_3d(smem_k + st * TILE_BYTES + hf * HALF_BYTES, &tmK, &k_full[st], head * TILE + hf * 64, key0,
                        seg, kEvictLast);
        }
        mbar_wait(&v_empty[st], ph ^ 1);
        if (leader) {
          mbar_arrive_expect_tx(&v_full[st], TILE_BYTES);
          for (int hf = 0; hf < 2; ++hf)
            tma_load_3d(smem_v + st * TILE_BYTES + hf * HALF_BYTES, &tmVT, &v_full[st], key0 + hf * 64, head * TILE,
                        seg, kEvictLast);
        }
      }
      }  // items
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------
    // The whole warp runs the control flow (warp-uniform values stay in uniform registers, which is what
    // UTCHMMA consumes); only the elected lane issues the tcgen05 instructions.
    {
      const bool leader = elect_one();
      constexpr uint32_t idesc = umma_idesc_bf16(TILE, TILE);
      const uint64_t d0 = umma_desc_sw128_kmajor(smem_u32(smem_q));
      const uint32_t desc_hi = static_cast<uint32_t>(d0 >> 32);
      const uint32_t q_lo = static_cast<uint32_t>(d0);
      const uint32_t k_lo = static_cast<uint32_t>(umma_desc_sw128_kmajor(smem_u32(smem_k)));
      const uint32_t v_lo = static_cast<uint32_t>(umma_desc_sw128_kmajor(smem_u32(smem_v)));
      auto issue_s = [&](int t, int kst) {
        const uint32_t a0 = q_lo + ((t * TILE_BYTES) >> 4);
        const uint32_t b0 = k_lo + ((kst * TILE_BYTES) >> 4);
        if (leader) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint32_t off = ((k >> 2) * HALF_BYTES + (k & 3) * 32) >> 4;
            umma_ss_lo(tmem_base + t * TILE, a0 + off, b0 + off, desc_hi, idesc, k > 0);
          }
        }
      };
      auto issue_pv = [&](int t, int vst, bool acc, int k0, int k1) {
        const uint32_t b0 = v_lo + ((vst * TILE_BYTES) >> 4);
        if (leader) {
#pragma unroll
          for (int k = k0; k < k1; ++k) {
            const uint32_t off = ((k >> 2) * HALF_BYTES + (k & 3) * 32) >> 4;
            // P tile: bf16 pairs packed in 32-bit columns, 16 keys = 8 columns per MMA
            umma_ts_lo(tmem_base + 256 + t * TILE, tmem_base + t * TILE + k * 8, b0 + off, desc_hi, idesc, acc || k > 0);
          }
        }
      };
      auto commit = [&](uint64_t* bar) {
        if (leader) umma_commit(bar);
      };
      int g0 = 0, it = 0;
      for (int item = item0; item < n_items; item += item_step, g0 += p.n_tiles, ++it) {
      mbar_wait(q_full, it & 1);
      mbar_wait(&k_full[g0 & 1], (g0 >> 1) & 1);
      tc_fence_after();
      issue_s(0, g0 & 1);
      commit(&s_full[0]);
      issue_s(1, g0 & 1);
      commit(&s_full[1]);
      commit(&k_empty[g0 & 1]);
      for (int j = 0; j < p.n_tiles; ++j) {
        const int gj = g0 + j;
        const int st = gj & 1;
        const uint32_t ph = (gj >> 1) & 1;
        const bool more = j + 1 < p.n_tiles;
        const int st1 = (gj + 1) & 1;
        const uint32_t ph1 = ((gj + 1) >> 1) & 1;
        mbar_wait(&v_full[st], ph);
        // P arrives in two parts (keys [0, 96) then [96, 128)): the tensor core starts on the first part
        // while the softmax warps are still exponentiating the last quarter
        mbar_wait(&p_full[0], gj & 1);
        if (kPersist && j == 0 && it > 0) mbar_wait(&o_free[0], (it - 1) & 1);  // the first PV of an item overwrites O
        tc_fence_after();
        issue_pv(0, st, j > 0, 0, kHeadSteps);
        if constexpr (kLazy) commit(&pv_head[0]);
        mbar_wait(&p_tail[0], gj & 1);
        tc_fence_after();
        issue_pv(0, st, true, kHeadSteps, 8);
        commit(&pv_done[0]);
        if (more) {
          mbar_wait(&k_full[st1], ph1);
          tc_fence_after();
          issue_s(0, st1);
          commit(&s_full[0]);
        }
        mbar_wait(&p_full[1], gj & 1);
        if (kPersist && j == 0 && it > 0) mbar_wait(&o_free[1], (it - 1) & 1);
        tc_fence_after();
        issue_pv(1, st, j > 0, 0, kHeadSteps);
        if constexpr (kLazy) commit(&pv_head[1]);
        mbar_wait(&p_tail[1], gj & 1);
        tc_fence_after();
        issue_pv(1, st, true, kHeadSteps, 8);
        commit(&pv_done[1]);
        commit(&v_empty[st]);
        if (more) {
          issue_s(1, st1);
          commit(&s_full[1]);
          commit(&k_empty[st1]);
        }
      }
      if (kPersist) commit(q_empty);  // every S MMA of this item has been issued: Q may be replaced once they retire
      }  // items
    }
  }
  } else {
    // ------------------------------- softmax + output ---------------------------
    reg_alloc<208>();
    const int t = (warp - 4) >> 2;
    const int quad = warp & 3;
    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_sel + t * TILE;
    const uint32_t o_addr = tmem_base + lane_sel + 256 + t * TILE;
    const float sc = p.scale_log2;
    int g0 = 0;  // tiles of earlier items: barrier parities continue across the items of a persistent CTA
    for (int item = item0; item < n_items; item += item_step, g0 += p.n_tiles) {
    const int head = item_head(item);
    const int row = item_q0(item) + t * TILE + quad * 32 + lane;
    float m_ref = 0.f;
    float l = 0.f;
    if constexpr (kLazy) {
      static_assert(!kLazy || kHeadChunks == 3, "the max-free path hands P over after chunk 2");
      float nms = 0.f;      // -(reference point) * scale_log2: P = exp2(s * sc + nms)
      bool grow = false;    // the last tile's row sum exceeded 2^8: raise the reference by 2^kk before the next tile
      int kk = 0;
      const unsigned long long sc2 = pk2(sc, sc);
      // multiply this thread's O row (and nothing else) by alpha; warp-collective TMEM traffic
      auto scale_o = [&](float alpha) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t raw[32];
          tmem_ld_x32(o_addr + c * 32, raw);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) raw[i] = __float_as_uint(__uint_as_float(raw[i]) * alpha);
          tmem_st_x32(o_addr + c * 32, raw);
        }
        tmem_wait_st();
      };
      for (int j = 0; j < p.n_tiles; ++j) {
        const int seg = j / p.tiles_per_seg;
        const int valid = p.seg_len - (j - seg * p.tiles_per_seg) * TILE;
        if (j > 0 && __any_sync(0xffffffffu, grow)) {
          mbar_wait(&pv_done[t], (j - 1) & 1);  // O_t must be quiescent
          tc_fence_after();
          const float alpha = grow ? __int_as_float((127 - kk) << 23) : 1.0f;  // exact 2^-kk
          if (grow) nms -= static_cast<float>(kk);
          l *= alpha;
          scale_o(alpha);
          grow = false;
        }
        mbar_wait(&s_full[t], j & 1);
        tc_fence_after();
        uint32_t s[TILE];
        tmem_ld_x32(s_addr, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
        tmem_ld_x32(s_addr + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
        tmem_ld_x32(s_addr + 64, *reinterpret_cast<uint32_t(*)[32]>(&s[64]));
        tmem_ld_x32(s_addr + 96, *reinterpret_cast<uint32_t(*)[32]>(&s[96]));
        tmem_wait_ld();
        if (valid < TILE) {
#pragma unroll
          for (int i = 0; i < TILE; ++i)
            if (i >= valid) s[i] = 0xff800000u;  // -inf
        }
        auto p_chunk = [&](auto cc, unsigned long long& a0, unsigned long long& a1) {
          constexpr int c = decltype(cc)::value;
          const unsigned long long nms2 = pk2(nms, nms);
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const unsigned long long x2 =
                fma2(pk2(__uint_as_float(s[c * 32 + i]), __uint_as_float(s[c * 32 + i + 1])), sc2, nms2);
            unsigned long long p2;
            if (((i >> 1) & 7) < kEmuEighths) {
              p2 = ex2_emu2_clamped(x2);
            } else {
              float x0, x1;
              upk2(x2, x0, x1);
              p2 = pk2(ex2_approx(x0), ex2_approx(x1));
            }
            if (i & 2)
              a1 = add2(a1, p2);
            else
              a0 = add2(a0, p2);
            float p0, p1;
            upk2(p2, p0, p1);
            pk[i >> 1] = pack_bf16x2(p0, p1);
          }
          tmem_st_x16(s_addr + c * 16, pk);
        };
        auto hsum2 = [&](unsigned long long a0, unsigned long long a1) {
          float x0, x1, y0, y1;
          upk2(a0, x0, x1);
          upk2(a1, y0, y1);
          return (x0 + x1) + (y0 + y1);
        };
        auto row_max = [&](int lo, int hi) {  // exact maximum of s[lo, hi) (slow paths and the first tile only)
          float m = __uint_as_float(s[lo]);
#pragma unroll
          for (int i = lo + 1; i + 1 < hi; i += 2) m = max3(m, __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
          return fmaxf(m, __uint_as_float(s[hi - 1]));
        };
        unsigned long long acc0 = 0ull, acc1 = 0ull;
        if (j == 0) nms = -row_max(0, TILE) * sc;
        p_chunk(std::integral_constant<int, 0>{}, acc0, acc1);
        p_chunk(std::integral_constant<int, 1>{}, acc0, acc1);
        p_chunk(std::integral_constant<int, 2>{}, acc0, acc1);
        float hsum = hsum2(acc0, acc1);
        if (j > 0 && __any_sync(0xffffffffu, !(hsum < 0x1p100f))) {
          // an element of this tile would leave the exponent range against the inherited reference: redo the tile
          // exactly against its true maximum (the reference only ever rises)
          const float nms_new = fminf(nms, -row_max(0, TILE) * sc);
          mbar_wait(&pv_done[t], (j - 1) & 1);
          tc_fence_after();
          const float alpha = ex2_approx(nms_new - nms);
          nms = nms_new;
          l *= alpha;
          scale_o(alpha);
          acc0 = 0ull;
          acc1 = 0ull;
          p_chunk(std::integral_constant<int, 0>{}, acc0, acc1);
          p_chunk(std::integral_constant<int, 1>{}, acc0, acc1);
          p_chunk(std::integral_constant<int, 2>{}, acc0, acc1);
          hsum = hsum2(acc0, acc1);
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
        unsigned long long t0 = 0ull, t1 = 0ull;
        p_chunk(std::integral_constant<int, 3>{}, t0, t1);
        float tsum = hsum2(t0, t1);
        if (j > 0 && __any_sync(0xffffffffu, !(tsum < 0x1p100f))) {
          // same for the tail chunk, after the head has been handed to the tensor core: wait until the head part of
          // O += P V has retired (which implies every earlier MMA has), then rescale O, l and the head's row sum
          const float nms_new = fminf(nms, -row_max(96, TILE) * sc);
          mbar_wait(&pv_head[t], j & 1);
          tc_fence_after();
          const float alpha = ex2_approx(nms_new - nms);
          nms = nms_new;
          l *= alpha;
          hsum *= alpha;
          scale_o(alpha);
          t0 = 0ull;
          t1 = 0ull;
          p_chunk(std::integral_constant<int, 3>{}, t0, t1);
          tsum = hsum2(t0, t1);
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_tail[t]);
        tsum += hsum;
        l += tsum;
        grow = tsum > 256.0f;
        kk = static_cast<int>((__float_as_uint(tsum) >> 23) & 0xffu) - 126;  // floor(log2(tsum)) + 1 (tsum <= 2^107 here)
      }
    } else
    for (int j = 0; j < p.n_tiles; ++j) {
      const int seg = j / p.tiles_per_seg;
      const int valid = p.seg_len - (j - seg * p.tiles_per_seg) * TILE;  // >= 1; < 128 only on a segment tail
      mbar_wait(&s_full[t], (g0 + j) & 1);
      tc_fence_after();
      // the whole 128-wide score row of this thread lives in registers (one TMEM round trip)
      uint32_t s[TILE];
      const unsigned long long sc2 = pk2(sc, sc);
      unsigned long long acc0 = 0ull, acc1 = 0ull;  // packed partial row sums
      // P chunk c = exp2(S*sc - m_ref*sc) of keys [32c, 32c+32) as bf16 over the head of the S tile; packed
      // fp32x2 math, part of the exponentials on the FMA pipe (ex2_emu2), the rest on the MUFU
      auto p_chunk = [&](auto cc, unsigned long long& a0, unsigned long long& a1) {
        constexpr int c = decltype(cc)::value;
        const float nms = -m_ref * sc;
        const unsigned long long nms2 = pk2(nms, nms);
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const unsigned long long x2 =
              fma2(pk2(__uint_as_float(s[c * 32 + i]), __uint_as_float(s[c * 32 + i + 1])), sc2, nms2);
          unsigned long long p2;
          if ((kWhatIf & 4) || ((kWhatIf & 2) && ((i >> 1) & 7) < 2)) {
            p2 = x2;
          } else if (((i >> 1) & 7) < kEmuEighths) {
            p2 = ex2_emu2(x2);
          } else {
            float x0, x1;
            upk2(x2, x0, x1);
            p2 = pk2(ex2_approx(x0), ex2_approx(x1));
          }
          if (i & 2)
            a1 = add2(a1, p2);
          else
            a0 = add2(a0, p2);
          float p0, p1;
          upk2(p2, p0, p1);
          pk[i >> 1] = pack_bf16x2(p0, p1);
        }
        tmem_st_x16(s_addr + c * 16, pk);
      };
      tmem_ld_x32(s_addr, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
      tmem_ld_x32(s_addr + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
      tmem_ld_x32(s_addr + 64, *reinterpret_cast<uint32_t(*)[32]>(&s[64]));
      tmem_ld_x32(s_addr + 96, *reinterpret_cast<uint32_t(*)[32]>(&s[96]));
      tmem_wait_ld();
      // row maximum: 8 independent chains (ALU pipe), chain c over scores [16c, 16c+16): mxa[c] starts at s[16c],
      // steps k = 0..6 fold two scores each (FMNMX3), step 7 the last one.  Step o of 64 = (chain o % 8, k = o / 8).
      float mxa[8];
      auto max_step = [&](int o) {
        const int c = o & 7, k = o >> 3;
        if (k < 7)
          mxa[c] = max3(mxa[c], __uint_as_float(s[16 * c + 1 + 2 * k]), __uint_as_float(s[16 * c + 2 + 2 * k]));
        else
          mxa[c] = fmaxf(mxa[c], __uint_as_float(s[16 * c + 15]));
      };
      if (!(kWhatIf & 1) || j == 0) {
        if (valid < TILE) {
#pragma unroll
          for (int i = 0; i < TILE; ++i)
            if (i >= valid) s[i] = 0xff800000u;  // -inf
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) mxa[c] = __uint_as_float(s[16 * c]);
#pragma unroll
        for (int o = 0; o < 64; ++o) max_step(o);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) mxa[i] = m_ref;
      }
      const float mx = fmaxf(max3(mxa[0], mxa[1], mxa[2]), max3(max3(mxa[3], mxa[4], mxa[5]), mxa[6], mxa[7]));
      if (j == 0) {
        m_ref = mx;
      } else {
        const bool need = (mx - m_ref) * sc > 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          // O_t must be quiescent: wait until PV_t(j-1) has retired
          mbar_wait(&pv_done[t], (g0 + j - 1) & 1);
          tc_fence_after();
          const float m_new = fmaxf(m_ref, mx);
          const float alpha = ex2_approx((m_ref - m_new) * sc);
          m_ref = m_new;
          l *= alpha;
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t raw[32];
            tmem_ld_x32(o_addr + c * 32, raw);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) raw[i] = __float_as_uint(__uint_as_float(raw[i]) * alpha);
            tmem_st_x32(o_addr + c * 32, raw);
          }
          tmem_wait_st();
        }
      }
      p_chunk(std::integral_constant<int, 0>{}, acc0, acc1);
      p_chunk(std::integral_constant<int, 1>{}, acc0, acc1);
      if constexpr (kHeadChunks == 3) p_chunk(std::integral_constant<int, 2>{}, acc0, acc1);
      // first part of P is in TMEM: let the tensor core start
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
      if constexpr (kHeadChunks == 2) p_chunk(std::integral_constant<int, 2>{}, acc0, acc1);
      p_chunk(std::integral_constant<int, 3>{}, acc0, acc1);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_tail[t]);
      {
        float a0, a1, b0, b1;
        upk2(acc0, a0, a1);
        upk2(acc1, b0, b1);
        l += (a0 + a1) + (b0 + b1);
      }
    }
    // epilogue: wait for the last PV, normalise, store
    mbar_wait(&pv_done[t], (g0 + p.n_tiles - 1) & 1);
    tc_fence_after();
    const float inv = 1.0f / l;
    __nv_bfloat16* dst = p.O + static_cast<size_t>(row) * p.ldo + head * TILE;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t raw[32];
      tmem_ld_x32(o_addr + c * 32, raw);
      tmem_wait_ld();
      if (row < p.Sq) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 pk;
          pk.x = pack_bf16x2(__uint_as_float(raw[i]) * inv, __uint_as_float(raw[i + 1]) * inv);
          pk.y = pack_bf16x2(__uint_as_float(raw[i + 2]) * inv, __uint_as_float(raw[i + 3]) * inv);
          pk.z = pack_bf16x2(__uint_as_float(raw[i + 4]) * inv, __uint_as_float(raw[i + 5]) * inv);
          pk.w = pack_bf16x2(__uint_as_float(raw[i + 6]) * inv, __uint_as_float(raw[i + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + c * 32 + i) = pk;
        }
      }
    }
    if (kPersist) {  // O of this item has left TMEM: the next item's first PV may overwrite it
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[t]);
    }
    }  // items
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace

int fmha_fwd(const __nv_bfloat16* Q, int ldq, const __nv_bfloat16* K, int ldk, long long k_seg_stride,
             const __nv_bfloat16* VT, int ldvt, long long vt_seg_stride, __nv_bfloat16* O, int ldo, int Sq,
             int seg_len, int n_seg, int n_heads, float softmax_scale, cudaStream_t stream, int seg_first,
             const unsigned* seg_ready, unsigned epoch) {
  if (Sq <= 0 || seg_len <= 0 || n_seg <= 0 || n_heads <= 0) return IC_ERR_INVALID;
  if (seg_first < 0 || seg_first >= n_seg) return IC_ERR_INVALID;
  if ((ldq % 8) || (ldk % 8) || (ldvt % 8) || (ldo % 8) || (k_seg_stride % 8) || (vt_seg_stride % 8))
    return IC_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(Q) | reinterpret_cast<uintptr_t>(K) | reinterpret_cast<uintptr_t>(VT) |
       reinterpret_cast<uintptr_t>(O)) & 15)
    return IC_ERR_INVALID;
  const int D = n_heads * TILE;

  CUtensorMap tmQ, tmK, tmVT;
  {
    const uint64_t dims[2] = {(uint64_t)D, (uint64_t)Sq};
    const uint64_t strides[1] = {(uint64_t)ldq * 2};
    const uint32_t box[2] = {64, TILE};
    int r = make_tmap_bf16(&tmQ, Q, 2, dims, strides, box);
    if (r) return r;
  }
  {
    const uint64_t dims[3] = {(uint64_t)D, (uint64_t)seg_len, (uint64_t)n_seg};
    const uint64_t strides[2] = {(uint64_t)ldk * 2, (uint64_t)(n_seg > 1 ? k_seg_stride : (long long)ldk * seg_len) * 2};
    const uint32_t box[3] = {64, TILE, 1};
    int r = make_tmap_bf16(&tmK, K, 3, dims, strides, box);
    if (r) return r;
  }
  {
    const uint64_t dims[3] = {(uint64_t)seg_len, (uint64_t)D, (uint64_t)n_seg};
    const uint64_t strides[2] = {(uint64_t)ldvt * 2, (uint64_t)(n_seg > 1 ? vt_seg_stride : (long long)ldvt * D) * 2};
    const uint32_t box[3] = {64, TILE, 1};
    int r = make_tmap_bf16(&tmVT, VT, 3, dims, strides, box);
    if (r) return r;
  }

  FmhaParams p;
  p.Sq = Sq;
  p.seg_len = seg_len;
  p.n_seg = n_seg;
  p.tiles_per_seg = (seg_len + TILE - 1) / TILE;
  p.n_tiles = p.tiles_per_seg * n_seg;
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  p.O = O;
  p.ldo = ldo;
  p.seg_first = seg_first;
  p.seg_ready = seg_ready;
  p.epoch = epoch;

  static int emu = -1, lazy = 0, whatif = 0;
  if (emu < 0) {
    // share of the exponentials computed on the FMA pipe (ex2_emu2), in eighths.  Measured on B200 after the max-tree
    // fix (profiles/r2_fmha_variants.json, S = 37 440): 0 -> 1418, 1 -> 1427, 2 -> 1437 TFLOP/s isolated; in the
    // denoising step 1192 / 1201 / 1212 TFLOP/s - MUFU and tensor pipe need the same 2 048 clocks per KV step, so
    // taking a quarter of the MUFU work away is what lets the two overlap.  Results agree to the same 1.9e-3 with
    // the fp32 oracle in every setting (the emulation's 9e-5 error is far below P's bf16 rounding).
    const char* e = getenv("ICB_FMHA_EMU");
    emu = e ? atoi(e) : 2;
    if (emu < 0 || emu > 4) emu = 2;
    if (const char* v = getenv("ICB_FMHA_LAZY")) lazy = atoi(v) != 0;   // max-free softmax: correct, measured slower
#ifdef ICB_FMHA_WHATIF_BUILD  // developer builds only (build.py: ICB_NVCC_EXTRA): measurement kernels, wrong results
    if (const char* w = getenv("ICB_FMHA_WHATIF")) whatif = atoi(w);
#endif
  }
  dim3 grid((Sq + 2 * TILE - 1) / (2 * TILE), n_heads);
#define ICB_FMHA_LAUNCH(...)                                                                                          \
  do {                                                                                                                \
    static bool configured = false;                                                                                   \
    if (!configured) {                                                                                                \
      ICB_CUDA_CHECK(cudaFuncSetAttribute(fmha_fwd_kernel<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                          FMHA_SMEM));                                                                \
      configured = true;                                                                                              \
    }                                                                                                                 \
    fmha_fwd_kernel<__VA_ARGS__><<<grid, FMHA_THREADS, FMHA_SMEM, stream>>>(tmQ, tmK, tmVT, p);                        \
  } while (0)
  const bool seg = seg_ready != nullptr;
  static int persist = -1;
  if (persist < 0) {
    const char* e = getenv("ICB_FMHA_PERSIST");
    persist = e ? atoi(e) : 1;
  }
  p.n_qblocks = static_cast<int>(grid.x);
  p.n_items = static_cast<int>(grid.x) * n_heads;
  if (persist && !seg && !whatif && !lazy && p.n_tiles <= 8 && p.n_items > num_sms()) {
    // short key sequences (cross-attention): one persistent CTA per SM walks the (query block, head) items
    grid = dim3(static_cast<unsigned>(num_sms()), 1, 1);
    if (emu == 0) ICB_FMHA_LAUNCH(0, false, 0, false, true); else ICB_FMHA_LAUNCH(2, false, 0, false, true);
#ifdef ICB_FMHA_WHATIF_BUILD
  } else if (whatif && !seg) {
    switch (whatif) {
      case 1: ICB_FMHA_LAUNCH(0, false, 1); break;
      case 2: ICB_FMHA_LAUNCH(0, false, 2); break;
      case 3: ICB_FMHA_LAUNCH(0, false, 3); break;
      case 4: ICB_FMHA_LAUNCH(0, false, 4); break;
      case 5: ICB_FMHA_LAUNCH(0, false, 5); break;
      default: return IC_ERR_INVALID;
    }
#endif
  } else if (lazy && kHeadChunks == 3) {
    if (seg) {
      if (emu == 0) ICB_FMHA_LAUNCH(0, true, 0, true); else ICB_FMHA_LAUNCH(2, true, 0, true);
    } else {
      if (emu == 0) ICB_FMHA_LAUNCH(0, false, 0, true);
      else if (emu == 1) ICB_FMHA_LAUNCH(1, false, 0, true);
      else ICB_FMHA_LAUNCH(2, false, 0, true);
    }
  } else if (seg) {  // peer-memory exchange variant: waits per remote segment
    if (emu == 0) ICB_FMHA_LAUNCH(0, true); else ICB_FMHA_LAUNCH(2, true);
  } else {
    switch (emu) {
      case 0: ICB_FMHA_LAUNCH(0, false); break;
      case 1: ICB_FMHA_LAUNCH(1, false); break;
      case 2: ICB_FMHA_LAUNCH(2, false); break;
      case 3: ICB_FMHA_LAUNCH(3, false); break;
      default: ICB_FMHA_LAUNCH(4, false); break;
    }
  }
#undef ICB_FMHA_LAUNCH
  ICB_CUDA_CHECK(cudaGetLastError());
  return IC_OK;
}

}  // namespace icb
