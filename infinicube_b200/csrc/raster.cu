// Voxel -> guidance-buffer rasteriser for sm_100a (HBM/L2-bound integer + fp32 work, no tensor cores).
//
// Replaces, in the reference checkout of nv-tlabs/InfiniCube:
//   points_to_fvdb (fvdb.gridbatch_from_points + per-label torch_scatter loop) . utils/fvdb_utils.py:71-216
//   CameraBase.get_zdepth_map_from_voxel / get_semantic_map_from_voxel .......... camera/base.py:520-618
//     (fvdb GridBatch.segments_along_rays / voxels_along_rays, three launches per frame)
//   semantic_to_color + generate_rgb_semantic_buffer (CPU numpy) ................ utils/semantic_utils.py:88-131
//   generate_coordinate_buffer_from_memory_global_norm + unproject_depth_torch .. utils/buffer_utils.py:180-265
//
// Data layout in HBM: the sparse grid is a dense array of 8^3 bricks over the voxel bounding box; each
// brick is 8 x uint64 occupancy words (word = local z, bit = local y*8 + local x) plus a prefix count
// `base` so that voxel index = base[brick] + rank(bit).  Labels are compact int32 arrays in voxel-index
// order.  One fused ray-march produces depth + semantic + instance for every camera in one launch.
//
// The fp32 operation order of the traversal is a specification shared with oracle/raster_oracle.c
// (written independently): every arithmetic step uses the explicit round-to-nearest intrinsics so that
// nvcc cannot contract to FMA, and the integer outputs must agree bit for bit.
#include <stdint.h>
#include <stdlib.h>

#include <vector>

#include "../../include/infinicube_b200.h"
#include "host_util.h"

using namespace icb;

struct ic_grid {
  float vs[3], org[3];
  int imin[3], imax[3];
  int bmin[3], bdim[3];
  long long n_bricks = 0, n_vox = 0;
  unsigned long long* mask = nullptr;  // [n_bricks][8]
  int* base = nullptr;                 // [n_bricks]
  unsigned char* occ = nullptr;        // [n_bricks] 1 = the brick holds at least one voxel
  int* sem = nullptr;                  // [n_vox]
  int* inst = nullptr;                 // [n_vox]
};

namespace {

struct GridView {
  float org[3], vs[3];
  int bmin[3], bdim[3];
  const unsigned long long* mask;
  const int* base;
  const unsigned char* occ;
  const int* sem;
  const int* inst;
};

__device__ __forceinline__ int floordiv8(int a) { return a >> 3; }

// ------------------------------------------------------------------------------------------------
// grid build
// ------------------------------------------------------------------------------------------------
__global__ void ijk_bbox_kernel(const float* __restrict__ pts, long long m, float3 org, float3 vs, int* __restrict__ ijk,
                                int* __restrict__ bbox /*[6]: min xyz, max xyz*/) {
  const long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  int v[3] = {INT_MAX, INT_MAX, INT_MAX};
  int w[3] = {INT_MIN, INT_MIN, INT_MIN};
  if (p < m) {
    const float o[3] = {org.x, org.y, org.z};
    const float s[3] = {vs.x, vs.y, vs.z};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float q = __fdiv_rn(__fsub_rn(pts[3 * p + a], o[a]), s[a]);
      const int i = static_cast<int>(rintf(q));  // round half to even (torch.round().long())
      ijk[3 * p + a] = i;
      v[a] = w[a] = i;
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      v[a] = min(v[a], __shfl_xor_sync(0xffffffffu, v[a], off));
      w[a] = max(w[a], __shfl_xor_sync(0xffffffffu, w[a], off));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      atomicMin(&bbox[a], v[a]);
      atomicMax(&bbox[3 + a], w[a]);
    }
  }
}

__device__ __forceinline__ long long brick_lin(const int* bmin, const int* bdim, int bx, int by, int bz) {
  return (static_cast<long long>(bz - bmin[2]) * bdim[1] + (by - bmin[1])) * bdim[0] + (bx - bmin[0]);
}

struct BrickGeom {
  int bmin[3], bdim[3];
};

__global__ void set_bits_kernel(const int* __restrict__ ijk, long long m, BrickGeom g, unsigned long long* __restrict__ mask) {
  const long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= m) return;
  const int i = ijk[3 * p], j = ijk[3 * p + 1], k = ijk[3 * p + 2];
  const long long b = brick_lin(g.bmin, g.bdim, floordiv8(i), floordiv8(j), floordiv8(k));
  atomicOr(&mask[b * 8 + (k & 7)], 1ull << ((j & 7) * 8 + (i & 7)));
}

// three-phase exclusive scan of per-brick popcounts (1024 bricks per block)
__global__ void brick_count_kernel(const unsigned long long* __restrict__ mask, long long n_bricks, int* __restrict__ base,
                                   int* __restrict__ block_sums, unsigned char* __restrict__ occ) {
  __shared__ int sh[1024];
  const long long b = static_cast<long long>(blockIdx.x) * 1024 + threadIdx.x;
  int c = 0;
  if (b < n_bricks) {
#pragma unroll
    for (int z = 0; z < 8; ++z) c += __popcll(mask[b * 8 + z]);
  }
  sh[threadIdx.x] = c;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {  // Hillis-Steele inclusive scan
    int v = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
    __syncthreads();
    sh[threadIdx.x] += v;
    __syncthreads();
  }
  if (b < n_bricks) {
    base[b] = sh[threadIdx.x] - c;  // exclusive within block
    occ[b] = c > 0;
  }
  if (threadIdx.x == 1023) block_sums[blockIdx.x] = sh[1023];
}
__global__ void scan_block_sums_kernel(int* __restrict__ block_sums, int n_blocks, long long* __restrict__ total) {
  // single thread block, sequential over chunks of 1024
  __shared__ int sh[1024];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int start = 0; start < n_blocks; start += 1024) {
    const int i = start + threadIdx.x;
    const int c = i < n_blocks ? block_sums[i] : 0;
    sh[threadIdx.x] = c;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
      int v = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
      __syncthreads();
      sh[threadIdx.x] += v;
      __syncthreads();
    }
    if (i < n_blocks) block_sums[i] = carry + sh[threadIdx.x] - c;
    __syncthreads();
    if (threadIdx.x == 0) carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}
__global__ void add_block_offsets_kernel(int* __restrict__ base, long long n_bricks, const int* __restrict__ block_sums) {
  const long long b = static_cast<long long>(blockIdx.x) * 1024 + threadIdx.x;
  if (b < n_bricks) base[b] += block_sums[blockIdx.x];
}

__device__ __forceinline__ int voxel_index(const GridView& g, int i, int j, int k) {
  const long long b = brick_lin(g.bmin, g.bdim, floordiv8(i), floordiv8(j), floordiv8(k));
  const unsigned long long* m = g.mask + b * 8;
  const int lz = k & 7, bit = (j & 7) * 8 + (i & 7);
  int r = 0;
  for (int z = 0; z < lz; ++z) r += __popcll(m[z]);
  r += __popcll(m[lz] & ((1ull << bit) - 1ull));
  return g.base[b] + r;
}

// per-(voxel,label) counting in an open-addressing table, then arg-max with ties -> smallest label
constexpr unsigned long long kEmptyKey = 0xFFFFFFFFFFFFFFFFull;

__global__ void label_count_kernel(const int* __restrict__ ijk, const int* __restrict__ lab, long long m, GridView g,
                                   unsigned long long* __restrict__ keys, unsigned int* __restrict__ cnt,
                                   unsigned long long cap_mask) {
  const long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= m) return;
  const int vox = voxel_index(g, ijk[3 * p], ijk[3 * p + 1], ijk[3 * p + 2]);
  const unsigned long long key = (static_cast<unsigned long long>(static_cast<unsigned int>(vox)) << 32) |
                                 static_cast<unsigned int>(lab[p]);
  unsigned long long h = key * 0x9E3779B97F4A7C15ull;
  h = (h ^ (h >> 29)) & cap_mask;
  for (;;) {
    const unsigned long long prev = atomicCAS(&keys[h], kEmptyKey, key);
    if (prev == kEmptyKey || prev == key) {
      atomicAdd(&cnt[h], 1u);
      return;
    }
    h = (h + 1) & cap_mask;
  }
}
__global__ void label_argmax_kernel(const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ cnt,
                                    unsigned long long cap, unsigned long long* __restrict__ best) {
  const unsigned long long h = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (h >= cap) return;
  const unsigned long long key = keys[h];
  if (key == kEmptyKey) return;
  const unsigned int vox = static_cast<unsigned int>(key >> 32);
  const unsigned int lab = static_cast<unsigned int>(key);
  // larger count wins; equal counts: smaller label wins (== larger ~label)
  atomicMax(&best[vox], (static_cast<unsigned long long>(cnt[h]) << 32) | (0xFFFFFFFFu - lab));
}
__global__ void label_finish_kernel(const unsigned long long* __restrict__ best, long long n_vox, int* __restrict__ out) {
  const long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (v < n_vox) out[v] = static_cast<int>(0xFFFFFFFFu - static_cast<unsigned int>(best[v]));
}

__global__ void export_kernel(GridView g, long long n_bricks, int* __restrict__ ijk) {
  const long long b = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (b >= n_bricks) return;
  const int bx = static_cast<int>(b % g.bdim[0]) + g.bmin[0];
  const int by = static_cast<int>((b / g.bdim[0]) % g.bdim[1]) + g.bmin[1];
  const int bz = static_cast<int>(b / (static_cast<long long>(g.bdim[0]) * g.bdim[1])) + g.bmin[2];
  long long v = g.base[b];
  for (int z = 0; z < 8; ++z) {
    unsigned long long w = g.mask[b * 8 + z];
    while (w) {
      const int bit = __ffsll(static_cast<long long>(w)) - 1;
      w &= w - 1;
      ijk[3 * v] = bx * 8 + (bit & 7);
      ijk[3 * v + 1] = by * 8 + (bit >> 3);
      ijk[3 * v + 2] = bz * 8 + z;
      ++v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// fused ray march: depth (first run >= 0.1), semantic + instance (first voxel with chord >= 0.01)
// ------------------------------------------------------------------------------------------------
struct Ray {
  float oi[3], di[3], inv[3];
  int step[3];
};

__device__ __forceinline__ float plane_t(const Ray& r, int a, int cell, int size) {
  if (r.step[a] == 0) return INFINITY;
  const float plane = static_cast<float>((cell + (r.step[a] > 0 ? 1 : 0)) * size);
  return __fmul_rn(__fsub_rn(plane, r.oi[a]), r.inv[a]);
}
// arg-min with ties to the lower axis (x < y < z) and the minimum itself; every index is a compile-time constant so
// that the three-element arrays of the traversal stay in registers (a run-time index sends them to local memory)
__device__ __forceinline__ int argmin3(const float* t, float& tmin) {
  int a = 0;
  tmin = t[0];
  if (t[1] < tmin) {
    a = 1;
    tmin = t[1];
  }
  if (t[2] < tmin) {
    a = 2;
    tmin = t[2];
  }
  return a;
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

struct RenderParams {
  GridView g;
  float kinv[9];
  const float* poses;  // [n_cam][16]
  int W, H;
  float* depth;
  int* sem;
  int* inst;
  int bg_sem, bg_inst;
};

// kStage (default): an empty brick costs one byte load (the per-brick occupancy array, L1-resident) instead of its
// 64-byte mask, and the masks of the brick a ray is inside live in the thread's column of a shared-memory tile
// ([word z][thread]: consecutive threads -> consecutive banks) instead of 16 registers, so the per-voxel test is one
// LDS.64 instead of an eight-way 64-bit select chain (ncu round 1: ALU pipe 77 %, the chain was a third of the
// voxel loop).  kStage = false is the round-1 kernel, kept for A/B (ICB_RASTER_STAGE=0).
template <bool kStage>
__global__ void __launch_bounds__(256, 3)
raymarch_kernel(const RenderParams p) {
  __shared__ unsigned long long sm_mask[kStage ? 8 : 1][256];
  const int u = blockIdx.x * 32 + (threadIdx.x & 31);
  const int v = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int cam = blockIdx.z;
  __shared__ float T[16];
  if (threadIdx.x < 16) T[threadIdx.x] = p.poses[cam * 16 + threadIdx.x];
  __syncthreads();
  if (u >= p.W || v >= p.H) return;
  const GridView& g = p.g;

  // ---- ray setup (camera/pinhole.py:110-138, camera/base.py:207-226) ----
  const float fu = static_cast<float>(u), fv = static_cast<float>(v);
  float rc[3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
    rc[a] = __fadd_rn(__fadd_rn(__fmul_rn(p.kinv[3 * a], fu), __fmul_rn(p.kinv[3 * a + 1], fv)), p.kinv[3 * a + 2]);
  const float nrm =
      __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(rc[0], rc[0]), __fmul_rn(rc[1], rc[1])), __fmul_rn(rc[2], rc[2])));
#pragma unroll
  for (int a = 0; a < 3; ++a) rc[a] = __fdiv_rn(rc[a], nrm);
  Ray r;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float d = __fadd_rn(__fadd_rn(__fmul_rn(T[4 * a], rc[0]), __fmul_rn(T[4 * a + 1], rc[1])),
                              __fmul_rn(T[4 * a + 2], rc[2]));
    r.oi[a] = __fadd_rn(__fdiv_rn(__fsub_rn(T[4 * a + 3], g.org[a]), g.vs[a]), 0.5f);
    r.di[a] = __fdiv_rn(d, g.vs[a]);
    r.step[a] = r.di[a] > 0.f ? 1 : (r.di[a] < 0.f ? -1 : 0);
    r.inv[a] = r.step[a] ? __fdiv_rn(1.0f, r.di[a]) : 0.f;
  }

  bool sem_done = false, dep_done = false, run_open = false;
  float run_t0 = 0.f, run_t1 = 0.f, depth_t = 0.f;
  int hit_vox = -1;

  // ---- clip against the brick-aligned bounding box ----
  float tnear = 0.f, tfar = INFINITY;
  bool miss = false;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float lo = static_cast<float>(g.bmin[a] * 8);
    const float hi = static_cast<float>((g.bmin[a] + g.bdim[a]) * 8);
    if (r.step[a] == 0) {
      if (r.oi[a] < lo || r.oi[a] >= hi) miss = true;
    } else {
      const float t1 = __fmul_rn(__fsub_rn(lo, r.oi[a]), r.inv[a]);
      const float t2 = __fmul_rn(__fsub_rn(hi, r.oi[a]), r.inv[a]);
      tnear = fmaxf(tnear, fminf(t1, t2));
      tfar = fminf(tfar, fmaxf(t1, t2));
    }
  }
  if (!miss && tnear < tfar) {
    int b[3];
    float tx[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float pos = __fadd_rn(r.oi[a], __fmul_rn(tnear, r.di[a]));
      b[a] = clampi(static_cast<int>(floorf(__fmul_rn(pos, 0.125f))), g.bmin[a], g.bmin[a] + g.bdim[a] - 1);
      tx[a] = plane_t(r, a, b[a], 8);
    }
    float t = tnear;
    for (;;) {
      float tx_min;
      const int ax = argmin3(tx, tx_min);
      float t_out = fminf(tx_min, tfar);
      if (t_out < t) t_out = t;
      if (t < t_out) {
        const long long bl = brick_lin(g.bmin, g.bdim, b[0], b[1], b[2]);
        const ulonglong2* m2 = reinterpret_cast<const ulonglong2*>(g.mask + bl * 8);
        unsigned long long m[kStage ? 1 : 8];
        bool nonempty;
        if constexpr (kStage) {
          nonempty = __ldg(g.occ + bl) != 0;
          if (nonempty) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const ulonglong2 w = __ldg(m2 + q);
              sm_mask[2 * q][threadIdx.x] = w.x;
              sm_mask[2 * q + 1][threadIdx.x] = w.y;
            }
          }
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const ulonglong2 w = __ldg(m2 + q);
            m[2 * q] = w.x;
            m[2 * q + 1] = w.y;
          }
          nonempty = (m[0] | m[1] | m[2] | m[3] | m[4] | m[5] | m[6] | m[7]) != 0ull;
        }
        if (!nonempty) {
          if (run_open) {  // a gap closes the current run
            if (!dep_done && __fsub_rn(run_t1, run_t0) >= 0.1f) {
              dep_done = true;
              depth_t = run_t0;
            }
            run_open = false;
          }
        } else {
          int c[3];
          float vx[3];
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            const float pos = __fadd_rn(r.oi[a], __fmul_rn(t, r.di[a]));
            c[a] = clampi(static_cast<int>(floorf(pos)), b[a] * 8, b[a] * 8 + 7);
            vx[a] = plane_t(r, a, c[a], 1);
          }
          float tc = t;
          for (;;) {
            float vx_min;
            const int va = argmin3(vx, vx_min);
            float t1 = fminf(vx_min, t_out);
            if (t1 < tc) t1 = tc;
            if (tc < t1) {
              const int lz = c[2] & 7, bit = (c[1] & 7) * 8 + (c[0] & 7);
              unsigned long long wz;
              if constexpr (kStage) {
                wz = sm_mask[lz][threadIdx.x];
              } else {  // select word lz without dynamic register indexing
                wz = m[0];
#pragma unroll
                for (int z = 1; z < 8; ++z) wz = (lz == z) ? m[z] : wz;
              }
              if ((wz >> bit) & 1ull) {
                if (!sem_done && __fsub_rn(t1, tc) >= 0.01f) {
                  int rk = 0;
                  if constexpr (kStage) {
                    for (int z = 0; z < lz; ++z) rk += __popcll(sm_mask[z][threadIdx.x]);
                  } else {
#pragma unroll
                    for (int z = 0; z < 8; ++z) rk += (z < lz) ? __popcll(m[z]) : 0;
                  }
                  rk += __popcll(wz & ((1ull << bit) - 1ull));
                  hit_vox = __ldg(g.base + bl) + rk;
                  sem_done = true;
                }
                if (!dep_done) {
                  if (!run_open) {
                    run_open = true;
                    run_t0 = tc;
                  }
                  run_t1 = t1;
                }
              } else if (run_open) {
                if (!dep_done && __fsub_rn(run_t1, run_t0) >= 0.1f) {
                  dep_done = true;
                  depth_t = run_t0;
                }
                run_open = false;
              }
            }
            if (sem_done && dep_done) break;
            tc = t1;
            if (!(tc < t_out)) break;
            bool left_brick = false;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
              if (a == va) {
                c[a] += r.step[a];
                left_brick = c[a] < b[a] * 8 || c[a] > b[a] * 8 + 7;
                if (!left_brick) vx[a] = plane_t(r, a, c[a], 1);
              }
            }
            if (left_brick) break;
          }
          if (sem_done && dep_done) break;
        }
      }
      t = t_out;
      if (!(t < tfar)) break;
      bool left_grid = false;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        if (a == ax) {
          b[a] += r.step[a];
          left_grid = b[a] < g.bmin[a] || b[a] >= g.bmin[a] + g.bdim[a];
          if (!left_grid) tx[a] = plane_t(r, a, b[a], 8);
        }
      }
      if (left_grid) break;
    }
    if (run_open && !dep_done && __fsub_rn(run_t1, run_t0) >= 0.1f) {
      dep_done = true;
      depth_t = run_t0;
    }
  }

  const size_t o = (static_cast<size_t>(cam) * p.H + v) * p.W + u;
  p.depth[o] = dep_done ? __fmul_rn(depth_t, rc[2]) : 0.f;  // zdepth = t0 * r_cam.z; miss = 0
  p.sem[o] = sem_done ? __ldg(g.sem + hit_vox) : p.bg_sem;  // background_semantic (default 0)
  p.inst[o] = sem_done ? __ldg(g.inst + hit_vox) : p.bg_inst;
}

// ------------------------------------------------------------------------------------------------
// guidance images
// ------------------------------------------------------------------------------------------------
// semantic palette LUT + instance colour overlay -> uint8 RGB (bit-exact integer path)
__global__ void semantic_rgb_kernel(const int* __restrict__ sem, const unsigned char* __restrict__ base_rgb,
                                    const int* __restrict__ inst, long long n,
                                    const unsigned char* __restrict__ palette, int n_classes,
                                    const int* __restrict__ inst_ids, const unsigned char* __restrict__ inst_colors,
                                    int n_ids, unsigned char* __restrict__ rgb) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned char* c;
  if (base_rgb) {
    c = base_rgb + 3 * i;
  } else {
    int s = sem[i];
    s = s < 0 ? 0 : (s >= n_classes ? 0 : s);
    c = palette + 3 * s;
  }
  const int id = inst ? inst[i] : 0;
  if (id > 0) {
    // ids are sorted ascending: binary search
    int lo = 0, hi = n_ids - 1;
    while (lo <= hi) {
      const int mid = (lo + hi) >> 1;
      const int v = __ldg(inst_ids + mid);
      if (v == id) {
        c = inst_colors + 3 * mid;
        break;
      }
      if (v < id)
        lo = mid + 1;
      else
        hi = mid - 1;
    }
  }
  const unsigned char c0 = c[0], c1 = c[1], c2 = c[2];
  rgb[3 * i] = c0;
  rgb[3 * i + 1] = c1;
  rgb[3 * i + 2] = c2;
}

// float LUT gather: out[i, :] = lut[idx[i], :]  (semantic_to_color's float palette lookup)
__global__ void lut_gather_kernel(const int* __restrict__ idx, long long n, const float* __restrict__ lut, int n_rows,
                                  float* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int s = idx[i];
  s = s < 0 ? 0 : (s >= n_rows ? 0 : s);
  out[3 * i] = __ldg(lut + 3 * s);
  out[3 * i + 1] = __ldg(lut + 3 * s + 1);
  out[3 * i + 2] = __ldg(lut + 3 * s + 2);
}

// get_instance_id_for_fvdb_scene_points (utils/fvdb_utils.py:299-385): a car-class point inside an (enlarged)
// oriented box takes that box's object_id_int; boxes are visited in dict order and the LAST match wins.
// One thread per point, the box table (3x4 world->object rows, half extents, id = 16 floats) staged in shared memory.
__global__ void instance_from_boxes_kernel(const float* __restrict__ pts, long long n, const int* __restrict__ sem,
                                           const float* __restrict__ boxes /*[nb][16]*/, int nb, unsigned car_mask,
                                           int* __restrict__ out) {
  extern __shared__ float sb[];
  for (int i = threadIdx.x; i < nb * 16; i += blockDim.x) sb[i] = boxes[i];
  __syncthreads();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = sem[i];
  int id = 0;
  if (s >= 0 && s < 32 && ((car_mask >> s) & 1u)) {
    const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    for (int b = 0; b < nb; ++b) {
      const float* m = sb + b * 16;
      const float lx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[0], x), __fmul_rn(m[1], y)), __fmul_rn(m[2], z)), m[3]);
      const float ly = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[4], x), __fmul_rn(m[5], y)), __fmul_rn(m[6], z)), m[7]);
      const float lz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[8], x), __fmul_rn(m[9], y)), __fmul_rn(m[10], z)), m[11]);
      if (fabsf(lx) <= m[12] && fabsf(ly) <= m[13] && fabsf(lz) <= m[14]) id = __float_as_int(m[15]);
    }
  }
  out[i] = id;
}

// X_cam0 = T_{0<-i} [depth * K^-1 (u,v,1); 1]  (utils/depth_utils.py:448-464); misses -> 1e7
__global__ void unproject_kernel(const float* __restrict__ depth, const float* __restrict__ c2c0 /*[n][16]*/,
                                 const float* __restrict__ kinv9, int n_cam, int H, int W, float* __restrict__ xyz) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long n = static_cast<long long>(n_cam) * H * W;
  if (i >= n) return;
  const int u = static_cast<int>(i % W);
  const int v = static_cast<int>((i / W) % H);
  const int cam = static_cast<int>(i / (static_cast<long long>(W) * H));
  const float d = depth[i];
  if (d == 0.f) {
    xyz[3 * i] = xyz[3 * i + 1] = xyz[3 * i + 2] = 1e7f;
    return;
  }
  const float fu = static_cast<float>(u), fv = static_cast<float>(v);
  float cp[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float un = __fadd_rn(__fadd_rn(__fmul_rn(kinv9[3 * a], fu), __fmul_rn(kinv9[3 * a + 1], fv)), kinv9[3 * a + 2]);
    cp[a] = __fmul_rn(d, un);
  }
  const float* T = c2c0 + cam * 16;
#pragma unroll
  for (int a = 0; a < 3; ++a)
    xyz[3 * i + a] = __fadd_rn(
        __fadd_rn(__fadd_rn(__fmul_rn(T[4 * a], cp[0]), __fmul_rn(T[4 * a + 1], cp[1])), __fmul_rn(T[4 * a + 2], cp[2])),
        T[4 * a + 3]);
}

// ((X - min)/range*2 - 1).clamp(-1,1) -> (.+1)/2 -> *255 -> uint8 (truncation); misses -> 255
// (utils/buffer_utils.py:251-262, guidance_buffer_generation.py:710)
__global__ void coord_normalize_kernel(const float* __restrict__ xyz, const float* __restrict__ depth, long long n,
                                       const float* __restrict__ mins, const float* __restrict__ ranges,
                                       float* __restrict__ out_f32, unsigned char* __restrict__ out_u8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool far = depth[i] == 0.f;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float c;
    if (far) {
      c = 1.0f;
    } else {
      float q = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(xyz[3 * i + a], mins[a]), ranges[a]), 2.0f), 1.0f);
      q = fminf(fmaxf(q, -1.0f), 1.0f);
      c = __fdiv_rn(__fadd_rn(q, 1.0f), 2.0f);
    }
    if (out_f32) out_f32[3 * i + a] = c;
    if (out_u8) out_u8[3 * i + a] = static_cast<unsigned char>(__fmul_rn(c, 255.0f));
  }
}

// ------------------------------------------------------------------------------------------------
// mesh -> voxels (fvdb.gridbatch_from_mesh as used for the CAD cars, utils/fvdb_utils.py:279-287): a voxel is
// active iff its cube [centre - vs/2, centre + vs/2] overlaps a triangle (separating-axis test, fp64, touching
// counts).  One thread per triangle marks bits in a dense mask over the mesh bounding box.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool axis_separates(double a0, double a1, double a2, double rad) {
  const double mn = fmin(a0, fmin(a1, a2)), mx = fmax(a0, fmax(a1, a2));
  return mn > rad || mx < -rad;
}

__device__ bool tri_box_overlap(const double* c, double h, const double* p0, const double* p1, const double* p2) {
  double v0[3], v1[3], v2[3];
  for (int a = 0; a < 3; ++a) {
    v0[a] = p0[a] - c[a];
    v1[a] = p1[a] - c[a];
    v2[a] = p2[a] - c[a];
  }
  const double e[3][3] = {{v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2]},
                          {v2[0] - v1[0], v2[1] - v1[1], v2[2] - v1[2]},
                          {v0[0] - v2[0], v0[1] - v2[1], v0[2] - v2[2]}};
  // 9 cross-product axes  (unit box axis u_i) x e_j
  for (int j = 0; j < 3; ++j) {
    const double ex = e[j][0], ey = e[j][1], ez = e[j][2];
    const double fx = fabs(ex), fy = fabs(ey), fz = fabs(ez);
    // axis (0, -ez, ey)
    if (axis_separates(-ez * v0[1] + ey * v0[2], -ez * v1[1] + ey * v1[2], -ez * v2[1] + ey * v2[2], (fz + fy) * h)) return false;
    // axis (ez, 0, -ex)
    if (axis_separates(ez * v0[0] - ex * v0[2], ez * v1[0] - ex * v1[2], ez * v2[0] - ex * v2[2], (fz + fx) * h)) return false;
    // axis (-ey, ex, 0)
    if (axis_separates(-ey * v0[0] + ex * v0[1], -ey * v1[0] + ex * v1[1], -ey * v2[0] + ex * v2[1], (fy + fx) * h)) return false;
  }
  // 3 box axes
  for (int a = 0; a < 3; ++a)
    if (axis_separates(v0[a], v1[a], v2[a], h)) return false;
  // triangle plane
  const double n[3] = {e[0][1] * e[1][2] - e[0][2] * e[1][1], e[0][2] * e[1][0] - e[0][0] * e[1][2],
                       e[0][0] * e[1][1] - e[0][1] * e[1][0]};
  const double d = n[0] * v0[0] + n[1] * v0[1] + n[2] * v0[2];
  const double r = h * (fabs(n[0]) + fabs(n[1]) + fabs(n[2]));
  return !(d > r || d < -r);
}

__global__ void mesh_voxelize_kernel(const double* __restrict__ verts, const int* __restrict__ faces, int nf, double vs,
                                     double origin, int3 ijk_min, int3 dims, unsigned int* __restrict__ mask) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  double p[3][3];
  for (int v = 0; v < 3; ++v)
    for (int a = 0; a < 3; ++a) p[v][a] = verts[3 * faces[3 * f + v] + a];
  int lo[3], hi[3];
  for (int a = 0; a < 3; ++a) {
    const double mn = fmin(p[0][a], fmin(p[1][a], p[2][a])), mx = fmax(p[0][a], fmax(p[1][a], p[2][a]));
    lo[a] = static_cast<int>(floor((mn - origin) / vs - 0.5)) - 1;
    hi[a] = static_cast<int>(ceil((mx - origin) / vs + 0.5)) + 1;
  }
  const int mn3[3] = {ijk_min.x, ijk_min.y, ijk_min.z};
  const int dm3[3] = {dims.x, dims.y, dims.z};
  for (int a = 0; a < 3; ++a) {
    lo[a] = max(lo[a], mn3[a]);
    hi[a] = min(hi[a], mn3[a] + dm3[a] - 1);
  }
  for (int k = lo[2]; k <= hi[2]; ++k)
    for (int j = lo[1]; j <= hi[1]; ++j)
      for (int i = lo[0]; i <= hi[0]; ++i) {
        const double c[3] = {origin + i * vs, origin + j * vs, origin + k * vs};
        if (tri_box_overlap(c, 0.5 * vs, p[0], p[1], p[2])) {
          const long long lin = (static_cast<long long>(k - mn3[2]) * dm3[1] + (j - mn3[1])) * dm3[0] + (i - mn3[0]);
          atomicOr(&mask[lin >> 5], 1u << (lin & 31));
        }
      }
}

GridView view_of(const ic_grid* g) {
  GridView v;
  for (int a = 0; a < 3; ++a) {
    v.org[a] = g->org[a];
    v.vs[a] = g->vs[a];
    v.bmin[a] = g->bmin[a];
    v.bdim[a] = g->bdim[a];
  }
  v.mask = g->mask;
  v.base = g->base;
  v.occ = g->occ;
  v.sem = g->sem;
  v.inst = g->inst;
  return v;
}

inline unsigned nblk(long long n, int t) { return static_cast<unsigned>((n + t - 1) / t); }

}  // namespace

extern "C" {

int ic_grid_destroy(ic_grid* g) {
  if (!g) return IC_OK;
  cudaFree(g->mask);
  cudaFree(g->base);
  cudaFree(g->occ);
  cudaFree(g->sem);
  cudaFree(g->inst);
  delete g;
  return IC_OK;
}

int ic_grid_build(const float* points, long long m, const float* vs_host, const float* origin_host, const int* sem,
                  const int* inst, ic_grid** out, void* stream) {
  if (!points || m <= 0 || !vs_host || !origin_host || !out) return IC_ERR_INVALID;
  int r = ic_device_check();
  if (r != IC_OK) return r;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ic_grid* g = new ic_grid();
  for (int a = 0; a < 3; ++a) {
    g->vs[a] = vs_host[a];
    g->org[a] = origin_host[a];
  }
  int* ijk = nullptr;
  int* bbox = nullptr;
  int* block_sums = nullptr;
  long long* total = nullptr;
  unsigned long long *keys = nullptr, *best = nullptr;
  unsigned int* cnt = nullptr;
  auto cleanup = [&]() {
    cudaFree(ijk);
    cudaFree(bbox);
    cudaFree(block_sums);
    cudaFree(total);
    cudaFree(keys);
    cudaFree(best);
    cudaFree(cnt);
  };
#define RB_CHECK(expr)                  \
  do {                                  \
    if ((expr) != cudaSuccess) {        \
      fprintf(stderr, "[icb] CUDA error in ic_grid_build: %s\n", cudaGetErrorString(cudaGetLastError())); \
      cleanup();                        \
      ic_grid_destroy(g);               \
      return IC_ERR_CUDA;               \
    }                                   \
  } while (0)
  RB_CHECK(cudaMalloc(&ijk, sizeof(int) * 3 * m));
  RB_CHECK(cudaMalloc(&bbox, sizeof(int) * 6));
  const int init[6] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN};
  RB_CHECK(cudaMemcpyAsync(bbox, init, sizeof(init), cudaMemcpyHostToDevice, st));
  ijk_bbox_kernel<<<nblk(m, 256), 256, 0, st>>>(points, m, make_float3(g->org[0], g->org[1], g->org[2]),
                                                make_float3(g->vs[0], g->vs[1], g->vs[2]), ijk, bbox);
  int hb[6];
  RB_CHECK(cudaMemcpyAsync(hb, bbox, sizeof(hb), cudaMemcpyDeviceToHost, st));
  RB_CHECK(cudaStreamSynchronize(st));
  g->n_bricks = 1;
  for (int a = 0; a < 3; ++a) {
    g->imin[a] = hb[a];
    g->imax[a] = hb[3 + a];
    g->bmin[a] = hb[a] >> 3;
    g->bdim[a] = (hb[3 + a] >> 3) - g->bmin[a] + 1;
    g->n_bricks *= g->bdim[a];
  }
  if (g->n_bricks > (1ll << 31)) {
    cleanup();
    ic_grid_destroy(g);
    return IC_ERR_UNSUPPORTED;
  }
  RB_CHECK(cudaMalloc(&g->mask, sizeof(unsigned long long) * 8 * g->n_bricks));
  RB_CHECK(cudaMalloc(&g->base, sizeof(int) * g->n_bricks));
  RB_CHECK(cudaMalloc(&g->occ, g->n_bricks));
  RB_CHECK(cudaMemsetAsync(g->mask, 0, sizeof(unsigned long long) * 8 * g->n_bricks, st));
  BrickGeom bg;
  for (int a = 0; a < 3; ++a) {
    bg.bmin[a] = g->bmin[a];
    bg.bdim[a] = g->bdim[a];
  }
  set_bits_kernel<<<nblk(m, 256), 256, 0, st>>>(ijk, m, bg, g->mask);
  const int n_scan_blocks = static_cast<int>((g->n_bricks + 1023) / 1024);
  RB_CHECK(cudaMalloc(&block_sums, sizeof(int) * n_scan_blocks));
  RB_CHECK(cudaMalloc(&total, sizeof(long long)));
  brick_count_kernel<<<n_scan_blocks, 1024, 0, st>>>(g->mask, g->n_bricks, g->base, block_sums, g->occ);
  scan_block_sums_kernel<<<1, 1024, 0, st>>>(block_sums, n_scan_blocks, total);
  add_block_offsets_kernel<<<n_scan_blocks, 1024, 0, st>>>(g->base, g->n_bricks, block_sums);
  long long n_vox = 0;
  RB_CHECK(cudaMemcpyAsync(&n_vox, total, sizeof(long long), cudaMemcpyDeviceToHost, st));
  RB_CHECK(cudaStreamSynchronize(st));
  g->n_vox = n_vox;
  RB_CHECK(cudaMalloc(&g->sem, sizeof(int) * n_vox));
  RB_CHECK(cudaMalloc(&g->inst, sizeof(int) * n_vox));
  RB_CHECK(cudaMemsetAsync(g->sem, 0, sizeof(int) * n_vox, st));
  RB_CHECK(cudaMemsetAsync(g->inst, 0, sizeof(int) * n_vox, st));
  if (sem || inst) {
    unsigned long long cap = 1;
    while (cap < static_cast<unsigned long long>(m) * 2) cap <<= 1;
    RB_CHECK(cudaMalloc(&keys, sizeof(unsigned long long) * cap));
    RB_CHECK(cudaMalloc(&cnt, sizeof(unsigned int) * cap));
    RB_CHECK(cudaMalloc(&best, sizeof(unsigned long long) * n_vox));
    GridView gv = view_of(g);
    for (int pass = 0; pass < 2; ++pass) {
      const int* lab = pass == 0 ? sem : inst;
      if (!lab) continue;
      RB_CHECK(cudaMemsetAsync(keys, 0xFF, sizeof(unsigned long long) * cap, st));
      RB_CHECK(cudaMemsetAsync(cnt, 0, sizeof(unsigned int) * cap, st));
      RB_CHECK(cudaMemsetAsync(best, 0, sizeof(unsigned long long) * n_vox, st));
      label_count_kernel<<<nblk(m, 256), 256, 0, st>>>(ijk, lab, m, gv, keys, cnt, cap - 1);
      label_argmax_kernel<<<nblk(static_cast<long long>(cap), 256), 256, 0, st>>>(keys, cnt, cap, best);
      label_finish_kernel<<<nblk(n_vox, 256), 256, 0, st>>>(best, n_vox, pass == 0 ? g->sem : g->inst);
    }
  }
  RB_CHECK(cudaStreamSynchronize(st));
  RB_CHECK(cudaGetLastError());
  cleanup();
#undef RB_CHECK
  *out = g;
  return IC_OK;
}

long long ic_grid_num_voxels(const ic_grid* g) { return g ? g->n_vox : 0; }
long long ic_grid_num_bricks(const ic_grid* g) { return g ? g->n_bricks : 0; }

int ic_grid_info(const ic_grid* g, int* imin, int* imax, int* bmin, int* bdim) {
  if (!g) return IC_ERR_INVALID;
  for (int a = 0; a < 3; ++a) {
    imin[a] = g->imin[a];
    imax[a] = g->imax[a];
    bmin[a] = g->bmin[a];
    bdim[a] = g->bdim[a];
  }
  return IC_OK;
}

int ic_grid_export(const ic_grid* g, int* ijk, int* sem, int* inst, void* stream) {
  if (!g || !ijk) return IC_ERR_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  export_kernel<<<nblk(g->n_bricks, 128), 128, 0, st>>>(view_of(g), g->n_bricks, ijk);
  if (sem) ICB_CUDA_CHECK(cudaMemcpyAsync(sem, g->sem, sizeof(int) * g->n_vox, cudaMemcpyDeviceToDevice, st));
  if (inst) ICB_CUDA_CHECK(cudaMemcpyAsync(inst, g->inst, sizeof(int) * g->n_vox, cudaMemcpyDeviceToDevice, st));
  ICB_CUDA_CHECK(cudaGetLastError());
  return IC_OK;
}

int ic_raster_render(const ic_grid* g, const float* kinv_host, const float* poses, int n_cam, int W, int H,
                     const int* attr0, const int* attr1, int background0, int background1, float* depth, int* sem,
                     int* inst, void* stream) {
  if (!g || !kinv_host || !poses || n_cam <= 0 || W <= 0 || H <= 0 || !depth || !sem || !inst) return IC_ERR_INVALID;
  RenderParams p;
  p.g = view_of(g);
  if (attr0) p.g.sem = attr0;
  if (attr1) p.g.inst = attr1;
  p.bg_sem = background0;
  p.bg_inst = background1;
  for (int i = 0; i < 9; ++i) p.kinv[i] = kinv_host[i];
  p.poses = poses;
  p.W = W;
  p.H = H;
  p.depth = depth;
  p.sem = sem;
  p.inst = inst;
  dim3 grid((W + 31) / 32, (H + 7) / 8, n_cam);
  static int stage = -1;
  if (stage < 0) {
    const char* e = getenv("ICB_RASTER_STAGE");
    stage = e ? atoi(e) : 1;
  }
  if (stage)
    raymarch_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  else
    raymarch_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  ICB_CUDA_CHECK(cudaGetLastError());
  return IC_OK;
}

int ic_semantic_rgb(const int* sem, const unsigned char* base_rgb, const int* inst, long long n,
                    const unsigned char* palette, int n_classes, const int* inst_ids_sorted,
                    const unsigned char* inst_colors, int n_ids, unsigned char* rgb, void* stream) {
  if ((!sem && !base_rgb) || (sem && !palette) || !rgb || n <= 0) return IC_ERR_INVALID;
  if (inst && n_ids > 0 && (!inst_ids_sorted || !inst_colors)) return IC_ERR_INVALID;
  semantic_rgb_kernel<<<nblk(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      sem, base_rgb, inst, n, palette, n_classes, inst_ids_sorted, inst_colors, n_ids, rgb);
  ICB_CUDA_CHECK(cudaGetLastError());
  return IC_OK;
}

int ic_lut_gather_f32(const int* idx, long long n, const float* lut, int n_rows, float* out, void* stream) {
  if (!idx || !lut || !out || n <= 0 || n_rows <= 0) return IC_ERR_INVALID;
  lut_gather_kernel<<<nblk(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(idx, n, lut, n_rows, out);
  ICB_CUDA_CHECK(cudaGetLastError());
  return IC_OK;
}

int ic_instance_from_boxes(const float* points, long long n, const int* sem, const float* boxes, int n_boxes,
                           unsigned int car_class_mask, int* instance_id, void* stream) {
  if (!points || !sem || !instance_id || n <= 0 || n_boxes < 0 || (n_boxes > 0 && !boxes)) return IC_ERR_INVALID;
  if (n_boxes * 64 > 200 * 1024) return IC_ERR_UNSUPPORTED;  // box table must fit in shared memory (3200 boxes)
  const size_t smem = static_cast<size_t>(n_boxes) * 64;
  if (smem > 48 * 1024)
    ICB_CUDA_CHECK(cudaFuncSetAttribute(instance_from_boxes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
  instance_from_boxes_kernel<<<nblk(n, 256), 256, smem, static_cast<cudaStream_t>(stream)>>>(points, n, sem, boxes, n_boxes,
                                                                                           car_class_mask, instance_id);
  ICB_CUDA_CHECK(cudaGetLastError());
  return IC_OK;
}

int ic_coord_unproject(const float* depth, const float* cam_to_cam0, const float* kinv9, int n_cam, int H, int W,
                       float* xyz, void* stream) {
  if (!depth || !cam_to_cam0 || !kinv9 || !xyz) return IC_ERR_INVALID;
  const long long n = static_cast<long long>(n_cam) * H * W;
  unproject_kernel<<<nblk(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(depth, cam_to_cam0, kinv9, n_cam, H, W, xyz);
  ICB_CUDA_CHECK(cudaGetLastError());
  return IC_OK;
}

int ic_coord_normalize(const float* xyz, const float* depth, long long n_pixels, const float* mins3, const float* ranges3,
                       float* out_f32, unsigned char* out_u8, void* stream) {
  if (!xyz || !depth || !mins3 || !ranges3 || (!out_f32 && !out_u8)) return IC_ERR_INVALID;
  coord_normalize_kernel<<<nblk(n_pixels, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(xyz, depth, n_pixels, mins3,
                                                                                            ranges3, out_f32, out_u8);
  ICB_CUDA_CHECK(cudaGetLastError());
  return IC_OK;
}

int ic_mesh_voxelize_mask(const double* verts, int nv, const int* faces, int nf, double voxel_size, double origin,
                          const int* ijk_min_host3, const int* dims_host3, unsigned int* mask, void* stream) {
  if (!verts || !faces || nv <= 0 || nf <= 0 || !ijk_min_host3 || !dims_host3 || !mask || voxel_size <= 0) return IC_ERR_INVALID;
  int r = ic_device_check();
  if (r != IC_OK) return r;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long nbits = static_cast<long long>(dims_host3[0]) * dims_host3[1] * dims_host3[2];
  ICB_CUDA_CHECK(cudaMemsetAsync(mask, 0, static_cast<size_t>((nbits + 31) / 32) * 4, st));
  mesh_voxelize_kernel<<<nblk(nf, 128), 128, 0, st>>>(verts, faces, nf, voxel_size, origin,
                                                      make_int3(ijk_min_host3[0], ijk_min_host3[1], ijk_min_host3[2]),
                                                      make_int3(dims_host3[0], dims_host3[1], dims_host3[2]), mask);
  ICB_CUDA_CHECK(cudaGetLastError());
  return IC_OK;
}

}  // extern "C"
