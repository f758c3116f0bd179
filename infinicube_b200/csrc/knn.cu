// Exact nearest-neighbour search over a uniform cell grid: the label transfer of stage-1's chunk merge
// (SURVEY.md §8f N4).  Replaces infinicube/voxelgen/ext/common/knn.cu:15-50 (knn_query_fast, a FLANN-style
// KD-tree: kdtree_cuda.cu:1070-1094) for k = 1 as used by semantic_from_points
// (infinicube/voxelgen/utils/color_util.py:52-60).
//
// B200 shape of the problem: millions of voxel centres against millions of voxel centres, bounded extent, near
// uniform density on surfaces.  A KD-tree walk is pointer chasing with a per-thread stack; a counting-sorted cell
// grid turns the search into a few contiguous row scans per query (cells are x-fastest, so one (y, z) row of a
// search shell is ONE contiguous range of the sorted point array), all L2-resident (16 B per point).
//   build : bbox (ordered-int atomics) -> per-cell counts -> 3-phase exclusive scan -> scatter (x, y, z, index)
//   query : one thread per query; shells of growing Chebyshev radius around the query's cell, stop as soon as
//           the best squared distance is below the squared distance to the nearest unscanned cell face
// Arithmetic spec (the CPU checker restates it independently): fp32, no FMA contraction,
//   d2 = ((qx-px)^2 + (qy-py)^2) + (qz-pz)^2 ; arg-min with ties -> smallest reference index.
#include <float.h>
#include <limits.h>
#include <math.h>

#include "host_util.h"

struct ic_knn {
  long long m = 0;
  float org[3] = {0, 0, 0};
  float cell = 1.f;
  float eps = 0.f;
  int dim[3] = {1, 1, 1};
  long long n_cells = 1;
  int* cell_start = nullptr;  // [n_cells + 1] exclusive prefix of the per-cell counts
  float4* pts = nullptr;      // [m] sorted by cell: (x, y, z, bit-cast reference index)
};

namespace icb {
namespace {

struct KnnGeom {
  float ox, oy, oz, cell, inv_cell;
  int dx, dy, dz;
  float eps;  // slack of the pruning bound: cell assignment rounds in fp32, so faces are only known to ~8 ulp
};

__device__ __forceinline__ unsigned f2ord(float f) {
  const unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
inline float ord2f(unsigned u) {
  const unsigned b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  float f;
  memcpy(&f, &b, 4);
  return f;
}

__global__ void knn_bbox_kernel(const float* __restrict__ ref, long long m, int stride, unsigned* __restrict__ bbox) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  unsigned lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
  if (i < m) {
#pragma unroll
    for (int a = 0; a < 3; ++a) lo[a] = hi[a] = f2ord(ref[i * stride + a]);
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    lo[a] = __reduce_min_sync(0xffffffffu, lo[a]);
    hi[a] = __reduce_max_sync(0xffffffffu, hi[a]);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      atomicMin(&bbox[a], lo[a]);
      atomicMax(&bbox[3 + a], hi[a]);
    }
  }
}

__device__ __forceinline__ int cell_coord(float p, float o, float inv_cell, int dim) {
  const int c = static_cast<int>(floorf(__fmul_rn(__fsub_rn(p, o), inv_cell)));
  return c < 0 ? 0 : (c >= dim ? dim - 1 : c);
}
__device__ __forceinline__ long long cell_lin(const KnnGeom& g, int x, int y, int z) {
  return (static_cast<long long>(z) * g.dy + y) * g.dx + x;
}

__global__ void knn_count_kernel(const float* __restrict__ ref, long long m, int stride, KnnGeom g,
                                 int* __restrict__ counts) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int x = cell_coord(ref[i * stride], g.ox, g.inv_cell, g.dx);
  const int y = cell_coord(ref[i * stride + 1], g.oy, g.inv_cell, g.dy);
  const int z = cell_coord(ref[i * stride + 2], g.oz, g.inv_cell, g.dz);
  atomicAdd(&counts[cell_lin(g, x, y, z)], 1);
}

// three-phase exclusive scan, 1024 elements per block (in place)
__global__ void knn_scan_local_kernel(int* __restrict__ a, long long n, int* __restrict__ block_sums) {
  __shared__ int sh[1024];
  const long long i = static_cast<long long>(blockIdx.x) * 1024 + threadIdx.x;
  const int c = i < n ? a[i] : 0;
  sh[threadIdx.x] = c;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    const int v = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
    __syncthreads();
    sh[threadIdx.x] += v;
    __syncthreads();
  }
  if (i < n) a[i] = sh[threadIdx.x] - c;
  if (threadIdx.x == 1023) block_sums[blockIdx.x] = sh[1023];
}
__global__ void knn_scan_sums_kernel(int* __restrict__ block_sums, int n_blocks) {
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
      const int v = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
      __syncthreads();
      sh[threadIdx.x] += v;
      __syncthreads();
    }
    if (i < n_blocks) block_sums[i] = carry + sh[threadIdx.x] - c;
    __syncthreads();
    if (threadIdx.x == 0) carry += sh[1023];
    __syncthreads();
  }
}
__global__ void knn_scan_add_kernel(int* __restrict__ a, long long n, const int* __restrict__ block_sums) {
  const long long i = static_cast<long long>(blockIdx.x) * 1024 + threadIdx.x;
  if (i < n) a[i] += block_sums[blockIdx.x];
}

__global__ void knn_fill_kernel(const float* __restrict__ ref, long long m, int stride, KnnGeom g,
                                const int* __restrict__ cell_start, int* __restrict__ cursor, float4* __restrict__ pts) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const float px = ref[i * stride], py = ref[i * stride + 1], pz = ref[i * stride + 2];
  const long long c = cell_lin(g, cell_coord(px, g.ox, g.inv_cell, g.dx), cell_coord(py, g.oy, g.inv_cell, g.dy),
                               cell_coord(pz, g.oz, g.inv_cell, g.dz));
  const int pos = cell_start[c] + atomicAdd(&cursor[c], 1);
  pts[pos] = make_float4(px, py, pz, __int_as_float(static_cast<int>(i)));
}

__device__ __forceinline__ void scan_range(const float4* __restrict__ pts, int a, int b, float qx, float qy, float qz,
                                           float& best, int& bi) {
  for (int t = a; t < b; ++t) {
    const float4 p = __ldg(pts + t);
    const float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
    const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    const int id = __float_as_int(p.w);
    if (d < best || (d == best && id < bi)) {
      best = d;
      bi = id;
    }
  }
}

__global__ void __launch_bounds__(128)
knn_query1_kernel(const float4* __restrict__ pts, const int* __restrict__ cell_start, KnnGeom g,
                  const float* __restrict__ q, long long n, int stride, const long long* __restrict__ labels,
                  int* __restrict__ out_idx, float* __restrict__ out_d2, long long* __restrict__ out_labels) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float qx = q[i * stride], qy = q[i * stride + 1], qz = q[i * stride + 2];
  const int cx = cell_coord(qx, g.ox, g.inv_cell, g.dx);
  const int cy = cell_coord(qy, g.oy, g.inv_cell, g.dy);
  const int cz = cell_coord(qz, g.oz, g.inv_cell, g.dz);
  float best = INFINITY;
  int bi = INT_MAX;
  const int rmax = max(g.dx, max(g.dy, g.dz));  // hard bound: a NaN query must not spin (its comparisons are all false)
  for (int r = 0; r <= rmax; ++r) {
    const int x0 = max(cx - r, 0), x1 = min(cx + r, g.dx - 1);
    const int y0 = max(cy - r, 0), y1 = min(cy + r, g.dy - 1);
    const int z0 = max(cz - r, 0), z1 = min(cz + r, g.dz - 1);
    for (int z = z0; z <= z1; ++z) {
      const bool zface = (z - cz == r) || (cz - z == r);
      for (int y = y0; y <= y1; ++y) {
        const long long row = cell_lin(g, 0, y, z);
        if (zface || (y - cy == r) || (cy - y == r)) {  // whole row of the shell: one contiguous point range
          scan_range(pts, cell_start[row + x0], cell_start[row + x1 + 1], qx, qy, qz, best, bi);
        } else {                                         // interior row: only the two x-faces of the shell
          if (cx - r >= 0) scan_range(pts, cell_start[row + cx - r], cell_start[row + cx - r + 1], qx, qy, qz, best, bi);
          if (r > 0 && cx + r < g.dx)
            scan_range(pts, cell_start[row + cx + r], cell_start[row + cx + r + 1], qx, qy, qz, best, bi);
        }
      }
    }
    // distance from the query to the nearest face of the scanned block that still has cells behind it
    float lb = INFINITY;
    if (cx - r > 0) lb = fminf(lb, qx - (g.ox + static_cast<float>(cx - r) * g.cell));
    if (cx + r < g.dx - 1) lb = fminf(lb, (g.ox + static_cast<float>(cx + r + 1) * g.cell) - qx);
    if (cy - r > 0) lb = fminf(lb, qy - (g.oy + static_cast<float>(cy - r) * g.cell));
    if (cy + r < g.dy - 1) lb = fminf(lb, (g.oy + static_cast<float>(cy + r + 1) * g.cell) - qy);
    if (cz - r > 0) lb = fminf(lb, qz - (g.oz + static_cast<float>(cz - r) * g.cell));
    if (cz + r < g.dz - 1) lb = fminf(lb, (g.oz + static_cast<float>(cz + r + 1) * g.cell) - qz);
    if (lb == INFINITY) break;  // the whole grid has been scanned
    lb -= g.eps;
    if (lb > 0.f && best < lb * lb) break;
  }
  const bool found = bi != INT_MAX;  // false only for NaN queries
  if (out_idx) out_idx[i] = found ? bi : -1;
  if (out_d2) out_d2[i] = best;
  if (out_labels) out_labels[i] = found ? labels[bi] : -1;
}

// k nearest neighbours (k <= K): the same shell walk with a sorted (distance, index) list in registers; the walk stops
// once the k-th best squared distance is below the squared distance to the nearest unscanned cell face.
template <int K>
__device__ __forceinline__ void knn_insert(float (&bd)[K], int (&bi)[K], float d, int id) {
  if (!(d < bd[K - 1] || (d == bd[K - 1] && id < bi[K - 1]))) return;
  bd[K - 1] = d;
  bi[K - 1] = id;
#pragma unroll
  for (int j = K - 1; j > 0; --j) {  // one bubble pass keeps the list sorted by (distance, index)
    const bool sw = bd[j] < bd[j - 1] || (bd[j] == bd[j - 1] && bi[j] < bi[j - 1]);
    const float td = bd[j], ud = bd[j - 1];
    const int ti = bi[j], ui = bi[j - 1];
    bd[j] = sw ? ud : td;
    bd[j - 1] = sw ? td : ud;
    bi[j] = sw ? ui : ti;
    bi[j - 1] = sw ? ti : ui;
  }
}

template <int K>
__device__ __forceinline__ void scan_range_k(const float4* __restrict__ pts, int s, int e, float qx, float qy, float qz,
                                             float (&bd)[K], int (&bi)[K]) {
  for (int j = s; j < e; ++j) {
    const float4 p = __ldg(pts + j);
    const float dx = __fsub_rn(p.x, qx), dy = __fsub_rn(p.y, qy), dz = __fsub_rn(p.z, qz);
    const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    knn_insert<K>(bd, bi, d, __float_as_int(p.w));
  }
}

template <int K>
__global__ void __launch_bounds__(128)
knn_queryk_kernel(const float4* __restrict__ pts, const int* __restrict__ cell_start, KnnGeom g,
                  const float* __restrict__ q, long long n, int stride, int k, int* __restrict__ out_idx,
                  float* __restrict__ out_d2) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float qx = q[i * stride], qy = q[i * stride + 1], qz = q[i * stride + 2];
  const int cx = cell_coord(qx, g.ox, g.inv_cell, g.dx);
  const int cy = cell_coord(qy, g.oy, g.inv_cell, g.dy);
  const int cz = cell_coord(qz, g.oz, g.inv_cell, g.dz);
  float bd[K];
  int bi[K];
#pragma unroll
  for (int j = 0; j < K; ++j) {
    bd[j] = INFINITY;
    bi[j] = INT_MAX;
  }
  const int rmax = max(g.dx, max(g.dy, g.dz));
  for (int r = 0; r <= rmax; ++r) {
    const int x0 = max(cx - r, 0), x1 = min(cx + r, g.dx - 1);
    const int y0 = max(cy - r, 0), y1 = min(cy + r, g.dy - 1);
    const int z0 = max(cz - r, 0), z1 = min(cz + r, g.dz - 1);
    for (int z = z0; z <= z1; ++z) {
      const bool zface = (z - cz == r) || (cz - z == r);
      for (int y = y0; y <= y1; ++y) {
        const long long row = cell_lin(g, 0, y, z);
        if (zface || (y - cy == r) || (cy - y == r)) {
          scan_range_k<K>(pts, cell_start[row + x0], cell_start[row + x1 + 1], qx, qy, qz, bd, bi);
        } else {
          if (cx - r >= 0) scan_range_k<K>(pts, cell_start[row + cx - r], cell_start[row + cx - r + 1], qx, qy, qz, bd, bi);
          if (r > 0 && cx + r < g.dx)
            scan_range_k<K>(pts, cell_start[row + cx + r], cell_start[row + cx + r + 1], qx, qy, qz, bd, bi);
        }
      }
    }
    float lb = INFINITY;
    if (cx - r > 0) lb = fminf(lb, qx - (g.ox + static_cast<float>(cx - r) * g.cell));
    if (cx + r < g.dx - 1) lb = fminf(lb, (g.ox + static_cast<float>(cx + r + 1) * g.cell) - qx);
    if (cy - r > 0) lb = fminf(lb, qy - (g.oy + static_cast<float>(cy - r) * g.cell));
    if (cy + r < g.dy - 1) lb = fminf(lb, (g.oy + static_cast<float>(cy + r + 1) * g.cell) - qy);
    if (cz - r > 0) lb = fminf(lb, qz - (g.oz + static_cast<float>(cz - r) * g.cell));
    if (cz + r < g.dz - 1) lb = fminf(lb, (g.oz + static_cast<float>(cz + r + 1) * g.cell) - qz);
    if (lb == INFINITY) break;
    lb -= g.eps;
    float kth = bd[0];
#pragma unroll
    for (int j = 1; j < K; ++j) kth = (j == k - 1) ? bd[j] : kth;
    if (k == 1) kth = bd[0];
    if (lb > 0.f && kth < lb * lb) break;
  }
#pragma unroll
  for (int j = 0; j < K; ++j) {
    if (j < k) {
      const bool found = bi[j] != INT_MAX;  // fewer than k reference points (or a NaN query): -1 / inf
      if (out_idx) out_idx[i * k + j] = found ? bi[j] : -1;
      if (out_d2) out_d2[i * k + j] = bd[j];
    }
  }
}

inline unsigned nblk(long long n, int t) { return static_cast<unsigned>((n + t - 1) / t); }

}  // namespace
}  // namespace icb

using namespace icb;

extern "C" {

int ic_knn_destroy(ic_knn* k) {
  if (!k) return IC_OK;
  cudaFree(k->cell_start);
  cudaFree(k->pts);
  delete k;
  return IC_OK;
}

int ic_knn_build(const float* ref_xyz, long long m, int stride, float cell_size, ic_knn** out, void* stream) {
  if (!ref_xyz || m <= 0 || m > INT_MAX || stride < 3 || !out) return IC_ERR_INVALID;
  int r = ic_device_check();
  if (r != IC_OK) return r;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ic_knn* k = new ic_knn();
  k->m = m;
  unsigned* bbox = nullptr;
  int* cursor = nullptr;
  int* block_sums = nullptr;
  auto cleanup = [&]() {
    cudaFree(bbox);
    cudaFree(cursor);
    cudaFree(block_sums);
  };
#define KB_CHECK(expr)                                                                                      \
  do {                                                                                                      \
    if ((expr) != cudaSuccess) {                                                                            \
      fprintf(stderr, "[icb] CUDA error in ic_knn_build: %s\n", cudaGetErrorString(cudaGetLastError())); \
      cleanup();                                                                                            \
      ic_knn_destroy(k);                                                                                    \
      return IC_ERR_CUDA;                                                                                   \
    }                                                                                                       \
  } while (0)
  KB_CHECK(cudaMalloc(&bbox, sizeof(unsigned) * 6));
  const unsigned init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
  KB_CHECK(cudaMemcpyAsync(bbox, init, sizeof(init), cudaMemcpyHostToDevice, st));
  knn_bbox_kernel<<<nblk(m, 256), 256, 0, st>>>(ref_xyz, m, stride, bbox);
  unsigned hb[6];
  KB_CHECK(cudaMemcpyAsync(hb, bbox, sizeof(hb), cudaMemcpyDeviceToHost, st));
  KB_CHECK(cudaStreamSynchronize(st));
  float lo[3], ext[3], cmax = 0.f;
  for (int a = 0; a < 3; ++a) {
    lo[a] = ord2f(hb[a]);
    const float hi = ord2f(hb[3 + a]);
    cmax = fmaxf(cmax, fmaxf(fabsf(lo[a]), fabsf(hi)));
    if (!isfinite(lo[a]) || !isfinite(hi)) {  // NaN / Inf coordinates: there is no nearest neighbour to speak of
      cleanup();
      ic_knn_destroy(k);
      return IC_ERR_INVALID;
    }
    ext[a] = hi - lo[a];
    k->org[a] = lo[a];
  }
  // cell size: caller's (the voxel size is the natural choice), else ~2 points per cell of the bounding volume
  double cell = cell_size;
  if (!(cell > 0)) {
    const double emax = fmax(ext[0], fmax(ext[1], ext[2]));
    double vol = 1;
    for (int a = 0; a < 3; ++a) vol *= fmax(ext[a], 1e-3 * emax);
    cell = emax > 0 ? cbrt(vol * 2.0 / static_cast<double>(m)) : 1.0;
    if (!(cell > 0)) cell = 1.0;
  }
  for (;;) {  // bound the directory: at most 2^27 cells (512 MB of int32)
    double cells = 1;
    for (int a = 0; a < 3; ++a) cells *= floor(ext[a] / cell) + 1;
    if (cells <= static_cast<double>(1 << 27)) break;
    cell *= 1.26;
  }
  k->cell = static_cast<float>(cell);
  k->eps = 1e-4f * k->cell + 1e-6f * cmax;
  k->n_cells = 1;
  for (int a = 0; a < 3; ++a) {
    k->dim[a] = static_cast<int>(floor(ext[a] / k->cell)) + 1;
    k->n_cells *= k->dim[a];
  }
  KnnGeom g{k->org[0], k->org[1], k->org[2], k->cell, 1.0f / k->cell, k->dim[0], k->dim[1], k->dim[2], k->eps};
  const long long n_scan = k->n_cells + 1;
  const int n_scan_blocks = static_cast<int>((n_scan + 1023) / 1024);
  KB_CHECK(cudaMalloc(&k->cell_start, sizeof(int) * n_scan));
  KB_CHECK(cudaMalloc(&k->pts, sizeof(float4) * m));
  KB_CHECK(cudaMalloc(&cursor, sizeof(int) * k->n_cells));
  KB_CHECK(cudaMalloc(&block_sums, sizeof(int) * n_scan_blocks));
  KB_CHECK(cudaMemsetAsync(k->cell_start, 0, sizeof(int) * n_scan, st));
  KB_CHECK(cudaMemsetAsync(cursor, 0, sizeof(int) * k->n_cells, st));
  knn_count_kernel<<<nblk(m, 256), 256, 0, st>>>(ref_xyz, m, stride, g, k->cell_start);
  knn_scan_local_kernel<<<n_scan_blocks, 1024, 0, st>>>(k->cell_start, n_scan, block_sums);
  knn_scan_sums_kernel<<<1, 1024, 0, st>>>(block_sums, n_scan_blocks);
  knn_scan_add_kernel<<<n_scan_blocks, 1024, 0, st>>>(k->cell_start, n_scan, block_sums);
  knn_fill_kernel<<<nblk(m, 256), 256, 0, st>>>(ref_xyz, m, stride, g, k->cell_start, cursor, k->pts);
  KB_CHECK(cudaStreamSynchronize(st));
  KB_CHECK(cudaGetLastError());
  cleanup();
#undef KB_CHECK
  *out = k;
  return IC_OK;
}

int ic_knn_info(const ic_knn* k, long long* n_points, long long* n_cells, float* cell_size, int* dims3_host) {
  if (!k) return IC_ERR_INVALID;
  if (n_points) *n_points = k->m;
  if (n_cells) *n_cells = k->n_cells;
  if (cell_size) *cell_size = k->cell;
  if (dims3_host)
    for (int a = 0; a < 3; ++a) dims3_host[a] = k->dim[a];
  return IC_OK;
}

int ic_knn_query1(const ic_knn* k, const float* queries, long long n, int stride, const long long* ref_labels,
                  int* out_idx, float* out_d2, long long* out_labels, void* stream) {
  if (!k || n < 0 || stride < 3 || (!out_idx && !out_d2 && !out_labels) || (out_labels && !ref_labels))
    return IC_ERR_INVALID;
  if (n == 0) return IC_OK;
  if (!queries) return IC_ERR_INVALID;
  int r = ic_device_check();
  if (r != IC_OK) return r;
  KnnGeom g{k->org[0], k->org[1], k->org[2], k->cell, 1.0f / k->cell, k->dim[0], k->dim[1], k->dim[2], k->eps};
  knn_query1_kernel<<<nblk(n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      k->pts, k->cell_start, g, queries, n, stride, ref_labels, out_idx, out_d2, out_labels);
  ICB_CUDA_CHECK(cudaGetLastError());
  return IC_OK;
}

int ic_knn_query(const ic_knn* k, const float* queries, long long n, int stride, int nb_points, int* out_idx,
                 float* out_d2, void* stream) {
  if (!k || n < 0 || stride < 3 || nb_points < 1 || nb_points > 32 || (!out_idx && !out_d2)) return IC_ERR_INVALID;
  if (n == 0) return IC_OK;
  if (!queries) return IC_ERR_INVALID;
  int r = ic_device_check();
  if (r != IC_OK) return r;
  KnnGeom g{k->org[0], k->org[1], k->org[2], k->cell, 1.0f / k->cell, k->dim[0], k->dim[1], k->dim[2], k->eps};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define ICB_KNN_K(K)                                                                                                  \
  knn_queryk_kernel<K><<<nblk(n, 128), 128, 0, st>>>(k->pts, k->cell_start, g, queries, n, stride, nb_points, out_idx, out_d2)
  if (nb_points <= 2)
    ICB_KNN_K(2);
  else if (nb_points <= 4)
    ICB_KNN_K(4);
  else if (nb_points <= 8)
    ICB_KNN_K(8);
  else if (nb_points <= 16)
    ICB_KNN_K(16);
  else
    ICB_KNN_K(32);
#undef ICB_KNN_K
  ICB_CUDA_CHECK(cudaGetLastError());
  return IC_OK;
}

}  // extern "C"
