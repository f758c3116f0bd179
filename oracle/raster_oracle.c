/*
 * CPU restatement of InfiniCube's voxel -> guidance-buffer rasteriser — TEST INFRASTRUCTURE ONLY.
 * (Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * compile, load or call this file; the product path in infinicube_b200/ never does.)
 *
 * Follows, in the reference checkout of nv-tlabs/InfiniCube:
 *   voxelisation + per-voxel arg-max label .... infinicube/utils/fvdb_utils.py:71-216  (points_to_fvdb)
 *       ijk = round((p - origin)/voxel_size)    fvdb_utils.py:155-157, utils/fvdb_test.py:76-79
 *       label = argmax count, ties -> smallest  fvdb_utils.py:174-191 (sorted unique + first argmax)
 *   pinhole rays at integer pixel centres ..... infinicube/camera/pinhole.py:110-138
 *   posed rays (d = R r_cam, o = t) ........... infinicube/camera/base.py:207-226
 *   depth: first run of active voxels >= 0.1 .. infinicube/camera/base.py:520-569 (segments_along_rays, eps=1e-1)
 *       zdepth = t0 * r_cam.z, miss = 0         camera/base.py:350-361, 545-550
 *   semantic / instance: first voxel whose
 *       chord >= 0.01, miss = background 0 .... infinicube/camera/base.py:571-618 (voxels_along_rays, eps=1e-2)
 *
 * PARITY UNPINNED for the traversal arithmetic: the ray/voxel intersection itself runs inside the
 * binary wheel fvdb==0.2.0+pt22cu121 (pyproject.toml:70), absent from /root/reference and this image,
 * and the reference has no golden vectors for it.  This file restates
 * fVDB's published HDDA semantics (hierarchical DDA in index space, voxel ijk spans
 * [ijk-1/2, ijk+1/2), front-to-back, eps on chord / run length) with the exact fp32 operation order
 * documented in DESIGN.md §Rasteriser; the CUDA kernel is written independently to the same spec and
 * must agree bit for bit.  ro_render_flat() is a second, brick-free traversal used by the tests to
 * cross-check the hierarchical walk.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC  (no FMA contraction: the operation
 * order below IS the specification).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  float vs[3], org[3];
  int imin[3], imax[3]; /* voxel-index bounding box (inclusive) */
  int bmin[3], bdim[3]; /* brick-grid origin (brick coords) and extent; brick = 8^3 voxels */
  int64_t n_bricks;
  uint64_t* mask; /* [n_bricks][8]: word = local z, bit = local y*8 + local x */
  int32_t* base;  /* [n_bricks]: voxel index of the brick's first active voxel */
  int64_t n_vox;
  int32_t* sem;  /* [n_vox] */
  int32_t* inst; /* [n_vox] */
} ro_grid;

static inline int floordiv8(int a) { return a >> 3; } /* arithmetic shift = floor for negatives */

static inline int64_t brick_lin(const ro_grid* g, int bx, int by, int bz) {
  return ((int64_t)(bz - g->bmin[2]) * g->bdim[1] + (by - g->bmin[1])) * g->bdim[0] + (bx - g->bmin[0]);
}

/* voxel index of ijk or -1 (fvdb ijk_to_index) */
static inline int64_t voxel_index(const ro_grid* g, int i, int j, int k) {
  if (i < g->imin[0] || i > g->imax[0] || j < g->imin[1] || j > g->imax[1] || k < g->imin[2] || k > g->imax[2])
    return -1;
  const int64_t b = brick_lin(g, floordiv8(i), floordiv8(j), floordiv8(k));
  const int lx = i & 7, ly = j & 7, lz = k & 7;
  const uint64_t* m = g->mask + b * 8;
  const int bit = ly * 8 + lx;
  if (!((m[lz] >> bit) & 1ull)) return -1;
  int r = 0;
  for (int z = 0; z < lz; ++z) r += __builtin_popcountll(m[z]);
  r += __builtin_popcountll(m[lz] & ((1ull << bit) - 1ull));
  return (int64_t)g->base[b] + r;
}

typedef struct {
  int64_t vox;
  int32_t label;
} pair_t;

static int pair_cmp(const void* a, const void* b) {
  const pair_t* x = (const pair_t*)a;
  const pair_t* y = (const pair_t*)b;
  if (x->vox != y->vox) return x->vox < y->vox ? -1 : 1;
  if (x->label != y->label) return x->label < y->label ? -1 : 1;
  return 0;
}

/* arg-max-count label per voxel, ties -> smallest label (fvdb_utils.py:174-191) */
static void argmax_labels(pair_t* pairs, int64_t m, int32_t* out) {
  qsort(pairs, (size_t)m, sizeof(pair_t), pair_cmp);
  int64_t s = 0;
  while (s < m) {
    const int64_t v = pairs[s].vox;
    int best_cnt = 0;
    int32_t best_lab = 0;
    int64_t e = s;
    while (e < m && pairs[e].vox == v) {
      int64_t r = e;
      while (r < m && pairs[r].vox == v && pairs[r].label == pairs[e].label) ++r;
      const int cnt = (int)(r - e);
      if (cnt > best_cnt) { /* strictly greater: the smallest label wins ties */
        best_cnt = cnt;
        best_lab = pairs[e].label;
      }
      e = r;
    }
    out[v] = best_lab;
    s = e;
  }
}

void ro_free(ro_grid* g) {
  if (!g) return;
  free(g->mask);
  free(g->base);
  free(g->sem);
  free(g->inst);
  free(g);
}

/* points_to_fvdb: points [m,3] fp32 (grid/world frame), per-point semantic + instance labels. */
ro_grid* ro_voxelize(const float* pts, int64_t m, const float* vs, const float* org, const int32_t* sem,
                     const int32_t* inst) {
  if (m <= 0) return NULL;
  ro_grid* g = (ro_grid*)calloc(1, sizeof(ro_grid));
  int32_t* ijk = (int32_t*)malloc(sizeof(int32_t) * 3 * (size_t)m);
  for (int a = 0; a < 3; ++a) {
    g->vs[a] = vs[a];
    g->org[a] = org[a];
    g->imin[a] = INT32_MAX;
    g->imax[a] = INT32_MIN;
  }
  for (int64_t p = 0; p < m; ++p)
    for (int a = 0; a < 3; ++a) {
      const float q = (pts[3 * p + a] - org[a]) / vs[a];
      const int v = (int)rintf(q); /* round half to even == torch.round().long() */
      ijk[3 * p + a] = v;
      if (v < g->imin[a]) g->imin[a] = v;
      if (v > g->imax[a]) g->imax[a] = v;
    }
  g->n_bricks = 1;
  for (int a = 0; a < 3; ++a) {
    g->bmin[a] = floordiv8(g->imin[a]);
    g->bdim[a] = floordiv8(g->imax[a]) - g->bmin[a] + 1;
    g->n_bricks *= g->bdim[a];
  }
  g->mask = (uint64_t*)calloc((size_t)g->n_bricks * 8, sizeof(uint64_t));
  g->base = (int32_t*)calloc((size_t)g->n_bricks, sizeof(int32_t));
  for (int64_t p = 0; p < m; ++p) {
    const int i = ijk[3 * p], j = ijk[3 * p + 1], k = ijk[3 * p + 2];
    const int64_t b = brick_lin(g, floordiv8(i), floordiv8(j), floordiv8(k));
    g->mask[b * 8 + (k & 7)] |= 1ull << ((j & 7) * 8 + (i & 7));
  }
  int64_t acc = 0;
  for (int64_t b = 0; b < g->n_bricks; ++b) {
    g->base[b] = (int32_t)acc;
    for (int z = 0; z < 8; ++z) acc += __builtin_popcountll(g->mask[b * 8 + z]);
  }
  g->n_vox = acc;
  g->sem = (int32_t*)calloc((size_t)acc, sizeof(int32_t));
  g->inst = (int32_t*)calloc((size_t)acc, sizeof(int32_t));
  pair_t* pairs = (pair_t*)malloc(sizeof(pair_t) * (size_t)m);
  for (int pass = 0; pass < 2; ++pass) {
    const int32_t* lab = pass == 0 ? sem : inst;
    if (!lab) continue;
    for (int64_t p = 0; p < m; ++p) {
      pairs[p].vox = voxel_index(g, ijk[3 * p], ijk[3 * p + 1], ijk[3 * p + 2]);
      pairs[p].label = lab[p];
    }
    argmax_labels(pairs, m, pass == 0 ? g->sem : g->inst);
  }
  free(pairs);
  free(ijk);
  return g;
}

int64_t ro_num_voxels(const ro_grid* g) { return g ? g->n_vox : 0; }
int64_t ro_num_bricks(const ro_grid* g) { return g ? g->n_bricks : 0; }
void ro_grid_info(const ro_grid* g, int* imin, int* imax, int* bmin, int* bdim) {
  for (int a = 0; a < 3; ++a) {
    imin[a] = g->imin[a];
    imax[a] = g->imax[a];
    bmin[a] = g->bmin[a];
    bdim[a] = g->bdim[a];
  }
}
/* export: ijk [n_vox,3] in voxel-index order + labels (for feeding the CUDA path the same grid) */
void ro_export(const ro_grid* g, int32_t* ijk, int32_t* sem, int32_t* inst) {
  for (int64_t b = 0; b < g->n_bricks; ++b) {
    const int bx = (int)(b % g->bdim[0]) + g->bmin[0];
    const int by = (int)((b / g->bdim[0]) % g->bdim[1]) + g->bmin[1];
    const int bz = (int)(b / ((int64_t)g->bdim[0] * g->bdim[1])) + g->bmin[2];
    int64_t v = g->base[b];
    for (int z = 0; z < 8; ++z)
      for (int bit = 0; bit < 64; ++bit)
        if ((g->mask[b * 8 + z] >> bit) & 1ull) {
          ijk[3 * v] = bx * 8 + (bit & 7);
          ijk[3 * v + 1] = by * 8 + (bit >> 3);
          ijk[3 * v + 2] = bz * 8 + z;
          ++v;
        }
  }
  memcpy(sem, g->sem, sizeof(int32_t) * (size_t)g->n_vox);
  memcpy(inst, g->inst, sizeof(int32_t) * (size_t)g->n_vox);
}

/* ---------------------------------------------------------------------------------------------
 * ray state shared by the two traversals
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  float oi[3], di[3], inv[3]; /* index-space origin (+0.5: cell = floor), direction, 1/direction */
  int step[3];
  float rcz;                  /* camera-space z of the unit ray: zdepth = t * rcz */
  /* hit bookkeeping */
  int sem_done, dep_done, run_open;
  float run_t0, run_t1, depth_t;
  int64_t hit_vox;
} ray_t;

static void ray_setup(ray_t* r, const ro_grid* g, const float* Kinv, const float* T, int u, int v) {
  const float fu = (float)u, fv = (float)v;
  float rc[3];
  for (int a = 0; a < 3; ++a) rc[a] = (Kinv[3 * a] * fu + Kinv[3 * a + 1] * fv) + Kinv[3 * a + 2];
  const float n = sqrtf((rc[0] * rc[0] + rc[1] * rc[1]) + rc[2] * rc[2]);
  for (int a = 0; a < 3; ++a) rc[a] = rc[a] / n;
  r->rcz = rc[2];
  for (int a = 0; a < 3; ++a) {
    const float d = (T[4 * a] * rc[0] + T[4 * a + 1] * rc[1]) + T[4 * a + 2] * rc[2];
    r->oi[a] = (T[4 * a + 3] - g->org[a]) / g->vs[a] + 0.5f;
    r->di[a] = d / g->vs[a];
    r->step[a] = r->di[a] > 0.f ? 1 : (r->di[a] < 0.f ? -1 : 0);
    r->inv[a] = r->step[a] ? 1.0f / r->di[a] : 0.f;
  }
  r->sem_done = r->dep_done = r->run_open = 0;
  r->run_t0 = r->run_t1 = r->depth_t = 0.f;
  r->hit_vox = -1;
}

/* clip against [lo, hi) per axis; returns 0 on miss */
static int ray_clip(const ray_t* r, const float* lo, const float* hi, float* tnear, float* tfar) {
  float tn = 0.f, tf = INFINITY;
  for (int a = 0; a < 3; ++a) {
    if (r->step[a] == 0) {
      if (r->oi[a] < lo[a] || r->oi[a] >= hi[a]) return 0;
    } else {
      const float t1 = (lo[a] - r->oi[a]) * r->inv[a];
      const float t2 = (hi[a] - r->oi[a]) * r->inv[a];
      tn = fmaxf(tn, fminf(t1, t2));
      tf = fminf(tf, fmaxf(t1, t2));
    }
  }
  *tnear = tn;
  *tfar = tf;
  return tn < tf;
}

static inline void close_run(ray_t* r) {
  if (r->run_open) {
    if (!r->dep_done && (r->run_t1 - r->run_t0) >= 0.1f) { /* eps = 1e-1, camera/base.py:543 */
      r->dep_done = 1;
      r->depth_t = r->run_t0;
    }
    r->run_open = 0;
  }
}

static inline void visit(ray_t* r, int64_t vox, float t0, float t1) {
  if (!(t0 < t1)) return; /* zero-length crossings neither hit nor break a run */
  if (vox >= 0) {
    if (!r->sem_done && (t1 - t0) >= 0.01f) { /* eps = 1e-2, camera/base.py:600 */
      r->sem_done = 1;
      r->hit_vox = vox;
    }
    if (!r->dep_done) {
      if (!r->run_open) {
        r->run_open = 1;
        r->run_t0 = t0;
      }
      r->run_t1 = t1;
    }
  } else {
    close_run(r);
  }
}

static inline float plane_t(const ray_t* r, int a, int cell, int size) {
  /* exit plane of `cell` (cells of `size` voxels) along axis a; +inf if the ray is parallel */
  if (r->step[a] == 0) return INFINITY;
  const float plane = (float)((cell + (r->step[a] > 0 ? 1 : 0)) * size);
  return (plane - r->oi[a]) * r->inv[a];
}

static inline int argmin3(const float* t) {
  int a = 0;
  if (t[1] < t[a]) a = 1;
  if (t[2] < t[a]) a = 2;
  return a;
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* hierarchical (brick -> voxel) walk */
static void trace_hdda(ray_t* r, const ro_grid* g) {
  float lo[3], hi[3], tnear, tfar;
  for (int a = 0; a < 3; ++a) {
    lo[a] = (float)(g->bmin[a] * 8);
    hi[a] = (float)((g->bmin[a] + g->bdim[a]) * 8);
  }
  if (!ray_clip(r, lo, hi, &tnear, &tfar)) return;
  int b[3];
  float tx[3];
  for (int a = 0; a < 3; ++a) {
    const float p = r->oi[a] + tnear * r->di[a];
    b[a] = clampi((int)floorf(p * 0.125f), g->bmin[a], g->bmin[a] + g->bdim[a] - 1);
    tx[a] = plane_t(r, a, b[a], 8);
  }
  float t = tnear;
  for (;;) {
    const int ax = argmin3(tx);
    float t_out = fminf(tx[ax], tfar);
    if (t_out < t) t_out = t; /* t never moves backwards */
    const int64_t bl = brick_lin(g, b[0], b[1], b[2]);
    const uint64_t* m = g->mask + bl * 8;
    const int nonempty = (m[0] | m[1] | m[2] | m[3] | m[4] | m[5] | m[6] | m[7]) != 0ull;
    if (!(t < t_out)) {
      /* zero-length brick crossing: ignored */
    } else if (!nonempty) {
      close_run(r);
    } else {
      /* voxel walk inside this brick over [t, t_out) */
      int c[3];
      float vx[3];
      for (int a = 0; a < 3; ++a) {
        const float p = r->oi[a] + t * r->di[a];
        c[a] = clampi((int)floorf(p), b[a] * 8, b[a] * 8 + 7);
        vx[a] = plane_t(r, a, c[a], 1);
      }
      float tc = t;
      for (;;) {
        const int va = argmin3(vx);
        float t1 = fminf(vx[va], t_out);
        if (t1 < tc) t1 = tc;
        const int lx = c[0] & 7, ly = c[1] & 7, lz = c[2] & 7;
        const int bit = ly * 8 + lx;
        int64_t vox = -1;
        if ((m[lz] >> bit) & 1ull) {
          int rk = 0;
          for (int z = 0; z < lz; ++z) rk += __builtin_popcountll(m[z]);
          rk += __builtin_popcountll(m[lz] & ((1ull << bit) - 1ull));
          vox = (int64_t)g->base[bl] + rk;
        }
        visit(r, vox, tc, t1);
        if (r->sem_done && r->dep_done) return;
        tc = t1;
        if (!(tc < t_out)) break;
        c[va] += r->step[va];
        if (c[va] < b[va] * 8 || c[va] > b[va] * 8 + 7) break;
        vx[va] = plane_t(r, va, c[va], 1);
      }
    }
    t = t_out;
    if (!(t < tfar)) break;
    b[ax] += r->step[ax];
    if (b[ax] < g->bmin[ax] || b[ax] >= g->bmin[ax] + g->bdim[ax]) break;
    tx[ax] = plane_t(r, ax, b[ax], 8);
  }
  close_run(r);
}

/* brick-free walk over the voxel bounding box (cross-check only) */
static void trace_flat(ray_t* r, const ro_grid* g) {
  float lo[3], hi[3], tnear, tfar;
  for (int a = 0; a < 3; ++a) {
    lo[a] = (float)(g->bmin[a] * 8);
    hi[a] = (float)((g->bmin[a] + g->bdim[a]) * 8);
  }
  if (!ray_clip(r, lo, hi, &tnear, &tfar)) return;
  int c[3];
  float vx[3];
  for (int a = 0; a < 3; ++a) {
    const float p = r->oi[a] + tnear * r->di[a];
    c[a] = clampi((int)floorf(p), g->bmin[a] * 8, (g->bmin[a] + g->bdim[a]) * 8 - 1);
    vx[a] = plane_t(r, a, c[a], 1);
  }
  float t = tnear;
  for (;;) {
    const int va = argmin3(vx);
    float t1 = fminf(vx[va], tfar);
    if (t1 < t) t1 = t;
    visit(r, voxel_index(g, c[0], c[1], c[2]), t, t1);
    if (r->sem_done && r->dep_done) return;
    t = t1;
    if (!(t < tfar)) break;
    c[va] += r->step[va];
    if (c[va] < g->bmin[va] * 8 || c[va] >= (g->bmin[va] + g->bdim[va]) * 8) break;
    vx[va] = plane_t(r, va, c[va], 1);
  }
  close_run(r);
}

static void render_impl(const ro_grid* g, const float* Kinv, const float* poses, int n_cam, int W, int H,
                        float* depth, int32_t* sem, int32_t* inst, int flat) {
  for (int n = 0; n < n_cam; ++n) {
    const float* T = poses + 16 * n;
    for (int v = 0; v < H; ++v)
      for (int u = 0; u < W; ++u) {
        ray_t r;
        ray_setup(&r, g, Kinv, T, u, v);
        if (flat)
          trace_flat(&r, g);
        else
          trace_hdda(&r, g);
        const size_t o = ((size_t)n * H + v) * W + u;
        depth[o] = r.dep_done ? r.depth_t * r.rcz : 0.f; /* camera/base.py:359-361, miss = 0 */
        sem[o] = r.sem_done ? g->sem[r.hit_vox] : 0;     /* background_semantic = 0, base.py:610 */
        inst[o] = r.sem_done ? g->inst[r.hit_vox] : 0;
      }
  }
}

void ro_render(const ro_grid* g, const float* Kinv, const float* poses, int n_cam, int W, int H, float* depth,
               int32_t* sem, int32_t* inst) {
  render_impl(g, Kinv, poses, n_cam, W, H, depth, sem, inst, 0);
}
void ro_render_flat(const ro_grid* g, const float* Kinv, const float* poses, int n_cam, int W, int H, float* depth,
                    int32_t* sem, int32_t* inst) {
  render_impl(g, Kinv, poses, n_cam, W, H, depth, sem, inst, 1);
}
/* rows [v0, v1) only — lets the CPU baseline time a bounded sample and use several threads */
void ro_render_rows(const ro_grid* g, const float* Kinv, const float* pose, int W, int H, int v0, int v1,
                    float* depth, int32_t* sem, int32_t* inst) {
  for (int v = v0; v < v1; ++v)
    for (int u = 0; u < W; ++u) {
      ray_t r;
      ray_setup(&r, g, Kinv, pose, u, v);
      trace_hdda(&r, g);
      const size_t o = (size_t)v * W + u;
      depth[o] = r.dep_done ? r.depth_t * r.rcz : 0.f;
      sem[o] = r.sem_done ? g->sem[r.hit_vox] : 0;
      inst[o] = r.sem_done ? g->inst[r.hit_vox] : 0;
    }
}

/* ---------------------------------------------------------------------------------------------
 * guidance images
 * ------------------------------------------------------------------------------------------- */
/* semantic_to_color + uint8 truncation (utils/semantic_utils.py:88-101, guidance_buffer_generation.py:693-695),
 * instance overlay with caller-supplied colours (utils/semantic_utils.py:104-131; the reference draws
 * them with unseeded np.random). palette_u8 [23][3]; inst_colors_u8 [n_ids][3]. */
void ro_semantic_rgb(const int32_t* sem, const int32_t* inst, int64_t n, const uint8_t* palette_u8,
                     const int32_t* inst_ids, const uint8_t* inst_colors_u8, int n_ids, uint8_t* rgb) {
  for (int64_t i = 0; i < n; ++i) {
    const uint8_t* c = palette_u8 + 3 * sem[i];
    if (inst[i] > 0) {
      for (int k = 0; k < n_ids; ++k)
        if (inst_ids[k] == inst[i]) {
          c = inst_colors_u8 + 3 * k;
          break;
        }
    }
    rgb[3 * i] = c[0];
    rgb[3 * i + 1] = c[1];
    rgb[3 * i + 2] = c[2];
  }
}
