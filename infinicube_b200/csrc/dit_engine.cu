// Wan2.1 DiT forward engine: owns weights + workspaces in HBM and strings the sm_100a kernels together
// (patch-embed GEMM + guidance-token add -> N x [AdaLN/QK-GEMM/V^T-GEMM/RMSNorm+RoPE/(all-gather)/FMHA/
// o-proj+gate residual, cross-attention, FFN] -> head).  Host code only; every FLOP runs in
// gemm_sm100.cu / fmha_sm100.cu / dit_ops.cu.
//
// Restates WanModel.forward / DiTBlock.forward of the reference's diffsynth dependency as driven by
// infinicube/videogen/inference.py:216-226.
#include <dlfcn.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/infinicube_b200.h"
#include "dit_ops.cuh"
#include "fmha_sm100.cuh"
#include "gemm_sm100.cuh"
#include "host_util.h"

using namespace icb;

namespace {

// ---------------------------------------------------------------------------------------------
// NCCL through dlopen: torch's bundled libnccl is already mapped into the process; no link-time dep.
// ---------------------------------------------------------------------------------------------
struct Id128 {  // ncclUniqueId (128 opaque bytes, passed by value)
  char b[128];
};
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, Id128, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.lib ? &api : nullptr;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) return nullptr;
  api.GetUniqueId = reinterpret_cast<int (*)(void*)>(dlsym(api.lib, "ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<int (*)(void**, int, Id128, int)>(dlsym(api.lib, "ncclCommInitRank"));
  api.AllGather = reinterpret_cast<int (*)(const void*, void*, size_t, int, void*, cudaStream_t)>(
      dlsym(api.lib, "ncclAllGather"));
  api.CommDestroy = reinterpret_cast<int (*)(void*)>(dlsym(api.lib, "ncclCommDestroy"));
  api.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(api.lib, "ncclGetErrorString"));
  if (!api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.CommDestroy) {
    api.lib = nullptr;
    return nullptr;
  }
  return &api;
}
constexpr int kNcclBfloat16 = 9;  // ncclDataType_t::ncclBfloat16

// Stream memory operations (driver API, resolved at run time like cuTensorMapEncodeTiled): used on LOCAL device
// memory only - "write epoch into a local staging word" and "wait until a local flag reached an epoch" - so that
// neither signalling nor flow control of the peer-memory exchange needs an SM.
struct MemOps {
  CUresult (*Write32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
  CUresult (*Wait32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
};
MemOps* memops() {
  static MemOps m;
  static bool tried = false;
  if (tried) return (m.Write32 && m.Wait32) ? &m : nullptr;
  tried = true;
  void* pw = nullptr;
  void* pq = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &pw, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &pq, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  m.Write32 = reinterpret_cast<CUresult (*)(CUstream, CUdeviceptr, cuuint32_t, unsigned int)>(pw);
  m.Wait32 = reinterpret_cast<CUresult (*)(CUstream, CUdeviceptr, cuuint32_t, unsigned int)>(pq);
  return &m;
}

// ---------------------------------------------------------------------------------------------
// small device kernels local to the engine
// ---------------------------------------------------------------------------------------------
__global__ void sinusoid_kernel(float t, float* out, int half) {
  // [cos(t*w_i) || sin(t*w_i)], w_i = 10000^(-i/half); fp64 like the reference (Appendix A.5)
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= half) return;
  const double w = pow(10000.0, -static_cast<double>(i) / static_cast<double>(half));
  const double a = static_cast<double>(t) * w;
  out[i] = static_cast<float>(cos(a));
  out[half + i] = static_cast<float>(sin(a));
}

__global__ void copy_convert_kernel(const void* src, int src_dtype, void* dst, int dst_dtype, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = src_dtype == IC_DTYPE_F32 ? static_cast<const float*>(src)[i]
                                      : __bfloat162float(static_cast<const __nv_bfloat16*>(src)[i]);
  if (dst_dtype == IC_DTYPE_F32)
    static_cast<float*>(dst)[i] = v;
  else
    static_cast<__nv_bfloat16*>(dst)[i] = __float2bfloat16(v);
}

struct Slot {
  void* ptr;
  int dtype;
  long long numel;
  bool loaded;
};

}  // namespace

struct ic_dit {
  ic_dit_config c;
  int D, F, H, L, S, Sall, hp, wp, CK;  // S = local tokens, CK = in_dim*4
  std::vector<void*> allocs;
  std::map<std::string, Slot> slots;
  long long bytes = 0;
  int launches = 0;

  // fused / converted weights
  struct Layer {
    __nv_bfloat16 *w_qk, *w_v, *w_o, *w_cq, *w_ck, *w_cv, *w_co, *w_f1, *w_f2;
    float *b_qk, *b_v, *b_o, *b_cq, *b_ck, *b_cv, *b_co, *b_f1, *b_f2;
    float *nq, *nk, *cnq, *cnk, *n3w, *n3b, *mod;  // mod [6, D]
  };
  std::vector<Layer> layers;
  __nv_bfloat16 *w_patch, *w_guide, *w_te0, *w_te2, *w_tm0, *w_tm2, *w_tp, *w_head;
  float *b_patch, *b_guide, *b_te0, *b_te2, *b_tm0, *b_tm2, *b_tp, *b_head, *head_mod;

  // workspaces
  float *x, *guide, *rowss, *sinus, *t_h, *t_emb, *t_mod, *emod, *head_e;
  __nv_bfloat16 *xn, *qk, *q, *attn, *hbuf, *patchA, *kv_all, *ctx_in, *ctx_h, *ctx_emb, *ctx_raw;
  __nv_bfloat16* ctx_k[2];   // [L][text_len, D]
  __nv_bfloat16* ctx_vt[2];  // [L][D, text_len]
  float *tab_f, *tab_h, *tab_w;
  RopeTables rope;
  bool has_guide = false;
  bool weights_checked = false;  // ic_dit_forward verified once that every registered tensor was loaded
  void* comm = nullptr;

  // optional in-stream CUDA-event profiling (bench.py roofline): pairs of events per launch, by kind
  struct ProfEvt {
    cudaEvent_t a, b;
    int kind;
  };
  std::vector<ProfEvt> prof;
  size_t prof_used = 0;
  int profiling = 0;        // bit k set: launches of kind k are bracketed by events
  bool prof_open = false;
  void prof_begin(int kind, cudaStream_t st) {
    prof_open = (profiling >> kind) & 1;
    if (!prof_open) return;
    if (prof_used == prof.size()) {
      ProfEvt e;
      cudaEventCreate(&e.a);
      cudaEventCreate(&e.b);
      prof.push_back(e);
    }
    prof[prof_used].kind = kind;
    cudaEventRecord(prof[prof_used].a, st);
  }
  void prof_end(cudaStream_t st) {
    if (!prof_open) return;
    prof_open = false;
    cudaEventRecord(prof[prof_used].b, st);
    ++prof_used;
  }

  template <typename T>
  T* alloc(long long n) {
    void* p = nullptr;
    if (cudaMalloc(&p, static_cast<size_t>(n) * sizeof(T)) != cudaSuccess) return nullptr;
    allocs.push_back(p);
    bytes += n * static_cast<long long>(sizeof(T));
    return static_cast<T*>(p);
  }
  void reg(const std::string& name, void* ptr, int dtype, long long numel) { slots[name] = Slot{ptr, dtype, numel, false}; }

  // Self-attention K / V^T gather buffer: [head group g][rank r][ K_g (S x Dg) || V^T_g (Dg x S) ].
  // One group by default (a single all-gather per attention); ICB_KV_GROUPS=2 splits it so that the all-gather
  // of group 1 overlaps the attention of group 0 on a separate NCCL stream.
  int n_groups = 1;

  // Peer-memory K / V^T exchange (opt-in, ic_dit_p2p_export / _attach): instead of an NCCL all-gather
  // every rank PUSHES its (K || V^T) segment into its peers' gather buffers with the copy engines (one stream per
  // peer, no SM involved) followed by a 4-byte epoch flag; the attention kernel starts on the local segment and
  // waits per remote segment (fmha_fwd's seg_ready), so the exchange overlaps the attention itself.  The buffer is
  // double-buffered by epoch parity and a peer's buffer is overwritten only after that peer has reported the
  // attention of two epochs ago done (`done` flags), all with full 32-bit monotonic epochs.
  struct P2P {
    static constexpr int kMaxRanks = 16;
    bool on = false;
    __nv_bfloat16* kv = nullptr;   // own [2 parity][world][chunk_elems], cudaMalloc'd and IPC-exported
    __nv_bfloat16* cur = nullptr;  // kv + parity * world * chunk_elems for the epoch in flight
    unsigned* flags = nullptr;     // own, IPC-exported: ready[2][world] | done[world] | scratch[2][world] (local staging)
    __nv_bfloat16* peer_kv[kMaxRanks] = {};
    unsigned* peer_flags[kMaxRanks] = {};
    cudaStream_t push[kMaxRanks] = {};
    bool serial = true;  // pushes to the peers run one after the other, in the order the peers consume them
    cudaEvent_t ev_pushed[2][kMaxRanks] = {};  // per parity and peer: the copy engine has finished reading the local segment
    cudaEvent_t ev_attn_done = nullptr;
    unsigned epoch = 0;
    unsigned* ready(int parity, int world) { return flags + parity * world; }
    unsigned* done(int world) { return flags + 2 * world; }
    unsigned* scratch(int kind, int world) { return flags + 3 * world + kind * world; }
    static int n_flags(int world) { return 5 * world; }
  } p2p;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_kv_ready = nullptr;
  cudaEvent_t ev_gathered[4] = {nullptr, nullptr, nullptr, nullptr};
  int Dg() const { return D / n_groups; }
  long long chunk_elems() const { return 2ll * S * Dg(); }
  __nv_bfloat16* group_base(int g) {
    return (p2p.on ? p2p.cur : kv_all) + static_cast<long long>(g) * c.world_size * chunk_elems();
  }
  __nv_bfloat16* k_local(int g) { return group_base(g) + static_cast<long long>(c.rank) * chunk_elems(); }
  __nv_bfloat16* vt_local(int g) { return k_local(g) + static_cast<long long>(S) * Dg(); }
};

namespace {

enum { PROF_FMHA_SELF = 0, PROF_FMHA_CROSS = 1, PROF_GEMM = 2, PROF_KINDS = 3 };

int gemm(ic_dit* h, const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb, int M, int N, int K,
         const GemmEpilogue& ep, cudaStream_t st) {
  h->launches++;
  h->prof_begin(PROF_GEMM, st);
  int r = gemm_bf16_tn(A, lda, B, ldb, M, N, K, ep, st);
  h->prof_end(st);
  return r;
}

#define IC_TRY(expr)          \
  do {                        \
    int _r = (expr);          \
    if (_r != IC_OK) return _r; \
  } while (0)

int build(ic_dit* h) {
  const ic_dit_config& c = h->c;
  const int D = h->D, F = h->F, L = h->L, S = h->S;
  auto wb = [&](long long n) { return h->alloc<__nv_bfloat16>(n); };
  auto wf = [&](long long n) { return h->alloc<float>(n); };
  bool ok = true;
  auto chk = [&](const void* p) { ok = ok && p != nullptr; };

  h->w_patch = wb(static_cast<long long>(D) * h->CK);
  h->b_patch = wf(D);
  h->reg("patch_embedding.weight", h->w_patch, IC_DTYPE_BF16, static_cast<long long>(D) * h->CK);
  h->reg("patch_embedding.bias", h->b_patch, IC_DTYPE_F32, D);
  if (c.guide_channels > 0) {
    h->w_guide = wb(static_cast<long long>(D) * c.guide_channels * 4);
    h->b_guide = wf(D);
    h->reg("buffer_embedder.weight", h->w_guide, IC_DTYPE_BF16, static_cast<long long>(D) * c.guide_channels * 4);
    h->reg("buffer_embedder.bias", h->b_guide, IC_DTYPE_F32, D);
  }
  h->w_te0 = wb(static_cast<long long>(D) * c.text_dim);
  h->b_te0 = wf(D);
  h->w_te2 = wb(static_cast<long long>(D) * D);
  h->b_te2 = wf(D);
  h->reg("text_embedding.0.weight", h->w_te0, IC_DTYPE_BF16, static_cast<long long>(D) * c.text_dim);
  h->reg("text_embedding.0.bias", h->b_te0, IC_DTYPE_F32, D);
  h->reg("text_embedding.2.weight", h->w_te2, IC_DTYPE_BF16, static_cast<long long>(D) * D);
  h->reg("text_embedding.2.bias", h->b_te2, IC_DTYPE_F32, D);
  h->w_tm0 = wb(static_cast<long long>(D) * c.freq_dim);
  h->b_tm0 = wf(D);
  h->w_tm2 = wb(static_cast<long long>(D) * D);
  h->b_tm2 = wf(D);
  h->w_tp = wb(6ll * D * D);
  h->b_tp = wf(6ll * D);
  h->reg("time_embedding.0.weight", h->w_tm0, IC_DTYPE_BF16, static_cast<long long>(D) * c.freq_dim);
  h->reg("time_embedding.0.bias", h->b_tm0, IC_DTYPE_F32, D);
  h->reg("time_embedding.2.weight", h->w_tm2, IC_DTYPE_BF16, static_cast<long long>(D) * D);
  h->reg("time_embedding.2.bias", h->b_tm2, IC_DTYPE_F32, D);
  h->reg("time_projection.1.weight", h->w_tp, IC_DTYPE_BF16, 6ll * D * D);
  h->reg("time_projection.1.bias", h->b_tp, IC_DTYPE_F32, 6ll * D);
  const int HO = 4 * c.out_dim;
  h->w_head = wb(static_cast<long long>(HO) * D);
  h->b_head = wf(HO);
  h->head_mod = wf(2ll * D);
  h->reg("head.head.weight", h->w_head, IC_DTYPE_BF16, static_cast<long long>(HO) * D);
  h->reg("head.head.bias", h->b_head, IC_DTYPE_F32, HO);
  h->reg("head.modulation", h->head_mod, IC_DTYPE_F32, 2ll * D);

  h->layers.resize(L);
  for (int i = 0; i < L; ++i) {
    ic_dit::Layer& l = h->layers[i];
    const long long DD = static_cast<long long>(D) * D;
    l.w_qk = wb(2 * DD);
    l.b_qk = wf(2 * D);
    l.w_v = wb(DD);
    l.b_v = wf(D);
    l.w_o = wb(DD);
    l.b_o = wf(D);
    l.w_cq = wb(DD);
    l.b_cq = wf(D);
    l.w_ck = wb(DD);
    l.b_ck = wf(D);
    l.w_cv = wb(DD);
    l.b_cv = wf(D);
    l.w_co = wb(DD);
    l.b_co = wf(D);
    l.w_f1 = wb(static_cast<long long>(F) * D);
    l.b_f1 = wf(F);
    l.w_f2 = wb(static_cast<long long>(D) * F);
    l.b_f2 = wf(D);
    l.nq = wf(D);
    l.nk = wf(D);
    l.cnq = wf(D);
    l.cnk = wf(D);
    l.n3w = wf(D);
    l.n3b = wf(D);
    l.mod = wf(6ll * D);
    chk(l.mod);
    const std::string p = "blocks." + std::to_string(i) + ".";
    h->reg(p + "self_attn.q.weight", l.w_qk, IC_DTYPE_BF16, DD);
    h->reg(p + "self_attn.k.weight", l.w_qk + DD, IC_DTYPE_BF16, DD);
    h->reg(p + "self_attn.q.bias", l.b_qk, IC_DTYPE_F32, D);
    h->reg(p + "self_attn.k.bias", l.b_qk + D, IC_DTYPE_F32, D);
    h->reg(p + "self_attn.v.weight", l.w_v, IC_DTYPE_BF16, DD);
    h->reg(p + "self_attn.v.bias", l.b_v, IC_DTYPE_F32, D);
    h->reg(p + "self_attn.o.weight", l.w_o, IC_DTYPE_BF16, DD);
    h->reg(p + "self_attn.o.bias", l.b_o, IC_DTYPE_F32, D);
    h->reg(p + "self_attn.norm_q.weight", l.nq, IC_DTYPE_F32, D);
    h->reg(p + "self_attn.norm_k.weight", l.nk, IC_DTYPE_F32, D);
    h->reg(p + "cross_attn.q.weight", l.w_cq, IC_DTYPE_BF16, DD);
    h->reg(p + "cross_attn.q.bias", l.b_cq, IC_DTYPE_F32, D);
    h->reg(p + "cross_attn.k.weight", l.w_ck, IC_DTYPE_BF16, DD);
    h->reg(p + "cross_attn.k.bias", l.b_ck, IC_DTYPE_F32, D);
    h->reg(p + "cross_attn.v.weight", l.w_cv, IC_DTYPE_BF16, DD);
    h->reg(p + "cross_attn.v.bias", l.b_cv, IC_DTYPE_F32, D);
    h->reg(p + "cross_attn.o.weight", l.w_co, IC_DTYPE_BF16, DD);
    h->reg(p + "cross_attn.o.bias", l.b_co, IC_DTYPE_F32, D);
    h->reg(p + "cross_attn.norm_q.weight", l.cnq, IC_DTYPE_F32, D);
    h->reg(p + "cross_attn.norm_k.weight", l.cnk, IC_DTYPE_F32, D);
    h->reg(p + "norm3.weight", l.n3w, IC_DTYPE_F32, D);
    h->reg(p + "norm3.bias", l.n3b, IC_DTYPE_F32, D);
    h->reg(p + "ffn.0.weight", l.w_f1, IC_DTYPE_BF16, static_cast<long long>(F) * D);
    h->reg(p + "ffn.0.bias", l.b_f1, IC_DTYPE_F32, F);
    h->reg(p + "ffn.2.weight", l.w_f2, IC_DTYPE_BF16, static_cast<long long>(D) * F);
    h->reg(p + "ffn.2.bias", l.b_f2, IC_DTYPE_F32, D);
    h->reg(p + "modulation", l.mod, IC_DTYPE_F32, 6ll * D);
  }

  // workspaces
  const long long SD = static_cast<long long>(S) * D;
  h->x = wf(SD);
  h->guide = c.guide_channels > 0 ? wf(SD) : nullptr;
  const int n_ss = 2 * ((D + 255) / 256) + 4;
  h->rowss = wf(static_cast<long long>(std::max(S, c.text_len)) * n_ss);
  h->sinus = wf(c.freq_dim);
  h->t_h = wf(D);
  h->t_emb = wf(D);
  h->t_mod = wf(6ll * D);
  h->emod = wf(6ll * D * L);
  h->head_e = wf(2ll * D);
  h->xn = wb(SD);
  h->qk = wb(2 * SD);
  h->q = wb(SD);
  h->attn = wb(SD);
  h->hbuf = wb(static_cast<long long>(S) * F);
  const int gk = std::max(h->CK, c.guide_channels * 4);
  h->patchA = wb(static_cast<long long>(S) * gk);
  h->kv_all = wb(h->chunk_elems() * h->n_groups * c.world_size);
  const long long TD = static_cast<long long>(c.text_len) * D;
  h->ctx_in = wb(static_cast<long long>(c.text_len) * c.text_dim);
  h->ctx_h = wb(TD);
  h->ctx_emb = wb(TD);
  h->ctx_raw = wb(TD);
  for (int s = 0; s < 2; ++s) {
    h->ctx_k[s] = wb(TD * L);
    h->ctx_vt[s] = wb(TD * L);
    chk(h->ctx_vt[s]);
  }
  chk(h->hbuf);
  chk(h->kv_all);
  chk(h->x);

  // RoPE tables (fp64 on host, Appendix A.4): pairs split 22 | 21 | 21, theta_j = 10000^(-2j/axis_dim)
  const int nf = c.lat_f, nh = h->hp, nw = h->wp;
  std::vector<float> tf(static_cast<size_t>(nf) * 22 * 2), th(static_cast<size_t>(nh) * 21 * 2),
      tw(static_cast<size_t>(nw) * 21 * 2);
  auto fill = [](std::vector<float>& t, int npos, int npair, int axis_dim) {
    for (int p = 0; p < npos; ++p)
      for (int j = 0; j < npair; ++j) {
        const double theta = pow(10000.0, -2.0 * j / static_cast<double>(axis_dim));
        const double a = p * theta;
        t[(static_cast<size_t>(p) * npair + j) * 2] = static_cast<float>(cos(a));
        t[(static_cast<size_t>(p) * npair + j) * 2 + 1] = static_cast<float>(sin(a));
      }
  };
  fill(tf, nf, 22, 44);
  fill(th, nh, 21, 42);
  fill(tw, nw, 21, 42);
  h->tab_f = wf(tf.size());
  h->tab_h = wf(th.size());
  h->tab_w = wf(tw.size());
  chk(h->tab_w);
  if (!ok) return IC_ERR_CUDA;
  ICB_CUDA_CHECK(cudaMemcpy(h->tab_f, tf.data(), tf.size() * 4, cudaMemcpyHostToDevice));
  ICB_CUDA_CHECK(cudaMemcpy(h->tab_h, th.data(), th.size() * 4, cudaMemcpyHostToDevice));
  ICB_CUDA_CHECK(cudaMemcpy(h->tab_w, tw.data(), tw.size() * 4, cudaMemcpyHostToDevice));
  h->rope = RopeTables{h->tab_f, h->tab_h, h->tab_w, nf, nh, nw};
  return IC_OK;
}

// time embedding -> t_emb, t_mod, per-layer modulation tables, head modulation
int time_stage(ic_dit* h, float t, cudaStream_t st) {
  const ic_dit_config& c = h->c;
  const int D = h->D;
  const int half = c.freq_dim / 2;
  sinusoid_kernel<<<(half + 127) / 128, 128, 0, st>>>(t, h->sinus, half);
  IC_TRY(gemv_bf16(h->w_tm0, c.freq_dim, h->sinus, h->b_tm0, h->t_h, D, c.freq_dim, 0, 1, st));
  IC_TRY(gemv_bf16(h->w_tm2, D, h->t_h, h->b_tm2, h->t_emb, D, D, 0, 0, st));
  IC_TRY(gemv_bf16(h->w_tp, D, h->t_emb, h->b_tp, h->t_mod, 6 * D, D, 1, 0, st));
  for (int i = 0; i < h->L; ++i)
    IC_TRY(add_bcast(h->layers[i].mod, h->t_mod, h->emod + 6ll * D * i, 6ll * D, 6 * D, st));
  // head: (shift, scale) = head.modulation[2, D] + t_emb[D]
  IC_TRY(add_bcast(h->head_mod, h->t_emb, h->head_e, 2ll * D, D, st));
  h->launches += 5 + h->L;
  return IC_OK;
}

int embed(ic_dit* h, const float* latents, float t, cudaStream_t st) {
  const ic_dit_config& c = h->c;
  IC_TRY(time_stage(h, t, st));
  IC_TRY(patchify(latents, h->patchA, c.in_dim, c.frames_local, c.lat_h, c.lat_w, h->CK, 0, st));
  GemmEpilogue ep;
  ep.bias = h->b_patch;
  ep.out_f32 = h->x;
  ep.ld_f32 = h->D;
  if (h->has_guide) {
    ep.addend = h->guide;
    ep.ld_add = h->D;
  }
  IC_TRY(gemm(h, h->patchA, h->CK, h->w_patch, h->CK, h->S, h->D, h->CK, ep, st));
  h->launches += 1;
  return IC_OK;
}

// One DiT block in three parts so that a test can stand in for the K / V^T exchange (ic_dit_run_block_phase):
//   BLOCK_PRODUCE  LN+mod, QK GEMM, V^T GEMM, RMSNorm+RoPE -> q and this rank's (K || V^T) segment
//   BLOCK_EXCHANGE peer-memory push or NCCL all-gather of the segments (nothing on one rank)
//   BLOCK_CONSUME  self-attention over all segments, o-proj, cross-attention, FFN
enum { BLOCK_PRODUCE = 1, BLOCK_EXCHANGE = 2, BLOCK_CONSUME = 4, BLOCK_ALL = 7 };

int run_block(ic_dit* h, int li, int slot, cudaStream_t st, int parts = BLOCK_ALL) {
  const ic_dit_config& c = h->c;
  const int D = h->D, F = h->F, S = h->S, H = h->H;
  ic_dit::Layer& l = h->layers[li];
  const float* e = h->emod + 6ll * D * li;  // shift1, scale1, gate1, shift2, scale2, gate2
  const int n_ss = 2 * ((D + 255) / 256) + 4;
  const int ss_per = (D + gemm_block_n(2 * D) - 1) / gemm_block_n(2 * D);
  const float scale = 1.0f / sqrtf(128.0f);
  if (h->p2p.on && parts != BLOCK_ALL) return IC_ERR_UNSUPPORTED;  // the push protocol spans the three parts

  const int G = h->n_groups, Dg = h->Dg();
  unsigned p2p_epoch = 0;

  // ---- self attention ----
  if (parts & BLOCK_PRODUCE) {
    IC_TRY(ln_modulate(h->x, D, e + D, e, 1, h->xn, D, S, D, c.eps, st));
    {
      GemmEpilogue ep;
      ep.bias = l.b_qk;
      ep.out_bf16 = h->qk;
      ep.ld_out = 2 * D;
      ep.rowss = h->rowss;
      ep.rowss_ld = n_ss;
      IC_TRY(gemm(h, h->xn, D, l.w_qk, D, S, 2 * D, D, ep, st));
    }
    if (h->p2p.on) {  // next epoch: producers and attention of this layer use the buffer of its parity
      p2p_epoch = ++h->p2p.epoch;
      h->p2p.cur = h->p2p.kv + static_cast<long long>(p2p_epoch & 1) * c.world_size * h->chunk_elems();
      if (p2p_epoch > 2) {  // the pushes of epoch - 2 read the local segment the producers below overwrite
        for (int r = 0; r < c.world_size; ++r)
          if (r != c.rank) ICB_CUDA_CHECK(cudaStreamWaitEvent(st, h->p2p.ev_pushed[p2p_epoch & 1][r], 0));
      }
    }
    for (int g = 0; g < G; ++g) {
      // V^T_g = W_v[g] * xn^T  (roles swapped so that the attention's second GEMM gets a K-major B operand)
      GemmEpilogue ep;
      ep.bias = l.b_v + g * Dg;
      ep.bias_per_row = 1;
      ep.out_bf16 = h->vt_local(g);
      ep.ld_out = S;
      IC_TRY(gemm(h, l.w_v + static_cast<long long>(g) * Dg * D, D, h->xn, D, Dg, S, D, ep, st));
    }
    IC_TRY(rmsnorm_rope(h->qk, 2 * D, h->rowss, n_ss, 0, ss_per, l.nq, h->q, D, S, D, c.eps, &h->rope, c.frame0, st));
    IC_TRY(rmsnorm_rope(h->qk + D, 2 * D, h->rowss, n_ss, ss_per, ss_per, l.nk, h->k_local(0), Dg, S, D, c.eps, &h->rope,
                        c.frame0, st, Dg, static_cast<long long>(c.world_size) * h->chunk_elems()));
    h->launches += 2;
  }  // BLOCK_PRODUCE
  if (h->p2p.on) {
    // push this rank's segment to every peer with the copy engines (the peer that consumes it first is served
    // first), each copy followed by the epoch flag of that (parity, segment) slot
    MemOps* mo = memops();
    if (!mo) return IC_ERR_UNSUPPORTED;
    ic_dit::P2P& P = h->p2p;
    const int W = c.world_size, me = c.rank, par = static_cast<int>(p2p_epoch & 1);
    const size_t seg_bytes = static_cast<size_t>(h->chunk_elems()) * sizeof(__nv_bfloat16);
    const long long seg_off = (static_cast<long long>(par) * W + me) * h->chunk_elems();
    ICB_CUDA_CHECK(cudaEventRecord(h->ev_kv_ready, st));
    for (int i = 1; i < W; ++i) {
      const int pr = (me - i + W) % W;
      // One stream for all peers by default: the copies then leave in the order the peers will reach this segment
      // (me-1 first), each at full NVLink rate.  With one stream per peer the W-1 copies share the egress and ALL
      // arrive late - the peer that needs this segment right after its local one waits for a third of the bytes of
      // everybody else's (measured at 8 GPUs: attention +0.13 ms per launch).
      cudaStream_t ps = P.serial ? P.push[(me + 1) % W] : P.push[pr];
      ICB_CUDA_CHECK(cudaStreamWaitEvent(ps, h->ev_kv_ready, 0));
      if (p2p_epoch > 2) {  // peer pr still reads this parity's buffer until its attention of epoch - 2 is done
        if (mo->Wait32(ps, reinterpret_cast<CUdeviceptr>(P.done(W) + pr), p2p_epoch - 2, CU_STREAM_WAIT_VALUE_GEQ) !=
            CUDA_SUCCESS)
          return IC_ERR_CUDA;
      }
      ICB_CUDA_CHECK(cudaMemcpyAsync(P.peer_kv[pr] + seg_off, P.kv + seg_off, seg_bytes, cudaMemcpyDeviceToDevice, ps));
      ICB_CUDA_CHECK(cudaEventRecord(P.ev_pushed[par][pr], ps));
      unsigned* stage = P.scratch(0, W) + pr;
      if (mo->Write32(ps, reinterpret_cast<CUdeviceptr>(stage), p2p_epoch, CU_STREAM_WRITE_VALUE_DEFAULT) != CUDA_SUCCESS)
        return IC_ERR_CUDA;
      ICB_CUDA_CHECK(cudaMemcpyAsync(P.peer_flags[pr] + par * W + me, stage, sizeof(unsigned), cudaMemcpyDeviceToDevice, ps));
    }
    h->prof_begin(PROF_FMHA_SELF, st);
    IC_TRY(fmha_fwd(h->q, D, h->group_base(0), Dg, h->chunk_elems(), h->group_base(0) + static_cast<long long>(S) * Dg, S,
                    h->chunk_elems(), h->attn, D, S, S, W, H, scale, st, me, P.ready(par, W), p2p_epoch));
    h->prof_end(st);
    h->launches += 1;
    // tell every peer that this rank no longer reads the epoch's buffer (flow control for epoch + 2)
    ICB_CUDA_CHECK(cudaEventRecord(P.ev_attn_done, st));
    for (int i = 1; i < W; ++i) {
      const int pr = (me - i + W) % W;
      cudaStream_t ps = P.serial ? P.push[(me + 1) % W] : P.push[pr];
      ICB_CUDA_CHECK(cudaStreamWaitEvent(ps, P.ev_attn_done, 0));
      unsigned* stage = P.scratch(1, W) + pr;
      if (mo->Write32(ps, reinterpret_cast<CUdeviceptr>(stage), p2p_epoch, CU_STREAM_WRITE_VALUE_DEFAULT) != CUDA_SUCCESS)
        return IC_ERR_CUDA;
      ICB_CUDA_CHECK(cudaMemcpyAsync(P.peer_flags[pr] + 2 * W + me, stage, sizeof(unsigned), cudaMemcpyDeviceToDevice, ps));
    }
  } else if (c.world_size > 1 && (parts & BLOCK_EXCHANGE)) {
    NcclApi* api = nccl_api();
    if (!api || !h->comm) return IC_ERR_NCCL;
    // per head group: in-place all-gather on the communication stream; attention of group g starts as soon as
    // its own gather has landed, while the next group's gather is still in flight
    ICB_CUDA_CHECK(cudaEventRecord(h->ev_kv_ready, st));
    ICB_CUDA_CHECK(cudaStreamWaitEvent(h->comm_stream, h->ev_kv_ready, 0));
    for (int g = 0; g < G; ++g) {
      int r = api->AllGather(h->k_local(g), h->group_base(g), static_cast<size_t>(h->chunk_elems()), kNcclBfloat16,
                             h->comm, h->comm_stream);
      if (r != 0) {
        fprintf(stderr, "[icb] ncclAllGather failed: %s\n", api->GetErrorString ? api->GetErrorString(r) : "?");
        return IC_ERR_NCCL;
      }
      ICB_CUDA_CHECK(cudaEventRecord(h->ev_gathered[g], h->comm_stream));
    }
  }
  if (!(parts & BLOCK_CONSUME)) return IC_OK;
  for (int g = 0; g < G && !h->p2p.on; ++g) {
    if (c.world_size > 1 && (parts & BLOCK_EXCHANGE)) ICB_CUDA_CHECK(cudaStreamWaitEvent(st, h->ev_gathered[g], 0));
    h->prof_begin(PROF_FMHA_SELF, st);
    IC_TRY(fmha_fwd(h->q + g * Dg, D, h->group_base(g), Dg, h->chunk_elems(),
                    h->group_base(g) + static_cast<long long>(S) * Dg, S, h->chunk_elems(), h->attn + g * Dg, D, S, S,
                    c.world_size, H / G, scale, st));
    h->prof_end(st);
    h->launches += 1;
  }
  {
    GemmEpilogue ep;
    ep.bias = l.b_o;
    ep.resid = h->x;
    ep.ld_res = D;
    ep.gate = e + 2 * D;
    IC_TRY(gemm(h, h->attn, D, l.w_o, D, S, D, D, ep, st));
  }

  // ---- cross attention ----
  IC_TRY(ln_modulate(h->x, D, l.n3w, l.n3b, 0, h->xn, D, S, D, c.eps, st));
  {
    GemmEpilogue ep;
    ep.bias = l.b_cq;
    ep.out_bf16 = h->qk;
    ep.ld_out = D;
    ep.rowss = h->rowss;
    ep.rowss_ld = n_ss;
    IC_TRY(gemm(h, h->xn, D, l.w_cq, D, S, D, D, ep, st));
  }
  const int ss_c = (D + gemm_block_n(D) - 1) / gemm_block_n(D);
  IC_TRY(rmsnorm_rope(h->qk, D, h->rowss, n_ss, 0, ss_c, l.cnq, h->q, D, S, D, c.eps, nullptr, 0, st));
  const long long TD = static_cast<long long>(c.text_len) * D;
  h->prof_begin(PROF_FMHA_CROSS, st);
  IC_TRY(fmha_fwd(h->q, D, h->ctx_k[slot] + TD * li, D, 0, h->ctx_vt[slot] + TD * li, c.text_len, 0, h->attn, D, S,
                  c.text_len, 1, H, scale, st));
  h->prof_end(st);
  h->launches += 3;
  {
    GemmEpilogue ep;
    ep.bias = l.b_co;
    ep.resid = h->x;
    ep.ld_res = D;
    IC_TRY(gemm(h, h->attn, D, l.w_co, D, S, D, D, ep, st));
  }

  // ---- feed forward ----
  IC_TRY(ln_modulate(h->x, D, e + 4 * D, e + 3 * D, 1, h->xn, D, S, D, c.eps, st));
  h->launches += 1;
  {
    GemmEpilogue ep;
    ep.bias = l.b_f1;
    ep.act = 1;
    ep.out_bf16 = h->hbuf;
    ep.ld_out = F;
    IC_TRY(gemm(h, h->xn, D, l.w_f1, D, S, F, D, ep, st));
  }
  {
    GemmEpilogue ep;
    ep.bias = l.b_f2;
    ep.resid = h->x;
    ep.ld_res = D;
    ep.gate = e + 5 * D;
    IC_TRY(gemm(h, h->hbuf, F, l.w_f2, F, S, D, F, ep, st));
  }
  return IC_OK;
}

int head(ic_dit* h, float* head_out, cudaStream_t st) {
  const int D = h->D;
  IC_TRY(ln_modulate(h->x, D, h->head_e + D, h->head_e, 1, h->xn, D, h->S, D, h->c.eps, st));
  h->launches += 1;
  GemmEpilogue ep;
  ep.bias = h->b_head;
  ep.out_f32 = head_out;
  ep.ld_f32 = 4 * h->c.out_dim;
  return gemm(h, h->xn, D, h->w_head, D, h->S, 4 * h->c.out_dim, D, ep, st);
}

}  // namespace

extern "C" {

int ic_dit_create(const ic_dit_config* cfg, ic_dit** out) {
  if (!cfg || !out) return IC_ERR_INVALID;
  const ic_dit_config& c = *cfg;
  if (c.dim <= 0 || c.dim % 128 || c.num_heads * 128 != c.dim || c.ffn_dim % 8 || c.num_layers <= 0) return IC_ERR_INVALID;
  if (c.lat_h % 2 || c.lat_w % 2 || c.frames_local <= 0 || c.frame0 < 0 || c.frame0 + c.frames_local > c.lat_f)
    return IC_ERR_INVALID;
  if (c.world_size < 1 || c.rank < 0 || c.rank >= c.world_size) return IC_ERR_INVALID;
  if (c.world_size > 1 && c.frames_local * c.world_size != c.lat_f) return IC_ERR_INVALID;  // equal shards
  if (c.text_len % 8 || c.text_dim % 8 || c.freq_dim % 8 || (c.in_dim * 4) % 8 || (c.out_dim * 4) % 8) return IC_ERR_INVALID;
  int r = ic_device_check();
  if (r != IC_OK) return r;
  ic_dit* h = new ic_dit();
  h->c = c;
  h->D = c.dim;
  h->F = c.ffn_dim;
  h->H = c.num_heads;
  h->L = c.num_layers;
  h->hp = c.lat_h / 2;
  h->wp = c.lat_w / 2;
  h->S = c.frames_local * h->hp * h->wp;
  h->Sall = c.lat_f * h->hp * h->wp;
  h->CK = c.in_dim * 4;
  h->n_groups = 1;  // measured: splitting the attention launch costs more than the overlap hides
  if (const char* e = getenv("ICB_KV_GROUPS")) {  // > 1: gather head groups separately, pipelined with attention
    const int g = atoi(e);
    if (g >= 1 && g <= 4 && c.num_heads % g == 0) h->n_groups = g;
  }
  if (c.world_size > 1) {
    cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&h->ev_kv_ready, cudaEventDisableTiming);
    for (int g = 0; g < h->n_groups; ++g) cudaEventCreateWithFlags(&h->ev_gathered[g], cudaEventDisableTiming);
  }
  if (h->S % 8) {
    delete h;
    return IC_ERR_INVALID;
  }
  r = build(h);
  if (r != IC_OK) {
    ic_dit_destroy(h);
    return r;
  }
  *out = h;
  return IC_OK;
}

int ic_dit_destroy(ic_dit* h) {
  if (!h) return IC_OK;
  if (h->comm) {
    NcclApi* api = nccl_api();
    if (api) api->CommDestroy(h->comm);
  }
  if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
  if (h->p2p.kv) {
    cudaDeviceSynchronize();  // no push may still target a peer, no peer copy may be mid-flight from this buffer
    for (int r = 0; r < h->c.world_size && r < ic_dit::P2P::kMaxRanks; ++r) {
      if (r != h->c.rank && h->p2p.peer_kv[r]) cudaIpcCloseMemHandle(h->p2p.peer_kv[r]);
      if (r != h->c.rank && h->p2p.peer_flags[r]) cudaIpcCloseMemHandle(h->p2p.peer_flags[r]);
      if (h->p2p.push[r]) cudaStreamDestroy(h->p2p.push[r]);
      for (int b = 0; b < 2; ++b)
        if (h->p2p.ev_pushed[b][r]) cudaEventDestroy(h->p2p.ev_pushed[b][r]);
    }
    if (h->p2p.ev_attn_done) cudaEventDestroy(h->p2p.ev_attn_done);
    cudaFree(h->p2p.kv);
    cudaFree(h->p2p.flags);
  }
  if (h->ev_kv_ready) cudaEventDestroy(h->ev_kv_ready);
  for (auto& ev : h->ev_gathered)
    if (ev) cudaEventDestroy(ev);
  for (auto& e : h->prof) {
    cudaEventDestroy(e.a);
    cudaEventDestroy(e.b);
  }
  for (void* p : h->allocs) cudaFree(p);
  delete h;
  return IC_OK;
}

long long ic_dit_workspace_bytes(const ic_dit* h) { return h ? h->bytes : 0; }

int ic_dit_load_tensor(ic_dit* h, const char* name, const void* src, int dtype, long long numel, void* stream) {
  if (!h || !name || !src) return IC_ERR_INVALID;
  auto it = h->slots.find(name);
  if (it == h->slots.end()) return IC_ERR_INVALID;
  Slot& s = it->second;
  if (s.numel != numel) {
    fprintf(stderr, "[icb] load_tensor %s: expected %lld elements, got %lld\n", name, s.numel, numel);
    return IC_ERR_INVALID;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  copy_convert_kernel<<<static_cast<unsigned>((numel + 255) / 256), 256, 0, st>>>(src, dtype, s.ptr, s.dtype, numel);
  ICB_CUDA_CHECK(cudaGetLastError());
  s.loaded = true;
  return IC_OK;
}

int ic_nccl_unique_id(void* id) {
  NcclApi* api = nccl_api();
  if (!api || !id) return IC_ERR_NCCL;
  return api->GetUniqueId(id) == 0 ? IC_OK : IC_ERR_NCCL;
}

int ic_dit_init_comm(ic_dit* h, const void* id) {
  if (!h || !id) return IC_ERR_INVALID;
  if (h->c.world_size == 1) return IC_OK;
  NcclApi* api = nccl_api();
  if (!api) return IC_ERR_NCCL;
  Id128 uid;
  memcpy(uid.b, id, 128);
  int r = api->CommInitRank(&h->comm, h->c.world_size, uid, h->c.rank);
  if (r != 0) {
    fprintf(stderr, "[icb] ncclCommInitRank failed: %s\n", api->GetErrorString ? api->GetErrorString(r) : "?");
    return IC_ERR_NCCL;
  }
  return IC_OK;
}

int ic_dit_p2p_export(ic_dit* h, void* handles_host_128B) {
  if (!h || !handles_host_128B) return IC_ERR_INVALID;
  const int W = h->c.world_size;
  if (W < 2 || W > ic_dit::P2P::kMaxRanks || h->n_groups != 1) return IC_ERR_UNSUPPORTED;
  if (!memops()) return IC_ERR_UNSUPPORTED;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  ic_dit::P2P& P = h->p2p;
  if (!P.kv) {
    const size_t bytes = sizeof(__nv_bfloat16) * 2 * W * static_cast<size_t>(h->chunk_elems());
    ICB_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&P.kv), bytes));
    ICB_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&P.flags), sizeof(unsigned) * ic_dit::P2P::n_flags(W)));
    ICB_CUDA_CHECK(cudaMemset(P.flags, 0, sizeof(unsigned) * ic_dit::P2P::n_flags(W)));
    ICB_CUDA_CHECK(cudaDeviceSynchronize());  // flags are zero before any peer can learn the handle
    h->bytes += static_cast<long long>(bytes);
  }
  cudaIpcMemHandle_t hk, hf;
  ICB_CUDA_CHECK(cudaIpcGetMemHandle(&hk, P.kv));
  ICB_CUDA_CHECK(cudaIpcGetMemHandle(&hf, P.flags));
  memcpy(handles_host_128B, &hk, 64);
  memcpy(static_cast<char*>(handles_host_128B) + 64, &hf, 64);
  return IC_OK;
}

int ic_dit_p2p_attach(ic_dit* h, const void* all_handles_host) {
  if (!h || !all_handles_host) return IC_ERR_INVALID;
  ic_dit::P2P& P = h->p2p;
  const int W = h->c.world_size, me = h->c.rank;
  if (!P.kv || P.on) return IC_ERR_INVALID;
  int lo = 0, hi = 0;
  ICB_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  for (int r = 0; r < W; ++r) {
    if (r == me) {
      P.peer_kv[r] = P.kv;
      P.peer_flags[r] = P.flags;
      continue;
    }
    cudaIpcMemHandle_t hk, hf;
    memcpy(&hk, static_cast<const char*>(all_handles_host) + 128 * r, 64);
    memcpy(&hf, static_cast<const char*>(all_handles_host) + 128 * r + 64, 64);
    ICB_CUDA_CHECK(cudaIpcOpenMemHandle(reinterpret_cast<void**>(&P.peer_kv[r]), hk, cudaIpcMemLazyEnablePeerAccess));
    ICB_CUDA_CHECK(cudaIpcOpenMemHandle(reinterpret_cast<void**>(&P.peer_flags[r]), hf, cudaIpcMemLazyEnablePeerAccess));
    ICB_CUDA_CHECK(cudaStreamCreateWithPriority(&P.push[r], cudaStreamNonBlocking, hi));
    for (int b = 0; b < 2; ++b) ICB_CUDA_CHECK(cudaEventCreateWithFlags(&P.ev_pushed[b][r], cudaEventDisableTiming));
  }
  ICB_CUDA_CHECK(cudaEventCreateWithFlags(&P.ev_attn_done, cudaEventDisableTiming));
  if (const char* e = getenv("ICB_P2P_SERIAL")) P.serial = atoi(e) != 0;
  P.cur = P.kv;
  P.epoch = 0;
  P.on = true;
  return IC_OK;
}

int ic_dit_p2p_enabled(const ic_dit* h) { return (h && h->p2p.on) ? 1 : 0; }

int ic_dit_set_context(ic_dit* h, int slot, const void* ctx, int dtype, void* stream) {
  if (!h || !ctx || slot < 0 || slot > 1) return IC_ERR_INVALID;
  const ic_dit_config& c = h->c;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int D = h->D, T = c.text_len;
  const long long n = static_cast<long long>(T) * c.text_dim;
  copy_convert_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(ctx, dtype, h->ctx_in, IC_DTYPE_BF16, n);
  ICB_CUDA_CHECK(cudaGetLastError());
  {
    GemmEpilogue ep;
    ep.bias = h->b_te0;
    ep.act = 1;
    ep.out_bf16 = h->ctx_h;
    ep.ld_out = D;
    IC_TRY(gemm(h, h->ctx_in, c.text_dim, h->w_te0, c.text_dim, T, D, c.text_dim, ep, st));
  }
  {
    GemmEpilogue ep;
    ep.bias = h->b_te2;
    ep.out_bf16 = h->ctx_emb;
    ep.ld_out = D;
    IC_TRY(gemm(h, h->ctx_h, D, h->w_te2, D, T, D, D, ep, st));
  }
  const int n_ss = 2 * ((D + 255) / 256) + 4;
  const int ss_c = (D + gemm_block_n(D) - 1) / gemm_block_n(D);
  const long long TD = static_cast<long long>(T) * D;
  for (int i = 0; i < h->L; ++i) {
    ic_dit::Layer& l = h->layers[i];
    {
      GemmEpilogue ep;
      ep.bias = l.b_ck;
      ep.out_bf16 = h->ctx_raw;
      ep.ld_out = D;
      ep.rowss = h->rowss;
      ep.rowss_ld = n_ss;
      IC_TRY(gemm(h, h->ctx_emb, D, l.w_ck, D, T, D, D, ep, st));
    }
    IC_TRY(rmsnorm_rope(h->ctx_raw, D, h->rowss, n_ss, 0, ss_c, l.cnk, h->ctx_k[slot] + TD * i, D, T, D, c.eps, nullptr,
                        0, st));
    {
      GemmEpilogue ep;
      ep.bias = l.b_cv;
      ep.bias_per_row = 1;
      ep.out_bf16 = h->ctx_vt[slot] + TD * i;
      ep.ld_out = T;
      IC_TRY(gemm(h, l.w_cv, D, h->ctx_emb, D, D, T, D, ep, st));
    }
  }
  return IC_OK;
}

int ic_dit_set_guidance(ic_dit* h, const float* guide_latents, void* stream) {
  if (!h) return IC_ERR_INVALID;
  if (!guide_latents) {
    h->has_guide = false;
    return IC_OK;
  }
  const ic_dit_config& c = h->c;
  if (c.guide_channels <= 0) return IC_ERR_INVALID;
  for (const char* n : {"buffer_embedder.weight", "buffer_embedder.bias"})
    if (!h->slots[n].loaded) {
      fprintf(stderr, "[icb] ic_dit_set_guidance: %s was never loaded\n", n);
      return IC_ERR_INVALID;
    }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int GK = c.guide_channels * 4;
  IC_TRY(patchify(guide_latents, h->patchA, c.guide_channels, c.frames_local, c.lat_h, c.lat_w, GK, 0, st));
  GemmEpilogue ep;
  ep.bias = h->b_guide;
  ep.out_f32 = h->guide;
  ep.ld_f32 = h->D;
  IC_TRY(gemm(h, h->patchA, GK, h->w_guide, GK, h->S, h->D, GK, ep, st));
  h->has_guide = true;
  return IC_OK;
}

int ic_dit_embed(ic_dit* h, const float* latents, float timestep, void* stream) {
  if (!h || !latents) return IC_ERR_INVALID;
  return embed(h, latents, timestep, static_cast<cudaStream_t>(stream));
}
int ic_dit_run_block(ic_dit* h, int layer, int ctx_slot, void* stream) {
  if (!h || layer < 0 || layer >= h->L || ctx_slot < 0 || ctx_slot > 1) return IC_ERR_INVALID;
  return run_block(h, layer, ctx_slot, static_cast<cudaStream_t>(stream));
}
int ic_dit_run_block_phase(ic_dit* h, int layer, int ctx_slot, int phase, void* stream) {
  if (!h || layer < 0 || layer >= h->L || ctx_slot < 0 || ctx_slot > 1 || (phase != 0 && phase != 1)) return IC_ERR_INVALID;
  return run_block(h, layer, ctx_slot, static_cast<cudaStream_t>(stream), phase == 0 ? BLOCK_PRODUCE : BLOCK_CONSUME);
}
int ic_dit_kv_segment(ic_dit* h, int rank, void** ptr, long long* bytes) {
  if (!h || !ptr || !bytes || rank < 0 || rank >= h->c.world_size || h->n_groups != 1 || h->p2p.on) return IC_ERR_INVALID;
  *ptr = h->kv_all + static_cast<long long>(rank) * h->chunk_elems();
  *bytes = h->chunk_elems() * static_cast<long long>(sizeof(__nv_bfloat16));
  return IC_OK;
}
int ic_dit_missing_tensors(const ic_dit* h, char* names_host, int cap) {
  if (!h || cap < 0 || (cap > 0 && !names_host)) return IC_ERR_INVALID;
  int n = 0, used = 0;
  if (cap > 0) names_host[0] = 0;
  for (const auto& kv : h->slots) {
    if (kv.second.loaded) continue;
    ++n;
    const int len = static_cast<int>(kv.first.size());
    if (used + len + 2 <= cap) {
      memcpy(names_host + used, kv.first.c_str(), len);
      used += len;
      names_host[used++] = '\n';
      names_host[used] = 0;
    }
  }
  return n;
}
int ic_dit_head(ic_dit* h, float* head_out, void* stream) {
  if (!h || !head_out) return IC_ERR_INVALID;
  return head(h, head_out, static_cast<cudaStream_t>(stream));
}
float* ic_dit_tokens(ic_dit* h) { return h ? h->x : nullptr; }

int ic_dit_forward(ic_dit* h, const float* latents, float timestep, int ctx_slot, float* head_out, void* stream) {
  if (!h || !latents || !head_out || ctx_slot < 0 || ctx_slot > 1) return IC_ERR_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!h->weights_checked) {  // never run on cudaMalloc'd weights nobody wrote
    for (const auto& kv : h->slots)
      if (!kv.second.loaded && !(kv.first.rfind("buffer_embedder.", 0) == 0 && !h->has_guide)) {
        fprintf(stderr, "[icb] ic_dit_forward: tensor %s was never loaded\n", kv.first.c_str());
        return IC_ERR_INVALID;
      }
    h->weights_checked = true;
  }
  h->launches = 0;
  IC_TRY(embed(h, latents, timestep, st));
  for (int i = 0; i < h->L; ++i) IC_TRY(run_block(h, i, ctx_slot, st));
  return head(h, head_out, st);
}

long long ic_dit_flops_per_forward(const ic_dit* h) {
  if (!h) return 0;
  const double N = h->Sall, D = h->D, F = h->F, L = h->c.text_len;
  const double per_layer = 8 * N * D * D + 4 * N * N * D + 4 * N * D * D + 4 * L * D * D + 4 * N * L * D + 4 * N * D * F;
  const double fwd = h->L * per_layer + 2 * N * h->CK * D + 2 * N * D * 4 * h->c.out_dim;
  return static_cast<long long>(fwd);
}
int ic_dit_launch_count(const ic_dit* h) { return h ? h->launches : 0; }

int ic_dit_set_profiling(ic_dit* h, int enable) {
  if (!h) return IC_ERR_INVALID;
  h->profiling = enable < 0 ? 0 : enable;
  h->prof_used = 0;
  return IC_OK;
}

int ic_dit_profile_collect(ic_dit* h, float* ms_by_kind, int* count_by_kind) {
  if (!h || !ms_by_kind || !count_by_kind) return IC_ERR_INVALID;
  ICB_CUDA_CHECK(cudaDeviceSynchronize());
  for (int k = 0; k < PROF_KINDS; ++k) {
    ms_by_kind[k] = 0.f;
    count_by_kind[k] = 0;
  }
  for (size_t i = 0; i < h->prof_used; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->prof[i].a, h->prof[i].b) == cudaSuccess) {
      ms_by_kind[h->prof[i].kind] += ms;
      count_by_kind[h->prof[i].kind] += 1;
    }
  }
  h->prof_used = 0;
  return IC_OK;
}

}  // extern "C"
