/*
 * infinicube_b200 — C ABI of the B200-native (sm_100a) implementation of InfiniCube's
 * video-generation hot path.  Plain C: opaque handles, raw device pointers, explicit cudaStream_t
 * (passed as void*), int return codes (0 = ok, negative = error; no exceptions cross the ABI).
 *
 * Every entry point names the reference interface it replaces (paths relative to the reference
 * checkout of nv-tlabs/InfiniCube).  The arithmetic of the DiT lives in the reference's un-vendored
 * `diffsynth` dependency (pyproject.toml:71); the call sites below are the reference's own.
 *
 * All pointers are DEVICE pointers unless the parameter name ends in `_host`.  The library never
 * allocates on behalf of a granular op; engine handles own their workspaces.
 */
#ifndef INFINICUBE_B200_H_
#define INFINICUBE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IC_OK 0
#define IC_ERR_INVALID (-1)
#define IC_ERR_CUDA (-2)
#define IC_ERR_NO_DEVICE (-3)
#define IC_ERR_UNSUPPORTED (-4)
#define IC_ERR_NCCL (-5)

#define IC_DTYPE_F32 0
#define IC_DTYPE_BF16 1

/* library / device introspection */
int ic_version(void);
const char* ic_error_string(int code);
int ic_device_check(void); /* IC_OK iff the current device is sm_100 (B200) */

/* ------------------------------------------------------------------------------------------------
 * Granular tensor-core ops
 * ---------------------------------------------------------------------------------------------- */

/* C = A[M,K] * B[N,K]^T with the fused epilogue of every nn.Linear in the Wan2.1 DiT block
 * (reference: diffsynth WanModel via infinicube/videogen/inference.py:216-226).
 *   v = acc + bias ; v = gelu_tanh(v) if act==1
 *   out_bf16 = bf16(v) ; rowss[r, n_tile] = sum bf16(v)^2 ; out_f32 = v + addend ; resid += gate * v
 * Null pointers disable the corresponding output. */
typedef struct {
  const float* bias;
  int bias_per_row;
  int act;
  void* out_bf16;
  int ld_out;
  float* rowss;
  int rowss_ld;
  float* out_f32;
  int ld_f32;
  const float* addend;
  int ld_add;
  float* resid;
  int ld_res;
  const float* gate;
} ic_gemm_epilogue;

int ic_gemm_bf16(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const ic_gemm_epilogue* ep,
                 void* stream);
int ic_gemm_block_n(int N);

/* Non-causal attention forward, head_dim 128 (reference: diffsynth flash_attention() used by
 * SelfAttention/CrossAttention of WanModel).
 * Q [Sq, ldq], K [n_seg][seg_len, ldk], VT [n_seg][n_heads*128, ldvt] (keys contiguous), O [Sq, ldo]; bf16. */
int ic_fmha_fwd(const void* Q, int ldq, const void* K, int ldk, long long k_seg_stride, const void* VT, int ldvt,
                long long vt_seg_stride, void* O, int ldo, int Sq, int seg_len, int n_seg, int n_heads,
                float softmax_scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Granular HBM-bound DiT ops
 * ---------------------------------------------------------------------------------------------- */
int ic_ln_modulate(const float* x, int ldx, const float* mul, const float* add, int mul_plus_one, void* out_bf16,
                   int ldo, int rows, int D, float eps, void* stream);
/* rope tables: (cos,sin) fp32 pairs tab_f [n_f][22], tab_h [n_h][21], tab_w [n_w][21]; pass NULLs for no RoPE */
int ic_rmsnorm_rope(const void* src_bf16, int ld_src, const float* rowss, int ss_ld, int ss_off, int ss_cnt,
                    const float* weight, void* dst_bf16, int ld_dst, int rows, int D, float eps, const float* tab_f,
                    const float* tab_h, const float* tab_w, int n_f, int n_h, int n_w, int frame0, void* stream);
int ic_patchify(const float* latents, void* out_bf16, int C, int F, int H, int W, int ld_out, int col_off,
                void* stream);
/* latents += (v_neg + cfg*(v_pos - v_neg)) * dsigma   (FlowMatchScheduler.step + CFG of WanVideoPipeline.__call__) */
int ic_unpatchify_cfg_step(float* latents, const float* head_pos, const float* head_neg, int C, int F, int H, int W,
                           float cfg_scale, float dsigma, float* v_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * DiT engine: the whole WanModel.forward behind one handle
 * (replaces `self.pipe.dit` as driven by infinicube/videogen/inference.py:216-226).
 * ---------------------------------------------------------------------------------------------- */
typedef struct ic_dit ic_dit;

typedef struct {
  int dim;         /* 1536 (1.3B) / 5120 (14B) */
  int ffn_dim;     /* 8960 / 13824 */
  int num_heads;   /* 12 / 40, head_dim fixed 128 */
  int num_layers;  /* 30 / 40 */
  int in_dim;      /* 16 latent channels */
  int out_dim;     /* 16 */
  int text_dim;    /* 4096 */
  int freq_dim;    /* 256 */
  int text_len;    /* 512 */
  int guide_channels; /* channels of the concatenated guidance latents (2 x buffer_channels), 0 = none */
  float eps;       /* 1e-6 */
  int lat_f, lat_h, lat_w; /* global latent grid, e.g. 24 x 60 x 104 */
  int frame0;        /* first latent frame owned by this rank */
  int frames_local;  /* latent frames owned by this rank (token shard = frames_local*(lat_h/2)*(lat_w/2)) */
  int world_size;    /* ranks sharing the token axis (1 = no collective) */
  int rank;
} ic_dit_config;

int ic_dit_create(const ic_dit_config* cfg, ic_dit** out);
int ic_dit_destroy(ic_dit* h);
long long ic_dit_workspace_bytes(const ic_dit* h);

/* Copy one state-dict tensor (official Wan2.1 key names, `dit.` prefix stripped; plus
 * `buffer_embedder.weight|bias`) into engine-owned storage.  Mirrors
 * WanVideoGenerator._load_checkpoint (infinicube/videogen/inference.py:101-128).
 * Returns IC_ERR_INVALID for an unknown key or a size mismatch. */
int ic_dit_load_tensor(ic_dit* h, const char* name, const void* src, int dtype, long long numel, void* stream);

/* Multi-GPU: attach an NCCL communicator for the per-layer (K || V^T) all-gather.
 * unique_id_host: the 128-byte ncclUniqueId produced by ic_nccl_unique_id on rank 0. */
int ic_nccl_unique_id(void* unique_id_host_128B);
int ic_dit_init_comm(ic_dit* h, const void* unique_id_host_128B);

/* Peer-memory alternative to the NCCL all-gather (opt-in; one process per GPU on ONE NVSwitch node): every rank
 * pushes its (K || V^T) segment into its peers' gather buffers with the copy engines and raises a per-segment epoch
 * flag; the attention kernel starts on the local segment and waits per remote segment, so the exchange overlaps the
 * attention.  Protocol: ic_dit_p2p_export on every rank (128 bytes: two cudaIpcMemHandle_t), all-gather the blobs in
 * rank order through the host, ic_dit_p2p_attach(world x 128 bytes) on every rank, then a barrier before the first
 * forward.  Needs world_size <= 16, one head group, and stream memory operations (driver). */
int ic_dit_p2p_export(ic_dit* h, void* handles_host_128B);
int ic_dit_p2p_attach(ic_dit* h, const void* all_handles_host);
int ic_dit_p2p_enabled(const ic_dit* h);

/* Text context (post umT5): ctx [text_len, text_dim]; runs text_embedding and caches the per-layer
 * cross-attention K / V^T.  slot 0 = prompt, 1 = negative prompt. */
int ic_dit_set_context(ic_dit* h, int slot, const void* ctx, int dtype, void* stream);

/* Guidance-buffer token injection: g = buffer_embedder(concat(sem_latents, coord_latents)); added to the
 * patch-embedded tokens in every forward (README.md:65, infinicube/videogen/inference.py:86-88).
 * guide_latents: fp32 [guide_channels, frames_local, lat_h, lat_w]; NULL clears the guidance. */
int ic_dit_set_guidance(ic_dit* h, const float* guide_latents, void* stream);

/* One WanModel.forward: latents fp32 [in_dim, frames_local, lat_h, lat_w], timestep in [0,1000],
 * ctx_slot selects the cached context; head_out fp32 [tokens_local, 4*out_dim] (patch layout,
 * column = (py*2+px)*out_dim + c). */
int ic_dit_forward(ic_dit* h, const float* latents, float timestep, int ctx_slot, float* head_out, void* stream);

/* Debug / parity hooks: run the pre-block stage only or a single block on the engine's token buffer. */
int ic_dit_embed(ic_dit* h, const float* latents, float timestep, void* stream);
int ic_dit_run_block(ic_dit* h, int layer, int ctx_slot, void* stream);
int ic_dit_head(ic_dit* h, float* head_out, void* stream);
/* Parity hook for the token shard: a block split at the K / V^T exchange so that a test can run the
 * ranks of one forward as several engines on ONE device and stand in for the all-gather with plain copies.
 * phase 0: LN+modulation, QK / V^T GEMMs, RMSNorm+RoPE -> q and this rank's (K || V^T) segment; phase 1:
 * self-attention over all segments (no collective is issued) and the rest of the block.
 * ic_dit_kv_segment: address and size of rank `rank`'s segment inside this engine's gather buffer. */
int ic_dit_run_block_phase(ic_dit* h, int layer, int ctx_slot, int phase, void* stream);
int ic_dit_kv_segment(ic_dit* h, int rank, void** ptr, long long* bytes);
/* Names (newline-separated, NUL-terminated) of registered tensors that ic_dit_load_tensor has not filled yet;
 * returns the number of missing tensors (the text is truncated to cap bytes, the count is not), < 0 on error.
 * WanVideoGenerator._load_checkpoint loads `dit.` non-strictly (videogen/inference.py:121-128) but an engine whose
 * weights were never written must not run. */
int ic_dit_missing_tensors(const ic_dit* h, char* names_host, int cap);
float* ic_dit_tokens(ic_dit* h); /* fp32 [tokens_local, dim] residual stream */
long long ic_dit_flops_per_forward(const ic_dit* h); /* algorithmic FLOPs, SURVEY §8(d) formula, global */
int ic_dit_launch_count(const ic_dit* h);            /* kernels launched by the last forward */
/* In-stream CUDA-event timing of the engine's own launches (measurement only; events are recorded on the
 * launching stream).  Kinds: 0 = self-attention FMHA, 1 = cross-attention FMHA, 2 = GEMM; `kind_mask` bit k enables
 * kind k (0 = off, 7 = all) - every bracketed launch costs two event records on the stream, so a timed region should
 * enable only the kind it reports.  ic_dit_profile_collect synchronises the device, sums elapsed ms / launch counts per
 * kind and resets. */
int ic_dit_set_profiling(ic_dit* h, int kind_mask);
int ic_dit_profile_collect(ic_dit* h, float* ms_by_kind3_host, int* count_by_kind3_host);

/* ------------------------------------------------------------------------------------------------
 * Voxel -> guidance-buffer rasteriser
 * ---------------------------------------------------------------------------------------------- */
typedef struct ic_grid ic_grid;

/* points_to_fvdb (infinicube/utils/fvdb_utils.py:71-216): ijk = round((p - origin)/voxel_size),
 * unique voxels, per-voxel arg-max-count label (ties -> smallest label) for `semantics` and `instance`.
 * points: device fp32 [m,3]; sem / inst: device int32 [m] (either may be NULL -> zeros).
 * Synchronises the stream twice (bounding box, voxel count). */
int ic_grid_build(const float* points, long long m, const float* voxel_size_host3, const float* origin_host3,
                  const int* sem, const int* inst, ic_grid** out, void* stream);
int ic_grid_destroy(ic_grid* g);
long long ic_grid_num_voxels(const ic_grid* g); /* GridBatch.total_voxels */
long long ic_grid_num_bricks(const ic_grid* g);
int ic_grid_info(const ic_grid* g, int* imin3_host, int* imax3_host, int* brick_min3_host, int* brick_dim3_host);
/* GridBatch.ijk.jdata in this grid's voxel-index order + the per-voxel labels (device int32) */
int ic_grid_export(const ic_grid* g, int* ijk, int* sem, int* inst, void* stream);

/* fvdb.gridbatch_from_mesh as used for the CAD car model (infinicube/utils/fvdb_utils.py:219-296): marks, in a
 * dense bit mask over the box [ijk_min, ijk_min + dims), every voxel whose cube overlaps a triangle
 * (separating-axis test in fp64).  verts: device fp64 [nv,3]; faces: device int32 [nf,3];
 * mask: device uint32 [(dims.x*dims.y*dims.z + 31)/32], bit index = ((k*dims.y + j)*dims.x + i). */
int ic_mesh_voxelize_mask(const double* verts, int nv, const int* faces, int nf, double voxel_size, double origin,
                          const int* ijk_min_host3, const int* dims_host3, unsigned int* mask, void* stream);

/* CameraBase.get_zdepth_map_from_voxel + 2 x get_semantic_map_from_voxel (infinicube/camera/base.py:520-618)
 * for n_cam poses in ONE launch.  kinv_host9: row-major inverse intrinsics (host); poses: device fp32
 * [n_cam,16] row-major camera->grid (OpenCV axes).  attr0 / attr1: optional device int32 [n_voxels]
 * per-voxel attributes in voxel-index order replacing the grid's own semantic / instance labels (the
 * `voxel_semantic` argument of get_semantic_map_from_voxel); background0/1 = value written on a miss.
 * Outputs [n_cam,H,W]: depth fp32 (0 = miss), attr0 image int32, attr1 image int32. */
int ic_raster_render(const ic_grid* g, const float* kinv_host9, const float* poses, int n_cam, int W, int H,
                     const int* attr0, const int* attr1, int background0, int background1, float* depth, int* sem,
                     int* inst, void* stream);

/* semantic_to_color -> uint8 (truncation) + instance overlay (infinicube/utils/semantic_utils.py:88-131,
 * inference/guidance_buffer_generation.py:690-698).  Base colour = palette[sem] (palette: device uint8
 * [n_classes,3]) or, when base_rgb != NULL, the given uint8 [n,3] image (generate_rgb_semantic_buffer's
 * `semantics_rgb`); pixels with inst > 0 take inst_colors[k] where inst_ids_sorted[k] == inst
 * (device int32 ascending / device uint8 [n_ids,3]).  inst may be NULL. */
int ic_semantic_rgb(const int* sem, const unsigned char* base_rgb, const int* inst, long long n,
                    const unsigned char* palette, int n_classes, const int* inst_ids_sorted,
                    const unsigned char* inst_colors, int n_ids, unsigned char* rgb, void* stream);
/* out[i,:] = lut[idx[i],:] with a float32 [n_rows,3] table (semantic_to_color, utils/semantic_utils.py:88-101) */
int ic_lut_gather_f32(const int* idx, long long n, const float* lut, int n_rows, float* out, void* stream);

/* get_instance_id_for_fvdb_scene_points (infinicube/utils/fvdb_utils.py:299-385): points device fp32 [n,3] (world),
 * sem device int32 [n]; boxes device fp32 [n_boxes,16] = rows 0-2 of world_to_object (12 floats), the half extents
 * lwh/2*enlarge (3 floats) and object_id_int stored bit-for-bit in the 16th float, in the dict order of
 * "000000.static_object_info.json" (a later box overwrites an earlier one).  A point whose class bit is set in
 * car_class_mask (bit c = class c, c < 32) and that lies inside a box (|local| <= half on every axis) gets that id,
 * everything else 0. */
int ic_instance_from_boxes(const float* points, long long n, const int* sem, const float* boxes, int n_boxes,
                           unsigned int car_class_mask, int* instance_id, void* stream);

/* unproject_depth_torch into the first camera's frame (infinicube/utils/depth_utils.py:402-466,
 * utils/buffer_utils.py:205-226): xyz [n_cam,H,W,3]; pixels with depth == 0 get the 1e7 sentinel. */
int ic_coord_unproject(const float* depth, const float* cam_to_cam0, const float* kinv9, int n_cam, int H, int W,
                       float* xyz, void* stream);
/* global-quantile normalisation to [0,1] (+ uint8 truncation) (utils/buffer_utils.py:246-262,
 * guidance_buffer_generation.py:710); mins3 / ranges3 are device fp32[3]. */
int ic_coord_normalize(const float* xyz, const float* depth, long long n_pixels, const float* mins3,
                       const float* ranges3, float* out_f32, unsigned char* out_u8, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Wan 3-D VAE ops: channels-last bf16 activations [T, H, W, C].
 * Replace the cuDNN conv3d/conv2d + PyTorch elementwise calls of diffsynth's WanVideoVAE, which the reference
 * drives through `self.pipe(...)` (infinicube/videogen/inference.py:216-226) for the two buffer encodes and the
 * final decode.
 * ---------------------------------------------------------------------------------------------- */
/* Implicit-GEMM convolution on tcgen05: out[t,h,w,:] = bias + sum_taps W[tap] in[t+dt,h+dh,w+dw,:] (+ resid);
 * out-of-range input reads are zero (spatial padding and causal temporal padding).  taps_host: ntaps x (dt,dh,dw);
 * weight bf16 [Cout, ntaps*Cin] with K index = tap*Cin + c; Cin % 32 == 0, Cout % 8 == 0. */
int ic_conv_cl(const void* in, int Tin, int Hin, int Win, int Cin, const void* weight, const float* bias,
               const int* taps_host, int ntaps, void* out, int T, int H, int W, int Cout, int ld_out, const void* resid,
               int ld_resid, void* stream);
/* RMS_norm over channels (x / max(|x|_2, 1e-12) * sqrt(C) * gamma) with optional SiLU */
int ic_rmsnorm_cl(const void* in, const float* gamma, void* out, long long npix, int C, int apply_silu, void* stream);
/* nearest-exact 2x spatial upsample */
int ic_upsample2x_cl(const void* in, void* out, int T, int H, int W, int C, void* stream);
/* Resample(upsample3d): out[0] = x0, out[1+2t+j] = y[t][..., j*C:(j+1)*C] for the 2C-channel time_conv output y */
int ic_time_interleave_cl(const void* x0, const void* y, void* out, int T1, long long HW, int C, void* stream);
/* Resample(downsample3d) operand: out[k] = concat_c(x[2k], x[2k+1], x[2k+2]) */
int ic_time_gather3_cl(const void* x, void* out, int K, long long HW, int C, void* stream);
/* space-to-depth (2x2 phases into channels) so the stride-2 3x3 conv becomes a stride-1 2x2-cell conv */
int ic_space_to_depth_cl(const void* in, void* out, int T, int H, int W, int C, void* stream);
/* row softmax of fp32 scores -> bf16 probabilities (VAE mid-block attention) */
int ic_softmax_rows(const float* s, int ld_s, void* p_bf16, int ld_p, int nrows, int ncols, float scale, void* stream);
/* uint8 frames [T,H,W,3] -> bf16 [T,H,W,Cpad] in [-1,1] (pixel*(2/255) - 1) */
int ic_frames_to_cl(const unsigned char* frames, void* out, long long npix, int Cpad, void* stream);
/* normalised latents fp32 [C,T,h,w] -> z*std + mean as bf16 [T,h,w,Cpad] */
int ic_latent_to_cl(const float* z, const float* mean, const float* stdv, void* out, int C, long long thw, int Cpad,
                    void* stream);
/* bf16 channels-last (first C channels, row pitch ld) -> fp32 channels-first, optional (x - mean)/std */
int ic_cl_to_cf(const void* in, int ld, const float* mean, const float* stdv, float* out, int C, long long npix,
                void* stream);
/* DiffSynth tiled encode/decode blending: values[c,t,h0+y,w0+x] += tile * ramp mask, weight += mask; then
 * values / weight (-> clamp -> uint8 frames [T,H,W,3] via ((x+1)*127.5).clip(0,255)).  bound_mask bits: 1 top,
 * 2 bottom, 4 left, 8 right edge tiles (no ramp on image borders). */
int ic_blend_accumulate(const void* tile, int ld, int C, int T, int th, int tw, float* values, float* weight, int H, int W,
                        int h0, int w0, int bound_mask, int border_h, int border_w, void* stream);
int ic_blend_finalize(const float* values, const float* weight, int C, int T, int H, int W, int clamp, float* out_f32,
                      unsigned char* frames, void* stream);

/* ------------------------------------------------------------------------------------------------
 * umT5-XXL prompt encoder pieces.  Replaces the text encoder that the reference
 * runs inside diffsynth's WanPrompter / WanTextEncoder when infinicube/videogen/inference.py:216-226
 * passes `prompt` / `negative_prompt` to the pipeline (model file named at :63-81).  The nn.Linear of
 * each T5 block run on ic_gemm_bf16; these are the kernels between them.
 * ---------------------------------------------------------------------------------------------- */
/* x[r, :] = float(table[ids[r], :]); ids int32 [L] on the device, table bf16 [vocab, D] */
int ic_t5_embed(const int* ids, int L, const void* table_bf16, int vocab, int D, float* x, int ldx, void* stream);
/* T5LayerNorm: out = bf16(weight * x * rsqrt(mean(x^2) + eps)); rows >= zero_from_row are written as zeros
 * (WanPrompter zeroes the rows past the prompt length); zero_from_row < 0 disables */
int ic_t5_rmsnorm(const float* x, int ldx, const float* weight, void* out_bf16, int ldo, int rows, int D, float eps,
                  int zero_from_row, void* stream);
/* self-attention of one block, head_dim 64, no 1/sqrt(d): q/k/v bf16 [L, ld] (head h = columns [64h, 64h+64)),
 * bias_by_offset fp32 [n_heads][2L-1] = relative-position bias indexed by (key - query + L - 1),
 * key_mask uint8 [L] (0 = padding key, NULL = none) */
int ic_t5_attention(const void* q, const void* k, const void* v, int ld, const float* bias_by_offset,
                    const unsigned char* key_mask, void* out, int ldo, int L, int n_heads, void* stream);
/* out = a * b elementwise on bf16 (gated-GELU product fc1(x) * gelu(gate(x))); n % 8 == 0 */
int ic_mul_bf16(const void* a, const void* b, void* out, long long n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Exact nearest-neighbour search / label transfer.  Replaces knn_query_fast(queries, ref, 1)
 * (infinicube/voxelgen/ext/common/knn.cu:15-50, KD-tree of kdtree_cuda.cu) as used by semantic_from_points
 * (infinicube/voxelgen/utils/color_util.py:52-60) in the stage-1 chunk merge
 * (infinicube/inference/voxel_generation_single_chunk.py:280, voxelgen/utils/extrap_util.py:272).
 * d2 = ((qx-px)^2 + (qy-py)^2) + (qz-pz)^2 in fp32 without FMA; ties -> smallest reference index.
 * ---------------------------------------------------------------------------------------------- */
typedef struct ic_knn ic_knn;
/* ref_xyz: device fp32, `stride` floats per point (>= 3); cell_size <= 0 picks one from the bounding box.
 * Allocates the index on the current device and synchronises `stream` once (bounding box read-back). */
int ic_knn_build(const float* ref_xyz, long long m, int stride, float cell_size, ic_knn** out, void* stream);
int ic_knn_destroy(ic_knn* k);
int ic_knn_info(const ic_knn* k, long long* n_points, long long* n_cells, float* cell_size, int* dims3_host);
/* out_idx int32 [n] (knn_query_fast's indices), out_d2 fp32 [n] (its squared distances),
 * out_labels int64 [n] = ref_labels[idx] (semantic_from_points fused); any output may be NULL. */
int ic_knn_query1(const ic_knn* k, const float* queries, long long n, int stride, const long long* ref_labels,
                  int* out_idx, float* out_d2, long long* out_labels, void* stream);
/* knn_query_fast(queries, ref, nb_points) for nb_points in [1, 32] (voxelgen/ext/common/knn.cu:15-51; the k = 8 use is
 * color_from_points, voxelgen/utils/color_util.py:21-49): out_idx int32 [n, nb_points], out_d2 fp32 [n, nb_points],
 * each row ascending by (squared distance, reference index); slots beyond the number of reference points hold
 * -1 / +inf.  Either output may be NULL. */
int ic_knn_query(const ic_knn* k, const float* queries, long long n, int stride, int nb_points, int* out_idx,
                 float* out_d2, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* INFINICUBE_B200_H_ */
