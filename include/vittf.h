/* libvittf_b200 -- C ABI of the B200-native vit-tf feature-volume hot path.
 *
 * The reference (xeTaiz/vit-tf) is pure Python and has no FFI layer of its own; the
 * boundary it exposes for this path is a set of Python functions.  Each
 * entry point below names the reference interface (file:line under /root/reference) whose
 * arithmetic it replaces; the modules in vittf_b200/ keep the Python signatures and call these through
 * ctypes (see INTEGRATION.md for the binding a maintainer would add to the reference).
 *
 * Conventions
 *   - every function returns VITTF_OK (0) or a negative vittf_status; the message of the
 *     last failure on the calling thread is available from vittf_last_error();
 *   - all pointers are DEVICE pointers owned by the caller (torch allocator) unless the
 *     parameter name ends in _host; nothing is allocated behind the caller's back except
 *     inside an explicit context object (vittf_vit_create / vittf_bls_create);
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it;
 *   - volumes are C-contiguous (X, Y, Z) with Z fastest, feature volumes (F, fX, fY, fZ),
 *     exactly the layouts of the reference's tensors.
 */
#ifndef VITTF_H
#define VITTF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    VITTF_OK = 0,
    VITTF_ERR_INVALID = -1, /* bad argument / unsupported shape */
    VITTF_ERR_CUDA = -2,    /* a CUDA runtime/driver call failed */
    VITTF_ERR_NOMEM = -3,   /* workspace too small */
    VITTF_ERR_STATE = -4    /* context used before it was fully initialised */
} vittf_status;

typedef enum { VITTF_U8 = 0, VITTF_F16 = 1, VITTF_BF16 = 2, VITTF_F32 = 3, VITTF_F64 = 4 } vittf_dtype;

const char* vittf_last_error(void);
int vittf_version(void);
/* compute capability of the current device, major*10+minor (100 on B200) */
int vittf_device_arch(int* out_arch);
/* number of kernels this library has launched in this process since the last reset
 * (bench.py reports it as `gpu_launches`) */
int64_t vittf_launch_count(void);
void vittf_launch_count_reset(void);

/* =====================================================================================
 * Stage 1 -- ViT K-feature extraction (infer.py:130-210 compute_qkv, :314-340 main loop)
 * ===================================================================================== */

/* Global intensity range of the volume: norm_minmax, infer.py:32-34 (used at :155).
 * out2 = {min, max} as float32. */
int vittf_minmax(const void* vol, int64_t n, int dtype, float* out2, void* stream);

typedef struct {
    int embed_dim;   /* D: 384 (ViT-S) / 768 (ViT-B) */
    int depth;       /* L: number of blocks of the hub model (12); the last block only
                        contributes norm1 + the K third of attn.qkv */
    int num_heads;   /* 6 / 12, head dim must be 64 */
    int patch;       /* 8 or 16 */
    int mlp_hidden;  /* 4*D */
} vittf_vit_config;

/* Device pointers to the weights of one transformer block, PyTorch Linear layout
 * (out_features, in_features), bf16 matrices, fp32 vectors. */
typedef struct {
    const float* ln1_w; const float* ln1_b;
    const void* qkv_w;  const float* qkv_b;   /* (3D, D) */
    const void* proj_w; const float* proj_b;  /* (D, D)  */
    const float* ln2_w; const float* ln2_b;
    const void* fc1_w;  const float* fc1_b;   /* (4D, D) */
    const void* fc2_w;  const float* fc2_b;   /* (D, 4D) */
    /* LayerNorm folded into the projections that follow it (vittf_gemm_bf16_ln).  Both NULL: qkv_w / fc1_w are the plain
     * weights and norm1 / norm2 run as vittf_layernorm.  Both set: qkv_w = bf16(W_qkv * ln1_w), qkv_b = b + W_qkv ln1_b,
     * qkv_colsum fp32 (3D) = row sums of that bf16 matrix; likewise fc1_* with ln2_*; ln*_w / ln*_b are then unused. */
    const float* qkv_colsum;
    const float* fc1_colsum;
} vittf_block_weights;

typedef struct vittf_vit vittf_vit; /* opaque */

/* Creates the engine: replaces `model = dino_model_fn(...).to(dev).eval()` (infer.py:323)
 * as far as the hooked output is concerned.
 *   patch_w   fp32 (p*p, D), tap-major: patch-embed conv weight with the 3 identical grey
 *             channels and the ImageNet mean/std affine folded in,
 *             patch_b fp32 (D)
 *   max_batch largest number of slices per forward; workspace is caller-provided.       */
int vittf_vit_create(vittf_vit** out, const vittf_vit_config* cfg, const vittf_block_weights* blocks_host,
                     const float* patch_w, const float* patch_b, int max_batch, int max_tokens);
void vittf_vit_destroy(vittf_vit* v);
/* bytes of scratch the engine needs for `batch` images of `tokens` tokens (incl. CLS) */
int64_t vittf_vit_workspace_bytes(const vittf_vit* v, int batch, int tokens);

/* One forward over slices [s0, s1) of `vol` along `axis` (0='x',1='y',2='z'):
 * slicing + min-max + NN resize (infer.py:137-155,177) -> patch embed + cls/pos -> L-1 blocks
 * -> norm1 + K projection of the last block -> fp16 K features of the PATCH tokens,
 * out_k (s1-s0, f0*f1, D) with D fastest (what infer.py:201-202 extracts from the hook).
 *   pos_embed fp32 (1+f0*f1, D) already interpolated for this image size (hub
 *             interpolate_pos_encoding), row 0 = cls_token + pos[0].
 *   im0, im1  network input size for this axis (image_sizes, infer.py:143-147).           */
int vittf_vit_k_features(vittf_vit* v, const void* vol, int vol_dtype, int X, int Y, int Z, int axis, int s0, int s1,
                         int im0, int im1, const float* minmax2, const float* pos_embed, void* out_k_f16,
                         void* workspace, int64_t workspace_bytes, void* stream);

/* Optional device-side timing of the kernels launched inside vittf_vit_k_features: cudaEvent pairs
 * on the caller's stream around every attention (kind 0) and GEMM (kind 1) launch.  bench.py uses
 * it for the roofline of the dominant kernel.  _read synchronises, returns and clears the totals. */
int vittf_vit_timing_enable(vittf_vit* v, int enable);
int vittf_vit_timing_read(vittf_vit* v, double* ms_by_kind2, int64_t* launches_by_kind2);

/* AdaptiveAvgPool3d along the slice axis only + permute to the reference's (D,fX,fY,fZ)
 * layout (infer.py:203 with pool_fn from :329): k (S, f0*f1, D) fp16 -> out fp16.
 * n_out == S reproduces the `_noop` pool of single-axis runs (infer.py:326).
 * accumulate != 0 adds (in fp16, rounding like infer.py:332) into `out` instead of storing.
 * Sharding: S is the GLOBAL slice count, k holds slices [slice0, slice0+n_local)
 * and only the output slabs [o0, o1) are produced (their windows must lie inside k's range).
 * compact != 0: `out` is this rank's block only, i.e. its slab axis has extent o1 - o0 (what the all-gather sends). */
int vittf_pool_axis(const void* k_f16, int S, int slice0, int n_local, int f0, int f1, int D, int axis, int n_out,
                    int o0, int o1, void* out_f16, int accumulate, int compact, void* stream);
/* Multi-GPU merge of one slicing axis: staging fp16 (world, D, e0, e1, e2) = the all-gathered compact blocks of the ranks
 * ((e0,e1,e2) = (fX,fY,fZ) with the slab axis -- 0 x, 1 y, 2 z -- divided by world) -> out fp16 (D,fX,fY,fZ), assigned
 * (accumulate == 0, the first axis) or added in fp16 like infer.py:332. */
int vittf_accumulate_gathered_f16(void* out_f16, const void* staging_f16, int world, int D, int fX, int fY, int fZ, int axis,
                                  int accumulate, void* stream);

/* out = fp16(out + in): the running sum over the three slicing axes (infer.py:332), used when
 * the per-axis volumes come from different GPUs. */
int vittf_accumulate_f16(void* out_f16, const void* in_f16, int64_t n, void* stream);

/* -------- individually exported building blocks (unit-tested against torch) ---------- */
enum { VITTF_EPI_BIAS_BF16 = 0, VITTF_EPI_BIAS_GELU_BF16 = 1, VITTF_EPI_BIAS_RESID_F32 = 2,
       VITTF_EPI_QKV_SPLIT = 3, VITTF_EPI_KFEAT_F16 = 4, VITTF_EPI_BIAS_RESID_LN = 5 };
/* C[M,N] = A[M,K] (bf16, row-major) x W[N,K]^T (bf16) + bias, tcgen05/TMEM/TMA.
 *   epi 0/1: out bf16 (M,N) [GELU(erf) for 1];  epi 2: out fp32 (M,N) += (residual stream);
 *   epi 3: N = 3D, out = qk bf16 (M, 2D), out2 = V^T bf16 (B*heads*64, tok_pad) with
 *          tokens/tok_pad describing the (image, token) split of M;
 *   epi 4: out fp16 ((M/tokens)*(tokens-1), N): CLS rows dropped (K features).             */
int vittf_gemm_bf16(const void* A, const void* W, const float* bias, void* out, void* out2, int M, int N, int K,
                    int epi, int tokens, int tok_pad, void* stream);
/* The same GEMM with the LayerNorm of the pre-LN blocks (hub Block.forward: x + attn(norm1(x)), x + mlp(norm2(x)),
 * called from infer.py:177) folded into the linear layers around it, so that no LayerNorm pass runs:
 *   LN(x) W^T + b = rstd * (x (W*gamma)^T - mean * colsum(W*gamma)) + (b + W beta).
 * The fp32 residual stream lives in a ROW-TILED layout xt[m_pad/32][D/4][32 rows][4 columns] (m_pad = M rounded up to 256;
 * element (r, c) at (((r/32)*(D/4) + c/4)*32 + r%32)*4 + c%4), together with its raw bf16 copy xb (M, D) row-major and
 * per-row partial sums stats[m_pad][VITTF_LN_SLOTS] = (sum, sum of squares) over the columns of a slot (unused slots: 0).
 *   consumer (colsum != NULL; epi 0, 1, 3, 4): A = xb, W = bf16(W*gamma), bias = b + W beta, `stats` describes the rows
 *            of x (K = D columns); the epilogue applies mean / rstd (eps) per row.
 *   producer (epi 5, VITTF_EPI_BIAS_RESID_LN): xt += A W^T + bias (read-modify-write of the row-tiled stream by the epilogue
 *            threads, one row each), out = xb bf16 (M, N) of the updated stream, stats_out slots [0, vittf_gemm_ln_slots(N))
 *            (one per 128 columns for N % 256 == 0, per 64 otherwise; N must need <= VITTF_LN_SLOTS of them).
 * vittf_ln_prepare builds xt / xb / stats (slot 0 = full sums, the other slots zero) from a row-major stream. */
#define VITTF_LN_SLOTS 8
typedef struct {
    const float* colsum;
    const float* stats;      /* (m_pad, VITTF_LN_SLOTS, 2) */
    float eps;
    float* xt;
    float* stats_out;        /* (m_pad, VITTF_LN_SLOTS, 2) */
    int64_t m_pad;
} vittf_ln_fold;
int vittf_gemm_ln_slots(int N);
int vittf_gemm_bf16_ln(const void* A, const void* W, const float* bias, void* out, void* out2, int M, int N, int K,
                       int epi, int tokens, int tok_pad, const vittf_ln_fold* ln, void* stream);
int vittf_ln_prepare(const float* x, float* xt, void* xb_bf16, float* stats, int64_t rows, int64_t m_pad, int D,
                     void* stream);
/* softmax(Q K^T / 8) V per (image, head); qk (B*tokens, 2D) bf16, vt as written by epi 3,
 * out bf16 (B*tokens, D). */
int vittf_attention(const void* qk, const void* vt, void* out, int B, int tokens, int tok_pad, int heads, void* stream);
/* Same operator for q ALREADY multiplied by hd^-0.5 * log2(e) (the engine folds the factor into the Q rows of the qkv
 * projection, vittf_b200/vit.py): a max-free first pass (p = 2^score, no running maximum) and a second, online-softmax pass
 * over the CTAs whose rows left the safe exponent range (flags in `workspace`, vittf_attention_workspace_bytes).          */
int64_t vittf_attention_workspace_bytes(int B, int tokens, int heads);
int vittf_attention_prescaled(const void* qk, const void* vt, void* out, int B, int tokens, int tok_pad, int heads,
                              void* workspace, int64_t workspace_bytes, void* stream);
/* LayerNorm(eps=1e-6) over the last dim: x fp32 (rows, D) -> y bf16 */
int vittf_layernorm(const float* x, const float* w, const float* b, void* y_bf16, int64_t rows, int D, void* stream);
/* slices -> tokens fp32 (B, 1+f0*f1, D) incl. cls/pos (see vittf_vit_k_features) */
int vittf_patch_embed(const void* vol, int vol_dtype, int X, int Y, int Z, int axis, int s0, int s1, int im0, int im1,
                      int patch, int D, const float* minmax2, const float* patch_w, const float* patch_b,
                      const float* pos_embed, float* out_tokens, void* stream);

/* =====================================================================================
 * Stage 2 -- prototype similarity (predict_ntf.py:24-101, old/cluster_dino.py:306-345)
 * ===================================================================================== */

/* sample_features3d (infer.py:48-72): grid_sample with align_corners=False, zero padding.
 *   feats (F,w,h,d) fp16|fp32, rel (A,3) fp32 in [-1,1] (X,Y,Z order), mode 0 nearest /
 *   1 trilinear ('bilinear'), out fp32 (A,F). */
int vittf_sample_prototypes(const void* feats, int feat_dtype, int F, int w, int h, int d, const float* rel, int A,
                            int mode, float* out, void* stream);

/* Low-resolution pass: reads the feature volume ONCE and produces
 *   dots  fp32 (A, n_lr)   <f_v, p_a>         (einsum 'fwhd,caf->cawhd', predict_ntf.py:65)
 *   gram  fp32 (14, n_lr)  <f_v, f_{v+o}> for o in {0} U 13 forward neighbours (may be
 *                             NULL): lets the up-sampling pass evaluate |interp(f)|^2
 *                             without materialising interp(f).
 *   dots_layout 0: dots fp32 (A, n_lr).  1: dots fp32 (n_lr, A4) with A4 = A rounded up to a multiple of 4 (voxel-major: the
 *   tcgen05 up-sampling kernel gathers the corner dots of 4 prototypes with one 16-byte load); only the fused tensor-core pass
 *   writes it -- vittf_sim_lowres_layout() tells whether it is available for an input (1) or not (0).
 */
int vittf_sim_lowres_layout(int feat_dtype, int F, int w, int h, int d, const void* feats);
/*   [xa, xb): low-res x planes to evaluate (slab sharding, SURVEY.md 8e: a rank only needs the planes under its output
 *   slab +- 1; the other planes of dots / gram are left untouched).  The generic fallback kernels evaluate every plane. */
int vittf_sim_lowres(const void* feats, int feat_dtype, int F, int w, int h, int d, const float* protos, int A,
                     float* dots, float* gram, int dots_layout, int xa, int xb, void* stream);

typedef enum {
    VITTF_SIM_NS = 0,     /* interp(features) -> L2 normalise -> dot -> clamp(0,1)^e -> class MAX */
    VITTF_SIM_REFNTF = 1, /* raw dot -> where(>=thr)^e -> class MEAN           (predict_ntf.py:71-72) */
    VITTF_SIM_LEGACY = 2, /* normalise at low res -> dot -> clamp(0,1)^e -> MAX (cluster_dino.py:307-322) */
    VITTF_SIM_CLAMP_MEAN = 3 /* raw dot -> clamp(0,1)^e -> group MEAN               (infer.py:104-106, resample_topk) */
} vittf_sim_mode;

/* Second pass: per OUTPUT voxel combine the 8 surrounding low-res dots (trilinear,
 * align_corners=False index rule of F.interpolate), normalise, non-linearity, per-class
 * reduction.  Output slab [x0,x1) x [z0,z1) of the (W,H,D) grid only (slab sharding, §8e): z-slabs as in north_star, x-slabs
 * (the slowest axis: every class map of a rank is one contiguous block, and pass 1 shards with it) or both:
 * out fp32 (C, x1-x0, H, z1-z0).  For REFNTF/LEGACY the output grid equals the low-res grid.
 *   class_offsets int32 (C+1) prefix offsets into the A prototypes.                        */
int vittf_sim_upsample(const float* dots, const float* gram, int w, int h, int d, int A, const int* class_offsets,
                       int C, int W, int H, int D, int x0, int x1, int z0, int z1, int mode, float threshold, float exponent,
                       int dots_layout, float* out, void* stream);

/* 0.99*max quantisation input (predict_ntf.py:95): per-class maximum, out fp32 (C) */
int vittf_class_max(const float* sims, int C, int64_t n, float* out, void* stream);
/* uint8 maps as compute_similarities returns them (predict_ntf.py:95-100): per class (255 / (0.99*class_max[c]) * sim) cast
 * to uint8 with the C-style wrap of the reference's CPU cast, then F.interpolate(mode='nearest') to (Wo,Ho,Do).
 * sims fp32 (C, W, H, z1-z0) is the z-slab [z0,z1) of a (W,H,D) grid; out uint8 (C, Wo, Ho, zo1-zo0) holds the output planes
 * [zo0,zo1), whose source planes min(floor(oz*D/Do), D-1) must lie inside the slab.  class_max: device, fp32 (C). */
int vittf_quantize_maps_u8(const float* sims, int C, int W, int H, int D, int z0, int z1, const float* class_max,
                           int Wo, int Ho, int Do, int zo0, int zo1, uint8_t* out, void* stream);
/* label composition (predict_ntf.py:203-215): thresholded running arg-max with strict '>'
 * on uint8 maps; thresholds_u8[i] = int(thr_i*255).  mode 1 = plain argmax(0) of float maps
 * (old/cluster_dino.py:345), writing int32 in that case is avoided: labels are uint8.      */
int vittf_labels(const void* sims, int sims_dtype, int C, int64_t n, const int* thresholds_u8, int mode,
                 uint8_t* out, void* stream);

/* =====================================================================================
 * Stage 3 -- 3-D bilateral solver (bilateral_solver3d.py:211-245)
 * ===================================================================================== */
typedef struct {
    int W, H, D;             /* shape of t / r[0] (crop-local) */
    double sigma_spatial;    /* grid_params, bilateral_solver3d.py:156-160 */
    double lam, A_diag_min, cg_tol; /* bs_params, :162-167 */
    int cg_maxiter;
    int luma_bins;           /* max(lut)+1 */
} vittf_bls_params;

/* bytes of fp64 grid scratch needed */
int64_t vittf_bls_workspace_bytes(const vittf_bls_params* p, int nrhs);
/* Solves `nrhs` targets that share one reference volume.
 *   t fp32 (nrhs,W,H,D) targets; r_u8 (W,H,D) grey reference; conf fp32 (W,H,D) or NULL for
 *   the Sobel confidence (:176-181,233-237); luma_lut int32[256] from the reference's exact
 *   numpy expression; out fp32 (nrhs,W,H,D); iters_out int32 (nrhs) PCG
 *   iterations actually run (device pointer, may be NULL).                                   */
int vittf_bls_solve(const vittf_bls_params* p, const float* t, const uint8_t* r_u8, const float* conf,
                    const int* luma_lut, int nrhs, float* out, int* iters_out, void* workspace,
                    int64_t workspace_bytes, void* stream);
/* ---- the same solver in stages, for z-slab sharding over several GPUs ----
 * Per-voxel arrays are slab-local (.., W, H, z1-z0); r_u8 is always the full (W,H,D) reference.
 *   rank-local : vittf_bls_sobel_slab  -> raw Sobel magnitude of the slab, atomic max into *c_max
 *   exchange   : all-reduce(max) of c_max                       (skipped when a confidence is given)
 *   rank-local : vittf_bls_splat_slab  -> acc[cell][m, wbar, b_0..b_{nrhs-1}] +=, ncell*(2+nrhs) fp64, cell-major
 *   exchange   : all-reduce(sum) of acc
 *   replicated : vittf_bls_grid_solve  -> y (ncell*nrhs fp64, cell-major): vertex compaction, bistochastisation + PCG on the
 *                                         occupied cells
 *   rank-local : vittf_bls_slice_slab  -> out fp32 slab
 * vittf_bls_solve is exactly this sequence with one slab.                                          */
int64_t vittf_bls_grid_cells(const vittf_bls_params* p);
int64_t vittf_bls_grid_workspace_bytes(const vittf_bls_params* p, int nrhs);
int vittf_bls_sobel_slab(const uint8_t* r_u8, int W, int H, int D, int z0, int z1, float* c_raw_slab, float* c_max,
                         void* stream);
/* conf_slab: the confidence itself (c_max == NULL) or the raw Sobel magnitude with the global maximum in *c_max */
int vittf_bls_splat_slab(const vittf_bls_params* p, const float* t_slab, const uint8_t* r_u8, const float* conf_slab,
                         const float* c_max, const int* luma_lut, int nrhs, int z0, int z1, double* acc, void* stream);
int vittf_bls_grid_solve(const vittf_bls_params* p, int nrhs, const double* acc, double* y, int* iters_out,
                         void* workspace, int64_t workspace_bytes, void* stream);
int vittf_bls_slice_slab(const vittf_bls_params* p, const uint8_t* r_u8, const int* luma_lut, const double* y, int nrhs,
                         int z0, int z1, float* out_slab, void* stream);
/* Sobel confidence alone: out fp32 (W,H,D) = max(c) - c (needs a float scratch of 1 elem) */
int vittf_sobel_confidence(const uint8_t* r_u8, int W, int H, int D, float* out, float* scratch_max, void* stream);

/* =====================================================================================
 * Annotation samplers (compare_feat_sampling.py:13-33) -- SURVEY.md 8f row 1
 * ===================================================================================== */
/* resample_topk (infer.py:90-94), per similarity map (n_maps x n fp32): thr = K-th largest value, out_idx = the first K
 * flat voxel indices in index order with value >= thr (the reference's `(s >= thr).nonzero()[:K]`); out_thr may be NULL. */
int vittf_topk_voxels(const float* maps, int n_maps, int64_t n, int K, long long* out_idx, float* out_thr, void* stream);
/* take_most_dissimilar (infer.py:118-121): out[i] = 1 - mean_j cos(f_i, f_j) (measure 0) or mean_j |f_i - f_j| (1) */
int vittf_mean_pairwise_distance(const float* feats, int N, int F, int measure, float* out, void* stream);
/* scipy.ndimage.binary_erosion(mask, generate_binary_structure(3, connectivity)) with border_value 0:
 * mask / out uint8 (W,H,D), non-zero = foreground; connectivity >= 3 is the full 3x3x3 box (:20-23). */
int vittf_binary_erosion(const uint8_t* mask, int W, int H, int D, int connectivity, uint8_t* out, void* stream);

/* =====================================================================================
 * Evaluation (evaluate_similarities.py:58-83) -- SURVEY.md 8f row 3
 * ===================================================================================== */
/* sklearn.metrics.confusion_matrix of two uint8 label volumes (n voxels, values < K <= 16): out[t * K + p] = number of
 * voxels with true label t and predicted label p (K*K uint64, zeroed by the call); out_of_range = voxels whose labels
 * were >= K (counted nowhere).  precision / recall / F1 / Jaccard / accuracy (:66-71) derive from this table. */
int vittf_confusion_matrix(const uint8_t* truth, const uint8_t* pred, int64_t n, int K, unsigned long long* out,
                           unsigned int* out_of_range, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VITTF_H */
