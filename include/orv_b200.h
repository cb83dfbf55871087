/* orv_b200.h — C ABI of liborv_b200.so, the B200 (sm_100a) implementation of ORV's denoising hot path.
 *
 * The reference has no FFI on this path: the boundary is the Python class
 * `CogVideoXTransformer3DModelTraj.forward` (reference orv/models/cogvideox_control.py:715-948) called once per
 * scheduler step from `CogVideoXImageToVideoPipelineTraj.__call__` (:1402-1473).  This header is what a binding
 * for that call would bind: plain pointers and sizes, no torch types.  The Python host
 * (`orv_b200/`) mirrors the reference's module API on top of it through ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - tensors are contiguous row-major; `bf16` = __nv_bfloat16 bits; weights keep torch's nn.Linear layout
 *     [out_features, in_features];
 *   - the library never allocates or frees caller tensors: activations live in a caller-provided workspace;
 *   - kernels are enqueued on the `stream` argument (a cudaStream_t passed as void*), never synchronise, and are
 *     CUDA-graph capturable;
 *   - every function returns ORVB_OK (0) or a negative error code; orvb_last_error() returns the thread-local
 *     message.  No C++ exception crosses this boundary.
 */
#ifndef ORV_B200_H_
#define ORV_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORVB_VERSION 104

enum {
  ORVB_OK = 0,
  ORVB_EINVAL = -1, /* bad argument (null pointer, unsupported option)            */
  ORVB_ESHAPE = -2, /* shape / alignment the kernels do not support               */
  ORVB_ECUDA = -3,  /* a CUDA runtime / driver call failed                        */
  ORVB_EARCH = -4,  /* the current device is not compute capability 10.x (B200)   */
  ORVB_ENOMEM = -5  /* caller workspace too small                                 */
};

int orvb_version(void);
const char* orvb_last_error(void);
/* ORVB_OK when the current CUDA device can run the sm_100a kernels. */
int orvb_check_device(void);

/* ------------------------------------------------------------------------------------------------------------
 * Granular operators (unit-testable pieces of the path)
 * ------------------------------------------------------------------------------------------------------------ */

/* Epilogue selector of orvb_gemm_bf16. */
enum {
  ORVB_EPI_BIAS = 0,      /* out = acc + bias                                                     */
  ORVB_EPI_GELU = 1,      /* out = gelu_tanh(acc + bias)           (FeedForward net.0, ref a9)    */
  ORVB_EPI_GATE_RESID = 2,/* out = resid + gate[g(row)] * (acc + bias)   (gated residual, ref a6) */
  ORVB_EPI_QKV = 3        /* out = acc + bias, then per-head LayerNorm(64) (+RoPE) on the Q and K
                             thirds of the row (attention processor, ref a8)                      */
};

/* Row -> modulation-group map of a joint [text | video] sequence (reference LayerNormZero semantics,
 * cogvideox_control.py:117-145): with s = row % seq_len and b = row / seq_len,
 *   group(row) = b * groups_per_batch + (s < text_len ? 0 : 1 + (s - text_len) / tokens_per_group).
 * Group 0 of every batch is the text group (modulated by the time embedding alone).  tokens_per_group <= 0 sends
 * the video rows to group 0 as well (no per-frame action modulation: MVBlock.norm1, :323-325).
 * seq_len == 0 disables the map (group = 0, every row counts as video). */
typedef struct orvb_rowmap {
  int32_t seq_len;
  int32_t text_len;
  int32_t tokens_per_group;
  int32_t groups_per_batch;
} orvb_rowmap;

typedef struct orvb_gemm_args {
  /* out[M,N] = epilogue(A[M,K] @ W[N,K]^T); A row pitch lda, W row pitch ldw (elements). */
  const void* a;     /* bf16 [M, lda]  */
  const void* w;     /* bf16 [N, ldw]  */
  void* out;         /* bf16 [*, ldo]  */
  const void* bias;  /* bf16 [N] or NULL */
  int32_t m, n, k;
  int32_t lda, ldw, ldo;
  int32_t epilogue;  /* ORVB_EPI_* */

  /* Output-row remap: out_row = (row / src_rows) * dst_rows + dst_offset + row % src_rows.
   * src_rows == 0 -> identity. Used to scatter per-batch blocks into the joint sequence. */
  int32_t src_rows, dst_rows, dst_offset;
  /* Multiview scatter (MVBlock, cogvideox_control.py:345): when mv_tokens > 0 the GEMM rows are in '(b f)(v s)' order
   * and out_row = (b*mv_views + v) * dst_rows + dst_offset + f*mv_tokens + i. */
  int32_t mv_tokens, mv_frames, mv_views;

  /* GATE_RESID: resid may alias out.  resid row = resid_mod > 0 ? (row % resid_mod) + resid_view_stride *
   * ((row / resid_mod) % resid_views) : out_row (a positional table broadcast over the batch when resid_mod>0).
   * gate (fp32, row pitch gate_ld) may be NULL (= 1.0).  For a row of group g the gate vector starts at
   * gate + g * gate_ld + (g % groups_per_batch == 0 ? gate_text_off : gate_video_off). */
  const void* resid; /* bf16 [*, ldr] */
  int32_t ldr;
  int32_t resid_mod, resid_views, resid_view_stride;
  const float* gate;
  int32_t gate_ld, gate_text_off, gate_video_off;
  orvb_rowmap rowmap;

  /* QKV: n == 3 * qk_dim; columns [0, 2*qk_dim) get LayerNorm over each 64-wide head with the q / k affine
   * parameters (bf16 [64]); rope_cos/sin (fp32 [video_tokens, 64], may be NULL) rotate rows whose position in
   * the sequence is >= rowmap.text_len (diffusers apply_rotary_emb, interleaved pairs). */
  int32_t qk_dim;
  const void* q_norm_w; const void* q_norm_b;
  const void* k_norm_w; const void* k_norm_b;
  float qk_eps;
  const float* rope_cos; const float* rope_sin;
  /* Tight-tolerance test mode: `out` is fp32 [*, ldo] (ldo in fp32 elements) and receives the epilogue's fp32 values
   * before the bf16 rounding of the product path (same mainloop, same epilogue arithmetic, direct stores).  With
   * bf16-exact operands the result must match an fp32 reference to rtol 1e-3 / atol 1e-4 (tests/test_gpu_tight.py). */
  int32_t out_f32;
  /* Optional DEVICE scalar added to every modulation-group index (the row of `gate`) — lets one captured launch
   * sequence walk through the per-step slices of a modulation schedule (orvb_forward_args.schedule).  NULL = 0. */
  const int32_t* group_offset;
  /* k_wrap > 0: W has only k_wrap columns and is walked cyclically along K (column kk of the contraction reads
   * W[:, kk % k_wrap]); k % k_wrap == 0, k_wrap % 64 == 0.  With A = [hi | lo] (an fp32 operand split into two bf16
   * halves, k = 2 * k_wrap) the product is the fp32-accurate A_fp32 @ W^T — the AdaLN table build uses this. */
  int32_t k_wrap;
} orvb_gemm_args;

/* tcgen05 / TMA GEMM.  Requirements: k % 8 == 0, n % 8 == 0, lda/ldw/ldo % 8 == 0, 16-byte aligned bases. */
int orvb_gemm_bf16(const orvb_gemm_args* args, void* stream);
/* Tile choice of orvb_gemm_bf16 for an [m, n] output (host arithmetic only): > 0 = single-CTA kernel, 128 x value tiles;
 * < 0 = CTA-pair kernel (tcgen05 cta_group::2), 256 x (-value) tiles, the width that fills the last wave of the
 * device's SM pairs best. */
int orvb_gemm_tile_width(int32_t m, int32_t n, int32_t epilogue);
/* CTA-pair kernel only: when cutting N into n / width full tiles plus ONE narrower tile per 256-row block balances the SM
 * pairs better than any uniform width (QKV of a 2B block: 22 x 256 + 128 instead of 30 x 192), the width of that narrower
 * tile; 0 = uniform tiles. */
int orvb_gemm_tile_remainder(int32_t m, int32_t n, int32_t epilogue);
/* The tile list orvb_gemm_bf16 runs for an [m, n] x k problem on the CTA-pair kernel, in the order the SM pairs walk it
 * (host arithmetic only): up to `capacity` records {sm_pair, first_row, first_column, width} of 4 x int32.  Returns the
 * number of tiles (0: single-CTA kernel) or a negative error code.  in_place_resid: a GATE_RESID call written over its own
 * residual.  Used by the host-side coverage test of the mixed tile list. */
int orvb_gemm_tile_list(int32_t m, int32_t n, int32_t k, int32_t epilogue, int32_t in_place_resid, int32_t* out,
                        int32_t capacity);
/* Test hook: orvb_gemm_bf16 with a forced tile.  bn in {64, 128, 192, 256} = single-CTA kernel with that N tile; -bn
 * (a multiple of 16 in 32..256) = CTA-pair kernel.  Used by the bit-identity test of the two kernels and by the tile
 * sweeps under tools/. */
int orvb_gemm_bf16_bn(const orvb_gemm_args* args, int bn, void* stream);
/* Non-causal multi-head attention over a packed QKV buffer (reference a8: F.scaled_dot_product_attention).
 * qkv: bf16 [batch * seq_len, 3 * heads * 64] with Q | K | V column blocks; out: bf16 [batch * seq_len, heads*64].
 * softmax scale = scale (1/sqrt(64) in the reference).  head_dim is fixed at 64 (every shipped config). */
int orvb_attention_bf16(const void* qkv, void* out, int32_t batch, int32_t seq_len, int32_t heads, float scale,
                        void* stream);
/* Full form.  The query window [q_row0, q_row0 + q_rows) of every sequence produces a COMPACT output
 * [batch * q_rows, heads*64] (MVBlock keeps only the video rows, cogvideox_control.py:333); q_rows <= 0 = all rows.
 * out_f32: tight-tolerance test mode, `out` is fp32 (the normalised accumulator before the bf16 rounding). */
typedef struct orvb_attention_args {
  const void* qkv; void* out;
  int32_t batch, seq_len, heads;
  float scale;
  int32_t q_row0, q_rows;
  int32_t out_f32;
} orvb_attention_args;
int orvb_attention(const orvb_attention_args* args, void* stream);
/* Test hooks of the attention kernel.  The running row max is raised lazily (only when a key tile exceeds it by more
 * than 2^threshold, default 8), which makes the rescale of the TMEM-resident output accumulators rare;
 * orvb_attention_set_rescale_threshold(0) forces that path on almost every tile (negative = restore the default).
 * orvb_attention_set_debug installs a device buffer for in-kernel clock stamps (builds with -DORVB_ATT_TIMELINE). */
void orvb_attention_set_rescale_threshold(float log2_units);
void orvb_attention_set_debug(void* dev_buf);

/* LayerNorm(eps, affine) followed by AdaLN modulation y = LN(x) * (1 + scale_g) + shift_g where g is the row's
 * group and (shift, scale) come from the fp32 table `mod` (row pitch mod_ld):
 *   group 0 (text):  shift = mod[g] + text_off,  scale = shift + dim
 *   video groups:    shift = mod[g] + video_off, scale = shift + dim      (`scale_first` swaps the two)
 * mod == NULL -> plain LayerNorm.  Output rows may be remapped like the GEMM (src_rows/dst_rows/dst_offset apply
 * to the OUTPUT row; with in_skip_text != 0 input row r reads x row (r / (seq-text)) * seq + text + r % (seq-text)).
 * Reference: CogVideoXLayerNormZero / AdaLayerNorm, cogvideox_control.py:41-197. */
typedef struct orvb_ln_args {
  const void* x; void* y;       /* bf16 [rows, dim] */
  const void* ln_w; const void* ln_b; /* bf16 [dim] or NULL */
  int32_t rows, dim;
  float eps;
  const float* mod; int32_t mod_ld, text_off, video_off, scale_first;
  orvb_rowmap rowmap;
  int32_t in_video_only;        /* gather only video rows of the joint sequence (norm_final / norm_out) */
  const void* pre_w; const void* pre_b; /* optional first LayerNorm (norm_final, cogvideox_control.py:909-916)
                                           applied before the modulated one; bf16 [dim] or NULL */
  float pre_eps;
  /* Optional pre-combined table (bf16, row pitch ab_ld >= 4*dim): per group row
   * [text A | text B | video A | video B] with A = ln_w*(1+scale), B = ln_b*(1+scale)+shift.  When set, ln_w/ln_b/mod
   * are ignored and y = xhat*A + B (what orvb_forward uses: one table build per forward instead of four vector
   * loads per element). */
  const void* ab; int32_t ab_ld;
  int32_t y_f32;                /* tight-tolerance test mode: y is fp32 [rows, dim] (values before the bf16 rounding) */
  const int32_t* group_offset;  /* optional DEVICE scalar added to every group index (row of mod / ab); NULL = 0 */
} orvb_ln_args;
int orvb_ln_modulate(const orvb_ln_args* args, void* stream);

/* y[r, n] = act(sum_k x[r,k] * W[n,k] + b[n]) for a handful of rows (AdaLN tables, time/ofs/action MLPs).
 * x, y fp32; W, b bf16.  act: 0 none, 1 silu, 2 gelu_tanh.  rows <= 64. */
int orvb_skinny_linear(const float* x, const void* w, const void* b, float* y, int32_t rows, int32_t n, int32_t k,
                       int32_t act, void* stream);

/* Patch gather (im2col of CogVideoXPatchEmbed):
 *   patch_t == 0: out[b, (f,i,j), c*p*p + kh*p + kw]           = x[b, f, c, i*p+kh, j*p+kw]
 *   patch_t  > 0: out[b, (f',i,j), c*pt*p*p + t*p*p + kh*p + kw] = x[b, f'*pt+t, c, i*p+kh, j*p+kw]
 * x: bf16 [B, F, C, H, W]; out: bf16 [B * F/pt * H/p * W/p, C*pt*p*p]. */
int orvb_patchify(const void* x, void* out, int32_t b, int32_t f, int32_t c, int32_t h, int32_t w, int32_t p,
                  int32_t patch_t, void* stream);
/* Inverse map for the model output (cogvideox_control.py:926-936).
 * y: bf16 [B * F/pt * H/p * W/p, C*pt*p*p] -> out: bf16 [B, F, C, H, W]. */
int orvb_unpatchify(const void* y, void* out, int32_t b, int32_t f, int32_t c, int32_t h, int32_t w, int32_t p,
                    int32_t patch_t, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Whole-model forward (reference a2: CogVideoXTransformer3DModelTraj.forward)
 * ------------------------------------------------------------------------------------------------------------ */

typedef struct orvb_config {
  int32_t dim;              /* num_attention_heads * attention_head_dim                     */
  int32_t heads;
  int32_t head_dim;         /* must be 64                                                   */
  int32_t layers;
  int32_t ff_dim;           /* 4 * dim                                                      */
  int32_t time_embed_dim;   /* 512                                                          */
  int32_t text_embed_dim;   /* 4096                                                         */
  int32_t in_channels;      /* 32 (latent 16 + image-condition 16)                          */
  int32_t out_channels;     /* 16                                                           */
  int32_t patch_size;       /* 2                                                            */
  int32_t patch_size_t;     /* 0 = None (2B family), 2 for CogVideoX1.5                     */
  int32_t use_rope;         /* use_rotary_positional_embeddings                             */
  int32_t has_ofs;          /* ofs_embed_dim is not None                                    */
  int32_t ofs_embed_dim;
  int32_t flip_sin_to_cos;
  float   freq_shift;
  float   norm_eps;
  int32_t visual_guidance;  /* initial_combine_linear present                               */
  int32_t num_control_keys;
  int32_t multiview;        /* mv_blocks present                                            */
  int32_t max_n_view;
  int32_t action_state_dim; /* 7                                                            */
  int32_t action_compress;  /* 4                                                            */
  int32_t action_hidden;    /* 4 * time_embed_dim                                           */
  int32_t modulate_text;    /* modulate_encoder_hidden_states: 1 = norm linears are [6D, T] and the text rows are
                               part of the joint sequence (every 2B / 5B ORV config); 0 = [3D, T], the text never
                               enters attention or the FFN (cogvideox_control.py:70-99, :404-424, the from-scratch
                               1.4B configs): the forward runs on the video rows alone, shape.text_len must be 0 */
} orvb_config;

/* Borrowed device pointers (bf16), one struct per CogVideoXBlock / MVBlock.  to_q/to_k/to_v are passed FUSED:
 * qkv_w = cat([to_q.weight, to_k.weight, to_v.weight], 0). */
typedef struct orvb_block_weights {
  const void* norm1_lin_w; const void* norm1_lin_b;   /* [6D, T], [6D] */
  const void* norm1_ln_w;  const void* norm1_ln_b;    /* [D] */
  const void* qkv_w;       const void* qkv_b;         /* [3D, D], [3D] */
  const void* q_norm_w;    const void* q_norm_b;      /* [64] */
  const void* k_norm_w;    const void* k_norm_b;      /* [64] */
  const void* out_w;       const void* out_b;         /* [D, D], [D] */
  const void* norm2_lin_w; const void* norm2_lin_b;   /* [6D, T], [6D]  (NULL in MVBlock) */
  const void* norm2_ln_w;  const void* norm2_ln_b;
  const void* ff1_w;       const void* ff1_b;         /* [FF, D], [FF] (NULL in MVBlock) */
  const void* ff2_w;       const void* ff2_b;         /* [D, FF], [D]  (NULL in MVBlock) */
  const void* proj_out_w;  const void* proj_out_b;    /* MVBlock only: [D, D], [D] */
} orvb_block_weights;

typedef struct orvb_weights {
  const void* patch_w;  const void* patch_b;      /* [D, C*pt*p*p] (conv weight flattened), [D] or NULL */
  const void* text_w;   const void* text_b;       /* [D, text_dim], [D] */
  const void* pos_embed;                          /* bf16 [views or 1, video_tokens, D] sin-cos rows (+ view table when views > 1), or NULL */
  const void* pos_embed_plain;                    /* bf16 [video_tokens, D] without the view table (control latents); NULL = pos_embed */
  const void* time1_w;  const void* time1_b;      /* [T, D], [T] */
  const void* time2_w;  const void* time2_b;      /* [T, T], [T] */
  const void* ofs1_w;   const void* ofs1_b;       /* [T, ofs], [T] or NULL */
  const void* ofs2_w;   const void* ofs2_b;
  const void* act1_w;   const void* act1_b;       /* [4T, align8(28*pt)] (columns zero-padded to a multiple of 8), [4T] */
  const void* act2_w;   const void* act2_b;       /* [T, 4T], [T] */
  const void* act_mask_embed;                     /* [T] */
  const void* combine_w; const void* combine_b;   /* [D, keys*D], [D] or NULL */
  const void* norm_final_w; const void* norm_final_b; /* [D] */
  const void* norm_out_lin_w; const void* norm_out_lin_b; /* [2D, T], [2D] */
  const void* norm_out_ln_w;  const void* norm_out_ln_b;  /* [D] */
  const void* proj_out_w; const void* proj_out_b; /* [p*p*pt*out_ch, D] */
  const orvb_block_weights* blocks_host;          /* HOST array [layers] */
  const orvb_block_weights* mv_blocks_host;       /* HOST array [layers] or NULL */
} orvb_weights;

typedef struct orvb_model orvb_model;

int orvb_model_create(const orvb_config* cfg, orvb_model** out);
void orvb_model_destroy(orvb_model* m);
/* Records the (borrowed) weight pointers; the caller keeps the tensors alive. */
int orvb_model_bind_weights(orvb_model* m, const orvb_weights* w);

typedef struct orvb_shape {
  int32_t batch;        /* transformer batch (clips x CFG copies x views folded by the caller: B * V) */
  int32_t views;        /* V (1 = single view)                                                     */
  int32_t frames;       /* latent frames per view                                                  */
  int32_t height;       /* latent height (pixels / 8)                                              */
  int32_t width;
  int32_t text_len;     /* 226                                                                     */
  int32_t action_frames;/* rows of the action embedding per sample (= frames / max(patch_t,1)); 0 = none */
} orvb_shape;

size_t orvb_workspace_bytes(const orvb_model* m, const orvb_shape* s);

typedef struct orvb_forward_args {
  orvb_shape shape;
  const void* hidden_states;   /* bf16 [batch, frames, in_channels, height, width]  (views already folded) */
  const void* text;            /* bf16 [batch, text_len, text_embed_dim]                                   */
  const float* timesteps;      /* fp32 [batch]                                                             */
  float ofs;                   /* ofs scalar (2.0 in the reference pipeline) when cfg.has_ofs              */
  const void* actions;         /* bf16 [batch, action_frames, state*compress*max(pt,1)] pre-padded/reshaped
                                  (host does the integer padding of cogvideox_control.py:805-811), or NULL  */
  const uint8_t* action_mask;  /* u8 [batch] 1 = replace by mask_embed (drawn by the caller), or NULL       */
  const void* depths;          /* bf16 [batch, frames, in_channels, height, width] or NULL                 */
  const void* labels;          /* same, or NULL                                                            */
  const float* rope_cos;       /* fp32 [video_tokens, 64] or NULL                                          */
  const float* rope_sin;
  void* out;                   /* bf16 [batch, frames, out_channels, height, width]                        */
  void* workspace; size_t workspace_bytes;
  /* optional taps for parity tests (bf16 [batch*seq, dim] joint hidden state after block `tap_layer`) */
  void* tap_hidden; int32_t tap_layer;
  /* 1: the AdaLN tables of this step were installed into `workspace` by orvb_modulation_select; the time / action
   * embeddings and the table build (timesteps, ofs, actions, action_mask) are skipped. */
  int32_t skip_modulation;
  /* Step-invariant inputs.  The text projection and the patch embeddings of the control latents (depths / labels)
   * do not depend on the noisy latents, yet the reference recomputes them in each of the 50 forwards of a clip
   * (cogvideox_control.py:788, :827-846).  ORVB_STATIC_SAVE computes them and also keeps a copy in the workspace;
   * ORVB_STATIC_REUSE restores that copy instead of recomputing (same bits; `text`, `depths`, `labels` are then only
   * checked for presence).  The caller uses SAVE on the first step of a clip and REUSE afterwards, on the same
   * workspace. */
  int32_t static_mode;
  /* Modulation schedule read in place (preferred over orvb_modulation_select's copies: 22 MB per step for the 2B
   * model).  schedule = the buffer orvb_modulation_schedule filled for `schedule_steps` steps of this shape;
   * schedule_row_offset = DEVICE int32 holding step * (batch * (action_frames + 1)), written by the caller before each
   * forward (a 4-byte device write instead of the table copies; the launch sequence itself — and a CUDA graph captured
   * from it — is the same for every step).  Implies skip_modulation. */
  const void* schedule; int32_t schedule_steps; const int32_t* schedule_row_offset;
} orvb_forward_args;
enum { ORVB_STATIC_COMPUTE = 0, ORVB_STATIC_SAVE = 1, ORVB_STATIC_REUSE = 2 };

int orvb_forward(orvb_model* m, const orvb_forward_args* a, void* stream);

/* Modulation schedule.  Every AdaLN shift / scale / gate of the model depends on (timestep, ofs, actions) only —
 * never on the latents — yet the reference recomputes them inside each of the 50 forwards of a clip
 * (cogvideox_control.py:117-130, :166-170 through :762-779), re-reading 0.7 GB of AdaLN weights per step.  With the
 * timesteps of the sampler known up front the tables of ALL steps are built once per clip:
 *   orvb_modulation_schedule: timesteps fp32 [steps * batch] (step-major), actions bf16 [steps * batch, ...] and
 *                             action_mask u8 [steps * batch] laid out the same way (NULL as in orvb_forward);
 *                             `tables` = caller buffer of orvb_modulation_bytes() bytes, 256-byte aligned
 *   orvb_modulation_select:   copies step `step`'s tables into the forward workspace (two strided device copies);
 *                             the following orvb_forward on that workspace is called with skip_modulation = 1.
 * Same kernels, same arithmetic as the per-step path: results are bit-identical. */
size_t orvb_modulation_bytes(const orvb_model* m, const orvb_shape* shape, int32_t steps);
int orvb_modulation_schedule(orvb_model* m, const orvb_shape* shape, int32_t steps, const float* timesteps, float ofs,
                             const void* actions, const uint8_t* action_mask, void* tables, size_t tables_bytes,
                             void* stream);
int orvb_modulation_select(const orvb_model* m, const orvb_shape* shape, int32_t steps, int32_t step,
                           const void* tables, void* workspace, void* stream);

/* Number of kernel launches orvb_forward enqueued in its most recent call on this model (for bench.py's
 * `gpu_launches`). */
int orvb_last_launch_count(const orvb_model* m);
/* ORVB_PC_* class (below) of every kernel that call launched, in launch order: writes min(count, capacity) entries and
 * returns the count.  bench.py joins this list with the device-side activity records (CUPTI, through torch.profiler)
 * of the graph-replayed forwards to attribute device time per kernel class. */
int orvb_last_launch_classes(const orvb_model* m, int32_t* classes_out, int32_t capacity);

/* Optional per-kernel-class timing of orvb_forward (CUDA events recorded on the launch stream around every
 * launch; the forward then synchronises at its end, so enable it for measurement only).  orvb_model_set_profile
 * also clears the accumulators; orvb_model_get_profile copies accumulated milliseconds / launch counts per class
 * into arrays of ORVB_PROFILE_CLASSES entries. */
enum {
  ORVB_PC_PROLOGUE = 0, /* time/ofs/action MLPs + AdaLN tables (skinny linears)  */
  ORVB_PC_EMBED = 1,    /* patchify, patch/text/control projections              */
  ORVB_PC_LN = 2,       /* LayerNorm + modulate                                  */
  ORVB_PC_QKV = 3,      /* QKV GEMM (+QK-LN, RoPE)                               */
  ORVB_PC_ATTN = 4,     /* attention                                             */
  ORVB_PC_OUT = 5,      /* attention out-projection GEMM (+gate, residual)       */
  ORVB_PC_FF1 = 6,      /* FF up GEMM (+GELU)                                    */
  ORVB_PC_FF2 = 7,      /* FF down GEMM (+gate, residual)                        */
  ORVB_PC_HEAD = 8,     /* norm_final/norm_out, proj_out, unpatchify             */
  ORVB_PROFILE_CLASSES = 9
};
int orvb_model_set_profile(orvb_model* m, int enable);
int orvb_model_get_profile(const orvb_model* m, float* ms_out, int32_t* launches_out);

/* ------------------------------------------------------------------------------------------------------------
 * Sampler step (reference a14: CFG combine + CogVideoXDDIMScheduler/CogVideoXDPMScheduler.step + bf16 cast)
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct orvb_sampler_step_args {
  const void* model_out;   /* bf16 [cfg_copies * n]  (uncond first, then cond, as torch.cat([latents]*2)) */
  void* latents;           /* bf16 [n] in/out                                                            */
  float* old_x0;           /* fp32 [n] in/out (DPM only; written every step)                             */
  const void* noise;       /* bf16 [n] (DPM only; the reference draws it in the latent dtype) or NULL    */
  int64_t n;
  int32_t cfg_copies;      /* 1 or 2 */
  float guidance_scale;
  /* v = CFG(model_out);  x0 = c_x * x + c_v * v;  d = d_cur * x0 + d_old * old_x0;
   * x_prev = k_x * x + k_d * d + k_noise * noise.   (DDIM: d_cur = 1, d_old = 0, k_noise = 0.)
   * Products with the bf16 tensors x and noise are rounded to bf16 first, as torch does for
   * `float64_scalar * bf16_tensor`; no FMA contraction, so the result matches the torch op sequence bit for bit. */
  float c_x, c_v, d_cur, d_old, k_x, k_d, k_noise;
  /* Optional: also scatter the new latents into the next iteration's transformer input
   * bf16 [cfg_copies*B, F, lat_channels + img_channels, h, w] (channels [0, lat_channels) of every copy), which
   * replaces the reference's per-step torch.cat (cogvideox_control.py:1409-1413).  NULL = skip. */
  void* next_input;
  int32_t lat_channels, img_channels, hw;
} orvb_sampler_step_args;
int orvb_sampler_step(const orvb_sampler_step_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Occupancy voxelization (replaces the reference extension orv/ops/voxelize:
 * `voxelization_op.dynamic_voxelize_forward` / `hard_voxelize_forward`, voxelization.cpp:104-150, bound at
 * voxelization.py:89-116) and the label vote of its caller (orv/dataset/prepare_dataset.py:137-198).
 * Points are fp32 [n, c] with xyz in the first three features; cells are int32 (z, y, x) like the reference's
 * `coors`; grid_size = round((range_max - range_min) / voxel_size) per axis, computed in float as the reference does.
 * ------------------------------------------------------------------------------------------------------------ */

/* coors[i] = (z, y, x) of point i, or (-1, -1, -1) when it falls outside the range (or is NaN).  The reference's CPU
 * implementation writes exactly this (voxelization_cpu.cpp:18-42); its CUDA kernel leaves the later components of
 * an out-of-range point at their zero initialisation (voxelization_kernel.cuh:24-41) — callers test coors[:, 0]. */
int orvb_dynamic_voxelize(const float* points, int32_t n, int32_t c, const float* voxel_size /*[3]*/,
                          const float* coors_range /*[6]*/, int32_t* coors /*[n, 3]*/, void* stream);

typedef struct orvb_voxelize_args {
  const float* points;            /* fp32 [n, c], c >= 3 (c == 4 and a 16-byte aligned base take the 128-bit path) */
  int32_t n, c;
  float voxel_size[3];
  float coors_range[6];           /* x_min, y_min, z_min, x_max, y_max, z_max */
  int32_t max_points;             /* > 0: points kept per voxel (the first ones in index order)              */
  int32_t max_voxels;             /* > 0: voxels kept (the first ones in order of first appearance)          */
  float* voxels;                  /* fp32 [max_voxels, max_points, c], ZERO-FILLED by the caller as the reference's
                                     new_zeros (voxelization.py:97-98); only occupied slots are written; NULL = skip */
  int32_t* coors;                 /* int32 [max_voxels, 3] (z, y, x) or NULL                                  */
  int32_t* num_points_per_voxel;  /* int32 [max_voxels] or NULL                                               */
  int64_t* voxel_num;             /* device scalar: number of voxels produced (<= max_voxels); required       */
  /* Optional fused label vote of points_to_voxels (prepare_dataset.py:176-196): float64 [max_voxels, 4] rows
   * (x, y, z, label) where label + 1 is the most frequent LAST feature among the voxel's max_points slots, empty
   * slots counting as 0 and 0 yielding to the runner-up; ties go to the smaller label.  Real points must carry a
   * last feature > 0 (the caller adds 1 to its labels, prepare_dataset.py:160-161). */
  double* voxel_labels;
  void* workspace; size_t workspace_bytes;  /* orvb_voxelize_workspace_bytes(n, max_voxels), 256-byte aligned */
} orvb_voxelize_args;

size_t orvb_voxelize_workspace_bytes(int32_t n, int32_t max_voxels);
/* Deterministic hard voxelization: the result of the reference's deterministic path (voxelization_cpu.cpp:48-108,
 * voxelization_kernel.cu:24-129), which is also a valid outcome of its non-deterministic one. */
int orvb_hard_voxelize(const orvb_voxelize_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * 3-D Gaussian rasteriser, forward pass (second half; replaces `_C.rasterize_gaussians` of the
 * reference extension orv/ops/diff-gaussian-rasterization — rasterize_points.cu:35-150, ext.cpp:14-18 — as called by
 * orv/dataset/gs_render.py:103-171: occupancy voxels rendered as Gaussians into RGB, 12 semantic channels, depth and
 * alpha).  All pointers are device pointers, fp32; images are [C, H, W]; precomputed colours only (the caller passes
 * `colors_precomp`, shs=None); backward is out of scope.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct orvb_gs_args {
  int32_t p;                   /* number of Gaussians                                                            */
  const float* means3d;        /* [p, 3]                                                                         */
  const float* colors;         /* [p, 3]   colors_precomp                                                        */
  const float* features;       /* [p, 12]  language_feature_precomp, 16-byte aligned; NULL = include_feature off  */
  const float* opacities;      /* [p]                                                                            */
  const float* scales;         /* [p, 3]   with rotations, or NULL when cov3d is given                           */
  const float* rotations;      /* [p, 4]   quaternion (r, x, y, z), used as given (the reference does not normalise) */
  const float* cov3d;          /* [p, 6]   precomputed upper triangle, or NULL                                   */
  float scale_modifier;
  const float* viewmatrix;     /* [16] as the reference passes it: world-to-camera, transposed (column-major)     */
  const float* projmatrix;     /* [16] full projection, same convention                                          */
  const float* background;     /* [3]                                                                            */
  float tan_fovx, tan_fovy;
  int32_t height, width;
  float* out_color;            /* [3, H, W]                                                                      */
  float* out_feature;          /* [12, H, W] or NULL                                                             */
  float* out_depth;            /* [1, H, W]                                                                      */
  float* out_alpha;            /* [1, H, W]                                                                      */
  int32_t* radii;              /* [p] screen-space radius, 0 = culled                                            */
  int32_t* num_rendered;       /* device scalar (may be NULL): number of (Gaussian, tile) instances; NEGATIVE (minus
                                  the count) when it exceeded max_instances — the excess was dropped, call again with a
                                  larger capacity                                                                  */
  int32_t max_instances;       /* capacity of the binning buffers in the workspace                               */
  void* workspace; size_t workspace_bytes;  /* orvb_gs_workspace_bytes(p, max_instances, height, width), 256-byte aligned */
} orvb_gs_args;
size_t orvb_gs_workspace_bytes(int32_t p, int32_t max_instances, int32_t height, int32_t width);
/* One stream-ordered launch sequence, no host synchronisation (the reference reads the instance count back to size
 * its buffers, rasterizer_impl.cu:276-281). */
int orvb_gs_rasterize(const orvb_gs_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * 3-D VAE decode: the operators behind `AutoencoderKLCogVideoX.decode` (diffusers, called by the
 * reference pipeline at orv/models/cogvideox_control.py:1095-1100 and :1476-1479, tiled + sliced by
 * orv/pipeline/inference_control_to_video.py:98-99).  Activations are CHANNELS-LAST bf16 [T, H, W, C] for one sample
 * (the reference decodes one sample at a time: slicing); the frame-batch / tile / cache orchestration is host code
 * (orv_b200/models/autoencoder_kl_cogvideox.py).
 * ------------------------------------------------------------------------------------------------------------ */

/* Causal convolution as an implicit GEMM on tcgen05 (CogVideoXCausalConv3d with pad_mode "constant"; kt = 1 gives
 * the per-frame Conv2d of CogVideoXUpsample3D and the 1x1x1 shortcut / spatial-norm convolutions):
 *   out[t,h,w,n] = bias[n] + sum_{kt,kh,kw,c} X[t+kt-(KT-1), h+kh-KH/2, w+kw-KW/2, c] * W[n, ((kt*KH+kh)*KW+kw)*c_in + c]
 *                  (+ resid[t,h,w,n])
 * Spatial borders are zero padded; frames in front of the first one come from `cache` (the last KT-1 input frames of
 * the previous frame batch, [KT-1, H, W, c_in]) or, when cache is NULL, repeat frame 0. */
typedef struct orvb_conv_args {
  const void* x;      /* bf16 [frames, height, width, c_in]                                   */
  const void* cache;  /* bf16 [kt-1, height, width, c_in] or NULL                             */
  const void* w;      /* bf16 [c_out, kt*kh*kw*c_in], tap-major then channel (K-major rows)   */
  const void* bias;   /* bf16 [c_out] or NULL                                                 */
  const void* resid;  /* bf16 [frames, height, width, c_out] or NULL; may alias out           */
  void* out;          /* bf16 [frames, height, width, c_out] (fp32 when out_f32)              */
  int32_t frames, height, width;
  int32_t c_in;       /* multiple of 64 (zero-pad the channels)                               */
  int32_t c_out;      /* multiple of 8                                                        */
  int32_t kt, kh, kw; /* kt in 1..4; kh, kw odd, <= 7                                         */
  int32_t out_f32;    /* tight-tolerance test mode (see orvb_gemm_args.out_f32)               */
  /* Optional: GroupNorm statistics of the OUTPUT (as stored, i.e. rounded to bf16) accumulated in the epilogue, so the
   * SpatialNorm3D that follows needs no extra pass over the tensor: gn_stats[g] = (mean, rstd) fp32 for gn_groups
   * (<= 32) groups of c_out / gn_groups (a power of two in 2..64) consecutive channels; c_out % 64 == 0.  gn_scratch:
   * orvb_conv_gn_scratch_bytes() bytes, 16-byte aligned.  Deterministic (fixed tile schedule, fp64 fold in a fixed
   * order).  NULL = off. */
  float* gn_stats; void* gn_scratch; int32_t gn_groups; float gn_eps;
} orvb_conv_args;
int orvb_conv_cl(const orvb_conv_args* args, void* stream);
size_t orvb_conv_gn_scratch_bytes(void);

/* GroupNorm statistics of a channels-last tensor over ALL its pixels (one sample): stats[g] = (mean, rstd) fp32 for
 * the `groups` groups of channels/groups consecutive channels.  Deterministic: fixed-size pixel chunks are summed in
 * fp32, the chunk partials in fp64 in chunk order by the last block to finish.  `scratch` holds
 * orvb_gn_scratch_bytes(pixels, groups) bytes; its first 4 bytes must be zero before the first call (the kernel
 * leaves them zero). */
size_t orvb_gn_scratch_bytes(int64_t pixels, int32_t groups);
int orvb_gn_stats_cl(const void* x, int64_t pixels, int32_t channels, int32_t groups, float eps, float* stats,
                     void* scratch, void* stream);

/* CogVideoXSpatialNorm3D + SiLU, fused:  y = act( GroupNorm(x; stats, gamma, beta) * Y[src] + B[src] )  where Y / B are
 * the 1x1x1 convolutions conv_y / conv_b of the latent `zq`, evaluated ONCE per latent pixel into `table` (a GEMM over
 * the latent pixels) and looked up through the nearest-neighbour map of F.interpolate:
 *   src(t,h,w) = (t_src[t] * lat_h + (h >> shift)) * lat_w + (w >> shift)
 * (every spatial upsampling of the decoder is an exact factor 2; the temporal map — first frame kept single for odd
 * frame counts — comes from the host as t_src). */
typedef struct orvb_spatial_norm_args {
  const void* x; void* y;          /* bf16 [frames, height, width, channels]; y fp32 when y_f32          */
  int32_t frames, height, width, channels, groups;
  const float* stats;              /* [groups, 2] from orvb_gn_stats_cl                                 */
  const void* gamma; const void* beta; /* bf16 [channels]                                               */
  const void* table;               /* bf16 [lat_frames*lat_h*lat_w, table_ld]                           */
  int32_t table_ld, y_off, b_off;  /* column of conv_y / conv_b channel 0 (multiples of 8)              */
  const int32_t* t_src;            /* DEVICE [frames] latent frame of every frame                       */
  int32_t lat_h, lat_w, shift;
  int32_t act;                     /* 0 none, 1 SiLU                                                    */
  int32_t y_f32;
} orvb_spatial_norm_args;
int orvb_spatial_norm_cl(const orvb_spatial_norm_args* args, void* stream);

/* Nearest-neighbour x2 spatial upsampling with a temporal source map (CogVideoXUpsample3D before its convolution):
 * out[t, h, w, :] = x[t_src[t], h >> 1, w >> 1, :];  x [frames_in, height, width, c], out [frames_out, 2h, 2w, c]. */
int orvb_upsample2x_cl(const void* x, void* out, int32_t frames_out, int32_t height, int32_t width, int32_t channels,
                       const int32_t* t_src, void* stream);
/* Channels-last -> planar: out[c, t, h, w] = x[t, h, w, c] for c < c_keep (x has c_ld channels per pixel). */
int orvb_cl_to_planar(const void* x, void* out, int64_t pixels, int32_t c_ld, int32_t c_keep, void* stream);


#ifdef __cplusplus
}
#endif
#endif /* ORV_B200_H_ */
