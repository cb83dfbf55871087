// Whole-model forward: the kernel schedule that replaces CogVideoXTransformer3DModelTraj.forward
// (reference orv/models/cogvideox_control.py:715-948) for modulate_encoder_hidden_states=True models.
//
// Data layout in HBM (all activations in the caller's workspace):
//   x    [B*S, D]   bf16  joint hidden state, per sample: text rows [0,St) then video rows [St,S) in (f,h,w) order
//   xn   [B*S, D]   bf16  LayerNorm+modulated copy (GEMM A operand)
//   qkv  [B*S, 3D]  bf16  Q | K | V, heads are 64-column blocks (read by the attention kernel through TMA)
//   att  [B*S, D]   bf16  attention output
//   ffh  [B*S, FF]  bf16  GELU(ff1) activations
//   mod  [sites][B*G][6D] fp32  AdaLN tables: row b*G is the text/time-only group, rows b*G+1+f the frame groups
// Per layer: LN+mod -> QKV GEMM (+QK-LN, RoPE) -> attention -> out GEMM (+gate, residual) -> LN+mod ->
//            FF1 GEMM (+GELU) -> FF2 GEMM (+gate, residual): 7 launches.
#include <string.h>

#include <vector>

#include "common.cuh"
#include <stdlib.h>

#include "pointwise.cuh"
#include "ptx.cuh"

namespace orvb {
int gemm_run(const orvb_gemm_args* a, cudaStream_t stream);
}

struct orvb_model {
  orvb_config cfg;
  orvb_weights w;
  std::vector<orvb_block_weights> blocks;
  std::vector<orvb_block_weights> mv_blocks;
  bool bound = false;
  // Device copies of the per-block weight pointers (written by orvb_model_bind_weights).  The AdaLN job table
  // (SkinnyJob per site) and the A/B-fold descriptors (AbSite per site) hold pointers INTO the modulation region they
  // describe, so they live in that region (carve_modulation) and are rebuilt from these copies by a one-block kernel
  // at the head of every table build: a captured CUDA graph bakes in a table's address, and a table is never
  // re-pointed at another workspace; nothing is uploaded from the host after bind time.
  orvb_block_weights* blocks_dev = nullptr;     // [layers]
  orvb_block_weights* mv_blocks_dev = nullptr;  // [layers] or null
  int launches = 0;
  std::vector<int> launch_cls;  // profile class of every kernel launch of the most recent call, in launch order
  // optional per-kernel-class timing (orvb_model_set_profile): CUDA events around every launch
  bool profile = false;
  int cur_cls = 0;
  std::vector<cudaEvent_t> ev;
  std::vector<int> ev_cls;
  size_t ev_used = 0;
  float cls_ms[ORVB_PROFILE_CLASSES] = {0};
  int cls_launches[ORVB_PROFILE_CLASSES] = {0};
};

namespace orvb {

static size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

struct Geometry {
  int B, V, F, Fp, H, W, Hp, Wp, St, Sv, S, R, D, FF, T, Kp, Nout, G, Fa, sites;
};

static int make_geometry(const orvb_config& c, const orvb_shape& s, Geometry* g) {
  ORVB_REQUIRE(s.batch > 0 && s.frames > 0 && s.height > 0 && s.width > 0 && s.text_len >= 0, ORVB_ESHAPE,
               "orvb_forward: bad shape batch=%d frames=%d h=%d w=%d text=%d", s.batch, s.frames, s.height, s.width,
               s.text_len);
  ORVB_REQUIRE(s.height % c.patch_size == 0 && s.width % c.patch_size == 0, ORVB_ESHAPE,
               "orvb_forward: latent height/width must be divisible by the patch size");
  const int pt = c.patch_size_t > 0 ? c.patch_size_t : 1;
  ORVB_REQUIRE(s.frames % pt == 0, ORVB_ESHAPE, "orvb_forward: frames (%d) not divisible by patch_size_t (%d)",
               s.frames, pt);
  ORVB_REQUIRE(s.views <= 1 || (c.multiview && s.batch % s.views == 0), ORVB_ESHAPE,
               "orvb_forward: views=%d needs a multiview model and batch (%d) divisible by views", s.views, s.batch);
  g->B = s.batch; g->V = s.views > 0 ? s.views : 1; g->F = s.frames; g->Fp = s.frames / pt;
  g->H = s.height; g->W = s.width; g->Hp = s.height / c.patch_size; g->Wp = s.width / c.patch_size;
  g->St = s.text_len; g->Sv = g->Fp * g->Hp * g->Wp; g->S = g->St + g->Sv; g->R = g->B * g->S;
  g->D = c.dim; g->FF = c.ff_dim; g->T = c.time_embed_dim;
  g->Kp = c.in_channels * pt * c.patch_size * c.patch_size;
  g->Nout = c.out_channels * pt * c.patch_size * c.patch_size;
  g->Fa = s.action_frames;
  ORVB_REQUIRE(g->Fa == 0 || g->Sv % g->Fa == 0, ORVB_ESHAPE,
               "orvb_forward: video tokens (%d) not divisible by action frames (%d)", g->Sv, g->Fa);
  g->G = g->Fa + 1;
  g->sites = 2 * c.layers + 1 + (c.multiview ? c.layers : 0);
  return ORVB_OK;
}

struct Workspace {
  bf16 *x, *xn, *qkv, *att, *ffh, *patches, *ctrl, *yout, *qkv_mv, *att_mv, *tmp_mv;
  float *tsin, *t1, *temb, *osin, *o1, *oemb, *act_in, *act_h, *act_emb, *emb, *mod;
  bf16* emb_hl;  // [groups, 2T]: the conditioning rows split into bf16 halves [hi | lo], A operand of the table GEMMs
  bf16* ab;
  SkinnyJob* jobs;   // [3 * layers] device table of the batched AdaLN linears (points into `mod`)
  AbSite* ab_sites;  // [3 * layers + 1] device table of the A/B folds (points into `mod` / `ab`)
  bf16 *text_cache, *ctrl_cache;  // step-invariant rows kept by ORVB_STATIC_SAVE (text projection, control embeddings)
  size_t bytes;
};

// Width of one AdaLN table row in units of D: [shift, scale, gate | enc_shift, enc_scale, enc_gate] when the text is
// modulated, [shift, scale, gate] otherwise (cogvideox_control.py:57-58).
static inline int mod_width(const orvb_config& c) { return c.modulate_text ? 6 : 3; }

// Scratch + tables of the modulation prologue (sections 1-2 of the forward) for g.B samples.  Also carved on its own
// for a whole schedule of timesteps (orvb_modulation_schedule: g.B = steps x batch virtual samples).
template <typename Take>
static void carve_modulation(const orvb_config& c, const Geometry& g, Take&& take, Workspace* ws) {
  const size_t D = g.D;
  ws->tsin = reinterpret_cast<float*>(take(static_cast<size_t>(g.B) * D * 4));
  ws->t1 = reinterpret_cast<float*>(take(static_cast<size_t>(g.B) * g.T * 4));
  ws->temb = reinterpret_cast<float*>(take(static_cast<size_t>(g.B) * g.T * 4));
  const int od = c.has_ofs ? c.ofs_embed_dim : 8;
  ws->osin = reinterpret_cast<float*>(take(static_cast<size_t>(od) * 4));
  ws->o1 = reinterpret_cast<float*>(take(static_cast<size_t>(g.T) * 4));
  ws->oemb = reinterpret_cast<float*>(take(static_cast<size_t>(g.T) * 4));
  const int pt = c.patch_size_t > 0 ? c.patch_size_t : 1;
  const int act_k = c.action_state_dim * c.action_compress * pt;
  const int fa = g.Fa > 0 ? g.Fa : 1;
  ws->act_in = reinterpret_cast<float*>(take(static_cast<size_t>(g.B) * fa * align_up(act_k, 8) * 4));
  ws->act_h = reinterpret_cast<float*>(take(static_cast<size_t>(g.B) * fa * c.action_hidden * 4));
  ws->act_emb = reinterpret_cast<float*>(take(static_cast<size_t>(g.B) * fa * g.T * 4));
  ws->emb = reinterpret_cast<float*>(take(static_cast<size_t>(g.B) * g.G * g.T * 4));
  ws->emb_hl = reinterpret_cast<bf16*>(take(static_cast<size_t>(g.B) * g.G * g.T * 2 * 2));
  ws->mod = reinterpret_cast<float*>(take(static_cast<size_t>(g.sites) * g.B * g.G * mod_width(c) * D * 4));
  ws->ab = reinterpret_cast<bf16*>(take(static_cast<size_t>(g.sites) * g.B * g.G * 4 * D * 2));
  ws->jobs = reinterpret_cast<SkinnyJob*>(take(sizeof(SkinnyJob) * 3 * c.layers));
  ws->ab_sites = reinterpret_cast<AbSite*>(take(sizeof(AbSite) * (3 * c.layers + 1)));
}

static void carve(const orvb_config& c, const Geometry& g, uint8_t* base, Workspace* ws) {
  size_t off = 0;
  auto take = [&](size_t n) {
    uint8_t* p = base ? base + off : nullptr;
    off += align_up(n);
    return p;
  };
  const size_t R = g.R, D = g.D;
  ws->x = reinterpret_cast<bf16*>(take(R * D * 2));
  ws->xn = reinterpret_cast<bf16*>(take(R * D * 2));
  ws->qkv = reinterpret_cast<bf16*>(take(R * 3 * D * 2));
  ws->att = reinterpret_cast<bf16*>(take(R * D * 2));
  ws->ffh = reinterpret_cast<bf16*>(take(R * static_cast<size_t>(g.FF) * 2));
  ws->patches = reinterpret_cast<bf16*>(take(static_cast<size_t>(g.B) * g.Sv * g.Kp * 2));
  const int keys = c.visual_guidance ? c.num_control_keys : 0;
  ws->ctrl = reinterpret_cast<bf16*>(take(static_cast<size_t>(g.B) * g.Sv * D * keys * 2));
  ws->yout = reinterpret_cast<bf16*>(take(static_cast<size_t>(g.B) * g.Sv * g.Nout * 2));
  {
    // multiview attention operands in '(b f)(v text | v s)' order (clips = B / V sequences of F frames)
    const size_t mv = (c.multiview && g.V > 1) ? 1 : 0;
    const size_t tok = static_cast<size_t>(g.Hp) * g.Wp;
    ws->qkv_mv = reinterpret_cast<bf16*>(take(mv * (g.B / g.V) * g.Fp * g.V * (g.St + tok) * 3 * D * 2));
    ws->att_mv = reinterpret_cast<bf16*>(take(mv * (g.B / g.V) * g.Fp * g.V * tok * D * 2));
    ws->tmp_mv = reinterpret_cast<bf16*>(take(mv * (g.B / g.V) * g.Fp * g.V * tok * D * 2));
  }
  carve_modulation(c, g, take, ws);
  ws->text_cache = reinterpret_cast<bf16*>(take(static_cast<size_t>(g.B) * g.St * D * 2));
  ws->ctrl_cache = reinterpret_cast<bf16*>(take(static_cast<size_t>(g.B) * g.Sv * D * keys * 2));
  ws->bytes = off;
}

// ---------------------------------------------------------------------------------------------------
// small prologue kernels
// ---------------------------------------------------------------------------------------------------
// diffusers get_timestep_embedding: emb = t * exp(-ln(1e4) * i / (half - shift)),
// [sin | cos], flipped to [cos | sin] when flip_sin_to_cos.  The reference casts the result to the model
// dtype (bf16) before time_embedding.linear_1 (cogvideox_control.py:768); that rounding is kept.
__global__ void timestep_sinusoid_kernel(const float* __restrict__ t, float* __restrict__ out, int rows, int dim,
                                         int flip, float shift, float scalar_t, int use_scalar) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim / 2;
  if (idx >= rows * half) return;
  const int r = idx / half, i = idx - r * half;
  const float tv = use_scalar ? scalar_t : t[r];
  const float freq = expf(-logf(10000.0f) * static_cast<float>(i) / (static_cast<float>(half) - shift));
  const float a = tv * freq;
  const float sn = __bfloat162float(__float2bfloat16(sinf(a)));
  const float cs = __bfloat162float(__float2bfloat16(cosf(a)));
  float* o = out + static_cast<size_t>(r) * dim;
  if (flip) {
    o[i] = cs;
    o[half + i] = sn;
  } else {
    o[i] = sn;
    o[half + i] = cs;
  }
}

// actions bf16 [rows, k] -> fp32 [rows, kpad] (zero padded so the skinny linear can use 8-wide loads)
__global__ void actions_to_f32_kernel(const bf16* __restrict__ a, float* __restrict__ out, int rows, int k, int kpad) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * kpad) return;
  const int r = idx / kpad, c = idx - r * kpad;
  out[idx] = c < k ? __bfloat162float(a[static_cast<size_t>(r) * k + c]) : 0.f;
}

// emb[b*G + 0]   = silu(temb[b] + ofs)                         (text / time-only group)
// emb[b*G + 1+f] = silu(temb[b] + ofs + action_emb[b,f])       (frame groups; masked samples use mask_embed)
// Reference: cogvideox_control.py:121-130 (LayerNormZero), :166-170 (AdaLayerNorm), components.py:66-69 (mask).
__global__ void build_emb_kernel(const float* __restrict__ temb, const float* __restrict__ oemb,
                                 const float* __restrict__ act_emb, const uint8_t* __restrict__ mask,
                                 const bf16* __restrict__ mask_embed, float* __restrict__ emb,
                                 bf16* __restrict__ emb_hl, int B, int G, int T) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * G * T) return;
  const int c = idx % T;
  const int g = (idx / T) % G;
  const int b = idx / (T * G);
  float v = temb[b * T + c] + (oemb != nullptr ? oemb[c] : 0.f);
  if (g > 0) {
    float a;
    if (mask != nullptr && mask[b] != 0) a = __bfloat162float(mask_embed[c]);
    else a = act_emb[(static_cast<size_t>(b) * (G - 1) + (g - 1)) * T + c];
    v += a;
  }
  const float e = silu(v);
  emb[idx] = e;
  // e = hi + lo + O(2^-17 e) with both halves bf16: [hi | lo] @ [W | W]^T on the tensor cores is e @ W^T to fp32 accuracy
  const bf16 hi = __float2bfloat16(e);
  const size_t row = static_cast<size_t>(idx / T);
  emb_hl[row * 2 * T + c] = hi;
  emb_hl[row * 2 * T + T + c] = __float2bfloat16(e - __bfloat162float(hi));
}

// ctrl[r, col_off + c] += x[video row r, c]   (hidden_states.repeat(1,1,keys) + controls, :853-855)
__global__ void add_hidden_kernel(bf16* __restrict__ ctrl, const bf16* __restrict__ x, int rows, int D, int ld_ctrl,
                                  int col_off, int Sv, int S, int St) {
  const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  const int chunks = D / 8;
  if (idx >= static_cast<long>(rows) * chunks) return;
  const int r = static_cast<int>(idx / chunks), c = static_cast<int>(idx - static_cast<long>(r) * chunks);
  const int b = r / Sv;
  const size_t xr = static_cast<size_t>(b) * S + St + (r - b * Sv);
  uint4 a = *reinterpret_cast<const uint4*>(ctrl + static_cast<size_t>(r) * ld_ctrl + col_off + c * 8);
  const uint4 h = *reinterpret_cast<const uint4*>(x + xr * D + c * 8);
  a.x = pack_bf16(bf16_lo(a.x) + bf16_lo(h.x), bf16_hi(a.x) + bf16_hi(h.x));
  a.y = pack_bf16(bf16_lo(a.y) + bf16_lo(h.y), bf16_hi(a.y) + bf16_hi(h.y));
  a.z = pack_bf16(bf16_lo(a.z) + bf16_lo(h.z), bf16_hi(a.z) + bf16_hi(h.z));
  a.w = pack_bf16(bf16_lo(a.w) + bf16_lo(h.w), bf16_hi(a.w) + bf16_hi(h.w));
  *reinterpret_cast<uint4*>(ctrl + static_cast<size_t>(r) * ld_ctrl + col_off + c * 8) = a;
}

static inline void note_launch(orvb_model* m) {
  ++m->launches;
  m->launch_cls.push_back(m->cur_cls);
}
static void prof_begin(orvb_model* m, cudaStream_t st) {
  if (!m->profile) return;
  if (m->ev_used + 2 > m->ev.size()) {
    size_t old = m->ev.size();
    m->ev.resize(old + 64);
    for (size_t i = old; i < m->ev.size(); ++i) cudaEventCreate(&m->ev[i]);
  }
  cudaEventRecord(m->ev[m->ev_used], st);
}
static void prof_end(orvb_model* m, cudaStream_t st) {
  if (!m->profile) return;
  cudaEventRecord(m->ev[m->ev_used + 1], st);
  m->ev_cls.push_back(m->cur_cls);
  m->ev_used += 2;
}

#define ORVB_TRY(expr)              \
  do {                              \
    prof_begin(m, st);              \
    int _rc = (expr);               \
    if (_rc != ORVB_OK) return _rc; \
    prof_end(m, st);                \
    note_launch(m);                 \
  } while (0)
#define ORVB_CLS(c) (m->cur_cls = (c))

static orvb_gemm_args gemm_base(const void* a, const void* w, const void* bias, void* out, int M, int N, int K,
                                int lda, int ldo, int epi) {
  orvb_gemm_args g;
  memset(&g, 0, sizeof(g));
  g.a = a; g.w = w; g.bias = bias; g.out = out;
  g.m = M; g.n = N; g.k = K; g.lda = lda; g.ldw = K; g.ldo = ldo; g.epilogue = epi;
  return g;
}

struct ModIn {
  const float* timesteps;
  float ofs;
  const void* actions;
  const uint8_t* action_mask;
};

// Sections 1-2 of the forward: time / ofs / action embeddings -> per-group conditioning rows -> every AdaLN table of
// the model (fp32 `mod`) and their folded LayerNorm form (bf16 `ab`), for g.B samples.  They depend on the timestep,
// ofs and the actions only, never on the latents.
// Fills the job / site tables of one modulation region (one thread per layer).
__global__ void fill_tables_kernel(const orvb_block_weights* __restrict__ blocks, const orvb_block_weights* __restrict__ mv_blocks,
                                   const bf16* norm_out_ln_w, const bf16* norm_out_ln_b, float* mod, bf16* ab,
                                   SkinnyJob* jobs, AbSite* sites, int layers, size_t site_stride, size_t ab_stride,
                                   int mod_ld, int text_off, int D) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l == 0)  // norm_out (AdaLayerNorm, shift first, pitch 2D): only the video variant is ever read
    sites[2 * layers] = AbSite{norm_out_ln_w, norm_out_ln_b, mod + static_cast<size_t>(2 * layers) * site_stride,
                               2 * D, 0, 0, ab + static_cast<size_t>(2 * layers) * ab_stride};
  if (l >= layers) return;
  const orvb_block_weights bw = blocks[l];
  float* m1 = mod + static_cast<size_t>(2 * l) * site_stride;
  float* m2 = m1 + site_stride;
  jobs[2 * l] = SkinnyJob{static_cast<const bf16*>(bw.norm1_lin_w), static_cast<const bf16*>(bw.norm1_lin_b), m1};
  jobs[2 * l + 1] = SkinnyJob{static_cast<const bf16*>(bw.norm2_lin_w), static_cast<const bf16*>(bw.norm2_lin_b), m2};
  sites[2 * l] = AbSite{static_cast<const bf16*>(bw.norm1_ln_w), static_cast<const bf16*>(bw.norm1_ln_b), m1, mod_ld,
                        text_off, 0, ab + static_cast<size_t>(2 * l) * ab_stride};
  sites[2 * l + 1] = AbSite{static_cast<const bf16*>(bw.norm2_ln_w), static_cast<const bf16*>(bw.norm2_ln_b), m2, mod_ld,
                            text_off, 0, ab + static_cast<size_t>(2 * l + 1) * ab_stride};
  if (mv_blocks != nullptr) {
    // MVBlock.norm1 sites live after the norm_out slot: site index 2L + 1 + l, job index 2L + l
    const orvb_block_weights mw = mv_blocks[l];
    const size_t site = static_cast<size_t>(2 * layers + 1 + l);
    jobs[2 * layers + l] = SkinnyJob{static_cast<const bf16*>(mw.norm1_lin_w), static_cast<const bf16*>(mw.norm1_lin_b),
                                     mod + site * site_stride};
    sites[site] = AbSite{static_cast<const bf16*>(mw.norm1_ln_w), static_cast<const bf16*>(mw.norm1_ln_b),
                         mod + site * site_stride, mod_ld, text_off, 0, ab + site * ab_stride};
  }
}

// ORVB_MOD_TABLES_TC=0 keeps the CUDA-core table build (A/B; fp32 activations instead of the [hi | lo] bf16 split).
static bool mod_tables_on_tensor_cores() {
  const char* e = getenv("ORVB_MOD_TABLES_TC");
  return !(e != nullptr && e[0] == '0');
}

static int build_modulation(orvb_model* m, const Geometry& g, const ModIn& in, const Workspace& ws, cudaStream_t st) {
  const orvb_config& c = m->cfg;
  const orvb_weights& w = m->w;
  const int D = g.D, T = g.T;
  const int pt = c.patch_size_t > 0 ? c.patch_size_t : 1;
  const ModIn* a = &in;
  ORVB_CLS(ORVB_PC_PROLOGUE);
  // ---- 1. time / ofs / action embeddings -> per-group conditioning rows -----------------------------
  {
    const int n = g.B * (D / 2);
    timestep_sinusoid_kernel<<<(n + 255) / 256, 256, 0, st>>>(a->timesteps, ws.tsin, g.B, D, c.flip_sin_to_cos,
                                                              c.freq_shift, 0.f, 0);
    ORVB_CHECK_CUDA(cudaGetLastError());
    note_launch(m);
    ORVB_TRY(skinny_linear_launch(ws.tsin, SkinnyJob{static_cast<const bf16*>(w.time1_w), static_cast<const bf16*>(w.time1_b), ws.t1},
                                  nullptr, 1, g.B, T, D, 1, st));
    ORVB_TRY(skinny_linear_launch(ws.t1, SkinnyJob{static_cast<const bf16*>(w.time2_w), static_cast<const bf16*>(w.time2_b), ws.temb},
                                  nullptr, 1, g.B, T, T, 0, st));
  }
  const float* oemb = nullptr;
  if (c.has_ofs) {
    ORVB_REQUIRE(w.ofs1_w && w.ofs2_w, ORVB_EINVAL, "orvb_forward: ofs embedding weights are not bound");
    const int od = c.ofs_embed_dim;
    timestep_sinusoid_kernel<<<(od / 2 + 255) / 256, 256, 0, st>>>(nullptr, ws.osin, 1, od, c.flip_sin_to_cos,
                                                                    c.freq_shift, a->ofs, 1);
    ORVB_CHECK_CUDA(cudaGetLastError());
    note_launch(m);
    ORVB_TRY(skinny_linear_launch(ws.osin, SkinnyJob{static_cast<const bf16*>(w.ofs1_w), static_cast<const bf16*>(w.ofs1_b), ws.o1},
                                  nullptr, 1, 1, T, od, 1, st));
    ORVB_TRY(skinny_linear_launch(ws.o1, SkinnyJob{static_cast<const bf16*>(w.ofs2_w), static_cast<const bf16*>(w.ofs2_b), ws.oemb},
                                  nullptr, 1, 1, T, T, 0, st));
    oemb = ws.oemb;
  }
  if (g.Fa > 0) {
    ORVB_REQUIRE(w.act1_w && w.act2_w, ORVB_EINVAL, "orvb_forward: action_embed weights are not bound");
    ORVB_REQUIRE(a->action_mask == nullptr || w.act_mask_embed != nullptr, ORVB_EINVAL,
                 "orvb_forward: action_mask given but mask_embed is not bound");
    const int act_k = c.action_state_dim * c.action_compress * pt;
    const int kpad = static_cast<int>(align_up(act_k, 8));
    const int rows = g.B * g.Fa;
    actions_to_f32_kernel<<<(rows * kpad + 255) / 256, 256, 0, st>>>(static_cast<const bf16*>(a->actions), ws.act_in,
                                                                     rows, act_k, kpad);
    ORVB_CHECK_CUDA(cudaGetLastError());
    note_launch(m);
    ORVB_TRY(skinny_linear_launch(ws.act_in, SkinnyJob{static_cast<const bf16*>(w.act1_w), static_cast<const bf16*>(w.act1_b), ws.act_h},
                                  nullptr, 1, rows, c.action_hidden, kpad, 2, st));
    ORVB_TRY(skinny_linear_launch(ws.act_h, SkinnyJob{static_cast<const bf16*>(w.act2_w), static_cast<const bf16*>(w.act2_b), ws.act_emb},
                                  nullptr, 1, rows, T, c.action_hidden, 0, st));
  }
  {
    const int n = g.B * g.G * T;
    build_emb_kernel<<<(n + 255) / 256, 256, 0, st>>>(ws.temb, oemb, ws.act_emb, a->action_mask,
                                                      static_cast<const bf16*>(w.act_mask_embed), ws.emb, ws.emb_hl, g.B, g.G, T);
    ORVB_CHECK_CUDA(cudaGetLastError());
    note_launch(m);
  }

  // ---- 2. all AdaLN tables of the forward in one batched launch (they depend only on emb) ----------
  const int mw = mod_width(c);
  const size_t site_stride = static_cast<size_t>(g.B) * g.G * mw * D;
  {
    const size_t ab_stride = static_cast<size_t>(g.B) * g.G * 4 * D;
    fill_tables_kernel<<<(c.layers + 63) / 64, 64, 0, st>>>(
        m->blocks_dev, c.multiview ? m->mv_blocks_dev : nullptr, static_cast<const bf16*>(w.norm_out_ln_w),
        static_cast<const bf16*>(w.norm_out_ln_b), ws.mod, ws.ab, ws.jobs, ws.ab_sites, c.layers, site_stride, ab_stride,
        mw * D, c.modulate_text ? 3 * D : 0, D);  // no text variant without text modulation: it aliases the video one
    ORVB_CHECK_CUDA(cudaGetLastError());
    note_launch(m);
  }
  float* mod_out = ws.mod + static_cast<size_t>(2 * c.layers) * site_stride;  // norm_out table, row pitch 2D
  const int rows = g.B * g.G;
  if (T % 64 == 0 && mod_tables_on_tensor_cores()) {
    // One tcgen05 GEMM per site: [rows, 2T] x [n, T (walked twice)]^T -> fp32 [rows, n] + bias.  The CUDA-core kernel
    // below streams every weight once per 8 rows; for the 300 rows of a 50-step schedule that is 38 passes and 7 ms per
    // clip, the GEMMs take < 1 ms.  The same kernels run for the 6 rows of a single forward (same K order in the
    // single-CTA and the CTA-pair kernel), so a scheduled step and a stand-alone forward see the same table bits.
    auto site_gemm = [&](const void* lw, const void* lb, float* out, int n) -> int {
      orvb_gemm_args ga = gemm_base(ws.emb_hl, lw, lb, out, rows, n, 2 * T, 2 * T, n, ORVB_EPI_BIAS);
      ga.ldw = T;
      ga.k_wrap = T;
      ga.out_f32 = 1;
      return gemm_run(&ga, st);
    };
    for (int l = 0; l < c.layers; ++l) {
      const orvb_block_weights& bw = m->blocks[l];
      ORVB_TRY(site_gemm(bw.norm1_lin_w, bw.norm1_lin_b, ws.mod + static_cast<size_t>(2 * l) * site_stride, mw * D));
      ORVB_TRY(site_gemm(bw.norm2_lin_w, bw.norm2_lin_b, ws.mod + static_cast<size_t>(2 * l + 1) * site_stride, mw * D));
    }
    if (c.multiview) {
      for (int l = 0; l < c.layers; ++l) {
        const orvb_block_weights& bw = m->mv_blocks[l];
        ORVB_TRY(site_gemm(bw.norm1_lin_w, bw.norm1_lin_b, ws.mod + static_cast<size_t>(2 * c.layers + 1 + l) * site_stride, mw * D));
      }
    }
    ORVB_TRY(site_gemm(w.norm_out_lin_w, w.norm_out_lin_b, mod_out, 2 * D));
  } else {
    ORVB_TRY(skinny_linear_launch(ws.emb, SkinnyJob{nullptr, nullptr, nullptr}, ws.jobs,
                                  2 * c.layers + (c.multiview ? c.layers : 0), rows, mw * D, T, 0, st));
    // norm_out.linear is [2D, T]; written with pitch 2D into its slot
    ORVB_TRY(skinny_linear_launch(ws.emb, SkinnyJob{static_cast<const bf16*>(w.norm_out_lin_w), static_cast<const bf16*>(w.norm_out_lin_b), mod_out},
                                  nullptr, 1, rows, 2 * D, T, 0, st));
  }
  // fold LayerNorm affine + (shift, scale) of every site into bf16 A/B tables for the LN kernels
  ORVB_TRY(ab_combine_launch(ws.ab_sites, g.sites, g.B * g.G, D, st));
  return ORVB_OK;
}

// Buffer layout: the modulation scratch + tables of carve_modulation() for steps x batch virtual samples (step-major).
static int schedule_layout(const orvb_model* m, const orvb_shape* s, int steps, uint8_t* base, Geometry* g1, Geometry* gv,
                           Workspace* ws) {
  ORVB_REQUIRE(m && s && steps > 0, ORVB_EINVAL, "orvb_modulation_*: bad arguments");
  int rc = make_geometry(m->cfg, *s, g1);
  if (rc != ORVB_OK) return rc;
  *gv = *g1;
  gv->B = g1->B * steps;
  size_t off = 0;
  auto take = [&](size_t n) {
    uint8_t* p = base ? base + off : nullptr;
    off += align_up(n);
    return p;
  };
  memset(ws, 0, sizeof(*ws));
  carve_modulation(m->cfg, *gv, take, ws);
  ws->bytes = off;
  return ORVB_OK;
}

static int forward_impl(orvb_model* m, const orvb_forward_args* a, cudaStream_t st) {
  const orvb_config& c = m->cfg;
  const orvb_weights& w = m->w;
  Geometry g;
  int rc = make_geometry(c, a->shape, &g);
  if (rc != ORVB_OK) return rc;
  ORVB_REQUIRE(c.modulate_text || a->shape.text_len == 0, ORVB_ESHAPE,
               "orvb_forward: a model without text modulation runs on the video rows alone (shape.text_len must be 0)");
  ORVB_REQUIRE(a->hidden_states && (a->text || a->shape.text_len == 0) && a->timesteps && a->out && a->workspace, ORVB_EINVAL,
               "orvb_forward: null input/output/workspace pointer");
  Workspace ws;
  carve(c, g, static_cast<uint8_t*>(a->workspace), &ws);
  ORVB_REQUIRE(a->workspace_bytes >= ws.bytes, ORVB_ENOMEM, "orvb_forward: workspace too small (%zu < %zu)",
               a->workspace_bytes, ws.bytes);
  ORVB_REQUIRE(reinterpret_cast<uintptr_t>(a->workspace) % 256 == 0, ORVB_ESHAPE,
               "orvb_forward: workspace must be 256-byte aligned");
  ORVB_REQUIRE(!c.use_rope || (a->rope_cos && a->rope_sin), ORVB_EINVAL,
               "orvb_forward: this model uses rotary embeddings but rope_cos/rope_sin are NULL");
  ORVB_REQUIRE(a->skip_modulation || (g.Fa > 0) == (a->actions != nullptr), ORVB_EINVAL,
               "orvb_forward: shape.action_frames and the actions pointer disagree");
  const int D = g.D, T = g.T;
  const int pt = c.patch_size_t > 0 ? c.patch_size_t : 1;
  m->launches = 0;
  m->launch_cls.clear();
  m->ev_used = 0;
  m->ev_cls.clear();
  ORVB_CLS(ORVB_PC_PROLOGUE);

  // ---- 1-2. modulation tables (skipped when the caller installed this step's slice of a schedule, or when the
  //           schedule is read in place through a device-side row offset) ----
  const float* tab_mod = ws.mod;
  const bf16* tab_ab = ws.ab;
  size_t tab_rows = static_cast<size_t>(g.B) * g.G;  // rows per site table
  const int32_t* goff = nullptr;
  if (a->schedule != nullptr) {
    ORVB_REQUIRE(a->schedule_steps > 0 && a->schedule_row_offset != nullptr, ORVB_EINVAL,
                 "orvb_forward: schedule given without schedule_steps / schedule_row_offset");
    Geometry g1, gv;
    Workspace sw;
    int src = schedule_layout(m, &a->shape, a->schedule_steps, const_cast<uint8_t*>(static_cast<const uint8_t*>(a->schedule)),
                              &g1, &gv, &sw);
    if (src != ORVB_OK) return src;
    tab_mod = sw.mod;
    tab_ab = sw.ab;
    tab_rows = static_cast<size_t>(gv.B) * gv.G;
    goff = a->schedule_row_offset;
  }
  if (!a->skip_modulation && a->schedule == nullptr) {
    ModIn in;
    in.timesteps = a->timesteps; in.ofs = a->ofs; in.actions = a->actions; in.action_mask = a->action_mask;
    int mrc = build_modulation(m, g, in, ws, st);
    if (mrc != ORVB_OK) return mrc;
  }
  const int mwid = mod_width(c);
  const int gate_text_off = c.modulate_text ? 5 * D : 2 * D;  // enc_gate, or (no text rows exist) the video gate
  const size_t site_stride = tab_rows * mwid * D;
  const size_t ab_stride = tab_rows * 4 * D;
  ORVB_REQUIRE(a->static_mode >= ORVB_STATIC_COMPUTE && a->static_mode <= ORVB_STATIC_REUSE, ORVB_EINVAL,
               "orvb_forward: unknown static_mode %d", a->static_mode);
  const bool st_save = a->static_mode == ORVB_STATIC_SAVE, st_reuse = a->static_mode == ORVB_STATIC_REUSE;

  orvb_rowmap rm;
  rm.seq_len = g.S; rm.text_len = g.St;
  rm.tokens_per_group = g.Fa > 0 ? g.Sv / g.Fa : 0;
  rm.groups_per_batch = g.G;

  // ---- 3. patch embed + text projection into the joint sequence -----------------------------------
  ORVB_CLS(ORVB_PC_EMBED);
  ORVB_TRY(patchify_launch(a->hidden_states, ws.patches, g.B, g.F, c.in_channels, g.H, g.W, c.patch_size,
                           c.patch_size_t, st));
  {
    orvb_gemm_args ga = gemm_base(ws.patches, w.patch_w, w.patch_b, ws.x, g.B * g.Sv, D, g.Kp, g.Kp, D,
                                  w.pos_embed ? ORVB_EPI_GATE_RESID : ORVB_EPI_BIAS);
    ga.src_rows = g.Sv; ga.dst_rows = g.S; ga.dst_offset = g.St;
    if (w.pos_embed) {
      ga.resid = w.pos_embed; ga.ldr = D; ga.resid_mod = g.Sv; ga.resid_views = g.V; ga.resid_view_stride = g.Sv;
    }
    ORVB_TRY(gemm_run(&ga, st));
  }
  {
    orvb_gemm_args ga = gemm_base(a->text, w.text_w, w.text_b, ws.x, g.B * g.St, D, c.text_embed_dim,
                                  c.text_embed_dim, D, ORVB_EPI_BIAS);
    ga.src_rows = g.St; ga.dst_rows = g.S; ga.dst_offset = 0;
    const size_t tw = static_cast<size_t>(g.St) * D * 2, xp = static_cast<size_t>(g.S) * D * 2;
    if (g.St > 0 && st_reuse) {
      ORVB_CHECK_CUDA(cudaMemcpy2DAsync(ws.x, xp, ws.text_cache, tw, tw, g.B, cudaMemcpyDeviceToDevice, st));
    } else if (g.St > 0) {
      ORVB_TRY(gemm_run(&ga, st));
      if (st_save) ORVB_CHECK_CUDA(cudaMemcpy2DAsync(ws.text_cache, tw, ws.x, xp, tw, g.B, cudaMemcpyDeviceToDevice, st));
    }
  }

  // ---- 4. visual controls (depth / semantic latents), cogvideox_control.py:827-858 -----------------
  if (c.visual_guidance && (a->depths != nullptr || a->labels != nullptr)) {
    const void* ctl[2] = {a->depths, a->labels};
    int present = (a->depths != nullptr) + (a->labels != nullptr);
    ORVB_REQUIRE(present == c.num_control_keys, ORVB_EINVAL,
                 "orvb_forward: Mismatched number of controls: %d given but num_control_keys=%d", present,
                 c.num_control_keys);
    ORVB_REQUIRE(w.combine_w != nullptr, ORVB_EINVAL, "orvb_forward: initial_combine_linear is not bound");
    const int ldc = c.num_control_keys * D;
    const size_t ctrl_bytes = static_cast<size_t>(g.B) * g.Sv * ldc * 2;
    // patch_embed(control) + pos: independent of the noisy latents (step-invariant)
    if (st_reuse) {
      ORVB_CHECK_CUDA(cudaMemcpyAsync(ws.ctrl, ws.ctrl_cache, ctrl_bytes, cudaMemcpyDeviceToDevice, st));
    } else {
      int slot = 0;
      for (int k = 0; k < 2; ++k) {
        if (ctl[k] == nullptr) continue;
        ORVB_TRY(patchify_launch(ctl[k], ws.patches, g.B, g.F, c.in_channels, g.H, g.W, c.patch_size, c.patch_size_t, st));
        orvb_gemm_args ga = gemm_base(ws.patches, w.patch_w, w.patch_b, ws.ctrl + slot * D, g.B * g.Sv, D, g.Kp, g.Kp,
                                      ldc, w.pos_embed ? ORVB_EPI_GATE_RESID : ORVB_EPI_BIAS);
        if (w.pos_embed) {
          ga.resid = w.pos_embed_plain ? w.pos_embed_plain : w.pos_embed; ga.ldr = D; ga.resid_mod = g.Sv; ga.resid_views = 1;
        }
        ORVB_TRY(gemm_run(&ga, st));
        ++slot;
      }
      if (st_save) ORVB_CHECK_CUDA(cudaMemcpyAsync(ws.ctrl_cache, ws.ctrl, ctrl_bytes, cudaMemcpyDeviceToDevice, st));
    }
    // + hidden_states (the step-dependent half, :853-855)
    for (int slot = 0; slot < c.num_control_keys; ++slot) {
      const long n = static_cast<long>(g.B) * g.Sv * (D / 8);
      add_hidden_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(ws.ctrl, ws.x, g.B * g.Sv, D, ldc,
                                                                                slot * D, g.Sv, g.S, g.St);
      ORVB_CHECK_CUDA(cudaGetLastError());
      note_launch(m);
    }
    orvb_gemm_args ga = gemm_base(ws.ctrl, w.combine_w, w.combine_b, ws.x, g.B * g.Sv, D, ldc, ldc, D,
                                  ORVB_EPI_GATE_RESID);
    ga.src_rows = g.Sv; ga.dst_rows = g.S; ga.dst_offset = g.St;
    ga.resid = ws.x; ga.ldr = D;  // hidden_states + combine(...)  (in place, row-aligned)
    ORVB_TRY(gemm_run(&ga, st));
  }

  // ---- 5. transformer blocks -----------------------------------------------------------------------
  const float scale = 1.0f / sqrtf(static_cast<float>(c.head_dim));
  for (int l = 0; l < c.layers; ++l) {
    if (c.multiview && g.V > 1) {
      // ---- MVBlock (cogvideox_control.py:313-348): cross-view attention per (clip, frame) ----
      const orvb_block_weights& mw = m->mv_blocks[l];
      const size_t site = static_cast<size_t>(2 * c.layers + 1 + l);
      const float* modv = tab_mod + site * site_stride;
      const int clips = g.B / g.V, tok = g.Hp * g.Wp, Smv = g.V * (g.St + tok);
      orvb_rowmap rm0 = rm;
      rm0.tokens_per_group = 0;  // norm1(temb) only: every row of a sample uses its time-only group
      orvb_ln_args lnv;
      memset(&lnv, 0, sizeof(lnv));
      lnv.x = ws.x; lnv.y = ws.xn; lnv.rows = g.R; lnv.dim = D; lnv.eps = c.norm_eps; lnv.rowmap = rm0;
      lnv.ab = tab_ab + site * ab_stride; lnv.ab_ld = 4 * D; lnv.group_offset = goff;
      ORVB_CLS(ORVB_PC_LN);
      ORVB_TRY(ln_modulate_launch(&lnv, st));
      orvb_gemm_args q = gemm_base(ws.xn, mw.qkv_w, mw.qkv_b, ws.qkv, g.R, 3 * D, D, D, 3 * D, ORVB_EPI_QKV);
      q.qk_dim = D; q.q_norm_w = mw.q_norm_w; q.q_norm_b = mw.q_norm_b; q.k_norm_w = mw.k_norm_w; q.k_norm_b = mw.k_norm_b;
      q.qk_eps = 1e-6f; q.rowmap = rm0;  // image_rotary_emb_view is never passed on the ORV path
      ORVB_CLS(ORVB_PC_QKV);
      ORVB_TRY(gemm_run(&q, st));
      ORVB_CLS(ORVB_PC_ATTN);
      ORVB_TRY(mv_gather_launch(ws.qkv, ws.qkv_mv, clips, g.V, g.Fp, g.St, tok, 3 * D, st));
      // only the video rows are queries: the reference discards the text outputs of this attention (:333)
      ORVB_TRY(attention_launch(ws.qkv_mv, ws.att_mv, clips * g.Fp, Smv, c.heads, scale, g.V * g.St, g.V * tok, st));
      const int Mv = clips * g.Fp * g.V * tok;
      ORVB_CLS(ORVB_PC_OUT);
      orvb_gemm_args o1 = gemm_base(ws.att_mv, mw.out_w, mw.out_b, ws.tmp_mv, Mv, D, D, D, D, ORVB_EPI_BIAS);
      ORVB_TRY(gemm_run(&o1, st));
      orvb_gemm_args o2 = gemm_base(ws.tmp_mv, mw.proj_out_w, mw.proj_out_b, ws.x, Mv, D, D, D, D, ORVB_EPI_GATE_RESID);
      o2.mv_tokens = tok; o2.mv_frames = g.Fp; o2.mv_views = g.V; o2.dst_rows = g.S; o2.dst_offset = g.St;
      o2.resid = ws.x; o2.ldr = D; o2.gate = modv; o2.gate_ld = mwid * D; o2.gate_text_off = gate_text_off; o2.gate_video_off = 2 * D;
      o2.rowmap = rm0; o2.group_offset = goff;
      ORVB_TRY(gemm_run(&o2, st));
    }
    const orvb_block_weights& bw = m->blocks[l];
    const float* mod1 = tab_mod + (2 * l) * site_stride;
    const float* mod2 = tab_mod + (2 * l + 1) * site_stride;
    orvb_ln_args ln;
    memset(&ln, 0, sizeof(ln));
    ln.x = ws.x; ln.y = ws.xn; ln.ln_w = bw.norm1_ln_w; ln.ln_b = bw.norm1_ln_b;
    ln.rows = g.R; ln.dim = D; ln.eps = c.norm_eps;
    ln.rowmap = rm;
    ln.ab = tab_ab + (2 * l) * ab_stride; ln.ab_ld = 4 * D; ln.group_offset = goff;
    ORVB_CLS(ORVB_PC_LN);
    ORVB_TRY(ln_modulate_launch(&ln, st));

    orvb_gemm_args q = gemm_base(ws.xn, bw.qkv_w, bw.qkv_b, ws.qkv, g.R, 3 * D, D, D, 3 * D, ORVB_EPI_QKV);
    q.qk_dim = D; q.q_norm_w = bw.q_norm_w; q.q_norm_b = bw.q_norm_b; q.k_norm_w = bw.k_norm_w; q.k_norm_b = bw.k_norm_b;
    q.qk_eps = 1e-6f; q.rowmap = rm;
    if (c.use_rope) { q.rope_cos = a->rope_cos; q.rope_sin = a->rope_sin; }
    ORVB_CLS(ORVB_PC_QKV);
    ORVB_TRY(gemm_run(&q, st));

    ORVB_CLS(ORVB_PC_ATTN);
    ORVB_TRY(attention_launch(ws.qkv, ws.att, g.B, g.S, c.heads, scale, 0, g.S, st));

    orvb_gemm_args o = gemm_base(ws.att, bw.out_w, bw.out_b, ws.x, g.R, D, D, D, D, ORVB_EPI_GATE_RESID);
    o.resid = ws.x; o.ldr = D; o.gate = mod1; o.gate_ld = mwid * D; o.gate_text_off = gate_text_off; o.gate_video_off = 2 * D;
    o.rowmap = rm; o.group_offset = goff;
    ORVB_CLS(ORVB_PC_OUT);
    ORVB_TRY(gemm_run(&o, st));

    ln.ab = tab_ab + (2 * l + 1) * ab_stride;
    ORVB_CLS(ORVB_PC_LN);
    ORVB_TRY(ln_modulate_launch(&ln, st));

    orvb_gemm_args f1 = gemm_base(ws.xn, bw.ff1_w, bw.ff1_b, ws.ffh, g.R, g.FF, D, D, g.FF, ORVB_EPI_GELU);
    orvb_gemm_args f2 = gemm_base(ws.ffh, bw.ff2_w, bw.ff2_b, ws.x, g.R, D, g.FF, g.FF, D, ORVB_EPI_GATE_RESID);
    f2.resid = ws.x; f2.ldr = D; f2.gate = mod2; f2.gate_ld = mwid * D; f2.gate_text_off = gate_text_off; f2.gate_video_off = 2 * D;
    f2.rowmap = rm; f2.group_offset = goff;
    ORVB_CLS(ORVB_PC_FF1);
    ORVB_TRY(gemm_run(&f1, st));
    ORVB_CLS(ORVB_PC_FF2);
    ORVB_TRY(gemm_run(&f2, st));

    if (a->tap_hidden != nullptr && a->tap_layer == l) {
      ORVB_CHECK_CUDA(cudaMemcpyAsync(a->tap_hidden, ws.x, static_cast<size_t>(g.R) * D * 2, cudaMemcpyDeviceToDevice, st));
    }
  }

  // ---- 6. norm_final -> norm_out (AdaLN, shift first) -> proj_out -> unpatchify --------------------
  {
    orvb_ln_args ln;
    memset(&ln, 0, sizeof(ln));
    ln.x = ws.x; ln.y = ws.xn; ln.rows = g.B * g.Sv; ln.dim = D;
    ln.pre_w = w.norm_final_w; ln.pre_b = w.norm_final_b; ln.pre_eps = c.norm_eps;
    ln.ln_w = w.norm_out_ln_w; ln.ln_b = w.norm_out_ln_b; ln.eps = c.norm_eps;
    ORVB_CLS(ORVB_PC_HEAD);
    ln.rowmap = rm; ln.in_video_only = 1;
    ln.ab = tab_ab + static_cast<size_t>(2 * c.layers) * ab_stride; ln.ab_ld = 4 * D; ln.group_offset = goff;
    ORVB_TRY(ln_modulate_launch(&ln, st));
    orvb_gemm_args po = gemm_base(ws.xn, w.proj_out_w, w.proj_out_b, ws.yout, g.B * g.Sv, g.Nout, D, D, g.Nout,
                                  ORVB_EPI_BIAS);
    ORVB_TRY(gemm_run(&po, st));
    ORVB_TRY(unpatchify_launch(ws.yout, a->out, g.B, g.F, c.out_channels, g.H, g.W, c.patch_size, c.patch_size_t, st));
  }
  if (m->profile && m->ev_used > 0) {
    ORVB_CHECK_CUDA(cudaEventSynchronize(m->ev[m->ev_used - 1]));
    for (size_t i = 0; i < m->ev_cls.size(); ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, m->ev[2 * i], m->ev[2 * i + 1]);
      m->cls_ms[m->ev_cls[i]] += ms;
      m->cls_launches[m->ev_cls[i]] += 1;
    }
  }
  return ORVB_OK;
}

}  // namespace orvb

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
extern "C" int orvb_model_create(const orvb_config* cfg, orvb_model** out) {
  using namespace orvb;
  ORVB_REQUIRE(cfg && out, ORVB_EINVAL, "orvb_model_create: null pointer");
  ORVB_REQUIRE(cfg->head_dim == 64, ORVB_ESHAPE, "orvb_model_create: attention_head_dim must be 64 (got %d)", cfg->head_dim);
  ORVB_REQUIRE(cfg->dim == cfg->heads * cfg->head_dim && cfg->dim % 64 == 0, ORVB_ESHAPE,
               "orvb_model_create: dim must equal heads*head_dim");
  ORVB_REQUIRE(cfg->layers > 0 && cfg->ff_dim % 8 == 0 && cfg->time_embed_dim % 8 == 0 && cfg->text_embed_dim % 8 == 0,
               ORVB_ESHAPE, "orvb_model_create: bad layer/ff/time/text dims");
  ORVB_REQUIRE(cfg->patch_size == 2, ORVB_ESHAPE, "orvb_model_create: patch_size must be 2");
  ORVB_REQUIRE(cfg->modulate_text || !cfg->multiview, ORVB_ESHAPE,
               "orvb_model_create: multiview without text modulation is not a configuration ORV ships");
  int rc = check_arch();
  if (rc != ORVB_OK) return rc;
  orvb_model* m = new orvb_model();
  m->cfg = *cfg;
  *out = m;
  return ORVB_OK;
}

extern "C" void orvb_model_destroy(orvb_model* m) {
  if (m == nullptr) return;
  if (m->blocks_dev) cudaFree(m->blocks_dev);
  if (m->mv_blocks_dev) cudaFree(m->mv_blocks_dev);
  for (cudaEvent_t e : m->ev) cudaEventDestroy(e);
  delete m;
}

extern "C" int orvb_model_bind_weights(orvb_model* m, const orvb_weights* w) {
  using namespace orvb;
  ORVB_REQUIRE(m && w && w->blocks_host, ORVB_EINVAL, "orvb_model_bind_weights: null pointer");
  ORVB_REQUIRE(w->patch_w && w->text_w && w->time1_w && w->time2_w && w->norm_final_w && w->norm_out_lin_w &&
                   w->norm_out_ln_w && w->proj_out_w,
               ORVB_EINVAL, "orvb_model_bind_weights: a required top-level weight pointer is NULL");
  m->w = *w;
  m->blocks.assign(w->blocks_host, w->blocks_host + m->cfg.layers);
  for (int l = 0; l < m->cfg.layers; ++l) {
    const orvb_block_weights& b = m->blocks[l];
    ORVB_REQUIRE(b.norm1_lin_w && b.norm1_ln_w && b.qkv_w && b.q_norm_w && b.k_norm_w && b.out_w && b.norm2_lin_w &&
                     b.norm2_ln_w && b.ff1_w && b.ff2_w,
                 ORVB_EINVAL, "orvb_model_bind_weights: block %d has a NULL weight pointer", l);
  }
  if (m->cfg.multiview) {
    ORVB_REQUIRE(w->mv_blocks_host != nullptr, ORVB_EINVAL, "orvb_model_bind_weights: multiview model without mv_blocks");
    m->mv_blocks.assign(w->mv_blocks_host, w->mv_blocks_host + m->cfg.layers);
    for (int l = 0; l < m->cfg.layers; ++l) {
      const orvb_block_weights& b = m->mv_blocks[l];
      ORVB_REQUIRE(b.norm1_lin_w && b.norm1_ln_w && b.qkv_w && b.q_norm_w && b.k_norm_w && b.out_w && b.proj_out_w,
                   ORVB_EINVAL, "orvb_model_bind_weights: mv block %d has a NULL weight pointer", l);
    }
  }
  m->w.blocks_host = nullptr;
  m->w.mv_blocks_host = nullptr;
  // device copies of the per-block pointer structs (read by fill_tables_kernel); synchronous, outside any capture
  const size_t bb = sizeof(orvb_block_weights) * m->cfg.layers;
  if (m->blocks_dev == nullptr) ORVB_CHECK_CUDA(cudaMalloc(&m->blocks_dev, bb));
  ORVB_CHECK_CUDA(cudaMemcpy(m->blocks_dev, m->blocks.data(), bb, cudaMemcpyHostToDevice));
  if (m->cfg.multiview) {
    if (m->mv_blocks_dev == nullptr) ORVB_CHECK_CUDA(cudaMalloc(&m->mv_blocks_dev, bb));
    ORVB_CHECK_CUDA(cudaMemcpy(m->mv_blocks_dev, m->mv_blocks.data(), bb, cudaMemcpyHostToDevice));
  }
  m->bound = true;
  return ORVB_OK;
}

extern "C" size_t orvb_workspace_bytes(const orvb_model* m, const orvb_shape* s) {
  using namespace orvb;
  if (m == nullptr || s == nullptr) return 0;
  Geometry g;
  if (make_geometry(m->cfg, *s, &g) != ORVB_OK) return 0;
  Workspace ws;
  carve(m->cfg, g, nullptr, &ws);
  return ws.bytes;
}

// ---- modulation schedule: the AdaLN tables of many timesteps at once ------------------------------------------
extern "C" size_t orvb_modulation_bytes(const orvb_model* m, const orvb_shape* s, int32_t steps) {
  using namespace orvb;
  Geometry g1, gv;
  Workspace ws;
  if (schedule_layout(m, s, steps, nullptr, &g1, &gv, &ws) != ORVB_OK) return 0;
  return ws.bytes;
}

extern "C" int orvb_modulation_schedule(orvb_model* m, const orvb_shape* s, int32_t steps, const float* timesteps,
                                        float ofs, const void* actions, const uint8_t* action_mask, void* tables,
                                        size_t tables_bytes, void* stream) {
  using namespace orvb;
  ORVB_REQUIRE(m && m->bound, ORVB_EINVAL, "orvb_modulation_schedule: weights are not bound");
  int rc = check_arch();
  if (rc != ORVB_OK) return rc;
  ORVB_REQUIRE(timesteps && tables, ORVB_EINVAL, "orvb_modulation_schedule: null pointer");
  ORVB_REQUIRE(reinterpret_cast<uintptr_t>(tables) % 256 == 0, ORVB_ESHAPE, "orvb_modulation_schedule: buffer must be 256-byte aligned");
  Geometry g1, gv;
  Workspace ws;
  rc = schedule_layout(m, s, steps, static_cast<uint8_t*>(tables), &g1, &gv, &ws);
  if (rc != ORVB_OK) return rc;
  ORVB_REQUIRE(tables_bytes >= ws.bytes, ORVB_ENOMEM, "orvb_modulation_schedule: buffer too small (%zu < %zu)", tables_bytes, ws.bytes);
  ORVB_REQUIRE((gv.Fa > 0) == (actions != nullptr), ORVB_EINVAL, "orvb_modulation_schedule: shape.action_frames and the actions pointer disagree");
  m->launches = 0;
  m->launch_cls.clear();
  m->ev_used = 0;
  m->ev_cls.clear();
  ModIn in;
  in.timesteps = timesteps; in.ofs = ofs; in.actions = actions; in.action_mask = action_mask;
  return build_modulation(m, gv, in, ws, static_cast<cudaStream_t>(stream));
}

extern "C" int orvb_modulation_select(const orvb_model* m, const orvb_shape* s, int32_t steps, int32_t step,
                                      const void* tables, void* workspace, void* stream) {
  using namespace orvb;
  ORVB_REQUIRE(tables && workspace && step >= 0 && step < steps, ORVB_EINVAL, "orvb_modulation_select: bad arguments");
  Geometry g1, gv;
  Workspace src, dst;
  int rc = schedule_layout(m, s, steps, const_cast<uint8_t*>(static_cast<const uint8_t*>(tables)), &g1, &gv, &src);
  if (rc != ORVB_OK) return rc;
  carve(m->cfg, g1, static_cast<uint8_t*>(workspace), &dst);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t rows1 = static_cast<size_t>(g1.B) * g1.G, rowsv = static_cast<size_t>(gv.B) * gv.G;
  const size_t D = g1.D;
  // every site: rows [step * B*G, (step+1) * B*G) of its [steps*B*G] row block, mod pitch 6D (3D without text modulation) fp32 / ab pitch 4D bf16
  const size_t mwD = static_cast<size_t>(mod_width(m->cfg)) * D;
  ORVB_CHECK_CUDA(cudaMemcpy2DAsync(dst.mod, rows1 * mwD * 4, src.mod + static_cast<size_t>(step) * rows1 * mwD,
                                    rowsv * mwD * 4, rows1 * mwD * 4, g1.sites, cudaMemcpyDeviceToDevice, st));
  ORVB_CHECK_CUDA(cudaMemcpy2DAsync(dst.ab, rows1 * 4 * D * 2, src.ab + static_cast<size_t>(step) * rows1 * 4 * D,
                                    rowsv * 4 * D * 2, rows1 * 4 * D * 2, g1.sites, cudaMemcpyDeviceToDevice, st));
  // norm_out's table is packed with row pitch 2D inside its slot
  const size_t so = static_cast<size_t>(2 * m->cfg.layers);
  ORVB_CHECK_CUDA(cudaMemcpyAsync(dst.mod + so * rows1 * mwD, src.mod + so * rowsv * mwD + static_cast<size_t>(step) * rows1 * 2 * D,
                                  rows1 * 2 * D * 4, cudaMemcpyDeviceToDevice, st));
  return ORVB_OK;
}

extern "C" int orvb_forward(orvb_model* m, const orvb_forward_args* a, void* stream) {
  using namespace orvb;
  ORVB_REQUIRE(m && a, ORVB_EINVAL, "orvb_forward: null pointer");
  ORVB_REQUIRE(m->bound, ORVB_EINVAL, "orvb_forward: weights are not bound");
  int rc = check_arch();
  if (rc != ORVB_OK) return rc;
  return forward_impl(m, a, static_cast<cudaStream_t>(stream));
}

extern "C" int orvb_last_launch_count(const orvb_model* m) { return m ? m->launches : 0; }

extern "C" int orvb_last_launch_classes(const orvb_model* m, int32_t* classes_out, int32_t capacity) {
  if (m == nullptr) return 0;
  const int n = static_cast<int>(m->launch_cls.size());
  for (int i = 0; i < n && i < capacity && classes_out != nullptr; ++i) classes_out[i] = m->launch_cls[i];
  return n;
}

extern "C" int orvb_model_set_profile(orvb_model* m, int enable) {
  using namespace orvb;
  ORVB_REQUIRE(m != nullptr, ORVB_EINVAL, "orvb_model_set_profile: null model");
  m->profile = enable != 0;
  for (int i = 0; i < ORVB_PROFILE_CLASSES; ++i) {
    m->cls_ms[i] = 0.f;
    m->cls_launches[i] = 0;
  }
  return ORVB_OK;
}

extern "C" int orvb_model_get_profile(const orvb_model* m, float* ms_out, int32_t* launches_out) {
  using namespace orvb;
  ORVB_REQUIRE(m != nullptr && ms_out != nullptr, ORVB_EINVAL, "orvb_model_get_profile: null pointer");
  for (int i = 0; i < ORVB_PROFILE_CLASSES; ++i) {
    ms_out[i] = m->cls_ms[i];
    if (launches_out) launches_out[i] = m->cls_launches[i];
  }
  return ORVB_OK;
}
