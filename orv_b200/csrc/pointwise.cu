// Memory-bound kernels of the path: LayerNorm + AdaLN modulation, the skinny (few-row) linears that build the
// modulation tables, patch gather / scatter, and the fused sampler step.  All of them are plain SIMT kernels
// with 128-bit global accesses; none is worth a tensor core.
#include "common.cuh"
#include "ptx.cuh"
#include "pointwise.cuh"

namespace orvb {

__device__ __forceinline__ int row_group_pw(const orvb_rowmap& rm, int row, int* is_text) {
  if (rm.seq_len <= 0) {
    *is_text = 0;
    return 0;
  }
  int b = row / rm.seq_len;
  int s = row - b * rm.seq_len;
  *is_text = (s < rm.text_len);
  int g = (s < rm.text_len || rm.tokens_per_group <= 0) ? 0 : 1 + (s - rm.text_len) / rm.tokens_per_group;
  return b * rm.groups_per_batch + g;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}

// ---------------------------------------------------------------------------------------------------
// LayerNorm (+ optional preceding LayerNorm) + AdaLN modulate.  One warp per row, row kept in registers.
// Reference: CogVideoXLayerNormZero.forward (cogvideox_control.py:117-145), AdaLayerNorm.forward (:153-197),
// norm_final (:909-916).
// ---------------------------------------------------------------------------------------------------
struct LnDev {
  const bf16* x;
  bf16* y;
  const bf16 *w, *b, *pre_w, *pre_b;
  int rows, dim;
  float eps, pre_eps;
  const float* mod;
  int mod_ld, text_off, video_off, scale_first;
  orvb_rowmap rm;
  int in_video_only;
  const bf16* ab;  // optional pre-combined table: per group row [text A | text B | video A | video B], each `dim`
  int ab_ld;
  int y_f32;       // test mode: y is fp32 (generic kernel only)
  const int* grp_off;  // optional device scalar added to every group index (schedule slice of this step)
};

__device__ __forceinline__ void ln_store8(const LnDev& p, int row, int c, const float (&o)[8]) {
  if (p.y_f32) {
    float* yf = reinterpret_cast<float*>(p.y) + static_cast<size_t>(row) * p.dim + c * 8;
    *reinterpret_cast<float4*>(yf) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(yf + 4) = make_float4(o[4], o[5], o[6], o[7]);
    return;
  }
  uint4 u;
  u.x = pack_bf16(o[0], o[1]); u.y = pack_bf16(o[2], o[3]);
  u.z = pack_bf16(o[4], o[5]); u.w = pack_bf16(o[6], o[7]);
  *reinterpret_cast<uint4*>(p.y + static_cast<size_t>(row) * p.dim + c * 8) = u;
}

template <int MAXC>
__device__ __forceinline__ void ln_stats(const float (&v)[MAXC][8], int nchunks, int lane, int dim, float eps,
                                         float* mean_out, float* rstd_out) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXC; ++i)
    if (lane + 32 * i < nchunks) {
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];
    }
  const float mean = warp_sum(s) / static_cast<float>(dim);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXC; ++i)
    if (lane + 32 * i < nchunks) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float d = v[i][j] - mean;
        q += d * d;
      }
    }
  const float var = warp_sum(q) / static_cast<float>(dim);
  *mean_out = mean;
  *rstd_out = rsqrtf(var + eps);
}

template <int MAXC>
__global__ void __launch_bounds__(256) ln_modulate_kernel(const LnDev p) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= p.rows) return;
  const int nchunks = p.dim >> 3;
  int in_row = row;
  if (p.in_video_only) {
    const int sv = p.rm.seq_len - p.rm.text_len;
    const int b = row / sv;
    in_row = b * p.rm.seq_len + p.rm.text_len + (row - b * sv);
  }
  const bf16* xr = p.x + static_cast<size_t>(in_row) * p.dim;
  float v[MAXC][8];
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    const int c = lane + 32 * i;
    if (c < nchunks) {
      uint4 u = *reinterpret_cast<const uint4*>(xr + c * 8);
      unpack8(u, v[i]);
    }
  }
  float mean, rstd;
  if (p.pre_w != nullptr) {
    ln_stats<MAXC>(v, nchunks, lane, p.dim, p.pre_eps, &mean, &rstd);
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunks) {
        float w[8], b[8];
        unpack8(*reinterpret_cast<const uint4*>(p.pre_w + c * 8), w);
        unpack8(*reinterpret_cast<const uint4*>(p.pre_b + c * 8), b);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          // kept in fp32 (the bf16 reference rounds here; the fp32 oracle does not)
          v[i][j] = (v[i][j] - mean) * rstd * w[j] + b[j];
        }
      }
    }
  }
  if (p.ab != nullptr) {
    // y = xhat * A_g + B_g with A = w * (1 + scale), B = b * (1 + scale) + shift folded once per forward
    // (ab_combine_kernel): one memory phase, every load issued before the reductions.
    int is_text;
    const int g = row_group_pw(p.rm, in_row, &is_text) + (p.grp_off != nullptr ? *p.grp_off : 0);
    const bf16* ap = p.ab + static_cast<size_t>(g) * p.ab_ld + (is_text ? 0 : 2 * p.dim);
    uint4 av[MAXC], bv[MAXC];
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunks) {
        av[i] = *reinterpret_cast<const uint4*>(ap + c * 8);
        bv[i] = *reinterpret_cast<const uint4*>(ap + p.dim + c * 8);
      }
    }
    ln_stats<MAXC>(v, nchunks, lane, p.dim, p.eps, &mean, &rstd);
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunks) {
        float a8[8], b8[8], o[8];
        unpack8(av[i], a8);
        unpack8(bv[i], b8);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * a8[j] + b8[j];
        ln_store8(p, row, c, o);
      }
    }
    return;
  }
  ln_stats<MAXC>(v, nchunks, lane, p.dim, p.eps, &mean, &rstd);

  const float* shift = nullptr;
  const float* scale = nullptr;
  if (p.mod != nullptr) {
    int is_text;
    const int g = row_group_pw(p.rm, in_row, &is_text) + (p.grp_off != nullptr ? *p.grp_off : 0);
    const float* base = p.mod + static_cast<size_t>(g) * p.mod_ld + (is_text ? p.text_off : p.video_off);
    shift = p.scale_first ? base + p.dim : base;
    scale = p.scale_first ? base : base + p.dim;
  }
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    const int c = lane + 32 * i;
    if (c < nchunks) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd;
      if (p.w != nullptr) {
        float w[8], b[8];
        unpack8(*reinterpret_cast<const uint4*>(p.w + c * 8), w);
        unpack8(*reinterpret_cast<const uint4*>(p.b + c * 8), b);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = o[j] * w[j] + b[j];
      }
      if (shift != nullptr) {
        const float4 s0 = *reinterpret_cast<const float4*>(scale + c * 8);
        const float4 s1 = *reinterpret_cast<const float4*>(scale + c * 8 + 4);
        const float4 h0 = *reinterpret_cast<const float4*>(shift + c * 8);
        const float4 h1 = *reinterpret_cast<const float4*>(shift + c * 8 + 4);
        o[0] = o[0] * (1.f + s0.x) + h0.x; o[1] = o[1] * (1.f + s0.y) + h0.y;
        o[2] = o[2] * (1.f + s0.z) + h0.z; o[3] = o[3] * (1.f + s0.w) + h0.w;
        o[4] = o[4] * (1.f + s1.x) + h1.x; o[5] = o[5] * (1.f + s1.y) + h1.y;
        o[6] = o[6] * (1.f + s1.z) + h1.z; o[7] = o[7] * (1.f + s1.w) + h1.w;
      }
      ln_store8(p, row, c, o);
    }
  }
}

// Hot variant of the kernel above for the block norms (60 launches per forward): y = xhat * A_g + B_g with the folded
// A/B tables.  Eight consecutive rows per CTA touch at most two (group, text|video) table rows (every segment of the
// sequence is far longer than 8 rows), so those are staged once in shared memory instead of being prefetched into
// every warp's registers; x stays packed (bf16) in registers.  ~60 registers per thread: all 3226 rows of a 2B
// forward are resident in one wave (the register-heavy generic kernel ran 1 CTA per SM, i.e. three waves).
template <int MAXC>
__global__ void __launch_bounds__(256, 4) ln_ab_kernel(const LnDev p) {
  extern __shared__ uint4 s_ab[];  // [2 slots][A: dim/8 chunks | B: dim/8 chunks]
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nchunks = p.dim >> 3;
  const int row0 = blockIdx.x * 8;
  const int row_last = min(row0 + 7, p.rows - 1);
  int t0, t1;
  const int g0 = row_group_pw(p.rm, row0, &t0);
  const int g1 = row_group_pw(p.rm, row_last, &t1);
  const int row = row0 + warp;
  const bool live = row < p.rows;
  pdl_launch_dependents();
  pdl_wait();  // x comes from the previous kernel (and the group offset from an earlier one)
  const int goff = p.grp_off != nullptr ? *p.grp_off : 0;
  const bf16* ap0 = p.ab + static_cast<size_t>(g0 + goff) * p.ab_ld + (t0 ? 0 : 2 * p.dim);
  const bf16* ap1 = p.ab + static_cast<size_t>(g1 + goff) * p.ab_ld + (t1 ? 0 : 2 * p.dim);
  // this warp's row first (longest latency), then the cooperative table copy
  uint4 xr[MAXC];
  if (live) {
    const bf16* xrow = p.x + static_cast<size_t>(row) * p.dim;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunks) xr[i] = *reinterpret_cast<const uint4*>(xrow + c * 8);
    }
  }
  const int two = 2 * nchunks;  // A then B, contiguous in the table row
  for (int i = threadIdx.x; i < two; i += 256) s_ab[i] = *reinterpret_cast<const uint4*>(ap0 + i * 8);
  if (ap1 != ap0)
    for (int i = threadIdx.x; i < two; i += 256) s_ab[two + i] = *reinterpret_cast<const uint4*>(ap1 + i * 8);
  __syncthreads();
  if (!live) return;
  int tt;
  const int g = row_group_pw(p.rm, row, &tt);
  const uint4* tab = s_ab + ((g == g0 && tt == t0) ? 0 : two);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXC; ++i)
    if (lane + 32 * i < nchunks) {
      float f[8];
      unpack8(xr[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += f[j];
    }
  const float mean = warp_sum(s) / static_cast<float>(p.dim);
  float qv = 0.f;
#pragma unroll
  for (int i = 0; i < MAXC; ++i)
    if (lane + 32 * i < nchunks) {
      float f[8];
      unpack8(xr[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = f[j] - mean;
        qv += d * d;
      }
    }
  const float rstd = rsqrtf(warp_sum(qv) / static_cast<float>(p.dim) + p.eps);
  bf16* yr = p.y + static_cast<size_t>(row) * p.dim;
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    const int c = lane + 32 * i;
    if (c < nchunks) {
      float f[8], a8[8], b8[8], o[8];
      unpack8(xr[i], f);
      unpack8(tab[c], a8);
      unpack8(tab[nchunks + c], b8);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (f[j] - mean) * rstd * a8[j] + b8[j];
      uint4 u;
      u.x = pack_bf16(o[0], o[1]); u.y = pack_bf16(o[2], o[3]);
      u.z = pack_bf16(o[4], o[5]); u.w = pack_bf16(o[6], o[7]);
      *reinterpret_cast<uint4*>(yr + c * 8) = u;
    }
  }
}

// A/B table for the LayerNorm+modulate kernel, built once per forward from the fp32 AdaLN tables:
//   text  variant: A = w (1 + mod[g][text_off + D ..]),  B = b (1 + ...) + mod[g][text_off ..]
//   video variant: A = w (1 + mod[g][video_off + D ..]), B = b (1 + ...) + mod[g][video_off ..]
// One block row per (site, group); sites are described by a device array (weights differ per site).
__global__ void ab_combine_kernel(const AbSite* __restrict__ sites, int groups, int dim) {
  const AbSite st = sites[blockIdx.z];
  const int g = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= dim) return;
  const float w = st.ln_w ? __bfloat162float(st.ln_w[c]) : 1.f;
  const float b = st.ln_b ? __bfloat162float(st.ln_b[c]) : 0.f;
  const float* m = st.mod + static_cast<size_t>(g) * st.mod_ld;
  bf16* o = st.ab + static_cast<size_t>(g) * 4 * dim;
  const float ts = 1.f + m[st.text_off + dim + c], th = m[st.text_off + c];
  const float vs = 1.f + m[st.video_off + dim + c], vh = m[st.video_off + c];
  o[c] = __float2bfloat16(w * ts);
  o[dim + c] = __float2bfloat16(b * ts + th);
  o[2 * dim + c] = __float2bfloat16(w * vs);
  o[3 * dim + c] = __float2bfloat16(b * vs + vh);
}

int ab_combine_launch(const AbSite* sites_dev, int num_sites, int groups, int dim, cudaStream_t stream) {
  dim3 grid((dim + 255) / 256, groups, num_sites);
  ab_combine_kernel<<<grid, 256, 0, stream>>>(sites_dev, groups, dim);
  ORVB_CHECK_CUDA(cudaGetLastError());
  return ORVB_OK;
}

int ln_modulate_launch(const orvb_ln_args* a, cudaStream_t stream) {
  ORVB_REQUIRE(a && a->x && a->y, ORVB_EINVAL, "orvb_ln_modulate: null pointer");
  ORVB_REQUIRE(a->rows > 0 && a->dim > 0 && a->dim % 8 == 0 && a->dim <= 4096, ORVB_ESHAPE,
               "orvb_ln_modulate: dim must be a multiple of 8 and <= 4096 (got %d)", a->dim);
  ORVB_REQUIRE((a->ln_w == nullptr) == (a->ln_b == nullptr), ORVB_EINVAL, "orvb_ln_modulate: ln_w/ln_b must pair");
  ORVB_REQUIRE(a->mod == nullptr || (a->mod_ld % 4 == 0 && a->text_off % 4 == 0 && a->video_off % 4 == 0), ORVB_ESHAPE,
               "orvb_ln_modulate: modulation pitch/offsets must be multiples of 4");
  ORVB_REQUIRE(!a->in_video_only || (a->rowmap.seq_len > a->rowmap.text_len), ORVB_EINVAL,
               "orvb_ln_modulate: in_video_only needs a row map");
  LnDev d;
  d.x = static_cast<const bf16*>(a->x); d.y = static_cast<bf16*>(a->y);
  d.w = static_cast<const bf16*>(a->ln_w); d.b = static_cast<const bf16*>(a->ln_b);
  d.pre_w = static_cast<const bf16*>(a->pre_w); d.pre_b = static_cast<const bf16*>(a->pre_b);
  d.rows = a->rows; d.dim = a->dim; d.eps = a->eps; d.pre_eps = a->pre_eps;
  d.mod = a->mod; d.mod_ld = a->mod_ld; d.text_off = a->text_off; d.video_off = a->video_off;
  d.scale_first = a->scale_first; d.rm = a->rowmap; d.in_video_only = a->in_video_only;
  d.ab = static_cast<const bf16*>(a->ab); d.ab_ld = a->ab_ld;
  d.y_f32 = a->y_f32 ? 1 : 0;
  d.grp_off = a->group_offset;
  ORVB_REQUIRE(d.ab == nullptr || (a->ab_ld % 8 == 0 && a->ab_ld >= 4 * a->dim), ORVB_ESHAPE,
               "orvb_ln_modulate: ab_ld must be a multiple of 8 and >= 4*dim");
  const int rows_per_block = 8;
  dim3 grid((a->rows + rows_per_block - 1) / rows_per_block);
  const int nchunks = a->dim / 8;
  // A then B must be adjacent in a table row (ab_combine writes them so) for the staged copy of the hot kernel
  if (d.ab != nullptr && !d.y_f32 && d.pre_w == nullptr && !d.in_video_only && d.rm.seq_len > 0 &&
      (d.rm.text_len >= 8 || d.rm.text_len == 0) &&
      (d.rm.tokens_per_group <= 0 || d.rm.tokens_per_group >= 8) && d.rm.seq_len - d.rm.text_len >= 8) {
    const int smem = 2 * 2 * nchunks * 16;
    if (nchunks <= 32 * 8) ORVB_CHECK_CUDA(launch_kernel(ln_ab_kernel<8>, grid, dim3(256), smem, stream, true, d));
    else if (nchunks <= 32 * 12) ORVB_CHECK_CUDA(launch_kernel(ln_ab_kernel<12>, grid, dim3(256), smem, stream, true, d));
    else ORVB_CHECK_CUDA(launch_kernel(ln_ab_kernel<16>, grid, dim3(256), smem, stream, true, d));
    return ORVB_OK;
  }
  if (nchunks <= 32 * 8) ln_modulate_kernel<8><<<grid, 256, 0, stream>>>(d);
  else if (nchunks <= 32 * 12) ln_modulate_kernel<12><<<grid, 256, 0, stream>>>(d);
  else ln_modulate_kernel<16><<<grid, 256, 0, stream>>>(d);
  ORVB_CHECK_CUDA(cudaGetLastError());
  return ORVB_OK;
}

// ---------------------------------------------------------------------------------------------------
// Skinny linear: y[r, n] = act(x[r, :] . W[n, :] + b[n]) for <= 8 rows per pass.  HBM-bound on W (the AdaLN
// linears hold 354 M parameters in the 2B model: ~0.7 GB read per forward), so the kernel is written for few
// instructions AND few shared-memory reads per weight byte: four lanes share a weight row (a warp reads 64 contiguous
// bytes of 8 rows per load), each lane quad works on SK_QCOLS weight rows at once so that one 32-byte read of the
// fp32 activations from shared memory feeds SK_QCOLS x 4 packed FFMA2 (ncu on the 1-row-per-quad version: 31 % issue
// activity, top stall short_scoreboard — the broadcast LDS.128 of x cost 4 LSU passes each and bounded the kernel at
// 2.2 TB/s), and the only cross-lane traffic is a two-step shuffle over the four lanes of a row.
// ---------------------------------------------------------------------------------------------------
constexpr int SK_ROWS = 8;
constexpr int SK_WARPS = 8;

// SK_QCOLS = weight rows per lane quad: 4 for the big batched launch, 1 for the small MLPs (few columns: spread them
// over more CTAs instead).
template <bool BATCHED, int SK_QCOLS>
__global__ void __launch_bounds__(256) skinny_linear_kernel(const float* __restrict__ x, SkinnyJob single,
                                                            const SkinnyJob* __restrict__ jobs, int rows, int n,
                                                            int k, int act) {
  extern __shared__ float sx[];  // [nr][k] activations of this row chunk, shared by all columns of the block
  const SkinnyJob job = BATCHED ? jobs[blockIdx.z] : single;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int r0 = blockIdx.y * SK_ROWS;
  const int nr = min(SK_ROWS, rows - r0);
  for (int i = threadIdx.x * 4; i < nr * k; i += blockDim.x * 4)
    *reinterpret_cast<float4*>(sx + i) = *reinterpret_cast<const float4*>(x + static_cast<size_t>(r0) * k + i);
  __syncthreads();
  constexpr int SK_COLS = 8 * SK_QCOLS;  // weight rows (output columns) per warp
  // quad g of the warp owns columns col0 + g + 8 * c, c < SK_QCOLS: a warp-wide load still touches 8 adjacent rows
  const int col0 = (blockIdx.x * SK_WARPS + warp) * SK_COLS + (lane >> 2);
  const int q = lane & 3;
  const bf16* wrow[SK_QCOLS];
#pragma unroll
  for (int c = 0; c < SK_QCOLS; ++c) {
    const int col = col0 + 8 * c;
    wrow[c] = job.w + static_cast<size_t>(col < n ? col : 0) * k;
  }
  f32x2 acc[SK_QCOLS][SK_ROWS];
#pragma unroll
  for (int c = 0; c < SK_QCOLS; ++c)
#pragma unroll
    for (int r = 0; r < SK_ROWS; ++r) acc[c][r] = pk2(0.f, 0.f);

  // Weight loads are issued U steps ahead of their use (explicitly: the row-count branches below keep the compiler
  // from hoisting them).  The small launches (one weight row per quad, a handful of CTAs) are pure load-latency
  // chains and take U = 8; the batched launch has 4 rows per quad in flight already (U = 1, loop unrolled twice: more
  // look-ahead costs registers, i.e. the second resident CTA, and measured slower).
  constexpr int U = (SK_QCOLS == 1) ? 8 : 1;
#pragma unroll(U == 1 ? 2 : 1)
  for (int kb = q * 8; kb < k; kb += 32 * U) {
    uint4 wq[U][SK_QCOLS];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (kb + 32 * u < k) {
#pragma unroll
        for (int c = 0; c < SK_QCOLS; ++c) wq[u][c] = *reinterpret_cast<const uint4*>(wrow[c] + kb + 32 * u);
      }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int k0 = kb + 32 * u;
      if (k0 < k) {
        f32x2 w[SK_QCOLS][4];
#pragma unroll
        for (int c = 0; c < SK_QCOLS; ++c) {
          w[c][0] = pk2(bf16_lo(wq[u][c].x), bf16_hi(wq[u][c].x));
          w[c][1] = pk2(bf16_lo(wq[u][c].y), bf16_hi(wq[u][c].y));
          w[c][2] = pk2(bf16_lo(wq[u][c].z), bf16_hi(wq[u][c].z));
          w[c][3] = pk2(bf16_lo(wq[u][c].w), bf16_hi(wq[u][c].w));
        }
#pragma unroll
        for (int r = 0; r < SK_ROWS; ++r) {
          if (r < nr) {
            const ulonglong2 xa = *reinterpret_cast<const ulonglong2*>(sx + r * k + k0);
            const ulonglong2 xb = *reinterpret_cast<const ulonglong2*>(sx + r * k + k0 + 4);
#pragma unroll
            for (int c = 0; c < SK_QCOLS; ++c) {
              acc[c][r] = fma2p(xa.x, w[c][0], acc[c][r]);
              acc[c][r] = fma2p(xa.y, w[c][1], acc[c][r]);
              acc[c][r] = fma2p(xb.x, w[c][2], acc[c][r]);
              acc[c][r] = fma2p(xb.y, w[c][3], acc[c][r]);
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < SK_QCOLS; ++c) {
    const int col = col0 + 8 * c;
#pragma unroll
    for (int r = 0; r < SK_ROWS; ++r) {
      if (r < nr) {
        float lo, hi;
        upk2(acc[c][r], lo, hi);
        float v = lo + hi;
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (q == 0 && col < n) {
          v += (job.b != nullptr ? __bfloat162float(job.b[col]) : 0.f);
          if (act == 1) v = silu(v);
          else if (act == 2) v = gelu_tanh(v);
          job.y[static_cast<size_t>(r0 + r) * n + col] = v;
        }
      }
    }
  }
}

int skinny_linear_launch(const float* x, const SkinnyJob& job, const SkinnyJob* jobs_dev, int num_jobs, int rows,
                         int n, int k, int act, cudaStream_t stream) {
  ORVB_REQUIRE(x != nullptr && rows > 0 && n > 0 && k > 0, ORVB_EINVAL, "orvb_skinny_linear: bad arguments");
  ORVB_REQUIRE(k % 8 == 0 && k <= 12288, ORVB_ESHAPE, "orvb_skinny_linear: k must be a multiple of 8 and <= 12288 (got %d)", k);
  const int njobs = jobs_dev ? num_jobs : 1;
  const bool wide = static_cast<long>(n) * njobs >= 16384;
  const int cols_per_block = SK_WARPS * 8 * (wide ? 4 : 1);
  const int smem = SK_ROWS * k * static_cast<int>(sizeof(float));
  dim3 grid((n + cols_per_block - 1) / cols_per_block, (rows + SK_ROWS - 1) / SK_ROWS, njobs);
  static bool attr_set = false;
  if (!attr_set) {
    ORVB_CHECK_CUDA(cudaFuncSetAttribute(skinny_linear_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    ORVB_CHECK_CUDA(cudaFuncSetAttribute(skinny_linear_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    ORVB_CHECK_CUDA(cudaFuncSetAttribute(skinny_linear_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    ORVB_CHECK_CUDA(cudaFuncSetAttribute(skinny_linear_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  ORVB_REQUIRE(smem <= 200 * 1024, ORVB_ESHAPE, "orvb_skinny_linear: k too large for the shared-memory staging");
  if (jobs_dev != nullptr) {
    if (wide) skinny_linear_kernel<true, 4><<<grid, 256, smem, stream>>>(x, job, jobs_dev, rows, n, k, act);
    else skinny_linear_kernel<true, 1><<<grid, 256, smem, stream>>>(x, job, jobs_dev, rows, n, k, act);
  } else {
    if (wide) skinny_linear_kernel<false, 4><<<grid, 256, smem, stream>>>(x, job, nullptr, rows, n, k, act);
    else skinny_linear_kernel<false, 1><<<grid, 256, smem, stream>>>(x, job, nullptr, rows, n, k, act);
  }
  ORVB_CHECK_CUDA(cudaGetLastError());
  return ORVB_OK;
}

// ---------------------------------------------------------------------------------------------------
// Patch gather / scatter (p = 2).  Index maps are the closed forms of SURVEY App. A.1 / A.6 and must be
// bit-exact with the reference's reshape/permute chains.
// ---------------------------------------------------------------------------------------------------
__global__ void patchify_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int B, int F, int C, int H, int W,
                                int pt) {
  // one thread per (token, channel, t): gathers the 2x2 spatial patch
  const int Hp = H >> 1, Wp = W >> 1, Fp = F / pt;
  const long total = static_cast<long>(B) * Fp * Hp * Wp * C * pt;
  const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  int j = idx % Wp;
  long r = idx / Wp;
  int i = r % Hp; r /= Hp;
  int t = r % pt; r /= pt;
  int c = r % C; r /= C;
  int fp = r % Fp;
  int b = r / Fp;
  const int f = fp * pt + t;
  const bf16* src = x + ((((static_cast<size_t>(b) * F + f) * C + c) * H + 2 * i) * W + 2 * j);
  const uint32_t top = *reinterpret_cast<const uint32_t*>(src);
  const uint32_t bot = *reinterpret_cast<const uint32_t*>(src + W);
  const size_t tok = (static_cast<size_t>(b) * Fp + fp) * Hp * Wp + static_cast<size_t>(i) * Wp + j;
  const int kdim = C * pt * 4;
  uint2 v;
  v.x = top;
  v.y = bot;
  *reinterpret_cast<uint2*>(out + tok * kdim + (c * pt + t) * 4) = v;
}

__global__ void unpatchify_kernel(const bf16* __restrict__ y, bf16* __restrict__ out, int B, int F, int C, int H,
                                  int W, int pt) {
  const int Hp = H >> 1, Wp = W >> 1, Fp = F / pt;
  const long total = static_cast<long>(B) * Fp * Hp * Wp * C * pt;
  const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  int j = idx % Wp;
  long r = idx / Wp;
  int i = r % Hp; r /= Hp;
  int t = r % pt; r /= pt;
  int c = r % C; r /= C;
  int fp = r % Fp;
  int b = r / Fp;
  const int f = fp * pt + t;
  const size_t tok = (static_cast<size_t>(b) * Fp + fp) * Hp * Wp + static_cast<size_t>(i) * Wp + j;
  const int kdim = C * pt * 4;
  const uint2 v = *reinterpret_cast<const uint2*>(y + tok * kdim + (c * pt + t) * 4);
  bf16* dst = out + ((((static_cast<size_t>(b) * F + f) * C + c) * H + 2 * i) * W + 2 * j);
  *reinterpret_cast<uint32_t*>(dst) = v.x;
  *reinterpret_cast<uint32_t*>(dst + W) = v.y;
}

static int patch_check(const void* a, const void* b, int B, int F, int C, int H, int W, int p, int patch_t,
                       const char* what) {
  ORVB_REQUIRE(a && b, ORVB_EINVAL, "%s: null pointer", what);
  ORVB_REQUIRE(p == 2, ORVB_ESHAPE, "%s: only patch_size 2 is supported (got %d)", what, p);
  ORVB_REQUIRE(B > 0 && F > 0 && C > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, ORVB_ESHAPE,
               "%s: bad geometry B=%d F=%d C=%d H=%d W=%d", what, B, F, C, H, W);
  ORVB_REQUIRE(patch_t == 0 || patch_t == 1 || F % patch_t == 0, ORVB_ESHAPE,
               "%s: frames (%d) must be divisible by patch_size_t (%d)", what, F, patch_t);
  return ORVB_OK;
}

int patchify_launch(const void* x, void* out, int B, int F, int C, int H, int W, int p, int patch_t,
                    cudaStream_t stream) {
  int rc = patch_check(x, out, B, F, C, H, W, p, patch_t, "orvb_patchify");
  if (rc != ORVB_OK) return rc;
  const int pt = patch_t > 0 ? patch_t : 1;
  const long total = static_cast<long>(B) * F * C * (H / 2) * (W / 2);
  patchify_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      static_cast<const bf16*>(x), static_cast<bf16*>(out), B, F, C, H, W, pt);
  ORVB_CHECK_CUDA(cudaGetLastError());
  return ORVB_OK;
}

int unpatchify_launch(const void* y, void* out, int B, int F, int C, int H, int W, int p, int patch_t,
                      cudaStream_t stream) {
  int rc = patch_check(y, out, B, F, C, H, W, p, patch_t, "orvb_unpatchify");
  if (rc != ORVB_OK) return rc;
  const int pt = patch_t > 0 ? patch_t : 1;
  const long total = static_cast<long>(B) * F * C * (H / 2) * (W / 2);
  unpatchify_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      static_cast<const bf16*>(y), static_cast<bf16*>(out), B, F, C, H, W, pt);
  ORVB_CHECK_CUDA(cudaGetLastError());
  return ORVB_OK;
}

// ---------------------------------------------------------------------------------------------------
// Multiview gather (reference MVBlock.forward, cogvideox_control.py:328-331):
//   video  '(b v) (f s) d -> (b f) (v s) d'   and   text  '(b v) n d -> b (v n) d' repeated over f,
// concatenated text-first, i.e. dst[(b f)][v*n_text + n | V*n_text + v*s + i] = src[(b v)][n | n_text + f*s + i].
// ---------------------------------------------------------------------------------------------------
__global__ void mv_gather_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int clips, int V, int F,
                                 int St, int s, int chunks) {
  const int S = St + F * s;         // source sequence length per (b, v)
  const int Smv = V * (St + s);     // destination sequence length per (b, f)
  const long total = static_cast<long>(clips) * F * Smv * chunks;
  const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % chunks);
  long r = idx / chunks;
  const int d = static_cast<int>(r % Smv);
  r /= Smv;
  const int f = static_cast<int>(r % F);
  const int b = static_cast<int>(r / F);
  long srow;
  if (d < V * St) {
    const int v = d / St, n = d - v * St;
    srow = static_cast<long>(b * V + v) * S + n;
  } else {
    const int e = d - V * St;
    const int v = e / s, i = e - v * s;
    srow = static_cast<long>(b * V + v) * S + St + f * s + i;
  }
  dst[(static_cast<long>(b * F + f) * Smv + d) * chunks + c] = src[srow * chunks + c];
}

int mv_gather_launch(const void* src, void* dst, int clips, int views, int frames, int text_len, int tokens,
                     int width, cudaStream_t stream) {
  ORVB_REQUIRE(src && dst && width % 8 == 0, ORVB_EINVAL, "mv_gather: bad arguments");
  const int chunks = width / 8;
  const long total = static_cast<long>(clips) * frames * views * (text_len + tokens) * chunks;
  mv_gather_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      static_cast<const uint4*>(src), static_cast<uint4*>(dst), clips, views, frames, text_len, tokens, chunks);
  ORVB_CHECK_CUDA(cudaGetLastError());
  return ORVB_OK;
}

// ---------------------------------------------------------------------------------------------------
// Sampler step: CFG combine + v-prediction -> x0 + DDIM / DPM-Solver++(2M, SDE) update + bf16 cast, one pass.
// Reference: cogvideox_control.py:1433-1459 and diffusers CogVideoX{DDIM,DPM}Scheduler.step.
// The reference keeps `latents` (and the DPM noise) in bf16 and multiplies them by 0-dim float64 coefficients,
// which rounds those products to bf16 before they meet the fp32 model output; `round_bf16_products` reproduces
// that, so the step is bit-identical to the torch sequence given the same model output.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float rbf(float v) { return __bfloat162float(__float2bfloat16(v)); }

__global__ void sampler_step_kernel(const orvb_sampler_step_args a) {
  const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (i >= a.n) return;
  const bf16* mo = static_cast<const bf16*>(a.model_out);
  bf16* lat = static_cast<bf16*>(a.latents);
  float v;
  if (a.cfg_copies == 2) {
    const float u = __bfloat162float(mo[i]);
    const float c = __bfloat162float(mo[a.n + i]);
    v = __fadd_rn(u, __fmul_rn(a.guidance_scale, __fsub_rn(c, u)));
  } else {
    v = __bfloat162float(mo[i]);
  }
  const float x = __bfloat162float(lat[i]);
  // pred_original_sample = sqrt(alpha_t) * sample - sqrt(1 - alpha_t) * model_output
  const float x0 = __fadd_rn(rbf(__fmul_rn(a.c_x, x)), __fmul_rn(a.c_v, v));
  float d = x0;
  if (a.d_old != 0.f) d = __fadd_rn(__fmul_rn(a.d_cur, x0), __fmul_rn(a.d_old, a.old_x0[i]));
  float prev = __fadd_rn(rbf(__fmul_rn(a.k_x, x)), __fmul_rn(a.k_d, d));
  if (a.noise != nullptr)
    prev = __fadd_rn(prev, rbf(__fmul_rn(a.k_noise, __bfloat162float(static_cast<const bf16*>(a.noise)[i]))));
  if (a.old_x0 != nullptr) a.old_x0[i] = x0;
  const bf16 pb = __float2bfloat16(prev);
  lat[i] = pb;
  if (a.next_input != nullptr) {
    // next iteration's transformer input: cat([latents] * cfg_copies) placed in channels [0, C) of
    // [cfg*B, F, C + C_img, h, w] (cogvideox_control.py:1409-1413); the image-latent channels are written once.
    bf16* ni = static_cast<bf16*>(a.next_input);
    const long chw = static_cast<long>(a.lat_channels) * a.hw;
    const long bf = i / chw;
    const long o = bf * (chw + static_cast<long>(a.img_channels) * a.hw) + (i - bf * chw);
    ni[o] = pb;
    if (a.cfg_copies == 2) ni[o + (a.n / chw) * (chw + static_cast<long>(a.img_channels) * a.hw)] = pb;
  }
}

int sampler_step_launch(const orvb_sampler_step_args* a, cudaStream_t stream) {
  ORVB_REQUIRE(a && a->model_out && a->latents && a->n > 0, ORVB_EINVAL, "orvb_sampler_step: bad arguments");
  ORVB_REQUIRE(a->cfg_copies == 1 || a->cfg_copies == 2, ORVB_EINVAL, "orvb_sampler_step: cfg_copies must be 1 or 2");
  ORVB_REQUIRE(a->d_old == 0.f || a->old_x0 != nullptr, ORVB_EINVAL, "orvb_sampler_step: d_old needs old_x0");
  sampler_step_kernel<<<static_cast<unsigned>((a->n + 255) / 256), 256, 0, stream>>>(*a);
  ORVB_CHECK_CUDA(cudaGetLastError());
  return ORVB_OK;
}

}  // namespace orvb

extern "C" int orvb_ln_modulate(const orvb_ln_args* args, void* stream) {
  int rc = orvb::check_arch();
  if (rc != ORVB_OK) return rc;
  return orvb::ln_modulate_launch(args, static_cast<cudaStream_t>(stream));
}

extern "C" int orvb_skinny_linear(const float* x, const void* w, const void* b, float* y, int32_t rows, int32_t n,
                                  int32_t k, int32_t act, void* stream) {
  int rc = orvb::check_arch();
  if (rc != ORVB_OK) return rc;
  ORVB_REQUIRE(w && y, ORVB_EINVAL, "orvb_skinny_linear: null pointer");
  orvb::SkinnyJob job{static_cast<const orvb::bf16*>(w), static_cast<const orvb::bf16*>(b), y};
  return orvb::skinny_linear_launch(x, job, nullptr, 1, rows, n, k, act, static_cast<cudaStream_t>(stream));
}

extern "C" int orvb_patchify(const void* x, void* out, int32_t b, int32_t f, int32_t c, int32_t h, int32_t w,
                             int32_t p, int32_t patch_t, void* stream) {
  int rc = orvb::check_arch();
  if (rc != ORVB_OK) return rc;
  return orvb::patchify_launch(x, out, b, f, c, h, w, p, patch_t, static_cast<cudaStream_t>(stream));
}

extern "C" int orvb_unpatchify(const void* y, void* out, int32_t b, int32_t f, int32_t c, int32_t h, int32_t w,
                               int32_t p, int32_t patch_t, void* stream) {
  int rc = orvb::check_arch();
  if (rc != ORVB_OK) return rc;
  return orvb::unpatchify_launch(y, out, b, f, c, h, w, p, patch_t, static_cast<cudaStream_t>(stream));
}

extern "C" int orvb_sampler_step(const orvb_sampler_step_args* a, void* stream) {
  int rc = orvb::check_arch();
  if (rc != ORVB_OK) return rc;
  return orvb::sampler_step_launch(a, static_cast<cudaStream_t>(stream));
}
