// Pieces shared by the GEMM kernels (gemm.cu) and the implicit-GEMM convolution built on the same CTA-pair mainloop
// (conv.cu): the device-side problem descriptor, tile constants and the fused epilogue over one 64-column unit.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace orvb {

struct GemmDev {
  int M, N, K;
  bf16* out;
  int ldo;
  const bf16* bias;
  int src_rows, dst_rows, dst_offset;
  int mv_tokens, mv_frames, mv_views;
  const bf16* resid;
  int ldr, resid_mod, resid_views, resid_view_stride;
  const float* gate;
  int gate_ld, gate_text_off, gate_video_off;
  orvb_rowmap rm;
  int qk_dim;
  const bf16 *qw, *qb, *kw, *kb;
  float qk_eps;
  const float *rope_cos, *rope_sin;
  int num_m_tiles, num_n_tiles;
  int tma_store;  // pair kernel: stage full 64-column units in shared memory and write them with TMA
  int out_f32;    // test mode: `out` is fp32, written before the bf16 rounding (direct stores only)
  const int* grp_off;  // optional device scalar added to the modulation-group index (schedule slice of this step)
  // ---- CTA-pair kernel shared-memory plan (gemm.cu; the convolution kernel has its own fixed plan) ----
  int stage_bytes;      // bytes per operand pipeline stage (A 16 KB + this CTA's half of B, rounded up to 1 KB)
  int out_stage_bytes;  // output staging area behind the stages
  int fast_resid;       // gated-residual epilogue with the residual tile prefetched by TMA (see gemm2_bf16_kernel)
  int k_wrap;           // > 0: W has k_wrap columns, walked cyclically along K (orvb_gemm_args.k_wrap)
  // CTA-pair kernel tile list: `tile_end` virtual tile indices, cluster c takes c, c + C, ...  With rem_width == 0 every
  // tile is `bn` wide (the last one of a row may hang over N).  With rem_width > 0 the N range is cut into n_full tiles
  // of bn columns plus ONE narrower tile of rem_width columns per 256-row block; the narrow tiles are numbered so that
  // they fall on the clusters that got one full tile less (g2_tile): QKV of config 2 (N = 5760) runs 4 x 256 columns per
  // cluster instead of 6 x 192.
  int tile_end, n_full, rem_width;
};

struct G2Tile {
  int m_blk, n0, w;  // 256-row block, first column, width; w == 0: this virtual index holds no tile
};
__host__ __device__ __forceinline__ G2Tile g2_tile(const GemmDev& p, int v, int bn, int num_clusters) {
  G2Tile t;
  const int mt = p.num_m_tiles;
  const int full = (p.rem_width > 0) ? mt * p.n_full : p.tile_end;
  if (v < full) {
    t.m_blk = v % mt;
    t.n0 = (v / mt) * bn;
    t.w = bn;
    return t;
  }
  const int wv = v - full;
  const int row = wv / num_clusters, col = wv - row * num_clusters;
  const int n_light = num_clusters - full % num_clusters;  // clusters with one full tile less (all of them if it divides)
  const int k = row * n_light + col;
  t.m_blk = k;
  t.n0 = p.n_full * bn;
  t.w = (col < n_light && k < mt) ? p.rem_width : 0;
  return t;
}

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle span
constexpr int GEMM_THREADS = 256;

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 192 ? 5 : (BN >= 128 ? 6 : 8));
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ int row_group(const orvb_rowmap& rm, int row, int* s_out) {
  if (rm.seq_len <= 0) {
    *s_out = row;
    return 0;
  }
  int b = row / rm.seq_len;
  int s = row - b * rm.seq_len;
  *s_out = s;
  int g = (s < rm.text_len || rm.tokens_per_group <= 0) ? 0 : 1 + (s - rm.text_len) / rm.tokens_per_group;
  return b * rm.groups_per_batch + g;
}

// Epilogue over one 64-column unit held by one thread (one output row).
// GATED = false compiles the gate operand out (the convolution kernel never has one and needs the registers).
template <int EPI, bool GATED = true>
__device__ __forceinline__ void epilogue_unit(const GemmDev& p, float (&v)[64], int row, int n0, int ncols,
                                              uint8_t* stage_row = nullptr, int sw = 0) {
  // ---- bias --------------------------------------------------------------------------------------
  if (p.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j * 8 < ncols) {
        uint4 bb = *reinterpret_cast<const uint4*>(p.bias + n0 + j * 8);
        v[j * 8 + 0] += bf16_lo(bb.x); v[j * 8 + 1] += bf16_hi(bb.x);
        v[j * 8 + 2] += bf16_lo(bb.y); v[j * 8 + 3] += bf16_hi(bb.y);
        v[j * 8 + 4] += bf16_lo(bb.z); v[j * 8 + 5] += bf16_hi(bb.z);
        v[j * 8 + 6] += bf16_lo(bb.w); v[j * 8 + 7] += bf16_hi(bb.w);
      }
    }
  }
  int out_row = row;
  if (p.mv_tokens > 0) {
    // '(b f) (v s) -> (b v) (text | f s)': row = ((b*F + f)*V + v)*s + i  ->  (b*V + v)*dst_rows + dst_offset + f*s + i
    const int blk = row / p.mv_tokens, i = row - blk * p.mv_tokens;
    const int v = blk % p.mv_views, bf = blk / p.mv_views;
    const int f = bf % p.mv_frames, b = bf / p.mv_frames;
    out_row = (b * p.mv_views + v) * p.dst_rows + p.dst_offset + f * p.mv_tokens + i;
  } else if (p.src_rows > 0) {
    int q = row / p.src_rows;
    out_row = q * p.dst_rows + p.dst_offset + (row - q * p.src_rows);
  }

  if (EPI == ORVB_EPI_GELU) {
#pragma unroll
    for (int j = 0; j < 64; j += 2) gelu_tanh2(v[j], v[j + 1]);
  }

  if (EPI == ORVB_EPI_QKV) {
    int s;
    (void)row_group(p.rm, row, &s);
    if (n0 < 2 * p.qk_dim) {  // Q or K head: LayerNorm over the 64 values of this head
      float mean = 0.f;
#pragma unroll
      for (int j = 0; j < 64; ++j) mean += v[j];
      mean *= (1.0f / 64.0f);
      float var = 0.f;
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        float d = v[j] - mean;
        var += d * d;
      }
      var *= (1.0f / 64.0f);
      float rstd = rsqrtf(var + p.qk_eps);
      const bf16* w = (n0 < p.qk_dim) ? p.qw : p.kw;
      const bf16* b = (n0 < p.qk_dim) ? p.qb : p.kb;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint4 ww = *reinterpret_cast<const uint4*>(w + j * 8);
        uint4 bb = *reinterpret_cast<const uint4*>(b + j * 8);
        const uint32_t wv[4] = {ww.x, ww.y, ww.z, ww.w};
        const uint32_t bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          v[j * 8 + 2 * t] = (v[j * 8 + 2 * t] - mean) * rstd * bf16_lo(wv[t]) + bf16_lo(bv[t]);
          v[j * 8 + 2 * t + 1] = (v[j * 8 + 2 * t + 1] - mean) * rstd * bf16_hi(wv[t]) + bf16_hi(bv[t]);
        }
      }
      if (p.rope_cos != nullptr && s >= p.rm.text_len) {
        // The reference rounds the LayerNorm output to bf16 before apply_rotary_emb upcasts it again.
        const float* cs = p.rope_cos + static_cast<size_t>(s - p.rm.text_len) * 64;
        const float* sn = p.rope_sin + static_cast<size_t>(s - p.rm.text_len) * 64;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float4 c4 = *reinterpret_cast<const float4*>(cs + j * 4);
          float4 s4 = *reinterpret_cast<const float4*>(sn + j * 4);
          float x0 = v[j * 4 + 0], x1 = v[j * 4 + 1], x2 = v[j * 4 + 2], x3 = v[j * 4 + 3];
          v[j * 4 + 0] = x0 * c4.x - x1 * s4.x;
          v[j * 4 + 1] = x1 * c4.y + x0 * s4.y;
          v[j * 4 + 2] = x2 * c4.z - x3 * s4.z;
          v[j * 4 + 3] = x3 * c4.w + x2 * s4.w;
        }
      }
    }
  }

  if (EPI == ORVB_EPI_GATE_RESID) {
    // out = resid + gate * (acc + bias): ONE fused multiply-add per element when both operands are present (the
    // TMA-prefetched epilogue of the CTA-pair kernel forms the same fma, so the two paths agree bit for bit).
    const float* gp = nullptr;
    if (GATED && p.gate != nullptr) {
      int s;
      int g = row_group(p.rm, out_row, &s);  // modulation group of the DESTINATION row
      if (p.grp_off != nullptr) g += *p.grp_off;
      int is_text = (p.rm.seq_len > 0) ? (s < p.rm.text_len) : 0;
      gp = p.gate + static_cast<size_t>(g) * p.gate_ld + (is_text ? p.gate_text_off : p.gate_video_off) + n0;
    }
    const bf16* rp = nullptr;
    if (p.resid != nullptr) {
      size_t rrow;
      if (p.resid_mod > 0) {
        int q = row / p.resid_mod;
        rrow = static_cast<size_t>(row - q * p.resid_mod) +
               static_cast<size_t>(p.resid_views > 1 ? (q % p.resid_views) : 0) * p.resid_view_stride;
      } else {
        rrow = static_cast<size_t>(out_row);
      }
      rp = p.resid + rrow * p.ldr + n0;
    }
    if (gp != nullptr && rp != nullptr) {
      // half a row of residual first (four independent 16-byte loads in flight together), then gate pieces + fma
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint4 rr[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if ((h * 4 + q) * 8 < ncols) rr[q] = *reinterpret_cast<const uint4*>(rp + (h * 4 + q) * 8);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int j = h * 4 + q;
          if (j * 8 < ncols) {
            const float4 g0 = *reinterpret_cast<const float4*>(gp + j * 8);
            const float4 g1 = *reinterpret_cast<const float4*>(gp + j * 8 + 4);
            v[j * 8 + 0] = __fmaf_rn(v[j * 8 + 0], g0.x, bf16_lo(rr[q].x)); v[j * 8 + 1] = __fmaf_rn(v[j * 8 + 1], g0.y, bf16_hi(rr[q].x));
            v[j * 8 + 2] = __fmaf_rn(v[j * 8 + 2], g0.z, bf16_lo(rr[q].y)); v[j * 8 + 3] = __fmaf_rn(v[j * 8 + 3], g0.w, bf16_hi(rr[q].y));
            v[j * 8 + 4] = __fmaf_rn(v[j * 8 + 4], g1.x, bf16_lo(rr[q].z)); v[j * 8 + 5] = __fmaf_rn(v[j * 8 + 5], g1.y, bf16_hi(rr[q].z));
            v[j * 8 + 6] = __fmaf_rn(v[j * 8 + 6], g1.z, bf16_lo(rr[q].w)); v[j * 8 + 7] = __fmaf_rn(v[j * 8 + 7], g1.w, bf16_hi(rr[q].w));
          }
        }
      }
    } else if (gp != nullptr) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (j * 4 < ncols) {
          float4 g4 = *reinterpret_cast<const float4*>(gp + j * 4);
          v[j * 4 + 0] = __fmul_rn(v[j * 4 + 0], g4.x); v[j * 4 + 1] = __fmul_rn(v[j * 4 + 1], g4.y);
          v[j * 4 + 2] = __fmul_rn(v[j * 4 + 2], g4.z); v[j * 4 + 3] = __fmul_rn(v[j * 4 + 3], g4.w);
        }
      }
    } else if (rp != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j * 8 < ncols) {
          uint4 rr = *reinterpret_cast<const uint4*>(rp + j * 8);
          v[j * 8 + 0] = __fadd_rn(v[j * 8 + 0], bf16_lo(rr.x)); v[j * 8 + 1] = __fadd_rn(v[j * 8 + 1], bf16_hi(rr.x));
          v[j * 8 + 2] = __fadd_rn(v[j * 8 + 2], bf16_lo(rr.y)); v[j * 8 + 3] = __fadd_rn(v[j * 8 + 3], bf16_hi(rr.y));
          v[j * 8 + 4] = __fadd_rn(v[j * 8 + 4], bf16_lo(rr.z)); v[j * 8 + 5] = __fadd_rn(v[j * 8 + 5], bf16_hi(rr.z));
          v[j * 8 + 6] = __fadd_rn(v[j * 8 + 6], bf16_lo(rr.w)); v[j * 8 + 7] = __fadd_rn(v[j * 8 + 7], bf16_hi(rr.w));
        }
      }
    }
  }

  // ---- store ---------------------------------------------------------------------------------------
  if (stage_row != nullptr) {
    // this thread's 128-byte row of a [32 rows x 64 cols] 128B-swizzled staging tile; the warp's elected lane then
    // writes the whole tile with one TMA store (full 128-byte lines instead of 32 scattered 16-byte pieces per STG)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint4 o;
      o.x = pack_bf16(v[j * 8 + 0], v[j * 8 + 1]);
      o.y = pack_bf16(v[j * 8 + 2], v[j * 8 + 3]);
      o.z = pack_bf16(v[j * 8 + 4], v[j * 8 + 5]);
      o.w = pack_bf16(v[j * 8 + 6], v[j * 8 + 7]);
      *reinterpret_cast<uint4*>(stage_row + ((j ^ sw) << 4)) = o;
    }
    return;
  }
  if (p.out_f32) {
    float* of = reinterpret_cast<float*>(p.out) + static_cast<size_t>(out_row) * p.ldo + n0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (j * 4 < ncols) *reinterpret_cast<float4*>(of + j * 4) = make_float4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
    }
    return;
  }
  bf16* op = p.out + static_cast<size_t>(out_row) * p.ldo + n0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (j * 8 < ncols) {
      uint4 o;
      o.x = pack_bf16(v[j * 8 + 0], v[j * 8 + 1]);
      o.y = pack_bf16(v[j * 8 + 2], v[j * 8 + 3]);
      o.z = pack_bf16(v[j * 8 + 4], v[j * 8 + 5]);
      o.w = pack_bf16(v[j * 8 + 6], v[j * 8 + 7]);
      *reinterpret_cast<uint4*>(op + j * 8) = o;
    }
  }
}


// ---- CTA-pair kernel constants (see gemm.cu) ----
constexpr int G2_STAGES = 6;
constexpr int G2_A_BYTES = BM * BK * 2;           // 16 KB: this CTA's 128 rows of A
constexpr int G2_B_BYTES = 128 * BK * 2;          // up to 16 KB: this CTA's BN/2 rows of B
constexpr int G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES;
constexpr int G2_OUT_STAGE_BYTES = 8 * 32 * 128;  // 32 KB of [32 x 64] bf16 output staging tiles, split over the epilogue warps
constexpr int G2_BAR_BYTES = 512;                 // mbarriers + the TMEM slot
constexpr int G2_SMEM_BYTES = G2_STAGES * G2_STAGE_BYTES + G2_OUT_STAGE_BYTES + 1024 + G2_BAR_BYTES;
constexpr int G2_SMEM_MAX = 232448;               // 227 KB: the most a CTA can own
constexpr int G2_GS_SLOT = 768;                   // fast gated-residual epilogue, per 64-column unit: gate rows A | B, bias (64 fp32 each)
constexpr int G2_ACC_COLS = 256;                  // TMEM columns per accumulator buffer (2 buffers = 512)

// Epilogue warps per CTA: eight — warps e and e + 4 share a TMEM lane quarter and take alternate 64-column units, which
// halves the epilogue of the LAST tile of a cluster, the only one that is not hidden behind a mainloop (the attn-out
// and FF2 GEMMs have just two tiles per cluster).  Measured: attn-out 41 -> 36 us, fused FF1 = plain FF1.  The
// register-heavy epilogues still fit (QKV LayerNorm 168, GELU 162 registers at 384 threads); a value of 4 keeps the
// code path for an epilogue that would not.
__host__ __device__ constexpr int g2_epi_warps(int epi) {
  (void)epi;
  return 8;
}
__host__ __device__ constexpr int g2_threads(int epi) { return (4 + g2_epi_warps(epi)) * 32; }

}  // namespace orvb
