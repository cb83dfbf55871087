// Non-causal multi-head attention (head_dim 64) for sm_100a with both contractions on tcgen05:
//   S   = Q K^T          (128 x 128 x 64,  A = Q smem K-major, B = K smem K-major, D in TMEM)
//   O_s += P_s V_s       (128 x 64 x 64 per column half s, A = P smem K-major (written by the softmax warps),
//                         B = V smem MN-major, D accumulates in TMEM)
// replacing F.scaled_dot_product_attention at reference orv/models/cogvideox_control.py:256-258.
//
// One CTA per (PAIR of 128-query tiles, head, batch), one CTA per SM.  Warp roles (576 threads):
//   warps 0-7   query tile 0: warps 0-3 softmax stream A (key columns [0,64) of every 128-key tile, one query row
//               per thread), warps 4-7 stream B (key columns [64,128))
//   warps 8-15  query tile 1, same split
//   warp  16    TMA producer (both Q tiles once, K/V triple-buffered, shared by the two query tiles)
//   warp  17    TMEM allocator + MMA issuer, alternating between the two query tiles
// The two streams of a tile are independent flash-attention accumulations (own running max / sum and own TMEM output
// accumulator, combined once at the end like a split-KV reduction), which doubles the number of softmax warps
// hiding MUFU / TMEM latency without any per-tile cross-thread reduction.  Output accumulators stay in TMEM; the
// running max is only raised when a tile exceeds it by more than 2^8 (lazy rescale), so the TMEM read-modify-write
// correction is rare.  Rows past seq_len are zero-filled by TMA and masked in the softmax.
//
// Measured structure notes (profiles/r01_attention_notes.md): at head_dim 64 the kernel is bound by the per-tile
// chain  QK^T -> s_full hop -> max pass -> exp pass (MUFU, 16/clk/SM) -> p_full hop -> MMA issue, not by the tensor
// pipe; two query tiles per SM give two chains in flight.
#include "common.cuh"
#include "pointwise.cuh"
#include "ptx.cuh"

namespace orvb {

constexpr int ATT_BQ = 128;
constexpr int ATT_BK = 128;
constexpr int ATT_D = 64;
constexpr int ATT_THREADS = 576;
constexpr int ATT_STAGES = 3;
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;  // 16 KB: one [128 x 64] bf16 tile
// per query tile: [Q 16 KB][P 32 KB] contiguous (reused for the end-of-kernel stream exchange), then K/V stages
constexpr int ATT_QP_BYTES = 3 * ATT_TILE_BYTES;
constexpr int ATT_SMEM_BYTES = 2 * ATT_QP_BYTES + ATT_STAGES * 2 * ATT_TILE_BYTES + 256;
constexpr int ATT_TMEM_COLS = 512;  // S_t: [t*128, +128)   O_{t,stream}: [256 + t*128 + stream*64, +64)
constexpr int ATT_XCH_STRIDE = 67;  // floats per row of the end-of-kernel stream exchange (conflict-free)
constexpr float ATT_RESCALE_THRESHOLD = 8.0f;  // log2 units

struct AttDev {
  bf16* out;
  int seq_len, heads, dim;  // dim = heads * 64
  float scale_log2;         // softmax scale * log2(e)
  int q_row0, q_rows;       // queries = rows [q_row0, q_row0 + q_rows) of every sequence; output is compact
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap tma_qkv, const AttDev p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // [Q0 | P0 (2 sub-tiles)] [Q1 | P1] [K stages] [V stages] [barriers]
  uint8_t* sK = smem + 2 * ATT_QP_BYTES;                // [stage]
  uint8_t* sV = sK + ATT_STAGES * ATT_TILE_BYTES;       // [stage]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATT_STAGES * ATT_TILE_BYTES);
  uint64_t* q_full = bars;                  // both Q tiles
  uint64_t* k_full = bars + 1;              // [3]
  uint64_t* v_full = bars + 4;              // [3]
  uint64_t* kv_empty = bars + 7;            // [3]
  uint64_t* s_full = bars + 10;             // [2] per query tile
  uint64_t* p_full = bars + 12;             // [2]
  uint64_t* o_full = bars + 14;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_pair = blockIdx.x;
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int n_kv = (p.seq_len + ATT_BK - 1) / ATT_BK;

  if ((smem_u32(smem) & 1023u) != 0) __trap();  // 128B-swizzle atoms need 1024-byte aligned tiles

  if (warp == 16 && lane == 0) {
    tma_prefetch_desc(&tma_qkv);
    mbar_init(q_full, 1);
    for (int i = 0; i < ATT_STAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 256);
      mbar_init(&o_full[t], 1);
    }
    fence_barrier_init();
  }
  if (warp == 17) {
    tmem_alloc(tmem_slot, ATT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 16) {
    // ======================================= TMA producer =======================================
    if (lane == 0) {
      const int q_col = head * ATT_D;
      const int k_col = p.dim + head * ATT_D;
      const int v_col = 2 * p.dim + head * ATT_D;
      mbar_expect_tx(q_full, 2 * ATT_TILE_BYTES);
      tma_load_3d(smem, &tma_qkv, q_full, q_col, p.q_row0 + (2 * q_pair) * ATT_BQ, batch);
      tma_load_3d(smem + ATT_QP_BYTES, &tma_qkv, q_full, q_col, p.q_row0 + (2 * q_pair + 1) * ATT_BQ, batch);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % ATT_STAGES;
        const uint32_t ph = static_cast<uint32_t>((j / ATT_STAGES) & 1);
        mbar_wait(&kv_empty[st], ph ^ 1);
        mbar_expect_tx(&k_full[st], ATT_TILE_BYTES);
        tma_load_3d(sK + st * ATT_TILE_BYTES, &tma_qkv, &k_full[st], k_col, j * ATT_BK, batch);
        mbar_expect_tx(&v_full[st], ATT_TILE_BYTES);
        tma_load_3d(sV + st * ATT_TILE_BYTES, &tma_qkv, &v_full[st], v_col, j * ATT_BK, batch);
      }
    }
  } else if (warp == 17) {
    // ======================================= MMA issuer =========================================
    // (One lane owns the loop here: with this kernel's short MMAs the warp-uniform form used in gemm.cu issues
    //  faster but lets the two query tiles fall into lockstep, which measured slower end to end.)
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, 0, 0);  // A, B K-major
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);   // B (= V) MN-major
      auto issue_qk = [&](int t, int j) {
        const int st = j % ATT_STAGES;
        mbar_wait(&k_full[st], static_cast<uint32_t>((j / ATT_STAGES) & 1));
        tc_fence_after();
        const uint64_t q_desc = umma_desc_sw128(smem_u32(smem + t * ATT_QP_BYTES));
        const uint64_t k_desc = umma_desc_sw128(smem_u32(sK + st * ATT_TILE_BYTES));
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k)
          umma_f16_ss(tmem_base + static_cast<uint32_t>(t * 128), q_desc + static_cast<uint64_t>(k * 2),
                      k_desc + static_cast<uint64_t>(k * 2), idesc_qk, static_cast<uint32_t>(k != 0));
        tc_commit(&s_full[t]);
      };
      mbar_wait(q_full, 0);
      issue_qk(0, 0);
      issue_qk(1, 0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % ATT_STAGES;
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&p_full[t], static_cast<uint32_t>(j & 1));  // P_{t,j} in smem; S_t read; O_t rescaled if needed
          tc_fence_after();
          // S_t is free again: queue the next QK^T of this query tile first, so its softmax warps can go on while
          // the tensor core still works on the P V products.
          if (j + 1 < n_kv) issue_qk(t, j + 1);
          mbar_wait(&v_full[st], static_cast<uint32_t>((j / ATT_STAGES) & 1));
          tc_fence_after();
          const uint64_t p_desc = umma_desc_sw128(smem_u32(smem + t * ATT_QP_BYTES + ATT_TILE_BYTES));
          const uint64_t v_desc = umma_desc_sw128(smem_u32(sV + st * ATT_TILE_BYTES));
#pragma unroll
          for (int s = 0; s < 2; ++s) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // A: P sub-tile s (16 KB apart), 32 bytes per K step.  B: keys s*64 + k*16 .. +16 = rows of 128 B.
              const uint64_t a = p_desc + static_cast<uint64_t>((s * ATT_TILE_BYTES + k * 32) >> 4);
              const uint64_t b = v_desc + static_cast<uint64_t>(((s * 64 + k * 16) * 128) >> 4);
              umma_f16_ss(tmem_base + static_cast<uint32_t>(256 + t * 128 + s * 64), a, b, idesc_pv,
                          static_cast<uint32_t>((j | k) != 0));
            }
          }
          tc_commit(&o_full[t]);
        }
        tc_commit(&kv_empty[st]);  // K_j / V_j consumed by both query tiles
      }
    }
  } else {
    // ======================================= softmax streams ====================================
    const int t = warp >> 3;                         // query tile of this warp
    const int stream = (warp >> 2) & 1;              // 0: key columns [0,64) of each tile, 1: [64,128)
    const int row_in_tile = (warp & 3) * 32 + lane;  // TMEM lane == query row
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t my_s = tmem_base + lane_off + static_cast<uint32_t>(t * 128 + stream * 64);
    const uint32_t my_o = tmem_base + lane_off + static_cast<uint32_t>(256 + t * 128 + stream * 64);
    uint64_t* my_s_full = &s_full[t];
    uint64_t* my_p_full = &p_full[t];
    uint64_t* my_o_full = &o_full[t];
    float m_run = -INFINITY;  // running (lazy) max in the scaled log2 domain
    float l_run = 0.f;
    uint8_t* p_row = smem + t * ATT_QP_BYTES + ATT_TILE_BYTES + stream * ATT_TILE_BYTES + row_in_tile * 128;
    const int sw = row_in_tile & 7;

    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(my_s_full, static_cast<uint32_t>(j & 1));
      tc_fence_after();
      const int valid = p.seq_len - (j * ATT_BK + stream * 64);  // my columns >= valid are padding
      // ---- pass 1: row max over my 64 columns ----
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 64; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(my_s + static_cast<uint32_t>(c), r);
        tmem_ld_wait();
        if (valid >= 64) {
          float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;  // 4 chains: no serial FMNMX latency
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            m0 = fmaxf(m0, __uint_as_float(r[i]));
            m1 = fmaxf(m1, __uint_as_float(r[i + 1]));
            m2 = fmaxf(m2, __uint_as_float(r[i + 2]));
            m3 = fmaxf(m3, __uint_as_float(r[i + 3]));
          }
          mx = fmaxf(mx, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c + i < valid) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
      }
      const float m_tile = mx * p.scale_log2;
      // PV_{t,j-1} must have retired before O is corrected or the P buffer is overwritten
      if (j > 0) {
        mbar_wait(my_o_full, static_cast<uint32_t>((j - 1) & 1));
        tc_fence_after();
      }
      // ---- lazy rescale: raise the running max only when this tile exceeds it by > 2^8 ----
      const bool grow = m_tile > m_run + ATT_RESCALE_THRESHOLD;
      if (__any_sync(0xffffffffu, grow)) {
        const float m_new = grow ? m_tile : m_run;
        const float alpha = (m_new == -INFINITY) ? 1.0f : ex2(m_run - m_new);
        if (j > 0) {
#pragma unroll 1
          for (int c = 0; c < 64; c += 32) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(my_o + static_cast<uint32_t>(c), o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32b_x32(my_o + static_cast<uint32_t>(c), o);
          }
          tmem_st_wait();
        }
        l_run *= alpha;
        m_run = m_new;
      }

      // ---- pass 2: P = exp2(S * scale_log2 - m_run), bf16, K-major / 128B-swizzled into my smem sub-tile ----
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
      const float neg_m = -m_run;
#pragma unroll 1
      for (int c = 0; c < 64; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(my_s + static_cast<uint32_t>(c), r);
        tmem_ld_wait();
        float pv[32];
        if (valid >= 64) {  // tile-uniform fast path: no per-element masking instructions
#pragma unroll
          for (int i = 0; i < 32; ++i) pv[i] = ex2(fmaf(__uint_as_float(r[i]), p.scale_log2, neg_m));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            pv[i] = (c + i < valid) ? ex2(fmaf(__uint_as_float(r[i]), p.scale_log2, neg_m)) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          l0 += pv[i];
          l1 += pv[i + 1];
          l2 += pv[i + 2];
          l3 += pv[i + 3];
        }
        const int chunk0 = c >> 3;  // first 16-byte chunk of this 32-column group inside the 128-byte row
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 v;
          v.x = pack_bf16(pv[q * 8 + 0], pv[q * 8 + 1]);
          v.y = pack_bf16(pv[q * 8 + 2], pv[q * 8 + 3]);
          v.z = pack_bf16(pv[q * 8 + 4], pv[q * 8 + 5]);
          v.w = pack_bf16(pv[q * 8 + 6], pv[q * 8 + 7]);
          *reinterpret_cast<uint4*>(p_row + (((chunk0 + q) ^ sw) << 4)) = v;
        }
      }
      l_run += (l0 + l1) + (l2 + l3);
      fence_proxy_async_smem();  // make the P stores visible to the tensor-core (async) proxy
      tc_fence_before();         // order our TMEM loads / stores before the MMAs that follow
      mbar_arrive(my_p_full);
    }

    // ---- combine the two streams of this query tile, normalise, store ----
    mbar_wait(my_o_full, static_cast<uint32_t>((n_kv - 1) & 1));
    tc_fence_after();
    // this tile's Q/P smem (48 KB) is idle once its last PV retired; O moves in 32-column halves
    float* xch = reinterpret_cast<float*>(smem + t * ATT_QP_BYTES) + row_in_tile * ATT_XCH_STRIDE;
    if (stream == 1) {
      xch[0] = m_run;
      xch[1] = l_run;
#pragma unroll 1
      for (int c = 0; c < ATT_D; c += 32) {
        uint32_t o[32];
        tmem_ld_32x32b_x32(my_o + static_cast<uint32_t>(c), o);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < 32; ++d) xch[2 + c + d] = __uint_as_float(o[d]);
      }
    }
    named_bar_sync(1 + t, 256);
    if (stream == 0) {
      const float m_b = xch[0], l_b = xch[1];
      const float m = fmaxf(m_run, m_b);
      const float wa = (l_run > 0.f) ? ex2(m_run - m) : 0.f;
      const float wb = (l_b > 0.f) ? ex2(m_b - m) : 0.f;
      const float inv = 1.0f / (l_run * wa + l_b * wb);
      const float ca = wa * inv, cb = wb * inv;
      const int q_row = (2 * q_pair + t) * ATT_BQ + row_in_tile;
      bf16* op = p.out + (static_cast<size_t>(batch) * p.q_rows + q_row) * p.dim + head * ATT_D;
#pragma unroll 1
      for (int c = 0; c < ATT_D; c += 32) {
        uint32_t o[32];
        tmem_ld_32x32b_x32(my_o + static_cast<uint32_t>(c), o);
        tmem_ld_wait();
        if (q_row < p.q_rows) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float f[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) f[u] = __uint_as_float(o[q * 8 + u]) * ca + xch[2 + c + q * 8 + u] * cb;
            uint4 v;
            v.x = pack_bf16(f[0], f[1]);
            v.y = pack_bf16(f[2], f[3]);
            v.z = pack_bf16(f[4], f[5]);
            v.w = pack_bf16(f[6], f[7]);
            *reinterpret_cast<uint4*>(op + c + q * 8) = v;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ATT_TMEM_COLS);
  }
}

int attention_launch(const void* qkv, void* out, int batch, int seq_len, int heads, float scale, int q_row0,
                     int q_rows, cudaStream_t stream) {
  ORVB_REQUIRE(qkv && out, ORVB_EINVAL, "orvb_attention_bf16: null pointer");
  ORVB_REQUIRE(batch > 0 && seq_len > 0 && heads > 0, ORVB_ESHAPE, "orvb_attention_bf16: empty problem");
  if (q_rows <= 0) {
    q_row0 = 0;
    q_rows = seq_len;
  }
  ORVB_REQUIRE(q_row0 >= 0 && q_row0 + q_rows <= seq_len, ORVB_ESHAPE, "orvb_attention_bf16: query window out of range");
  const int dim = heads * ATT_D;
  CUtensorMap tm;
  int rc = make_tmap_3d_bf16(&tm, qkv, batch, seq_len, 3 * dim, 3 * dim, static_cast<uint64_t>(seq_len) * 3 * dim,
                             ATT_BK, ATT_D);
  if (rc != ORVB_OK) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    ORVB_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
    attr_set = true;
  }
  AttDev p;
  p.out = static_cast<bf16*>(out);
  p.seq_len = seq_len;
  p.heads = heads;
  p.dim = dim;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.q_row0 = q_row0;
  p.q_rows = q_rows;
  dim3 grid((q_rows + 2 * ATT_BQ - 1) / (2 * ATT_BQ), heads, batch);
  attention_kernel<<<grid, ATT_THREADS, ATT_SMEM_BYTES, stream>>>(tm, p);
  ORVB_CHECK_CUDA(cudaGetLastError());
  return ORVB_OK;
}

}  // namespace orvb

extern "C" int orvb_attention_bf16(const void* qkv, void* out, int32_t batch, int32_t seq_len, int32_t heads,
                                   float scale, void* stream) {
  int rc = orvb::check_arch();
  if (rc != ORVB_OK) return rc;
  return orvb::attention_launch(qkv, out, batch, seq_len, heads, scale, 0, seq_len, static_cast<cudaStream_t>(stream));
}

