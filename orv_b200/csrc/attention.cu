// Non-causal multi-head attention (head_dim 64) for sm_100a with both contractions on tcgen05:
//   S   = Q K^T          (128 x 128 x 64,  A = Q smem K-major, B = K smem K-major, D in TMEM)
//   O_s += P_s V_s       (128 x 64 x 64 per column half s, A = P smem K-major (written by the softmax warps),
//                         B = V smem MN-major, D accumulates in TMEM)
// replacing F.scaled_dot_product_attention at reference orv/models/cogvideox_control.py:256-258.
//
// One CTA per (128-query tile, head, batch); two CTAs are co-resident per SM.  Warp roles (320 threads):
//   warps 0-3  softmax stream A: key columns [0,64) of every 128-key tile, one query row per thread
//   warps 4-7  softmax stream B: key columns [64,128)
//   warp  8    TMA producer (Q once, K/V double-buffered through one 3-D tensor map over the packed QKV buffer)
//   warp  9    TMEM allocator + single-thread MMA issuer
// The two streams are independent flash-attention accumulations (own running max / sum and own TMEM output
// accumulator, combined once at the end like a split-KV reduction), which doubles the number of softmax warps
// hiding MUFU / TMEM latency without any per-tile cross-thread reduction.  Output accumulators stay in TMEM; the
// running max is only raised when a tile exceeds it by more than 2^8 (lazy rescale), so the TMEM read-modify-write
// correction is rare.
//
// Rows past seq_len are zero-filled by TMA and masked in the softmax.
#include "common.cuh"
#include "pointwise.cuh"
#include "ptx.cuh"

namespace orvb {

constexpr int ATT_BQ = 128;
constexpr int ATT_BK = 128;
constexpr int ATT_D = 64;
constexpr int ATT_THREADS = 320;
constexpr int ATT_STAGES = 2;
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;  // 16 KB: one [128 x 64] bf16 tile
constexpr int ATT_SMEM_BYTES = ATT_TILE_BYTES /*Q*/ + ATT_STAGES * 2 * ATT_TILE_BYTES /*K,V*/ + 2 * ATT_TILE_BYTES /*P*/ + 256;
constexpr int ATT_TMEM_COLS = 256;  // S: [0,128)  O_A: [128,192)  O_B: [192,256)
constexpr int ATT_XCH_STRIDE = 67;  // floats per row of the end-of-kernel stream exchange (conflict-free)
constexpr float ATT_RESCALE_THRESHOLD = 8.0f;  // log2 units

struct AttDev {
  bf16* out;
  int seq_len, heads, dim;  // dim = heads * 64
  float scale_log2;         // softmax scale * log2(e)
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_kernel(const __grid_constant__ CUtensorMap tma_qkv, const AttDev p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + ATT_TILE_BYTES;                    // [stage]
  uint8_t* sV = sK + ATT_STAGES * ATT_TILE_BYTES;       // [stage]
  uint8_t* sP = sV + ATT_STAGES * ATT_TILE_BYTES;       // two [128 x 64] K-major sub-tiles (stream A, stream B)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * ATT_TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;              // [2]
  uint64_t* v_full = bars + 3;              // [2]
  uint64_t* kv_empty = bars + 5;            // [2]
  uint64_t* s_full = bars + 7;
  uint64_t* p_full = bars + 8;
  uint64_t* o_full = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x;
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int n_kv = (p.seq_len + ATT_BK - 1) / ATT_BK;

  if ((smem_u32(smem) & 1023u) != 0) __trap();  // 128B-swizzle atoms need 1024-byte aligned tiles

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tma_qkv);
    mbar_init(q_full, 1);
    for (int i = 0; i < ATT_STAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 256);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, ATT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;        // columns [0,128)
  const uint32_t tmem_o = tmem_base + 128;  // columns [128,256): O_A | O_B

  if (warp == 8) {
    // ======================================= TMA producer =======================================
    if (lane == 0) {
      const int q_col = head * ATT_D;
      const int k_col = p.dim + head * ATT_D;
      const int v_col = 2 * p.dim + head * ATT_D;
      mbar_expect_tx(q_full, ATT_TILE_BYTES);
      tma_load_3d(sQ, &tma_qkv, q_full, q_col, q_tile * ATT_BQ, batch);
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
  } else if (warp == 9) {
    // ======================================= MMA issuer =========================================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, 0, 0);  // A, B K-major
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);   // B (= V) MN-major
      const uint64_t q_desc = umma_desc_sw128(smem_u32(sQ));
      const uint64_t p_desc = umma_desc_sw128(smem_u32(sP));
      auto issue_qk = [&](int j) {
        const int st = j % ATT_STAGES;
        mbar_wait(&k_full[st], static_cast<uint32_t>((j / ATT_STAGES) & 1));
        tc_fence_after();
        const uint64_t k_desc = umma_desc_sw128(smem_u32(sK + st * ATT_TILE_BYTES));
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k)
          umma_f16_ss(tmem_s, q_desc + static_cast<uint64_t>(k * 2), k_desc + static_cast<uint64_t>(k * 2), idesc_qk,
                      static_cast<uint32_t>(k != 0));
        tc_commit(s_full);
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % ATT_STAGES;
        mbar_wait(p_full, static_cast<uint32_t>(j & 1));  // P_j in smem; S_j read; O rescaled if needed
        tc_fence_after();
        // S is free again: start the next QK^T first so the next softmax is not held up by P_j V_j.
        if (j + 1 < n_kv) issue_qk(j + 1);
        mbar_wait(&v_full[st], static_cast<uint32_t>((j / ATT_STAGES) & 1));
        tc_fence_after();
        const uint64_t v_desc = umma_desc_sw128(smem_u32(sV + st * ATT_TILE_BYTES));
#pragma unroll
        for (int s = 0; s < 2; ++s) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // A: P sub-tile s (16 KB apart), 32 bytes per K step.  B: keys s*64 + k*16 .. +16 = rows of 128 B.
            const uint64_t a = p_desc + static_cast<uint64_t>((s * ATT_TILE_BYTES + k * 32) >> 4);
            const uint64_t b = v_desc + static_cast<uint64_t>(((s * 64 + k * 16) * 128) >> 4);
            umma_f16_ss(tmem_o + static_cast<uint32_t>(s * 64), a, b, idesc_pv, static_cast<uint32_t>((j | k) != 0));
          }
        }
        tc_commit(o_full);
        tc_commit(&kv_empty[st]);
      }
    }
  } else {
    // ======================================= softmax streams ====================================
    const int stream = warp >> 2;                    // 0: key columns [0,64) of each tile, 1: [64,128)
    const int row_in_tile = (warp & 3) * 32 + lane;  // TMEM lane == query row
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t my_s = tmem_s + lane_off + static_cast<uint32_t>(stream * 64);
    const uint32_t my_o = tmem_o + lane_off + static_cast<uint32_t>(stream * 64);
    float m_run = -INFINITY;  // running (lazy) max in the scaled log2 domain
    float l_run = 0.f;
    uint8_t* p_row = sP + stream * ATT_TILE_BYTES + row_in_tile * 128;
    const int sw = row_in_tile & 7;

    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(s_full, static_cast<uint32_t>(j & 1));
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
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c + i < valid) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
      }
      const float m_tile = mx * p.scale_log2;
      // PV_{j-1} must have retired before O is corrected or the P buffer is overwritten
      if (j > 0) {
        mbar_wait(o_full, static_cast<uint32_t>((j - 1) & 1));
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
      float l_add = 0.f;
#pragma unroll 1
      for (int c = 0; c < 64; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(my_s + static_cast<uint32_t>(c), r);
        tmem_ld_wait();
        float pv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float e = ex2(fmaf(__uint_as_float(r[i]), p.scale_log2, -m_run));
          if (valid < 64 && c + i >= valid) e = 0.f;
          pv[i] = e;
          l_add += e;
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
      l_run += l_add;
      fence_proxy_async_smem();  // make the P stores visible to the tensor-core (async) proxy
      tc_fence_before();         // order our TMEM loads / stores before the MMAs that follow
      mbar_arrive(p_full);
    }

    // ---- combine the two streams, normalise, store ----
    mbar_wait(o_full, static_cast<uint32_t>((n_kv - 1) & 1));
    tc_fence_after();
    float acc[ATT_D];
    {
      uint32_t o0[32], o1[32];
      tmem_ld_32x32b_x32(my_o, o0);
      tmem_ld_32x32b_x32(my_o + 32, o1);
      tmem_ld_wait();
#pragma unroll
      for (int d = 0; d < 32; ++d) {
        acc[d] = __uint_as_float(o0[d]);
        acc[32 + d] = __uint_as_float(o1[d]);
      }
    }
    float* xch = reinterpret_cast<float*>(sK) + row_in_tile * ATT_XCH_STRIDE;  // K/V smem is idle by now
    if (stream == 1) {
      xch[0] = m_run;
      xch[1] = l_run;
#pragma unroll
      for (int d = 0; d < ATT_D; ++d) xch[2 + d] = acc[d];
    }
    named_bar_sync(1, 256);
    if (stream == 0) {
      const float m_b = xch[0], l_b = xch[1];
      const float m = fmaxf(m_run, m_b);
      const float wa = (l_run > 0.f) ? ex2(m_run - m) : 0.f;
      const float wb = (l_b > 0.f) ? ex2(m_b - m) : 0.f;
      const float inv = 1.0f / (l_run * wa + l_b * wb);
      const int q_row = q_tile * ATT_BQ + row_in_tile;
      if (q_row < p.seq_len) {
        bf16* op = p.out + (static_cast<size_t>(batch) * p.seq_len + q_row) * p.dim + head * ATT_D;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float o[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) o[t] = (acc[q * 8 + t] * wa + xch[2 + q * 8 + t] * wb) * inv;
          uint4 v;
          v.x = pack_bf16(o[0], o[1]);
          v.y = pack_bf16(o[2], o[3]);
          v.z = pack_bf16(o[4], o[5]);
          v.w = pack_bf16(o[6], o[7]);
          *reinterpret_cast<uint4*>(op + q * 8) = v;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ATT_TMEM_COLS);
  }
}

int attention_launch(const void* qkv, void* out, int batch, int seq_len, int heads, float scale,
                     cudaStream_t stream) {
  ORVB_REQUIRE(qkv && out, ORVB_EINVAL, "orvb_attention_bf16: null pointer");
  ORVB_REQUIRE(batch > 0 && seq_len > 0 && heads > 0, ORVB_ESHAPE, "orvb_attention_bf16: empty problem");
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
  dim3 grid((seq_len + ATT_BQ - 1) / ATT_BQ, heads, batch);
  attention_kernel<<<grid, ATT_THREADS, ATT_SMEM_BYTES, stream>>>(tm, p);
  ORVB_CHECK_CUDA(cudaGetLastError());
  return ORVB_OK;
}

}  // namespace orvb

extern "C" int orvb_attention_bf16(const void* qkv, void* out, int32_t batch, int32_t seq_len, int32_t heads,
                                   float scale, void* stream) {
  int rc = orvb::check_arch();
  if (rc != ORVB_OK) return rc;
  return orvb::attention_launch(qkv, out, batch, seq_len, heads, scale, static_cast<cudaStream_t>(stream));
}
