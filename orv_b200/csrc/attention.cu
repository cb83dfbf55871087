// Non-causal multi-head attention (head_dim 64) for sm_100a with both contractions on tcgen05:
//   S   = Q K^T          (128 x 128 x 64,  A = Q smem K-major, B = K smem K-major, D in TMEM)
//   O_s += P_s V_s       (128 x 64 x 64 per column half s, A = P smem K-major (written by the softmax warps),
//                         B = V smem MN-major, D accumulates in TMEM)
// replacing F.scaled_dot_product_attention at reference orv/models/cogvideox_control.py:256-258.
//
// One CTA per (PAIR of 128-query tiles, head, batch), one CTA per SM.  Warp roles:
//   warps 0-7   query tile 0: warps 0-3 softmax stream A (key columns [0,64) of every 128-key tile, one query row
//               per thread), warps 4-7 stream B (key columns [64,128))
//   warps 8-15  query tile 1, same split
//   warp  16    TMA producer (both Q tiles once; K and V double-buffered with separate release barriers)
//   warp  17    TMEM allocator + MMA issuer of query tile 0 (of both tiles in the single-issuer variant)
//   warp  18    MMA issuer of query tile 1
// The two streams of a tile are independent flash-attention accumulations (own running max / sum and own TMEM output
// accumulator, combined once at the end like a split-KV reduction): twice the softmax warps to hide MUFU / TMEM
// latency, no per-tile cross-thread reduction.  Output accumulators stay in TMEM; the running max is only raised when
// a tile exceeds it by more than 2^8 (lazy rescale), so the TMEM read-modify-write correction is rare.  Rows past
// seq_len are zero-filled by TMA and masked in the softmax.
//
// Per 128-key tile a softmax thread copies its 64 scores TMEM -> registers ONCE and releases S_t right away (s_free):
// the next QK^T is queued while the exponentials of this tile are still being computed, so in steady state the
// softmax warps never wait for the tensor core.  Max, exp2 (FFMA2 + MUFU.EX2 + FADD2) and the bf16 P stores then
// run from registers; P is double-buffered in shared memory so writing P_j never waits for PV_{j-1}.
//
// What bounds it (profiles/r01_attention_notes.md): MUFU.EX2 at 16/clk/SM = 2048 clk per 2 x (128 x 128) scores;
// measured ~2600 clk per key tile in the shipped variant, plus 2.64 -> 3 wave quantisation of the 390 CTAs.
#include <stdlib.h>

#include "common.cuh"
#include "pointwise.cuh"
#include "ptx.cuh"

namespace orvb {

constexpr int ATT_BQ = 128;
constexpr int ATT_BK = 128;
constexpr int ATT_D = 64;
constexpr int ATT_TMEM_COLS = 512;  // S_t: [t*128, +128)   O_{t,stream}: [256 + t*128 + stream*64, +64)
constexpr int ATT_XCH_STRIDE = 67;  // floats per row of the end-of-kernel stream exchange (conflict-free)
constexpr float ATT_RESCALE_THRESHOLD = 8.0f;  // log2 units
// 1 = ONE MMA-issuing warp, warp-uniform issue.  The one-issuer-per-query-tile variants (3, 11, 19) are ~4 % faster but
// NOT safe: when the rare lazy-rescale path rewrites an O accumulator (tcgen05.ld -> scale -> tcgen05.st) while a
// DIFFERENT thread is issuing tcgen05.mma, a few rows come out wrong (forced rescale: 28/30 launches bad with two
// issuers, 0/60 with one; same signature with two CTAs per SM — profiles/r01_attention_notes.md).  They are compiled
// only into measurement builds (-DORVB_EXPERIMENTAL, ORVB_ATT_VARIANT); the product library does not contain them.
constexpr int ATT_DEFAULT_VARIANT = 1;

struct AttDev {
  bf16* out;
  int seq_len, heads, dim;  // dim = heads * 64
  float scale_log2;         // softmax scale * log2(e)
  int q_row0, q_rows;       // queries = rows [q_row0, q_row0 + q_rows) of every sequence; output is compact
  float rescale_threshold;  // log2 units by which a tile max must exceed the running max before it is raised
  long long* dbg;           // optional timeline buffer (tools/profile_attention_timeline.py); nullptr in production
  int out_f32;              // test mode: `out` is fp32 (the normalised accumulator before the bf16 rounding)
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Template switches (ORVB_ATT_VARIANT picks one per process for A/B timing): DUAL = one MMA-issuing warp per query
// tile, UNIFORM = warp-uniform issue loop (elect.sync around the tcgen05 instructions only), EMU = share of the
// exponentials taken on the FMA pipe instead of MUFU.
constexpr int A4_TILE = 128 * 64 * 2;                      // 16 KB
constexpr int A4_OFF_P = 2 * A4_TILE;                      // P[t][buf]: 32 KB each (two 16 KB stream sub-tiles)
constexpr int A4_OFF_K = A4_OFF_P + 4 * 2 * A4_TILE;       // K[2]
constexpr int A4_OFF_V = A4_OFF_K + 2 * A4_TILE;           // V[2]
constexpr int A4_OFF_BAR = A4_OFF_V + 2 * A4_TILE;         // 224 KB
constexpr int A4_SMEM_BYTES = A4_OFF_BAR + 256;

__device__ __forceinline__ void fma2(float& d0, float& d1, float a0, float a1, float b, float c) {
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %4};\n\tmov.b64 rc, {%5, %5};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b), "f"(c));
}
__device__ __forceinline__ void add2(float& d0, float& d1, float a0, float a1) {
  asm("{\n\t.reg .b64 ra, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rd, {%0, %1};\n\t"
      "add.rn.f32x2 rd, rd, ra;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "+f"(d0), "+f"(d1)
      : "f"(a0), "f"(a1));
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float ex2_nv(float x) {  // non-volatile: may be scheduled freely
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Row max of one thread's 64 raw scores (already in registers).
template <bool MASKED>
__device__ __forceinline__ float row_max64(const uint32_t (&r0)[32], const uint32_t (&r1)[32], int valid) {
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
  if (MASKED) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if (i < valid) m0 = fmaxf(m0, __uint_as_float(r0[i]));
      if (32 + i < valid) m1 = fmaxf(m1, __uint_as_float(r1[i]));
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      m0 = max3(m0, __uint_as_float(r0[i]), __uint_as_float(r0[i + 1]));
      m1 = max3(m1, __uint_as_float(r0[i + 2]), __uint_as_float(r0[i + 3]));
      m2 = max3(m2, __uint_as_float(r1[i]), __uint_as_float(r1[i + 1]));
      m3 = max3(m3, __uint_as_float(r1[i + 2]), __uint_as_float(r1[i + 3]));
    }
  }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

// exp2 of a pair on the FMA / ALU pipes instead of the MUFU (the unit that bounds this kernel): round-to-nearest
// split x = i + f with the 1.5 * 2^23 trick, cubic minimax polynomial for 2^f on [-0.5, 0.5] (max relative error
// 7.5e-5, 50x below one bf16 ulp), exponent inserted with one integer multiply-add.  x <= ~8 always holds (lazy
// rescale threshold); x is clamped at -125 so the exponent field cannot wrap.
__device__ __forceinline__ void ex2_emulated2(f32x2 x, float& p0, float& p1) {
  const float kMagic = 12582912.0f;  // 1.5 * 2^23
  float x0, x1;
  upk2(x, x0, x1);
  x0 = fmaxf(x0, -125.0f);
  x1 = fmaxf(x1, -125.0f);
  const f32x2 xc = pk2(x0, x1);
  const f32x2 t = add2p(xc, pk2(kMagic, kMagic));
  const f32x2 i = add2p(t, pk2(-kMagic, -kMagic));
  const f32x2 f = fma2p(i, pk2(-1.0f, -1.0f), xc);
  f32x2 q = fma2p(f, pk2(0.055171408f, 0.055171408f), pk2(0.24261075f, 0.24261075f));
  q = fma2p(q, f, pk2(0.69326097f, 0.69326097f));
  q = fma2p(q, f, pk2(0.99992812f, 0.99992812f));
  float t0, t1, q0, q1;
  upk2(t, t0, t1);
  upk2(q, q0, q1);
  p0 = __uint_as_float(__float_as_uint(t0) * 0x800000u + __float_as_uint(q0));
  p1 = __uint_as_float(__float_as_uint(t1) * 0x800000u + __float_as_uint(q1));
}

// Exponentials of one thread's 64 scores: P = exp2(S * scale - m) as bf16 into the thread's 128-byte row of the
// (128B-swizzled, K-major) P sub-tile; returns the fp32 row sum.  EMU is a 4-bit mask over the four score pairs of
// every 8-column group: pairs whose bit is set take the FMA-pipe exp2 instead of MUFU.EX2.
template <bool MASKED, int EMU>
__device__ __forceinline__ float softmax_pass(const uint32_t (&r0)[32], const uint32_t (&r1)[32], uint8_t* p_row, int sw,
                                              float scale, float neg_m, int valid) {
  f32x2 sa = pk2(0.f, 0.f), sb = pk2(0.f, 0.f);
  const f32x2 scale2 = pk2(scale, scale), negm2 = pk2(neg_m, neg_m);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const uint32_t(&r)[32] = h ? r1 : r0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = q * 8 + e * 2;
        const f32x2 x = fma2p(pk2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), scale2, negm2);
        float p0, p1;
        if ((EMU >> e) & 1) {
          ex2_emulated2(x, p0, p1);
        } else {
          upk2(x, p0, p1);
          p0 = ex2_nv(p0);
          p1 = ex2_nv(p1);
        }
        if (MASKED) {
          const int col = h * 32 + i;
          if (col >= valid) p0 = 0.f;
          if (col + 1 >= valid) p1 = 0.f;
        }
        if (e & 1) sb = add2p(sb, pk2(p0, p1));
        else sa = add2p(sa, pk2(p0, p1));
        w[e] = pack_bf16(p0, p1);
      }
      *reinterpret_cast<uint4*>(p_row + (((h * 4 + q) ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
  float s0, s1, s2, s3;
  upk2(sa, s0, s1);
  upk2(sb, s2, s3);
  return (s0 + s1) + (s2 + s3);
}

// Timeline stamps of CTA (0,0,0): row `who`, slot = 4 * j + k, taken when a debug buffer is installed
// (tools/profile_attention_timeline.py).  Compiled in only with -DORVB_ATT_TIMELINE: the eight predicated stamps per
// key tile cost issue slots the softmax warps are short of.
#ifdef ORVB_ATT_TIMELINE
#define A4_STAMP(who, j, k)                                                                             \
  do {                                                                                                  \
    if (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 && (j) < 32) \
      p.dbg[(who) * 128 + (j) * 4 + (k)] = clock64();                                                   \
  } while (0)
#else
#define A4_STAMP(who, j, k) \
  do {                      \
  } while (0)
#endif

template <bool DUAL, bool UNIFORM, int EMU>
__global__ void __launch_bounds__(DUAL ? 608 : 576, 1)
attention_kernel(const __grid_constant__ CUtensorMap tma_qkv, const AttDev p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A4_OFF_BAR);
  uint64_t* q_full = bars;        // 1
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* v_full = bars + 3;    // [2]
  uint64_t* k_empty = bars + 5;   // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2] per query tile
  // p_full: [2 tiles][2 P buffers].  One barrier per P buffer, not per tile: the single issuing warp serves the two
  // tiles in turn, and a tile whose next QK^T is already queued may finish key tile j+1 while the issuer is still
  // blocked on the OTHER tile's P_j.  With one barrier per tile its P_j wait would then be two phases behind, which a
  // parity wait cannot tell from "not yet" (dead-lock until the watchdog trap); per buffer it is at most one.
  uint64_t* p_full = bars + 11;   // [4]
  uint64_t* o_full = bars + 15;   // [2 tiles][2 P buffers]: PV_{t,j} retired (j & 1 selects the barrier)
  uint64_t* s_free = bars + 19;   // [2] S_t copied to registers by all 8 softmax warps of the tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_pair = blockIdx.x;
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int n_kv = (p.seq_len + ATT_BK - 1) / ATT_BK;

  if ((smem_u32(smem) & 1023u) != 0) __trap();

  pdl_launch_dependents();
  if (warp == 16 && lane == 0) {
    tma_prefetch_desc(&tma_qkv);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&k_empty[i], DUAL ? 2 : 1);
      mbar_init(&v_empty[i], DUAL ? 2 : 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[2 * i], 256);  // every softmax thread of the tile, after its own proxy fence
      mbar_init(&p_full[2 * i + 1], 256);
      mbar_init(&o_full[2 * i], 1);
      mbar_init(&o_full[2 * i + 1], 1);
      mbar_init(&s_free[i], 256);
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
  pdl_wait();  // the QKV GEMM's output is visible from here on (prologue above overlapped its tail)

  if (warp == 16) {
    // ======================================= TMA producer =======================================
    if (lane == 0) {
      const int q_col = head * ATT_D;
      const int k_col = p.dim + head * ATT_D;
      const int v_col = 2 * p.dim + head * ATT_D;
      mbar_expect_tx(q_full, 2 * A4_TILE);
      tma_load_3d(smem, &tma_qkv, q_full, q_col, p.q_row0 + (2 * q_pair) * ATT_BQ, batch);
      tma_load_3d(smem + A4_TILE, &tma_qkv, q_full, q_col, p.q_row0 + (2 * q_pair + 1) * ATT_BQ, batch);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        const uint32_t ph = static_cast<uint32_t>((j >> 1) & 1);
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_expect_tx(&k_full[st], A4_TILE);
        tma_load_3d(smem + A4_OFF_K + st * A4_TILE, &tma_qkv, &k_full[st], k_col, j * ATT_BK, batch);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_expect_tx(&v_full[st], A4_TILE);
        tma_load_3d(smem + A4_OFF_V + st * A4_TILE, &tma_qkv, &v_full[st], v_col, j * ATT_BK, batch);
      }
    }
  } else if (warp >= 17) {
    // ======================================= MMA issuer(s) ======================================
    // DUAL: warp 17 drives query tile 0, warp 18 tile 1.  Otherwise warp 17 alternates between the two tiles.
    const int t_lo = DUAL ? (warp - 17) : 0;
    const int t_hi = DUAL ? (warp - 17) : 1;
    if (UNIFORM || lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, 0, 0);  // A, B K-major
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);   // B (= V) MN-major
      auto leader = [&]() -> bool { return UNIFORM ? elect_one() : true; };
      auto issue_qk = [&](int t, int j, bool last_of_tile_pair) {
        const int st = j & 1;
        if (t == t_lo) {  // K_j landed: observed once per key tile, the second query tile of this warp relies on it
          mbar_wait(&k_full[st], static_cast<uint32_t>((j >> 1) & 1));
          tc_fence_after();
        }
        const uint64_t q_desc = umma_desc_sw128(smem_u32(smem + t * A4_TILE));
        const uint64_t k_desc = umma_desc_sw128(smem_u32(smem + A4_OFF_K + st * A4_TILE));
        if (leader()) {
#pragma unroll
          for (int k = 0; k < ATT_D / 16; ++k)
            umma_f16_ss(tmem_base + static_cast<uint32_t>(t * 128), q_desc + static_cast<uint64_t>(k * 2),
                        k_desc + static_cast<uint64_t>(k * 2), idesc_qk, static_cast<uint32_t>(k != 0));
          tc_commit(&s_full[t]);
          if (last_of_tile_pair) tc_commit(&k_empty[st]);  // K_j consumed (by every tile this warp drives)
        }
        if (UNIFORM) __syncwarp();
      };
      mbar_wait(q_full, 0);
      for (int t = t_lo; t <= t_hi; ++t) issue_qk(t, 0, t == t_hi);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        // S_{t,j} lives in the softmax warps' registers as soon as they have loaded it: the next QK^T is queued right
        // then, so S_{t,j+1} is complete long before the exponentials of tile j are — the softmax warps never wait
        // for the tensor core in steady state.
        if (j + 1 < n_kv) {
#pragma unroll 1
          for (int t = t_lo; t <= t_hi; ++t) {
            mbar_wait(&s_free[t], static_cast<uint32_t>(j & 1));
            tc_fence_after();
            A4_STAMP(16 + t, j, 0);
            issue_qk(t, j + 1, t == t_hi);
            A4_STAMP(16 + t, j, 1);
          }
        }
#pragma unroll 1
        for (int t = t_lo; t <= t_hi; ++t) {
          mbar_wait(&p_full[2 * t + (j & 1)], static_cast<uint32_t>((j >> 1) & 1));  // P_{t,j} in smem, O_t rescaled
          tc_fence_after();
          A4_STAMP(16 + t, j, 2);
          if (t == t_lo) {  // V_j landed (once per key tile, as for K)
            mbar_wait(&v_full[st], static_cast<uint32_t>((j >> 1) & 1));
            tc_fence_after();
          }
          const uint64_t p_desc = umma_desc_sw128(smem_u32(smem + A4_OFF_P + (t * 2 + (j & 1)) * 2 * A4_TILE));
          const uint64_t v_desc = umma_desc_sw128(smem_u32(smem + A4_OFF_V + st * A4_TILE));
          if (leader()) {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                // A: P sub-tile s (16 KB apart), 32 bytes per K step.  B: keys s*64 + k*16 .. +16 = rows of 128 B.
                const uint64_t a = p_desc + static_cast<uint64_t>((s * A4_TILE + k * 32) >> 4);
                const uint64_t b = v_desc + static_cast<uint64_t>(((s * 64 + k * 16) * 128) >> 4);
                umma_f16_ss(tmem_base + static_cast<uint32_t>(256 + t * 128 + s * 64), a, b, idesc_pv,
                            static_cast<uint32_t>((j | k) != 0));
              }
            }
            tc_commit(&o_full[2 * t + (j & 1)]);
            if (t == t_hi) tc_commit(&v_empty[st]);  // V_j consumed
          }
          if (UNIFORM) __syncwarp();
          A4_STAMP(16 + t, j, 3);
        }
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
    uint64_t* my_p_full = &p_full[2 * t];  // [j & 1]
    uint64_t* my_o_full = &o_full[2 * t];  // [j & 1]; completion k of barrier b belongs to PV_{t, 2k + b}
    uint64_t* my_s_free = &s_free[t];
    float m_run = -INFINITY;  // running (lazy) max in the scaled log2 domain
    float l_run = 0.f;
    uint8_t* p_row0 = smem + A4_OFF_P + (t * 2) * 2 * A4_TILE + stream * A4_TILE + row_in_tile * 128;
    const int sw = row_in_tile & 7;
    const float scale = p.scale_log2;

    for (int j = 0; j < n_kv; ++j) {
      A4_STAMP(warp, j, 0);
      mbar_wait(my_s_full, static_cast<uint32_t>(j & 1));
      tc_fence_after();
      A4_STAMP(warp, j, 1);
      const int valid = p.seq_len - (j * ATT_BK + stream * 64);  // my columns >= valid are padding
      uint8_t* p_row = p_row0 + (j & 1) * 2 * A4_TILE;
      uint32_t r0[32], r1[32];
      tmem_ld_32x32b_x16(my_s, r0);
      tmem_ld_32x32b_x16(my_s + 16u, r0 + 16);
      tmem_ld_32x32b_x16(my_s + 32u, r1);
      tmem_ld_32x32b_x16(my_s + 48u, r1 + 16);
      tmem_ld_wait();
      tc_fence_before();  // our reads of S_t are ordered before the QK^T that overwrites it
      mbar_arrive(my_s_free);
      A4_STAMP(warp, j, 2);
      const float mx = (valid >= 64) ? row_max64<false>(r0, r1, valid) : row_max64<true>(r0, r1, valid);
      const float m_tile = mx * scale;
      // ---- lazy rescale: raise the running max only when this tile exceeds it by > 2^8 ----
      const bool grow = m_tile > m_run + p.rescale_threshold;
      if (__any_sync(0xffffffffu, grow)) {
        const float m_new = grow ? m_tile : m_run;
        const float alpha = grow ? ex2(m_run - m_new) : 1.0f;
        if (j > 0) {
          mbar_wait(&my_o_full[(j - 1) & 1], static_cast<uint32_t>(((j - 1) >> 1) & 1));  // PV_{t,j-1} retired
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < 64; c += 8) {
            uint32_t o[8];
            tmem_ld_32x32b_x8(my_o + static_cast<uint32_t>(c), o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32b_x8(my_o + static_cast<uint32_t>(c), o);
          }
          tmem_st_wait();
        }
        l_run *= alpha;
        m_run = m_new;
      }
      const float sum = (valid >= 64) ? softmax_pass<false, EMU>(r0, r1, p_row, sw, scale, -m_run, valid)
                                      : softmax_pass<true, 0>(r0, r1, p_row, sw, scale, -m_run, valid);
      l_run += sum;
      A4_STAMP(warp, j, 3);
      fence_proxy_async_smem();  // make the P stores visible to the tensor-core (async) proxy
      tc_fence_before();         // order our TMEM loads / stores before the MMAs that follow
      mbar_arrive(&my_p_full[j & 1]);
    }

    // ---- combine the two streams of this query tile, normalise, store ----
    // The loop above never waits for the PV MMAs: wait for the last use of each P buffer (PV_{n-2} and PV_{n-1}); with
    // one barrier per buffer a parity wait is never more than one phase behind.
    if (n_kv >= 2) mbar_wait(&my_o_full[(n_kv - 2) & 1], static_cast<uint32_t>(((n_kv - 2) >> 1) & 1));
    mbar_wait(&my_o_full[(n_kv - 1) & 1], static_cast<uint32_t>(((n_kv - 1) >> 1) & 1));
    tc_fence_after();
    // this tile's P buffers (64 KB) are idle once its last PV retired; O moves in 32-column halves
    float* xch = reinterpret_cast<float*>(smem + A4_OFF_P + (t * 2) * 2 * A4_TILE) + row_in_tile * ATT_XCH_STRIDE;
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
            if (p.out_f32) {
              float* of = reinterpret_cast<float*>(p.out) +
                          (static_cast<size_t>(batch) * p.q_rows + q_row) * p.dim + head * ATT_D + c + q * 8;
              *reinterpret_cast<float4*>(of) = make_float4(f[0], f[1], f[2], f[3]);
              *reinterpret_cast<float4*>(of + 4) = make_float4(f[4], f[5], f[6], f[7]);
              continue;
            }
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

template <bool DUAL, bool UNIFORM, int EMU>
static int launch_attention_v4(const CUtensorMap& tm, const AttDev& p, dim3 grid, cudaStream_t stream) {
  static bool attr_set = false;
  auto kern = attention_kernel<DUAL, UNIFORM, EMU>;
  if (!attr_set) {
    ORVB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, A4_SMEM_BYTES));
    attr_set = true;
  }
  ORVB_CHECK_CUDA(launch_kernel(kern, grid, dim3(DUAL ? 608 : 576), A4_SMEM_BYTES, stream, true, tm, p));
  return ORVB_OK;
}

static long long* g_att_dbg = nullptr;
static float g_att_threshold = -1.f;  // < 0: take ORVB_ATT_THRESHOLD or the default on the next launch

int attention_launch(const void* qkv, void* out, int batch, int seq_len, int heads, float scale, int q_row0,
                     int q_rows, cudaStream_t stream, int out_f32) {
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
  AttDev p;
  p.out = static_cast<bf16*>(out);
  p.seq_len = seq_len;
  p.heads = heads;
  p.dim = dim;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.q_row0 = q_row0;
  p.q_rows = q_rows;
  p.dbg = g_att_dbg;
  p.out_f32 = out_f32 ? 1 : 0;
  // Test knob (ORVB_ATT_THRESHOLD or orvb_attention_set_rescale_threshold): 0 raises the running max on (almost) every key tile, i.e. forces the otherwise rare
  // TMEM read-modify-write of the O accumulators (tests/test_gpu_ops.py::test_attention_forced_rescale).
  if (g_att_threshold < 0.f) {
    const char* e = getenv("ORVB_ATT_THRESHOLD");
    g_att_threshold = e ? static_cast<float>(atof(e)) : ATT_RESCALE_THRESHOLD;
  }
  p.rescale_threshold = g_att_threshold;
  dim3 grid((q_rows + 2 * ATT_BQ - 1) / (2 * ATT_BQ), heads, batch);
#ifdef ORVB_EXPERIMENTAL
  // Measurement builds only (-DORVB_EXPERIMENTAL; absent from the product library): ORVB_ATT_VARIANT = issue mode
  // (bit 0 uniform, bit 1 one issuer per query tile — NOT safe, see ATT_DEFAULT_VARIANT) + 8 * (exp2 emulation: 0 none,
  // 1 = 25 %, 2 = 50 % of the scores).
  static int variant = -1;
  if (variant < 0) {
    const char* e = getenv("ORVB_ATT_VARIANT");
    variant = e ? atoi(e) : ATT_DEFAULT_VARIANT;
  }
  switch (variant) {
    case 1: return launch_attention_v4<false, true, 0>(tm, p, grid, stream);
    case 3: return launch_attention_v4<true, true, 0>(tm, p, grid, stream);
    case 9: return launch_attention_v4<false, true, 0x8>(tm, p, grid, stream);
    case 11: return launch_attention_v4<true, true, 0x8>(tm, p, grid, stream);
    case 17: return launch_attention_v4<false, true, 0xA>(tm, p, grid, stream);
    case 19: return launch_attention_v4<true, true, 0xA>(tm, p, grid, stream);
    default: break;
  }
  set_error("orvb_attention_bf16: unknown ORVB_ATT_VARIANT %d (1, 3, 9, 11, 17, 19)", variant);
  return ORVB_EINVAL;
#else
  // The product library holds exactly one attention kernel: one warp-uniform MMA issuer, no exp2 emulation.
  return launch_attention_v4<false, true, 0>(tm, p, grid, stream);
#endif
}

}  // namespace orvb

extern "C" int orvb_attention_bf16(const void* qkv, void* out, int32_t batch, int32_t seq_len, int32_t heads,
                                   float scale, void* stream) {
  int rc = orvb::check_arch();
  if (rc != ORVB_OK) return rc;
  return orvb::attention_launch(qkv, out, batch, seq_len, heads, scale, 0, seq_len, static_cast<cudaStream_t>(stream), 0);
}

extern "C" int orvb_attention(const orvb_attention_args* a, void* stream) {
  int rc = orvb::check_arch();
  if (rc != ORVB_OK) return rc;
  ORVB_REQUIRE(a != nullptr, ORVB_EINVAL, "orvb_attention: null argument struct");
  return orvb::attention_launch(a->qkv, a->out, a->batch, a->seq_len, a->heads, a->scale, a->q_row0, a->q_rows,
                                static_cast<cudaStream_t>(stream), a->out_f32);
}


// Measurement hook: installs (or clears, with NULL) a device buffer of 18 x 128 int64 that CTA (0,0,0) of the next
// attention launches fills with clock64() stamps (tools/profile_attention_timeline.py).
extern "C" void orvb_attention_set_debug(void* dev_buf) { orvb::g_att_dbg = static_cast<long long*>(dev_buf); }

// Test hook: log2-unit threshold of the lazy row-max update (default 8; 0 forces the O-accumulator rescale path on
// almost every key tile; a negative value restores the default / ORVB_ATT_THRESHOLD).
extern "C" void orvb_attention_set_rescale_threshold(float log2_units) { orvb::g_att_threshold = log2_units; }
