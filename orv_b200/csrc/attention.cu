// Non-causal multi-head attention (head_dim 64) over a packed QKV buffer for sm_100a, replacing
// F.scaled_dot_product_attention at reference orv/models/cogvideox_control.py:256-258.  Both contractions run on tcgen05
// and the probabilities never touch shared memory:
//   S_t = Q_t K^T        SS MMA, 128 x 128 x 64, D in TMEM columns [t*128, +128)
//   P_t = exp2(S_t - m)  one thread per query row copies its 128 scores TMEM -> registers, releases S_t at once (the next
//                        QK^T is queued while the exponentials run), writes bf16 pairs back with tcgen05.st into the
//                        tile's own P region, TMEM columns [384 + t*64, +64)
//   O_t += P_t V         TS MMA (A operand read from TMEM), 128 x 64 x 128, D in TMEM columns [256 + t*64, +64)
// Why (round 2; the round-1 kernel staged P in shared memory): that kernel moved 256 KB through shared memory per 128-key step (QK^T operands 64,
// P written 64 + read 64 by the MMA, V 32, TMA fills 32) = 2048 clk at 128 B/clk — as long as the MUFU.EX2 floor itself
// (2 x 128 x 128 exponentials at 16 / clk); the single MMA-issuing thread spent ~2000 of 2977 clk per key step inside
// issue blocks (profiles/r02m_attn_timeline.log).  Here the shared-memory traffic is 128 KB per step, the P round trip
// (8 x 16-byte stores per thread + proxy fence) is two tcgen05.st, and the freed 128 KB of shared memory deepen the K / V
// ring from 2 to 4 stages.
//
// One CTA per (pair of 128-query tiles, head, batch), one CTA per SM.  Warps 0-3: softmax of query tile 0 (thread =
// query row = TMEM lane, all 128 key columns of the step), warps 4-7: tile 1; warp 8: TMA producer; warp 9: TMEM
// allocator + the ONE MMA-issuing thread (a second issuer is not safe next to tcgen05.st, profiles/r01_attention_notes.md).
// One running max / sum / O accumulator per row: no stream combine at the end.  The running max is raised lazily (only
// when a step exceeds it by 2^8), so the TMEM read-modify-write of O is rare.
// Two measured refinements (profiles/r02p_attn5_sweep.log, r02o_attn5_timeline.log): query tile 1 starts ~600 clk behind
// tile 0, so the MUFU-free phases of the two warps that share an SM sub-partition (barrier wait, S copy, P store) do not
// coincide (138.5 -> 127.2 us on config 2; the offset is neutrally stable, nothing pulls the tiles back together once the
// issuer has slack); and one score pair in eight takes its exponential on the FMA pipe (cubic minimax, 7.5e-5 relative:
// 50x below the bf16 rounding of P) instead of MUFU.EX2, the unit that bounds the kernel (127.2 -> 119.9 us; two in
// eight is level, three in eight is issue-bound and slower).
#include <stdlib.h>

#include "common.cuh"
#include "pointwise.cuh"
#include "ptx.cuh"

namespace orvb {

constexpr int A5_BQ = 128;
constexpr int A5_BK = 128;
constexpr int A5_D = 64;
constexpr int A5_STAGES = 4;                                  // K and V ring depth
constexpr int A5_TILE = 128 * 64 * 2;                         // 16 KB
constexpr int A5_OFF_K = 2 * A5_TILE;                         // Q[2] | K[STAGES] | V[STAGES]
constexpr int A5_OFF_V = A5_OFF_K + A5_STAGES * A5_TILE;
constexpr int A5_OFF_BAR = A5_OFF_V + A5_STAGES * A5_TILE;    // 160 KB
constexpr int A5_SMEM_BYTES = A5_OFF_BAR + 256;
constexpr int A5_THREADS = 320;
constexpr int A5_EMU = 1;            // score pairs per eight whose exp2 runs on the FMA pipe
constexpr int A5_STAGGER_CLK = 600;  // head start of query tile 0
constexpr float ATT_RESCALE_THRESHOLD = 8.0f;  // log2 units
constexpr uint32_t A5_COL_S = 0, A5_COL_O = 256, A5_COL_P = 384;

struct AttDev {
  bf16* out;
  int seq_len, heads, dim;
  float scale_log2;
  int q_row0, q_rows;
  float rescale_threshold;
  long long* dbg;
  int out_f32;
  int stagger;  // clocks by which query tile 1 starts behind tile 0 (keeps the two tiles' MUFU-free phases apart)
};

#ifdef ORVB_ATT_TIMELINE
#define A5_STAMP(who, j, k)                                                                             \
  do {                                                                                                  \
    if (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 && (j) < 32) \
      p.dbg[(who) * 128 + (j) * 4 + (k)] = clock64();                                                   \
  } while (0)
#else
#define A5_STAMP(who, j, k) \
  do {                      \
  } while (0)
#endif

__device__ __forceinline__ float a5_max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float a5_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// max of 32 raw scores; columns >= valid (relative to r[0]) are padding
template <bool MASKED>
__device__ __forceinline__ float a5_max32(const uint32_t (&r)[32], int valid) {
  float m0 = -INFINITY, m1 = -INFINITY;
  if (MASKED) {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < valid) m0 = fmaxf(m0, __uint_as_float(r[i]));
  } else {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      m0 = a5_max3(m0, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
      m1 = a5_max3(m1, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
    }
  }
  return fmaxf(m0, m1);
}

// exp2 of a pair on the FMA / ALU pipes instead of the MUFU (the unit that bounds this kernel): round-to-nearest split
// x = i + f with the 1.5 * 2^23 trick, cubic minimax polynomial for 2^f on [-0.5, 0.5] (max relative error 7.5e-5, 50x
// below one bf16 ulp of P), exponent inserted with one integer multiply-add.  x <= threshold always holds (lazy rescale);
// x is clamped at -125 so the exponent field cannot wrap.
__device__ __forceinline__ void a5_ex2_fma(f32x2 x, float& p0, float& p1) {
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

// P = exp2(S * scale - m) for 32 scores, packed as 16 bf16 pairs (element 2i in the low half of word i); adds the row sum.
// EMU of every 8 score pairs take the FMA-pipe exponential instead of MUFU.EX2.
template <bool MASKED, int EMU>
__device__ __forceinline__ void a5_exp32(const uint32_t (&r)[32], uint32_t* w, f32x2 scale2, f32x2 negm2, int valid,
                                         f32x2& sa, f32x2& sb) {
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    const f32x2 x = fma2p(pk2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), scale2, negm2);
    float p0, p1;
    if (!MASKED && ((i >> 1) & 7) < EMU) {
      a5_ex2_fma(x, p0, p1);
    } else {
      upk2(x, p0, p1);
      p0 = a5_ex2(p0);
      p1 = a5_ex2(p1);
    }
    if (MASKED) {
      if (i >= valid) p0 = 0.f;
      if (i + 1 >= valid) p1 = 0.f;
    }
    if (i & 2) sb = add2p(sb, pk2(p0, p1));
    else sa = add2p(sa, pk2(p0, p1));
    w[i >> 1] = pack_bf16(p0, p1);
  }
}

template <int EMU>
__global__ void __launch_bounds__(A5_THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap tma_qkv, const AttDev p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A5_OFF_BAR);
  uint64_t* q_full = bars;                     // 1
  uint64_t* k_full = bars + 1;                 // [STAGES]
  uint64_t* v_full = k_full + A5_STAGES;
  uint64_t* k_empty = v_full + A5_STAGES;
  uint64_t* v_empty = k_empty + A5_STAGES;
  uint64_t* s_full = v_empty + A5_STAGES;      // [2] QK^T of the tile retired
  uint64_t* s_free = s_full + 2;               // [2] S_t copied to registers by the tile's 128 threads
  uint64_t* p_full = s_free + 2;               // [2] P_t written (and O_t rescaled) by the tile's 128 threads
  uint64_t* o_full = p_full + 2;               // [2] PV of the tile retired: P_t may be overwritten, O_t may be read
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_pair = blockIdx.x;
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int n_kv = (p.seq_len + A5_BK - 1) / A5_BK;

  if ((smem_u32(smem) & 1023u) != 0) __trap();

  pdl_launch_dependents();
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tma_qkv);
    mbar_init(q_full, 1);
    for (int i = 0; i < A5_STAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 128);
      mbar_init(&p_full[i], 128);
      mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // the QKV GEMM's output is visible from here on
  A5_STAMP(warp, 30, 0);

  if (warp == 8) {
    // ======================================= TMA producer =======================================
    if (lane == 0) {
      const int q_col = head * A5_D;
      const int k_col = p.dim + head * A5_D;
      const int v_col = 2 * p.dim + head * A5_D;
      mbar_expect_tx(q_full, 2 * A5_TILE);
      tma_load_3d(smem, &tma_qkv, q_full, q_col, p.q_row0 + (2 * q_pair) * A5_BQ, batch);
      tma_load_3d(smem + A5_TILE, &tma_qkv, q_full, q_col, p.q_row0 + (2 * q_pair + 1) * A5_BQ, batch);
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_expect_tx(&k_full[st], A5_TILE);
        tma_load_3d(smem + A5_OFF_K + st * A5_TILE, &tma_qkv, &k_full[st], k_col, j * A5_BK, batch);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_expect_tx(&v_full[st], A5_TILE);
        tma_load_3d(smem + A5_OFF_V + st * A5_TILE, &tma_qkv, &v_full[st], v_col, j * A5_BK, batch);
        if (++st == A5_STAGES) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 9) {
    // ======================================= MMA issuer ==========================================
    constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, 0, 0);  // A, B K-major
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);   // A from TMEM, B (= V) MN-major
    auto issue_qk = [&](int t, int j) {
      const int st = j % A5_STAGES;
      if (t == 0) {
        mbar_wait(&k_full[st], static_cast<uint32_t>((j / A5_STAGES) & 1));
        tc_fence_after();
      }
      const uint64_t q_desc = umma_desc_sw128(smem_u32(smem + t * A5_TILE));
      const uint64_t k_desc = umma_desc_sw128(smem_u32(smem + A5_OFF_K + st * A5_TILE));
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < A5_D / 16; ++k)
          umma_f16_ss(tmem_base + A5_COL_S + static_cast<uint32_t>(t * 128), q_desc + static_cast<uint64_t>(k * 2),
                      k_desc + static_cast<uint64_t>(k * 2), idesc_qk, static_cast<uint32_t>(k != 0));
        tc_commit(&s_full[t]);
        if (t == 1) tc_commit(&k_empty[st]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    issue_qk(0, 0);
    if (p.stagger > 0) {
      const long long t0 = clock64();
      while (clock64() - t0 < p.stagger) {
      }
    }
    issue_qk(1, 0);
    auto issue_pv = [&](int t, int j) {
      const int st = j % A5_STAGES;
      mbar_wait(&p_full[t], static_cast<uint32_t>(j & 1));
      tc_fence_after();
      A5_STAMP(16 + t, j, 2);
      if (t == 0) {  // tile 0 is the first user of V_j (tile 1's PV of the same step is issued one iteration later)
        mbar_wait(&v_full[st], static_cast<uint32_t>((j / A5_STAGES) & 1));
        tc_fence_after();
      }
      const uint64_t v_desc = umma_desc_sw128(smem_u32(smem + A5_OFF_V + st * A5_TILE));
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < A5_BK / 16; ++k) {
          // A: 16 keys = 8 TMEM columns of bf16 pairs.  B: keys k*16 .. +16 = 16 rows of 128 bytes.
          umma_f16_ts(tmem_base + A5_COL_O + static_cast<uint32_t>(t * 64),
                      tmem_base + A5_COL_P + static_cast<uint32_t>(t * 64 + k * 8),
                      v_desc + static_cast<uint64_t>((k * 16 * 128) >> 4), idesc_pv, static_cast<uint32_t>((j | k) != 0));
        }
        tc_commit(&o_full[t]);
        if (t == 1) tc_commit(&v_empty[st]);
      }
      __syncwarp();
      A5_STAMP(16 + t, j, 3);
    };
    auto issue_next_qk = [&](int t, int j) {  // QK^T of step j + 1 as soon as the tile has copied S_j out of TMEM
      if (j + 1 >= n_kv) return;
      mbar_wait(&s_free[t], static_cast<uint32_t>(j & 1));
      tc_fence_after();
      A5_STAMP(16 + t, j, 0);
      issue_qk(t, j + 1);
      A5_STAMP(16 + t, j, 1);
    };
    // The single issuing thread works through its waits IN PROGRAM ORDER.  Query tile 1 runs about half a key step behind
    // tile 0 (the stagger), so the events arrive as S_0(j) copied -> P_1(j-1) written -> S_1(j) copied -> P_0(j) written,
    // and the natural program order below (both QK^T of the next step, then both PV of this one) keeps QK_0(j+1) waiting
    // behind PV_1(j-1): a softmax warp finds its next S 100 - 260 clk late in every key step
    // (profiles/r02o_attn5_timeline.log, column "s_full wait").  Issuing in arrival order (-DORVB_ATT_EVENT_ORDER) removes
    // that wait and is 1.2 % faster in isolation (121.0 -> 119.6 us) but 0.4 % SLOWER inside the power-capped step, twice
    // in an alternating A/B on one box (attention 145.2 -> 146.6 us, clip 596.2 -> 598.9 ms, profiles/r03c_*): the
    // natural order stays.
    for (int j = 0; j < n_kv; ++j) {
#ifndef ORVB_ATT_EVENT_ORDER
      issue_next_qk(0, j);
      issue_next_qk(1, j);
      issue_pv(0, j);
      issue_pv(1, j);
#else
      issue_next_qk(0, j);
      if (j > 0) issue_pv(1, j - 1);
      issue_next_qk(1, j);
      issue_pv(0, j);
#endif
    }
#ifdef ORVB_ATT_EVENT_ORDER
    issue_pv(1, n_kv - 1);
#endif
  } else {
    // ======================================= softmax (one thread per query row) ==================
    const int t = warp >> 2;
    const int row_in_tile = (warp & 3) * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t my_s = tmem_base + lane_off + A5_COL_S + static_cast<uint32_t>(t * 128);
    const uint32_t my_o = tmem_base + lane_off + A5_COL_O + static_cast<uint32_t>(t * 64);
    const uint32_t my_p = tmem_base + lane_off + A5_COL_P + static_cast<uint32_t>(t * 64);
    float m_run = -INFINITY, l_run = 0.f;
    const float scale = p.scale_log2;

    for (int j = 0; j < n_kv; ++j) {
      const uint32_t par = static_cast<uint32_t>(j & 1);
      A5_STAMP(warp, j, 0);
      mbar_wait(&s_full[t], par);
      tc_fence_after();
      A5_STAMP(warp, j, 1);
      uint32_t r0[32], r1[32], r2[32], r3[32];
      tmem_ld_32x32b_x32(my_s, r0);
      tmem_ld_32x32b_x32(my_s + 32u, r1);
      tmem_ld_32x32b_x32(my_s + 64u, r2);
      tmem_ld_32x32b_x32(my_s + 96u, r3);
      tmem_ld_wait();
      tc_fence_before();  // our reads of S_t are ordered before the QK^T that overwrites it
      mbar_arrive(&s_free[t]);
      A5_STAMP(warp, j, 2);
      const int valid = p.seq_len - j * A5_BK;  // key columns >= valid are padding (only in the last step)
      float mx;
      if (valid >= A5_BK) {
        mx = fmaxf(fmaxf(a5_max32<false>(r0, 32), a5_max32<false>(r1, 32)),
                   fmaxf(a5_max32<false>(r2, 32), a5_max32<false>(r3, 32)));
      } else {
        mx = fmaxf(fmaxf(a5_max32<true>(r0, valid), a5_max32<true>(r1, valid - 32)),
                   fmaxf(a5_max32<true>(r2, valid - 64), a5_max32<true>(r3, valid - 96)));
      }
      const float m_tile = mx * scale;
      // ---- lazy rescale: raise the running max only when this step exceeds it by > 2^threshold ----
      const bool grow = m_tile > m_run + p.rescale_threshold;
      if (__any_sync(0xffffffffu, grow)) {
        const float m_new = grow ? m_tile : m_run;
        const float alpha = grow ? a5_ex2(m_run - m_new) : 1.0f;
        if (j > 0) {
          mbar_wait(&o_full[t], par ^ 1);  // PV_{t,j-1} retired
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < A5_D; c += 8) {
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
      uint32_t w0[32], w1[32];
      f32x2 sa = pk2(0.f, 0.f), sb = pk2(0.f, 0.f);
      const f32x2 scale2 = pk2(scale, scale), negm2 = pk2(-m_run, -m_run);
      if (valid >= A5_BK) {
        a5_exp32<false, EMU>(r0, w0, scale2, negm2, 32, sa, sb);
        a5_exp32<false, EMU>(r1, w0 + 16, scale2, negm2, 32, sa, sb);
        a5_exp32<false, EMU>(r2, w1, scale2, negm2, 32, sa, sb);
        a5_exp32<false, EMU>(r3, w1 + 16, scale2, negm2, 32, sa, sb);
      } else {
        a5_exp32<true, 0>(r0, w0, scale2, negm2, valid, sa, sb);
        a5_exp32<true, 0>(r1, w0 + 16, scale2, negm2, valid - 32, sa, sb);
        a5_exp32<true, 0>(r2, w1, scale2, negm2, valid - 64, sa, sb);
        a5_exp32<true, 0>(r3, w1 + 16, scale2, negm2, valid - 96, sa, sb);
      }
      float s0, s1, s2, s3;
      upk2(sa, s0, s1);
      upk2(sb, s2, s3);
      l_run += (s0 + s1) + (s2 + s3);
      A5_STAMP(warp, j, 3);
      if (j > 0) {  // the tile's single P region is free once PV_{t,j-1} retired
        mbar_wait(&o_full[t], par ^ 1);
        tc_fence_after();
      }
      tmem_st_32x32b_x32(my_p, w0);
      tmem_st_32x32b_x32(my_p + 32u, w1);
      tmem_st_wait();
      tc_fence_before();  // our TMEM stores are ordered before the PV MMAs issued after the barrier
      mbar_arrive(&p_full[t]);
    }

    // ---- normalise and store this row ----
    mbar_wait(&o_full[t], static_cast<uint32_t>((n_kv - 1) & 1));
    tc_fence_after();
    const float inv = (l_run > 0.f) ? 1.0f / l_run : 0.f;
    const int q_row = (2 * q_pair + t) * A5_BQ + row_in_tile;
    const size_t orow = (static_cast<size_t>(batch) * p.q_rows + q_row) * p.dim + head * A5_D;
#pragma unroll 1
    for (int c = 0; c < A5_D; c += 32) {
      uint32_t o[32];
      tmem_ld_32x32b_x32(my_o + static_cast<uint32_t>(c), o);
      tmem_ld_wait();
      if (q_row < p.q_rows) {
        if (p.out_f32) {
          float* of = reinterpret_cast<float*>(p.out) + orow + c;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(of + q * 4) =
                make_float4(__uint_as_float(o[q * 4]) * inv, __uint_as_float(o[q * 4 + 1]) * inv,
                            __uint_as_float(o[q * 4 + 2]) * inv, __uint_as_float(o[q * 4 + 3]) * inv);
        } else {
          bf16* op = p.out + orow + c;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 v;
            v.x = pack_bf16(__uint_as_float(o[q * 8 + 0]) * inv, __uint_as_float(o[q * 8 + 1]) * inv);
            v.y = pack_bf16(__uint_as_float(o[q * 8 + 2]) * inv, __uint_as_float(o[q * 8 + 3]) * inv);
            v.z = pack_bf16(__uint_as_float(o[q * 8 + 4]) * inv, __uint_as_float(o[q * 8 + 5]) * inv);
            v.w = pack_bf16(__uint_as_float(o[q * 8 + 6]) * inv, __uint_as_float(o[q * 8 + 7]) * inv);
            *reinterpret_cast<uint4*>(op + q * 8) = v;
          }
        }
      }
    }
  }

  A5_STAMP(warp, 30, 1);
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int EMU>
static int launch_attention(const CUtensorMap& tm, const AttDev& p, dim3 grid, cudaStream_t stream) {
  static bool attr_set = false;
  auto kern = attention_kernel<EMU>;
  if (!attr_set) {
    ORVB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, A5_SMEM_BYTES));
    attr_set = true;
  }
  ORVB_CHECK_CUDA(launch_kernel(kern, grid, dim3(A5_THREADS), A5_SMEM_BYTES, stream, true, tm, p));
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
  const int dim = heads * A5_D;
  CUtensorMap tm;
  int rc = make_tmap_3d_bf16(&tm, qkv, batch, seq_len, 3 * dim, 3 * dim, static_cast<uint64_t>(seq_len) * 3 * dim, A5_BK,
                             A5_D);
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
  // Test knob (ORVB_ATT_THRESHOLD or orvb_attention_set_rescale_threshold): 0 raises the running max on (almost) every
  // key step, i.e. forces the otherwise rare TMEM read-modify-write of the O accumulators
  // (tests/test_gpu_ops.py::test_attention_forced_rescale).
  if (g_att_threshold < 0.f) {
    const char* e = getenv("ORVB_ATT_THRESHOLD");
    g_att_threshold = e ? static_cast<float>(atof(e)) : ATT_RESCALE_THRESHOLD;
  }
  p.rescale_threshold = g_att_threshold;
  p.stagger = A5_STAGGER_CLK;
  dim3 grid((q_rows + 2 * A5_BQ - 1) / (2 * A5_BQ), heads, batch);
#ifdef ORVB_EXPERIMENTAL
  // Measurement builds only (-DORVB_EXPERIMENTAL; absent from the product library): ORVB_ATT_STAGGER (clocks) and
  // ORVB_ATT_EMU (0..3 score pairs in eight on the FMA pipe) for the sweeps of tools/sweep_attention.py.
  static int stagger = -1, emu = -1;
  if (stagger < 0) {
    const char* e = getenv("ORVB_ATT_STAGGER");
    stagger = e ? atoi(e) : A5_STAGGER_CLK;
    e = getenv("ORVB_ATT_EMU");
    emu = e ? atoi(e) : A5_EMU;
  }
  p.stagger = stagger;
  switch (emu) {
    case 0: return launch_attention<0>(tm, p, grid, stream);
    case 1: return launch_attention<1>(tm, p, grid, stream);
    case 2: return launch_attention<2>(tm, p, grid, stream);
    case 3: return launch_attention<3>(tm, p, grid, stream);
  }
  set_error("attention: ORVB_ATT_EMU must be 0..3");
  return ORVB_EINVAL;
#else
  // The product library holds exactly one attention kernel.
  return launch_attention<A5_EMU>(tm, p, grid, stream);
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
// attention launches fills with clock64() stamps (tools/profile_attention_timeline.py; builds with -DORVB_ATT_TIMELINE).
extern "C" void orvb_attention_set_debug(void* dev_buf) { orvb::g_att_dbg = static_cast<long long*>(dev_buf); }

// Test hook: log2-unit threshold of the lazy row-max update (default 8; 0 forces the O-accumulator rescale path on
// almost every key step; a negative value restores the default / ORVB_ATT_THRESHOLD).
extern "C" void orvb_attention_set_rescale_threshold(float log2_units) { orvb::g_att_threshold = log2_units; }
