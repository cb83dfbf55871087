// Persistent, warp-specialised bf16 GEMMs for sm_100a:  out = epilogue(A[M,K] · W[N,K]^T)
//
// Two kernels share the fused epilogues below: the single-CTA kernel (128 x BN tiles, used for M <= 128 and as the
// bit-exactness reference) and the CTA-pair kernel further down (256 x BN tiles, tcgen05 cta_group::2 — the one the
// forward runs).  Roles of the single-CTA kernel:
//
//   warp 0      TMA producer   (cp.async.bulk.tensor → 128B-swizzled smem ring)
//   warp 1      MMA issuer     (one thread, tcgen05.mma cta_group::1, 128 x BN x 16 atoms, fp32 accum in TMEM)
//   warp 2      TMEM allocator
//   warps 4-7   epilogue       (tcgen05.ld → registers → fused epilogue → 16-byte global stores)
//
// Two TMEM accumulator buffers let the epilogue of tile i overlap the mainloop of tile i+1.
// The epilogues fuse what the reference executes as separate ATen kernels around each nn.Linear
// (reference orv/models/cogvideox_control.py): bias; GELU-tanh (FeedForward, :431-440); gated residual
// `hidden + gate * linear(x)` (:419-421, :442-443); per-head QK LayerNorm + RoPE (:243-254); positional-table add
// (CogVideoXPatchEmbed).
#include "common.cuh"
#include "gemm_common.cuh"
#include "ptx.cuh"

namespace orvb {

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const GemmDev p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int num_k = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // The producer and MMA warps run their loops warp-uniformly and elect one lane only around the issuing
  // instructions: descriptor / coordinate arithmetic then stays in the uniform datapath (UR registers feed
  // UTMALDG / UTCHMMA directly).  Wrapping the whole loop in `if (lane == 0)` makes every operand a per-thread
  // value that must be moved with R2UR before each issue, which costs ~100 clk per tcgen05.mma (measured).
  if (warp == 0) {
    // ================================ TMA producer ================================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile % p.num_m_tiles;
      const int n_blk = tile / p.num_m_tiles;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
        uint8_t* sb = sa + Cfg::A_BYTES;
        if (elect_one()) {
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_2d(sa, &tma_a, &full_bar[stage], kb * BK, m_blk * BM);
          tma_load_2d(sb, &tma_b, &full_bar[stage], p.k_wrap > 0 ? (kb * BK) % p.k_wrap : kb * BK, n_blk * BN);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
        const uint64_t a_desc = umma_desc_sw128(sa);
        const uint64_t b_desc = umma_desc_sw128(sa + Cfg::A_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the swizzle span: +2 in the (addr >> 4) field
            umma_f16_ss(d_tmem, a_desc + static_cast<uint64_t>(k * 2), b_desc + static_cast<uint64_t>(k * 2), idesc,
                        static_cast<uint32_t>((kb | k) != 0));
          }
          tc_commit(&empty_bar[stage]);  // frees the smem stage once these MMAs retire
          if (kb == num_k - 1) tc_commit(&tfull_bar[acc]);  // accumulator complete
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue ====================================
    const int ew = warp - 4;  // == warp % 4: TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile % p.num_m_tiles;
      const int n_blk = tile / p.num_m_tiles;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * BM + ew * 32 + lane;
      const uint32_t taddr = tmem_base + static_cast<uint32_t>(acc * BN) + (static_cast<uint32_t>(ew * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BN; c += 64) {
        const int n0 = n_blk * BN + c;
        if (n0 >= p.N) break;  // warp-uniform
        uint32_t r0[32], r1[32];
        tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(c), r0);
        tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(c + 32), r1);
        tmem_ld_wait();
        if (row < p.M) {
          float v[64];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = __uint_as_float(r0[j]);
            v[32 + j] = __uint_as_float(r1[j]);
          }
          const int ncols = (p.N - n0 < 64) ? (p.N - n0) : 64;
          epilogue_unit<EPI>(p, v, row, n0, ncols);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  // ---- teardown ----------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------
// CTA-pair kernel: one 256 x BN output tile per cluster of two CTAs (tcgen05.mma cta_group::2, M = 256).
//
// The 1-CTA kernel above streams 16 KB of A + BN*128 B of B from L2 per 128 x BN x 64 step; on B200 that L2 -> SM
// stream, not the tensor pipe, is what bounds it (ncu: ~12 TB/s of TMA traffic at 52 % tensor-pipe activity).  With
// a pair, each CTA fetches only its own 128 rows of A and its own BN/2 rows of B: 2/3 of the bytes per FLOP at
// BN = 256.  BN is a run-time value (any multiple of 16 up to 256), so the host can pick the width that fills the
// last wave of the 74 clusters best.
//
// Roles per CTA: warp 0 TMA producer (both CTAs; byte counts are credited to the LEADER's full barrier), warp 1
// MMA issuer (leader only), warp 2 TMEM allocator, warps 4-11 epilogue over this CTA's 128 accumulator rows (eight
// warps: e and e + 4 share a TMEM lane quarter and alternate 64-column units; tcgen05.ld -> fused epilogue ->
// 128B-swizzled staging tile -> one TMA store per warp and unit).
// ---------------------------------------------------------------------------------------------------

// In-kernel timeline of the first and the last cluster (tools/profile_gemm_timeline.py): where the ~20 us per launch
// go in which the tensor pipe is idle (prologue, first-TMA fill, last epilogue, teardown).  Compiled in only with
// -DORVB_GEMM_TIMELINE; the default build contains none of it.  Layout: [first | last cluster][CTA rank][32 slots].
#ifdef ORVB_GEMM_TIMELINE
__device__ long long* g_gemm_dbg = nullptr;
#define G2_STAMP_VALUE(slot, value)                                                                          \
  do {                                                                                                       \
    long long* d_ = g_gemm_dbg;                                                                              \
    if (d_ != nullptr && lane == 0 && (cluster_id == 0 || cluster_id == num_clusters - 1))                   \
      d_[(cluster_id == 0 ? 0 : 64) + static_cast<int>(rank) * 32 + (slot)] = static_cast<long long>(value); \
  } while (0)
#define G2_STAMP(slot) G2_STAMP_VALUE(slot, clock64())
#else
#define G2_STAMP_VALUE(slot, value) \
  do {                              \
  } while (0)
#define G2_STAMP(slot) \
  do {                 \
  } while (0)
#endif

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(g2_threads(EPI), 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                  const __grid_constant__ CUtensorMap tma_o, const __grid_constant__ CUtensorMap tma_t, const GemmDev p,
                  const int bn) {
  constexpr int STAGES = G2_STAGES;
  extern __shared__ uint8_t smem_raw[];
  // Both CTAs of the pair must use identical offsets (the MMA applies the leader's descriptors to the peer's shared
  // memory), which holds because the dynamic shared window starts at the same shared::cta address in every CTA.
  // Plan (host: g2_plan): [stages][output staging][fast epilogue: gate / bias rows][barriers]; a stage holds 16 KB of A
  // and this CTA's half of B (the full 16 KB unless the host shrank it to make room for the fast epilogue's buffers).
  // (pointer arithmetic on the __shared__ array, not an integer round trip: the compiler keeps the address space and
  //  emits LDS / STS instead of generic loads and stores for the epilogue's staging traffic)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stage_bytes = p.stage_bytes;
  uint8_t* out_stage = smem + STAGES * stage_bytes;
  const bool fast_resid = (EPI == ORVB_EPI_GATE_RESID) && p.fast_resid != 0;
  uint8_t* gstage = out_stage + p.out_stage_bytes;  // fast epilogue only: 4 * units slots of G2_GS_SLOT bytes
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(gstage + (fast_resid ? 4 * ((bn + 63) >> 6) * G2_GS_SLOT : 0));
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* rbar = tempty_bar + 2;  // [8 epilogue warps][2 units]: residual tile landed (fast epilogue)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rbar + 16);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int num_tiles = p.tile_end;  // virtual tile indices (g2_tile); num_m_tiles counts 256-row pairs here
  const int num_k = (p.K + BK - 1) / BK;
  const int half_bn = bn >> 1;

  if (warp == 0) G2_STAMP(0);  // kernel entry
  pdl_launch_dependents();  // the next kernel's CTAs may be scheduled (and run their prologue) as SMs free up
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    if (p.tma_store) tma_prefetch_desc(&tma_o);
    if (fast_resid && (bn & 63)) tma_prefetch_desc(&tma_t);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);   // leader's producer arrive.expect_tx (both CTAs' bytes)
      mbar_init(&empty_bar[i], 1);  // multicast tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);   // multicast tcgen05.commit
      mbar_init(&tempty_bar[i], 2 * g2_epi_warps(EPI));  // epilogue warps x 2 CTAs (leader's copy is waited on)
    }
    for (int i = 0; i < 16; ++i) mbar_init(&rbar[i], 1);  // the issuing lane's arrive.expect_tx
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_slot, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of BOTH CTAs initialised before any remote arrive / TMA credit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) G2_STAMP(1);  // prologue done (barriers, TMEM, cluster sync)
  pdl_wait();  // everything above overlapped the previous kernel's tail; its outputs are visible from here on
  if (warp == 0) G2_STAMP(2);  // previous kernel's outputs visible

  if (warp == 0) {
    // ================================ TMA producer (both CTAs) ================================
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t stage_tx = 2u * static_cast<uint32_t>(G2_A_BYTES + half_bn * BK * 2);
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const G2Tile tl = g2_tile(p, tile, bn, num_clusters);
      if (tl.w == 0) continue;
      const int a_row = tl.m_blk * 256 + static_cast<int>(rank) * BM;
      // (a narrower tile still loads boxes of bn / 2 rows of W: the MMA reads the first w / 2 of them in each CTA)
      const int b_row = tl.n0 + static_cast<int>(rank) * (tl.w >> 1);
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * stage_bytes;
        uint8_t* sb = sa + G2_A_BYTES;
        const uint32_t leader_full = mapa_u32(smem_u32(&full_bar[stage]), 0);
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx(&full_bar[stage], stage_tx);
          tma_load_2d_pair(sa, &tma_a, leader_full, kb * BK, a_row);
          tma_load_2d_pair(sb, &tma_b, leader_full, p.k_wrap > 0 ? (kb * BK) % p.k_wrap : kb * BK, b_row);
        }
        __syncwarp();
        if (tile == cluster_id && kb == 0) G2_STAMP(3);  // first stage requested
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    G2_STAMP(4);  // last stage requested
  } else if (warp == 1 && rank == 0) {
    // ================================ MMA issuer (leader CTA) ================================
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const G2Tile tl = g2_tile(p, tile, bn, num_clusters);
      if (tl.w == 0) continue;
      const uint32_t idesc = umma_idesc_bf16(256, tl.w, 0, 0);
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * G2_ACC_COLS);
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (tile == cluster_id && kb == 0) G2_STAMP(5);  // first operands landed
        const uint32_t sa = smem_u32(smem + stage * stage_bytes);
        const uint64_t a_desc = umma_desc_sw128(sa);
        const uint64_t b_desc = umma_desc_sw128(sa + G2_A_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            umma_f16_ss_pair(d_tmem, a_desc + static_cast<uint64_t>(k * 2), b_desc + static_cast<uint64_t>(k * 2), idesc,
                             static_cast<uint32_t>((kb | k) != 0));
          }
          tc_commit_pair(&empty_bar[stage], 3);                     // frees this stage in both CTAs
          if (kb == num_k - 1) tc_commit_pair(&tfull_bar[acc], 3);  // accumulator complete in both CTAs
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (tile == cluster_id) G2_STAMP(6);  // first tile's MMAs issued
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    G2_STAMP(7);  // all MMAs issued
  } else if (warp >= 4) {
    // ================================ epilogue (both CTAs) ====================================
    constexpr int EW = g2_epi_warps(EPI);
    constexpr int NBUF = 8 / EW;             // staging tiles per warp: 32 KB split over the epilogue warps
    const int ew = (warp - 4) & 3;           // == warp % 4: TMEM lane quarter this warp may access
    const int unit_par = (warp - 4) >> 2;    // with 8 warps: which of the alternating 64-column units are mine
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t stores = 0;  // TMA stores issued by this warp (staging buffer = stores % NBUF)
    uint8_t* my_stage = out_stage + (warp - 4) * (NBUF * 32 * 128);
    if (fast_resid) {
      // ---- gated residual written in place (out == resid), everything but the accumulator fetched ahead ----------
      // The generic path below reads, per thread and AFTER the accumulator has landed, 128 bytes of its own residual
      // row and 256 bytes of a gate row from global memory: with 64 accumulator registers live the compiler can keep
      // only a few of those loads in flight, and the unit takes ~6 k clocks, all of it exposed behind the last tile of a
      // cluster (profiles/r02a_gemm_timeline.log).  Here every 64-column unit of the tile has its own staging tile:
      // BEFORE waiting for the accumulator the warp's elected lane fetches the [32 x 64] residual tile into it with one
      // TMA load (same tensor map and swizzle as the output store) and the warp copies the (at most two distinct) gate
      // rows and the bias of the unit into shared memory.  After the accumulator wait a thread only touches TMEM and
      // shared memory: out = resid + gate * (acc + bias) is formed in place in the staging tile and written back with
      // the usual TMA store.  A narrower last unit (tile width not a multiple of 64) goes through a second,
      // unswizzled tensor map whose box is exactly that wide, so it never touches the neighbouring tile's columns.
      const int U = (bn + 63) >> 6;
      const int tail = bn & 63;
      uint64_t* my_rbar = rbar + (warp - 4) * 2;
      uint32_t rphase = 0;       // bit i: parity of my_rbar[i]
      bool stores_out = false;   // this warp has TMA stores whose source reads may still be in flight
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const G2Tile tl = g2_tile(p, tile, bn, num_clusters);  // (always full-width tiles on this path: host)
        if (tl.w == 0) continue;
        const int m_blk = tl.m_blk;
        const int tile_n0 = tl.n0;
        const int row0 = m_blk * 256 + static_cast<int>(rank) * BM + ew * 32;
        const int row = row0 + lane;
        const bool rows_ok = row0 < p.M;  // warp-uniform
        const uint32_t taddr = tmem_base + static_cast<uint32_t>(acc * G2_ACC_COLS) + (static_cast<uint32_t>(ew * 32) << 16);
        // gate row of this thread's output row; the rows of a warp normally share one or two of them
        const float* gpl = nullptr;
        if (p.gate != nullptr && row < p.M) {
          int sq;
          int g = row_group(p.rm, row, &sq);
          if (p.grp_off != nullptr) g += *p.grp_off;
          const int is_text = (p.rm.seq_len > 0) ? (sq < p.rm.text_len) : 0;
          gpl = p.gate + static_cast<size_t>(g) * p.gate_ld + (is_text ? p.gate_text_off : p.gate_video_off);
        }
        int lv = p.M - 1 - row0;
        lv = lv > 31 ? 31 : (lv < 0 ? 0 : lv);
        const float* gA = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(gpl), 0));
        const float* gB = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(gpl), lv));
        const bool gate_staged = __all_sync(0xffffffffu, row >= p.M || gpl == gA || gpl == gB);
        const int gsel = (gpl == gA) ? 0 : 64;
        if (rows_ok) {
          if (stores_out) {
            if (lane == 0) bulk_wait_group_read<0>();  // last tile's stores have read my staging tiles
            stores_out = false;
          }
#pragma unroll 1
          for (int i = 0; i < 2; ++i) {
            const int u = unit_par + 2 * i;
            const int n0 = tile_n0 + u * 64;
            if (u >= U || n0 >= p.N) break;  // warp-uniform
            const bool is_tail = tail != 0 && u == U - 1;
            uint8_t* buf = out_stage + (ew * U + u) * 4096;
            if (lane == 0) {
              mbar_expect_tx(&my_rbar[i], is_tail ? static_cast<uint32_t>(32 * tail * 2) : 4096u);
              tma_load_2d(buf, is_tail ? &tma_t : &tma_o, &my_rbar[i], n0, row0);
            }
            float* gs = reinterpret_cast<float*>(gstage + (ew * U + u) * G2_GS_SLOT);
            const int cc = n0 + 2 * lane;
            float2 a2 = make_float2(1.f, 1.f), b2 = make_float2(1.f, 1.f);
            uint32_t bb = 0;
            if (cc < p.N) {
              if (p.gate != nullptr && gate_staged) {
                a2 = *reinterpret_cast<const float2*>(gA + cc);
                b2 = *reinterpret_cast<const float2*>(gB + cc);
              }
              if (p.bias != nullptr) bb = *reinterpret_cast<const uint32_t*>(p.bias + cc);
            }
            reinterpret_cast<float2*>(gs)[lane] = a2;
            reinterpret_cast<float2*>(gs + 64)[lane] = b2;
            reinterpret_cast<float2*>(gs + 128)[lane] = make_float2(bf16_lo(bb), bf16_hi(bb));
          }
          __syncwarp();
        }
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        if (warp == 4) {
          if (tile == cluster_id) G2_STAMP(8);
          G2_STAMP(10);
        }
        if (rows_ok) {
#pragma unroll 1
          for (int i = 0; i < 2; ++i) {
            const int u = unit_par + 2 * i;
            const int c = u * 64;
            const int n0 = tile_n0 + c;
            if (u >= U || n0 >= p.N) break;  // warp-uniform
            const bool is_tail = tail != 0 && u == U - 1;
            uint8_t* buf = out_stage + (ew * U + u) * 4096;
            const float* gs = reinterpret_cast<const float*>(gstage + (ew * U + u) * G2_GS_SLOT);
            uint32_t r0[32], r1[32];
            tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(c), r0);
            tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(c + 32), r1);  // may run past bn: still inside the buffer
            // while the accumulator columns are on their way: this thread's residual row out of the staging tile
            mbar_wait(&my_rbar[i], (rphase >> i) & 1u);
            rphase ^= 1u << i;
            int nchunk = is_tail ? (tail >> 3) : 8;                    // 16-byte pieces of a row
            if (((p.N - n0) >> 3) < nchunk) nchunk = (p.N - n0) >> 3;  // columns past N: clipped by the store anyway
            if (row >= p.M) nchunk = 0;
            uint8_t* rowp = buf + lane * (is_tail ? tail * 2 : 128);
            uint4 rr[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (j < nchunk) rr[j] = *reinterpret_cast<const uint4*>(rowp + (is_tail ? (j << 4) : ((j ^ (lane & 7)) << 4)));
            }
            tmem_ld_wait();
            {
              const float* bsrc = gs + 128;
              const bool has_bias = p.bias != nullptr;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (j < nchunk) {
                  float4 g0, g1;
                  if (gate_staged) {  // warp-uniform
                    g0 = *reinterpret_cast<const float4*>(gs + gsel + j * 8);
                    g1 = *reinterpret_cast<const float4*>(gs + gsel + j * 8 + 4);
                  } else {  // three or more gate rows inside 32 output rows (tiny test shapes): straight from global
                    g0 = *reinterpret_cast<const float4*>(gpl + n0 + j * 8);
                    g1 = *reinterpret_cast<const float4*>(gpl + n0 + j * 8 + 4);
                  }
                  // packed fp32 pairs (FADD2 / FFMA2: the same round-to-nearest results as the scalar fadd / fma of
                  // the generic epilogue, half the issue slots)
                  f32x2 x0 = pk2(__uint_as_float(j < 4 ? r0[(j * 8 + 0) & 31] : r1[(j * 8 + 0) & 31]),
                                 __uint_as_float(j < 4 ? r0[(j * 8 + 1) & 31] : r1[(j * 8 + 1) & 31]));
                  f32x2 x1 = pk2(__uint_as_float(j < 4 ? r0[(j * 8 + 2) & 31] : r1[(j * 8 + 2) & 31]),
                                 __uint_as_float(j < 4 ? r0[(j * 8 + 3) & 31] : r1[(j * 8 + 3) & 31]));
                  f32x2 x2 = pk2(__uint_as_float(j < 4 ? r0[(j * 8 + 4) & 31] : r1[(j * 8 + 4) & 31]),
                                 __uint_as_float(j < 4 ? r0[(j * 8 + 5) & 31] : r1[(j * 8 + 5) & 31]));
                  f32x2 x3 = pk2(__uint_as_float(j < 4 ? r0[(j * 8 + 6) & 31] : r1[(j * 8 + 6) & 31]),
                                 __uint_as_float(j < 4 ? r0[(j * 8 + 7) & 31] : r1[(j * 8 + 7) & 31]));
                  if (has_bias) {
                    const float4 b0 = *reinterpret_cast<const float4*>(bsrc + j * 8);
                    const float4 b1 = *reinterpret_cast<const float4*>(bsrc + j * 8 + 4);
                    x0 = add2p(x0, pk2(b0.x, b0.y)); x1 = add2p(x1, pk2(b0.z, b0.w));
                    x2 = add2p(x2, pk2(b1.x, b1.y)); x3 = add2p(x3, pk2(b1.z, b1.w));
                  }
                  x0 = fma2p(x0, pk2(g0.x, g0.y), pk2(bf16_lo(rr[j].x), bf16_hi(rr[j].x)));
                  x1 = fma2p(x1, pk2(g0.z, g0.w), pk2(bf16_lo(rr[j].y), bf16_hi(rr[j].y)));
                  x2 = fma2p(x2, pk2(g1.x, g1.y), pk2(bf16_lo(rr[j].z), bf16_hi(rr[j].z)));
                  x3 = fma2p(x3, pk2(g1.z, g1.w), pk2(bf16_lo(rr[j].w), bf16_hi(rr[j].w)));
                  float lo, hi;
                  uint4 o;
                  upk2(x0, lo, hi); o.x = pack_bf16(lo, hi);
                  upk2(x1, lo, hi); o.y = pack_bf16(lo, hi);
                  upk2(x2, lo, hi); o.z = pack_bf16(lo, hi);
                  upk2(x3, lo, hi); o.w = pack_bf16(lo, hi);
                  sts128_nobarrier(rowp + (is_tail ? (j << 4) : ((j ^ (lane & 7)) << 4)), o);
                }
              }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(is_tail ? &tma_t : &tma_o, buf, n0, row0);
              bulk_commit_group();
            }
            stores_out = true;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
        if (warp == 4) {
          if (tile == cluster_id) G2_STAMP(9);
          G2_STAMP(11);
          G2_STAMP_VALUE(14, (tile - cluster_id) / num_clusters + 1);
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    } else
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const G2Tile tl = g2_tile(p, tile, bn, num_clusters);
      if (tl.w == 0) continue;
      const int m_blk = tl.m_blk;
      const int tw = tl.w;  // this tile's width
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if (warp == 4) {
        if (tile == cluster_id) G2_STAMP(8);  // first accumulator complete
        G2_STAMP(10);                         // (overwritten per tile: the last accumulator complete)
      }
      const int row0 = m_blk * 256 + static_cast<int>(rank) * BM + ew * 32;
      const int row = row0 + lane;
      const uint32_t taddr = tmem_base + static_cast<uint32_t>(acc * G2_ACC_COLS) + (static_cast<uint32_t>(ew * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < tw; c += 64) {
        const int n0 = tl.n0 + c;
        if (n0 >= p.N || row0 >= p.M) break;  // warp-uniform
        if (EW == 8 && ((c >> 6) & 1) != unit_par) continue;
        uint32_t r0[32], r1[32];
        tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(c), r0);
        tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(c + 32), r1);  // may run past the tile: still inside the buffer
        tmem_ld_wait();
        // Units that lie completely inside this tile go through shared memory + TMA (columns past N and rows past M are
        // clipped by the tensor map); the narrower last unit of a tile whose width is not a multiple of 64 must not
        // touch its neighbour's columns and is stored directly.
        const bool staged = p.tma_store && (tw - c >= 64);
        uint8_t* sbuf = my_stage + (stores % NBUF) * (32 * 128);
        if (staged && stores >= NBUF) {
          if (lane == 0) bulk_wait_group_read<NBUF - 1>();  // the store that last used this buffer has read it
          __syncwarp();
        }
        if (row < p.M) {
          float v[64];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = __uint_as_float(r0[j]);
            v[32 + j] = __uint_as_float(r1[j]);
          }
          int ncols = tw - c;
          if (ncols > 64) ncols = 64;
          if (p.N - n0 < ncols) ncols = p.N - n0;
          epilogue_unit<EPI>(p, v, row, n0, ncols, staged ? sbuf + lane * 128 : nullptr, lane & 7);
        }
        if (staged) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tma_o, sbuf, n0, row0);
            bulk_commit_group();
          }
          ++stores;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
      if (warp == 4) {
        if (tile == cluster_id) G2_STAMP(9);  // first tile's epilogue units done (this warp)
        G2_STAMP(11);                         // (overwritten per tile: the last tile's)
        G2_STAMP_VALUE(14, (tile - cluster_id) / num_clusters + 1);  // tiles this cluster has processed
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  // ---- teardown: nobody may exit (or free TMEM) while the peer can still touch this CTA's memory ----
  // (the staged output tiles must have been READ out of shared memory before the CTA may go away; the writes
  //  themselves complete like any other store before the grid counts as finished)
  if (warp >= 4 && lane == 0) bulk_wait_group_read<0>();
  if (warp == 4) G2_STAMP(12);  // this warp's output stores complete
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) G2_STAMP(13);  // teardown barrier passed
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}


// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
template <int BN, int EPI>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmDev& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  auto kern = gemm_bf16_kernel<BN, EPI>;
  if (!attr_set) {
    ORVB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  int tiles = p.num_m_tiles * p.num_n_tiles;
  int grid = tiles < sm_count() ? tiles : sm_count();
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(ta, tb, p);
  ORVB_CHECK_CUDA(cudaGetLastError());
  return ORVB_OK;
}

// Shared-memory bytes of a CTA-pair launch with this plan (mirrors the pointer arithmetic at the top of the kernel).
static int g2_smem_bytes(const GemmDev& p, int bn) {
  const int gs = p.fast_resid ? 4 * ((bn + 63) / 64) * G2_GS_SLOT : 0;
  return 1024 + G2_STAGES * p.stage_bytes + p.out_stage_bytes + gs + G2_BAR_BYTES;
}

template <int EPI>
static int launch_gemm2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tt,
                        const GemmDev& p, int bn, cudaStream_t stream) {
  static bool attr_set = false;
  auto kern = gemm2_bf16_kernel<EPI>;
  if (!attr_set) {
    ORVB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_MAX));
    attr_set = true;
  }
  const int tiles = p.rem_width > 0 ? p.num_m_tiles * (p.n_full + 1) : p.num_m_tiles * p.num_n_tiles;
  const int clusters = sm_count() / 2;
  const int grid = 2 * (tiles < clusters ? tiles : clusters);  // (a mixed tile list has at least `clusters` full tiles)
  const int smem = g2_smem_bytes(p, bn);
  ORVB_REQUIRE(smem <= G2_SMEM_MAX, ORVB_EINVAL, "gemm: shared-memory plan of %d bytes does not fit", smem);
  ORVB_CHECK_CUDA(launch_kernel(kern, dim3(grid), dim3(g2_threads(EPI)), smem, stream, true, ta, tb, to, tt, p, bn));
  return ORVB_OK;
}

static int launch_gemm2_epi(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to,
                            const CUtensorMap& tt, const GemmDev& p, int bn, cudaStream_t stream) {
  switch (epi) {
    case ORVB_EPI_BIAS: return launch_gemm2<ORVB_EPI_BIAS>(ta, tb, to, tt, p, bn, stream);
    case ORVB_EPI_GELU: return launch_gemm2<ORVB_EPI_GELU>(ta, tb, to, tt, p, bn, stream);
    case ORVB_EPI_GATE_RESID: return launch_gemm2<ORVB_EPI_GATE_RESID>(ta, tb, to, tt, p, bn, stream);
    case ORVB_EPI_QKV: return launch_gemm2<ORVB_EPI_QKV>(ta, tb, to, tt, p, bn, stream);
  }
  set_error("orvb_gemm_bf16: unknown epilogue %d", epi);
  return ORVB_EINVAL;
}

template <int BN>
static int launch_gemm_epi(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmDev& p,
                           cudaStream_t stream) {
  switch (epi) {
    case ORVB_EPI_BIAS: return launch_gemm<BN, ORVB_EPI_BIAS>(ta, tb, p, stream);
    case ORVB_EPI_GELU: return launch_gemm<BN, ORVB_EPI_GELU>(ta, tb, p, stream);
    case ORVB_EPI_GATE_RESID: return launch_gemm<BN, ORVB_EPI_GATE_RESID>(ta, tb, p, stream);
    case ORVB_EPI_QKV: return launch_gemm<BN, ORVB_EPI_QKV>(ta, tb, p, stream);
  }
  set_error("orvb_gemm_bf16: unknown epilogue %d", epi);
  return ORVB_EINVAL;
}

// ORVB_GEMM_MIXED_TILES=0: uniform tile widths only (A/B of the full + remainder tile list).
static bool mixed_tiles_enabled() {
  const char* e = getenv("ORVB_GEMM_MIXED_TILES");
  return !(e != nullptr && e[0] == '0');
}
static bool fast_resid_enabled() {
  const char* e = getenv("ORVB_GEMM_FAST_RESID");  // read per call: the tests switch it inside one process
  return !(e != nullptr && e[0] == '0');
}

// Picks the N tile that minimises (waves x tile cost) on this GPU for a persistent 1-CTA/SM launch.
int gemm_pick_bn(int m, int n) {
  // Cost model fitted to B200 measurements of this kernel (profiles/r01_gemm_tiles.md): one wave of 128 x BN
  // tiles costs ~ (7.6 + 0.0184 * BN) units, nearly independent of how full the wave is (the mainloop is
  // L2->SM bandwidth bound: 16 KB of A + BN*128 B of B per 128 x BN x 64 step).
  const int sms = sm_count();
  const int mt = (m + BM - 1) / BM;
  int best_bn = 64;
  double best = 1e30;
  const int cands[4] = {256, 192, 128, 64};
  for (int i = 0; i < 4; ++i) {
    int bn = cands[i];
    if (bn > 64 && n < bn) continue;
    int nt = (n + bn - 1) / bn;
    long tiles = static_cast<long>(mt) * nt;
    long waves = (tiles + sms - 1) / sms;
    double cost = waves * (7.6 + 0.0184 * bn);
    if (cost < best) {
      best = cost;
      best_bn = bn;
    }
  }
  return best_bn;
}

// Tile width for the CTA-pair kernel: 256-row tiles over `clusters` pairs.  A wave of 256 x bn tiles costs about
// (G2_WAVE_FIXED + bn) column units (operand streaming of the fixed 256 rows of A + per-tile pipeline turnaround);
// widths are multiples of 16 (64 when the epilogue normalises whole 64-wide heads).
constexpr int G2_WAVE_FIXED = 96;
// `rem_out` (may be NULL = uniform tiles only): width of the ONE narrower tile per 256-row block when cutting N into
// n / bn full tiles plus a remainder beats every uniform width (GemmDev::rem_width); 0 = uniform.  `fast_ok`: the GEMM is
// eligible for the prefetching gated-residual epilogue, which needs a tile of at most three 64-column units: a width
// <= 192 within 3 % of the best modelled cost is preferred (the model does not see the epilogue).
int gemm_pick_bn_pair2(int m, int n, int k, int epi, bool fast_ok, int* rem_out) {
  const int clusters = sm_count() / 2;
  const int mt = (m + 255) / 256;
  const int step = (epi == ORVB_EPI_QKV) ? 64 : 16;
  int best_bn = 256, best_rem = 0;
  double best = 1e30;
  int fast_bn = 0;
  double fast_cost = 1e30;
  for (int bn = 256; bn >= 64; bn -= step) {
    const int nt = (n + bn - 1) / bn;
    const long tiles = static_cast<long>(mt) * nt;
    const long waves = (tiles + clusters - 1) / clusters;
    const double cost = static_cast<double>(waves) * (G2_WAVE_FIXED + bn);
    if (cost < best - 1e-9) {
      best = cost;
      best_bn = bn;
      best_rem = 0;
    }
    if (fast_ok && bn <= 192 && cost < fast_cost - 1e-9) {
      fast_cost = cost;
      fast_bn = bn;
    }
    const int q = n / bn, r = n % bn;
    if (rem_out != nullptr && !fast_ok && r >= 32 && r % step == 0 && static_cast<long>(mt) * q >= clusters) {
      const long full = static_cast<long>(mt) * q;
      const long base = full / clusters, extra = full % clusters;
      const long n_light = clusters - extra;
      const long per_light = (mt + n_light - 1) / n_light;
      const double heavy = extra > 0 ? static_cast<double>(base + 1) * (G2_WAVE_FIXED + bn) : 0.0;
      const double light = static_cast<double>(base) * (G2_WAVE_FIXED + bn) + static_cast<double>(per_light) * (G2_WAVE_FIXED + r);
      const double c2 = heavy > light ? heavy : light;
      if (c2 < best - 1e-9) {
        best = c2;
        best_bn = bn;
        best_rem = r;
      }
    }
  }
  // How much modelled mainloop cost the prefetching epilogue is worth: with a short K the generic epilogue of a 4-unit
  // tile is as long as the tile's mainloop (config 4's attn-out, K = 3072: 72.8 us at 240 columns, 66.0 us at 192 with the
  // prefetching epilogue, profiles/r02zz_cfg4_pref*.json); with K = 4 D it hides behind the mainloop and the wider tile
  // wins (FF2: 226.7 vs 228.1 us).  ORVB_GEMM_FAST_PREF overrides both.
  static const double pref_env = [] {
    const char* e = getenv("ORVB_GEMM_FAST_PREF");
    const double v = e != nullptr ? atof(e) : 0.0;
    return v >= 1.0 ? v : 0.0;
  }();
  const double fast_pref = pref_env > 0.0 ? pref_env : (k <= 4096 ? 1.20 : 1.03);
  if (fast_ok && fast_bn > 0 && fast_cost <= fast_pref * best) {
    best_bn = fast_bn;
    best_rem = 0;
  }
  if (rem_out != nullptr) *rem_out = best_rem;
  return best_bn;
}
int gemm_pick_bn_pair(int m, int n, int epi) { return gemm_pick_bn_pair2(m, n, 1 << 30, epi, false, nullptr); }

// bn > 0: 1-CTA kernel with that N tile; bn < 0: CTA-pair kernel with N tile -bn.
int gemm_launch_prepared(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tt,
                         const GemmDev& p, int bn, int epi, cudaStream_t stream) {
  if (bn < 0) return launch_gemm2_epi(epi, ta, tb, to, tt, p, -bn, stream);
  switch (bn) {
    case 256: return launch_gemm_epi<256>(epi, ta, tb, p, stream);
    case 192: return launch_gemm_epi<192>(epi, ta, tb, p, stream);
    case 128: return launch_gemm_epi<128>(epi, ta, tb, p, stream);
    case 64: return launch_gemm_epi<64>(epi, ta, tb, p, stream);
  }
  set_error("gemm: unsupported BN %d", bn);
  return ORVB_EINVAL;
}

int gemm_prepare(const orvb_gemm_args* a, int bn_override, CUtensorMap* ta, CUtensorMap* tb, CUtensorMap* to, CUtensorMap* tt,
                 GemmDev* p, int* bn_out) {
  ORVB_REQUIRE(a != nullptr && a->a && a->w && a->out, ORVB_EINVAL, "orvb_gemm_bf16: null pointer");
  ORVB_REQUIRE(a->m > 0 && a->n > 0 && a->k > 0, ORVB_ESHAPE, "orvb_gemm_bf16: empty problem m=%d n=%d k=%d", a->m,
               a->n, a->k);
  ORVB_REQUIRE(a->k % 8 == 0 && a->n % 8 == 0 && a->lda % 8 == 0 && a->ldw % 8 == 0 && a->ldo % 8 == 0, ORVB_ESHAPE,
               "orvb_gemm_bf16: k, n and row pitches must be multiples of 8 (k=%d n=%d lda=%d ldw=%d ldo=%d)", a->k,
               a->n, a->lda, a->ldw, a->ldo);
  ORVB_REQUIRE((reinterpret_cast<uintptr_t>(a->a) | reinterpret_cast<uintptr_t>(a->w) |
                reinterpret_cast<uintptr_t>(a->out)) % 16 == 0,
               ORVB_ESHAPE, "orvb_gemm_bf16: base pointers must be 16-byte aligned");
  if (a->epilogue == ORVB_EPI_QKV) {
    ORVB_REQUIRE(a->qk_dim > 0 && a->qk_dim % 64 == 0 && a->n == 3 * a->qk_dim, ORVB_ESHAPE,
                 "orvb_gemm_bf16: QKV epilogue needs n == 3*qk_dim, qk_dim %% 64 == 0");
    ORVB_REQUIRE(a->q_norm_w && a->q_norm_b && a->k_norm_w && a->k_norm_b, ORVB_EINVAL,
                 "orvb_gemm_bf16: QKV epilogue needs the q/k norm parameters");
  }
  if (a->epilogue == ORVB_EPI_GATE_RESID) {
    ORVB_REQUIRE(a->resid == nullptr || a->ldr % 8 == 0, ORVB_ESHAPE, "orvb_gemm_bf16: ldr must be a multiple of 8");
    ORVB_REQUIRE(a->gate == nullptr || (a->gate_ld % 4 == 0 && a->gate_text_off % 4 == 0 && a->gate_video_off % 4 == 0),
                 ORVB_ESHAPE, "orvb_gemm_bf16: gate pitch/offsets must be multiples of 4");
  }
  // More than one 128-row tile: CTA pairs (M = 256 MMAs); otherwise the 1-CTA kernel.
  int bn, rem = 0;
  const bool fast_ok = a->epilogue == ORVB_EPI_GATE_RESID && a->resid != nullptr && a->resid == a->out && a->ldr == a->ldo &&
                       a->resid_mod == 0 && a->src_rows == 0 && a->mv_tokens == 0 && !a->out_f32 && fast_resid_enabled();
  if (bn_override != 0) bn = bn_override;
  else if (a->m > BM) bn = -gemm_pick_bn_pair2(a->m, a->n, a->k, a->epilogue, fast_ok, mixed_tiles_enabled() ? &rem : nullptr);
  else bn = gemm_pick_bn(a->m, a->n);
  const bool pair = bn < 0;
  if (pair) {
    const int w = -bn;
    ORVB_REQUIRE(w >= 32 && w <= 256 && w % 16 == 0, ORVB_EINVAL, "gemm: pair tile width %d must be a multiple of 16 in [32, 256]", w);
    ORVB_REQUIRE(a->epilogue != ORVB_EPI_QKV || w % 64 == 0, ORVB_EINVAL, "gemm: the QKV epilogue needs a tile width that is a multiple of 64 (got %d)", w);
  }
  const bool maps = ta != nullptr;  // planning only (orvb_gemm_tile_list, no driver needed) when the maps are not asked for
  int rc = maps ? make_tmap_2d_bf16(ta, a->a, a->m, a->k, a->lda, BM, BK) : ORVB_OK;
  if (rc != ORVB_OK) return rc;
  ORVB_REQUIRE(a->k_wrap == 0 || (a->k_wrap > 0 && a->k_wrap % BK == 0 && a->k % a->k_wrap == 0), ORVB_ESHAPE,
               "orvb_gemm_bf16: k_wrap %d must be a multiple of %d dividing k = %d", a->k_wrap, BK, a->k);
  if (maps) rc = make_tmap_2d_bf16(tb, a->w, a->n, a->k_wrap > 0 ? a->k_wrap : a->k, a->ldw, pair ? (-bn) / 2 : bn, BK);
  if (rc != ORVB_OK) return rc;
  // Output through TMA (pair kernel, rows written in place): [32 rows x 64 cols] boxes of the [M, N] output
  const bool tma_store = pair && a->src_rows == 0 && a->mv_tokens == 0 && !a->out_f32;
  if (maps) {
    if (tma_store) {
      rc = make_tmap_2d_bf16(to, a->out, a->m, a->n, a->ldo, 32, 64);
      if (rc != ORVB_OK) return rc;
    } else {
      *to = *ta;  // unused placeholder
    }
  }
  GemmDev d;
  d.k_wrap = a->k_wrap;
  d.stage_bytes = G2_STAGE_BYTES;
  d.out_stage_bytes = G2_OUT_STAGE_BYTES;
  d.fast_resid = 0;
  if (maps) *tt = *to;  // (placeholder unless the fast epilogue needs a tail map)
  // Gated residual written in place over its own residual (attn-out / FF2 of every block): the epilogue that prefetches
  // the residual tiles by TMA.  One staging tile per 64-column unit of the CTA's half tile -> needs the room a narrower
  // B stage leaves (tile widths up to 192).  ORVB_GEMM_FAST_RESID=0 keeps the generic epilogue (A/B, bit-identical).
  if (pair && tma_store && fast_ok) {
    const int w = -bn, units = (w + 63) / 64, tail = w % 64;
    GemmDev f = d;
    f.stage_bytes = G2_A_BYTES + (((w / 2) * BK * 2 + 1023) & ~1023);
    f.out_stage_bytes = 4 * units * 4096;
    f.fast_resid = 1;
    if (units <= 3 && g2_smem_bytes(f, w) <= G2_SMEM_MAX) {
      d.stage_bytes = f.stage_bytes;
      d.out_stage_bytes = f.out_stage_bytes;
      d.fast_resid = 1;
      if (tail != 0 && maps) {
        rc = make_tmap_2d_bf16_plain(tt, a->out, a->m, a->n, a->ldo, 32, tail);
        if (rc != ORVB_OK) return rc;
      }
    }
  }
  d.tma_store = tma_store ? 1 : 0;
  d.out_f32 = a->out_f32 ? 1 : 0;
  d.grp_off = a->group_offset;
  d.M = a->m; d.N = a->n; d.K = a->k;
  d.out = static_cast<bf16*>(a->out); d.ldo = a->ldo;
  d.bias = static_cast<const bf16*>(a->bias);
  d.src_rows = a->src_rows; d.dst_rows = a->dst_rows; d.dst_offset = a->dst_offset;
  d.mv_tokens = a->mv_tokens; d.mv_frames = a->mv_frames; d.mv_views = a->mv_views;
  d.resid = static_cast<const bf16*>(a->resid); d.ldr = a->ldr;
  d.resid_mod = a->resid_mod; d.resid_views = a->resid_views; d.resid_view_stride = a->resid_view_stride;
  d.gate = a->gate; d.gate_ld = a->gate_ld; d.gate_text_off = a->gate_text_off; d.gate_video_off = a->gate_video_off;
  d.rm = a->rowmap;
  d.qk_dim = a->qk_dim;
  d.qw = static_cast<const bf16*>(a->q_norm_w); d.qb = static_cast<const bf16*>(a->q_norm_b);
  d.kw = static_cast<const bf16*>(a->k_norm_w); d.kb = static_cast<const bf16*>(a->k_norm_b);
  d.qk_eps = a->qk_eps;
  d.rope_cos = a->rope_cos; d.rope_sin = a->rope_sin;
  d.num_m_tiles = pair ? (a->m + 2 * BM - 1) / (2 * BM) : (a->m + BM - 1) / BM;
  d.num_n_tiles = pair ? (a->n - bn - 1) / (-bn) : (a->n + bn - 1) / bn;
  d.tile_end = d.num_m_tiles * d.num_n_tiles;
  d.n_full = 0;
  d.rem_width = 0;
  if (pair && rem > 0 && !d.fast_resid) {
    // n / bn full tiles + one narrower tile per 256-row block, the narrow ones placed on the clusters with one full tile
    // less (g2_tile): virtual indices [0, full) then rows of `clusters` slots of which the first n_light hold a tile
    const int clusters = sm_count() / 2;
    const int full = d.num_m_tiles * (a->n / -bn);
    const int n_light = clusters - full % clusters;
    d.n_full = a->n / -bn;
    d.rem_width = rem;
    d.tile_end = full + ((d.num_m_tiles + n_light - 1) / n_light) * clusters;
  }
  *p = d;
  *bn_out = bn;
  return ORVB_OK;
}

// The planning half of gemm_prepare alone: kernel / tile choice and the tile list, no tensor maps (no driver call).
int gemm_plan(const orvb_gemm_args* a, int bn_override, GemmDev* p, int* bn_out) {
  return gemm_prepare(a, bn_override, nullptr, nullptr, nullptr, nullptr, p, bn_out);
}

int gemm_run(const orvb_gemm_args* a, cudaStream_t stream) {
  CUtensorMap ta, tb, to, tt;
  GemmDev p;
  int bn;
  int rc = gemm_prepare(a, 0, &ta, &tb, &to, &tt, &p, &bn);
  if (rc != ORVB_OK) return rc;
  return gemm_launch_prepared(ta, tb, to, tt, p, bn, a->epilogue, stream);
}


}  // namespace orvb

extern "C" int orvb_gemm_bf16(const orvb_gemm_args* args, void* stream) {
  using namespace orvb;
  int rc = check_arch();
  if (rc != ORVB_OK) return rc;
  CUtensorMap ta, tb, to, tt;
  GemmDev p;
  int bn;
  rc = gemm_prepare(args, 0, &ta, &tb, &to, &tt, &p, &bn);
  if (rc != ORVB_OK) return rc;
  return gemm_launch_prepared(ta, tb, to, tt, p, bn, args->epilogue, static_cast<cudaStream_t>(stream));
}

// Test hook: same as orvb_gemm_bf16 with a forced tile: bn in {64,128,192,256} = 1-CTA kernel, -bn (multiple of 16,
// 32..256) = CTA-pair kernel.
extern "C" int orvb_gemm_bf16_bn(const orvb_gemm_args* args, int bn, void* stream) {
  using namespace orvb;
  int rc = check_arch();
  if (rc != ORVB_OK) return rc;
  CUtensorMap ta, tb, to, tt;
  GemmDev p;
  int bn_used;
  rc = gemm_prepare(args, bn, &ta, &tb, &to, &tt, &p, &bn_used);
  if (rc != ORVB_OK) return rc;
  return gemm_launch_prepared(ta, tb, to, tt, p, bn_used, args->epilogue, static_cast<cudaStream_t>(stream));
}

#ifdef ORVB_GEMM_TIMELINE
// Measurement hook (timeline builds only): installs (or clears, with NULL) a device buffer of 128 int64 that the first
// and the last cluster of the next CTA-pair GEMM launches fill with clock64() stamps.
extern "C" int orvb_gemm_set_debug(void* dev_buf) {
  long long* ptr = static_cast<long long*>(dev_buf);
  ORVB_CHECK_CUDA(cudaMemcpyToSymbol(orvb::g_gemm_dbg, &ptr, sizeof(ptr)));
  return ORVB_OK;
}
#endif

// Which kernel / tile width orvb_gemm_bf16 picks for an [m, n] output on this device: > 0 = single-CTA kernel with that
// N tile, < 0 = CTA-pair kernel with N tile -value.  Pure host arithmetic (148 SMs assumed when no GPU is present).
extern "C" int orvb_gemm_tile_width(int32_t m, int32_t n, int32_t epilogue) {
  using namespace orvb;
  if (m <= 0 || n <= 0) return 0;
  int rem = 0;
  return m > BM ? -gemm_pick_bn_pair2(m, n, 1 << 30, epilogue, false, mixed_tiles_enabled() ? &rem : nullptr) : gemm_pick_bn(m, n);
}

// The tile list orvb_gemm_bf16 would run for this problem on the CTA-pair kernel, in the order the clusters walk it
// (host arithmetic only; the same g2_tile() the kernel decodes its virtual tile indices with): up to `capacity` records
// of {cluster, first row, first column, width}.  Returns the number of tiles (0 for the single-CTA kernel), or a negative
// error code.  `in_place_resid` = the GATE_RESID call writes over its own residual (eligible for the prefetching epilogue).
extern "C" int orvb_gemm_tile_list(int32_t m, int32_t n, int32_t k, int32_t epilogue, int32_t in_place_resid, int32_t* out,
                                   int32_t capacity) {
  using namespace orvb;
  if (m <= BM || n <= 0) return 0;
  orvb_gemm_args a;
  memset(&a, 0, sizeof(a));
  // pointers are only compared / checked for alignment, never dereferenced, by the planning part of gemm_prepare
  static uint8_t dummy[64] __attribute__((aligned(64)));
  a.a = dummy; a.w = dummy; a.out = dummy;
  a.m = m; a.n = n; a.k = k; a.lda = k; a.ldw = k; a.ldo = n; a.epilogue = epilogue;
  if (epilogue == ORVB_EPI_QKV) {
    a.qk_dim = n / 3;
    a.q_norm_w = a.q_norm_b = a.k_norm_w = a.k_norm_b = dummy;
  }
  if (in_place_resid) { a.resid = dummy; a.ldr = n; }
  GemmDev p;
  int bn = 0;
  const int rc = gemm_plan(&a, 0, &p, &bn);
  if (rc != ORVB_OK) return rc;
  if (bn >= 0) return 0;
  const int clusters_all = sm_count() / 2;
  const int real_tiles = p.rem_width > 0 ? p.num_m_tiles * (p.n_full + 1) : p.num_m_tiles * p.num_n_tiles;
  const int clusters = real_tiles < clusters_all ? real_tiles : clusters_all;
  int count = 0;
  for (int c = 0; c < clusters; ++c) {
    for (int v = c; v < p.tile_end; v += clusters) {
      const G2Tile t = g2_tile(p, v, -bn, clusters);
      if (t.w == 0) continue;
      if (out != nullptr && count < capacity) {
        out[4 * count + 0] = c;
        out[4 * count + 1] = t.m_blk * 256;
        out[4 * count + 2] = t.n0;
        out[4 * count + 3] = t.w;
      }
      ++count;
    }
  }
  return count;
}

// Width of the one narrower tile per 256-row block when orvb_gemm_bf16 cuts N into n / width full tiles plus a remainder
// (0 = uniform tiles; always 0 for the single-CTA kernel).
extern "C" int orvb_gemm_tile_remainder(int32_t m, int32_t n, int32_t epilogue) {
  using namespace orvb;
  if (m <= BM || n <= 0 || !mixed_tiles_enabled()) return 0;
  int rem = 0;
  (void)gemm_pick_bn_pair2(m, n, 1 << 30, epilogue, false, &rem);
  return rem;
}
