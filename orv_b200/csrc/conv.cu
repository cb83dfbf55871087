// Causal 3-D (and per-frame 2-D) convolution over channels-last bf16 activations as an IMPLICIT GEMM on the CTA-pair
// tcgen05 mainloop of gemm.cu — the contraction of the 3-D VAE decoder (diffusers
// AutoencoderKLCogVideoX.decode, called by the reference at orv/models/cogvideox_control.py:1095-1100, :1476-1479).
//
//   out[t, h, w, n] = bias[n] + sum_{kt, kh, kw, c} x[t + kt - (KT-1), h + kh - KH/2, w + kw - KW/2, c] * W[n, (kt, kh, kw, c)]
//
// GEMM view: M = output pixels, N = C_out, K = taps * C_in.  An M tile of 128 rows is an 8 x 16 PATCH of one frame, so
// the A operand of one (tap, 64-channel) K step is ONE 4-D TMA box {64 c, 16 w, 8 h, 1 t} of the activation tensor at
// the tap's offset: out-of-bounds pixels are zero-filled by the TMA unit (= the convolution's spatial zero padding),
// the box lands in shared memory as 128 rows of 128 bytes, 128B-swizzled — exactly the K-major tile tcgen05 consumes.
// No im2col buffer exists anywhere.  Temporal padding is causal (CogVideoXCausalConv3d, pad_mode "constant"): a tap
// that reaches in front of the first frame reads the convolution cache (the last KT-1 input frames of the previous frame
// batch, a second tensor map) or, without a cache, the first frame again.
// Everything else — 6-stage ring, cta_group::2 MMAs of M = 256 (two patches per cluster), two TMEM accumulators, eight
// epilogue warps, swizzled staging + TMA store (a {64, 16, 2, 1} box per warp and 64-column unit, clipped at the image
// border) — is the GEMM kernel's structure (gemm.cu); the fused epilogues are the shared ones (bias, bias + residual).
#include <stdlib.h>

#include "common.cuh"
#include "gemm_common.cuh"
#include "ptx.cuh"

namespace orvb {

int gemm_pick_bn_pair(int m, int n, int epi);

struct ConvDev {
  int T, H, W, HB, WB;     // frames, image size, patch grid (HB = ceil(H/8), WB = ceil(W/16))
  int KT, KH, KW, kc;      // kernel taps, K steps per tap (C_in / 64)
  int num_patches;         // T * HB * WB
  int has_cache;
  // GroupNorm statistics of the OUTPUT, accumulated in the epilogue (gn_gs_log2 < 0: off): channels per group = 2^gs_log2,
  // gn_part[cta * 32 + group] = (sum, sum of squares) over the pixels that CTA stored
  int gn_gs_log2;
  double2* gn_part;
};

constexpr int CV_PH = 8, CV_PW = 16;  // patch = 8 rows x 16 columns = 128 pixels = one M tile

// "Halo" variant for 3 x 3 spatial taps and C_out <= 128.  With N = 128 a K step of the plain kernel moves 24 KB of fills
// + 24 KB of operand reads through a CTA's shared memory in the 256 clk its MMAs take = 192 B/clk against the 128 B/clk
// the SM has: those layers (42 % of the decoder's FLOPs) ran at 65 % of the tensor peak.  Here the M tile is a 16 x 8
// patch (16 image rows of 8 pixels = 16 eight-row core groups of 1024 bytes), and ONE TMA box of 18 rows x 8 pixels per
// (temporal tap, 64-channel chunk, horizontal tap) serves the three vertical taps: the A operand of vertical tap dh is
// the same shared-memory tile read from row dh on — a descriptor start address moved by dh x 1024 bytes, which keeps the
// 128-byte swizzle phase, so nothing but the standard descriptor is needed.  A fills drop from 16 to 6 KB per tap.
constexpr int CH_PH = 16, CH_PW = 8;
constexpr int CH_STAGES = 4;
constexpr int CH_A_BYTES = (CH_PH + 2) * CH_PW * 128;      // 18 KB: rows h0 - 1 .. h0 + 16 of the patch's 8 columns
constexpr int CH_B_SLOT = 64 * 128;                        // 8 KB: this CTA's <= 64 weight rows of one tap
constexpr int CH_STAGE_BYTES = CH_A_BYTES + 3 * CH_B_SLOT;  // 42 KB
constexpr int CH_SMEM_BYTES = CH_STAGES * CH_STAGE_BYTES + G2_OUT_STAGE_BYTES + 1024 + 256;

__device__ __forceinline__ float cv_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Sum and sum of squares of one 64-column unit, per group of GS channels, over the warp's 32 pixels — of the values AS
// STORED (rounded to bf16: the next operator normalises the bf16 tensor).  Lane g keeps group g's running totals in
// fp64; which tiles a warp sees and in which order is fixed by the static tile schedule, so the result is deterministic.
template <int GS>
__device__ __forceinline__ void cv_gn_unit(const float (&v)[64], bool valid, int n0, int ncols, int lane, double& acc_s,
                                           double& acc_q) {
  constexpr int G = 64 / GS;
  const int g0 = n0 / GS;
#pragma unroll
  for (int k = 0; k < G; ++k) {
    if (k * GS >= ncols) break;  // (warp-uniform) the narrower last unit of a tile: its tail belongs to the next tile
    float s = 0.f, q = 0.f;
    if (valid) {
#pragma unroll
      for (int c = 0; c < GS; ++c) {
        const float r = __bfloat162float(__float2bfloat16(v[k * GS + c]));
        s += r;
        q = fmaf(r, r, q);
      }
    }
    s = cv_warp_sum(s);
    q = cv_warp_sum(q);
    if (lane == g0 + k) {
      acc_s += static_cast<double>(s);
      acc_q += static_cast<double>(q);
    }
  }
}

template <int EPI, bool HALO>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(g2_threads(EPI), 1)
conv2_bf16_kernel(const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_c,
                  const __grid_constant__ CUtensorMap tma_b, const __grid_constant__ CUtensorMap tma_o, const GemmDev p,
                  const ConvDev cv, const int bn) {
  constexpr int STAGES = HALO ? CH_STAGES : G2_STAGES;
  constexpr int STAGE_BYTES = HALO ? CH_STAGE_BYTES : G2_STAGE_BYTES;
  constexpr int PH = HALO ? CH_PH : CV_PH, PW = HALO ? CH_PW : CV_PW;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* out_stage = smem + STAGES * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(out_stage + G2_OUT_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;  // num_m_tiles counts patch PAIRS
  // K steps per tile: one per (tap, channel chunk); the halo variant handles the three vertical taps of a step together
  const int num_k = HALO ? cv.KT * cv.KW * cv.kc : cv.KT * cv.KH * cv.KW * cv.kc;
  const int half_bn = bn >> 1;

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_x);
    tma_prefetch_desc(&tma_b);
    if (cv.has_cache) tma_prefetch_desc(&tma_c);
    if (p.tma_store) tma_prefetch_desc(&tma_o);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * g2_epi_warps(EPI));
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_slot, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  // patch of this CTA inside tile `m_blk`: (frame, patch row, patch column); past the end -> every load is out of
  // bounds (zeros, but the byte count the barrier expects still arrives) and nothing is stored
  auto patch_of = [&](int m_blk, int& t, int& h0, int& w0) -> bool {
    const int patch = 2 * m_blk + static_cast<int>(rank);
    const int pp = patch < cv.num_patches ? patch : 0;
    t = pp / (cv.HB * cv.WB);
    const int r = pp - t * (cv.HB * cv.WB);
    h0 = (r / cv.WB) * PH;
    w0 = (r % cv.WB) * PW;
    return patch < cv.num_patches;
  };

  if (warp == 0) {
    // ================================ TMA producer (both CTAs) ================================
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t stage_tx = HALO ? 2u * static_cast<uint32_t>(CH_A_BYTES + 3 * half_bn * BK * 2)
                                   : 2u * static_cast<uint32_t>(G2_A_BYTES + half_bn * BK * 2);
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int m_blk = tile % p.num_m_tiles;
      const int n_blk = tile / p.num_m_tiles;
      int t, h0, w0;
      const bool live = patch_of(m_blk, t, h0, w0);
      const int b_row = n_blk * bn + static_cast<int>(rank) * half_bn;
      if (HALO) {
        // step order: temporal tap, channel chunk, horizontal tap; the three vertical taps share the step's A tile
        int kt_i = 0, cc = 0, kw_i = 0;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + CH_A_BYTES;
          const uint32_t leader_full = mapa_u32(smem_u32(&full_bar[stage]), 0);
          int ts = t + kt_i - (cv.KT - 1);
          const CUtensorMap* mp = &tma_x;
          if (ts < 0) {
            if (cv.has_cache) {
              mp = &tma_c;
              ts += cv.KT - 1;
            } else {
              ts = 0;
            }
          }
          if (!live) ts = cv.T + cv.KT;  // out of bounds in either map: zero fill
          if (elect_one()) {
            if (rank == 0) mbar_expect_tx(&full_bar[stage], stage_tx);
            tma_load_4d_pair(sa, mp, leader_full, cc * BK, w0 + kw_i - 1, h0 - 1, ts);
#pragma unroll
            for (int kh_i = 0; kh_i < 3; ++kh_i) {
              const int tap = (kt_i * 3 + kh_i) * 3 + kw_i;
              tma_load_2d_pair(sb + kh_i * CH_B_SLOT, &tma_b, leader_full, (tap * cv.kc + cc) * BK, b_row);
            }
          }
          __syncwarp();
          if (++kw_i == 3) {
            kw_i = 0;
            if (++cc == cv.kc) {
              cc = 0;
              ++kt_i;
            }
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        continue;
      }
      int tap = 0, cc = 0;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * STAGE_BYTES;
        uint8_t* sb = sa + G2_A_BYTES;
        const uint32_t leader_full = mapa_u32(smem_u32(&full_bar[stage]), 0);
        const int kt_i = tap / (cv.KH * cv.KW);
        const int rem = tap - kt_i * (cv.KH * cv.KW);
        const int kh_i = rem / cv.KW, kw_i = rem - kh_i * cv.KW;
        int ts = t + kt_i - (cv.KT - 1);
        const CUtensorMap* mp = &tma_x;
        if (ts < 0) {
          if (cv.has_cache) {
            mp = &tma_c;
            ts += cv.KT - 1;
          } else {
            ts = 0;
          }
        }
        if (!live) ts = cv.T + cv.KT;  // out of bounds in either map: zero fill
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx(&full_bar[stage], stage_tx);
          tma_load_4d_pair(sa, mp, leader_full, cc * BK, w0 + kw_i - (cv.KW >> 1), h0 + kh_i - (cv.KH >> 1), ts);
          tma_load_2d_pair(sb, &tma_b, leader_full, kb * BK, b_row);
        }
        __syncwarp();
        if (++cc == cv.kc) {
          cc = 0;
          ++tap;
        }
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ================================ MMA issuer (leader CTA) ================================
    const uint32_t idesc = umma_idesc_bf16(256, bn, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * G2_ACC_COLS);
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
        if (HALO) {
          if (elect_one()) {
#pragma unroll
            for (int kh_i = 0; kh_i < 3; ++kh_i) {
              // vertical tap kh_i: the patch's rows start kh_i image rows (= kh_i x 1024 bytes) into the haloed tile
              const uint64_t a_desc = umma_desc_sw128(sa + kh_i * (CH_PW * 128));
              const uint64_t b_desc = umma_desc_sw128(sa + CH_A_BYTES + kh_i * CH_B_SLOT);
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                umma_f16_ss_pair(d_tmem, a_desc + static_cast<uint64_t>(k * 2), b_desc + static_cast<uint64_t>(k * 2), idesc,
                                 static_cast<uint32_t>((kb | kh_i | k) != 0));
              }
            }
            tc_commit_pair(&empty_bar[stage], 3);
            if (kb == num_k - 1) tc_commit_pair(&tfull_bar[acc], 3);
          }
        } else {
          const uint64_t a_desc = umma_desc_sw128(sa);
          const uint64_t b_desc = umma_desc_sw128(sa + G2_A_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              umma_f16_ss_pair(d_tmem, a_desc + static_cast<uint64_t>(k * 2), b_desc + static_cast<uint64_t>(k * 2), idesc,
                               static_cast<uint32_t>((kb | k) != 0));
            }
            tc_commit_pair(&empty_bar[stage], 3);
            if (kb == num_k - 1) tc_commit_pair(&tfull_bar[acc], 3);
          }
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
    // ================================ epilogue (both CTAs) ====================================
    constexpr int EW = g2_epi_warps(EPI);
    constexpr int NBUF = 8 / EW;
    const int ew = (warp - 4) & 3;         // TMEM lane quarter = patch rows 2 ew, 2 ew + 1 (halo variant: 4 ew .. 4 ew + 3)
    const int unit_par = (warp - 4) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t stores = 0;
    double gn_s = 0.0, gn_q = 0.0;  // lane g: group g of the output's GroupNorm (cv.gn_part)
    uint8_t* my_stage = out_stage + (warp - 4) * (NBUF * 32 * 128);
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int m_blk = tile % p.num_m_tiles;
      const int n_blk = tile / p.num_m_tiles;
      int t, h0, w0;
      const bool live = patch_of(m_blk, t, h0, w0);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int hq = h0 + (HALO ? 4 : 2) * ew;                              // first image row of this lane quarter
      const int hh = hq + (HALO ? (lane >> 3) : (lane >> 4)), ww = w0 + (HALO ? (lane & 7) : (lane & 15));
      const bool pix_ok = live && hh < cv.H && ww < cv.W;
      const int row = pix_ok ? (t * cv.H + hh) * cv.W + ww : p.M;          // linear pixel index (row of out / resid)
      const bool warp_ok = live && hq < cv.H;                               // warp-uniform: any row of mine inside?
      const uint32_t taddr = tmem_base + static_cast<uint32_t>(acc * G2_ACC_COLS) + (static_cast<uint32_t>(ew * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < bn; c += 64) {
        const int n0 = n_blk * bn + c;
        if (n0 >= p.N || !warp_ok) break;  // warp-uniform
        if (EW == 8 && ((c >> 6) & 1) != unit_par) continue;
        uint32_t r0[32], r1[32];
        tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(c), r0);
        tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(c + 32), r1);
        tmem_ld_wait();
        const bool staged = p.tma_store && (bn - c >= 64);
        uint8_t* sbuf = my_stage + (stores % NBUF) * (32 * 128);
        if (staged && stores >= NBUF) {
          if (lane == 0) bulk_wait_group_read<NBUF - 1>();
          __syncwarp();
        }
        float v[64];
        if (row < p.M || staged) {  // (staged: rows outside the image still fill their slot; the store clips them)
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = __uint_as_float(r0[j]);
            v[32 + j] = __uint_as_float(r1[j]);
          }
          int ncols = bn - c;
          if (ncols > 64) ncols = 64;
          if (p.N - n0 < ncols) ncols = p.N - n0;
          if (row < p.M) {
            epilogue_unit<EPI, false>(p, v, row, n0, ncols, staged ? sbuf + lane * 128 : nullptr, lane & 7);
          }
        }
        if (cv.gn_gs_log2 >= 0) {  // warp-uniform: statistics of the values just stored, for the next GroupNorm
          const bool valid = row < p.M;
          int nc = bn - c;
          if (nc > 64) nc = 64;
          if (p.N - n0 < nc) nc = p.N - n0;
          switch (cv.gn_gs_log2) {
            case 1: cv_gn_unit<2>(v, valid, n0, nc, lane, gn_s, gn_q); break;
            case 2: cv_gn_unit<4>(v, valid, n0, nc, lane, gn_s, gn_q); break;
            case 3: cv_gn_unit<8>(v, valid, n0, nc, lane, gn_s, gn_q); break;
            case 4: cv_gn_unit<16>(v, valid, n0, nc, lane, gn_s, gn_q); break;
            case 5: cv_gn_unit<32>(v, valid, n0, nc, lane, gn_s, gn_q); break;
            default: cv_gn_unit<64>(v, valid, n0, nc, lane, gn_s, gn_q); break;
          }
        }
        if (staged) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&tma_o, sbuf, n0, w0, hq, t);
            bulk_commit_group();
          }
          ++stores;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (cv.gn_part != nullptr) {
      // one partial per CTA: the eight epilogue warps meet in the output staging area (every TMA store that read it has
      // finished) and warp 4 adds them up in warp order
      if (lane == 0) bulk_wait_group_read<0>();
      __syncwarp();
      named_bar_sync(1, 8 * 32);  // nobody stages output any more (a warp without units would otherwise get here early)
      double2* sh = reinterpret_cast<double2*>(out_stage);
      sh[(warp - 4) * 32 + lane] = make_double2(gn_s, gn_q);
      named_bar_sync(1, 8 * 32);
      if (warp == 4) {
        double s = 0.0, q = 0.0;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) {
          s += sh[w8 * 32 + lane].x;
          q += sh[w8 * 32 + lane].y;
        }
        cv.gn_part[static_cast<size_t>(blockIdx.x) * 32 + lane] = make_double2(s, q);
      }
    }
  }

  if (warp >= 4 && lane == 0) bulk_wait_group<0>();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// Folds the per-warp partials of the convolution's epilogue into (mean, rstd) per group: fp64, fixed order (lanes stride
// over the partials, xor tree), one block.
__global__ void __launch_bounds__(1024) conv_gn_finalize_kernel(const double2* __restrict__ part, int nparts, int groups,
                                                                double count, float eps, float* __restrict__ stats) {
  pdl_launch_dependents();
  pdl_wait();
  const int g = threadIdx.x >> 5, lane = threadIdx.x & 31;  // one warp per group
  if (g >= groups) return;
  double s = 0.0, q = 0.0;
  int i = lane;
  for (; i + 3 * 32 < nparts; i += 4 * 32) {  // four partials requested per trip
    double2 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = part[static_cast<size_t>(i + k * 32) * 32 + g];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s += v[k].x;
      q += v[k].y;
    }
  }
  for (; i < nparts; i += 32) {
    const double2 v = part[static_cast<size_t>(i) * 32 + g];
    s += v.x;
    q += v.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (lane == 0) {
    const double mean = s / count;
    double var = q / count - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[2 * g] = static_cast<float>(mean);
    stats[2 * g + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
}

template <int EPI, bool HALO>
static int launch_conv(const CUtensorMap& tx, const CUtensorMap& tc, const CUtensorMap& tb, const CUtensorMap& to,
                       const GemmDev& p, const ConvDev& cv, int bn, cudaStream_t stream, int* grid_out) {
  static bool attr_set = false;
  auto kern = conv2_bf16_kernel<EPI, HALO>;
  constexpr int SMEM = HALO ? CH_SMEM_BYTES : G2_SMEM_BYTES;
  if (!attr_set) {
    ORVB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_set = true;
  }
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int clusters = sm_count() / 2;
  const int grid = 2 * (tiles < clusters ? tiles : clusters);
  ORVB_CHECK_CUDA(launch_kernel(kern, dim3(grid), dim3(g2_threads(EPI)), SMEM, stream, true, tx, tc, tb, to, p, cv, bn));
  if (grid_out != nullptr) *grid_out = grid;
  return ORVB_OK;
}

}  // namespace orvb

extern "C" int orvb_conv_cl(const orvb_conv_args* a, void* stream) {
  using namespace orvb;
  int rc = check_arch();
  if (rc != ORVB_OK) return rc;
  ORVB_REQUIRE(a != nullptr && a->x && a->w && a->out, ORVB_EINVAL, "orvb_conv_cl: null pointer");
  ORVB_REQUIRE(a->frames > 0 && a->height > 0 && a->width > 0, ORVB_ESHAPE, "orvb_conv_cl: empty activation tensor");
  ORVB_REQUIRE(a->c_in > 0 && a->c_in % 64 == 0, ORVB_ESHAPE,
               "orvb_conv_cl: c_in (%d) must be a multiple of 64 (pad the channels with zeros)", a->c_in);
  ORVB_REQUIRE(a->c_out > 0 && a->c_out % 8 == 0, ORVB_ESHAPE, "orvb_conv_cl: c_out (%d) must be a multiple of 8", a->c_out);
  ORVB_REQUIRE(a->kt >= 1 && a->kt <= 4 && a->kh >= 1 && a->kw >= 1 && (a->kh & 1) && (a->kw & 1) && a->kh <= 7 && a->kw <= 7,
               ORVB_ESHAPE, "orvb_conv_cl: kernel %dx%dx%d not supported (odd spatial extents <= 7, kt <= 4)", a->kt, a->kh, a->kw);
  ORVB_REQUIRE(a->cache == nullptr || a->kt > 1, ORVB_EINVAL, "orvb_conv_cl: a cache needs a temporal kernel extent > 1");
  const long long pixels = static_cast<long long>(a->frames) * a->height * a->width;
  ORVB_REQUIRE(pixels < (1ll << 30), ORVB_ESHAPE, "orvb_conv_cl: too many pixels");
  const int K = a->kt * a->kh * a->kw * a->c_in;
  // 3 x 3 spatial taps and a narrow output: the halo variant (one activation tile per horizontal tap serves the three
  // vertical taps; see CH_* above).  ORVB_CONV_HALO=0 keeps the plain kernel (A/B measurements).
  static int halo_env = -1;
  if (halo_env < 0) {
    const char* e = getenv("ORVB_CONV_HALO");
    halo_env = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  const bool halo = halo_env && a->kh == 3 && a->kw == 3 && a->c_out <= 128;
  const int ph = halo ? CH_PH : CV_PH, pw = halo ? CH_PW : CV_PW;
  ConvDev cv;
  cv.T = a->frames; cv.H = a->height; cv.W = a->width;
  cv.HB = (a->height + ph - 1) / ph; cv.WB = (a->width + pw - 1) / pw;
  cv.KT = a->kt; cv.KH = a->kh; cv.KW = a->kw; cv.kc = a->c_in / 64;
  cv.num_patches = cv.T * cv.HB * cv.WB;
  cv.has_cache = a->cache != nullptr ? 1 : 0;
  const int epi = a->resid != nullptr ? ORVB_EPI_GATE_RESID : ORVB_EPI_BIAS;
  const int num_m_tiles = (cv.num_patches + 1) / 2;
  int bn = gemm_pick_bn_pair(num_m_tiles * 256, a->c_out, epi);
  if (halo && bn > 128) bn = 128;
  CUtensorMap tx, tc, tb, to;
  // activation boxes: the patch itself, or (halo) the patch's 8 columns with one row of halo above and below
  const int bw = halo ? CH_PW : CV_PW, bh = halo ? CH_PH + 2 : CV_PH;
  rc = make_tmap_4d_bf16(&tx, a->x, a->c_in, a->width, a->height, a->frames, bw, bh);
  if (rc != ORVB_OK) return rc;
  if (cv.has_cache) {
    rc = make_tmap_4d_bf16(&tc, a->cache, a->c_in, a->width, a->height, a->kt - 1, bw, bh);
    if (rc != ORVB_OK) return rc;
  } else {
    tc = tx;
  }
  rc = make_tmap_2d_bf16(&tb, a->w, a->c_out, K, K, bn / 2, BK);
  if (rc != ORVB_OK) return rc;
  const bool tma_store = !a->out_f32;
  if (tma_store) {
    // one epilogue warp stores 32 pixels: 2 rows of 16 (plain) or 4 rows of 8 (halo)
    rc = make_tmap_4d_bf16(&to, a->out, a->c_out, a->width, a->height, a->frames, pw, halo ? 4 : 2);
    if (rc != ORVB_OK) return rc;
  } else {
    to = tx;
  }
  GemmDev d;
  memset(&d, 0, sizeof(d));
  d.M = static_cast<int>(pixels); d.N = a->c_out; d.K = K;
  d.out = static_cast<bf16*>(a->out); d.ldo = a->c_out;
  d.bias = static_cast<const bf16*>(a->bias);
  d.resid = static_cast<const bf16*>(a->resid); d.ldr = a->c_out;
  d.num_m_tiles = num_m_tiles;
  d.num_n_tiles = (a->c_out + bn - 1) / bn;
  d.tma_store = tma_store ? 1 : 0;
  d.out_f32 = a->out_f32 ? 1 : 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cv.gn_gs_log2 = -1;
  cv.gn_part = nullptr;
  if (a->gn_stats != nullptr) {
    const int gs = a->gn_groups > 0 ? a->c_out / a->gn_groups : 0;
    ORVB_REQUIRE(a->gn_groups > 0 && a->gn_groups <= 32 && gs * a->gn_groups == a->c_out && gs >= 2 && gs <= 64 &&
                     (gs & (gs - 1)) == 0 && a->c_out % 64 == 0,
                 ORVB_ESHAPE, "orvb_conv_cl: fused GroupNorm statistics need c_out %% 64 == 0, <= 32 groups of 2..64 "
                 "(power of two) channels; got c_out %d, %d groups", a->c_out, a->gn_groups);
    ORVB_REQUIRE(a->gn_scratch != nullptr && reinterpret_cast<uintptr_t>(a->gn_scratch) % 16 == 0, ORVB_EINVAL,
                 "orvb_conv_cl: gn_scratch (orvb_conv_gn_scratch_bytes(), 16-byte aligned) is required with gn_stats");
    ORVB_REQUIRE(!a->out_f32, ORVB_EINVAL, "orvb_conv_cl: gn_stats describes the bf16 output (not available with out_f32)");
    int l2 = 0;
    while ((1 << l2) < gs) ++l2;
    cv.gn_gs_log2 = l2;
    cv.gn_part = static_cast<double2*>(a->gn_scratch);
  }
  int grid = 0;
  if (halo)
    rc = (epi == ORVB_EPI_GATE_RESID) ? launch_conv<ORVB_EPI_GATE_RESID, true>(tx, tc, tb, to, d, cv, bn, st, &grid)
                                      : launch_conv<ORVB_EPI_BIAS, true>(tx, tc, tb, to, d, cv, bn, st, &grid);
  else
    rc = (epi == ORVB_EPI_GATE_RESID) ? launch_conv<ORVB_EPI_GATE_RESID, false>(tx, tc, tb, to, d, cv, bn, st, &grid)
                                      : launch_conv<ORVB_EPI_BIAS, false>(tx, tc, tb, to, d, cv, bn, st, &grid);
  if (rc != ORVB_OK) return rc;
  if (a->gn_stats != nullptr) {
    const double count = static_cast<double>(pixels) * (a->c_out / a->gn_groups);
    ORVB_CHECK_CUDA(launch_kernel(conv_gn_finalize_kernel, dim3(1), dim3(1024), 0, st, true,
                                  static_cast<const double2*>(a->gn_scratch), grid, static_cast<int>(a->gn_groups), count,
                                  a->gn_eps, a->gn_stats));
  }
  return ORVB_OK;
}

extern "C" size_t orvb_conv_gn_scratch_bytes(void) {
  return static_cast<size_t>(orvb::sm_count() > 0 ? orvb::sm_count() : 160) * 32 * sizeof(double2);
}
