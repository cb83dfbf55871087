// 3-D Gaussian rasteriser of the occupancy chain (second half): the B200 replacement of the forward
// pass of the reference's `orv/ops/diff-gaussian-rasterization` extension (Inria's rasteriser extended by a 12-channel
// semantic feature, a depth and an alpha output), which `orv/dataset/gs_render.py:103-171` uses to turn occupancy
// voxels into depth / semantic maps.  Forward only — the reference's backward is a training facility and out of scope.
//
// Floating-point, HBM / latency bound; nothing here is GEMM-shaped.  Same algorithm as the reference
// (cuda_rasterizer/forward.cu:156-262 preprocess, :267-398 render; rasterizer_impl.cu:54-318 binning), re-designed:
//
//   * no host round trip: the reference copies the instance count to the host to size its buffers and sorts 64-bit
//     (tile | depth) keys of every (Gaussian, tile) instance with cub.  Here the P Gaussians are depth-sorted ONCE
//     (32-bit keys, stable LSD radix, values = Gaussian index), instances are emitted in that order and then only the
//     tile id (10 bits for a 320x480 frame) is sorted, stably: ONE 10-bit pass over the instances instead of a 64-bit
//     sort.  The result is the reference's order — by tile, then depth, ties by Gaussian index.  Buffers have a
//     caller-given capacity; every kernel reads the real count from device memory, so the whole call is one
//     stream-ordered launch sequence (capturable).
//   * the render kernel stages, per batch of 256 instances, not only position / conic / opacity but also the 16
//     blended channels (RGB, depth, 12 features) in shared memory: the reference fetches them from global memory inside
//     the per-pixel loop (16 loads per Gaussian and pixel).
//
// The arithmetic of every formula follows the reference expression by expression (same operand order, expf, no
// fast-math), so results agree to fp32 rounding; they are not bit-identical because glm's matrix products and nvcc's FMA
// contraction are not reproduced instruction by instruction (tests: tolerance written there).
#include "common.cuh"
#include "sortscan.cuh"

#include <math.h>

namespace orvb {
namespace {

constexpr int GS_BLOCK_X = 16, GS_BLOCK_Y = 16;  // config.h:16-17
constexpr int GS_BLOCK = GS_BLOCK_X * GS_BLOCK_Y;
constexpr int GS_COLOR = 3, GS_FEAT = 12;        // config.h:14-15
constexpr int GS_PAYLOAD = 16;                   // RGB, depth, 12 features

struct GsParams {
  int P, W, H, gx, gy;
  float tan_fovx, tan_fovy, focal_x, focal_y, scale_mod;
  const float* view;
  const float* proj;
};

// auxiliary.h:47-57
__device__ __forceinline__ void gs_rect(float px, float py, int radius, int gx, int gy, int& x0, int& y0, int& x1,
                                        int& y1) {
  x0 = min(gx, max(0, static_cast<int>((px - radius) / GS_BLOCK_X)));
  y0 = min(gy, max(0, static_cast<int>((py - radius) / GS_BLOCK_Y)));
  x1 = min(gx, max(0, static_cast<int>((px + radius + GS_BLOCK_X - 1) / GS_BLOCK_X)));
  y1 = min(gy, max(0, static_cast<int>((py + radius + GS_BLOCK_Y - 1) / GS_BLOCK_Y)));
}

// auxiliary.h:42-45 (the arithmetic is done in double there: 1.0 and 0.5 are double literals)
__device__ __forceinline__ float gs_ndc2pix(float v, int S) {
  return static_cast<float>(((static_cast<double>(v) + 1.0) * S - 1.0) * 0.5);
}

// ---- 1. per-Gaussian preprocessing (forward.cu:156-262) ---------------------------------------------------------------
__global__ void __launch_bounds__(256) gs_preprocess_kernel(GsParams g, const float* __restrict__ means,
                                                            const float* __restrict__ scales,
                                                            const float* __restrict__ rots,
                                                            const float* __restrict__ cov_pre,
                                                            const float* __restrict__ opac, int32_t* __restrict__ radii,
                                                            float2* __restrict__ xy, float* __restrict__ depth,
                                                            float4* __restrict__ conic_op, uint32_t* __restrict__ tiles,
                                                            uint32_t* __restrict__ dkey) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= g.P) return;
  radii[idx] = 0;
  tiles[idx] = 0;
  dkey[idx] = 0xFFFFFFFFu;  // invisible Gaussians sort last and emit no instance
  const float* V = g.view;
  const float* M = g.proj;
  const float px = means[3 * idx], py = means[3 * idx + 1], pz = means[3 * idx + 2];
  // in_frustum (auxiliary.h:139-161): near culling on the view-space depth only
  const float hx = M[0] * px + M[4] * py + M[8] * pz + M[12];
  const float hy = M[1] * px + M[5] * py + M[9] * pz + M[13];
  const float hw = M[3] * px + M[7] * py + M[11] * pz + M[15];
  const float p_w = 1.0f / (hw + 0.0000001f);
  const float projx = hx * p_w, projy = hy * p_w;
  float tx = V[0] * px + V[4] * py + V[8] * pz + V[12];
  float ty = V[1] * px + V[5] * py + V[9] * pz + V[13];
  const float tz = V[2] * px + V[6] * py + V[10] * pz + V[14];
  if (tz <= 0.01f) return;

  // 3-D covariance (computeCov3D, forward.cu:118-154): Sigma = (S R)^T (S R); the quaternion is used as given
  float c3[6];
  if (cov_pre != nullptr) {
#pragma unroll
    for (int i = 0; i < 6; ++i) c3[i] = cov_pre[6 * idx + i];
  } else {
    const float sx = g.scale_mod * scales[3 * idx], sy = g.scale_mod * scales[3 * idx + 1],
                sz = g.scale_mod * scales[3 * idx + 2];
    const float r = rots[4 * idx], x = rots[4 * idx + 1], y = rots[4 * idx + 2], z = rots[4 * idx + 3];
    // glm::mat3 R(...) is filled column by column, so these nine literals are R's COLUMNS; M = S * R scales row i by s_i
    const float R00 = 1.f - 2.f * (y * y + z * z), R01 = 2.f * (x * y - r * z), R02 = 2.f * (x * z + r * y);
    const float R10 = 2.f * (x * y + r * z), R11 = 1.f - 2.f * (x * x + z * z), R12 = 2.f * (y * z - r * x);
    const float R20 = 2.f * (x * z - r * y), R21 = 2.f * (y * z + r * x), R22 = 1.f - 2.f * (x * x + y * y);
    // column c of M (glm M[c]) = (sx R[c][0], sy R[c][1], sz R[c][2]); Sigma = M^T M, Sigma[i][j] = dot(row_i(M^T) ...)
    // glm: Sigma[c][r] = sum_k (M^T)[k][r] * M[c][k] = sum_k M[r][k] * M[c][k]  = dot(column r of M, column c of M)
    const float m0[3] = {sx * R00, sy * R01, sz * R02};
    const float m1[3] = {sx * R10, sy * R11, sz * R12};
    const float m2[3] = {sx * R20, sy * R21, sz * R22};
    c3[0] = m0[0] * m0[0] + m0[1] * m0[1] + m0[2] * m0[2];
    c3[1] = m0[0] * m1[0] + m0[1] * m1[1] + m0[2] * m1[2];
    c3[2] = m0[0] * m2[0] + m0[1] * m2[1] + m0[2] * m2[2];
    c3[3] = m1[0] * m1[0] + m1[1] * m1[1] + m1[2] * m1[2];
    c3[4] = m1[0] * m2[0] + m1[1] * m2[1] + m1[2] * m2[2];
    c3[5] = m2[0] * m2[0] + m2[1] * m2[1] + m2[2] * m2[2];
  }

  // 2-D covariance (computeCov2D, forward.cu:74-116; EWA splatting eqs. 29 / 31)
  const float limx = 1.3f * g.tan_fovx, limy = 1.3f * g.tan_fovy;
  const float txtz = tx / tz, tytz = ty / tz;
  tx = fminf(limx, fmaxf(-limx, txtz)) * tz;
  ty = fminf(limy, fmaxf(-limy, tytz)) * tz;
  // J (math rows): [fx/tz, 0, -fx tx / tz^2; 0, fy/tz, -fy ty / tz^2; 0 0 0];  Wr = rotation part of the view matrix
  const float j00 = g.focal_x / tz, j02 = -(g.focal_x * tx) / (tz * tz);
  const float j11 = g.focal_y / tz, j12 = -(g.focal_y * ty) / (tz * tz);
  // rows of A = J * Wr (2 x 3), Wr[r][c] = V[r + 4 c]
  const float a00 = j00 * V[0] + j02 * V[2], a01 = j00 * V[4] + j02 * V[6], a02 = j00 * V[8] + j02 * V[10];
  const float a10 = j11 * V[1] + j12 * V[2], a11 = j11 * V[5] + j12 * V[6], a12 = j11 * V[9] + j12 * V[10];
  // cov2d = A Sigma A^T
  const float b00 = a00 * c3[0] + a01 * c3[1] + a02 * c3[2];
  const float b01 = a00 * c3[1] + a01 * c3[3] + a02 * c3[4];
  const float b02 = a00 * c3[2] + a01 * c3[4] + a02 * c3[5];
  const float b10 = a10 * c3[0] + a11 * c3[1] + a12 * c3[2];
  const float b11 = a10 * c3[1] + a11 * c3[3] + a12 * c3[4];
  const float b12 = a10 * c3[2] + a11 * c3[4] + a12 * c3[5];
  const float cxx = b00 * a00 + b01 * a01 + b02 * a02 + 0.3f;  // low-pass: at least one pixel wide
  const float cxy = b00 * a10 + b01 * a11 + b02 * a12;
  const float cyy = b10 * a10 + b11 * a11 + b12 * a12 + 0.3f;

  const float det = cxx * cyy - cxy * cxy;
  if (det == 0.0f) return;
  const float det_inv = 1.f / det;
  const float mid = 0.5f * (cxx + cyy);
  const float lambda1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
  const float lambda2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
  const float my_radius = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));
  const float ix = gs_ndc2pix(projx, g.W), iy = gs_ndc2pix(projy, g.H);
  int x0, y0, x1, y1;
  gs_rect(ix, iy, static_cast<int>(my_radius), g.gx, g.gy, x0, y0, x1, y1);
  if ((x1 - x0) * (y1 - y0) == 0) return;

  depth[idx] = tz;
  radii[idx] = static_cast<int>(my_radius);
  xy[idx] = make_float2(ix, iy);
  conic_op[idx] = make_float4(cyy * det_inv, -cxy * det_inv, cxx * det_inv, opac[idx]);
  tiles[idx] = static_cast<uint32_t>((y1 - y0) * (x1 - x0));
  dkey[idx] = __float_as_uint(tz);  // tz > 0.01: the bit pattern of a positive float orders like the value
}

// tiles touched, read in depth order
struct LoadTilesSorted {
  const uint32_t* tiles;
  const uint32_t* order;
  __device__ uint32_t operator()(int64_t i) const { return tiles[order[i]]; }
};

// ---- 2. one (tile, Gaussian) instance per overlapped tile, emitted in depth order (rasterizer_impl.cu:74-115) --------
__global__ void __launch_bounds__(256) gs_expand_kernel(int P, int gx, int gy, const uint32_t* __restrict__ order,
                                                        const uint32_t* __restrict__ off,
                                                        const uint32_t* __restrict__ tiles,
                                                        const float2* __restrict__ xy, const int32_t* __restrict__ radii,
                                                        uint32_t cap, uint32_t* __restrict__ tkey,
                                                        uint32_t* __restrict__ tval) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const uint32_t gidx = order[i];
  const uint32_t n = tiles[gidx];
  if (n == 0) return;
  uint32_t o = off[i];
  int x0, y0, x1, y1;
  const float2 p = xy[gidx];
  gs_rect(p.x, p.y, radii[gidx], gx, gy, x0, y0, x1, y1);
  // instances past the capacity are dropped (every slot below it stays valid); the caller sees total > capacity in
  // num_rendered and retries with more room
  for (int y = y0; y < y1; ++y)
    for (int x = x0; x < x1 && o < cap; ++x) {
      tkey[o] = static_cast<uint32_t>(y * gx + x);
      tval[o] = gidx;
      ++o;
    }
}

// ---- 3. per-tile ranges of the tile-sorted instance list (rasterizer_impl.cu:117-141) --------------------------------
__global__ void __launch_bounds__(256) gs_ranges_kernel(const uint32_t* __restrict__ tkey, uint32_t cap,
                                                        const uint32_t* __restrict__ total, uint2* __restrict__ ranges,
                                                        int32_t* __restrict__ num_rendered) {
  const uint32_t L = min(*total, cap);
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && num_rendered != nullptr)
    *num_rendered = (*total > cap) ? -static_cast<int32_t>(min(*total, 0x7FFFFFFFu)) : static_cast<int32_t>(*total);
  if (i >= L) return;
  const uint32_t cur = tkey[i];
  if (i == 0) {
    ranges[cur].x = 0;
  } else {
    const uint32_t prev = tkey[i - 1];
    if (cur != prev) {
      ranges[prev].y = i;
      ranges[cur].x = i;
    }
  }
  if (i == L - 1) ranges[cur].y = L;
}

// ---- 4. tile renderer (forward.cu:267-398) ----------------------------------------------------------------------------
template <bool FEAT>
__global__ void __launch_bounds__(GS_BLOCK) gs_render_kernel(const uint2* __restrict__ ranges,
                                                             const uint32_t* __restrict__ list, int W, int H,
                                                             const float2* __restrict__ xy,
                                                             const float* __restrict__ colors,
                                                             const float* __restrict__ feats,
                                                             const float* __restrict__ depth,
                                                             const float4* __restrict__ conic_op,
                                                             const float* __restrict__ bg, float* __restrict__ out_color,
                                                             float* __restrict__ out_feat, float* __restrict__ out_depth,
                                                             float* __restrict__ out_alpha) {
  __shared__ float2 s_xy[GS_BLOCK];
  __shared__ float4 s_co[GS_BLOCK];
  __shared__ float s_pay[GS_BLOCK][FEAT ? GS_PAYLOAD : 4];  // RGB, depth (, 12 features) of the batch
  const int tid = threadIdx.y * GS_BLOCK_X + threadIdx.x;
  const int hb = (W + GS_BLOCK_X - 1) / GS_BLOCK_X;
  const int pxi = blockIdx.x * GS_BLOCK_X + threadIdx.x, pyi = blockIdx.y * GS_BLOCK_Y + threadIdx.y;
  const bool inside = pxi < W && pyi < H;
  const int pix_id = W * pyi + pxi;
  const float pxf = static_cast<float>(pxi), pyf = static_cast<float>(pyi);
  bool done = !inside;
  const uint2 range = ranges[blockIdx.y * hb + blockIdx.x];
  const int rounds = (static_cast<int>(range.y - range.x) + GS_BLOCK - 1) / GS_BLOCK;
  int todo = static_cast<int>(range.y - range.x);

  float T = 1.0f;
  float C[GS_COLOR] = {0.f, 0.f, 0.f};
  float F[GS_FEAT] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float D = 0.f;
  for (int i = 0; i < rounds; ++i, todo -= GS_BLOCK) {
    if (__syncthreads_count(done) == GS_BLOCK) break;  // (also orders the previous batch's reads before the refill)
    const int progress = i * GS_BLOCK + tid;
    if (static_cast<int>(range.x) + progress < static_cast<int>(range.y)) {
      const uint32_t id = list[range.x + progress];
      s_xy[tid] = xy[id];
      s_co[tid] = conic_op[id];
      s_pay[tid][0] = colors[id * GS_COLOR + 0];
      s_pay[tid][1] = colors[id * GS_COLOR + 1];
      s_pay[tid][2] = colors[id * GS_COLOR + 2];
      s_pay[tid][3] = depth[id];
      if (FEAT) {
        const float4* f4 = reinterpret_cast<const float4*>(feats + static_cast<size_t>(id) * GS_FEAT);
        const float4 f0 = f4[0], f1 = f4[1], f2 = f4[2];
        s_pay[tid][4] = f0.x; s_pay[tid][5] = f0.y; s_pay[tid][6] = f0.z; s_pay[tid][7] = f0.w;
        s_pay[tid][8] = f1.x; s_pay[tid][9] = f1.y; s_pay[tid][10] = f1.z; s_pay[tid][11] = f1.w;
        s_pay[tid][12] = f2.x; s_pay[tid][13] = f2.y; s_pay[tid][14] = f2.z; s_pay[tid][15] = f2.w;
      }
    }
    __syncthreads();
    const int nb = min(GS_BLOCK, todo);
    for (int j = 0; !done && j < nb; ++j) {
      const float2 p = s_xy[j];
      const float dx = p.x - pxf, dy = p.y - pyf;
      const float4 co = s_co[j];
      const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
      if (power > 0.0f) continue;
      const float alpha = fminf(0.99f, co.w * expf(power));
      if (alpha < 1.0f / 255.0f) continue;
      const float test_T = T * (1 - alpha);
      if (test_T < 0.0001f) {
        done = true;
        continue;
      }
#pragma unroll
      for (int ch = 0; ch < GS_COLOR; ++ch) C[ch] += s_pay[j][ch] * alpha * T;
      D += s_pay[j][3] * alpha * T;
      if (FEAT) {
#pragma unroll
        for (int ch = 0; ch < GS_FEAT; ++ch) F[ch] += s_pay[j][4 + ch] * alpha * T;
      }
      T = test_T;
    }
  }
  if (inside) {
    const size_t HW = static_cast<size_t>(H) * W;
#pragma unroll
    for (int ch = 0; ch < GS_COLOR; ++ch) out_color[ch * HW + pix_id] = C[ch] + T * bg[ch];
    out_alpha[pix_id] = 1 - T;
    out_depth[pix_id] = D;
    if (FEAT) {
#pragma unroll
      for (int ch = 0; ch < GS_FEAT; ++ch) out_feat[ch * HW + pix_id] = F[ch];
    }
  }
}

inline int gs_depth_items(int P) { return P < (1 << 20) ? 2 : 8; }
inline int gs_tile_digit_bits(size_t tiles) { return tiles <= 1024 ? 10 : 9; }

struct GsWorkspace {
  size_t off_xy, off_depth, off_conic, off_tiles, off_dka, off_dva, off_dkb, off_dvb, off_off, off_total, off_hist,
      off_scan, off_tka, off_tva, off_tkb, off_tvb, off_ranges, bytes;
};

GsWorkspace gs_plan(int P, int cap, int H, int W) {
  GsWorkspace w;
  size_t off = 0;
  auto take = [&](size_t n) {
    const size_t o = off;
    off += align256(n);
    return o;
  };
  const size_t p = static_cast<size_t>(P > 0 ? P : 1), c = static_cast<size_t>(cap > 0 ? cap : 1);
  const size_t tiles = static_cast<size_t>((W + GS_BLOCK_X - 1) / GS_BLOCK_X) * ((H + GS_BLOCK_Y - 1) / GS_BLOCK_Y);
  w.off_xy = take(p * 8);
  w.off_depth = take(p * 4);
  w.off_conic = take(p * 16);
  w.off_tiles = take(p * 4);
  w.off_dka = take(p * 4);
  w.off_dva = take(p * 4);
  w.off_dkb = take(p * 4);
  w.off_dvb = take(p * 4);
  w.off_off = take(p * 4);
  w.off_total = take(256);
  // depth sort: 8-bit digits, small tiles below a million Gaussians; tile sort: one 10-bit pass when the frame has at
  // most 1024 tiles (two 9-bit passes otherwise), 2048-element tiles
  size_t he = radix_hist_entries(static_cast<int64_t>(p), 8, gs_depth_items(P));
  const size_t he_t = radix_hist_entries(static_cast<int64_t>(c), gs_tile_digit_bits(tiles), 8);
  if (he_t > he) he = he_t;
  w.off_hist = take(he * 4);
  w.off_scan = take(scan_scratch_entries(static_cast<int64_t>(he > p ? he : p)) * 4);
  w.off_tka = take(c * 4);
  w.off_tva = take(c * 4);
  w.off_tkb = take(c * 4);
  w.off_tvb = take(c * 4);
  w.off_ranges = take(tiles * 8);
  w.bytes = off;
  return w;
}

}  // namespace
}  // namespace orvb

extern "C" size_t orvb_gs_workspace_bytes(int32_t p, int32_t max_instances, int32_t height, int32_t width) {
  if (p < 0 || max_instances < 0 || height <= 0 || width <= 0) return 0;
  return orvb::gs_plan(p, max_instances, height, width).bytes;
}

extern "C" int orvb_gs_rasterize(const orvb_gs_args* a, void* stream) {
  using namespace orvb;
  int rc = check_arch();
  if (rc != ORVB_OK) return rc;
  ORVB_REQUIRE(a != nullptr, ORVB_EINVAL, "gs_rasterize: null args");
  ORVB_REQUIRE(a->height > 0 && a->width > 0 && a->p >= 0 && a->p < (1 << 28), ORVB_ESHAPE,
               "gs_rasterize: bad sizes p=%d height=%d width=%d", a->p, a->height, a->width);
  ORVB_REQUIRE(a->out_color && a->out_depth && a->out_alpha && a->background && a->viewmatrix && a->projmatrix,
               ORVB_EINVAL, "gs_rasterize: null output / camera pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int W = a->width, H = a->height, P = a->p;
  const int gx = (W + GS_BLOCK_X - 1) / GS_BLOCK_X, gy = (H + GS_BLOCK_Y - 1) / GS_BLOCK_Y;
  ORVB_REQUIRE(static_cast<long long>(gx) * gy < (1ll << 27), ORVB_ESHAPE, "gs_rasterize: image too large");
  const int cap = a->max_instances;
  ORVB_REQUIRE(cap > 0 && a->workspace != nullptr, ORVB_EINVAL, "gs_rasterize: workspace / max_instances missing");
  ORVB_REQUIRE(reinterpret_cast<uintptr_t>(a->workspace) % 256 == 0, ORVB_ESHAPE, "gs_rasterize: workspace must be 256-byte aligned");
  const GsWorkspace w = gs_plan(P, cap, H, W);
  ORVB_REQUIRE(a->workspace_bytes >= w.bytes, ORVB_ENOMEM, "gs_rasterize: workspace %zu < %zu bytes", a->workspace_bytes, w.bytes);
  ORVB_REQUIRE(P == 0 || (a->means3d && a->colors && a->opacities && a->radii), ORVB_EINVAL, "gs_rasterize: null Gaussian tensor");
  ORVB_REQUIRE(P == 0 || a->cov3d != nullptr || (a->scales != nullptr && a->rotations != nullptr), ORVB_EINVAL,
               "gs_rasterize: provide either scales + rotations or a precomputed 3-D covariance");
  const bool feat = a->features != nullptr && a->out_feature != nullptr;
  ORVB_REQUIRE(!feat || reinterpret_cast<uintptr_t>(a->features) % 16 == 0, ORVB_ESHAPE, "gs_rasterize: features must be 16-byte aligned");
  char* base = static_cast<char*>(a->workspace);
  float2* xy = reinterpret_cast<float2*>(base + w.off_xy);
  float* depth = reinterpret_cast<float*>(base + w.off_depth);
  float4* conic = reinterpret_cast<float4*>(base + w.off_conic);
  uint32_t* tiles = reinterpret_cast<uint32_t*>(base + w.off_tiles);
  uint32_t* dka = reinterpret_cast<uint32_t*>(base + w.off_dka);
  uint32_t* dva = reinterpret_cast<uint32_t*>(base + w.off_dva);
  uint32_t* dkb = reinterpret_cast<uint32_t*>(base + w.off_dkb);
  uint32_t* dvb = reinterpret_cast<uint32_t*>(base + w.off_dvb);
  uint32_t* off = reinterpret_cast<uint32_t*>(base + w.off_off);
  uint32_t* total = reinterpret_cast<uint32_t*>(base + w.off_total);
  uint32_t* hist = reinterpret_cast<uint32_t*>(base + w.off_hist);
  uint32_t* scan = reinterpret_cast<uint32_t*>(base + w.off_scan);
  uint32_t* tka = reinterpret_cast<uint32_t*>(base + w.off_tka);
  uint32_t* tva = reinterpret_cast<uint32_t*>(base + w.off_tva);
  uint32_t* tkb = reinterpret_cast<uint32_t*>(base + w.off_tkb);
  uint32_t* tvb = reinterpret_cast<uint32_t*>(base + w.off_tvb);
  uint2* ranges = reinterpret_cast<uint2*>(base + w.off_ranges);

  ORVB_CHECK_CUDA(cudaMemsetAsync(ranges, 0, static_cast<size_t>(gx) * gy * sizeof(uint2), st));
  ORVB_CHECK_CUDA(cudaMemsetAsync(total, 0, 4, st));
  const uint32_t* list = tva;
  if (P > 0) {
    GsParams g;
    g.P = P; g.W = W; g.H = H; g.gx = gx; g.gy = gy;
    g.tan_fovx = a->tan_fovx; g.tan_fovy = a->tan_fovy;
    g.focal_y = H / (2.0f * a->tan_fovy);  // rasterizer_impl.cu:229-230
    g.focal_x = W / (2.0f * a->tan_fovx);
    g.scale_mod = a->scale_modifier;
    g.view = a->viewmatrix; g.proj = a->projmatrix;
    const unsigned pb = static_cast<unsigned>((P + 255) / 256);
    gs_preprocess_kernel<<<pb, 256, 0, st>>>(g, a->means3d, a->scales, a->rotations, a->cov3d, a->opacities, a->radii, xy,
                                             depth, conic, tiles, dka);
    ORVB_CHECK_CUDA(cudaGetLastError());
    // depth order of the Gaussians (stable: ties keep index order)
    const uint32_t *dk, *order;
    rc = radix_sort_pairs(dka, dva, dkb, dvb, true, P, 32, 8, gs_depth_items(P), hist, scan, &dk, &order, st);
    if (rc != ORVB_OK) return rc;
    // instance offsets in depth order, total instance count
    rc = scan_any(LoadTilesSorted{tiles, order}, off, P, scan, total, st);
    if (rc != ORVB_OK) return rc;
    gs_expand_kernel<<<pb, 256, 0, st>>>(P, gx, gy, order, off, tiles, xy, a->radii, static_cast<uint32_t>(cap), tka, tva);
    ORVB_CHECK_CUDA(cudaGetLastError());
    // stable sort by tile id only: (tile, depth, index) order, the reference's 64-bit key order
    int bits = 1;
    while ((1 << bits) < gx * gy) ++bits;
    const uint32_t *tk, *tv;
    rc = radix_sort_pairs(tka, tva, tkb, tvb, false, cap, bits, gs_tile_digit_bits(static_cast<size_t>(gx) * gy), 8, hist,
                          scan, &tk, &tv, st, total);
    if (rc != ORVB_OK) return rc;
    list = tv;
    const unsigned rb = static_cast<unsigned>((cap + 255) / 256);
    gs_ranges_kernel<<<rb, 256, 0, st>>>(tk, static_cast<uint32_t>(cap), total, ranges, a->num_rendered);
    ORVB_CHECK_CUDA(cudaGetLastError());
  } else if (a->num_rendered != nullptr) {
    ORVB_CHECK_CUDA(cudaMemsetAsync(a->num_rendered, 0, 4, st));
  }
  const dim3 grid(gx, gy), block(GS_BLOCK_X, GS_BLOCK_Y);
  if (feat)
    gs_render_kernel<true><<<grid, block, 0, st>>>(ranges, list, W, H, xy, a->colors, a->features, depth, conic,
                                                   a->background, a->out_color, a->out_feature, a->out_depth, a->out_alpha);
  else
    gs_render_kernel<false><<<grid, block, 0, st>>>(ranges, list, W, H, xy, a->colors, nullptr, depth, conic, a->background,
                                                    a->out_color, nullptr, a->out_depth, a->out_alpha);
  ORVB_CHECK_CUDA(cudaGetLastError());
  return ORVB_OK;
}
