// Memory-bound operators of the 3-D VAE decoder (diffusers AutoencoderKLCogVideoX.decode as called by the
// reference at orv/models/cogvideox_control.py:1095-1100) over channels-last bf16 activations [T, H, W, C]:
//   gn_stats         GroupNorm statistics over all pixels of a sample (two-level deterministic reduction, one launch)
//   spatial_norm     CogVideoXSpatialNorm3D + SiLU: GroupNorm(x) * conv_y(zq) + conv_b(zq) with the two 1x1x1
//                    convolutions of the latent read from a per-latent-pixel table through the nearest-neighbour map
//   upsample2x       nearest x2 in H, W with a temporal source map (CogVideoXUpsample3D in front of its convolution)
//   cl_to_planar     [T, H, W, C] -> [C, T, H, W] for the decoded frames
// All are HBM-bound: 128-bit loads / stores, one pass over the tensor each.  The contractions are in conv.cu.
#include "common.cuh"
#include "ptx.cuh"

namespace orvb {

constexpr int VAE_THREADS = 256;

typedef float2 GnPartial;  // (sum, sum of squares) of one chunk and group

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x);
  f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z);
  f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}

// One block per chunk of `chunk` pixels.  Thread layout: vp = C / 8 threads cover one pixel (16 bytes each), so a
// block covers `lanes` = 256 / vp pixels per iteration; per-channel fp32 sums over <= chunk / lanes pixels.
__global__ void __launch_bounds__(VAE_THREADS)
gn_stats_kernel(const uint4* __restrict__ x, long long pixels, int C, int groups, int chunk, int nchunks, float eps,
                GnPartial* __restrict__ partial, unsigned int* __restrict__ counter, float* __restrict__ stats) {
  __shared__ float s_sum[VAE_THREADS * 8];
  __shared__ float s_sq[VAE_THREADS * 8];
  __shared__ bool s_last;
  pdl_launch_dependents();
  pdl_wait();  // x comes from the previous kernel of the chain (programmatic dependent launch)
  const int vp = C >> 3;
  const int lanes = VAE_THREADS / vp;
  const int tid = threadIdx.x;
  const int lane = tid / vp, v = tid - lane * vp;
  float a[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = q[j] = 0.f;
  if (lane < lanes) {
    const long long p0 = static_cast<long long>(blockIdx.x) * chunk;
    long long p1 = p0 + chunk;
    if (p1 > pixels) p1 = pixels;
    // four pixels per trip: the loads are issued together (one 16-byte load in flight per thread left the kernel at
    // 3.0 TB/s); the accumulation order per channel stays pixel order, so the result does not change
    long long p = p0 + lane;
    for (; p + 3LL * lanes < p1; p += 4LL * lanes) {
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = __ldg(x + (p + static_cast<long long>(k) * lanes) * vp + v);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float f[8];
        unpack8(u[k], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          a[j] += f[j];
          q[j] = fmaf(f[j], f[j], q[j]);
        }
      }
    }
    for (; p < p1; p += lanes) {
      float f[8];
      unpack8(__ldg(x + p * vp + v), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a[j] += f[j];
        q[j] = fmaf(f[j], f[j], q[j]);
      }
    }
  }
  // s_*[lane][channel]
  if (lane < lanes) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s_sum[lane * C + v * 8 + j] = a[j];
      s_sq[lane * C + v * 8 + j] = q[j];
    }
  }
  __syncthreads();
  const int gs = C / groups;
  for (int g = tid; g < groups; g += VAE_THREADS) {
    float s = 0.f, t = 0.f;
    for (int l = 0; l < lanes; ++l) {
      for (int c = 0; c < gs; ++c) {
        s += s_sum[l * C + g * gs + c];
        t += s_sq[l * C + g * gs + c];
      }
    }
    partial[static_cast<size_t>(blockIdx.x) * groups + g] = make_float2(s, t);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(counter, 1u) == static_cast<unsigned int>(nchunks - 1));
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // the last block to finish folds the chunk partials in fp64, in chunk order per lane and a fixed shuffle tree
  const int warp = tid >> 5, wl = tid & 31;
  const double n = static_cast<double>(pixels) * gs;
  for (int g = warp; g < groups; g += VAE_THREADS / 32) {
    double s = 0.0, t = 0.0;
    int c = wl;
    // eight partials requested per trip: one dependent L2 round trip per partial made this fold (19 per lane for the
    // decoder's largest tensors, four groups per warp) as long as the streaming pass itself
    for (; c + 7 * 32 < nchunks; c += 8 * 32) {
      GnPartial pp[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) pp[k] = __ldcg(partial + static_cast<size_t>(c + k * 32) * groups + g);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        s += pp[k].x;
        t += pp[k].y;
      }
    }
    for (; c < nchunks; c += 32) {
      const GnPartial pp = __ldcg(partial + static_cast<size_t>(c) * groups + g);
      s += pp.x;
      t += pp.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    if (wl == 0) {
      const double mean = s / n;
      double var = t / n - mean * mean;
      if (var < 0.0) var = 0.0;
      stats[2 * g] = static_cast<float>(mean);
      stats[2 * g + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    }
  }
  if (tid == 0) *counter = 0u;
}

static int gn_chunk(long long pixels) {
  // enough chunks to fill the machine, few enough that the serial fold stays short
  long long c = pixels / 592;
  if (c < 32) c = 32;
  if (c > 2048) c = 2048;
  return static_cast<int>(c);
}

struct SnDev {
  const uint4* x;
  void* y;
  long long pixels;
  int H, W, C, groups;
  const float* stats;
  const bf16 *gamma, *beta;
  const bf16* table;
  int table_ld, y_off, b_off;
  const int* t_src;
  int lat_h, lat_w, shift;
  int act, y_f32;
};

__global__ void __launch_bounds__(VAE_THREADS) spatial_norm_kernel(const SnDev p) {
  extern __shared__ float s_ab[];  // [2][C]: a = rstd * gamma, b = beta - mean * a
  pdl_launch_dependents();
  pdl_wait();  // stats / x come from the previous kernels of the chain
  const int C = p.C;
  const int gs = C / p.groups;
  for (int c = threadIdx.x; c < C; c += VAE_THREADS) {
    const int g = c / gs;
    const float a = p.stats[2 * g + 1] * __bfloat162float(p.gamma[c]);
    s_ab[c] = a;
    s_ab[C + c] = __bfloat162float(p.beta[c]) - p.stats[2 * g] * a;
  }
  __syncthreads();
  const int vp = C >> 3;
  const long long total = p.pixels * vp;
  const long long stride = static_cast<long long>(gridDim.x) * VAE_THREADS;
  const int hw = p.H * p.W;
  // The grid stride is a multiple of vp whenever vp divides the block size (C = 64 .. 512 and every power of two), so a
  // thread keeps ONE channel chunk for the whole loop: its eight (a, b) pairs live in registers, and the pixel index
  // advances by a constant.  Two vectors per trip: six independent 16-byte loads in flight per thread.
  const unsigned uvp = static_cast<unsigned>(vp);
  const bool fixed_chunk = (VAE_THREADS % vp) == 0;
  const unsigned first = blockIdx.x * VAE_THREADS + threadIdx.x;
  float ra[8], rb[8];
  if (fixed_chunk) {
    const int c0 = static_cast<int>(first % uvp) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ra[j] = s_ab[c0 + j];
      rb[j] = s_ab[C + c0 + j];
    }
  }
  for (long long i0 = first; i0 < total; i0 += 2 * stride) {
    long long idx[2] = {i0, i0 + stride};
    uint4 ux[2], uy[2], ub[2];
    int c0s[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const long long i = idx[k] < total ? idx[k] : i0;
      // 32-bit index arithmetic (the host checks pixels * C / 8 < 2^31): 64-bit divisions cost ~100 instructions each
      const unsigned iu = static_cast<unsigned>(i);
      const unsigned pix = iu / uvp;
      const int c0 = static_cast<int>(iu - pix * uvp) * 8;
      const unsigned t = pix / static_cast<unsigned>(hw);
      const unsigned r = pix - t * static_cast<unsigned>(hw);
      const unsigned h = r / static_cast<unsigned>(p.W), w = r - h * static_cast<unsigned>(p.W);
      const size_t src = (static_cast<size_t>(__ldg(p.t_src + t)) * p.lat_h + (h >> p.shift)) * p.lat_w + (w >> p.shift);
      const bf16* trow = p.table + src * p.table_ld;
      c0s[k] = c0;
      ux[k] = __ldg(p.x + i);
      uy[k] = __ldg(reinterpret_cast<const uint4*>(trow + p.y_off + c0));
      ub[k] = __ldg(reinterpret_cast<const uint4*>(trow + p.b_off + c0));
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (idx[k] >= total) break;
      const long long i = idx[k];
      const int c0 = c0s[k];
      float f[8], yy[8], bb[8], o[8];
      unpack8(ux[k], f);
      unpack8(uy[k], yy);
      unpack8(ub[k], bb);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float a = fixed_chunk ? ra[j] : s_ab[c0 + j];
        const float b = fixed_chunk ? rb[j] : s_ab[C + c0 + j];
        const float nrm = fmaf(f[j], a, b);
        float v = fmaf(nrm, yy[j], bb[j]);
        if (p.act == 1) v = __fdividef(v, 1.0f + __expf(-v));  // SiLU: MUFU.EX2 + MUFU.RCP (2 ulp), no division routine
        o[j] = v;
      }
      if (p.y_f32) {
        float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.y) + i * 8);
        op[0] = make_float4(o[0], o[1], o[2], o[3]);
        op[1] = make_float4(o[4], o[5], o[6], o[7]);
      } else {
        uint4 u;
        u.x = pack_bf16(o[0], o[1]); u.y = pack_bf16(o[2], o[3]);
        u.z = pack_bf16(o[4], o[5]); u.w = pack_bf16(o[6], o[7]);
        reinterpret_cast<uint4*>(p.y)[i] = u;
      }
    }
  }
}

__global__ void __launch_bounds__(VAE_THREADS)
upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int frames_out, int H, int W, int vp,
                  const int* __restrict__ t_src) {
  pdl_launch_dependents();
  pdl_wait();
  const int H2 = 2 * H, W2 = 2 * W;
  const long long total = static_cast<long long>(frames_out) * H2 * W2 * vp;
  const long long stride = static_cast<long long>(gridDim.x) * VAE_THREADS;
  for (long long i = static_cast<long long>(blockIdx.x) * VAE_THREADS + threadIdx.x; i < total; i += stride) {
    const unsigned iu = static_cast<unsigned>(i);  // (the host checks total < 2^31)
    const unsigned pix = iu / static_cast<unsigned>(vp);
    const unsigned v = iu - pix * static_cast<unsigned>(vp);
    const unsigned t = pix / static_cast<unsigned>(H2 * W2);
    const unsigned r = pix - t * static_cast<unsigned>(H2 * W2);
    const unsigned h = r / static_cast<unsigned>(W2), w = r - h * static_cast<unsigned>(W2);
    const size_t src = (static_cast<size_t>(__ldg(t_src + t)) * H + (h >> 1)) * W + (w >> 1);
    out[i] = __ldg(x + src * vp + v);
  }
}

__global__ void __launch_bounds__(VAE_THREADS)
cl_to_planar_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, long long pixels, int c_ld, int c_keep) {
  pdl_launch_dependents();
  pdl_wait();
  const long long stride = static_cast<long long>(gridDim.x) * VAE_THREADS;
  for (long long p = static_cast<long long>(blockIdx.x) * VAE_THREADS + threadIdx.x; p < pixels; p += stride) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + p * c_ld));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    for (int c = 0; c < c_keep && c < 8; ++c) {
      const uint16_t bits = static_cast<uint16_t>((c & 1) ? (w[c >> 1] >> 16) : (w[c >> 1] & 0xFFFFu));
      reinterpret_cast<uint16_t*>(out)[static_cast<size_t>(c) * pixels + p] = bits;
    }
  }
}

static int grid_for(long long items) {
  long long blocks = (items + VAE_THREADS - 1) / VAE_THREADS;
  const long long cap = static_cast<long long>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace orvb

extern "C" size_t orvb_gn_scratch_bytes(int64_t pixels, int32_t groups) {
  if (pixels <= 0 || groups <= 0) return 256;
  const int chunk = orvb::gn_chunk(pixels);
  const long long nchunks = (pixels + chunk - 1) / chunk;
  return 256 + static_cast<size_t>(nchunks) * groups * sizeof(orvb::GnPartial);
}

extern "C" int orvb_gn_stats_cl(const void* x, int64_t pixels, int32_t channels, int32_t groups, float eps, float* stats,
                                void* scratch, void* stream) {
  using namespace orvb;
  int rc = check_arch();
  if (rc != ORVB_OK) return rc;
  ORVB_REQUIRE(x && stats && scratch, ORVB_EINVAL, "orvb_gn_stats_cl: null pointer");
  ORVB_REQUIRE(pixels > 0, ORVB_ESHAPE, "orvb_gn_stats_cl: empty tensor");
  ORVB_REQUIRE(channels > 0 && channels % 8 == 0 && channels <= 2048 && groups > 0 && channels % groups == 0, ORVB_ESHAPE,
               "orvb_gn_stats_cl: channels (%d) must be a multiple of 8 (<= 2048) and of groups (%d)", channels, groups);
  const int chunk = gn_chunk(pixels);
  const int nchunks = static_cast<int>((pixels + chunk - 1) / chunk);
  ORVB_CHECK_CUDA(launch_kernel(gn_stats_kernel, dim3(nchunks), dim3(VAE_THREADS), 0, static_cast<cudaStream_t>(stream), true,
                                static_cast<const uint4*>(x), static_cast<long long>(pixels), static_cast<int>(channels),
                                static_cast<int>(groups), chunk, nchunks, eps,
                                reinterpret_cast<GnPartial*>(static_cast<uint8_t*>(scratch) + 256),
                                static_cast<unsigned int*>(scratch), stats));
  return ORVB_OK;
}

extern "C" int orvb_spatial_norm_cl(const orvb_spatial_norm_args* a, void* stream) {
  using namespace orvb;
  int rc = check_arch();
  if (rc != ORVB_OK) return rc;
  ORVB_REQUIRE(a && a->x && a->y && a->stats && a->gamma && a->beta && a->table && a->t_src, ORVB_EINVAL,
               "orvb_spatial_norm_cl: null pointer");
  ORVB_REQUIRE(a->frames > 0 && a->height > 0 && a->width > 0, ORVB_ESHAPE, "orvb_spatial_norm_cl: empty tensor");
  ORVB_REQUIRE(a->channels > 0 && a->channels % 8 == 0 && a->channels <= 2048 && a->groups > 0 &&
                   a->channels % a->groups == 0,
               ORVB_ESHAPE, "orvb_spatial_norm_cl: channels (%d) must be a multiple of 8 (<= 2048) and of groups (%d)",
               a->channels, a->groups);
  ORVB_REQUIRE(a->table_ld % 8 == 0 && a->y_off % 8 == 0 && a->b_off % 8 == 0 && a->y_off >= 0 && a->b_off >= 0 &&
                   a->y_off + a->channels <= a->table_ld && a->b_off + a->channels <= a->table_ld,
               ORVB_ESHAPE, "orvb_spatial_norm_cl: table columns must be 16-byte aligned and inside the row");
  ORVB_REQUIRE(a->shift >= 0 && a->shift < 8 && a->lat_h > 0 && a->lat_w > 0 &&
                   ((a->height - 1) >> a->shift) < a->lat_h && ((a->width - 1) >> a->shift) < a->lat_w,
               ORVB_ESHAPE, "orvb_spatial_norm_cl: %dx%d >> %d does not fit the %dx%d latent", a->height, a->width,
               a->shift, a->lat_h, a->lat_w);
  SnDev d;
  d.x = static_cast<const uint4*>(a->x); d.y = a->y;
  d.pixels = static_cast<long long>(a->frames) * a->height * a->width;
  d.H = a->height; d.W = a->width; d.C = a->channels; d.groups = a->groups;
  d.stats = a->stats;
  d.gamma = static_cast<const bf16*>(a->gamma); d.beta = static_cast<const bf16*>(a->beta);
  d.table = static_cast<const bf16*>(a->table);
  d.table_ld = a->table_ld; d.y_off = a->y_off; d.b_off = a->b_off;
  d.t_src = a->t_src; d.lat_h = a->lat_h; d.lat_w = a->lat_w; d.shift = a->shift;
  d.act = a->act; d.y_f32 = a->y_f32;
  const long long items = d.pixels * (d.C >> 3);
  ORVB_REQUIRE(items < (1ll << 31), ORVB_ESHAPE, "orvb_spatial_norm_cl: tensor too large (%lld 16-byte vectors)", items);
  ORVB_CHECK_CUDA(launch_kernel(spatial_norm_kernel, dim3(grid_for(items)), dim3(VAE_THREADS), 2 * d.C * sizeof(float),
                                static_cast<cudaStream_t>(stream), true, d));
  return ORVB_OK;
}

extern "C" int orvb_upsample2x_cl(const void* x, void* out, int32_t frames_out, int32_t height, int32_t width,
                                  int32_t channels, const int32_t* t_src, void* stream) {
  using namespace orvb;
  int rc = check_arch();
  if (rc != ORVB_OK) return rc;
  ORVB_REQUIRE(x && out && t_src, ORVB_EINVAL, "orvb_upsample2x_cl: null pointer");
  ORVB_REQUIRE(frames_out > 0 && height > 0 && width > 0 && channels > 0 && channels % 8 == 0, ORVB_ESHAPE,
               "orvb_upsample2x_cl: bad shape (channels must be a multiple of 8)");
  const long long items = static_cast<long long>(frames_out) * 4 * height * width * (channels >> 3);
  ORVB_REQUIRE(items < (1ll << 31), ORVB_ESHAPE, "orvb_upsample2x_cl: tensor too large (%lld 16-byte vectors)", items);
  ORVB_CHECK_CUDA(launch_kernel(upsample2x_kernel, dim3(grid_for(items)), dim3(VAE_THREADS), 0,
                                static_cast<cudaStream_t>(stream), true, static_cast<const uint4*>(x),
                                static_cast<uint4*>(out), static_cast<int>(frames_out), static_cast<int>(height),
                                static_cast<int>(width), static_cast<int>(channels >> 3), static_cast<const int*>(t_src)));
  return ORVB_OK;
}

extern "C" int orvb_cl_to_planar(const void* x, void* out, int64_t pixels, int32_t c_ld, int32_t c_keep, void* stream) {
  using namespace orvb;
  int rc = check_arch();
  if (rc != ORVB_OK) return rc;
  ORVB_REQUIRE(x && out, ORVB_EINVAL, "orvb_cl_to_planar: null pointer");
  ORVB_REQUIRE(pixels > 0 && c_ld >= 8 && c_ld % 8 == 0 && c_keep >= 1 && c_keep <= 8, ORVB_ESHAPE,
               "orvb_cl_to_planar: c_ld must be a multiple of 8 and 1 <= c_keep <= 8");
  ORVB_CHECK_CUDA(launch_kernel(cl_to_planar_kernel, dim3(grid_for(pixels)), dim3(VAE_THREADS), 0,
                                static_cast<cudaStream_t>(stream), true, static_cast<const bf16*>(x), static_cast<bf16*>(out),
                                static_cast<long long>(pixels), static_cast<int>(c_ld), static_cast<int>(c_keep)));
  return ORVB_OK;
}
