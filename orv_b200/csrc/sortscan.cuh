// Scan and stable radix-sort building blocks shared by the integer / index kernels of the occupancy chain
// (voxelize.cu: voxel numbering and point grouping; gs_render.cu: depth order and tile binning of the Gaussians).
// Header-only on purpose: every translation unit gets its own copy inside an anonymous namespace.
#pragma once
#include "common.cuh"

namespace orvb {
namespace {

constexpr int kThreads = 256;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kItems = 8;                    // elements per thread in the scan / sort tiles
constexpr int kTile = kThreads * kItems;     // 2048
constexpr int kRadixBits = 9;                // 17-bit keys (max_voxels = 1e5, the occupancy caller) sort in two passes
constexpr int kBins = 1 << kRadixBits;       // 512

struct LoadU32 {
  const uint32_t* in;
  __device__ __forceinline__ uint32_t operator()(int64_t i) const { return in[i]; }
};

// ---- block-wide exclusive scan of one value per thread (256 threads) ------------------------------------------------
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t& total) {
  __shared__ uint32_t warp_sums[kThreads / 32 + 1];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(kFull, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const uint32_t w = (lane < kThreads / 32) ? warp_sums[lane] : 0u;
    uint32_t winc = w;
#pragma unroll
    for (int o = 1; o < kThreads / 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(kFull, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < kThreads / 32) warp_sums[lane] = winc - w;
    if (lane == kThreads / 32 - 1) warp_sums[kThreads / 32] = winc;
  }
  __syncthreads();
  const uint32_t res = warp_sums[warp] + inc - v;
  total = warp_sums[kThreads / 32];
  __syncthreads();  // warp_sums may be reused by the caller's next scan
  return res;
}

// tile sums: sums[b] = sum of in[b*kTile .. (b+1)*kTile)
template <class Load>
__global__ void __launch_bounds__(kThreads) scan_reduce_kernel(Load in, int64_t n, uint32_t* __restrict__ sums) {
  const int64_t base = static_cast<int64_t>(blockIdx.x) * kTile + static_cast<int64_t>(threadIdx.x) * kItems;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    if (base + k < n) s += in(base + k);
  }
  uint32_t total;
  block_exclusive_scan(s, total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// exclusive scan of one tile (+ offsets[b]); in == out allowed (each thread reads its items before it writes them)
template <class Load>
__global__ void __launch_bounds__(kThreads) scan_tile_kernel(Load in, uint32_t* out, int64_t n,
                                                             const uint32_t* __restrict__ offsets,
                                                             uint32_t* __restrict__ total_out) {
  const int64_t base = static_cast<int64_t>(blockIdx.x) * kTile + static_cast<int64_t>(threadIdx.x) * kItems;
  uint32_t v[kItems];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    v[k] = (base + k < n) ? in(base + k) : 0u;
    s += v[k];
  }
  uint32_t total;
  uint32_t run = block_exclusive_scan(s, total);
  if (offsets != nullptr) run += offsets[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    if (base + k < n) out[base + k] = run;
    run += v[k];
  }
  if (total_out != nullptr && threadIdx.x == 0 && gridDim.x == 1) *total_out = total;
}

__host__ __device__ inline int64_t tiles_of(int64_t n) { return (n + kTile - 1) / kTile; }
inline size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

// scratch (uint32 entries) needed by scan_u32 for n inputs
size_t scan_scratch_entries(int64_t n) {
  size_t e = 0;
  while (n > kTile) {
    n = tiles_of(n);
    e += align256(static_cast<size_t>(n) * 4) / 4;
  }
  return e + 64;
}

// exclusive scan of n values produced by `in` (a LoadU32 over `out` itself is allowed: each thread reads its items
// before it writes them); *total_out (device) = sum of all inputs
template <class Load>
int scan_any(Load in, uint32_t* out, int64_t n, uint32_t* scratch, uint32_t* total_out, cudaStream_t st) {
  const int64_t nb = tiles_of(n);
  if (nb <= 1) {
    scan_tile_kernel<Load><<<1, kThreads, 0, st>>>(in, out, n, nullptr, total_out);
    ORVB_CHECK_CUDA(cudaGetLastError());
    return ORVB_OK;
  }
  scan_reduce_kernel<Load><<<static_cast<unsigned>(nb), kThreads, 0, st>>>(in, n, scratch);
  ORVB_CHECK_CUDA(cudaGetLastError());
  uint32_t* next = scratch + align256(static_cast<size_t>(nb) * 4) / 4;
  const int rc = scan_any(LoadU32{scratch}, scratch, nb, next, total_out, st);
  if (rc != ORVB_OK) return rc;
  scan_tile_kernel<Load><<<static_cast<unsigned>(nb), kThreads, 0, st>>>(in, out, n, scratch, nullptr);
  ORVB_CHECK_CUDA(cudaGetLastError());
  return ORVB_OK;
}
inline int scan_u32(const uint32_t* in, uint32_t* out, int64_t n, uint32_t* scratch, uint32_t* total_out,
                    cudaStream_t st) {
  return scan_any(LoadU32{in}, out, n, scratch, total_out, st);
}

// ---- single-pass exclusive scan (decoupled look-back) ------------------------------------------------------------------
// One launch instead of reduce + scan-of-sums + scan: a tile publishes its aggregate, then walks back over its
// predecessors' descriptors, 32 at a time, until it meets one that already knows its inclusive prefix.  A descriptor is
// ONE 64-bit word (flag in bits 63:62, value in the low 32), so flag and value always arrive together.  Flag 3 = not
// ready: the descriptor array and the ticket counter are initialised by a memset with 0xFF bytes — the same memset that
// empties the voxel hash table — and tiles are handed out by an atomic ticket in scheduling order, so every predecessor
// of a running tile is running or done (forward progress).  The element order of the sum is fixed (integers): results
// are bit-identical run to run.
//   Load:  uint32_t operator()(int64_t i)                       element i
//   Store: void operator()(int64_t i, uint32_t exclusive, uint32_t value)
constexpr unsigned long long kDescAggregate = 1ull << 62, kDescInclusive = 2ull << 62, kDescFlagMask = 3ull << 62;

//          void total(uint32_t sum)                              called once, by the last tile, with the sum of all elements
struct StoreU32 {
  uint32_t* out;
  __device__ __forceinline__ void operator()(int64_t i, uint32_t exclusive, uint32_t) const { out[i] = exclusive; }
  __device__ __forceinline__ void total(uint32_t) const {}
};

template <class Load, class Store>
__global__ void __launch_bounds__(kThreads) scan_lookback_kernel(Load in, Store out, int64_t n,
                                                                 unsigned long long* __restrict__ desc,
                                                                 uint32_t* __restrict__ ticket,
                                                                 uint32_t* __restrict__ total_out) {
  __shared__ uint32_t s_tile, s_prefix;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u) + 1u;  // the counter starts at 0xFFFFFFFF: first ticket = 0
  __syncthreads();
  const uint32_t tile = s_tile;
  const int64_t base = static_cast<int64_t>(tile) * kTile + static_cast<int64_t>(threadIdx.x) * kItems;
  uint32_t v[kItems];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    v[k] = (base + k < n) ? in(base + k) : 0u;
    s += v[k];
  }
  uint32_t total;
  uint32_t run = block_exclusive_scan(s, total);
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    uint32_t prefix = 0;
    if (tile == 0) {
      if (lane == 0) atomicExch(&desc[0], kDescInclusive | total);
    } else {
      if (lane == 0) atomicExch(&desc[tile], kDescAggregate | total);
      int64_t look = static_cast<int64_t>(tile) - 1;
      while (true) {
        const int64_t idx = look - lane;
        unsigned long long d = kDescInclusive;  // in front of tile 0: an inclusive prefix of zero
        if (idx >= 0) d = *reinterpret_cast<volatile unsigned long long*>(desc + idx);
        const unsigned notready = __ballot_sync(kFull, (d & kDescFlagMask) == kDescFlagMask);
        const unsigned inclusive = __ballot_sync(kFull, (d & kDescFlagMask) == kDescInclusive);
        const int first_inc = inclusive ? __ffs(inclusive) - 1 : 32;
        const unsigned need = (first_inc >= 31) ? kFull : ((2u << first_inc) - 1u);
        if (notready & need) continue;  // a predecessor inside the window has not published yet: poll again
        uint32_t c = (lane <= first_inc) ? static_cast<uint32_t>(d) : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(kFull, c, o);
        prefix += c;
        if (inclusive) break;
        look -= 32;
      }
      if (lane == 0) atomicExch(&desc[tile], kDescInclusive | static_cast<unsigned long long>(prefix + total));
    }
    if (lane == 0) {
      s_prefix = prefix;
      if (static_cast<int64_t>(tile) == tiles_of(n) - 1) {
        if (total_out != nullptr) *total_out = prefix + total;
        out.total(prefix + total);
      }
    }
  }
  __syncthreads();
  run += s_prefix;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    if (base + k < n) out(base + k, run, v[k]);
    run += v[k];
  }
}

// desc: tiles_of(n) 64-bit words, ticket: one uint32 — both filled with 0xFF bytes before the launch
template <class Load, class Store>
int scan_lookback(Load in, Store out, int64_t n, unsigned long long* desc, uint32_t* ticket, uint32_t* total_out,
                  cudaStream_t st) {
  scan_lookback_kernel<Load, Store><<<static_cast<unsigned>(tiles_of(n)), kThreads, 0, st>>>(in, out, n, desc, ticket,
                                                                                             total_out);
  ORVB_CHECK_CUDA(cudaGetLastError());
  return ORVB_OK;
}

// ---- 4. stable LSD radix sort, BITS bits per pass, tiles of kThreads * ITEMS elements ------------------------------------
// hist[d * nblocks + b] = number of keys of tile b whose digit is d (digit-major, so one exclusive scan over the
// whole array yields the global start of (digit d, tile b)).
// n_dev (optional): device scalar clamping n — lets a launch sized for a capacity run on a count only the GPU knows.
// Defaults (9 bits, 8 items per thread) are the voxelizer's: 2048-element tiles keep the histogram array (512 entries
// per tile) far smaller than the data at millions of points.  Small inputs want small tiles (more CTAs in flight: the
// scatter is a latency chain of 2 * ITEMS match / shared-memory rounds per warp) and fewer bins.
// KeyFn: uint32_t operator()(int64_t i) — the key of element i (LoadU32 for a key array; the voxelizer's first pass
// computes it from the hash table instead of materialising it).
template <int BITS = kRadixBits, int ITEMS = kItems, class KeyFn = LoadU32>
__global__ void __launch_bounds__(kThreads) radix_hist_kernel(KeyFn keys, int n, int shift,
                                                              uint32_t* __restrict__ hist, int nblocks,
                                                              const uint32_t* __restrict__ n_dev = nullptr) {
  constexpr int BINS = 1 << BITS;
  __shared__ uint32_t h[BINS];
  if (n_dev != nullptr) n = static_cast<int>(min(static_cast<uint32_t>(n), *n_dev));
  for (int d = threadIdx.x; d < BINS; d += kThreads) h[d] = 0;
  __syncthreads();
  const int64_t base = static_cast<int64_t>(blockIdx.x) * (kThreads * ITEMS);
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const int64_t idx = base + k * kThreads + threadIdx.x;
    if (idx < n) atomicAdd(&h[(keys(idx) >> shift) & (BINS - 1)], 1u);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < BINS; d += kThreads) hist[static_cast<size_t>(d) * nblocks + blockIdx.x] = h[d];
}

// Scatter of one tile.  Warp w owns the 32 * ITEMS consecutive elements [tile + 32 ITEMS w, tile + 32 ITEMS (w+1)) and
// walks them 32 at a time, so (warp, iteration, lane) order = element order: ranks by __match_any_sync peers below the
// lane keep the sort stable.  vin == nullptr means "value = element index" (first pass).
template <int BITS = kRadixBits, int ITEMS = kItems, class KeyFn = LoadU32>
__global__ void __launch_bounds__(kThreads) radix_scatter_kernel(KeyFn kin,
                                                                 const uint32_t* __restrict__ vin,
                                                                 uint32_t* __restrict__ kout, uint32_t* __restrict__ vout,
                                                                 int n, int shift, const uint32_t* __restrict__ offs,
                                                                 int nblocks, const uint32_t* __restrict__ n_dev = nullptr) {
  constexpr int BINS = 1 << BITS;
  __shared__ uint32_t wh[kThreads / 32][BINS];
  if (n_dev != nullptr) n = static_cast<int>(min(static_cast<uint32_t>(n), *n_dev));
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  for (int w = 0; w < kThreads / 32; ++w)
    for (int d = threadIdx.x; d < BINS; d += kThreads) wh[w][d] = 0;
  __syncthreads();
  const int64_t wbase = static_cast<int64_t>(blockIdx.x) * (kThreads * ITEMS) + warp * (32 * ITEMS);
  // this lane's keys and values, requested together up front: with a key functor that chases pointers (the voxelizer's
  // first pass: slot -> table) a load per phase and iteration was a chain of 2 * ITEMS dependent round trips
  uint32_t kreg[ITEMS], vreg[ITEMS];
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    const int64_t idx = wbase + it * 32 + lane;
    kreg[it] = (idx < n) ? kin(idx) : 0u;
    vreg[it] = (idx < n) ? ((vin != nullptr) ? vin[idx] : static_cast<uint32_t>(idx)) : 0u;
  }
  // A: per-warp digit histogram
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    const int64_t idx = wbase + it * 32 + lane;
    const bool act = idx < n;
    const uint32_t d = act ? ((kreg[it] >> shift) & (BINS - 1)) : (BINS + lane);  // inactive lanes match nobody
    const unsigned peers = __match_any_sync(kFull, d);
    if (act && lane == __ffs(peers) - 1) wh[warp][d] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  // B: per digit, turn the per-warp counts into global start positions
  for (int d = threadIdx.x; d < BINS; d += kThreads) {
    uint32_t run = offs[static_cast<size_t>(d) * nblocks + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
      const uint32_t t = wh[w][d];
      wh[w][d] = run;
      run += t;
    }
  }
  __syncthreads();
  // C: stable scatter
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    const int64_t idx = wbase + it * 32 + lane;
    const bool act = idx < n;
    const uint32_t key = kreg[it];
    const uint32_t d = act ? ((key >> shift) & (BINS - 1)) : (BINS + lane);
    const unsigned peers = __match_any_sync(kFull, d);
    const uint32_t start = act ? wh[warp][d] : 0u;
    __syncwarp();  // every lane has read its start before a leader advances it
    if (act) {
      const uint32_t dest = start + __popc(peers & ((1u << lane) - 1u));
      kout[dest] = key;
      vout[dest] = vreg[it];
      if (lane == __ffs(peers) - 1) wh[warp][d] = start + __popc(peers);
    }
    __syncwarp();
  }
}

// uint32 entries of the histogram array one pass needs
inline size_t radix_hist_entries(int64_t n, int digit_bits, int items) {
  const int64_t tile = static_cast<int64_t>(kThreads) * items;
  return static_cast<size_t>(((n + tile - 1) / tile) << digit_bits);
}

template <int BITS, int ITEMS>
inline int radix_sort_pairs_t(uint32_t* ka, uint32_t* va, uint32_t* kb, uint32_t* vb, bool identity_values, int n, int bits,
                              uint32_t* hist, uint32_t* scratch, const uint32_t** k_sorted, const uint32_t** v_sorted,
                              cudaStream_t st, const uint32_t* n_dev) {
  constexpr int tile = kThreads * ITEMS;
  const int nblocks = (n + tile - 1) / tile;
  const uint32_t* kin = ka;
  const uint32_t* vin = identity_values ? nullptr : va;
  uint32_t* kout = kb;
  uint32_t* vout = vb;
  const int passes = (bits + BITS - 1) / BITS;
  for (int pass = 0; pass < passes; ++pass) {
    const int shift = BITS * pass;
    radix_hist_kernel<BITS, ITEMS><<<nblocks, kThreads, 0, st>>>(LoadU32{kin}, n, shift, hist, nblocks, n_dev);
    ORVB_CHECK_CUDA(cudaGetLastError());
    const int rc = scan_u32(hist, hist, static_cast<int64_t>(nblocks) << BITS, scratch, nullptr, st);
    if (rc != ORVB_OK) return rc;
    radix_scatter_kernel<BITS, ITEMS><<<nblocks, kThreads, 0, st>>>(LoadU32{kin}, vin, kout, vout, n, shift, hist, nblocks, n_dev);
    ORVB_CHECK_CUDA(cudaGetLastError());
    kin = kout;
    vin = vout;
    kout = (kout == kb) ? ka : kb;
    vout = (vout == vb) ? va : vb;
  }
  *k_sorted = kin;
  *v_sorted = vin;
  return ORVB_OK;
}

// Stable LSD radix sort of (key, value) pairs over the low `bits` bits of the keys, ping-ponging between (ka, va) and
// (kb, vb).  identity_values: values start as the element index.  digit_bits in {8, 9, 10}, items in {2, 8}.
// hist: radix_hist_entries(n, digit_bits, items) uint32; scratch: scan_scratch_entries(that).  Returns through
// *k_sorted / *v_sorted which buffer holds the result.
inline int radix_sort_pairs(uint32_t* ka, uint32_t* va, uint32_t* kb, uint32_t* vb, bool identity_values, int n, int bits,
                            int digit_bits, int items, uint32_t* hist, uint32_t* scratch, const uint32_t** k_sorted,
                            const uint32_t** v_sorted, cudaStream_t st, const uint32_t* n_dev = nullptr) {
#define ORVB_RADIX_CASE(B, I)                                                                                          \
  if (digit_bits == B && items == I)                                                                                   \
    return radix_sort_pairs_t<B, I>(ka, va, kb, vb, identity_values, n, bits, hist, scratch, k_sorted, v_sorted, st, n_dev);
  ORVB_RADIX_CASE(8, 2)
  ORVB_RADIX_CASE(8, 8)
  ORVB_RADIX_CASE(9, 2)
  ORVB_RADIX_CASE(9, 8)
  ORVB_RADIX_CASE(10, 2)
  ORVB_RADIX_CASE(10, 8)
#undef ORVB_RADIX_CASE
  set_error("radix_sort_pairs: unsupported digit width %d / items %d", digit_bits, items);
  return ORVB_EINVAL;
}

}  // namespace
}  // namespace orvb
