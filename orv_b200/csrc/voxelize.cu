// Point-cloud voxelization for the occupancy condition: the B200 replacement of the reference's
// `orv/ops/voxelize` extension (voxelization.py:41-122 -> voxelization_kernel.cu / voxelization_cpu.cpp) and of the
// label vote its caller runs on the result (orv/dataset/prepare_dataset.py:137-198).
//
// Integer / index work, HBM-bound.  The reference's deterministic CUDA path scans, for every point, ALL earlier
// points for a duplicate voxel (point_to_voxelidx_kernel, voxelization_kernel.cuh:90-132: O(N^2) loads) and then
// numbers the voxels in a <<<1,1>>> kernel (determin_voxel_num, :134-162).  Here the same result — voxels numbered
// in order of first appearance, points kept in index order, max_points / max_voxels caps — comes from O(N) passes:
//
//   0. ONE memset (0xFF bytes) empties the hash table and arms the look-back descriptors / tickets of the three scans
//   1. voxel_insert_kernel   128-bit point loads -> (z,y,x) cell -> open-addressing hash insert of the 64-bit cell
//                            key, atomicMin of the point index per occupied slot (= first point of the voxel); the only
//                            per-point output is the slot number (4 bytes)
//   2. ONE single-pass (decoupled look-back) exclusive scan of the flags "point i is the first of its voxel", evaluated
//      on the fly; the scan's store step writes the voxel number (order of first appearance, or the drop sentinel
//      beyond max_voxels) into the table slot of every first point — no per-point order / key arrays exist
//   3. stable LSD radix sort of (voxel number, point index), 9 bits per pass, ceil(log2(cap+1)/9) passes, each pass =
//      histogram + single-pass scan + scatter; the first pass reads its keys through the table
//      (key(i) = t_vox[slot_of[i]]): points of a voxel end up contiguous and in index order, so "position in voxel" =
//      sorted position - segment start
//   4. segment_kernel / gather_kernel   segment bounds per voxel, feature copy of the first max_points points; the
//                            voxel coordinates are recomputed from the voxel's first point (same arithmetic as step 1)
//   5. label_vote_kernel     (optional) one warp per voxel: most frequent label among the max_points slots of the
//                            voxel (empty slots count as label 0, which yields to the runner-up), without ever
//                            materialising the reference's [max_voxels, max_points, C] tensor (160 MB at its
//                            1e5 x 100 x 4 call site) or copying it to the host.
//
// No atomics decide an output value except atomicMin / atomicCAS whose results are order-independent, so the output
// is bit-identical run to run and identical to the reference's deterministic path.
#include "common.cuh"
#include "sortscan.cuh"

#include <math.h>

namespace orvb {
namespace {

constexpr unsigned long long kEmptyKey = ~0ull;
#ifndef ORVB_VOX_ITEMS
#define ORVB_VOX_ITEMS 8
#endif
constexpr int kVoxItems = ORVB_VOX_ITEMS;  // elements per thread of a radix-sort tile (tile = 256 x this)

struct VoxGeom {
  float lo[3];
  float vs[3];
  int32_t grid[3];
};

// floor((p - lo) / vs) per axis with IEEE subtraction / division (what the reference's CPU loop and its CUDA kernel
// both evaluate: voxelization_cpu.cpp:21, voxelization_kernel.cuh:23-41); false when any axis leaves [0, grid).
// NaN coordinates are invalid (CPU-reference behaviour: the cast of NaN to int is negative on x86).
__device__ __forceinline__ bool voxel_cell(const VoxGeom& g, float px, float py, float pz, int& cx, int& cy, int& cz) {
  const float fx = floorf(__fdiv_rn(__fsub_rn(px, g.lo[0]), g.vs[0]));
  const float fy = floorf(__fdiv_rn(__fsub_rn(py, g.lo[1]), g.vs[1]));
  const float fz = floorf(__fdiv_rn(__fsub_rn(pz, g.lo[2]), g.vs[2]));
  const float lim = 2147483648.f;
  if (!(fx >= 0.f && fx < lim && fy >= 0.f && fy < lim && fz >= 0.f && fz < lim)) return false;
  cx = static_cast<int>(fx);
  cy = static_cast<int>(fy);
  cz = static_cast<int>(fz);
  return cx < g.grid[0] && cy < g.grid[1] && cz < g.grid[2];
}

template <bool VEC4>
__device__ __forceinline__ void load_xyz(const float* __restrict__ points, int64_t i, int c, float& x, float& y,
                                         float& z) {
  if (VEC4) {
    const float4 p = __ldg(reinterpret_cast<const float4*>(points) + i);  // one coalesced 128-bit load per point
    x = p.x; y = p.y; z = p.z;
  } else {
    const float* p = points + i * c;
    x = __ldg(p); y = __ldg(p + 1); z = __ldg(p + 2);
  }
}

// ---- dynamic voxelization: coors[i] = (z, y, x) or (-1, -1, -1) ---------------------------------------------------
template <bool VEC4>
__global__ void __launch_bounds__(kThreads) dynamic_voxelize_kernel(const float* __restrict__ points, int n, int c,
                                                                    VoxGeom g, int32_t* __restrict__ coors) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
  if (i >= n) return;
  float x, y, z;
  load_xyz<VEC4>(points, i, c, x, y, z);
  int cx, cy, cz;
  const bool ok = voxel_cell(g, x, y, z, cx, cy, cz);
  int32_t* o = coors + i * 3;
  o[0] = ok ? cz : -1;
  o[1] = ok ? cy : -1;
  o[2] = ok ? cx : -1;
}

// ---- 1. cell + hash insert ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hash_slot(unsigned long long key, int log2_slots) {
  return static_cast<uint32_t>((key * 0x9E3779B97F4A7C15ull) >> (64 - log2_slots));
}

template <bool VEC4>
__global__ void __launch_bounds__(kThreads) voxel_insert_kernel(const float* __restrict__ points, int n, int c, VoxGeom g,
                                                                int32_t* __restrict__ slot_of /*[n]*/,
                                                                unsigned long long* __restrict__ t_keys,
                                                                uint32_t* __restrict__ t_first, int log2_slots) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
  if (i >= n) return;
  float x, y, z;
  load_xyz<VEC4>(points, i, c, x, y, z);
  int cx, cy, cz;
  if (!voxel_cell(g, x, y, z, cx, cy, cz)) {
    slot_of[i] = -1;
    return;
  }
  const unsigned long long key =
      (static_cast<unsigned long long>(cz) * static_cast<unsigned long long>(g.grid[1]) + cy) *
          static_cast<unsigned long long>(g.grid[0]) + cx;
  const uint32_t mask = (1u << log2_slots) - 1u;
  uint32_t s = hash_slot(key, log2_slots);
  while (true) {  // the table has >= 2n slots, so an empty one is always reachable
    const unsigned long long prev = atomicCAS(&t_keys[s], kEmptyKey, key);
    if (prev == kEmptyKey || prev == key) break;
    s = (s + 1u) & mask;
  }
  // (measured: reading the slot first and skipping the atomics when an earlier point already holds it is SLOWER —
  // 57.8 vs 38.1 us for 2 M points: two dependent L2 round trips in front of the atomic instead of one fire-and-forget)
  atomicMin(&t_first[s], static_cast<uint32_t>(i));  // empty = 0xFFFFFFFF (the table memset)
  slot_of[i] = static_cast<int32_t>(s);
}

// ---- 2. first-of-voxel flags (evaluated inside the scan, never stored) and voxel numbers -----------------------------
struct LoadFirstFlag {  // 1 when point i is the first (lowest-index) point of its voxel
  const int32_t* slot_of;
  const uint32_t* t_first;
  __device__ __forceinline__ uint32_t operator()(int64_t i) const {
    const int32_t s = __ldg(slot_of + i);
    return (s >= 0 && t_first[s] == static_cast<uint32_t>(i)) ? 1u : 0u;
  }
};
// The exclusive prefix of a first point IS its voxel's number (order of first appearance); voxels beyond max_voxels
// get the drop sentinel `cap`, which sorts last (voxelization_cpu.cpp:83: a new voxel is refused once voxel_num >=
// max_voxels, and so are its later points).
struct StoreVoxelNumber {
  const int32_t* slot_of;
  uint32_t* t_vox;
  uint32_t cap;
  long long* voxel_num;  // the caller's count: min(number of voxels, max_voxels), written by the last tile
  __device__ __forceinline__ void operator()(int64_t i, uint32_t exclusive, uint32_t flag) const {
    if (flag) t_vox[slot_of[i]] = exclusive < cap ? exclusive : cap;
  }
  __device__ __forceinline__ void total(uint32_t sum) const { *voxel_num = static_cast<long long>(sum < cap ? sum : cap); }
};
// ---- 3. sort key of a point = number of its voxel (read through the table; invalid points sort last) -------------------
struct KeyOfPoint {
  const int32_t* slot_of;
  const uint32_t* t_vox;
  uint32_t cap;
  __device__ __forceinline__ uint32_t operator()(int64_t i) const {
    const int32_t s = __ldg(slot_of + i);
    return s >= 0 ? t_vox[s] : cap;
  }
};
// ---- 5. segments + gather ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) segment_kernel(const uint32_t* __restrict__ skey, int n, uint32_t cap,
                                                           int32_t* __restrict__ seg_start,
                                                           int32_t* __restrict__ seg_end) {
  const int64_t p = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
  if (p >= n) return;
  const uint32_t v = skey[p];
  if (v >= cap) return;
  if (p == 0 || skey[p - 1] != v) seg_start[v] = static_cast<int32_t>(p);
  if (p == n - 1 || skey[p + 1] != v) seg_end[v] = static_cast<int32_t>(p + 1);
}

template <bool VEC4>
__global__ void __launch_bounds__(kThreads) gather_kernel(const float* __restrict__ points, int n, int c,
                                                          const uint32_t* __restrict__ skey,
                                                          const uint32_t* __restrict__ sval, uint32_t cap,
                                                          const int32_t* __restrict__ seg_start,
                                                          const int32_t* __restrict__ seg_end,
                                                          VoxGeom g, int max_points,
                                                          float* __restrict__ voxels, int32_t* __restrict__ coors,
                                                          int32_t* __restrict__ num_points) {
  const int64_t p = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
  if (p >= n) return;
  const uint32_t v = skey[p];
  if (v >= cap) return;
  const int32_t start = seg_start[v];
  const int32_t rank = static_cast<int32_t>(p) - start;
  const uint32_t idx = sval[p];
  if (rank == 0) {
    if (coors != nullptr) {  // (z, y, x) of the voxel = the cell of its first point
      float x, y, z;
      load_xyz<VEC4>(points, idx, c, x, y, z);
      int cx = 0, cy = 0, cz = 0;
      voxel_cell(g, x, y, z, cx, cy, cz);
      coors[3 * static_cast<size_t>(v) + 0] = cz;
      coors[3 * static_cast<size_t>(v) + 1] = cy;
      coors[3 * static_cast<size_t>(v) + 2] = cx;
    }
    if (num_points != nullptr) num_points[v] = min(seg_end[v] - start, max_points);
  }
  if (voxels == nullptr || rank >= max_points) return;
  const size_t dst = (static_cast<size_t>(v) * max_points + rank) * c;
  if (VEC4) {
    *reinterpret_cast<float4*>(voxels + dst) = __ldg(reinterpret_cast<const float4*>(points) + idx);
  } else {
    const float* src = points + static_cast<size_t>(idx) * c;
    for (int k = 0; k < c; ++k) voxels[dst + k] = __ldg(src + k);
  }
}

// ---- 6. label vote (points_to_voxels, prepare_dataset.py:176-196) ------------------------------------------------------
// One warp per voxel.  The label of a point is its LAST feature; the reference counts labels over the max_points
// slots of the voxel (unfilled slots hold 0), sorts the counts in descending order (stable on the host: ties go to the
// smaller label) and takes the runner-up when the winner is 0.  out[v] = (x, y, z, label - 1) as float64, the dtype
// numpy gives the reference's concatenate of int32 coordinates and float32 labels.
struct Vote {
  int count;
  float value;
};
__device__ __forceinline__ bool vote_better(const Vote& a, const Vote& b) {  // a strictly ahead of b
  return a.count > b.count || (a.count == b.count && a.count > 0 && a.value < b.value);
}
constexpr int kVoteGroup = 8;    // lanes per voxel: the occupancy caller's voxels hold ~20 points, a full warp idles
constexpr int kVoteStage = 128;  // labels of a voxel staged in shared memory (per group); the rest re-read from L2

__device__ __forceinline__ Vote vote_group_best(Vote v) {
#pragma unroll
  for (int o = kVoteGroup / 2; o > 0; o >>= 1) {
    Vote t;
    t.count = __shfl_xor_sync(kFull, v.count, o);
    t.value = __shfl_xor_sync(kFull, v.value, o);
    if (vote_better(t, v)) v = t;
  }
  return v;
}

// Eight lanes per voxel (four voxels per warp): the kernel is a chain of dependent loads per voxel (segment bounds ->
// point indices -> labels -> first point), so more voxels in flight per warp is what shortens it.
__global__ void __launch_bounds__(kThreads) label_vote_kernel(const float* __restrict__ points, int c,
                                                              const uint32_t* __restrict__ sval,
                                                              const int32_t* __restrict__ seg_start,
                                                              const int32_t* __restrict__ seg_end,
                                                              VoxGeom g, int max_points,
                                                              const long long* __restrict__ voxel_num,
                                                              double* __restrict__ out) {
  __shared__ float stage[kThreads / kVoteGroup][kVoteStage];
  const int gl = threadIdx.x & (kVoteGroup - 1);
  const int grp = threadIdx.x / kVoteGroup;
  const long long nv = *voxel_num;
  const long long v = static_cast<long long>(blockIdx.x) * (kThreads / kVoteGroup) + grp;
  const bool live = v < nv;  // (all lanes stay for the shuffles)
  const int32_t start = live ? seg_start[v] : 0;
  const int kept = live ? min(seg_end[v] - start, max_points) : 0;
  const int pad = max_points - kept;
  float* my = stage[grp];
  for (int e = gl; e < kept && e < kVoteStage; e += kVoteGroup)
    my[e] = __ldg(points + static_cast<size_t>(sval[start + e]) * c + (c - 1));
  __syncwarp();
  auto label_at = [&](int e) -> float {
    return e < kVoteStage ? my[e] : __ldg(points + static_cast<size_t>(sval[start + e]) * c + (c - 1));
  };
  Vote best_all = {0, 0.f}, best_nz = {0, 0.f};
  if (live && gl == 0 && pad > 0) {  // the empty slots: label 0
    int cnt = pad;
    for (int f = 0; f < kept; ++f) cnt += (label_at(f) == 0.f) ? 1 : 0;
    best_all.count = cnt;
    best_all.value = 0.f;
  }
  for (int e = gl; e < kept; e += kVoteGroup) {
    const float le = label_at(e);
    int cnt = (le == 0.f) ? pad : 0;
    for (int f = 0; f < kept; ++f) cnt += (label_at(f) == le) ? 1 : 0;
    Vote cand = {cnt, le};
    if (vote_better(cand, best_all)) best_all = cand;
    if (le != 0.f && vote_better(cand, best_nz)) best_nz = cand;
  }
  best_all = vote_group_best(best_all);
  best_nz = vote_group_best(best_nz);
  if (live && gl == 0) {
    const float top = (best_all.value == 0.f && best_nz.count > 0) ? best_nz.value : best_all.value;
    const uint32_t idx = sval[start];
    const float* pp = points + static_cast<size_t>(idx) * c;
    int cx = 0, cy = 0, cz = 0;
    voxel_cell(g, __ldg(pp), __ldg(pp + 1), __ldg(pp + 2), cx, cy, cz);
    double* o = out + 4 * v;
    o[0] = static_cast<double>(cx);
    o[1] = static_cast<double>(cy);
    o[2] = static_cast<double>(cz);
    o[3] = static_cast<double>(top - 1.f);
  }
}

// ---- host side ------------------------------------------------------------------------------------------------------------
int make_geom(const float* voxel_size, const float* coors_range, VoxGeom* g) {
  for (int i = 0; i < 3; ++i) {
    ORVB_REQUIRE(voxel_size[i] > 0.f, ORVB_EINVAL, "voxelize: voxel_size[%d] = %g must be positive", i, voxel_size[i]);
    g->lo[i] = coors_range[i];
    g->vs[i] = voxel_size[i];
    // grid_size = round((max - min) / voxel) in float, as voxelization_cpu.cpp:120-123 / voxelization_kernel.cu:28-30
    const float extent = (coors_range[3 + i] - coors_range[i]) / voxel_size[i];
    const double r = round(static_cast<double>(extent));
    ORVB_REQUIRE(r >= 0 && r < 2147483648.0, ORVB_ESHAPE, "voxelize: grid extent %g on axis %d is out of range", r, i);
    g->grid[i] = static_cast<int32_t>(r);
  }
  // the 64-bit cell key (z * gy + y) * gx + x must stay below the empty-slot marker
  ORVB_REQUIRE(static_cast<double>(g->grid[0]) * g->grid[1] * g->grid[2] < 1.8e19, ORVB_ESHAPE,
               "voxelize: grid %d x %d x %d has more than 2^64 cells", g->grid[0], g->grid[1], g->grid[2]);
  return ORVB_OK;
}

struct VoxWorkspace {
  int log2_slots;
  size_t slots;
  uint32_t cap;
  int passes;
  int nblocks;
  // [off_keys, off_keys + init_bytes) is what the one memset (0xFF) covers: table keys, first indices, the descriptors
  // and tickets of the three look-back scans
  size_t off_keys, off_first, off_desc, off_ticket, init_bytes;
  size_t desc_stride;  // 64-bit words per scan
  size_t off_slot, off_vox, off_total, off_ka, off_va, off_kb, off_vb, off_hist, off_seg_start, off_seg_end, bytes;
};

VoxWorkspace plan_workspace(int32_t n, int32_t max_voxels) {
  VoxWorkspace w = {};
  const size_t nn = static_cast<size_t>(n > 0 ? n : 1);
  w.log2_slots = 10;
  while ((static_cast<size_t>(1) << w.log2_slots) < 2 * nn) ++w.log2_slots;
  w.slots = static_cast<size_t>(1) << w.log2_slots;
  w.cap = static_cast<uint32_t>(n < max_voxels ? n : max_voxels);
  int bits = 1;
  while ((1ull << bits) <= w.cap) ++bits;  // keys take values 0..cap
  w.passes = (bits + kRadixBits - 1) / kRadixBits;
  w.nblocks = static_cast<int>((nn + static_cast<size_t>(kThreads) * kVoxItems - 1) / (static_cast<size_t>(kThreads) * kVoxItems));
  const size_t hist_entries = static_cast<size_t>(kBins) * w.nblocks;
  w.desc_stride = static_cast<size_t>(tiles_of(static_cast<int64_t>(hist_entries > nn ? hist_entries : nn))) + 1;
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t at = o; o += align256(bytes); return at; };
  w.off_keys = take(w.slots * 8);
  w.off_first = take(w.slots * 4);
  w.off_desc = take((1 + static_cast<size_t>(w.passes)) * w.desc_stride * 8);
  w.off_ticket = take(256);
  w.init_bytes = o - w.off_keys;
  w.off_slot = take(nn * 4);
  w.off_vox = take(w.slots * 4);
  w.off_total = take(256);
  w.off_ka = take(nn * 4);
  w.off_va = take(nn * 4);
  w.off_kb = take(nn * 4);
  w.off_vb = take(nn * 4);
  w.off_hist = take(hist_entries * 4);
  w.off_seg_start = take((static_cast<size_t>(w.cap) + 1) * 4);
  w.off_seg_end = take((static_cast<size_t>(w.cap) + 1) * 4);
  w.bytes = o;
  return w;
}

}  // namespace
}  // namespace orvb

using namespace orvb;

extern "C" int orvb_dynamic_voxelize(const float* points, int32_t n, int32_t c, const float* voxel_size,
                                     const float* coors_range, int32_t* coors, void* stream) {
  int rc = check_arch();
  if (rc != ORVB_OK) return rc;
  ORVB_REQUIRE(voxel_size != nullptr && coors_range != nullptr, ORVB_EINVAL, "dynamic_voxelize: null geometry");
  ORVB_REQUIRE(n >= 0 && c >= 3, ORVB_ESHAPE, "dynamic_voxelize: points must be [n, >= 3], got [%d, %d]", n, c);
  if (n == 0) return ORVB_OK;
  ORVB_REQUIRE(points != nullptr && coors != nullptr, ORVB_EINVAL, "dynamic_voxelize: null pointer");
  VoxGeom g;
  rc = make_geom(voxel_size, coors_range, &g);
  if (rc != ORVB_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned blocks = static_cast<unsigned>((static_cast<int64_t>(n) + kThreads - 1) / kThreads);
  const bool vec4 = (c == 4) && (reinterpret_cast<uintptr_t>(points) % 16 == 0);
  if (vec4)
    dynamic_voxelize_kernel<true><<<blocks, kThreads, 0, st>>>(points, n, c, g, coors);
  else
    dynamic_voxelize_kernel<false><<<blocks, kThreads, 0, st>>>(points, n, c, g, coors);
  ORVB_CHECK_CUDA(cudaGetLastError());
  return ORVB_OK;
}

extern "C" size_t orvb_voxelize_workspace_bytes(int32_t n, int32_t max_voxels) {
  if (n < 0 || max_voxels <= 0) return 0;
  return plan_workspace(n, max_voxels).bytes;
}

extern "C" int orvb_hard_voxelize(const orvb_voxelize_args* a, void* stream) {
  int rc = check_arch();
  if (rc != ORVB_OK) return rc;
  ORVB_REQUIRE(a != nullptr, ORVB_EINVAL, "hard_voxelize: null args");
  ORVB_REQUIRE(a->n >= 0 && a->n < (1 << 30) && a->c >= 3, ORVB_ESHAPE,
               "hard_voxelize: points must be [n < 2^30, >= 3], got [%d, %d]", a->n, a->c);
  ORVB_REQUIRE(a->max_points > 0 && a->max_voxels > 0, ORVB_EINVAL,
               "hard_voxelize: max_points = %d and max_voxels = %d must be positive (use orvb_dynamic_voxelize for -1)",
               a->max_points, a->max_voxels);
  ORVB_REQUIRE(a->voxel_num != nullptr, ORVB_EINVAL, "hard_voxelize: voxel_num is required");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->n == 0) {
    ORVB_CHECK_CUDA(cudaMemsetAsync(a->voxel_num, 0, sizeof(int64_t), st));
    return ORVB_OK;
  }
  ORVB_REQUIRE(a->points != nullptr && a->workspace != nullptr, ORVB_EINVAL, "hard_voxelize: null pointer");
  VoxGeom g;
  rc = make_geom(a->voxel_size, a->coors_range, &g);
  if (rc != ORVB_OK) return rc;
  const VoxWorkspace w = plan_workspace(a->n, a->max_voxels);
  ORVB_REQUIRE(a->workspace_bytes >= w.bytes, ORVB_ENOMEM, "hard_voxelize: workspace %zu < %zu bytes",
               a->workspace_bytes, w.bytes);
  ORVB_REQUIRE(reinterpret_cast<uintptr_t>(a->workspace) % 256 == 0, ORVB_ESHAPE,
               "hard_voxelize: workspace must be 256-byte aligned");
  char* base = static_cast<char*>(a->workspace);
  int32_t* slot_of = reinterpret_cast<int32_t*>(base + w.off_slot);
  unsigned long long* t_keys = reinterpret_cast<unsigned long long*>(base + w.off_keys);
  uint32_t* t_first = reinterpret_cast<uint32_t*>(base + w.off_first);
  uint32_t* t_vox = reinterpret_cast<uint32_t*>(base + w.off_vox);
  unsigned long long* desc = reinterpret_cast<unsigned long long*>(base + w.off_desc);
  uint32_t* ticket = reinterpret_cast<uint32_t*>(base + w.off_ticket);
  uint32_t* total = reinterpret_cast<uint32_t*>(base + w.off_total);
  uint32_t* ka = reinterpret_cast<uint32_t*>(base + w.off_ka);
  uint32_t* va = reinterpret_cast<uint32_t*>(base + w.off_va);
  uint32_t* kb = reinterpret_cast<uint32_t*>(base + w.off_kb);
  uint32_t* vb = reinterpret_cast<uint32_t*>(base + w.off_vb);
  uint32_t* hist = reinterpret_cast<uint32_t*>(base + w.off_hist);
  int32_t* seg_start = reinterpret_cast<int32_t*>(base + w.off_seg_start);
  int32_t* seg_end = reinterpret_cast<int32_t*>(base + w.off_seg_end);

  const int n = a->n;
  const unsigned blocks = static_cast<unsigned>((static_cast<int64_t>(n) + kThreads - 1) / kThreads);
  const bool vec4 = (a->c == 4) && (reinterpret_cast<uintptr_t>(a->points) % 16 == 0);
  const bool vec4_out = vec4 && (a->voxels == nullptr || reinterpret_cast<uintptr_t>(a->voxels) % 16 == 0);

  // empty table (key = all ones, first index = 0xFFFFFFFF) + armed scan descriptors / tickets: one memset
  ORVB_CHECK_CUDA(cudaMemsetAsync(t_keys, 0xff, w.init_bytes, st));
  if (vec4)
    voxel_insert_kernel<true><<<blocks, kThreads, 0, st>>>(a->points, n, a->c, g, slot_of, t_keys, t_first, w.log2_slots);
  else
    voxel_insert_kernel<false><<<blocks, kThreads, 0, st>>>(a->points, n, a->c, g, slot_of, t_keys, t_first, w.log2_slots);
  ORVB_CHECK_CUDA(cudaGetLastError());
  // voxel numbers (order of first appearance) into the table, voxel count
  rc = scan_lookback(LoadFirstFlag{slot_of, t_first},
                     StoreVoxelNumber{slot_of, t_vox, w.cap, reinterpret_cast<long long*>(a->voxel_num)}, n, desc, ticket,
                     total, st);
  if (rc != ORVB_OK) return rc;

  // stable LSD radix sort of (voxel number, point index); the first pass reads its keys through the table
  const uint32_t* kin = nullptr;
  const uint32_t* vin = nullptr;  // identity
  uint32_t* kout = kb;
  uint32_t* vout = vb;
  const KeyOfPoint key0{slot_of, t_vox, w.cap};
  const int64_t hist_entries = static_cast<int64_t>(kBins) * w.nblocks;
  for (int pass = 0; pass < w.passes; ++pass) {
    const int shift = kRadixBits * pass;
    if (pass == 0) radix_hist_kernel<kRadixBits, kVoxItems, KeyOfPoint><<<w.nblocks, kThreads, 0, st>>>(key0, n, shift, hist, w.nblocks);
    else radix_hist_kernel<kRadixBits, kVoxItems><<<w.nblocks, kThreads, 0, st>>>(LoadU32{kin}, n, shift, hist, w.nblocks);
    ORVB_CHECK_CUDA(cudaGetLastError());
    rc = scan_lookback(LoadU32{hist}, StoreU32{hist}, hist_entries, desc + (1 + pass) * w.desc_stride, ticket + 1 + pass,
                       nullptr, st);
    if (rc != ORVB_OK) return rc;
    if (pass == 0)
      radix_scatter_kernel<kRadixBits, kVoxItems, KeyOfPoint><<<w.nblocks, kThreads, 0, st>>>(key0, vin, kout, vout, n, shift, hist,
                                                                                         w.nblocks);
    else
      radix_scatter_kernel<kRadixBits, kVoxItems><<<w.nblocks, kThreads, 0, st>>>(LoadU32{kin}, vin, kout, vout, n, shift, hist, w.nblocks);
    ORVB_CHECK_CUDA(cudaGetLastError());
    kin = kout;
    vin = vout;
    kout = (kout == kb) ? ka : kb;
    vout = (vout == vb) ? va : vb;
  }
  const uint32_t* skey = kin;
  const uint32_t* sval = vin;

  segment_kernel<<<blocks, kThreads, 0, st>>>(skey, n, w.cap, seg_start, seg_end);
  ORVB_CHECK_CUDA(cudaGetLastError());
  if (a->voxels != nullptr || a->coors != nullptr || a->num_points_per_voxel != nullptr) {
    if (vec4_out)
      gather_kernel<true><<<blocks, kThreads, 0, st>>>(a->points, n, a->c, skey, sval, w.cap, seg_start, seg_end, g,
                                                      a->max_points, a->voxels, a->coors, a->num_points_per_voxel);
    else
      gather_kernel<false><<<blocks, kThreads, 0, st>>>(a->points, n, a->c, skey, sval, w.cap, seg_start, seg_end, g,
                                                       a->max_points, a->voxels, a->coors, a->num_points_per_voxel);
    ORVB_CHECK_CUDA(cudaGetLastError());
  }
  if (a->voxel_labels != nullptr) {
    const unsigned vblocks = (w.cap + kThreads / kVoteGroup - 1) / (kThreads / kVoteGroup);
    if (vblocks > 0) {
      label_vote_kernel<<<vblocks, kThreads, 0, st>>>(a->points, a->c, sval, seg_start, seg_end, g, a->max_points,
                                                     reinterpret_cast<const long long*>(a->voxel_num),
                                                     a->voxel_labels);
      ORVB_CHECK_CUDA(cudaGetLastError());
    }
  }
  return ORVB_OK;
}
