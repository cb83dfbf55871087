// Micro-benchmarks behind the attention kernel's design (run under gpurun): per-SM throughput of tcgen05.ld (TMEM ->
// registers) and of MUFU.EX2 as a function of the number of warps issuing them.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/microbench_tmem tools/microbench_tmem.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../orv_b200/csrc/ptx.cuh"
using namespace orvb;

__global__ void __launch_bounds__(512, 1) k_tmem(int iters, int nwarps, int mode, long long* out, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + static_cast<uint32_t>((warp >> 2) * 128);
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps) {
    for (int it = 0; it < iters; ++it) {
      if (mode == 0) {          // two x32 loads, one wait (what the softmax warps do per key tile)
        uint32_t r0[32], r1[32];
        tmem_ld_32x32b_x32(base, r0);
        tmem_ld_32x32b_x32(base + 32, r1);
        tmem_ld_wait();
        uint32_t x = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) x ^= r0[i] ^ r1[i];
        acc += __uint_as_float(x);
      } else if (mode == 2) {   // 64 ex2 per thread as 32 x ex2.approx.f16x2 (compiles to 2 MUFU.EX2.F16 + PRMT each)
        uint32_t x = __float_as_uint(acc) | 0x3c003c00u;
#pragma unroll
        for (int i = 0; i < 32; ++i) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x + i)); x ^= y; }
        acc += __uint_as_float(x);
      } else if (mode == 1) {   // 64 ex2 per thread
        float x = acc;
#pragma unroll
        for (int i = 0; i < 64; ++i) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x + i)); acc += y; }
      }
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}


template <int N> struct LdN;
#define DEF_LDN(N, REGS, ...)                                                                          \
  template <> struct LdN<N> {                                                                          \
    static __device__ __forceinline__ void ld(uint32_t a, uint32_t* r) {                               \
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x" #N ".b32 {" REGS "}, [%" #N "];" : __VA_ARGS__ : "r"(a) : "memory"); \
    }                                                                                                  \
  };
DEF_LDN(8, "%0,%1,%2,%3,%4,%5,%6,%7", "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]))
DEF_LDN(16, "%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15", "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]))

template <int N, int REPS>
__global__ void __launch_bounds__(512, 1) k_ldn(int iters, int nwarps, long long* out, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + static_cast<uint32_t>((warp >> 2) * 128);
  {  // initialise this warp's 128 columns
    uint32_t z[32];
    for (int i = 0; i < 32; ++i) z[i] = 0x3f800000u + i;
    for (int c = 0; c < 128; c += 32) tmem_st_32x32b_x32(base + c, z);
    tmem_st_wait();
  }
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps) {
    for (int it = 0; it < iters; ++it) {
      uint32_t r[REPS][N];
#pragma unroll
      for (int q = 0; q < REPS; ++q) LdN<N>::ld(base + q * N, r[q]);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < REPS; ++q) {
        uint32_t x = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) x ^= r[q][i];
        acc += __uint_as_float(x);
      }
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}
template <int N, int REPS>
void run_ldn(long long* out, float* sink) {
  const int iters = 2000;
  for (int nw : {1, 4, 16}) {
    k_ldn<N, REPS><<<148, 512>>>(iters, nw, out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    const double per_it = double(h) / iters;
    printf("initialised TMEM, %d x ld.x%d + wait: %2d warps: %.1f clk/iter/warp -> %.1f B/clk/SM (%s)\n", REPS, N, nw, per_it,
           nw * REPS * N * 128.0 / per_it, cudaGetErrorString(e));
  }
}

int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&sink, 4);
  const int iters = 2000;
  for (int mode = 0; mode < 3; ++mode)
    for (int nw : {1, 4, 8, 16}) {
      k_tmem<<<148, 512>>>(iters, nw, mode, out, sink);
      cudaError_t e = cudaDeviceSynchronize();
      long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
      const double per_it = double(h) / iters;
      if (mode == 0) printf("tcgen05.ld 2 x (32 lanes x 32 cols): %2d warps: %.1f clk/iter/warp -> %.1f B/clk/SM  (%s)\n", nw, per_it, nw * 8192.0 / per_it, cudaGetErrorString(e));
      else if (mode == 2) printf("ex2.f16x2 x32 (=64 values)/thread: %2d warps: %.1f clk/iter -> %.2f ex2/clk/SM  (%s)\n", nw, per_it, nw * 64 * 32.0 / per_it, cudaGetErrorString(e));
      else printf("ex2 x64/thread: %2d warps: %.1f clk/iter -> %.2f ex2/clk/SM  (%s)\n", nw, per_it, nw * 64 * 32.0 / per_it, cudaGetErrorString(e));
    }
  run_ldn<8, 8>(out, sink);
  run_ldn<16, 4>(out, sink);
  run_ldn<16, 1>(out, sink);
  return 0;
}
