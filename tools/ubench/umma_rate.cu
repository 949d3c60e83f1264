// Microbenchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16, SS operands, 128B swizzle) for N = 64 / 128 / 256,
// issued back to back by one thread per CTA on every SM.  nvcc -gencode arch=compute_100a,code=sm_100a -o umma_rate umma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../videometamaterials_b200/csrc/sm100_ptx.cuh"
using namespace vmm;

__global__ void __launch_bounds__(128, 1) k(int N, int iters, int a_step, int b_step, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tb;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 1) { tmem_alloc(&tb, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tb;
  const uint32_t idesc = make_idesc_f16(128, N, 1, 0, 0);
  if (warp == 0) {
    const bool leader = elect_one();
    long long t0 = clock64();
    if (leader) {
      const uint32_t a0 = smem_u32(smem), b0 = a0 + 32 * 1024;
      for (int i = 0; i < iters; i += 4) {
        const uint64_t ad = make_smem_desc_sw128(a0 + ((i / 4) % 4) * a_step, 16, 1024);
        const uint64_t bd = make_smem_desc_sw128(b0 + ((i / 4) % 2) * b_step, 16, 1024);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_f16(tmem, ad + kk * 2, bd + kk * 2, idesc, 1);
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 8192;
  for (int N : {64, 128, 256}) {
    for (int astep : {0, 16384}) {
      k<<<148, 128, 98 * 1024>>>(N, iters, astep, astep ? N * 128 : 0, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0, mn = 1LL << 60; for (int i = 0; i < 148; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
      printf("N=%3d a_step=%5d: %s  cycles/MMA min %.1f max %.1f  -> %.0f FLOP/clk/SM\n", N, astep, cudaGetErrorString(e), double(mn) / iters, double(mx) / iters,
             2.0 * 128 * N * 16 / (double(mx) / iters));
    }
  }
  return 0;
}
