// Probe for the swap-AB node kernel: tcgen05.mma kind::tf32 M = 128, K = 8 with SMALL N (16 / 32 / 64), both operands in
// shared memory -- clocks per MMA back to back, and with one tcgen05.commit per 8 MMAs (the per-K-panel pattern of
// the GEMM kernels; the commits go to a ring of mbarriers nobody waits on).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o profiles/ubench/umma_small_n profiles/ubench/umma_small_n.cu
#include <cuda.h>
#include <stdio.h>
#include "../../tsdiff_b200/csrc/tc_common.cuh"
using namespace tc;

__global__ void __launch_bounds__(128, 1) k_rate(int N, int iters, int commit_every, int two_acc, long long* out) {
  extern __shared__ uint8_t smem_dyn[];
  __shared__ uint64_t bar;
  __shared__ uint64_t ring[8];
  __shared__ uint32_t tmem_base_s;
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < (32768 + 32768) / 4; i += 128) reinterpret_cast<float*>(smem_dyn + (base - smem_u32(smem_dyn)))[i] = 0.f;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    for (int i = 0; i < 8; ++i) mbar_init(&ring[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_tf32(N);
    const uint64_t adesc = umma_desc_sw128(base), bdesc = umma_desc_sw128(base + 32768);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        umma_tf32(tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, 1u);
        if (two_acc) umma_tf32(tmem + 256, adesc + 1024 + 2 * kk, bdesc + 2 * kk, idesc, 1u);  // second M half: A rows 128..255
      }
      if (commit_every && (it % commit_every) == commit_every - 1) umma_commit(&ring[(it / commit_every) & 7]);
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long* out;
  cudaMalloc(&out, 16);
  const int iters = 256;
  cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  for (int two_acc = 0; two_acc < 2; ++two_acc)
    for (int commit_every : {0, 1, 2})
      for (int N : {16, 32, 64, 128, 256}) {
        for (int rep = 0; rep < 2; ++rep) {
          k_rate<<<1, 128, 70 * 1024>>>(N, iters, commit_every, two_acc, out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        }
        long long h[2];
        cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        const int mmas = iters * 4 * (two_acc ? 2 : 1);
        printf("M128 N %3d halves %d commit every %d x4 MMAs: issue %.1f clk/MMA, complete %.1f clk/MMA\n", N, two_acc + 1,
               commit_every, (double)h[0] / mmas, (double)h[1] / mmas);
      }
  return 0;
}
