// Microbenchmark: back-to-back tcgen05.mma issue rate from resident shared-memory operands
// (no loads in the loop) -- cycles per MMA for kind::tf32 (K = 8) and kind::f16/bf16 (K = 16) at
// M = 128, N = 64 / 128 / 256, one CTA per SM on `grid` SMs.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o profiles/ubench/umma_rate profiles/ubench/umma_rate.cu
#include <cuda.h>
#include <stdio.h>
#include "../../tsdiff_b200/csrc/tc_common.cuh"
using namespace tc;

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

template <int KIND>  // 0 tf32, 1 bf16
__global__ void __launch_bounds__(128, 1) k_rate(int N, int iters, long long* out) {
  extern __shared__ uint8_t smem_dyn[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<float*>(smem_dyn + (base - smem_u32(smem_dyn)))[i] = 0.f;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    uint32_t idesc = KIND == 0 ? umma_idesc_tf32(N)
                               : ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24));
    const uint64_t adesc = umma_desc_sw128(base), bdesc = umma_desc_sw128(base + 16384);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (KIND == 0) umma_tf32(tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, 1u);
        else umma_bf16(tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, 1u);
      }
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
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

int main() {
  long long* out;
  cudaMalloc(&out, 16);
  const int iters = 256;  // x4 MMAs
  cudaFuncSetAttribute(k_rate<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
  cudaFuncSetAttribute(k_rate<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
  for (int grid : {1, 148})
    for (int kind = 0; kind < 2; ++kind)
      for (int N : {64, 128, 256}) {
        for (int rep = 0; rep < 2; ++rep) {
          if (kind == 0) k_rate<0><<<grid, 128, 50 * 1024>>>(N, iters, out);
          else k_rate<1><<<grid, 128, 50 * 1024>>>(N, iters, out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        }
        long long h[2];
        cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        printf("grid %3d kind %s N %3d : issue %.1f clk/MMA, complete %.1f clk/MMA (%.0f%% of %s peak at 1.9 GHz)\n", grid,
               kind ? "bf16(K16)" : "tf32(K8) ", N, (double)h[0] / (iters * 4), (double)h[1] / (iters * 4),
               100.0 * (2.0 * 128 * N * (kind ? 16 : 8)) / ((double)h[1] / (iters * 4)) / (kind ? 8192.0 : 4096.0), kind ? "bf16" : "tf32");
      }
  return 0;
}
