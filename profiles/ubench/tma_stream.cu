// Probe: how fast can ONE SM stream a weight matrix L2 -> shared memory, and how does it scale with the number of
// CTAs that stream the SAME matrix at the same time?
//   mode 0: cp.async.bulk.tensor.2d, box {32 floats, 256 rows}, SWIZZLE_128B  (what the GEMM kernels did in round 1:
//           every 128-byte box row is its own L2 request, rows are 1 KiB apart)
//   mode 1: cp.async.bulk (1-D), 32 KiB contiguous chunks of a matrix PRE-PACKED panel by panel
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o profiles/ubench/tma_stream profiles/ubench/tma_stream.cu
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../tsdiff_b200/csrc/tc_common.cuh"
using namespace tc;

constexpr int PANEL = 32 * 1024;

__global__ void __launch_bounds__(128, 1) k_stream(const __grid_constant__ CUtensorMap map, const float* packed, int mode,
                                                   int slots, int panels_per_matrix, int repeats, long long* clk) {
  extern __shared__ uint8_t smem_dyn[];
  __shared__ uint64_t full[8];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  uint8_t* gen = smem_dyn + (base - smem_u32(smem_dyn));
  if (threadIdx.x == 0) {
    for (int s = 0; s < slots; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int total = panels_per_matrix * repeats;
    unsigned long long t0 = gtimer();
    // keep `slots` panels in flight; a panel is "consumed" as soon as it has landed
    for (int g = 0; g < total + slots; ++g) {
      if (g >= slots) mbar_wait(&full[g % slots], (uint32_t)(((g - slots) / slots) & 1));
      if (g < total) {
        const int s = g % slots, kb = g % panels_per_matrix;
        mbar_arrive_expect_tx(&full[s], PANEL);
        if (mode == 0) {
          tma_load_2d(gen + (size_t)s * PANEL, &map, &full[s], kb * 32, 0);
        } else {
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           smem_u32(gen + (size_t)s * PANEL)),
                       "l"(reinterpret_cast<uint64_t>(packed) + (uint64_t)kb * PANEL), "r"(PANEL), "r"(smem_u32(&full[s]))
                       : "memory");
        }
      }
    }
    clk[blockIdx.x] = (long long)(gtimer() - t0);
  }
}

int main() {
  const int H = 256;
  float *w, *packed;
  cudaMalloc(&w, H * H * 4);
  cudaMalloc(&packed, H * H * 4);
  cudaMemset(w, 0, H * H * 4);
  cudaMemset(packed, 0, H * H * 4);
  CUtensorMap map;
  if (!make_tensor_map(&map, w, H, H, H)) return 1;
  long long* clk;
  cudaMalloc(&clk, 148 * sizeof(long long));
  cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int clock_khz = 0;
  cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
  printf("mode slots ctas | MB per CTA | us | GB/s per SM | TB/s total   (clock %d MHz)\n", clock_khz / 1000);
  const int repeats = 16;
  for (int mode = 0; mode < 2; ++mode)
    for (int slots : {2, 4, 6}) {
      for (int ctas : {1, 14, 28, 56, 112, 148}) {
        for (int it = 0; it < 3; ++it) {
          k_stream<<<ctas, 128, (size_t)slots * PANEL + 1024>>>(map, packed, mode, slots, 8, repeats, clk);
          cudaDeviceSynchronize();
        }
        long long h[148];
        cudaMemcpy(h, clk, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
        const double us = (double)mx * 1e-3, mb = 8.0 * repeats * PANEL / 1e6;
        printf("%d %d %3d | %.2f | %7.2f | %7.1f | %6.2f\n", mode, slots, ctas, mb, us, mb * 1e3 / us, mb * ctas / us * 1e-3);
      }
    }
  cudaError_t e = cudaGetLastError();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
