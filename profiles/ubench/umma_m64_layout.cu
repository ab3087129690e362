// Probe: where does tcgen05.mma cta_group::1 kind::tf32 with M = 64 put the rows of D in TMEM, and how
// fast does it issue?  A[i][k] = i + k/16 (K = 8), B[n][k] = delta(n, k) for n < 8  =>  D[i][n] = A[i][n].
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o profiles/ubench/umma_m64_layout profiles/ubench/umma_m64_layout.cu
#include <cuda.h>
#include <stdio.h>
#include "../../tsdiff_b200/csrc/tc_common.cuh"
using namespace tc;

__global__ void __launch_bounds__(128, 1) k_probe(int M, int N, int iters, float* out, long long* clk) {
  extern __shared__ uint8_t smem_dyn[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  uint8_t* gen = smem_dyn + (base - smem_u32(smem_dyn));
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<float*>(gen)[i] = 0.f;
  __syncthreads();
  // A panel: 128 rows x 32 floats (only k < 8 used), B panel at +16384: N rows x 32 floats
  for (int idx = threadIdx.x; idx < 128 * 8; idx += 128) {
    int r = idx / 8, k = idx % 8;
    *reinterpret_cast<float*>(gen + sw128_off(r, k / 4) + (k % 4) * 4) = (float)r + (float)k / 16.f;
  }
  for (int idx = threadIdx.x; idx < 8; idx += 128)
    *reinterpret_cast<float*>(gen + 16384 + sw128_off(idx, idx / 4) + (idx % 4) * 4) = 1.f;
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
  // clear the accumulator region we will read (all 128 lanes, 32 columns)
  {
    uint32_t z[32];
    for (int i = 0; i < 32; ++i) z[i] = __float_as_uint(-1.f);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                 ::"r"(tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16)), "r"(z[0]), "r"(z[1]), "r"(z[2]), "r"(z[3]), "r"(z[4]), "r"(z[5]), "r"(z[6]), "r"(z[7]),
                   "r"(z[8]), "r"(z[9]), "r"(z[10]), "r"(z[11]), "r"(z[12]), "r"(z[13]), "r"(z[14]), "r"(z[15]), "r"(z[16]), "r"(z[17]), "r"(z[18]), "r"(z[19]),
                   "r"(z[20]), "r"(z[21]), "r"(z[22]), "r"(z[23]), "r"(z[24]), "r"(z[25]), "r"(z[26]), "r"(z[27]), "r"(z[28]), "r"(z[29]), "r"(z[30]), "r"(z[31]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint64_t adesc = umma_desc_sw128(base), bdesc = umma_desc_sw128(base + 16384);
    long long t0 = clock64();
    umma_tf32(tmem, adesc, bdesc, idesc, 0u);
    for (int it = 1; it < iters; ++it) umma_tf32(tmem, adesc, bdesc, idesc, 0u);
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    clk[0] = clock64() - t0;
  }
  __syncthreads();
  tc_fence_after();
  uint32_t v[32];
  tmem_ld32(tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16), v);
  for (int j = 0; j < 32; ++j) out[threadIdx.x * 32 + j] = __uint_as_float(v[j]);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

int main() {
  float* out;
  long long* clk;
  cudaMalloc(&out, 128 * 32 * 4);
  cudaMalloc(&clk, 8);
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
  static float h[128 * 32];
  for (int M : {128, 64}) {
    k_probe<<<1, 128, 50 * 1024>>>(M, 256, 1, out, clk);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("M %d error %s\n", M, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("M = %d: TMEM lane -> (D row recovered from column 0; column 1 value)\n", M);
    for (int lane = 0; lane < 128; ++lane) {
      float c0 = h[lane * 32 + 0], c1 = h[lane * 32 + 1];
      if (lane % 8 == 0) printf("  lanes %3d..%3d:", lane, lane + 7);
      printf(" %6.2f/%5.3f", c0, c1 - c0);
      if (lane % 8 == 7) printf("\n");
    }
    long long hc;
    k_probe<<<1, 128, 50 * 1024>>>(M, 256, 1024, out, clk);
    cudaDeviceSynchronize();
    cudaMemcpy(&hc, clk, 8, cudaMemcpyDeviceToHost);
    printf("M = %d N = 256: %.1f clk per MMA (1024 back to back)\n", M, (double)hc / 1024);
  }
  return 0;
}
