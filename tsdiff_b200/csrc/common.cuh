// Shared device helpers for the tsdiff_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tsdiff_b200.h"

// compile-time switches of profiles/scripts/variants.py (A/B measurements behind DESIGN.md section 8)
#ifndef TSD_EXP_OLD_NODE
#define TSD_EXP_OLD_NODE 0     // 1: node side of the encoder as in round 1 (aggregation kernel + 128-row chained kernels)
#endif
#ifndef TSD_EXP_FILTER_POOL
#define TSD_EXP_FILTER_POOL 1  // one filter buffer per block: the filter kernels run ahead of the node-side chain
#endif
#define TSD_WARP 32
#define TSD_FULL_MASK 0xffffffffu

int tsd_record_cuda_error(cudaError_t e);  // api.cu

#define TSD_CUDA(expr)                                     \
  do {                                                     \
    cudaError_t _e = (expr);                               \
    if (_e != cudaSuccess) return tsd_record_cuda_error(_e); \
  } while (0)

void tsd_count_launch();  // api.cu
#define TSD_LAUNCH_CHECK()         \
  do {                             \
    tsd_count_launch();            \
    TSD_CUDA(cudaGetLastError());  \
  } while (0)

#define TSD_REQUIRE(cond)              \
  do {                                 \
    if (!(cond)) return TSD_ERR_INVALID; \
  } while (0)

static inline cudaStream_t tsd_cu(tsd_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int tsd_ceil_div(int a, int b) { return (a + b - 1) / b; }

// float32(log(2)) as torch.log(torch.tensor(2.)).item() gives it (schnet.py:68)
#define TSD_SSP_SHIFT 0.693147182464599609375f

__device__ __forceinline__ float tsd_softplus(float x) {
  // F.softplus(beta=1, threshold=20) in the branch-free stable form max(x,0) + log1p(exp(-|x|)):
  // identical to log1p(exp(x)) up to rounding, returns x exactly for x > 20 (the correction is
  // < 2^-29 x), and -- unlike `x > 20 ? x : ...` -- compiles without a per-element branch, so
  // the unrolled epilogues keep their instruction-level parallelism.
  return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
}

__device__ __forceinline__ float tsd_act(int act, float x) {
  switch (act) {
    case TSD_ACT_RELU: return fmaxf(x, 0.f);
    case TSD_ACT_SWISH: return x * (1.f / (1.f + expf(-x)));  // x * sigmoid(x)
    case TSD_ACT_SSP: return tsd_softplus(x) - TSD_SSP_SHIFT;
    case TSD_ACT_SOFTPLUS: return tsd_softplus(x);
    default: return x;
  }
}

// compile-time selected activation (accurate math): keeps unrolled epilogues small
template <int ACT>
__device__ __forceinline__ float tsd_act_t(float x) {
  if (ACT == TSD_ACT_RELU) return fmaxf(x, 0.f);
  if (ACT == TSD_ACT_SWISH) return x * (1.f / (1.f + expf(-x)));
  if (ACT == TSD_ACT_SSP) return tsd_softplus(x) - TSD_SSP_SHIFT;
  if (ACT == TSD_ACT_SOFTPLUS) return tsd_softplus(x);
  return x;
}

// Canonical fp32 squared distance: (dx*dx + dy*dy) + dz*dz, every op rounded (no FMA
// contraction) so the neighbour test is bit-identical to the oracle (third_party.pair_dist2).
__device__ __forceinline__ float tsd_dist2(float ax, float ay, float az, float bx, float by, float bz) {
  float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// SchNet cutoff envelope, schnet.py:91-97
__device__ __forceinline__ float tsd_cutoff_fn(float len, float cutoff, int smooth) {
  if (smooth) {
    float c = 0.5f * (cosf(len * 3.14159265358979323846f / cutoff) + 1.0f);
    return (len <= cutoff && len >= 0.f) ? c : 0.f;
  }
  return len <= cutoff ? 1.f : 0.f;
}

__device__ __forceinline__ unsigned tsd_lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}
