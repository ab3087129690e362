// Tensor-core realisation of the fused linear-layer GEMM (gemm.cuh) for sm_100a:
// tcgen05.mma kind::tf32 (fp32 operands read as TF32, fp32 accumulation in TMEM).
//
//   CTA tile   : 128 rows x N (N = 64 / 128 / 256 = the whole output row, so the epilogue
//                can fuse bias / activation / cutoff / row-dot), K streamed in 32-float
//                (128-byte) panels through a 3-stage shared-memory ring
//   operands   : A panel (128 x 32) is PRODUCED by the threads (the prologue of gemm.cuh:
//                RBF-free edge MLP layer 0, bond-embedding gating, pair products) and
//                written straight into the UMMA canonical K-major SWIZZLE_128B layout;
//                W panel (N x 32) is copied from the live nn.Linear weight the same way
//   MMA        : one elected thread issues 4 x (M128 x N x K8) tcgen05.mma per panel and
//                tcgen05.commit's the stage's "empty" mbarrier; accumulator = N TMEM columns
//   epilogue   : 8 warps tcgen05.ld their lane quarter (32 rows) x half of the columns,
//                apply the epilogue in registers and store / row-reduce
//
// The generic-proxy shared-memory writes are made visible to the tensor core (async proxy)
// with fence.proxy.async before the CTA barrier that precedes the MMA issue.
#include "gemm.cuh"

namespace {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;  // floats per panel row = 128 bytes = one swizzle atom row
constexpr int TC_THREADS = 256;
constexpr int TC_STAGES = 2;
constexpr int TC_A_PANEL_BYTES = TC_BM * TC_BK * 4;  // 16 KiB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// SM100 shared-memory matrix descriptor, K-major, SWIZZLE_128B: 8-row atoms of 128 B rows,
// atoms 1024 B apart (SBO), LBO unused (=1), descriptor version 1.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// kind::tf32 instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// byte offset of the 16-byte chunk `c` (4 floats) of row `r` inside a K-major SW128 panel
__device__ __forceinline__ uint32_t sw128_off(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}

// fast-math activations: this arithmetic mode already carries the TF32 bound, so the SFU
// approximations (__expf / __logf, ~2 ulp) are far below it
template <int ACT>
__device__ __forceinline__ float tc_act(float x) {
  if (ACT == TSD_ACT_RELU) return fmaxf(x, 0.f);
  if (ACT == TSD_ACT_SWISH) return __fdividef(x, 1.f + __expf(-x));
  if (ACT == TSD_ACT_SSP) return (x > 15.f ? x : __logf(1.f + __expf(x))) - TSD_SSP_SHIFT;
  if (ACT == TSD_ACT_SOFTPLUS) return x > 15.f ? x : __logf(1.f + __expf(x));
  return x;
}

enum { TC_EPI_PLAIN = 0, TC_EPI_SCALE = 1, TC_EPI_MULEMB = 2, TC_EPI_DOT = 3 };

// Software pipeline: two shared-memory stages + one panel of register prefetch.  Iteration kb
//   waits until the MMAs that read stage kb%2 retired, stores the prefetched panel, issues the
//   global loads of panel kb+1 (in flight across the CTA barrier and the MMA issue), then one
//   thread issues the 4 MMAs of panel kb.  ~98 KB smem / CTA -> two CTAs per SM hide each
//   other's load latency; 2 x 256 TMEM columns fill the SM's 512.
template <int ACT, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 2) k_gemm_tf32(const GemmArgs p, int tmem_cols) {
  extern __shared__ uint8_t smem_dyn[];
  __shared__ uint64_t bar_empty[TC_STAGES];
  __shared__ uint64_t bar_accum;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_dot[TC_BM];

  const int M = p.M_ptr ? min(*p.M_ptr, p.M_cap) : p.M_cap;
  const int m0 = blockIdx.x * TC_BM;
  if (m0 >= M) return;  // uniform across the CTA, before any allocation
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N, K = p.K;
  const uint32_t stage_bytes = TC_A_PANEL_BYTES + (uint32_t)N * TC_BK * 4;
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024 B alignment
  uint8_t* smem_gen = smem_dyn + (smem_base - smem_u32(smem_dyn));

  if (tid == 0) {
    for (int s = 0; s < TC_STAGES; ++s) mbar_init(&bar_empty[s], 1);
    mbar_init(&bar_accum, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"((uint32_t)tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = umma_idesc_tf32(N);

  constexpr int A_LD = (TC_BM * 8) / TC_THREADS;  // 4 float4 per thread per panel
  constexpr int B_LD = (256 * 8) / TC_THREADS;    // up to 8 (N = 256)
  float4 ra[A_LD], rb[B_LD];
  const int b_items = N * 8;
  auto load_regs = [&](int k0) {
#pragma unroll
    for (int i = 0; i < A_LD; ++i) {
      int idx = tid + i * TC_THREADS;
      int m = m0 + (idx >> 3);
      ra[i] = m < M ? tsd_load_a4(p, m, k0 + ((idx & 7) << 2)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < B_LD; ++i) {
      int idx = tid + i * TC_THREADS;
      if (idx < b_items)
        rb[i] = __ldg(reinterpret_cast<const float4*>(p.W + (size_t)(idx >> 3) * K + k0 + ((idx & 7) << 2)));
    }
  };
  auto store_smem = [&](uint8_t* a_panel, uint8_t* b_panel) {
#pragma unroll
    for (int i = 0; i < A_LD; ++i) {
      int idx = tid + i * TC_THREADS;
      *reinterpret_cast<float4*>(a_panel + sw128_off(idx >> 3, idx & 7)) = ra[i];
    }
#pragma unroll
    for (int i = 0; i < B_LD; ++i) {
      int idx = tid + i * TC_THREADS;
      if (idx < b_items) *reinterpret_cast<float4*>(b_panel + sw128_off(idx >> 3, idx & 7)) = rb[i];
    }
  };

  const int num_kb = K / TC_BK;
  load_regs(0);
  for (int kb = 0; kb < num_kb; ++kb) {
    const int s = kb % TC_STAGES, round = kb / TC_STAGES;
    if (round > 0) mbar_wait(&bar_empty[s], (uint32_t)((round - 1) & 1));  // MMAs that read this stage retired
    uint8_t* a_panel = smem_gen + (size_t)s * stage_bytes;
    store_smem(a_panel, a_panel + TC_A_PANEL_BYTES);
    if (kb + 1 < num_kb) load_regs((kb + 1) * TC_BK);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async proxy (UMMA)
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint64_t adesc = umma_desc_sw128(smem_base + (uint32_t)s * stage_bytes);
      const uint64_t bdesc = umma_desc_sw128(smem_base + (uint32_t)s * stage_bytes + TC_A_PANEL_BYTES);
#pragma unroll
      for (int kk = 0; kk < TC_BK / 8; ++kk)  // UMMA K = 8 tf32 = 32 bytes: advance the start address by 2 (x16 B)
        umma_tf32(tmem, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc, (kb | kk) != 0 ? 1u : 0u);
      umma_commit(&bar_empty[s]);
      if (kb == num_kb - 1) umma_commit(&bar_accum);
    }
  }

  // ------------------------------------------------------------------ epilogue
  mbar_wait(&bar_accum, 0);
  tc_fence_after();
  const int q = warp & 3, half = warp >> 2;  // TMEM lane quarter of this warp, column half
  const int row = q * 32 + lane, m = m0 + row;
  const bool live = m < M;
  const int cols_per_half = N >> 1;
  float cscale = 1.f;
  if (EPI == TC_EPI_SCALE && live) cscale = tsd_cutoff_fn(p.scale_len[m], p.cutoff, p.smooth);
  const float* emb_row = nullptr;
  if (EPI == TC_EPI_MULEMB && live) emb_row = p.mul_emb + (size_t)(p.mul_code[m] & 0xffff) * N;
  const float* res_row = (EPI == TC_EPI_PLAIN && p.residual && live) ? p.residual + (size_t)m * p.ldr : nullptr;
  float dot = 0.f;
  for (int cc = 0; cc < cols_per_half; cc += 32) {
    const int c0 = half * cols_per_half + cc;
    uint32_t v[32];
    tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
    if (live) {
      float* dst = (EPI == TC_EPI_DOT) ? nullptr : p.C + (size_t)m * p.ldc + c0;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 b = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + c0 + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 o;
        o.x = tc_act<ACT>(__uint_as_float(v[j + 0]) + b.x);
        o.y = tc_act<ACT>(__uint_as_float(v[j + 1]) + b.y);
        o.z = tc_act<ACT>(__uint_as_float(v[j + 2]) + b.z);
        o.w = tc_act<ACT>(__uint_as_float(v[j + 3]) + b.w);
        if (EPI == TC_EPI_SCALE) {
          o.x *= cscale; o.y *= cscale; o.z *= cscale; o.w *= cscale;
        }
        if (EPI == TC_EPI_MULEMB) {
          float4 e = __ldg(reinterpret_cast<const float4*>(emb_row + c0 + j));
          o.x *= e.x; o.y *= e.y; o.z *= e.z; o.w *= e.w;
        }
        if (EPI == TC_EPI_PLAIN && res_row) {
          float4 r = *reinterpret_cast<const float4*>(res_row + c0 + j);
          o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        if (EPI == TC_EPI_DOT) {
          float4 w = __ldg(reinterpret_cast<const float4*>(p.w3 + c0 + j));
          dot = fmaf(o.x, w.x, dot); dot = fmaf(o.y, w.y, dot); dot = fmaf(o.z, w.z, dot); dot = fmaf(o.w, w.w, dot);
        } else {
          *reinterpret_cast<float4*>(dst + j) = o;
        }
      }
    }
  }
  if (EPI == TC_EPI_DOT) {
    if (half == 1) s_dot[row] = dot;
    __syncthreads();
    if (half == 0 && live) {
      float r = (dot + s_dot[row]) + (p.b3 ? p.b3[0] : 0.f);
      p.out_vec[m] = p.accumulate ? p.out_vec[m] + r : r;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)tmem_cols)
                 : "memory");
  }
}

template <int ACT, int EPI>
int tc_launch(const GemmArgs& g, int tmem_cols, size_t smem, cudaStream_t stream) {
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    TSD_CUDA(cudaFuncSetAttribute(k_gemm_tf32<ACT, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr_set = true;
  }
  k_gemm_tf32<ACT, EPI><<<tsd_ceil_div(g.M_cap, TC_BM), TC_THREADS, smem, stream>>>(g, tmem_cols);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

template <int EPI>
int tc_dispatch_act(const GemmArgs& g, int tmem_cols, size_t smem, cudaStream_t stream) {
  switch (g.act) {
    case TSD_ACT_NONE: return tc_launch<TSD_ACT_NONE, EPI>(g, tmem_cols, smem, stream);
    case TSD_ACT_RELU: return tc_launch<TSD_ACT_RELU, EPI>(g, tmem_cols, smem, stream);
    case TSD_ACT_SWISH: return tc_launch<TSD_ACT_SWISH, EPI>(g, tmem_cols, smem, stream);
    case TSD_ACT_SSP: return tc_launch<TSD_ACT_SSP, EPI>(g, tmem_cols, smem, stream);
    default: return TSD_ERR_UNSUPPORTED;
  }
}

}  // namespace

int tsd_gemm_tf32(const GemmArgs& g, cudaStream_t stream) {
  // shapes outside the tensor-core kernel's envelope go back to the FFMA kernel (tsd_gemm)
  if (!(g.N == 64 || g.N == 128 || g.N == 256) || g.K % TC_BK != 0 || g.K <= 0) return TSD_ERR_UNSUPPORTED;
  if (g.M_cap < 1024) return TSD_ERR_UNSUPPORTED;
  TSD_REQUIRE(g.W && (g.a_kind == TSD_A_EDGE_MLP0 || g.A) && (g.out_vec || g.C));
  const int tmem_cols = g.N < 32 ? 32 : g.N;  // power of two >= 32
  const size_t smem = (size_t)TC_STAGES * (TC_A_PANEL_BYTES + (size_t)g.N * TC_BK * 4) + 1024;
  const int n_epi = (g.scale_len != nullptr) + (g.mul_emb != nullptr) + (g.out_vec != nullptr);
  if (n_epi > 1 || (n_epi == 1 && g.residual)) return TSD_ERR_UNSUPPORTED;
  if (g.out_vec) return tc_dispatch_act<TC_EPI_DOT>(g, tmem_cols, smem, stream);
  if (g.scale_len) return tc_dispatch_act<TC_EPI_SCALE>(g, tmem_cols, smem, stream);
  if (g.mul_emb) return tc_dispatch_act<TC_EPI_MULEMB>(g, tmem_cols, smem, stream);
  return tc_dispatch_act<TC_EPI_PLAIN>(g, tmem_cols, smem, stream);
}
