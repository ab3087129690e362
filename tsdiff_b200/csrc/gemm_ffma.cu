// FP32 FFMA realisation of the fused linear-layer GEMM (gemm.cuh): the strict-parity path
// (north_star: "plain FFMA wherever tolerance demands it").  Classic shared-memory tiled
// SGEMM, register-staged double buffering, 256 threads, BK = 16; thread tiles are split in
// two half-tile groups so shared-memory float4 reads are conflict free.  Activation and
// epilogue flavour are template parameters (see the code-size note in gemm.cuh).
#include "gemm.cuh"

namespace {

constexpr int BK = 16;
constexpr int NTHREADS = 256;

template <int BM, int BN, int TM, int TN, int ACT, int EPI>
__global__ void __launch_bounds__(NTHREADS, 2) k_gemm_ffma(const GemmArgs p) {
  static_assert((BM / TM) * (BN / TN) == NTHREADS, "tile/thread mismatch");
  static_assert(TM == 4 || TM == 8, "TM");
  static_assert(TN == 4 || TN == 8, "TN");
  constexpr int LDA_S = BM + 4, LDB_S = BN + 4;
  constexpr int A_LD = (BM * 4 + NTHREADS - 1) / NTHREADS;  // float4 loads per thread per tile
  constexpr int B_LD = (BN * 4 + NTHREADS - 1) / NTHREADS;
  constexpr int TX = BN / TN;
  constexpr int GM = TM / 4, GN = TN / 4;  // float4 groups per thread

  __shared__ __align__(16) float As[2][BK][LDA_S];
  __shared__ __align__(16) float Bs[2][BK][LDB_S];

  const int M = p.M_ptr ? min(*p.M_ptr, p.M_cap) : p.M_cap;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  if (m0 >= M) return;
  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;

  float4 a_reg[A_LD], b_reg[B_LD];
  auto load_regs = [&](int k0) {
#pragma unroll
    for (int i = 0; i < A_LD; ++i) {
      int idx = tid + i * NTHREADS;
      if (BM * 4 >= NTHREADS * (i + 1) || idx < BM * 4) {
        int m = m0 + (idx >> 2), k = k0 + ((idx & 3) << 2);
        a_reg[i] = m < M ? tsd_load_a4(p, m, k) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int i = 0; i < B_LD; ++i) {
      int idx = tid + i * NTHREADS;
      if (BN * 4 >= NTHREADS * (i + 1) || idx < BN * 4) {
        int n = n0 + (idx >> 2), k = k0 + ((idx & 3) << 2);
        b_reg[i] = __ldg(reinterpret_cast<const float4*>(p.W + (size_t)n * p.K + k));
      }
    }
  };
  auto store_smem = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_LD; ++i) {
      int idx = tid + i * NTHREADS;
      if (BM * 4 >= NTHREADS * (i + 1) || idx < BM * 4) {
        int m = idx >> 2, kq = (idx & 3) << 2;
        As[buf][kq + 0][m] = a_reg[i].x;
        As[buf][kq + 1][m] = a_reg[i].y;
        As[buf][kq + 2][m] = a_reg[i].z;
        As[buf][kq + 3][m] = a_reg[i].w;
      }
    }
#pragma unroll
    for (int i = 0; i < B_LD; ++i) {
      int idx = tid + i * NTHREADS;
      if (BN * 4 >= NTHREADS * (i + 1) || idx < BN * 4) {
        int n = idx >> 2, kq = (idx & 3) << 2;
        Bs[buf][kq + 0][n] = b_reg[i].x;
        Bs[buf][kq + 1][n] = b_reg[i].y;
        Bs[buf][kq + 2][n] = b_reg[i].z;
        Bs[buf][kq + 3][n] = b_reg[i].w;
      }
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int num_tiles = p.K / BK;
  load_regs(0);
  store_smem(0);
  __syncthreads();
#pragma unroll 1
  for (int t = 0; t < num_tiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < num_tiles) load_regs((t + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int gm = 0; gm < GM; ++gm) {
        float4 v = *reinterpret_cast<const float4*>(&As[buf][kk][gm * (BM / 2) + ty * 4]);
        a[gm * 4 + 0] = v.x; a[gm * 4 + 1] = v.y; a[gm * 4 + 2] = v.z; a[gm * 4 + 3] = v.w;
      }
#pragma unroll
      for (int gn = 0; gn < GN; ++gn) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[buf][kk][gn * (BN / 2) + tx * 4]);
        b[gn * 4 + 0] = v.x; b[gn * 4 + 1] = v.y; b[gn * 4 + 2] = v.z; b[gn * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (t + 1 < num_tiles) store_smem(buf ^ 1);
    __syncthreads();
  }

  // ------------------------------------------------------------------ epilogue
  float4 bias4[GN];
#pragma unroll
  for (int gn = 0; gn < GN; ++gn)
    bias4[gn] = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n0 + gn * (BN / 2) + tx * 4))
                       : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + (i >> 2) * (BM / 2) + ty * 4 + (i & 3);
    const bool live = m < M;
    float cscale = 1.f;
    if (EPI == TSD_EPI_SCALE && live) cscale = tsd_cutoff_fn(p.scale_len[m], p.cutoff, p.smooth);
    const float* emb_row = (EPI == TSD_EPI_MULEMB && live) ? p.mul_emb + (size_t)(p.mul_code[m] & 0xffff) * p.N : nullptr;
    float dot = 0.f;
#pragma unroll
    for (int gn = 0; gn < GN; ++gn) {
      const int n = n0 + gn * (BN / 2) + tx * 4;
      float4 o;
      o.x = tsd_act_t<ACT>(acc[i][gn * 4 + 0] + bias4[gn].x);
      o.y = tsd_act_t<ACT>(acc[i][gn * 4 + 1] + bias4[gn].y);
      o.z = tsd_act_t<ACT>(acc[i][gn * 4 + 2] + bias4[gn].z);
      o.w = tsd_act_t<ACT>(acc[i][gn * 4 + 3] + bias4[gn].w);
      if (EPI == TSD_EPI_SCALE) {
        o.x *= cscale; o.y *= cscale; o.z *= cscale; o.w *= cscale;
      }
      if (EPI == TSD_EPI_MULEMB && live) {
        const float4 e = __ldg(reinterpret_cast<const float4*>(emb_row + n));
        o.x *= e.x; o.y *= e.y; o.z *= e.z; o.w *= e.w;
      }
      if (EPI == TSD_EPI_PLAIN && p.residual && live) {
        const float4 r = *reinterpret_cast<const float4*>(p.residual + (size_t)m * p.ldr + n);
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      if (EPI == TSD_EPI_DOT) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(p.w3 + n));
        if (live) {
          dot = fmaf(o.x, w.x, dot); dot = fmaf(o.y, w.y, dot); dot = fmaf(o.z, w.z, dot); dot = fmaf(o.w, w.w, dot);
        }
      } else if (live) {
        *reinterpret_cast<float4*>(p.C + (size_t)m * p.ldc + n) = o;
      }
    }
    if (EPI == TSD_EPI_DOT) {
      // the TX threads that share row m are consecutive lanes (TX is 16 or 32): butterfly sum
#pragma unroll
      for (int o = TX / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(TSD_FULL_MASK, dot, o);
      if (tx == 0 && live) {
        float r = dot + (p.b3 ? p.b3[0] : 0.f);
        p.out_vec[m] = p.accumulate ? p.out_vec[m] + r : r;
      }
    }
  }
}

template <int BM, int BN, int TM, int TN, int ACT, int EPI>
int launch(const GemmArgs& g, cudaStream_t stream) {
  dim3 grid(tsd_ceil_div(g.M_cap, BM), g.N / BN);
  k_gemm_ffma<BM, BN, TM, TN, ACT, EPI><<<grid, NTHREADS, 0, stream>>>(g);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

template <int BM, int BN, int TM, int TN, int EPI>
int launch_act(const GemmArgs& g, cudaStream_t stream) {
  switch (g.act) {
    case TSD_ACT_NONE: return launch<BM, BN, TM, TN, TSD_ACT_NONE, EPI>(g, stream);
    case TSD_ACT_RELU: return launch<BM, BN, TM, TN, TSD_ACT_RELU, EPI>(g, stream);
    case TSD_ACT_SWISH: return launch<BM, BN, TM, TN, TSD_ACT_SWISH, EPI>(g, stream);
    case TSD_ACT_SSP: return launch<BM, BN, TM, TN, TSD_ACT_SSP, EPI>(g, stream);
    default: return TSD_ERR_UNSUPPORTED;
  }
}

template <int BM, int BN, int TM, int TN>
int launch_epi(const GemmArgs& g, int epi, cudaStream_t stream) {
  switch (epi) {
    case TSD_EPI_PLAIN: return launch_act<BM, BN, TM, TN, TSD_EPI_PLAIN>(g, stream);
    case TSD_EPI_SCALE: return launch_act<BM, BN, TM, TN, TSD_EPI_SCALE>(g, stream);
    case TSD_EPI_MULEMB: return launch_act<BM, BN, TM, TN, TSD_EPI_MULEMB>(g, stream);
    default: return TSD_ERR_UNSUPPORTED;
  }
}

}  // namespace

int tsd_gemm_ffma(const GemmArgs& g, cudaStream_t stream) {
  TSD_REQUIRE(g.W && g.K % BK == 0 && g.K > 0 && g.N > 0 && g.M_cap >= 0);
  TSD_REQUIRE(g.a_kind == TSD_A_EDGE_MLP0 || g.A);
  const int epi = tsd_gemm_epi_kind(g);
  TSD_REQUIRE(epi >= 0);
  if (g.M_cap == 0) return TSD_OK;
  if (epi == TSD_EPI_DOT) {
    // the row-dot epilogue needs the whole output row inside one CTA
    if (g.N == 128) return launch_act<128, 128, 8, 8, TSD_EPI_DOT>(g, stream);
    if (g.N == 64) return launch_act<128, 64, 8, 4, TSD_EPI_DOT>(g, stream);
    return TSD_ERR_UNSUPPORTED;
  }
  TSD_REQUIRE(g.C);
  const bool small_m = g.M_cap <= 8192;  // node-level GEMMs: favour more CTAs over tile reuse
  if (g.N % 128 == 0)
    return small_m ? launch_epi<32, 128, 4, 4>(g, epi, stream) : launch_epi<128, 128, 8, 8>(g, epi, stream);
  if (g.N % 64 == 0)
    return small_m ? launch_epi<64, 64, 4, 4>(g, epi, stream) : launch_epi<128, 64, 8, 4>(g, epi, stream);
  return TSD_ERR_UNSUPPORTED;
}
