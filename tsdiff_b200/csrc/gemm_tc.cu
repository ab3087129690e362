// Tensor-core realisation of the fused linear-layer GEMM (gemm.cuh) for sm_100a:
// tcgen05.mma kind::tf32 (fp32 operands read as TF32, fp32 accumulation in TMEM), operands
// staged in shared memory by TMA.
//
//   CTA tile   : 128 rows x N (N = 64 / 128 / 256 = the whole output row, so the epilogue can
//                fuse bias / activation / cutoff / bond-embedding gate / row-dot), K streamed
//                in 32-float (128-byte) panels through a 2- or 4-stage shared-memory ring
//   CTA shapes : 256 threads, 2 stages, 2 CTAs/SM (more tiles than SMs)  |  512 threads, 4 stages,
//                one CTA per SM (grids that cannot fill the GPU: latency matters, not occupancy)
//   warp roles : warp 0  TMA producer  (one lane: cp.async.bulk.tensor of the W panel and, for
//                         a dense A, of the A panel; SWIZZLE_128B tensor maps write the UMMA
//                         canonical K-major layout directly; mbarrier complete_tx)
//                warp 1  MMA issuer    (one lane: 4 x tcgen05.mma M128 x N x K8 per panel,
//                         tcgen05.commit releases the stage)
//                warps 2..  A producers, two groups alternating panels (only when the A operand is
//                         COMPUTED by a prologue of
//                         gemm.cuh -- edge MLP layer 0, bond-embedding gating, pair products:
//                         values are written with the same swizzle, fence.proxy.async, arrive)
//                all warps epilogue    (tcgen05.ld of their TMEM lane quarter x column slice,
//                         epilogue in registers, shared-memory transpose, coalesced float4
//                         stores or row-dot)
//   pipelines  : full[s]  (TMA bytes + producer arrivals -> MMA), empty[s] (MMA commit ->
//                producers), accum (last commit -> epilogue)
#include <stdio.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace {
using namespace tc;

// 256-thread shape: 2 stages (~97 KB at N = 256) so TWO CTAs share an SM (2 x 256 TMEM columns);
// 512-thread shape (one CTA per SM): 4 stages, so four panels are in flight behind the TMA latency
__host__ __device__ constexpr int tc_stages(int nt) { return nt == 512 ? 4 : 2; }
// A-producer groups of the computed-operand kernels.  A group strides TC_GROUPS panels, and the
// parity wait on a stage's `empty` barrier is only unambiguous when the group cannot fall two
// phases behind, i.e. when TC_GROUPS <= number of stages: 2 stages -> 2 groups of three warps.
constexpr int TC_GROUPS = 2;
// Two CTA shapes: 256 threads (2 CTAs/SM: for batches with more tiles than SMs) and 512 threads (one
// CTA per SM: when the tiles do not fill the GPU anyway, 16 warps halve the latency-bound epilogue
// and 14 producer warps more than halve the time to compute an A panel).
__host__ __device__ constexpr int tc_group_threads(int nt) { return (nt - 64) / TC_GROUPS; }  // 96 or 224

// A-operand prologue with the fast activations of this arithmetic mode.  A producer thread
// always serves the same 16 rows (and the same 16-byte column chunk), so the per-row metadata
// (length / packed bond codes / row+col atom ids) is loaded ONCE before the K loop: the panel
// loads then have no dependent index load in front of them.
template <int AKIND>
__device__ __forceinline__ void tc_row_meta(const GemmArgs& p, int m, int& meta0, int& meta1, int& meta2) {
  meta0 = 0;
  meta1 = 0;
  meta2 = -1;
  if (AKIND == TSD_A_EDGE_MLP0) meta0 = __float_as_int(p.len[m]);
  if (AKIND == TSD_A_CAT) {
    meta1 = p.row_index ? p.row_index[m] : m;  // source row of d_emb / code (compact row lists)
    meta0 = p.code[meta1];
  }
  if (AKIND == TSD_A_PAIR) {
    meta0 = p.row[m];
    meta1 = p.col[m];
    if (p.alt_pos) meta2 = p.alt_pos[m];       // >= 0: the row's edge_attr lives in alt_A
  }
}

template <int AKIND>
__device__ __forceinline__ float4 tc_load_a4(const GemmArgs& p, int m, int k, int meta0, int meta1, int meta2) {
  if (AKIND == TSD_A_EDGE_MLP0) {
    const float l = __int_as_float(meta0);
    const float4 w = __ldg(reinterpret_cast<const float4*>(p.w0 + k));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.b0 + k));
    const float4 x = make_float4(fmaf(l, w.x, b.x), fmaf(l, w.y, b.y), fmaf(l, w.z, b.z), fmaf(l, w.w, b.w));
    switch (p.act0) {  // uniform
      case TSD_ACT_SWISH:
        return make_float4(tc_act<TSD_ACT_SWISH>(x.x), tc_act<TSD_ACT_SWISH>(x.y), tc_act<TSD_ACT_SWISH>(x.z),
                           tc_act<TSD_ACT_SWISH>(x.w));
      case TSD_ACT_RELU:
        return make_float4(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f), fmaxf(x.z, 0.f), fmaxf(x.w, 0.f));
      case TSD_ACT_SSP:
        return make_float4(tc_act<TSD_ACT_SSP>(x.x), tc_act<TSD_ACT_SSP>(x.y), tc_act<TSD_ACT_SSP>(x.z),
                           tc_act<TSD_ACT_SSP>(x.w));
      default:
        return x;
    }
  }
  if (AKIND == TSD_A_CAT) {
    const int hi = k >= p.H;
    const int kk = k - (hi ? p.H : 0);
    const int r = hi ? ((unsigned)meta0 >> 16) : (meta0 & 0xffff);
    const float4 d = *reinterpret_cast<const float4*>(p.A + (size_t)meta1 * p.lda + kk);
    const float4 e = __ldg(reinterpret_cast<const float4*>(p.emb + (size_t)r * p.H + kk));
    return make_float4(d.x * e.x, d.y * e.y, d.z * e.z, d.w * e.w);
  }
  if (AKIND == TSD_A_PAIR) {
    if (k < p.H) {
      const float4 a = *reinterpret_cast<const float4*>(p.h + (size_t)meta0 * p.H + k);
      const float4 b = *reinterpret_cast<const float4*>(p.h + (size_t)meta1 * p.H + k);
      return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
    }
    const float* src = meta2 >= 0 ? p.alt_A + (size_t)meta2 * p.lda : p.A + (size_t)m * p.lda;
    return *reinterpret_cast<const float4*>(src + (k - p.H));
  }
  return *reinterpret_cast<const float4*>(p.A + (size_t)m * p.lda + k);
}

// CHAIN: a second layer W2 (N2 x N) on the same tile.  After the first main loop every thread rewrites its part of the
// accumulator IN PLACE as tf32(act(acc + bias)) (tcgen05.ld / tcgen05.st), which the second main loop reads as its A
// operand straight from tensor memory; W2 streams through the same shared-memory ring (its first panels already while
// the first layer's last MMAs drain); the second accumulator sits in the TMEM columns behind the first.
template <int ACT, int EPI, int AKIND, int NT, bool CHAIN>
__global__ void __launch_bounds__(NT, NT == 256 ? 2 : 1)
    k_gemm_tf32(const GemmArgs p, int tmem_cols, const __grid_constant__ CUtensorMap map_a,
                const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_w2) {
  extern __shared__ uint8_t smem_dyn[];
  constexpr int TC_STAGES = tc_stages(NT);
  __shared__ uint64_t bar_full[TC_STAGES];
  __shared__ uint64_t bar_empty[TC_STAGES];
  __shared__ uint64_t bar_accum;
  __shared__ uint64_t bar_full2[TC_STAGES];  // CHAIN: W2 panels (TMA bytes only)
  __shared__ uint64_t bar_accum2;
  __shared__ uint32_t tmem_base_s;
  constexpr int TC_GROUP_THREADS = tc_group_threads(NT);
  constexpr int NSL = NT / 128;  // column slices of the epilogue (warps per TMEM lane quarter)
  __shared__ float s_dot[NSL - 1][TC_BM];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr bool dense_a = AKIND == TSD_A_PLAIN;
  // predecessor-independent setup first (see launch_pdl): barriers and descriptor prefetch
  if (tid == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&bar_full[s], dense_a ? 1u : 1u + TC_GROUP_THREADS);
      mbar_init(&bar_empty[s], 1);
    }
    mbar_init(&bar_accum, 1);
    if (CHAIN) {
      for (int s = 0; s < TC_STAGES; ++s) mbar_init(&bar_full2[s], 1);
      mbar_init(&bar_accum2, 1);
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w2)) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
    if (dense_a) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
  }
  pdl_wait();
  pdl_trigger();
  const int M = p.M_ptr ? min(*p.M_ptr, p.M_cap) : p.M_cap;
  const int m0 = blockIdx.x * TC_BM;
  if (m0 >= M) return;  // uniform across the CTA, before any allocation
  const int N = p.N, K = p.K;
  const uint32_t b_panel_bytes = (uint32_t)N * TC_BK * 4;
  const uint32_t stage_bytes = TC_A_PANEL_BYTES + b_panel_bytes;
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024 B alignment
  uint8_t* smem_gen = smem_dyn + (smem_base - smem_u32(smem_dyn));

  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"((uint32_t)tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const int num_kb = K / TC_BK;
  // CHAIN: the second layer's K panels (K2 = N) continue the ring's panel count; W2 panel j goes to slot (num_kb + j) %
  // stages once the MMAs that read the slot before have retired
  const int num_kb2 = CHAIN ? N / TC_BK : 0;
  const int N2 = CHAIN ? p.N2 : 0;
  auto load_w2_panel = [&](int j) {
    const int g = num_kb + j, s = g % TC_STAGES, round = g / TC_STAGES;
    if (round > 0) mbar_wait(&bar_empty[s], (uint32_t)((round - 1) & 1));
    mbar_arrive_expect_tx(&bar_full2[s], (uint32_t)N2 * TC_BK * 4);
    tma_load_2d(smem_gen + (size_t)s * stage_bytes + TC_A_PANEL_BYTES, &map_w2, &bar_full2[s], j * TC_BK, 0);
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const uint32_t tx = b_panel_bytes + (dense_a ? TC_A_PANEL_BYTES : 0);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % TC_STAGES, round = kb / TC_STAGES;
        if (round > 0) mbar_wait(&bar_empty[s], (uint32_t)((round - 1) & 1));
        uint8_t* a_panel = smem_gen + (size_t)s * stage_bytes;
        mbar_arrive_expect_tx(&bar_full[s], tx);
        if (dense_a) tma_load_2d(a_panel, &map_a, &bar_full[s], kb * TC_BK, m0);
        tma_load_2d(a_panel + TC_A_PANEL_BYTES, &map_w, &bar_full[s], kb * TC_BK, 0);
      }
      // the first W2 panels only wait for MMAs of the FIRST layer: issued before this warp joins its epilogue
      for (int j = 0; j < num_kb2 && j < TC_STAGES; ++j) load_w2_panel(j);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(N);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % TC_STAGES, round = kb / TC_STAGES;
        mbar_wait(&bar_full[s], (uint32_t)(round & 1));
        tc_fence_after();
        const uint64_t adesc = umma_desc_sw128(smem_base + (uint32_t)s * stage_bytes);
        const uint64_t bdesc = umma_desc_sw128(smem_base + (uint32_t)s * stage_bytes + TC_A_PANEL_BYTES);
#pragma unroll
        for (int kk = 0; kk < TC_BK / 8; ++kk)  // UMMA K = 8 tf32 = 32 bytes: start address += 2 (x16 B)
          umma_tf32(tmem, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc, (kb | kk) != 0 ? 1u : 0u);
        umma_commit(&bar_empty[s]);
        if (kb == num_kb - 1) umma_commit(&bar_accum);
      }
    }
  } else if (warp >= 2 && !dense_a) {
    // ------------------------------------------------------------ A producers (computed operand)
    // the groups take panels round-robin: the global-load / SFU latency of one panel (which the
    // proxy fence of its own threads cannot overlap) hides behind the other group's panel and the
    // co-resident CTA
    const int group = (warp - 2) / (TC_GROUP_THREADS / 32);
    const int t = tid - 64 - group * TC_GROUP_THREADS;
    constexpr int ROW_STEP = TC_GROUP_THREADS / 8;                              // rows between a thread's items
    constexpr int ITEMS = (TC_BM * 8 + TC_GROUP_THREADS - 1) / TC_GROUP_THREADS;  // float4 per thread per panel
    const int chunk = t & 7, row0 = t >> 3;  // item i -> row row0 + ROW_STEP i, 16-byte chunk `chunk`
    int meta0[ITEMS], meta1[ITEMS], meta2[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i)  // rows past M are clamped: they only feed output rows that are never stored
      tc_row_meta<AKIND>(p, min(m0 + min(row0 + ROW_STEP * i, TC_BM - 1), M - 1), meta0[i], meta1[i], meta2[i]);
    for (int kb = group; kb < num_kb; kb += TC_GROUPS) {
      const int s = kb % TC_STAGES, round = kb / TC_STAGES;
      if (round > 0) mbar_wait(&bar_empty[s], (uint32_t)((round - 1) & 1));
      uint8_t* a_panel = smem_gen + (size_t)s * stage_bytes;
      const int k = kb * TC_BK + (chunk << 2);
      float4 v[ITEMS];
#pragma unroll
      for (int i = 0; i < ITEMS; ++i)  // branch-free on purpose: a per-item `m < M` branch serialises the loads
        v[i] = tf32_rn4(
            tc_load_a4<AKIND>(p, min(m0 + min(row0 + ROW_STEP * i, TC_BM - 1), M - 1), k, meta0[i], meta1[i], meta2[i]));
#pragma unroll
      for (int i = 0; i < ITEMS; ++i)
        if ((TC_BM * 8) % TC_GROUP_THREADS == 0 || i < ITEMS - 1 || row0 + ROW_STEP * i < TC_BM)
          *reinterpret_cast<float4*>(a_panel + sw128_off(row0 + ROW_STEP * i, chunk)) = v[i];
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async proxy (UMMA)
      mbar_arrive(&bar_full[s]);
    }
  }
  __syncwarp();

  // ------------------------------------------------------------------ epilogue
  // tcgen05.ld hands every thread ONE output row (32 consecutive columns per load).  Storing
  // that straight to global memory makes each warp store touch 32 different 128-byte lines, so
  // each warp transposes its 32x32 chunk through the now idle pipeline shared memory and writes
  // 4 full 128-byte line segments per instruction.  (The activations are branch-free on
  // purpose: a `x > t ? x : f(x)` softplus compiled to one branch per element and serialised
  // the MUFU latency -- 2.1 us per 32-column chunk in the globaltimer timeline, now 0.4-1.0.)
  mbar_wait(&bar_accum, 0);
  tc_fence_after();
  const int q = warp & 3, half = warp >> 2;  // TMEM lane quarter of this warp, column half
  const int row = q * 32 + lane, m = m0 + row;
  const bool live = m < M;
  if (CHAIN) {
    // ---- first layer's epilogue, in place: X = tf32(act(acc + bias)) becomes the A operand of the second layer
    const int cols1 = N / NSL;
    for (int cc = 0; cc < cols1; cc += 32) {
      const int c0 = half * cols1 + cc;
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
      uint32_t v[32];
      tmem_ld32(taddr, v);
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + c0 + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[j + 0] = __float_as_uint(tf32_rn(tc_act<ACT>(__uint_as_float(v[j + 0]) + b.x)));
        v[j + 1] = __float_as_uint(tf32_rn(tc_act<ACT>(__uint_as_float(v[j + 1]) + b.y)));
        v[j + 2] = __float_as_uint(tf32_rn(tc_act<ACT>(__uint_as_float(v[j + 2]) + b.z)));
        v[j + 3] = __float_as_uint(tf32_rn(tc_act<ACT>(__uint_as_float(v[j + 3]) + b.w)));
      }
      tmem_st32(taddr, v);
    }
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0 && lane == 0) {
      for (int j = TC_STAGES; j < num_kb2; ++j) load_w2_panel(j);  // slots freed by the second layer's own MMAs
    } else if (warp == 1 && lane == 0) {
      const uint32_t idesc2 = umma_idesc_tf32(N2);
      for (int j = 0; j < num_kb2; ++j) {
        const int g = num_kb + j, s = g % TC_STAGES;
        mbar_wait(&bar_full2[s], (uint32_t)((j / TC_STAGES) & 1));
        tc_fence_after();
        const uint64_t bdesc = umma_desc_sw128(smem_base + (uint32_t)s * stage_bytes + TC_A_PANEL_BYTES);
#pragma unroll
        for (int kk = 0; kk < TC_BK / 8; ++kk)
          umma_tf32_ts(tmem + (uint32_t)N, tmem + (uint32_t)(j * TC_BK + kk * 8), bdesc + (uint64_t)(2 * kk), idesc2,
                       (j | kk) != 0 ? 1u : 0u);
        umma_commit(&bar_empty[s]);
        if (j == num_kb2 - 1) umma_commit(&bar_accum2);
      }
    }
    __syncwarp();
    mbar_wait(&bar_accum2, 0);
    tc_fence_after();
  }
  // the (last) layer's epilogue: N_out columns of the accumulator at TMEM column acc0
  const int N_out = CHAIN ? N2 : N;
  const uint32_t acc0 = CHAIN ? (uint32_t)N : 0u;
  const float* const bias_out = CHAIN ? p.bias2 : p.bias;
  constexpr int ACT_OUT = (CHAIN && EPI != TSD_EPI_DOT) ? TSD_ACT_NONE : ACT;
  const int cols_per_half = N_out / NSL;
  constexpr int TLD = 36;  // padded row stride (floats) of the transpose tile: conflict-free float4 access
  float* tile = reinterpret_cast<float*>(smem_gen) + warp * (32 * TLD);
  float cscale = 1.f;
  if (EPI == TSD_EPI_SCALE && live) cscale = tsd_cutoff_fn(p.scale_len[m], p.cutoff, p.smooth);
  const float* emb_row = nullptr;
  if (EPI == TSD_EPI_MULEMB) emb_row = p.mul_emb + (size_t)(live ? (p.mul_code[m] & 0xffff) : 0) * N_out;
  float dot = 0.f;
  for (int cc = 0; cc < cols_per_half; cc += 32) {
    const int c0 = half * cols_per_half + cc;
    uint32_t v[32];
    tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + acc0 + (uint32_t)c0, v);
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float4 b = bias_out ? __ldg(reinterpret_cast<const float4*>(bias_out + c0 + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 o;
      o.x = tc_act<ACT_OUT>(__uint_as_float(v[j + 0]) + b.x);
      o.y = tc_act<ACT_OUT>(__uint_as_float(v[j + 1]) + b.y);
      o.z = tc_act<ACT_OUT>(__uint_as_float(v[j + 2]) + b.z);
      o.w = tc_act<ACT_OUT>(__uint_as_float(v[j + 3]) + b.w);
      if (EPI == TSD_EPI_SCALE) {
        o.x *= cscale; o.y *= cscale; o.z *= cscale; o.w *= cscale;
      }
      if (EPI == TSD_EPI_MULEMB) {
        float4 e = __ldg(reinterpret_cast<const float4*>(emb_row + c0 + j));
        o.x *= e.x; o.y *= e.y; o.z *= e.z; o.w *= e.w;
      }
      if (EPI != TSD_EPI_DOT && p.round_out) o = tf32_rn4(o);
      if (EPI == TSD_EPI_DOT) {
        float4 w = __ldg(reinterpret_cast<const float4*>(p.w3 + c0 + j));
        dot = fmaf(o.x, w.x, dot); dot = fmaf(o.y, w.y, dot); dot = fmaf(o.z, w.z, dot); dot = fmaf(o.w, w.w, dot);
      } else {
        *reinterpret_cast<float4*>(tile + lane * TLD + j) = o;
      }
    }
    if (EPI != TSD_EPI_DOT) {
      __syncwarp();
      const int cg = (lane & 7) << 2;  // 8 lanes cover the 32 columns of one row: a full 128-byte line
#pragma unroll
      float4 xo[8], rs[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {  // loads first, branch-free (clamped row), so they are all in flight together
        const int r = i * 4 + (lane >> 3);
        xo[i] = *reinterpret_cast<const float4*>(tile + r * TLD + cg);
        rs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (EPI == TSD_EPI_PLAIN && p.residual)
          rs[i] = *reinterpret_cast<const float4*>(p.residual + (size_t)min(m0 + q * 32 + r, M - 1) * p.ldr + c0 + cg);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int mr = m0 + q * 32 + i * 4 + (lane >> 3);
        xo[i].x += rs[i].x; xo[i].y += rs[i].y; xo[i].z += rs[i].z; xo[i].w += rs[i].w;
        if (mr < M) *reinterpret_cast<float4*>(p.C + (size_t)mr * p.ldc + c0 + cg) = xo[i];
      }
      __syncwarp();
    }
  }
  if (EPI == TSD_EPI_DOT) {
    if (half > 0) s_dot[half - 1][row] = dot;
    __syncthreads();
    if (half == 0 && live) {
      float r = dot;
#pragma unroll
      for (int h = 0; h < NSL - 1; ++h) r += s_dot[h][row];
      r += (p.b3 ? p.b3[0] : 0.f);
      p.out_vec[m] = p.accumulate ? p.out_vec[m] + r : r;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)tmem_cols)
                 : "memory");
  }
}

template <int ACT, int EPI, int AKIND, int NT, bool CHAIN = false>
int tc_launch_nt(const GemmArgs& g0, int tmem_cols, size_t smem, const CUtensorMap& map_a, const CUtensorMap& map_w,
                 const CUtensorMap& map_w2, cudaStream_t stream) {
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    TSD_CUDA(cudaFuncSetAttribute(k_gemm_tf32<ACT, EPI, AKIND, NT, CHAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  200 * 1024));
    attr_set = true;
  }
  smem = (size_t)tc_stages(NT) * (TC_A_PANEL_BYTES + (size_t)g0.N * TC_BK * 4) + 1024;
  const size_t tiles = (size_t)(NT / 32) * 32 * 36 * sizeof(float) + 1024;  // epilogue transpose tiles reuse the pipeline
  if (smem < tiles) smem = tiles;
  TSD_CUDA(launch_pdl(k_gemm_tf32<ACT, EPI, AKIND, NT, CHAIN>, dim3(tsd_ceil_div(g0.M_cap, TC_BM)), dim3(NT), smem, stream,
                      g0, tmem_cols, map_a, map_w, map_w2));
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

// 512-thread CTAs when the tiles cannot fill the GPU at two per SM anyway (and every epilogue warp still
// gets a whole 32-column chunk)
template <int ACT, int EPI, int AKIND>
int tc_launch(const GemmArgs& g, int tmem_cols, size_t smem, const CUtensorMap& map_a, const CUtensorMap& map_w,
              cudaStream_t stream) {
  const bool wide = g.N >= 128 && tsd_ceil_div(g.M_cap, TC_BM) <= 148;
  if (wide) return tc_launch_nt<ACT, EPI, AKIND, 512>(g, tmem_cols, smem, map_a, map_w, map_w, stream);
  return tc_launch_nt<ACT, EPI, AKIND, 256>(g, tmem_cols, smem, map_a, map_w, map_w, stream);
}

template <int EPI, int AKIND>
int tc_dispatch_act(const GemmArgs& g, int tmem_cols, size_t smem, const CUtensorMap& map_a, const CUtensorMap& map_w,
                    cudaStream_t stream) {
  switch (g.act) {
    case TSD_ACT_NONE: return tc_launch<TSD_ACT_NONE, EPI, AKIND>(g, tmem_cols, smem, map_a, map_w, stream);
    case TSD_ACT_RELU: return tc_launch<TSD_ACT_RELU, EPI, AKIND>(g, tmem_cols, smem, map_a, map_w, stream);
    case TSD_ACT_SWISH: return tc_launch<TSD_ACT_SWISH, EPI, AKIND>(g, tmem_cols, smem, map_a, map_w, stream);
    case TSD_ACT_SSP: return tc_launch<TSD_ACT_SSP, EPI, AKIND>(g, tmem_cols, smem, map_a, map_w, stream);
    default: return TSD_ERR_UNSUPPORTED;
  }
}

}  // namespace

int tsd_gemm_tf32(const GemmArgs& g, cudaStream_t stream) {
  // shapes outside the tensor-core kernel's envelope go back to the FFMA kernel (tsd_gemm)
  if (!(g.N == 64 || g.N == 128 || g.N == 256) || g.K % TC_BK != 0 || g.K <= 0) return TSD_ERR_UNSUPPORTED;
  if (g.M_cap < 1024) return TSD_ERR_UNSUPPORTED;
  TSD_REQUIRE(g.W && (g.a_kind == TSD_A_EDGE_MLP0 || g.A) && (g.out_vec || g.C));
  const int epi = tsd_gemm_epi_kind(g);
  if (epi < 0) return TSD_ERR_UNSUPPORTED;
  if (g.a_kind == TSD_A_PLAIN && g.lda != g.K) return TSD_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(g.W) & 15) || (g.a_kind == TSD_A_PLAIN && (reinterpret_cast<uintptr_t>(g.A) & 15)))
    return TSD_ERR_UNSUPPORTED;  // TMA needs 16-byte aligned global addresses
  CUtensorMap map_a, map_w;
  if (!make_tensor_map(&map_w, g.W, (uint64_t)g.N, (uint64_t)g.K, (uint32_t)g.N)) return TSD_ERR_UNSUPPORTED;
  if (g.a_kind == TSD_A_PLAIN) {
    if (!make_tensor_map(&map_a, g.A, (uint64_t)g.M_cap, (uint64_t)g.K, TC_BM)) return TSD_ERR_UNSUPPORTED;
  } else {
    map_a = map_w;  // unused by the kernel
  }
  const int tmem_cols = g.N < 32 ? 32 : g.N;  // power of two >= 32
  size_t smem = 0;  // sized per CTA shape in tc_launch_nt
#define TC_GO(EPI, AKIND) return tc_dispatch_act<EPI, AKIND>(g, tmem_cols, smem, map_a, map_w, stream)
  switch (g.a_kind) {
    case TSD_A_PLAIN:
      if (epi == TSD_EPI_DOT) TC_GO(TSD_EPI_DOT, TSD_A_PLAIN);
      if (epi == TSD_EPI_SCALE) TC_GO(TSD_EPI_SCALE, TSD_A_PLAIN);
      if (epi == TSD_EPI_PLAIN) TC_GO(TSD_EPI_PLAIN, TSD_A_PLAIN);
      return TSD_ERR_UNSUPPORTED;
    case TSD_A_EDGE_MLP0:  // d_emb (cat path) or d_emb * bond_emb (no-cat path)
      if (epi == TSD_EPI_PLAIN) TC_GO(TSD_EPI_PLAIN, TSD_A_EDGE_MLP0);
      if (epi == TSD_EPI_MULEMB) TC_GO(TSD_EPI_MULEMB, TSD_A_EDGE_MLP0);
      return TSD_ERR_UNSUPPORTED;
    case TSD_A_CAT:
      if (epi == TSD_EPI_PLAIN) TC_GO(TSD_EPI_PLAIN, TSD_A_CAT);
      return TSD_ERR_UNSUPPORTED;
    case TSD_A_PAIR:
      if (epi == TSD_EPI_PLAIN) TC_GO(TSD_EPI_PLAIN, TSD_A_PAIR);
      return TSD_ERR_UNSUPPORTED;
    default:
      return TSD_ERR_UNSUPPORTED;
  }
#undef TC_GO
}

// Two chained layers on one tile (GemmArgs.W2): the computed-operand kernels of the edge embedding (cat0 -> cat2) and of
// the pair MLP (l0 -> l1 -> row-dot).  One 512-thread CTA per SM (the two accumulators take all 512 TMEM columns).
template <int EPI, int AKIND>
static int tc_chain_dispatch(const GemmArgs& g, const CUtensorMap& map_w, const CUtensorMap& map_w2, cudaStream_t stream) {
  switch (g.act) {
    case TSD_ACT_NONE: return tc_launch_nt<TSD_ACT_NONE, EPI, AKIND, 512, true>(g, 512, 0, map_w, map_w, map_w2, stream);
    case TSD_ACT_RELU: return tc_launch_nt<TSD_ACT_RELU, EPI, AKIND, 512, true>(g, 512, 0, map_w, map_w, map_w2, stream);
    case TSD_ACT_SWISH: return tc_launch_nt<TSD_ACT_SWISH, EPI, AKIND, 512, true>(g, 512, 0, map_w, map_w, map_w2, stream);
    case TSD_ACT_SSP: return tc_launch_nt<TSD_ACT_SSP, EPI, AKIND, 512, true>(g, 512, 0, map_w, map_w, map_w2, stream);
    default: return TSD_ERR_UNSUPPORTED;
  }
}

int tsd_gemm_chain2_tf32(const GemmArgs& g, cudaStream_t stream) {
  if (!g.W2 || g.N != 256 || !(g.N2 == 128 || g.N2 == 256) || g.K % TC_BK != 0 || g.K < 4 * TC_BK || g.M_cap < 1024)
    return TSD_ERR_UNSUPPORTED;
  if (!(g.a_kind == TSD_A_CAT || g.a_kind == TSD_A_PAIR) || g.residual) return TSD_ERR_UNSUPPORTED;
  TSD_REQUIRE(g.W && g.A && (g.out_vec || g.C));
  const int epi = tsd_gemm_epi_kind(g);
  if (!(epi == TSD_EPI_PLAIN || epi == TSD_EPI_DOT)) return TSD_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(g.W) | reinterpret_cast<uintptr_t>(g.W2)) & 15) return TSD_ERR_UNSUPPORTED;
  CUtensorMap map_w, map_w2;
  if (!make_tensor_map(&map_w, g.W, (uint64_t)g.N, (uint64_t)g.K, (uint32_t)g.N) ||
      !make_tensor_map(&map_w2, g.W2, (uint64_t)g.N2, (uint64_t)g.N, (uint32_t)g.N2))
    return TSD_ERR_UNSUPPORTED;
  if (g.a_kind == TSD_A_CAT) {
    if (epi == TSD_EPI_PLAIN) return tc_chain_dispatch<TSD_EPI_PLAIN, TSD_A_CAT>(g, map_w, map_w2, stream);
    return TSD_ERR_UNSUPPORTED;
  }
  if (epi == TSD_EPI_DOT) return tc_chain_dispatch<TSD_EPI_DOT, TSD_A_PAIR>(g, map_w, map_w2, stream);
  return TSD_ERR_UNSUPPORTED;
}

// elementwise round-to-nearest fp32 -> TF32 (kept in an fp32 container): weight shadow copies
__global__ void k_round_tf32(const float* __restrict__ src, float* __restrict__ dst, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(src[i]));
    dst[i] = __uint_as_float(r);
  }
}

extern "C" int tsd_round_tf32(const float* src, float* dst, int64_t n, tsd_stream_t stream) {
  TSD_REQUIRE(src && dst && n >= 0);
  if (n == 0) return TSD_OK;
  k_round_tf32<<<(unsigned)((n + 255) / 256), 256, 0, tsd_cu(stream)>>>(src, dst, (long long)n);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}
