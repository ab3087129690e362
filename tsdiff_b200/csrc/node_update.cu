// Node side of a SchNet interaction block on the tensor cores, formulated for a SMALL number of
// rows (atoms): the segmented CFConv aggregation and the chained node linears in ONE kernel.
//
//   agg_i = sum_{j->i} x1_j * filt_ji                    (schnet.py:102-107, in-CSR, ascending source order)
//   y     = ssp(lin2(agg))                               (schnet.py:103-104 + :126)
//   h'    = h + lin(y)                                   (schnet.py:127-128)
//   x1'   = lin1_next(h')                                (the next block's schnet.py:101)
//
// At batch 100 there are ~1750 atoms: 14 row tiles of 128 for a conventional (rows = M) GEMM, i.e.
// 15 CTAs on 148 SMs, each streaming every 256 KiB weight matrix through its own L2 port while the
// separately launched aggregation kernel waits in front of it (27.5 us per block, the critical path
// of the whole Langevin step; profiles/r1_kineto_step_tf32.txt).  Here the GEMMs are transposed
// ("swap AB"):  Y^T (H x NT) = W (H x H) . X^T (H x NT)  with  A = W  (M = 128 output features per
// tcgen05.mma, K-major as nn.Linear stores it, streamed by TMA through a deep ring) and  B = NT atoms
// (N = 16 / 32 / 64), so a CTA owns only NT atoms, there are N/NT = 55..110 CTAs, an MMA costs NT/2
// clocks instead of 128, the accumulator of one stage is H x NT fp32 in TMEM (2 x NT columns), and
// every epilogue access is coalesced without a transpose: thread = one output feature, so for a
// fixed atom a warp touches 32 consecutive floats (bias = one register per thread).
//
//   warps 0..15  workers: (1) the B operand of stage 0 -- either the fused aggregation (one warp
//                per (atom, 128-channel slab); the tile's in-CSR segment is staged in shared memory
//                first so the row gathers have no dependent index load in front of them) or a dense
//                (N, H) input -- TF32-rounded (RNE) into UMMA K-major SWIZZLE_128B panels;
//                (2) the epilogues: tcgen05.ld of their TMEM lane quarter, bias / activation /
//                residual, coalesced global store and / or the next stage's B operand
//   warp 16      TMA producer: W panels (H rows x 32 floats) of ALL stages through the ring; it does
//                not depend on the atoms, so it starts streaming at kernel start, behind the aggregation
//   warp 17      MMA issuer (one lane): per K panel 4 x (H/128) tcgen05.mma M128 x NT x K8 kind::tf32
#include <string.h>

#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int NU_WORKERS = 16;
constexpr int NU_THREADS = (NU_WORKERS + 2) * 32;
constexpr int NU_MAX_SLOTS = 16;

struct NodeMaps {
  CUtensorMap w[3];
};

template <int CW>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&v)[CW]) {
  static_assert(CW == 8 || CW == 16 || CW == 32, "column count per warp");
  if constexpr (CW == 8) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
  } else if constexpr (CW == 16) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
  } else {
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
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void nu_fma_rn4(float4& acc, const float4& x, const float4& w) {
  acc.x = __fadd_rn(acc.x, __fmul_rn(x.x, w.x));
  acc.y = __fadd_rn(acc.y, __fmul_rn(x.y, w.y));
  acc.z = __fadd_rn(acc.z, __fmul_rn(x.z, w.z));
  acc.w = __fadd_rn(acc.w, __fmul_rn(x.w, w.w));
}

// agg of one (atom, slab) item: the in-edges [beg, end) are read from `eids` / `srcs` (shared-memory staged
// copies or the global arrays, both indexed by the global in-CSR position), sums in ascending source order
__device__ __forceinline__ float4 nu_aggregate_item(const NodeArgs& p, const int* eids, const int* srcs, int beg, int end,
                                                    int off, int lane) {
  constexpr int UNROLL = 8;
  const int H = p.H;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int base = beg; base < end; base += 32) {
    const int cnt = min(32, end - base);
    const int e_l = lane < cnt ? eids[base + lane] : 0;
    const int r_l = lane < cnt ? srcs[base + lane] : 0;
    int j = 0;
    for (; j + UNROLL <= cnt; j += UNROLL) {
      float4 w[UNROLL], x[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int e = __shfl_sync(TSD_FULL_MASK, e_l, j + u), r = __shfl_sync(TSD_FULL_MASK, r_l, j + u);
        w[u] = __ldg(reinterpret_cast<const float4*>(p.filt + (size_t)e * H + off));
        x[u] = __ldg(reinterpret_cast<const float4*>(p.x1 + (size_t)r * H + off));
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) nu_fma_rn4(acc, x[u], w[u]);
    }
    // tail: loads of the remaining (< UNROLL) edges issued together, summed in order
    if (j < cnt) {
      float4 w[UNROLL], x[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int jj = min(j + u, cnt - 1);
        const int e = __shfl_sync(TSD_FULL_MASK, e_l, jj), r = __shfl_sync(TSD_FULL_MASK, r_l, jj);
        w[u] = __ldg(reinterpret_cast<const float4*>(p.filt + (size_t)e * H + off));
        x[u] = __ldg(reinterpret_cast<const float4*>(p.x1 + (size_t)r * H + off));
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
        if (j + u < cnt) nu_fma_rn4(acc, x[u], w[u]);
    }
  }
  return acc;
}

template <int H, int NT>
__global__ void __launch_bounds__(NU_THREADS, 1) k_node_update(const NodeArgs p, const __grid_constant__ NodeMaps maps,
                                                               int num_slots, int tmem_cols) {
  constexpr int MH = H / 128;                 // M = 128 halves of the output features
  constexpr int NUM_KB = H / TC_BK;           // K panels per stage
  constexpr int W_PANEL = H * TC_BK * 4;      // bytes of one W panel (H rows x 128 B)
  constexpr int X_PANEL = NT * TC_BK * 4;     // bytes of one B-operand panel (NT rows x 128 B)
  constexpr int X_BYTES = NUM_KB * X_PANEL;   // one B operand: NT x H floats
  constexpr int CW = NT / 2;                  // accumulator columns per epilogue warp
  extern __shared__ uint8_t smem_dyn[];
  __shared__ uint64_t bar_full[NU_MAX_SLOTS];
  __shared__ uint64_t bar_empty[NU_MAX_SLOTS];
  __shared__ uint64_t bar_x[2];    // B operand buffer b written (all worker threads arrive)
  __shared__ uint64_t bar_acc[2];  // accumulator set b complete (tcgen05.commit)
  __shared__ uint32_t tmem_base_s;
  __shared__ int s_ptr[NT + 1];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int node0 = blockIdx.x * NT;
  const int N = p.num_nodes;
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_dyn + (smem_base - smem_u32(smem_dyn));
  uint8_t* xbuf[2] = {smem_gen, smem_gen + X_BYTES};
  uint8_t* ring = smem_gen + 2 * X_BYTES;
  const uint32_t ring_base = smem_base + 2 * X_BYTES;
  const int total_panels = p.num_stages * NUM_KB;

  if (tid == 0) {
    for (int s = 0; s < num_slots; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bar_x[b], NU_WORKERS * 32);
      mbar_init(&bar_acc[b], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == NU_WORKERS + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"((uint32_t)tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == NU_WORKERS) {
    // ------------------------------------------------------------------ TMA producer: the weights
    if (lane == 0) {
      for (int s = 0; s < p.num_stages; ++s)
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.w[s])) : "memory");
      for (int g = 0; g < total_panels; ++g) {
        const int slot = g % num_slots, round = g / num_slots;
        if (round > 0) mbar_wait(&bar_empty[slot], (uint32_t)((round - 1) & 1));
        mbar_arrive_expect_tx(&bar_full[slot], (uint32_t)W_PANEL);
        tma_load_2d(ring + (size_t)slot * W_PANEL, &maps.w[g / NUM_KB], &bar_full[slot], (g % NUM_KB) * TC_BK, 0);
      }
    }
  } else if (warp == NU_WORKERS + 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(NT);
      for (int s = 0; s < p.num_stages; ++s) {
        const int b = s & 1;
        mbar_wait(&bar_x[b], (uint32_t)((s >> 1) & 1));
        tc_fence_after();
        for (int kb = 0; kb < NUM_KB; ++kb) {
          const int g = s * NUM_KB + kb;
          const int slot = g % num_slots, round = g / num_slots;
          mbar_wait(&bar_full[slot], (uint32_t)(round & 1));
          tc_fence_after();
          const uint64_t bdesc = umma_desc_sw128(smem_base + (uint32_t)(b * X_BYTES + kb * X_PANEL));
#pragma unroll
          for (int half = 0; half < MH; ++half) {
            const uint64_t adesc = umma_desc_sw128(ring_base + (uint32_t)(slot * W_PANEL + half * TC_A_PANEL_BYTES));
            const uint32_t acc = tmem + (uint32_t)((b * MH + half) * NT);
#pragma unroll
            for (int kk = 0; kk < TC_BK / 8; ++kk)
              umma_tf32(acc, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc, (kb | kk) != 0 ? 1u : 0u);
          }
          umma_commit(&bar_empty[slot]);
        }
        umma_commit(&bar_acc[b]);
      }
    }
  } else {
    // ------------------------------------------------------------------ workers
    // (1) B operand of stage 0 -> xbuf[0]
    constexpr int SL = H / 128;          // 128-channel slabs per atom
    constexpr int ITEMS = NT * SL;
    if (p.x == nullptr) {
      // stage the tile's in-CSR segment (contiguous: the in-CSR is sorted by target atom) in xbuf[1], which
      // is idle until the first epilogue: the row gathers then have no dependent global index load in front
      constexpr int CAP = X_BYTES / 8;   // entries of each of the two staged id arrays
      int* s_eid = reinterpret_cast<int*>(xbuf[1]);
      int* s_src = s_eid + CAP;
      for (int i = tid; i <= NT; i += NU_WORKERS * 32) s_ptr[i] = p.in_ptr[min(node0 + i, N)];
      asm volatile("bar.sync 1, %0;" ::"r"(NU_WORKERS * 32) : "memory");
      const int seg0 = s_ptr[0], seg_n = s_ptr[NT] - seg0;
      const bool staged = seg_n <= CAP;
      if (staged) {
        for (int i = tid; i < seg_n; i += NU_WORKERS * 32) {
          s_eid[i] = p.in_eid[seg0 + i];
          s_src[i] = p.in_src[seg0 + i];
        }
        asm volatile("bar.sync 1, %0;" ::"r"(NU_WORKERS * 32) : "memory");
      }
      const int* eids = staged ? s_eid - seg0 : p.in_eid;
      const int* srcs = staged ? s_src - seg0 : p.in_src;
      for (int item = warp; item < ITEMS; item += NU_WORKERS) {
        const int n = item / SL, slab = item - n * SL;
        const int off = slab * 128 + lane * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (node0 + n < N) acc = nu_aggregate_item(p, eids, srcs, s_ptr[n], s_ptr[n + 1], off, lane);
        *reinterpret_cast<float4*>(xbuf[0] + (size_t)(off >> 5) * X_PANEL + sw128_off(n, (off & 31) >> 2)) = tf32_rn4(acc);
      }
      // every worker has finished READING the staged ids before any epilogue overwrites xbuf[1]
      asm volatile("bar.sync 1, %0;" ::"r"(NU_WORKERS * 32) : "memory");
    } else {
      for (int item = warp; item < ITEMS; item += NU_WORKERS) {
        const int n = item / SL, slab = item - n * SL;
        const int off = slab * 128 + lane * 4;
        const float4 v = __ldg(reinterpret_cast<const float4*>(p.x + (size_t)min(node0 + n, N - 1) * H + off));
        *reinterpret_cast<float4*>(xbuf[0] + (size_t)(off >> 5) * X_PANEL + sw128_off(n, (off & 31) >> 2)) = tf32_rn4(v);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async proxy (UMMA)
    mbar_arrive(&bar_x[0]);

    // (2) epilogues.  warp -> TMEM lane quarter q (hardware: warp_id % 4), output-feature half, column half
    const int q = warp & 3, half = (warp >> 2) & 1, cs = warp >> 3;
    const bool epi_warp = half < MH;  // H = 128: one M half, warps with half == 1 only take part in the barriers
    const int f = half * 128 + q * 32 + lane;
    const int n0 = cs * CW;
    for (int s = 0; s < p.num_stages; ++s) {
      const int b = s & 1;
      const NodeStage& st = p.st[s];
      const bool feeds = s + 1 < p.num_stages;
      float res[CW];
      float bias = 0.f;
      if (epi_warp) {
        // operands of the epilogue that do not depend on the accumulator: in flight behind the MMA
        if (st.bias) bias = __ldg(st.bias + f);
        if (st.residual) {
#pragma unroll
          for (int j = 0; j < CW; ++j) res[j] = st.residual[(size_t)min(node0 + n0 + j, N - 1) * H + f];
        }
      }
      mbar_wait(&bar_acc[b], (uint32_t)((s >> 1) & 1));
      tc_fence_after();
      if (epi_warp) {
        uint32_t v[CW];
        tmem_ld_cols<CW>(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((b * MH + half) * NT + n0), v);
        uint8_t* xn = xbuf[b ^ 1] + (size_t)(f >> 5) * X_PANEL + (size_t)((f & 3) << 2);
        const int chunk = (f & 31) >> 2;
#pragma unroll
        for (int j = 0; j < CW; ++j) {
          float r = __uint_as_float(v[j]) + bias;
          if (st.act == TSD_ACT_SSP) r = tc_act<TSD_ACT_SSP>(r);
          if (st.residual) r += res[j];
          const int n = n0 + j;
          if (st.store && node0 + n < N) st.store[(size_t)(node0 + n) * H + f] = r;
          if (feeds) *reinterpret_cast<float*>(xn + n * 128 + ((chunk ^ (n & 7)) << 4)) = tf32_rn(r);
        }
      }
      if (feeds) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        mbar_arrive(&bar_x[b ^ 1]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NU_WORKERS + 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)tmem_cols)
                 : "memory");
  }
}

template <int H, int NT>
int node_launch(const NodeArgs& a, const NodeMaps& maps, cudaStream_t stream) {
  constexpr int W_PANEL = H * TC_BK * 4, X_BYTES = NT * H * 4, NUM_KB = H / TC_BK;
  const int budget = 227 * 1024 - 2 * X_BYTES - 1024 - 2048;  // alignment slack + static shared memory
  int slots = budget / W_PANEL;
  if (slots > a.num_stages * NUM_KB) slots = a.num_stages * NUM_KB;
  if (slots > NU_MAX_SLOTS) slots = NU_MAX_SLOTS;
  if (slots < 2) return TSD_ERR_UNSUPPORTED;
  const size_t smem = (size_t)2 * X_BYTES + (size_t)slots * W_PANEL + 1024;
  static size_t attr_smem = 0;  // per instantiation
  if (smem > attr_smem) {
    TSD_CUDA(cudaFuncSetAttribute(k_node_update<H, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  int tmem_cols = 2 * (H / 128) * NT;
  if (tmem_cols < 32) tmem_cols = 32;
  k_node_update<H, NT><<<tsd_ceil_div(a.num_nodes, NT), NU_THREADS, smem, stream>>>(a, maps, slots, tmem_cols);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

}  // namespace

// Rows per CTA: enough CTAs to spread the aggregation's gathers over the GPU, few enough that the weight
// stream (every CTA reads every W once) stays small against them.
static int g_node_tile_override = 0;
// tuning hook of profiles/scripts (not part of the C-ABI header): 0 restores the built-in choice
extern "C" void tsd_tune_node_tile(int tile) { g_node_tile_override = tile; }

int tsd_node_tile(int num_nodes) {
  if (g_node_tile_override == 16 || g_node_tile_override == 32 || g_node_tile_override == 64) return g_node_tile_override;
  if (num_nodes <= 16 * 148) return 16;
  if (num_nodes <= 32 * 148) return 32;
  return 64;
}

int tsd_node_update_tf32(const NodeArgs& a, int tile, cudaStream_t stream) {
  using namespace tc;
  if (!(a.H == 128 || a.H == 256) || a.num_stages < 1 || a.num_stages > 3) return TSD_ERR_UNSUPPORTED;
  if (a.num_nodes <= 0) return TSD_OK;
  TSD_REQUIRE(a.x || (a.in_ptr && a.in_eid && a.in_src && a.x1 && a.filt));
  NodeMaps maps;
  for (int s = 0; s < 3; ++s) {
    const float* w = a.st[s < a.num_stages ? s : 0].W;
    TSD_REQUIRE(w);
    if (reinterpret_cast<uintptr_t>(w) & 15) return TSD_ERR_UNSUPPORTED;
    if (!make_tensor_map(&maps.w[s], w, (uint64_t)a.H, (uint64_t)a.H, (uint32_t)a.H)) return TSD_ERR_UNSUPPORTED;
  }
  if (a.H == 256) {
    if (tile == 16) return node_launch<256, 16>(a, maps, stream);
    if (tile == 32) return node_launch<256, 32>(a, maps, stream);
    return node_launch<256, 64>(a, maps, stream);
  }
  if (tile == 16) return node_launch<128, 16>(a, maps, stream);
  if (tile == 32) return node_launch<128, 32>(a, maps, stream);
  return node_launch<128, 64>(a, maps, stream);
}
