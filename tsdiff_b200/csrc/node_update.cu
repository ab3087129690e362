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
//
// Two kernels: k_node_update<H, NT, C> (one GEMM CTA per tile, optionally C - 1 helper CTAs that only aggregate) and,
// further down, k_node_pair<NT> (H = 256: both CTAs of a 2-CTA cluster run the GEMMs, each on half of the output
// features -- what the Langevin step uses).  node_chain.cu holds the persistent variant over all blocks (off by default).
#include <string.h>

#include "tc_common.cuh"

namespace {
using namespace tc;

// -DTSD_NODE_DBG: %globaltimer stamps of cluster 0's leader (profiles/scripts/node_timeline.py); off in the product build
#ifdef TSD_NODE_DBG
__device__ unsigned long long g_node_dbg[64];
__device__ unsigned long long g_node_cta[256 * 4];  // per CTA: start, aggregation done, end, in-edges
#define NU_STAMP(slot)                                                                         \
  do {                                                                                         \
    if (blockIdx.x == 0) g_node_dbg[slot] = gtimer();                                          \
    if ((slot) == 0 && blockIdx.x < 256) g_node_cta[blockIdx.x * 4 + 0] = gtimer();            \
    if ((slot) == 2 && blockIdx.x < 256) g_node_cta[blockIdx.x * 4 + 1] = gtimer();            \
    if ((slot) == 6 && blockIdx.x < 256) g_node_cta[blockIdx.x * 4 + 2] = gtimer();            \
  } while (0)
#else
#define NU_STAMP(slot) do {} while (0)
#endif

constexpr int NU_WORKERS = 16;
constexpr int NU_THREADS = (NU_WORKERS + 2) * 32;
constexpr int NU_MAX_SLOTS = 16;
constexpr int NU_SLOT_PANELS = 2;

struct NodeMaps {
  CUtensorMap w[3];
};

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

__device__ __forceinline__ uint32_t nu_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of `local` (a shared::cta address of THIS CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t nu_mapa(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void nu_st_cluster4(uint32_t addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void nu_arrive_cluster(uint32_t addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void nu_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}

// tuning hook of profiles/scripts (not part of the C-ABI header): programmatic dependent launch of the node chain
int g_node_pdl = 1;

constexpr int NU_STAGE_BYTES = 16 * 1024;  // staged in-CSR ids of a CTA's atoms: 2 x 2048 entries

// One cluster of C CTAs owns NT atoms.  Every CTA aggregates NT / C of them (the gathers want many SMs) and
// stores the TF32-rounded rows straight into the LEADER's B-operand buffer through distributed shared memory;
// then only the leader keeps its SM for the chained GEMMs (the weight stream wants few CTAs: every GEMM CTA
// reads every W once), the other CTAs exit and free their SMs for the edge-side kernels.
template <int H, int NT, int C>
__global__ void __launch_bounds__(NU_THREADS, 1) k_node_update(const NodeArgs p, const __grid_constant__ NodeMaps maps,
                                                               int num_slots, int tmem_cols) {
  constexpr int MH = H / 128;                 // M = 128 halves of the output features
  constexpr int NUM_KB = H / TC_BK;           // K panels per stage
  constexpr int W_PANEL = H * TC_BK * 4;      // bytes of one W panel (H rows x 128 B)
  constexpr int KP = NU_SLOT_PANELS;          // K panels per ring slot: one tcgen05.commit (~250 clk of tensor-pipe
  constexpr int W_SLOT = KP * W_PANEL;        // time when the MMAs are short) releases KP panels
  constexpr int NUM_KS = NUM_KB / KP;         // slots per stage
  constexpr int X_PANEL = NT * TC_BK * 4;     // bytes of one B-operand panel (NT rows x 128 B)
  constexpr int X_BYTES = NUM_KB * X_PANEL;   // the B operand: NT x H floats
  constexpr int CW = NT / 2;                  // accumulator columns per epilogue warp
  constexpr int CWC = CW % 32 == 0 ? 32 : (CW % 16 == 0 ? 16 : 8);  // ... processed in chunks of 32 / 16 / 8
  constexpr int NA = NT / C;                  // atoms aggregated by one CTA
  constexpr int NW = NU_WORKERS * 32;
  extern __shared__ uint8_t smem_dyn[];
  __shared__ uint64_t bar_full[NU_MAX_SLOTS];
  __shared__ uint64_t bar_empty[NU_MAX_SLOTS];
  __shared__ uint64_t bar_x0;      // stage-0 B operand complete: every worker thread of every CTA of the cluster arrives
  __shared__ uint64_t bar_xn;      // B operand of a later stage written by the leader's epilogue
  __shared__ uint64_t bar_acc[2];  // accumulator set b complete (tcgen05.commit)
  __shared__ uint32_t tmem_base_s;
  __shared__ int s_ptr[NA + 1];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) NU_STAMP(0);
  const uint32_t rank = C > 1 ? nu_cluster_rank() : 0u;
  const bool leader = rank == 0;
  // a cluster owns `npc` <= NT atoms (the MMA shape N = NT is padded): the host picks npc so that the node CTAs fit on the
  // SMs the concurrently running filter kernel leaves free, which matters more than an exactly filled tile
  const int npc = p.nodes_per_cluster > 0 ? min(p.nodes_per_cluster, NT) : NT;
  const int node0 = (blockIdx.x / C) * npc;  // first atom of the cluster's tile
  const int N = min(p.num_nodes, node0 + npc);  // atoms past the tile belong to the next cluster
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_dyn + (smem_base - smem_u32(smem_dyn));
  uint8_t* xbuf = smem_gen;                       // B operand (leader); rewritten in place by every epilogue
  uint8_t* stage_area = smem_gen + X_BYTES;       // this CTA's staged in-CSR ids
  uint8_t* ring = stage_area + NU_STAGE_BYTES;    // W panels (leader)
  const uint32_t ring_base = smem_base + X_BYTES + NU_STAGE_BYTES;
  const int total_slots = p.num_stages * NUM_KS;

  if (tid == 0) {
    for (int s = 0; s < num_slots; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    mbar_init(&bar_x0, C * NW);
    mbar_init(&bar_xn, NW);
    mbar_init(&bar_acc[0], 1);
    mbar_init(&bar_acc[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (leader && warp == NU_WORKERS + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"((uint32_t)tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (C > 1) {  // the leader's barriers exist before any CTA of the cluster arrives on them remotely
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) NU_STAMP(1);

  if (warp == NU_WORKERS) {
    // ------------------------------------------------------------------ TMA producer: the weights (leader)
    if (leader && lane == 0) {
      for (int s = 0; s < p.num_stages; ++s)
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.w[s])) : "memory");
      for (int g = 0; g < total_slots; ++g) {
        const int slot = g % num_slots, round = g / num_slots;
        if (round > 0) mbar_wait(&bar_empty[slot], (uint32_t)((round - 1) & 1));
        mbar_arrive_expect_tx(&bar_full[slot], (uint32_t)W_SLOT);
#pragma unroll
        for (int k = 0; k < KP; ++k)
          tma_load_2d(ring + (size_t)slot * W_SLOT + (size_t)k * W_PANEL, &maps.w[g / NUM_KS], &bar_full[slot],
                      ((g % NUM_KS) * KP + k) * TC_BK, 0);
      }
    }
  } else if (warp == NU_WORKERS + 1) {
    // ------------------------------------------------------------------ MMA issuer (leader)
    if (leader && lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(NT);
      for (int s = 0; s < p.num_stages; ++s) {
        const int b = s & 1;
        if (s == 0) nu_wait_cluster(&bar_x0, 0);
        else mbar_wait(&bar_xn, (uint32_t)((s - 1) & 1));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // (remote) generic-proxy writes -> UMMA reads
        tc_fence_after();
        NU_STAMP(8 + 4 * s);
        for (int ks = 0; ks < NUM_KS; ++ks) {
          const int g = s * NUM_KS + ks;
          const int slot = g % num_slots, round = g / num_slots;
          mbar_wait(&bar_full[slot], (uint32_t)(round & 1));
          tc_fence_after();
          if (ks == 0) NU_STAMP(9 + 4 * s);
#pragma unroll
          for (int k = 0; k < KP; ++k) {
            const int kb = ks * KP + k;
            const uint64_t bdesc = umma_desc_sw128(smem_base + (uint32_t)(kb * X_PANEL));
            // consecutive MMAs alternate between the accumulators of the two M halves (MMAs into the same
            // accumulator serialise: 94 clk each instead of 47 at small N; profiles/r2_umma_small_n.txt)
#pragma unroll
            for (int kk = 0; kk < TC_BK / 8; ++kk) {
#pragma unroll
              for (int half = 0; half < MH; ++half) {
                const uint64_t adesc =
                    umma_desc_sw128(ring_base + (uint32_t)(slot * W_SLOT + k * W_PANEL + half * TC_A_PANEL_BYTES));
                umma_tf32(tmem + (uint32_t)((b * MH + half) * NT), adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk),
                          idesc, (kb | kk) != 0 ? 1u : 0u);
              }
            }
          }
          umma_commit(&bar_empty[slot]);
        }
        umma_commit(&bar_acc[b]);
        NU_STAMP(10 + 4 * s);
      }
    }
  } else {
    // ------------------------------------------------------------------ workers
    // Programmatic dependent launch: everything above (barriers, TMEM, the weight stream -- weights are written by no
    // kernel of the step) ran while the previous node kernel was still finishing; the atoms' data is read from here on.
    pdl_wait();
    pdl_trigger();
    // (1) this CTA's NA rows of the stage-0 B operand -> the leader's xbuf
    constexpr int SL = H / 128;          // 128-channel slabs per atom
    constexpr int ITEMS = NA * SL;
    const int my0 = (int)rank * NA;      // first row (within the tile) of this CTA
    const uint32_t xdst = nu_mapa(smem_base, 0);  // the leader's xbuf (same offset in every CTA)
    if (p.x == nullptr) {
      // stage the in-CSR segment of this CTA's atoms (contiguous: the in-CSR is sorted by target atom) in shared
      // memory: the row gathers then have no dependent global index load in front of them
      constexpr int CAP = NU_STAGE_BYTES / 8;
      int* s_eid = reinterpret_cast<int*>(stage_area);
      int* s_src = s_eid + CAP;
      for (int i = tid; i <= NA; i += NW) s_ptr[i] = p.in_ptr[min(node0 + my0 + i, N)];
      asm volatile("bar.sync 1, %0;" ::"r"(NW) : "memory");
      const int seg0 = s_ptr[0], seg_n = s_ptr[NA] - seg0;
#ifdef TSD_NODE_DBG
      if (tid == 0 && blockIdx.x < 256) g_node_cta[blockIdx.x * 4 + 3] = (unsigned long long)seg_n;
#endif
      const bool staged = seg_n <= CAP;
      if (staged) {
        for (int i = tid; i < seg_n; i += NW) {
          s_eid[i] = p.in_eid[seg0 + i];
          s_src[i] = p.in_src[seg0 + i];
        }
        asm volatile("bar.sync 1, %0;" ::"r"(NW) : "memory");
      }
      const int* eids = staged ? s_eid - seg0 : p.in_eid;
      const int* srcs = staged ? s_src - seg0 : p.in_src;
      for (int item = warp; item < ITEMS; item += NU_WORKERS) {
        const int n = item / SL, slab = item - n * SL;
        const int off = slab * 128 + lane * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);  // rows past the tile: zeros (their output columns are never stored)
        if (node0 + my0 + n < N) acc = nu_aggregate_item(p, eids, srcs, s_ptr[n], s_ptr[n + 1], off, lane);
        nu_st_cluster4(xdst + (uint32_t)((off >> 5) * X_PANEL) + sw128_off(my0 + n, (off & 31) >> 2), tf32_rn4(acc));
      }
    } else {
      for (int item = warp; item < ITEMS; item += NU_WORKERS) {
        const int n = item / SL, slab = item - n * SL;
        const int off = slab * 128 + lane * 4;
        const float4 v = __ldg(reinterpret_cast<const float4*>(p.x + (size_t)min(node0 + my0 + n, N - 1) * H + off));
        nu_st_cluster4(xdst + (uint32_t)((off >> 5) * X_PANEL) + sw128_off(my0 + n, (off & 31) >> 2), tf32_rn4(v));
      }
    }
    asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy writes -> async proxy (the leader's UMMA)
    nu_arrive_cluster(nu_mapa(smem_u32(&bar_x0), 0));
    if (tid == 0) NU_STAMP(2);

    if (leader) {
      // (2) epilogues.  warp -> TMEM lane quarter q (hardware: warp_id % 4), output-feature half, column half
      const int q = warp & 3, half = (warp >> 2) & 1, cs = warp >> 3;
      const bool epi_warp = half < MH;  // H = 128: one M half, warps with half == 1 only take part in the barriers
      const int f = half * 128 + q * 32 + lane;
      for (int s = 0; s < p.num_stages; ++s) {
        const int b = s & 1;
        // the stage description in REGISTERS: indexing p.st[] at run time puts it in local memory, and every
        // generic-address store below would force a reload of each field (measured: ~200 clk per element)
        const float* const st_bias = p.st[s].bias;
        const float* const st_res = p.st[s].residual;
        float* const st_store = p.st[s].store;
        const bool st_ssp = p.st[s].act == TSD_ACT_SSP;
        const bool feeds = s + 1 < p.num_stages;
        float bias = 0.f;
        if (epi_warp && st_bias) bias = __ldg(st_bias + f);
        bool waited = false;
#pragma unroll 1
        for (int cc = 0; cc < CW; cc += CWC) {
          const int n0 = cs * CW + cc;
          if (waited && node0 + n0 >= N) break;  // columns of atoms past the tile (padding of the MMA shape)
          float res[CWC];
          if (epi_warp && st_res) {  // independent of the accumulator: in flight behind the MMA
#pragma unroll
            for (int j = 0; j < CWC; ++j) res[j] = st_res[(size_t)min(node0 + n0 + j, N - 1) * H + f];
          }
          if (!waited) {
            mbar_wait(&bar_acc[b], (uint32_t)((s >> 1) & 1));
            tc_fence_after();
            waited = true;
            if (tid == 0) NU_STAMP(11 + 4 * s);
          }
          if (epi_warp) {
            uint32_t v[CWC];
            tmem_ld_cols<CWC>(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((b * MH + half) * NT + n0), v);
            uint8_t* xn = xbuf + (size_t)(f >> 5) * X_PANEL + (size_t)((f & 3) << 2);
            const int chunk = (f & 31) >> 2;
#pragma unroll
            for (int j = 0; j < CWC; ++j) {
              float r = __uint_as_float(v[j]) + bias;
              if (st_ssp) r = tc_act<TSD_ACT_SSP>(r);
              if (st_res) r += res[j];
              const int n = n0 + j;
              if (st_store && node0 + n < N) st_store[(size_t)(node0 + n) * H + f] = r;
              if (feeds) *reinterpret_cast<float*>(xn + n * 128 + ((chunk ^ (n & 7)) << 4)) = tf32_rn(r);
            }
          }
        }
        if (feeds) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          tc_fence_before();
          mbar_arrive(&bar_xn);
        }
        if (tid == 0) NU_STAMP(3 + s);
      }
    }
  }
  if (!leader) return;  // the other CTAs of the cluster only aggregate
  tc_fence_before();
  __syncthreads();
  if (tid == 0) NU_STAMP(6);
  if (warp == NU_WORKERS + 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)tmem_cols)
                 : "memory");
  }
}

template <int H, int NT, int C>
int node_launch(const NodeArgs& a, const NodeMaps& maps, cudaStream_t stream) {
  constexpr int W_SLOT = NU_SLOT_PANELS * H * TC_BK * 4, X_BYTES = NT * H * 4, NUM_KS = H / TC_BK / NU_SLOT_PANELS;
  const int budget = 227 * 1024 - X_BYTES - NU_STAGE_BYTES - 1024 - 2048;  // alignment slack + static shared memory
  int slots = budget / W_SLOT;
  if (slots > a.num_stages * NUM_KS) slots = a.num_stages * NUM_KS;
  if (slots > NU_MAX_SLOTS) slots = NU_MAX_SLOTS;
  if (slots < 2) return TSD_ERR_UNSUPPORTED;
  const size_t smem = (size_t)X_BYTES + NU_STAGE_BYTES + (size_t)slots * W_SLOT + 1024;
  static size_t attr_smem = 0;  // per instantiation
  if (smem > attr_smem) {
    TSD_CUDA(cudaFuncSetAttribute(k_node_update<H, NT, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  int tmem_cols = 32;  // a power of two >= 32 that holds 2 accumulator sets x H/128 halves x NT columns
  while (tmem_cols < 2 * (H / 128) * NT) tmem_cols *= 2;
  cudaLaunchConfig_t cfg = {};
  const int npc = a.nodes_per_cluster > 0 && a.nodes_per_cluster < NT ? a.nodes_per_cluster : NT;
  cfg.gridDim = dim3(tsd_ceil_div(a.num_nodes, npc) * C);
  cfg.blockDim = dim3(NU_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_node_pdl ? 2 : 1;
  TSD_CUDA(cudaLaunchKernelEx(&cfg, k_node_update<H, NT, C>, a, maps, slots, tmem_cols));
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

__device__ __forceinline__ void nu_st_cluster1(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// ---- the pair variant (H = 256): BOTH CTAs of a cluster run the chained GEMMs, each on 128 of the 256 output features.
// The single-leader kernel above streams every 256 KiB weight matrix through one SM's L2 port (127 GB/s: 2.1 us per
// stage, three stages per block, on the critical path of the step).  Here a CTA streams only ITS half of every W
// (128 rows: 1 us per stage), aggregates half of the tile's atoms, and after every stage stores its 128 features of the
// NT atoms TF32-rounded into BOTH CTAs' B-operand buffers through distributed shared memory (16 KiB per stage).
//   * the B operand ping-pongs between two buffers by stage parity: a CTA's epilogue of stage s writes the peer's input
//     of stage s + 1 while the peer's MMAs may still be reading its input of stage s;
//   * one M half per CTA means consecutive MMAs would hit the same accumulator and serialise (94 instead of 47 clk each,
//     profiles/r2_umma_small_n.txt): the K steps alternate between two accumulators, summed in the epilogue;
//   * the in-CSR ids of the CTA's atoms are staged BEFORE griddepcontrol.wait: they are written by the edge build at the
//     start of the step, only x1 comes from the preceding node kernel.  (Also staging the x1 rows of the tile's reactions
//     -- a contiguous range -- in shared memory with one bulk copy and gathering only the filter rows measured SLOWER,
//     289 vs 279 us per step: the phase is bound by dependent latencies, not bytes; profiles/r3_node_pair_timeline.txt);
//   * no CTA touches the peer's shared memory after the peer's last wait (its bar_x of the final stage), so no
//     cluster barrier is needed before exit.
template <int NT>
__global__ void __launch_bounds__(NU_THREADS, 1) k_node_pair(const NodeArgs p, const __grid_constant__ NodeMaps maps,
                                                             int tmem_cols) {
  constexpr int H = 256;
  constexpr int NUM_KB = H / TC_BK;            // K panels per stage
  constexpr int W_PANEL = 128 * TC_BK * 4;     // this CTA's 128 rows of one K panel
  constexpr int KP = 4;                        // K panels per ring slot (one tcgen05.commit per 16 MMAs)
  constexpr int W_SLOT = KP * W_PANEL;         // 64 KiB
  constexpr int NUM_KS = NUM_KB / KP;          // slots per stage
  constexpr int NSLOT = 2;                     // = one stage's half matrix in flight
  constexpr int X_PANEL = NT * TC_BK * 4;
  constexpr int X_BYTES = NUM_KB * X_PANEL;    // one B operand: NT x H floats
  constexpr int NA = NT / 2;                   // atoms aggregated by one CTA
  constexpr int NW = NU_WORKERS * 32;
  constexpr int CW = NT / 4;                   // accumulator columns (atoms) per epilogue warp
  static_assert(CW == 8, "NT must be 32 (two NT x 256 B operands + the weight ring fill the shared memory)");
  extern __shared__ uint8_t smem_dyn[];
  __shared__ uint64_t bar_full[NSLOT];
  __shared__ uint64_t bar_empty[NSLOT];
  __shared__ uint64_t bar_x[3];    // B operand of stage s complete: one arrival per CTA, after its workers' barrier
  __shared__ uint64_t bar_acc[2];  // accumulator set b complete (tcgen05.commit)
  __shared__ uint32_t tmem_base_s;
  __shared__ int s_ptr[NA + 1];   // staged (local) positions of this CTA's atoms' in-edges
  __shared__ int s_gbeg[NA];      // ... and where they start in the global in-CSR

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) NU_STAMP(0);
  const uint32_t rank = nu_cluster_rank();
  // Row n of cluster c's tile is atom c + n * (number of clusters): a tile of CONSECUTIVE atoms lies inside one or two
  // reactions, whose in-degree (n_g - 1 when every pair is inside the cutoff) varies 9 .. 24 at batch 100, so the CTAs'
  // gather volumes differed by a factor of two and every block waited for the slowest one (8 of 24 us in the in-kernel
  // timeline profiles/r3_node_chain_timeline.txt); strided rows sample ~14 reactions per CTA.
  const int npc = p.nodes_per_cluster > 0 ? min(p.nodes_per_cluster, NT) : NT;
  const int cluster = blockIdx.x / 2, T = gridDim.x / 2;
  const int N = p.num_nodes;
  auto atom_of = [&](int n) { return cluster + n * T; };  // valid while n < npc and the result is < N
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_dyn + (smem_base - smem_u32(smem_dyn));
  // layout: B operand 0 | staged in-CSR ids | ring slot 0 | ring slot 1 | B operand 1
  uint8_t* stage_area = smem_gen + X_BYTES;
  uint8_t* ring = stage_area + NU_STAGE_BYTES;
  const uint32_t ring_base = smem_base + X_BYTES + NU_STAGE_BYTES;
  const uint32_t xbuf0 = smem_base, xbuf1 = ring_base + 2 * W_SLOT;  // (an indexed array would live in local memory)
  const int total_slots = p.num_stages * NUM_KS;

  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    for (int s = 0; s < 3; ++s) mbar_init(&bar_x[s], 2);  // one elected arrival per CTA
    mbar_init(&bar_acc[0], 1);
    mbar_init(&bar_acc[1], 1);
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
  // the peer's barriers exist before this CTA arrives on them
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) NU_STAMP(1);

  if (warp == NU_WORKERS) {
    // ------------------------------------------------------------------ TMA producer: this CTA's half of every W
    if (lane == 0) {
      for (int s = 0; s < p.num_stages; ++s)
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.w[s])) : "memory");
      for (int g = 0; g < total_slots; ++g) {
        const int slot = g % NSLOT, round = g / NSLOT;
        if (round > 0) mbar_wait(&bar_empty[slot], (uint32_t)((round - 1) & 1));
        mbar_arrive_expect_tx(&bar_full[slot], (uint32_t)W_SLOT);
#pragma unroll
        for (int k = 0; k < KP; ++k)
          tma_load_2d(ring + (size_t)slot * W_SLOT + (size_t)k * W_PANEL, &maps.w[g / NUM_KS], &bar_full[slot],
                      ((g % NUM_KS) * KP + k) * TC_BK, (int)rank * 128);
      }
    }
  } else if (warp == NU_WORKERS + 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(NT);
      for (int s = 0; s < p.num_stages; ++s) {
        const int b = s & 1;
        nu_wait_cluster(&bar_x[s], 0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // (remote) generic-proxy writes -> UMMA reads
        tc_fence_after();
        NU_STAMP(8 + 4 * s);
        const uint32_t xaddr = b ? xbuf1 : xbuf0;
        for (int ks = 0; ks < NUM_KS; ++ks) {
          const int g = s * NUM_KS + ks;
          const int slot = g % NSLOT, round = g / NSLOT;
          mbar_wait(&bar_full[slot], (uint32_t)(round & 1));
          tc_fence_after();
          if (ks == 0) NU_STAMP(9 + 4 * s);
#pragma unroll
          for (int k = 0; k < KP; ++k) {
            const int kb = ks * KP + k;
            const uint64_t bdesc = umma_desc_sw128(xaddr + (uint32_t)(kb * X_PANEL));
            const uint64_t adesc = umma_desc_sw128(ring_base + (uint32_t)(slot * W_SLOT + k * W_PANEL));
#pragma unroll
            for (int kk = 0; kk < TC_BK / 8; ++kk)
              umma_tf32(tmem + (uint32_t)((b * 2 + (kk & 1)) * NT), adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk),
                        idesc, (kb != 0 || kk >= 2) ? 1u : 0u);
          }
          umma_commit(&bar_empty[slot]);
        }
        umma_commit(&bar_acc[b]);
        NU_STAMP(10 + 4 * s);
      }
    }
  } else {
    // ------------------------------------------------------------------ workers
    const uint32_t peer = rank ^ 1u;
    // (1) this CTA's NA atoms of the stage-0 B operand -> both CTAs' buffer 0
    {
      constexpr int CAP = NU_STAGE_BYTES / 8;
      int* s_eid = reinterpret_cast<int*>(stage_area);
      int* s_src = s_eid + CAP;
      // the cluster's npc atoms are split evenly between the two CTAs (npc < NT when the host sizes the tiles to the SMs)
      const int na0 = (npc + 1) >> 1;
      const int my0 = rank == 0 ? 0 : na0, my_n = rank == 0 ? na0 : npc - na0;
      // the in-CSR ids: written by the edge build at the start of the step (complete before the first node kernel
      // started), not by the preceding node kernel -- no need to wait for it
      // one warp per atom: its in-edge range, then (first warp) the prefix of the counts = the staged positions
      if (warp == 0) {
        int cnt = 0;
        if (lane < my_n) {
          const int a = atom_of(my0 + lane);
          const int beg = a < N ? p.in_ptr[a] : 0;
          cnt = a < N ? p.in_ptr[a + 1] - beg : 0;
          s_gbeg[lane] = beg;
        }
        int inc = cnt;
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(TSD_FULL_MASK, inc, o);
          if (lane >= o) inc += t;
        }
        if (lane < my_n) s_ptr[lane + 1] = inc;
        if (lane == 0) s_ptr[0] = 0;
      }
      asm volatile("bar.sync 1, %0;" ::"r"(NW) : "memory");
      const int seg_n = s_ptr[my_n];
      const bool staged = seg_n <= CAP;
      if (staged) {
        for (int n = warp; n < my_n; n += NU_WORKERS) {
          const int gb = s_gbeg[n], lb = s_ptr[n], cnt = s_ptr[n + 1] - lb;
          for (int i = lane; i < cnt; i += 32) {
            const int e = p.in_eid[gb + i];
            s_eid[lb + i] = e;
            s_src[lb + i] = p.in_src[gb + i];
            // the filter rows were written by the filter stack tens of microseconds (and ~90 MB) ago: pull them back
            // into L2 while the preceding node kernel is still running (profiles/r3_pair_gather_cost.txt)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.filt + (size_t)e * H), "r"(H * 4) : "memory");
          }
        }
        asm volatile("bar.sync 1, %0;" ::"r"(NW) : "memory");
      }
      const uint32_t x_own = smem_base, x_peer = nu_mapa(smem_base, peer);  // own copy: plain shared-memory stores
      pdl_wait();  // see k_node_update
      pdl_trigger();
      if (tid == 0) NU_STAMP(7);
      // rows [npc, NT) of the B operand (padding of the MMA shape) are never written: their accumulator columns are
      // garbage that no epilogue stores
      for (int item = warp; item < my_n * 2; item += NU_WORKERS) {
        const int n = item >> 1, slab = item & 1;
        const int off = slab * 128 + lane * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);  // rows past the last atom: zeros (their output columns are never stored)
        if (atom_of(my0 + n) < N) {
          if (staged) acc = nu_aggregate_item(p, s_eid, s_src, s_ptr[n], s_ptr[n + 1], off, lane);
          else acc = nu_aggregate_item(p, p.in_eid, p.in_src, s_gbeg[n], s_gbeg[n] + s_ptr[n + 1] - s_ptr[n], off, lane);
        }
        const float4 r = tf32_rn4(acc);
        const uint32_t o = (uint32_t)((off >> 5) * X_PANEL) + sw128_off(my0 + n, (off & 31) >> 2);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(x_own + o), "f"(r.x), "f"(r.y), "f"(r.z), "f"(r.w) : "memory");
        nu_st_cluster4(x_peer + o, r);
      }
      // every thread fences its own writes, the workers meet, ONE thread publishes with cluster-scope release (instead of
      // 512 release arrivals per hand-over: 15 % of the kernel's stall samples in ncu, though neutral on the step time)
      asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy writes -> async proxy (both CTAs' UMMA)
      asm volatile("bar.sync 1, %0;" ::"r"(NW) : "memory");
      if (tid == 0) {
        nu_arrive_cluster(nu_mapa(smem_u32(&bar_x[0]), rank));
        nu_arrive_cluster(nu_mapa(smem_u32(&bar_x[0]), peer));
        NU_STAMP(2);
      }
    }
    // (2) epilogues.  warp -> TMEM lane quarter q (hardware: warp_id % 4) = 32 of this CTA's 128 features, atom group cs
    const int q = warp & 3, cs = warp >> 2;
    const int f = (int)rank * 128 + q * 32 + lane;
    const int n0 = cs * CW;
    for (int s = 0; s < p.num_stages; ++s) {
      const int b = s & 1;
      const float* const st_bias = p.st[s].bias;  // in registers: see k_node_update
      const float* const st_res = p.st[s].residual;
      float* const st_store = p.st[s].store;
      const bool st_ssp = p.st[s].act == TSD_ACT_SSP;
      const bool feeds = s + 1 < p.num_stages;
      const float bias = st_bias ? __ldg(st_bias + f) : 0.f;
      float res[CW];
      if (st_res) {  // independent of the accumulator: in flight behind the MMA
#pragma unroll
        for (int j = 0; j < CW; ++j) res[j] = st_res[(size_t)min(atom_of(n0 + j), N - 1) * H + f];
      }
      mbar_wait(&bar_acc[b], (uint32_t)((s >> 1) & 1));
      tc_fence_after();
      if (tid == 0) NU_STAMP(11 + 4 * s);
      uint32_t v0[CW], v1[CW];
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * 2 * NT + n0);
      tmem_ld_cols_async<CW>(taddr, v0);
      tmem_ld_cols_async<CW>(taddr + (uint32_t)NT, v1);
      tmem_wait_ld();
      const uint32_t xo = (((s + 1) & 1) ? xbuf1 : xbuf0) + (uint32_t)((f >> 5) * X_PANEL + ((f & 3) << 2));
      const uint32_t x_own = xo, x_peer = nu_mapa(xo, peer);  // own copy: plain shared-memory stores
      const int chunk = (f & 31) >> 2;
      float r[CW];
#pragma unroll
      for (int j = 0; j < CW; ++j) {
        r[j] = (__uint_as_float(v0[j]) + __uint_as_float(v1[j])) + bias;
        if (st_ssp) r[j] = tc_act<TSD_ACT_SSP>(r[j]);
        if (st_res) r[j] += res[j];
      }
      // the next stage's operand first: the release of the arrivals below would otherwise wait for the acknowledgement of
      // the global stores as well (1.1 us per stage measured, profiles/r3_node_pair_timeline.txt)
      if (feeds) {
#pragma unroll
        for (int j = 0; j < CW; ++j) {
          const int n = n0 + j;
          const uint32_t o = (uint32_t)(n * 128 + ((chunk ^ (n & 7)) << 4));
          const float t = tf32_rn(r[j]);
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(x_own + o), "f"(t) : "memory");
          nu_st_cluster1(x_peer + o, t);
        }
        asm volatile("fence.proxy.async;" ::: "memory");
        tc_fence_before();
        asm volatile("bar.sync 1, %0;" ::"r"(NW) : "memory");
        if (tid == 0) {
          nu_arrive_cluster(nu_mapa(smem_u32(&bar_x[s + 1]), rank));
          nu_arrive_cluster(nu_mapa(smem_u32(&bar_x[s + 1]), peer));
        }
      }
      if (st_store) {
#pragma unroll
        for (int j = 0; j < CW; ++j)
          if (n0 + j < npc && atom_of(n0 + j) < N) st_store[(size_t)atom_of(n0 + j) * H + f] = r[j];
      }
      if (tid == 0) NU_STAMP(3 + s);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) NU_STAMP(6);
  if (warp == NU_WORKERS + 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)tmem_cols)
                 : "memory");
  }
}

template <int NT>
int node_pair_launch(const NodeArgs& a, const NodeMaps& maps, cudaStream_t stream) {
  constexpr int X_BYTES = NT * 256 * 4, W_SLOT = 4 * 128 * TC_BK * 4;
  const size_t smem = 1024 + (size_t)2 * X_BYTES + NU_STAGE_BYTES + (size_t)2 * W_SLOT;
  static size_t attr_smem = 0;  // per instantiation
  if (smem > attr_smem) {
    TSD_CUDA(cudaFuncSetAttribute(k_node_pair<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  int tmem_cols = 32;  // a power of two >= 32 that holds 2 accumulator sets x 2 K parities x NT columns
  while (tmem_cols < 4 * NT) tmem_cols *= 2;
  cudaLaunchConfig_t cfg = {};
  const int npc = a.nodes_per_cluster > 0 && a.nodes_per_cluster < NT ? a.nodes_per_cluster : NT;
  cfg.gridDim = dim3(tsd_ceil_div(a.num_nodes, npc) * 2);
  cfg.blockDim = dim3(NU_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_node_pdl ? 2 : 1;
  TSD_CUDA(cudaLaunchKernelEx(&cfg, k_node_pair<NT>, a, maps, tmem_cols));
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

}  // namespace

// Rows per CTA: enough CTAs to spread the aggregation's gathers over the GPU, few enough that the weight
// stream (every CTA reads every W once) stays small against them.
#ifdef TSD_NODE_DBG
extern "C" void tsd_node_dbg_read(unsigned long long* out) { cudaMemcpyFromSymbol(out, g_node_dbg, sizeof(g_node_dbg)); }
extern "C" void tsd_node_cta_read(unsigned long long* out) { cudaMemcpyFromSymbol(out, g_node_cta, sizeof(g_node_cta)); }
#endif
extern "C" void tsd_tune_node_pdl(int on) { g_node_pdl = on; }
static int g_node_tile_override = 0, g_node_npc_override = 0;
extern "C" void tsd_tune_node_npc(int atoms) { g_node_npc_override = atoms; }
// tuning hook of profiles/scripts (not part of the C-ABI header): code = atoms per cluster * 10 + CTAs per cluster,
// 0 restores the built-in choice
extern "C" void tsd_tune_node_tile(int code) { g_node_tile_override = code; }

// Kernel shape (atoms per cluster * 10 + CTAs per cluster) and the atoms a cluster really takes (0 = all of the shape).
// Measured at batch 100 (profiles/r2_variants_*.txt, profiles/r3_stack_modes.txt): fewer atoms per CTA shorten the
// aggregation phase (bound by one SM's L2 ingest, 127 GB/s) but multiply the weight streams and the CTAs that compete
// with concurrently running filter kernels for SMs; more atoms per CTA (48, 64) lengthen the serial node chain.  Next to
// per-block filter kernels one CTA per 32 atoms is best; when the node chain runs alone (behind the filter stack) the
// pair kernel is (two CTAs share a 32-atom tile's gathers AND split the output features of the GEMMs; H = 128 and
// dense inputs fall back to 322 / 321).
// The pair kernel's clusters take fewer than 32 atoms when that spreads the gathers over more SMs: as many clusters as
// fit 90 % of the SMs at one CTA each (the rest is where the next node kernel's CTAs set up under programmatic dependent
// launch); measured at batch 100 (1826 atoms): 28 atoms per cluster 273 us per step, 32: 277, 26: 280, 25 (all SMs): 291.
int tsd_node_tile(bool alone, int num_nodes, int* nodes_per_cluster) {
  *nodes_per_cluster = g_node_npc_override;
  if (g_node_tile_override > 0) return g_node_tile_override;
  if (!alone) return 321;
  if (g_node_npc_override == 0) {
    static int num_sms = 0;
    if (num_sms == 0) {
      int dev = 0;
      if (cudaGetDevice(&dev) != cudaSuccess ||
          cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        num_sms = 148;
    }
    const int clusters = num_sms * 9 / 20;
    int npc = clusters > 0 ? tsd_ceil_div(num_nodes, clusters) : 32;
    npc = npc < 8 ? 8 : npc;
    *nodes_per_cluster = npc >= 32 ? 0 : npc;
  }
  return 323;
}

int tsd_node_update_tf32(const NodeArgs& a, int tile, cudaStream_t stream) {
  using namespace tc;
  if (!(a.H == 128 || a.H == 256) || a.num_stages < 1 || a.num_stages > 3) return TSD_ERR_UNSUPPORTED;
  if (a.num_nodes <= 0) return TSD_OK;
  TSD_REQUIRE(a.x || (a.in_ptr && a.in_eid && a.in_src && a.x1 && a.filt));
  // the pair variant (code NT * 10 + 3) needs the fused aggregation and H = 256; other inputs take the one-CTA shape
  if (tile % 10 == 3 && (a.x || a.H != 256)) tile = a.x ? 321 : 322;
  NodeMaps maps;
  for (int s = 0; s < 3; ++s) {
    const float* w = a.st[s < a.num_stages ? s : 0].W;
    TSD_REQUIRE(w);
    if (reinterpret_cast<uintptr_t>(w) & 15) return TSD_ERR_UNSUPPORTED;
    if (!make_tensor_map(&maps.w[s], w, (uint64_t)a.H, (uint64_t)a.H, tile % 10 == 3 ? 128u : (uint32_t)a.H))
      return TSD_ERR_UNSUPPORTED;
  }
  if (tile == 323) return node_pair_launch<32>(a, maps, stream);
#define NU_GO(NT, C)                                                  \
  if (tile == NT * 10 + C)                                            \
    return a.H == 256 ? node_launch<256, NT, C>(a, maps, stream) : node_launch<128, NT, C>(a, maps, stream)
  NU_GO(16, 1);
  NU_GO(32, 1);
  NU_GO(48, 1);
  NU_GO(64, 1);
  NU_GO(32, 2);
  NU_GO(64, 4);
#undef NU_GO
  return TSD_ERR_UNSUPPORTED;
}
