// The node side of ALL interaction blocks of a SchNet encoder in ONE persistent kernel (tf32, H = 256).
//
//   per block l:  agg_i = sum_{j->i} x1_j * filt_l,ji ;  y = ssp(lin2_l(agg)) ;  h' = h + lin_l(y) ;  x1' = lin1_{l+1}(h')
//
// The per-block kernel (k_node_pair, node_update.cu) spends 18 us per block at batch 100, of which only ~6.5 us are the
// three GEMM stages: the rest is gathers (filter rows 3.4 us, x1 rows 1.3 us: profiles/r3_pair_gather_cost.txt), the
// per-launch set-up, the staging of the in-CSR ids and the kernel-to-kernel hand-over.  Programmatic dependent launch
// cannot hide that work: one kernel's CTAs already take 90 % of the SMs, so the next kernel's CTAs only become resident
// when their predecessors exit (landing the filter rows in shared memory before the dependency wait made the step SLOWER
// for that reason: 284 vs 263 us).  Here the same CTAs stay resident for the whole encoder:
//
//   * the same two-CTA cluster as k_node_pair per tile of <= 32 atoms (each CTA: half of the atoms' gathers, 128 of the 256
//     output features of every GEMM stage, activations exchanged through distributed shared memory);
//   * set-up, TMEM allocation and the in-CSR ids once per step instead of once per block;
//   * between two blocks a grid barrier (one release / acquire counter in global memory): while a CTA waits there for the
//     slowest cluster, its FILTER ROWS of the next block -- the filter stack's output, independent of the node chain --
//     arrive in a shared-memory landing zone by 1 KiB bulk copies (the weight ring, the second B operand and 24 spare
//     KiB, all idle between the last MMA of a block and the first of the next: 184 rows); after the barrier only the x1
//     rows are gathered (ld.global.cg: other SMs wrote them during this kernel), 16 per warp in flight;
//   * the weight stream of a block's first stage starts when the workers have left the landing zone, behind the hand-over
//     of the aggregated rows.
//
// All clusters must be resident at once (the host checks cudaOccupancyMaxActiveClusters; otherwise one k_node_pair per
// block); the barrier spins with a 2 s timeout that raises an error flag instead of hanging the GPU.
#include <string.h>

#include "tc_common.cuh"

namespace {
using namespace tc;

// -DTSD_NODE_DBG: %globaltimer stamps of CTA 0 (profiles/scripts/node_chain_timeline.py); off in the product build
#ifdef TSD_NODE_DBG
__device__ unsigned long long g_nc_dbg[128];
#define NC_STAMP(slot)                               \
  do {                                               \
    if (blockIdx.x == 0) g_nc_dbg[slot] = gtimer();  \
  } while (0)
#else
#define NC_STAMP(slot) do {} while (0)
#endif

constexpr int NC_WORKERS = 16;
constexpr int NC_THREADS = (NC_WORKERS + 2) * 32;
constexpr int NC_NT = 32;                        // atoms per cluster tile (MMA N)
constexpr int NC_IDS_BYTES = 8 * 1024;           // staged in-CSR ids: 2 x 1024 entries
constexpr int NC_LAND_EXTRA = 24 * 1024;         // spare shared memory appended to the landing zone

struct NodeChainMaps {
  CUtensorMap w[3 * TSD_NC_MAX_BLOCKS];  // block l: lin2, lin, lin1 of block l + 1 (box = 128 rows x 32 floats)
};

// what the kernel needs of NodeChainArgs without the weight pointers (they live in the tensor maps)
struct NcBlockDev {
  const float* filt;
  const float* b_lin2;
  const float* b_lin;
};
struct NcArgsDev {
  int num_nodes, num_blocks, nodes_per_cluster;
  const int* in_ptr;
  const int* in_eid;
  const int* in_src;
  const float* x1_first;
  float* x1buf[2];
  const float* h_in;
  float* h_out;
  unsigned int* barrier;
  int* error_flag;
  NcBlockDev blk[TSD_NC_MAX_BLOCKS];
};

__device__ __forceinline__ void nc_fma_rn4(float4& acc, const float4& x, const float4& w) {
  acc.x = __fadd_rn(acc.x, __fmul_rn(x.x, w.x));
  acc.y = __fadd_rn(acc.y, __fmul_rn(x.y, w.y));
  acc.z = __fadd_rn(acc.z, __fmul_rn(x.z, w.z));
  acc.w = __fadd_rn(acc.w, __fmul_rn(x.w, w.w));
}
__device__ __forceinline__ uint32_t nc_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t nc_mapa(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void nc_arrive_cluster(uint32_t addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void nc_wait_cluster(uint64_t* bar, uint32_t parity) {
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

// agg of one (atom, slab) item: x1 rows gathered from global memory (L2: written by other SMs during this kernel), U at
// a time; filter rows from the landing zone (row li of the CTA's in-CSR segment, li < land_n) or, past its capacity,
// from global memory.  Ascending source order (the same sum as a sequential scatter_add).
template <int U>
__device__ __forceinline__ float4 nc_aggregate_item(const float* __restrict__ x1, const float* __restrict__ filt,
                                                    const int* eids, const int* srcs, int beg, int end, int off, int lane,
                                                    const uint8_t* land, int seg0, int land_n) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int base = beg; base < end; base += U) {
    const int cnt = min(U, end - base);
    const int r_l = lane < cnt ? srcs[base + lane] : 0;
    const int e_l = lane < cnt ? eids[base + lane] : 0;
    float4 x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = __shfl_sync(TSD_FULL_MASK, r_l, min(u, cnt - 1));
      x[u] = __ldcg(reinterpret_cast<const float4*>(x1 + (size_t)r * 256 + off));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int e = __shfl_sync(TSD_FULL_MASK, e_l, min(u, cnt - 1));
      const int li = base + u - seg0;
      if (u < cnt) {
        const float4 w = li < land_n ? *reinterpret_cast<const float4*>(land + (size_t)li * 1024 + (size_t)off * 4)
                                     : __ldg(reinterpret_cast<const float4*>(filt + (size_t)e * 256 + off));
        nc_fma_rn4(acc, x[u], w);
      }
    }
  }
  return acc;
}

__global__ void __launch_bounds__(NC_THREADS, 1) k_node_chain(const NcArgsDev p, const __grid_constant__ NodeChainMaps maps,
                                                              int tmem_cols) {
  constexpr int H = 256, NT = NC_NT;
  constexpr int NUM_KB = H / TC_BK;            // K panels per stage
  constexpr int W_PANEL = 128 * TC_BK * 4;     // this CTA's 128 rows of one K panel
  constexpr int KP = 4;                        // K panels per ring slot (one tcgen05.commit per 16 MMAs)
  constexpr int W_SLOT = KP * W_PANEL;         // 64 KiB
  constexpr int NUM_KS = NUM_KB / KP;          // slots per stage
  constexpr int NSLOT = 2;
  constexpr int X_PANEL = NT * TC_BK * 4;
  constexpr int X_BYTES = NUM_KB * X_PANEL;    // one B operand: NT x H floats
  constexpr int NA = NT / 2;
  constexpr int NW = NC_WORKERS * 32;
  constexpr int CW = NT / 4;                   // accumulator columns (atoms) per epilogue warp
  constexpr int LAND_ROWS = (2 * W_SLOT + X_BYTES + NC_LAND_EXTRA) / 1024;
  constexpr int CAP = NC_IDS_BYTES / 8;
  constexpr int U = 16;
  extern __shared__ uint8_t smem_dyn[];
  __shared__ uint64_t bar_full[NSLOT];
  __shared__ uint64_t bar_empty[NSLOT];
  __shared__ uint64_t bar_x[3];    // B operand of stage s complete: one arrival per CTA
  __shared__ uint64_t bar_acc[2];  // accumulator set complete (tcgen05.commit)
  __shared__ uint64_t bar_land;    // the landed filter rows of a block have arrived (bulk-copy bytes)
  __shared__ uint64_t bar_agg;     // every worker has left the landing zone: the weight ring may use it
  __shared__ uint32_t tmem_base_s;
  __shared__ int s_ptr[NA + 1];   // staged (local) positions of this CTA's atoms' in-edges
  __shared__ int s_gbeg[NA];      // ... and where they start in the global in-CSR

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = nc_cluster_rank();
  const uint32_t peer = rank ^ 1u;
  const int npc = p.nodes_per_cluster > 0 ? min(p.nodes_per_cluster, NT) : NT;
  // row n of cluster c's tile is atom c + n * (number of clusters): see k_node_pair (load balance of the gathers)
  const int cluster = blockIdx.x / 2, T = gridDim.x / 2;
  const int N = p.num_nodes;
  auto atom_of = [&](int n) { return cluster + n * T; };
  const int L = p.num_blocks;
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_dyn + (smem_base - smem_u32(smem_dyn));
  // layout: B operand 0 | staged in-CSR ids | ring slot 0 | ring slot 1 | B operand 1 | spare; the landing zone of the
  // filter rows is everything from the ring on
  uint8_t* ids_area = smem_gen + X_BYTES;
  uint8_t* ring = ids_area + NC_IDS_BYTES;
  const uint32_t ring_base = smem_base + X_BYTES + NC_IDS_BYTES;
  const uint32_t xbuf0 = smem_base, xbuf1 = ring_base + 2 * W_SLOT;

  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    for (int s = 0; s < 3; ++s) mbar_init(&bar_x[s], 2);
    mbar_init(&bar_acc[0], 1);
    mbar_init(&bar_acc[1], 1);
    mbar_init(&bar_land, 1);
    mbar_init(&bar_agg, NW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == NC_WORKERS + 1) {
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

  if (warp == NC_WORKERS) {
    // ------------------------------------------------------------------ TMA producer: this CTA's half of every W
    if (lane == 0) {
      int g = 0;
      for (int l = 0; l < L; ++l) {
        const int stages = l + 1 < L ? 3 : 2;
        for (int s = 0; s < stages; ++s)
          asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.w[3 * l + s])) : "memory");
        mbar_wait(&bar_agg, (uint32_t)(l & 1));  // the ring is the landing zone until the block's aggregation is done
        for (int gs = 0; gs < stages * NUM_KS; ++gs, ++g) {
          const int slot = g % NSLOT, round = g / NSLOT;
          if (round > 0) mbar_wait(&bar_empty[slot], (uint32_t)((round - 1) & 1));
          mbar_arrive_expect_tx(&bar_full[slot], (uint32_t)W_SLOT);
#pragma unroll
          for (int k = 0; k < KP; ++k)
            tma_load_2d(ring + (size_t)slot * W_SLOT + (size_t)k * W_PANEL, &maps.w[3 * l + gs / NUM_KS], &bar_full[slot],
                        ((gs % NUM_KS) * KP + k) * TC_BK, (int)rank * 128);
        }
      }
    }
  } else if (warp == NC_WORKERS + 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(NT);
      int g = 0, gstage = 0;
      for (int l = 0; l < L; ++l) {
        const int stages = l + 1 < L ? 3 : 2;
        for (int s = 0; s < stages; ++s, ++gstage) {
          const int b = gstage & 1;
          nc_wait_cluster(&bar_x[s], (uint32_t)(l & 1));
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // (remote) generic-proxy writes -> UMMA reads
          tc_fence_after();
          const uint32_t xaddr = (s & 1) ? xbuf1 : xbuf0;
          for (int ks = 0; ks < NUM_KS; ++ks, ++g) {
            const int slot = g % NSLOT, round = g / NSLOT;
            mbar_wait(&bar_full[slot], (uint32_t)(round & 1));
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < KP; ++k) {
              const int kb = ks * KP + k;
              const uint64_t bdesc = umma_desc_sw128(xaddr + (uint32_t)(kb * X_PANEL));
              const uint64_t adesc = umma_desc_sw128(ring_base + (uint32_t)(slot * W_SLOT + k * W_PANEL));
              // K steps alternate between two accumulators (consecutive MMAs into one accumulator serialise)
#pragma unroll
              for (int kk = 0; kk < TC_BK / 8; ++kk)
                umma_tf32(tmem + (uint32_t)((b * 2 + (kk & 1)) * NT), adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk),
                          idesc, (kb != 0 || kk >= 2) ? 1u : 0u);
            }
            umma_commit(&bar_empty[slot]);
          }
          umma_commit(&bar_acc[b]);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ workers
    int* s_eid = reinterpret_cast<int*>(ids_area);
    int* s_src = s_eid + CAP;
    // the cluster's npc atoms are split evenly between the two CTAs
    const int na0 = (npc + 1) >> 1;
    const int my0 = rank == 0 ? 0 : na0, my_n = rank == 0 ? na0 : npc - na0;
    // in-CSR ids of this CTA's atoms: the same for every block
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
    const int seg_n = min(s_ptr[my_n], CAP);  // atoms whose ids do not fit the staging area read them from global memory
    const int land_n = min(seg_n, LAND_ROWS);
    for (int n = warp; n < my_n; n += NC_WORKERS) {
      const int gb = s_gbeg[n], lb = s_ptr[n], cnt = s_ptr[n + 1] - lb;
      for (int i = lane; i < cnt && lb + i < CAP; i += 32) {
        s_eid[lb + i] = p.in_eid[gb + i];
        s_src[lb + i] = p.in_src[gb + i];
      }
    }
    asm volatile("bar.sync 1, %0;" ::"r"(NW) : "memory");
    const int seg0 = 0;
    const int* eids = s_eid;
    const int* srcs = s_src;
    // the filter rows of block l -> landing zone (rows past its capacity: at least back into L2)
    auto land_block = [&](int l) {
      if (tid == 0) mbar_arrive_expect_tx(&bar_land, (uint32_t)(land_n * 1024));
      const float* filt = p.blk[l].filt;
      for (int i = tid; i < seg_n; i += NW) {
        const float* src = filt + (size_t)eids[seg0 + i] * H;
        if (i < land_n)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           ring_base + (uint32_t)i * 1024u),
                       "l"(src), "r"(1024), "r"(smem_u32(&bar_land))
                       : "memory");
        else
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(1024) : "memory");
      }
    };
    land_block(0);

    const int q = warp & 3, cs = warp >> 2;  // epilogue: TMEM lane quarter = 32 of this CTA's 128 features, atom group
    const int f = (int)rank * 128 + q * 32 + lane;
    const int n0 = cs * CW;
    int gstage = 0;
    for (int l = 0; l < L; ++l) {
      const int stages = l + 1 < L ? 3 : 2;
      if (tid == 0) NC_STAMP(8 * l + 0);
      if (l > 0) {
        // ---- grid barrier: every cluster has finished block l - 1 (its x1 / h rows are visible)
        asm volatile("bar.sync 1, %0;" ::"r"(NW) : "memory");
        if (tid == 0) {
          __threadfence();
          atomicAdd(p.barrier, 1u);
          const unsigned int target = (unsigned int)l * gridDim.x;
          const unsigned long long t0 = gtimer();
          unsigned int seen;
          do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.barrier) : "memory");
            if (seen < target) {
              __nanosleep(64);  // 132 CTAs polling one L2 line back to back starve the arrivals of the late ones
              if (gtimer() - t0 > 2000000000ull) {  // a cluster was never scheduled: do not hang the GPU
                if (p.error_flag) atomicOr(p.error_flag, 4);
                break;
              }
            }
          } while (seen < target);
          __threadfence();
        }
        asm volatile("bar.sync 1, %0;" ::"r"(NW) : "memory");
      }
      // ---- (1) aggregation: this CTA's atoms of the stage-0 B operand -> both CTAs' buffer 0
      const float* x1 = l == 0 ? p.x1_first : p.x1buf[l & 1];
      if (tid == 0) NC_STAMP(8 * l + 1);
      mbar_wait(&bar_land, (uint32_t)(l & 1));
      if (tid == 0) NC_STAMP(8 * l + 2);
      const uint32_t x_peer0 = nc_mapa(xbuf0, peer);
      for (int item = warp; item < my_n * 2; item += NC_WORKERS) {
        const int n = item >> 1, slab = item & 1;
        const int off = slab * 128 + lane * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);  // rows past the tile: zeros (their output columns are never stored)
        if (atom_of(my0 + n) < N) {
          const int lb = s_ptr[n], le = s_ptr[n + 1];
          if (le <= CAP) acc = nc_aggregate_item<U>(x1, p.blk[l].filt, eids, srcs, lb, le, off, lane, ring, seg0, land_n);
          else acc = nc_aggregate_item<U>(x1, p.blk[l].filt, p.in_eid, p.in_src, s_gbeg[n], s_gbeg[n] + (le - lb), off, lane,
                                          ring, 0, 0);
        }
        const float4 r = tf32_rn4(acc);
        const uint32_t o = (uint32_t)((off >> 5) * X_PANEL) + sw128_off(my0 + n, (off & 31) >> 2);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(xbuf0 + o), "f"(r.x), "f"(r.y), "f"(r.z), "f"(r.w)
                     : "memory");
        asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(x_peer0 + o), "f"(r.x), "f"(r.y), "f"(r.z),
                     "f"(r.w)
                     : "memory");
      }
      if (tid == 0) NC_STAMP(8 * l + 3);
      mbar_arrive(&bar_agg);  // this thread is done with the landing zone: the weight stream may start
      asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy writes -> async proxy (both CTAs' UMMA)
      asm volatile("bar.sync 1, %0;" ::"r"(NW) : "memory");
      if (tid == 0) {
        nc_arrive_cluster(nc_mapa(smem_u32(&bar_x[0]), rank));
        nc_arrive_cluster(nc_mapa(smem_u32(&bar_x[0]), peer));
      }
      // ---- (2) the block's GEMM stages: lin2 + ssp | lin + residual -> h | next lin1 -> x1
      for (int s = 0; s < stages; ++s, ++gstage) {
        const int b = gstage & 1;
        const float* const st_bias = s == 0 ? p.blk[l].b_lin2 : (s == 1 ? p.blk[l].b_lin : nullptr);
        const float* const st_res = s == 1 ? (l == 0 ? p.h_in : p.h_out) : nullptr;
        float* const st_store = s == 0 ? nullptr : (s == 1 ? p.h_out : p.x1buf[(l + 1) & 1]);
        const bool st_ssp = s == 0;
        const bool feeds = s + 1 < stages;
        const float bias = st_bias ? __ldg(st_bias + f) : 0.f;
        float res[CW];
        if (st_res) {  // independent of the accumulator: in flight behind the MMA (rows this very thread wrote last block)
#pragma unroll
          for (int j = 0; j < CW; ++j) res[j] = st_res[(size_t)min(atom_of(n0 + j), N - 1) * H + f];
        }
        mbar_wait(&bar_acc[b], (uint32_t)((gstage >> 1) & 1));
        tc_fence_after();
        if (tid == 0) NC_STAMP(8 * l + 4 + s);
        uint32_t v0[CW], v1[CW];
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * 2 * NT + n0);
        tmem_ld_cols_async<CW>(taddr, v0);
        tmem_ld_cols_async<CW>(taddr + (uint32_t)NT, v1);
        tmem_wait_ld();
        float r[CW];
#pragma unroll
        for (int j = 0; j < CW; ++j) {
          r[j] = (__uint_as_float(v0[j]) + __uint_as_float(v1[j])) + bias;
          if (st_ssp) r[j] = tc_act<TSD_ACT_SSP>(r[j]);
          if (st_res) r[j] += res[j];
        }
        if (feeds) {  // the next stage's operand first, then the global stores
          const uint32_t xo = (((s + 1) & 1) ? xbuf1 : xbuf0) + (uint32_t)((f >> 5) * X_PANEL + ((f & 3) << 2));
          const uint32_t x_peer = nc_mapa(xo, peer);
          const int chunk = (f & 31) >> 2;
#pragma unroll
          for (int j = 0; j < CW; ++j) {
            const int n = n0 + j;
            const uint32_t o = (uint32_t)(n * 128 + ((chunk ^ (n & 7)) << 4));
            const float t = tf32_rn(r[j]);
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(xo + o), "f"(t) : "memory");
            asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(x_peer + o), "f"(t) : "memory");
          }
          asm volatile("fence.proxy.async;" ::: "memory");
          tc_fence_before();
          asm volatile("bar.sync 1, %0;" ::"r"(NW) : "memory");
          if (tid == 0) {
            nc_arrive_cluster(nc_mapa(smem_u32(&bar_x[s + 1]), rank));
            nc_arrive_cluster(nc_mapa(smem_u32(&bar_x[s + 1]), peer));
          }
        }
        if (st_store) {
#pragma unroll
          for (int j = 0; j < CW; ++j)
            if (n0 + j < npc && atom_of(n0 + j) < N) st_store[(size_t)atom_of(n0 + j) * H + f] = r[j];
        }
      }
      // ---- the next block's filter rows while this CTA waits for the others: the ring and the second B operand are
      // idle (this block's last accumulator was complete before the epilogue above; the peer's last write into this CTA
      // preceded this CTA's last MMA)
      if (tid == 0) NC_STAMP(8 * l + 7);
      if (l + 1 < L) {
        tc_fence_before();
        asm volatile("bar.sync 1, %0;" ::"r"(NW) : "memory");  // every worker's TMEM / shared-memory reads of this block are done
        land_block(l + 1);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NC_WORKERS + 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)tmem_cols) : "memory");
  }
}

}  // namespace

#ifdef TSD_NODE_DBG
extern "C" void tsd_node_chain_dbg_read(unsigned long long* out) { cudaMemcpyFromSymbol(out, g_nc_dbg, sizeof(g_nc_dbg)); }
#endif

int tsd_node_chain_tf32(const NodeChainArgs& a, cudaStream_t stream) {
  using namespace tc;
  static_assert(sizeof(NcArgsDev) + sizeof(NodeChainMaps) + 16 <= 4096, "kernel parameter space");
  if (a.H != 256 || a.num_blocks < 1 || a.num_blocks > TSD_NC_MAX_BLOCKS || a.num_nodes <= 0) return TSD_ERR_UNSUPPORTED;
  TSD_REQUIRE(a.in_ptr && a.in_eid && a.in_src && a.x1_first && a.x1buf[0] && a.x1buf[1] && a.h_in && a.h_out && a.barrier);
  constexpr int X_BYTES = NC_NT * 256 * 4, W_SLOT = 4 * 128 * TC_BK * 4;
  const size_t smem = 1024 + (size_t)2 * X_BYTES + NC_IDS_BYTES + (size_t)2 * W_SLOT + NC_LAND_EXTRA;
  static bool attr_set = false;
  if (!attr_set) {
    TSD_CUDA(cudaFuncSetAttribute(k_node_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int npc = a.nodes_per_cluster > 0 && a.nodes_per_cluster < NC_NT ? a.nodes_per_cluster : NC_NT;
  const int clusters = tsd_ceil_div(a.num_nodes, npc);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * 2);
  cfg.blockDim = dim3(NC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // the grid barrier needs every cluster resident at once
  static int max_clusters = -1;
  if (max_clusters < 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, k_node_chain, &cfg) != cudaSuccess) {
      (void)cudaGetLastError();
      n = 0;
    }
    max_clusters = n;
  }
  if (clusters > max_clusters) return TSD_ERR_UNSUPPORTED;
  NodeChainMaps maps;
  NcArgsDev d;
  memset(&d, 0, sizeof(d));
  d.num_nodes = a.num_nodes;
  d.num_blocks = a.num_blocks;
  d.nodes_per_cluster = a.nodes_per_cluster;
  d.in_ptr = a.in_ptr;
  d.in_eid = a.in_eid;
  d.in_src = a.in_src;
  d.x1_first = a.x1_first;
  d.x1buf[0] = a.x1buf[0];
  d.x1buf[1] = a.x1buf[1];
  d.h_in = a.h_in;
  d.h_out = a.h_out;
  d.barrier = a.barrier;
  d.error_flag = a.error_flag;
  for (int l = 0; l < TSD_NC_MAX_BLOCKS; ++l) {
    const NodeChainBlock& b = a.blk[l < a.num_blocks ? l : 0];
    const bool last = l + 1 >= a.num_blocks;
    const float* w[3] = {b.w_lin2, b.w_lin, (l < a.num_blocks && !last) ? b.w_lin1_next : b.w_lin};
    TSD_REQUIRE(b.filt && b.w_lin2 && b.w_lin && (last || l >= a.num_blocks || b.w_lin1_next));
    for (int s = 0; s < 3; ++s) {
      if (reinterpret_cast<uintptr_t>(w[s]) & 15) return TSD_ERR_UNSUPPORTED;
      if (!make_tensor_map(&maps.w[3 * l + s], w[s], 256, 256, 128)) return TSD_ERR_UNSUPPORTED;
    }
    d.blk[l].filt = b.filt;
    d.blk[l].b_lin2 = b.b_lin2;
    d.blk[l].b_lin = b.b_lin;
  }
  int tmem_cols = 32;  // a power of two >= 32 that holds 2 accumulator sets x 2 K parities x NT columns
  while (tmem_cols < 4 * NC_NT) tmem_cols *= 2;
  TSD_CUDA(cudaLaunchKernelEx(&cfg, k_node_chain, d, maps, tmem_cols));
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}
