// The CFConv filter networks of ALL interaction blocks in one kernel.
//
//   filt_l = nn2_l(ssp(nn0_l(edge_attr))) * C(len)        l = 0 .. L-1        (schnet.py:91-98)
//
// Every block's filter network reads the same edge_attr and nothing else, so the work items (block, 128-row tile)
// are independent: the kernel is persistent, one CTA per SM, CTA b takes items b, b + gridDim.x, ... (block-major;
// 658 items on 148 SMs at batch 100).  Round 1/2 launched one two-GEMM kernel per block (gemm_chain.cu,
// k_chain_tf32<0>): 94 CTAs, main loop -> epilogue -> main loop -> epilogue in strict sequence, 7 launches of 22-30 us
// each (profiles/r2_kineto_step_tf32.txt).  Here, per item, the roles run concurrently and only meet at mbarriers:
//
//   warp 16      TMA producer: per item the tile's 8 edge_attr panels + the 8 panels of W0, then the
//                8 panels of W2, through a ring of {W panel, A panel} slots.  One SM ingests 64 B/clk
//                from L2 (profiles/r2_tma_stream.txt), exactly what the tensor pipe consumes per
//                M128 x N256 x K8 tf32 MMA of streamed W -- the ring must never drain.
//   warp 17      MMA issuer: acc1 = A . W0^T (TMEM columns [0,H)), then acc2 = X . W2^T (columns [H,2H))
//                panel by panel as soon as the activation warps have produced the matching K panel of X.
//   warps 0..15  epilogues.  epi1: acc1 -> + b0 -> shifted softplus -> TF32 (RNE) -> written back IN PLACE over
//                acc1 (tcgen05.st): X never leaves tensor memory, the second GEMM takes its A operand from
//                TMEM.  Handed to the MMA issuer K panel by K panel (32 columns), so the second GEMM trails the
//                activation by a panel instead of waiting for all of it.
//   warps 18..29 epi2 (three warps per TMEM lane quarter = 32 rows, 16-column units u % 3): acc2 -> + b2 -> * C(len)
//                -> 32 x 16 block in the warp's own 2 KiB staging -> TMA store (1.2 us per 32 x 32 block: the stores queue
//                behind the weight loads of the same SM; reading the block back and storing whole 128-byte lines
//                with st.global measured slower still, 1.75 us per block:
//                profiles/r3_filter_stack_timeline_{tma_store,st_global}.txt).  No CTA-level barrier anywhere; it
//                overlaps the NEXT item's first GEMM and activation (the MMA issuer waits on `acc2_free`
//                before it overwrites acc2).
//
// Shared memory (H = 256): 4 ring slots x 48 KiB + 12 x 2 KiB staging (the ring depth is what the weight
// stream's throughput hangs on: with X in shared memory there was room for 2 slots = 96 KiB in flight and the
// stream reached 61 of the 127 GB/s one SM can ingest; profiles/r3_filter_stack_timeline_smemX.txt).
// TMEM: all 512 columns.
#include <string.h>

#include "tc_common.cuh"

namespace {
using namespace tc;

// -DTSD_FS_DBG: %globaltimer stamps of CTA 0 (profiles/scripts/filter_stack_timeline.py); off in the product build
#ifdef TSD_FS_DBG
__device__ unsigned long long g_fs_dbg[256];
#define FS_STAMP(slot)                                     \
  do {                                                     \
    if (blockIdx.x == 0) g_fs_dbg[slot] = gtimer();        \
  } while (0)
#else
#define FS_STAMP(slot) do {} while (0)
#endif

constexpr int FS_EPI_WARPS = 16;
constexpr int FS_EPI_THREADS = FS_EPI_WARPS * 32;
constexpr int FS_STORE_WARPS = 12;
constexpr int FS_THREADS = (FS_EPI_WARPS + 2 + FS_STORE_WARPS) * 32;
constexpr int FS_STAGE_BYTES = FS_STORE_WARPS * 32 * 16 * 4;  // per store warp 32 rows x 64 B (half an output panel)
constexpr int FS_MAX_SLOTS = 6;
int g_fs_grid = 0;  // tuning hook of profiles/scripts: CTA count (0 = one per SM)

struct FilterStackMaps {
  CUtensorMap a;
  CUtensorMap w[2 * TSD_FS_MAX_LAYERS];
  CUtensorMap out[TSD_FS_MAX_LAYERS];
};

// the part of FilterStackArgs the kernel needs (kernel parameters + 25 tensor maps must stay below 4 KiB)
struct FsLayerDev {
  const float* b0;
  const float* b2;
  float cutoff;
  int smooth;
};
struct FsArgsDev {
  int M_cap;
  const int* M_ptr;
  int num_layers;
  const float* len;
  FsLayerDev layer[TSD_FS_MAX_LAYERS];
};

struct FsBars {
  uint64_t full[FS_MAX_SLOTS];
  uint64_t empty[FS_MAX_SLOTS];
  uint64_t acc1_full, acc2_full, acc2_free;
  uint64_t x_ready[8];  // one per K panel of X
  uint32_t tmem_base;
};

// epi1: CQ accumulator columns of this thread's row -> + bias -> shifted softplus -> TF32 (RNE), written back over
// the same TMEM columns: the A operand of the second GEMM
template <int CQ>
__device__ __forceinline__ void fs_activate_in_place(uint32_t taddr, const float* __restrict__ bias) {
  float4 b[CQ / 4];
#pragma unroll
  for (int j = 0; j < CQ / 4; ++j)
    b[j] = bias ? __ldg(reinterpret_cast<const float4*>(bias + 4 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
  uint32_t v[CQ];
  tmem_ld_cols<CQ>(taddr, v);
#pragma unroll
  for (int j = 0; j < CQ / 4; ++j) {
    v[4 * j + 0] = __float_as_uint(tf32_rn(tc_act<TSD_ACT_SSP>(__uint_as_float(v[4 * j + 0]) + b[j].x)));
    v[4 * j + 1] = __float_as_uint(tf32_rn(tc_act<TSD_ACT_SSP>(__uint_as_float(v[4 * j + 1]) + b[j].y)));
    v[4 * j + 2] = __float_as_uint(tf32_rn(tc_act<TSD_ACT_SSP>(__uint_as_float(v[4 * j + 2]) + b[j].z)));
    v[4 * j + 3] = __float_as_uint(tf32_rn(tc_act<TSD_ACT_SSP>(__uint_as_float(v[4 * j + 3]) + b[j].w)));
  }
  tmem_st_cols<CQ>(taddr, v);
}

// one K panel (32 floats = 4 MMAs of K = 8) of the first GEMM: A = the tile's edge_attr panel, B = the W0 panel (shared memory)
__device__ __forceinline__ void fs_mma_panel_ss(uint32_t acc, uint32_t a_addr, uint32_t w_addr, uint32_t idesc, bool first) {
  const uint64_t adesc = umma_desc_sw128(a_addr), bdesc = umma_desc_sw128(w_addr);
#pragma unroll
  for (int kk = 0; kk < TC_BK / 8; ++kk)
    umma_tf32(acc, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc, (!first || kk != 0) ? 1u : 0u);
}
// ... of the second GEMM: A = 32 TMEM columns of X (row m = lane m, one 32-bit column per K element), B = the W2 panel
__device__ __forceinline__ void fs_mma_panel_ts(uint32_t acc, uint32_t x_tmem, uint32_t w_addr, uint32_t idesc, bool first) {
  const uint64_t bdesc = umma_desc_sw128(w_addr);
#pragma unroll
  for (int kk = 0; kk < TC_BK / 8; ++kk)
    umma_tf32_ts(acc, x_tmem + (uint32_t)(8 * kk), bdesc + (uint64_t)(2 * kk), idesc, (!first || kk != 0) ? 1u : 0u);
}

template <int H>
__global__ void __launch_bounds__(FS_THREADS, 1) k_filter_stack(const FsArgsDev p, const __grid_constant__ FilterStackMaps maps,
                                                                int num_slots) {
  constexpr int NKB = H / TC_BK;                 // K panels per GEMM = 32-column panels of the output
  constexpr int W_PANEL = H * TC_BK * 4;         // bytes of one W panel: H rows x 128 B
  constexpr int SLOT = W_PANEL + TC_A_PANEL_BYTES;
  constexpr int CQ = TC_BK / 4;                  // X is handed over panel by panel: 8 of its 32 columns per activation warp
  static_assert(NKB <= 8 && NKB % 4 == 0, "H must be 128 or 256");
  extern __shared__ uint8_t smem_dyn[];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int M = p.M_ptr ? min(*p.M_ptr, p.M_cap) : p.M_cap;
  // Work items = (layer, 128-row tile), layer-major; CTA b takes items b, b + gridDim.x, ...  A tile's layers are
  // independent of each other (each reads edge_attr again), so the items spread over ALL SMs whatever the tile count.
  const int tiles = (M + TC_BM - 1) / TC_BM;
  const int items = tiles * p.num_layers;
  const int stride = (int)gridDim.x;
  if ((int)blockIdx.x >= items) return;
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_dyn + (smem_base - smem_u32(smem_dyn));
  uint8_t* stage = smem_gen;                     // per store warp 32 rows x 32 floats, SWIZZLE_128B
  uint8_t* ring = smem_gen + FS_STAGE_BYTES;
  const uint32_t ring_base = smem_base + FS_STAGE_BYTES;
  FsBars* bars = reinterpret_cast<FsBars*>(ring + (size_t)num_slots * SLOT);

  if (tid == 0) {
    for (int s = 0; s < num_slots; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    mbar_init(&bars->acc1_full, 1);
    mbar_init(&bars->acc2_full, 1);
    mbar_init(&bars->acc2_free, FS_STORE_WARPS);
    for (int t = 0; t < NKB; ++t) mbar_init(&bars->x_ready[t], FS_EPI_THREADS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.w[0])) : "memory");
  }
  if (warp == FS_EPI_WARPS + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                 "r"((uint32_t)(2 * H))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  const uint32_t acc1 = tmem, acc2 = tmem + (uint32_t)H;

  if (warp == FS_EPI_WARPS) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      int g = 0;
      for (int item = blockIdx.x; item < items; item += stride) {
        const int l = item / tiles, m0 = (item - l * tiles) * TC_BM;
        for (int st = 0; st < 2; ++st) {
          const CUtensorMap* wmap = &maps.w[2 * l + st];
          if (st == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(wmap + 1)) : "memory");
          else if (item + stride < items)
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.w[2 * ((item + stride) / tiles)])) : "memory");
          for (int kb = 0; kb < NKB; ++kb, ++g) {
            const int s = g % num_slots, round = g / num_slots;
            if (round > 0) mbar_wait(&bars->empty[s], (uint32_t)((round - 1) & 1));
            uint8_t* slot = ring + (size_t)s * SLOT;
            mbar_arrive_expect_tx(&bars->full[s], (uint32_t)(W_PANEL + (st == 0 ? TC_A_PANEL_BYTES : 0)));
            if (st == 0) tma_load_2d(slot + W_PANEL, &maps.a, &bars->full[s], kb * TC_BK, m0);
            tma_load_2d(slot, wmap, &bars->full[s], kb * TC_BK, 0);
            if (kb == 0 && st == 0) FS_STAMP(16 * l + 13);
            if (kb == NKB - 1) FS_STAMP(16 * l + 14 + st);
          }
        }
      }
    }
  } else if (warp == FS_EPI_WARPS + 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(H);
      int g = 0, it = 0;
      for (int item = blockIdx.x; item < items; item += stride, ++it) {
        const int l = item / tiles;
        (void)l;
        // acc1 = A . W0^T.  acc1 (= the previous layer's X) is free: the MMAs of one thread execute in issue order, so
        // these follow the previous layer's second GEMM, the only reader of X.
        for (int kb = 0; kb < NKB; ++kb, ++g) {
          const int s = g % num_slots, round = g / num_slots;
          mbar_wait(&bars->full[s], (uint32_t)(round & 1));
          tc_fence_after();
          if (kb == 0) FS_STAMP(16 * l + 9);
          const uint32_t slot = ring_base + (uint32_t)(s * SLOT);
          fs_mma_panel_ss(acc1, slot + (uint32_t)W_PANEL, slot, idesc, kb == 0);
          umma_commit(&bars->empty[s]);
        }
        umma_commit(&bars->acc1_full);
        FS_STAMP(16 * l + 10);
        // acc2 = X . W2^T, panel by panel of X as the activation warps hand them over, once the store warps have read the
        // previous item's acc2
        if (it > 0) {
          mbar_wait(&bars->acc2_free, (uint32_t)((it - 1) & 1));
          tc_fence_after();
        }
        for (int kb = 0; kb < NKB; ++kb, ++g) {
          mbar_wait(&bars->x_ready[kb], (uint32_t)(it & 1));
          tc_fence_after();
          if (kb == 0) FS_STAMP(16 * l + 11);
          const int s = g % num_slots, round = g / num_slots;
          mbar_wait(&bars->full[s], (uint32_t)(round & 1));
          tc_fence_after();
          fs_mma_panel_ts(acc2, acc1 + (uint32_t)(kb * TC_BK), ring_base + (uint32_t)(s * SLOT), idesc, kb == 0);
          umma_commit(&bars->empty[s]);
        }
        umma_commit(&bars->acc2_full);
        FS_STAMP(16 * l + 12);
      }
    }
  } else if (warp < FS_EPI_WARPS) {
    // ---------------------------------------------------------------- activation warps (epi1)
    const int q = warp & 3, cg = warp >> 2;  // TMEM lane quarter (hardware: warp id % 4), column group
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    int it = 0;
    for (int item = blockIdx.x; item < items; item += stride, ++it) {
      const int l = item / tiles;
      const float* const b0 = p.layer[l].b0;
      // X = tf32(ssp(acc1 + b0)), in place, one K panel of the second GEMM at a time
      mbar_wait(&bars->acc1_full, (uint32_t)(it & 1));
      tc_fence_after();
      if (tid == 0) FS_STAMP(16 * l + 0);
#pragma unroll
      for (int t = 0; t < NKB; ++t) {
        const int c0 = t * TC_BK + cg * CQ;
        fs_activate_in_place<CQ>(acc1 + lane_addr + (uint32_t)c0, b0 ? b0 + c0 : nullptr);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&bars->x_ready[t]);
        if (tid == 0 && (t & 1)) FS_STAMP(16 * l + 1 + (t >> 1));
      }
    }
  } else {
    // ---------------------------------------------------------------- store warps (epi2)
    const int q = warp & 3;  // TMEM lane quarter = rows [32 q, 32 q + 32) of the tile
    const bool stamp = warp == FS_EPI_WARPS + 2 && lane == 0;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    uint8_t* const buf = stage + (size_t)(warp - (FS_EPI_WARPS + 2)) * (32 * 16 * 4);
    const int u0 = (warp - (FS_EPI_WARPS + 2)) >> 2;  // this warp's 16-column units of the output: u0, u0 + 3, ...
    int it = 0;
    for (int item = blockIdx.x; item < items; item += stride, ++it) {
      const int l = item / tiles, m0 = (item - l * tiles) * TC_BM;
      const float* const b2 = p.layer[l].b2;
      const bool valid = m0 + row < M;
      const float cscale = valid ? tsd_cutoff_fn(p.len[min(m0 + row, M - 1)], p.layer[l].cutoff, p.layer[l].smooth) : 0.f;
      mbar_wait(&bars->acc2_full, (uint32_t)(it & 1));
      tc_fence_after();
      if (stamp) FS_STAMP(16 * l + 5);
#pragma unroll 1
      for (int u = u0; u < H / 16; u += FS_STORE_WARPS / 4) {
        uint32_t v[16];
        tmem_ld_cols<16>(acc2 + lane_addr + (uint32_t)(u * 16), v);
        if (lane == 0) tma_store_wait_read();  // this warp's previous store has read the staging block
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 b = b2 ? __ldg(reinterpret_cast<const float4*>(b2 + u * 16 + 4 * c)) : make_float4(0.f, 0.f, 0.f, 0.f);
          float4 o = make_float4((__uint_as_float(v[4 * c + 0]) + b.x) * cscale, (__uint_as_float(v[4 * c + 1]) + b.y) * cscale,
                                 (__uint_as_float(v[4 * c + 2]) + b.z) * cscale, (__uint_as_float(v[4 * c + 3]) + b.w) * cscale);
          if (!valid) o = make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(buf + sw64_off(lane, c)) = o;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> TMA store reads
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&maps.out[l], buf, u * 16, m0 + q * 32);
          tma_store_commit();
        }
      }
      tc_fence_before();  // this warp's acc2 reads precede the MMA issuer's next writes
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->acc2_free);
      if (stamp) FS_STAMP(16 * l + 6);
    }
    if (lane == 0) tma_store_wait_read();  // the staging buffers outlive the last stores' reads
    if (stamp) FS_STAMP(16 * (p.num_layers - 1) + 8);
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == FS_EPI_WARPS + 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)(2 * H)) : "memory");
  }
}

template <int H>
int fs_launch(const FsArgsDev& d, const FilterStackMaps& maps, int max_ctas, cudaStream_t stream) {
  constexpr int SLOT = H * TC_BK * 4 + TC_A_PANEL_BYTES;
  const int budget = 227 * 1024 - 1024 - FS_STAGE_BYTES - 256;  // alignment slack, staging, barriers
  int slots = budget / SLOT;
  if (slots > FS_MAX_SLOTS) slots = FS_MAX_SLOTS;
  if (slots < 2) return TSD_ERR_UNSUPPORTED;
  const size_t smem = 1024 + (size_t)FS_STAGE_BYTES + (size_t)slots * SLOT + 256;
  static size_t attr_smem = 0;  // per instantiation
  if (smem > attr_smem) {
    TSD_CUDA(cudaFuncSetAttribute(k_filter_stack<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  // one CTA per SM (the shared memory allows no more), never more CTAs than work items of the largest possible pair count
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    TSD_CUDA(cudaGetDevice(&dev));
    TSD_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int max_items = tsd_ceil_div(d.M_cap, TC_BM) * d.num_layers;
  int grid = g_fs_grid > 0 ? g_fs_grid : (max_items < num_sms ? max_items : num_sms);
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  k_filter_stack<H><<<dim3(grid), dim3(FS_THREADS), smem, stream>>>(d, maps, slots);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

}  // namespace

#ifdef TSD_FS_DBG
extern "C" void tsd_fs_dbg_read(unsigned long long* out) { cudaMemcpyFromSymbol(out, g_fs_dbg, sizeof(g_fs_dbg)); }
#endif
extern "C" void tsd_tune_filter_stack_grid(int ctas) { g_fs_grid = ctas; }

int tsd_filter_stack_tf32(const FilterStackArgs& a, cudaStream_t stream) {
  using namespace tc;
  static_assert(sizeof(FsArgsDev) + sizeof(FilterStackMaps) + 16 <= 4096, "kernel parameter space");
  if (!(a.H == 128 || a.H == 256) || a.num_layers < 1 || a.num_layers > TSD_FS_MAX_LAYERS || a.M_cap < 1024)
    return TSD_ERR_UNSUPPORTED;
  TSD_REQUIRE(a.A && a.len);
  if (reinterpret_cast<uintptr_t>(a.A) & 15) return TSD_ERR_UNSUPPORTED;
  FilterStackMaps maps;
  FsArgsDev d;
  memset(&d, 0, sizeof(d));
  d.M_cap = a.M_cap;
  d.M_ptr = a.M_ptr;
  d.num_layers = a.num_layers;
  d.len = a.len;
  if (!make_tensor_map(&maps.a, a.A, (uint64_t)a.M_cap, (uint64_t)a.H, TC_BM)) return TSD_ERR_UNSUPPORTED;
  for (int l = 0; l < TSD_FS_MAX_LAYERS; ++l) {
    const FilterStackLayer& y = a.layer[l < a.num_layers ? l : 0];
    TSD_REQUIRE(y.W0 && y.W2 && y.out);
    if ((reinterpret_cast<uintptr_t>(y.W0) | reinterpret_cast<uintptr_t>(y.W2) | reinterpret_cast<uintptr_t>(y.out) |
         reinterpret_cast<uintptr_t>(y.b0) | reinterpret_cast<uintptr_t>(y.b2)) & 15)
      return TSD_ERR_UNSUPPORTED;
    if (!make_tensor_map(&maps.w[2 * l], y.W0, (uint64_t)a.H, (uint64_t)a.H, (uint32_t)a.H) ||
        !make_tensor_map(&maps.w[2 * l + 1], y.W2, (uint64_t)a.H, (uint64_t)a.H, (uint32_t)a.H) ||
        !make_tensor_map(&maps.out[l], y.out, (uint64_t)a.M_cap, (uint64_t)a.H, 32, 16))  // one store warp's block
      return TSD_ERR_UNSUPPORTED;
    d.layer[l].b0 = y.b0;
    d.layer[l].b2 = y.b2;
    d.layer[l].cutoff = y.cutoff;
    d.layer[l].smooth = y.smooth;
  }
  return a.H == 256 ? fs_launch<256>(d, maps, a.max_ctas, stream) : fs_launch<128>(d, maps, a.max_ctas, stream);
}
