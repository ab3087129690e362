// The CFConv filter networks of ALL interaction blocks in one kernel.
//
//   filt_l = nn2_l(ssp(nn0_l(edge_attr))) * C(len)        l = 0 .. L-1        (schnet.py:91-98)
//
// Every block's filter network reads the same edge_attr and nothing else, so one CTA keeps a 128-row
// tile and runs the 2 L chained GEMMs back to back.  Round 1/2 launched one two-GEMM kernel per block
// (gemm_chain.cu, k_chain_tf32<0>): main loop -> epilogue -> main loop -> epilogue in strict sequence,
// 7 launches of 22-30 us each on the step's critical path (profiles/r2_kineto_step_tf32.txt).  Here the
// roles run concurrently and only meet at mbarriers:
//
//   warp 16      TMA producer: per layer the tile's 8 edge_attr panels + the 8 panels of W0, then the
//                8 panels of W2, through a ring of {W panel, A panel} slots.  One SM ingests 64 B/clk
//                from L2 (profiles/r2_tma_stream.txt), exactly what the tensor pipe consumes per
//                M128 x N256 x K8 tf32 MMA of streamed W -- the ring must never drain.
//   warp 17      MMA issuer: acc1 = A . W0^T (TMEM columns [0,H)), then acc2 = X . W2^T (columns [H,2H))
//                panel by panel as soon as the epilogue warps have produced the matching QUARTER of X.
//   warps 0..15  epilogues.  epi1: acc1 -> + b0 -> shifted softplus -> TF32 (RNE) -> X in the UMMA K-major
//                SWIZZLE_128B layout (shared memory), handed to the MMA issuer in four column quarters, so
//                the second GEMM trails the activation by a quarter instead of waiting for all of it.
//                epi2: acc2 -> + b2 -> * C(len) -> staged in the same buffer (idle by then) -> TMA store;
//                it overlaps the NEXT layer's first GEMM.
//
// Shared memory (H = 256): X / staging 128 KiB + 2 ring slots x 48 KiB; TMEM: all 512 columns.
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
constexpr int FS_THREADS = (FS_EPI_WARPS + 2) * 32;
constexpr int FS_MAX_SLOTS = 6;

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
  uint64_t acc1_full, acc2_full;
  uint64_t x_ready[4];
  uint32_t tmem_base;
};

// CQ accumulator columns of one row: + bias, activation / scale, into the SW128 panel buffer
template <int CQ, bool SSP>
__device__ __forceinline__ void fs_columns(uint32_t taddr, const float* __restrict__ bias, int c0, float scale, bool valid,
                                           uint8_t* xbuf, int row) {
  float4 b[CQ / 4];
#pragma unroll
  for (int j = 0; j < CQ / 4; ++j)
    b[j] = bias ? __ldg(reinterpret_cast<const float4*>(bias + c0 + 4 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
  uint32_t v[CQ];
  tmem_ld_cols<CQ>(taddr + (uint32_t)c0, v);
  uint8_t* panel = xbuf + (size_t)(c0 >> 5) * TC_A_PANEL_BYTES;
  const int cb = (c0 & 31) >> 2;
#pragma unroll
  for (int j = 0; j < CQ / 4; ++j) {
    float4 o = make_float4(__uint_as_float(v[4 * j + 0]) + b[j].x, __uint_as_float(v[4 * j + 1]) + b[j].y,
                           __uint_as_float(v[4 * j + 2]) + b[j].z, __uint_as_float(v[4 * j + 3]) + b[j].w);
    if (SSP) {
      o = tf32_rn4(make_float4(tc_act<TSD_ACT_SSP>(o.x), tc_act<TSD_ACT_SSP>(o.y), tc_act<TSD_ACT_SSP>(o.z),
                               tc_act<TSD_ACT_SSP>(o.w)));
    } else {
      o = valid ? make_float4(o.x * scale, o.y * scale, o.z * scale, o.w * scale) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    *reinterpret_cast<float4*>(panel + sw128_off(row, cb + j)) = o;
  }
}

// SPLIT: every K step as two N = H/2 MMAs on disjoint accumulator columns (consecutive MMAs into the SAME accumulator
// serialise: 171 instead of 128 clk per M128 x N256 x K8, profiles/r2_umma_small_n.txt) at the price of reading the
// A-side panel twice from shared memory.
template <int H, bool SPLIT>
__device__ __forceinline__ void fs_mma_panel(uint32_t acc, uint32_t a_addr, uint32_t w_addr, uint32_t idesc, bool first) {
  const uint64_t adesc = umma_desc_sw128(a_addr);
  if (!SPLIT) {
    const uint64_t bdesc = umma_desc_sw128(w_addr);
#pragma unroll
    for (int kk = 0; kk < TC_BK / 8; ++kk)
      umma_tf32(acc, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc, (!first || kk != 0) ? 1u : 0u);
  } else {
    const uint64_t bdesc0 = umma_desc_sw128(w_addr);
    const uint64_t bdesc1 = umma_desc_sw128(w_addr + (uint32_t)(H / 2) * 128u);
#pragma unroll
    for (int kk = 0; kk < TC_BK / 8; ++kk) {
      umma_tf32(acc, adesc + (uint64_t)(2 * kk), bdesc0 + (uint64_t)(2 * kk), idesc, (!first || kk != 0) ? 1u : 0u);
      umma_tf32(acc + (uint32_t)(H / 2), adesc + (uint64_t)(2 * kk), bdesc1 + (uint64_t)(2 * kk), idesc,
                (!first || kk != 0) ? 1u : 0u);
    }
  }
}

template <int H, bool SPLIT>
__global__ void __launch_bounds__(FS_THREADS, 1) k_filter_stack(const FsArgsDev p, const __grid_constant__ FilterStackMaps maps,
                                                                int num_slots) {
  constexpr int NKB = H / TC_BK;                 // K panels per GEMM
  constexpr int W_PANEL = H * TC_BK * 4;         // bytes of one W panel: H rows x 128 B
  constexpr int SLOT = W_PANEL + TC_A_PANEL_BYTES;
  constexpr int X_BYTES = NKB * TC_A_PANEL_BYTES;
  constexpr int CQ = H / 16;                     // accumulator columns per epilogue warp and quarter
  constexpr int KBQ = NKB / 4;                   // K panels per quarter of X
  static_assert(NKB % 4 == 0 && (CQ == 8 || CQ == 16), "H must be 128 or 256");
  extern __shared__ uint8_t smem_dyn[];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int M = p.M_ptr ? min(*p.M_ptr, p.M_cap) : p.M_cap;
  const int m0 = blockIdx.x * TC_BM;
  if (m0 >= M) return;
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_dyn + (smem_base - smem_u32(smem_dyn));
  uint8_t* xbuf = smem_gen;
  uint8_t* ring = smem_gen + X_BYTES;
  const uint32_t ring_base = smem_base + X_BYTES;
  FsBars* bars = reinterpret_cast<FsBars*>(ring + (size_t)num_slots * SLOT);
  const int L = p.num_layers;

  if (tid == 0) {
    for (int s = 0; s < num_slots; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    mbar_init(&bars->acc1_full, 1);
    mbar_init(&bars->acc2_full, 1);
    for (int t = 0; t < 4; ++t) mbar_init(&bars->x_ready[t], FS_EPI_THREADS);
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
      for (int l = 0; l < L; ++l) {
        for (int st = 0; st < 2; ++st) {
          const CUtensorMap* wmap = &maps.w[2 * l + st];
          if (g + NKB < 2 * L * NKB) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(wmap + 1)) : "memory");
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
      const uint32_t idesc = umma_idesc_tf32(SPLIT ? H / 2 : H);
      int g = 0;
      for (int l = 0; l < L; ++l) {
        // acc1 = A . W0^T.  acc1 is free: the last quarter of X of the previous layer (waited on below) is signalled
        // after every epilogue thread's last read of it.
        for (int kb = 0; kb < NKB; ++kb, ++g) {
          const int s = g % num_slots, round = g / num_slots;
          mbar_wait(&bars->full[s], (uint32_t)(round & 1));
          tc_fence_after();
          if (kb == 0) FS_STAMP(16 * l + 9);
          const uint32_t slot = ring_base + (uint32_t)(s * SLOT);
          fs_mma_panel<H, SPLIT>(acc1, slot + (uint32_t)W_PANEL, slot, idesc, kb == 0);
          umma_commit(&bars->empty[s]);
        }
        umma_commit(&bars->acc1_full);
        FS_STAMP(16 * l + 10);
        // acc2 = X . W2^T, a quarter of X (KBQ panels) at a time.  acc2 is free: x_ready[0] of THIS layer needs every
        // epilogue thread's arrival, which comes after its reads of the previous layer's acc2.
        for (int kb = 0; kb < NKB; ++kb, ++g) {
          if (kb % KBQ == 0) {
            mbar_wait(&bars->x_ready[kb / KBQ], (uint32_t)(l & 1));
            tc_fence_after();
            if (kb == 0) FS_STAMP(16 * l + 11);
          }
          const int s = g % num_slots, round = g / num_slots;
          mbar_wait(&bars->full[s], (uint32_t)(round & 1));
          tc_fence_after();
          fs_mma_panel<H, SPLIT>(acc2, smem_base + (uint32_t)(kb * TC_A_PANEL_BYTES), ring_base + (uint32_t)(s * SLOT), idesc,
                                 kb == 0);
          umma_commit(&bars->empty[s]);
        }
        umma_commit(&bars->acc2_full);
        FS_STAMP(16 * l + 12);
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue warps
    const int q = warp & 3, cg = warp >> 2;  // TMEM lane quarter (hardware: warp id % 4), column group within a quarter
    const int row = q * 32 + lane;
    const bool valid = m0 + row < M;
    const float len = p.len[min(m0 + row, M - 1)];
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    for (int l = 0; l < L; ++l) {
      const float* const b0 = p.layer[l].b0;
      const float* const b2 = p.layer[l].b2;
      const float cscale = tsd_cutoff_fn(len, p.layer[l].cutoff, p.layer[l].smooth);
      // epi1: X = tf32(ssp(acc1 + b0)).  X is free: the previous layer's second GEMM has retired (acc2_full was waited
      // on in epi2) and its staged output has been read by the TMA store (named barrier below).
      mbar_wait(&bars->acc1_full, (uint32_t)(l & 1));
      tc_fence_after();
      if (tid == 0) FS_STAMP(16 * l + 0);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        fs_columns<CQ, true>(acc1 + lane_addr, b0, t * (H / 4) + cg * CQ, 1.f, true, xbuf, row);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> UMMA reads
        tc_fence_before();
        mbar_arrive(&bars->x_ready[t]);
        if (tid == 0) FS_STAMP(16 * l + 1 + t);
      }
      // epi2: filt = (acc2 + b2) * C(len), staged in X's buffer, stored by TMA
      mbar_wait(&bars->acc2_full, (uint32_t)(l & 1));
      tc_fence_after();
      if (tid == 0) FS_STAMP(16 * l + 5);
#pragma unroll
      for (int t = 0; t < 4; ++t) fs_columns<CQ, false>(acc2 + lane_addr, b2, t * (H / 4) + cg * CQ, cscale, valid, xbuf, row);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> TMA store reads
      if (tid == 0) FS_STAMP(16 * l + 6);
      asm volatile("bar.sync 1, %0;" ::"r"(FS_EPI_THREADS) : "memory");
      if (tid == 0) {
        FS_STAMP(16 * l + 7);
#pragma unroll 1
        for (int kb = 0; kb < NKB; ++kb) tma_store_2d(&maps.out[l], xbuf + (size_t)kb * TC_A_PANEL_BYTES, kb * TC_BK, m0);
        tma_store_commit();
        tma_store_wait_read();
        FS_STAMP(16 * l + 8);
      }
      asm volatile("bar.sync 1, %0;" ::"r"(FS_EPI_THREADS) : "memory");
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == FS_EPI_WARPS + 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)(2 * H)) : "memory");
  }
}

template <int H, bool SPLIT>
int fs_launch(const FsArgsDev& d, const FilterStackMaps& maps, cudaStream_t stream) {
  constexpr int SLOT = H * TC_BK * 4 + TC_A_PANEL_BYTES, X_BYTES = (H / TC_BK) * TC_A_PANEL_BYTES;
  const int budget = 227 * 1024 - 1024 - X_BYTES - 256;  // alignment slack, barriers
  int slots = budget / SLOT;
  if (slots > FS_MAX_SLOTS) slots = FS_MAX_SLOTS;
  if (slots < 2) return TSD_ERR_UNSUPPORTED;
  const size_t smem = 1024 + (size_t)X_BYTES + (size_t)slots * SLOT + 256;
  static size_t attr_smem = 0;  // per instantiation
  if (smem > attr_smem) {
    TSD_CUDA(cudaFuncSetAttribute(k_filter_stack<H, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  k_filter_stack<H, SPLIT><<<dim3(tsd_ceil_div(d.M_cap, TC_BM)), dim3(FS_THREADS), smem, stream>>>(d, maps, slots);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

}  // namespace

#ifdef TSD_FS_DBG
extern "C" void tsd_fs_dbg_read(unsigned long long* out) { cudaMemcpyFromSymbol(out, g_fs_dbg, sizeof(g_fs_dbg)); }
#endif
static int g_fs_split_mma = 0;
// tuning hook of profiles/scripts (not part of the C-ABI header): 1 = two N/2 MMAs per K step
extern "C" void tsd_tune_filter_stack_mma(int split) { g_fs_split_mma = split; }

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
        !make_tensor_map(&maps.out[l], y.out, (uint64_t)a.M_cap, (uint64_t)a.H, TC_BM))
      return TSD_ERR_UNSUPPORTED;
    d.layer[l].b0 = y.b0;
    d.layer[l].b2 = y.b2;
    d.layer[l].cutoff = y.cutoff;
    d.layer[l].smooth = y.smooth;
  }
  if (a.H == 256) return g_fs_split_mma ? fs_launch<256, true>(d, maps, stream) : fs_launch<256, false>(d, maps, stream);
  return g_fs_split_mma ? fs_launch<128, true>(d, maps, stream) : fs_launch<128, false>(d, maps, stream);
}
