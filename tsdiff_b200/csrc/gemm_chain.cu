// Chained linear layers on the tensor cores: up to three H x H nn.Linear layers applied to a
// 128-row tile without the intermediates ever leaving the SM.
//
//   stage 0 : acc0 = A . W0^T            A, W0 panels by TMA (3-slot weight ring), tcgen05.mma kind::tf32
//   epi   0 : x = act0(acc0 + b0)        TMEM -> registers -> TF32-rounded, written in the UMMA
//                                        K-major SWIZZLE_128B layout into `abuf` (the next A operand)
//   stage 1 : acc1 = abuf . W1^T         only W1 streams through the ring
//   epi   1 : y = act1(acc1 + b1) [* C(len)] [+ residual] -> global and/or abuf
//   stage 2 : acc0 = abuf . W2^T ; epi 2 -> global
//
// Used for (a) the SchNet filter network  W = nn2(ssp(nn0(edge_attr))) * C(len)
//              (schnet.py:91-98; saves the (E,H) intermediate's round trip), and
//          (b) the node update            h' = h + lin(ssp(lin2(agg))),  x1' = lin1_next(h')
//              (schnet.py:103-104,124-128 + the next block's :101; 3 launches -> 1).
// Shared memory: abuf H/32 panels x 16 KiB + ring of 3 W panels (H*128 B each): 224 KiB at H = 256;
// TMEM: two H-column accumulators used alternately.  One CTA per SM.
#include <stdio.h>
#include <stdlib.h>

#include "tc_common.cuh"



namespace {
using namespace tc;

constexpr int CH_RING = 3;  // W-panel slots; the A operand of every stage lives in abuf
constexpr int CH_THREADS = 512;  // 16 warps: the epilogues are latency bound, 4 warps per TMEM lane quarter

// The chains the encoders need, as compile-time stage tables.  (A single kernel that decided
// activation / residual / store / scale per stage at run time was 24.5k SASS instructions with
// the executed epilogues scattered over it: 37 % of the last epilogue's stall samples were
// instruction-fetch misses.  One instantiation per chain keeps each kernel's code compact.)
enum { CH_FILTER = 0, CH_NODE3 = 1, CH_NODE2 = 2 };
struct StageCfg {
  int act;
  bool scale, resid, store;
};
__host__ __device__ constexpr int ch_num_stages(int kind) { return kind == CH_NODE3 ? 3 : 2; }
__host__ __device__ constexpr StageCfg ch_stage(int kind, int s) {
  if (s == 0) return {TSD_ACT_SSP, false, false, false};                   // nn0 / lin2 + shifted softplus
  if (kind == CH_FILTER) return {TSD_ACT_NONE, true, false, true};         // nn2, * C(len) -> filter
  if (s == 1) return {TSD_ACT_NONE, false, true, true};                    // lin, + h -> h'
  return {TSD_ACT_NONE, false, false, true};                               // next block's lin1 -> x1
}

// One stage's epilogue.  VIA_TILE: the result (plus residual) goes to global memory -> transpose
// each warp's 32x32 chunk through a shared-memory tile so global accesses are 128-byte coalesced.
// FEEDS_NEXT: the result is also the next stage's A operand -> TF32-rounded into `abuf` (UMMA
// K-major SWIZZLE_128B panels).
template <int ACT, bool RESID, bool STORE, bool FEEDS_NEXT>
__device__ __forceinline__ void chain_epilogue(const ChainArgs& p, const ChainStage& st, const float* sbias,
                                               float cscale, uint32_t tmem_acc, uint8_t* abuf, float* tile, int m0,
                                               int M, int warp, int lane) {
  constexpr bool VIA_TILE = RESID || STORE;
  const int H = p.H;
  const int q = warp & 3, half = warp >> 2;  // TMEM lane quarter, column slice (CH_THREADS/128 slices)
  const int row = q * 32 + lane;
  const int cols_per_half = H / (CH_THREADS / 128);
  constexpr int TLD = 36;
  if (RESID) {
    // the later chunks' residual rows: pull them into L1 now, their loads are issued only after the
    // previous chunk's stores (one L2 round trip per chunk otherwise)
    for (int cc = 32; cc < cols_per_half; cc += 32) {
      const int c0 = half * cols_per_half + cc;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float* ptr = st.residual + (size_t)min(m0 + q * 32 + i * 4 + (lane >> 3), M - 1) * H + c0 + 4 * (lane & 7);
        asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr));
      }
    }
  }
  for (int cc = 0; cc < cols_per_half; cc += 32) {
    const int c0 = half * cols_per_half + cc;
    // residual rows first: their L2 latency hides behind the TMEM load and the activation
    // (branch-free, clamped rows: a per-row `if (mr < M)` serialised 8 L2 round trips per chunk)
    float4 rs[8];
    if (RESID) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        rs[i] = *reinterpret_cast<const float4*>(st.residual + (size_t)min(m0 + q * 32 + i * 4 + (lane >> 3), M - 1) * H +
                                                 c0 + 4 * (lane & 7));
    }
    uint32_t v[32];
    tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
    float4 o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = *reinterpret_cast<const float4*>(sbias + c0 + 4 * j);  // staged in shared memory per stage
      o[j].x = tc_act<ACT>(__uint_as_float(v[4 * j + 0]) + b.x) * cscale;
      o[j].y = tc_act<ACT>(__uint_as_float(v[4 * j + 1]) + b.y) * cscale;
      o[j].z = tc_act<ACT>(__uint_as_float(v[4 * j + 2]) + b.z) * cscale;
      o[j].w = tc_act<ACT>(__uint_as_float(v[4 * j + 3]) + b.w) * cscale;
    }
    if (!VIA_TILE) {
      // thread = row: panel c0/32, 16-byte chunk j
      uint8_t* panel = abuf + (size_t)(c0 >> 5) * TC_A_PANEL_BYTES;
#pragma unroll
      for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(panel + sw128_off(row, j)) = tf32_rn4(o[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(tile + lane * TLD + 4 * j) = o[j];
      __syncwarp();
      const int cg = lane & 7;  // 8 lanes cover the 32 columns of one row: a full 128-byte line
      uint8_t* panel = abuf + (size_t)(c0 >> 5) * TC_A_PANEL_BYTES;
      float4 x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = *reinterpret_cast<const float4*>(tile + (i * 4 + (lane >> 3)) * TLD + 4 * cg);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = i * 4 + (lane >> 3);
        const int mr = m0 + q * 32 + r;
        if (RESID) {
          x[i].x += rs[i].x; x[i].y += rs[i].y; x[i].z += rs[i].z; x[i].w += rs[i].w;
        }
        if (STORE && mr < M) *reinterpret_cast<float4*>(st.store + (size_t)mr * H + c0 + 4 * cg) = x[i];
        if (FEEDS_NEXT) *reinterpret_cast<float4*>(panel + sw128_off(q * 32 + r, cg)) = tf32_rn4(x[i]);
      }
      __syncwarp();
    }
  }
}

struct ChainMaps {
  CUtensorMap a;
  CUtensorMap w[3];
};

struct ChainCtx {
  uint8_t* abuf;
  uint8_t* ring;
  uint32_t smem_base, ring_base, w_panel_bytes, tmem, idesc;
  uint64_t *bar_full, *bar_empty, *bar_accum;
  float* s_bias;
  int num_kb, m0, M, tid, warp, lane;
  int issued;  // TMA thread: W panels issued so far (global index over all stages)
};

// ring bookkeeping: panel g (global over all stages) lives in slot g % CH_RING, round g / CH_RING
__device__ __forceinline__ void chain_tma_issue(ChainCtx& c, const ChainMaps& maps, int stage, int kb) {
  const int g = stage * c.num_kb + kb;
  const int s = g % CH_RING, round = g / CH_RING;
  if (round > 0) mbar_wait(&c.bar_empty[s], (uint32_t)((round - 1) & 1));
  uint8_t* slot = c.ring + (size_t)s * c.w_panel_bytes;
  mbar_arrive_expect_tx(&c.bar_full[s], c.w_panel_bytes + (stage == 0 ? TC_A_PANEL_BYTES : 0));
  if (stage == 0) tma_load_2d(c.abuf + (size_t)kb * TC_A_PANEL_BYTES, &maps.a, &c.bar_full[s], kb * TC_BK, c.m0);
  tma_load_2d(slot, &maps.w[stage], &c.bar_full[s], kb * TC_BK, 0);
}

template <int KIND, int STAGE>
__device__ __forceinline__ void chain_stage_run(const ChainArgs& p, const ChainMaps& maps, ChainCtx& c) {
  constexpr int NS = ch_num_stages(KIND);
  constexpr StageCfg cfg = ch_stage(KIND, STAGE);
  constexpr bool last = STAGE == NS - 1;
  constexpr bool via_tile = cfg.resid || cfg.store;
  const ChainStage& st = p.st[STAGE];
  const int H = p.H, num_kb = c.num_kb;
  // this stage's bias -> shared memory; the barrier orders it against the previous epilogue's
  // reads and this epilogue's reads (all 512 threads arrive at once: the role loops start after it)
  if (c.tid < H) c.s_bias[c.tid] = st.bias ? st.bias[c.tid] : 0.f;
  float cscale = 1.f;  // loaded before the accumulator wait: the latency hides behind the main loop
  if (cfg.scale) {
    const int m = c.m0 + (c.warp & 3) * 32 + c.lane;
    cscale = tsd_cutoff_fn(st.scale_len[min(m, c.M - 1)], st.cutoff, st.smooth);
  }
  __syncthreads();
  if (c.warp == 0 && c.lane == 0) {
    for (int g = max(c.issued, STAGE * num_kb); g < (STAGE + 1) * num_kb; ++g) chain_tma_issue(c, maps, STAGE, g - STAGE * num_kb);
    c.issued = (STAGE + 1) * num_kb;
  } else if (c.warp == 1 && c.lane == 0) {
    const uint32_t acc = c.tmem + (uint32_t)((STAGE & 1) * H);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int g = STAGE * num_kb + kb;
      const int s = g % CH_RING, round = g / CH_RING;
      mbar_wait(&c.bar_full[s], (uint32_t)(round & 1));
      tc_fence_after();
      const uint32_t slot = c.ring_base + (uint32_t)s * c.w_panel_bytes;
      const uint64_t adesc = umma_desc_sw128(c.smem_base + (uint32_t)kb * TC_A_PANEL_BYTES);
      const uint64_t bdesc = umma_desc_sw128(slot);
      // (two N = H/2 MMAs per K step on disjoint accumulator columns issue faster in isolation -- 128 instead of
      // 171 clk per K step, profiles/r2_umma_small_n.txt -- but read the A panel twice: 10 us slower per Langevin step)
#pragma unroll
      for (int kk = 0; kk < TC_BK / 8; ++kk)
        umma_tf32(acc, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), c.idesc, (kb | kk) != 0 ? 1u : 0u);
      umma_commit(&c.bar_empty[s]);
      if (kb == num_kb - 1) umma_commit(&c.bar_accum[STAGE]);
    }
  }
  __syncwarp();
  mbar_wait(&c.bar_accum[STAGE], 0);
  tc_fence_after();
  if (!last && !via_tile && c.warp == 0 && c.lane == 0) {
    // every ring slot is free now (this stage's MMAs retired) and this epilogue does not need
    // the ring as a transpose tile: start the next stage's weights so they land meanwhile
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.w[STAGE + 1])) : "memory");
    for (int kb = 0; kb < CH_RING && kb < num_kb; ++kb) chain_tma_issue(c, maps, STAGE + 1, kb);
    c.issued = (STAGE + 1) * num_kb + min(CH_RING, num_kb);
  }
  __syncwarp();
  const uint32_t acc = c.tmem + (uint32_t)((STAGE & 1) * H);
  float* tile = reinterpret_cast<float*>(c.ring) + c.warp * (32 * 36);  // the ring is idle whenever via_tile
  chain_epilogue<cfg.act, cfg.resid, cfg.store, !last>(p, st, c.s_bias, cscale, acc, c.abuf, tile, c.m0, c.M, c.warp, c.lane);
  if (!last) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // abuf writes -> UMMA
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}

template <int KIND>
__global__ void __launch_bounds__(CH_THREADS, 1) k_chain_tf32(const ChainArgs p, const __grid_constant__ ChainMaps maps,
                                                              int tmem_cols) {
  extern __shared__ uint8_t smem_dyn[];
  __shared__ uint64_t bar_full[CH_RING];
  __shared__ uint64_t bar_empty[CH_RING];
  __shared__ uint64_t bar_accum[3];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_bias[256];  // bias of the stage in flight (224 KiB + this must stay < 227 KiB)

  ChainCtx c;
  c.tid = threadIdx.x, c.warp = c.tid >> 5, c.lane = c.tid & 31;
  // predecessor-independent setup first (see launch_pdl): barriers and descriptor prefetch
  if (c.tid == 0) {
    for (int s = 0; s < CH_RING; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    for (int s = 0; s < 3; ++s) mbar_init(&bar_accum[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.w[0])) : "memory");
  }
  pdl_wait();
  pdl_trigger();
  c.M = p.M_ptr ? min(*p.M_ptr, p.M_cap) : p.M_cap;
  c.m0 = blockIdx.x * TC_BM;
  if (c.m0 >= c.M) return;
  const int H = p.H;
  c.num_kb = H / TC_BK;
  c.w_panel_bytes = (uint32_t)H * TC_BK * 4;
  c.smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_dyn + (c.smem_base - smem_u32(smem_dyn));
  // layout: [abuf: num_kb A panels (stage 0: loaded by TMA; later stages: written by the epilogue)]
  //         [ring: CH_RING slots of one W panel]
  c.abuf = smem_gen;
  const uint32_t abuf_bytes = (uint32_t)c.num_kb * TC_A_PANEL_BYTES;
  c.ring = smem_gen + abuf_bytes;
  c.ring_base = c.smem_base + abuf_bytes;
  c.bar_full = bar_full, c.bar_empty = bar_empty, c.bar_accum = bar_accum, c.s_bias = s_bias;
  c.issued = 0;

  if (c.warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"((uint32_t)tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.tmem = tmem_base_s;
  c.idesc = umma_idesc_tf32(H);

  chain_stage_run<KIND, 0>(p, maps, c);
  chain_stage_run<KIND, 1>(p, maps, c);
  if (ch_num_stages(KIND) == 3) chain_stage_run<KIND, ch_num_stages(KIND) == 3 ? 2 : 1>(p, maps, c);
  if (c.warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(c.tmem), "r"((uint32_t)tmem_cols)
                 : "memory");
  }
}

// does the runtime description match the compile-time stage table of `kind`?
bool chain_matches(const ChainArgs& c, int kind) {
  if (c.num_stages != ch_num_stages(kind)) return false;
  for (int s = 0; s < c.num_stages; ++s) {
    const StageCfg cfg = ch_stage(kind, s);
    const ChainStage& st = c.st[s];
    if (st.act != cfg.act || (st.scale_len != nullptr) != cfg.scale || (st.residual != nullptr) != cfg.resid ||
        (st.store != nullptr) != cfg.store)
      return false;
  }
  return true;
}

template <int KIND>
int chain_launch(const ChainArgs& cc, const ChainMaps& maps, int tmem_cols, size_t smem, cudaStream_t stream) {
  static size_t attr_smem = 0;  // static + dynamic shared memory must stay <= 227 KiB: ask for exactly what is used
  if (smem > attr_smem) {
    TSD_CUDA(cudaFuncSetAttribute(k_chain_tf32<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  TSD_CUDA(launch_pdl(k_chain_tf32<KIND>, dim3(tsd_ceil_div(cc.M_cap, TC_BM)), dim3(CH_THREADS), smem, stream, cc, maps,
                      tmem_cols));
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

}  // namespace

int tsd_chain_tf32(const ChainArgs& c, cudaStream_t stream) {
  const float* A = c.A;
  using namespace tc;
  if (!(c.H == 128 || c.H == 256) || c.num_stages < 2 || c.num_stages > 3) return TSD_ERR_UNSUPPORTED;
  if (c.M_cap < 1024) return TSD_ERR_UNSUPPORTED;
  TSD_REQUIRE(A && c.st[c.num_stages - 1].store);
  if (reinterpret_cast<uintptr_t>(A) & 15) return TSD_ERR_UNSUPPORTED;
  ChainMaps maps;
  if (!make_tensor_map(&maps.a, A, (uint64_t)c.M_cap, (uint64_t)c.H, TC_BM)) return TSD_ERR_UNSUPPORTED;
  for (int s = 0; s < 3; ++s) {
    const float* w = c.st[s < c.num_stages ? s : 0].W;
    TSD_REQUIRE(w);
    if (reinterpret_cast<uintptr_t>(w) & 15) return TSD_ERR_UNSUPPORTED;
    if (!make_tensor_map(&maps.w[s], w, (uint64_t)c.H, (uint64_t)c.H, (uint32_t)c.H)) return TSD_ERR_UNSUPPORTED;
  }
  const int num_kb = c.H / TC_BK;
  size_t ring_bytes = (size_t)CH_RING * ((size_t)c.H * TC_BK * 4);
  const size_t tile_bytes = (size_t)(CH_THREADS / 32) * 32 * 36 * sizeof(float);  // epilogue transpose tiles live in the ring
  if (ring_bytes < tile_bytes) ring_bytes = tile_bytes;
  const size_t smem = (size_t)num_kb * TC_A_PANEL_BYTES + ring_bytes + 1024;
  const int tmem_cols = 2 * c.H;  // 512 or 256: both powers of two
  int rc = TSD_ERR_UNSUPPORTED;
  const int kind = chain_matches(c, CH_FILTER) ? CH_FILTER : chain_matches(c, CH_NODE3) ? CH_NODE3 : chain_matches(c, CH_NODE2) ? CH_NODE2 : -1;
  if (kind == CH_FILTER) rc = chain_launch<CH_FILTER>(c, maps, tmem_cols, smem, stream);
  else if (kind == CH_NODE3) rc = chain_launch<CH_NODE3>(c, maps, tmem_cols, smem, stream);
  else if (kind == CH_NODE2) rc = chain_launch<CH_NODE2>(c, maps, tmem_cols, smem, stream);
  return rc;
}
