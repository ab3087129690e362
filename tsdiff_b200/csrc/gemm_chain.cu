// Chained linear layers on the tensor cores: up to three H x H nn.Linear layers applied to a
// 128-row tile without the intermediates ever leaving the SM.
//
//   stage 0 : acc0 = A . W0^T            A, W0 panels by TMA (2-slot ring), tcgen05.mma kind::tf32
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

// One stage's epilogue.  VIA_TILE: the result (plus residual) goes to global memory and/or the
// residual must be read -> transpose each warp's 32x32 chunk through a shared-memory tile so
// global accesses are 128-byte coalesced.  FEEDS_NEXT: the result is also the next stage's A
// operand -> TF32-rounded into `abuf` (UMMA K-major SWIZZLE_128B panels).
template <int ACT, bool VIA_TILE, bool FEEDS_NEXT>
__device__ __forceinline__ void chain_epilogue(const ChainArgs& p, const ChainStage& st, const float* sbias,
                                               uint32_t tmem_acc, uint8_t* abuf, float* tile, int m0, int M, int warp,
                                               int lane) {
  const int H = p.H;
  const int q = warp & 3, half = warp >> 2;  // TMEM lane quarter, column slice (CH_THREADS/128 slices)
  const int row = q * 32 + lane, m = m0 + row;
  const int cols_per_half = H / (CH_THREADS / 128);
  constexpr int TLD = 36;
  float cscale = 1.f;
  if (st.scale_len && m < M) cscale = tsd_cutoff_fn(st.scale_len[m], st.cutoff, st.smooth);
  for (int cc = 0; cc < cols_per_half; cc += 32) {
    const int c0 = half * cols_per_half + cc;
    // residual rows first: their L2 latency hides behind the TMEM load and the activation
    float4 rs[8];
    if (VIA_TILE) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        rs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (st.residual)
          rs[i] = *reinterpret_cast<const float4*>(st.residual + (size_t)min(m0 + q * 32 + i * 4 + (lane >> 3), M - 1) * H +
                                                   c0 + 4 * (lane & 7));
      }
    }
    uint32_t v[32];
    tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
    float4 o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = *reinterpret_cast<const float4*>(sbias + c0 + 4 * j);  // staged in shared memory at kernel start
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
      // (residual rows were loaded branch-free at the top of the chunk: a per-row `if (mr < M)`
      // around them serialised 8 L2 round trips per chunk; only the STORE is predicated)
      float4 x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = *reinterpret_cast<const float4*>(tile + (i * 4 + (lane >> 3)) * TLD + 4 * cg);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = i * 4 + (lane >> 3);
        const int mr = m0 + q * 32 + r;
        x[i].x += rs[i].x; x[i].y += rs[i].y; x[i].z += rs[i].z; x[i].w += rs[i].w;
        if (st.store && mr < M) *reinterpret_cast<float4*>(st.store + (size_t)mr * H + c0 + 4 * cg) = x[i];
        if (FEEDS_NEXT) *reinterpret_cast<float4*>(panel + sw128_off(q * 32 + r, cg)) = tf32_rn4(x[i]);
      }
      __syncwarp();
    }
  }
}

template <bool VIA_TILE, bool FEEDS_NEXT>
__device__ __forceinline__ void chain_epilogue_act(const ChainArgs& p, const ChainStage& st, const float* sbias,
                                                   uint32_t acc, uint8_t* abuf, float* tile, int m0, int M, int warp,
                                                   int lane) {
  switch (st.act) {  // uniform
    case TSD_ACT_SSP: chain_epilogue<TSD_ACT_SSP, VIA_TILE, FEEDS_NEXT>(p, st, sbias, acc, abuf, tile, m0, M, warp, lane); break;
    case TSD_ACT_RELU: chain_epilogue<TSD_ACT_RELU, VIA_TILE, FEEDS_NEXT>(p, st, sbias, acc, abuf, tile, m0, M, warp, lane); break;
    case TSD_ACT_SWISH: chain_epilogue<TSD_ACT_SWISH, VIA_TILE, FEEDS_NEXT>(p, st, sbias, acc, abuf, tile, m0, M, warp, lane); break;
    default: chain_epilogue<TSD_ACT_NONE, VIA_TILE, FEEDS_NEXT>(p, st, sbias, acc, abuf, tile, m0, M, warp, lane); break;
  }
}

struct ChainMaps {
  CUtensorMap a;
  CUtensorMap w[3];
};

__global__ void __launch_bounds__(CH_THREADS, 1) k_chain_tf32(const ChainArgs p, const __grid_constant__ ChainMaps maps,
                                                              int tmem_cols) {
  extern __shared__ uint8_t smem_dyn[];
  __shared__ uint64_t bar_full[CH_RING];
  __shared__ uint64_t bar_empty[CH_RING];
  __shared__ uint64_t bar_accum[3];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_bias[256];  // bias of the stage in flight (224 KiB + this must stay < 227 KiB)

  if (threadIdx.x == 0) TC_STAMP(0);
  const int M = p.M_ptr ? min(*p.M_ptr, p.M_cap) : p.M_cap;
  const int m0 = blockIdx.x * TC_BM;
  if (m0 >= M) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = p.H, ns = p.num_stages;
  const int num_kb = H / TC_BK;
  const uint32_t w_panel_bytes = (uint32_t)H * TC_BK * 4;
  const uint32_t slot_bytes = w_panel_bytes;
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_dyn + (smem_base - smem_u32(smem_dyn));
  // layout: [abuf: num_kb A panels (stage 0: loaded by TMA; later stages: written by the epilogue)]
  //         [ring: CH_RING slots of one W panel]
  uint8_t* abuf = smem_gen;
  const uint32_t abuf_bytes = (uint32_t)num_kb * TC_A_PANEL_BYTES;
  uint8_t* ring = smem_gen + abuf_bytes;
  const uint32_t ring_base = smem_base + abuf_bytes;

  if (tid == 0) {
    for (int s = 0; s < CH_RING; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    for (int s = 0; s < 3; ++s) mbar_init(&bar_accum[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.w[0])) : "memory");
  }
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
  const uint32_t idesc = umma_idesc_tf32(H);

  // ring bookkeeping: panel g (global over all stages) lives in slot g % CH_RING, round g / CH_RING
  int issued = 0;  // TMA thread: panels issued so far (global index)
  auto tma_issue = [&](int stage, int kb) {
    const int g = stage * num_kb + kb;
    const int s = g % CH_RING, round = g / CH_RING;
    if (round > 0) mbar_wait(&bar_empty[s], (uint32_t)((round - 1) & 1));
    uint8_t* slot = ring + (size_t)s * slot_bytes;
    mbar_arrive_expect_tx(&bar_full[s], w_panel_bytes + (stage == 0 ? TC_A_PANEL_BYTES : 0));
    if (stage == 0) tma_load_2d(abuf + (size_t)kb * TC_A_PANEL_BYTES, &maps.a, &bar_full[s], kb * TC_BK, m0);
    tma_load_2d(slot, &maps.w[stage], &bar_full[s], kb * TC_BK, 0);
  };

#pragma unroll
  for (int stage = 0; stage < 3; ++stage) {
    if (stage >= ns) break;
    const ChainStage& st = p.st[stage];
    const bool last = stage == ns - 1;
    const bool via_tile = last || st.store != nullptr || st.residual != nullptr;
    // this stage's bias -> shared memory (read by the epilogue; the previous epilogue ended with a CTA barrier)
    if (tid < H) s_bias[tid] = st.bias ? st.bias[tid] : 0.f;
    if (warp == 0 && lane == 0) {
      for (int g = max(issued, stage * num_kb); g < (stage + 1) * num_kb; ++g) tma_issue(stage, g - stage * num_kb);
      issued = (stage + 1) * num_kb;
    } else if (warp == 1 && lane == 0) {
      const uint32_t acc = tmem + (uint32_t)((stage & 1) * H);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int g = stage * num_kb + kb;
        const int s = g % CH_RING, round = g / CH_RING;
        mbar_wait(&bar_full[s], (uint32_t)(round & 1));
        tc_fence_after();
        const uint32_t slot = ring_base + (uint32_t)s * slot_bytes;
        const uint64_t adesc = umma_desc_sw128(smem_base + (uint32_t)kb * TC_A_PANEL_BYTES);
        const uint64_t bdesc = umma_desc_sw128(slot);
#pragma unroll
        for (int kk = 0; kk < TC_BK / 8; ++kk)
          umma_tf32(acc, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc, (kb | kk) != 0 ? 1u : 0u);
        umma_commit(&bar_empty[s]);
        if (kb == num_kb - 1) umma_commit(&bar_accum[stage]);
      }
    }
    __syncwarp();
    mbar_wait(&bar_accum[stage], 0);
    tc_fence_after();
    if (tid == 0) TC_STAMP(1 + 2 * stage);
    if (!last && !via_tile && warp == 0 && lane == 0) {
      // every ring slot is free now (this stage's MMAs retired) and this epilogue does not need
      // the ring as a transpose tile: start the next stage's weights so they land meanwhile
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.w[stage + 1])) : "memory");
      for (int kb = 0; kb < CH_RING && kb < num_kb; ++kb) tma_issue(stage + 1, kb);
      issued = (stage + 1) * num_kb + min(CH_RING, num_kb);
    }
    __syncwarp();
    const uint32_t acc = tmem + (uint32_t)((stage & 1) * H);
    float* tile = reinterpret_cast<float*>(ring) + warp * (32 * 36);  // ring is idle whenever via_tile
    const float* sb = s_bias;
    if (!via_tile) chain_epilogue_act<false, true>(p, st, sb, acc, abuf, tile, m0, M, warp, lane);
    else if (!last) chain_epilogue_act<true, true>(p, st, sb, acc, abuf, tile, m0, M, warp, lane);
    else chain_epilogue_act<true, false>(p, st, sb, acc, abuf, tile, m0, M, warp, lane);
    if (!last) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // abuf writes -> UMMA
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) TC_STAMP(2 + 2 * stage);
  }
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)tmem_cols)
                 : "memory");
  }
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
  static size_t attr_smem = 0;  // static + dynamic shared memory must stay <= 227 KiB: ask for exactly what is used
  if (smem > attr_smem) {
    TSD_CUDA(cudaFuncSetAttribute(k_chain_tf32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  const int tmem_cols = 2 * c.H;  // 512 or 256: both powers of two
  static unsigned long long* dbg = nullptr;
  static bool dbg_checked = false;
  if (!dbg_checked) {
    dbg_checked = true;
    const char* e = getenv("TSD_GEMM_DBG");
    if (e && e[0] == '1') cudaMalloc(&dbg, 64 * sizeof(unsigned long long));
  }
  ChainArgs cc = c;
  cc.dbg = dbg;
  if (dbg) cudaMemsetAsync(dbg, 0, 64 * sizeof(unsigned long long), stream);
  k_chain_tf32<<<tsd_ceil_div(c.M_cap, TC_BM), CH_THREADS, smem, stream>>>(cc, maps, tmem_cols);
  TSD_LAUNCH_CHECK();
  if (dbg) {
    unsigned long long hb[64];
    cudaStreamSynchronize(stream);
    cudaMemcpy(hb, dbg, sizeof(hb), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[chain dbg] M_cap=%d H=%d stages=%d |", c.M_cap, c.H, c.num_stages);
    for (int i = 1; i <= 2 * c.num_stages; ++i) fprintf(stderr, " %s%llu", (i & 1) ? "acc " : "epi ", hb[i] - hb[0]);
    fprintf(stderr, "\n");
  }
  return TSD_OK;
}
