// K7: eq_transform + clip_norm + Langevin update + NaN flag + per-graph centring in ONE
// launch, one CTA per reaction graph, one thread per atom.  Reads the per-step scalars
// from a device table indexed by a device step counter, so the same captured CUDA graph is
// replayed for every step with no host round-trip.
//
// Replaces models/geometry.py:22-30 (eq_transform), models/sampler.py:208-254 (LD branch,
// NaN check, center_pos, clip_pos, trajectory append), :260-268 and the two-channel
// variant of models/epsnet/dualenc.py:827-849,946-965.
//
// Per atom the two scatter_add sums of eq_transform are accumulated sequentially in edge
// order (out-edges by col, in-edges by row) -- the association order of a sequential
// scatter over the row-major sorted edge list -- so results are deterministic.
#include "common.cuh"

// ---------------------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ uint4 tsd_philox4x32_10(uint4 c, uint2 k) {
  const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    unsigned hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    unsigned hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

// three standard normals for (seed, step, global atom id): counter = (atom_lo, atom_hi, step, 0),
// key = (seed_lo, seed_hi); Box-Muller on 24-bit uniforms in (0, 1).
__device__ __forceinline__ float3 tsd_philox_normal3(uint64_t seed, int step, int64_t atom) {
  uint4 r = tsd_philox4x32_10(make_uint4((unsigned)atom, (unsigned)((uint64_t)atom >> 32), (unsigned)step, 0u),
                              make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
  const float s = 1.0f / 16777216.0f;
  float u0 = ((float)(r.x >> 8) + 0.5f) * s, u1 = ((float)(r.y >> 8) + 0.5f) * s;
  float u2 = ((float)(r.z >> 8) + 0.5f) * s, u3 = ((float)(r.w >> 8) + 0.5f) * s;
  const float two_pi = 6.283185307179586f;
  float ra = sqrtf(-2.0f * logf(u0)), rb = sqrtf(-2.0f * logf(u2));
  return make_float3(ra * cosf(two_pi * u1), ra * sinf(two_pi * u1), rb * cosf(two_pi * u3));
}

__global__ void k_philox_normal(int num_nodes, uint64_t seed, int step, int64_t atom_offset, float* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num_nodes) return;
  float3 z = tsd_philox_normal3(seed, step, atom_offset + i);
  out[3 * i] = z.x;
  out[3 * i + 1] = z.y;
  out[3 * i + 2] = z.z;
}

extern "C" int tsd_philox_normal(int32_t num_nodes, uint64_t seed, int32_t step, int64_t atom_offset, float* out,
                                 tsd_stream_t stream) {
  TSD_REQUIRE(out);
  if (num_nodes == 0) return TSD_OK;
  k_philox_normal<<<tsd_ceil_div(num_nodes, 256), 256, 0, tsd_cu(stream)>>>(num_nodes, seed, step, atom_offset, out);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

// ------------------------------------------------------------------------------ eq_transform
__device__ __forceinline__ bool tsd_edge_selected(const tsd_score_channel_t& ch, int e) {
  if (ch.mask_mode == 0 || ch.mask == nullptr) return true;
  int v = ch.mask[e];
  return ch.mask_mode == 1 ? (v != 0) : (v == 0);
}

// score of atom `i` (global index): sum_{row=i} u*s - sum_{col=i} u*s, u = (p_row - p_col)/len.
// spos holds the graph's positions (local index = global - n0).
__device__ float3 tsd_node_score(const tsd_score_channel_t& ch, const tsd_edges_t& e, const float* spos, int n0, int i,
                                 float inv_div, float inv_mul = 1.0f) {
  const float px = spos[3 * (i - n0)], py = spos[3 * (i - n0) + 1], pz = spos[3 * (i - n0) + 2];
  float ax = 0.f, ay = 0.f, az = 0.f, bx = 0.f, by = 0.f, bz = 0.f;
  for (int k = e.row_ptr[i]; k < e.row_ptr[i + 1]; ++k) {
    if (!tsd_edge_selected(ch, k)) continue;
    int c = e.col[k] - n0;
    float inv_len = __fdiv_rn(1.0f, e.length[k]);
    float s = __fdiv_rn(__fmul_rn(ch.inv[ch.inv_index ? ch.inv_index[k] : k], inv_mul), inv_div);
    ax = __fadd_rn(ax, __fmul_rn(__fmul_rn(inv_len, __fsub_rn(px, spos[3 * c])), s));
    ay = __fadd_rn(ay, __fmul_rn(__fmul_rn(inv_len, __fsub_rn(py, spos[3 * c + 1])), s));
    az = __fadd_rn(az, __fmul_rn(__fmul_rn(inv_len, __fsub_rn(pz, spos[3 * c + 2])), s));
  }
  for (int k = e.in_ptr[i]; k < e.in_ptr[i + 1]; ++k) {
    int id = e.in_eid[k];
    if (!tsd_edge_selected(ch, id)) continue;
    int r = e.row[id] - n0;
    float inv_len = __fdiv_rn(1.0f, e.length[id]);
    float s = __fdiv_rn(__fmul_rn(ch.inv[ch.inv_index ? ch.inv_index[id] : id], inv_mul), inv_div);
    bx = __fadd_rn(bx, __fmul_rn(-__fmul_rn(inv_len, __fsub_rn(spos[3 * r], px)), s));
    by = __fadd_rn(by, __fmul_rn(-__fmul_rn(inv_len, __fsub_rn(spos[3 * r + 1], py)), s));
    bz = __fadd_rn(bz, __fmul_rn(-__fmul_rn(inv_len, __fsub_rn(spos[3 * r + 2], pz)), s));
  }
  return make_float3(__fadd_rn(ax, bx), __fadd_rn(ay, by), __fadd_rn(az, bz));
}

// sampler.py:265-268
__device__ __forceinline__ float3 tsd_clip_norm(float3 v, float limit) {
  if (limit <= 0.f) return v;
  float norm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)), __fmul_rn(v.z, v.z)));
  float denom = norm > limit ? __fdiv_rn(limit, norm) : 1.0f;
  return make_float3(__fmul_rn(v.x, denom), __fmul_rn(v.y, denom), __fmul_rn(v.z, denom));
}

__global__ void __launch_bounds__(TSD_MAX_GRAPH_NODES) k_eq_transform(tsd_batch_t b, tsd_edges_t e,
                                                                     const float* __restrict__ pos,
                                                                     tsd_score_channel_t ch, float inv_div,
                                                                     float* __restrict__ node_eq) {
  __shared__ float spos[3 * TSD_MAX_GRAPH_NODES];
  const int n0 = b.graph_ptr[blockIdx.x], n = b.graph_ptr[blockIdx.x + 1] - n0;
  for (int i = threadIdx.x; i < 3 * n; i += blockDim.x) spos[i] = pos[(size_t)3 * n0 + i];
  __syncthreads();
  for (int li = threadIdx.x; li < n; li += blockDim.x) {
    float3 s = tsd_node_score(ch, e, spos, n0, n0 + li, inv_div);
    node_eq[3 * (n0 + li)] = s.x;
    node_eq[3 * (n0 + li) + 1] = s.y;
    node_eq[3 * (n0 + li) + 2] = s.z;
  }
}

extern "C" int tsd_eq_transform(const tsd_batch_t* batch, const tsd_edges_t* edges, const float* pos,
                                const tsd_score_channel_t* ch, float inv_div, float* node_eq, tsd_stream_t stream) {
  TSD_REQUIRE(batch && edges && pos && ch && ch->inv && node_eq);
  if (batch->num_graphs == 0) return TSD_OK;
  k_eq_transform<<<batch->num_graphs, TSD_MAX_GRAPH_NODES, 0, tsd_cu(stream)>>>(*batch, *edges, pos, *ch, inv_div, node_eq);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

// ------------------------------------------------------------------------------------- K7
// Per-edge terms t_e = ((1/len_e) (p_row - p_col)) * s_e are computed ONCE per edge by all the
// threads (coalesced loads) into shared memory; every atom then adds its out-edge terms and
// subtracts its in-edge terms sequentially in edge order -- the same association order (and
// the same bits) as the per-atom global-memory walk of tsd_node_score, without its chain of
// dependent global loads.  Edges masked out of a channel contribute an exact +0.
struct LdSmem {
  float* term0;   // [cap][3]
  float* term1;   // [cap][3] (two-channel variant)
  int* in_local;  // [cap] in-slot -> edge id local to the graph
};

__device__ __forceinline__ void k7_edge_terms(const tsd_score_channel_t& ch, const tsd_edges_t& e, const float* spos,
                                              int n0, int e0, int count, float inv_div, float inv_mul, float* term) {
  for (int k = threadIdx.x; k < count; k += blockDim.x) {
    const int id = e0 + k;
    float tx = 0.f, ty = 0.f, tz = 0.f;
    if (tsd_edge_selected(ch, id)) {
      const int r = e.row[id] - n0, c = e.col[id] - n0;
      const float inv_len = __fdiv_rn(1.0f, e.length[id]);
      const float s = __fdiv_rn(__fmul_rn(ch.inv[ch.inv_index ? ch.inv_index[id] : id], inv_mul), inv_div);
      tx = __fmul_rn(__fmul_rn(inv_len, __fsub_rn(spos[3 * r], spos[3 * c])), s);
      ty = __fmul_rn(__fmul_rn(inv_len, __fsub_rn(spos[3 * r + 1], spos[3 * c + 1])), s);
      tz = __fmul_rn(__fmul_rn(inv_len, __fsub_rn(spos[3 * r + 2], spos[3 * c + 2])), s);
    }
    term[3 * k] = tx;
    term[3 * k + 1] = ty;
    term[3 * k + 2] = tz;
  }
}

__device__ __forceinline__ float3 k7_node_sum(const tsd_edges_t& e, const float* term, const int* in_local, int e0,
                                              int i) {
  float ax = 0.f, ay = 0.f, az = 0.f, bx = 0.f, by = 0.f, bz = 0.f;
  for (int k = e.row_ptr[i] - e0, end = e.row_ptr[i + 1] - e0; k < end; ++k) {
    ax = __fadd_rn(ax, term[3 * k]);
    ay = __fadd_rn(ay, term[3 * k + 1]);
    az = __fadd_rn(az, term[3 * k + 2]);
  }
  for (int k = e.in_ptr[i] - e0, end = e.in_ptr[i + 1] - e0; k < end; ++k) {
    const int id = in_local[k];
    bx = __fadd_rn(bx, -term[3 * id]);
    by = __fadd_rn(by, -term[3 * id + 1]);
    bz = __fadd_rn(bz, -term[3 * id + 2]);
  }
  return make_float3(__fadd_rn(ax, bx), __fadd_rn(ay, by), __fadd_rn(az, bz));
}

// ------------------------------------------------------------ one-shot score exchange between ensemble ranks
__device__ __forceinline__ void k7_st_release_sys(int* addr, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ int k7_ld_acquire_sys(const int* addr) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long k7_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Reaction g of this rank: push the partial scores in snew[3 n] to every rank's buffer, publish the flag, wait
// for every rank's contribution of this step and leave their rank-ordered sum in snew.  All CTAs of all ranks are
// co-resident (one small CTA per reaction), every CTA pushes before it waits, and the dependencies are per
// reaction: no rank can wait on a kernel that has not been launched, and nobody overwrites a buffer its peer may
// still be reading (a rank can get at most one step ahead of the slowest peer, hence two buffers).
__device__ void k7_exchange(const tsd_exchange_t& ex, int g, int n0, int n, int num_nodes, int step, float* snew,
                            int* nan_flag) {
  const int parity = step & 1;
  const int expect = *ex.epoch_base + step + 1;
  const size_t slab = (size_t)num_nodes * 3;  // one (rank, parity) block of scores
  for (int p = 0; p < ex.world; ++p) {        // push: coalesced stores straight into peer p's memory
    float* dst = ex.peer_data[p] + ((size_t)parity * ex.world + ex.rank) * slab + (size_t)3 * n0;
    for (int i = threadIdx.x; i < 3 * n; i += blockDim.x) __stcg(dst + i, snew[i]);
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < ex.world) {
    const int p = threadIdx.x;
    k7_st_release_sys(ex.peer_flags[p] + ((size_t)parity * ex.world + ex.rank) * ex.num_graphs + g, expect);
    // wait for rank p's contribution to reaction g (its flag lives in OUR memory: a local spin)
    const int* mine = ex.peer_flags[ex.rank] + ((size_t)parity * ex.world + p) * ex.num_graphs + g;
    const unsigned long long t0 = k7_globaltimer();
    while (k7_ld_acquire_sys(mine) != expect) {
      if (k7_globaltimer() - t0 > 5000000000ull) {  // a peer died or was never launched: do not hang the GPU
        atomicOr(nan_flag, 2);
        break;
      }
    }
  }
  __syncthreads();
  const float* src = ex.peer_data[ex.rank] + (size_t)parity * ex.world * slab + (size_t)3 * n0;
  for (int i = threadIdx.x; i < 3 * n; i += blockDim.x) {
    float s = __ldcg(src + i);  // L2: the peers' stores never pass through this SM's L1
    for (int r = 1; r < ex.world; ++r) s = __fadd_rn(s, __ldcg(src + (size_t)r * slab + i));
    snew[i] = s;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(TSD_MAX_GRAPH_NODES) k_ld_step(tsd_batch_t b, tsd_edges_t e, float* __restrict__ pos,
                                                                tsd_score_channel_t ch0, tsd_score_channel_t ch1,
                                                                tsd_ld_params_t ld, tsd_exchange_t ex, int smem_edge_cap) {
  extern __shared__ float k7_dyn[];
  __shared__ float spos[3 * TSD_MAX_GRAPH_NODES];
  __shared__ float snew[3 * TSD_MAX_GRAPH_NODES];
  __shared__ float smean[3];
  __shared__ int sstep;
  const int g = blockIdx.x;
  const int n0 = b.graph_ptr[g], n = b.graph_ptr[g + 1] - n0;
  if (threadIdx.x == 0) sstep = *reinterpret_cast<volatile int*>(ld.step_counter);
  for (int i = threadIdx.x; i < 3 * n; i += blockDim.x) spos[i] = pos[(size_t)3 * n0 + i];
  __syncthreads();
  const int step = sstep;
  if (step < ld.num_steps) {
    const int rule = ld.rule;
    const bool wide = rule == TSD_RULE_DDPM || rule == TSD_RULE_DDPM_DUALENC;  // 8-column tables
    const float* sc = ld.sched + (size_t)(wide ? 8 : 4) * step;
    const float step_size = sc[0], sigma = sc[1], nscale = sc[2];
    const float use1_flag = rule == TSD_RULE_DDPM ? 0.f : (rule == TSD_RULE_DDPM_DUALENC ? sc[6] : sc[3]);
    // DSM (dualenc.py:305-309,353-361): the edge scores are multiplied by 1 / sigma(noise level) before eq_transform
    const float inv_mul = rule == TSD_RULE_DSM ? __fdiv_rn(1.0f, sigma) : 1.0f;
    const bool use1 = ch1.inv != nullptr && use1_flag != 0.f;
    const int e0 = e.row_ptr[n0], count = e.row_ptr[n0 + n] - e0;  // this graph's edges are contiguous
    const bool external = ld.node_score != nullptr;  // per-atom scores already reduced over the ensemble ranks
    const bool staged = !external && count <= smem_edge_cap;
    LdSmem sm;
    sm.term0 = k7_dyn;
    sm.term1 = k7_dyn + 3 * (size_t)smem_edge_cap;
    sm.in_local = reinterpret_cast<int*>(k7_dyn + (ch1.inv ? 6 : 3) * (size_t)smem_edge_cap);
    if (staged) {
      k7_edge_terms(ch0, e, spos, n0, e0, count, ld.inv_div, inv_mul, sm.term0);
      if (use1) k7_edge_terms(ch1, e, spos, n0, e0, count, ld.inv_div, inv_mul, sm.term1);
      for (int k = threadIdx.x; k < count; k += blockDim.x) sm.in_local[k] = e.in_eid[e0 + k] - e0;
      __syncthreads();
    }
    const bool exchanged = ex.world > 0;
    if (exchanged) {
      // this rank's partial scores -> snew, exchanged in place for the rank-ordered sum over the ensemble
      for (int li = threadIdx.x; li < n; li += blockDim.x) {
        const float3 part = staged ? k7_node_sum(e, sm.term0, sm.in_local, e0, n0 + li)
                                   : tsd_node_score(ch0, e, spos, n0, n0 + li, ld.inv_div, inv_mul);
        snew[3 * li] = part.x;
        snew[3 * li + 1] = part.y;
        snew[3 * li + 2] = part.z;
      }
      __syncthreads();
      k7_exchange(ex, g, n0, n, b.num_nodes, step, snew, ld.nan_flag);
    }
    for (int li = threadIdx.x; li < n; li += blockDim.x) {
      const int i = n0 + li;
      float3 eps;
      if (exchanged) eps = make_float3(snew[3 * li], snew[3 * li + 1], snew[3 * li + 2]);
      else if (external) eps = make_float3(ld.node_score[3 * (size_t)i], ld.node_score[3 * (size_t)i + 1], ld.node_score[3 * (size_t)i + 2]);
      else eps = staged ? k7_node_sum(e, sm.term0, sm.in_local, e0, i) : tsd_node_score(ch0, e, spos, n0, i, ld.inv_div, inv_mul);
      eps = tsd_clip_norm(eps, ch0.clip);
      if (use1) {
        float3 g1 = staged ? k7_node_sum(e, sm.term1, sm.in_local, e0, i) : tsd_node_score(ch1, e, spos, n0, i, ld.inv_div, inv_mul);
        g1 = tsd_clip_norm(g1, ch1.clip);
        eps.x = __fadd_rn(eps.x, __fmul_rn(g1.x, ch1.weight));
        eps.y = __fadd_rn(eps.y, __fmul_rn(g1.y, ch1.weight));
        eps.z = __fadd_rn(eps.z, __fmul_rn(g1.z, ch1.weight));
      }
      float3 z;
      if (ld.noise) {
        const float* zp = ld.noise + ((size_t)step * b.num_nodes + i) * 3;
        z = make_float3(zp[0], zp[1], zp[2]);
      } else {
        z = tsd_philox_normal3(ld.seed, step, ld.atom_offset + i);
      }
      float nx, ny, nz;
      if (rule == TSD_RULE_DDPM_DUALENC) {
        // dualenc.py:906-944 (ddpm_noisy / ddpm_det differ only in the noise coefficient)
        const float ev[3] = {-eps.x, -eps.y, -eps.z}, zv[3] = {z.x, z.y, z.z};
        float out[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const float p = spos[3 * li + d];
          const float pos0 = __fsub_rn(__fmul_rn(sc[0], p), __fmul_rn(sc[1], ev[d]));
          const float mean = __fdiv_rn(__fadd_rn(__fmul_rn(sc[2], pos0), __fmul_rn(sc[3], p)), sc[4]);
          out[d] = __fadd_rn(mean, __fmul_rn(sc[5], zv[d]));
        }
        nx = out[0], ny = out[1], nz = out[2];
      } else if (rule == TSD_RULE_DSM) {
        // dualenc.py:1183-1184: pos + step_size * eps_pos + randn * sqrt(2 step_size)
        nx = __fadd_rn(__fadd_rn(spos[3 * li], __fmul_rn(step_size, eps.x)), __fmul_rn(z.x, nscale));
        ny = __fadd_rn(__fadd_rn(spos[3 * li + 1], __fmul_rn(step_size, eps.y)), __fmul_rn(z.y, nscale));
        nz = __fadd_rn(__fadd_rn(spos[3 * li + 2], __fmul_rn(step_size, eps.z)), __fmul_rn(z.z, nscale));
      } else if (rule == TSD_RULE_GENERALIZED) {
        // dualenc.py:904: pos - et * step_size_pos + noise * step_size_noise, et = -eps
        nx = __fadd_rn(__fsub_rn(spos[3 * li], __fmul_rn(-eps.x, sc[0])), __fmul_rn(z.x, sc[1]));
        ny = __fadd_rn(__fsub_rn(spos[3 * li + 1], __fmul_rn(-eps.y, sc[0])), __fmul_rn(z.y, sc[1]));
        nz = __fadd_rn(__fsub_rn(spos[3 * li + 2], __fmul_rn(-eps.z, sc[0])), __fmul_rn(z.z, sc[1]));
      } else if (rule == TSD_RULE_DDPM) {
        // sampler.py:223-236, one rounded operation per reference tensor op
        const float ev[3] = {-eps.x, -eps.y, -eps.z}, zv[3] = {z.x, z.y, z.z};
        float out[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const float pos_c = __fmul_rn(sc[0], spos[3 * li + d]);
          const float pos0 = __fsub_rn(__fmul_rn(sc[1], pos_c), __fmul_rn(sc[2], ev[d]));
          const float mean = __fdiv_rn(__fadd_rn(__fmul_rn(sc[3], pos0), __fmul_rn(sc[4], pos_c)), sc[5]);
          out[d] = __fdiv_rn(__fadd_rn(mean, __fmul_rn(sc[6], zv[d])), sc[7]);
        }
        nx = out[0], ny = out[1], nz = out[2];
      } else {
        // pos + step_size * eps / sigma + noise * sqrt(2 step_size)   (sampler.py:239-244)
        nx = __fadd_rn(__fadd_rn(spos[3 * li], __fdiv_rn(__fmul_rn(step_size, eps.x), sigma)), __fmul_rn(z.x, nscale));
        ny = __fadd_rn(__fadd_rn(spos[3 * li + 1], __fdiv_rn(__fmul_rn(step_size, eps.y), sigma)), __fmul_rn(z.y, nscale));
        nz = __fadd_rn(__fadd_rn(spos[3 * li + 2], __fdiv_rn(__fmul_rn(step_size, eps.z), sigma)), __fmul_rn(z.z, nscale));
      }
      if (isnan(nx) || isnan(ny) || isnan(nz)) atomicOr(ld.nan_flag, 1);
      snew[3 * li] = nx;
      snew[3 * li + 1] = ny;
      snew[3 * li + 2] = nz;
    }
    __syncthreads();
    // center_pos: subtract the per-graph mean (sequential sum in atom order, like scatter_mean)
    if (threadIdx.x < 3) {
      float s = 0.f;
      for (int li = 0; li < n; ++li) s = __fadd_rn(s, snew[3 * li + threadIdx.x]);
      smean[threadIdx.x] = __fdiv_rn(s, (float)max(n, 1));
    }
    __syncthreads();
    const int slot = step - ld.traj_base_step;
    float* traj = (ld.traj && slot >= 0 && slot < ld.traj_steps) ? ld.traj + (size_t)slot * b.num_nodes * 3 : nullptr;
    for (int i = threadIdx.x; i < 3 * n; i += blockDim.x) {
      float v = __fsub_rn(snew[i], smean[i % 3]);
      if (ld.clip_pos > 0.f) v = fminf(fmaxf(v, -ld.clip_pos), ld.clip_pos);
      pos[(size_t)3 * n0 + i] = v;
      if (traj) traj[(size_t)3 * n0 + i] = v;
    }
  }
  // last CTA out advances the step counter (every CTA has read it by then)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    int t = atomicAdd(ld.ticket, 1);
    if (t == (int)gridDim.x - 1) {
      *ld.ticket = 0;
      *ld.step_counter = step + 1;
      __threadfence();
    }
  }
}

extern "C" int tsd_ld_step(const tsd_batch_t* batch, const tsd_edges_t* edges, float* pos, const tsd_score_channel_t* ch0,
                           const tsd_score_channel_t* ch1, const tsd_ld_params_t* ld, tsd_stream_t stream) {
  TSD_REQUIRE(batch && edges && pos && ch0 && (ch0->inv || ld->node_score) && ld && ld->sched && ld->step_counter &&
              ld->ticket && ld->nan_flag);
  TSD_REQUIRE(!(ld->node_score && ch1 && ch1->inv));  // the reduced-score mode is single channel
  tsd_exchange_t ex;
  memset(&ex, 0, sizeof(ex));
  if (ld->exchange) {
    ex = *ld->exchange;
    TSD_REQUIRE(ex.world >= 1 && ex.world <= TSD_MAX_EXCHANGE_RANKS && ex.rank >= 0 && ex.rank < ex.world && ex.epoch_base &&
                ex.num_graphs == batch->num_graphs && ch0->inv && !ld->node_score && !(ch1 && ch1->inv));
    for (int p = 0; p < ex.world; ++p) TSD_REQUIRE(ex.peer_data[p] && ex.peer_flags[p]);
  }
  TSD_REQUIRE(batch->max_graph_nodes <= TSD_MAX_GRAPH_NODES);
  TSD_REQUIRE(ld->rule >= TSD_RULE_LD && ld->rule <= TSD_RULE_DSM);
  if (batch->num_graphs == 0) return TSD_OK;
  tsd_score_channel_t off;
  memset(&off, 0, sizeof(off));
  int threads = ((batch->max_graph_nodes + 31) / 32) * 32;
  if (threads < 128) threads = 128;  // extra warps only help the per-edge staging loops
  // stage the per-edge terms of one graph in shared memory when they fit (<= 160 KB)
  const int nch = (ch1 && ch1->inv) ? 2 : 1;
  const long long max_edges = (long long)batch->max_graph_nodes * (batch->max_graph_nodes - 1);
  const long long bytes = max_edges * (12 * nch + 4);
  int cap = 0;
  size_t smem = 0;
  if (max_edges > 0 && bytes <= 160 * 1024) {
    cap = (int)max_edges;
    smem = (size_t)bytes;
    if (smem > 40 * 1024) TSD_CUDA(cudaFuncSetAttribute(k_ld_step, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  }
  k_ld_step<<<batch->num_graphs, threads, smem, tsd_cu(stream)>>>(*batch, *edges, pos, *ch0, ch1 ? *ch1 : off, *ld, ex, cap);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

// ------------------------------------------------------------------------- peer-mapped memory (CUDA IPC)
extern "C" int tsd_peer_alloc(uint64_t bytes, void** ptr, unsigned char* handle64) {
  TSD_REQUIRE(ptr && handle64 && bytes > 0);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  TSD_CUDA(cudaMalloc(ptr, bytes));
  TSD_CUDA(cudaMemset(*ptr, 0, bytes));
  cudaIpcMemHandle_t h;
  TSD_CUDA(cudaIpcGetMemHandle(&h, *ptr));
  memcpy(handle64, &h, 64);
  return TSD_OK;
}

extern "C" int tsd_peer_open(const unsigned char* handle64, void** ptr) {
  TSD_REQUIRE(ptr && handle64);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  TSD_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return TSD_OK;
}

extern "C" int tsd_peer_close(void* ptr) {
  TSD_REQUIRE(ptr);
  TSD_CUDA(cudaIpcCloseMemHandle(ptr));
  return TSD_OK;
}

extern "C" int tsd_peer_free(void* ptr) {
  TSD_REQUIRE(ptr);
  TSD_CUDA(cudaFree(ptr));
  return TSD_OK;
}
