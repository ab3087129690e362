// Message-passing kernels: deterministic segmented reductions over the dst-sorted in-CSR
// (no atomics), plus the small node-embedding kernels.
//   k_cfconv_aggregate : schnet.py:102-107  agg_i = sum_{j->i} x1_j * W_ji
//   k_gine_aggregate   : gin.py:61-73       out_i = sum_{j->i} relu(h_j + e_ji) + (1+eps) h_i
// The in-CSR lists the edges of every target in ascending source order, which is the order
// a sequential scatter_add over the row-major sorted edge list visits them, so sums are
// reproducible run to run and match the oracle's association order.
#include <stdlib.h>

#include "common.cuh"

// HBM-bound: per target node stream its in-edges' filter rows (E*H*4 bytes in total, each
// read exactly once, 16 B per lane, fully coalesced 512-byte slabs) and gather x1 rows
// (N*H*4 bytes, L2 resident).  Algorithmic bytes: E*H*4 + 2*N*H*4 + E*8 + (N+1)*4.
// One warp per (target node, 128-channel slab).  The lanes first fetch the segment's edge
// ids and source ids coalesced, then broadcast them with shuffles, so the row loads of
// consecutive edges are independent and UNROLL of them are in flight per lane.
__device__ __forceinline__ void tsd_fma_rn4(float4& acc, const float4& x, const float4& w) {
  acc.x = __fadd_rn(acc.x, __fmul_rn(x.x, w.x));
  acc.y = __fadd_rn(acc.y, __fmul_rn(x.y, w.y));
  acc.z = __fadd_rn(acc.z, __fmul_rn(x.z, w.z));
  acc.w = __fadd_rn(acc.w, __fmul_rn(x.w, w.w));
}

template <int UNROLL>
__global__ void __launch_bounds__(256) k_cfconv_aggregate(int num_nodes, int H, int slabs,
                                                          const int* __restrict__ in_ptr,
                                                          const int* __restrict__ in_eid,
                                                          const int* __restrict__ in_src, const float* __restrict__ x1,
                                                          const float* __restrict__ filt, float* __restrict__ agg) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int node = gw / slabs, slab = gw - node * slabs;
  if (node >= num_nodes) return;  // warp uniform
  const int off = slab * 128 + lane * 4;
  const bool active = off < H;
  const int beg = in_ptr[node], end = in_ptr[node + 1];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int base = beg; base < end; base += 32) {
    const int cnt = min(32, end - base);
    const int e_l = lane < cnt ? in_eid[base + lane] : 0;
    const int r_l = lane < cnt ? in_src[base + lane] : 0;
    int j = 0;
    for (; j + UNROLL <= cnt; j += UNROLL) {
      float4 w[UNROLL], x[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int e = __shfl_sync(TSD_FULL_MASK, e_l, j + u), r = __shfl_sync(TSD_FULL_MASK, r_l, j + u);
        if (active) {
          w[u] = __ldcs(reinterpret_cast<const float4*>(filt + (size_t)e * H + off));  // streamed once
          x[u] = __ldg(reinterpret_cast<const float4*>(x1 + (size_t)r * H + off));
        }
      }
      if (active) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) tsd_fma_rn4(acc, x[u], w[u]);
      }
    }
    for (; j < cnt; ++j) {
      const int e = __shfl_sync(TSD_FULL_MASK, e_l, j), r = __shfl_sync(TSD_FULL_MASK, r_l, j);
      if (active) {
        float4 w = __ldcs(reinterpret_cast<const float4*>(filt + (size_t)e * H + off));
        float4 x = __ldg(reinterpret_cast<const float4*>(x1 + (size_t)r * H + off));
        tsd_fma_rn4(acc, x, w);
      }
    }
  }
  if (active) *reinterpret_cast<float4*>(agg + (size_t)node * H + off) = acc;
}

// Graph-staged variant (large edge lists, see tsd_aggregate in api.cu): one CTA per (reaction, 128-channel slab).  The x1 rows of the
// reaction (<= 256 x 512 B) are staged in shared memory once, so the per-edge gather of the
// source features never leaves the SM; only the filter rows stream from L2/HBM, each exactly
// once.  (The node-parallel kernel above re-reads a 512-byte x1 slab per edge from L2: E*H*4
// bytes of gather traffic on top of the E*H*4 filter stream.)
template <int UNROLL>
__global__ void __launch_bounds__(256) k_cfconv_aggregate_staged(int H, const int* __restrict__ graph_ptr,
                                                                 const int* __restrict__ in_ptr,
                                                                 const int* __restrict__ in_eid,
                                                                 const int* __restrict__ in_src,
                                                                 const float* __restrict__ x1,
                                                                 const float* __restrict__ filt, float* __restrict__ agg) {
  extern __shared__ float4 sx[];  // [n][32] float4 = this slab of the reaction's x1 rows
  const int g = blockIdx.x, slab = blockIdx.y;
  const int n0 = graph_ptr[g], n = graph_ptr[g + 1] - n0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int off = slab * 128 + lane * 4;
  const bool active = off < H;
  for (int i = warp; i < n; i += nwarps)
    sx[i * 32 + lane] = active ? __ldg(reinterpret_cast<const float4*>(x1 + (size_t)(n0 + i) * H + off))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
  // in-CSR bounds of all of this warp's nodes in one coalesced round trip (lane k <-> node warp + k nwarps),
  // and the edge / source ids of the NEXT node prefetched while the current node's rows stream:
  // the dependent chain per node is then just the row loads.
  const int my_nodes = n > warp ? (n - warp + nwarps - 1) / nwarps : 0;  // <= 32 for n <= 256, 8 warps
  int beg_l = 0, end_l = 0;
  if (lane < my_nodes) {
    beg_l = in_ptr[n0 + warp + lane * nwarps];
    end_l = in_ptr[n0 + warp + lane * nwarps + 1];
  }
  __syncthreads();
  int beg = __shfl_sync(TSD_FULL_MASK, beg_l, 0), end = __shfl_sync(TSD_FULL_MASK, end_l, 0);
  int e_l = (my_nodes > 0 && lane < end - beg) ? in_eid[beg + lane] : 0;
  int r_l = (my_nodes > 0 && lane < end - beg) ? in_src[beg + lane] - n0 : 0;
  for (int k = 0; k < my_nodes; ++k) {
    const int node = n0 + warp + k * nwarps;
    const int nbeg = __shfl_sync(TSD_FULL_MASK, beg_l, min(k + 1, 31)), nend = __shfl_sync(TSD_FULL_MASK, end_l, min(k + 1, 31));
    const bool has_next = k + 1 < my_nodes;
    const int e_n = (has_next && lane < nend - nbeg) ? in_eid[nbeg + lane] : 0;   // prefetch (first 32 in-edges)
    const int r_n = (has_next && lane < nend - nbeg) ? in_src[nbeg + lane] - n0 : 0;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int base = beg; base < end; base += 32) {
      const int cnt = min(32, end - base);
      if (base != beg) {  // degree > 32: later batches are fetched on demand
        e_l = lane < cnt ? in_eid[base + lane] : 0;
        r_l = lane < cnt ? in_src[base + lane] - n0 : 0;
      }
      int j = 0;
      for (; j + UNROLL <= cnt; j += UNROLL) {
        float4 w[UNROLL];
        int r[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          const int e = __shfl_sync(TSD_FULL_MASK, e_l, j + u);
          r[u] = __shfl_sync(TSD_FULL_MASK, r_l, j + u);
          if (active) w[u] = __ldcs(reinterpret_cast<const float4*>(filt + (size_t)e * H + off));  // streamed once
        }
        if (active) {
#pragma unroll
          for (int u = 0; u < UNROLL; ++u) tsd_fma_rn4(acc, sx[r[u] * 32 + lane], w[u]);
        }
      }
      for (; j < cnt; ++j) {
        const int e = __shfl_sync(TSD_FULL_MASK, e_l, j), r = __shfl_sync(TSD_FULL_MASK, r_l, j);
        if (active) tsd_fma_rn4(acc, sx[r * 32 + lane], __ldcs(reinterpret_cast<const float4*>(filt + (size_t)e * H + off)));
      }
    }
    if (active) *reinterpret_cast<float4*>(agg + (size_t)node * H + off) = acc;
    beg = nbeg, end = nend, e_l = e_n, r_l = r_n;
  }
}

int tsd_launch_cfconv_aggregate_staged(const tsd_batch_t* b, int H, const int* in_ptr, const int* in_eid,
                                       const int* in_src, const float* x1, const float* filt, float* agg,
                                       cudaStream_t s) {
  TSD_REQUIRE(H % 4 == 0 && H >= 4 && H <= 4096);
  if (b->num_graphs == 0) return TSD_OK;
  const int slabs = tsd_ceil_div(H, 128);
  const size_t smem = (size_t)b->max_graph_nodes * 32 * sizeof(float4);
  if (smem > 48 * 1024) TSD_CUDA(cudaFuncSetAttribute(k_cfconv_aggregate_staged<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_cfconv_aggregate_staged<8><<<dim3(b->num_graphs, slabs), 256, smem, s>>>(H, b->graph_ptr, in_ptr, in_eid, in_src, x1, filt, agg);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

int tsd_launch_cfconv_aggregate(int num_nodes, int H, const int* in_ptr, const int* in_eid, const int* in_src,
                                const float* x1, const float* filt, float* agg, cudaStream_t s) {
  TSD_REQUIRE(H % 4 == 0 && H >= 4 && H <= 4096);
  if (num_nodes == 0) return TSD_OK;
  const int slabs = tsd_ceil_div(H, 128);
  const long long warps = (long long)num_nodes * slabs;
  // measured best at batch 100 (profiles/r1_aggregate_hbm.txt): UNROLL 4, 60 registers, 256 threads
  k_cfconv_aggregate<4><<<(unsigned)((warps + 7) / 8), 256, 0, s>>>(num_nodes, H, slabs, in_ptr, in_eid, in_src, x1, filt, agg);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

__global__ void __launch_bounds__(256) k_gine_aggregate(int num_nodes, int H, int slabs,
                                                        const int* __restrict__ in_ptr,
                                                        const int* __restrict__ in_eid, const int* __restrict__ in_src,
                                                        const int* __restrict__ local_tab,
                                                        const float* __restrict__ h, const float* __restrict__ ea,
                                                        const float* __restrict__ eps, float* __restrict__ out) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int node = gw / slabs, slab = gw - node * slabs;
  if (node >= num_nodes) return;
  const int off = slab * 128 + lane * 4;
  const bool active = off < H;
  const int beg = in_ptr[node], end = in_ptr[node + 1];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int base = beg; base < end; base += 32) {
    const int cnt = min(32, end - base);
    int e_l = 0, r_l = 0, loc_l = 0;
    if (lane < cnt) {
      e_l = in_eid[base + lane];
      r_l = in_src[base + lane];
      loc_l = local_tab[e_l];  // GIN runs on edge_index[:, edge_type > 0] (dualenc.py:343-347)
    }
    for (int j = 0; j < cnt; ++j) {
      const int e = __shfl_sync(TSD_FULL_MASK, e_l, j), r = __shfl_sync(TSD_FULL_MASK, r_l, j);
      const int loc = __shfl_sync(TSD_FULL_MASK, loc_l, j);
      if (loc != 0 && active) {
        float4 a = __ldg(reinterpret_cast<const float4*>(ea + (size_t)e * H + off));
        float4 x = __ldg(reinterpret_cast<const float4*>(h + (size_t)r * H + off));
        acc.x = __fadd_rn(acc.x, fmaxf(__fadd_rn(x.x, a.x), 0.f));
        acc.y = __fadd_rn(acc.y, fmaxf(__fadd_rn(x.y, a.y), 0.f));
        acc.z = __fadd_rn(acc.z, fmaxf(__fadd_rn(x.z, a.z), 0.f));
        acc.w = __fadd_rn(acc.w, fmaxf(__fadd_rn(x.w, a.w), 0.f));
      }
    }
  }
  if (active) {
    const float s = __fadd_rn(1.f, eps[0]);
    float4 x = *reinterpret_cast<const float4*>(h + (size_t)node * H + off);
    acc.x = __fadd_rn(acc.x, __fmul_rn(s, x.x));
    acc.y = __fadd_rn(acc.y, __fmul_rn(s, x.y));
    acc.z = __fadd_rn(acc.z, __fmul_rn(s, x.z));
    acc.w = __fadd_rn(acc.w, __fmul_rn(s, x.w));
    *reinterpret_cast<float4*>(out + (size_t)node * H + off) = acc;
  }
}

int tsd_launch_gine_aggregate(int num_nodes, int H, const int* in_ptr, const int* in_eid, const int* in_src,
                              const int* local_tab, const float* h, const float* ea, const float* eps, float* out,
                              cudaStream_t s) {
  TSD_REQUIRE(H % 4 == 0 && H >= 4 && H <= 4096);
  if (num_nodes == 0) return TSD_OK;
  const int slabs = tsd_ceil_div(H, 128);
  const long long warps = (long long)num_nodes * slabs;
  k_gine_aggregate<<<(unsigned)((warps + 7) / 8), 256, 0, s>>>(num_nodes, H, slabs, in_ptr, in_eid, in_src, local_tab,
                                                                h, ea, eps, out);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

// ----------------------------------------------------------------------- node embeddings
// condensenc.py:193-198: z = cat[emb[Z] + W r, W p - W r]
__global__ void k_condensed_node_embed(int num_nodes, const int64_t* __restrict__ atom_type,
                                       const int64_t* __restrict__ r_feat, const int64_t* __restrict__ p_feat,
                                       int feat_dim, const float* __restrict__ atom_emb,
                                       const float* __restrict__ feat_w, int half, float* __restrict__ z) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= num_nodes * half) return;
  int n = idx / half, j = idx - n * half;
  const float* w = feat_w + (size_t)j * feat_dim;
  float sr = 0.f, sp = 0.f;
  for (int f = 0; f < feat_dim; ++f) {
    sr = fmaf((float)r_feat[(size_t)n * feat_dim + f], w[f], sr);
    sp = fmaf((float)p_feat[(size_t)n * feat_dim + f], w[f], sp);
  }
  float e = atom_emb[(size_t)atom_type[n] * half + j];
  z[(size_t)n * 2 * half + j] = e + sr;
  z[(size_t)n * 2 * half + half + j] = sp - sr;
}

extern "C" int tsd_condensed_node_embed(int32_t num_nodes, const int64_t* atom_type, const int64_t* r_feat,
                                        const int64_t* p_feat, int32_t feat_dim, const float* atom_emb,
                                        const float* feat_weight, int32_t half, float* z, tsd_stream_t stream) {
  TSD_REQUIRE(atom_type && r_feat && p_feat && atom_emb && feat_weight && z && half > 0 && feat_dim > 0);
  if (num_nodes == 0) return TSD_OK;
  int total = num_nodes * half;
  k_condensed_node_embed<<<tsd_ceil_div(total, 256), 256, 0, tsd_cu(stream)>>>(num_nodes, atom_type, r_feat, p_feat, feat_dim,
                                                                              atom_emb, feat_weight, half, z);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

// nn.Embedding(max_norm): rows that are looked up and whose L2 norm exceeds max_norm are
// rescaled in place by max_norm / (norm + 1e-7) before the gather (schnet.py:152).
__global__ void k_embedding_renorm(int num_nodes, const int64_t* __restrict__ index, float* weight, int num_rows,
                                   int dim, float max_norm) {
  const int r = blockIdx.x;  // one warp per embedding row
  const int lane = threadIdx.x;
  bool used = false;
  for (int i = lane; i < num_nodes; i += 32) used |= (index[i] == r);
  if (!__any_sync(TSD_FULL_MASK, used)) return;
  float ss = 0.f;
  for (int d = lane; d < dim; d += 32) {
    float v = weight[(size_t)r * dim + d];
    ss = fmaf(v, v, ss);
  }
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(TSD_FULL_MASK, ss, o);
  float norm = sqrtf(ss);
  if (norm > max_norm) {
    float scale = max_norm / (norm + 1e-7f);
    for (int d = lane; d < dim; d += 32) weight[(size_t)r * dim + d] *= scale;
  }
}

__global__ void k_embedding_gather(int num_nodes, const int64_t* __restrict__ index, const float* __restrict__ weight,
                                   int dim, float* __restrict__ out) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= num_nodes * dim) return;
  int n = idx / dim, d = idx - n * dim;
  out[idx] = weight[(size_t)index[n] * dim + d];
}

extern "C" int tsd_embedding(int32_t num_nodes, const int64_t* index, float* weight, int32_t num_rows, int32_t dim,
                             float max_norm, float* out, tsd_stream_t stream) {
  TSD_REQUIRE(index && weight && out && num_rows > 0 && dim > 0);
  if (num_nodes == 0) return TSD_OK;
  cudaStream_t s = tsd_cu(stream);
  if (max_norm > 0.f) {
    k_embedding_renorm<<<num_rows, 32, 0, s>>>(num_nodes, index, weight, num_rows, dim, max_norm);
    TSD_LAUNCH_CHECK();
  }
  k_embedding_gather<<<tsd_ceil_div(num_nodes * dim, 256), 256, 0, s>>>(num_nodes, index, weight, dim, out);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}
