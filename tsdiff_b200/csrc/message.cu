// Message-passing kernels: deterministic segmented reductions over the dst-sorted in-CSR
// (no atomics), plus the small node-embedding kernels.
//   k_cfconv_aggregate : schnet.py:102-107  agg_i = sum_{j->i} x1_j * W_ji
//   k_gine_aggregate   : gin.py:61-73       out_i = sum_{j->i} relu(h_j + e_ji) + (1+eps) h_i
// The in-CSR lists the edges of every target in ascending source order, which is the order
// a sequential scatter_add over the row-major sorted edge list visits them, so sums are
// reproducible run to run and match the oracle's association order.
#include "common.cuh"

// HBM-bound: per target node stream its in-edges' filter rows (E*H*4 bytes in total, each
// read exactly once, 16 B per lane, fully coalesced 4*LPN-byte rows) and gather x1 rows
// (N*H*4 bytes, L2 resident).  Algorithmic bytes: E*H*4 + 2*N*H*4 + E*8 + (N+1)*4.
template <int UNROLL>
__global__ void __launch_bounds__(256) k_cfconv_aggregate(int num_nodes, int H, const int* __restrict__ in_ptr,
                                                          const int* __restrict__ in_eid,
                                                          const int* __restrict__ row, const float* __restrict__ x1,
                                                          const float* __restrict__ filt, float* __restrict__ agg) {
  const int lpn = H >> 2;  // lanes per node, one float4 each
  const int npb = blockDim.x / lpn;
  const int node = blockIdx.x * npb + threadIdx.x / lpn;
  const int lane = threadIdx.x % lpn;
  if (node >= num_nodes || threadIdx.x >= npb * lpn) return;
  const int beg = in_ptr[node], end = in_ptr[node + 1];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int k = beg;
  for (; k + UNROLL <= end; k += UNROLL) {
    int e[UNROLL], r[UNROLL];
    float4 w[UNROLL], x[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) e[u] = in_eid[k + u];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) r[u] = row[e[u]];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      w[u] = __ldcs(reinterpret_cast<const float4*>(filt + (size_t)e[u] * H) + lane);  // streamed once
      x[u] = __ldg(reinterpret_cast<const float4*>(x1 + (size_t)r[u] * H) + lane);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      acc.x = __fadd_rn(acc.x, __fmul_rn(x[u].x, w[u].x));
      acc.y = __fadd_rn(acc.y, __fmul_rn(x[u].y, w[u].y));
      acc.z = __fadd_rn(acc.z, __fmul_rn(x[u].z, w[u].z));
      acc.w = __fadd_rn(acc.w, __fmul_rn(x[u].w, w[u].w));
    }
  }
  for (; k < end; ++k) {
    int e = in_eid[k];
    int r = row[e];
    float4 w = __ldcs(reinterpret_cast<const float4*>(filt + (size_t)e * H) + lane);
    float4 x = __ldg(reinterpret_cast<const float4*>(x1 + (size_t)r * H) + lane);
    acc.x = __fadd_rn(acc.x, __fmul_rn(x.x, w.x));
    acc.y = __fadd_rn(acc.y, __fmul_rn(x.y, w.y));
    acc.z = __fadd_rn(acc.z, __fmul_rn(x.z, w.z));
    acc.w = __fadd_rn(acc.w, __fmul_rn(x.w, w.w));
  }
  reinterpret_cast<float4*>(agg + (size_t)node * H)[lane] = acc;
}

int tsd_launch_cfconv_aggregate(int num_nodes, int H, const int* in_ptr, const int* in_eid, const int* row,
                                const float* x1, const float* filt, float* agg, cudaStream_t s) {
  TSD_REQUIRE(H % 4 == 0 && H >= 4 && H <= 1024);
  if (num_nodes == 0) return TSD_OK;
  int lpn = H / 4, npb = 256 / lpn;
  k_cfconv_aggregate<4><<<tsd_ceil_div(num_nodes, npb), 256, 0, s>>>(num_nodes, H, in_ptr, in_eid, row, x1, filt, agg);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

__global__ void __launch_bounds__(256) k_gine_aggregate(int num_nodes, int H, const int* __restrict__ in_ptr,
                                                        const int* __restrict__ in_eid, const int* __restrict__ row,
                                                        const int* __restrict__ local_tab,
                                                        const float* __restrict__ h, const float* __restrict__ ea,
                                                        const float* __restrict__ eps, float* __restrict__ out) {
  const int lpn = H >> 2;
  const int npb = blockDim.x / lpn;
  const int node = blockIdx.x * npb + threadIdx.x / lpn;
  const int lane = threadIdx.x % lpn;
  if (node >= num_nodes || threadIdx.x >= npb * lpn) return;
  const int beg = in_ptr[node], end = in_ptr[node + 1];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = beg; k < end; ++k) {
    int e = in_eid[k];
    if (local_tab[e] == 0) continue;  // GIN runs on edge_index[:, edge_type > 0] (dualenc.py:343-347)
    int r = row[e];
    float4 a = __ldg(reinterpret_cast<const float4*>(ea + (size_t)e * H) + lane);
    float4 x = __ldg(reinterpret_cast<const float4*>(h + (size_t)r * H) + lane);
    acc.x = __fadd_rn(acc.x, fmaxf(__fadd_rn(x.x, a.x), 0.f));
    acc.y = __fadd_rn(acc.y, fmaxf(__fadd_rn(x.y, a.y), 0.f));
    acc.z = __fadd_rn(acc.z, fmaxf(__fadd_rn(x.z, a.z), 0.f));
    acc.w = __fadd_rn(acc.w, fmaxf(__fadd_rn(x.w, a.w), 0.f));
  }
  const float s = __fadd_rn(1.f, eps[0]);
  float4 x = reinterpret_cast<const float4*>(h + (size_t)node * H)[lane];
  acc.x = __fadd_rn(acc.x, __fmul_rn(s, x.x));
  acc.y = __fadd_rn(acc.y, __fmul_rn(s, x.y));
  acc.z = __fadd_rn(acc.z, __fmul_rn(s, x.z));
  acc.w = __fadd_rn(acc.w, __fmul_rn(s, x.w));
  reinterpret_cast<float4*>(out + (size_t)node * H)[lane] = acc;
}

int tsd_launch_gine_aggregate(int num_nodes, int H, const int* in_ptr, const int* in_eid, const int* row,
                              const int* local_tab, const float* h, const float* ea, const float* eps, float* out,
                              cudaStream_t s) {
  TSD_REQUIRE(H % 4 == 0 && H >= 4 && H <= 1024);
  if (num_nodes == 0) return TSD_OK;
  int lpn = H / 4, npb = 256 / lpn;
  k_gine_aggregate<<<tsd_ceil_div(num_nodes, npb), 256, 0, s>>>(num_nodes, H, in_ptr, in_eid, row, local_tab, h, ea, eps, out);
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
