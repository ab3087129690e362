// K1 (bond-order pair tables, once per batch) and K2 (per-step radius + bond-order edge
// list in CSR form).  One CTA per reaction graph; graphs never interact, so everything is
// block-diagonal by construction: adjacency / reachability / neighbour sets are bit masks
// in shared memory (n_g <= 256 atoms -> <= 8 words per row).
//
// Replaces: models/common.py:115-202, :255-325 (K1); :205-223, :328-417 and
// models/epsnet/condensenc.py:117-154 (K2).
#include "common.cuh"

#define K1_TYPE_BITS 20
#define K1_TYPE_MASK ((1 << K1_TYPE_BITS) - 1)

// ------------------------------------------------------------------------------------ K1
// scatter bond entries into per-graph dense n_g x n_g tables: low 20 bits = sum of types
// (to_dense_adj sums duplicates), high bits = number of entries (adjacency).
__global__ void k_bond_scatter(int mode, int num_bonds, int num_nodes, const int64_t* __restrict__ bond_index,
                               const int64_t* __restrict__ bond_type, const int* __restrict__ graph_ptr,
                               const int* __restrict__ pair_ptr, const int* __restrict__ node_graph,
                               int* tab_r, int* tab_p, int* error_flag) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= num_bonds) return;
  long long r = bond_index[b], c = bond_index[(size_t)num_bonds + b], t = bond_type[b];
  if (r < 0 || r >= num_nodes || c < 0 || c >= num_nodes || t < 0 || t > K1_TYPE_MASK) {
    atomicOr(error_flag, 1);
    return;
  }
  int g = node_graph[r];
  int n0 = graph_ptr[g], n = graph_ptr[g + 1] - n0;
  int li = (int)r - n0, lj = (int)c - n0;
  if (lj < 0 || lj >= n) {  // bond across two reactions
    atomicOr(error_flag, 2);
    return;
  }
  if (li == lj) return;  // self loops are not part of the data contract
  int idx = pair_ptr[g] + li * n + lj;
  if (mode == 0) {
    int tr = (int)(t / TSD_NUM_BOND_TYPES), tp = (int)(t % TSD_NUM_BOND_TYPES);
    if (tr) atomicAdd(&tab_r[idx], tr + (1 << K1_TYPE_BITS));
    if (tp) atomicAdd(&tab_p[idx], tp + (1 << K1_TYPE_BITS));
  } else {
    atomicAdd(&tab_r[idx], (int)t + (1 << K1_TYPE_BITS));
  }
}

__device__ __forceinline__ int k1_type_from(int hop, int bond, int order, int hi_base) {
  if (hop == 1) return bond;
  return (hop >= 2 && hop <= order) ? hi_base + hop - 1 : 0;
}

// k-hop reachability by bit-parallel frontier expansion: reach_k[i] = OR_{m in reach_{k-1}[i]} adj1[m]
// with adj1 = binarize(adj + I) -- the same sets as adj_mats[k] in common.py:131-137.
__global__ void __launch_bounds__(256) k_hop_types(int mode, const int* __restrict__ graph_ptr,
                                                   const int* __restrict__ pair_ptr,
                                                   const int* __restrict__ tab_r, const int* __restrict__ tab_p,
                                                   int order_a, int order_b, int ts_decode, int* __restrict__ table_a,
                                                   int* __restrict__ table_b, int W) {
  extern __shared__ unsigned smem_u[];
  const int g = blockIdx.x;
  const int n0 = graph_ptr[g], n = graph_ptr[g + 1] - n0, base = pair_ptr[g];
  const int nsides = mode == 0 ? 2 : 1;
  const int maxo = mode == 0 ? max(order_a, order_b) : order_a;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
#define REACH(side, k, i, w) smem_u[(((side) * maxo + ((k) - 1)) * n + (i)) * W + (w)]

  for (int side = 0; side < nsides; ++side) {
    const int* tab = side ? tab_p : tab_r;
    for (int i = warp; i < n; i += nwarps)
      for (int w = 0; w < W; ++w) {
        int j = w * 32 + lane;
        bool a = j < n && (j == i || (tab[base + i * n + j] >> K1_TYPE_BITS) > 0);
        unsigned bits = __ballot_sync(TSD_FULL_MASK, a);
        if (lane == 0) REACH(side, 1, i, w) = bits;
      }
  }
  __syncthreads();
  for (int k = 2; k <= maxo; ++k) {
    for (int side = 0; side < nsides; ++side)
      for (int idx = threadIdx.x; idx < n * W; idx += blockDim.x) {
        int i = idx / W, w = idx - i * W;
        unsigned acc = 0;
        for (int w2 = 0; w2 < W; ++w2) {
          unsigned bits = REACH(side, k - 1, i, w2);
          while (bits) {
            int m = w2 * 32 + __ffs(bits) - 1;
            bits &= bits - 1;
            acc |= REACH(side, 1, m, w);
          }
        }
        REACH(side, k, i, w) = acc;
      }
    __syncthreads();
  }
  const int nb = TSD_NUM_BOND_TYPES;
  for (int p = threadIdx.x; p < n * n; p += blockDim.x) {
    int i = p / n, j = p - i * n;
    int code_a = 0, code_b = 0;
    if (i != j) {
      int hop[2] = {0, 0}, bond[2] = {0, 0};
      for (int side = 0; side < nsides; ++side) {
        const int* tab = side ? tab_p : tab_r;
        bond[side] = tab[base + p] & K1_TYPE_MASK;
        for (int k = 1; k <= maxo; ++k)
          if ((REACH(side, k, i, j >> 5) >> (j & 31)) & 1u) {
            hop[side] = k;
            break;
          }
      }
      if (mode == 0) {
        code_a = k1_type_from(hop[0], bond[0], order_a, nb) | (k1_type_from(hop[1], bond[1], order_a, nb) << 16);
        code_b = k1_type_from(hop[0], bond[0], order_b, nb) | (k1_type_from(hop[1], bond[1], order_b, nb) << 16);
      } else {
        int raw = k1_type_from(hop[0], bond[0], order_a, nb * nb);
        code_a = raw;
        // dualenc.py:270-293: rows of bond_emb for the (one or two) edge-encoder passes
        bool low = raw / (nb * nb) == 0;
        int high = low ? 0 : raw % (nb * nb) + nb;
        int row1, row2 = 0;
        if (ts_decode) {
          row1 = (low ? raw / nb : 0) + high;
          row2 = (low ? raw % nb : 0) + high;
        } else {
          row1 = (low ? raw % nb : 0) + high;
        }
        code_b = raw ? (row1 | (row2 << 16)) : 0;
      }
    }
    table_a[base + p] = code_a;
    table_b[base + p] = code_b;
  }
#undef REACH
}

extern "C" int tsd_bond_order_build(int mode, const tsd_batch_t* batch, int32_t num_bonds, const int64_t* bond_index,
                                    const int64_t* bond_type, int32_t order_a, int32_t order_b, int32_t ts_decode,
                                    int32_t* table_a, int32_t* table_b, int32_t* scratch, int32_t* error_flag,
                                    tsd_stream_t stream) {
  TSD_REQUIRE(batch && table_a && table_b && scratch && error_flag);
  TSD_REQUIRE(mode == 0 || mode == 1);
  TSD_REQUIRE(order_a >= 1 && order_a <= 8 && (mode == 1 || (order_b >= 1 && order_b <= 8)));
  TSD_REQUIRE(batch->max_graph_nodes >= 1 && batch->max_graph_nodes <= TSD_MAX_GRAPH_NODES);
  if (batch->num_graphs == 0) return TSD_OK;
  cudaStream_t s = tsd_cu(stream);
  // total pair count = pair_ptr[G]; the host passes it through edge_capacity + num_nodes
  size_t total_pairs = (size_t)batch->edge_capacity + (size_t)batch->num_nodes;
  int* tab_r = scratch;
  int* tab_p = scratch + total_pairs;
  TSD_CUDA(cudaMemsetAsync(scratch, 0, 2 * total_pairs * sizeof(int), s));
  if (num_bonds > 0) {
    TSD_REQUIRE(bond_index && bond_type);
    k_bond_scatter<<<tsd_ceil_div(num_bonds, 256), 256, 0, s>>>(mode, num_bonds, batch->num_nodes, bond_index, bond_type,
                                                                batch->graph_ptr, batch->pair_ptr, batch->node_graph,
                                                                tab_r, tab_p, error_flag);
    TSD_LAUNCH_CHECK();
  }
  int W = tsd_ceil_div(batch->max_graph_nodes, 32);
  int maxo = mode == 0 ? (order_a > order_b ? order_a : order_b) : order_a;
  size_t smem = (size_t)(mode == 0 ? 2 : 1) * maxo * batch->max_graph_nodes * W * sizeof(unsigned);
  if (smem > 48 * 1024) TSD_CUDA(cudaFuncSetAttribute(k_hop_types, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_hop_types<<<batch->num_graphs, 256, smem, s>>>(mode, batch->graph_ptr, batch->pair_ptr, tab_r, tab_p, order_a,
                                                   order_b, ts_decode, table_a, table_b, W);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

// ------------------------------------------------------------------------------------ K2
struct GraphTile {
  int n0, n, base, W;
  float* spos;      // 3n
  unsigned* radin;  // n*W : radin[c] bit r  <=>  radius edge (row = r, col = c)
  unsigned* out_a;  // n*W : out_a[r] bit c  <=>  edge (r, c) in graph a
  unsigned* in_a;   // n*W : in_a[c] bit r   <=>  edge (r, c) in graph a
};

__device__ __forceinline__ GraphTile k2_tile(const tsd_batch_t& b, int g, unsigned* smem_u) {
  GraphTile t;
  t.n0 = b.graph_ptr[g];
  t.n = b.graph_ptr[g + 1] - t.n0;
  t.base = b.pair_ptr[g];
  t.W = (b.max_graph_nodes + 31) >> 5;
  int nmax = b.max_graph_nodes;
  t.spos = reinterpret_cast<float*>(smem_u);
  t.radin = smem_u + 3 * nmax;
  t.out_a = t.radin + nmax * t.W;
  t.in_a = t.out_a + nmax * t.W;
  return t;
}

__device__ void k2_build_masks(const GraphTile& t, const float* __restrict__ pos, float r2, int cap,
                               const int* __restrict__ table0) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int n = t.n, W = t.W;
  for (int i = threadIdx.x; i < 3 * n; i += blockDim.x) t.spos[i] = pos[(size_t)3 * t.n0 + i];
  __syncthreads();
  // phase 1: torch_cluster radius rule -- per centre c keep the first `cap` in-range atoms (index
  // order, self included), then drop self.
  const unsigned lt = tsd_lanemask_lt();
  for (int c = warp; c < n; c += nwarps) {
    float cx = t.spos[3 * c], cy = t.spos[3 * c + 1], cz = t.spos[3 * c + 2];
    int running = 0;
    for (int w = 0; w < W; ++w) {
      int j = w * 32 + lane;
      bool inr = false;
      if (j < n) inr = tsd_dist2(t.spos[3 * j], t.spos[3 * j + 1], t.spos[3 * j + 2], cx, cy, cz) < r2;
      unsigned bits = __ballot_sync(TSD_FULL_MASK, inr);
      int rank = running + __popc(bits & lt);
      bool keep = inr && rank < cap && j != c;
      unsigned kb = __ballot_sync(TSD_FULL_MASK, keep);
      if (lane == 0) t.radin[c * W + w] = kb;
      running += __popc(bits);
    }
  }
  __syncthreads();
  // phase 2: union with the local (bond-order) pairs of table 0
  for (int r = warp; r < n; r += nwarps)
    for (int w = 0; w < W; ++w) {
      int c = w * 32 + lane;
      bool o = false, in_ = false;
      if (c < n && c != r) {
        bool rad_out = (t.radin[c * W + (r >> 5)] >> (r & 31)) & 1u;  // r -> c
        bool rad_in = (t.radin[r * W + w] >> lane) & 1u;              // c -> r
        o = rad_out || table0[t.base + r * n + c] != 0;
        in_ = rad_in || table0[t.base + c * n + r] != 0;
      }
      unsigned ob = __ballot_sync(TSD_FULL_MASK, o), ib = __ballot_sync(TSD_FULL_MASK, in_);
      if (lane == 0) {
        t.out_a[r * W + w] = ob;
        t.in_a[r * W + w] = ib;
      }
    }
  __syncthreads();
}

__device__ __forceinline__ int k2_row_popc(const unsigned* m, int W) {
  int s = 0;
  for (int w = 0; w < W; ++w) s += __popc(m[w]);
  return s;
}

// bits j > r of word w of row r: the upper triangle of the pair matrix
__device__ __forceinline__ unsigned k2_upper_mask(int r, int w) {
  const int rw = r >> 5;
  if (w < rw) return 0u;
  if (w > rw) return 0xffffffffu;
  return ~((2u << (r & 31)) - 1u);  // r & 31 == 31: 2u << 31 wraps to 0 -> mask 0
}

// word w of the undirected pair mask of row r: j > r with (r -> j) or (j -> r) in graph a
__device__ __forceinline__ unsigned k2_upair_word(const GraphTile& t, int r, int w) {
  return (t.out_a[r * t.W + w] | t.in_a[r * t.W + w]) & k2_upper_mask(r, w);
}

// number of set bits of the row mask below column c
__device__ __forceinline__ int k2_rank_below(const unsigned* row_mask, int c) {
  int rank = 0;
  for (int w2 = 0; w2 < (c >> 5); ++w2) rank += __popc(row_mask[w2]);
  return rank + __popc(row_mask[c >> 5] & ((1u << (c & 31)) - 1u));
}

__device__ int k2_block_sum(int v, int* s_red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(TSD_FULL_MASK, v, o);
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  int tot = 0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_red[w];
  return tot;
}

// pass 1: edges per graph
__global__ void __launch_bounds__(256) k_edge_count(tsd_batch_t b, const float* __restrict__ pos, float r2, int cap,
                                                    const int* __restrict__ table0, int* __restrict__ graph_count,
                                                    int* __restrict__ graph_ucount) {
  extern __shared__ unsigned smem_u[];
  __shared__ int s_red[8];
  GraphTile t = k2_tile(b, blockIdx.x, smem_u);
  k2_build_masks(t, pos, r2, cap, table0);
  int local = 0, ulocal = 0;
  for (int r = threadIdx.x; r < t.n; r += blockDim.x) {
    local += k2_row_popc(t.out_a + r * t.W, t.W);
    if (graph_ucount)
      for (int w = 0; w < t.W; ++w) ulocal += __popc(k2_upair_word(t, r, w));
  }
  int tot = k2_block_sum(local, s_red);
  if (threadIdx.x == 0) graph_count[blockIdx.x] = tot;
  if (graph_ucount) {
    int utot = k2_block_sum(ulocal, s_red);
    if (threadIdx.x == 0) graph_ucount[blockIdx.x] = utot;
  }
}

// exclusive scan of deg(r) over the rows of one graph by one warp
__device__ void k2_warp_scan_rows(const unsigned* mask, int n, int W, int* off) {
  const int lane = threadIdx.x & 31;
  int carry = 0;
  for (int chunk = 0; chunk < n; chunk += 32) {
    int r = chunk + lane;
    int d = r < n ? k2_row_popc(mask + r * W, W) : 0;
    int incl = d;
    for (int o = 1; o < 32; o <<= 1) {
      int v = __shfl_up_sync(TSD_FULL_MASK, incl, o);
      if (lane >= o) incl += v;
    }
    if (r < n) off[r] = carry + incl - d;
    carry += __shfl_sync(TSD_FULL_MASK, incl, 31);
  }
  if (lane == 0) off[n] = carry;
}

// pass 2: emit the row-major sorted edge list + out-CSR + dst-sorted in-CSR
__global__ void __launch_bounds__(256) k_edge_emit(tsd_batch_t b, const float* __restrict__ pos, float r2, int cap,
                                                   const int* __restrict__ table0, const int* __restrict__ table1,
                                                   int tab1_is_graph, tsd_edges_t e) {
  extern __shared__ unsigned smem_u[];
  __shared__ int s_red[8];
  const int g = blockIdx.x;
  GraphTile t = k2_tile(b, g, smem_u);
  int* row_off = reinterpret_cast<int*>(t.in_a + b.max_graph_nodes * t.W);
  int* in_off = row_off + b.max_graph_nodes + 1;
  int* u_off = in_off + b.max_graph_nodes + 1;
  unsigned* pm = reinterpret_cast<unsigned*>(u_off + b.max_graph_nodes + 1);  // n*W undirected pair masks
  k2_build_masks(t, pos, r2, cap, table0);
  const bool upairs = e.num_upairs != nullptr;

  int part = 0, upart = 0;
  for (int i = threadIdx.x; i < g; i += blockDim.x) {
    part += e.graph_count[i];
    if (upairs) upart += e.graph_ucount[i];
  }
  const int gbase = k2_block_sum(part, s_red);
  const int ubase = upairs ? k2_block_sum(upart, s_red) : 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int n = t.n, W = t.W;
  if (upairs) {
    for (int idx = threadIdx.x; idx < n * W; idx += blockDim.x) pm[idx] = k2_upair_word(t, idx / W, idx % W);
    __syncthreads();
  }
  if (warp == 0) k2_warp_scan_rows(t.out_a, n, W, row_off);
  if (warp == 1) k2_warp_scan_rows(t.in_a, n, W, in_off);
  if (warp == 2 && upairs) k2_warp_scan_rows(pm, n, W, u_off);
  __syncthreads();
  for (int r = threadIdx.x; r < n; r += blockDim.x) {
    e.row_ptr[t.n0 + r] = gbase + row_off[r];
    e.in_ptr[t.n0 + r] = gbase + in_off[r];
  }
  if (g == b.num_graphs - 1 && threadIdx.x == 0) {
    int total = gbase + row_off[n];
    e.row_ptr[b.num_nodes] = total;
    e.in_ptr[b.num_nodes] = total;
    e.num_edges[0] = total;
    if (upairs) e.num_upairs[0] = ubase + u_off[n];
  }
  const unsigned lt = tsd_lanemask_lt();
  // undirected pairs {i < j}, warp per row i: the same length / table expressions as the directed
  // edges below, so a pair's values are bit-identical to those of both of its directions
  if (upairs) {
    for (int i = warp; i < n; i += nwarps) {
      float ix = t.spos[3 * i], iy = t.spos[3 * i + 1], iz = t.spos[3 * i + 2];
      int running = ubase + u_off[i];
      for (int w = 0; w < W; ++w) {
        unsigned bits = pm[i * W + w];
        if ((bits >> lane) & 1u) {
          int j = w * 32 + lane;
          int uid = running + __popc(bits & lt);
          e.u_row[uid] = t.n0 + i;
          e.u_col[uid] = t.n0 + j;
          e.u_length[uid] = __fsqrt_rn(tsd_dist2(ix, iy, iz, t.spos[3 * j], t.spos[3 * j + 1], t.spos[3 * j + 2]));
          int pidx = t.base + i * n + j;
          e.u_tab0[uid] = table0[pidx];
          if (e.u_tab1) e.u_tab1[uid] = table1 ? table1[pidx] : 0;
        }
        running += __popc(bits);
      }
    }
  }
  // out-edges, warp per row
  for (int r = warp; r < n; r += nwarps) {
    float rx = t.spos[3 * r], ry = t.spos[3 * r + 1], rz = t.spos[3 * r + 2];
    int running = gbase + row_off[r];
    for (int w = 0; w < W; ++w) {
      unsigned bits = t.out_a[r * W + w];
      if ((bits >> lane) & 1u) {
        int c = w * 32 + lane;
        int eid = running + __popc(bits & lt);
        float d2 = tsd_dist2(rx, ry, rz, t.spos[3 * c], t.spos[3 * c + 1], t.spos[3 * c + 2]);
        e.row[eid] = t.n0 + r;
        e.col[eid] = t.n0 + c;
        e.length[eid] = __fsqrt_rn(d2);
        int pidx = t.base + r * n + c;
        e.tab0[eid] = table0[pidx];
        int t1 = table1 ? table1[pidx] : 0;
        if (e.tab1) e.tab1[eid] = t1;
        if (e.in_b) {
          bool rad_out = (t.radin[c * W + (r >> 5)] >> (r & 31)) & 1u;
          e.in_b[eid] = tab1_is_graph ? ((t1 != 0 || rad_out) ? 1 : 0) : 1;
        }
        if (upairs) {
          const int i = min(r, c), j = max(r, c);
          e.edge_upair[eid] = ubase + u_off[i] + k2_rank_below(pm + i * W, j);
        }
      }
      running += __popc(bits);
    }
  }
  // in-CSR, warp per target
  for (int c = warp; c < n; c += nwarps) {
    int running = gbase + in_off[c];
    for (int w = 0; w < W; ++w) {
      unsigned bits = t.in_a[c * W + w];
      if ((bits >> lane) & 1u) {
        int r = w * 32 + lane;
        int k = running + __popc(bits & lt);
        // position of c inside row r's out list
        const unsigned* orow = t.out_a + r * W;
        int rank = 0;
        for (int w2 = 0; w2 < (c >> 5); ++w2) rank += __popc(orow[w2]);
        rank += __popc(orow[c >> 5] & ((1u << (c & 31)) - 1u));
        e.in_eid[k] = gbase + row_off[r] + rank;
        e.in_src[k] = t.n0 + r;
        if (upairs) {
          const int i = min(r, c), j = max(r, c);
          e.in_upair[k] = ubase + u_off[i] + k2_rank_below(pm + i * W, j);
        }
      }
      running += __popc(bits);
    }
  }
}

extern "C" int tsd_edge_build(const tsd_batch_t* batch, const float* pos, double cutoff, int32_t max_neighbors,
                              const int32_t* table0, const int32_t* table1, int32_t tab1_is_graph,
                              const tsd_edges_t* edges, tsd_stream_t stream) {
  TSD_REQUIRE(batch && pos && table0 && edges);
  TSD_REQUIRE(edges->num_edges && edges->row && edges->col && edges->length && edges->tab0 && edges->row_ptr &&
              edges->in_ptr && edges->in_eid && edges->in_src && edges->graph_count);
  TSD_REQUIRE(batch->max_graph_nodes >= 1 && batch->max_graph_nodes <= TSD_MAX_GRAPH_NODES);
  TSD_REQUIRE(!tab1_is_graph || table1);
  if (batch->num_graphs == 0) return TSD_OK;
  cudaStream_t s = tsd_cu(stream);
  int nmax = batch->max_graph_nodes, W = tsd_ceil_div(nmax, 32);
  // torch_cluster squares the double radius on the host and casts to float
  float r2 = (float)(cutoff * cutoff);
  int cap = max_neighbors + 1;  // radius_graph(loop=False) asks for max_num_neighbors + 1 incl. self
  const bool upairs = edges->num_upairs != nullptr;
  if (upairs)
    TSD_REQUIRE(edges->u_row && edges->u_col && edges->u_length && edges->u_tab0 && edges->edge_upair && edges->in_upair &&
                edges->graph_ucount);
  size_t smem = (size_t)(3 * nmax + 4 * nmax * W + 3 * (nmax + 1)) * sizeof(unsigned);
  k_edge_count<<<batch->num_graphs, 256, smem, s>>>(*batch, pos, r2, cap, table0, edges->graph_count,
                                                    upairs ? edges->graph_ucount : nullptr);
  TSD_LAUNCH_CHECK();
  k_edge_emit<<<batch->num_graphs, 256, smem, s>>>(*batch, pos, r2, cap, table0, table1, tab1_is_graph, *edges);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}
