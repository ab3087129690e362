// C-ABI entry points that compose the kernels into the reference's operator granularity
// (edge embedding, one SchNet interaction, one GINE conv, the output MLP).  See
// include/tsdiff_b200.h for the contract and the reference lines each call replaces.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "gemm.cuh"

int tsd_launch_cfconv_aggregate(int num_nodes, int H, const int* in_ptr, const int* in_eid, const int* row,
                                const float* x1, const float* filt, float* agg, cudaStream_t s);
int tsd_launch_cfconv_aggregate_staged(const tsd_batch_t* b, int H, const int* in_ptr, const int* in_eid,
                                       const int* in_src, const float* x1, const float* filt, float* agg,
                                       cudaStream_t s);

// CFConv aggregation, two kernels for two regimes (profiles/scripts/time_agg.py):
//   node-parallel (warp per target node and 128-channel slab): most memory-level parallelism for
//     a small edge list -- 20 us cold at batch 100 (E = 33.5k), where the graph-staged kernel
//     needs 33 us (only G x H/128 CTAs);
//   graph-staged (CTA per reaction and slab, the reaction's x1 rows in shared memory, in-CSR
//     bounds and next-node ids prefetched): the x1 gather stays on the SM, so only the filter
//     rows stream -- 4.64 TB/s = 72 % of the measured HBM copy bandwidth at BASELINE config 5
//     (1000 reactions x ~60 atoms, E = 2.3M) against 3.9 TB/s for the node-parallel kernel.
// The switch uses the host-side edge capacity (the valid count lives on the device).
static int tsd_aggregate(const tsd_batch_t* batch, const tsd_edges_t* edges, int H, const float* x1, const float* filt,
                         float* agg, cudaStream_t s) {
  const bool staged = batch->edge_capacity >= (1 << 18) && batch->max_graph_nodes <= 256;
  if (staged) return tsd_launch_cfconv_aggregate_staged(batch, H, edges->in_ptr, edges->in_eid, edges->in_src, x1, filt, agg, s);
  return tsd_launch_cfconv_aggregate(batch->num_nodes, H, edges->in_ptr, edges->in_eid, edges->in_src, x1, filt, agg, s);
}

int tsd_linear_generic(int M, int N, int K, const float* x, const float* w, const float* b, float* out, cudaStream_t s);

int tsd_launch_gine_aggregate(int num_nodes, int H, const int* in_ptr, const int* in_eid, const int* row,
                              const int* local_tab, const float* h, const float* ea, const float* eps, float* out,
                              cudaStream_t s);

#include <atomic>
#include <mutex>
static thread_local int g_last_cuda_error = 0;
static std::atomic<long long> g_launches{0};
void tsd_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" int64_t tsd_launch_count(void) { return (int64_t)g_launches.load(); }

int tsd_record_cuda_error(cudaError_t e) {
  g_last_cuda_error = (int)e;
  return TSD_ERR_CUDA;
}

extern "C" int tsd_last_cuda_error(void) { return g_last_cuda_error; }
extern "C" int tsd_version(void) { return 100; }

extern "C" int tsd_workspace_bytes(int32_t num_nodes, int32_t edge_capacity, int32_t hidden, int32_t network,
                                   int32_t math, uint64_t* edge_buffer_bytes, uint64_t* node_buffer_bytes,
                                   int32_t* edge_buffers, int32_t* node_buffers) {
  TSD_REQUIRE(num_nodes >= 0 && edge_capacity >= 0 && hidden > 0 && (network == 0 || network == 1));
  const uint64_t e = (uint64_t)(edge_capacity > 0 ? edge_capacity : 1), n = (uint64_t)(num_nodes > 0 ? num_nodes : 1);
  if (edge_buffer_bytes) *edge_buffer_bytes = e * (uint64_t)hidden * sizeof(float);
  if (node_buffer_bytes) *node_buffer_bytes = n * (uint64_t)hidden * sizeof(float);
  // edge: d_emb, tmp, edge_attr (graph a), edge_attr (graph b / local), two filter buffers, tmp of the second embedding
  if (edge_buffers) *edge_buffers = 7;
  // node: h, x1, agg, spare (+ h_local, x1_local, agg_local for the GIN branch of path A) + encoder pool (tf32)
  if (node_buffers) *node_buffers = (network == 0 ? 4 : 7) + (math == TSD_MATH_TF32 ? 2 : 0);
  return TSD_OK;
}

extern "C" const char* tsd_error_string(int code) {
  switch (code) {
    case TSD_OK: return "ok";
    case TSD_ERR_INVALID: return "invalid argument";
    case TSD_ERR_CUDA: return cudaGetErrorString((cudaError_t)g_last_cuda_error);
    case TSD_ERR_UNSUPPORTED: return "unsupported shape or mode";
    default: return "unknown error";
  }
}

int g_tsd_gemm_chain2 = 1;
extern "C" void tsd_tune_gemm_chain2(int on) { g_tsd_gemm_chain2 = on; }

int tsd_gemm(const GemmArgs& g, int math, cudaStream_t stream) {
  if (math == TSD_MATH_TF32) {
    int rc = tsd_gemm_tf32(g, stream);
    if (rc != TSD_ERR_UNSUPPORTED) return rc;  // shapes the tensor-core kernel does not take run on FFMA
  }
  return tsd_gemm_ffma(g, stream);
}

// tf32: layer `g` and the layer `next` chained on the tile (gemm_tc.cu; the intermediate never leaves tensor memory).
// `g` describes the first layer with the SECOND layer's output fields (C / ldc / round_out or w3 / b3 / out_vec).
// Returns TSD_ERR_UNSUPPORTED when the caller has to run the two layers as separate kernels.
static int tsd_gemm_chain2(GemmArgs g, const tsd_linear_t& next, int math, cudaStream_t stream) {
  if (math != TSD_MATH_TF32 || !g_tsd_gemm_chain2 || next.in_features != g.N) return TSD_ERR_UNSUPPORTED;
  g.W2 = next.weight;
  g.bias2 = next.bias;
  g.N2 = next.out_features;
  g.ldc = g.N2;
  return tsd_gemm_chain2_tf32(g, stream);
}

#define TSD_TRY(expr)        \
  do {                       \
    int _rc = (expr);        \
    if (_rc != TSD_OK) return _rc; \
  } while (0)

static GemmArgs edge_gemm(const tsd_batch_t* b, const tsd_edges_t* e, const tsd_linear_t& lin) {
  GemmArgs g = tsd_gemm_args();
  g.M_cap = b->edge_capacity;
  g.M_ptr = e->num_edges;
  g.N = lin.out_features;
  g.K = lin.in_features;
  g.W = lin.weight;
  g.bias = lin.bias;
  g.ldc = g.N;
  g.lda = g.K;
  return g;
}

static GemmArgs node_gemm(const tsd_batch_t* b, const tsd_linear_t& lin) {
  GemmArgs g = tsd_gemm_args();
  g.M_cap = b->num_nodes;
  g.N = lin.out_features;
  g.K = lin.in_features;
  g.W = lin.weight;
  g.bias = lin.bias;
  g.ldc = g.N;
  g.lda = g.K;
  return g;
}

extern "C" int tsd_edge_embed(const tsd_batch_t* batch, const tsd_edges_t* edges, const int32_t* code,
                              const tsd_edge_encoder_t* enc, int32_t reuse_d_emb, float* d_emb, float* tmp, float* out,
                              int32_t math, tsd_stream_t stream) {
  TSD_REQUIRE(batch && edges && code && enc && out && enc->bond_emb && enc->lin0.weight && enc->lin0.bias &&
              enc->lin1.weight);
  const int H = enc->lin1.out_features;
  TSD_REQUIRE(enc->lin0.in_features == 1 && enc->lin0.out_features == enc->lin1.in_features);
  cudaStream_t s = tsd_cu(stream);
  const bool cat = enc->cat0 != nullptr;
  if (!cat || !reuse_d_emb) {
    GemmArgs g = edge_gemm(batch, edges, enc->lin1);
    g.a_kind = TSD_A_EDGE_MLP0;
    g.len = edges->length;
    g.w0 = enc->lin0.weight;
    g.b0 = enc->lin0.bias;
    g.act0 = enc->act;
    g.H = H;
    if (cat) {
      TSD_REQUIRE(d_emb);
      g.C = d_emb;
    } else {  // edge.py:66-68: d_emb * bond_emb[type]
      g.mul_emb = enc->bond_emb;
      g.mul_code = code;
      g.C = out;
      g.round_out = 1;
    }
    TSD_TRY(tsd_gemm(g, math, s));
  }
  if (cat) {
    TSD_REQUIRE(enc->cat2 && tmp && d_emb);
    TSD_REQUIRE(enc->cat0->in_features == 2 * H && enc->cat0->out_features == H);
    GemmArgs g = edge_gemm(batch, edges, *enc->cat0);
    g.a_kind = TSD_A_CAT;
    g.A = d_emb;
    g.lda = H;
    g.H = H;
    g.emb = enc->bond_emb;
    g.code = code;
    g.act = enc->cat_act;
    g.round_out = 1;  // feeds cat2 / edge_attr feeds the filter networks and the pair MLP
    g.C = out;
    int rc = tsd_gemm_chain2(g, *enc->cat2, math, s);  // cat0 -> cat2 on the tile
    if (rc != TSD_ERR_UNSUPPORTED) return rc;
    g.C = tmp;
    TSD_TRY(tsd_gemm(g, math, s));
    GemmArgs g2 = edge_gemm(batch, edges, *enc->cat2);
    g2.A = tmp;
    g2.C = out;
    g2.round_out = 1;
    TSD_TRY(tsd_gemm(g2, math, s));
  }
  return TSD_OK;
}

// Rows whose two packed type codes differ, in ascending order: diff_rows[0 .. *diff_count), and for every row its
// position in that list or -1.  One CTA walks the rows in chunks of 1024 (ballot ranks + a 32-entry scan per chunk).
__global__ void __launch_bounds__(1024) k_code_delta(const int* __restrict__ rows_dev, int rows_cap,
                                                     const int* __restrict__ code0, const int* __restrict__ code1,
                                                     int* __restrict__ diff_rows, int* __restrict__ diff_pos,
                                                     int* __restrict__ diff_count) {
  __shared__ int warp_off[32];
  __shared__ int chunk_total;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int M = min(*rows_dev, rows_cap);
  int base = 0;
  for (int m0 = 0; m0 < M; m0 += 1024) {
    const int m = m0 + tid;
    const bool d = m < M && code0[m] != code1[m];
    const unsigned bal = __ballot_sync(TSD_FULL_MASK, d);
    if (lane == 0) warp_off[warp] = __popc(bal);
    __syncthreads();
    if (warp == 0) {
      const int c = warp_off[lane];
      int inc = c;
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(TSD_FULL_MASK, inc, o);
        if (lane >= o) inc += t;
      }
      warp_off[lane] = inc - c;
      if (lane == 31) chunk_total = inc;
    }
    __syncthreads();
    if (m < M) {
      const int pos = base + warp_off[warp] + __popc(bal & tsd_lanemask_lt());
      diff_pos[m] = d ? pos : -1;
      if (d) diff_rows[pos] = m;
    }
    base += chunk_total;
    __syncthreads();
  }
  if (tid == 0) *diff_count = base;
}

// Second graph of path B (condensenc.py:219-234): its edge embedding differs from the first graph's only on the rows
// whose type codes differ (the 4-hop pairs when edge_order = 4 and pred_edge_order = 3: a fifth of the pairs), so
// edge_cat runs on the compact list of those rows and consumers read `out_compact[diff_pos[m]]` where diff_pos[m] >= 0
// and the first graph's edge_attr elsewhere (tsd_pair_mlp_delta).  Same arithmetic per row as tsd_edge_embed.
extern "C" int tsd_edge_embed_delta(const tsd_batch_t* batch, const tsd_edges_t* edges, const int32_t* code0,
                                    const int32_t* code1, const tsd_edge_encoder_t* enc, const float* d_emb, float* tmp,
                                    float* out_compact, int32_t* diff_rows, int32_t* diff_pos, int32_t* diff_count,
                                    int32_t math, tsd_stream_t stream) {
  TSD_REQUIRE(batch && edges && code0 && code1 && enc && enc->cat0 && enc->cat2 && enc->bond_emb && d_emb && tmp &&
              out_compact && diff_rows && diff_pos && diff_count);
  const int H = enc->lin1.out_features;
  TSD_REQUIRE(enc->cat0->in_features == 2 * H && enc->cat0->out_features == H);
  cudaStream_t s = tsd_cu(stream);
  if (batch->edge_capacity == 0) return TSD_OK;
  k_code_delta<<<1, 1024, 0, s>>>(edges->num_edges, batch->edge_capacity, code0, code1, diff_rows, diff_pos, diff_count);
  TSD_LAUNCH_CHECK();
  GemmArgs g = edge_gemm(batch, edges, *enc->cat0);
  g.M_ptr = diff_count;
  g.a_kind = TSD_A_CAT;
  g.A = d_emb;
  g.lda = H;
  g.H = H;
  g.emb = enc->bond_emb;
  g.code = code1;
  g.row_index = diff_rows;
  g.act = enc->cat_act;
  g.round_out = 1;
  g.C = out_compact;
  int rc = tsd_gemm_chain2(g, *enc->cat2, math, s);
  if (rc != TSD_ERR_UNSUPPORTED) return rc;
  g.C = tmp;
  TSD_TRY(tsd_gemm(g, math, s));
  GemmArgs g2 = edge_gemm(batch, edges, *enc->cat2);
  g2.M_ptr = diff_count;
  g2.A = tmp;
  g2.C = out_compact;
  g2.round_out = 1;
  return tsd_gemm(g2, math, s);
}

extern "C" int tsd_cfconv_layer(const tsd_batch_t* batch, const tsd_edges_t* edges, const float* edge_attr,
                                const tsd_interaction_t* blk, const float* h_in, float* h_out, float* ef0, float* ef1,
                                float* nf0, float* nf1, float* nf2, int32_t math, tsd_stream_t stream) {
  TSD_REQUIRE(batch && edges && edge_attr && blk && h_in && h_out && ef0 && ef1 && nf0 && nf1 && nf2);
  cudaStream_t s = tsd_cu(stream);
  const int F = blk->nn2.out_features;
  TSD_REQUIRE(blk->lin1.out_features == F && blk->lin2.in_features == F);
  // filter network on the edges: W = nn2(ssp(nn0(edge_attr))) * C(len)
  GemmArgs g = edge_gemm(batch, edges, blk->nn0);
  g.A = edge_attr;
  g.act = TSD_ACT_SSP;
  g.C = ef0;
  g.round_out = 1;  // feeds nn2
  TSD_TRY(tsd_gemm(g, math, s));
  g = edge_gemm(batch, edges, blk->nn2);
  g.A = ef0;
  g.scale_len = edges->length;
  g.cutoff = blk->cutoff;
  g.smooth = blk->smooth;
  g.C = ef1;
  TSD_TRY(tsd_gemm(g, math, s));
  // x1 = lin1(h) (no bias)
  g = node_gemm(batch, blk->lin1);
  g.A = h_in;
  g.C = nf0;
  TSD_TRY(tsd_gemm(g, math, s));
  TSD_TRY(tsd_aggregate(batch, edges, F, nf0, ef1, nf1, s));
  // h_out = h_in + lin(ssp(lin2(agg)))
  g = node_gemm(batch, blk->lin2);
  g.A = nf1;
  g.act = TSD_ACT_SSP;
  g.C = nf2;
  TSD_TRY(tsd_gemm(g, math, s));
  g = node_gemm(batch, blk->lin);
  g.A = nf2;
  g.residual = h_in;
  g.ldr = blk->lin.out_features;
  g.C = h_out;
  TSD_TRY(tsd_gemm(g, math, s));
  return TSD_OK;
}

extern "C" int tsd_gine_layer(const tsd_batch_t* batch, const tsd_edges_t* edges, const float* edge_attr,
                              const tsd_gine_t* conv, const float* h_in, float* h_out, float* nf0, float* nf1,
                              int32_t math, tsd_stream_t stream) {
  TSD_REQUIRE(batch && edges && edge_attr && conv && conv->eps && h_in && h_out && nf0 && nf1);
  cudaStream_t s = tsd_cu(stream);
  const int H = conv->nn0.in_features;
  TSD_TRY(tsd_launch_gine_aggregate(batch->num_nodes, H, edges->in_ptr, edges->in_eid, edges->in_src, edges->tab0, h_in,
                                    edge_attr, conv->eps, nf0, s));
  GemmArgs g = node_gemm(batch, conv->nn0);
  g.A = nf0;
  g.act = TSD_ACT_RELU;
  g.C = nf1;
  TSD_TRY(tsd_gemm(g, math, s));
  g = node_gemm(batch, conv->nn1);
  g.A = nf1;
  g.act = conv->relu_after ? TSD_ACT_RELU : TSD_ACT_NONE;
  g.residual = h_in;
  g.ldr = conv->nn1.out_features;
  g.C = h_out;
  TSD_TRY(tsd_gemm(g, math, s));
  return TSD_OK;
}

extern "C" int tsd_pair_mlp(const tsd_batch_t* batch, const tsd_edges_t* edges, const float* h, const float* edge_attr,
                            const tsd_pair_mlp_t* mlp, int32_t accumulate, float* ef0, float* edge_inv, int32_t math,
                            tsd_stream_t stream) {
  return tsd_pair_mlp_delta(batch, edges, h, edge_attr, nullptr, nullptr, mlp, accumulate, ef0, edge_inv, math, stream);
}

extern "C" int tsd_pair_mlp_delta(const tsd_batch_t* batch, const tsd_edges_t* edges, const float* h,
                                  const float* edge_attr, const float* alt_attr, const int32_t* alt_pos,
                                  const tsd_pair_mlp_t* mlp, int32_t accumulate, float* ef0, float* edge_inv, int32_t math,
                                  tsd_stream_t stream) {
  TSD_REQUIRE(batch && edges && h && edge_attr && mlp && ef0 && edge_inv && ((alt_attr == nullptr) == (alt_pos == nullptr)));
  TSD_REQUIRE(mlp->l2.out_features == 1 && mlp->l2.in_features == mlp->l1.out_features);
  cudaStream_t s = tsd_cu(stream);
  const int H = mlp->l0.in_features / 2;
  GemmArgs g = edge_gemm(batch, edges, mlp->l0);
  g.a_kind = TSD_A_PAIR;
  g.A = edge_attr;
  g.lda = H;
  g.H = H;
  g.h = h;
  g.row = edges->row;
  g.col = edges->col;
  g.alt_A = alt_attr;
  g.alt_pos = alt_pos;
  g.act = mlp->act;
  {
    GemmArgs c = g;  // l0 -> l1 -> row-dot with l2 on the tile
    c.w3 = mlp->l2.weight;
    c.b3 = mlp->l2.bias;
    c.out_vec = edge_inv;
    c.accumulate = accumulate;
    int rc = tsd_gemm_chain2(c, mlp->l1, math, s);
    if (rc != TSD_ERR_UNSUPPORTED) return rc;
  }
  g.C = ef0;
  g.round_out = 1;  // feeds l1
  TSD_TRY(tsd_gemm(g, math, s));
  g = edge_gemm(batch, edges, mlp->l1);
  g.A = ef0;
  g.act = mlp->act;
  g.w3 = mlp->l2.weight;
  g.b3 = mlp->l2.bias;
  g.out_vec = edge_inv;
  g.accumulate = accumulate;
  TSD_TRY(tsd_gemm(g, math, s));
  return TSD_OK;
}

extern "C" int tsd_linear(int32_t rows, const int32_t* rows_dev, const float* x, const tsd_linear_t* lin, int32_t act,
                          float* out, int32_t math, tsd_stream_t stream) {
  TSD_REQUIRE(x && lin && lin->weight && out && rows >= 0);
  if (lin->out_features % 64 != 0 || lin->in_features % 16 != 0) {
    // odd shapes (K = 1 or 25, N = 1: edge MLP layer 0, the feature embedding, the last output layer and their data
    // gradients in the training step): a plain fp32 kernel, one thread per output
    TSD_REQUIRE(rows_dev == nullptr && act == TSD_ACT_NONE);
    return tsd_linear_generic(rows, lin->out_features, lin->in_features, x, lin->weight, lin->bias, out, tsd_cu(stream));
  }
  GemmArgs g = tsd_gemm_args();
  g.M_cap = rows;
  g.M_ptr = rows_dev;
  g.N = lin->out_features;
  g.K = lin->in_features;
  g.W = lin->weight;
  g.bias = lin->bias;
  g.A = x;
  g.lda = g.K;
  g.C = out;
  g.ldc = g.N;
  g.act = act;
  return tsd_gemm(g, math, tsd_cu(stream));
}

extern "C" int tsd_cfconv_aggregate(const tsd_batch_t* batch, const tsd_edges_t* edges, int32_t channels,
                                    const float* x1, const float* filt, float* agg, tsd_stream_t stream) {
  TSD_REQUIRE(batch && edges && x1 && filt && agg);
  return tsd_aggregate(batch, edges, channels, x1, filt, agg, tsd_cu(stream));
}

// The node side of one interaction block as ONE kernel (node_update.cu): aggregation fused in front of the linears.
extern "C" int tsd_interaction_node_update(const tsd_batch_t* batch, const tsd_edges_t* edges, const tsd_interaction_t* blk,
                                           const tsd_linear_t* next_lin1, const float* x1, const float* filt,
                                           const float* h_in, float* h_out, float* x1_next, tsd_stream_t stream) {
  TSD_REQUIRE(batch && edges && blk && x1 && filt && h_in && h_out && ((next_lin1 == nullptr) == (x1_next == nullptr)));
  TSD_REQUIRE(x1_next != x1);  // other CTAs still gather x1 rows while this one writes
  const int H = blk->lin.out_features;
  TSD_REQUIRE(blk->lin2.in_features == H && blk->lin2.out_features == H && blk->lin.in_features == H);
  NodeArgs na;
  memset(&na, 0, sizeof(na));
  int npc = 0;
  const int tile = tsd_node_tile(true, batch->num_nodes, &npc);  // a single launch has the GPU to itself
  na.num_nodes = batch->num_nodes;
  na.nodes_per_cluster = npc;
  na.H = H;
  na.in_ptr = edges->in_ptr;
  na.in_eid = edges->in_eid;
  na.in_src = edges->in_src;
  na.x1 = x1;
  na.filt = filt;
  na.st[0].W = blk->lin2.weight;
  na.st[0].bias = blk->lin2.bias;
  na.st[0].act = TSD_ACT_SSP;
  na.st[1].W = blk->lin.weight;
  na.st[1].bias = blk->lin.bias;
  na.st[1].residual = h_in;
  na.st[1].store = h_out;
  na.num_stages = 2;
  if (next_lin1) {
    TSD_REQUIRE(next_lin1->in_features == H && next_lin1->out_features == H);
    na.st[2].W = next_lin1->weight;
    na.st[2].bias = next_lin1->bias;
    na.st[2].store = x1_next;
    na.num_stages = 3;
  }
  return tsd_node_update_tf32(na, tile, tsd_cu(stream));
}

// library-owned side stream + events for the fork/join inside tsd_schnet_encoder
struct EncoderFork {
  static const int MAX_BLOCKS = 32;
  cudaStream_t side = nullptr, side2 = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr, join2 = nullptr, xh_init = nullptr;
  cudaEvent_t edge_done[MAX_BLOCKS], agg_done[MAX_BLOCKS], n2_done[MAX_BLOCKS];
  bool ready = false;
  unsigned int* nc_words = nullptr;  // node chain (node_chain.cu): [0] grid barrier counter, [1] error flag
  int init(int num_blocks) {
    if (num_blocks > MAX_BLOCKS) return TSD_ERR_UNSUPPORTED;
    if (ready) return TSD_OK;

    // the node-side chain is serial and short (15-CTA kernels): give it the highest priority so its
    // CTAs are scheduled as soon as an SM frees up instead of queueing behind the edge tiles
    int prio_lo = 0, prio_hi = 0;
    TSD_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    TSD_CUDA(cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, prio_hi));
    TSD_CUDA(cudaStreamCreateWithPriority(&side2, cudaStreamNonBlocking, prio_hi));
    TSD_CUDA(cudaEventCreateWithFlags(&join2, cudaEventDisableTiming));
    TSD_CUDA(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
    TSD_CUDA(cudaEventCreateWithFlags(&join, cudaEventDisableTiming));
    TSD_CUDA(cudaEventCreateWithFlags(&xh_init, cudaEventDisableTiming));
    for (int i = 0; i < MAX_BLOCKS; ++i) {
      TSD_CUDA(cudaEventCreateWithFlags(&edge_done[i], cudaEventDisableTiming));
      TSD_CUDA(cudaEventCreateWithFlags(&agg_done[i], cudaEventDisableTiming));
      TSD_CUDA(cudaEventCreateWithFlags(&n2_done[i], cudaEventDisableTiming));
    }
    ready = true;
    return TSD_OK;
  }
};

static ChainStage chain_stage(const tsd_linear_t& lin, int act) {
  ChainStage st;
  memset(&st, 0, sizeof(st));
  st.W = lin.weight;
  st.bias = lin.bias;
  st.act = act;
  return st;
}

// Filter network of one CFConv on its own (schnet.py:91-98): filt = nn2(ssp(nn0(edge_attr))) * C(len).
// tf32 mode: one chained tensor-core kernel; fp32 mode: two FFMA GEMMs through `tmp`.
extern "C" int tsd_filter_network(const tsd_batch_t* batch, const tsd_edges_t* edges, const float* edge_attr,
                                  const tsd_interaction_t* blk, float* tmp, float* filt, int32_t math,
                                  tsd_stream_t stream) {
  TSD_REQUIRE(batch && edges && edge_attr && blk && tmp && filt);
  cudaStream_t s = tsd_cu(stream);
  const int H = blk->nn2.out_features;
  if (math == TSD_MATH_TF32 && (H == 128 || H == 256) && blk->nn0.in_features == H && blk->nn0.out_features == H &&
      blk->nn2.in_features == H && batch->edge_capacity >= 1024) {
    ChainArgs c;
    memset(&c, 0, sizeof(c));
    c.M_cap = batch->edge_capacity;
    c.M_ptr = edges->num_edges;
    c.H = H;
    c.A = edge_attr;
    c.num_stages = 2;
    c.st[0] = chain_stage(blk->nn0, TSD_ACT_SSP);
    c.st[1] = chain_stage(blk->nn2, TSD_ACT_NONE);
    c.st[1].scale_len = edges->length;
    c.st[1].cutoff = blk->cutoff;
    c.st[1].smooth = blk->smooth;
    c.st[1].store = filt;
    return tsd_chain_tf32(c, s);
  }
  GemmArgs g = edge_gemm(batch, edges, blk->nn0);
  g.A = edge_attr;
  g.act = TSD_ACT_SSP;
  g.C = tmp;
  g.round_out = 1;
  TSD_TRY(tsd_gemm(g, math, s));
  g = edge_gemm(batch, edges, blk->nn2);
  g.A = tmp;
  g.scale_len = edges->length;
  g.cutoff = blk->cutoff;
  g.smooth = blk->smooth;
  g.C = filt;
  return tsd_gemm(g, math, s);
}

// The filter networks of blocks [0, num_blocks) in ONE kernel (filter_stack.cu; tf32 only): filt[l] = nn2_l(ssp(nn0_l(
// edge_attr))) * C_l(len).  TSD_ERR_UNSUPPORTED when the shapes do not fit the kernel (callers fall back to one
// tsd_filter_network per block).
static int filter_stack_launch(const tsd_batch_t* batch, const tsd_edges_t* edges, const float* edge_attr,
                               const tsd_interaction_t* blocks, int32_t num_blocks, float* const* filt, int max_ctas,
                               tsd_stream_t stream) {
  TSD_REQUIRE(batch && edges && edge_attr && blocks && filt && num_blocks >= 1);
  if (num_blocks > TSD_FS_MAX_LAYERS) return TSD_ERR_UNSUPPORTED;
  const int H = blocks[0].nn2.out_features;
  FilterStackArgs a;
  memset(&a, 0, sizeof(a));
  a.M_cap = batch->edge_capacity;
  a.M_ptr = edges->num_edges;
  a.H = H;
  a.num_layers = num_blocks;
  a.A = edge_attr;
  a.len = edges->length;
  a.max_ctas = max_ctas;
  for (int l = 0; l < num_blocks; ++l) {
    const tsd_interaction_t& b = blocks[l];
    if (b.nn0.in_features != H || b.nn0.out_features != H || b.nn2.in_features != H || b.nn2.out_features != H)
      return TSD_ERR_UNSUPPORTED;
    TSD_REQUIRE(filt[l]);
    a.layer[l].W0 = b.nn0.weight;
    a.layer[l].b0 = b.nn0.bias;
    a.layer[l].W2 = b.nn2.weight;
    a.layer[l].b2 = b.nn2.bias;
    a.layer[l].cutoff = b.cutoff;
    a.layer[l].smooth = b.smooth;
    a.layer[l].out = filt[l];
  }
  return tsd_filter_stack_tf32(a, tsd_cu(stream));
}

extern "C" int tsd_filter_stack(const tsd_batch_t* batch, const tsd_edges_t* edges, const float* edge_attr,
                                const tsd_interaction_t* blocks, int32_t num_blocks, float* const* filt,
                                tsd_stream_t stream) {
  return filter_stack_launch(batch, edges, edge_attr, blocks, num_blocks, filt, 0, stream);
}

// tuning hook of profiles/scripts (not part of the C-ABI header): -1 = one filter kernel per block (round-2 path),
// 0 = all blocks in one launch, k > 0 = two launches, blocks [0, k) and [k, L)
static int g_filter_stack_mode = 0, g_filter_stack_ctas2 = 0, g_node_chain = 0;
extern "C" void tsd_tune_node_chain(int on) { g_node_chain = on; }
static unsigned int* g_nc_words_for_flag = nullptr;
// debugging aid: the persistent node chain's error flag (4 = its grid barrier timed out); synchronises the device
extern "C" int tsd_node_chain_flag(void) {
  unsigned int v = 0;
  if (g_nc_words_for_flag && cudaMemcpy(&v, g_nc_words_for_flag + 1, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (int)v;
}
extern "C" void tsd_tune_filter_stack_ctas2(int ctas) { g_filter_stack_ctas2 = ctas; }
// programmatic dependent launch of the edge-side GEMM kernels (gemm_tc.cu, gemm_chain.cu)
int g_tsd_gemm_pdl = 1;
extern "C" void tsd_tune_gemm_pdl(int on) { g_tsd_gemm_pdl = on; }
extern "C" void tsd_tune_filter_stack(int code) { g_filter_stack_mode = code; }

// Whole SchNet encoder (schnet.py:203-225).  fp32 mode: one tsd_cfconv_layer per block.  tf32
// mode: per block ONE chained filter-network kernel on the edges, the segmented aggregation, and
// chained node kernels -- no (E,H) / (N,H) intermediate round trips.
extern "C" int tsd_schnet_encoder(const tsd_batch_t* batch, const tsd_edges_t* edges, const float* edge_attr,
                                  const tsd_interaction_t* blocks, int32_t num_blocks, const float* h_in, float* h_out,
                                  float* ef0, float* ef1, float* nf0, float* nf1, float* nf2, float* nf_pool,
                                  int32_t nf_pool_count, float* ef_pool, int32_t ef_pool_count, float* x1_first,
                                  int32_t x1_first_valid, int32_t math, tsd_stream_t stream) {
  TSD_REQUIRE(batch && edges && edge_attr && blocks && num_blocks >= 1 && h_in && h_out && ef0 && ef1 && nf0 && nf1 && nf2);
  cudaStream_t s = tsd_cu(stream);
  const int H = blocks[0].lin.out_features;
  bool chain_ok = math == TSD_MATH_TF32 && (H == 128 || H == 256) && batch->num_nodes >= 1024 &&
                  batch->edge_capacity >= 1024;
  for (int l = 0; l < num_blocks && chain_ok; ++l) {
    const tsd_interaction_t& b = blocks[l];
    chain_ok = b.nn0.in_features == H && b.nn0.out_features == H && b.nn2.in_features == H && b.nn2.out_features == H &&
               b.lin1.in_features == H && b.lin1.out_features == H && b.lin2.in_features == H &&
               b.lin2.out_features == H && b.lin.in_features == H && b.lin.out_features == H;
  }
  if (!chain_ok) {
    const float* h = h_in;
    for (int l = 0; l < num_blocks; ++l) {
      TSD_TRY(tsd_cfconv_layer(batch, edges, edge_attr, &blocks[l], h, h_out, ef0, ef1, nf0, nf1, nf2, math, stream));
      h = h_out;
    }
    return TSD_OK;
  }
  // Dependency structure of a block l:   filter_l (edges; depends on edge_attr only)
  //                                      agg_l    (needs filter_l and x1_l)
  //                                      node_l   (needs agg_l; produces h_{l+1} and x1_{l+1})
  // The serial chain is agg_l -> node_l -> agg_{l+1}: 7 x (aggregate + 3 chained GEMMs).  Streams:
  //   caller's stream : the filter kernels back to back, alternating between two filter buffers
  //                     (their second wave leaves most SMs idle, which is where the node side runs)
  //   side            : x1_0, then per block agg_l and the CRITICAL node kernel
  //   side2           : the node work that is NOT needed by the next aggregation (split mode)
  // Split mode (fused weights + two extra node buffers): the next aggregation only needs
  //   x1_{l+1} = lin1_{l+1}(h_l + lin_l(y)) = xh_l + fused_w y + fused_b,  y = ssp(lin2_l(agg_l)),
  // where xh_l = lin1_{l+1}(h_l) is known one block EARLIER.  So the critical kernel is two chained
  // GEMMs (lin2, fused), and a second kernel on side2 computes h_{l+1} = h_l + lin_l(y) and
  // xh_{l+1} = lin1_{l+2}(h_{l+1}) beside agg_{l+1}.  Inside a CUDA-graph capture the event waits
  // become graph edges.
  // one set of library-owned side streams / events per process: the enqueue of a whole encoder is serialised
  // (it is host work of ~60 launches); a later call may re-record the events, waits already enqueued keep
  // the state they captured
  static EncoderFork fk;
  static std::mutex fk_mutex;
  std::lock_guard<std::mutex> fk_lock(fk_mutex);
  TSD_TRY(fk.init(num_blocks));
  cudaStream_t side = fk.side, side2 = fk.side2;
  if (!TSD_EXP_OLD_NODE && tsd_ceil_div(batch->num_nodes, 16) <= 4 * 148) {
    // Few atoms (batch 100: ~1750): the node side of every block is ONE kernel of N/NT small CTAs -- the
    // aggregation fused in front of the three linears as transposed (swap-AB) tensor-core GEMMs
    // (node_update.cu).  x1 ping-pongs between nf0 and nf1: a CTA's aggregation gathers x1 rows of atoms
    // that other CTAs own, so the next block's x1 must not overwrite them.
    int npc = 0;
    float* x1buf[2] = {nf0, nf1};
    // filter buffers: ef1, ef0, then the optional pool (one per block lets every filter kernel run ahead)
    float* fbuf[EncoderFork::MAX_BLOCKS];
    int nbuf = 0;
    fbuf[nbuf++] = ef1;
    fbuf[nbuf++] = ef0;
    const size_t edge_elems = (size_t)(batch->edge_capacity > 0 ? batch->edge_capacity : 1) * H;
    for (int i = 0; TSD_EXP_FILTER_POOL && ef_pool && i < ef_pool_count && nbuf < num_blocks; ++i)
      fbuf[nbuf++] = ef_pool + (size_t)i * edge_elems;
    TSD_CUDA(cudaEventRecord(fk.fork, s));
    TSD_CUDA(cudaStreamWaitEvent(side, fk.fork, 0));
    NodeArgs na;
    memset(&na, 0, sizeof(na));
    na.num_nodes = batch->num_nodes;
    na.nodes_per_cluster = npc;
    na.H = H;
    na.x = h_in;  // x1 of block 0
    na.num_stages = 1;
    na.st[0].W = blocks[0].lin1.weight;
    float* const x1_block0 = x1_first ? x1_first : x1buf[0];
    na.st[0].store = x1_block0;
    // With one filter buffer per block the filter networks of all blocks run as ONE kernel (filter_stack.cu) -- or as
    // two launches, so that the node chain starts after the first few blocks' filters -- ahead of the node side.
    bool stacked = g_filter_stack_mode >= 0 && nbuf >= num_blocks && num_blocks <= TSD_FS_MAX_LAYERS;
    int stack_cut = 0;
    if (stacked) {
      stack_cut = g_filter_stack_mode > 0 && g_filter_stack_mode < num_blocks ? g_filter_stack_mode : num_blocks;
      int rc = tsd_filter_stack(batch, edges, edge_attr, blocks, stack_cut, fbuf, stream);
      if (rc == TSD_ERR_UNSUPPORTED) {
        stacked = false;
      } else {
        TSD_TRY(rc);
        TSD_CUDA(cudaEventRecord(fk.edge_done[0], s));
        if (stack_cut < num_blocks) {
          TSD_TRY(filter_stack_launch(batch, edges, edge_attr, blocks + stack_cut, num_blocks - stack_cut, fbuf + stack_cut,
                                      g_filter_stack_ctas2, stream));
          TSD_CUDA(cudaEventRecord(fk.edge_done[stack_cut], s));
        }
      }
    }
    // Node kernel shape: behind the filter stack the node chain has the GPU to itself, so two CTAs share a 32-atom tile's
    // gathers; next to per-block filter kernels one CTA per tile competes least for SMs (profiles/r3_stack_modes.txt).
    int npc0 = 0;
    if (!(x1_first && x1_first_valid))  // x1 of block 0: dense input (skipped when the caller holds it from an earlier call)
      TSD_TRY(tsd_node_update_tf32(na, tsd_node_tile(false, batch->num_nodes, &npc0), side));
    const int tile = tsd_node_tile(stacked && stack_cut == num_blocks, batch->num_nodes, &npc);
    // Behind the complete filter stack the node side of ALL blocks is one persistent kernel (node_chain.cu) when every
    // cluster can be resident at once; otherwise one node kernel per block below.
    if (stacked && stack_cut == num_blocks && tile == 323 && H == 256 && g_node_chain && !fk.nc_words) {
      // the chain's two device words, allocated on first use -- never under stream capture (an allocation would
      // invalidate the capture): a first call inside a capture keeps one kernel per block
      cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
      if (cudaStreamIsCapturing(s, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone) {
        if (cudaMalloc(&fk.nc_words, 2 * sizeof(unsigned int)) != cudaSuccess ||
            cudaMemset(fk.nc_words, 0, 2 * sizeof(unsigned int)) != cudaSuccess) {
          (void)cudaGetLastError();
          fk.nc_words = nullptr;
        }
      }
    }
    if (stacked && stack_cut == num_blocks && tile == 323 && H == 256 && g_node_chain && fk.nc_words &&
        num_blocks <= TSD_NC_MAX_BLOCKS) {
      NodeChainArgs ca;
      memset(&ca, 0, sizeof(ca));
      ca.num_nodes = batch->num_nodes;
      ca.H = H;
      ca.num_blocks = num_blocks;
      ca.nodes_per_cluster = npc;
      ca.in_ptr = edges->in_ptr;
      ca.in_eid = edges->in_eid;
      ca.in_src = edges->in_src;
      ca.x1_first = x1_block0;
      ca.x1buf[0] = x1buf[0];
      ca.x1buf[1] = x1buf[1];
      ca.h_in = h_in;
      ca.h_out = h_out;
      ca.barrier = fk.nc_words;
      ca.error_flag = reinterpret_cast<int*>(fk.nc_words + 1);
      for (int l = 0; l < num_blocks; ++l) {
        ca.blk[l].filt = fbuf[l];
        ca.blk[l].w_lin2 = blocks[l].lin2.weight;
        ca.blk[l].b_lin2 = blocks[l].lin2.bias;
        ca.blk[l].w_lin = blocks[l].lin.weight;
        ca.blk[l].b_lin = blocks[l].lin.bias;
        ca.blk[l].w_lin1_next = l + 1 < num_blocks ? blocks[l + 1].lin1.weight : nullptr;
      }
      g_nc_words_for_flag = fk.nc_words;
      TSD_CUDA(cudaStreamWaitEvent(side, fk.edge_done[0], 0));
      TSD_CUDA(cudaMemsetAsync(fk.nc_words, 0, sizeof(unsigned int), side));
      const int rc = tsd_node_chain_tf32(ca, side);
      if (rc == TSD_OK) {
        TSD_CUDA(cudaEventRecord(fk.join, side));
        TSD_CUDA(cudaStreamWaitEvent(s, fk.join, 0));
        return TSD_OK;
      }
      if (rc != TSD_ERR_UNSUPPORTED) return rc;
    }
    for (int l = 0; l < num_blocks; ++l) {
      const tsd_interaction_t& b = blocks[l];
      float* filt = fbuf[l % nbuf];
      if (stacked) {
        if (l == 0 || l == stack_cut) TSD_CUDA(cudaStreamWaitEvent(side, fk.edge_done[l], 0));
      } else {
        if (l >= nbuf) TSD_CUDA(cudaStreamWaitEvent(s, fk.agg_done[l - nbuf], 0));  // buffer reuse: that block has read it
        ChainArgs c;
        memset(&c, 0, sizeof(c));
        c.M_cap = batch->edge_capacity;
        c.M_ptr = edges->num_edges;
        c.H = H;
        c.A = edge_attr;
        c.num_stages = 2;
        c.st[0] = chain_stage(b.nn0, TSD_ACT_SSP);
        c.st[1] = chain_stage(b.nn2, TSD_ACT_NONE);
        c.st[1].scale_len = edges->length;
        c.st[1].cutoff = b.cutoff;
        c.st[1].smooth = b.smooth;
        c.st[1].store = filt;
        TSD_TRY(tsd_chain_tf32(c, s));
        TSD_CUDA(cudaEventRecord(fk.edge_done[l], s));
        TSD_CUDA(cudaStreamWaitEvent(side, fk.edge_done[l], 0));
      }
      memset(&na, 0, sizeof(na));
      na.num_nodes = batch->num_nodes;
      na.nodes_per_cluster = npc;
      na.H = H;
      na.in_ptr = edges->in_ptr;
      na.in_eid = edges->in_eid;
      na.in_src = edges->in_src;
      na.x1 = l == 0 ? x1_block0 : x1buf[l & 1];
      na.filt = filt;
      na.st[0].W = b.lin2.weight;
      na.st[0].bias = b.lin2.bias;
      na.st[0].act = TSD_ACT_SSP;
      na.st[1].W = b.lin.weight;
      na.st[1].bias = b.lin.bias;
      na.st[1].residual = l == 0 ? h_in : h_out;
      na.st[1].store = h_out;
      na.num_stages = 2;
      if (l + 1 < num_blocks) {
        na.st[2].W = blocks[l + 1].lin1.weight;
        na.st[2].store = x1buf[(l + 1) & 1];
        na.num_stages = 3;
      }
      TSD_TRY(tsd_node_update_tf32(na, tile, side));
      TSD_CUDA(cudaEventRecord(fk.agg_done[l], side));
    }
    TSD_CUDA(cudaEventRecord(fk.join, side));
    TSD_CUDA(cudaStreamWaitEvent(s, fk.join, 0));
    return TSD_OK;
  }
  bool split = nf_pool && nf_pool_count >= 2 && num_blocks >= 2;
  for (int l = 0; l + 1 < num_blocks && split; ++l) split = blocks[l].fused_w && blocks[l].fused_b;
  const size_t node_elems = (size_t)batch->num_nodes * H;
  float* aggbuf[2] = {nf1, split ? nf2 : nf1};
  float* xh[2] = {nf_pool, split ? nf_pool + node_elems : nullptr};
  TSD_CUDA(cudaEventRecord(fk.fork, s));
  TSD_CUDA(cudaStreamWaitEvent(side, fk.fork, 0));
  TSD_CUDA(cudaStreamWaitEvent(side2, fk.fork, 0));
  GemmArgs g = node_gemm(batch, blocks[0].lin1);  // x1 of block 0
  g.A = h_in;
  g.C = nf0;
  TSD_TRY(tsd_gemm(g, math, side));
  if (split) {
    g = node_gemm(batch, blocks[1].lin1);  // xh_0 = lin1_1(h_0)
    g.A = h_in;
    g.C = xh[0];
    TSD_TRY(tsd_gemm(g, math, side2));
    TSD_CUDA(cudaEventRecord(fk.xh_init, side2));
  }
  const float* h = h_in;
  for (int l = 0; l < num_blocks; ++l) {
    const tsd_interaction_t& b = blocks[l];
    float* filt = (l & 1) ? ef0 : ef1;
    // (measured: letting consecutive blocks' filter kernels run concurrently on two streams with one
    // buffer per block was SLOWER, 33-36 vs 38.7 samples/s: the queued edge tiles starve the serial
    // node-side chain even with a high-priority stream.)
    if (l >= 2) TSD_CUDA(cudaStreamWaitEvent(s, fk.agg_done[l - 2], 0));  // buffer reuse: agg_{l-2} has read it
    ChainArgs c;
    memset(&c, 0, sizeof(c));
    c.M_cap = batch->edge_capacity;
    c.M_ptr = edges->num_edges;
    c.H = H;
    c.A = edge_attr;
    c.num_stages = 2;
    c.st[0] = chain_stage(b.nn0, TSD_ACT_SSP);
    c.st[1] = chain_stage(b.nn2, TSD_ACT_NONE);
    c.st[1].scale_len = edges->length;
    c.st[1].cutoff = b.cutoff;
    c.st[1].smooth = b.smooth;
    c.st[1].store = filt;
    TSD_TRY(tsd_chain_tf32(c, s));
    TSD_CUDA(cudaEventRecord(fk.edge_done[l], s));
    // node side of block l (issued in the same loop iteration so that, under stream capture, every
    // event is recorded in the capture before anything waits on it)
    float* agg = aggbuf[l & 1];
    TSD_CUDA(cudaStreamWaitEvent(side, fk.edge_done[l], 0));
    TSD_TRY(tsd_aggregate(batch, edges, H, nf0, filt, agg, side));
    TSD_CUDA(cudaEventRecord(fk.agg_done[l], side));
    memset(&c, 0, sizeof(c));
    c.M_cap = batch->num_nodes;
    c.H = H;
    c.A = agg;
    c.st[0] = chain_stage(b.lin2, TSD_ACT_SSP);
    if (!split) {
      // node update: h' = h + lin(ssp(lin2(agg))) and, unless this is the last block, x1' = lin1_next(h')
      c.st[1] = chain_stage(b.lin, TSD_ACT_NONE);
      c.st[1].residual = h;
      c.st[1].store = h_out;
      if (l + 1 < num_blocks) {
        c.num_stages = 3;
        c.st[2] = chain_stage(blocks[l + 1].lin1, TSD_ACT_NONE);
        c.st[2].store = nf0;
      } else {
        c.num_stages = 2;
      }
      TSD_TRY(tsd_chain_tf32(c, side));
    } else if (l + 1 < num_blocks) {
      // side2: h_{l+1} = h_l + lin(y) and xh_{l+1} = lin1_{l+2}(h_{l+1}); needs agg_l and (stream order) h_l
      ChainArgs c2 = c;
      c2.st[1] = chain_stage(b.lin, TSD_ACT_NONE);
      c2.st[1].residual = h;
      c2.st[1].store = h_out;
      if (l + 2 < num_blocks) {
        c2.num_stages = 3;
        c2.st[2] = chain_stage(blocks[l + 2].lin1, TSD_ACT_NONE);
        c2.st[2].store = xh[(l + 1) & 1];
      } else {
        c2.num_stages = 2;
      }
      TSD_CUDA(cudaStreamWaitEvent(side2, fk.agg_done[l], 0));
      TSD_TRY(tsd_chain_tf32(c2, side2));
      TSD_CUDA(cudaEventRecord(fk.n2_done[l], side2));
      // side (critical): x1_{l+1} = xh_l + fused_w y + fused_b
      tsd_linear_t fused = b.lin;
      fused.weight = b.fused_w;
      fused.bias = b.fused_b;
      c.num_stages = 2;
      c.st[1] = chain_stage(fused, TSD_ACT_NONE);
      c.st[1].residual = xh[l & 1];
      c.st[1].store = nf0;
      TSD_CUDA(cudaStreamWaitEvent(side, l == 0 ? fk.xh_init : fk.n2_done[l - 1], 0));
      TSD_TRY(tsd_chain_tf32(c, side));
    } else {
      // last block: only h_out is needed; h_{L-1} comes from side2
      c.num_stages = 2;
      c.st[1] = chain_stage(b.lin, TSD_ACT_NONE);
      c.st[1].residual = h;
      c.st[1].store = h_out;
      TSD_CUDA(cudaStreamWaitEvent(side, fk.n2_done[l - 1], 0));
      TSD_TRY(tsd_chain_tf32(c, side));
    }
    h = h_out;
  }
  TSD_CUDA(cudaEventRecord(fk.join2, side2));
  TSD_CUDA(cudaStreamWaitEvent(s, fk.join2, 0));
  TSD_CUDA(cudaEventRecord(fk.join, side));
  TSD_CUDA(cudaStreamWaitEvent(s, fk.join, 0));
  return TSD_OK;
}
