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
int tsd_launch_gine_aggregate(int num_nodes, int H, const int* in_ptr, const int* in_eid, const int* row,
                              const int* local_tab, const float* h, const float* ea, const float* eps, float* out,
                              cudaStream_t s);

#include <atomic>
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

extern "C" const char* tsd_error_string(int code) {
  switch (code) {
    case TSD_OK: return "ok";
    case TSD_ERR_INVALID: return "invalid argument";
    case TSD_ERR_CUDA: return cudaGetErrorString((cudaError_t)g_last_cuda_error);
    case TSD_ERR_UNSUPPORTED: return "unsupported shape or mode";
    default: return "unknown error";
  }
}

int tsd_gemm(const GemmArgs& g, int math, cudaStream_t stream) {
  if (math == TSD_MATH_TF32) {
    int rc = tsd_gemm_tf32(g, stream);
    if (rc != TSD_ERR_UNSUPPORTED) return rc;  // shapes the tensor-core kernel does not take run on FFMA
  }
  return tsd_gemm_ffma(g, stream);
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
    g.C = tmp;
    g.round_out = 1;  // feeds cat2
    TSD_TRY(tsd_gemm(g, math, s));
    GemmArgs g2 = edge_gemm(batch, edges, *enc->cat2);
    g2.A = tmp;
    g2.C = out;
    g2.round_out = 1;  // edge_attr feeds the filter networks and the pair MLP
    TSD_TRY(tsd_gemm(g2, math, s));
  }
  return TSD_OK;
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
  TSD_TRY(tsd_launch_cfconv_aggregate(batch->num_nodes, F, edges->in_ptr, edges->in_eid, edges->in_src, nf0, ef1, nf1, s));
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
  TSD_REQUIRE(batch && edges && h && edge_attr && mlp && ef0 && edge_inv);
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
  g.act = mlp->act;
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
  // TSD_AGG_VARIANT=5 selects the graph-staged kernel (x1 rows in shared memory).  It was measured
  // SLOWER than the node-parallel kernel at batch 100 (30 us vs 18 us warm at E = 33.5k: only
  // G x H/128 CTAs, too little memory-level parallelism), so it is kept for experiments only.
  const char* v = getenv("TSD_AGG_VARIANT");
  if (v && atoi(v) == 5)
    return tsd_launch_cfconv_aggregate_staged(batch, channels, edges->in_ptr, edges->in_eid, edges->in_src, x1, filt, agg,
                                              tsd_cu(stream));
  return tsd_launch_cfconv_aggregate(batch->num_nodes, channels, edges->in_ptr, edges->in_eid, edges->in_src, x1, filt,
                                     agg, tsd_cu(stream));
}

// library-owned side stream + events for the fork/join inside tsd_schnet_encoder
struct EncoderFork {
  static const int MAX_BLOCKS = 32;
  cudaStream_t side = nullptr, edge2 = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr, join2 = nullptr;
  cudaEvent_t edge_done[MAX_BLOCKS], agg_done[MAX_BLOCKS];
  bool ready = false;
  bool two_edge_streams = false;
  int init(int num_blocks) {
    if (num_blocks > MAX_BLOCKS) return TSD_ERR_UNSUPPORTED;
    if (ready) return TSD_OK;
    // the node-side chain is serial and short (15-CTA kernels): give it the highest priority so its
    // CTAs are scheduled as soon as an SM frees up instead of queueing behind the edge tiles
    int prio_lo = 0, prio_hi = 0;
    TSD_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    TSD_CUDA(cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, prio_hi));
    TSD_CUDA(cudaStreamCreateWithFlags(&edge2, cudaStreamNonBlocking));
    TSD_CUDA(cudaEventCreateWithFlags(&join2, cudaEventDisableTiming));
    TSD_CUDA(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
    TSD_CUDA(cudaEventCreateWithFlags(&join, cudaEventDisableTiming));
    for (int i = 0; i < MAX_BLOCKS; ++i) {
      TSD_CUDA(cudaEventCreateWithFlags(&edge_done[i], cudaEventDisableTiming));
      TSD_CUDA(cudaEventCreateWithFlags(&agg_done[i], cudaEventDisableTiming));
    }
    const char* e = getenv("TSD_ENCODER_EDGE_STREAMS");
    two_edge_streams = e && atoi(e) == 2;
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

// Whole SchNet encoder (schnet.py:203-225).  fp32 mode: one tsd_cfconv_layer per block.  tf32
// mode: per block ONE chained filter-network kernel on the edges, the segmented aggregation, and
// ONE chained node kernel that also produces the next block's x1 = lin1(h') -- 3 launches per
// block instead of 6 and no (E,H) / (N,H) intermediate round trips.
extern "C" int tsd_schnet_encoder(const tsd_batch_t* batch, const tsd_edges_t* edges, const float* edge_attr,
                                  const tsd_interaction_t* blocks, int32_t num_blocks, const float* h_in, float* h_out,
                                  float* ef0, float* ef1, float* nf0, float* nf1, float* nf2, float* filt_pool,
                                  int32_t filt_pool_count, int32_t math, tsd_stream_t stream) {
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
  // The filter networks depend only on edge_attr, not on h: block l+1's edge kernel can run while
  // block l's aggregation / node update is in flight.  Fork a side stream for the node-side chain
  // (x1_0, agg_l, node kernel_l); the main stream runs the edge kernels back to back, alternating
  // between two filter buffers.  The edge kernel's second wave leaves most SMs idle (188-263 tiles
  // on 148 SMs), which is where the 15-CTA node kernels and the aggregation execute.  Inside a
  // CUDA-graph capture the event waits become graph edges.
  static EncoderFork fk;
  TSD_TRY(fk.init(num_blocks));
  cudaStream_t side = fk.side, edge2 = fk.edge2;
  TSD_CUDA(cudaEventRecord(fk.fork, s));
  TSD_CUDA(cudaStreamWaitEvent(side, fk.fork, 0));
  TSD_CUDA(cudaStreamWaitEvent(edge2, fk.fork, 0));
  GemmArgs g = node_gemm(batch, blocks[0].lin1);  // x1 of block 0
  g.A = h_in;
  g.C = nf0;
  TSD_TRY(tsd_gemm(g, math, side));
  const float* h = h_in;
  // filter buffers: a caller-provided pool (one per block: no reuse waits) or the two scratch buffers
  const int nbuf = (filt_pool && filt_pool_count >= 2) ? (filt_pool_count < num_blocks ? filt_pool_count : num_blocks) : 2;
  const size_t buf_elems = (size_t)(batch->edge_capacity > 0 ? batch->edge_capacity : 1) * H;
  auto filt_of = [&](int l) -> float* {
    if (filt_pool && filt_pool_count >= 2) return filt_pool + (size_t)(l % nbuf) * buf_elems;
    return (l & 1) ? ef0 : ef1;
  };
  // The edge kernels of consecutive blocks are independent of each other too: they alternate
  // between two streams so block l+1's tiles fill the SMs that block l's second wave leaves idle.
  for (int l = 0; l < num_blocks; ++l) {
    const tsd_interaction_t& b = blocks[l];
    float* filt = filt_of(l);
    // NOTE measured: letting consecutive blocks' edge kernels run concurrently (alternating streams,
    // one filter buffer per block) was SLOWER (33-36 vs 38.7 samples/s): the queued edge tiles
    // starve the serial node-side chain even with a high-priority stream.  Edge kernels therefore
    // stay on the caller's stream; TSD_ENCODER_EDGE_STREAMS=2 re-enables the experiment.
    cudaStream_t es = (fk.two_edge_streams && (l & 1)) ? edge2 : s;
    if (l >= nbuf) TSD_CUDA(cudaStreamWaitEvent(es, fk.agg_done[l - nbuf], 0));  // buffer reuse: agg_{l-nbuf} has read it
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
    TSD_TRY(tsd_chain_tf32(c, es));
    TSD_CUDA(cudaEventRecord(fk.edge_done[l], es));
    // node side of block l (issued in the same loop iteration so that, under stream capture, every
    // event is recorded in the capture before anything waits on it)
    TSD_CUDA(cudaStreamWaitEvent(side, fk.edge_done[l], 0));
    TSD_TRY(tsd_launch_cfconv_aggregate(batch->num_nodes, H, edges->in_ptr, edges->in_eid, edges->in_src, nf0, filt, nf1,
                                        side));
    TSD_CUDA(cudaEventRecord(fk.agg_done[l], side));
    // node update: h' = h + lin(ssp(lin2(agg))) and, unless this is the last block, x1' = lin1_next(h')
    memset(&c, 0, sizeof(c));
    c.M_cap = batch->num_nodes;
    c.H = H;
    c.A = nf1;
    c.st[0] = chain_stage(b.lin2, TSD_ACT_SSP);
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
    h = h_out;
  }
  TSD_CUDA(cudaEventRecord(fk.join2, edge2));
  TSD_CUDA(cudaStreamWaitEvent(s, fk.join2, 0));
  TSD_CUDA(cudaEventRecord(fk.join, side));
  TSD_CUDA(cudaStreamWaitEvent(s, fk.join, 0));
  return TSD_OK;
}
