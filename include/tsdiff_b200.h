/* tsdiff_b200 C-ABI -- the drop-in boundary for the LD eps-net hot path.
 *
 * Built from tsdiff_b200/csrc/ into tsdiff_b200/libtsdiff_b200.so for sm_100a.
 * The reference (seonghann/tsdiff) has no FFI of its own: its hot path is Python calling
 * third-party CUDA wheels.  Each entry point below replaces the group of reference
 * functions cited next to it (paths relative to the reference root); INTEGRATION.md shows
 * the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless noted;
 *  - nothing here allocates or frees caller memory; scratch comes in through the structs;
 *  - every call is stream-ordered on `stream` (a cudaStream_t), never synchronises, and is
 *    legal inside CUDA-graph stream capture (tsd_bond_order_build excepted: it memsets);
 *  - return value: 0 = TSD_OK, otherwise a TSD_ERR_* code (tsd_error_string explains);
 *  - edge counts live on the device (`num_edges`), kernels are launched at capacity and
 *    exit early, so one captured graph serves every step of a sampling run;
 *  - host threads: entry points keep no per-call state except tsd_schnet_encoder's library-owned
 *    side streams and events (one set per process); its enqueue is serialised internally, so calls
 *    from several threads / streams are safe, their node-side work shares the side streams.
 */
#ifndef TSDIFF_B200_H
#define TSDIFF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* tsd_stream_t; /* cudaStream_t */

enum {
  TSD_OK = 0,
  TSD_ERR_INVALID = 1,   /* bad argument (null pointer, unsupported size) */
  TSD_ERR_CUDA = 2,      /* a CUDA runtime call failed; see tsd_last_cuda_error */
  TSD_ERR_UNSUPPORTED = 3
};

enum { TSD_ACT_NONE = 0, TSD_ACT_RELU = 1, TSD_ACT_SWISH = 2, TSD_ACT_SSP = 3, TSD_ACT_SOFTPLUS = 4 };

/* GEMM arithmetic for the per-edge / per-node linear layers */
enum {
  TSD_MATH_FP32 = 0, /* FFMA, fp32 in / fp32 accumulate: the 1e-4 parity path            */
  TSD_MATH_TF32 = 1  /* tcgen05.mma kind::tf32, fp32 accumulate in TMEM: looser bound    */
};

#define TSD_MAX_GRAPH_NODES 256 /* atoms per reaction graph supported by the tile builders */
#define TSD_NUM_BOND_TYPES 22   /* len(utils.chem.BOND_TYPES), utils/chem.py:21 */

/* nn.Linear: weight (out_features, in_features) row-major, bias (out_features) or NULL.
 * Pointers go straight into the live nn.Parameter storage. */
typedef struct {
  const float* weight;
  const float* bias;
  int32_t in_features;
  int32_t out_features;
} tsd_linear_t;

/* Static description of a batch of reaction graphs (built once per batch). */
typedef struct {
  int32_t num_nodes;        /* N */
  int32_t num_graphs;       /* G */
  int32_t max_graph_nodes;  /* max_g n_g (<= TSD_MAX_GRAPH_NODES) */
  int32_t edge_capacity;    /* sum_g n_g (n_g - 1) */
  const int32_t* graph_ptr; /* (G+1) first node of each graph; `batch` must be sorted */
  const int32_t* pair_ptr;  /* (G+1) prefix of n_g^2: offset of graph g's n_g x n_g pair tables */
  const int32_t* node_graph;/* (N) graph id of each node (= batch) */
} tsd_batch_t;

/* Per-step edge list in CSR form, written by tsd_edge_build. All arrays have
 * `edge_capacity` entries; the first num_edges[0] are valid, sorted by (row, col). */
typedef struct {
  int32_t* num_edges;  /* (2) [0] = E of graph a, [1] = E of graph b (edges with in_b) */
  int32_t* row;        /* edge_index[0] */
  int32_t* col;        /* edge_index[1] */
  float* length;       /* ||pos[row] - pos[col]|| */
  int32_t* tab0;       /* pair table 0 value of the edge (0 for radius-only edges) */
  int32_t* tab1;       /* pair table 1 value */
  int32_t* in_b;       /* 1 if the edge also belongs to graph b */
  int32_t* row_ptr;    /* (N+1) out-CSR over `row` */
  int32_t* in_ptr;     /* (N+1) in-CSR: edges grouped by `col` (dst-sorted) */
  int32_t* in_eid;     /* (E) edge ids sorted by (col, row) */
  int32_t* in_src;     /* (E) source node of in_eid[k] (= row[in_eid[k]]), saves a dependent load */
  int32_t* graph_count;/* (G) scratch: edges per graph */
  /* Undirected pair list (optional; num_upairs == NULL skips it).  The directed edges (i,j) and
   * (j,i) have bit-identical length and -- for a symmetric bond list -- table values, so everything
   * the networks compute per edge (edge embedding, CFConv filters, the pair MLP) is identical for
   * both directions.  K2 also emits each unordered pair {i<j} that has at least one directed edge
   * once, sorted by (i, j), plus the pair id of every directed edge / in-CSR slot; the per-edge
   * kernels then run over U ~ E/2 rows (pass a view whose num_edges, row, col, length, tab0, tab1, in_eid
   * are num_upairs, u_row, u_col, u_length, u_tab0, u_tab1, in_upair) and consumers index results
   * through edge_upair. */
  int32_t* num_upairs; /* (1) U */
  int32_t* u_row;      /* (U_cap) i, the smaller node; U_cap = sum_g n_g (n_g - 1) / 2 */
  int32_t* u_col;      /* (U_cap) j */
  float* u_length;     /* (U_cap) */
  int32_t* u_tab0;     /* (U_cap) table 0 value at (i, j) */
  int32_t* u_tab1;     /* (U_cap) */
  int32_t* edge_upair; /* (E) pair id of directed edge e */
  int32_t* in_upair;   /* (E) pair id of in-CSR slot k (= edge_upair[in_eid[k]]) */
  int32_t* graph_ucount;/* (G) scratch: pairs per graph */
} tsd_edges_t;

const char* tsd_error_string(int code);
int tsd_last_cuda_error(void);  /* cudaError_t of the last TSD_ERR_CUDA on this thread */
int tsd_version(void);
/* number of kernel launches this library has issued (host counter; bench.py's gpu_launches) */
int64_t tsd_launch_count(void);

/* Scratch a caller has to provide for one eps-net evaluation (the library never allocates caller
 * memory).  network: 0 = condensenc (path B), 1 = dualenc (path A).
 * Outputs (any may be NULL): bytes of ONE (E_cap, hidden) edge buffer and of ONE (N, hidden)
 * node buffer, and how many of each the kernel sequence of the network uses, including the two
 * pooled node buffers of the tf32 encoder (tsd_schnet_encoder's nf_pool). */
int tsd_workspace_bytes(int32_t num_nodes, int32_t edge_capacity, int32_t hidden, int32_t network, int32_t math,
                        uint64_t* edge_buffer_bytes, uint64_t* node_buffer_bytes, int32_t* edge_buffers,
                        int32_t* node_buffers);

/* ---- K1: bond-order (k-hop) pair tables; position independent, once per batch -------------
 * mode 0 (path B, replaces models/common.py:115-202 `_extend_ts_graph_order`): reactant
 *   (type / 22) and product (type % 22) bond graphs are extended separately; table value =
 *   type_r | type_p << 16 with bonds keeping their type and k-hop pairs (2 <= k <= order)
 *   getting 22 + k - 1; 0 = not a local pair.  table_a uses order_a, table_b order_b.
 * mode 1 (path A, replaces models/common.py:255-325 `_extend_graph_order`): one graph;
 *   table_a = raw extended type (bond type, or 484 + k - 1); table_b = bond-embedding rows
 *   row1 | row2 << 16 decoded as models/epsnet/dualenc.py:270-293 (ts_decode selects the TS
 *   branch).  order_b is ignored.
 * scratch: 2 * pair_ptr[G] int32.  Duplicate bond entries are summed like to_dense_adj. */
int tsd_bond_order_build(int mode, const tsd_batch_t* batch, int32_t num_bonds,
                         const int64_t* bond_index /* (2, num_bonds) */,
                         const int64_t* bond_type /* (num_bonds) */,
                         int32_t order_a, int32_t order_b, int32_t ts_decode,
                         int32_t* table_a, int32_t* table_b, int32_t* scratch,
                         int32_t* error_flag /* (1) device, set !=0 on cross-graph bonds */,
                         tsd_stream_t stream);

/* ---- K2: per-step edge list (replaces models/common.py:328-384 `_extend_to_radius_graph`,
 * :205-223, :387-417 and models/epsnet/condensenc.py:117-154): radius graph per reaction
 * (d^2 < cutoff^2 strict, first max_neighbors+1 in-range atoms per centre in index order,
 * self included then dropped -- the torch_cluster CUDA rule) united with the local pairs of
 * table 0, emitted row-major sorted with out- and in-CSR.  If tab1_is_graph, `in_b` marks
 * edges of the second graph (local pairs of table 1 united with the same radius graph). */
int tsd_edge_build(const tsd_batch_t* batch, const float* pos /* (N,3) */, double cutoff,
                   int32_t max_neighbors, const int32_t* table0, const int32_t* table1,
                   int32_t tab1_is_graph, const tsd_edges_t* edges, tsd_stream_t stream);

/* ---- node embedding of path B (models/epsnet/condensenc.py:193-198):
 * z = cat[emb[Z] + W r, W p - W r]  -> (N, 2*half) */
int tsd_condensed_node_embed(int32_t num_nodes, const int64_t* atom_type, const int64_t* r_feat,
                             const int64_t* p_feat, int32_t feat_dim, const float* atom_emb /* (100, half) */,
                             const float* feat_weight /* (half, feat_dim) */, int32_t half, float* z,
                             tsd_stream_t stream);

/* nn.Embedding lookup (+ optional max_norm renormalisation IN PLACE on the looked-up rows,
 * models/encoder/schnet.py:152; scale = max_norm / (norm + 1e-7)). max_norm <= 0 disables. */
int tsd_embedding(int32_t num_nodes, const int64_t* index, float* weight /* (num_rows, dim) */,
                  int32_t num_rows, int32_t dim, float max_norm, float* out /* (N, dim) */,
                  tsd_stream_t stream);

/* ---- K3: edge embedding (replaces models/encoder/edge.py:58-68 `MLPEdgeEncoder.forward`,
 * and with `cat0` != NULL models/epsnet/condensenc.py:156-176 / dualenc.py:270-285):
 *   d_emb = lin1(act(lin0(len)));  without cat: out = d_emb * bond_emb[code & 0xffff]
 *   with cat: out = cat2(cat_act(cat0(cat[d_emb*bond_emb[code&0xffff], d_emb*bond_emb[code>>16]])))
 * `d_emb` (E_cap, H) is scratch that is also an output: pass reuse_d_emb = 1 to skip
 * recomputing it when only the codes changed (second graph of path B). */
typedef struct {
  tsd_linear_t lin0, lin1;       /* edge_encoder.mlp.layers.{0,1} */
  const float* bond_emb;         /* edge_encoder.bond_emb.weight (100, H) */
  int32_t act;                   /* config.mlp_act */
  const tsd_linear_t* cat0;      /* edge_cat.0 (2H -> H) or NULL */
  const tsd_linear_t* cat2;      /* edge_cat.2 (H -> H) */
  int32_t cat_act;               /* config.edge_cat_act */
} tsd_edge_encoder_t;

int tsd_edge_embed(const tsd_batch_t* batch, const tsd_edges_t* edges, const int32_t* code,
                   const tsd_edge_encoder_t* enc, int32_t reuse_d_emb, float* d_emb, float* tmp,
                   float* out, int32_t math, tsd_stream_t stream);

/* ---- K4: one SchNet InteractionBlock (replaces models/encoder/schnet.py:90-128):
 *   W = (nn2(ssp(nn0(edge_attr)))) * C(len);  x1 = lin1(h_in);  agg_i = sum_{j->i} x1_j * W_ji
 *   h_out = h_in + lin(ssp(lin2(agg)))
 * C = [len <= cutoff] or the cosine cutoff when smooth.  Deterministic segmented reduction
 * over the dst-sorted in-CSR (no atomics). scratch: ef (E_cap,H) x2, nf (N,H) x3. */
typedef struct {
  tsd_linear_t nn0, nn2, lin1, lin2, lin;
  float cutoff;
  int32_t smooth;
  /* optional (tf32 encoder only, NULL otherwise): the NEXT block's lin1 folded into this block's
   * lin -- fused_w (H,H) = lin1_next.W @ lin.W, fused_b (H) = lin1_next.W @ lin.b -- so that
   * x1_next = lin1_next(h + lin(y)) = lin1_next(h) + fused_w y + fused_b needs one GEMM, not two,
   * after the aggregation (see tsd_schnet_encoder). */
  const float* fused_w;
  const float* fused_b;
} tsd_interaction_t;

int tsd_cfconv_layer(const tsd_batch_t* batch, const tsd_edges_t* edges, const float* edge_attr,
                     const tsd_interaction_t* blk, const float* h_in, float* h_out, float* ef0,
                     float* ef1, float* nf0, float* nf1, float* nf2, int32_t math,
                     tsd_stream_t stream);

/* The filter network of one CFConv on its own (models/encoder/schnet.py:91-98):
 * filt = nn2(ssp(nn0(edge_attr))) * C(len).  tmp is (E_cap, H) scratch (fp32 mode only). */
int tsd_filter_network(const tsd_batch_t* batch, const tsd_edges_t* edges, const float* edge_attr,
                       const tsd_interaction_t* blk, float* tmp, float* filt, int32_t math, tsd_stream_t stream);

/* The filter networks of `num_blocks` consecutive interaction blocks in ONE tensor-core kernel (tf32 only;
 * models/encoder/schnet.py:91-98 for every block of :203-225): filt[l] = nn2_l(ssp(nn0_l(edge_attr))) * C_l(len).
 * The filters depend on edge_attr only, so a CTA keeps a 128-row tile and runs the 2 x num_blocks chained GEMMs with
 * the TMA weight stream, the tcgen05 main loops and both epilogues overlapped (what tsd_schnet_encoder launches when it
 * has one filter buffer per block).  `filt`: num_blocks device buffers (E_cap, H).  num_blocks <= 8, H in {128, 256},
 * E_cap >= 1024; otherwise TSD_ERR_UNSUPPORTED (use tsd_filter_network per block). */
int tsd_filter_stack(const tsd_batch_t* batch, const tsd_edges_t* edges, const float* edge_attr,
                     const tsd_interaction_t* blocks, int32_t num_blocks, float* const* filt, tsd_stream_t stream);

/* The whole SchNet encoder (models/encoder/schnet.py:203-225): `num_blocks` interaction blocks
 * applied in sequence, h_out = SchNet(h_in).  Same scratch as tsd_cfconv_layer.  In tf32 mode the
 * blocks run as chained tensor-core kernels (filter network fused on the edges; lin2 -> lin ->
 * next block's lin1 fused on the nodes) and the edge kernels of different blocks overlap with each
 * other and with the node side on library-owned side streams (graph branches under capture).
 * nf_pool (optional): nf_pool_count x (N, H) extra node buffers.  With >= 2 of them and
 * fused_w / fused_b set on every block but the last, the node update is split: the critical
 * kernel of a block computes only x1_next = lin1_next(h) + fused_w ssp(lin2(agg)) + fused_b (two
 * chained GEMMs), while h' = h + lin(ssp(lin2(agg))) and lin1 of the block after next run beside
 * the next aggregation.  h_in is not modified; h_out may not alias h_in.
 * ef_pool (optional): ef_pool_count x (E_cap, H) extra filter buffers.  The filter network of a block only depends
 * on edge_attr, so with one buffer per block (2 + ef_pool_count >= num_blocks) the edge-side kernels of ALL blocks
 * run ahead of the serial node-side chain instead of waiting for the block that last used their buffer.
 * x1_first (optional, tf32 fast path only): an (N, H) buffer for the first block's x1 = lin1_0(h_in).  With
 * x1_first_valid = 0 the encoder computes it into the buffer, with 1 it trusts the buffer: a caller whose h_in does not
 * change between calls (the Langevin loop: the node embedding is position independent, the reference recomputes
 * lin1(h) every step, schnet.py:101) takes that kernel out of every step. */
int tsd_schnet_encoder(const tsd_batch_t* batch, const tsd_edges_t* edges, const float* edge_attr,
                       const tsd_interaction_t* blocks, int32_t num_blocks, const float* h_in, float* h_out,
                       float* ef0, float* ef1, float* nf0, float* nf1, float* nf2, float* nf_pool,
                       int32_t nf_pool_count, float* ef_pool, int32_t ef_pool_count, float* x1_first,
                       int32_t x1_first_valid, int32_t math, tsd_stream_t stream);

/* The node side of one InteractionBlock as ONE tensor-core kernel (tf32; what tsd_schnet_encoder launches per block when
 * the batch has few atoms -- schnet.py:101-104,124-128 and the next block's :101):
 *   agg_i = sum_{j->i} x1_j * filt_ji;  h_out = h_in + lin(ssp(lin2(agg)));  x1_next = next_lin1(h_out)
 * The segmented aggregation (in-CSR, ascending source order, no atomics) is the B-operand producer of transposed
 * ("swap AB") tcgen05 GEMMs: M = 128 output features, N = 32 atoms per CTA, weights streamed by TMA.
 * next_lin1 / x1_next may both be NULL (last block).  x1_next must not alias x1; h_out may alias h_in. */
int tsd_interaction_node_update(const tsd_batch_t* batch, const tsd_edges_t* edges, const tsd_interaction_t* blk,
                                const tsd_linear_t* next_lin1, const float* x1, const float* filt, const float* h_in,
                                float* h_out, float* x1_next, tsd_stream_t stream);

/* The two building blocks of K4, exposed on their own for unit tests and for the per-kernel
 * roofline timing in bench.py:
 *  tsd_linear          : out = act(x W^T + b) for a dense (rows, in) x; `rows_dev` (device int,
 *                        may be NULL) overrides `rows` like the per-step edge count does.
 *  tsd_cfconv_aggregate: agg_i = sum_{j->i} x1_j * filt_ji over the dst-sorted in-CSR. */
int tsd_linear(int32_t rows, const int32_t* rows_dev, const float* x, const tsd_linear_t* lin, int32_t act,
               float* out, int32_t math, tsd_stream_t stream);
int tsd_cfconv_aggregate(const tsd_batch_t* batch, const tsd_edges_t* edges, int32_t channels,
                         const float* x1, const float* filt, float* agg, tsd_stream_t stream);

/* fp32 -> TF32 round-to-nearest (result kept in an fp32 container).  The tensor-core path reads
 * fp32 operands by truncation; the host keeps RNE-rounded shadow copies of the weights. */
int tsd_round_tf32(const float* src, float* dst, int64_t n, tsd_stream_t stream);

/* ---- K5: one GINEConv + GINEncoder glue (replaces models/encoder/gin.py:42-73,:136-143):
 *   out_i = sum_{j->i, edge local} relu(h_j + e_ji) + (1 + eps) h_i;  hid = nn1(relu(nn0(out)))
 *   h_out = (relu_after ? relu(hid) : hid) + h_in
 * Only edges with edges->tab0 != 0 (local edges) take part. */
typedef struct {
  tsd_linear_t nn0, nn1;
  const float* eps; /* (1) buffer */
  int32_t relu_after;
} tsd_gine_t;

int tsd_gine_layer(const tsd_batch_t* batch, const tsd_edges_t* edges, const float* edge_attr,
                   const tsd_gine_t* conv, const float* h_in, float* h_out, float* nf0, float* nf1,
                   int32_t math, tsd_stream_t stream);

/* ---- K6: pair features + output MLP (replaces models/common.py:226-229 and :78-90 as
 * grad_dist_mlp / grad_{global,local}_dist_mlp):
 *   edge_inv = l2(act(l1(act(l0(cat[h_row * h_col, edge_attr])))))
 * accumulate != 0 adds into edge_inv (ensemble sum, models/sampler.py:96-109). */
typedef struct {
  tsd_linear_t l0, l1, l2;
  int32_t act;
} tsd_pair_mlp_t;

int tsd_pair_mlp(const tsd_batch_t* batch, const tsd_edges_t* edges, const float* h,
                 const float* edge_attr, const tsd_pair_mlp_t* mlp, int32_t accumulate, float* ef0,
                 float* edge_inv, int32_t math, tsd_stream_t stream);

/* The second graph of path B costs a fraction of the first: its edge embedding differs only on the rows whose packed
 * type codes differ (condensenc.py:219-234; the 4-hop pairs for edge_order 4 / pred_edge_order 3).
 * tsd_edge_embed_delta compacts those rows (diff_rows ascending, diff_pos[m] = position or -1, *diff_count on the
 * device), runs edge_cat on them only (d_emb from the first call, codes code1) into out_compact; tsd_pair_mlp_delta
 * is tsd_pair_mlp whose edge_attr of row m is alt_attr[alt_pos[m]] where alt_pos[m] >= 0.  Bit-identical per row to
 * the full evaluation.  All index arrays have edge_capacity entries. */
int tsd_edge_embed_delta(const tsd_batch_t* batch, const tsd_edges_t* edges, const int32_t* code0, const int32_t* code1,
                         const tsd_edge_encoder_t* enc, const float* d_emb, float* tmp, float* out_compact,
                         int32_t* diff_rows, int32_t* diff_pos, int32_t* diff_count, int32_t math, tsd_stream_t stream);
int tsd_pair_mlp_delta(const tsd_batch_t* batch, const tsd_edges_t* edges, const float* h, const float* edge_attr,
                       const float* alt_attr, const int32_t* alt_pos, const tsd_pair_mlp_t* mlp, int32_t accumulate,
                       float* ef0, float* edge_inv, int32_t math, tsd_stream_t stream);

/* ---- K7: eq_transform + clip + position update + centring + NaN flag, one launch
 * (replaces models/geometry.py:22-30, models/sampler.py:208-254 -- the `ld` branch :238-244 and the
 * `ddpm` branch :215-236 --, :260-268 and models/epsnet/dualenc.py:827-849,946-965).
 * Per step k (read from *step_counter, which the kernel post-increments):
 *   score_c = clip_c(eq_transform(inv_c / inv_div, edges selected by mask_c))  for channel c
 *   eps = score_0 + (use1[k] ? w1 * score_1 : 0)
 *   LD:   pos = center(pos + step_size[k] * eps / sigma[k] + noise * noise_scale[k])
 *   DDPM: pos_c = c0 pos; pos0 = c1 pos_c - c2 (-eps); mean = (c3 pos0 + c4 pos_c) / c5;
 *         pos = center((mean + c6 noise) / c7)     (channel 0 only; every product / sum rounded)
 *   DDPM_DUALENC: pos0 = c0 pos - c1 (-eps); pos = center((c2 pos0 + c3 pos) / c4 + c5 noise)
 *   GENERALIZED:  pos = center(pos - (-eps) c0 + noise c1)
 *   DSM:  the edge scores are multiplied by 1 / sigma[k] first (dualenc.py:305-309);
 *         pos = center(pos + step_size[k] * eps + noise * noise_scale[k])      (sched columns as for LD)
 * noise: external tensor (n_steps, N, 3) if given, else Philox4x32-10 keyed by
 * (seed, step, atom_offset + atom) + Box-Muller.  mask mode: 0 all edges, 1 tab != 0, 2 tab == 0. */
#define TSD_RULE_LD 0
#define TSD_RULE_DDPM 1          /* EnsembleSampler `ddpm`, sampler.py:215-236 */
#define TSD_RULE_DDPM_DUALENC 2  /* dualenc `ddpm_noisy` / `ddpm_det`, dualenc.py:906-944 */
#define TSD_RULE_GENERALIZED 3   /* dualenc `generalized`, dualenc.py:872-904 */
#define TSD_RULE_DSM 4           /* dualenc model type `dsm`, annealed Langevin dynamics, dualenc.py:1102-1203 */

typedef struct {
  const float* inv;     /* (E) -- or (U) when inv_index is given -- or NULL to disable the channel */
  const int32_t* mask;  /* (E) */
  int32_t mask_mode;
  float clip;           /* <= 0: no clipping */
  float weight;
  const int32_t* inv_index; /* (E) or NULL: score of edge e = inv[inv_index[e]] (edges->edge_upair) */
} tsd_score_channel_t;

/* One-shot exchange of the per-atom partial scores between the ranks of an ensemble whose members live on
 * different GPUs (BASELINE config 3; the reference averages edge_inv over the members every step,
 * models/sampler.py:96-111, and eq_transform is linear in edge_inv).  Fused into tsd_ld_step: the CTA of reaction g
 * computes this rank's partial scores, STORES them into every peer's buffer over NVLink (peer-mapped memory),
 * publishes a per-reaction flag (release, system scope), spins on its own flags until every rank's contribution
 * of this step has arrived (acquire), sums the contributions in RANK ORDER (identical on all ranks, so positions
 * stay bit-equal without a broadcast) and goes on with the update.  Double buffered by step parity; flag values are
 * *epoch_base + step + 1, so buffers are never reset while a peer may still write them. */
#define TSD_MAX_EXCHANGE_RANKS 8
typedef struct {
  int32_t world, rank;
  int32_t num_graphs;                            /* G */
  float* peer_data[TSD_MAX_EXCHANGE_RANKS];      /* rank p's buffer (2, world, N, 3); peer_data[rank] is local */
  int32_t* peer_flags[TSD_MAX_EXCHANGE_RANKS];   /* rank p's flags (2, world, G), zero-initialised once */
  const int32_t* epoch_base;                     /* (1) device: bumped by the caller before every new trajectory */
} tsd_exchange_t;

typedef struct {
  const float* sched;      /* rule LD:   (num_steps, 4): step_size, sigma, noise_scale, use_channel1
                            * rule DDPM: (num_steps, 8): sqrt(at), sqrt(1/at), sqrt(1/at - 1), sqrt(atm1) beta_t,
                            *   sqrt(1 - beta_t) (1 - atm1), 1 - at, mask exp(0.5 log beta_t), sqrt(atm1)
                            *   (sampler.py:216-236; the caller evaluates them with the reference's fp32 ops)
                            * rule DDPM_DUALENC: (num_steps, 8): sqrt(1/at), sqrt(1/at - 1), sqrt(atm1) beta_t,
                            *   sqrt(1 - beta_t) (1 - atm1), 1 - at, mask exp(0.5 logvar), use_channel1, 0
                            * rule GENERALIZED: (num_steps, 4): step_size_pos, step_size_noise, 0, use_channel1 */
  int32_t num_steps;       /* rows of sched / noise; steps beyond it are no-ops */
  int32_t* step_counter;   /* (1) device */
  int32_t* ticket;         /* (1) device scratch, zero-initialised once */
  int32_t* nan_flag;       /* (1) device, sticky */
  const float* noise;      /* (n_steps, N, 3) or NULL -> Philox */
  uint64_t seed;
  int64_t atom_offset;     /* global index of this shard's first atom */
  float inv_div;           /* ensemble size M (edge_inv /= M, sampler.py:111) */
  float clip_pos;          /* <= 0: none */
  float* traj;             /* (traj_steps, N, 3) or NULL */
  int32_t traj_steps;      /* rows of traj */
  int32_t traj_base_step;  /* traj slot = step - traj_base_step (skipped when out of range) */
  int32_t rule;            /* TSD_RULE_* */
  const float* node_score; /* (N,3) or NULL.  When set, the per-atom score eq_transform(edge_inv / inv_div) is NOT
                            * recomputed from channel 0 but read from here (then clipped with ch0->clip): the
                            * ensemble-member-per-GPU mode, where every rank runs tsd_eq_transform on its own
                            * members' edge_inv and the (N,3) partial scores are summed across ranks first
                            * (eq_transform is linear in edge_inv, sampler.py:96-111,208-209). */
  const tsd_exchange_t* exchange; /* HOST pointer or NULL.  When set, channel 0 holds THIS rank's members' sum and
                            * inv_div the TOTAL ensemble size; the partial scores are exchanged inside the kernel (see
                            * tsd_exchange_t).  nan_flag bit 1 is set if a peer's contribution did not arrive within ~5 s. */
} tsd_ld_params_t;

int tsd_ld_step(const tsd_batch_t* batch, const tsd_edges_t* edges, float* pos,
                const tsd_score_channel_t* ch0, const tsd_score_channel_t* ch1,
                const tsd_ld_params_t* ld, tsd_stream_t stream);

/* eq_transform alone (models/geometry.py:22-30), used by the API-level tests. */
int tsd_eq_transform(const tsd_batch_t* batch, const tsd_edges_t* edges, const float* pos,
                     const tsd_score_channel_t* ch, float inv_div, float* node_eq /* (N,3) */,
                     tsd_stream_t stream);

/* Philox4x32-10 + Box-Muller normals exactly as tsd_ld_step draws them: out (N,3). */
int tsd_philox_normal(int32_t num_nodes, uint64_t seed, int32_t step, int64_t atom_offset,
                      float* out, tsd_stream_t stream);

/* ---- training step (BASELINE config 4): the backward kernels behind
 * `get_loss(...).mean().backward()` (models/epsnet/condensenc.py:267-328, train.py:124-152).  The training forward runs
 * the operators above UNFUSED (pre-activations kept); tsdiff_b200/training.py wraps forward + backward of every operator
 * in a torch.autograd.Function.  fp32, deterministic reductions (no float atomics).
 *  tsd_act_forward / _backward       y = act(x);  dx = dy * act'(x)   (utils/activation_functions.py, schnet.py:65-71)
 *  tsd_row_scale, tsd_cutoff_envelope  out[m,:] = x[m,:] * s[m], s = C(len) (schnet.py:91-98); own backward with dy
 *  tsd_gate_rows                      out[m,:] = a[m,:] * table[(code[m] >> shift) & 0xffff, :] (edge.py:66-68);
 *                                     a == NULL gathers the rows; backward w.r.t. a = the same op on dy
 *  tsd_onehot                         (rows, classes) one-hot of the codes: d table = wgrad(onehot, dy * a)
 *  tsd_transpose                      W (out,in) -> W^T, so the data gradient dx = dy W is tsd_linear with weight W^T
 *  tsd_linear_wgrad                   dW (N,K) = dy^T x, db (N) = column sums of dy; dy (M,N), x (M,K); scratch floats
 *                                     from tsd_linear_wgrad_scratch
 *  tsd_cfconv_aggregate_backward      dx1, dfilt of agg_i = sum_{j->i} x1_j * filt_ji (schnet.py:102-107)
 *  tsd_pair_features / _backward      cat[h_row * h_col, ea] (E, 2H) (common.py:226-229); dh (N, H)
 *  tsd_eq_transform_backward          d inv of geometry.py:22-30 (mask as in tsd_score_channel_t)
 *  tsd_condensed_node_embed_backward  d atom_embedding.weight (num_types, half), d atom_feat_embedding.weight (half, F)
 *  tsd_sqerr_forward / _backward      loss[n] = sum_d (a - b)^2 (condensenc.py:324-326)
 *  tsd_add, tsd_mul                   out = a + b (residual), out = a * b */
int tsd_act_forward(int64_t n, const float* x, int32_t act, float* y, tsd_stream_t stream);
int tsd_act_backward(int64_t n, const float* x, const float* dy, int32_t act, float* dx, tsd_stream_t stream);
int tsd_row_scale(int32_t rows, int32_t H, const float* x, const float* s, float* out, tsd_stream_t stream);
int tsd_cutoff_envelope(int32_t rows, const float* len, float cutoff, int32_t smooth, float* s, tsd_stream_t stream);
int tsd_gate_rows(int32_t rows, int32_t H, const float* a, const float* table, const int32_t* code, int32_t shift,
                  float* out, tsd_stream_t stream);
int tsd_onehot(int32_t rows, int32_t classes, const int32_t* code, int32_t shift, float* out, tsd_stream_t stream);
int tsd_transpose(int32_t rows, int32_t cols, const float* src, float* dst, tsd_stream_t stream);
int tsd_linear_wgrad_scratch(int32_t M, int32_t N, int32_t K, uint64_t* floats);
int tsd_linear_wgrad(int32_t M, int32_t N, int32_t K, const float* dy, const float* x, float* dW, float* db,
                     float* scratch, tsd_stream_t stream);
int tsd_cfconv_aggregate_backward(const tsd_batch_t* batch, const tsd_edges_t* edges, int32_t num_edges, int32_t H,
                                  const float* x1, const float* filt, const float* dagg, float* dx1, float* dfilt,
                                  tsd_stream_t stream);
int tsd_pair_features(const tsd_edges_t* edges, int32_t num_edges, int32_t H, const float* h, const float* ea, float* out,
                      tsd_stream_t stream);
int tsd_pair_features_backward(const tsd_batch_t* batch, const tsd_edges_t* edges, int32_t H, const float* h,
                               const float* dout, float* dh, tsd_stream_t stream);
int tsd_eq_transform_backward(const tsd_edges_t* edges, int32_t num_edges, const float* pos, const int32_t* mask,
                              int32_t mask_mode, float inv_div, const float* dnode, float* dinv, tsd_stream_t stream);
int tsd_condensed_node_embed_backward(int32_t num_nodes, const int64_t* atom_type, const int64_t* r_feat,
                                      const int64_t* p_feat, int32_t feat_dim, int32_t half, int32_t num_types,
                                      const float* dz, float* d_atom_emb, float* d_feat_weight, tsd_stream_t stream);
int tsd_sqerr_forward(int32_t n, const float* a, const float* b, float* loss, tsd_stream_t stream);
int tsd_sqerr_backward(int32_t n, const float* a, const float* b, const float* dloss, float* da, tsd_stream_t stream);
int tsd_add(int64_t n, const float* a, const float* b, float* out, tsd_stream_t stream);
int tsd_mul(int64_t n, const float* a, const float* b, float* out, tsd_stream_t stream);

/* Peer-mapped device memory for tsd_exchange_t: cudaMalloc + zero fill + cudaIpcGetMemHandle on the owner,
 * cudaIpcOpenMemHandle on the peers (the 64-byte handle travels through the caller's process group). */
int tsd_peer_alloc(uint64_t bytes, void** ptr, unsigned char* handle64);
int tsd_peer_open(const unsigned char* handle64, void** ptr);
int tsd_peer_close(void* ptr);
int tsd_peer_free(void* ptr);

/* ---- post-sampling geometry metrics, fp64 like the reference's numpy / scipy code.
 * tsd_dmae (clustering.py:98-105 `calc_DMAE`): out[b] = sum over i < j of |dm_ref - dm_guess[b]| (mape: / dm_ref),
 *   divided by n (n - 1) / 2; dm_ref (n, n), dm_guess (batch, n, n).  tsd_dmae_pos: the same from positions
 *   pos_ref (n, 3), pos_guess (batch, n, 3).
 * tsd_min_match (clustering.py:123-135 `get_minimum_matches` with its default metric): for every probe geometry b
 *   out_val[b] = min over the permutations m of sum_{i<j} (|ref_i - ref_j| - |prb[b][match[m][i]] - prb[b][match[m][j]]|)^2,
 *   out_idx[b] = the first arg-min (list.index(min(...))).  matches (num_matches, n) int32.  Scratch sizes come
 *   from tsd_min_match_scratch (elements, not bytes). */
int tsd_dmae(int32_t num_atoms, int32_t batch, const double* dm_ref, const double* dm_guess, int32_t mape, double* out,
             tsd_stream_t stream);
int tsd_dmae_pos(int32_t num_atoms, int32_t batch, const double* pos_ref, const double* pos_guess, int32_t mape,
                 double* out, tsd_stream_t stream);
int tsd_min_match_scratch(int32_t batch, int32_t num_matches, int64_t* doubles, int64_t* ints);
int tsd_min_match(int32_t num_atoms, int32_t batch, int32_t num_matches, const double* pos_ref, const double* pos_prb,
                  const int32_t* matches, double* scratch_val, int32_t* scratch_idx, double* out_val, int32_t* out_idx,
                  tsd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TSDIFF_B200_H */
