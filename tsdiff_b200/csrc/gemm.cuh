// Fused "linear layer" GEMM used by every per-edge / per-node MLP of the path:
//   C[m, n] = epilogue( sum_k Aop[m, k] * W[n, k] + bias[n] )
// W is the nn.Linear weight as stored (out x in, row-major = K-major), read in place.
// The A operand is produced on the fly by a prologue so intermediates never hit memory:
//   A_PLAIN     : a dense (M, K) matrix
//   A_EDGE_MLP0 : act0(w0[k] * len[m] + b0[k])                        (edge.py:50-52 layer 0)
//   A_CAT       : d_emb[m, k%H] * bond_emb[code(m, k/H)][k%H], K = 2H  (edge.py:66-68 + cat)
//   A_PAIR      : k < H ? h[row[m], k] * h[col[m], k] : ea[m, k-H]     (common.py:226-229)
// The number of rows is read from device memory (M_ptr) so one captured launch serves any
// per-step edge count up to M_cap.
//
// Code-size note: the kernels are specialised at compile time on the activation, the
// epilogue flavour and (tensor-core kernel) the A kind.  A first version selected them at
// run time inside fully unrolled loops and grew to 10-16k SASS instructions per kernel,
// which made every launch instruction-cache bound (ncu: stall_no_inst).
#pragma once
#include "common.cuh"

enum { TSD_A_PLAIN = 0, TSD_A_EDGE_MLP0 = 1, TSD_A_CAT = 2, TSD_A_PAIR = 3 };
enum { TSD_EPI_PLAIN = 0, TSD_EPI_SCALE = 1, TSD_EPI_MULEMB = 2, TSD_EPI_DOT = 3 };

struct GemmArgs {
  int M_cap;
  const int* M_ptr;  // device row count (NULL: M = M_cap)
  int N, K, H;
  int a_kind;
  const float* A;
  int lda;
  const float* len;  // A_EDGE_MLP0
  const float* w0;
  const float* b0;
  int act0;
  const float* emb;  // A_CAT: (100, H) bond embedding; code = row_lo | row_hi << 16
  const int* code;
  const int* row_index;  // A_CAT, optional: logical row m reads d_emb / code of row row_index[m] (compact row lists)
  const float* alt_A;    // A_PAIR, optional: rows with alt_pos[m] >= 0 take their edge_attr from alt_A[alt_pos[m]]
  const int* alt_pos;
  const float* h;    // A_PAIR
  const int* row;
  const int* col;
  const float* W;
  const float* bias;
  int act;
  const float* scale_len;  // EPI_SCALE: *= C(len[m]) (cutoff envelope)
  float cutoff;
  int smooth;
  const float* mul_emb;    // EPI_MULEMB: *= mul_emb[(mul_code[m] & 0xffff) * N + n]
  const int* mul_code;
  const float* residual;   // EPI_PLAIN only: += residual[m * ldr + n]
  int ldr;
  float* C;
  int ldc;
  const float* w3;         // EPI_DOT: out_vec[m] (+)= sum_n v[m,n] * w3[n] + b3
  const float* b3;
  float* out_vec;
  int accumulate;
  int round_out;           // tf32 mode: round stored outputs to TF32 (RNE) because a tensor-core GEMM consumes them
  // tf32 tensor-core kernel only (tsd_gemm_chain2_tf32): a SECOND linear layer chained on the same 128-row tile,
  //   out = epilogue( act2( act(Aop W^T + bias) W2^T + bias2 ) ),   W2: (N2, N) row-major, act2 = act for the row-dot
  // epilogue and none otherwise.  The first layer's output stays in tensor memory (the A operand of the second MMA);
  // C / ldc / round_out / w3 / b3 / out_vec / accumulate describe the second layer's output.
  const float* W2;
  const float* bias2;
  int N2;
};

static inline GemmArgs tsd_gemm_args() {
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  return g;
}

// which epilogue flavour a call needs; -1 if the combination is not expressible
static inline int tsd_gemm_epi_kind(const GemmArgs& g) {
  const int n = (g.scale_len != nullptr) + (g.mul_emb != nullptr) + (g.out_vec != nullptr);
  if (n > 1 || (n == 1 && g.residual)) return -1;
  if (g.out_vec) return TSD_EPI_DOT;
  if (g.scale_len) return TSD_EPI_SCALE;
  if (g.mul_emb) return TSD_EPI_MULEMB;
  return TSD_EPI_PLAIN;
}

// run-time selected activation on 4 values; deliberately NOT inlined so the prologue of the
// edge-MLP layer 0 (the only place that needs a run-time activation) costs one copy of the code
static __device__ __noinline__ float4 tsd_act4_rt(int act, float4 v) {
  v.x = tsd_act(act, v.x);
  v.y = tsd_act(act, v.y);
  v.z = tsd_act(act, v.z);
  v.w = tsd_act(act, v.w);
  return v;
}

// 4 consecutive k of the A operand of row m (m < M, k % 4 == 0), A kind known at compile time
template <int AKIND>
__device__ __forceinline__ float4 tsd_load_a4_t(const GemmArgs& p, int m, int k) {
  if (AKIND == TSD_A_EDGE_MLP0) {
    const float l = p.len[m];
    const float4 w = __ldg(reinterpret_cast<const float4*>(p.w0 + k));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.b0 + k));
    return tsd_act4_rt(p.act0, make_float4(fmaf(l, w.x, b.x), fmaf(l, w.y, b.y), fmaf(l, w.z, b.z), fmaf(l, w.w, b.w)));
  }
  if (AKIND == TSD_A_CAT) {
    const int src = p.row_index ? p.row_index[m] : m;
    const int code = p.code[src];
    const int hi = k >= p.H;
    const int kk = k - (hi ? p.H : 0);
    const int r = hi ? ((unsigned)code >> 16) : (code & 0xffff);
    const float4 d = *reinterpret_cast<const float4*>(p.A + (size_t)src * p.lda + kk);
    const float4 e = __ldg(reinterpret_cast<const float4*>(p.emb + (size_t)r * p.H + kk));
    return make_float4(d.x * e.x, d.y * e.y, d.z * e.z, d.w * e.w);
  }
  if (AKIND == TSD_A_PAIR) {
    if (k < p.H) {
      const float4 a = *reinterpret_cast<const float4*>(p.h + (size_t)p.row[m] * p.H + k);
      const float4 b = *reinterpret_cast<const float4*>(p.h + (size_t)p.col[m] * p.H + k);
      return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
    }
    const int alt = p.alt_pos ? p.alt_pos[m] : -1;
    if (alt >= 0) return *reinterpret_cast<const float4*>(p.alt_A + (size_t)alt * p.lda + (k - p.H));
    return *reinterpret_cast<const float4*>(p.A + (size_t)m * p.lda + (k - p.H));
  }
  return *reinterpret_cast<const float4*>(p.A + (size_t)m * p.lda + k);
}

// run-time A kind (FFMA kernel: one uniform branch per 16-byte load)
__device__ __forceinline__ float4 tsd_load_a4(const GemmArgs& p, int m, int k) {
  switch (p.a_kind) {
    case TSD_A_EDGE_MLP0: return tsd_load_a4_t<TSD_A_EDGE_MLP0>(p, m, k);
    case TSD_A_CAT: return tsd_load_a4_t<TSD_A_CAT>(p, m, k);
    case TSD_A_PAIR: return tsd_load_a4_t<TSD_A_PAIR>(p, m, k);
    default: return tsd_load_a4_t<TSD_A_PLAIN>(p, m, k);
  }
}

// chained H x H linear layers on one 128-row tile (gemm_chain.cu)
struct ChainStage {
  const float* W;         // (H, H) row-major (TF32-rounded shadow of the nn.Linear weight)
  const float* bias;      // (H) or NULL
  int act;
  const float* scale_len; // (M) or NULL: multiply by the cutoff envelope C(len[m])
  float cutoff;
  int smooth;
  const float* residual;  // (M, H) or NULL: added after the activation
  float* store;           // (M, H) or NULL: result written to global memory
};

struct ChainArgs {
  int M_cap;
  const int* M_ptr;
  int H;
  const float* A;         // (M_cap, H) dense input of stage 0
  int num_stages;         // 2 or 3
  ChainStage st[3];
};

// node side of an interaction block for few rows: fused aggregation + transposed chained linears (node_update.cu)
struct NodeStage {
  const float* W;         // (H, H) row-major, TF32-rounded shadow
  const float* bias;      // (H) or NULL
  int act;                // TSD_ACT_SSP or TSD_ACT_NONE
  const float* residual;  // (N, H) or NULL: added after the activation
  float* store;           // (N, H) or NULL
};

struct NodeArgs {
  int num_nodes, H, num_stages;
  int nodes_per_cluster;  // atoms a cluster owns (<= the kernel's NT: the MMA shape is padded); 0 = NT
  const float* x;  // dense (N, H) input of stage 0, or NULL -> fused CFConv aggregation of the fields below
  const int* in_ptr;
  const int* in_eid;
  const int* in_src;
  const float* x1;    // (N, H)
  const float* filt;  // (rows, H)
  NodeStage st[3];
};

// the node side of ALL interaction blocks in one persistent kernel (node_chain.cu)
constexpr int TSD_NC_MAX_BLOCKS = 8;
struct NodeChainBlock {
  const float* filt;   // (rows, H) filter of this block
  const float* w_lin2; // (H, H) TF32-rounded shadows
  const float* b_lin2;
  const float* w_lin;
  const float* b_lin;
  const float* w_lin1_next;  // the NEXT block's lin1 (no bias); NULL for the last block
};
struct NodeChainArgs {
  int num_nodes, H, num_blocks;
  int nodes_per_cluster;     // atoms a cluster owns (<= 32); 0 = 32
  const int* in_ptr;
  const int* in_eid;
  const int* in_src;
  const float* x1_first;     // (N, H) x1 of block 0
  float* x1buf[2];           // block l > 0 reads x1buf[l & 1]; block l writes x1buf[(l + 1) & 1]
  const float* h_in;         // residual of block 0
  float* h_out;              // written by every block, residual of the next one
  unsigned int* barrier;     // grid barrier counter, zero at launch
  int* error_flag;           // |= 4 if the grid barrier timed out (a CTA was never scheduled)
  NodeChainBlock blk[TSD_NC_MAX_BLOCKS];
};
// TSD_ERR_UNSUPPORTED when the clusters cannot all be resident at once (callers launch one node kernel per block)
int tsd_node_chain_tf32(const NodeChainArgs& a, cudaStream_t stream);

// the CFConv filter networks of ALL interaction blocks on one 128-row tile, layer after layer (filter_stack.cu)
constexpr int TSD_FS_MAX_LAYERS = 8;
struct FilterStackLayer {
  const float* W0;  // nn.0 weight (H, H), TF32-rounded shadow
  const float* b0;  // (H) or NULL
  const float* W2;  // nn.2 weight
  const float* b2;
  float cutoff;
  int smooth;
  float* out;       // (M_cap, H) filter of this block
};
struct FilterStackArgs {
  int M_cap;
  const int* M_ptr;
  int H, num_layers;
  const float* A;    // edge_attr (M_cap, H)
  const float* len;  // (M_cap)
  int max_ctas;      // 0 = one CTA per SM; otherwise at most this many (a launch that shares the GPU with the node chain)
  FilterStackLayer layer[TSD_FS_MAX_LAYERS];
};
int tsd_filter_stack_tf32(const FilterStackArgs& a, cudaStream_t stream);

int tsd_node_tile(bool alone, int num_nodes, int* nodes_per_cluster);  // node_update.cu: kernel shape
int tsd_node_update_tf32(const NodeArgs& a, int tile, cudaStream_t stream);
int tsd_chain_tf32(const ChainArgs& c, cudaStream_t stream);           // gemm_chain.cu
int tsd_gemm(const GemmArgs& g, int math, cudaStream_t stream);        // dispatch (api.cu)
int tsd_gemm_ffma(const GemmArgs& g, cudaStream_t stream);             // gemm_ffma.cu
int tsd_gemm_tf32(const GemmArgs& g, cudaStream_t stream);             // gemm_tc.cu
int tsd_gemm_chain2_tf32(const GemmArgs& g, cudaStream_t stream);      // gemm_tc.cu: two chained layers (GemmArgs.W2)
