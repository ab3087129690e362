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
#pragma once
#include "common.cuh"

enum { TSD_A_PLAIN = 0, TSD_A_EDGE_MLP0 = 1, TSD_A_CAT = 2, TSD_A_PAIR = 3 };

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
  const float* h;    // A_PAIR
  const int* row;
  const int* col;
  const float* W;
  const float* bias;
  int act;
  const float* scale_len;  // epilogue: *= C(len[m]) (cutoff envelope)
  float cutoff;
  int smooth;
  const float* mul_emb;    // epilogue: *= mul_emb[(mul_code[m] & 0xffff) * N + n]
  const int* mul_code;
  const float* residual;   // epilogue: += residual[m * ldr + n]
  int ldr;
  float* C;
  int ldc;
  const float* w3;         // final-dot epilogue: out_vec[m] (+)= sum_n v[m,n] * w3[n] + b3
  const float* b3;
  float* out_vec;
  int accumulate;
};

static inline GemmArgs tsd_gemm_args() {
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  return g;
}

// 4 consecutive k of the A operand of row m (m < M, k % 4 == 0)
__device__ __forceinline__ float4 tsd_load_a4(const GemmArgs& p, int m, int k) {
  switch (p.a_kind) {
    case TSD_A_EDGE_MLP0: {
      float l = p.len[m];
      float4 w = *reinterpret_cast<const float4*>(p.w0 + k);
      float4 b = *reinterpret_cast<const float4*>(p.b0 + k);
      return make_float4(tsd_act(p.act0, fmaf(l, w.x, b.x)), tsd_act(p.act0, fmaf(l, w.y, b.y)),
                         tsd_act(p.act0, fmaf(l, w.z, b.z)), tsd_act(p.act0, fmaf(l, w.w, b.w)));
    }
    case TSD_A_CAT: {
      int code = p.code[m];
      int hi = k >= p.H;
      int kk = k - (hi ? p.H : 0);
      int r = hi ? ((unsigned)code >> 16) : (code & 0xffff);
      float4 d = *reinterpret_cast<const float4*>(p.A + (size_t)m * p.lda + kk);
      float4 e = *reinterpret_cast<const float4*>(p.emb + (size_t)r * p.H + kk);
      return make_float4(d.x * e.x, d.y * e.y, d.z * e.z, d.w * e.w);
    }
    case TSD_A_PAIR: {
      if (k < p.H) {
        float4 a = *reinterpret_cast<const float4*>(p.h + (size_t)p.row[m] * p.H + k);
        float4 b = *reinterpret_cast<const float4*>(p.h + (size_t)p.col[m] * p.H + k);
        return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
      }
      return *reinterpret_cast<const float4*>(p.A + (size_t)m * p.lda + (k - p.H));
    }
    default:
      return *reinterpret_cast<const float4*>(p.A + (size_t)m * p.lda + k);
  }
}

// element (m, n) of the epilogue given the raw accumulator
__device__ __forceinline__ float tsd_epilogue(const GemmArgs& p, int m, int n, float acc, float cscale) {
  float v = acc;
  if (p.bias) v += p.bias[n];
  v = tsd_act(p.act, v);
  if (p.scale_len) v *= cscale;
  if (p.mul_emb) v *= p.mul_emb[(size_t)(p.mul_code[m] & 0xffff) * p.N + n];
  if (p.residual) v += p.residual[(size_t)m * p.ldr + n];
  return v;
}

int tsd_gemm(const GemmArgs& g, int math, cudaStream_t stream);        // dispatch
int tsd_gemm_ffma(const GemmArgs& g, cudaStream_t stream);             // gemm_ffma.cu
int tsd_gemm_tf32(const GemmArgs& g, cudaStream_t stream);             // gemm_tc.cu
