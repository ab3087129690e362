// Backward (and the few extra forward) kernels of the training step -- SURVEY.md section 8(f)-2, BASELINE config 4:
// `get_loss(...).mean().backward()` of CondenseEncoderEpsNetwork (models/epsnet/condensenc.py:267-328, train.py:124-152).
//
// The training forward runs the same operators as sampling but UNFUSED (every pre-activation is kept for the
// backward); the host side (tsdiff_b200/training.py) strings these entry points together behind torch.autograd
// Functions, so autograd is the tape and every arithmetic operation is a kernel of this library.  fp32 throughout.
// Reductions are deterministic (fixed split + ordered second pass, segmented sums over the CSR) -- no float atomics.
#include "common.cuh"

// ----------------------------------------------------------------------------- elementwise
__device__ __forceinline__ float tr_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void k_act_forward(long long n, const float* __restrict__ x, int act, float* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = tsd_act(act, x[i]);
}

// dx = dy * act'(x): swish' = s (1 + x (1 - s)), ssp' = softplus' = s, relu' = [x > 0]   (s = sigmoid(x))
__global__ void k_act_backward(long long n, const float* __restrict__ x, const float* __restrict__ dy, int act,
                               float* __restrict__ dx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i], s = tr_sigmoid(v);
  float d = 1.f;
  if (act == TSD_ACT_SWISH) d = s * (1.f + v * (1.f - s));
  else if (act == TSD_ACT_SSP || act == TSD_ACT_SOFTPLUS) d = s;
  else if (act == TSD_ACT_RELU) d = v > 0.f ? 1.f : 0.f;
  dx[i] = dy[i] * d;
}

extern "C" int tsd_act_forward(int64_t n, const float* x, int32_t act, float* y, tsd_stream_t stream) {
  TSD_REQUIRE(x && y && n >= 0);
  if (n == 0) return TSD_OK;
  k_act_forward<<<(unsigned)((n + 255) / 256), 256, 0, tsd_cu(stream)>>>(n, x, act, y);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

extern "C" int tsd_act_backward(int64_t n, const float* x, const float* dy, int32_t act, float* dx, tsd_stream_t stream) {
  TSD_REQUIRE(x && dy && dx && n >= 0);
  if (n == 0) return TSD_OK;
  k_act_backward<<<(unsigned)((n + 255) / 256), 256, 0, tsd_cu(stream)>>>(n, x, dy, act, dx);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

// out[m, :] = x[m, :] * s[m]   (the cutoff envelope C(len) of schnet.py:91-98; its own backward with dy for x)
__global__ void k_row_scale(int rows, int H, const float* __restrict__ x, const float* __restrict__ s, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (long long)rows * H) out[i] = x[i] * s[i / H];
}

__global__ void k_cutoff_envelope(int rows, const float* __restrict__ len, float cutoff, int smooth, float* __restrict__ s) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows) s[i] = tsd_cutoff_fn(len[i], cutoff, smooth);
}

extern "C" int tsd_row_scale(int32_t rows, int32_t H, const float* x, const float* s, float* out, tsd_stream_t stream) {
  TSD_REQUIRE(x && s && out && rows >= 0 && H > 0);
  if (rows == 0) return TSD_OK;
  const long long n = (long long)rows * H;
  k_row_scale<<<(unsigned)((n + 255) / 256), 256, 0, tsd_cu(stream)>>>(rows, H, x, s, out);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

extern "C" int tsd_cutoff_envelope(int32_t rows, const float* len, float cutoff, int32_t smooth, float* s,
                                   tsd_stream_t stream) {
  TSD_REQUIRE(len && s && rows >= 0);
  if (rows == 0) return TSD_OK;
  k_cutoff_envelope<<<tsd_ceil_div(rows, 256), 256, 0, tsd_cu(stream)>>>(rows, len, cutoff, smooth, s);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

// out[m, :] = a[m, :] * table[(code[m] >> shift) & 0xffff, :]   (edge.py:66-68: d_emb * bond_emb[type]); with
// a == NULL: out = the gathered rows.  Its backward w.r.t. `a` is the same op applied to dy; the gradient of the table
// is a weight-gradient GEMM against the one-hot matrix (tsd_onehot + tsd_linear_wgrad), which keeps it deterministic.
__global__ void k_gate_rows(int rows, int H, const float* __restrict__ a, const float* __restrict__ table,
                            const int* __restrict__ code, int shift, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * H) return;
  const int m = (int)(i / H), h = (int)(i - (long long)m * H);
  const float t = table[(size_t)((code[m] >> shift) & 0xffff) * H + h];
  out[i] = a ? a[i] * t : t;
}

__global__ void k_onehot(int rows, int classes, const int* __restrict__ code, int shift, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * classes) return;
  const int m = (int)(i / classes), c = (int)(i - (long long)m * classes);
  out[i] = ((code[m] >> shift) & 0xffff) == c ? 1.f : 0.f;
}

extern "C" int tsd_gate_rows(int32_t rows, int32_t H, const float* a, const float* table, const int32_t* code,
                             int32_t shift, float* out, tsd_stream_t stream) {
  TSD_REQUIRE(table && code && out && rows >= 0 && H > 0);
  if (rows == 0) return TSD_OK;
  const long long n = (long long)rows * H;
  k_gate_rows<<<(unsigned)((n + 255) / 256), 256, 0, tsd_cu(stream)>>>(rows, H, a, table, code, shift, out);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

extern "C" int tsd_onehot(int32_t rows, int32_t classes, const int32_t* code, int32_t shift, float* out,
                          tsd_stream_t stream) {
  TSD_REQUIRE(code && out && rows >= 0 && classes > 0);
  if (rows == 0) return TSD_OK;
  const long long n = (long long)rows * classes;
  k_onehot<<<(unsigned)((n + 255) / 256), 256, 0, tsd_cu(stream)>>>(rows, classes, code, shift, out);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

// ------------------------------------------------------------------------------ small / odd-shaped linear layers
// out[m, n] = sum_k x[m, k] w[n, k] + b[n] for shapes the tiled FFMA kernel does not take (K = 1, 25; N = 1):
// edge MLP layer 0, the feature embedding, the last layer of grad_dist_mlp and their data gradients.
__global__ void k_linear_generic(int M, int N, int K, const float* __restrict__ x, const float* __restrict__ w,
                                 const float* __restrict__ b, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)M * N) return;
  const int m = (int)(i / N), n = (int)(i - (long long)m * N);
  float s = 0.f;
  for (int k = 0; k < K; ++k) s = fmaf(x[(size_t)m * K + k], w[(size_t)n * K + k], s);
  out[i] = s + (b ? b[n] : 0.f);
}

int tsd_linear_generic(int M, int N, int K, const float* x, const float* w, const float* b, float* out, cudaStream_t s) {
  if (M == 0) return TSD_OK;
  const long long n = (long long)M * N;
  k_linear_generic<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(M, N, K, x, w, b, out);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

__global__ void k_transpose(int rows, int cols, const float* __restrict__ src, float* __restrict__ dst) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[i][threadIdx.x] = src[(size_t)r * cols + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[(size_t)c * rows + r] = tile[threadIdx.x][i];
  }
}

extern "C" int tsd_transpose(int32_t rows, int32_t cols, const float* src, float* dst, tsd_stream_t stream) {
  TSD_REQUIRE(src && dst && rows >= 0 && cols >= 0);
  if (rows == 0 || cols == 0) return TSD_OK;
  k_transpose<<<dim3(tsd_ceil_div(cols, 32), tsd_ceil_div(rows, 32)), dim3(32, 8), 0, tsd_cu(stream)>>>(rows, cols, src, dst);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

// ------------------------------------------------------------------------------ weight gradient
// dW[n, k] = sum_m dy[m, n] x[m, k],  db[n] = sum_m dy[m, n].  The rows are cut into `splits` contiguous ranges; CTA
// (n tile, k tile, split) accumulates a 64 x 64 tile of its range (16 x 16 threads, 4 x 4 outputs each, 16-row
// chunks through shared memory) into scratch[split]; the second kernel adds the splits in order.
constexpr int WG_T = 64, WG_CH = 16;

__global__ void __launch_bounds__(256) k_wgrad_partial(int M, int N, int K, int rows_per_split, const float* __restrict__ dy,
                                                       const float* __restrict__ x, float* __restrict__ part_w,
                                                       float* __restrict__ part_b) {
  __shared__ float sa[WG_CH][WG_T + 1], sb[WG_CH][WG_T + 1];
  const int n0 = blockIdx.x * WG_T, k0 = blockIdx.y * WG_T, split = blockIdx.z;
  const int m_beg = split * rows_per_split, m_end = min(M, m_beg + rows_per_split);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  float bsum[4] = {};
  for (int m0 = m_beg; m0 < m_end; m0 += WG_CH) {
    for (int i = threadIdx.x; i < WG_CH * WG_T; i += 256) {
      const int r = i / WG_T, c = i - r * WG_T, m = m0 + r;
      sa[r][c] = (m < m_end && n0 + c < N) ? dy[(size_t)m * N + n0 + c] : 0.f;
      sb[r][c] = (m < m_end && k0 + c < K) ? x[(size_t)m * K + k0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < WG_CH; ++r) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = sa[r][ty * 4 + i];
        b[i] = sb[r][tx * 4 + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        bsum[i] += a[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
  float* pw = part_w + (size_t)split * N * K;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < K) pw[(size_t)n * K + k] = acc[i][j];
    }
    if (part_b && blockIdx.y == 0 && tx == 0) part_b[(size_t)split * N + n] = bsum[i];
  }
}

__global__ void k_wgrad_reduce(int count, int splits, const float* __restrict__ part, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float s = 0.f;
  for (int p = 0; p < splits; ++p) s += part[(size_t)p * count + i];
  out[i] = s;
}

static int wgrad_splits(int M) {
  int s = (M + 511) / 512;
  return s < 1 ? 1 : (s > 64 ? 64 : s);
}

extern "C" int tsd_linear_wgrad_scratch(int32_t M, int32_t N, int32_t K, uint64_t* floats) {
  TSD_REQUIRE(floats && M >= 0 && N > 0 && K > 0);
  *floats = (uint64_t)wgrad_splits(M) * ((uint64_t)N * K + N);
  return TSD_OK;
}

extern "C" int tsd_linear_wgrad(int32_t M, int32_t N, int32_t K, const float* dy, const float* x, float* dW, float* db,
                                float* scratch, tsd_stream_t stream) {
  TSD_REQUIRE(dy && x && dW && scratch && M >= 0 && N > 0 && K > 0);
  cudaStream_t s = tsd_cu(stream);
  if (M == 0) {
    TSD_CUDA(cudaMemsetAsync(dW, 0, (size_t)N * K * sizeof(float), s));
    if (db) TSD_CUDA(cudaMemsetAsync(db, 0, (size_t)N * sizeof(float), s));
    return TSD_OK;
  }
  const int splits = wgrad_splits(M);
  const int rows_per_split = (tsd_ceil_div(M, splits) + WG_CH - 1) / WG_CH * WG_CH;
  float* part_w = scratch;
  float* part_b = db ? scratch + (size_t)splits * N * K : nullptr;
  k_wgrad_partial<<<dim3(tsd_ceil_div(N, WG_T), tsd_ceil_div(K, WG_T), splits), 256, 0, s>>>(M, N, K, rows_per_split, dy, x,
                                                                                               part_w, part_b);
  TSD_LAUNCH_CHECK();
  k_wgrad_reduce<<<tsd_ceil_div(N * K, 256), 256, 0, s>>>(N * K, splits, part_w, dW);
  TSD_LAUNCH_CHECK();
  if (db) {
    k_wgrad_reduce<<<tsd_ceil_div(N, 256), 256, 0, s>>>(N, splits, part_b, db);
    TSD_LAUNCH_CHECK();
  }
  return TSD_OK;
}

// ------------------------------------------------------------------------------ CFConv aggregation, backward
// forward: agg_i = sum_{e: col[e] = i} x1[row[e]] * filt[e]            (schnet.py:102-107)
// dfilt[e] = dagg[col[e]] * x1[row[e]];   dx1[j] = sum_{e: row[e] = j} dagg[col[e]] * filt[e]  -- the edges are
// sorted by row, so the second sum runs over the contiguous out-CSR segment of j in edge order.
__global__ void k_cfconv_backward_filt(int num_edges, int H, const int* __restrict__ row, const int* __restrict__ col,
                                       const float* __restrict__ x1, const float* __restrict__ dagg,
                                       float* __restrict__ dfilt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)num_edges * (H / 4)) return;
  const int e = (int)(i / (H / 4)), c = (int)(i - (long long)e * (H / 4)) * 4;
  const float4 a = *reinterpret_cast<const float4*>(dagg + (size_t)col[e] * H + c);
  const float4 b = *reinterpret_cast<const float4*>(x1 + (size_t)row[e] * H + c);
  *reinterpret_cast<float4*>(dfilt + (size_t)e * H + c) = make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
}

__global__ void __launch_bounds__(256) k_cfconv_backward_x1(int num_nodes, int H, int slabs, const int* __restrict__ row_ptr,
                                                            const int* __restrict__ col, const float* __restrict__ filt,
                                                            const float* __restrict__ dagg, float* __restrict__ dx1) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int node = gw / slabs, slab = gw - node * slabs;
  if (node >= num_nodes) return;
  const int off = slab * 128 + lane * 4;
  if (off >= H) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int e = row_ptr[node]; e < row_ptr[node + 1]; ++e) {
    const float4 g = *reinterpret_cast<const float4*>(dagg + (size_t)col[e] * H + off);
    const float4 w = *reinterpret_cast<const float4*>(filt + (size_t)e * H + off);
    acc.x = fmaf(g.x, w.x, acc.x);
    acc.y = fmaf(g.y, w.y, acc.y);
    acc.z = fmaf(g.z, w.z, acc.z);
    acc.w = fmaf(g.w, w.w, acc.w);
  }
  *reinterpret_cast<float4*>(dx1 + (size_t)node * H + off) = acc;
}

extern "C" int tsd_cfconv_aggregate_backward(const tsd_batch_t* batch, const tsd_edges_t* edges, int32_t num_edges,
                                             int32_t H, const float* x1, const float* filt, const float* dagg, float* dx1,
                                             float* dfilt, tsd_stream_t stream) {
  TSD_REQUIRE(batch && edges && x1 && filt && dagg && dx1 && dfilt && H % 4 == 0 && H > 0 && num_edges >= 0);
  cudaStream_t s = tsd_cu(stream);
  if (num_edges > 0) {
    const long long n = (long long)num_edges * (H / 4);
    k_cfconv_backward_filt<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(num_edges, H, edges->row, edges->col, x1, dagg, dfilt);
    TSD_LAUNCH_CHECK();
  }
  if (batch->num_nodes > 0) {
    const int slabs = tsd_ceil_div(H, 128);
    const long long warps = (long long)batch->num_nodes * slabs;
    k_cfconv_backward_x1<<<(unsigned)((warps + 7) / 8), 256, 0, s>>>(batch->num_nodes, H, slabs, edges->row_ptr, edges->col, filt,
                                                                     dagg, dx1);
    TSD_LAUNCH_CHECK();
  }
  return TSD_OK;
}

// ------------------------------------------------------------------------------ pair features (common.py:226-229)
// out[e] = cat[h[row[e]] * h[col[e]], ea[e]]  (E, 2H);  backward: dea = dout[:, H:] (a view on the host side) and
// dh[i] = sum_{e: row[e] = i} dout[e, :H] * h[col[e]] + sum_{e: col[e] = i} dout[e, :H] * h[row[e]]  (out-CSR, then in-CSR)
__global__ void k_pair_features(int num_edges, int H, const int* __restrict__ row, const int* __restrict__ col,
                                const float* __restrict__ h, const float* __restrict__ ea, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)num_edges * (H / 4)) return;
  const int e = (int)(i / (H / 4)), c = (int)(i - (long long)e * (H / 4)) * 4;
  const float4 a = *reinterpret_cast<const float4*>(h + (size_t)row[e] * H + c);
  const float4 b = *reinterpret_cast<const float4*>(h + (size_t)col[e] * H + c);
  *reinterpret_cast<float4*>(out + (size_t)e * 2 * H + c) = make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
  *reinterpret_cast<float4*>(out + (size_t)e * 2 * H + H + c) = *reinterpret_cast<const float4*>(ea + (size_t)e * H + c);
}

__global__ void __launch_bounds__(256) k_pair_features_backward(int num_nodes, int H, int slabs, const int* __restrict__ row_ptr,
                                                                const int* __restrict__ col, const int* __restrict__ in_ptr,
                                                                const int* __restrict__ in_eid, const int* __restrict__ in_src,
                                                                const float* __restrict__ h, const float* __restrict__ dout,
                                                                float* __restrict__ dh) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int node = gw / slabs, slab = gw - node * slabs;
  if (node >= num_nodes) return;
  const int off = slab * 128 + lane * 4;
  if (off >= H) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int e = row_ptr[node]; e < row_ptr[node + 1]; ++e) {
    const float4 g = *reinterpret_cast<const float4*>(dout + (size_t)e * 2 * H + off);
    const float4 o = *reinterpret_cast<const float4*>(h + (size_t)col[e] * H + off);
    acc.x = fmaf(g.x, o.x, acc.x);
    acc.y = fmaf(g.y, o.y, acc.y);
    acc.z = fmaf(g.z, o.z, acc.z);
    acc.w = fmaf(g.w, o.w, acc.w);
  }
  for (int k = in_ptr[node]; k < in_ptr[node + 1]; ++k) {
    const float4 g = *reinterpret_cast<const float4*>(dout + (size_t)in_eid[k] * 2 * H + off);
    const float4 o = *reinterpret_cast<const float4*>(h + (size_t)in_src[k] * H + off);
    acc.x = fmaf(g.x, o.x, acc.x);
    acc.y = fmaf(g.y, o.y, acc.y);
    acc.z = fmaf(g.z, o.z, acc.z);
    acc.w = fmaf(g.w, o.w, acc.w);
  }
  *reinterpret_cast<float4*>(dh + (size_t)node * H + off) = acc;
}

extern "C" int tsd_pair_features(const tsd_edges_t* edges, int32_t num_edges, int32_t H, const float* h, const float* ea,
                                 float* out, tsd_stream_t stream) {
  TSD_REQUIRE(edges && h && ea && out && H % 4 == 0 && H > 0 && num_edges >= 0);
  if (num_edges == 0) return TSD_OK;
  const long long n = (long long)num_edges * (H / 4);
  k_pair_features<<<(unsigned)((n + 255) / 256), 256, 0, tsd_cu(stream)>>>(num_edges, H, edges->row, edges->col, h, ea, out);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

extern "C" int tsd_pair_features_backward(const tsd_batch_t* batch, const tsd_edges_t* edges, int32_t H, const float* h,
                                          const float* dout, float* dh, tsd_stream_t stream) {
  TSD_REQUIRE(batch && edges && h && dout && dh && H % 4 == 0 && H > 0);
  if (batch->num_nodes == 0) return TSD_OK;
  const int slabs = tsd_ceil_div(H, 128);
  const long long warps = (long long)batch->num_nodes * slabs;
  k_pair_features_backward<<<(unsigned)((warps + 7) / 8), 256, 0, tsd_cu(stream)>>>(
      batch->num_nodes, H, slabs, edges->row_ptr, edges->col, edges->in_ptr, edges->in_eid, edges->in_src, h, dout, dh);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

// ------------------------------------------------------------------------------ eq_transform, backward w.r.t. the edge scores
// forward (geometry.py:22-30): node_i = sum_{row = i} u_e s_e - sum_{col = i} u_e s_e,  u_e = (p_row - p_col) / len_e,
// s_e = inv_e / inv_div on the selected edges.  d inv_e = u_e . (dnode[row] - dnode[col]) / inv_div  (0 when not selected).
__global__ void k_eq_transform_backward(int num_edges, const int* __restrict__ row, const int* __restrict__ col,
                                        const float* __restrict__ length, const int* __restrict__ mask, int mask_mode,
                                        const float* __restrict__ pos, const float* __restrict__ dnode, float inv_div,
                                        float* __restrict__ dinv) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= num_edges) return;
  bool sel = true;
  if (mask_mode != 0 && mask) sel = mask_mode == 1 ? mask[e] != 0 : mask[e] == 0;
  float g = 0.f;
  if (sel) {
    const int r = row[e], c = col[e];
    const float il = 1.0f / length[e];
    g = (il * (pos[3 * r] - pos[3 * c])) * (dnode[3 * r] - dnode[3 * c]) +
        (il * (pos[3 * r + 1] - pos[3 * c + 1])) * (dnode[3 * r + 1] - dnode[3 * c + 1]) +
        (il * (pos[3 * r + 2] - pos[3 * c + 2])) * (dnode[3 * r + 2] - dnode[3 * c + 2]);
    g /= inv_div;
  }
  dinv[e] = g;
}

extern "C" int tsd_eq_transform_backward(const tsd_edges_t* edges, int32_t num_edges, const float* pos, const int32_t* mask,
                                         int32_t mask_mode, float inv_div, const float* dnode, float* dinv,
                                         tsd_stream_t stream) {
  TSD_REQUIRE(edges && pos && dnode && dinv && num_edges >= 0);
  if (num_edges == 0) return TSD_OK;
  k_eq_transform_backward<<<tsd_ceil_div(num_edges, 256), 256, 0, tsd_cu(stream)>>>(num_edges, edges->row, edges->col,
                                                                                    edges->length, mask, mask_mode, pos, dnode,
                                                                                    inv_div, dinv);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

// ------------------------------------------------------------------------------ node embedding of path B, backward
// forward (condensenc.py:193-198): z[n] = cat[emb[Z_n] + W r_n, W p_n - W r_n],  W (half, F).
// d emb[a] = sum_{n: Z_n = a} dz[n, :half]   (one CTA per embedding row, atoms in index order: deterministic)
// d W[j, f] = sum_n dz[n, j] r[n, f] + dz[n, half + j] (p[n, f] - r[n, f])
__global__ void k_node_embed_backward_emb(int num_nodes, int half, const int64_t* __restrict__ atom_type,
                                          const float* __restrict__ dz, float* __restrict__ demb) {
  const int a = blockIdx.x;
  for (int j = threadIdx.x; j < half; j += blockDim.x) {
    float s = 0.f;
    for (int n = 0; n < num_nodes; ++n)
      if (atom_type[n] == a) s += dz[(size_t)n * 2 * half + j];
    demb[(size_t)a * half + j] = s;
  }
}

__global__ void k_node_embed_backward_w(int num_nodes, int half, int F, const int64_t* __restrict__ r_feat,
                                        const int64_t* __restrict__ p_feat, const float* __restrict__ dz,
                                        float* __restrict__ dw) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= half * F) return;
  const int j = idx / F, f = idx - j * F;
  float s = 0.f;
  for (int n = 0; n < num_nodes; ++n) {
    const float r = (float)r_feat[(size_t)n * F + f], p = (float)p_feat[(size_t)n * F + f];
    s = fmaf(dz[(size_t)n * 2 * half + j], r, s);
    s = fmaf(dz[(size_t)n * 2 * half + half + j], p - r, s);
  }
  dw[idx] = s;
}

extern "C" int tsd_condensed_node_embed_backward(int32_t num_nodes, const int64_t* atom_type, const int64_t* r_feat,
                                                 const int64_t* p_feat, int32_t feat_dim, int32_t half, int32_t num_types,
                                                 const float* dz, float* d_atom_emb, float* d_feat_weight,
                                                 tsd_stream_t stream) {
  TSD_REQUIRE(atom_type && r_feat && p_feat && dz && d_atom_emb && d_feat_weight && half > 0 && feat_dim > 0 && num_types > 0);
  cudaStream_t s = tsd_cu(stream);
  k_node_embed_backward_emb<<<num_types, 128, 0, s>>>(num_nodes, half, atom_type, dz, d_atom_emb);
  TSD_LAUNCH_CHECK();
  k_node_embed_backward_w<<<tsd_ceil_div(half * feat_dim, 128), 128, 0, s>>>(num_nodes, half, feat_dim, r_feat, p_feat, dz,
                                                                            d_feat_weight);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

// ------------------------------------------------------------------------------ per-atom squared error (condensenc.py:324-326)
// loss[n] = sum_d (a[n, d] - b[n, d])^2;  backward: da[n, d] = 2 (a - b) dloss[n]
__global__ void k_sqerr_forward(int n, const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ loss) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int d = 0; d < 3; ++d) {
    const float t = a[3 * i + d] - b[3 * i + d];
    s += t * t;
  }
  loss[i] = s;
}

__global__ void k_sqerr_backward(int n, const float* __restrict__ a, const float* __restrict__ b,
                                 const float* __restrict__ dloss, float* __restrict__ da) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * n) return;
  da[i] = 2.f * (a[i] - b[i]) * dloss[i / 3];
}

extern "C" int tsd_sqerr_forward(int32_t n, const float* a, const float* b, float* loss, tsd_stream_t stream) {
  TSD_REQUIRE(a && b && loss && n >= 0);
  if (n == 0) return TSD_OK;
  k_sqerr_forward<<<tsd_ceil_div(n, 256), 256, 0, tsd_cu(stream)>>>(n, a, b, loss);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

extern "C" int tsd_sqerr_backward(int32_t n, const float* a, const float* b, const float* dloss, float* da,
                                  tsd_stream_t stream) {
  TSD_REQUIRE(a && b && dloss && da && n >= 0);
  if (n == 0) return TSD_OK;
  k_sqerr_backward<<<tsd_ceil_div(3 * n, 256), 256, 0, tsd_cu(stream)>>>(n, a, b, dloss, da);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

// out = a + b (the residual h + interaction(h), schnet.py:212-213; backward: the gradient passes to both)
__global__ void k_add(long long n, const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}

extern "C" int tsd_add(int64_t n, const float* a, const float* b, float* out, tsd_stream_t stream) {
  TSD_REQUIRE(a && b && out && n >= 0);
  if (n == 0) return TSD_OK;
  k_add<<<(unsigned)((n + 255) / 256), 256, 0, tsd_cu(stream)>>>(n, a, b, out);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

// out = a * b (elementwise; the bond-embedding gradient multiplies dy by the gated operand before the one-hot wgrad)
__global__ void k_mul(long long n, const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] * b[i];
}

extern "C" int tsd_mul(int64_t n, const float* a, const float* b, float* out, tsd_stream_t stream) {
  TSD_REQUIRE(a && b && out && n >= 0);
  if (n == 0) return TSD_OK;
  k_mul<<<(unsigned)((n + 255) / 256), 256, 0, tsd_cu(stream)>>>(n, a, b, out);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}
