// Post-sampling geometry metrics on the GPU, fp64 like the reference's numpy / scipy code:
//   tsd_dmae          clustering.py:98-105   calc_DMAE: mean |D_ref - D_guess| (or the relative error) over the
//                                            strict upper triangle, for a batch of guess geometries
//   tsd_min_match     clustering.py:123-135  get_minimum_matches: min over atom permutations of
//                                            sum_{i<j} (d_ref(i,j) - d_prb(match[i], match[j]))^2, value and arg-min
// One CTA per (geometry, slice of the permutation list); pair terms are summed in pdist order (i < j, row-major)
// by one thread per permutation, so the result does not depend on the launch shape.
#include "common.cuh"

__global__ void k_dmae(int n, int batch, const double* __restrict__ pos_ref, const double* __restrict__ pos_guess,
                       int mape, double* __restrict__ out) {
  // pair p = (i, j), i < j, row-major; every thread sums a strided subset in fp64, block tree reduction in a fixed order
  __shared__ double red[256];
  const int b = blockIdx.x;
  const double* g = pos_guess + (size_t)b * n * 3;
  double s = 0.0;
  const int pairs = n * (n - 1) / 2;
  for (int p = threadIdx.x; p < pairs; p += blockDim.x) {
    int i = 0, rem = p;
    while (rem >= n - 1 - i) {
      rem -= n - 1 - i;
      ++i;
    }
    const int j = i + 1 + rem;
    const double rx = pos_ref[3 * i] - pos_ref[3 * j], ry = pos_ref[3 * i + 1] - pos_ref[3 * j + 1],
                 rz = pos_ref[3 * i + 2] - pos_ref[3 * j + 2];
    const double gx = g[3 * i] - g[3 * j], gy = g[3 * i + 1] - g[3 * j + 1], gz = g[3 * i + 2] - g[3 * j + 2];
    const double dr = sqrt(rx * rx + ry * ry + rz * rz), dg = sqrt(gx * gx + gy * gy + gz * gz);
    s += mape ? fabs(dr - dg) / dr : fabs(dr - dg);
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[b] = red[0] / n / (n - 1) * 2;
}

extern "C" int tsd_dmae_pos(int32_t num_atoms, int32_t batch, const double* pos_ref, const double* pos_guess, int32_t mape,
                            double* out, tsd_stream_t stream) {
  TSD_REQUIRE(pos_ref && pos_guess && out && num_atoms >= 2 && batch >= 0);
  if (batch == 0) return TSD_OK;
  k_dmae<<<batch, 256, 0, tsd_cu(stream)>>>(num_atoms, batch, pos_ref, pos_guess, mape, out);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

// the reference's signature: distance MATRICES in (clustering.py:98-105)
__global__ void k_dmae_matrix(int n, const double* __restrict__ dm_ref, const double* __restrict__ dm_guess, int mape,
                              double* __restrict__ out) {
  __shared__ double red[256];
  const double* g = dm_guess + (size_t)blockIdx.x * n * n;
  double s = 0.0;
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
    const int i = idx / n, j = idx - i * n;
    if (j > i) s += mape ? fabs(dm_ref[idx] - g[idx]) / dm_ref[idx] : fabs(dm_ref[idx] - g[idx]);
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = red[0] / n / (n - 1) * 2;
}

extern "C" int tsd_dmae(int32_t num_atoms, int32_t batch, const double* dm_ref, const double* dm_guess, int32_t mape,
                        double* out, tsd_stream_t stream) {
  TSD_REQUIRE(dm_ref && dm_guess && out && num_atoms >= 2 && batch >= 0);
  if (batch == 0) return TSD_OK;
  k_dmae_matrix<<<batch, 256, 0, tsd_cu(stream)>>>(num_atoms, dm_ref, dm_guess, mape, out);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}

// partial[(b, chunk)] = (min value, arg-min) over the chunk's permutations; a second tiny kernel reduces the chunks
__global__ void k_min_match(int n, int num_matches, const double* __restrict__ pos_ref, const double* __restrict__ pos_prb,
                            const int32_t* __restrict__ matches, double* __restrict__ part_val, int* __restrict__ part_idx) {
  extern __shared__ double sm[];  // d_ref in pdist order, then this geometry's positions
  const int b = blockIdx.y, pairs = n * (n - 1) / 2;
  double* dref = sm;
  double* prb = sm + pairs;
  for (int i = threadIdx.x; i < 3 * n; i += blockDim.x) prb[i] = pos_prb[(size_t)b * 3 * n + i];
  for (int p = threadIdx.x; p < pairs; p += blockDim.x) {
    int i = 0, rem = p;
    while (rem >= n - 1 - i) {
      rem -= n - 1 - i;
      ++i;
    }
    const int j = i + 1 + rem;
    const double x = pos_ref[3 * i] - pos_ref[3 * j], y = pos_ref[3 * i + 1] - pos_ref[3 * j + 1],
                 z = pos_ref[3 * i + 2] - pos_ref[3 * j + 2];
    dref[p] = sqrt(x * x + y * y + z * z);
  }
  __syncthreads();
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  double best = INFINITY;
  int best_m = 0x7fffffff;
  if (m < num_matches) {
    const int32_t* perm = matches + (size_t)m * n;
    double s = 0.0;
    int p = 0;
    for (int i = 0; i < n; ++i) {
      const int a = perm[i];
      for (int j = i + 1; j < n; ++j, ++p) {
        const int c = perm[j];
        const double x = prb[3 * a] - prb[3 * c], y = prb[3 * a + 1] - prb[3 * c + 1], z = prb[3 * a + 2] - prb[3 * c + 2];
        const double d = dref[p] - sqrt(x * x + y * y + z * z);
        s += d * d;
      }
    }
    best = s;
    best_m = m;
  }
  // block arg-min; ties go to the SMALLER permutation index (list.index(min(...)) in the reference)
  __shared__ double rv[256];
  __shared__ int ri[256];
  rv[threadIdx.x] = best;
  ri[threadIdx.x] = best_m;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const double v = rv[threadIdx.x + o];
      const int k = ri[threadIdx.x + o];
      if (v < rv[threadIdx.x] || (v == rv[threadIdx.x] && k < ri[threadIdx.x])) {
        rv[threadIdx.x] = v;
        ri[threadIdx.x] = k;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    part_val[(size_t)b * gridDim.x + blockIdx.x] = rv[0];
    part_idx[(size_t)b * gridDim.x + blockIdx.x] = ri[0];
  }
}

__global__ void k_min_match_reduce(int batch, int chunks, const double* __restrict__ part_val,
                                   const int* __restrict__ part_idx, double* __restrict__ out_val,
                                   int* __restrict__ out_idx) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  double v = part_val[(size_t)b * chunks];
  int k = part_idx[(size_t)b * chunks];
  for (int c = 1; c < chunks; ++c) {
    const double w = part_val[(size_t)b * chunks + c];
    const int kk = part_idx[(size_t)b * chunks + c];
    if (w < v || (w == v && kk < k)) {
      v = w;
      k = kk;
    }
  }
  out_val[b] = v;
  out_idx[b] = k;
}

extern "C" int tsd_min_match_scratch(int32_t batch, int32_t num_matches, int64_t* doubles, int64_t* ints) {
  const int64_t chunks = (num_matches + 255) / 256;
  if (doubles) *doubles = (int64_t)batch * chunks;
  if (ints) *ints = (int64_t)batch * chunks;
  return TSD_OK;
}

extern "C" int tsd_min_match(int32_t num_atoms, int32_t batch, int32_t num_matches, const double* pos_ref,
                             const double* pos_prb, const int32_t* matches, double* scratch_val, int32_t* scratch_idx,
                             double* out_val, int32_t* out_idx, tsd_stream_t stream) {
  TSD_REQUIRE(pos_ref && pos_prb && matches && scratch_val && scratch_idx && out_val && out_idx);
  TSD_REQUIRE(num_atoms >= 2 && num_atoms <= 512 && num_matches >= 1 && batch >= 0);
  if (batch == 0) return TSD_OK;
  const int chunks = (num_matches + 255) / 256;
  const size_t smem = ((size_t)num_atoms * (num_atoms - 1) / 2 + 3 * (size_t)num_atoms) * sizeof(double);
  if (smem > 48 * 1024) TSD_CUDA(cudaFuncSetAttribute(k_min_match, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_min_match<<<dim3(chunks, batch), 256, smem, tsd_cu(stream)>>>(num_atoms, num_matches, pos_ref, pos_prb, matches,
                                                                   scratch_val, scratch_idx);
  TSD_LAUNCH_CHECK();
  k_min_match_reduce<<<tsd_ceil_div(batch, 128), 128, 0, tsd_cu(stream)>>>(batch, chunks, scratch_val, scratch_idx, out_val,
                                                                            out_idx);
  TSD_LAUNCH_CHECK();
  return TSD_OK;
}
