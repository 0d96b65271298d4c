// memproj_kernel.cuh — softmax-weighted projection of queries onto the memory bank:
//
//     out[q] = normalise( softmax(t * q . B^T) . B )          t = 100 in the reference
//
// the reference's map2memory (predict_prompt.py:23-29; defined, its call at :134 is commented
// out).  It is used with ONE audio embedding per call, so the work is a single pass over the
// bank: an HBM-bound stream, not a tensor-core problem.  The bank is read in fp32 straight from
// the caller's tensor (4*d bytes per row): with t = 100 a bf16 similarity error of 4e-4 would
// move a weight by 4 %.
//
// memproj_stream_kernel: one warp owns rows j = warp, warp + W, ...; per row it computes the QB
// similarities (warp reduction) and keeps, per query, a running maximum m, normaliser Z and
// weighted sum acc[d] (online softmax: acc is only rescaled when the maximum moves).  Every
// block folds its 8 warps and writes one partial (m, Z, acc).  memproj_combine_kernel folds the
// per-block partials per query with the usual exp(m_b - M) rescale and divides by Z;
// memproj_finalize_kernel L2-normalises.
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace zs {

constexpr int MEMPROJ_THREADS = 256;
constexpr int MEMPROJ_MAX_D = 1024;               // 8 float4 per lane
constexpr int MEMPROJ_VEC = MEMPROJ_MAX_D / 128;  // float4 slots per lane

// partial layout per (warp, query): [0] = m, [1] = Z, [2 .. 2+d) = acc
__host__ __device__ inline int64_t memproj_partial_stride(int d) { return d + 4; }

template <int QB>
__global__ void __launch_bounds__(MEMPROJ_THREADS)
memproj_stream_kernel(const float* __restrict__ queries,   // [QB_valid, d]
                      const float* __restrict__ bank,      // [n_rows, d] fp32
                      int64_t n_rows, int d, int n_valid_queries, float temperature,
                      float* __restrict__ partials) {      // [gridDim.x, QB, d + 4]
  __shared__ float4 q_s[QB][MEMPROJ_MAX_D / 4];
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const int64_t warp = static_cast<int64_t>(blockIdx.x) * (MEMPROJ_THREADS / 32) + warp_in_block;
  const int64_t total_warps = static_cast<int64_t>(gridDim.x) * (MEMPROJ_THREADS / 32);
  const int nvec = d / 4;                                   // float4 per row

  for (int i = threadIdx.x; i < QB * nvec; i += MEMPROJ_THREADS) {
    const int q = i / nvec, c = i % nvec;
    q_s[q][c] = (q < n_valid_queries)
                    ? reinterpret_cast<const float4*>(queries + static_cast<int64_t>(q) * d)[c]
                    : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();

  float m[QB], z[QB];
  float4 acc[QB][MEMPROJ_VEC];
#pragma unroll
  for (int q = 0; q < QB; ++q) {
    m[q] = -CUDART_INF_F;
    z[q] = 0.f;
#pragma unroll
    for (int v = 0; v < MEMPROJ_VEC; ++v) acc[q][v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }

  // lane owns float4 columns lane, lane + 32, ... (coalesced 512-byte warp loads)
  float4 cur[MEMPROJ_VEC], nxt[MEMPROJ_VEC];
  auto load_row = [&](int64_t row, float4 (&dst)[MEMPROJ_VEC]) {
    const float4* src = reinterpret_cast<const float4*>(bank + row * d);
#pragma unroll
    for (int v = 0; v < MEMPROJ_VEC; ++v) {
      const int c = lane + 32 * v;
      dst[v] = (c < nvec) ? __ldcs(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);   // streaming: read once
    }
  };

  int64_t row = warp;
  if (row < n_rows) load_row(row, cur);
  while (row < n_rows) {
    const int64_t next = row + total_warps;
    if (next < n_rows) load_row(next, nxt);                 // next row in flight while this one is used
#pragma unroll
    for (int q = 0; q < QB; ++q) {
      float s = 0.f;
#pragma unroll
      for (int v = 0; v < MEMPROJ_VEC; ++v) {
        const int c = lane + 32 * v;
        if (c < nvec) {
          const float4 qa = q_s[q][c];
          s = fmaf(qa.x, cur[v].x, s); s = fmaf(qa.y, cur[v].y, s);
          s = fmaf(qa.z, cur[v].z, s); s = fmaf(qa.w, cur[v].w, s);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float logit = temperature * s;                  // warp-uniform
      if (logit > m[q]) {                                    // new maximum: rescale what we have
        const float r = __expf(m[q] - logit);                // exp(-inf) = 0 on the first row
        z[q] *= r;
#pragma unroll
        for (int v = 0; v < MEMPROJ_VEC; ++v) {
          acc[q][v].x *= r; acc[q][v].y *= r; acc[q][v].z *= r; acc[q][v].w *= r;
        }
        m[q] = logit;
      }
      const float w = __expf(logit - m[q]);
      z[q] += w;
#pragma unroll
      for (int v = 0; v < MEMPROJ_VEC; ++v) {
        acc[q][v].x = fmaf(w, cur[v].x, acc[q][v].x); acc[q][v].y = fmaf(w, cur[v].y, acc[q][v].y);
        acc[q][v].z = fmaf(w, cur[v].z, acc[q][v].z); acc[q][v].w = fmaf(w, cur[v].w, acc[q][v].w);
      }
    }
#pragma unroll
    for (int v = 0; v < MEMPROJ_VEC; ++v) cur[v] = nxt[v];
    row = next;
  }

  // fold the 8 warps of the block into one partial per query (one after the other through smem:
  // runs once per launch) so that the combine step reads gridDim.x partials, not 8x as many
  __shared__ float m_s[MEMPROJ_THREADS / 32];
  __shared__ __align__(16) float acc_s[MEMPROJ_MAX_D];
  const int64_t stride = memproj_partial_stride(d);
#pragma unroll
  for (int q = 0; q < QB; ++q) {
    if (lane == 0) m_s[warp_in_block] = m[q];
    __syncthreads();
    float mb = -CUDART_INF_F;
#pragma unroll
    for (int w = 0; w < MEMPROJ_THREADS / 32; ++w) mb = fmaxf(mb, m_s[w]);
    const float r = (m[q] == -CUDART_INF_F) ? 0.f : __expf(m[q] - mb);    // warps without rows add 0
    float zb = z[q] * r;
    for (int w = 0; w < MEMPROJ_THREADS / 32; ++w) {
      if (warp_in_block == w) {
#pragma unroll
        for (int v = 0; v < MEMPROJ_VEC; ++v) {
          const int c = lane + 32 * v;
          if (c < nvec) {
            float4* a = reinterpret_cast<float4*>(acc_s) + c;
            float4 cur4 = (w == 0) ? make_float4(0.f, 0.f, 0.f, 0.f) : *a;
            cur4.x = fmaf(r, acc[q][v].x, cur4.x); cur4.y = fmaf(r, acc[q][v].y, cur4.y);
            cur4.z = fmaf(r, acc[q][v].z, cur4.z); cur4.w = fmaf(r, acc[q][v].w, cur4.w);
            *a = cur4;
          }
        }
      }
      __syncthreads();
    }
    // Z of the block: fixed-order sum over the warps
    __shared__ float z_s[MEMPROJ_THREADS / 32];
    if (lane == 0) z_s[warp_in_block] = zb;
    __syncthreads();
    float* dst = partials + (static_cast<int64_t>(blockIdx.x) * QB + q) * stride;
    if (threadIdx.x == 0) {
      float zt = 0.f;
      for (int w = 0; w < MEMPROJ_THREADS / 32; ++w) zt += z_s[w];
      dst[0] = mb;
      dst[1] = zt;
    }
    for (int c = threadIdx.x; c < nvec; c += MEMPROJ_THREADS)
      reinterpret_cast<float4*>(dst + 4)[c] = reinterpret_cast<const float4*>(acc_s)[c];
    __syncthreads();
  }
}

// Combine the per-block partials of one query: grid (d / 64 column groups, queries), 256 threads
// = 64 columns x 4 slices of the partial list.  Writes the un-normalised softmax-weighted sum.
constexpr int MEMPROJ_COMBINE_COLS = 64;
__global__ void __launch_bounds__(256)
memproj_combine_kernel(const float* __restrict__ partials, int n_partials, int qb, int d,
                       float* __restrict__ out /* [qb_valid, d] */) {
  const int q = blockIdx.y;
  const int64_t stride = memproj_partial_stride(d);
  __shared__ float red[256];
  __shared__ float colsum[4][MEMPROJ_COMBINE_COLS];

  float mx = -CUDART_INF_F;
  for (int b = threadIdx.x; b < n_partials; b += blockDim.x)
    mx = fmaxf(mx, partials[(static_cast<int64_t>(b) * qb + q) * stride]);
  red[threadIdx.x] = mx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + o]);
    __syncthreads();
  }
  const float M = red[0];
  __syncthreads();
  float zsum = 0.f;
  for (int b = threadIdx.x; b < n_partials; b += blockDim.x) {
    const float* p = partials + (static_cast<int64_t>(b) * qb + q) * stride;
    zsum += (p[0] == -CUDART_INF_F) ? 0.f : __expf(p[0] - M) * p[1];
  }
  red[threadIdx.x] = zsum;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  const float Z = red[0];

  const int col = blockIdx.x * MEMPROJ_COMBINE_COLS + (threadIdx.x & (MEMPROJ_COMBINE_COLS - 1));
  const int slice = threadIdx.x / MEMPROJ_COMBINE_COLS;          // 0..3
  float a = 0.f;
  if (col < d) {
    for (int b = slice; b < n_partials; b += 4) {
      const float* p = partials + (static_cast<int64_t>(b) * qb + q) * stride;
      const float e = (p[0] == -CUDART_INF_F) ? 0.f : __expf(p[0] - M);
      a = fmaf(e, p[4 + col], a);
    }
  }
  colsum[slice][threadIdx.x & (MEMPROJ_COMBINE_COLS - 1)] = a;
  __syncthreads();
  if (slice == 0 && col < d) {
    const int c = threadIdx.x;
    out[static_cast<int64_t>(q) * d + col] = (colsum[0][c] + colsum[1][c] + colsum[2][c] + colsum[3][c]) / Z;
  }
}

// out[q] /= ||out[q]||  (predict_prompt.py:28 — no epsilon there either; an all-zero row stays 0)
__global__ void __launch_bounds__(256)
memproj_finalize_kernel(float* __restrict__ out, int d) {
  const int q = blockIdx.x;
  __shared__ float red[256];
  float ss = 0.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const float v = out[static_cast<int64_t>(q) * d + c];
    ss = fmaf(v, v, ss);
  }
  red[threadIdx.x] = ss;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  const float norm = sqrtf(red[0]);
  if (norm > 0.f)
    for (int c = threadIdx.x; c < d; c += blockDim.x) out[static_cast<int64_t>(q) * d + c] /= norm;
}

}  // namespace zs
