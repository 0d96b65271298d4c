// memproj_tc_kernels.cuh — the batched memory projection on the tensor cores:
//
//     out[q] = normalise( softmax(t * q . B^T) . B )     (reference predict_prompt.py:23-29, t = 100)
//
// for batches of queries (Q >= 8), where the streaming kernel of memproj_kernel.cuh — one pass
// over the fp32 bank per query — loses to two library GEMMs.  Both contractions run on the fused
// kernel's TMA + tcgen05 pipeline (zs_simtopk_kernel in DUMP mode: plain fp32 tile stores):
//
//   S = Q . B^T      [Q, N]   K = d          "queries" Q', "bank" B'
//   O = P . B        [Q, d]   K = N          "queries" P'', "bank" Bt'' (the bank transposed), split-K
//
// With t = 100 a bf16 similarity error of 4e-4 would move a weight by 4 %, so every operand is
// split into two bf16 terms, x = hi + lo (|x - hi - lo| <= 2^-18 |x|), and the three significant
// products are obtained from ONE bf16 contraction over a 3x longer K by concatenation:
//
//     [a_hi | a_lo | a_hi] . [b_hi | b_hi | b_lo]  =  a_hi.b_hi + a_lo.b_hi + a_hi.b_lo  ~  a.b
//
// Measured on B200 (profiles/r02/exp_split_precision.json): max score error 1.2e-6 for unit
// vectors at d = 1024 (bf16: 3.5e-4, cuBLAS fp32: 2e-7).  The tensor core adds into its fp32
// accumulator with truncation (mean relative error -5e-9 per accumulated term on all-positive
// data), so the second contraction, whose K is the bank length, is cut into K chunks of at most
// 32,768 terms whose partial sums are added in fp32 by memproj_reduce_kernel.
//
// The kernels here are the HBM-bound glue: operand splitting (once per bank / per call), the
// row softmax between the two contractions, and the final reduction + L2 normalisation.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "aux_kernels.cuh"

namespace zs {

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// Segment order of the 3-way concatenation: the "A side" of a contraction carries
// [hi | lo | hi], the "B side" [hi | hi | lo].
enum : int { SPLIT_A_SIDE = 0, SPLIT_B_SIDE = 1 };

// out[row, seg * cols + c] for row-major fp32 in[rows_in, cols]; rows [rows_in, rows_out) are
// zero (query-side TMA boxes never leave the tensor).  One warp per output row.
__global__ void __launch_bounds__(256)
split_rows_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int64_t rows_in,
                  int64_t rows_out, int cols, int side) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows_out) return;
  __nv_bfloat16* dst = out + row * 3 * cols;
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.0f);
  for (int c = lane; c < cols; c += 32) {
    __nv_bfloat16 hi = zero, lo = zero;
    if (row < rows_in) split_bf16(in[row * cols + c], hi, lo);
    dst[c] = hi;
    dst[cols + c] = (side == SPLIT_A_SIDE) ? lo : hi;
    dst[2 * cols + c] = (side == SPLIT_A_SIDE) ? hi : lo;
  }
}

// The bank transposed and split for the second contraction (B side: [hi | hi | lo] along K = bank
// rows): out[c, seg * n_pad + j] = split(in[j, c]); columns [n_rows, n_pad) of every segment are
// zero.  32 x 32 tiles through shared memory, coalesced on both sides.
__global__ void __launch_bounds__(256)
transpose_split_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int64_t n_rows,
                       int64_t n_pad, int d) {
  __shared__ float tile[32][33];
  const int64_t j0 = static_cast<int64_t>(blockIdx.x) * 32;   // bank rows of this tile
  const int c0 = blockIdx.y * 32;                              // embedding columns of this tile
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int64_t j = j0 + r;
    tile[r][tx] = (j < n_rows && c0 + tx < d) ? in[j * d + c0 + tx] : 0.0f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r;
    const int64_t j = j0 + tx;
    if (c < d && j < n_pad) {
      __nv_bfloat16 hi, lo;
      split_bf16(tile[tx][r], hi, lo);
      __nv_bfloat16* dst = out + static_cast<int64_t>(c) * 3 * n_pad + j;
      dst[0] = hi;
      dst[n_pad] = hi;
      dst[2 * n_pad] = lo;
    }
  }
}

__device__ __forceinline__ float block_reduce_max(float v, float* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  v = (lane < n_warps) ? scratch[lane] : -CUDART_INF_F;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float block_reduce_sum(float v, float* scratch) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  v = (lane < n_warps) ? scratch[lane] : 0.0f;
  return warp_sum(v);
}

// One block per query row: p = softmax(t * s) over the bank, written as the A side of the second
// contraction, [hi | lo | hi] along K = bank rows (pads zeroed).  The row (4 N bytes) stays in L2
// between the three passes.
constexpr int SOFTMAX_THREADS = 1024;
__global__ void __launch_bounds__(SOFTMAX_THREADS)
softmax_split_kernel(const float* __restrict__ scores, __nv_bfloat16* __restrict__ out, int64_t n_rows,
                     int64_t n_pad, float temperature) {
  __shared__ float scratch[32];
  const int64_t q = blockIdx.x;
  const float* s = scores + q * n_rows;
  float m = -CUDART_INF_F;
  for (int64_t j = threadIdx.x; j < n_rows; j += SOFTMAX_THREADS) m = fmaxf(m, s[j]);
  m = block_reduce_max(m, scratch) * temperature;
  float z = 0.0f;
  for (int64_t j = threadIdx.x; j < n_rows; j += SOFTMAX_THREADS) z += expf(s[j] * temperature - m);
  z = block_reduce_sum(z, scratch);
  const float inv_z = 1.0f / z;
  __nv_bfloat16* dst = out + q * 3 * n_pad;
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.0f);
  for (int64_t j = threadIdx.x; j < n_pad; j += SOFTMAX_THREADS) {
    __nv_bfloat16 hi = zero, lo = zero;
    if (j < n_rows) split_bf16(expf(s[j] * temperature - m) * inv_z, hi, lo);
    dst[j] = hi;
    dst[n_pad + j] = lo;
    dst[2 * n_pad + j] = hi;
  }
}

// out[q, :] = (sum over K chunks of partial[kc, q, :]) / its L2 norm  (reference :28, no epsilon:
// a zero vector stays zero here instead of becoming NaN).  One block per query.
__global__ void __launch_bounds__(256)
memproj_reduce_kernel(const float* __restrict__ partial, int k_chunks, int64_t n_queries, int d,
                      float* __restrict__ out) {
  __shared__ float scratch[32];
  const int64_t q = blockIdx.x;
  float ss = 0.0f;
  for (int c = threadIdx.x; c < d; c += 256) {
    float v = 0.0f;
    for (int kc = 0; kc < k_chunks; ++kc) v += partial[(static_cast<int64_t>(kc) * n_queries + q) * d + c];
    out[q * d + c] = v;
    ss = fmaf(v, v, ss);
  }
  ss = block_reduce_sum(ss, scratch);
  const float norm = sqrtf(ss);
  if (norm > 0.0f)
    for (int c = threadIdx.x; c < d; c += 256) out[q * d + c] = out[q * d + c] / norm;
}

}  // namespace zs
