// exact_f32_kernels.cuh — fp32 similarity for SMALL banks, where the reference's own arithmetic
// (fp32 matmul) costs microseconds and bf16 tensor-core ranking would only add near-tie noise:
//   * utils.sound_effect_choice           prefix @ label_bank.T -> topk        (utils.py:133-135)
//   * zero-shot classification            audio_emb @ text_embeds.t() -> argmax
//                                         (retrieval/zero_shot_classification.py:97-103)
//   * retrieval metrics a2t / t2a         cos_sim -> argsort -> position of the ground truth
//                                         (retrieval/tools/utils.py:169-251)
// Scores are fp32 FMA dot products with a fixed summation order (lane-strided, then xor
// butterfly), i.e. the same class of arithmetic as torch's fp32 matmul: results agree with the
// reference except where two fp32 scores are within rounding of each other.
//
//   exact_scores_kernel    every warp owns a 4 x 4 (bank rows x queries) register tile and walks
//                          the embedding dimension with 16-byte loads served by L1; writes the
//                          [Q, N] score matrix and, for small batches, lets the LAST block to
//                          finish select the top-k of every query in the same launch
//   exact_topk_kernel      top-k of every row of a score matrix (one warp per query)
//   exact_rank_kernel      position of given columns in every row's ordering (one warp per
//                          (query, target)), order (score desc, index asc) like the search
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "aux_kernels.cuh"

namespace zs {

constexpr int EXACT_THREADS = 256;
constexpr int EXACT_QB = 4;                         // queries per register tile
constexpr int EXACT_SMALL_BANK = 8192;              // up to here: one bank row per warp (latency), else four
constexpr int EXACT_FUSED_MAX_Q = 64;               // single-launch top-k up to this many queries

// (s, i) ranks before (ts, ti) in the result order (score desc, index asc)
__device__ __forceinline__ bool ranks_before(float s, long long i, float ts, long long ti) {
  return (s > ts) || (s == ts && i < ti);
}

// Top-k of one score row by one warp: k rounds; round r finds the best element that comes
// strictly after the element chosen in round r-1, so no "taken" flags are needed.  `load(c)`
// returns score c (from shared memory when the row has been staged there, else through L2: the
// scores may have been written by other blocks of the same launch); loads are issued in
// independent batches of 8 so that a round costs one memory round trip per 256 columns.
template <typename Load>
__device__ __forceinline__ void warp_topk_row(Load load, int64_t n, int k, long long self_col,
                                              long long index_offset, float* out_s, long long* out_i,
                                              int lane) {
  const long long SENT = 0x7fffffffffffffffll;
  float last_s = CUDART_INF_F;
  long long last_i = -1;
  for (int r = 0; r < k; ++r) {
    float best_s = -CUDART_INF_F;
    long long best_i = SENT;
    for (int64_t c0 = lane; c0 < n; c0 += 32 * 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int64_t c = c0 + 32 * u;
        v[u] = (c < n) ? load(c) : CUDART_NAN_F;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int64_t c = c0 + 32 * u;
        const float s = v[u];
        const bool usable = (s == s) && c != self_col;                 // NaN never ranks
        const bool after_last = (s < last_s) || (s == last_s && c > last_i);
        if (usable && after_last && ranks_before(s, c, best_s, best_i)) { best_s = s; best_i = c; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, best_s, o);
      const long long oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ranks_before(os, oi, best_s, best_i)) { best_s = os; best_i = oi; }
    }
    if (lane == 0) {
      out_s[r] = best_s;
      out_i[r] = (best_i == SENT) ? -1ll : best_i + index_offset;
    }
    if (best_i == SENT) {            // fewer than k rankable elements: the rest stays empty
      for (int rr = r + 1; rr < k; ++rr)
        if (lane == 0) { out_s[rr] = -CUDART_INF_F; out_i[rr] = -1ll; }
      return;
    }
    last_s = best_s;
    last_i = best_i;
  }
}

constexpr int EXACT_STAGE_COLS = 1024;      // rows up to this long are staged in shared memory (per warp)

// RB bank rows per warp: 4 for throughput on large problems (the rows are re-read from L1 for
// every block of 4 queries), 1 for the small banks of the latency-bound callers (label bank, class
// prompts: more CTAs, and the warp's row stays in registers).  D1024: d <= 1024, the embedding
// loop is fully unrolled so that all loads of a query block are in flight together.
template <int RB, bool D1024>
__global__ void __launch_bounds__(EXACT_THREADS)
exact_scores_kernel(const float* __restrict__ queries, int Q, const float* __restrict__ bank,
                    int64_t n_bank, int d, int normalize, float* scores,
                    // fused top-k (k > 0): the last block to finish selects for every query
                    unsigned int* done_counter, int k, const long long* __restrict__ self_index,
                    long long index_offset, float* out_scores, long long* out_idx) {
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t row0 = (static_cast<int64_t>(blockIdx.x) * (EXACT_THREADS / 32) + warp) * RB;
  constexpr int STEPS = D1024 ? 8 : 1;          // 128-element steps held in registers at once
  if (row0 < n_bank) {
    const float* brow[RB];
    float norm_b[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r)
      brow[r] = bank + min(row0 + r, n_bank - 1) * d;      // clamped: the duplicates are not stored
    float4 breg[RB][STEPS];
    if constexpr (D1024) {
#pragma unroll
      for (int t = 0; t < STEPS; ++t) {
        const int e = lane * 4 + t * 128;
#pragma unroll
        for (int r = 0; r < RB; ++r)
          breg[r][t] = (e < d) ? *reinterpret_cast<const float4*>(brow[r] + e) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    if (normalize) {
      float bb[RB] = {};
      if constexpr (D1024) {
#pragma unroll
        for (int t = 0; t < STEPS; ++t) {
#pragma unroll
          for (int r = 0; r < RB; ++r) {
            const float4 b = breg[r][t];
            bb[r] = fmaf(b.x, b.x, bb[r]); bb[r] = fmaf(b.y, b.y, bb[r]);
            bb[r] = fmaf(b.z, b.z, bb[r]); bb[r] = fmaf(b.w, b.w, bb[r]);
          }
        }
      } else {
        for (int e = lane * 4; e < d; e += 128) {
#pragma unroll
          for (int r = 0; r < RB; ++r) {
            const float4 b = *reinterpret_cast<const float4*>(brow[r] + e);
            bb[r] = fmaf(b.x, b.x, bb[r]); bb[r] = fmaf(b.y, b.y, bb[r]);
            bb[r] = fmaf(b.z, b.z, bb[r]); bb[r] = fmaf(b.w, b.w, bb[r]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < RB; ++r) norm_b[r] = fmaxf(sqrtf(warp_sum(bb[r])), 1e-12f);
    }
    for (int q0 = 0; q0 < Q; q0 += EXACT_QB) {
      const float* qrow[EXACT_QB];
#pragma unroll
      for (int j = 0; j < EXACT_QB; ++j) qrow[j] = queries + static_cast<int64_t>(min(q0 + j, Q - 1)) * d;
      float acc[EXACT_QB][RB] = {};
      float qq[EXACT_QB] = {};
      auto step = [&](const float4 (&b)[RB], const float4 (&a)[EXACT_QB]) {
#pragma unroll
        for (int j = 0; j < EXACT_QB; ++j) {
          if (normalize) {
            qq[j] = fmaf(a[j].x, a[j].x, qq[j]); qq[j] = fmaf(a[j].y, a[j].y, qq[j]);
            qq[j] = fmaf(a[j].z, a[j].z, qq[j]); qq[j] = fmaf(a[j].w, a[j].w, qq[j]);
          }
#pragma unroll
          for (int r = 0; r < RB; ++r) {
            acc[j][r] = fmaf(a[j].x, b[r].x, acc[j][r]); acc[j][r] = fmaf(a[j].y, b[r].y, acc[j][r]);
            acc[j][r] = fmaf(a[j].z, b[r].z, acc[j][r]); acc[j][r] = fmaf(a[j].w, b[r].w, acc[j][r]);
          }
        }
      };
      if constexpr (D1024) {
        float4 a[STEPS][EXACT_QB];
#pragma unroll
        for (int t = 0; t < STEPS; ++t) {
          const int e = lane * 4 + t * 128;
#pragma unroll
          for (int j = 0; j < EXACT_QB; ++j)
            a[t][j] = (e < d) ? *reinterpret_cast<const float4*>(qrow[j] + e) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int t = 0; t < STEPS; ++t) {
          float4 b[RB];
#pragma unroll
          for (int r = 0; r < RB; ++r) b[r] = breg[r][t];
          step(b, a[t]);
        }
      } else {
        for (int e = lane * 4; e < d; e += 128) {
          float4 b[RB], a[EXACT_QB];
#pragma unroll
          for (int r = 0; r < RB; ++r) b[r] = *reinterpret_cast<const float4*>(brow[r] + e);
#pragma unroll
          for (int j = 0; j < EXACT_QB; ++j) a[j] = *reinterpret_cast<const float4*>(qrow[j] + e);
          step(b, a);
        }
      }
#pragma unroll
      for (int j = 0; j < EXACT_QB; ++j) {
        float norm_q = 1.0f;
        if (normalize) norm_q = fmaxf(sqrtf(warp_sum(qq[j])), 1e-12f);
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          float s = warp_sum(acc[j][r]);
          if (normalize) s = s / (norm_q * norm_b[r]);
          if (lane == 0 && q0 + j < Q && row0 + r < n_bank)
            scores[static_cast<int64_t>(q0 + j) * n_bank + row0 + r] = s;
        }
      }
    }
  }
  if (k <= 0) return;
  // ---- fused selection: classic "last block done" (writers fence, one thread counts)
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int ticket = atomicAdd(done_counter, 1u);
    is_last = (ticket == gridDim.x - 1);
    if (is_last) *done_counter = 0u;          // ready for the next launch on this stream
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  __shared__ float staged[EXACT_THREADS / 32][EXACT_STAGE_COLS];
  for (int q = warp; q < Q; q += EXACT_THREADS / 32) {
    long long self_col = -1;
    if (self_index != nullptr) {
      const long long g = self_index[q];
      if (g >= 0) self_col = g - index_offset;
    }
    const float* row = scores + static_cast<int64_t>(q) * n_bank;
    float* out_s = out_scores + static_cast<int64_t>(q) * k;
    long long* out_i = out_idx + static_cast<int64_t>(q) * k;
    if (n_bank <= EXACT_STAGE_COLS) {
      float* mine = staged[warp];
      for (int c = lane; c < n_bank; c += 32) mine[c] = __ldcg(row + c);   // one round trip, all loads in flight
      __syncwarp();
      warp_topk_row([&](int64_t c) { return mine[c]; }, n_bank, k, self_col, index_offset, out_s, out_i, lane);
      __syncwarp();
    } else {
      warp_topk_row([&](int64_t c) { return __ldcg(row + c); }, n_bank, k, self_col, index_offset, out_s, out_i, lane);
    }
  }
}

// One warp per query (many queries: the selection is its own launch).
__global__ void __launch_bounds__(256)
exact_topk_kernel(const float* scores, int Q, int64_t n_bank, int k,
                  const long long* __restrict__ self_index, long long index_offset,
                  float* out_scores, long long* out_idx) {
  const int q = static_cast<int>((static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= Q) return;
  long long self_col = -1;
  if (self_index != nullptr) {
    const long long g = self_index[q];
    if (g >= 0) self_col = g - index_offset;
  }
  const float* row = scores + static_cast<int64_t>(q) * n_bank;
  warp_topk_row([&](int64_t c) { return row[c]; }, n_bank, k, self_col, index_offset,
                out_scores + static_cast<int64_t>(q) * k, out_idx + static_cast<int64_t>(q) * k, lane);
}

// One warp per (query, target): position of the target in the query's ordering = number of bank
// rows that rank before it under (score desc, index asc); -1 for unused / absent targets.
__global__ void __launch_bounds__(256)
exact_rank_kernel(const float* scores, int Q, int64_t n_bank, const long long* __restrict__ target_index,
                  int n_targets, long long index_offset, float* out_target_scores,
                  long long* out_ranks) {
  const int64_t pair = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (pair >= static_cast<int64_t>(Q) * n_targets) return;
  const int64_t q = pair / n_targets;
  const long long g = target_index[pair];
  const long long col = g - index_offset;
  if (g < 0 || col < 0 || col >= n_bank) {
    if (lane == 0) {
      out_ranks[pair] = -1;
      if (out_target_scores) out_target_scores[pair] = CUDART_INF_F;
    }
    return;
  }
  const float* row = scores + q * n_bank;
  const float ts = row[col];
  long long count = 0;
  for (int64_t c = lane; c < n_bank; c += 32) {
    const float s = row[c];
    if (s != s) continue;
    count += ranks_before(s, c, ts, col) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
  if (lane == 0) {
    out_ranks[pair] = count;
    if (out_target_scores) out_target_scores[pair] = ts;
  }
}

}  // namespace zs
