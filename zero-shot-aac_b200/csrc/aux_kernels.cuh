// aux_kernels.cuh — the small HBM-bound kernels around the fused similarity/top-k kernel:
//   normalize_cast_kernel : F.normalize(x, dim=-1) (eps 1e-12) + cast to bf16 (or fp32 out)
//                           (reference embeddings_related_generator.py:17 for the bank, :21 per query)
//   merge_lists_kernel    : k-way merge of sorted (score, index) lists, order (score desc, index asc)
//   gather_rows_kernel    : out[i] = src[idx[i]]      (reference embeddings_related_generator.py:23)
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "ptx_sm100.cuh"

namespace zs {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One warp per row.  d is a multiple of 4 (64 on the search path), so every lane handles whole
// 8-byte/16-byte vectors.
// The sum of squares is accumulated in a fixed order (lane-strided, then xor butterfly), so the
// result does not depend on the launch geometry (nor on which kernel calls this).
template <typename InT>
__device__ __forceinline__ void load4(const InT* __restrict__ p, float& x0, float& x1, float& x2, float& x3) {
  if constexpr (sizeof(InT) == 4) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    x0 = v.x; x1 = v.y; x2 = v.z; x3 = v.w;
  } else {
    const uint2 raw = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
    x0 = __low2float(a); x1 = __high2float(a); x2 = __low2float(b); x3 = __high2float(b);
  }
}

template <typename OutT>
__device__ __forceinline__ void store4(OutT* __restrict__ p, float y0, float y1, float y2, float y3) {
  if constexpr (sizeof(OutT) == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(y0, y1, y2, y3);
  } else {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(y0, y1);
    const __nv_bfloat162 hi = __floats2bfloat162_rn(y2, y3);
    uint2 packed;
    packed.x = *reinterpret_cast<const uint32_t*>(&lo);
    packed.y = *reinterpret_cast<const uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(p) = packed;
  }
}

// Rows up to 1024 elements (the path's d) are held in registers between the sum of squares and
// the scaling, so the row is read once and all its loads are in flight together (one memory
// round trip instead of two); longer rows take two passes.  Same summation order either way.
template <typename InT, typename OutT>
__device__ __forceinline__ void normalize_cast_row(const InT* __restrict__ src, OutT* __restrict__ dst,
                                                   int d, int normalize, int lane) {
  if (d <= 1024) {
    float x[8][4];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int c = lane * 4 + t * 128;
      if (c < d) load4(src + c, x[t][0], x[t][1], x[t][2], x[t][3]);
      else x[t][0] = x[t][1] = x[t][2] = x[t][3] = 0.0f;
    }
    float denom = 1.0f;   // F.normalize divides by max(||x||, eps); x / 1 is exact when not normalising
    if (normalize) {
      float ss = 0.0f;
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        ss = fmaf(x[t][0], x[t][0], ss); ss = fmaf(x[t][1], x[t][1], ss);
        ss = fmaf(x[t][2], x[t][2], ss); ss = fmaf(x[t][3], x[t][3], ss);
      }
      ss = warp_sum(ss);
      denom = fmaxf(sqrtf(ss), 1e-12f);
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int c = lane * 4 + t * 128;
      if (c < d) store4(dst + c, x[t][0] / denom, x[t][1] / denom, x[t][2] / denom, x[t][3] / denom);
    }
    return;
  }
  float denom = 1.0f;
  if (normalize) {
    float ss = 0.0f;
    for (int c = lane * 4; c < d; c += 128) {
      float x0, x1, x2, x3;
      load4(src + c, x0, x1, x2, x3);
      ss = fmaf(x0, x0, ss); ss = fmaf(x1, x1, ss); ss = fmaf(x2, x2, ss); ss = fmaf(x3, x3, ss);
    }
    ss = warp_sum(ss);
    denom = fmaxf(sqrtf(ss), 1e-12f);
  }
  for (int c = lane * 4; c < d; c += 128) {
    float x0, x1, x2, x3;
    load4(src + c, x0, x1, x2, x3);
    store4(dst + c, x0 / denom, x1 / denom, x2 / denom, x3 / denom);
  }
}

// Zero row (query workspace padding: the consumer's TMA boxes never leave the tensor).
template <typename OutT>
__device__ __forceinline__ void zero_row(OutT* __restrict__ dst, int d, int lane) {
  for (int c = lane * 4; c < d; c += 128) {
    if constexpr (sizeof(OutT) == 4) *reinterpret_cast<float4*>(dst + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    else *reinterpret_cast<uint2*>(dst + c) = make_uint2(0u, 0u);
  }
}

template <typename InT, typename OutT>
__global__ void __launch_bounds__(256)
normalize_cast_kernel(const InT* __restrict__ in, OutT* __restrict__ out, int64_t n_rows,
                      int64_t n_rows_out, int d, int normalize) {
  // the consumer (fused search kernel) may begin its prologue now; it waits for this grid to
  // finish before it reads `out`
  ptx::pdl_launch_dependents();
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_rows_out) return;
  if (row >= n_rows) { zero_row(out + row * d, d, lane); return; }   // rows [n_rows, n_rows_out)
  normalize_cast_row(in + row * d, out + row * d, d, normalize, lane);
}

// ---------------------------------------------------------------------------------------------
// Merge: one warp per query; lane l owns lists l, l+32, ... (LPL = lists per lane, a template
// parameter so that the common small merges keep a small register footprint).
constexpr int MERGE_MAX_LISTS_PER_LANE = 16;
constexpr int MERGE_MAX_LISTS = 32 * MERGE_MAX_LISTS_PER_LANE;

struct Cand {
  float s;
  long long i;
  int src;  // lane * LPL + slot : makes the order strict even for duplicates
};

// int32 lists (chunk partials of the fused kernel) mark empty slots with 0x7fffffff
template <typename IdxT>
__device__ __forceinline__ long long widen_index(IdxT v) {
  if (sizeof(IdxT) == 4 && static_cast<long long>(v) == 0x7fffffffll) return 0x7fffffffffffffffll;
  return static_cast<long long>(v);
}

__device__ __forceinline__ bool cand_better(const Cand& a, const Cand& b) {
  if (a.s != b.s) return a.s > b.s;
  if (a.i != b.i) return a.i < b.i;
  return a.src < b.src;
}

// k-way merge of `n_lists` sorted lists by one warp: lane l owns lists l, l+32, ...; k rounds of a
// warp arg-max over the list heads (xor butterfly on a strict total order, so every lane agrees).
// load(list, pos, s, i) reads element `pos` of a list, store(r, s, i) receives result r (lane 0).
// The successor of every head is fetched one round ahead, so a round never waits on a dependent
// load.
template <int LPL, typename Load, typename Store>
__device__ __forceinline__ void warp_merge(int n_lists, int k, int lane, Load load, Store store) {
  const long long SENT = 0x7fffffffffffffffll;
  int head[LPL];
  float hs[LPL], ns[LPL];
  long long hi[LPL], ni[LPL];
#pragma unroll
  for (int l = 0; l < LPL; ++l) {
    const int list = lane + 32 * l;
    head[l] = 0;
    hs[l] = ns[l] = -CUDART_INF_F;
    hi[l] = ni[l] = SENT;
    if (list < n_lists) {
      load(list, 0, hs[l], hi[l]);
      if (k > 1) load(list, 1, ns[l], ni[l]);
    } else {
      head[l] = k;  // exhausted
    }
  }
  for (int r = 0; r < k; ++r) {
    Cand best{-CUDART_INF_F, SENT, 0x7fffffff};
#pragma unroll
    for (int l = 0; l < LPL; ++l) {
      if (head[l] < k) {
        const Cand c{hs[l], hi[l], lane * LPL + l};
        if (cand_better(c, best)) best = c;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      Cand other;
      other.s = __shfl_xor_sync(0xffffffffu, best.s, o);
      other.i = __shfl_xor_sync(0xffffffffu, best.i, o);
      other.src = __shfl_xor_sync(0xffffffffu, best.src, o);
      if (cand_better(other, best)) best = other;
    }
    if (lane == 0) store(r, best.s, best.i);
    // the owner of the winning list advances it: the prefetched successor becomes the head and
    // the element after that is requested
#pragma unroll
    for (int l = 0; l < LPL; ++l) {
      if (best.src == lane * LPL + l) {
        ++head[l];
        hs[l] = ns[l];
        hi[l] = ni[l];
        if (head[l] + 1 < k) load(lane + 32 * l, head[l] + 1, ns[l], ni[l]);
      }
    }
  }
}

// The same merge on PACKED keys, for lists whose indices fit 32 bits (the chunk partials of the
// fused kernel: columns within a shard).  key = (order-preserving score bits << 32) | ~column, so
// "better under (score desc, column asc)" is one unsigned 64-bit compare and a pick is a 64-bit
// max: ~30 dependent instructions per round instead of ~120 — what counts when ONE warp merges
// one query's lists while nothing else is left to run (small searches, single-launch mode).
// Exhausted lists present key 0; empty slots (-inf, IDX 0x7fffffff) present a small non-zero key.
__device__ __forceinline__ unsigned long long pack_key(float s, int col) {
  unsigned int u = __float_as_uint(s);
  if (u == 0x80000000u) u = 0u;                 // -0 == +0 in the float order the lists were built with
  const unsigned int sk = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return (static_cast<unsigned long long>(sk) << 32) | (0xffffffffu - static_cast<unsigned int>(col));
}
__device__ __forceinline__ void unpack_key(unsigned long long key, float& s, long long& col) {
  const unsigned int sk = static_cast<unsigned int>(key >> 32);
  s = __uint_as_float((sk & 0x80000000u) ? (sk & 0x7fffffffu) : ~sk);
  const unsigned int c = 0xffffffffu - static_cast<unsigned int>(key);
  col = (c == 0x7fffffffu) ? 0x7fffffffffffffffll : static_cast<long long>(c);
}

// load_key(list, pos) -> packed key; store(r, score, column-or-SENT)
template <int LPL, typename LoadKey, typename Store>
__device__ __forceinline__ void warp_merge_keys(int n_lists, int k, int lane, LoadKey load_key, Store store) {
  int head[LPL];
  unsigned long long hk[LPL], nk[LPL];
#pragma unroll
  for (int l = 0; l < LPL; ++l) {
    const int list = lane + 32 * l;
    head[l] = 0;
    hk[l] = nk[l] = 0ull;
    if (list < n_lists) {
      hk[l] = load_key(list, 0);
      if (k > 1) nk[l] = load_key(list, 1);
    } else {
      head[l] = k;
    }
  }
  for (int r = 0; r < k; ++r) {
    unsigned long long best = 0ull;
#pragma unroll
    for (int l = 0; l < LPL; ++l) best = max(best, hk[l]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) {
      float s = -CUDART_INF_F;
      long long col = 0x7fffffffffffffffll;
      if (best != 0ull) unpack_key(best, s, col);
      store(r, s, col);
    }
#pragma unroll
    for (int l = 0; l < LPL; ++l) {
      if (best != 0ull && hk[l] == best) {       // the owner(s) of the pick advance
        ++head[l];
        hk[l] = (head[l] < k) ? nk[l] : 0ull;
        nk[l] = (head[l] + 1 < k) ? load_key(lane + 32 * l, head[l] + 1) : 0ull;
      }
    }
  }
}

// One warp per query: enough parallelism whenever there are many queries.
template <typename IdxT, int LPL>
__global__ void __launch_bounds__(256)
merge_lists_kernel(const float* __restrict__ scores, const IdxT* __restrict__ idx, int S,
                   int64_t score_stride, int64_t index_stride, int64_t Q, int k, long long idx_offset,
                   float* __restrict__ out_scores, long long* __restrict__ out_idx, int64_t out_stride) {
  ptx::pdl_wait();   // lists are written by the preceding grid (no-op without the PDL attribute)
  const int64_t q = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q >= Q) return;
  const long long SENT = 0x7fffffffffffffffll;
  auto store = [&](int r, float s, long long i) {
    out_scores[q * out_stride + r] = s;
    out_idx[q * out_stride + r] = (i == SENT) ? -1ll : i + idx_offset;
  };
  if constexpr (sizeof(IdxT) == 4) {
    warp_merge_keys<LPL>(
        S, k, lane,
        [&](int list, int pos) {
          return pack_key(scores[static_cast<int64_t>(list) * score_stride + q * k + pos],
                          static_cast<int>(idx[static_cast<int64_t>(list) * index_stride + q * k + pos]));
        },
        store);
  } else {
    warp_merge<LPL>(
        S, k, lane,
        [&](int list, int pos, float& s, long long& i) {
          s = scores[static_cast<int64_t>(list) * score_stride + q * k + pos];
          i = widen_index(idx[static_cast<int64_t>(list) * index_stride + q * k + pos]);
        },
        store);
  }
}

// One block per query, for few queries with many lists (a small batch against a bank split into
// a chunk per SM: up to 512 lists per query): the 8 warps each merge an eighth of the lists into
// shared memory, warp 0 merges those 8.  ncu: 24 -> 12 us for 128 queries x 286 lists, 21 -> 11 us
// for one query.
constexpr int MERGE_BLOCK_WARPS = 8;
constexpr int MERGE_BLOCK_MAX_K = 32;
template <typename IdxT>
__global__ void __launch_bounds__(32 * MERGE_BLOCK_WARPS)
merge_lists_block_kernel(const float* __restrict__ scores, const IdxT* __restrict__ idx, int S,
                         int64_t score_stride, int64_t index_stride, int64_t Q, int k,
                         long long idx_offset, float* __restrict__ out_scores,
                         long long* __restrict__ out_idx, int64_t out_stride) {
  __shared__ float part_s[MERGE_BLOCK_WARPS][MERGE_BLOCK_MAX_K];
  __shared__ long long part_i[MERGE_BLOCK_WARPS][MERGE_BLOCK_MAX_K];
  ptx::pdl_wait();
  const int64_t q = blockIdx.x;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long SENT = 0x7fffffffffffffffll;
  const int per_warp = (S + MERGE_BLOCK_WARPS - 1) / MERGE_BLOCK_WARPS;   // <= 64: two lists per lane
  const int first = warp * per_warp;
  const int mine = max(0, min(per_warp, S - first));
  auto load_mine = [&](int list, int pos, float& s, long long& i) {
    s = scores[static_cast<int64_t>(first + list) * score_stride + q * k + pos];
    i = widen_index(idx[static_cast<int64_t>(first + list) * index_stride + q * k + pos]);
  };
  auto store_part = [&](int r, float s, long long i) {
    part_s[warp][r] = s;
    part_i[warp][r] = i;
  };
  if constexpr (sizeof(IdxT) == 4) {
    warp_merge_keys<2>(
        mine, k, lane,
        [&](int list, int pos) {
          return pack_key(scores[static_cast<int64_t>(first + list) * score_stride + q * k + pos],
                          static_cast<int>(idx[static_cast<int64_t>(first + list) * index_stride + q * k + pos]));
        },
        store_part);
  } else {
    warp_merge<2>(mine, k, lane, load_mine, store_part);
  }
  __syncthreads();
  if (warp == 0) {
    warp_merge<1>(
        MERGE_BLOCK_WARPS, k, lane,
        [&](int list, int pos, float& s, long long& i) {
          s = part_s[list][pos];
          i = part_i[list][pos];
        },
        [&](int r, float s, long long i) {
          out_scores[q * out_stride + r] = s;
          out_idx[q * out_stride + r] = (i == SENT) ? -1ll : i + idx_offset;
        });
  }
}

// ---------------------------------------------------------------------------------------------
// fp32 re-scoring of the candidates the bf16 search returned (zs_rescore_f32).  The fused kernel
// ranks with bf16-rounded operands (score error <= ~1e-4), so two bank rows closer than that can
// swap places against the reference's fp32 torch.cosine_similarity(...).topk(k)
// (embeddings_related_generator.py:22).  Searching for k + margin candidates and re-scoring them
// from the caller's fp32 bank restores the fp32 ranking wherever the true top-k lies inside the
// candidate set.
//
// One block (4 warps) per query: warps stride over the kc candidates, each computes q.b, q.q and
// b.b of one (query, candidate) pair in fp32 (16-byte loads, fixed summation order: lane-strided,
// then xor butterfly) -> cosine = q.b / (max(|q|, eps) max(|b|, eps)), or the raw q.b with
// normalize == 0; the block then sorts the candidates in shared memory (bitonic network over
// P = next power of two >= kc entries) under (score desc, index asc) and writes the first k.
constexpr int RESCORE_THREADS = 128;
constexpr int RESCORE_MAX_CAND = 2048;
__global__ void __launch_bounds__(RESCORE_THREADS)
rescore_f32_kernel(const float* __restrict__ queries, const float* __restrict__ bank, int64_t n_bank,
                   int d, int normalize, const long long* __restrict__ cand, int kc, int k,
                   long long idx_offset, float* __restrict__ out_scores,
                   long long* __restrict__ out_idx, int P) {
  extern __shared__ __align__(16) unsigned char rescore_smem[];
  long long* s_idx = reinterpret_cast<long long*>(rescore_smem);   // [P]
  float* s_score = reinterpret_cast<float*>(s_idx + P);            // [P]
  const int64_t q = blockIdx.x;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long SENT = 0x7fffffffffffffffll;
  const float* qrow = queries + q * d;
  for (int c = warp; c < kc; c += RESCORE_THREADS / 32) {
    const long long g = cand[q * kc + c];
    const long long col = g - idx_offset;
    float score = -CUDART_INF_F;
    long long keep = SENT;
    if (g >= 0 && col >= 0 && col < n_bank) {          // (-1 marks an empty slot)
      const float* brow = bank + col * d;
      float qb = 0.f, qq = 0.f, bb = 0.f;
      for (int e = lane * 4; e < d; e += 128) {
        const float4 a = *reinterpret_cast<const float4*>(qrow + e);
        const float4 b = *reinterpret_cast<const float4*>(brow + e);
        qb = fmaf(a.x, b.x, qb); qb = fmaf(a.y, b.y, qb); qb = fmaf(a.z, b.z, qb); qb = fmaf(a.w, b.w, qb);
        qq = fmaf(a.x, a.x, qq); qq = fmaf(a.y, a.y, qq); qq = fmaf(a.z, a.z, qq); qq = fmaf(a.w, a.w, qq);
        bb = fmaf(b.x, b.x, bb); bb = fmaf(b.y, b.y, bb); bb = fmaf(b.z, b.z, bb); bb = fmaf(b.w, b.w, bb);
      }
      qb = warp_sum(qb); qq = warp_sum(qq); bb = warp_sum(bb);
      score = normalize ? qb / (fmaxf(sqrtf(qq), 1e-12f) * fmaxf(sqrtf(bb), 1e-12f)) : qb;
      if (score != score) score = -CUDART_INF_F;       // NaN never ranks (as in the search)
      keep = g;
    }
    if (lane == 0) { s_score[c] = score; s_idx[c] = keep; }
  }
  for (int c = kc + static_cast<int>(threadIdx.x); c < P; c += RESCORE_THREADS) {
    s_score[c] = -CUDART_INF_F;
    s_idx[c] = SENT;
  }
  __syncthreads();
  // bitonic sort, "better first" = (score desc, index asc)
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < P / 2; i += RESCORE_THREADS) {
        const int lo = ((i & ~(stride - 1)) << 1) | (i & (stride - 1));
        const int hi = lo | stride;
        const bool up = (lo & size) == 0;
        const float sa = s_score[lo], sb = s_score[hi];
        const long long ia = s_idx[lo], ib = s_idx[hi];
        const bool b_better = (sb > sa) || (sb == sa && ib < ia);
        if (b_better == up) {
          s_score[lo] = sb; s_score[hi] = sa;
          s_idx[lo] = ib; s_idx[hi] = ia;
        }
      }
      __syncthreads();
    }
  }
  for (int r = threadIdx.x; r < k; r += RESCORE_THREADS) {
    out_scores[q * k + r] = s_score[r];
    out_idx[q * k + r] = (s_idx[r] == SENT) ? -1ll : s_idx[r];
  }
}

// ---------------------------------------------------------------------------------------------
// Rank-of-target support (retrieval metrics a2t / t2a, reference retrieval/tools/utils.py:169-251)

// One warp per (query, target): score = <bf16 query row, bf16 bank row> in fp32, i.e. the same
// operands the fused kernel multiplies (accumulation order differs, which only matters for
// candidates within an ulp of the target; the kernel never counts the target against itself).
// Also converts the global target index to a column of this bank (-1 = absent / unused).
__global__ void __launch_bounds__(256)
target_scores_kernel(const __nv_bfloat16* __restrict__ queries, const __nv_bfloat16* __restrict__ bank,
                     const long long* __restrict__ target_index, int64_t n_pairs, int n_targets,
                     int d, long long index_offset, int64_t n_bank, float* __restrict__ out_scores,
                     int* __restrict__ out_cols) {
  const int64_t pair = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (pair >= n_pairs) return;
  const int64_t q = pair / n_targets;
  const long long g = target_index[pair];
  const long long col = g - index_offset;
  if (g < 0 || col < 0 || col >= n_bank) {
    if (lane == 0) { out_scores[pair] = CUDART_INF_F; out_cols[pair] = -1; }
    return;
  }
  const __nv_bfloat16* a = queries + q * d;
  const __nv_bfloat16* b = bank + col * d;
  float acc = 0.0f;
  for (int c = lane * 4; c < d; c += 128) {
    const uint2 ra = *reinterpret_cast<const uint2*>(a + c);
    const uint2 rb = *reinterpret_cast<const uint2*>(b + c);
    const __nv_bfloat162 a0 = *reinterpret_cast<const __nv_bfloat162*>(&ra.x);
    const __nv_bfloat162 a1 = *reinterpret_cast<const __nv_bfloat162*>(&ra.y);
    const __nv_bfloat162 b0 = *reinterpret_cast<const __nv_bfloat162*>(&rb.x);
    const __nv_bfloat162 b1 = *reinterpret_cast<const __nv_bfloat162*>(&rb.y);
    acc = fmaf(__low2float(a0), __low2float(b0), acc);
    acc = fmaf(__high2float(a0), __high2float(b0), acc);
    acc = fmaf(__low2float(a1), __low2float(b1), acc);
    acc = fmaf(__high2float(a1), __high2float(b1), acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) { out_scores[pair] = acc; out_cols[pair] = static_cast<int>(col); }
}

// ranks[i] = sum over the partial lists of counts[list, i]; absent targets get -1.
__global__ void __launch_bounds__(256)
sum_counts_kernel(const int* __restrict__ part_counts, int n_lists, int64_t n_pairs,
                  const int* __restrict__ cols, long long* __restrict__ out_ranks) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_pairs) return;
  if (cols[i] < 0) { out_ranks[i] = -1; return; }
  long long total = 0;
  for (int l = 0; l < n_lists; ++l) total += part_counts[static_cast<int64_t>(l) * n_pairs + i];
  out_ranks[i] = total;
}

// One warp per output row, 16-byte vectors.  Out-of-range indices produce a zero row.
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ src, int64_t n_src_rows, int d,
                   const long long* __restrict__ idx, int64_t n_idx, float* __restrict__ out) {
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n_idx) return;
  const long long r = idx[i];
  float4* dst = reinterpret_cast<float4*>(out + i * d);
  if (r < 0 || r >= n_src_rows) {
    for (int c = lane; c < d / 4; c += 32) dst[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const float4* s = reinterpret_cast<const float4*>(src + r * d);
  for (int c = lane; c < d / 4; c += 32) dst[c] = s[c];
}

}  // namespace zs
