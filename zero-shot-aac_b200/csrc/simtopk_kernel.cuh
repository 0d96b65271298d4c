// simtopk_kernel.cuh — the hot kernel: fused cosine-similarity GEMM + running per-query top-k.
//
// Replaces, for a whole batch of queries at once, what the reference does per item with
//   torch.cosine_similarity(q, bank).topk(k)            (embeddings_related_generator.py:22)
//   prefix @ bank.T -> softmax -> topk                   (utils.py:133-135)
// The [Q, N] similarity matrix lives only in tensor memory (TMEM), 128 x 256 fp32 at a time.
//
// Shape of the contraction:  M = query rows, N = bank rows, K = embedding dim (bf16, fp32 acc).
//   * warp 0   : TMA producer  — streams 128x64 query blocks and 256x64 bank blocks (128-byte
//                swizzle) through a STAGES-deep shared-memory ring
//   * warp 1   : MMA issuer    — issues tcgen05.mma (M=128*CG, N=256, K=16) into one of two
//                256-column TMEM accumulators
//                (both walk their loops with the whole warp, warp-uniformly, and elect one lane
//                per issue: the operands then live in uniform registers)
//   * warp 2   : TMEM allocator
//   * warps 4-11: epilogue     — two warps per TMEM lane quarter; thread t of warp w owns query row
//                32*(w%4)+t and the 128-column half (w-4)/4 of every bank tile: it reads its
//                scores with tcgen05.ld and maintains a sorted top-k list in registers; the
//                epilogue of bank tile j overlaps the MMAs of tile j+1 (double-buffered TMEM)
// A work unit is (query tile, bank chunk); units are walked persistently with a static stride.
// Every unit writes, per row and column half, its k best (score, column) pairs;
// merge_lists_kernel reduces the 2 x chunks lists.  CG == 2 pairs two CTAs of a cluster on a
// 256-row query tile (cta_group::2): each CTA loads its own 128 query rows and half of the bank
// tile, the leader issues the MMAs for both.
#pragma once

#include <math_constants.h>

#include "aux_kernels.cuh"
#include "ptx_sm100.cuh"

namespace zs {

constexpr int BLOCK_M = 128;   // query rows per CTA
constexpr int BLOCK_N = 256;   // bank rows per accumulator tile
constexpr int BLOCK_K = 64;    // bf16 elements per 128-byte swizzled row
constexpr int UMMA_K = 16;
// smem ring depth: a slot holds one K-step of operands, 48 KiB for a single CTA (16 KiB queries +
// 32 KiB bank) and 32 KiB per CTA of a pair (16 + 16), so a pair can run 6 stages deep
#ifndef ZS_PAIR_STAGES
#define ZS_PAIR_STAGES 6
#endif
template <int CG>
constexpr int num_stages() { return CG == 2 ? ZS_PAIR_STAGES : 4; }
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = ACC_STAGES * BLOCK_N;  // 512: the whole tensor memory of the SM
constexpr int EPI_WARP0 = 4;
constexpr int NUM_EPI_WARPS = 8;                 // 2 per TMEM lane quarter (column halves)
constexpr int EPI_HALVES = NUM_EPI_WARPS / 4;
constexpr int EPI_COLS = BLOCK_N / EPI_HALVES;   // columns of a tile each epilogue thread scans
constexpr int NUM_THREADS = 32 * (EPI_WARP0 + NUM_EPI_WARPS);  // 384

constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KiB
template <int CG>
constexpr int b_stage_bytes() { return (BLOCK_N / CG) * BLOCK_K * 2; }  // 32 KiB, or 16 KiB per CTA of a pair
template <int CG>
constexpr int stage_bytes() { return A_STAGE_BYTES + b_stage_bytes<CG>(); }
constexpr int BARRIER_BYTES = 256;
template <int CG>
constexpr int smem_bytes() { return num_stages<CG>() * stage_bytes<CG>() + BARRIER_BYTES + 1024; }

constexpr int IDX_SENTINEL = 0x7fffffff;

// error codes written to the device flag by a timed-out wait
enum : int { ERR_PRODUCER = 101, ERR_MMA_FULL = 102, ERR_MMA_TEMPTY = 103, ERR_EPILOGUE = 104,
              ERR_GRID_CAST = 105, ERR_GRID_DONE = 106 };

struct SimTopkParams {
  int Q;                 // query rows
  int n_bank;            // bank rows of this shard
  int num_k_blocks;      // d / 64
  int num_m_tiles;       // ceil(Q / (128 * CG))
  int num_n_tiles;       // ceil(n_bank / 256)
  int tiles_per_chunk;   // bank tiles per work unit
  int num_chunks;
  int k;                 // requested list length (<= KCAP)
  const long long* self_index;  // nullable [Q]: global bank index to skip
  long long index_offset;       // global index of bank row 0 of this shard
  float* part_scores;    // [num_chunks * EPI_HALVES, Q, k]
  int* part_idx;         // [num_chunks * EPI_HALVES, Q, k]  column within the shard
  float* dump;           // DUMP mode: [Q, n_bank] (or [k chunks, Q, n_bank] partial sums with kb_per_unit > 0)
  // DUMP mode as a split-K GEMM (the batched memory projection's second contraction, K = bank
  // rows): with kb_per_unit > 0 a "chunk" of the unit list is (bank tile, K chunk) — chunk =
  // k_chunk * num_n_tiles + n_tile — and the unit accumulates only K blocks
  // [k_chunk * kb_per_unit, +kb_per_unit) of that one tile into dump[k_chunk].
  int kb_per_unit;
  // RANK mode: per (row, target) the target's score and column within the shard (-1 = unused),
  // and the per-(chunk, half) partial counts
  const float* tgt_scores;   // [Q, n_targets]
  const int* tgt_cols;       // [Q, n_targets]
  int n_targets;
  int* part_counts;          // [num_chunks * EPI_HALVES, Q, n_targets]
  int* err_flag;
  unsigned long long* trace;  // nullable: [gridDim.x, 16] globaltimer stamps (zs_debug_trace)
  // Soft lock-step of the bank stream (nullable = off).  All workers walk units of identical
  // length in the same order, so "window w" (sync_window consecutive bank tiles of a unit
  // iteration) covers the same tile positions for everyone.  A worker starts loading window w
  // only after every worker has issued the loads of window w-1, which keeps the co-running
  // workers within ~2 windows of each other: the bank tiles one worker pulled from HBM are still
  // in L2 when the others ask for them.  Purely a performance hint: the wait is bounded and a
  // worker that times out stops waiting (it keeps counting so nobody waits for it).
  unsigned int* sync_cnt;     // [max_iters * windows_per_unit], zeroed by the host per launch
  int sync_window;            // tiles per window
  int windows_per_unit;       // ceil(tiles_per_chunk / sync_window)
  int max_iters;              // ceil(num_units / num_workers)
  // Shared admission threshold per query row (nullable = off), TOPK mode.  Every epilogue thread
  // publishes the k-th score of its list once the list is full (atomicMax on an order-preserving
  // key) and starts each bank tile from the largest value published so far for its row — by the
  // other column half, by concurrently running chunks and, above all, by chunks of the same query
  // tile that ran earlier, so a unit that starts late does not warm its list up from -inf again.
  // Exactness: if some list holds k entries >= t, no element scoring strictly below t can be in
  // the global top-k, so units admit only v > pred(t) (the float just below t: elements EQUAL to
  // t may still win the index tie-break).  Lists then may hold fewer than k entries (the rest
  // stay -inf / IDX_SENTINEL); the merged result is the exact top-k whatever the timing.
  // Entries are 64-bit: (epoch << 32) | key.  Every search (every pass of a k > 32 search) uses a
  // new epoch, so stale entries compare lower than any entry of the running search and are
  // ignored by readers: the array never needs a reset.
  unsigned long long* row_thr;   // [Q padded]
  unsigned int epoch;
  // Bootstrap of the admission threshold (nullable = off).  When all units of a query row start
  // at the same time (small searches: one unit per CTA) nobody has a full list to publish yet.
  // Every list therefore scans its FIRST bank tile twice: pass A only takes the maximum of its
  // scores and publishes it to one of `boot_slots` (<= 32) slots of its row; the k-th largest of
  // the row's slot maxima is a valid admission threshold at once (k distinct bank rows score at
  // least that much), and about as tight as the k-th entry of a list that has seen
  // boot_slots x more columns.  Pass B then scans the tile with that threshold.
  unsigned long long* boot;      // [Q padded, BOOT_SLOTS] epoch-tagged keys
  int boot_slots;                // min(number of lists per row, BOOT_SLOTS)
  // k > 32: the search runs in passes of <= 32; pass p admits only elements strictly AFTER the
  // last element of pass p-1 under (score desc, index asc).  bound_* point at that element for
  // row 0 (inside the caller's output arrays), bound_stride = elements between rows.
  const float* bound_scores;     // nullable
  const long long* bound_idx;    // global indices
  long long bound_stride;
  // Single-launch ("solo") mode for small searches: the kernel casts (and normalises) the queries
  // itself in a distributed prologue and merges the partial lists itself after a grid-wide
  // arrival counter, so a search is ONE launch instead of three.  Needs every CTA of the grid
  // resident at once: at most one CTA per SM is launched (see launch_simtopk).
  int solo;
  const void* q_src;             // raw queries [Q, d] (nullptr: the bf16 workspace is already filled)
  int q_src_bf16;                // dtype of q_src
  int q_normalize;
  __nv_bfloat16* q_ws;           // bf16 query workspace [q_pad, d] the query tensor map points at
  int q_pad;
  int d;
  unsigned long long* grid_cnt;  // [2] monotonic arrival counters: queries cast, lists written
  unsigned long long cast_target, done_target;
  float* out_scores;             // [Q, out_stride]
  long long* out_idx;
  long long out_stride;
};

// Order-preserving float -> uint32 key (unsigned compare == float compare, -0 < +0), so that
// atomicMax works on scores of either sign.  Key 0 (the zero-initialised state) and everything up
// to key(-inf) decode to "no threshold yet".
constexpr unsigned int KEY_NEG_INF = 0x007fffffu;
__device__ __forceinline__ unsigned int score_key(float f) {
  const unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
// The largest float strictly below the score a key encodes (-inf when nothing is published yet).
__device__ __forceinline__ float seed_below(unsigned int key) {
  if (key <= KEY_NEG_INF) return -CUDART_INF_F;
  const unsigned int kk = key - 1u;
  const float f = __uint_as_float((kk & 0x80000000u) ? (kk & 0x7fffffffu) : ~kk);
  // pred(+0) is -0, which still compares equal to +0: step once more, to the negative denormal
  return (f == 0.0f) ? __uint_as_float(0x80000001u) : f;
}
__device__ __forceinline__ float key_to_score(unsigned int key) {
  return __uint_as_float((key & 0x80000000u) ? (key & 0x7fffffffu) : ~key);
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// key of the running search, or 0 ("nothing published") for entries of older searches
__device__ __forceinline__ unsigned int epoch_key(unsigned long long e, unsigned int epoch) {
  return (static_cast<unsigned int>(e >> 32) == epoch) ? static_cast<unsigned int>(e) : 0u;
}
__device__ __forceinline__ unsigned long long make_epoch_key(unsigned int epoch, float f) {
  return (static_cast<unsigned long long>(epoch) << 32) | score_key(f);
}

constexpr long long SYNC_WAIT_LIMIT_CYCLES = 600000;   // ~0.3-0.4 ms: then give up lock-step

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

constexpr int TRACE_SLOTS = 16;
__device__ __forceinline__ void trace_stamp(const SimTopkParams& p, int slot) {
  if (p.trace != nullptr) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[static_cast<size_t>(blockIdx.x) * TRACE_SLOTS + slot] = t;
  }
}

// Sorted (descending score; equal scores keep arrival order = ascending column) list in registers.
// The list always has KCAP physical slots; for a requested length k < KCAP the top KCAP-k slots
// are pinned with +inf so that the live entries are slots [KCAP-k, KCAP) and the admission
// threshold is always the LAST slot — every register index stays a compile-time constant.
template <int KCAP>
struct TopkList {
  float s[KCAP];
  int i[KCAP];

  __device__ __forceinline__ void init(int k) {
#pragma unroll
    for (int j = 0; j < KCAP; ++j) {
      s[j] = (j < KCAP - k) ? CUDART_INF_F : -CUDART_INF_F;
      i[j] = IDX_SENTINEL;
    }
  }

  __device__ __forceinline__ float threshold() const { return s[KCAP - 1]; }

  // Insert (v, idx) with v > threshold(): shifts the tail down by one, dropping the last slot.
  __device__ __forceinline__ void insert(float v, int idx) {
    bool gt[KCAP];
#pragma unroll
    for (int j = 0; j < KCAP; ++j) gt[j] = v > s[j];
#pragma unroll
    for (int j = KCAP - 1; j > 0; --j) {
      s[j] = gt[j - 1] ? s[j - 1] : (gt[j] ? v : s[j]);
      i[j] = gt[j - 1] ? i[j - 1] : (gt[j] ? idx : i[j]);
    }
    s[0] = gt[0] ? v : s[0];
    i[0] = gt[0] ? idx : i[0];
  }
};

// r[j] for a run-time j without local memory: a 5-level select tree (31 SEL).
__device__ __forceinline__ uint32_t select32(const uint32_t (&r)[32], int j) {
  uint32_t a[16], b[8], c[4], d[2];
#pragma unroll
  for (int t = 0; t < 16; ++t) a[t] = (j & 1) ? r[2 * t + 1] : r[2 * t];
#pragma unroll
  for (int t = 0; t < 8; ++t) b[t] = (j & 2) ? a[2 * t + 1] : a[2 * t];
#pragma unroll
  for (int t = 0; t < 4; ++t) c[t] = (j & 4) ? b[2 * t + 1] : b[2 * t];
#pragma unroll
  for (int t = 0; t < 2; ++t) d[t] = (j & 8) ? c[2 * t + 1] : c[2 * t];
  return (j & 16) ? d[1] : d[0];
}

// ---------------------------------------------------------------------------------------------
// Bootstrap of the admission threshold (SimTopkParams::boot)
constexpr int BOOT_SLOTS = 32;
constexpr long long BOOT_WAIT_LIMIT_CYCLES = 6000;    // ~3 us: then take what has been published

// Descending bitonic sort of 32 floats in registers (fully unrolled, every index a constant).
__device__ __forceinline__ void sort32_desc(float (&a)[32]) {
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int l = i ^ stride;
        if (l > i) {
          const bool desc = (i & size) == 0;
          const float lo = fminf(a[i], a[l]);
          const float hi = fmaxf(a[i], a[l]);
          a[i] = desc ? hi : lo;
          a[l] = desc ? lo : hi;
        }
      }
    }
  }
}

// Pass A over the first bank tile of a unit (warp-collective: tcgen05.ld is .sync.aligned).
//   taddr       tensor-memory address of this warp's lane quarter, first column of its half
//   col_first   bank column of that first column
// Publishes the maximum of the thread's EPI_COLS scores (ragged tail and the self column masked)
// to slot `slot` of its row, waits a bounded time for `expect` slots of the row, and returns the
// admission seed: the largest float strictly below the k-th largest slot maximum (-inf while
// fewer than k slots are filled).  Valid because the slot maxima are scores of distinct bank rows.
__device__ __noinline__ float boot_threshold(uint32_t taddr, int col_first, int n_bank, int self_col,
                                             bool row_ok, unsigned long long* slots_row, int slot,
                                             int expect, int k, unsigned int epoch) {
  float m = -CUDART_INF_F;
#pragma unroll 1
  for (int c = 0; c < EPI_COLS; c += 32) {
    uint32_t r[32];
    ptx::tmem_ld_32x32(taddr + c, r);
    ptx::tmem_ld_wait();
    const int col0 = col_first + c;
    const bool clean = (col0 + 32 <= n_bank) && !(self_col >= col0 && self_col < col0 + 32);
    if (clean) {
#pragma unroll
      for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(r[j]));
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < n_bank && col0 + j != self_col) m = fmaxf(m, __uint_as_float(r[j]));
    }
  }
  if (row_ok) atomicMax(slots_row + slot, make_epoch_key(epoch, m));
  float v[32];
  const long long w0 = clock64();
  while (true) {
    int filled = 0;
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      unsigned long long e0, e1;
      asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];"
                   : "=l"(e0), "=l"(e1) : "l"(slots_row + j) : "memory");
      const unsigned int k0 = epoch_key(e0, epoch), k1 = epoch_key(e1, epoch);
      filled += (k0 != 0u) + (k1 != 0u);
      v[j] = k0 ? key_to_score(k0) : -CUDART_INF_F;
      v[j + 1] = k1 ? key_to_score(k1) : -CUDART_INF_F;
    }
    const bool ok = (filled >= expect) || !row_ok || (clock64() - w0 > BOOT_WAIT_LIMIT_CYCLES);
    if (__all_sync(0xffffffffu, ok)) break;
    __nanosleep(64);
  }
  sort32_desc(v);
  uint32_t vb[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) vb[j] = __float_as_uint(v[j]);
  const float kth = __uint_as_float(select32(vb, k - 1));
  return seed_below(score_key(kth));
}

// ---------------------------------------------------------------------------------------------
// Solo mode helpers
// All lanes poll (one coalesced request per try); straight-line asm so that the callers keep
// warp-uniform control flow.  The grid has at most one CTA per SM, so every CTA becomes resident
// and the wait ends; the bound (~a minute) only turns a bug into a trap.
__device__ __forceinline__ void grid_wait(const unsigned long long* cnt, unsigned long long target,
                                          int* err, int code) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q, has_err;\n\t"
      ".reg .u64 v;\n\t"
      ".reg .u32 tries;\n\t"
      "mov.u32 tries, 0;\n\t"
      "ZS_GRID_WAIT_LOOP:\n\t"
      "ld.acquire.gpu.global.u64 v, [%0];\n\t"
      "setp.ge.u64 p, v, %1;\n\t"
      "@p bra ZS_GRID_WAIT_DONE;\n\t"
      "nanosleep.u32 32;\n\t"
      "add.u32 tries, tries, 1;\n\t"
      "setp.lt.u32 q, tries, 0x4000000;\n\t"
      "@q bra ZS_GRID_WAIT_LOOP;\n\t"
      "setp.ne.u64 has_err, %2, 0;\n\t"
      "@has_err st.volatile.global.u32 [%2], %3;\n\t"
      "fence.sc.sys;\n\t"
      "trap;\n\t"
      "ZS_GRID_WAIT_DONE:\n\t"
      "}" ::"l"(cnt), "l"(target), "l"(reinterpret_cast<uint64_t>(err)), "r"(code)
      : "memory");
}

// Rows are dealt round-robin to all warps of the grid.  One-shot code: kept small on purpose — a
// larger variant that staged the keys in shared memory first was SLOWER (profiles/r02/SUMMARY.md:
// this phase is dominated by fetching cold instructions, not by its loads).
__device__ __noinline__ void solo_merge(const float* part_scores, const int* part_idx, int n_lists,
                                        int Q, int k, long long index_offset, float* out_scores,
                                        long long* out_idx, long long out_stride, int gwarp,
                                        int n_gwarps, int lane) {
  const size_t list_stride = static_cast<size_t>(Q) * k;
  const long long SENT = 0x7fffffffffffffffll;
  for (int row = gwarp; row < Q; row += n_gwarps) {
    // the lists were written by other CTAs of this same launch: read them through L2 (.cg),
    // never through the non-coherent path
    auto load_key = [&](int list, int pos) {
      const size_t o = static_cast<size_t>(list) * list_stride + static_cast<size_t>(row) * k + pos;
      return pack_key(__ldcg(part_scores + o), __ldcg(part_idx + o));
    };
    auto store = [&](int r, float sc, long long ix) {
      out_scores[static_cast<size_t>(row) * out_stride + r] = sc;
      out_idx[static_cast<size_t>(row) * out_stride + r] = (ix == SENT) ? -1ll : ix + index_offset;
    };
    if (n_lists <= 64) warp_merge_keys<2>(n_lists, k, lane, load_key, store);
    else if (n_lists <= 256) warp_merge_keys<8>(n_lists, k, lane, load_key, store);
    else warp_merge_keys<16>(n_lists, k, lane, load_key, store);
  }
}

// MODE_TOPK: running top-k (KCAP list slots).  MODE_DUMP: write the score matrix (test hook).
// MODE_RANK: count, per query row and per target (KCAP = max targets per row), the bank rows
// whose score is strictly greater than the target's score — the rank of the ground truth that
// the reference's retrieval metrics obtain from a full argsort (retrieval/tools/utils.py:183,236).
enum : int { MODE_TOPK = 0, MODE_DUMP = 1, MODE_RANK = 2 };

// Tile and K-block ranges of a work unit's chunk (see SimTopkParams::kb_per_unit).
struct UnitRange { int t0, t1, kb0, kb1, k_chunk; };
__device__ __forceinline__ UnitRange unit_range(const SimTopkParams& p, int chunk) {
  UnitRange r;
  if (p.kb_per_unit > 0) {
    r.k_chunk = chunk / p.num_n_tiles;
    r.t0 = chunk - r.k_chunk * p.num_n_tiles;
    r.t1 = r.t0 + 1;
    r.kb0 = r.k_chunk * p.kb_per_unit;
    r.kb1 = min(r.kb0 + p.kb_per_unit, p.num_k_blocks);
  } else {
    r.k_chunk = 0;
    r.t0 = chunk * p.tiles_per_chunk;
    r.t1 = min(r.t0 + p.tiles_per_chunk, p.num_n_tiles);
    r.kb0 = 0;
    r.kb1 = p.num_k_blocks;
  }
  return r;
}

template <int KCAP, int CG, int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
zs_simtopk_kernel(const __grid_constant__ CUtensorMap tmap_q,
                  const __grid_constant__ CUtensorMap tmap_b, const SimTopkParams p) {
  constexpr bool DUMP = (MODE == MODE_DUMP);
  constexpr bool RANK = (MODE == MODE_RANK);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = ptx::smem_u32(smem_raw);
  const uint32_t base_u32 = (raw_u32 + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = smem_raw + (base_u32 - raw_u32);

  constexpr int STAGES = num_stages<CG>();
  constexpr int STAGE_STRIDE = stage_bytes<CG>();
  constexpr int B_BYTES = b_stage_bytes<CG>();
  constexpr uint32_t TX_BYTES = static_cast<uint32_t>(CG) * (A_STAGE_BYTES + B_BYTES);
  static_assert(8 * (2 * STAGES + 2 * ACC_STAGES) + 4 <= BARRIER_BYTES, "barrier area");

  const uint32_t bar_base = base_u32 + STAGES * STAGE_STRIDE;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + ACC_STAGES + a); };
  uint32_t* tmem_slot =
      reinterpret_cast<uint32_t*>(smem + STAGES * STAGE_STRIDE + 8 * (2 * STAGES + 2 * ACC_STAGES));

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = static_cast<int>(threadIdx.x & 31);
  // (shuffles from lane 0 tell the compiler these values are warp-uniform)
  const uint32_t cta_rank = (CG == 2) ? __shfl_sync(0xffffffffu, ptx::cluster_ctarank(), 0) : 0u;
  const bool is_leader = cta_rank == 0;

  if (threadIdx.x == 0) {
    trace_stamp(p, 0);                                   // kernel entry
    if (p.trace != nullptr) p.trace[static_cast<size_t>(blockIdx.x) * TRACE_SLOTS + 6] = clock64();
    ptx::prefetch_tensormap(&tmap_q);
    ptx::prefetch_tensormap(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < ACC_STAGES; ++a) {
      ptx::mbar_init(tfull_bar(a), 1);
      ptx::mbar_init(tempty_bar(a), NUM_EPI_WARPS * CG);   // one arrival per epilogue warp
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc<CG>(ptx::smem_u32(tmem_slot), TMEM_COLS);
  ptx::tc_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync_all(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base =
      __shfl_sync(0xffffffffu, *reinterpret_cast<volatile uint32_t*>(tmem_slot), 0);
  if (threadIdx.x == 0) trace_stamp(p, 1);               // barriers + TMEM ready
  // Everything above overlaps the tail of the query normalise/cast kernel (programmatic
  // dependent launch); its output (the bf16 query workspace) is only touched below.
  ptx::pdl_wait();

  // static persistent schedule: all roles walk the same unit list
  const int worker = static_cast<int>(blockIdx.x) / CG;
  const int num_workers = static_cast<int>(gridDim.x) / CG;
  const int num_units = p.num_m_tiles * p.num_chunks;

  // Solo mode, distributed prologue: the warps of the whole grid share the rows of the query
  // batch (F.normalize + bf16 cast into the workspace the query tensor map points at); the TMA
  // producers wait for the grid-wide arrival counter before their first query load.
  // pair: operand bytes of both CTAs are accounted on the LEADER's full barrier
  uint32_t full_leader0 = full_bar(0);
  if constexpr (CG == 2) {
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(full_leader0) : "r"(full_bar(0)));
    full_leader0 = __shfl_sync(0xffffffffu, full_leader0, 0);
  }
  int pre_issued = 0;     // producer: ring slots whose bank half is already in flight
  if (MODE == MODE_TOPK && p.solo != 0 && p.q_src != nullptr) {
    // The bank does not depend on the cast: the producer puts the bank half of its first ring
    // slots in flight now (each slot's barrier is armed for the full slot, so the MMA still waits
    // for the query half) and the cold HBM latency of the first loads hides behind the cast.
    if (warp == 0 && worker < num_units) {
      const int b_row = (worker / p.num_m_tiles) * p.tiles_per_chunk * BLOCK_N +
                        static_cast<int>(cta_rank) * (BLOCK_N / CG);
      pre_issued = min(STAGES, p.num_k_blocks);
      for (int st = 0; st < pre_issued; ++st) {
        const uint32_t b_dst = base_u32 + st * STAGE_STRIDE + A_STAGE_BYTES;
        if (ptx::elect_one()) {
          if constexpr (CG == 1) {
            ptx::mbar_arrive_expect_tx(full_bar(st), TX_BYTES);
            ptx::tma_load_2d(b_dst, &tmap_b, full_bar(st), st * BLOCK_K, b_row);
          } else {
            if (is_leader) ptx::mbar_arrive_expect_tx(full_bar(st), TX_BYTES);
            ptx::tma_load_2d_cg2(b_dst, &tmap_b, full_leader0 + 8u * st, st * BLOCK_K, b_row);
          }
        }
      }
    }
    const int n_gwarps = static_cast<int>(gridDim.x) * (NUM_THREADS / 32);
    for (int r = static_cast<int>(blockIdx.x) * (NUM_THREADS / 32) + warp; r < p.q_pad; r += n_gwarps) {
      __nv_bfloat16* dst = p.q_ws + static_cast<size_t>(r) * p.d;
      if (r >= p.Q) {
        zero_row(dst, p.d, lane);
      } else if (p.q_src_bf16) {
        normalize_cast_row(static_cast<const __nv_bfloat16*>(p.q_src) + static_cast<size_t>(r) * p.d,
                           dst, p.d, p.q_normalize, lane);
      } else {
        normalize_cast_row(static_cast<const float*>(p.q_src) + static_cast<size_t>(r) * p.d, dst,
                           p.d, p.q_normalize, lane);
      }
    }
    // generic-proxy writes, read by other CTAs' TMA (async proxy): CTA barrier + one cumulative
    // gpu-scope fence on the writer side, counter acquire + proxy fence on the reader side
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();          // cumulative: orders the whole CTA's row writes (seen through the barrier)
      atomicAdd(p.grid_cnt, 1ull);
      trace_stamp(p, 8);                                   // this CTA's share of the queries is cast
    }
  }

  // The producer and the MMA issuer run their loops with the WHOLE warp (all values warp-uniform)
  // and elect one lane only around the instructions that must be issued once.  Inside an
  // `if (lane == 0)` region the compiler cannot use the uniform datapath, and every
  // tcgen05.mma / TMA operand (they live in uniform registers) then costs an elect + R2UR
  // "waterfall" loop: ~90 dependent instructions per K-step on the issuing thread, which —
  // not the tensor pipe — capped the kernel at 80-88 % tensor duty (profiles/r01/SUMMARY.md).
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    int issued = 0;
    const bool sync_on = (p.sync_cnt != nullptr) && is_leader;   // the leader paces the pair
    bool sync_wait = sync_on;
    int iter = 0;
    if (MODE == MODE_TOPK && p.solo != 0 && p.q_src != nullptr) {
      grid_wait(p.grid_cnt, p.cast_target, p.err_flag, ERR_GRID_CAST);   // every CTA has cast its rows
      asm volatile("fence.proxy.async;" ::: "memory");
      if (lane == 0) trace_stamp(p, 9);                    // the whole batch is cast: loads may start
    }
    for (int u = worker; u < num_units; u += num_workers, ++iter) {
      const int m_tile = u % p.num_m_tiles;
      const int chunk = u / p.num_m_tiles;
      const UnitRange ur = unit_range(p, chunk);
      const int t0 = ur.t0, t1 = ur.t1;
      const int q_row = (m_tile * CG + static_cast<int>(cta_rank)) * BLOCK_M;
      for (int t = t0; t < t1; ++t) {
        const int b_row = t * BLOCK_N + static_cast<int>(cta_rank) * (BLOCK_N / CG);
        if (sync_on && (t - t0) % p.sync_window == 0) {
          const int win = iter * p.windows_per_unit + (t - t0) / p.sync_window;
          if (sync_wait && win > 0) {
            int in_step = 1;
            if (lane == 0) {
              const unsigned int* prev = p.sync_cnt + (win - 1);
              const long long w0 = clock64();
              while (ld_acquire_u32(prev) < static_cast<unsigned int>(num_workers)) {
                if (clock64() - w0 > SYNC_WAIT_LIMIT_CYCLES) { in_step = 0; break; }
                __nanosleep(256);
              }
            }
            if (__shfl_sync(0xffffffffu, in_step, 0) == 0) sync_wait = false;
          }
        }
        for (int kb = ur.kb0; kb < ur.kb1; ++kb) {
          // (an L2 prefetch of the next bank tile via cp.async.bulk.prefetch.tensor was measured
          //  twice: it halved the HBM-bound throughput in round 1 and cost 4-65 % on every small
          //  shape in round 2 — profiles/r02/SUMMARY.md — so the ring is the only look-ahead)
          const bool bank_in_flight = issued < pre_issued;   // armed + bank half loaded in the prologue
          ++issued;
          if (!bank_in_flight) ptx::mbar_wait(empty_bar(stage), phase ^ 1u, p.err_flag, ERR_PRODUCER);
          const uint32_t a_dst = base_u32 + stage * STAGE_STRIDE;
          const uint32_t b_dst = a_dst + A_STAGE_BYTES;
          if (ptx::elect_one()) {
            if constexpr (CG == 1) {
              if (!bank_in_flight) ptx::mbar_arrive_expect_tx(full_bar(stage), TX_BYTES);
              ptx::tma_load_2d(a_dst, &tmap_q, full_bar(stage), kb * BLOCK_K, q_row);
              if (!bank_in_flight) ptx::tma_load_2d(b_dst, &tmap_b, full_bar(stage), kb * BLOCK_K, b_row);
            } else {
              if (is_leader && !bank_in_flight) ptx::mbar_arrive_expect_tx(full_bar(stage), TX_BYTES);
              const uint32_t full_leader = full_leader0 + 8u * stage;
              ptx::tma_load_2d_cg2(a_dst, &tmap_q, full_leader, kb * BLOCK_K, q_row);
              if (!bank_in_flight) ptx::tma_load_2d_cg2(b_dst, &tmap_b, full_leader, kb * BLOCK_K, b_row);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        if (sync_on && lane == 0 &&
            ((t - t0) % p.sync_window == p.sync_window - 1 || t == t1 - 1)) {
          // loads of this window are issued: count this worker in
          atomicAdd(p.sync_cnt + iter * p.windows_per_unit + (t - t0) / p.sync_window, 1u);
        }
      }
      if (sync_on && lane == 0) {   // a ragged (shorter) last chunk: count the windows this unit does not have
        for (int w = (t1 - t0 + p.sync_window - 1) / p.sync_window; w < p.windows_per_unit; ++w)
          atomicAdd(p.sync_cnt + iter * p.windows_per_unit + w, 1u);
      }
    }
    if (sync_on && lane == 0) {     // workers with one unit fewer: count the iterations they do not run
      for (; iter < p.max_iters; ++iter)
        for (int w = 0; w < p.windows_per_unit; ++w)
          atomicAdd(p.sync_cnt + iter * p.windows_per_unit + w, 1u);
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (is_leader) {
      constexpr uint32_t IDESC = ptx::make_idesc_bf16_f32(BLOCK_M * CG, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t tile_count = 0;
      for (int u = worker; u < num_units; u += num_workers) {
        const int chunk = u / p.num_m_tiles;
        const UnitRange ur = unit_range(p, chunk);
        const int t0 = ur.t0, t1 = ur.t1;
        for (int t = t0; t < t1; ++t, ++tile_count) {
          const uint32_t acc = tile_count & 1u;
          const uint32_t acc_phase = (tile_count >> 1) & 1u;
          ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u, p.err_flag, ERR_MMA_TEMPTY);
          ptx::tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
          for (int kb = ur.kb0; kb < ur.kb1; ++kb) {
            ptx::mbar_wait(full_bar(stage), phase, p.err_flag, ERR_MMA_FULL);
            ptx::tc_fence_after();
            const uint32_t a_src = base_u32 + stage * STAGE_STRIDE;
            const uint64_t a_desc = ptx::make_smem_desc_sw128(a_src);
            const uint64_t b_desc = ptx::make_smem_desc_sw128(a_src + A_STAGE_BYTES);
            if (ptx::elect_one()) {
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                // advance 16 bf16 = 32 bytes inside the 128-byte swizzle atom: +2 in the >>4 address field
                ptx::umma_bf16<CG>(tmem_d, a_desc + 2u * k, b_desc + 2u * k, IDESC,
                                   static_cast<uint32_t>(kb != ur.kb0 || k != 0));
              }
              if constexpr (CG == 1) ptx::umma_commit(empty_bar(stage));
              else ptx::umma_commit_cg2(empty_bar(stage), 0b11);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
          if (ptx::elect_one()) {
            if constexpr (CG == 1) ptx::umma_commit(tfull_bar(acc));
            else ptx::umma_commit_cg2(tfull_bar(acc), 0b11);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp >= EPI_WARP0) {
    // ------------------------------------------------------------------ epilogue: running top-k
    const int quarter = warp & 3;                   // TMEM lanes [32*quarter, +32) belong to this warp
    const int half = (warp - EPI_WARP0) >> 2;       // which EPI_COLS-wide slice of every tile
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t tmem_lane = static_cast<uint32_t>(quarter * 32) << 16;
    uint32_t tile_count = 0;
    TopkList<KCAP> list;
    for (int u = worker; u < num_units; u += num_workers) {
      const int m_tile = u % p.num_m_tiles;
      const int chunk = u / p.num_m_tiles;
      const UnitRange ur = unit_range(p, chunk);
      const int t0 = ur.t0, t1 = ur.t1;
      const int row = (m_tile * CG + static_cast<int>(cta_rank)) * BLOCK_M + row_in_tile;
      int self_col = -1;
      if (MODE == MODE_TOPK && p.self_index != nullptr && row < p.Q) {
        const long long g = p.self_index[row];
        const long long c = g - p.index_offset;
        if (g >= 0 && c >= 0 && c < p.n_bank) self_col = static_cast<int>(c);
      }
      list.init(p.k);
      // admission threshold = max(k-th score of this list, what the row's other lists published)
      const bool share = (MODE == MODE_TOPK) && p.row_thr != nullptr && row < p.Q;
      float seed = -CUDART_INF_F;
      float published = -CUDART_INF_F;
      unsigned long long seed_entry = 0;
      if (share) seed_entry = ld_relaxed_u64(p.row_thr + row);
      float thr = list.threshold();
      // k > 32, pass >= 2: only elements strictly after (bound_s, bound_col) in the result order
      const bool bounded = (MODE == MODE_TOPK) && p.bound_scores != nullptr;
      float bound_s = CUDART_INF_F;
      long long bound_col = -1;
      if (bounded && row < p.Q) {
        bound_s = p.bound_scores[static_cast<size_t>(row) * p.bound_stride];
        bound_col = p.bound_idx[static_cast<size_t>(row) * p.bound_stride] - p.index_offset;
      }
      // bootstrap (first wave only: later units inherit thresholds through row_thr)
      int boot_expect = 0;
      if (MODE == MODE_TOPK && p.boot != nullptr && u == worker && m_tile < num_workers) {
        const int chunks_now = min((num_workers - 1 - m_tile) / p.num_m_tiles + 1, p.num_chunks);
        boot_expect = min(p.boot_slots, chunks_now * EPI_HALVES);
        if (boot_expect < p.k) boot_expect = 0;    // fewer concurrent lists than k: nothing to gain
      }
      // RANK mode state (KCAP = target slots)
      float ts[KCAP];
      int tc[KCAP];
      int cnt[KCAP];
      if constexpr (RANK) {
#pragma unroll
        for (int g = 0; g < KCAP; ++g) {
          ts[g] = CUDART_INF_F;      // nothing is greater than +inf: unused slots count 0
          tc[g] = -1;
          cnt[g] = 0;
          if (row < p.Q && g < p.n_targets) {
            tc[g] = p.tgt_cols[static_cast<size_t>(row) * p.n_targets + g];
            if (tc[g] >= 0) ts[g] = p.tgt_scores[static_cast<size_t>(row) * p.n_targets + g];
          }
        }
      }
      for (int t = t0; t < t1; ++t, ++tile_count) {
        const uint32_t acc = tile_count & 1u;
        const uint32_t acc_phase = (tile_count >> 1) & 1u;
        ptx::mbar_wait(tfull_bar(acc), acc_phase, p.err_flag, ERR_EPILOGUE);
        ptx::tc_fence_after();
        if (tile_count == 0 && warp == EPI_WARP0 && lane == 0) trace_stamp(p, 2);  // first tile's MMAs done
        const int col_tile = t * BLOCK_N;
        if constexpr (MODE == MODE_TOPK) {
          // the key was requested one tile ago (at unit start for the first tile), so its latency
          // is hidden behind a whole tile of scanning; the next one is requested right away
          seed = fmaxf(seed, seed_below(epoch_key(seed_entry, p.epoch)));
          if (boot_expect > 0 && t == t0) {
            __syncwarp();
            const int list_id = chunk * EPI_HALVES + half;
            const int row_c = min(row, p.q_pad - 1);
            seed = fmaxf(seed, boot_threshold(tmem_base + tmem_lane + acc * BLOCK_N + half * EPI_COLS,
                                              col_tile + half * EPI_COLS, p.n_bank, self_col, row < p.Q,
                                              p.boot + static_cast<size_t>(row_c) * BOOT_SLOTS,
                                              list_id % p.boot_slots, boot_expect, p.k, p.epoch));
          }
          thr = fmaxf(thr, seed);
          if (share) seed_entry = ld_relaxed_u64(p.row_thr + row);
        }
#pragma unroll 1
        for (int c0 = half * EPI_COLS; c0 < (half + 1) * EPI_COLS; c0 += 32) {
          uint32_t r[32];
          __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the divergent insert path
          ptx::tmem_ld_32x32(tmem_base + tmem_lane + acc * BLOCK_N + c0, r);
          ptx::tmem_ld_wait();
          const int col0 = col_tile + c0;
          if constexpr (DUMP) {
            if (row < p.Q) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.n_bank)
                  p.dump[(static_cast<size_t>(ur.k_chunk) * p.Q + row) * p.n_bank + col0 + j] = __uint_as_float(r[j]);
            }
          } else if constexpr (RANK) {
            if (col0 + 32 <= p.n_bank) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float v = __uint_as_float(r[j]);
#pragma unroll
                for (int g = 0; g < KCAP; ++g) cnt[g] += (v > ts[g]) ? 1 : 0;
              }
            } else {   // ragged bank tail: zero-filled columns beyond the bank do not count
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float v = __uint_as_float(r[j]);
                const bool valid = col0 + j < p.n_bank;
#pragma unroll
                for (int g = 0; g < KCAP; ++g) cnt[g] += (valid && v > ts[g]) ? 1 : 0;
              }
            }
            // a target never counts against itself, whatever the rounding of its own MMA score
#pragma unroll
            for (int g = 0; g < KCAP; ++g) {
              if (tc[g] >= col0 && tc[g] < col0 + 32) {
                const float own = __uint_as_float(select32(r, tc[g] - col0));
                if (own > ts[g]) cnt[g] -= 1;
              }
            }
          } else {
            // fast path: two instructions per score, no branch
            uint32_t cand = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (__uint_as_float(r[j]) > thr) cand |= (1u << j);
            if (cand != 0) {
              // per-thread rare path (but some lane of the warp takes it for most chunks)
              while (cand != 0) {
                const int j = __ffs(cand) - 1;
                cand &= cand - 1;
                const float v = __uint_as_float(select32(r, j));
                const int col = col0 + j;
                if (v > thr && col < p.n_bank && col != self_col &&
                    (!bounded || v < bound_s || (v == bound_s && static_cast<long long>(col) > bound_col))) {
                  list.insert(v, col);
                  thr = fmaxf(list.threshold(), seed);
                }
              }
            }
          }
        }
        // this warp is done with accumulator `acc` (tcgen05.wait::ld is warp-collective): hand
        // it back to the MMA issuer with one arrival per warp
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 1) ptx::mbar_arrive(tempty_bar(acc));
          else ptx::mbar_arrive_cluster(tempty_bar(acc), 0);
        }
        if constexpr (MODE == MODE_TOPK) {
          if (share) {
            const float kth = list.threshold();      // > -inf only once the list holds k entries
            if (kth > published) {
              atomicMax(p.row_thr + row, make_epoch_key(p.epoch, kth));   // result unused: a fire-and-forget RED
              published = kth;
            }
          }
        }
      }
      if (warp == EPI_WARP0 && lane == 0) trace_stamp(p, 3);   // last tile of the unit scanned
      if constexpr (RANK) {
        if (row < p.Q) {
          const size_t o = (static_cast<size_t>(chunk * EPI_HALVES + half) * p.Q + row) * p.n_targets;
#pragma unroll
          for (int g = 0; g < KCAP; ++g)
            if (g < p.n_targets) p.part_counts[o + g] = cnt[g];
        }
      }
      if constexpr (MODE == MODE_TOPK) {
        if (row < p.Q) {
          const size_t o = (static_cast<size_t>(chunk * EPI_HALVES + half) * p.Q + row) * p.k;
          const int pinned = KCAP - p.k;   // slots [pinned, KCAP) are the live entries
#pragma unroll
          for (int j = 0; j < KCAP; ++j) {
            if (j >= pinned) {
              p.part_scores[o + j - pinned] = list.s[j];
              p.part_idx[o + j - pinned] = list.i[j];
            }
          }
        }
        if (p.solo != 0) __threadfence();   // the lists are merged by other CTAs of this launch
      }
    }
  }

  if (warp == EPI_WARP0 && lane == 0) trace_stamp(p, 4);     // partial lists written
  if (threadIdx.x == 0) ptx::pdl_launch_dependents();        // the merge kernel may start its prologue
  ptx::tc_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync_all(); else __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<CG>(tmem_base, TMEM_COLS);
  if constexpr (MODE == MODE_TOPK) {
    if (p.solo != 0) {
      // Solo mode, distributed merge: once every CTA has written (and fenced) its partial lists,
      // the warps of the whole grid share the query rows and merge the lists of each.
      if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(p.grid_cnt + 1, 1ull);
      }
      grid_wait(p.grid_cnt + 1, p.done_target, p.err_flag, ERR_GRID_DONE);
      if (threadIdx.x == 0) trace_stamp(p, 10);            // every CTA's lists are written
      solo_merge(p.part_scores, p.part_idx, p.num_chunks * EPI_HALVES, p.Q, p.k, p.index_offset,
                 p.out_scores, p.out_idx, p.out_stride,
                 static_cast<int>(blockIdx.x) * (NUM_THREADS / 32) + warp,
                 static_cast<int>(gridDim.x) * (NUM_THREADS / 32), lane);
    }
  }
  if (threadIdx.x == 0) {
    trace_stamp(p, 5);                                       // exit
    if (p.trace != nullptr) p.trace[static_cast<size_t>(blockIdx.x) * TRACE_SLOTS + 7] = clock64();
  }
}

}  // namespace zs
