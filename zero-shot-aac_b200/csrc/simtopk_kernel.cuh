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
enum : int { ERR_PRODUCER = 101, ERR_MMA_FULL = 102, ERR_MMA_TEMPTY = 103, ERR_EPILOGUE = 104 };

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
  float* dump;           // DUMP mode: [Q, n_bank]
  // RANK mode: per (row, target) the target's score and column within the shard (-1 = unused),
  // and the per-(chunk, half) partial counts
  const float* tgt_scores;   // [Q, n_targets]
  const int* tgt_cols;       // [Q, n_targets]
  int n_targets;
  int* part_counts;          // [num_chunks * EPI_HALVES, Q, n_targets]
  int* err_flag;
  unsigned long long* trace;  // nullable: [gridDim.x, 8] globaltimer stamps (zs_debug_trace)
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
  unsigned int* row_thr;      // [Q padded], keys; zeroed per search by the query cast kernel
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
__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

constexpr long long SYNC_WAIT_LIMIT_CYCLES = 600000;   // ~0.3-0.4 ms: then give up lock-step

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void trace_stamp(const SimTopkParams& p, int slot) {
  if (p.trace != nullptr) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[static_cast<size_t>(blockIdx.x) * 8 + slot] = t;
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

// MODE_TOPK: running top-k (KCAP list slots).  MODE_DUMP: write the score matrix (test hook).
// MODE_RANK: count, per query row and per target (KCAP = max targets per row), the bank rows
// whose score is strictly greater than the target's score — the rank of the ground truth that
// the reference's retrieval metrics obtain from a full argsort (retrieval/tools/utils.py:183,236).
enum : int { MODE_TOPK = 0, MODE_DUMP = 1, MODE_RANK = 2 };

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
    if (p.trace != nullptr) p.trace[static_cast<size_t>(blockIdx.x) * 8 + 6] = clock64();
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
    // pair: operand bytes of both CTAs are accounted on the LEADER's full barrier
    uint32_t full_leader0 = full_bar(0);
    if constexpr (CG == 2) {
      asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(full_leader0) : "r"(full_bar(0)));
      full_leader0 = __shfl_sync(0xffffffffu, full_leader0, 0);
    }
    const bool sync_on = (p.sync_cnt != nullptr) && is_leader;   // the leader paces the pair
    bool sync_wait = sync_on;
    int iter = 0;
    for (int u = worker; u < num_units; u += num_workers, ++iter) {
      const int m_tile = u % p.num_m_tiles;
      const int chunk = u / p.num_m_tiles;
      const int t0 = chunk * p.tiles_per_chunk;
      const int t1 = min(t0 + p.tiles_per_chunk, p.num_n_tiles);
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
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          // (an L2 prefetch of the next bank tile via cp.async.bulk.prefetch.tensor was measured
          //  and halved the HBM-bound throughput — profiles/r01/SUMMARY.md — so the ring is the
          //  only look-ahead)
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u, p.err_flag, ERR_PRODUCER);
          const uint32_t a_dst = base_u32 + stage * STAGE_STRIDE;
          const uint32_t b_dst = a_dst + A_STAGE_BYTES;
          if (ptx::elect_one()) {
            if constexpr (CG == 1) {
              ptx::mbar_arrive_expect_tx(full_bar(stage), TX_BYTES);
              ptx::tma_load_2d(a_dst, &tmap_q, full_bar(stage), kb * BLOCK_K, q_row);
              ptx::tma_load_2d(b_dst, &tmap_b, full_bar(stage), kb * BLOCK_K, b_row);
            } else {
              if (is_leader) ptx::mbar_arrive_expect_tx(full_bar(stage), TX_BYTES);
              const uint32_t full_leader = full_leader0 + 8u * stage;
              ptx::tma_load_2d_cg2(a_dst, &tmap_q, full_leader, kb * BLOCK_K, q_row);
              ptx::tma_load_2d_cg2(b_dst, &tmap_b, full_leader, kb * BLOCK_K, b_row);
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
        const int t0 = chunk * p.tiles_per_chunk;
        const int t1 = min(t0 + p.tiles_per_chunk, p.num_n_tiles);
        for (int t = t0; t < t1; ++t, ++tile_count) {
          const uint32_t acc = tile_count & 1u;
          const uint32_t acc_phase = (tile_count >> 1) & 1u;
          ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u, p.err_flag, ERR_MMA_TEMPTY);
          ptx::tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
          for (int kb = 0; kb < p.num_k_blocks; ++kb) {
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
                                   static_cast<uint32_t>((kb | k) != 0));
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
      const int t0 = chunk * p.tiles_per_chunk;
      const int t1 = min(t0 + p.tiles_per_chunk, p.num_n_tiles);
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
      unsigned int seed_key = 0;
      if (share) seed_key = ld_relaxed_u32(p.row_thr + row);
      float thr = list.threshold();
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
          seed = fmaxf(seed, seed_below(seed_key));
          thr = fmaxf(thr, seed);
          if (share) seed_key = ld_relaxed_u32(p.row_thr + row);
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
                  p.dump[static_cast<size_t>(row) * p.n_bank + col0 + j] = __uint_as_float(r[j]);
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
                if (v > thr && col < p.n_bank && col != self_col) {
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
              atomicMax(p.row_thr + row, score_key(kth));   // result unused: a fire-and-forget RED
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
      }
    }
  }

  if (warp == EPI_WARP0 && lane == 0) trace_stamp(p, 4);     // partial lists written
  if (threadIdx.x == 0) ptx::pdl_launch_dependents();        // the merge kernel may start its prologue
  ptx::tc_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync_all(); else __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<CG>(tmem_base, TMEM_COLS);
  if (threadIdx.x == 0) {
    trace_stamp(p, 5);                                       // exit
    if (p.trace != nullptr) p.trace[static_cast<size_t>(blockIdx.x) * 8 + 7] = clock64();
  }
}

}  // namespace zs
