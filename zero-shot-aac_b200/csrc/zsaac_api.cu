// zsaac_api.cu — host side of libzsaac_b200.so: the C ABI declared in include/zsaac.h.
//
// Owns the bf16 bank copy, the TMA descriptors and the search workspaces; picks the launch
// geometry (bank chunks per query tile) and enqueues  normalize/cast -> fused similarity+top-k
// -> chunk merge  on the caller's stream.  No CPU path exists: without an sm_100 device
// zs_create fails.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "../../include/zsaac.h"
#include "aux_kernels.cuh"
#include "exact_f32_kernels.cuh"
#include "memproj_kernel.cuh"
#include "memproj_tc_kernels.cuh"
#include "simtopk_kernel.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define ZS_CUDA(expr)                                                                          \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      return fail(ZS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),        \
                  __FILE__, __LINE__);                                                         \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Every kernel moves rows with 8- or 16-byte vector accesses; torch allocations are 256-byte
// aligned and row offsets are multiples of 2*d (d % 64 == 0), but a raw pointer handed in over
// the C ABI might not be.
bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    int cur = -1;
    cudaGetDevice(&cur);
    if (prev >= 0 && cur != prev) cudaSetDevice(prev);
  }
};

}  // namespace

struct zs_ctx {
  int device = 0;
  int sm_count = 0;
  int cta_group_override = 0;  // 0 = choose per search; env ZSAAC_CTA_GROUP=1|2 pins it (tests)
  EncodeTiledFn encode = nullptr;

  __nv_bfloat16* bank = nullptr;
  int64_t bank_rows = 0;
  int bank_d = 0;
  CUtensorMap bank_map[2];     // [0]: box {64, 256} (cta_group::1)  [1]: box {64, 128} (cta_group::2)
  // search window (zs_bank_window): zs_search ranks rows [win_lo, win_lo + win_rows) only
  int64_t win_lo = 0, win_rows = 0;   // win_rows == 0: no window, the whole bank
  CUtensorMap win_map[2];

  __nv_bfloat16* q_ws = nullptr;  // bf16 (normalised) queries
  int64_t q_ws_rows = 0;
  unsigned long long* row_thr = nullptr;  // [q_ws_rows] shared admission thresholds (SimTopkParams::row_thr)
  unsigned long long* boot = nullptr;     // [q_ws_rows, BOOT_SLOTS] bootstrap slot maxima (SimTopkParams::boot)
  unsigned int epoch = 0;                 // tags row_thr / boot entries; bumped per search pass
  unsigned long long* grid_cnt = nullptr; // [2] monotonic arrival counters of the single-launch mode
  unsigned long long cnt_base[2] = {0, 0};
  int solo_state = 0;                     // 0 = not probed, 1 = single-launch mode works, -1 = unavailable
  int solo_override = -1;                 // env ZSAAC_SOLO=0|1 (tests / A-B runs); -1 = choose per search
  bool cooperative = false;               // env ZSAAC_COOPERATIVE=1: launch single-launch searches cooperatively
  int boot_override = -1;                 // env ZSAAC_BOOT=0|1: bootstrap off / on wherever it is valid; -1 = per search
  float* part_scores = nullptr;   // [chunks * EPI_HALVES, Q, k]
  int* part_idx = nullptr;
  int64_t part_elems = 0;
  int* err_flag = nullptr;            // device alias of err_host
  int* err_host = nullptr;            // pinned + mapped: readable after a kernel trap
  float* tgt_scores = nullptr;        // rank mode: [Q, T]
  int* tgt_cols = nullptr;
  int* part_counts = nullptr;         // [chunks * EPI_HALVES, Q, T]
  int64_t tgt_elems = 0, count_elems = 0;
  float* memproj_partials = nullptr;  // zs_memory_project: [blocks, 2, d + 4]
  int64_t memproj_elems = 0;
  // batched memory projection on the tensor cores (zs_memory_bank_prepare / zs_memory_project_batched)
  __nv_bfloat16* mp_bank = nullptr;    // B'   [n, 3d]      bank rows, B side [hi | hi | lo]
  __nv_bfloat16* mp_bank_t = nullptr;  // Bt'' [d, 3 n_pad] bank transposed, B side
  int64_t mp_rows = 0, mp_pad = 0;
  int mp_d = 0;
  __nv_bfloat16* mp_q = nullptr;       // Q'   [q_pad, 3d]      A side [hi | lo | hi]
  float* mp_scores = nullptr;          // S    [Q, n]
  __nv_bfloat16* mp_p = nullptr;       // P''  [q_pad, 3 n_pad] A side
  float* mp_partial = nullptr;         // [k chunks, Q, d]
  int64_t mp_q_rows = 0, mp_partial_elems = 0;
  unsigned int* sync_cnt = nullptr;   // lock-step window counters (see SimTopkParams)
  int64_t sync_cnt_elems = 0;
  float* exact_scores = nullptr;      // zs_exact_*: [Q, n_rows] fp32 score scratch
  int64_t exact_elems = 0;
  unsigned int* exact_counter = nullptr;   // "last block done" ticket of the fused exact top-k

  int64_t launches = 0;

  unsigned long long* trace = nullptr;  // zs_debug_trace: caller-owned [ctas, 8] device buffer
  bool pdl_enabled = true;         // env ZSAAC_PDL=0 switches programmatic dependent launch off
  bool pdl_next = false;           // next fused-kernel launch directly follows its producer kernel
  bool profiling = false;          // zs_profile_enable
  cudaEvent_t* prof_ev = nullptr;  // 2 * ZS_PROFILE_RING events (start, stop)
  int prof_count = 0;              // launches recorded since enable (ring wraps)
};

namespace {

int encode_rows_map(zs_ctx* ctx, CUtensorMap* map, const void* base, int64_t rows, int64_t d,
                    int box_rows) {
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(d), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(d) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(zs::BLOCK_K), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = ctx->encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim,
                           gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(ZS_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld d=%lld box=%d)",
                static_cast<int>(r), static_cast<long long>(rows), static_cast<long long>(d), box_rows);
  return ZS_OK;
}

// Physical slots of the register-resident top-k list: an insert costs ~5 instructions per slot,
// so the list is sized to the request (k = 5 -> 8, k = 10 -> 12, k = 32 -> 32).
int kcap_for(int k) {
  if (const char* e = getenv("ZSAAC_KCAP_POW2")) {    // tuning hook: the former 8 / 16 / 32 only
    if (e[0] == '1') return k <= 8 ? 8 : (k <= 16 ? 16 : 32);
  }
  return k <= 8 ? 8 : (k <= 12 ? 12 : (k <= 16 ? 16 : (k <= 24 ? 24 : 32)));
}

// rows zs_search ranks: the window if one is set, else the whole bank
int64_t search_rows(const zs_ctx* ctx) { return ctx->win_rows > 0 ? ctx->win_rows : ctx->bank_rows; }

// zs_rank_count / zs_debug_scores always work on the whole bank: the window is lifted for the call
struct WholeBank {
  zs_ctx* ctx;
  int64_t lo, rows;
  explicit WholeBank(zs_ctx* c) : ctx(c), lo(c->win_lo), rows(c->win_rows) { c->win_lo = c->win_rows = 0; }
  ~WholeBank() { ctx->win_lo = lo; ctx->win_rows = rows; }
};

struct Plan {
  int cg, m_tiles, n_tiles, chunks, tiles_per_chunk, ctas;
  int sync_window, windows_per_unit, max_iters;   // lock-step of the bank stream (0 = off)
  double cost;                                    // planner's estimate, in CG=2 bank-tile times
};

constexpr int kSyncWindowTiles = 32;   // 32 tiles x 512 KiB = 16 MiB of bank per window

// A single CTA (cta_group::1) has to pull 768 KiB of operands through shared memory per bank tile,
// a CTA of a pair only 512 KiB, for the same 8192 cycles of MMA: measured 4-9 % more time per
// tile (profiles/r02/bench_small_variants.jsonl).
constexpr double kSingleCtaTileCost = 1.08;

// Split the bank into `chunks` contiguous runs of 256-row tiles so that (query tiles x chunks)
// work units fill the SMs in whole waves with the least padded work.
Plan make_plan_for(const zs_ctx* ctx, int64_t Q, int k, int cg) {
  Plan pl{};
  pl.cg = cg;
  const int workers = std::max(1, ctx->sm_count / cg);
  pl.m_tiles = static_cast<int>((Q + zs::BLOCK_M * cg - 1) / (zs::BLOCK_M * cg));
  pl.n_tiles = static_cast<int>((search_rows(ctx) + zs::BLOCK_N - 1) / zs::BLOCK_N);
  const int max_chunks = std::min(pl.n_tiles, zs::MERGE_MAX_LISTS / zs::EPI_HALVES);
  double best_cost = 1e300;
  int best_s = 1;
  for (int s = 1; s <= max_chunks; ++s) {
    const int tpc = (pl.n_tiles + s - 1) / s;
    const int s_eff = (pl.n_tiles + tpc - 1) / tpc;
    if (s_eff != s) continue;  // same split as a smaller s
    const int64_t units = static_cast<int64_t>(pl.m_tiles) * s_eff;
    const int64_t waves = (units + workers - 1) / workers;
    // per unit: tpc tiles + pipeline fill / partial write-out (~0.75 tile) + what is left of the
    // top-k list warm-up once units share their thresholds (grows with k; fitted to the chunk
    // sweeps in profiles/r01/sweep_chunks_*.jsonl)
    const double cost = static_cast<double>(waves) * (tpc + 0.75 + 0.15 * k);
    if (cost < best_cost - 1e-9) { best_cost = cost; best_s = s_eff; }
  }
  if (const char* forced = getenv("ZSAAC_CHUNKS")) {   // tuning hook: pin the number of bank chunks
    const int s = atoi(forced);
    if (s >= 1 && s <= max_chunks) {
      const int tpc = (pl.n_tiles + s - 1) / s;
      best_s = (pl.n_tiles + tpc - 1) / tpc;
    }
  }
  pl.chunks = best_s;
  pl.tiles_per_chunk = (pl.n_tiles + best_s - 1) / best_s;
  pl.cost = best_cost * (cg == 1 ? kSingleCtaTileCost : 1.0);
  const int64_t units = static_cast<int64_t>(pl.m_tiles) * pl.chunks;
  const int n_workers = static_cast<int>(std::min<int64_t>(units, workers));
  pl.ctas = n_workers * cg;
  // Lock-step pays when several workers stream the same long chunk (m_tiles > 1) and the chunk
  // is much larger than what stays in L2 between the fastest and the slowest worker.
  const char* sync_env = getenv("ZSAAC_LOCKSTEP");
  const bool sync_allowed = !(sync_env && sync_env[0] == '0');
  int window = kSyncWindowTiles;
  if (const char* w = getenv("ZSAAC_SYNC_WINDOW")) {   // tuning hook: tiles per lock-step window
    if (atoi(w) >= 1) window = atoi(w);
  }
  if (sync_allowed && pl.m_tiles > 1 && n_workers > 1 && pl.tiles_per_chunk >= 4 * window) {
    pl.sync_window = window;
    pl.windows_per_unit = (pl.tiles_per_chunk + window - 1) / window;
    pl.max_iters = static_cast<int>((units + n_workers - 1) / n_workers);
  }
  return pl;
}

// One 128-row query tile streams the bank fastest from independent CTAs (an HBM-bound search).
// Beyond that: CTA pairs (cta_group::2, 256-row query tiles) for large batches, where they run at
// the MMA rate; for the few-tile batches in between whichever geometry the cost model prefers
// (e.g. 1,045 queries are 9 tiles of 128 rows but 5 tiles of 256: 18 % padding for pairs).
Plan make_plan(const zs_ctx* ctx, int64_t Q, int k) {
  if (ctx->cta_group_override) return make_plan_for(ctx, Q, k, ctx->cta_group_override);
  if (Q <= zs::BLOCK_M) return make_plan_for(ctx, Q, k, 1);
  const Plan pair = make_plan_for(ctx, Q, k, 2);
  if (Q > 16 * zs::BLOCK_M) return pair;
  const Plan single = make_plan_for(ctx, Q, k, 1);
  return single.cost < pair.cost ? single : pair;
}

// Query workspace rows are padded to whole 256-row tile pairs (zero rows), so the query-side TMA
// boxes are always fully inside the tensor (out-of-bounds boxes are served markedly slower).
int64_t padded_query_rows(int64_t Q) { return (Q + 2 * zs::BLOCK_M - 1) / (2 * zs::BLOCK_M) * (2 * zs::BLOCK_M); }

int ensure_workspace(zs_ctx* ctx, int64_t Q, int k) {
  const int64_t q_pad = padded_query_rows(Q);
  if (q_pad > ctx->q_ws_rows) {
    if (ctx->q_ws) { ZS_CUDA(cudaFree(ctx->q_ws)); ctx->q_ws = nullptr; ctx->q_ws_rows = 0; }
    if (ctx->row_thr) { ZS_CUDA(cudaFree(ctx->row_thr)); ctx->row_thr = nullptr; }
    if (ctx->boot) { ZS_CUDA(cudaFree(ctx->boot)); ctx->boot = nullptr; }
    ZS_CUDA(cudaMalloc(&ctx->q_ws, static_cast<size_t>(q_pad) * ctx->bank_d * sizeof(__nv_bfloat16)));
    ZS_CUDA(cudaMalloc(&ctx->row_thr, static_cast<size_t>(q_pad) * sizeof(unsigned long long)));
    ZS_CUDA(cudaMalloc(&ctx->boot, static_cast<size_t>(q_pad) * zs::BOOT_SLOTS * sizeof(unsigned long long)));
    // entries are epoch-tagged: zero once, never reset again
    ZS_CUDA(cudaMemset(ctx->row_thr, 0, static_cast<size_t>(q_pad) * sizeof(unsigned long long)));
    ZS_CUDA(cudaMemset(ctx->boot, 0, static_cast<size_t>(q_pad) * zs::BOOT_SLOTS * sizeof(unsigned long long)));
    ctx->q_ws_rows = q_pad;
  }
  const Plan pl = make_plan(ctx, Q, k);
  const int64_t sync_need = static_cast<int64_t>(pl.max_iters) * pl.windows_per_unit;
  if (sync_need > ctx->sync_cnt_elems) {
    if (ctx->sync_cnt) { ZS_CUDA(cudaFree(ctx->sync_cnt)); ctx->sync_cnt = nullptr; ctx->sync_cnt_elems = 0; }
    ZS_CUDA(cudaMalloc(&ctx->sync_cnt, static_cast<size_t>(sync_need) * sizeof(unsigned int)));
    ctx->sync_cnt_elems = sync_need;
  }
  const int64_t need = static_cast<int64_t>(pl.chunks) * zs::EPI_HALVES * Q * k;
  if (need > ctx->part_elems) {
    if (ctx->part_scores) { ZS_CUDA(cudaFree(ctx->part_scores)); ctx->part_scores = nullptr; }
    if (ctx->part_idx) { ZS_CUDA(cudaFree(ctx->part_idx)); ctx->part_idx = nullptr; }
    ctx->part_elems = 0;
    ZS_CUDA(cudaMalloc(&ctx->part_scores, static_cast<size_t>(need) * sizeof(float)));
    ZS_CUDA(cudaMalloc(&ctx->part_idx, static_cast<size_t>(need) * sizeof(int)));
    ctx->part_elems = need;
  }
  return ZS_OK;
}

template <typename InT>
void launch_normalize(const void* in, __nv_bfloat16* out, int64_t rows, int64_t rows_out, int d,
                      int normalize, cudaStream_t st) {
  const int64_t blocks = (rows_out * 32 + 255) / 256;
  zs::normalize_cast_kernel<InT, __nv_bfloat16><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
      static_cast<const InT*>(in), out, rows, rows_out, d, normalize);
}



// pdl: launch as a programmatic dependent of the kernel enqueued just before (its prologue may
// overlap that kernel's tail; merge_lists_kernel waits with griddepcontrol.wait before reading)
template <typename IdxT>
cudaError_t launch_merge(const float* scores, const IdxT* idx, int S, int64_t score_stride,
                         int64_t index_stride, int64_t Q, int k, long long idx_offset,
                         float* out_scores, long long* out_idx, int64_t out_stride, bool pdl,
                         cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>((Q * 32 + 255) / 256));
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  // few queries, many lists (a small batch against a bank split into a chunk per SM): a block per
  // query instead of a warp per query
  if (S > 64 && Q <= 4096 && k <= zs::MERGE_BLOCK_MAX_K) {
    cfg.gridDim = dim3(static_cast<unsigned>(Q));
    cfg.blockDim = dim3(32 * zs::MERGE_BLOCK_WARPS);
    return cudaLaunchKernelEx(&cfg, zs::merge_lists_block_kernel<IdxT>, scores, idx, S, score_stride,
                              index_stride, Q, k, idx_offset, out_scores, out_idx, out_stride);
  }
  if (S <= 64)
    return cudaLaunchKernelEx(&cfg, zs::merge_lists_kernel<IdxT, 2>, scores, idx, S, score_stride,
                              index_stride, Q, k, idx_offset, out_scores, out_idx, out_stride);
  if (S <= 256)
    return cudaLaunchKernelEx(&cfg, zs::merge_lists_kernel<IdxT, 8>, scores, idx, S, score_stride,
                              index_stride, Q, k, idx_offset, out_scores, out_idx, out_stride);
  return cudaLaunchKernelEx(&cfg, zs::merge_lists_kernel<IdxT, 16>, scores, idx, S, score_stride,
                            index_stride, Q, k, idx_offset, out_scores, out_idx, out_stride);
}

template <int KCAP, int CG, int MODE>
int launch_simtopk(zs_ctx* ctx, const CUtensorMap& qmap, const CUtensorMap& bmap, const zs::SimTopkParams& p,
                   int ctas, cudaStream_t st) {
  auto kern = zs::zs_simtopk_kernel<KCAP, CG, MODE>;
  const int smem = zs::smem_bytes<CG>();
  static bool smem_opt_in[64] = {};   // per instantiation and device: opt in to > 48 KiB once
  if (ctx->device >= 64 || !smem_opt_in[ctx->device]) {
    ZS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (ctx->device < 64) smem_opt_in[ctx->device] = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(ctas));
  cfg.blockDim = dim3(zs::NUM_THREADS);
  cfg.dynamicSmemBytes = static_cast<size_t>(smem);
  cfg.stream = st;
  cudaLaunchAttribute attr[3];
  int n_attr = 0;
  // Single-launch mode waits on grid-wide arrival counters inside the kernel, so every CTA has to
  // become resident while the others spin.  The grid never exceeds one CTA per SM (198 KiB of
  // shared memory each), so on a GPU that is not running another spinning grid they all are;
  // kernels of other streams only delay the last CTAs.  The cooperative-launch attribute would
  // make that a guarantee, but profilers cannot replay cooperative cluster launches (ncu 2025.2:
  // "LaunchFailed"), so it is opt-in (ZSAAC_COOPERATIVE=1); processes that run searches of one
  // GPU concurrently from several streams should set it, or ZSAAC_SOLO=0.
  if (p.solo != 0 && ctx->cooperative) {
    attr[n_attr].id = cudaLaunchAttributeCooperative;
    attr[n_attr].val.cooperative = 1;
    ++n_attr;
  }
  if (CG == 2) {
    attr[n_attr].id = cudaLaunchAttributeClusterDimension;
    attr[n_attr].val.clusterDim.x = 2;
    attr[n_attr].val.clusterDim.y = 1;
    attr[n_attr].val.clusterDim.z = 1;
    ++n_attr;
  }
  if (ctx->pdl_next && p.solo == 0) {   // the query normalise/cast kernel was enqueued just before this launch
    attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
    ++n_attr;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n_attr;
  const bool prof = ctx->profiling && MODE != zs::MODE_DUMP;
  const int slot = ctx->prof_count % ZS_PROFILE_RING;
  if (prof) ZS_CUDA(cudaEventRecord(ctx->prof_ev[2 * slot], st));
  {
    cudaError_t le = cudaLaunchKernelEx(&cfg, kern, qmap, bmap, p);
    if (le != cudaSuccess && p.solo != 0) {
      cudaGetLastError();          // not sticky: the caller falls back to the three-launch path
      return fail(ZS_ERR_STATE, "single-launch search could not be launched: %s", cudaGetErrorString(le));
    }
    if (le != cudaSuccess)
      return fail(ZS_ERR_CUDA, "cudaLaunchKernelEx(zs_simtopk_kernel) failed: %s", cudaGetErrorString(le));
  }
  if (prof) {
    ZS_CUDA(cudaEventRecord(ctx->prof_ev[2 * slot + 1], st));
    ctx->prof_count += 1;
  }
  ctx->launches += 1;
  return ZS_OK;
}

template <int CG>
int dispatch_simtopk(zs_ctx* ctx, const CUtensorMap& qmap, const zs::SimTopkParams& p, int ctas,
                     bool dump, cudaStream_t st, const CUtensorMap* bank_map = nullptr) {
  const CUtensorMap& bmap = bank_map ? *bank_map : ctx->bank_map[CG - 1];
  if (dump) return launch_simtopk<8, CG, zs::MODE_DUMP>(ctx, qmap, bmap, p, ctas, st);
  if (p.part_counts != nullptr)   // rank mode: KCAP = target slots
    return p.n_targets <= 1 ? launch_simtopk<1, CG, zs::MODE_RANK>(ctx, qmap, bmap, p, ctas, st)
                            : launch_simtopk<8, CG, zs::MODE_RANK>(ctx, qmap, bmap, p, ctas, st);
  switch (kcap_for(p.k)) {
    case 8: return launch_simtopk<8, CG, zs::MODE_TOPK>(ctx, qmap, bmap, p, ctas, st);
    case 12: return launch_simtopk<12, CG, zs::MODE_TOPK>(ctx, qmap, bmap, p, ctas, st);
    case 16: return launch_simtopk<16, CG, zs::MODE_TOPK>(ctx, qmap, bmap, p, ctas, st);
    case 24: return launch_simtopk<24, CG, zs::MODE_TOPK>(ctx, qmap, bmap, p, ctas, st);
    default: return launch_simtopk<32, CG, zs::MODE_TOPK>(ctx, qmap, bmap, p, ctas, st);
  }
}

// Shared front half of zs_search / zs_debug_scores: cast queries, build the query map + params.
int prepare_queries(zs_ctx* ctx, const void* queries, int64_t Q, int q_dtype, int normalize,
                    CUtensorMap* qmap, cudaStream_t st) {
  const int64_t q_pad = padded_query_rows(Q);
  if (q_dtype == ZS_F32)
    launch_normalize<float>(queries, ctx->q_ws, Q, q_pad, ctx->bank_d, normalize, st);
  else
    launch_normalize<__nv_bfloat16>(queries, ctx->q_ws, Q, q_pad, ctx->bank_d, normalize, st);
  ZS_CUDA(cudaGetLastError());
  ctx->launches += 1;
  return encode_rows_map(ctx, qmap, ctx->q_ws, q_pad, ctx->bank_d, zs::BLOCK_M);
}

}  // namespace

extern "C" {

int zs_abi_version(void) { return ZS_ABI_VERSION; }
const char* zs_last_error(void) { return g_err; }
const char* zs_kernel_name(void) { return "zs_simtopk_kernel"; }

int zs_create(zs_ctx** out, int device) {
  if (!out) return fail(ZS_ERR_INVALID, "zs_create: out is null");
  *out = nullptr;
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0) {
    cudaGetLastError();
    return fail(ZS_ERR_NO_DEVICE, "zs_create: no CUDA device (%s); this library has no CPU path",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  if (device < 0 || device >= n_dev)
    return fail(ZS_ERR_INVALID, "zs_create: device %d out of range [0, %d)", device, n_dev);
  cudaDeviceProp prop;
  ZS_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(ZS_ERR_NO_DEVICE, "zs_create: device %d is sm_%d%d; the kernels are sm_100a only",
                device, prop.major, prop.minor);
  zs_ctx* ctx = new (std::nothrow) zs_ctx();
  if (!ctx) return fail(ZS_ERR_INVALID, "zs_create: out of host memory");
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  DeviceGuard guard(device);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    delete ctx;
    return fail(ZS_ERR_CUDA, "zs_create: cuTensorMapEncodeTiled entry point not available");
  }
  ctx->encode = reinterpret_cast<EncodeTiledFn>(fn);
  const char* cg = getenv("ZSAAC_CTA_GROUP");
  if (cg && (cg[0] == '1' || cg[0] == '2')) ctx->cta_group_override = cg[0] - '0';
  const char* pdl = getenv("ZSAAC_PDL");
  if (pdl && pdl[0] == '0') ctx->pdl_enabled = false;
  const char* solo = getenv("ZSAAC_SOLO");
  if (solo && (solo[0] == '0' || solo[0] == '1')) ctx->solo_override = solo[0] - '0';
  const char* boot = getenv("ZSAAC_BOOT");
  if (boot && (boot[0] == '0' || boot[0] == '1')) ctx->boot_override = boot[0] - '0';
  const char* coop_env = getenv("ZSAAC_COOPERATIVE");
  if (coop_env && coop_env[0] == '1') {
    int coop = 0;
    if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device) == cudaSuccess && coop)
      ctx->cooperative = true;
  }
  // the role code of a timed-out pipeline wait goes to mapped host memory, so that it can still
  // be read after the trap has poisoned the CUDA context
  e = cudaHostAlloc(&ctx->err_host, sizeof(int), cudaHostAllocMapped);
  if (e == cudaSuccess) {
    *ctx->err_host = 0;
    e = cudaHostGetDevicePointer(&ctx->err_flag, ctx->err_host, 0);
  }
  if (e != cudaSuccess) {
    if (ctx->err_host) cudaFreeHost(ctx->err_host);
    delete ctx;
    return fail(ZS_ERR_CUDA, "zs_create: mapped host allocation failed: %s", cudaGetErrorString(e));
  }
  e = cudaMalloc(&ctx->grid_cnt, 2 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemset(ctx->grid_cnt, 0, 2 * sizeof(unsigned long long));
  if (e != cudaSuccess) {
    cudaFreeHost(ctx->err_host);
    if (ctx->grid_cnt) cudaFree(ctx->grid_cnt);
    delete ctx;
    return fail(ZS_ERR_CUDA, "zs_create: counter allocation failed: %s", cudaGetErrorString(e));
  }
  *out = ctx;
  return ZS_OK;
}

int zs_destroy(zs_ctx* ctx) {
  if (!ctx) return ZS_OK;
  DeviceGuard guard(ctx->device);
  cudaFree(ctx->bank);
  cudaFree(ctx->q_ws);
  cudaFree(ctx->row_thr);
  cudaFree(ctx->boot);
  cudaFree(ctx->grid_cnt);
  cudaFree(ctx->part_scores);
  cudaFree(ctx->part_idx);
  cudaFreeHost(ctx->err_host);
  cudaFree(ctx->sync_cnt);
  cudaFree(ctx->exact_scores);
  cudaFree(ctx->exact_counter);
  cudaFree(ctx->memproj_partials);
  cudaFree(ctx->mp_bank);
  cudaFree(ctx->mp_bank_t);
  cudaFree(ctx->mp_q);
  cudaFree(ctx->mp_scores);
  cudaFree(ctx->mp_p);
  cudaFree(ctx->mp_partial);
  cudaFree(ctx->tgt_scores);
  cudaFree(ctx->tgt_cols);
  cudaFree(ctx->part_counts);
  if (ctx->prof_ev) {
    for (int i = 0; i < 2 * ZS_PROFILE_RING; ++i) cudaEventDestroy(ctx->prof_ev[i]);
    delete[] ctx->prof_ev;
  }
  delete ctx;
  return ZS_OK;
}

int zs_bank_alloc(zs_ctx* ctx, int64_t n_rows, int d) {
  if (!ctx) return fail(ZS_ERR_INVALID, "zs_bank_alloc: ctx is null");
  if (n_rows < 1 || n_rows > 0x7fffff00ll)
    return fail(ZS_ERR_INVALID, "zs_bank_alloc: n_rows=%lld outside [1, 2^31)", (long long)n_rows);
  if (d < ZS_DIM_MULTIPLE || d > ZS_MAX_DIM || d % ZS_DIM_MULTIPLE != 0)
    return fail(ZS_ERR_INVALID, "zs_bank_alloc: d=%d must be a multiple of %d in [%d, %d]", d,
                ZS_DIM_MULTIPLE, ZS_DIM_MULTIPLE, ZS_MAX_DIM);
  DeviceGuard guard(ctx->device);
  if (ctx->bank && (ctx->bank_rows != n_rows || ctx->bank_d != d)) {
    ZS_CUDA(cudaFree(ctx->bank));
    ctx->bank = nullptr;
    ctx->bank_rows = 0;
    ctx->bank_d = 0;
  }
  if (ctx->bank_d != d && ctx->q_ws) {  // query workspace is sized in rows of d
    ZS_CUDA(cudaFree(ctx->q_ws));
    ctx->q_ws = nullptr;
    ctx->q_ws_rows = 0;
  }
  if (!ctx->bank) {
    ZS_CUDA(cudaMalloc(&ctx->bank, static_cast<size_t>(n_rows) * d * sizeof(__nv_bfloat16)));
    ctx->bank_rows = n_rows;
    ctx->bank_d = d;
  }
  ctx->win_lo = ctx->win_rows = 0;
  int rc = encode_rows_map(ctx, &ctx->bank_map[0], ctx->bank, n_rows, d, zs::BLOCK_N);
  if (rc) return rc;
  return encode_rows_map(ctx, &ctx->bank_map[1], ctx->bank, n_rows, d, zs::BLOCK_N / 2);
}

int zs_bank_window(zs_ctx* ctx, int64_t row_lo, int64_t n_rows) {
  if (!ctx || !ctx->bank) return fail(ZS_ERR_STATE, "zs_bank_window: no bank");
  if (n_rows == 0 && row_lo == 0) {          // back to the whole bank
    ctx->win_lo = ctx->win_rows = 0;
    return ZS_OK;
  }
  if (row_lo < 0 || n_rows < 1 || row_lo + n_rows > ctx->bank_rows)
    return fail(ZS_ERR_INVALID, "zs_bank_window: rows [%lld, %lld) outside the bank of %lld rows",
                (long long)row_lo, (long long)(row_lo + n_rows), (long long)ctx->bank_rows);
  DeviceGuard guard(ctx->device);
  const __nv_bfloat16* base = ctx->bank + row_lo * ctx->bank_d;   // d % 64 == 0: 128-byte aligned rows
  CUtensorMap maps[2];
  int rc = encode_rows_map(ctx, &maps[0], base, n_rows, ctx->bank_d, zs::BLOCK_N);
  if (rc) return rc;
  rc = encode_rows_map(ctx, &maps[1], base, n_rows, ctx->bank_d, zs::BLOCK_N / 2);
  if (rc) return rc;
  ctx->win_map[0] = maps[0];
  ctx->win_map[1] = maps[1];
  ctx->win_lo = row_lo;
  ctx->win_rows = n_rows;
  return ZS_OK;
}

int zs_bank_upload(zs_ctx* ctx, const void* rows, int64_t n_rows, int64_t dst_row, int in_dtype,
                   int normalize, void* stream) {
  if (!ctx || !ctx->bank) return fail(ZS_ERR_STATE, "zs_bank_upload: call zs_bank_alloc first");
  if (!rows && n_rows > 0) return fail(ZS_ERR_INVALID, "zs_bank_upload: rows is null");
  if (n_rows < 0 || dst_row < 0 || dst_row + n_rows > ctx->bank_rows)
    return fail(ZS_ERR_INVALID, "zs_bank_upload: rows [%lld, %lld) outside the bank of %lld rows",
                (long long)dst_row, (long long)(dst_row + n_rows), (long long)ctx->bank_rows);
  if (in_dtype != ZS_F32 && in_dtype != ZS_BF16)
    return fail(ZS_ERR_INVALID, "zs_bank_upload: unknown dtype %d", in_dtype);
  if (n_rows == 0) return ZS_OK;
  if (!aligned16(rows)) return fail(ZS_ERR_INVALID, "zs_bank_upload: rows must be 16-byte aligned");
  DeviceGuard guard(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* dst = ctx->bank + dst_row * ctx->bank_d;
  if (in_dtype == ZS_F32) launch_normalize<float>(rows, dst, n_rows, n_rows, ctx->bank_d, normalize, st);
  else launch_normalize<__nv_bfloat16>(rows, dst, n_rows, n_rows, ctx->bank_d, normalize, st);
  ZS_CUDA(cudaGetLastError());
  ctx->launches += 1;
  return ZS_OK;
}

int64_t zs_bank_rows(const zs_ctx* ctx) { return ctx ? ctx->bank_rows : 0; }
int zs_bank_dim(const zs_ctx* ctx) { return ctx ? ctx->bank_d : 0; }

int zs_reserve(zs_ctx* ctx, int64_t Q, int k) {
  if (!ctx || !ctx->bank) return fail(ZS_ERR_STATE, "zs_reserve: no bank");
  if (Q < 1 || k < 1 || k > ZS_MAX_K) return fail(ZS_ERR_INVALID, "zs_reserve: Q=%lld k=%d", (long long)Q, k);
  DeviceGuard guard(ctx->device);
  return ensure_workspace(ctx, Q, std::min(k, (int)ZS_PASS_K));
}

int zs_plan(const zs_ctx* ctx, int64_t Q, int k, int* n_chunks, int* tiles_per_chunk, int* n_ctas) {
  if (!ctx || !ctx->bank) return fail(ZS_ERR_STATE, "zs_plan: no bank");
  if (Q < 1 || k < 1) return fail(ZS_ERR_INVALID, "zs_plan: Q=%lld k=%d", (long long)Q, k);
  const Plan pl = make_plan(ctx, Q, std::min(k, (int)ZS_PASS_K));
  if (n_chunks) *n_chunks = pl.chunks;
  if (tiles_per_chunk) *tiles_per_chunk = pl.tiles_per_chunk;
  if (n_ctas) *n_ctas = pl.ctas;
  return ZS_OK;
}

int zs_plan_dry(int sm_count, int64_t bank_rows, int64_t Q, int k, int cta_group, int* n_chunks,
                int* tiles_per_chunk, int* n_ctas, int* lockstep_window, int* cta_group_chosen) {
  if (sm_count < 1 || bank_rows < 1 || Q < 1 || k < 1 || k > ZS_MAX_K || cta_group < 0 || cta_group > 2)
    return fail(ZS_ERR_INVALID, "zs_plan_dry: sm_count=%d bank_rows=%lld Q=%lld k=%d cta_group=%d",
                sm_count, (long long)bank_rows, (long long)Q, k, cta_group);
  zs_ctx shape;                      // never touches a device: only the planner's inputs are set
  shape.sm_count = sm_count;
  shape.bank_rows = bank_rows;
  shape.cta_group_override = cta_group;
  const Plan pl = make_plan(&shape, Q, std::min(k, (int)ZS_PASS_K));
  if (n_chunks) *n_chunks = pl.chunks;
  if (tiles_per_chunk) *tiles_per_chunk = pl.tiles_per_chunk;
  if (n_ctas) *n_ctas = pl.ctas;
  if (lockstep_window) *lockstep_window = pl.sync_window;
  if (cta_group_chosen) *cta_group_chosen = pl.cg;
  return ZS_OK;
}

int64_t zs_launch_count(const zs_ctx* ctx) { return ctx ? ctx->launches : 0; }

int zs_kernel_error(const zs_ctx* ctx) { return (ctx && ctx->err_host) ? *ctx->err_host : 0; }

int zs_profile_enable(zs_ctx* ctx, int enable) {
  if (!ctx) return fail(ZS_ERR_INVALID, "zs_profile_enable: ctx is null");
  DeviceGuard guard(ctx->device);
  if (enable && !ctx->prof_ev) {
    ctx->prof_ev = new (std::nothrow) cudaEvent_t[2 * ZS_PROFILE_RING];
    if (!ctx->prof_ev) return fail(ZS_ERR_INVALID, "zs_profile_enable: out of host memory");
    for (int i = 0; i < 2 * ZS_PROFILE_RING; ++i) ZS_CUDA(cudaEventCreate(&ctx->prof_ev[i]));
  }
  ctx->profiling = enable != 0;
  ctx->prof_count = 0;
  return ZS_OK;
}

int zs_profile_read(zs_ctx* ctx, float* ms_out, int max_entries, int* n_entries) {
  if (!ctx || !ms_out || !n_entries || max_entries < 0)
    return fail(ZS_ERR_INVALID, "zs_profile_read: bad argument");
  *n_entries = 0;
  if (!ctx->prof_ev) return ZS_OK;
  DeviceGuard guard(ctx->device);
  const int have = std::min(ctx->prof_count, (int)ZS_PROFILE_RING);
  const int n = std::min(have, max_entries);
  const int first = ctx->prof_count - have;   // oldest launch still in the ring
  for (int j = 0; j < n; ++j) {
    const int slot = (first + j) % ZS_PROFILE_RING;
    ZS_CUDA(cudaEventSynchronize(ctx->prof_ev[2 * slot + 1]));
    ZS_CUDA(cudaEventElapsedTime(&ms_out[j], ctx->prof_ev[2 * slot], ctx->prof_ev[2 * slot + 1]));
  }
  *n_entries = n;
  return ZS_OK;
}

// A pipeline wait that timed out traps and poisons the CUDA context; the role code survives in
// mapped host memory.  Every search-type entry point reports it instead of enqueueing more work.
#define ZS_CHECK_KERNEL_FLAG(ctx, fn)                                                            \
  do {                                                                                           \
    if ((ctx)->err_host && *(ctx)->err_host != 0)                                                \
      return fail(ZS_ERR_KERNEL, fn ": an earlier kernel timed out in pipeline role %d and "     \
                  "trapped; the CUDA context of this process is unusable", *(ctx)->err_host);    \
  } while (0)

namespace {

// Small searches run as ONE cooperative launch (in-kernel query cast + in-kernel merge).
constexpr int64_t kSoloMaxQueries = 4096;

bool solo_wanted(zs_ctx* ctx, int64_t Q, const Plan& pl) {
  if (ctx->solo_state < 0 || ctx->solo_override == 0) return false;
  if (ctx->profiling && ctx->solo_override != 1) {
    // (profiling brackets the fused kernel alone; solo is still profiled as a whole when forced)
  }
  if (pl.ctas > ctx->sm_count) return false;
  if (ctx->solo_override == 1) return true;
  // One 128-row tile against a long bank is an HBM-bound stream: there three launches chained
  // by programmatic dependent launch overlap the next search's prologue with this one's tail
  // and win by ~10 % (profiles/r02/bench_small_variants.jsonl); everything else up to
  // kSoloMaxQueries is faster, or as fast, as one cooperative launch.
  if (pl.cg == 1 && pl.m_tiles == 1) return false;
  return Q <= kSoloMaxQueries;
}

// One pass of a search: the k_pass best elements that come strictly after the element
// (bound_scores[q], bound_idx[q]) of every query (no bound on the first pass), written to
// out_scores / out_indices with `out_stride` elements between rows.
int search_pass(zs_ctx* ctx, const void* queries, int64_t Q, int q_dtype, int k_pass, int normalize,
                bool cast_queries, const int64_t* self_index, int64_t index_offset,
                const float* bound_scores, const int64_t* bound_idx, float* out_scores,
                int64_t* out_indices, int64_t out_stride, const Plan& pl, cudaStream_t st) {
  const int64_t q_pad = padded_query_rows(Q);
  bool solo = solo_wanted(ctx, Q, pl);
  ctx->epoch += 1;
  if (ctx->epoch == 0) {   // 2^32 passes later: stale entries could alias a live epoch, so wipe them once
    ZS_CUDA(cudaMemsetAsync(ctx->row_thr, 0, static_cast<size_t>(ctx->q_ws_rows) * sizeof(unsigned long long), st));
    ZS_CUDA(cudaMemsetAsync(ctx->boot, 0, static_cast<size_t>(ctx->q_ws_rows) * zs::BOOT_SLOTS * sizeof(unsigned long long), st));
    ctx->epoch = 1;
  }
  if (pl.sync_window > 0) {   // before the cast kernel, so that cast -> fused kernel stay adjacent
    const size_t n_cnt = static_cast<size_t>(pl.max_iters) * pl.windows_per_unit;
    ZS_CUDA(cudaMemsetAsync(ctx->sync_cnt, 0, n_cnt * sizeof(unsigned int), st));
  }

  zs::SimTopkParams p{};
  p.Q = static_cast<int>(Q);
  p.n_bank = static_cast<int>(search_rows(ctx));
  p.num_k_blocks = ctx->bank_d / zs::BLOCK_K;
  p.num_m_tiles = pl.m_tiles;
  p.num_n_tiles = pl.n_tiles;
  p.tiles_per_chunk = pl.tiles_per_chunk;
  p.num_chunks = pl.chunks;
  p.k = k_pass;
  p.self_index = reinterpret_cast<const long long*>(self_index);
  // with a window the kernel sees rows [win_lo, win_lo + win_rows) as columns 0 .. win_rows - 1
  const bool windowed = ctx->win_rows > 0;
  index_offset += windowed ? ctx->win_lo : 0;
  const CUtensorMap* wmap1 = windowed ? &ctx->win_map[0] : nullptr;
  const CUtensorMap* wmap2 = windowed ? &ctx->win_map[1] : nullptr;
  p.index_offset = index_offset;
  p.part_scores = ctx->part_scores;
  p.part_idx = ctx->part_idx;
  p.dump = nullptr;
  p.err_flag = ctx->err_flag;
  p.trace = ctx->trace;
  const char* share_env = getenv("ZSAAC_SHARE_THR");   // tuning hook: 0 = every unit warms up alone
  p.row_thr = (share_env && share_env[0] == '0') ? nullptr : ctx->row_thr;
  p.epoch = ctx->epoch;
  const int n_lists = pl.chunks * zs::EPI_HALVES;
  // (not for a single 128-row tile: that search is HBM-bound, its epilogue has slack anyway)
  const bool hbm_stream = pl.cg == 1 && pl.m_tiles == 1;
  if (ctx->boot_override != 0 && p.row_thr != nullptr && bound_scores == nullptr && n_lists >= k_pass &&
      (!hbm_stream || ctx->boot_override == 1)) {
    p.boot = ctx->boot;
    p.boot_slots = std::min(n_lists, zs::BOOT_SLOTS);
  }
  if (bound_scores != nullptr) {
    p.bound_scores = bound_scores;
    p.bound_idx = reinterpret_cast<const long long*>(bound_idx);
    p.bound_stride = out_stride;
  }
  if (pl.sync_window > 0) {
    p.sync_cnt = ctx->sync_cnt;
    p.sync_window = pl.sync_window;
    p.windows_per_unit = pl.windows_per_unit;
    p.max_iters = pl.max_iters;
  }
  p.q_ws = ctx->q_ws;
  p.q_pad = static_cast<int>(q_pad);
  p.d = ctx->bank_d;

  CUtensorMap qmap;
  int rc = encode_rows_map(ctx, &qmap, ctx->q_ws, q_pad, ctx->bank_d, zs::BLOCK_M);
  if (rc) return rc;

  if (solo) {
    p.solo = 1;
    p.q_src = cast_queries ? queries : nullptr;
    p.q_src_bf16 = (q_dtype == ZS_BF16) ? 1 : 0;
    p.q_normalize = normalize;
    p.grid_cnt = ctx->grid_cnt;
    p.cast_target = ctx->cnt_base[0] + (cast_queries ? pl.ctas : 0);
    p.done_target = ctx->cnt_base[1] + pl.ctas;
    p.out_scores = out_scores;
    p.out_idx = reinterpret_cast<long long*>(out_indices);
    p.out_stride = out_stride;
    rc = (pl.cg == 2) ? dispatch_simtopk<2>(ctx, qmap, p, pl.ctas, false, st, wmap2)
                      : dispatch_simtopk<1>(ctx, qmap, p, pl.ctas, false, st, wmap1);
    if (rc == ZS_OK) {
      ctx->solo_state = 1;
      ctx->cnt_base[0] = p.cast_target;
      ctx->cnt_base[1] = p.done_target;
      return ZS_OK;
    }
    if (rc != ZS_ERR_STATE) return rc;
    ctx->solo_state = -1;        // launch refused: three launches from now on
    p.solo = 0;
    p.q_src = nullptr;
    p.grid_cnt = nullptr;
    p.out_scores = nullptr;
    p.out_idx = nullptr;
  }

  if (cast_queries) {
    if (q_dtype == ZS_F32)
      launch_normalize<float>(queries, ctx->q_ws, Q, q_pad, ctx->bank_d, normalize, st);
    else
      launch_normalize<__nv_bfloat16>(queries, ctx->q_ws, Q, q_pad, ctx->bank_d, normalize, st);
    ZS_CUDA(cudaGetLastError());
    ctx->launches += 1;
  }
  // programmatic dependent launch: cast kernel -> fused kernel -> merge (not while profiling,
  // the timing events would sit between the kernels)
  ctx->pdl_next = cast_queries && ctx->pdl_enabled && !ctx->profiling;
  rc = (pl.cg == 2) ? dispatch_simtopk<2>(ctx, qmap, p, pl.ctas, false, st, wmap2)
                    : dispatch_simtopk<1>(ctx, qmap, p, pl.ctas, false, st, wmap1);
  ctx->pdl_next = false;
  if (rc) return rc;
  ZS_CUDA(launch_merge<int>(ctx->part_scores, ctx->part_idx, n_lists, Q * k_pass, Q * k_pass, Q, k_pass,
                            index_offset, out_scores, reinterpret_cast<long long*>(out_indices), out_stride,
                            /*pdl=*/ctx->pdl_enabled && !ctx->profiling, st));
  ctx->launches += 1;
  return ZS_OK;
}

}  // namespace

int zs_search(zs_ctx* ctx, const void* queries, int64_t Q, int q_dtype, int k, int normalize_queries,
              const int64_t* self_index, int64_t index_offset, float* out_scores,
              int64_t* out_indices, void* stream) {
  if (!ctx) return fail(ZS_ERR_INVALID, "zs_search: ctx is null");
  ZS_CHECK_KERNEL_FLAG(ctx, "zs_search");
  if (!ctx->bank) return fail(ZS_ERR_STATE, "zs_search: no bank uploaded");
  if (Q < 0 || Q > 0x7fffff00ll) return fail(ZS_ERR_INVALID, "zs_search: Q=%lld", (long long)Q);
  if (q_dtype != ZS_F32 && q_dtype != ZS_BF16)
    return fail(ZS_ERR_INVALID, "zs_search: unknown query dtype %d", q_dtype);
  const int64_t avail = search_rows(ctx) - (self_index ? 1 : 0);
  if (k < 1 || k > ZS_MAX_K || k > avail)
    return fail(ZS_ERR_INVALID,
                "zs_search: selected index k out of range (k=%d, bank rows=%lld%s, max k=%d)", k,
                (long long)search_rows(ctx), self_index ? " minus the excluded row" : "", ZS_MAX_K);
  if (Q == 0) return ZS_OK;
  if (!queries || !out_scores || !out_indices)
    return fail(ZS_ERR_INVALID, "zs_search: null queries / output pointer");
  if (!aligned16(queries)) return fail(ZS_ERR_INVALID, "zs_search: queries must be 16-byte aligned");
  DeviceGuard guard(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int k_first = std::min(k, (int)ZS_PASS_K);
  int rc = ensure_workspace(ctx, Q, k_first);
  if (rc) return rc;
  const Plan pl = make_plan(ctx, Q, k_first);
  // k <= 32: one pass.  k > 32: passes of 32; pass p continues strictly after the last element
  // pass p-1 wrote (read from the caller's output arrays, where the merge has just put it).
  for (int done = 0; done < k; done += ZS_PASS_K) {
    const int k_pass = std::min(k - done, (int)ZS_PASS_K);
    const float* bs = done ? out_scores + (done - 1) : nullptr;
    const int64_t* bi = done ? out_indices + (done - 1) : nullptr;
    rc = search_pass(ctx, queries, Q, q_dtype, k_pass, normalize_queries, /*cast_queries=*/done == 0,
                     self_index, index_offset, bs, bi, out_scores + done, out_indices + done, k, pl, st);
    if (rc) return rc;
  }
  return ZS_OK;
}

int zs_rank_count(zs_ctx* ctx, const void* queries, int64_t Q, int q_dtype, int normalize_queries,
                  const int64_t* target_index, int n_targets, int64_t index_offset,
                  float* out_target_scores, int64_t* out_ranks, void* stream) {
  if (!ctx) return fail(ZS_ERR_INVALID, "zs_rank_count: ctx is null");
  ZS_CHECK_KERNEL_FLAG(ctx, "zs_rank_count");
  if (!ctx->bank) return fail(ZS_ERR_STATE, "zs_rank_count: no bank uploaded");
  if (Q < 0 || Q > 0x7fffff00ll) return fail(ZS_ERR_INVALID, "zs_rank_count: Q=%lld", (long long)Q);
  if (n_targets < 1 || n_targets > ZS_MAX_TARGETS)
    return fail(ZS_ERR_INVALID, "zs_rank_count: n_targets=%d outside [1, %d]", n_targets, ZS_MAX_TARGETS);
  if (q_dtype != ZS_F32 && q_dtype != ZS_BF16)
    return fail(ZS_ERR_INVALID, "zs_rank_count: unknown query dtype %d", q_dtype);
  if (Q == 0) return ZS_OK;
  if (!queries || !target_index || !out_ranks)
    return fail(ZS_ERR_INVALID, "zs_rank_count: null pointer");
  if (!aligned16(queries)) return fail(ZS_ERR_INVALID, "zs_rank_count: queries must be 16-byte aligned");
  DeviceGuard guard(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  WholeBank whole(ctx);
  int rc = ensure_workspace(ctx, Q, 1);
  if (rc) return rc;
  const Plan pl = make_plan(ctx, Q, 1);
  const int64_t n_pairs = Q * n_targets;
  const int n_lists = pl.chunks * zs::EPI_HALVES;
  if (n_pairs > ctx->tgt_elems) {
    if (ctx->tgt_scores) { ZS_CUDA(cudaFree(ctx->tgt_scores)); ctx->tgt_scores = nullptr; }
    if (ctx->tgt_cols) { ZS_CUDA(cudaFree(ctx->tgt_cols)); ctx->tgt_cols = nullptr; }
    ctx->tgt_elems = 0;
    ZS_CUDA(cudaMalloc(&ctx->tgt_scores, static_cast<size_t>(n_pairs) * sizeof(float)));
    ZS_CUDA(cudaMalloc(&ctx->tgt_cols, static_cast<size_t>(n_pairs) * sizeof(int)));
    ctx->tgt_elems = n_pairs;
  }
  if (n_pairs * n_lists > ctx->count_elems) {
    if (ctx->part_counts) { ZS_CUDA(cudaFree(ctx->part_counts)); ctx->part_counts = nullptr; }
    ctx->count_elems = 0;
    ZS_CUDA(cudaMalloc(&ctx->part_counts, static_cast<size_t>(n_pairs) * n_lists * sizeof(int)));
    ctx->count_elems = n_pairs * n_lists;
  }
  CUtensorMap qmap;
  rc = prepare_queries(ctx, queries, Q, q_dtype, normalize_queries, &qmap, st);
  if (rc) return rc;
  {
    const int64_t blocks = (n_pairs * 32 + 255) / 256;
    zs::target_scores_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(
        ctx->q_ws, ctx->bank, reinterpret_cast<const long long*>(target_index), n_pairs, n_targets,
        ctx->bank_d, index_offset, ctx->bank_rows, ctx->tgt_scores, ctx->tgt_cols);
    ZS_CUDA(cudaGetLastError());
    ctx->launches += 1;
  }
  zs::SimTopkParams p{};
  p.Q = static_cast<int>(Q);
  p.n_bank = static_cast<int>(ctx->bank_rows);
  p.num_k_blocks = ctx->bank_d / zs::BLOCK_K;
  p.num_m_tiles = pl.m_tiles;
  p.num_n_tiles = pl.n_tiles;
  p.tiles_per_chunk = pl.tiles_per_chunk;
  p.num_chunks = pl.chunks;
  p.k = 1;
  p.index_offset = index_offset;
  p.err_flag = ctx->err_flag;
  p.tgt_scores = ctx->tgt_scores;
  p.tgt_cols = ctx->tgt_cols;
  p.n_targets = n_targets;
  p.part_counts = ctx->part_counts;
  rc = (pl.cg == 2) ? dispatch_simtopk<2>(ctx, qmap, p, pl.ctas, false, st)
                    : dispatch_simtopk<1>(ctx, qmap, p, pl.ctas, false, st);
  if (rc) return rc;
  {
    const int64_t blocks = (n_pairs + 255) / 256;
    zs::sum_counts_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(
        ctx->part_counts, n_lists, n_pairs, ctx->tgt_cols, reinterpret_cast<long long*>(out_ranks));
    ZS_CUDA(cudaGetLastError());
    ctx->launches += 1;
  }
  if (out_target_scores)
    ZS_CUDA(cudaMemcpyAsync(out_target_scores, ctx->tgt_scores, static_cast<size_t>(n_pairs) * sizeof(float),
                            cudaMemcpyDeviceToDevice, st));
  return ZS_OK;
}

int zs_memory_project(zs_ctx* ctx, const float* queries, int64_t Q, const float* bank, int64_t n_rows,
                      int d, float temperature, float* out, void* stream) {
  if (!ctx) return fail(ZS_ERR_INVALID, "zs_memory_project: ctx is null");
  if (Q < 0 || n_rows < 1) return fail(ZS_ERR_INVALID, "zs_memory_project: Q=%lld n_rows=%lld",
                                       (long long)Q, (long long)n_rows);
  if (d < 4 || d % 4 != 0 || d > zs::MEMPROJ_MAX_D)
    return fail(ZS_ERR_INVALID, "zs_memory_project: d=%d must be a multiple of 4, at most %d", d,
                zs::MEMPROJ_MAX_D);
  if (Q == 0) return ZS_OK;
  if (!queries || !bank || !out) return fail(ZS_ERR_INVALID, "zs_memory_project: null pointer");
  if (!aligned16(queries) || !aligned16(bank) || !aligned16(out))
    return fail(ZS_ERR_INVALID, "zs_memory_project: queries, bank and out must be 16-byte aligned");
  DeviceGuard guard(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // queries per pass over the bank.  Measured (profiles/r01/bench_map2memory.jsonl): against a
  // 400 k-row bank one query per pass streams at 5.7 TB/s (0.29 ms), two per pass take 0.69 ms —
  // the second accumulator set (157 registers) costs more than the second pass; the 4-query
  // instantiation (225 registers) was slower still.  Small banks are launch-bound (three kernels
  // per pass), so there two queries share a pass.
  const int QB = (static_cast<double>(n_rows) * d * sizeof(float) >= 256e6) ? 1 : 2;
  constexpr int QB_MAX = 2;
  const int blocks = ctx->sm_count * 2;
  const int64_t need = static_cast<int64_t>(blocks) * QB_MAX * zs::memproj_partial_stride(d);
  if (need > ctx->memproj_elems) {
    if (ctx->memproj_partials) { ZS_CUDA(cudaFree(ctx->memproj_partials)); ctx->memproj_partials = nullptr; }
    ctx->memproj_elems = 0;
    ZS_CUDA(cudaMalloc(&ctx->memproj_partials, static_cast<size_t>(need) * sizeof(float)));
    ctx->memproj_elems = need;
  }
  for (int64_t q0 = 0; q0 < Q; q0 += QB) {
    const int nq = static_cast<int>(std::min<int64_t>(QB, Q - q0));
    const float* qptr = queries + q0 * d;
    // 1 or 2 queries use the narrower instantiations (fewer accumulator registers, more warps in flight)
    if (nq == 1)
      zs::memproj_stream_kernel<1><<<blocks, zs::MEMPROJ_THREADS, 0, st>>>(qptr, bank, n_rows, d, nq,
                                                                          temperature, ctx->memproj_partials);
    else
      zs::memproj_stream_kernel<2><<<blocks, zs::MEMPROJ_THREADS, 0, st>>>(qptr, bank, n_rows, d, nq,
                                                                          temperature, ctx->memproj_partials);
    ZS_CUDA(cudaGetLastError());
    const int qb = nq;
    const dim3 cgrid((d + zs::MEMPROJ_COMBINE_COLS - 1) / zs::MEMPROJ_COMBINE_COLS, nq);
    zs::memproj_combine_kernel<<<cgrid, 256, 0, st>>>(ctx->memproj_partials, blocks, qb, d, out + q0 * d);
    zs::memproj_finalize_kernel<<<nq, 256, 0, st>>>(out + q0 * d, d);
    ZS_CUDA(cudaGetLastError());
    ctx->launches += 3;
  }
  return ZS_OK;
}

namespace {

constexpr int kMemprojMaxKBlocksPerUnit = 512;     // 32,768 terms per tensor-memory accumulation

// C[M, N] (+ K chunks) = A[M, K] . B[N, K]^T on the fused kernel's pipeline in DUMP mode.
//   a      bf16 [a_rows_pad, K] (rows beyond `m` are padding, a_rows_pad a multiple of 256)
//   b      bf16 [n, K]
//   out    fp32 [k_chunks, m, n]: k_chunks = 1 unless split_k
int run_dump_gemm(zs_ctx* ctx, const __nv_bfloat16* a, int64_t a_rows_pad, int64_t m, const __nv_bfloat16* b,
                  int64_t n, int64_t K, bool split_k, float* out, int* k_chunks_out, cudaStream_t st) {
  const int cg = m > zs::BLOCK_M ? 2 : 1;
  const int workers = std::max(1, ctx->sm_count / cg);
  zs::SimTopkParams p{};
  p.Q = static_cast<int>(m);
  p.n_bank = static_cast<int>(n);
  p.num_k_blocks = static_cast<int>(K / zs::BLOCK_K);
  p.num_m_tiles = static_cast<int>((m + zs::BLOCK_M * cg - 1) / (zs::BLOCK_M * cg));
  p.num_n_tiles = static_cast<int>((n + zs::BLOCK_N - 1) / zs::BLOCK_N);
  p.k = 1;
  p.dump = out;
  p.err_flag = ctx->err_flag;
  int k_chunks = 1;
  if (split_k) {
    const int by_work = (workers + p.num_m_tiles * p.num_n_tiles - 1) / (p.num_m_tiles * p.num_n_tiles);
    const int by_precision = (p.num_k_blocks + kMemprojMaxKBlocksPerUnit - 1) / kMemprojMaxKBlocksPerUnit;
    k_chunks = std::max(1, std::min(p.num_k_blocks, std::max(by_work, by_precision)));
    p.kb_per_unit = (p.num_k_blocks + k_chunks - 1) / k_chunks;
    k_chunks = (p.num_k_blocks + p.kb_per_unit - 1) / p.kb_per_unit;
    p.tiles_per_chunk = 1;
    p.num_chunks = k_chunks * p.num_n_tiles;
  } else {
    p.tiles_per_chunk = std::max<int>(1, static_cast<int>((static_cast<int64_t>(p.num_n_tiles) * p.num_m_tiles + workers - 1) / workers));
    p.num_chunks = (p.num_n_tiles + p.tiles_per_chunk - 1) / p.tiles_per_chunk;
  }
  if (k_chunks_out) *k_chunks_out = k_chunks;
  const int64_t units = static_cast<int64_t>(p.num_m_tiles) * p.num_chunks;
  const int ctas = static_cast<int>(std::min<int64_t>(units, workers)) * cg;
  CUtensorMap amap, bmap;
  int rc = encode_rows_map(ctx, &amap, a, a_rows_pad, K, zs::BLOCK_M);
  if (rc) return rc;
  rc = encode_rows_map(ctx, &bmap, b, n, K, zs::BLOCK_N / cg);
  if (rc) return rc;
  return cg == 2 ? dispatch_simtopk<2>(ctx, amap, p, ctas, true, st, &bmap)
                 : dispatch_simtopk<1>(ctx, amap, p, ctas, true, st, &bmap);
}

}  // namespace

int zs_memory_bank_prepare(zs_ctx* ctx, const float* bank, int64_t n_rows, int d, void* stream) {
  if (!ctx) return fail(ZS_ERR_INVALID, "zs_memory_bank_prepare: ctx is null");
  if (n_rows < 1 || n_rows > 0x3fffff00ll)
    return fail(ZS_ERR_INVALID, "zs_memory_bank_prepare: n_rows=%lld", (long long)n_rows);
  if (d < ZS_DIM_MULTIPLE || d % ZS_DIM_MULTIPLE != 0 || d > ZS_MAX_DIM)
    return fail(ZS_ERR_INVALID, "zs_memory_bank_prepare: d=%d must be a multiple of %d in [%d, %d]", d,
                ZS_DIM_MULTIPLE, ZS_DIM_MULTIPLE, ZS_MAX_DIM);
  if (!bank) return fail(ZS_ERR_INVALID, "zs_memory_bank_prepare: bank is null");
  DeviceGuard guard(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t n_pad = (n_rows + zs::BLOCK_K - 1) / zs::BLOCK_K * zs::BLOCK_K;
  if (ctx->mp_rows != n_rows || ctx->mp_d != d) {
    if (ctx->mp_bank) { ZS_CUDA(cudaFree(ctx->mp_bank)); ctx->mp_bank = nullptr; }
    if (ctx->mp_bank_t) { ZS_CUDA(cudaFree(ctx->mp_bank_t)); ctx->mp_bank_t = nullptr; }
    if (ctx->mp_scores) { ZS_CUDA(cudaFree(ctx->mp_scores)); ctx->mp_scores = nullptr; }
    if (ctx->mp_p) { ZS_CUDA(cudaFree(ctx->mp_p)); ctx->mp_p = nullptr; }
    if (ctx->mp_q) { ZS_CUDA(cudaFree(ctx->mp_q)); ctx->mp_q = nullptr; }
    ctx->mp_rows = 0;
    ctx->mp_q_rows = 0;
    ZS_CUDA(cudaMalloc(&ctx->mp_bank, static_cast<size_t>(n_rows) * 3 * d * sizeof(__nv_bfloat16)));
    ZS_CUDA(cudaMalloc(&ctx->mp_bank_t, static_cast<size_t>(d) * 3 * n_pad * sizeof(__nv_bfloat16)));
    ctx->mp_rows = n_rows;
    ctx->mp_pad = n_pad;
    ctx->mp_d = d;
  }
  zs::split_rows_kernel<<<static_cast<unsigned>((n_rows * 32 + 255) / 256), 256, 0, st>>>(
      bank, ctx->mp_bank, n_rows, n_rows, d, zs::SPLIT_B_SIDE);
  const dim3 tgrid(static_cast<unsigned>(n_pad / 32), static_cast<unsigned>((d + 31) / 32));
  zs::transpose_split_kernel<<<tgrid, 256, 0, st>>>(bank, ctx->mp_bank_t, n_rows, n_pad, d);
  ZS_CUDA(cudaGetLastError());
  ctx->launches += 2;
  return ZS_OK;
}

int zs_memory_project_batched(zs_ctx* ctx, const float* queries, int64_t Q, float temperature, float* out,
                              void* stream) {
  if (!ctx) return fail(ZS_ERR_INVALID, "zs_memory_project_batched: ctx is null");
  ZS_CHECK_KERNEL_FLAG(ctx, "zs_memory_project_batched");
  if (!ctx->mp_bank) return fail(ZS_ERR_STATE, "zs_memory_project_batched: call zs_memory_bank_prepare first");
  if (Q < 0 || Q > 0x3fffff00ll) return fail(ZS_ERR_INVALID, "zs_memory_project_batched: Q=%lld", (long long)Q);
  if (Q == 0) return ZS_OK;
  if (!queries || !out) return fail(ZS_ERR_INVALID, "zs_memory_project_batched: null pointer");
  DeviceGuard guard(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int d = ctx->mp_d;
  const int64_t n = ctx->mp_rows, n_pad = ctx->mp_pad;
  const int64_t q_pad = padded_query_rows(Q);
  if (q_pad > ctx->mp_q_rows) {
    if (ctx->mp_q) { ZS_CUDA(cudaFree(ctx->mp_q)); ctx->mp_q = nullptr; }
    if (ctx->mp_scores) { ZS_CUDA(cudaFree(ctx->mp_scores)); ctx->mp_scores = nullptr; }
    if (ctx->mp_p) { ZS_CUDA(cudaFree(ctx->mp_p)); ctx->mp_p = nullptr; }
    ctx->mp_q_rows = 0;
    ZS_CUDA(cudaMalloc(&ctx->mp_q, static_cast<size_t>(q_pad) * 3 * d * sizeof(__nv_bfloat16)));
    ZS_CUDA(cudaMalloc(&ctx->mp_scores, static_cast<size_t>(q_pad) * n * sizeof(float)));
    ZS_CUDA(cudaMalloc(&ctx->mp_p, static_cast<size_t>(q_pad) * 3 * n_pad * sizeof(__nv_bfloat16)));
    // padding rows of P'' are never written again: zero them once
    ZS_CUDA(cudaMemsetAsync(ctx->mp_p, 0, static_cast<size_t>(q_pad) * 3 * n_pad * sizeof(__nv_bfloat16), st));
    ctx->mp_q_rows = q_pad;
  }
  // 1. queries -> A side [hi | lo | hi] (padding rows zero)
  zs::split_rows_kernel<<<static_cast<unsigned>((q_pad * 32 + 255) / 256), 256, 0, st>>>(
      queries, ctx->mp_q, Q, q_pad, d, zs::SPLIT_A_SIDE);
  ZS_CUDA(cudaGetLastError());
  // 2. S = Q . B^T  (K = 3d)
  int rc = run_dump_gemm(ctx, ctx->mp_q, q_pad, Q, ctx->mp_bank, n, 3ll * d, false, ctx->mp_scores, nullptr, st);
  if (rc) return rc;
  // 3. P = softmax(t S), as the A side of the second contraction
  zs::softmax_split_kernel<<<static_cast<unsigned>(Q), zs::SOFTMAX_THREADS, 0, st>>>(
      ctx->mp_scores, ctx->mp_p, n, n_pad, temperature);
  ZS_CUDA(cudaGetLastError());
  // 4. O = P . B  (K = 3 n_pad, split into chunks) -> partial sums
  const int n_tiles_out = (d + zs::BLOCK_N - 1) / zs::BLOCK_N;
  const int m_tiles = static_cast<int>((Q + (Q > zs::BLOCK_M ? 2 : 1) * zs::BLOCK_M - 1) / ((Q > zs::BLOCK_M ? 2 : 1) * zs::BLOCK_M));
  const int64_t kb_total = 3 * n_pad / zs::BLOCK_K;
  const int64_t max_chunks = std::max<int64_t>((kb_total + kMemprojMaxKBlocksPerUnit - 1) / kMemprojMaxKBlocksPerUnit,
                                               (ctx->sm_count + m_tiles * n_tiles_out - 1) / (m_tiles * n_tiles_out)) + 1;
  const int64_t need = max_chunks * Q * d;
  if (need > ctx->mp_partial_elems) {
    if (ctx->mp_partial) { ZS_CUDA(cudaFree(ctx->mp_partial)); ctx->mp_partial = nullptr; ctx->mp_partial_elems = 0; }
    ZS_CUDA(cudaMalloc(&ctx->mp_partial, static_cast<size_t>(need) * sizeof(float)));
    ctx->mp_partial_elems = need;
  }
  int k_chunks = 1;
  rc = run_dump_gemm(ctx, ctx->mp_p, q_pad, Q, ctx->mp_bank_t, d, 3 * n_pad, true, ctx->mp_partial, &k_chunks, st);
  if (rc) return rc;
  if (static_cast<int64_t>(k_chunks) * Q * d > ctx->mp_partial_elems)
    return fail(ZS_ERR_STATE, "zs_memory_project_batched: partial buffer too small (%d chunks)", k_chunks);
  // 5. sum the K chunks, L2-normalise
  zs::memproj_reduce_kernel<<<static_cast<unsigned>(Q), 256, 0, st>>>(ctx->mp_partial, k_chunks, Q, d, out);
  ZS_CUDA(cudaGetLastError());
  ctx->launches += 3;
  return ZS_OK;
}

int zs_merge(zs_ctx* ctx, const float* scores, const int64_t* indices, int S, int64_t score_stride,
             int64_t index_stride, int64_t Q, int k, float* out_scores, int64_t* out_indices,
             void* stream) {
  if (!ctx) return fail(ZS_ERR_INVALID, "zs_merge: ctx is null");
  if (S < 1 || S > zs::MERGE_MAX_LISTS)
    return fail(ZS_ERR_INVALID, "zs_merge: S=%d outside [1, %d]", S, zs::MERGE_MAX_LISTS);
  if (k < 1 || Q < 0 || score_stride < Q * k || index_stride < Q * k)
    return fail(ZS_ERR_INVALID, "zs_merge: Q=%lld k=%d score_stride=%lld index_stride=%lld",
                (long long)Q, k, (long long)score_stride, (long long)index_stride);
  if (Q == 0) return ZS_OK;
  if (!scores || !indices || !out_scores || !out_indices)
    return fail(ZS_ERR_INVALID, "zs_merge: null pointer");
  DeviceGuard guard(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ZS_CUDA(launch_merge<long long>(scores, reinterpret_cast<const long long*>(indices), S, score_stride,
                                  index_stride, Q, k, 0ll, out_scores,
                                  reinterpret_cast<long long*>(out_indices), k, /*pdl=*/false, st));
  ctx->launches += 1;
  return ZS_OK;
}

int zs_rescore_f32(zs_ctx* ctx, const float* queries, int64_t Q, int normalize, const float* bank,
                   int64_t n_rows, int d, int64_t index_offset, const int64_t* candidates, int kc,
                   int k, float* out_scores, int64_t* out_indices, void* stream) {
  if (!ctx) return fail(ZS_ERR_INVALID, "zs_rescore_f32: ctx is null");
  if (Q < 0 || n_rows < 1) return fail(ZS_ERR_INVALID, "zs_rescore_f32: Q=%lld n_rows=%lld",
                                       (long long)Q, (long long)n_rows);
  if (d < 4 || d % 4 != 0) return fail(ZS_ERR_INVALID, "zs_rescore_f32: d=%d must be a multiple of 4", d);
  if (kc < 1 || kc > zs::RESCORE_MAX_CAND || k < 1 || k > kc)
    return fail(ZS_ERR_INVALID, "zs_rescore_f32: need 1 <= k <= kc <= %d (k=%d, kc=%d)",
                zs::RESCORE_MAX_CAND, k, kc);
  if (Q == 0) return ZS_OK;
  if (!queries || !bank || !candidates || !out_scores || !out_indices)
    return fail(ZS_ERR_INVALID, "zs_rescore_f32: null pointer");
  if (!aligned16(queries) || !aligned16(bank))
    return fail(ZS_ERR_INVALID, "zs_rescore_f32: queries and bank must be 16-byte aligned");
  DeviceGuard guard(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int P = 2;
  while (P < kc) P <<= 1;
  const size_t smem = static_cast<size_t>(P) * (sizeof(long long) + sizeof(float));   // <= 24 KiB
  zs::rescore_f32_kernel<<<static_cast<unsigned>(Q), zs::RESCORE_THREADS, smem, st>>>(
      queries, bank, n_rows, d, normalize, reinterpret_cast<const long long*>(candidates), kc, k,
      index_offset, out_scores, reinterpret_cast<long long*>(out_indices), P);
  ZS_CUDA(cudaGetLastError());
  ctx->launches += 1;
  return ZS_OK;
}

int zs_gather_rows_f32(zs_ctx* ctx, const float* src, int64_t n_src_rows, int d,
                       const int64_t* indices, int64_t n_idx, float* out, void* stream) {
  if (!ctx) return fail(ZS_ERR_INVALID, "zs_gather_rows_f32: ctx is null");
  if (d < 4 || d % 4 != 0) return fail(ZS_ERR_INVALID, "zs_gather_rows_f32: d=%d must be a multiple of 4", d);
  if (n_idx < 0 || n_src_rows < 0) return fail(ZS_ERR_INVALID, "zs_gather_rows_f32: negative size");
  if (n_idx == 0) return ZS_OK;
  if (!src || !indices || !out) return fail(ZS_ERR_INVALID, "zs_gather_rows_f32: null pointer");
  if (!aligned16(src) || !aligned16(out))
    return fail(ZS_ERR_INVALID, "zs_gather_rows_f32: src and out must be 16-byte aligned");
  DeviceGuard guard(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t blocks = (n_idx * 32 + 255) / 256;
  zs::gather_rows_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(
      src, n_src_rows, d, reinterpret_cast<const long long*>(indices), n_idx, out);
  ZS_CUDA(cudaGetLastError());
  ctx->launches += 1;
  return ZS_OK;
}

int zs_normalize_rows_f32(zs_ctx* ctx, const float* in, float* out, int64_t n_rows, int d,
                          void* stream) {
  if (!ctx) return fail(ZS_ERR_INVALID, "zs_normalize_rows_f32: ctx is null");
  if (d < 4 || d % 4 != 0 || n_rows < 0)
    return fail(ZS_ERR_INVALID, "zs_normalize_rows_f32: n_rows=%lld d=%d (d must be a multiple of 4)",
                (long long)n_rows, d);
  if (n_rows == 0) return ZS_OK;
  if (!in || !out) return fail(ZS_ERR_INVALID, "zs_normalize_rows_f32: null pointer");
  if (!aligned16(in) || !aligned16(out))
    return fail(ZS_ERR_INVALID, "zs_normalize_rows_f32: in and out must be 16-byte aligned");
  DeviceGuard guard(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t blocks = (n_rows * 32 + 255) / 256;
  zs::normalize_cast_kernel<float, float><<<static_cast<unsigned>(blocks), 256, 0, st>>>(in, out, n_rows, n_rows, d, 1);
  ZS_CUDA(cudaGetLastError());
  ctx->launches += 1;
  return ZS_OK;
}

namespace {

constexpr int64_t kExactMaxScores = 1ll << 30;    // 4 GiB of fp32 scores

void launch_exact_scores(zs_ctx* ctx, const float* queries, int64_t Q, const float* bank, int64_t n_rows,
                         int d, int normalize, int k_fused, const int64_t* self_index,
                         int64_t index_offset, float* out_scores, int64_t* out_indices, cudaStream_t st) {
  const bool small = n_rows <= zs::EXACT_SMALL_BANK;
  const int rb = small ? 1 : 4;
  const int rows_per_block = (zs::EXACT_THREADS / 32) * rb;
  const unsigned blocks = static_cast<unsigned>((n_rows + rows_per_block - 1) / rows_per_block);
  const long long* self = reinterpret_cast<const long long*>(self_index);
  long long* out_i = reinterpret_cast<long long*>(out_indices);
#define ZS_EXACT_LAUNCH(RB, D1024)                                                                 \
  zs::exact_scores_kernel<RB, D1024><<<blocks, zs::EXACT_THREADS, 0, st>>>(                          \
      queries, static_cast<int>(Q), bank, n_rows, d, normalize, ctx->exact_scores, ctx->exact_counter, \
      k_fused, self, index_offset, out_scores, out_i)
  if (small && d <= 1024) ZS_EXACT_LAUNCH(1, true);
  else if (small) ZS_EXACT_LAUNCH(1, false);
  else ZS_EXACT_LAUNCH(4, false);
#undef ZS_EXACT_LAUNCH
}

int exact_prepare(zs_ctx* ctx, const char* fn, const float* queries, int64_t Q, const float* bank,
                  int64_t n_rows, int d) {
  if (Q < 0 || n_rows < 1) return fail(ZS_ERR_INVALID, "%s: Q=%lld n_rows=%lld", fn, (long long)Q, (long long)n_rows);
  if (d < 4 || d % 4 != 0) return fail(ZS_ERR_INVALID, "%s: d=%d must be a multiple of 4", fn, d);
  if (Q > 0x7fffff00ll || Q * n_rows > kExactMaxScores)
    return fail(ZS_ERR_INVALID, "%s: %lld x %lld scores exceed the fp32 scratch limit; this entry point "
                "is for small banks (use zs_search / zs_rank_count)", fn, (long long)Q, (long long)n_rows);
  if (Q == 0) return ZS_OK;
  if (!queries || !bank) return fail(ZS_ERR_INVALID, "%s: null pointer", fn);
  if (!aligned16(queries) || !aligned16(bank))
    return fail(ZS_ERR_INVALID, "%s: queries and bank must be 16-byte aligned", fn);
  if (Q * n_rows > ctx->exact_elems) {
    if (ctx->exact_scores) { ZS_CUDA(cudaFree(ctx->exact_scores)); ctx->exact_scores = nullptr; ctx->exact_elems = 0; }
    ZS_CUDA(cudaMalloc(&ctx->exact_scores, static_cast<size_t>(Q * n_rows) * sizeof(float)));
    ctx->exact_elems = Q * n_rows;
  }
  if (!ctx->exact_counter) {
    ZS_CUDA(cudaMalloc(&ctx->exact_counter, sizeof(unsigned int)));
    ZS_CUDA(cudaMemset(ctx->exact_counter, 0, sizeof(unsigned int)));
  }
  return ZS_OK;
}

}  // namespace

int zs_exact_topk_f32(zs_ctx* ctx, const float* queries, int64_t Q, const float* bank, int64_t n_rows,
                      int d, int normalize, int k, const int64_t* self_index, int64_t index_offset,
                      float* out_scores, int64_t* out_indices, void* stream) {
  if (!ctx) return fail(ZS_ERR_INVALID, "zs_exact_topk_f32: ctx is null");
  const int64_t avail = n_rows - (self_index ? 1 : 0);
  if (k < 1 || k > avail)
    return fail(ZS_ERR_INVALID, "zs_exact_topk_f32: selected index k out of range (k=%d, bank rows=%lld%s)",
                k, (long long)n_rows, self_index ? " minus the excluded row" : "");
  DeviceGuard guard(ctx->device);
  int rc = exact_prepare(ctx, "zs_exact_topk_f32", queries, Q, bank, n_rows, d);
  if (rc || Q == 0) return rc;
  if (!out_scores || !out_indices) return fail(ZS_ERR_INVALID, "zs_exact_topk_f32: null output pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool fused = Q <= zs::EXACT_FUSED_MAX_Q;     // one launch: the last block selects
  launch_exact_scores(ctx, queries, Q, bank, n_rows, d, normalize, fused ? k : 0, self_index, index_offset,
                      out_scores, out_indices, st);
  ZS_CUDA(cudaGetLastError());
  ctx->launches += 1;
  if (!fused) {
    zs::exact_topk_kernel<<<static_cast<unsigned>((Q * 32 + 255) / 256), 256, 0, st>>>(
        ctx->exact_scores, static_cast<int>(Q), n_rows, k, reinterpret_cast<const long long*>(self_index),
        index_offset, out_scores, reinterpret_cast<long long*>(out_indices));
    ZS_CUDA(cudaGetLastError());
    ctx->launches += 1;
  }
  return ZS_OK;
}

int zs_exact_rank_f32(zs_ctx* ctx, const float* queries, int64_t Q, const float* bank, int64_t n_rows,
                      int d, int normalize, const int64_t* target_index, int n_targets,
                      int64_t index_offset, float* out_target_scores, int64_t* out_ranks, void* stream) {
  if (!ctx) return fail(ZS_ERR_INVALID, "zs_exact_rank_f32: ctx is null");
  if (n_targets < 1) return fail(ZS_ERR_INVALID, "zs_exact_rank_f32: n_targets=%d", n_targets);
  DeviceGuard guard(ctx->device);
  int rc = exact_prepare(ctx, "zs_exact_rank_f32", queries, Q, bank, n_rows, d);
  if (rc || Q == 0) return rc;
  if (!target_index || !out_ranks) return fail(ZS_ERR_INVALID, "zs_exact_rank_f32: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  launch_exact_scores(ctx, queries, Q, bank, n_rows, d, normalize, 0, nullptr, 0, nullptr, nullptr, st);
  ZS_CUDA(cudaGetLastError());
  const int64_t n_pairs = Q * n_targets;
  zs::exact_rank_kernel<<<static_cast<unsigned>((n_pairs * 32 + 255) / 256), 256, 0, st>>>(
      ctx->exact_scores, static_cast<int>(Q), n_rows, reinterpret_cast<const long long*>(target_index),
      n_targets, index_offset, out_target_scores, reinterpret_cast<long long*>(out_ranks));
  ZS_CUDA(cudaGetLastError());
  ctx->launches += 2;
  return ZS_OK;
}

int zs_debug_trace(zs_ctx* ctx, void* stamps) {
  if (!ctx) return fail(ZS_ERR_INVALID, "zs_debug_trace: ctx is null");
  ctx->trace = static_cast<unsigned long long*>(stamps);
  return ZS_OK;
}

int zs_debug_scores(zs_ctx* ctx, const void* queries, int64_t Q, int q_dtype, int normalize_queries,
                    float* out_scores, void* stream) {
  if (!ctx || !ctx->bank) return fail(ZS_ERR_STATE, "zs_debug_scores: no bank uploaded");
  if (Q < 1 || !queries || !out_scores) return fail(ZS_ERR_INVALID, "zs_debug_scores: bad argument");
  if (q_dtype != ZS_F32 && q_dtype != ZS_BF16)
    return fail(ZS_ERR_INVALID, "zs_debug_scores: unknown query dtype %d", q_dtype);
  DeviceGuard guard(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  WholeBank whole(ctx);
  int rc = ensure_workspace(ctx, Q, 1);
  if (rc) return rc;
  CUtensorMap qmap;
  rc = prepare_queries(ctx, queries, Q, q_dtype, normalize_queries, &qmap, st);
  if (rc) return rc;
  const Plan pl = make_plan(ctx, Q, 1);
  zs::SimTopkParams p{};
  p.Q = static_cast<int>(Q);
  p.n_bank = static_cast<int>(ctx->bank_rows);
  p.num_k_blocks = ctx->bank_d / zs::BLOCK_K;
  p.num_m_tiles = pl.m_tiles;
  p.num_n_tiles = pl.n_tiles;
  p.tiles_per_chunk = pl.tiles_per_chunk;
  p.num_chunks = pl.chunks;
  p.k = 1;
  p.dump = out_scores;
  p.err_flag = ctx->err_flag;
  return (pl.cg == 2) ? dispatch_simtopk<2>(ctx, qmap, p, pl.ctas, true, st)
                      : dispatch_simtopk<1>(ctx, qmap, p, pl.ctas, true, st);
}

}  // extern "C"
