/*
 * zsaac.h — C ABI of the B200-native related-caption retrieval library (libzsaac_b200.so).
 *
 * The reference (XinMing0411/zero-shot-AAC) has no FFI of its own: its hot path is four
 * module-level Python functions that call torch.  Each entry point below replaces the torch
 * calls made at one reference call site; the Python host (zero-shot-aac_b200/) keeps the
 * reference's function names and signatures and binds these symbols through ctypes
 * (INTEGRATION.md shows the stub a maintainer of the reference would add).
 *
 * Conventions
 *   - every function returns 0 (ZS_OK) or a negative zs_status; the message of the last
 *     failure on the calling thread is available from zs_last_error()
 *   - all tensor pointers are DEVICE pointers on the context's device unless the name says host
 *   - work is enqueued on the cudaStream_t passed as `stream` (NULL = legacy default stream);
 *     no entry point synchronises the host except where stated (workspace growth)
 *   - there is no CPU fallback: on a machine without an sm_100 GPU zs_create fails
 *   - a context is not thread-safe and owns one set of workspaces: use one context per host
 *     thread / per concurrent stream (zs_last_error() is thread-local)
 *   - plain C types only; no torch / C++ types cross this boundary
 */
#ifndef ZSAAC_H_
#define ZSAAC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZS_ABI_VERSION 1

typedef struct zs_ctx zs_ctx;

typedef enum zs_status {
  ZS_OK = 0,
  ZS_ERR_INVALID = -1,     /* bad argument (k out of range, d not supported, null pointer ...) */
  ZS_ERR_CUDA = -2,        /* a CUDA runtime / driver call failed                              */
  ZS_ERR_NO_DEVICE = -3,   /* no sm_100 device: this library has no fallback path              */
  ZS_ERR_STATE = -4,       /* call order violated (search before the bank was uploaded ...)    */
  ZS_ERR_KERNEL = -5       /* a kernel reported a pipeline time-out through the device flag    */
} zs_status;

typedef enum zs_dtype {
  ZS_F32 = 0,
  ZS_BF16 = 1
} zs_dtype;

/* Limits of the fused kernel. */
#define ZS_PASS_K 32       /* top-k list length held in registers per query row and pass     */
#define ZS_MAX_K 1024      /* k > ZS_PASS_K runs ceil(k / 32) passes over the bank           */
#define ZS_DIM_MULTIPLE 64 /* embedding dim must be a multiple of the 128-byte bf16 K block  */
#define ZS_MAX_DIM 4096

/* ---- context ---------------------------------------------------------------------------- */

/* Create a context bound to CUDA device `device`.  Owns the bf16 bank copy and the search
 * workspaces.  One context per GPU; multi-GPU runs use one process (and one context) per GPU. */
int zs_create(zs_ctx** out, int device);
int zs_destroy(zs_ctx* ctx);

/* ---- bank (replaces load_data's `F.normalize(torch.cat(...).to('cuda'), dim=-1)`,
 *      reference data_handing/embeddings_related_generator.py:15-17, _wavcaps.py:16-18) ---- */

/* Allocate (or re-allocate) library-owned storage for `n_rows` bank rows of dimension `d`
 * in bf16.  Synchronises the device if storage has to be (re)allocated. */
int zs_bank_alloc(zs_ctx* ctx, int64_t n_rows, int d);

/* Convert `n_rows` rows starting at `rows` (row-major, leading dimension d, dtype `in_dtype`)
 * to bf16 and store them at bank rows [dst_row, dst_row + n_rows).  With normalize != 0 every
 * row is scaled by 1 / max(||row||_2, 1e-12) in fp32 before the cast (F.normalize semantics,
 * eps = 1e-12).  May be called repeatedly to fill the bank block by block. */
int zs_bank_upload(zs_ctx* ctx, const void* rows, int64_t n_rows, int64_t dst_row,
                   int in_dtype, int normalize, void* stream);

/* Restrict zs_search to bank rows [row_lo, row_lo + n_rows) of the stored bank ("search
 * window"); returned indices stay global (index_offset + row within the stored bank).  Host-side
 * only (two tensor maps are re-encoded): cheap enough to call before every search.  Used by the
 * multi-GPU path, where every rank stores a superset of its shard and the shard boundaries
 * follow the measured speed of the GPUs (sharded.py); scores do not depend on the window, so
 * any partition of the bank into windows merges to the same bits.  (0, 0) lifts the window;
 * zs_bank_alloc lifts it too.  zs_rank_count / zs_debug_scores always see the whole bank. */
int zs_bank_window(zs_ctx* ctx, int64_t row_lo, int64_t n_rows);

/* out[i, :] = in[i, :] / max(||in[i, :]||_2, 1e-12), fp32 in and out (in-place allowed): the
 * fp32 unit-row bank load_data returns to its caller (reference
 * embeddings_related_generator.py:17).  d must be a multiple of 4.  Needs no bank. */
int zs_normalize_rows_f32(zs_ctx* ctx, const float* in, float* out, int64_t n_rows, int d,
                          void* stream);

/* Number of rows / dimension currently allocated (0 when no bank). */
int64_t zs_bank_rows(const zs_ctx* ctx);
int zs_bank_dim(const zs_ctx* ctx);

/* ---- search (replaces process_data's per-item
 *      `torch.cosine_similarity(F.normalize(q), bank).topk(k)`,
 *      reference data_handing/embeddings_related_generator.py:21-22, and
 *      `prefix @ bank.T -> softmax -> topk` of utils.py:133-135) ------------------------- */

/* Pre-size the workspaces for searches of up to Q queries with top-k `k` so that zs_search
 * does not allocate (and therefore never synchronises). */
int zs_reserve(zs_ctx* ctx, int64_t Q, int k);

/* Fused similarity + top-k of Q queries against the uploaded bank.
 *   queries          [Q, d] row-major, dtype q_dtype
 *   normalize_queries != 0: each query is L2-normalised (eps 1e-12) in fp32 before the bf16 cast
 *   self_index       nullable [Q] int64: GLOBAL bank index that query i must not return
 *                    (self-exclusion); entries < 0 mean "no exclusion"
 *   index_offset     added to every returned index (first global row of this bank shard)
 *   out_scores       [Q, k] fp32, descending; ties broken by ascending index
 *   out_indices      [Q, k] int64 global indices
 * Requires 1 <= k <= min(ZS_MAX_K, bank rows [- 1 with self exclusion]).
 * Deterministic: the result is the exact top-k of the bf16 scores under (score desc, index asc)
 * whatever the launch geometry or timing (the work units exchange admission thresholds while
 * they run, which changes only how much each unit contributes, never the merged result).
 * One kernel (129 ... 4,096 queries: cast and merge run inside the fused kernel) or three (cast,
 * fused similarity/top-k, merge) are enqueued on `stream` per pass of 32; the context's
 * workspaces are in use until they finish, so searches on one context must not overlap.
 * Returns ZS_ERR_KERNEL (and enqueues nothing) once a kernel of this context has reported a
 * pipeline time-out (zs_kernel_error). */
int zs_search(zs_ctx* ctx, const void* queries, int64_t Q, int q_dtype, int k,
              int normalize_queries, const int64_t* self_index, int64_t index_offset,
              float* out_scores, int64_t* out_indices, void* stream);

/* Rank of ground-truth items (replaces the per-query `np.argsort(cos_sim(...))[::-1]` +
 * `np.where(inds == i)` of the retrieval metrics a2t / t2a, reference
 * retrieval/tools/utils.py:182-192,232-237).  Same GEMM as zs_search with a counting epilogue:
 *   target_index  [Q, n_targets] int64 global bank indices (< 0 = unused slot)
 *   out_ranks     [Q, n_targets] int64: number of bank rows, other than the target itself, whose
 *                 similarity to query q is STRICTLY greater than the target's (0 = retrieved
 *                 first); -1 for unused slots / targets outside this bank
 *   out_target_scores  nullable [Q, n_targets] fp32 similarity of each target (+inf if unused)
 * Exact score ties are resolved in the target's favour (np.argsort leaves them unspecified). */
#define ZS_MAX_TARGETS 8
int zs_rank_count(zs_ctx* ctx, const void* queries, int64_t Q, int q_dtype, int normalize_queries,
                  const int64_t* target_index, int n_targets, int64_t index_offset,
                  float* out_target_scores, int64_t* out_ranks, void* stream);

/* Softmax-weighted projection onto the memory bank (replaces map2memory, reference
 * predict_prompt.py:23-29: sim = q @ B.T; p = softmax(100 * sim); out = p @ B; out /= ||out||).
 *   queries [Q, d] fp32, bank [n_rows, d] fp32 (the caller's text_features tensor, read in place:
 *   no bank upload needed), out [Q, d] fp32.  d a multiple of 4, <= 1024.  One streaming pass
 *   over the bank per query on banks of 256 MB or more, per pair of queries on smaller ones (the
 *   reference calls it with one audio embedding); batches go to zs_memory_project_batched. */
int zs_memory_project(zs_ctx* ctx, const float* queries, int64_t Q, const float* bank, int64_t n_rows,
                      int d, float temperature, float* out, void* stream);

/* Exact fp32 top-k for SMALL banks (replaces `prefix @ bank.T -> softmax -> topk` of
 * utils.py:133-135 — softmax is monotone — and the zero-shot classification step
 * `audio_emb @ text_embeds.t() -> argmax`, retrieval/zero_shot_classification.py:97-103).
 * Scores are fp32 FMA dot products (cosine with normalize != 0), i.e. the reference's own
 * arithmetic class: indices equal the reference's except at fp32 rounding ties.  No bf16 bank:
 *   queries [Q, d] fp32, bank [n_rows, d] fp32 (the caller's tensor, read in place), d % 4 == 0
 *   self_index nullable [Q] global index to skip; index_offset = global index of bank row 0
 *   out_scores [Q, k] fp32 descending, out_indices [Q, k] int64, ties by ascending index
 * Up to 64 queries are ONE launch (the last block to finish selects); Q * n_rows <= 2^30. */
int zs_exact_topk_f32(zs_ctx* ctx, const float* queries, int64_t Q, const float* bank, int64_t n_rows,
                      int d, int normalize, int k, const int64_t* self_index, int64_t index_offset,
                      float* out_scores, int64_t* out_indices, void* stream);

/* Exact fp32 rank of ground-truth items (the retrieval metrics a2t / t2a at their real sizes,
 * reference retrieval/tools/utils.py:182-192,232-237: cos_sim -> argsort -> np.where).
 *   target_index [Q, n_targets] int64 global bank indices (< 0 = unused slot)
 *   out_ranks    [Q, n_targets] int64: number of bank rows ranking BEFORE the target under
 *                (score desc, index asc) — the target's position in zs_exact_topk_f32's order,
 *                so tied ground truths get distinct positions; -1 for unused / absent targets
 *   out_target_scores nullable [Q, n_targets] fp32 (+inf if unused)
 * Same limits as zs_exact_topk_f32. */
int zs_exact_rank_f32(zs_ctx* ctx, const float* queries, int64_t Q, const float* bank, int64_t n_rows,
                      int d, int normalize, const int64_t* target_index, int n_targets,
                      int64_t index_offset, float* out_target_scores, int64_t* out_ranks, void* stream);

/* The same projection for BATCHES of queries on the tensor cores (Q >= 8 is where it beats the
 * streaming pass per query above).  Both contractions — S = Q B^T and O = softmax(t S) B — run on
 * the fused kernel's TMA + tcgen05 pipeline with every operand split into two bf16 terms
 * (x = hi + lo; three products per contraction via a 3x longer K), which keeps the scores within
 * ~1e-6 of fp32; the second contraction is split along K = bank rows into chunks of <= 32,768
 * terms that are summed in fp32.
 *   zs_memory_bank_prepare    splits the fp32 bank [n_rows, d] (d a multiple of 64) into the two
 *                             operand layouts (12 n_rows d bytes, library-owned); once per bank
 *   zs_memory_project_batched queries [Q, d] fp32 -> out [Q, d] fp32 unit rows
 * Scratch per call (library-owned, grown on demand): 4 Q n_rows bytes of scores and 6 Q n_rows
 * bytes of weights. */
int zs_memory_bank_prepare(zs_ctx* ctx, const float* bank, int64_t n_rows, int d, void* stream);
int zs_memory_project_batched(zs_ctx* ctx, const float* queries, int64_t Q, float temperature, float* out,
                              void* stream);

/* k-way merge of S sorted top-k lists per query (shard-local results gathered from S GPUs, or
 * bank chunks) under the total order (score desc, index asc).
 *   scores  list s of query q starts at scores  + s*score_stride + q*k   (float elements)
 *   indices list s of query q starts at indices + s*index_stride + q*k   (int64 elements)
 * (two strides so that one all-gathered byte buffer holding [scores | indices] per rank can be
 * merged in place).  Output [Q, k].  Stand-alone: needs no bank. */
int zs_merge(zs_ctx* ctx, const float* scores, const int64_t* indices, int S, int64_t score_stride,
             int64_t index_stride, int64_t Q, int k, float* out_scores, int64_t* out_indices,
             void* stream);

/* fp32 re-scoring of search candidates: the fused kernel ranks with bf16-rounded operands, so
 * bank rows whose scores differ by less than ~1e-4 may swap places against the reference's fp32
 * `torch.cosine_similarity(text_embs, valid_text_embs).topk(topnumber)`
 * (data_handing/embeddings_related_generator.py:22).  Search for kc = k + margin candidates,
 * then call this with the caller's fp32 bank (the tensor load_data returned): every candidate is
 * re-scored in fp32 — cosine with normalize != 0 (q.b / (max(|q|,1e-12) max(|b|,1e-12))), the raw
 * dot product otherwise — and the k best under (score desc, index asc) are returned.
 *   queries     [Q, d] fp32 (raw: normalisation happens here), bank [n_rows, d] fp32
 *   candidates  [Q, kc] int64 global indices (index_offset = global index of bank row 0;
 *               entries < 0 or outside the bank are ignored), k <= kc <= 2048
 *   out_scores  [Q, k] fp32, out_indices [Q, k] int64 (-1 where fewer than k candidates exist)
 * Stand-alone: needs no bf16 bank. */
int zs_rescore_f32(zs_ctx* ctx, const float* queries, int64_t Q, int normalize, const float* bank,
                   int64_t n_rows, int d, int64_t index_offset, const int64_t* candidates, int kc,
                   int k, float* out_scores, int64_t* out_indices, void* stream);

/* Gather rows: out[i, :] = src[indices[i], :] for fp32 [*, d] row-major src (replaces
 * `valid_text_embs[ids]`, reference embeddings_related_generator.py:23). */
int zs_gather_rows_f32(zs_ctx* ctx, const float* src, int64_t n_src_rows, int d,
                       const int64_t* indices, int64_t n_idx, float* out, void* stream);

/* ---- introspection ---------------------------------------------------------------------- */

/* Launch geometry chosen for a (Q, k) search on the current bank: number of bank chunks a
 * query tile is split into, 256-row bank tiles per chunk, CTAs launched.  For tests / bench. */
int zs_plan(const zs_ctx* ctx, int64_t Q, int k, int* n_chunks, int* tiles_per_chunk, int* n_ctas);

/* The same planner for a hypothetical device and bank (no context, no GPU needed): sm_count SMs,
 * bank_rows rows, cta_group 0 = choose per search, 1 / 2 = pinned.  lockstep_window receives the
 * bank tiles per lock-step window (0 = lock-step off for this shape), cta_group_chosen the CTA
 * group the plan uses (1 = single CTAs, 2 = cta_group::2 pairs).  For host-side tests. */
int zs_plan_dry(int sm_count, int64_t bank_rows, int64_t Q, int k, int cta_group, int* n_chunks,
                int* tiles_per_chunk, int* n_ctas, int* lockstep_window, int* cta_group_chosen);

/* Kernels launched by this context since creation (bench.py's gpu_launches). */
int64_t zs_launch_count(const zs_ctx* ctx);

/* Role code written by a fused-kernel pipeline wait that timed out (the kernel then traps, which
 * poisons the CUDA context — ZS_ERR_KERNEL territory): 0 none, 101 TMA producer, 102 MMA issuer
 * waiting for operands, 103 MMA issuer waiting for an accumulator, 104 epilogue.  The code lives
 * in mapped host memory, so it stays readable after the trap. */
int zs_kernel_error(const zs_ctx* ctx);

/* Per-launch device timing of the fused similarity+top-k kernel.  With enable != 0 every
 * zs_search brackets that kernel with CUDA events on the caller's stream (a ring of
 * ZS_PROFILE_RING launches; enabling resets the ring).  zs_profile_read synchronises on the
 * recorded events and returns the durations in milliseconds, oldest first. */
#define ZS_PROFILE_RING 256
int zs_profile_enable(zs_ctx* ctx, int enable);
int zs_profile_read(zs_ctx* ctx, float* ms_out, int max_entries, int* n_entries);

/* Test hook: run the same TMA + tcgen05 pipeline but write the full fp32 score matrix
 * [Q, n_bank] (row-major) instead of the top-k.  Small shapes only. */
int zs_debug_scores(zs_ctx* ctx, const void* queries, int64_t Q, int q_dtype,
                    int normalize_queries, float* out_scores, void* stream);

/* Test / tuning hook: while `stamps` is non-NULL every CTA of the fused kernel writes uint64
 * %globaltimer values (ns) to stamps[cta*16 + i]: 0 entry, 1 barriers+TMEM ready, 2 first bank
 * tile's MMAs complete, 3 last tile of the last unit scanned, 4 lists written, 5 exit (after the
 * in-kernel merge in single-launch mode); 6 / 7 are clock64() at entry / exit (SM cycles, so
 * (7-6)/(5-0) is the SM clock in GHz during the kernel); single-launch mode only: 8 this CTA's
 * query rows cast, 9 all queries cast (first load may start), 10 all CTAs' lists written (merge
 * starts); 11-15 unused.  The device buffer [n_ctas, 16] is caller-owned; pass NULL to switch
 * tracing off. */
int zs_debug_trace(zs_ctx* ctx, void* stamps);

/* Name of the dominant kernel (for ncu -k) and ABI version. */
const char* zs_kernel_name(void);
int zs_abi_version(void);

/* Thread-local message describing the last non-zero status returned on this thread. */
const char* zs_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* ZSAAC_H_ */
