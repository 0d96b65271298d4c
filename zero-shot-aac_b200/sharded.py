"""Bank row-sharded over the GPUs of one box: one process per GPU, torch.distributed for plumbing.

The reference's hot path is single-process / single-GPU; sharding is this repo's
extension for banks that outgrow one GPU's time budget (BASELINE config 4: 10 M rows).

  rank r holds bank rows [lo_r, hi_r) (contiguous; ceil(N/G) rows each, or sized by
  shard_weights = measured speed of each GPU) as bf16
  every rank gets the full query batch
  each rank: fused similarity + top-k over its shard  -> (score, GLOBAL index) [Q, k]
  ONE all-gather of the packed [scores | indices] byte buffer over NCCL / NVLink
  each rank: k-way merge of the G lists under (score desc, index asc)

Scores of a (query, bank row) pair do not depend on the shard layout (same K-loop order in the
kernel), so the merged result is bit-identical to the single-GPU result.

Adaptive shard boundaries (overlap > 0, SearchPipeline(balance_every=n)): every step ends in an
exchange, so the slowest GPU of the box sets the pace, and the GPUs of one box differ by 3-7 %
under the power cap.  Each rank therefore STORES a superset of its shard (its rows plus `overlap`
of a shard on either side: 180 GB of HBM make that free), searches only a window of it
(zs_bank_window), and the windows' boundaries follow the measured time per row of every rank:
because scores do not depend on the partition, moving a boundary never changes a result bit.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


SHARD_ALIGN = 256       # weighted shards start on a bank-tile boundary of the fused kernel


def shard_bounds(n_rows: int, world: int, weights: Optional[Sequence[float]] = None
                 ) -> List[Tuple[int, int]]:
    """Row range [lo, hi) per rank: contiguous blocks of ceil(n_rows / world) rows, or — with
    `weights` (one positive number per rank, e.g. measured rows per millisecond) — blocks
    proportional to the weights, cut at multiples of SHARD_ALIGN rows."""
    if n_rows < 0 or world < 1:
        raise ValueError(f"shard_bounds(n_rows={n_rows}, world={world})")
    if weights is None:
        per = -(-n_rows // world)
        return [(min(r * per, n_rows), min((r + 1) * per, n_rows)) for r in range(world)]
    w = [float(x) for x in weights]
    if len(w) != world or min(w) <= 0 or not all(x == x and x != float("inf") for x in w):
        raise ValueError(f"shard weights must be {world} positive finite numbers, got {weights}")
    total = sum(w)
    align = SHARD_ALIGN if n_rows >= 64 * SHARD_ALIGN * world else 1    # small banks: exact cuts
    cuts, acc = [0], 0.0
    for r in range(world - 1):
        acc += w[r]
        cut = int(round(n_rows * acc / total / align)) * align
        # keep at least one row for this rank and for every rank after it where the bank allows
        cut = max(cut, min(cuts[-1] + 1, n_rows))
        cut = min(cut, max(n_rows - (world - 1 - r), cuts[-1]))
        cuts.append(cut)
    cuts.append(n_rows)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def rebalanced_bounds(bounds: Sequence[Tuple[int, int]], times_ms: Sequence[float],
                      stores: Sequence[Tuple[int, int]], *, damping: float = 0.6,
                      align: int = SHARD_ALIGN, min_rows: int = 1) -> List[Tuple[int, int]]:
    """New contiguous row ranges after one control step.

    bounds    current [lo, hi) of every rank (contiguous, covering [0, N))
    times_ms  what every rank took for its current range (same list on every rank)
    stores    rows [s_lo, s_hi) every rank holds in memory: rank r can only be given rows inside
    Rank r's speed is rows / ms; the target size is N x its share of the summed speeds, approached
    by `damping` per step; cuts are rounded to `align` rows and clamped so that every range stays
    inside its rank's stored rows and keeps at least `min_rows`.  Deterministic: every rank
    computes the same answer from the same inputs.  Invalid timings leave the bounds unchanged."""
    world = len(bounds)
    n_rows = bounds[-1][1]
    sizes = [hi - lo for lo, hi in bounds]
    t = [float(x) for x in times_ms]
    if len(t) != world or any(not (x > 0.0) or x == float("inf") for x in t):
        return [tuple(b) for b in bounds]
    speed = [n / x for n, x in zip(sizes, t)]
    total = sum(speed)
    target = [n_rows * v / total for v in speed]
    new = [n + damping * (g - n) for n, g in zip(sizes, target)]
    cuts, acc = [0], 0.0
    for r in range(1, world):
        acc += new[r - 1]
        cut = int(round(acc / align)) * align
        # rank r-1 holds rows below stores[r-1][1], rank r rows from stores[r][0] on
        cut = min(cut, stores[r - 1][1], n_rows - (world - r) * min_rows)
        cut = max(cut, stores[r][0], cuts[-1] + min_rows)
        cuts.append(cut)
    cuts.append(n_rows)
    out = [(cuts[r], cuts[r + 1]) for r in range(world)]
    ok = all(hi - lo >= min_rows and lo >= stores[r][0] and hi <= stores[r][1] for r, (lo, hi) in enumerate(out))
    return out if ok else [tuple(b) for b in bounds]


def _default_local_bank(rows: int, dim: int, device, index_offset: int):
    from .retrieval import RelatedBank   # needs CUDA + the native library: no fallback
    return RelatedBank(rows, dim, device=device, index_offset=index_offset)


class ShardedRelatedBank:
    """Per-rank handle of a bank sharded by rows over `group` (default: the world group).

    local_bank_factory(rows, dim, device, index_offset) builds the object that holds this rank's
    shard; the default is the CUDA RelatedBank.  (The hook exists so the host logic — bounds,
    packing, gather layout — can be exercised with the gloo backend on CPU in tests.)
    """

    def __init__(self, n_rows: int, dim: int, *, device=None, group=None,
                 local_bank_factory: Optional[Callable] = None,
                 shard_weights: Optional[Sequence[float]] = None, overlap: float = 0.0):
        if not dist.is_initialized():
            raise RuntimeError("ShardedRelatedBank needs an initialised torch.distributed process group")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.n_rows = int(n_rows)
        self.dim = int(dim)
        # shard_weights (identical on every rank): relative speed of the ranks' GPUs.  Every search
        # ends in an all-gather, so the slowest GPU sets the pace; the GPUs of one box differ by
        # several per cent under the power cap, and shards sized by measured speed even that out.
        self.base_bounds = shard_bounds(self.n_rows, self.world, shard_weights)
        if min(hi - lo for lo, hi in self.base_bounds) < 1:
            raise ValueError(f"bank of {n_rows} rows cannot be sharded over {self.world} ranks")
        # overlap: every rank stores this fraction of a shard beyond either end of its own, so
        # that set_bounds() can move the boundaries later without moving a byte
        if not 0.0 <= overlap <= 1.0:
            raise ValueError(f"overlap must be a fraction of a shard in [0, 1], got {overlap}")
        per = -(-self.n_rows // self.world)
        margin = -(-int(overlap * per) // SHARD_ALIGN) * SHARD_ALIGN if overlap > 0 else 0
        self.stores = [(max(0, lo - margin), min(self.n_rows, hi + margin)) for lo, hi in self.base_bounds]
        self.store_lo, self.store_hi = self.stores[self.rank]
        self.bounds = [tuple(b) for b in self.base_bounds]
        self.lo, self.hi = self.bounds[self.rank]
        factory = local_bank_factory or _default_local_bank
        self.local = factory(self.store_hi - self.store_lo, self.dim, device, self.store_lo)
        self.device = self.local.device
        self.adaptive = margin > 0
        self.mpr_sum, self.mpr_n = 0.0, 0      # running mean of this rank's milliseconds per bank row
        self._cpu_group = None
        if self.adaptive:
            self.local.window(self.lo - self.store_lo, self.hi - self.lo)

    # ---------------------------------------------------------------- adaptive boundaries
    def set_bounds(self, bounds: Sequence[Tuple[int, int]]) -> None:
        """Make [lo_r, hi_r) the rows rank r searches from now on (same list on every rank:
        contiguous, covering the bank, inside every rank's stored rows).  Host-side only."""
        bounds = [(int(lo), int(hi)) for lo, hi in bounds]
        if len(bounds) != self.world or bounds[0][0] != 0 or bounds[-1][1] != self.n_rows or \
                any(bounds[r][1] != bounds[r + 1][0] for r in range(self.world - 1)):
            raise ValueError(f"bounds must be {self.world} contiguous ranges covering [0, {self.n_rows})")
        for r, (lo, hi) in enumerate(bounds):
            if hi - lo < 1 or lo < self.stores[r][0] or hi > self.stores[r][1]:
                raise ValueError(f"rank {r}: rows [{lo}, {hi}) are not inside its stored rows "
                                 f"[{self.stores[r][0]}, {self.stores[r][1]})")
        self.bounds = bounds
        self.lo, self.hi = bounds[self.rank]
        if self.adaptive:
            self.local.window(self.lo - self.store_lo, self.hi - self.lo)

    def cpu_group(self):
        """A gloo group over the same ranks: the timings that steer the boundaries are exchanged
        on the host, beside the GPU work (an NCCL collective would queue behind the running search)."""
        if self._cpu_group is None:
            ranks = dist.get_process_group_ranks(self.group) if self.group is not None else None
            self._cpu_group = dist.new_group(ranks=ranks, backend="gloo")
        return self._cpu_group

    def rebalance(self, my_ms: float, *, damping: float = 0.6) -> List[Tuple[int, int]]:
        """One control step: exchange what every rank takes for its CURRENT rows (host-side
        all-gather), move the boundaries towards equal times, apply them.  Collective."""
        times = [None] * self.world
        dist.all_gather_object(times, float(my_ms), group=self.cpu_group())
        small = self.n_rows < 64 * SHARD_ALIGN * self.world           # small banks: exact cuts
        new = rebalanced_bounds(self.bounds, times, self.stores, damping=damping,
                                align=1 if small else SHARD_ALIGN,
                                min_rows=max(1, min(SHARD_ALIGN, min(hi - lo for lo, hi in self.base_bounds) // 4)))
        self.set_bounds(new)
        return new

    # ---------------------------------------------------------------- bank
    def upload_local(self, rows: torch.Tensor, dst_row: int = 0, *, normalize: bool = True) -> None:
        """Store `rows` at LOCAL rows [dst_row, dst_row + n) of this rank's STORED rows
        [store_lo, store_hi) (= its shard [lo, hi) unless the bank was built with overlap)."""
        self.local.upload(rows, dst_row, normalize=normalize)

    def upload_global(self, bank: torch.Tensor, *, normalize: bool = True) -> None:
        """Every rank passes the same full [N, d] bank; each keeps its own row range."""
        if bank.shape[0] != self.n_rows:
            raise ValueError(f"bank has {bank.shape[0]} rows, expected {self.n_rows}")
        self.local.upload(bank[self.store_lo:self.store_hi], 0, normalize=normalize)

    # ---------------------------------------------------------------- queries
    def replicate_from_host(self, host_queries: torch.Tensor) -> torch.Tensor:
        """Full [Q, d] query batch on this rank's device from a host batch every rank holds.

        Every rank needs all queries (the bank, not the batch, is sharded).  Copying the whole
        batch over each GPU's PCIe link costs G x the bytes; here every rank uploads only rows
        [r*ceil(Q/G), (r+1)*ceil(Q/G)) from (pinned) host memory and the slices are all-gathered
        over NVLink — at 8 GPUs and 65,536 fp32 queries 33 MB of PCIe + < 1 ms of all-gather
        instead of 268 MB of PCIe per rank.
        """
        if host_queries.dim() != 2:
            raise ValueError(f"queries must be [Q, d], got {tuple(host_queries.shape)}")
        if self.world == 1:
            return host_queries.to(self.device, non_blocking=True)
        q, d = host_queries.shape
        per = -(-q // self.world)
        full = torch.empty((self.world * per, d), dtype=host_queries.dtype, device=self.device)
        lo = min(self.rank * per, q)
        hi = min(lo + per, q)
        mine = full[self.rank * per:(self.rank + 1) * per]
        if hi > lo:
            mine[:hi - lo].copy_(host_queries[lo:hi], non_blocking=True)
        dist.all_gather_into_tensor(full.view(-1), mine.reshape(-1), group=self.group)
        return full[:q]

    # ---------------------------------------------------------------- search
    def search(self, queries: torch.Tensor, k: int, *, normalize_queries: bool = True,
               self_index: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Global top-k on every rank: (scores [Q, k] float32, indices [Q, k] int64)."""
        k = int(k)
        q = queries.shape[0]
        min_rows = min(hi - lo for lo, hi in self.bounds) - (1 if self_index is not None else 0)
        if k > min_rows:
            raise RuntimeError(f"selected index k out of range: k={k} exceeds the smallest shard "
                               f"({min_rows} usable rows over {self.world} ranks)")
        # packed per-rank payload: [ scores fp32 (padded to 8 B) | indices int64 ]
        score_bytes = (q * k * 4 + 7) // 8 * 8
        total_bytes = score_bytes + q * k * 8
        if self.world == 1:
            return self.local.search(queries, k, normalize_queries=normalize_queries,
                                     self_index=self_index)
        gathered = torch.empty((self.world, total_bytes), dtype=torch.uint8, device=self.device)
        mine = gathered[self.rank]
        my_scores = mine[:q * k * 4].view(torch.float32).view(q, k)
        my_index = mine[score_bytes:].view(torch.int64).view(q, k)
        self.local.search(queries, k, normalize_queries=normalize_queries, self_index=self_index,
                          out=(my_scores, my_index))
        # in-place all-gather: rank r's slot of `gathered` is the send buffer
        dist.all_gather_into_tensor(gathered.view(-1), mine, group=self.group)
        all_scores = gathered[:, :q * k * 4].view(torch.float32).view(self.world, q, k)
        all_index = gathered[:, score_bytes:].view(torch.int64).view(self.world, q, k)
        return self.local.merge(all_scores, all_index)


class SearchPipeline:
    """Software pipeline of searches against one (sharded) bank: three CUDA streams per rank.

        input   : the host query batch -> device, pinned H2D on the copy engine      [batch i+1]
        search  : normalise + cast -> fused similarity / top-k over the shard ->
                  exchange of the shard-local lists (NCCL) -> k-way merge            [batch i]
        output  : merged rows -> pinned host memory, D2H on the copy engine          [batch i-1]

    What overlaps is what the copy engines do.  The fused kernel is persistent — one CTA per SM,
    all of the SM's registers — so nothing that needs SMs can run beside it: round 2 measured an
    exchange + merge issued on a side stream under the next batch's kernel at 2 GPUs, and the NCCL
    kernels (spinning on SMs until the slower rank arrives) kept part of the next grid from
    starting, its lock-step windows timed out, and the step went from 446 to 602 ms
    (profiles/r02/SUMMARY.md).  Collectives therefore stay on the search stream, between fused
    kernels, where they cost their own ~1 ms; uploads and read-backs cost nothing.

    Queries: input="replicate" uploads only this rank's 1/G slice and all-gathers the batch over
    NVLink on the search stream (G x less PCIe traffic, ~0.5 ms of NVLink time per 268 MB);
    input="full" (default when from_host) has every rank upload the whole batch over its own PCIe
    link, which the copy engine hides entirely under the previous batch's kernel.

    result="replicated": all-gather + merge of all Q rows on every rank (what
    ShardedRelatedBank.search returns).  result="row_slice": rank r ends up with the global top-k
    of query rows [r*ceil(Q/G), (r+1)*ceil(Q/G)) only — an all-to-all moves 1/G of the lists to
    each rank, which merges (and copies out) 1/G of the rows: the natural form when each process
    hands its share of the result to the host.

    rescore_from: the fp32 rows of THIS rank's shard ([local rows, d], unit rows or raw — cosine
    is recomputed).  The search stage then fetches k + rescore_margin bf16 candidates from the
    shard, re-scores them in fp32 (zs_rescore_f32) and passes its k best ON: shard lists carry
    fp32 scores, so the merged result is the fp32 ranking (reference
    embeddings_related_generator.py:22) wherever the true top-k lies inside the candidate sets.

    balance_every = n > 0 (sharded bank built with overlap > 0): before every n-th batch the ranks
    exchange (on the host, over gloo, while the previous batch is still running on the GPUs) their
    running mean time per bank row, and the shard boundaries move to where those predict equal
    times (ShardedRelatedBank.rebalance).  The result does not depend on the
    boundaries; the slowest GPU of the box stops setting the pace of every batch.
    """

    def __init__(self, bank, n_queries: int, k: int, *, depth: int = 2, from_host: bool = True,
                 to_host: bool = True, result: str = "replicated",
                 self_index: Optional[torch.Tensor] = None, normalize_queries: bool = True,
                 query_dtype: torch.dtype = torch.float32,
                 rescore_from: Optional[torch.Tensor] = None, rescore_margin: int = 8,
                 excludes_self: bool = False, input: str = "full", balance_every: int = 0,
                 balance_damping: float = 1.0):
        if result not in ("replicated", "row_slice"):
            raise ValueError(f"result must be 'replicated' or 'row_slice', got {result!r}")
        if input not in ("full", "replicate", "slice"):
            raise ValueError(f"input must be 'full', 'replicate' or 'slice', got {input!r}")
        # "slice": like "replicate", and only rows [lo, hi) of the host batch are meaningful
        self.input = input
        self.sharded = isinstance(bank, ShardedRelatedBank)
        self.bank = bank
        self.local = bank.local if self.sharded else bank
        self.world = bank.world if self.sharded else 1
        self.rank = bank.rank if self.sharded else 0
        self.group = bank.group if self.sharded else None
        self.device = self.local.device
        self.q, self.k, self.dim = int(n_queries), int(k), self.local.dim
        self.from_host, self.to_host, self.result = from_host, to_host, result
        self.self_index = self_index
        self.normalize_queries = normalize_queries
        self.rescore_from = rescore_from
        self.kc = self.k
        if rescore_from is not None:
            if rescore_from.dtype != torch.float32 or rescore_from.shape[0] != self.local.rows:
                raise ValueError("rescore_from must be the float32 rows of this rank's shard")
            if not from_host and query_dtype != torch.float32:
                raise TypeError("fp32 re-scoring needs float32 queries")
            avail = self.local.rows - (1 if (self_index is not None or excludes_self) else 0)
            from . import _abi
            self.kc = max(self.k, min(self.k + int(rescore_margin), _abi.ZS_MAX_K, avail))
        G, q, dev = self.world, self.q, self.device
        self.per = -(-q // G)                          # query rows per rank (input slices, row_slice results)
        self.lo = min(self.rank * self.per, q)
        self.hi = min(self.lo + self.per, q)
        self.s_in, self.s_main, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
        rows_out = self.per if (result == "row_slice" and G > 1) else q
        self.rows_out = rows_out
        self.slots = []
        for _ in range(depth):
            slot = {
                "q_full": torch.empty((G * self.per, self.dim), dtype=query_dtype, device=dev) if from_host else None,
                "local_s": torch.empty((G * self.per, self.k), dtype=torch.float32, device=dev),
                "local_i": torch.empty((G * self.per, self.k), dtype=torch.int64, device=dev),
                "out_s": torch.empty((rows_out, self.k), dtype=torch.float32, device=dev),
                "out_i": torch.empty((rows_out, self.k), dtype=torch.int64, device=dev),
                "in_done": torch.cuda.Event(), "search_done": torch.cuda.Event(),
                "out_done": torch.cuda.Event(), "used": False,
            }
            if rescore_from is not None:
                slot["cand_s"] = torch.empty((self.q, self.kc), dtype=torch.float32, device=dev)
                slot["cand_i"] = torch.empty((self.q, self.kc), dtype=torch.int64, device=dev)
            if G > 1:
                slot["recv_s"] = torch.empty((G, rows_out, self.k), dtype=torch.float32, device=dev)
                slot["recv_i"] = torch.empty((G, rows_out, self.k), dtype=torch.int64, device=dev)
            if to_host:
                slot["host_s"] = torch.empty((rows_out, self.k), dtype=torch.float32).pin_memory()
                slot["host_i"] = torch.empty((rows_out, self.k), dtype=torch.int64).pin_memory()
            self.slots.append(slot)
        self.local.reserve(self.q, self.kc)
        self._next = 0
        self.balance_every = int(balance_every) if (self.sharded and G > 1 and getattr(bank, "adaptive", False)) else 0
        self.balance_damping = float(balance_damping)
        if rescore_from is not None and self.balance_every:
            raise ValueError("balance_every and rescore_from cannot be combined (rescore_from is sized to the shard)")
        self._step = 0
        self._timed = []          # (step, start event, end event, rows) of the shard-local searches
        self.rebalances = 0
        self.rows_log = []        # bank rows this rank searched, per submitted batch
        up_rows = q if input == "full" else (self.hi - self.lo)
        self.h2d_bytes = up_rows * self.dim * torch.empty((), dtype=query_dtype).element_size() if from_host else 0
        self.d2h_bytes = rows_out * self.k * 12 if to_host else 0

    @property
    def out_rows(self) -> Tuple[int, int]:
        """[lo, hi) query rows whose global top-k this rank's outputs hold."""
        if self.result == "row_slice" and self.world > 1:
            return self.lo, self.hi
        return 0, self.q

    def submit(self, queries: torch.Tensor, *, self_index: Optional[torch.Tensor] = None) -> int:
        """Enqueue one batch ([Q, d]: pinned host tensor if from_host — only rows [lo, hi) of this
        rank are read — else a device tensor that must stay untouched until the slot is waited
        for).  self_index overrides the pipeline's for this batch.  Returns the slot holding the
        result."""
        if self_index is None:
            self_index = self.self_index
        if self_index is not None:
            self_index = self_index.detach().to(device=self.device, dtype=torch.int64).contiguous()
        if self.balance_every and self._step >= self.balance_every + 1 and self._step % self.balance_every == 0:
            self._rebalance()
        idx = self._next
        self._next = (self._next + 1) % len(self.slots)
        slot = self.slots[idx]
        cur = torch.cuda.current_stream(self.device)
        G = self.world
        replicate = self.from_host and G > 1 and self.input != "full"
        if self.from_host:
            with torch.cuda.stream(self.s_in):
                if slot["used"]:
                    self.s_in.wait_event(slot["search_done"])      # the previous search read q_full
                else:
                    self.s_in.wait_stream(cur)
                q_full = slot["q_full"]
                if replicate:
                    mine = q_full[self.rank * self.per:(self.rank + 1) * self.per]
                    if self.hi > self.lo:
                        mine[:self.hi - self.lo].copy_(queries[self.lo:self.hi], non_blocking=True)
                else:
                    q_full[:self.q].copy_(queries[:self.q], non_blocking=True)
                slot["in_done"].record(self.s_in)
            q_dev = q_full[:self.q]
        else:
            q_dev = queries
        with torch.cuda.stream(self.s_main):
            if self.from_host:
                self.s_main.wait_event(slot["in_done"])
            elif not slot["used"]:
                self.s_main.wait_stream(cur)
            if slot["used"]:
                self.s_main.wait_event(slot["out_done"])          # result buffers are free again
            if replicate:
                mine = q_full[self.rank * self.per:(self.rank + 1) * self.per]
                dist.all_gather_into_tensor(q_full.view(-1), mine.reshape(-1), group=self.group)
            if self.rescore_from is None:
                if self.balance_every:
                    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    ev0.record(self.s_main)
                self.local.search(q_dev, self.k, normalize_queries=self.normalize_queries,
                                  self_index=self_index,
                                  out=(slot["local_s"][:self.q], slot["local_i"][:self.q]))
                if self.balance_every:
                    ev1.record(self.s_main)
                    self._timed.append((self._step, ev0, ev1, self.bank.hi - self.bank.lo))
            else:
                self.local.search(q_dev, self.kc, normalize_queries=self.normalize_queries,
                                  self_index=self_index, out=(slot["cand_s"], slot["cand_i"]))
                self.local.rescore(q_dev, self.rescore_from, slot["cand_i"], self.k,
                                   normalize=self.normalize_queries, index_offset=self.local.index_offset,
                                   out=(slot["local_s"][:self.q], slot["local_i"][:self.q]))
            if G == 1:
                res_s, res_i = slot["local_s"][:self.q], slot["local_i"][:self.q]
            else:
                if self.result == "row_slice":
                    dist.all_to_all_single(slot["recv_s"].view(-1), slot["local_s"].view(-1), group=self.group)
                    dist.all_to_all_single(slot["recv_i"].view(-1), slot["local_i"].view(-1), group=self.group)
                else:
                    dist.all_gather_into_tensor(slot["recv_s"].view(-1), slot["local_s"][:self.q].reshape(-1), group=self.group)
                    dist.all_gather_into_tensor(slot["recv_i"].view(-1), slot["local_i"][:self.q].reshape(-1), group=self.group)
                self.local.merge(slot["recv_s"], slot["recv_i"], out=(slot["out_s"], slot["out_i"]))
                res_s, res_i = slot["out_s"], slot["out_i"]
            slot["res"] = (res_s, res_i)
            slot["search_done"].record(self.s_main)
        if self.to_host:
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(slot["search_done"])
                slot["host_s"].copy_(res_s, non_blocking=True)
                slot["host_i"].copy_(res_i, non_blocking=True)
                slot["out_done"].record(self.s_out)
        else:
            slot["out_done"].record(self.s_main)
        slot["keepalive"] = (queries, self_index)     # inputs of kernels still queued on the side streams
        slot["used"] = True
        self._step += 1
        self.rows_log.append((self.bank.hi - self.bank.lo) if self.sharded else self.local.rows)
        return idx

    def _rebalance(self) -> None:
        """Called before batch n is enqueued (n a multiple of balance_every): batches up to n - 2
        have finished or are about to — waiting for the end of batch n - 2 keeps the host at most
        two batches ahead of the GPU and costs the GPU nothing, batch n - 1 is queued behind it.

        What steers the boundaries is each rank's time PER ROW averaged over every batch finished
        so far (the first one excepted): under the power cap the step-to-step noise of one GPU
        (sigma ~ 2 % of a 110 ms kernel) is as large as the differences between the GPUs of a box,
        so a controller that follows the last few batches only adds its own jitter; the running
        mean gets better by 1 / sqrt(n) and the boundaries settle."""
        done = [t for t in self._timed if 1 <= t[0] <= self._step - 2]
        if not done:
            return
        done[-1][2].synchronize()
        for _, a, b, rows in done:           # (kept on the bank: every pipeline over it adds to the same mean)
            self.bank.mpr_sum += a.elapsed_time(b) / rows
            self.bank.mpr_n += 1
        self._timed = [t for t in self._timed if t[0] > done[-1][0]]
        mean_ms = self.bank.mpr_sum / self.bank.mpr_n * (self.bank.hi - self.bank.lo)
        self.bank.rebalance(mean_ms, damping=self.balance_damping)
        self.rebalances += 1

    def wait_stream(self, idx: Optional[int] = None) -> None:
        """Make the current stream wait for slot `idx` (all slots if None) — no host sync."""
        cur = torch.cuda.current_stream(self.device)
        for j, slot in enumerate(self.slots):
            if slot["used"] and (idx is None or idx == j):
                cur.wait_event(slot["out_done"])

    def result_of(self, idx: int, *, host: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
        """Result tensors of slot `idx` (device, or the pinned host copies); the caller must have
        waited (wait_stream, or a host synchronise for the host copies).  Rows out_rows."""
        slot = self.slots[idx]
        if host:
            n = self.out_rows[1] - self.out_rows[0]
            return slot["host_s"][:n], slot["host_i"][:n]
        s, i = slot["res"]
        n = self.out_rows[1] - self.out_rows[0]
        return s[:n], i[:n]
