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
                 shard_weights: Optional[Sequence[float]] = None):
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
        self.bounds = shard_bounds(self.n_rows, self.world, shard_weights)
        self.lo, self.hi = self.bounds[self.rank]
        if min(hi - lo for lo, hi in self.bounds) < 1:
            raise ValueError(f"bank of {n_rows} rows cannot be sharded over {self.world} ranks")
        factory = local_bank_factory or _default_local_bank
        self.local = factory(self.hi - self.lo, self.dim, device, self.lo)
        self.device = self.local.device

    # ---------------------------------------------------------------- bank
    def upload_local(self, rows: torch.Tensor, dst_row: int = 0, *, normalize: bool = True) -> None:
        """Store `rows` at LOCAL rows [dst_row, dst_row + n) of this rank's shard."""
        self.local.upload(rows, dst_row, normalize=normalize)

    def upload_global(self, bank: torch.Tensor, *, normalize: bool = True) -> None:
        """Every rank passes the same full [N, d] bank; each keeps its own row range."""
        if bank.shape[0] != self.n_rows:
            raise ValueError(f"bank has {bank.shape[0]} rows, expected {self.n_rows}")
        self.local.upload(bank[self.lo:self.hi], 0, normalize=normalize)

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
