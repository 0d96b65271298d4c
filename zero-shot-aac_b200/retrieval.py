"""Related-caption retrieval on one B200: cosine similarity against a caption-embedding bank and
per-query top-k, through the C ABI of libzsaac_b200.so.

`related_topk` is the batched form of what the reference computes one item at a time in
process_data (data_handing/embeddings_related_generator.py:21-22):

    text_embs = F.normalize(item['text_embedding'], dim=-1).to('cuda')
    ids = torch.cosine_similarity(text_embs, valid_text_embs).topk(topnumber)[1]

and of utils.sound_effect_choice (utils.py:133-135, no normalisation there).  torch is used for
device memory and streams only; every kernel on the path lives in the native library.
"""
from __future__ import annotations

import ctypes
import weakref
from typing import Optional, Tuple

import torch

from . import _abi


def _require_cuda() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "zsaac_b200 needs a CUDA device (B200, sm_100a): there is no CPU fallback for the "
            "related-caption retrieval path")


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return _abi.ZS_F32
    if t.dtype == torch.bfloat16:
        return _abi.ZS_BF16
    raise TypeError(f"embeddings must be float32 or bfloat16, got {t.dtype}")


class RelatedBank:
    """A caption-embedding memory bank resident on one GPU as bf16 rows (library-owned copy).

    rows         number of bank rows held by this object (a shard when index_offset > 0)
    dim          embedding dimension (multiple of 64; 1024 in the reference)
    index_offset global index of local row 0; added to every index `search` returns
    """

    def __init__(self, rows: int, dim: int, *, device=None, index_offset: int = 0):
        _require_cuda()
        self._lib = _abi.load_library()
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.type != "cuda":
            raise ValueError(f"RelatedBank lives on a CUDA device, got {dev}")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self.rows = int(rows)
        self.dim = int(dim)
        self.index_offset = int(index_offset)
        self.window_rows = None             # (row_lo, n_rows) while a search window is set
        handle = ctypes.c_void_p()
        _abi.check(self._lib.zs_create(ctypes.byref(handle), dev.index))
        self._handle = handle
        self._finalizer = weakref.finalize(self, self._lib.zs_destroy, handle)
        _abi.check(self._lib.zs_bank_alloc(self._ctx, self.rows, self.dim))

    @property
    def _ctx(self) -> ctypes.c_void_p:
        """The native context; raises once the bank has been closed (never a dangling pointer)."""
        if self._handle is None:
            raise RuntimeError("this RelatedBank has been closed")
        return self._handle

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_tensor(cls, bank: torch.Tensor, *, normalize: bool = True, index_offset: int = 0,
                    device=None) -> "RelatedBank":
        """Build from a [N, d] float32 / bfloat16 tensor (moved to the GPU if it is not there)."""
        if bank.dim() != 2:
            raise ValueError(f"bank must be [N, d], got {tuple(bank.shape)}")
        _require_cuda()
        if not bank.is_cuda:
            bank = bank.to(device if device is not None else "cuda", non_blocking=False)
        obj = cls(bank.shape[0], bank.shape[1], device=bank.device, index_offset=index_offset)
        obj.upload(bank, 0, normalize=normalize)
        return obj

    def upload(self, rows: torch.Tensor, dst_row: int = 0, *, normalize: bool = True) -> None:
        """Cast (and L2-normalise, F.normalize eps=1e-12) `rows` into bank rows [dst_row, ...)."""
        if rows.dim() != 2 or rows.shape[1] != self.dim:
            raise ValueError(f"rows must be [n, {self.dim}], got {tuple(rows.shape)}")
        rows = rows.detach()
        if rows.device != self.device:
            rows = rows.to(self.device)
        rows = rows.contiguous()
        with torch.cuda.device(self.device):
            _abi.check(self._lib.zs_bank_upload(
                self._ctx, rows.data_ptr(), rows.shape[0], int(dst_row), _dtype_code(rows),
                1 if normalize else 0, _stream_ptr(self.device)))
        # `rows` must stay alive until the enqueued kernel has read it
        rows.record_stream(torch.cuda.current_stream(self.device))

    # ------------------------------------------------------------------ search
    def window(self, row_lo: int = 0, n_rows: int = 0) -> None:
        """Restrict search() to stored rows [row_lo, row_lo + n_rows) (indices stay global);
        window() without arguments lifts it.  Host-side only, no synchronisation."""
        _abi.check(self._lib.zs_bank_window(self._ctx, int(row_lo), int(n_rows)))
        self.window_rows = (int(row_lo), int(n_rows)) if n_rows else None

    def reserve(self, n_queries: int, k: int) -> None:
        _abi.check(self._lib.zs_reserve(self._ctx, int(n_queries), int(k)))

    def search(self, queries: torch.Tensor, k: int, *, normalize_queries: bool = True,
               self_index: Optional[torch.Tensor] = None,
               out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
               ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Top-k of every query against this bank.

        queries [Q, d] float32 / bfloat16 on this device.  Returns (scores [Q, k] float32
        descending, indices [Q, k] int64 global), ties ordered by ascending index.
        self_index: optional [Q] int64 global bank index each query must not return.
        """
        if queries.dim() != 2 or queries.shape[1] != self.dim:
            raise ValueError(f"queries must be [Q, {self.dim}], got {tuple(queries.shape)}")
        if queries.device != self.device:
            raise ValueError(f"queries are on {queries.device}, the bank is on {self.device}")
        k = int(k)
        queries = queries.detach().contiguous()
        q = queries.shape[0]
        if out is None:
            scores = torch.empty((q, k), dtype=torch.float32, device=self.device)
            indices = torch.empty((q, k), dtype=torch.int64, device=self.device)
        else:
            scores, indices = out
            if (scores.shape != (q, k) or indices.shape != (q, k) or scores.dtype != torch.float32
                    or indices.dtype != torch.int64 or not scores.is_contiguous()
                    or not indices.is_contiguous()):
                raise ValueError("out must be contiguous (float32 [Q,k], int64 [Q,k])")
        self_ptr = None
        if self_index is not None:
            self_index = self_index.detach().to(device=self.device, dtype=torch.int64).contiguous()
            if self_index.shape != (q,):
                raise ValueError(f"self_index must be [{q}], got {tuple(self_index.shape)}")
            self_ptr = self_index.data_ptr()
        with torch.cuda.device(self.device):
            _abi.check(self._lib.zs_search(
                self._ctx, queries.data_ptr(), q, _dtype_code(queries), k,
                1 if normalize_queries else 0, self_ptr, self.index_offset,
                scores.data_ptr(), indices.data_ptr(), _stream_ptr(self.device)))
        return scores, indices

    def rank_of(self, queries: torch.Tensor, targets: torch.Tensor, *, normalize_queries: bool = True
                ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Rank of given bank rows in every query's similarity ordering, without sorting.

        targets [Q, T] (T <= 8) int64 global bank indices (< 0 = unused slot).  Returns
        (ranks [Q, T] int64, target_scores [Q, T] float32): ranks[q, t] = number of other bank
        rows scoring strictly higher than targets[q, t] for query q (0 = retrieved first; -1 for
        unused slots).  Same fused GEMM as search() with a counting epilogue.
        """
        if queries.dim() != 2 or queries.shape[1] != self.dim:
            raise ValueError(f"queries must be [Q, {self.dim}], got {tuple(queries.shape)}")
        if queries.device != self.device:
            raise ValueError(f"queries are on {queries.device}, the bank is on {self.device}")
        queries = queries.detach().contiguous()
        q = queries.shape[0]
        targets = targets.detach().to(device=self.device, dtype=torch.int64)
        if targets.dim() == 1:
            targets = targets.unsqueeze(1)
        if targets.dim() != 2 or targets.shape[0] != q:
            raise ValueError(f"targets must be [{q}, T], got {tuple(targets.shape)}")
        targets = targets.contiguous()
        t = targets.shape[1]
        ranks = torch.empty((q, t), dtype=torch.int64, device=self.device)
        scores = torch.empty((q, t), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _abi.check(self._lib.zs_rank_count(
                self._ctx, queries.data_ptr(), q, _dtype_code(queries), 1 if normalize_queries else 0,
                targets.data_ptr(), t, self.index_offset, scores.data_ptr(), ranks.data_ptr(),
                _stream_ptr(self.device)))
        return ranks, scores

    def rescore(self, queries: torch.Tensor, bank_f32: torch.Tensor, candidates: torch.Tensor, k: int,
                *, normalize: bool = True, index_offset: int = 0,
                out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
                ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Re-score search candidates in fp32 and keep the k best (score desc, index asc).

        queries [Q, d] float32 (raw), bank_f32 [N, d] float32 on this device, candidates [Q, kc]
        int64 global indices (kc <= 2048; bank_f32 row 0 has global index `index_offset`).  With
        normalize=True the score is the cosine similarity computed entirely in fp32 — the
        arithmetic of the reference's torch.cosine_similarity (embeddings_related_generator.py:22)
        — so searching for k + margin candidates and re-scoring them removes the bf16 near-tie
        swaps of the fused kernel.
        """
        if queries.dtype != torch.float32 or bank_f32.dtype != torch.float32:
            raise TypeError("rescore expects float32 queries and a float32 bank")
        if queries.dim() != 2 or bank_f32.dim() != 2 or queries.shape[1] != bank_f32.shape[1]:
            raise ValueError("rescore expects queries [Q, d] and bank [N, d]")
        if queries.device != self.device or bank_f32.device != self.device:
            raise ValueError(f"rescore expects tensors on {self.device}")
        queries = queries.detach().contiguous()
        bank_f32 = bank_f32.detach().contiguous()
        cand = candidates.detach().to(device=self.device, dtype=torch.int64).contiguous()
        q = queries.shape[0]
        if cand.dim() != 2 or cand.shape[0] != q:
            raise ValueError(f"candidates must be [{q}, kc], got {tuple(cand.shape)}")
        kc, k = cand.shape[1], int(k)
        if out is None:
            out_s = torch.empty((q, k), dtype=torch.float32, device=self.device)
            out_i = torch.empty((q, k), dtype=torch.int64, device=self.device)
        else:
            out_s, out_i = out
            if (tuple(out_s.shape) != (q, k) or tuple(out_i.shape) != (q, k)
                    or out_s.dtype != torch.float32 or out_i.dtype != torch.int64
                    or not out_s.is_contiguous() or not out_i.is_contiguous()):
                raise ValueError("out must be contiguous (float32 [Q,k], int64 [Q,k])")
        with torch.cuda.device(self.device):
            _abi.check(self._lib.zs_rescore_f32(
                self._ctx, queries.data_ptr(), q, 1 if normalize else 0, bank_f32.data_ptr(),
                bank_f32.shape[0], bank_f32.shape[1], int(index_offset), cand.data_ptr(), kc, k,
                out_s.data_ptr(), out_i.data_ptr(), _stream_ptr(self.device)))
        return out_s, out_i

    def debug_scores(self, queries: torch.Tensor, *, normalize_queries: bool = True) -> torch.Tensor:
        """Full [Q, rows] similarity matrix out of the same tcgen05 pipeline (tests only)."""
        queries = queries.detach().contiguous()
        out = torch.empty((queries.shape[0], self.rows), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _abi.check(self._lib.zs_debug_scores(
                self._ctx, queries.data_ptr(), queries.shape[0], _dtype_code(queries),
                1 if normalize_queries else 0, out.data_ptr(), _stream_ptr(self.device)))
        return out

    # ------------------------------------------------------------------ helpers on the same ctx
    def merge(self, scores: torch.Tensor, indices: torch.Tensor,
              out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
              ) -> Tuple[torch.Tensor, torch.Tensor]:
        """k-way merge of [S, Q, k] sorted lists -> [Q, k] under (score desc, index asc).

        The S lists may be strided views (dim 0 stride arbitrary, [Q, k] blocks contiguous), e.g.
        slices of one all-gathered byte buffer.  `out`: optional preallocated contiguous
        (float32 [Q, k], int64 [Q, k]) result tensors."""
        if scores.dim() != 3 or scores.shape != indices.shape:
            raise ValueError("merge expects scores and indices of identical shape [S, Q, k]")
        s, q, k = scores.shape
        if scores.dtype != torch.float32 or indices.dtype != torch.int64:
            raise TypeError("merge expects float32 scores and int64 indices")

        def blocks_contiguous(t):
            return s == 0 or q == 0 or (t.stride(2) == 1 and t.stride(1) == k)

        if not blocks_contiguous(scores):
            scores = scores.contiguous()
        if not blocks_contiguous(indices):
            indices = indices.contiguous()
        if out is None:
            out_s = torch.empty((q, k), dtype=torch.float32, device=self.device)
            out_i = torch.empty((q, k), dtype=torch.int64, device=self.device)
        else:
            out_s, out_i = out
            if (tuple(out_s.shape) != (q, k) or tuple(out_i.shape) != (q, k)
                    or out_s.dtype != torch.float32 or out_i.dtype != torch.int64
                    or not out_s.is_contiguous() or not out_i.is_contiguous()):
                raise ValueError("out must be contiguous (float32 [Q,k], int64 [Q,k])")
        s_stride = scores.stride(0) if s > 1 else q * k
        i_stride = indices.stride(0) if s > 1 else q * k
        with torch.cuda.device(self.device):
            _abi.check(self._lib.zs_merge(
                self._ctx, scores.data_ptr(), indices.data_ptr(), s, s_stride, i_stride, q, k,
                out_s.data_ptr(), out_i.data_ptr(), _stream_ptr(self.device)))
        return out_s, out_i

    def gather_rows(self, src: torch.Tensor, indices: torch.Tensor) -> torch.Tensor:
        """src[indices] for a float32 [N, d] tensor on this device -> [*indices.shape, d]."""
        if src.dtype != torch.float32 or src.dim() != 2 or src.device != self.device:
            raise ValueError("gather_rows expects a float32 [N, d] tensor on the bank's device")
        src = src.contiguous()
        flat = indices.to(device=self.device, dtype=torch.int64).contiguous().view(-1)
        out = torch.empty((flat.numel(), src.shape[1]), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _abi.check(self._lib.zs_gather_rows_f32(
                self._ctx, src.data_ptr(), src.shape[0], src.shape[1], flat.data_ptr(),
                flat.numel(), out.data_ptr(), _stream_ptr(self.device)))
        return out.view(*indices.shape, src.shape[1])

    def normalize_rows(self, x: torch.Tensor) -> torch.Tensor:
        """F.normalize(x, dim=-1) for a float32 [N, d] tensor on this device (new tensor)."""
        if x.dtype != torch.float32 or x.dim() != 2 or x.device != self.device:
            raise ValueError("normalize_rows expects a float32 [N, d] tensor on the bank's device")
        x = x.contiguous()
        out = torch.empty_like(x)
        with torch.cuda.device(self.device):
            _abi.check(self._lib.zs_normalize_rows_f32(
                self._ctx, x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1],
                _stream_ptr(self.device)))
        return out

    def plan(self, n_queries: int, k: int) -> Tuple[int, int, int]:
        """(bank chunks, 256-row tiles per chunk, CTAs) the library would launch."""
        a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _abi.check(self._lib.zs_plan(self._ctx, int(n_queries), int(k), ctypes.byref(a),
                                     ctypes.byref(b), ctypes.byref(c)))
        return a.value, b.value, c.value

    def trace(self, stamps: Optional[torch.Tensor]) -> None:
        """Per-CTA %globaltimer stamps of the fused kernel into a uint64-sized [ctas, 16] int64
        tensor on this device (None switches tracing off).  Tuning hook."""
        self._trace_keepalive = stamps
        _abi.check(self._lib.zs_debug_trace(self._ctx, None if stamps is None else stamps.data_ptr()))

    def profile(self, enable: bool) -> None:
        """Bracket every fused-kernel launch with CUDA events (ring of 256 launches)."""
        _abi.check(self._lib.zs_profile_enable(self._ctx, 1 if enable else 0))

    def kernel_times_ms(self) -> list:
        """Device durations of the fused kernel launches recorded since profile(True)."""
        buf = (ctypes.c_float * 256)()
        n = ctypes.c_int()
        _abi.check(self._lib.zs_profile_read(self._ctx, buf, 256, ctypes.byref(n)))
        return [float(buf[j]) for j in range(n.value)]

    @property
    def kernel_error(self) -> int:
        """Role code of a timed-out pipeline wait (0 = none); readable even after a kernel trap."""
        return int(self._lib.zs_kernel_error(self._ctx))

    @property
    def launch_count(self) -> int:
        return int(self._lib.zs_launch_count(self._ctx))

    def close(self) -> None:
        """Destroy the native context now.  Every later call on this object raises."""
        self._handle = None
        self._finalizer()

    @property
    def closed(self) -> bool:
        return self._handle is None


# ------------------------------------------------------------------------------------------------
# Bank cache: process_data / sound_effect_choice receive the fp32 bank tensor on every call, as
# in the reference; the bf16 copy is rebuilt only when that tensor (identity, version) changes.
# The cache only ever DROPS its reference to an evicted RelatedBank: a caller (e.g. a suspended
# process_data generator) may still hold the object, and the native context is destroyed by the
# object's finalizer once the last reference is gone — never under a live user.
_BANK_CACHE: "dict[tuple, tuple]" = {}   # key -> (weakref to the source tensor, RelatedBank)
_BANK_CACHE_MAX = 4
_REBUILDS = {"count": 0, "warned": False}


def _cache_key(bank: torch.Tensor, normalize: bool) -> tuple:
    return (bank.data_ptr(), tuple(bank.shape), bank.dtype, str(bank.device), bank._version,
            bool(normalize))


def _cache_insert(key: tuple, bank: torch.Tensor, obj: RelatedBank) -> None:
    for stale in [k for k, (ref, _) in _BANK_CACHE.items() if ref() is None]:
        _BANK_CACHE.pop(stale)               # the source tensor is gone: nobody can hit this entry
    while len(_BANK_CACHE) >= _BANK_CACHE_MAX:
        _BANK_CACHE.pop(next(iter(_BANK_CACHE)))
    _BANK_CACHE[key] = (weakref.ref(bank), obj)


def bank_for(bank: torch.Tensor, *, normalize: bool) -> RelatedBank:
    key = _cache_key(bank, normalize)
    hit = _BANK_CACHE.get(key)
    if hit is not None and hit[0]() is bank and not hit[1].closed:
        return hit[1]
    if hit is not None:                      # same address, different tensor object: stale
        _BANK_CACHE.pop(key)
    obj = RelatedBank.from_tensor(bank, normalize=normalize)
    _REBUILDS["count"] += 1
    if _REBUILDS["count"] > 64 and not _REBUILDS["warned"]:
        import warnings
        _REBUILDS["warned"] = True
        warnings.warn("zsaac_b200: the bf16 search copy of the bank has been rebuilt more than 64 "
                      "times; pass the SAME bank tensor object on every call (a fresh temporary "
                      "such as `bank.cuda()` defeats the cache), or hold a RelatedBank yourself")
    _cache_insert(key, bank, obj)
    return obj


def clear_bank_cache() -> None:
    _BANK_CACHE.clear()


# ------------------------------------------------------------------------------------------------
# Exact fp32 path for small banks (label banks, zero-shot prompts, retrieval metrics): the
# reference computes these with a plain fp32 matmul, so the mirror does too — on the GPU, through
# zs_exact_topk_f32 / zs_exact_rank_f32 — instead of ranking bf16 products.
EXACT_MAX_BANK_ROWS = 65536          # beyond this the tensor-core search (+ fp32 re-scoring) takes over
EXACT_MAX_SCORES = 1 << 28           # Q * N fp32 scores of scratch (1 GiB)

_HELPERS: "dict[tuple, RelatedBank]" = {}


def helper_context(device: torch.device) -> RelatedBank:
    """A context for the bank-less entry points (normalise, exact fp32, memory projection)."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    h = _HELPERS.get(key)
    if h is None or h.closed:
        h = RelatedBank(1, 64, device=torch.device("cuda", key[1]))
        _HELPERS[key] = h
    return h


def exact_fits(n_queries: int, n_rows: int) -> bool:
    return n_rows <= EXACT_MAX_BANK_ROWS and n_queries * n_rows <= EXACT_MAX_SCORES


def _f32_on(x: torch.Tensor, device: torch.device) -> torch.Tensor:
    return x.detach().to(device=device, dtype=torch.float32).contiguous()


def exact_topk(queries: torch.Tensor, bank: torch.Tensor, k: int, *, normalize: bool = False,
               self_index: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """fp32 scores (raw dot product, or cosine with normalize=True) of queries [Q, d] against a
    small bank [N, d] and the k best per query under (score desc, index asc).  One launch for up
    to 64 queries.  Returns (float32 [Q, k], int64 [Q, k]) on the bank's CUDA device."""
    _require_cuda()
    dev = bank.device if bank.is_cuda else torch.device("cuda", torch.cuda.current_device())
    b = _f32_on(bank, dev)
    q = _f32_on(queries, dev)
    if q.dim() != 2 or b.dim() != 2 or q.shape[1] != b.shape[1]:
        raise ValueError(f"exact_topk expects queries [Q, d] and bank [N, d], got {tuple(q.shape)} "
                         f"and {tuple(b.shape)}")
    k = int(k)
    nq = q.shape[0]
    out_s = torch.empty((nq, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
    self_ptr = None
    if self_index is not None:
        self_index = self_index.detach().to(device=dev, dtype=torch.int64).contiguous()
        self_ptr = self_index.data_ptr()
    h = helper_context(dev)
    with torch.cuda.device(dev):
        _abi.check(h._lib.zs_exact_topk_f32(
            h._ctx, q.data_ptr(), nq, b.data_ptr(), b.shape[0], b.shape[1], 1 if normalize else 0, k,
            self_ptr, 0, out_s.data_ptr(), out_i.data_ptr(), _stream_ptr(dev)))
    for t in (q, b):
        t.record_stream(torch.cuda.current_stream(dev))
    return out_s, out_i


def exact_rank(queries: torch.Tensor, bank: torch.Tensor, targets: torch.Tensor, *,
               normalize: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """Position of bank rows `targets` [Q, T] (int64, < 0 = unused) in every query's fp32
    similarity ordering (score desc, index asc): (ranks int64 [Q, T], target scores fp32 [Q, T])."""
    _require_cuda()
    dev = bank.device if bank.is_cuda else torch.device("cuda", torch.cuda.current_device())
    b = _f32_on(bank, dev)
    q = _f32_on(queries, dev)
    t = targets.detach().to(device=dev, dtype=torch.int64)
    if t.dim() == 1:
        t = t.unsqueeze(1)
    t = t.contiguous()
    if q.dim() != 2 or b.dim() != 2 or q.shape[1] != b.shape[1] or t.shape[0] != q.shape[0]:
        raise ValueError("exact_rank expects queries [Q, d], bank [N, d], targets [Q, T]")
    ranks = torch.empty(t.shape, dtype=torch.int64, device=dev)
    scores = torch.empty(t.shape, dtype=torch.float32, device=dev)
    h = helper_context(dev)
    with torch.cuda.device(dev):
        _abi.check(h._lib.zs_exact_rank_f32(
            h._ctx, q.data_ptr(), q.shape[0], b.data_ptr(), b.shape[0], b.shape[1],
            1 if normalize else 0, t.data_ptr(), t.shape[1], 0, scores.data_ptr(), ranks.data_ptr(),
            _stream_ptr(dev)))
    for x in (q, b, t):
        x.record_stream(torch.cuda.current_stream(dev))
    return ranks, scores


RESCORE_MARGIN = 8      # extra candidates fetched for fp32 re-scoring


def related_topk(queries: torch.Tensor, bank: torch.Tensor, k: int, *, exclude_self: bool = False,
                 self_index: Optional[torch.Tensor] = None, normalize: bool = True,
                 rescore_fp32: bool = False, precision: str = "bf16"
                 ) -> Tuple[torch.Tensor, torch.Tensor]:
    """scores, indices = top-k of cosine_similarity(queries, bank).

    queries [Q, d], bank [N, d] (float32 or bfloat16).  With normalize=True both sides are
    L2-normalised first (cosine similarity, reference embeddings_related_generator.py:21-22);
    with normalize=False the raw dot product is ranked (reference utils.py:133).
    exclude_self=True drops bank row i from the result of query i (self_index defaults to
    arange(Q)); the reference itself never excludes (slot 0 of its output is the item).
    rescore_fp32=True re-scores k + 8 bf16 candidates in fp32 from `bank` (float32 inputs only).
    precision="fp32" ranks exact fp32 scores on CUDA cores instead (small banks only: N <= 65,536
    and Q * N <= 2^28); the default "bf16" is the fused tensor-core search.  k may exceed 32: the
    search then runs ceil(k / 32) passes over the bank (k <= 1024).
    Returns float32 [Q, k] scores (descending) and int64 [Q, k] bank indices on the GPU.
    """
    _require_cuda()
    if precision not in ("bf16", "fp32"):
        raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
    if precision == "fp32":
        q2 = queries.detach()
        if q2.dim() == 1:
            q2 = q2.unsqueeze(0)
        if not exact_fits(q2.shape[0], bank.shape[0]):
            raise ValueError("precision='fp32' is for small banks (N <= 65,536, Q * N <= 2^28); "
                             "use rescore_fp32=True on the tensor-core search instead")
        if exclude_self and self_index is None:
            self_index = torch.arange(q2.shape[0], dtype=torch.int64)
        return exact_topk(q2, bank, k, normalize=normalize, self_index=self_index)
    rb = bank_for(bank, normalize=normalize)   # a CPU bank is copied to the current GPU once
    q = queries.detach()
    if q.dim() == 1:
        q = q.unsqueeze(0)
    q = q.to(rb.device)
    if exclude_self and self_index is None:
        self_index = torch.arange(q.shape[0], dtype=torch.int64, device=rb.device)
    if not rescore_fp32:
        return rb.search(q, k, normalize_queries=normalize, self_index=self_index)
    return search_rescored(rb, q, bank, k, normalize=normalize, self_index=self_index)


def search_rescored(rb: RelatedBank, queries: torch.Tensor, bank_f32: torch.Tensor, k: int, *,
                    normalize: bool = True, self_index: Optional[torch.Tensor] = None
                    ) -> Tuple[torch.Tensor, torch.Tensor]:
    """bf16 search for k + RESCORE_MARGIN candidates, then fp32 re-scoring from `bank_f32` (the
    float32 tensor the bf16 bank `rb` was built from) down to k: scores and order are those of an
    fp32 cosine similarity wherever the true top-k lies inside the candidate set."""
    if bank_f32.dtype != torch.float32 or queries.dtype != torch.float32:
        raise TypeError("rescore_fp32 needs float32 queries and a float32 bank")
    bank_dev = bank_f32 if bank_f32.device == rb.device else bank_f32.to(rb.device)
    avail = rb.rows - (1 if self_index is not None else 0)
    # (the margin only shrinks where the bank, or ZS_MAX_K = 1024, leaves no room for it)
    kc = max(int(k), min(int(k) + RESCORE_MARGIN, _abi.ZS_MAX_K, avail))
    _, cand = rb.search(queries, kc, normalize_queries=normalize, self_index=self_index)
    return rb.rescore(queries, bank_dev, cand, k, normalize=normalize, index_offset=rb.index_offset)
