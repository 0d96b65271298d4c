"""Shared implementation of the related-caption generator scripts.

The reference has two near-identical scripts (data_handing/embeddings_related_generator.py and
embeddings_related_generator_wavcaps.py); they differ only in load_data taking one path or a
list.  Both module mirrors in data_handing/ delegate here.

What changes against the reference: the N-iteration Python loop of process_data
(embeddings_related_generator.py:19-28 — per item: CPU normalise, H2D, cosine_similarity over the
whole bank, topk, gather, two D2H syncs) becomes batches of queries through one fused
similarity+top-k launch.  What does not change: function names, arguments, lazy generator
semantics, the mutated item dicts and the append-mode stream of per-record pickles.
"""
from __future__ import annotations

import pickle
from typing import Iterable, Iterator, List, Sequence, Tuple, Union

import torch

from .retrieval import RelatedBank, _require_cuda, bank_for, search_rescored

try:  # the reference wraps the writer loop in tqdm (embeddings_related_generator.py:33)
    from tqdm import tqdm
except Exception:  # pragma: no cover - tqdm is present in the image
    def tqdm(it, total=None):
        return it

# queries per fused launch; 16384 x k=5 gathered rows = 335 MB of pinned staging at d=1024
QUERY_BATCH = 16384


# ------------------------------------------------------------------------------------------------
# Reading the reference's input files.  A plain-pickled torch tensor carries its storage as
# `torch.storage._load_from_bytes(<bytes written by torch.save(storage), legacy format>)`, and
# unpickling it runs a full torch.load per tensor (~100 us: it even probes for a tar archive).
# The unpickler below resolves that one global to a direct parser of the legacy stream — header
# pickles, the persistent id (storage type, key, location, element count), the key list, then
# `int64 count | raw little-endian data` — and copies the raw bytes into a fresh storage.  Anything
# unexpected (other magic / protocol, several storages, a non-CPU location, a length that does not
# add up) goes to torch's own loader, so the records are identical either way; only the time
# differs.  The tensor itself is then rebuilt as a strided view of that storage instead of through
# torch._utils._rebuild_tensor_v2 (whose fake-mode detection costs another ~20 us per tensor).  ZSAAC_FAST_UNPICKLE=0 switches it off.
_LEGACY_MAGIC = 0x1950a86a20f9469cfc6c
_LEGACY_PROTOCOL = 1001
_STORAGE_DTYPES = {
    "FloatStorage": torch.float32, "DoubleStorage": torch.float64, "HalfStorage": torch.float16,
    "BFloat16Storage": torch.bfloat16, "LongStorage": torch.int64, "IntStorage": torch.int32,
    "ShortStorage": torch.int16, "CharStorage": torch.int8, "ByteStorage": torch.uint8,
    "BoolStorage": torch.bool,
}


class _PersistentIdUnpickler(pickle.Unpickler):
    def persistent_load(self, pid):          # ('storage', storage class, key, location, count, view)
        return pid


def _storage_from_legacy_bytes(b: bytes):
    import io
    import struct
    slow = torch.storage._load_from_bytes
    try:
        f = io.BytesIO(b)
        if pickle.load(f) != _LEGACY_MAGIC or pickle.load(f) != _LEGACY_PROTOCOL:
            return slow(b)
        sys_info = pickle.load(f)
        if not (isinstance(sys_info, dict) and sys_info.get("little_endian", False)):
            return slow(b)
        pid = _PersistentIdUnpickler(f).load()
        keys = pickle.load(f)
        if (not isinstance(pid, tuple) or len(pid) < 5 or pid[0] != "storage" or pid[3] != "cpu"
                or (len(pid) > 5 and pid[5] is not None) or list(keys) != [pid[2]]):
            return slow(b)
        dtype = _STORAGE_DTYPES.get(getattr(pid[1], "__name__", ""))
        if dtype is None:
            return slow(b)
        count, pos = int(pid[4]), f.tell()
        size = torch._utils._element_size(dtype)
        if len(b) != pos + 8 + count * size or struct.unpack("<q", b[pos:pos + 8])[0] != count:
            return slow(b)
        if count == 0:
            flat = torch.empty(0, dtype=dtype)
        else:
            flat = torch.frombuffer(bytearray(memoryview(b)[pos + 8:]), dtype=dtype)
        return _FlatStorage(flat)
    except Exception:
        return slow(b)


class _FlatStorage:
    """What _storage_from_legacy_bytes hands to the tensor rebuild: the storage's elements as a
    flat tensor."""
    __slots__ = ("flat",)

    def __init__(self, flat: torch.Tensor):
        self.flat = flat

    def typed_storage(self):
        return torch.storage.TypedStorage(wrap_storage=self.flat.untyped_storage(),
                                          dtype=self.flat.dtype, _internal=True)


def _rebuild_tensor_fast(storage, storage_offset, size, stride, requires_grad=False,
                         backward_hooks=None, metadata=None):
    """torch._utils._rebuild_tensor_v2 for the common case (plain tensor, no autograd state):
    a strided view of the flat tensor, without torch's per-tensor fake-mode detection."""
    if isinstance(storage, _FlatStorage):
        if not requires_grad and not backward_hooks and not metadata:
            return storage.flat.as_strided(tuple(size), tuple(stride), storage_offset)
        storage = storage.typed_storage()
    return torch._utils._rebuild_tensor_v2(storage, storage_offset, size, stride, requires_grad,
                                           backward_hooks, metadata)


class _FastTensorUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == "torch.storage" and name == "_load_from_bytes":
            return _storage_from_legacy_bytes
        if module == "torch._utils" and name == "_rebuild_tensor_v2":
            return _rebuild_tensor_fast
        return super().find_class(module, name)


def _read_records(paths: Sequence[str]) -> List[dict]:
    import os
    fast = os.environ.get("ZSAAC_FAST_UNPICKLE", "1") != "0"
    all_data: List[dict] = list()
    for path in paths:
        with open(path, "rb") as f:
            # reference :11-12 / wavcaps :11-13 (pickle.load of one list per file)
            all_data = all_data + (_FastTensorUnpickler(f).load() if fast else pickle.load(f))
    return all_data


def load_data(raw_path: Union[str, Sequence[str]]) -> Tuple[torch.Tensor, List[dict]]:
    """(bank, all_data): bank = fp32 [N, d] unit rows on the GPU, all_data = the pickled records.

    Reference: embeddings_related_generator.py:9-17 (raw_path: str) and
    embeddings_related_generator_wavcaps.py:9-18 (raw_path: list of str).  The reference routes
    the rows through `set()` (:15), which keeps value-duplicates (tensors hash by identity) and
    scrambles the order; here the bank keeps input order, which only affects how exact ties are
    ordered.  Val/test records carry `text_embedding: 0` (embeddings_generator.py:72) and fail
    here exactly as they do in the reference (int has no .cpu()).
    """
    _require_cuda()
    paths = [raw_path] if isinstance(raw_path, (str, bytes)) else list(raw_path)
    all_data = _read_records(paths)
    all_captions = [raw_data["text_embedding"].cpu() for raw_data in all_data]      # :14
    host = torch.cat(all_captions, dim=0).to(torch.float32).contiguous().pin_memory()
    dev = host.to("cuda", non_blocking=True)                                        # :15
    # F.normalize(..., dim=-1) (:17) through the native library; the bf16 search copy is built
    # from the same tensor lazily by process_data
    rb = RelatedBank(dev.shape[0], dev.shape[1], device=dev.device)
    bank = rb.normalize_rows(dev)
    rb.upload(bank, 0, normalize=True)
    _register_bank(bank, rb)
    return bank, all_data


def _register_bank(bank: torch.Tensor, rb: RelatedBank) -> None:
    """Let process_data find the bf16 copy load_data already built for this tensor."""
    from . import retrieval
    retrieval._cache_insert(retrieval._cache_key(bank, True), bank, rb)


def process_data(valid_text_embs: torch.Tensor, all_data: Iterable[dict], topnumber: int,
                 *, exclude_self: bool = False, rescore_fp32: bool = True) -> Iterator[dict]:
    """Yield every item with `related_embeddings` = its top-`topnumber` bank rows, best first.

    Reference: embeddings_related_generator.py:19-28.  valid_text_embs is the fp32 bank returned
    by load_data (CUDA).  For each item the query is F.normalize(item['text_embedding']) (:21),
    ranked by cosine similarity against the bank (:22); the k rows are gathered from
    valid_text_embs itself (:23): exact copies of the caller's bank rows, whatever precision the
    search ran in.  item['text_embedding'] is moved to the CPU (:25).  Like the reference there is no
    self-exclusion unless exclude_self=True (opt-in; assumes item i is bank row i).
    rescore_fp32 (default on): the fused kernel ranks bf16-rounded operands; its k + 8 best
    candidates are re-scored in fp32 from valid_text_embs, so the k rows kept are the reference's
    fp32 choice also where two captions are closer than the bf16 resolution (~1e-4).
    """
    _require_cuda()
    if not valid_text_embs.is_cuda:
        raise ValueError("valid_text_embs must be the CUDA bank returned by load_data "
                         "(no CPU path exists)")
    topnumber = int(topnumber)
    rb = bank_for(valid_text_embs, normalize=True)
    device = valid_text_embs.device
    d = valid_text_embs.shape[1]
    batch: List[dict] = []
    base = 0
    # pinned staging buffers, allocated once (page-locking 64 + 335 MB per batch of 16,384 queries
    # costs more than the search); every flush synchronises the stream before it yields, so the
    # next flush may overwrite them
    stage = {}

    def pinned(name: str, shape) -> torch.Tensor:
        buf = stage.get(name)
        if buf is None or buf.shape[0] < shape[0] or tuple(buf.shape[1:]) != tuple(shape[1:]):
            # the first batch is the largest one (full batches first, then the ragged tail)
            buf = torch.empty(tuple(shape), dtype=torch.float32).pin_memory()
            stage[name] = buf
        return buf[:shape[0]]

    def flush(items: List[dict], first: int) -> Iterator[dict]:
        q_rows = torch.cat([it["text_embedding"].detach().cpu().reshape(1, -1) for it in items], dim=0)
        q_host = pinned("queries", q_rows.shape)
        q_host.copy_(q_rows)                               # casts to fp32 if the records are not
        q_dev = q_host.to(device, non_blocking=True)
        self_index = None
        if exclude_self:
            self_index = torch.arange(first, first + len(items), dtype=torch.int64, device=device)
        if rescore_fp32:
            _, ids = search_rescored(rb, q_dev, valid_text_embs, topnumber, normalize=True,
                                     self_index=self_index)
        else:
            _, ids = rb.search(q_dev, topnumber, normalize_queries=True, self_index=self_index)
        related = rb.gather_rows(valid_text_embs, ids)                 # [B, k, d] fp32 on the GPU
        related_host = pinned("related", related.shape)
        related_host.copy_(related, non_blocking=True)
        torch.cuda.current_stream(device).synchronize()
        for j, item in enumerate(items):
            item["text_embedding"] = item["text_embedding"].cpu()      # :25
            # own storage per record: a view would pickle the whole batch buffer with each item
            item["related_embeddings"] = related_host[j].clone()       # :26  [k, d] fp32 CPU
            yield item

    for item in all_data:
        batch.append(item)
        if len(batch) == QUERY_BATCH:
            yield from flush(batch, base)
            base += len(batch)
            batch = []
    if batch:
        yield from flush(batch, base)


# records pickled per parallel round of the optional multi-process writer
WRITER_BATCH = 8192
_WRITER_ITEMS: List[dict] = []      # read by forked writer processes (copy-on-write, never sent)
_WRITER_FAST = False

_NUMPY_DTYPES = (torch.float32, torch.float64, torch.float16, torch.int64, torch.int32, torch.int16,
                 torch.int8, torch.uint8, torch.bool)


class _FastTensorPickler(pickle.Pickler):
    """Opt-in record writer: plain CPU tensors are pickled as `torch.from_numpy(ndarray)` instead
    of torch's own reduction (which runs torch.save on every storage: ~100 us per tensor).  The
    stream is still one pickle per record and `pickle.load` still returns the same dicts of
    torch.Tensor (same dtype / shape / values, writable) — the reference's reader
    (dataset/dataset.py:64-78) needs no change and reads such a file ~4x faster too — but the
    bytes on disk are not the ones torch's reduction would have produced, hence opt-in."""

    def reducer_override(self, obj):
        if (isinstance(obj, torch.Tensor) and type(obj) is torch.Tensor and obj.device.type == "cpu"
                and obj.layout == torch.strided and not obj.requires_grad
                and obj.dtype in _NUMPY_DTYPES):
            return (torch.from_numpy, (obj.detach().numpy(),))
        return NotImplemented


def _dump_record(item: dict, file, fast: bool) -> None:
    if fast:
        _FastTensorPickler(file, protocol=pickle.DEFAULT_PROTOCOL).dump(item)
    else:
        pickle.dump(item, file)                          # reference :34


def _pickle_slice(args) -> str:
    """Worker: pickle items[lo:hi] of the inherited batch into a part file; returns its path."""
    lo, hi, part_path = args
    with open(part_path, "wb") as f:
        for item in _WRITER_ITEMS[lo:hi]:
            _dump_record(item, f, _WRITER_FAST)
    return part_path


def _write_batch_parallel(items: List[dict], file, workers: int, tmp_prefix: str,
                          fast: bool = False) -> None:
    """Pickle `items` in `workers` forked processes (pickling tensors is GIL-bound Python, ~80 us
    per record) and append the parts to `file` in order.  The bytes are exactly what the serial
    loop would have written."""
    import multiprocessing as mp
    import os
    import shutil
    global _WRITER_ITEMS, _WRITER_FAST
    _WRITER_ITEMS = items
    _WRITER_FAST = fast
    n = len(items)
    per = -(-n // workers)
    jobs = [(lo, min(lo + per, n), f"{tmp_prefix}.part{j}") for j, lo in enumerate(range(0, n, per))]
    try:
        # fork: children see the batch through copy-on-write memory; they only touch CPU tensors
        with mp.get_context("fork").Pool(len(jobs)) as pool:
            parts = pool.map(_pickle_slice, jobs)
        for part in parts:
            with open(part, "rb") as src:
                shutil.copyfileobj(src, file, length=16 << 20)
    finally:
        _WRITER_ITEMS = []
        for _, _, part in jobs:
            if os.path.exists(part):
                os.remove(part)


def save_data_to_hdf5(processed_data_gen: Iterable[dict], output_path: str, total_items: int,
                      *, workers: int = None, fast_pickle: bool = None) -> None:
    """Append one pickle per record to output_path (reference :30-34; the name is historical —
    the format is a pickle stream, read back by dataset/dataset.py:64-78).

    workers (default: env ZSAAC_WRITER_PROCS, else 0): 0 reproduces the reference's serial loop;
    N > 0 pickles batches of records in N forked processes — same bytes, same order — because
    once the search takes milliseconds the per-record pickling dominates the script.
    fast_pickle (default: env ZSAAC_FAST_PICKLE=1, else off): pickle tensors through numpy
    (_FastTensorPickler): ~3x faster to write and ~4x faster to read back, same objects after
    pickle.load, different bytes on disk."""
    import os
    if workers is None:
        workers = int(os.environ.get("ZSAAC_WRITER_PROCS", "0"))
    if fast_pickle is None:
        fast_pickle = os.environ.get("ZSAAC_FAST_PICKLE", "0") == "1"
    with open(output_path, "ab") as file:
        if workers <= 0:
            for i, item in enumerate(tqdm(processed_data_gen, total=total_items)):
                _dump_record(item, file, fast_pickle)
            return
        batch: List[dict] = []
        progress = tqdm(total=total_items)
        for item in processed_data_gen:
            batch.append(item)
            if len(batch) == WRITER_BATCH:
                _write_batch_parallel(batch, file, workers, output_path, fast_pickle)
                progress.update(len(batch))
                batch = []
        if batch:
            _write_batch_parallel(batch, file, workers, output_path, fast_pickle)
            progress.update(len(batch))
        progress.close()
