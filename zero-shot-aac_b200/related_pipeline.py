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

from .retrieval import RelatedBank, _require_cuda, bank_for

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


# Streams already parsed once, as (head, middle, tail, dtype, element count, element size): what
# precedes, separates and follows the two occurrences of the storage key.  A stream that matches
# one of them byte for byte needs no unpickling at all — only the key and the data differ between
# the tensors of a file.  Most recently matched first; a handful of entries (one per dtype and
# element count seen).
_READ_TEMPLATES: list = []
_READ_TEMPLATES_MAX = 16


def _match_read_template(b: bytes):
    """The flat tensor of a stream that matches a known template, else None."""
    for slot, (head, middle, tail, dtype, count, size) in enumerate(_READ_TEMPLATES):
        if not b.startswith(head):
            continue
        p = len(head)
        if len(b) < p + 5 or b[p] != 0x58:                         # BINUNICODE: 'X' + uint32 length
            continue
        field_end = p + 5 + int.from_bytes(b[p + 1:p + 5], "little")
        field = b[p:field_end]
        if not field[5:].isdigit():
            continue
        q = field_end + len(middle)
        r = q + len(field) + len(tail)
        if (len(b) != r + 8 + count * size or b[field_end:q] != middle or b[q:q + len(field)] != field
                or b[q + len(field):r] != tail or int.from_bytes(b[r:r + 8], "little") != count):
            continue
        if slot:
            _READ_TEMPLATES.insert(0, _READ_TEMPLATES.pop(slot))
        if count == 0:
            return torch.empty(0, dtype=dtype)
        return torch.frombuffer(bytearray(memoryview(b)[r + 8:]), dtype=dtype)
    return None


def _remember_read_template(b: bytes, key: str, pos: int, dtype, count: int, size: int) -> None:
    import struct
    if not isinstance(key, str) or not key.isdigit():
        return
    ident = key.encode()
    parts = b[:pos].split(b"X" + struct.pack("<I", len(ident)) + ident)
    if len(parts) == 3:
        _READ_TEMPLATES.insert(0, (parts[0], parts[1], parts[2], dtype, count, size))
        del _READ_TEMPLATES[_READ_TEMPLATES_MAX:]


def _storage_from_legacy_bytes(b: bytes):
    import io
    import struct
    slow = torch.storage._load_from_bytes
    try:
        flat = _match_read_template(b)
        if flat is not None:
            return _FlatStorage(flat)
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
        _remember_read_template(b, pid[2], pos, dtype, count, size)
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


def _dist_info():
    """(torch.distributed module or None, rank, world) — multi-GPU mode is on when the script was
    launched with one process per GPU (torchrun) and the process group is up."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def item_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of items rank `rank` processes (and writes) in multi-GPU mode: the rank
    files concatenated in rank order are the single-GPU stream."""
    per = -(-n_items // world)
    return min(rank * per, n_items), min((rank + 1) * per, n_items)


def load_data(raw_path: Union[str, Sequence[str]]) -> Tuple[torch.Tensor, List[dict]]:
    """(bank, all_data): bank = fp32 [N, d] unit rows on the GPU, all_data = the pickled records.

    Reference: embeddings_related_generator.py:9-17 (raw_path: str) and
    embeddings_related_generator_wavcaps.py:9-18 (raw_path: list of str).  The reference routes
    the rows through `set()` (:15), which keeps value-duplicates (tensors hash by identity) and
    scrambles the order; here the bank keeps input order, which only affects how exact ties are
    ordered.  Val/test records carry `text_embedding: 0` (embeddings_generator.py:72) and fail
    here exactly as they do in the reference (int has no .cpu()).
    In multi-GPU mode every rank reads the file(s) and holds the fp32 bank (rows are gathered and
    re-scored from it); the bf16 search copy is built per shard by process_data.
    """
    _require_cuda()
    paths = [raw_path] if isinstance(raw_path, (str, bytes)) else list(raw_path)
    all_data = _read_records(paths)
    all_captions = [raw_data["text_embedding"].cpu() for raw_data in all_data]      # :14
    host = torch.cat(all_captions, dim=0).to(torch.float32).contiguous().pin_memory()
    dev = host.to("cuda", non_blocking=True)                                        # :15
    # F.normalize(..., dim=-1) (:17) through the native library
    if _dist_info()[2] > 1:
        from .retrieval import helper_context
        return helper_context(dev.device).normalize_rows(dev), all_data
    # single GPU: the bf16 search copy is built from the same tensor right away
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
                 *, exclude_self: bool = False, rescore_fp32: bool = True,
                 dtype: str = "bf16") -> Iterator[dict]:
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
    dtype="fp32" ranks exact fp32 scores instead (banks up to 65,536 rows, single GPU).

    Batches of queries go through a three-stream pipeline (sharded.SearchPipeline): while the
    caller pickles the records of batch i, batch i+1 is already being uploaded and searched.
    Multi-GPU mode (one process per GPU, process group up): the bank is sharded by rows over the
    ranks, rank r processes — and yields — the contiguous block item_range(len(all_data), r, G);
    each step all-gathers every rank's query slice, every rank searches (and fp32-re-scores) its
    shard, an all-to-all returns to each rank the shard lists of its own queries, merged there.
    """
    _require_cuda()
    if not valid_text_embs.is_cuda:
        raise ValueError("valid_text_embs must be the CUDA bank returned by load_data "
                         "(no CPU path exists)")
    if dtype not in ("bf16", "fp32"):
        raise ValueError(f"dtype must be 'bf16' or 'fp32', got {dtype!r}")
    topnumber = int(topnumber)
    dist, rank, world = _dist_info()
    if dtype == "fp32":
        if world > 1:
            raise ValueError("dtype='fp32' (exact small-bank path) runs on one GPU")
        yield from _process_exact(valid_text_embs, all_data, topnumber, exclude_self)
        return
    device = valid_text_embs.device
    n_bank, d = valid_text_embs.shape
    from .sharded import SearchPipeline, ShardedRelatedBank

    if world > 1:
        all_data = all_data if isinstance(all_data, list) else list(all_data)
        first, last = item_range(len(all_data), rank, world)
        per = max(1, min(QUERY_BATCH // world, -(-len(all_data) // world)))
        n_steps = -(-(-(-len(all_data) // world)) // per)          # same on every rank: collectives inside
        bank = ShardedRelatedBank(n_bank, d, device=device)
        bank.upload_global(valid_text_embs, normalize=True)
        shard_f32 = valid_text_embs[bank.lo:bank.hi]
        source = iter(all_data[first:last])
    else:
        first = 0
        known = len(all_data) if hasattr(all_data, "__len__") else QUERY_BATCH
        per = max(1, min(QUERY_BATCH, known))
        n_steps = None                                             # until the iterable runs dry
        bank = bank_for(valid_text_embs, normalize=True)
        shard_f32 = valid_text_embs
        source = iter(all_data)
    helper = bank.local if world > 1 else bank
    pipe = SearchPipeline(bank, world * per, topnumber, depth=2, from_host=True, to_host=False,
                          result="row_slice", rescore_from=shard_f32 if rescore_fp32 else None,
                          excludes_self=exclude_self, input="slice")
    # pinned staging, allocated once (page-locking costs more than the search); a step's buffers
    # are free again once its records have been yielded
    q_stage = [torch.zeros((world * per, d), dtype=torch.float32).pin_memory() for _ in range(2)]
    r_stage = [torch.empty((per, topnumber, d), dtype=torch.float32).pin_memory() for _ in range(2)]

    def take(n: int) -> List[dict]:
        out = []
        for item in source:
            out.append(item)
            if len(out) == n:
                break
        return out

    def submit(step: int, items: List[dict], base: int):
        buf = q_stage[step % 2]
        lo = rank * per
        if items:
            rows = torch.cat([it["text_embedding"].detach().cpu().reshape(1, -1) for it in items], dim=0)
            buf[lo:lo + len(items)].copy_(rows)                   # casts to fp32 if the records are not
        if len(items) < per:
            buf[lo + len(items):lo + per].zero_()                  # ragged tail: padding queries, results dropped
        self_index = None
        if exclude_self:
            if world > 1:
                # global item index of every query row of the all-gathered batch: rank g's slice
                # holds its items [first_g + step*per, ...)
                starts = torch.tensor([item_range(len(all_data), g, world)[0] for g in range(world)])
                self_index = (starts.unsqueeze(1) + step * per + torch.arange(per).unsqueeze(0)).reshape(-1)
            else:
                self_index = base + torch.arange(per)
        return pipe.submit(buf, self_index=self_index)

    step, base = 0, first
    items = take(per)
    slot = submit(0, items, base) if (items or world > 1) else None
    while slot is not None:
        nxt_items = take(per)
        more = (step + 1 < n_steps) if n_steps is not None else bool(nxt_items)
        nxt_slot = submit(step + 1, nxt_items, base + len(items)) if more else None
        pipe.wait_stream(slot)
        _, ids = pipe.result_of(slot)                                  # [per, k] global indices, this rank's queries
        if items:
            related = helper.gather_rows(valid_text_embs, ids[:len(items)])   # [B, k, d] fp32 on the GPU
            related_host = r_stage[step % 2][:len(items)]
            related_host.copy_(related, non_blocking=True)
            torch.cuda.current_stream(device).synchronize()
            for j, item in enumerate(items):
                item["text_embedding"] = item["text_embedding"].cpu()      # :25
                # own storage per record: a view would pickle the whole batch buffer with each item
                item["related_embeddings"] = related_host[j].clone()       # :26  [k, d] fp32 CPU
                yield item
        base += len(items)
        step, items, slot = step + 1, nxt_items, nxt_slot
    torch.cuda.synchronize(device)
    if world > 1:
        bank.local.close()


def _process_exact(valid_text_embs: torch.Tensor, all_data: Iterable[dict], topnumber: int,
                   exclude_self: bool) -> Iterator[dict]:
    """dtype='fp32': exact fp32 ranking on CUDA cores (zs_exact_topk_f32), small banks only."""
    from .retrieval import exact_fits, exact_topk, helper_context
    device = valid_text_embs.device
    batch_rows = max(1, min(4096, (1 << 28) // max(1, valid_text_embs.shape[0])))
    if not exact_fits(batch_rows, valid_text_embs.shape[0]):
        raise ValueError("dtype='fp32' is limited to banks of at most 65,536 rows; use the default "
                         "bf16 search with fp32 re-scoring")
    helper = helper_context(device)
    batch: List[dict] = []
    base = 0

    def flush(items: List[dict], first: int) -> Iterator[dict]:
        q = torch.cat([it["text_embedding"].detach().cpu().reshape(1, -1) for it in items], dim=0)
        self_index = torch.arange(first, first + len(items)) if exclude_self else None
        _, ids = exact_topk(q.to(device), valid_text_embs, topnumber, normalize=True, self_index=self_index)
        related_host = helper.gather_rows(valid_text_embs, ids).cpu()
        for j, item in enumerate(items):
            item["text_embedding"] = item["text_embedding"].cpu()
            item["related_embeddings"] = related_host[j].clone()
            yield item

    for item in all_data:
        batch.append(item)
        if len(batch) == batch_rows:
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


# ------------------------------------------------------------------------------------------------
# The default writer.  `pickle.dump(item, file)` (reference :34) spends ~100 us per tensor inside
# torch's reduction: Tensor.__reduce_ex__ wraps the storage in a TypedStorage whose __reduce__ runs
# a complete legacy torch.save (three header pickles, a persistent-id pickler, a key list) into a
# BytesIO, and the resulting bytes are what the record pickle embeds.  For a plain CPU tensor that
# byte string is a fixed template around two things: the storage key (str(storage._cdata), twice)
# and the raw data.  The reducer below takes the template ONCE per (dtype, element count) from
# torch.save itself — so it follows whatever the installed torch writes — and then fills it in:
# the bytes on disk are the very bytes `pickle.dump` writes (tests/test_host_cpu.py compares the
# streams), ~5x faster.  Everything that is not a plain tensor (subclasses, autograd state, Python
# attributes, names, conj/neg bits, other devices or layouts, dtypes of the newer storage format)
# goes through torch's own reduction, and a template that does not reproduce torch's bytes for the
# first tensor it is used on is dropped.  ZSAAC_TEMPLATE_PICKLE=0 restores the literal
# `pickle.dump`.
_TEMPLATE_DTYPES = frozenset(_STORAGE_DTYPES.values())
# beyond this the ~100 us of torch's own path no longer matter, and a template costs a probe storage
# of the same size
_TEMPLATE_MAX_BYTES = 4 << 20
_STORAGE_TEMPLATES: dict = {}       # (dtype, element count) -> (head, middle, tail) | False


class _LegacyStorageBytes:
    """Pickles as `torch.storage._load_from_bytes(payload)`, like torch's TypedStorage.__reduce__."""
    __slots__ = ("payload",)

    def __init__(self, payload: bytes):
        self.payload = payload


def _reduce_storage_bytes(stub: _LegacyStorageBytes):
    return (torch.storage._load_from_bytes, (stub.payload,))


def _key_field(cdata: int, _pack=__import__("struct").Struct("<I").pack) -> bytes:
    ident = str(cdata).encode()
    return b"X" + _pack(len(ident)) + ident                      # BINUNICODE, as protocol 2 writes it


def _storage_template(dtype: torch.dtype, numel: int):
    """Split what torch.save(storage) writes for a CPU storage of this dtype and element count
    around the two occurrences of the storage key; False if the stream does not have that shape."""
    import io
    probe = torch.empty(numel, dtype=dtype)._typed_storage()
    buf = io.BytesIO()
    torch.save(probe, buf, _use_new_zipfile_serialization=False)
    raw = buf.getvalue()
    nbytes = numel * torch._utils._element_size(dtype)
    parts = raw[:len(raw) - nbytes].split(_key_field(probe._cdata))
    return tuple(parts) if len(parts) == 3 else False


# (looked up once: this runs three times per record)
_get_obj_state = torch._utils._get_obj_state
_get_tensor_metadata = torch._C._get_tensor_metadata          # what torch._utils.get_tensor_metadata calls
_serialization_tls = torch.serialization._serialization_tls
_rebuild_tensor_v2 = torch._utils._rebuild_tensor_v2
_strided = torch.strided


def _reduce_plain_tensor(t: torch.Tensor, _string_at=__import__("ctypes").string_at,
                         _ordered_dict=__import__("collections").OrderedDict):
    """dispatch_table entry for exactly torch.Tensor: the tuple Tensor.__reduce_ex__ returns, with
    the storage bytes filled into the template instead of produced by a torch.save call."""
    if (t.device.type == "cpu" and t.layout is _strided and t.dtype in _TEMPLATE_DTYPES
            and not t.requires_grad and not t.has_names() and not _get_obj_state(t)
            and not _get_tensor_metadata(t) and not _serialization_tls.skip_data):
        storage = t.untyped_storage()
        nbytes = storage.nbytes()
        if nbytes > _TEMPLATE_MAX_BYTES:
            return t.__reduce_ex__(pickle.DEFAULT_PROTOCOL)
        size = t.element_size()
        numel = nbytes // size
        key = (t.dtype, numel)
        template = _STORAGE_TEMPLATES.get(key)
        first_use = template is None
        if first_use:
            template = _storage_template(t.dtype, numel) if numel * size == nbytes else False
        if template:
            field = _key_field(storage._cdata)
            payload = b"".join((template[0], field, template[1], field, template[2],
                                _string_at(storage.data_ptr(), nbytes) if nbytes else b""))
            if first_use and payload != t._typed_storage().__reduce__()[1][0]:
                template = False                                   # not torch's bytes: never used
        if first_use:
            if len(_STORAGE_TEMPLATES) >= 4096:       # records of ever-changing sizes: stay bounded
                _STORAGE_TEMPLATES.clear()
            _STORAGE_TEMPLATES[key] = template
        if template:
            return (_rebuild_tensor_v2,
                    (_LegacyStorageBytes(payload), t.storage_offset(), tuple(t.size()), t.stride(),
                     False, _ordered_dict()))
    return t.__reduce_ex__(pickle.DEFAULT_PROTOCOL)


def _template_dump(item, file, copyreg=__import__("copyreg")) -> None:
    pickler = pickle.Pickler(file)
    table = dict(copyreg.dispatch_table)
    table[torch.Tensor] = _reduce_plain_tensor
    table[_LegacyStorageBytes] = _reduce_storage_bytes
    pickler.dispatch_table = table
    pickler.dump(item)


def _template_pickle_enabled(_environ=__import__("os").environ) -> bool:
    return _environ.get("ZSAAC_TEMPLATE_PICKLE", "1") != "0"


def _dump_record(item: dict, file, fast: bool) -> None:
    if fast:
        _FastTensorPickler(file, protocol=pickle.DEFAULT_PROTOCOL).dump(item)
    elif _template_pickle_enabled():
        _template_dump(item, file)                       # the bytes of reference :34, faster
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
    loop would have written.  The children are forked per batch so that they see the batch through
    copy-on-write memory (sending the records to a long-lived pool would cost the very pickling
    this is meant to parallelise); they touch CPU tensors only, never CUDA.  Opt-in: forking a
    multi-threaded process is at the caller's risk — `--gpus N` gets the same parallelism from
    the N rank processes without forking."""
    import multiprocessing as mp
    import os
    import shutil
    import tempfile
    global _WRITER_ITEMS, _WRITER_FAST
    _WRITER_ITEMS = items
    _WRITER_FAST = fast
    n = len(items)
    per = -(-n // workers)
    # part files in a private directory next to the output (two jobs sharing an output prefix
    # must not collide: ADVICE r1)
    part_dir = tempfile.mkdtemp(prefix=".zsaac_parts_", dir=os.path.dirname(os.path.abspath(tmp_prefix)) or ".")
    jobs = [(lo, min(lo + per, n), os.path.join(part_dir, f"part{j}")) for j, lo in enumerate(range(0, n, per))]
    try:
        # fork: children see the batch through copy-on-write memory; they only touch CPU tensors
        with mp.get_context("fork").Pool(len(jobs)) as pool:
            parts = pool.map(_pickle_slice, jobs)
        for part in parts:
            with open(part, "rb") as src:
                shutil.copyfileobj(src, file, length=16 << 20)
    finally:
        _WRITER_ITEMS = []
        shutil.rmtree(part_dir, ignore_errors=True)


def save_data_to_hdf5(processed_data_gen: Iterable[dict], output_path: str, total_items: int,
                      *, workers: int = None, fast_pickle: bool = None) -> None:
    """Append one pickle per record to output_path (reference :30-34; the name is historical —
    the format is a pickle stream, read back by dataset/dataset.py:64-78).

    workers (default: env ZSAAC_WRITER_PROCS, else 0): 0 reproduces the reference's serial loop;
    N > 0 pickles batches of records in N forked processes — same bytes, same order — because
    once the search takes milliseconds the per-record pickling dominates the script.
    fast_pickle (default: env ZSAAC_FAST_PICKLE=1, else off): pickle tensors through numpy
    (_FastTensorPickler): ~3x faster to write and ~4x faster to read back, same objects after
    pickle.load, different bytes on disk.
    Multi-GPU mode: every rank writes the records it processed (process_data yields only those)
    to `<output_path>.rank<r>`, the sizes are exchanged, and every rank copies its file to its own
    offset of output_path (rank order = item order: the same stream one GPU writes; an existing
    file is appended to, like the reference's 'ab').  The G ranks pickle and copy in parallel,
    which is what makes the script scale (pickling, not the search, is the wall clock)."""
    import os
    import shutil
    if workers is None:
        workers = int(os.environ.get("ZSAAC_WRITER_PROCS", "0"))
    if fast_pickle is None:
        fast_pickle = os.environ.get("ZSAAC_FAST_PICKLE", "0") == "1"
    dist, rank, world = _dist_info()
    if world > 1:
        part = f"{output_path}.rank{rank:03d}"
        first, last = item_range(total_items, rank, world)
        _write_stream(processed_data_gen, part, "wb", last - first, workers, fast_pickle)
        # every rank copies its own file to its own offset of the output (appended after whatever
        # the file already holds, like the reference's 'ab'): the G copies run in parallel
        sizes = [None] * world
        dist.all_gather_object(sizes, os.path.getsize(part))
        base = [0]
        if rank == 0:
            base[0] = os.path.getsize(output_path) if os.path.exists(output_path) else 0
            with open(output_path, "ab") as out:
                out.truncate(base[0] + sum(sizes))
        dist.broadcast_object_list(base, src=0)
        with open(part, "rb") as src, open(output_path, "r+b") as out:
            out.seek(base[0] + sum(sizes[:rank]))
            shutil.copyfileobj(src, out, length=64 << 20)
        os.remove(part)
        dist.barrier()
        return
    _write_stream(processed_data_gen, output_path, "ab", total_items, workers, fast_pickle)


def _write_stream(processed_data_gen: Iterable[dict], path: str, mode: str, total_items: int,
                  workers: int, fast_pickle: bool) -> None:
    with open(path, mode) as file:
        if workers <= 0:
            for i, item in enumerate(tqdm(processed_data_gen, total=total_items)):
                _dump_record(item, file, fast_pickle)
            return
        batch: List[dict] = []
        progress = tqdm(total=total_items)
        for item in processed_data_gen:
            batch.append(item)
            if len(batch) == WRITER_BATCH:
                _write_batch_parallel(batch, file, workers, path, fast_pickle)
                progress.update(len(batch))
                batch = []
        if batch:
            _write_batch_parallel(batch, file, workers, path, fast_pickle)
            progress.update(len(batch))
        progress.close()


# ------------------------------------------------------------------------------------------------
# CLI shared by the two module mirrors (reference :41-53 / _wavcaps.py:42-54)
def add_extension_flags(parser) -> None:
    """Flags the reference does not have (all optional; without them the script behaves like the
    reference's: one GPU, no self-exclusion, per-record pickles)."""
    import argparse
    parser.add_argument('--gpus', type=int, default=1,
                        help="GPUs of this node to shard the bank over (one process per GPU; the script "
                             "re-launches itself under torch.distributed.run when it is not already)")
    parser.add_argument('--exclude_self', action='store_true',
                        help="do not return item i's own bank row i (the reference keeps it in slot 0)")
    parser.add_argument('--rescore_fp32', action=argparse.BooleanOptionalAction, default=True,
                        help="re-score the k + 8 bf16 candidates in fp32 (the reference's precision)")
    parser.add_argument('--dtype', choices=['bf16', 'fp32'], default='bf16',
                        help="bf16 tensor-core search (default) or exact fp32 ranking (banks <= 65,536 rows)")
    # pickle the output records in N forked processes (byte-identical stream)
    parser.add_argument('--writer_procs', type=int, default=None)
    # pickle tensors through numpy (same objects after pickle.load, ~3x faster)
    parser.add_argument('--fast_pickle', action='store_true', default=None)


def run_cli(args, module: str, argv, load_fn) -> None:
    """Body of main() for both mirrors: optional self-launch on N GPUs, then the reference's four
    lines (load_data -> process_data -> save_data_to_hdf5)."""
    import os
    import sys
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world_env == 1:
        import socket
        import subprocess
        with socket.socket() as sock:
            sock.bind(("127.0.0.1", 0))
            port = sock.getsockname()[1]
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), "-m", module,
               *(sys.argv[1:] if argv is None else list(argv))]
        raise SystemExit(subprocess.run(cmd).returncode)
    if world_env > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    import time
    t0 = time.perf_counter()
    valid_text_embs, all_data = load_fn(args.input_path)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    processed_data_gen = process_data(valid_text_embs, all_data, args.topnumber,
                                      exclude_self=args.exclude_self, rescore_fp32=args.rescore_fp32,
                                      dtype=args.dtype)
    total_items = len(all_data)
    save_data_to_hdf5(processed_data_gen, args.output_path, total_items, workers=args.writer_procs,
                      fast_pickle=args.fast_pickle)
    t2 = time.perf_counter()
    if os.environ.get("ZSAAC_TIMING") == "1" and int(os.environ.get("RANK", "0")) == 0:
        import json
        print(json.dumps({"zsaac_timing": {"records": total_items, "gpus": world_env,
                                           "load_data_s": round(t1 - t0, 3),
                                           "process_and_save_s": round(t2 - t1, 3)}}), flush=True)
    if world_env > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
