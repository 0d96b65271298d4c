"""Mirror of the memory-projection helpers of the reference's predict_prompt.py.

`map2memory` (predict_prompt.py:23-29) and `construct_support_memory` (:30-56) are defined in the
reference but the call is commented out (:134); they are the training-free alternative to the
related-caption prefix: project the audio embedding onto the span of the caption memory with
softmax(100 * similarity) weights.  Same contraction as the retrieval path with a softmax-weighted
reduction instead of a top-k, so it lives on the same library (zs_memory_project).
"""
from __future__ import annotations

from typing import List, Sequence

import torch

from . import _abi
from .retrieval import RelatedBank, _require_cuda, _stream_ptr, helper_context

TEMPERATURE = 100.0   # predict_prompt.py:26  `(sim*100).softmax(dim=-1)`

def _helper(device: torch.device) -> RelatedBank:
    """A context for the bank-less entry points (normalise, memory projection) on `device`."""
    return helper_context(device)


BATCHED_MIN_QUERIES = 8       # from here on the tensor-core path wins over one streaming pass per query
_MEMORY_BANKS: "dict[tuple, tuple]" = {}     # key -> (weakref to text_features, prepared context)
_MEMORY_BANKS_MAX = 2


def _prepared_bank(tf: torch.Tensor) -> RelatedBank:
    """A context holding the split-bf16 operand copies of `tf` (zs_memory_bank_prepare), cached on
    the tensor's identity and version like the search bank."""
    import weakref
    key = (tf.data_ptr(), tuple(tf.shape), str(tf.device), tf._version)
    hit = _MEMORY_BANKS.get(key)
    if hit is not None and hit[0]() is tf and not hit[1].closed:
        return hit[1]
    ctx = RelatedBank(1, 64, device=tf.device)
    with torch.cuda.device(tf.device):
        _abi.check(ctx._lib.zs_memory_bank_prepare(ctx._ctx, tf.data_ptr(), tf.shape[0], tf.shape[1],
                                                   _stream_ptr(tf.device)))
    for stale in [k for k, (ref, _) in _MEMORY_BANKS.items() if ref() is None]:
        _MEMORY_BANKS.pop(stale)
    while len(_MEMORY_BANKS) >= _MEMORY_BANKS_MAX:
        _MEMORY_BANKS.pop(next(iter(_MEMORY_BANKS)))
    _MEMORY_BANKS[key] = (weakref.ref(tf), ctx)
    return ctx


def map2memory(audio_embed: torch.Tensor, text_features: torch.Tensor) -> torch.Tensor:
    """prefix_embedding = normalise(softmax(100 * audio_embed @ text_features.T) @ text_features).

    audio_embed [Q, d] (the reference passes one embedding, [1, 1024]); text_features [N, d]
    float32 unit rows (construct_support_memory).  Returns float32 [Q, d] on the GPU.  CPU inputs
    are copied to the current CUDA device; there is no CPU path.
    Up to 7 queries stream the fp32 bank once per query (HBM-bound, zs_memory_project); batches of
    8 or more run both contractions on the tensor cores with split-bf16 operands
    (zs_memory_project_batched; the operand copies of `text_features` — 12 bytes per element — are
    built on the first call with that tensor and cached)."""
    _require_cuda()
    tf = text_features.detach()
    if not tf.is_cuda:
        tf = tf.to("cuda")
    if tf.dtype != torch.float32 or not tf.is_contiguous():
        tf = tf.to(torch.float32).contiguous()
    q = audio_embed.detach().to(device=tf.device, dtype=torch.float32)
    lead = tuple(q.shape[:-1])
    q = q.reshape(-1, q.shape[-1]).contiguous()
    if q.shape[1] != tf.shape[1]:
        raise ValueError(f"audio_embed has d={q.shape[1]}, text_features d={tf.shape[1]}")
    out = torch.empty_like(q)
    if q.shape[0] >= BATCHED_MIN_QUERIES and tf.shape[1] % _abi.ZS_DIM_MULTIPLE == 0:
        # cached on the caller's tensor object when it was usable as is (a converted temporary
        # would defeat the cache: pass a contiguous float32 CUDA tensor)
        same = text_features.is_cuda and tf.data_ptr() == text_features.data_ptr()
        h = _prepared_bank(text_features if same else tf)
        with torch.cuda.device(tf.device):
            _abi.check(h._lib.zs_memory_project_batched(
                h._ctx, q.data_ptr(), q.shape[0], TEMPERATURE, out.data_ptr(), _stream_ptr(tf.device)))
        return out.reshape(*lead, q.shape[1])
    h = _helper(tf.device)
    with torch.cuda.device(tf.device):
        _abi.check(h._lib.zs_memory_project(
            h._ctx, q.data_ptr(), q.shape[0], tf.data_ptr(), tf.shape[0], tf.shape[1],
            TEMPERATURE, out.data_ptr(), _stream_ptr(tf.device)))
    return out.reshape(*lead, q.shape[1])


def construct_support_memory(text_json: Sequence[str]) -> torch.Tensor:
    """All `text_embedding`s of the given pickle streams as unit rows, float32 [N, d] on the GPU.

    Reference predict_prompt.py:30-56: every path is read with `pickle.load` until EOF; a list
    object is spliced as is, a dict is kept only if its caption has 8..20 words (:43); the
    embeddings are concatenated and divided by their norms (:54-55; no epsilon — a zero row
    would be NaN there, it stays zero here)."""
    _require_cuda()
    # the reader loop of :33-47, over the direct parser of torch's per-tensor storage stream
    # (same records as pickle.load, ~4x faster: dataset.read_related_records)
    from .dataset.dataset import read_related_records
    all_data: List[dict] = read_related_records(list(text_json), caption_words=(8, 20))
    rows = [item["text_embedding"].detach().cpu().reshape(1, -1) for item in all_data]
    host = torch.cat(rows, dim=0).to(torch.float32).contiguous().pin_memory()
    dev = host.to("cuda", non_blocking=True)
    return _helper(dev.device).normalize_rows(dev)
