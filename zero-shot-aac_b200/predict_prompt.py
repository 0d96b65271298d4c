"""Mirror of the memory-projection helpers of the reference's predict_prompt.py.

`map2memory` (predict_prompt.py:23-29) and `construct_support_memory` (:30-56) are defined in the
reference but the call is commented out (:134); they are the training-free alternative to the
related-caption prefix: project the audio embedding onto the span of the caption memory with
softmax(100 * similarity) weights.  Same contraction as the retrieval path with a softmax-weighted
reduction instead of a top-k, so it lives on the same library (zs_memory_project).
"""
from __future__ import annotations

import pickle
from typing import List, Sequence

import torch

from . import _abi
from .retrieval import RelatedBank, _require_cuda, _stream_ptr, helper_context

TEMPERATURE = 100.0   # predict_prompt.py:26  `(sim*100).softmax(dim=-1)`

def _helper(device: torch.device) -> RelatedBank:
    """A context for the bank-less entry points (normalise, memory projection) on `device`."""
    return helper_context(device)


def map2memory(audio_embed: torch.Tensor, text_features: torch.Tensor) -> torch.Tensor:
    """prefix_embedding = normalise(softmax(100 * audio_embed @ text_features.T) @ text_features).

    audio_embed [Q, d] (the reference passes one embedding, [1, 1024]); text_features [N, d]
    float32 unit rows (construct_support_memory).  Returns float32 [Q, d] on the GPU.  CPU inputs
    are copied to the current CUDA device; there is no CPU path."""
    _require_cuda()
    tf = text_features.detach()
    if not tf.is_cuda:
        tf = tf.to("cuda")
    tf = tf.to(torch.float32).contiguous()
    q = audio_embed.detach().to(device=tf.device, dtype=torch.float32)
    lead = tuple(q.shape[:-1])
    q = q.reshape(-1, q.shape[-1]).contiguous()
    if q.shape[1] != tf.shape[1]:
        raise ValueError(f"audio_embed has d={q.shape[1]}, text_features d={tf.shape[1]}")
    out = torch.empty_like(q)
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
    all_data: List[dict] = list()
    for dp in text_json:
        with open(dp, "rb") as f:
            while True:
                try:
                    item = pickle.load(f)
                    if type(item) is list:
                        all_data = all_data + item
                    else:
                        if len(item["caption"].split()) >= 8 and len(item["caption"].split()) <= 20:
                            all_data.append(item)
                except EOFError:
                    break
    rows = [item["text_embedding"].detach().cpu().reshape(1, -1) for item in all_data]
    host = torch.cat(rows, dim=0).to(torch.float32).contiguous().pin_memory()
    dev = host.to("cuda", non_blocking=True)
    return _helper(dev.device).normalize_rows(dev)
