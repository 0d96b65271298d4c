"""Mirror of the retrieval helper in the reference's utils.py.

Only `sound_effect_choice` (utils.py:131-137) — and its method copies in the caption models
(models/caption_model.py:15-21, :277-283, :414-420: same ranking, returning the chosen label
embeddings instead of their indices) — is on the hot path; the rest of the reference's utils.py
(logging, COCO scoring, prompt text assembly) is out of scope and stays reference code.
"""
from __future__ import annotations

import torch

from .retrieval import (_require_cuda, bank_for, exact_fits, exact_topk, search_rescored)


def sound_effect_choice(prefix, sound_effect_embeddings, choice_num):
    """Indices of the `choice_num` label embeddings most similar to each prefix row.

    Reference (utils.py:131-137): similarity = prefix @ bank.T; softmax over the bank on the CPU;
    topk -> index.  softmax is strictly monotone per row, so the indices are the top-k of the raw
    similarity; no normalisation is applied (inputs are unit-norm CLAP embeddings,
    retrieval/models/ase_model.py:54,59).  Returns an int64 CPU tensor of shape
    prefix.shape[:-1] + (choice_num,), like the reference.

    The label bank is small (L = 527 AudioSet labels), so the scores are computed in fp32 like the
    reference's own matmul (zs_exact_topk_f32: one launch for up to 64 prefixes; the indices equal
    the reference's except at fp32 rounding ties).  A bank too large for that route goes through
    the tensor-core search with fp32 re-scoring.  CPU inputs are copied to the current device.
    There is no CPU fallback: inside a forked DataLoader worker, where CUDA cannot be initialised,
    this raises — batch the call in the collate step of the main process instead
    (zsaac_b200.dataset.collate_with_sound_effects).
    """
    return _choice_index(prefix, sound_effect_embeddings, choice_num).cpu()


def _choice_index(prefix, sound_effect_embeddings, choice_num) -> torch.Tensor:
    """int64 [prefix.shape[:-1] + (choice_num,)] on the GPU: the ranking both mirrors share."""
    _require_cuda()
    lead = tuple(prefix.shape[:-1])
    q = prefix.detach().reshape(-1, prefix.shape[-1])
    k = int(choice_num)
    if exact_fits(q.shape[0], sound_effect_embeddings.shape[0]):
        _, index = exact_topk(q, sound_effect_embeddings, k, normalize=False)
    else:
        rb = bank_for(sound_effect_embeddings, normalize=False)
        q = q.to(rb.device)
        if q.dtype == torch.float32 and sound_effect_embeddings.dtype == torch.float32:
            _, index = search_rescored(rb, q, sound_effect_embeddings, k, normalize=False)
        else:
            if q.dtype not in (torch.float32, torch.bfloat16):
                q = q.float()
            _, index = rb.search(q, k, normalize_queries=False)
    return index.reshape(*lead, k)


def sound_effect_embeddings_choice(prefix, sound_effect_embeddings, choice_num):
    """The `choice_num` label embeddings most similar to each prefix row, best first.

    Mirror of the `sound_effect_choice` METHOD of the reference's caption models
    (models/caption_model.py:15-21, :277-283, :414-420), which their `clap_to_gpt` calls in every
    forward pass with the batch of prefixes (:75, :127, :173, :265): same ranking as
    utils.sound_effect_choice, but it returns `sound_effect_embeddings[index].squeeze(1)` — shape
    prefix.shape[:-1] + (choice_num, d), dim 1 dropped if it is 1 — on the bank's device and in the
    bank's dtype.  The reference moves the similarity to the CPU for softmax + topk and indexes the
    bank with a CPU index (a D2H and an H2D synchronisation per forward pass); here the ranking is
    one launch (zs_exact_topk_f32) and the indices never leave the GPU.  The gather is torch's own
    indexing, so autograd sees exactly what it sees in the reference (a gradient into
    `sound_effect_embeddings` if that is a parameter, none into `prefix`)."""
    index = _choice_index(prefix, sound_effect_embeddings, choice_num)
    if index.device != sound_effect_embeddings.device:      # a CPU bank was ranked on the GPU
        index = index.to(sound_effect_embeddings.device)
    return sound_effect_embeddings[index].squeeze(1)
