"""Mirror of the retrieval helper in the reference's utils.py.

Only `sound_effect_choice` (utils.py:131-137) is on the hot path; the rest of the reference's
utils.py (logging, COCO scoring, prompt text assembly) is out of scope and stays reference code.
"""
from __future__ import annotations

import torch

from .retrieval import _require_cuda, bank_for


def sound_effect_choice(prefix, sound_effect_embeddings, choice_num):
    """Indices of the `choice_num` label embeddings most similar to each prefix row.

    Reference (utils.py:131-137): similarity = prefix @ bank.T; softmax over the bank on the CPU;
    topk -> index.  softmax is strictly monotone per row, so the indices are the top-k of the raw
    similarity; no normalisation is applied (inputs are unit-norm CLAP embeddings,
    retrieval/models/ase_model.py:54,59).  Returns an int64 CPU tensor of shape
    prefix.shape[:-1] + (choice_num,), like the reference.

    The similarity + top-k run on the GPU through libzsaac_b200 for CPU and CUDA inputs alike
    (CPU tensors are copied to the current device; the label bank is converted once and cached).
    There is no CPU fallback: inside a forked DataLoader worker, where CUDA cannot be initialised,
    this raises — call it from the main process / collate step instead.
    """
    _require_cuda()
    rb = bank_for(sound_effect_embeddings, normalize=False)
    lead = tuple(prefix.shape[:-1])
    q = prefix.detach().reshape(-1, prefix.shape[-1]).to(rb.device)
    if q.dtype not in (torch.float32, torch.bfloat16):
        q = q.float()
    _, index = rb.search(q, int(choice_num), normalize_queries=False)
    return index.reshape(*lead, int(choice_num)).cpu()
