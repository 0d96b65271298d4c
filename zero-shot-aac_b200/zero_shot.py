"""The similarity step of the reference's zero-shot classification script.

retrieval/zero_shot_classification.py is a module-level script; inside its loop over the test
clips it computes (:97-98, :103)

    score = audio_emb @ text_embeds.t()                 # [1, 1024] x [1024, n_classes]
    pred = torch.argmax(F.softmax(score, dim=-1), dim=-1)

(softmax is monotone, so the prediction is the arg-max of the raw similarity).  `predict` is that
step for one clip or a batch, in fp32 on the GPU through zs_exact_topk_f32 — the class-prompt
bank has a few dozen rows, one launch per call.
"""
from __future__ import annotations

import torch

from .retrieval import _require_cuda, exact_topk


def similarity_top1(audio_emb: torch.Tensor, text_embeds: torch.Tensor):
    """(best score [...], predicted class index [...]) for audio_emb [..., d] vs text_embeds [C, d]."""
    _require_cuda()
    lead = tuple(audio_emb.shape[:-1])
    q = audio_emb.detach().reshape(-1, audio_emb.shape[-1])
    score, index = exact_topk(q, text_embeds, 1, normalize=False)
    return score.reshape(lead), index.reshape(lead)


def predict(audio_emb: torch.Tensor, text_embeds: torch.Tensor) -> torch.Tensor:
    """pred of reference :103 — int64 tensor shaped like audio_emb.shape[:-1], on the GPU."""
    return similarity_top1(audio_emb, text_embeds)[1]
