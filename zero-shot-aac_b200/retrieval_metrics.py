"""Retrieval metrics of the CLAP validation loop, without the per-query argsort.

Mirror of `a2t` / `t2a` in the reference's retrieval/tools/utils.py:169-251 (same arguments,
same returned tuple).  The reference loops over audios on the CPU: `util.cos_sim` against all
embeddings, a full `np.argsort`, then `np.where(inds == i)` to find where the ground truth
landed — O(Q N log N).  Here the rank of each ground-truth item is *counted* in the epilogue of
the same fused similarity kernel that serves the related-caption search (zs_rank_count), and the
R@k / medR / meanR / mAP10 arithmetic that follows is the reference's, on the resulting ranks.

Conventions inherited from the reference: embeddings come in groups of 5 captions per audio
(audio_embs repeats each audio 5 times, cap_embs holds the 5 captions).  Exact score ties are
resolved in favour of the ground truth (np.argsort leaves them unspecified).
"""
from __future__ import annotations

import numpy as np
import torch

from .retrieval import RelatedBank, _require_cuda

CAPTIONS_PER_AUDIO = 5


def _to_cuda(x) -> torch.Tensor:
    t = torch.as_tensor(np.asarray(x) if not isinstance(x, torch.Tensor) else x)
    return t.detach().to(device="cuda", dtype=torch.float32).contiguous()


def _summary(ranks: np.ndarray, ap10_sum: float):
    r1 = 100.0 * len(np.where(ranks < 1)[0]) / len(ranks)
    r5 = 100.0 * len(np.where(ranks < 5)[0]) / len(ranks)
    r10 = 100.0 * len(np.where(ranks < 10)[0]) / len(ranks)
    r50 = 100.0 * len(np.where(ranks < 50)[0]) / len(ranks)
    mAP10 = 100.0 * ap10_sum / len(ranks)
    medr = np.floor(np.median(ranks)) + 1
    meanr = ranks.mean() + 1
    return r1, r5, r10, r50, medr, meanr, mAP10


def a2t(audio_embs, cap_embs, return_ranks=False):
    """Audio-to-caption retrieval (reference retrieval/tools/utils.py:169-213).

    Query i is audio_embs[5*i]; its ground truth is captions 5*i .. 5*i+4.  rank = best (lowest)
    position of the five; AP@10 over the five as in the reference."""
    _require_cuda()
    audio = _to_cuda(audio_embs)
    caps = _to_cuda(cap_embs)
    num_audios = int(audio.shape[0] / CAPTIONS_PER_AUDIO)
    queries = audio[0:CAPTIONS_PER_AUDIO * num_audios:CAPTIONS_PER_AUDIO].contiguous()
    bank = RelatedBank.from_tensor(caps, normalize=True)          # util.cos_sim normalises both sides
    targets = (torch.arange(num_audios, device=caps.device).unsqueeze(1) * CAPTIONS_PER_AUDIO
               + torch.arange(CAPTIONS_PER_AUDIO, device=caps.device).unsqueeze(0))
    pos, _ = bank.rank_of(queries, targets)                      # [num_audios, 5] positions in the ordering
    top1 = None
    if return_ranks:
        _, top1_idx = bank.search(queries, 1)
        top1 = top1_idx[:, 0].cpu().numpy().astype(np.float64)
    pos = pos.cpu().numpy()
    bank.close()
    # within the five ground truths of one audio the positions must be distinct: a caption that
    # scores lower than a sibling also has that sibling ahead of it, which the count includes
    ranks = pos.min(axis=1).astype(np.float64)
    AP10 = np.zeros(num_audios)
    for index in range(num_audios):
        inds_map = np.sort(pos[index][pos[index] < 10] + 1)      # reference :190-193
        if len(inds_map) != 0:
            AP10[index] = np.sum(np.arange(1, len(inds_map) + 1) / inds_map) / CAPTIONS_PER_AUDIO
    out = _summary(ranks, float(np.sum(AP10)))
    if return_ranks:
        return (*out, ranks, top1)
    return out


def t2a(audio_embs, cap_embs, return_ranks=False):
    """Caption-to-audio retrieval (reference retrieval/tools/utils.py:216-251).

    Every caption is a query against the num_audios distinct audios (every 5th embedding); its
    ground truth is audio index // 5."""
    _require_cuda()
    audio = _to_cuda(audio_embs)
    caps = _to_cuda(cap_embs)
    num_audios = int(audio.shape[0] / CAPTIONS_PER_AUDIO)
    audios = audio[0:audio.shape[0]:CAPTIONS_PER_AUDIO].contiguous()
    queries = caps[:CAPTIONS_PER_AUDIO * num_audios].contiguous()
    bank = RelatedBank.from_tensor(audios, normalize=True)
    targets = (torch.arange(CAPTIONS_PER_AUDIO * num_audios, device=caps.device) // CAPTIONS_PER_AUDIO).unsqueeze(1)
    pos, _ = bank.rank_of(queries, targets)
    top1 = None
    if return_ranks:
        _, top1_idx = bank.search(queries, 1)
        top1 = top1_idx[:, 0].cpu().numpy().astype(np.float64)
    ranks = pos[:, 0].cpu().numpy().astype(np.float64)
    bank.close()
    ap10_sum = float(np.sum(1 / (ranks[np.where(ranks < 10)[0]] + 1)))   # reference :246
    out = _summary(ranks, ap10_sum)
    if return_ranks:
        return (*out, ranks, top1)
    return out
