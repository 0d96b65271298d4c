"""Retrieval metrics of the CLAP validation loop, without the per-query argsort.

Mirror of `a2t` / `t2a` in the reference's retrieval/tools/utils.py:169-251 (same arguments,
same returned tuple).  The reference loops over audios on the CPU: `util.cos_sim` against all
embeddings, a full `np.argsort`, then `np.where(inds == i)` to find where the ground truth
landed — O(Q N log N).  Here the position of each ground-truth item is *counted* on the GPU and
the R@k / medR / meanR / mAP10 arithmetic that follows is the reference's, on those positions.

Precision: at the sizes the reference evaluates (Clotho: 1,045 audios x 5,225 captions;
AudioCaps: 975 x 4,875) the scores are computed in fp32 (zs_exact_rank_f32), the arithmetic of
the reference's cos_sim, so positions — and therefore every metric — equal the reference's except
where two fp32 scores are within rounding of each other.  Only evaluation sets too large for the
fp32 score scratch (Q x N > 2^28) go through the bf16 tensor-core counting epilogue
(zs_rank_count), whose near-tie band is ~1e-4.

Ties: positions follow the library's total order (score desc, index asc) — np.argsort leaves
tied elements in unspecified order — so tied items, e.g. two identical captions of one audio,
occupy distinct consecutive positions exactly as they do in any argsort.

Conventions inherited from the reference: embeddings come in groups of 5 captions per audio
(audio_embs repeats each audio 5 times, cap_embs holds the 5 captions).
"""
from __future__ import annotations

import numpy as np
import torch

from .retrieval import RelatedBank, _require_cuda, exact_fits, exact_rank, exact_topk

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


def _positions(queries: torch.Tensor, bank: torch.Tensor, targets: torch.Tensor, want_top1: bool):
    """(positions [Q, T] int64 numpy, top-1 index per query or None)."""
    if exact_fits(queries.shape[0], bank.shape[0]):
        pos, _ = exact_rank(queries, bank, targets, normalize=True)   # util.cos_sim normalises both sides
        top1 = exact_topk(queries, bank, 1, normalize=True)[1][:, 0] if want_top1 else None
    else:
        rb = RelatedBank.from_tensor(bank, normalize=True)
        pos, _ = rb.rank_of(queries, targets)
        top1 = rb.search(queries, 1)[1][:, 0] if want_top1 else None
        pos = _distinct_positions(pos)
        torch.cuda.current_stream(bank.device).synchronize()
        rb.close()
    pos = pos.cpu().numpy()
    return pos, (None if top1 is None else top1.cpu().numpy().astype(np.float64))


def _distinct_positions(pos: torch.Tensor) -> torch.Tensor:
    """bf16 counting epilogue only: it counts rows scoring STRICTLY higher, so ground truths of one
    query that tie exactly share a count; an ordering gives them consecutive positions.  Sort the
    T positions of every query and push each tied one behind its predecessor."""
    if pos.shape[1] == 1:
        return pos
    srt, order = torch.sort(pos, dim=1)
    for j in range(1, srt.shape[1]):
        bump = (srt[:, j] <= srt[:, j - 1]) & (srt[:, j] >= 0) & (srt[:, j - 1] >= 0)
        srt[:, j] = torch.where(bump, srt[:, j - 1] + 1, srt[:, j])
    out = torch.empty_like(pos)
    out.scatter_(1, order, srt)
    return out


def a2t(audio_embs, cap_embs, return_ranks=False):
    """Audio-to-caption retrieval (reference retrieval/tools/utils.py:169-213).

    Query i is audio_embs[5*i]; its ground truth is captions 5*i .. 5*i+4.  rank = best (lowest)
    position of the five; AP@10 over the five as in the reference."""
    _require_cuda()
    audio = _to_cuda(audio_embs)
    caps = _to_cuda(cap_embs)
    num_audios = int(audio.shape[0] / CAPTIONS_PER_AUDIO)
    queries = audio[0:CAPTIONS_PER_AUDIO * num_audios:CAPTIONS_PER_AUDIO].contiguous()
    targets = (torch.arange(num_audios, device=caps.device).unsqueeze(1) * CAPTIONS_PER_AUDIO
               + torch.arange(CAPTIONS_PER_AUDIO, device=caps.device).unsqueeze(0))
    pos, top1 = _positions(queries, caps, targets, return_ranks)   # [num_audios, 5] positions in the ordering
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
    targets = (torch.arange(CAPTIONS_PER_AUDIO * num_audios, device=caps.device) // CAPTIONS_PER_AUDIO).unsqueeze(1)
    pos, top1 = _positions(queries, audios, targets, return_ranks)
    ranks = pos[:, 0].astype(np.float64)
    ap10_sum = float(np.sum(1 / (ranks[np.where(ranks < 10)[0]] + 1)))   # reference :246
    out = _summary(ranks, ap10_sum)
    if return_ranks:
        return (*out, ranks, top1)
    return out
