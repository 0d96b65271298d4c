"""CPU oracle for the related-caption retrieval path — TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (zero-shot-aac_b200/) never does and has no CPU path.

It restates, on the CPU, what XinMing0411/zero-shot-AAC computes for this path.  The arithmetic
of the reference lives in PyTorch (pinned there to pytorch=1.11.0, retrieval/work.yaml:62), so
the restatement uses torch CPU fp32 ops for the literal forms and numpy float64 for the exact
scorer used to judge near-ties.  Each function cites the reference lines it follows.

Parity pin: the reference ships no tests, golden vectors or fixtures for this path (SURVEY.md
§4, §8c).  The oracle is pinned instead against outputs of the reference's own functions run in
the build container (tests/golden/make_golden.py imports them from /root/reference with a
'cuda'->'cpu' shim and commits the results as tests/golden/*.npz); tests/test_oracle.py checks
every function here against those files.
"""
from __future__ import annotations

import math
from typing import Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-12  # F.normalize default eps


# ------------------------------------------------------------------------------------------------
# bank construction — data_handing/embeddings_related_generator.py:9-17, _wavcaps.py:9-18
def build_bank(all_data: Sequence[dict]) -> torch.Tensor:
    """F.normalize(torch.cat([item['text_embedding'] ...]), dim=-1) in INPUT order.

    The reference passes the list through set() (:15), which neither removes value-duplicates
    (tensors hash by identity) nor keeps the order; row order only matters for tie-breaking, so
    the oracle keeps input order.
    """
    rows = [d["text_embedding"].detach().cpu().float().reshape(1, -1) for d in all_data]
    return F.normalize(torch.cat(rows, dim=0), dim=-1)


def normalize_rows(x: torch.Tensor) -> torch.Tensor:
    """F.normalize(x, dim=-1): x / max(||x||_2, 1e-12) — generator.py:17,21."""
    return F.normalize(x.detach().cpu().float(), dim=-1)


# ------------------------------------------------------------------------------------------------
# the hot loop, literally — data_handing/embeddings_related_generator.py:19-28
def process_data_literal(valid_text_embs: torch.Tensor, all_data: List[dict], topnumber: int,
                         device: str = "cpu") -> Iterator[dict]:
    """One query per iteration exactly as the reference writes it.  device='cuda' reproduces the
    reference's own device placement (bank on the GPU, per-item H2D / D2H); the default 'cpu' is
    what the build container can run."""
    valid_text_embs = valid_text_embs.to(device)
    for item in all_data:
        text_embs = F.normalize(item["text_embedding"].cpu().float(), dim=-1).to(device)   # :21
        ids = torch.cosine_similarity(text_embs, valid_text_embs).topk(topnumber)[1]    # :22
        related_embs = valid_text_embs[ids.cpu()].cpu()                                 # :23
        item["text_embedding"] = item["text_embedding"].cpu()                           # :25
        item["related_embeddings"] = related_embs                                       # :26
        yield item


# ------------------------------------------------------------------------------------------------
# the same math, batched — what the CUDA path is compared with
def stable_topk(scores: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Top-k along dim 1 under the total order (score desc, index asc)."""
    order = torch.sort(scores, dim=1, descending=True, stable=True)  # stable: ties keep low index first
    return order.values[:, :k].contiguous(), order.indices[:, :k].contiguous()


def cosine_topk(queries: torch.Tensor, bank: torch.Tensor, k: int, *, normalize: bool = True,
                self_index: Optional[torch.Tensor] = None, chunk: int = 1024
                ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Batched fp32 restatement of generator.py:21-22: normalise both sides, q @ bank.T, top-k.

    self_index [Q] (global row to skip, <0 = none) is the opt-in self-exclusion extension.
    Returns (scores [Q,k] fp32 desc, indices [Q,k] int64), ties by ascending index.
    """
    q = queries.detach().cpu().float()
    b = bank.detach().cpu().float()
    if normalize:
        q = F.normalize(q, dim=-1)
        b = F.normalize(b, dim=-1)
    out_s, out_i = [], []
    for lo in range(0, q.shape[0], chunk):
        s = q[lo:lo + chunk] @ b.T
        if self_index is not None:
            si = self_index[lo:lo + chunk].cpu().long()
            rows = torch.arange(s.shape[0])[si >= 0]
            s[rows, si[si >= 0]] = -math.inf
        vs, vi = stable_topk(s, k)
        out_s.append(vs)
        out_i.append(vi)
    return torch.cat(out_s), torch.cat(out_i)


def fast_topk(queries: torch.Tensor, bank_normalized: torch.Tensor, k: int, *, chunk: int = 1024
              ) -> Tuple[torch.Tensor, torch.Tensor]:
    """The CPU baseline that is timed: F.normalize(q) @ bank.T -> torch.topk, all host threads.
    (torch.topk leaves tie order unspecified; use cosine_topk when checking.)"""
    q = F.normalize(queries.float(), dim=-1)
    vs, vi = [], []
    for lo in range(0, q.shape[0], chunk):
        s, i = (q[lo:lo + chunk] @ bank_normalized.T).topk(k, dim=1)
        vs.append(s)
        vi.append(i)
    return torch.cat(vs), torch.cat(vi)


def exact_scores(queries: torch.Tensor, bank: torch.Tensor, *, normalize: bool = True) -> np.ndarray:
    """float64 cosine similarities [Q, N] (small shapes): the arbiter for near-ties."""
    q = queries.detach().cpu().double().numpy()
    b = bank.detach().cpu().double().numpy()
    if normalize:
        q = q / np.maximum(np.linalg.norm(q, axis=1, keepdims=True), EPS)
        b = b / np.maximum(np.linalg.norm(b, axis=1, keepdims=True), EPS)
    return q @ b.T


# ------------------------------------------------------------------------------------------------
# utils.py:131-137
def sound_effect_choice(prefix: torch.Tensor, sound_effect_embeddings: torch.Tensor, choice_num: int
                        ) -> torch.Tensor:
    similarity = prefix.float().cpu() @ sound_effect_embeddings.float().cpu().t()      # :133
    similarity_softmax = F.softmax(similarity.detach().cpu(), dim=-1)                   # :134
    _, index = torch.topk(similarity_softmax, choice_num, dim=-1)                       # :135
    return index


# models/caption_model.py:15-21 (copies at :277-283, :414-420): the method form, called by clap_to_gpt
def sound_effect_choice_method(prefix: torch.Tensor, sound_effect_embeddings: torch.Tensor, choice_num: int
                               ) -> torch.Tensor:
    index = sound_effect_choice(prefix, sound_effect_embeddings, choice_num)            # :17-19
    return sound_effect_embeddings[index].squeeze(1)                                    # :21


# utils.py:138-208 — prompt text assembly around the label retrieval (mask_probability = 0 branch:
# every selected label is kept) and the padding of the token lists, as dataset/dataset.py's
# __getitem__ (:365-368) and collate (:632-647) use them
def parse_entities(tokenizer, detected_entities: Sequence[str], mask_probability=0) -> torch.Tensor:
    if mask_probability != 0:
        raise NotImplementedError("the oracle restates the deterministic branch only")
    if len(detected_entities) == 0:
        prompt = "There are something in this audio."                                   # :162-163
    else:
        prompt = "There are" + ",".join(" " + e for e in detected_entities) + " in this audio."   # :165-169
    return torch.tensor(tokenizer.encode(prompt))                                       # :171


def padding_captions(hard_prompts: Sequence[torch.Tensor], hard_prompts_length: Sequence[int]
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
    max_length = max(hard_prompts_length)                                               # :190
    out = []
    for h in hard_prompts:
        pad = max_length - h.shape[0]
        out.append(torch.cat((h, torch.zeros(pad, dtype=torch.int64) - 1)) if pad >= 0 else h[:max_length])
    out = torch.stack(out)                                                              # :201
    mask = out.ge(0)
    out[~mask] = 0
    return out, mask.float()                                                            # :202-208


# utils.py:19-31 (Gaussian branch) — used to make BASELINE config-5 style queries
def noise_injection(x: torch.Tensor, variance: float = 0.001, generator: Optional[torch.Generator] = None
                    ) -> torch.Tensor:
    if variance == 0.0:
        return x
    std = math.sqrt(variance)
    x = F.normalize(x, dim=-1)                                                          # :26
    x = x + torch.randn(x.shape, generator=generator) * std                             # :30
    return F.normalize(x, dim=-1)                                                       # :32


# retrieval/zero_shot_classification.py:97-98,103 — top-1 special case
def zero_shot_predict(audio_emb: torch.Tensor, text_embeds: torch.Tensor) -> torch.Tensor:
    similarity = audio_emb.float().cpu() @ text_embeds.float().cpu().t()
    return F.softmax(similarity, dim=1).argmax(dim=1)


# ------------------------------------------------------------------------------------------------
# shard merge (the new repo's only exchange step; no reference counterpart)
def merge_lists(scores: torch.Tensor, indices: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """[S, Q, k] sorted lists -> [Q, k] under (score desc, index asc)."""
    s, q, k = scores.shape
    fs = scores.permute(1, 0, 2).reshape(q, s * k).cpu()
    fi = indices.permute(1, 0, 2).reshape(q, s * k).cpu().long()
    # sort by index asc first, then stable by score desc => (score desc, index asc)
    o1 = torch.sort(fi, dim=1, stable=True)
    fs1 = fs.gather(1, o1.indices)
    o2 = torch.sort(fs1, dim=1, descending=True, stable=True)
    return o2.values[:, :k].contiguous(), o1.values.gather(1, o2.indices)[:, :k].contiguous()


def shard_bounds(n_rows: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous row ranges [lo, hi) of ceil(N/G) rows per rank."""
    per = -(-n_rows // world)
    return [(min(r * per, n_rows), min((r + 1) * per, n_rows)) for r in range(world)]


# ------------------------------------------------------------------------------------------------
# comparators (north_star: scores within 1e-3; index sets identical except at near-ties)
def check_topk(got_scores, got_indices, queries, bank, k, *, normalize=True, self_index=None,
               score_tol: float = 1e-3, tie_tol: float = 1e-3) -> dict:
    """Compare a top-k result with the fp32 oracle; returns a report dict with `ok`.

    Rules: |score - oracle score of the same index| <= score_tol; every returned
    index must have oracle score >= (oracle k-th score - tie_tol); every oracle index whose score
    is > (oracle k-th score + tie_tol) must be returned; returned scores are non-increasing;
    indices within a row are distinct and honour self-exclusion.
    """
    gs = torch.as_tensor(got_scores).detach().cpu().float()
    gi = torch.as_tensor(got_indices).detach().cpu().long()
    q = queries.detach().cpu().float()
    b = bank.detach().cpu().float()
    if normalize:
        q = F.normalize(q, dim=-1)
        b = F.normalize(b, dim=-1)
    full = q @ b.T
    if self_index is not None:
        si = self_index.cpu().long()
        rows = torch.arange(full.shape[0])[si >= 0]
        full[rows, si[si >= 0]] = -math.inf
    os_, oi = stable_topk(full, k)
    rep = {}
    ref_at_got = full.gather(1, gi)
    rep["max_score_err"] = float((gs - ref_at_got).abs().max())
    rep["exact_index_match"] = float((gi == oi).float().mean())
    kth = os_[:, -1:]
    rep["all_good_enough"] = bool((ref_at_got >= kth - tie_tol).all())
    must = os_ > kth + tie_tol                      # clearly-inside oracle entries
    present = (oi.unsqueeze(2) == gi.unsqueeze(1)).any(dim=2)
    rep["all_clear_winners_present"] = bool((present | ~must).all())
    rep["sorted"] = bool((gs[:, 1:] <= gs[:, :-1]).all())
    srt = torch.sort(gi, dim=1).values
    rep["distinct"] = bool((srt[:, 1:] != srt[:, :-1]).all()) if k > 1 else True
    rep["in_range"] = bool(((gi >= 0) & (gi < b.shape[0])).all())
    rep["ok"] = (rep["max_score_err"] <= score_tol and rep["all_good_enough"]
                 and rep["all_clear_winners_present"] and rep["sorted"] and rep["distinct"]
                 and rep["in_range"])
    return rep


# ------------------------------------------------------------------------------------------------
# retrieval metrics — retrieval/tools/utils.py:169-251.  `util.cos_sim` comes from
# sentence-transformers==2.2.2 (retrieval/work.yaml:159), absent from /root/reference; its
# published definition is: normalise both operands along dim 1 (F.normalize, p=2), then mm.
def cos_sim(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    a = torch.as_tensor(a, dtype=torch.float32)
    b = torch.as_tensor(b, dtype=torch.float32)
    if a.dim() == 1:
        a = a.unsqueeze(0)
    if b.dim() == 1:
        b = b.unsqueeze(0)
    return torch.mm(F.normalize(a, p=2, dim=1), F.normalize(b, p=2, dim=1).transpose(0, 1))


def _metric_summary(ranks: np.ndarray, ap10_sum: float):
    r1 = 100.0 * len(np.where(ranks < 1)[0]) / len(ranks)
    r5 = 100.0 * len(np.where(ranks < 5)[0]) / len(ranks)
    r10 = 100.0 * len(np.where(ranks < 10)[0]) / len(ranks)
    r50 = 100.0 * len(np.where(ranks < 50)[0]) / len(ranks)
    return (r1, r5, r10, r50, np.floor(np.median(ranks)) + 1, ranks.mean() + 1,
            100.0 * ap10_sum / len(ranks))


def a2t(audio_embs, cap_embs):
    """Literal restatement of utils.py:169-213; returns (metrics tuple, ranks, top1, positions)."""
    num_audios = int(audio_embs.shape[0] / 5)
    ranks, top1, AP10 = np.zeros(num_audios), np.zeros(num_audios), np.zeros(num_audios)
    positions = np.zeros((num_audios, 5), np.int64)
    for index in range(num_audios):
        d = cos_sim(torch.Tensor(audio_embs[5 * index]), torch.Tensor(cap_embs)).squeeze(0).numpy()  # :182
        inds = np.argsort(d)[::-1]                                                                    # :183
        inds_map, rank = [], 1e20
        for slot, i in enumerate(range(5 * index, 5 * index + 5)):
            tmp = np.where(inds == i)[0][0]                                                           # :189
            positions[index, slot] = tmp
            rank = min(rank, tmp)
            if tmp < 10:
                inds_map.append(tmp + 1)
        inds_map = np.sort(np.array(inds_map))
        AP10[index] = np.sum(np.arange(1, len(inds_map) + 1) / inds_map) / 5 if len(inds_map) else 0.0
        ranks[index], top1[index] = rank, inds[0]
    return _metric_summary(ranks, float(np.sum(AP10))), ranks, top1, positions


def t2a(audio_embs, cap_embs):
    """Literal restatement of utils.py:216-251; returns (metrics tuple, ranks, top1)."""
    num_audios = int(audio_embs.shape[0] / 5)
    audios = np.array([audio_embs[i] for i in range(0, audio_embs.shape[0], 5)])
    ranks, top1 = np.zeros(5 * num_audios), np.zeros(5 * num_audios)
    for index in range(num_audios):
        d = cos_sim(torch.Tensor(cap_embs[5 * index: 5 * index + 5]), torch.Tensor(audios)).numpy()   # :232
        for i in range(d.shape[0]):
            inds = np.argsort(d[i])[::-1]
            ranks[5 * index + i] = np.where(inds == index)[0][0]                                      # :237
            top1[5 * index + i] = inds[0]
    ap10_sum = float(np.sum(1 / (ranks[np.where(ranks < 10)[0]] + 1)))
    return _metric_summary(ranks, ap10_sum), ranks, top1


# ------------------------------------------------------------------------------------------------
# predict_prompt.py:23-29 (defined in the reference, its call at :134 is commented out)
def map2memory(audio_embed: torch.Tensor, text_features: torch.Tensor) -> torch.Tensor:
    audio_embed = audio_embed.detach().cpu().float()
    text_features = text_features.detach().cpu()
    sim = audio_embed @ text_features.T.float()                                          # :25
    sim = (sim * 100).softmax(dim=-1)                                                    # :26
    prefix_embedding = sim @ text_features.float()                                       # :27
    prefix_embedding = prefix_embedding / prefix_embedding.norm(dim=-1, keepdim=True)    # :28
    return prefix_embedding


# predict_prompt.py:30-56
def construct_support_memory(paths) -> torch.Tensor:
    import pickle
    all_data = []
    for dp in paths:
        with open(dp, "rb") as f:
            while True:
                try:
                    item = pickle.load(f)
                    if type(item) is list:
                        all_data = all_data + item
                    elif 8 <= len(item["caption"].split()) <= 20:
                        all_data.append(item)
                except EOFError:
                    break
    text_features = torch.cat([item["text_embedding"] for item in all_data], dim=0)
    return text_features / text_features.norm(dim=-1, keepdim=True).float()
