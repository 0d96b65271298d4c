"""CPU: the arithmetic retrieval_metrics.a2t / t2a do AROUND the kernels (reference
retrieval/tools/utils.py:169-251: R@k, medR, meanR, mAP10 from the positions of the ground truth),
with the position counting — zs_exact_rank_f32 / zs_rank_count on the GPU — replaced by a torch
stand-in.  Checked against the reference's own outputs (tests/golden/retrieval_metrics*.npz)."""
import numpy as np
import pytest
import torch

import helpers
from helpers import recipes
from oracle import oracle


def _scores(queries, bank):
    return (torch.nn.functional.normalize(queries.float(), dim=-1)
            @ torch.nn.functional.normalize(bank.float(), dim=-1).T)


def _exact_positions(queries, bank, targets, want_top1):
    """What zs_exact_rank_f32 returns: rows ranking before the target under (score desc, index asc)."""
    s = _scores(queries, bank)
    ts = s.gather(1, targets)
    cols = torch.arange(bank.shape[0])
    before = (s.unsqueeze(1) > ts.unsqueeze(2)) | ((s.unsqueeze(1) == ts.unsqueeze(2))
                                                   & (cols.view(1, 1, -1) < targets.unsqueeze(2)))
    pos = before.sum(dim=2)
    top1 = torch.sort(s, dim=1, descending=True, stable=True).indices[:, 0] if want_top1 else None
    return pos.numpy(), (None if top1 is None else top1.numpy().astype(np.float64))


def _patch(monkeypatch, positions):
    import zsaac_b200  # noqa: F401
    from zsaac_b200 import retrieval_metrics as rm
    monkeypatch.setattr(rm, "_require_cuda", lambda: None)
    monkeypatch.setattr(rm, "_to_cuda", lambda x: torch.as_tensor(np.asarray(x)).float().contiguous())
    monkeypatch.setattr(rm, "_positions", positions)
    return rm


@pytest.mark.parametrize("name", list(recipes.RETRIEVAL_CASES))
def test_metric_arithmetic_matches_the_reference_golden(monkeypatch, name):
    rm = _patch(monkeypatch, _exact_positions)
    audio, caps = recipes.make_retrieval_inputs(recipes.RETRIEVAL_CASES[name])
    g = helpers.golden(name)
    a_t, c_t = torch.from_numpy(audio), torch.from_numpy(caps)
    n_audio = audio.shape[0] // 5
    for fn, key in ((rm.a2t, "a2t"), (rm.t2a, "t2a")):
        full = fn(audio, caps, return_ranks=True)
        assert len(full) == 9 and len(fn(audio, caps)) == 7            # the reference's two return forms
        ranks, top1 = full[7], full[8]
        ref_ranks, ref_top1, ref_metrics = g[key + "_ranks"], g[key + "_top1"], g[key + "_metrics"]
        if key == "a2t":
            tg = torch.arange(n_audio).unsqueeze(1) * 5 + torch.arange(5).unsqueeze(0)
            s = torch.from_numpy(oracle.exact_scores(a_t[0:5 * n_audio:5], c_t))
        else:
            tg = (torch.arange(5 * n_audio) // 5).unsqueeze(1)
            s = torch.from_numpy(oracle.exact_scores(c_t[:5 * n_audio], a_t[0::5]))
        ts = s.gather(1, tg)
        # ground truths tied (to fp32 rounding) with other rows: np.argsort's order is arbitrary there
        tied = (((s.unsqueeze(1) - ts.unsqueeze(2)).abs() < 2e-6).sum(dim=2) - 1).max(dim=1).values.numpy()
        assert (np.abs(ranks - ref_ranks) <= tied).all()
        if tied.max() == 0:
            assert np.array_equal(ranks, ref_ranks) and np.array_equal(top1, ref_top1)
            np.testing.assert_allclose(np.array(full[:7]), ref_metrics, rtol=0, atol=1e-9)
        else:
            np.testing.assert_allclose(np.array(full[:4]), ref_metrics[:4], atol=100.0 * (tied > 0).mean() + 1e-9)


def test_counting_path_gives_tied_siblings_consecutive_positions(monkeypatch):
    """Beyond 2^28 scores the positions come from the bf16 counting epilogue (rows scoring STRICTLY
    higher), where bit-identical sibling captions share a count; _distinct_positions must turn
    those into the consecutive positions an argsort gives (AP10 saw a position twice)."""
    import zsaac_b200  # noqa: F401
    from zsaac_b200 import retrieval_metrics as rm

    def counting_positions(queries, bank, targets, want_top1):
        s = _scores(queries, bank)
        ts = s.gather(1, targets)
        pos = (s.unsqueeze(1) > ts.unsqueeze(2)).sum(dim=2)            # what zs_rank_count returns
        return rm._distinct_positions(pos).numpy(), None

    _patch(monkeypatch, counting_positions)
    rng = np.random.RandomState(5)
    audio = np.repeat(rng.randn(40, 1024).astype(np.float32), 5, axis=0)
    caps = (audio + 0.3 * rng.randn(200, 1024)).astype(np.float32)
    caps[1] = caps[0]                       # siblings of audio 0, bit-identical
    caps[3] = caps[0]
    caps[7] = caps[6]                       # and of audio 1
    got = rm.a2t(audio, caps)
    want = oracle.a2t(audio, caps)          # the reference's argsort + np.where, restated
    np.testing.assert_allclose(np.array(got), np.array(want[0]), atol=1e-9)
    # the helper itself: ties bumped behind their predecessor, unused slots (-1) left alone
    pos = torch.tensor([[4, 4, 9, 4, 0], [2, -1, 2, 7, 7], [5, 6, 7, 8, 9]])
    out = rm._distinct_positions(pos)
    assert sorted(out[0].tolist()) == [0, 4, 5, 6, 9] and out[0, 4] == 0 and out[0, 2] == 9
    assert sorted(out[1].tolist()) == [-1, 2, 3, 7, 8] and out[1, 1] == -1
    assert out[2].tolist() == [5, 6, 7, 8, 9]
    assert torch.equal(rm._distinct_positions(torch.tensor([[3], [3]])), torch.tensor([[3], [3]]))
