"""GPU (-m gpu): rank-of-ground-truth epilogue (zs_rank_count) and the a2t / t2a mirrors against
the reference's own a2t / t2a outputs (golden) and the oracle."""
import numpy as np
import pytest
import torch

import helpers
from helpers import recipes
from oracle import oracle

pytestmark = pytest.mark.gpu


def _exact_positions(queries, bank, targets):
    """float64 count of strictly greater scores, and the count of near-ties (|ds| < 1e-3)."""
    s = torch.from_numpy(oracle.exact_scores(queries, bank))
    ts = s.gather(1, targets)
    greater = (s.unsqueeze(1) > ts.unsqueeze(2)).sum(dim=2)
    near = ((s.unsqueeze(1) - ts.unsqueeze(2)).abs() < 1e-3).sum(dim=2) - 1     # minus the target itself
    return greater, near, ts


@pytest.mark.parametrize("Q,N,T", [(300, 5000, 5), (1, 300, 1), (1045, 19195, 5), (257, 70001, 3), (64, 255, 8)])
def test_rank_count_matches_exact_scores(Q, N, T, monkeypatch):
    import zsaac_b200
    q = helpers.seeded((Q, 1024), Q + N + T)
    b = helpers.seeded((N, 1024), 7 * N + T)
    gen = torch.Generator().manual_seed(Q * N)
    targets = torch.randint(0, N, (Q, T), generator=gen)
    # make some targets genuinely good matches so that small ranks are exercised too
    q[: Q // 2] = b[targets[: Q // 2, 0]] + 1.5 * q[: Q // 2]
    if T > 1:
        targets[::7, T - 1] = -1                         # unused slots
    for cg in ("1", "2"):
        monkeypatch.setenv("ZSAAC_CTA_GROUP", cg)
        rb = zsaac_b200.RelatedBank.from_tensor(b.cuda())
        ranks, scores = rb.rank_of(q.cuda(), targets.cuda())
        torch.cuda.synchronize()
        rb.close()
        ranks, scores = ranks.cpu(), scores.cpu()
        used = targets >= 0
        greater, near, ts = _exact_positions(q, b, targets.clamp(min=0))
        assert (ranks[~used] == -1).all()
        assert (scores[used] - ts[used].float()).abs().max().item() < 1e-3
        # the count can differ from the float64 count only by candidates within the 1e-3 band
        assert ((ranks - greater).abs()[used] <= near[used]).all()
        assert (ranks[used] >= 0).all() and (ranks[used] < N).all()
        # planted matches are retrieved first or nearly so
        assert (ranks[: Q // 2, 0] <= near[: Q // 2, 0]).all()


def test_rank_count_consistent_with_topk():
    """Items returned by search() at position p have rank p (up to exact ties)."""
    import zsaac_b200
    q = helpers.seeded((200, 1024), 11).cuda()
    b = helpers.seeded((30000, 1024), 12).cuda()
    rb = zsaac_b200.RelatedBank.from_tensor(b)
    s, i = rb.search(q, 8)
    ranks, ts = rb.rank_of(q, i)
    torch.cuda.synchronize()
    assert torch.equal(ranks, torch.arange(8, device="cuda").expand(200, 8))
    assert (ts - s).abs().max().item() < 1e-5
    rb.close()


@pytest.mark.parametrize("name", list(recipes.RETRIEVAL_CASES))
def test_a2t_t2a_match_reference_golden(name):
    import zsaac_b200  # noqa: F401
    from zsaac_b200 import retrieval_metrics
    audio, caps = recipes.make_retrieval_inputs(recipes.RETRIEVAL_CASES[name])
    g = helpers.golden(name)
    for fn, key in ((retrieval_metrics.a2t, "a2t"), (retrieval_metrics.t2a, "t2a")):
        plain = fn(audio, caps)
        full = fn(audio, caps, return_ranks=True)
        assert len(plain) == 7 and len(full) == 9 and tuple(plain) == tuple(full[:7])
        ranks, top1 = full[7], full[8]
        ref_ranks, ref_top1 = g[key + "_ranks"], g[key + "_top1"]
        assert ranks.shape == ref_ranks.shape and top1.shape == ref_top1.shape
        # bf16 operands, and np.argsort's arbitrary order among exact ties (the easy fixture has
        # duplicate audios): a rank may differ from the reference's by at most the number of
        # candidates within 1e-3 of the ground truth's score
        a_t, c_t = torch.from_numpy(audio), torch.from_numpy(caps)
        n_audio = audio.shape[0] // 5
        if key == "a2t":
            tg = torch.arange(n_audio).unsqueeze(1) * 5 + torch.arange(5).unsqueeze(0)
            _, near, _ = _exact_positions(a_t[0:5 * n_audio:5], c_t, tg)
            allowed = near.max(dim=1).values.numpy()
        else:
            tg = (torch.arange(5 * n_audio) // 5).unsqueeze(1)
            _, near, _ = _exact_positions(c_t[:5 * n_audio], a_t[0::5], tg)
            allowed = near[:, 0].numpy()
        diff = np.abs(ranks - ref_ranks)
        assert (diff <= allowed).all(), (diff.max(), allowed.max())
        if name == "retrieval_metrics":           # no duplicates there: nearly everything is exact
            assert diff.max() <= 2 and (diff > 0).mean() < 0.05, (diff.max(), (diff > 0).mean())
        if name == "retrieval_metrics":
            assert (top1 == ref_top1).mean() > 0.97
            np.testing.assert_allclose(np.array(plain[:4]), g[key + "_metrics"][:4], atol=2.0)   # R@k in %
            assert plain[4] == g[key + "_metrics"][4]                                            # medR
            np.testing.assert_allclose(plain[5], g[key + "_metrics"][5], atol=0.05)               # meanR
            np.testing.assert_allclose(plain[6], g[key + "_metrics"][6], atol=2.0)                # mAP10
        else:
            # ties go to the ground truth here, to an arbitrary item in the reference: the
            # metrics can only be equal or better
            assert plain[0] >= g[key + "_metrics"][0] - 1e-9 and plain[5] <= g[key + "_metrics"][5] + 1e-9


# ------------------------------------------------------------------------------ map2memory (row a8)
@pytest.mark.parametrize("name", list(recipes.MEMORY_CASES))
def test_map2memory_matches_reference_golden(name, tmp_path):
    import pickle
    import zsaac_b200  # noqa: F401
    from zsaac_b200.predict_prompt import construct_support_memory, map2memory
    q, bank = recipes.make_memory_inputs(recipes.MEMORY_CASES[name])
    g = helpers.golden(name)
    for dev in ("cpu", "cuda"):
        out = map2memory(torch.from_numpy(q).to(dev), torch.from_numpy(bank).to(dev))
        assert out.is_cuda and out.dtype == torch.float32 and tuple(out.shape) == q.shape
        # fp32 everywhere; the fast exp and a different summation order leave ~1e-6
        assert (out.cpu() - torch.from_numpy(g["out"])).abs().max().item() < 2e-5
        assert (out.norm(dim=-1) - 1).abs().max().item() < 1e-5
    p = tmp_path / "mem.pkl"
    caps = ["too short", "a caption that has exactly eight words in it", " ".join(["w"] * 25),
            "another caption with nine words in it right here now"]
    with open(p, "wb") as f:
        for i, c in enumerate(caps):
            pickle.dump({"caption": c, "text_embedding": torch.from_numpy(bank[i:i + 1] * (i + 2.0))}, f)
        pickle.dump([{"caption": "listed", "text_embedding": torch.from_numpy(bank[9:10] * 3.0)}], f)
    mem = construct_support_memory([str(p)])
    assert mem.is_cuda and (mem.cpu() - torch.from_numpy(g["memory"])).abs().max().item() < 1e-6


@pytest.mark.parametrize("Q,N,d", [(1, 400_000, 1024), (3, 50_001, 1024), (9, 1000, 512), (2, 7, 64),
                                   # batches: both contractions on the tensor cores (split-bf16 operands)
                                   (8, 50_001, 1024), (64, 400_000, 1024), (130, 30_000, 1024),
                                   (300, 5_000, 256), (1045, 19_195, 1024), (16, 63, 64)])
def test_map2memory_against_oracle_at_scale(Q, N, d):
    import zsaac_b200  # noqa: F401
    from zsaac_b200.predict_prompt import map2memory
    gen = torch.Generator(device="cuda").manual_seed(Q + N)
    bank = torch.nn.functional.normalize(torch.randn(N, d, device="cuda", generator=gen), dim=-1)
    q = torch.nn.functional.normalize(torch.randn(Q, d, device="cuda", generator=gen), dim=-1)
    q[0] = torch.nn.functional.normalize(bank[N // 2] + 0.02 * q[0], dim=-1)     # one peaked query
    if Q > 4:
        q[1] = torch.nn.functional.normalize(bank[3] + bank[N // 3] + 0.01 * q[1], dim=-1)   # two near-equal peaks
    out = map2memory(q, bank)
    torch.cuda.synchronize()
    want = oracle.map2memory(q.cpu(), bank.cpu())
    assert (out.cpu() - want).abs().max().item() < 5e-5
    assert (out.norm(dim=-1) - 1).abs().max().item() < 1e-5
    out2 = map2memory(q, bank)                       # second call: cached operands, same bits
    torch.cuda.synchronize()
    assert torch.equal(out, out2)
