"""GPU (-m gpu): round-2 additions — k > 32 (multi-pass search), the single-launch ("solo") mode
and the threshold bootstrap (both must be result-neutral), the exact fp32 small-bank path
(sound_effect_choice, zero-shot top-1, retrieval metrics), the collate hook and the bank cache."""
import numpy as np
import pytest
import torch

import helpers
from helpers import recipes
from oracle import oracle
from test_gpu_parity import BF16_TOL, SCORE_TOL, bf16_scores, run_search

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def zs():
    import zsaac_b200
    assert torch.cuda.is_available()
    zsaac_b200.load_library()
    return zsaac_b200


# ------------------------------------------------------------------------------------------ k > 32
@pytest.mark.parametrize("Q,N,k", [(260, 9000, 33), (975, 49838, 50), (64, 3000, 100), (7, 1100, 1024),
                                   (300, 20000, 64), (130, 70, 65)])
@pytest.mark.parametrize("solo", ["0", "1"])
def test_k_above_32_matches_oracle(zs, monkeypatch, solo, Q, N, k):
    """embeddings_related_generator.py:22,46: --topnumber is an unconstrained int."""
    monkeypatch.setenv("ZSAAC_SOLO", solo)
    q, b = helpers.seeded((Q, 1024), 3 * Q + N + k), helpers.seeded((N, 1024), 5 * N + k)
    if N > 200:
        b[N // 2:N // 2 + 40] = b[10:50]                      # exact ties across a pass boundary
    s, i = run_search(zs, q, b, k)
    ws, wi = oracle.stable_topk(bf16_scores(q, b), k)
    assert (s - ws).abs().max().item() < BF16_TOL
    assert (i == wi).float().mean().item() > 0.999
    assert (s[:, :-1] >= s[:, 1:]).all()
    assert all(len(set(row.tolist())) == k for row in i)           # no index twice across passes
    rep = oracle.check_topk(s, i, q, b, k, score_tol=SCORE_TOL, tie_tol=1e-3)
    assert rep["ok"], rep


def test_k_above_32_self_exclusion_and_shards(zs):
    b = helpers.seeded((6000, 1024), 5)
    q = b[:500] + 0.05 * helpers.seeded((500, 1024), 6)
    self_index = torch.arange(500)
    s, i = run_search(zs, q, b, 70, self_index=self_index.cuda())
    assert not (i == self_index[:, None]).any()
    sc = bf16_scores(q, b)
    sc[torch.arange(500), self_index] = -float("inf")
    ws, wi = oracle.stable_topk(sc, 70)
    assert (i == wi).float().mean().item() > 0.999
    # two shards merged == one bank
    parts = []
    for lo, hi in ((0, 3000), (3000, 6000)):
        rb = zs.RelatedBank.from_tensor(b[lo:hi].cuda(), index_offset=lo)
        parts.append(rb.search(q.cuda(), 70, self_index=self_index.cuda()))
        torch.cuda.synchronize()
        rb.close()
    helper = zs.RelatedBank(1, 1024)
    ms, mi = helper.merge(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
    assert torch.equal(ms.cpu(), s) and torch.equal(mi.cpu(), i)
    helper.close()


@pytest.mark.parametrize("k", [40, 100])
def test_rescore_fp32_beyond_32(zs, k):
    b = helpers.clustered(8000, 1024, 256, 0.05, 3)
    q = b[::40] + 0.01 * helpers.seeded((200, 1024), 4)
    s, i = zs.related_topk(q.cuda(), b.cuda(), k, rescore_fp32=True)
    torch.cuda.synchronize()
    rep = oracle.check_topk(s.cpu(), i.cpu(), q, b, k, score_tol=5e-6, tie_tol=5e-6)
    assert rep["ok"], rep


# ------------------------------------------------------------------- solo mode / bootstrap neutrality
NEUTRAL_SHAPES = [(1, 527, 3), (32, 400_000, 10), (128, 120_000, 10), (200, 30_000, 5), (975, 49838, 10),
                  (1045, 19195, 5), (2000, 9000, 32), (257, 70001, 17), (64, 40, 32), (4096, 5000, 10)]


@pytest.mark.parametrize("Q,N,k", NEUTRAL_SHAPES)
def test_solo_and_bootstrap_are_result_neutral(zs, monkeypatch, Q, N, k):
    """One cooperative launch (in-kernel cast + merge) and the threshold bootstrap change how the
    work is scheduled, never the result: bit-identical to the three-launch path without them."""
    q, b = helpers.seeded((Q, 1024), Q + k), helpers.seeded((N, 1024), N + k)
    b[N // 3:N // 3 + 8] = b[:8]
    results = {}
    for solo in ("0", "1"):
        for boot in ("0", "1"):
            monkeypatch.setenv("ZSAAC_SOLO", solo)
            monkeypatch.setenv("ZSAAC_BOOT", boot)
            results[(solo, boot)] = run_search(zs, q, b, k)
    ref = results[("0", "0")]
    for key, (s, i) in results.items():
        assert torch.equal(s, ref[0]) and torch.equal(i, ref[1]), key
    ws, wi = oracle.stable_topk(bf16_scores(q, b), k)
    assert (ref[1] == wi).float().mean().item() > 0.999


def test_solo_is_one_launch(zs, monkeypatch):
    monkeypatch.delenv("ZSAAC_SOLO", raising=False)
    b = helpers.seeded((50_000, 1024), 1).cuda()
    q = helpers.seeded((975, 1024), 2).cuda()
    rb = zs.RelatedBank.from_tensor(b)
    rb.search(q, 10)
    n0 = rb.launch_count
    rb.search(q, 10)
    rb.search(q[:300], 10)
    torch.cuda.synchronize()
    assert rb.launch_count - n0 == 2          # one kernel per search
    # one 128-row tile against a long bank is an HBM-bound stream: three launches chained by
    # programmatic dependent launch are faster there (profiles/r02), unless single-launch is forced
    rb.search(q[:1], 10)
    torch.cuda.synchronize()
    assert rb.launch_count - n0 == 5
    rb.close()


def test_solo_bf16_queries_and_raw_dot(zs, monkeypatch):
    b = helpers.seeded((3000, 1024), 3)
    q = helpers.seeded((77, 1024), 4)
    out = {}
    for solo in ("0", "1"):
        monkeypatch.setenv("ZSAAC_SOLO", solo)
        out[solo] = (run_search(zs, q.bfloat16(), b, 9), run_search(zs, q, b, 9, normalize=False,
                                                                  normalize_queries=False))
    for a, c in zip(out["0"], out["1"]):
        assert torch.equal(a[0], c[0]) and torch.equal(a[1], c[1])


# ------------------------------------------------------------------------------ exact fp32 small banks
def _check_exact(s, i, q, b, k, normalize):
    ref = torch.from_numpy(oracle.exact_scores(q, b, normalize=normalize))       # float64
    got_scores = ref.gather(1, i)
    assert (s.double() - got_scores).abs().max().item() < 2e-6
    kth = torch.sort(ref, dim=1, descending=True).values[:, k - 1:k]
    # every returned index is a true top-k member up to fp32 rounding, none is missing
    assert (got_scores >= kth - 2e-6).all()
    clear = ref > kth + 2e-6
    hit = torch.zeros_like(clear)
    hit.scatter_(1, i, True)
    assert (hit | ~clear).all()
    assert (s[:, :-1] >= s[:, 1:]).all() if k > 1 else True


@pytest.mark.parametrize("Q,N,k,normalize", [(1, 527, 3, False), (32, 527, 3, False), (64, 527, 5, False),
                                             (65, 527, 3, False), (1, 10, 1, False), (200, 5000, 10, True),
                                             (5, 33, 33, True), (3, 1, 1, False), (1045, 5225, 1, True)])
def test_exact_topk_matches_float64(zs, Q, N, k, normalize):
    from zsaac_b200.retrieval import exact_topk
    q, b = helpers.seeded((Q, 1024), Q + N), helpers.seeded((N, 1024), N + k)
    if not normalize:
        q, b = torch.nn.functional.normalize(q, dim=-1), torch.nn.functional.normalize(b, dim=-1)
    s, i = exact_topk(q.cuda(), b.cuda(), k, normalize=normalize)
    torch.cuda.synchronize()
    _check_exact(s.cpu(), i.cpu(), q, b, k, normalize)
    # CPU inputs are accepted (copied), result identical
    s2, i2 = exact_topk(q, b, k, normalize=normalize)
    assert torch.equal(s2.cpu(), s.cpu()) and torch.equal(i2.cpu(), i.cpu())


def test_exact_topk_ties_self_and_errors(zs):
    from zsaac_b200.retrieval import exact_topk
    b = helpers.seeded((300, 256), 1)
    b[200] = b[7]
    b[250] = b[7]
    q = b[7:8].clone()
    s, i = exact_topk(q, b, 4, normalize=True)
    assert i[0, :3].tolist() == [7, 200, 250]                    # exact ties: ascending index
    s, i = exact_topk(q, b, 3, normalize=True, self_index=torch.tensor([7]))
    assert i[0, :2].tolist() == [200, 250]
    with pytest.raises(RuntimeError, match="out of range"):
        exact_topk(q, b, 301)
    s, i = zs.related_topk(q, b, 2, precision="fp32")
    assert i[0].tolist() == [7, 200]
    with pytest.raises(ValueError):
        zs.related_topk(q, helpers.seeded((70_000, 256), 2), 2, precision="fp32")


def test_sound_effect_choice_is_fp32_exact(zs):
    """utils.py:131-137 at its real shapes: indices equal torch's fp32 formula (no near-tie band)."""
    from zsaac_b200.utils import sound_effect_choice
    labels = torch.nn.functional.normalize(helpers.seeded((527, 1024), 31), dim=-1)
    for shape in ((1, 1024), (32, 1, 1024), (8, 1024), (100, 1024)):
        prefix = torch.nn.functional.normalize(helpers.seeded(shape, 32 + len(shape)), dim=-1)
        got = sound_effect_choice(prefix, labels, 3)
        want = oracle.sound_effect_choice(prefix, labels, 3)
        assert got.dtype == torch.int64 and not got.is_cuda and got.shape == want.shape
        assert torch.equal(got, want)
    # a bank too large for the fp32 route: tensor-core search + fp32 re-scoring
    big = torch.nn.functional.normalize(helpers.seeded((70_000, 1024), 33), dim=-1)
    prefix = torch.nn.functional.normalize(helpers.seeded((4, 1024), 34), dim=-1)
    assert torch.equal(sound_effect_choice(prefix, big, 3), oracle.sound_effect_choice(prefix, big, 3))


def test_zero_shot_predict(zs):
    from zsaac_b200 import zero_shot
    text = torch.nn.functional.normalize(helpers.seeded((10, 1024), 41), dim=-1)
    audio = torch.nn.functional.normalize(helpers.seeded((50, 1, 1024), 42), dim=-1)
    pred = zero_shot.predict(audio, text)
    assert pred.shape == (50, 1)
    want = torch.stack([oracle.zero_shot_predict(a, text) for a in audio])
    assert torch.equal(pred.cpu().reshape(-1), want.reshape(-1))


# ------------------------------------------------------------------------------------- metrics, exact
@pytest.mark.parametrize("name", list(recipes.RETRIEVAL_CASES))
def test_metrics_equal_the_reference_except_fp32_ties(name):
    """retrieval/tools/utils.py:169-251 — ranks, top1 and every metric equal the reference's own
    outputs (golden); a difference is only tolerated where the float64 scores show a tie within
    fp32 rounding (2e-6) with the ground truth, where np.argsort's order is arbitrary."""
    import zsaac_b200  # noqa: F401
    from zsaac_b200 import retrieval_metrics
    audio, caps = recipes.make_retrieval_inputs(recipes.RETRIEVAL_CASES[name])
    g = helpers.golden(name)
    a_t, c_t = torch.from_numpy(audio), torch.from_numpy(caps)
    n_audio = audio.shape[0] // 5
    for fn, key in ((retrieval_metrics.a2t, "a2t"), (retrieval_metrics.t2a, "t2a")):
        full = fn(audio, caps, return_ranks=True)
        ranks, top1 = full[7], full[8]
        ref_ranks, ref_top1, ref_metrics = g[key + "_ranks"], g[key + "_top1"], g[key + "_metrics"]
        if key == "a2t":
            tg = torch.arange(n_audio).unsqueeze(1) * 5 + torch.arange(5).unsqueeze(0)
            s = torch.from_numpy(oracle.exact_scores(a_t[0:5 * n_audio:5], c_t))
        else:
            tg = (torch.arange(5 * n_audio) // 5).unsqueeze(1)
            s = torch.from_numpy(oracle.exact_scores(c_t[:5 * n_audio], a_t[0::5]))
        ts = s.gather(1, tg)
        tied = (((s.unsqueeze(1) - ts.unsqueeze(2)).abs() < 2e-6).sum(dim=2) - 1).max(dim=1).values.numpy()
        diff = np.abs(ranks - ref_ranks)
        assert (diff <= tied).all(), (diff.max(), tied.max())
        if tied.max() == 0:
            assert np.array_equal(ranks, ref_ranks) and np.array_equal(top1, ref_top1)
            np.testing.assert_allclose(np.array(full[:7]), ref_metrics, rtol=0, atol=1e-9)
        else:
            loose = 100.0 * (tied > 0).mean() + 1e-9
            np.testing.assert_allclose(np.array(full[:4]), ref_metrics[:4], atol=loose)


def test_metrics_positions_are_distinct_for_tied_siblings():
    """Two identical captions of one audio take consecutive positions (AP10 must not
    see the same position twice)."""
    import zsaac_b200  # noqa: F401
    from zsaac_b200 import retrieval_metrics
    from zsaac_b200.retrieval import exact_rank
    rng = np.random.RandomState(5)
    audio = np.repeat(rng.randn(40, 1024).astype(np.float32), 5, axis=0)
    caps = (audio + 0.3 * rng.randn(200, 1024)).astype(np.float32)
    caps[1] = caps[0]                       # siblings of audio 0, bit-identical
    caps[3] = caps[0]
    tg = torch.arange(40).unsqueeze(1) * 5 + torch.arange(5).unsqueeze(0)
    pos, _ = exact_rank(torch.from_numpy(audio[::5]), torch.from_numpy(caps), tg)
    pos = pos.cpu().numpy()
    assert all(len(set(r.tolist())) == 5 for r in pos)
    want = oracle.a2t(audio, caps)
    got = retrieval_metrics.a2t(audio, caps)
    np.testing.assert_allclose(np.array(got), np.array(want[0]), atol=1e-9)


# ------------------------------------------------------------------------------------- collate hook
def test_collate_with_sound_effects_matches_per_sample_reference(zs):
    """dataset/dataset.py:365-368 + :632-647: one batched retrieval in collate == the reference's
    per-sample calls in __getitem__ followed by its collate."""
    from zsaac_b200.dataset import collate_with_sound_effects
    labels = [f"Label {j}" for j in range(527)]
    bank = torch.nn.functional.normalize(helpers.seeded((527, 1024), 51), dim=-1)

    def parse_entities(tokenizer, selected, mask_probability):          # stand-in for utils.py:178-188
        return torch.tensor([int(x.split()[1]) + 1 for x in selected] + [0] * (len(selected[0]) % 3))

    def padding_captions(hard_prompts, lengths):                         # restated utils.py:190-208
        m = max(lengths)
        out = torch.stack([torch.cat((h, torch.zeros(m - h.shape[0], dtype=torch.int64) - 1)) for h in hard_prompts])
        mask = out.ge(0)
        out[~mask] = 0
        return out, mask.float()

    prefixes = torch.nn.functional.normalize(helpers.seeded((32, 1, 1024), 52), dim=-1)
    batch = [(torch.arange(4) + n, torch.ones(4), prefixes[n]) for n in range(32)]
    tokens, mask, prefix, hp, hp_mask = collate_with_sound_effects(
        batch, sound_effect_embeddings=bank, sound_effect_labels=labels, sound_effect_num=3,
        tokenizer=None, parse_entities=parse_entities, padding_captions=padding_captions)
    want_hp = []
    for n in range(32):
        idx = oracle.sound_effect_choice(prefixes[n], bank, 3).squeeze(0)
        want_hp.append(parse_entities(None, [labels[i].lower() for i in list(idx)], 0))
    w, wm = padding_captions(want_hp, [len(h) for h in want_hp])
    assert torch.equal(hp, w) and torch.equal(hp_mask, wm)
    assert prefix.shape == (32, 1, 1024) and tokens.shape == (32, 4) and mask.shape == (32, 4)
    ev = collate_with_sound_effects([(f"id{n}", prefixes[n]) for n in range(5)], sound_effect_embeddings=bank,
                                    sound_effect_labels=labels, sound_effect_num=3, tokenizer=None,
                                    parse_entities=parse_entities, padding_captions=padding_captions)
    assert len(ev) == 4 and ev[0] == tuple(f"id{n}" for n in range(5)) and torch.equal(ev[2], w[:5][:, :ev[2].shape[1]])


def test_sound_effect_embeddings_choice_matches_the_model_method(zs):
    """models/caption_model.py:15-21 (the method clap_to_gpt calls in every forward pass): the
    chosen label embeddings, `bank[index].squeeze(1)`, on the bank's device."""
    from zsaac_b200.utils import sound_effect_choice, sound_effect_embeddings_choice
    bank = torch.nn.functional.normalize(helpers.seeded((527, 1024), 51), dim=-1)
    prefixes = torch.nn.functional.normalize(helpers.seeded((32, 1, 1024), 52), dim=-1)

    def reference(prefix, k):                                            # the method's body, on the CPU
        return bank[oracle.sound_effect_choice(prefix, bank, k)].squeeze(1)

    bank_gpu = bank.cuda()
    for prefix, k in ((prefixes, 3), (prefixes[:, 0], 3), (prefixes[:, 0], 1), (prefixes[:5], 1)):
        want = reference(prefix, k)
        got = sound_effect_embeddings_choice(prefix.cuda(), bank_gpu, k)
        assert got.is_cuda and got.dtype == torch.float32 and got.shape == want.shape
        assert torch.equal(got.cpu(), want)
        # the same rows the index-returning mirror names
        idx = sound_effect_choice(prefix, bank, k)
        assert torch.equal(got.cpu(), bank[idx].squeeze(1))
    # a CPU bank is ranked on the GPU, the rows come back on the bank's device
    got = sound_effect_embeddings_choice(prefixes[:4], bank, 3)
    assert got.device.type == "cpu" and torch.equal(got, reference(prefixes[:4], 3))
    # autograd as in the reference: into the label embeddings (if they are trained), not the prefix
    param = torch.nn.Parameter(bank_gpu.clone())
    out = sound_effect_embeddings_choice(prefixes[:2].cuda().requires_grad_(), param, 3)
    out.sum().backward()
    assert param.grad is not None and int((param.grad.abs().sum(dim=1) > 0).sum()) >= 3


# ------------------------------------------------------------------------------------- bank cache
def test_bank_cache_eviction_never_frees_a_bank_in_use(zs):
    """an evicted RelatedBank may still be held by a suspended process_data generator."""
    from zsaac_b200 import retrieval
    zs.clear_bank_cache()
    first = helpers.seeded((500, 1024), 61).cuda()
    held = retrieval.bank_for(first, normalize=True)
    others = [helpers.seeded((300, 1024), 62 + j).cuda() for j in range(6)]
    for t in others:                                    # pushes `first` out of the 4-entry cache
        retrieval.bank_for(t, normalize=True)
    assert not held.closed
    s, i = held.search(first[:10], 1)                   # still a live native context
    torch.cuda.synchronize()
    assert i[:, 0].tolist() == list(range(10))
    held.close()
    with pytest.raises(RuntimeError, match="closed"):
        held.search(first[:10], 1)
    zs.clear_bank_cache()
