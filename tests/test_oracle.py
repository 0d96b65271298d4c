"""CPU: the oracle (oracle/oracle.py) against the golden vectors produced by the reference's own
functions (tests/golden/make_golden.py), plus its internal consistency."""
import numpy as np
import pytest
import torch

import helpers
from helpers import recipes
from oracle import oracle


@pytest.mark.parametrize("name", list(recipes.CASES))
def test_literal_loop_matches_reference_golden(name):
    case = recipes.CASES[name]
    x, recs = helpers.case_records(case)
    g = helpers.golden(name)
    bank = oracle.build_bank(recs)
    # bank rows: same values as the reference's (its order is a permutation, stored in the fixture)
    np.testing.assert_allclose(bank.double().sum(dim=1).numpy()[g["bank_order"]], g["bank_rowsum"],
                               rtol=0, atol=1e-5)
    out = list(oracle.process_data_literal(bank, recs, case["k"]))
    assert len(out) == case["n"]
    # same scores as the reference picked, and the same rows, except where two candidates are tied
    # to within fp32 rounding (clustered case, duplicate rows): then either row is a correct answer
    # (torch.topk may return fp32-tied neighbours in either order: 2e-6 of slack on "best first")
    same = helpers.check_related_rows(out, x, g, case["k"], score_atol=2e-6, tie_tol=5e-6, order_tol=2e-6,
                                      bank_cpu=bank)        # rows are copies of the bank's rows (:23)
    assert same.mean() > 0.9


@pytest.mark.parametrize("name", list(recipes.CASES))
def test_batched_oracle_matches_reference_golden(name):
    case = recipes.CASES[name]
    x, recs = helpers.case_records(case)
    g = helpers.golden(name)
    q = torch.from_numpy(x)
    s, idx = oracle.cosine_topk(q, q, case["k"])
    np.testing.assert_allclose(s.numpy(), g["related_score"], atol=2e-6)
    exact = oracle.exact_scores(q, q)
    a, b = idx.numpy(), g["related_index"]
    rows = np.arange(case["n"])[:, None].repeat(case["k"], axis=1)
    # identical / fp32-tied rows: the index choice among them is arbitrary in torch.topk
    assert ((a == b) | (np.abs(exact[rows, a] - exact[rows, b]) < 5e-6)).all()
    rep = oracle.check_topk(s, idx, q, q, case["k"])
    assert rep["ok"], rep


def test_reference_never_excludes_self():
    # SURVEY §3A [probe]: slot 0 of related_embeddings is the item itself (score 1.0)
    g = helpers.golden("generator_gauss")
    assert (g["related_index"][:, 0] == np.arange(g["related_index"].shape[0])).all()
    np.testing.assert_allclose(g["related_score"][:, 0], 1.0, atol=1e-6)


@pytest.mark.parametrize("name", list(recipes.SEC_CASES))
def test_sound_effect_choice_matches_reference_golden(name):
    case = recipes.SEC_CASES[name]
    prefix, bank = recipes.make_sec_inputs(case)
    g = helpers.golden(name)
    idx = oracle.sound_effect_choice(torch.from_numpy(prefix), torch.from_numpy(bank), case["k"])
    assert idx.dtype == torch.int64
    assert idx.tolist() == g["index"].tolist()
    # softmax is monotone: same as the top-k of the raw similarity
    s, i2 = oracle.cosine_topk(torch.from_numpy(prefix), torch.from_numpy(bank), case["k"], normalize=False)
    assert i2.tolist() == g["index"].tolist()
    np.testing.assert_allclose(s.numpy(), g["score"], atol=1e-6)


def test_stable_topk_tie_order_and_self_exclusion():
    bank = torch.eye(8, 64)
    bank[5] = bank[2]                                   # exact duplicate rows
    q = bank[2:3].clone()
    s, i = oracle.cosine_topk(q, bank, 3)
    assert i[0, :2].tolist() == [2, 5] and s[0, 0] == s[0, 1] == 1.0
    s, i = oracle.cosine_topk(q, bank, 2, self_index=torch.tensor([2]))
    assert i[0, 0].item() == 5 and 2 not in i[0].tolist()


def test_zero_norm_rows_follow_f_normalize_eps():
    x = torch.zeros(2, 64)
    x[1, 0] = 3.0
    n = oracle.normalize_rows(x)
    assert torch.equal(n[0], torch.zeros(64)) and n[1, 0] == 1.0


def test_merge_lists_equals_global_topk():
    q = helpers.seeded((17, 64), 5)
    b = helpers.seeded((1000, 64), 6)
    k = 7
    gs, gi = oracle.cosine_topk(q, b, k)
    parts_s, parts_i = [], []
    for lo, hi in oracle.shard_bounds(1000, 3):
        s, i = oracle.cosine_topk(q, b[lo:hi], k)
        parts_s.append(s)
        parts_i.append(i + lo)
    ms, mi = oracle.merge_lists(torch.stack(parts_s), torch.stack(parts_i))
    assert torch.equal(mi, gi)
    # merging never changes a score's bits
    assert torch.equal(ms, torch.stack(parts_s).permute(1, 0, 2).reshape(17, -1).gather(
        1, torch.stack(parts_i).permute(1, 0, 2).reshape(17, -1).argsort(1).gather(
            1, torch.searchsorted(torch.stack(parts_i).permute(1, 0, 2).reshape(17, -1).sort(1).values, mi))))


def test_noise_injection_cosine_to_source():
    # variance 0.001 at d=1024 gives cos(query, own row) ~ 0.70
    x = torch.nn.functional.normalize(helpers.seeded((64, 1024), 9), dim=-1)
    y = oracle.noise_injection(x, 0.001, generator=torch.Generator().manual_seed(3))
    cos = (x * y).sum(dim=1)
    assert 0.6 < cos.mean().item() < 0.8


def test_check_topk_flags_wrong_answers():
    q = helpers.seeded((4, 64), 1)
    b = helpers.seeded((50, 64), 2)
    s, i = oracle.cosine_topk(q, b, 5)
    assert oracle.check_topk(s, i, q, b, 5)["ok"]
    bad_i = i.clone()
    bad_i[0, 0] = (i[0, 0] + 1) % 50 if ((i[0, 0] + 1) % 50) not in i[0].tolist() else (i[0, 0] + 7) % 50
    assert not oracle.check_topk(s, bad_i, q, b, 5)["ok"]
    assert not oracle.check_topk(s + 0.01, i, q, b, 5)["ok"]


@pytest.mark.parametrize("name", list(recipes.RETRIEVAL_CASES))
def test_retrieval_metrics_match_reference_golden(name):
    """a2t / t2a restated (oracle) vs the reference's own functions (golden)."""
    audio, caps = recipes.make_retrieval_inputs(recipes.RETRIEVAL_CASES[name])
    g = helpers.golden(name)
    m, ranks, top1, positions = oracle.a2t(audio, caps)
    np.testing.assert_allclose(np.array(m, np.float64), g["a2t_metrics"], rtol=0, atol=1e-9)
    np.testing.assert_array_equal(ranks, g["a2t_ranks"])
    np.testing.assert_array_equal(top1, g["a2t_top1"])
    assert (positions.min(axis=1) == ranks).all()
    m, ranks, top1 = oracle.t2a(audio, caps)
    np.testing.assert_allclose(np.array(m, np.float64), g["t2a_metrics"], rtol=0, atol=1e-9)
    np.testing.assert_array_equal(ranks, g["t2a_ranks"])
    np.testing.assert_array_equal(top1, g["t2a_top1"])
    assert 0 < g["a2t_metrics"][0] < 100 and 0 < g["t2a_metrics"][0] < 100   # a non-trivial fixture


@pytest.mark.parametrize("name", list(recipes.MEMORY_CASES))
def test_map2memory_matches_reference_golden(name, tmp_path):
    """map2memory / construct_support_memory restated (oracle) vs the reference's own functions."""
    import pickle
    q, bank = recipes.make_memory_inputs(recipes.MEMORY_CASES[name])
    g = helpers.golden(name)
    out = oracle.map2memory(torch.from_numpy(q), torch.from_numpy(bank))
    np.testing.assert_allclose(out.numpy(), g["out"], atol=1e-6)
    assert (g["max_weight"] < 0.9).all()         # the fixture is not a one-hot softmax
    p = tmp_path / "mem.pkl"
    caps = ["too short", "a caption that has exactly eight words in it", " ".join(["w"] * 25),
            "another caption with nine words in it right here now"]
    with open(p, "wb") as f:
        for i, c in enumerate(caps):
            pickle.dump({"caption": c, "text_embedding": torch.from_numpy(bank[i:i + 1] * (i + 2.0))}, f)
        pickle.dump([{"caption": "listed", "text_embedding": torch.from_numpy(bank[9:10] * 3.0)}], f)
    mem = oracle.construct_support_memory([str(p)])
    np.testing.assert_allclose(mem.numpy(), g["memory"], atol=1e-7)


@pytest.mark.parametrize("name", list(recipes.SEC_METHOD_CASES))
def test_sound_effect_method_matches_reference_golden(name, monkeypatch):
    """models/caption_model.py:15-21 — the caption models' method returns the chosen label
    EMBEDDINGS, `bank[index].squeeze(1)`.  Golden from the reference's own method; checked here for
    the oracle's restatement and for the product mirror's shape / gather logic (its ranking — one
    CUDA launch — replaced by the oracle's)."""
    case = recipes.SEC_METHOD_CASES[name]
    prefix, bank = recipes.make_sec_inputs(case)
    g = helpers.golden(name)
    prefix_t = torch.from_numpy(prefix).reshape(*case["lead"], helpers.D)
    bank_t = torch.from_numpy(bank)
    import zsaac_b200  # noqa: F401
    from zsaac_b200 import utils
    monkeypatch.setattr(utils, "_choice_index", lambda p, b, k: oracle.sound_effect_choice(p, b, k))
    for out in (oracle.sound_effect_choice_method(prefix_t, bank_t, case["k"]),
                utils.sound_effect_embeddings_choice(prefix_t, bank_t, case["k"])):
        assert list(out.shape) == g["shape"].tolist() and out.dtype == torch.float32
        flat = out.reshape(-1, helpers.D)
        assert torch.equal(flat, bank_t[torch.from_numpy(g["index"])])        # the reference's rows, bit for bit
        np.testing.assert_allclose(flat.double().sum(dim=1).numpy(), g["rowsum"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("name", list(recipes.COLLATE_CASES))
def test_collate_mirror_matches_the_reference_collate_golden(name, monkeypatch):
    """Golden: the reference's own __getitem__ lines (dataset/dataset.py:365-368: sound_effect_choice
    + parse_entities per sample) followed by its own collate (:632-647).  Checked here: the oracle's
    restatement of the two text helpers, and zsaac_b200.dataset.collate_with_sound_effects — ONE
    batched retrieval in collate — with its ranking (one CUDA launch) replaced by the oracle's."""
    case = recipes.COLLATE_CASES[name]
    g = helpers.golden(name)
    prefixes, bank, labels, lead = recipes.collate_samples(case)
    tok = recipes.WordTokenizer()
    # the reference's order of work, through the oracle's restatements
    hard = []
    for n in range(case["q"]):
        idx = oracle.sound_effect_choice(prefixes[n], bank, case["k"]).squeeze(0)
        hard.append(oracle.parse_entities(tok, [labels[i].lower() for i in list(idx)], 0))
    hp, hm = oracle.padding_captions(hard, [len(h) for h in hard])
    assert np.array_equal(hp.numpy(), g["hard_prompt"]) and np.array_equal(hm.numpy(), g["mask"])
    # the product's collate on samples WITHOUT the two trailing elements
    import zsaac_b200  # noqa: F401
    from zsaac_b200 import utils
    from zsaac_b200.dataset import collate_with_sound_effects
    from zsaac_b200.dataset import dataset as ds_mod
    fake = lambda p, b, k: oracle.sound_effect_choice(p, b, k)     # noqa: E731  (indices on the CPU, like the mirror)
    monkeypatch.setattr(utils, "sound_effect_choice", fake)
    monkeypatch.setattr(ds_mod, "sound_effect_choice", fake)
    batch = [(*lead[n], prefixes[n]) for n in range(case["q"])]
    out = collate_with_sound_effects(batch, sound_effect_embeddings=bank, sound_effect_labels=labels,
                                     sound_effect_num=case["k"], tokenizer=tok,
                                     parse_entities=oracle.parse_entities, padding_captions=oracle.padding_captions)
    assert len(out) == (5 if case["training"] else 4)
    assert np.array_equal(out[-2].numpy(), g["hard_prompt"]) and np.array_equal(out[-1].numpy(), g["mask"])
    assert tuple(out[-3].shape) == (case["q"], 1, helpers.D)
    assert abs(out[-3].double().sum().item() - float(g["prefix_sum"])) < 1e-9
    if case["training"]:
        assert torch.equal(out[0], torch.stack([lead[n][0] for n in range(case["q"])]))
        assert torch.equal(out[1], torch.stack([lead[n][1] for n in range(case["q"])]))
    else:
        assert out[0] == tuple(lead[n][0] for n in range(case["q"]))


@pytest.mark.parametrize("name", list(recipes.MEMORY_CASES))
def test_construct_support_memory_host_side_matches_golden(name, tmp_path, monkeypatch):
    """predict_prompt.construct_support_memory's HOST side (the reader loop of predict_prompt.py:30-47
    through dataset.read_related_records, the concatenation) on the stream the reference's golden
    was made from; the normalisation — a CUDA kernel in the product — done by torch here."""
    import pickle
    import zsaac_b200  # noqa: F401
    from zsaac_b200 import predict_prompt as pp
    _, bank = recipes.make_memory_inputs(recipes.MEMORY_CASES[name])
    g = helpers.golden(name)
    p = tmp_path / "mem.pkl"
    caps = ["too short", "a caption that has exactly eight words in it", " ".join(["w"] * 25),
            "another caption with nine words in it right here now"]
    with open(p, "wb") as f:
        for i, c in enumerate(caps):
            pickle.dump({"caption": c, "text_embedding": torch.from_numpy(bank[i:i + 1] * (i + 2.0))}, f)
        pickle.dump([{"caption": "listed", "text_embedding": torch.from_numpy(bank[9:10] * 3.0)}], f)

    class Normalizer:
        def normalize_rows(self, x):
            return torch.nn.functional.normalize(x, dim=-1)

    real_to = torch.Tensor.to
    monkeypatch.setattr(pp, "_require_cuda", lambda: None)
    monkeypatch.setattr(pp, "_helper", lambda device: Normalizer())
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.Tensor, "to", lambda self, *a, **k: self if (a and a[0] == "cuda") else real_to(self, *a, **k))
    mem = pp.construct_support_memory([str(p)])
    assert tuple(mem.shape) == tuple(g["memory"].shape)
    assert (mem - torch.from_numpy(g["memory"])).abs().max().item() < 1e-6
