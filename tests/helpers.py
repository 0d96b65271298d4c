"""Shared test helpers: input recipes (same as tests/golden/make_golden.py) and record builders."""
import os
import pickle
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as recipes  # noqa: E402  (import-safe: touches /root/reference only when run)

GOLDEN = os.path.join(HERE, "golden")
D = recipes.D


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def case_records(case):
    x = recipes.make_embeddings(case)
    recs = [{"caption": f"caption number {i} of the synthetic set", "text_id": i,
             "text_embedding": torch.from_numpy(x[i:i + 1].copy())} for i in range(case["n"])]
    return x, recs


def write_case_files(case, recs, tmpdir):
    paths, lo = [], 0
    for fi, cnt in enumerate(case["files"]):
        p = os.path.join(str(tmpdir), f"in{fi}.pkl")
        with open(p, "wb") as f:
            pickle.dump(recs[lo:lo + cnt], f)
        paths.append(p)
        lo += cnt
    return paths


def read_related_stream(path):
    """The reader loop of the reference's dataset/dataset.py:64-78 / :401-417, restated (that
    module does not import under transformers 5.x): pickle.load until EOFError, splice lists."""
    all_data = []
    with open(path, "rb") as f:
        while True:
            try:
                item = pickle.load(f)
                if type(item) is list:
                    all_data = all_data + item
                else:
                    all_data.append(item)
            except EOFError:
                break
    return all_data


def seeded(shape, seed, device="cpu"):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(*shape, generator=g).to(device)


def clustered(n, d, centres, sigma, seed):
    g = torch.Generator().manual_seed(seed)
    c = torch.randn(centres, d, generator=g)
    a = torch.randint(0, centres, (n,), generator=g)
    return c[a] + sigma * torch.randn(n, d, generator=g)


def check_related_rows(items, x, g, k, *, score_atol, tie_tol, order_tol=0.0, bank_cpu=None):
    """The related_embeddings of every record against a golden fixture of the reference, all
    records at once (the per-record loop costs a minute at n = 2048).

    items[i]['related_embeddings'] [k, d] must be rows of the normalised input, best first
    (adjacent scores may rise by at most order_tol), with the scores the reference's rows have
    (score_atol); where a slot holds another row than the reference's, the two must be tied to
    within tie_tol in float64 (either is then a correct answer); where it holds the same row, the
    row's float64 checksum must be the reference's — and, given the caller's fp32 bank, the row
    must be a bit-identical copy of that bank row.  Returns the [n, k] bool array "same row"."""
    from oracle import oracle
    n = len(items)
    xt = torch.from_numpy(x)
    xn = torch.nn.functional.normalize(xt, dim=-1)
    exact = oracle.exact_scores(xt, xt)
    rel = torch.stack([it["related_embeddings"] for it in items])            # [n, k, d]
    assert tuple(rel.shape) == (n, k, D) and rel.dtype == torch.float32
    mine = (rel.reshape(n * k, D) @ xn.T).argmax(dim=1).reshape(n, k).numpy()     # which input row is it?
    theirs = g["related_index"]
    score = torch.einsum("nd,nkd->nk", xn, rel).numpy()
    assert (np.diff(score, axis=1) <= order_tol).all()
    np.testing.assert_allclose(score, g["related_score"], atol=score_atol)
    same = mine == theirs
    rows = np.arange(n)[:, None].repeat(k, axis=1)
    tie_gap = np.abs(exact[rows, mine] - exact[rows, theirs])
    assert (tie_gap[~same] < tie_tol).all(), np.argwhere(~same & ~(tie_gap < tie_tol))[:5]
    rowsum = rel.double().sum(dim=2).numpy()
    np.testing.assert_allclose(rowsum[same], g["related_rowsum"][same], atol=1e-5)
    if bank_cpu is not None:
        picked = bank_cpu[torch.from_numpy(mine)]                            # [n, k, d]
        identical = (picked == rel).all(dim=2).numpy()
        assert identical[same].all()      # same index => a bit-identical copy of the caller's bank row
    return same
