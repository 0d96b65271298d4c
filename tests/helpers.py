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
