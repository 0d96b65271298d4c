"""CPU model of the fused kernel's selection logic (simtopk_kernel.cuh): per-(chunk, column half)
sorted lists, the shared per-row admission threshold (order-preserving key, atomicMax, pred(t)
seeding, one-tile-stale reads) and the final k-way merge — executed under RANDOM unit schedules.

The claim under test ("shared admission threshold"): whatever the interleaving of
units and however stale the threshold a unit reads, the merged result is the exact top-k under the
total order (score desc, index asc) — including exact ties, +0.0 / -0.0 and denormals, where the
pred() step matters.  This is a model of the algorithm, not of the CUDA code; the CUDA code is
checked against the oracle in tests/test_gpu_parity.py.
"""
import math
import random
import struct

import numpy as np
import pytest

TILE = 256          # bank rows per tile
HALF = 128          # columns per epilogue thread and tile
NEG_INF = float("-inf")
KEY_NEG_INF = 0x007FFFFF


def f32_bits(x: float) -> int:
    return struct.unpack("<I", struct.pack("<f", x))[0]


def bits_f32(u: int) -> float:
    return struct.unpack("<f", struct.pack("<I", u & 0xFFFFFFFF))[0]


def score_key(x: float) -> int:
    """score_key() of the kernel: unsigned compare == float compare, -0 < +0."""
    u = f32_bits(x)
    return (~u) & 0xFFFFFFFF if u & 0x80000000 else u | 0x80000000


def seed_below(key: int) -> float:
    """seed_below() of the kernel: the largest float strictly below the published score."""
    if key <= KEY_NEG_INF:
        return NEG_INF
    kk = key - 1
    f = bits_f32(kk & 0x7FFFFFFF) if kk & 0x80000000 else bits_f32(~kk)
    return bits_f32(0x80000001) if f == 0.0 else f


def test_key_is_order_preserving_and_pred_steps_below():
    vals = [NEG_INF, -3.5, -1e-38, bits_f32(0x80000001), -0.0, 0.0, bits_f32(1), 1e-38, 0.25, 1.0, 3e38]
    keys = [score_key(v) for v in vals]
    assert keys == sorted(keys) and len(set(keys)) == len(keys)
    assert score_key(NEG_INF) == KEY_NEG_INF
    assert seed_below(0) == NEG_INF and seed_below(KEY_NEG_INF) == NEG_INF
    for v in vals[1:]:
        s = seed_below(score_key(v))
        assert s < v, (v, s)                      # strictly below, also across the signed zeros
        assert not (v > s) is False
    assert seed_below(score_key(0.0)) < 0.0 and seed_below(score_key(-0.0)) < 0.0


class UnitList:
    """One epilogue thread's register list for one work unit."""

    def __init__(self, k):
        self.k = k
        self.entries = []                        # (score, column), best first, arrival order on ties

    def kth(self):
        return self.entries[-1][0] if len(self.entries) == self.k else NEG_INF

    def insert(self, v, col):
        pos = len(self.entries)
        while pos > 0 and v > self.entries[pos - 1][0]:      # strict: equal scores keep arrival order
            pos -= 1
        self.entries.insert(pos, (v, col))
        del self.entries[self.k:]


def run_model(scores_row: np.ndarray, k: int, chunks: int, rng: random.Random, share: bool,
              self_col: int = -1):
    """One query row against n = len(scores_row) bank rows, split into `chunks` chunks of whole
    tiles x 2 column halves; units advance tile by tile in a random interleaving."""
    n = len(scores_row)
    n_tiles = -(-n // TILE)
    tpc = -(-n_tiles // chunks)
    units = []
    for c in range(chunks):
        t0, t1 = c * tpc, min((c + 1) * tpc, n_tiles)
        if t0 >= t1:
            continue
        for half in range(2):
            units.append({"t": t0, "t1": t1, "half": half, "list": UnitList(k), "seed": NEG_INF,
                          "published": NEG_INF, "next_key": 0, "started": False})
    row_key = 0                                  # the shared threshold, zeroed per search
    active = list(range(len(units)))
    max_running = rng.randint(1, max(1, len(units)))          # how many units run "at once"
    running = []
    while active or running:
        while active and len(running) < max_running:
            running.append(active.pop(rng.randrange(len(active)) if rng.random() < 0.5 else 0))
        u = units[running[rng.randrange(len(running))]]
        if not u["started"]:                     # unit start: first read of the shared key
            u["next_key"] = row_key if share else 0
            u["started"] = True
        # tile start: consume the key read one tile ago, request the next one
        u["seed"] = max(u["seed"], seed_below(u["next_key"]))
        if share:
            u["next_key"] = row_key if rng.random() < 0.8 else u["next_key"]   # sometimes staler
        lst = u["list"]
        thr = max(lst.kth(), u["seed"])
        c0 = u["t"] * TILE + u["half"] * HALF
        for col in range(c0, min(c0 + HALF, n)):
            v = float(scores_row[col])
            if v > thr and col != self_col and not math.isnan(v):
                lst.insert(v, col)
                thr = max(lst.kth(), u["seed"])
        kth = lst.kth()
        if share and kth > u["published"]:
            row_key = max(row_key, score_key(kth))            # atomicMax
            u["published"] = kth
        u["t"] += 1
        if u["t"] == u["t1"]:
            running.remove(units.index(u))
    # k-way merge of all partial lists under (score desc, index asc)
    pool = [e for u in units for e in u["list"].entries]
    pool.sort(key=lambda e: (-e[0], e[1]))
    return pool[:k]


def exact_topk(scores_row, k, self_col=-1):
    cand = [(float(v), c) for c, v in enumerate(scores_row) if c != self_col and not math.isnan(float(v))]
    cand.sort(key=lambda e: (-e[0], e[1]))
    return cand[:k]


@pytest.mark.parametrize("kind", ["gauss", "quantised", "zeros", "constant", "ascending", "tiny"])
@pytest.mark.parametrize("share", [True, False])
def test_merged_result_is_exact_under_any_schedule(kind, share):
    rng = random.Random(hash((kind, share)) & 0xFFFF)
    np_rng = np.random.RandomState(rng.randrange(1 << 30))
    for trial in range(12):
        n = rng.choice([rng.randint(1, 600), rng.randint(600, 6000)])
        k = rng.randint(1, min(32, n))
        chunks = rng.randint(1, max(1, -(-n // TILE)))
        if kind == "gauss":
            row = np_rng.standard_normal(n).astype(np.float32)
        elif kind == "quantised":                               # many exact ties
            row = (np_rng.randint(-3, 4, n) / 4.0).astype(np.float32)
        elif kind == "zeros":                                   # +0 / -0 / denormals around the seed step
            row = np_rng.choice(np.array([0.0, -0.0, 1e-45, -1e-45, 1e-3], dtype=np.float32), n)
        elif kind == "constant":
            row = np.full(n, 0.37, dtype=np.float32)
        elif kind == "ascending":                               # every element a new maximum
            row = np.sort(np_rng.standard_normal(n).astype(np.float32))
        else:
            row = (np_rng.standard_normal(n) * 1e-40).astype(np.float32)   # denormal scores
        self_col = rng.randrange(n) if n > k and rng.random() < 0.3 else -1
        if self_col >= 0:
            k = min(k, n - 1)
        got = run_model(row, k, chunks, rng, share, self_col)
        want = exact_topk(row, k, self_col)
        assert [c for _, c in got] == [c for _, c in want], (kind, n, k, chunks, trial)
        assert [s for s, _ in got] == [s for s, _ in want]
