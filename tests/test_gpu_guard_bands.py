"""GPU (-m gpu): no entry point writes outside the output arrays it was given.

compute-sanitizer is closed on the GPU pool ("runs under it have left GPUs needing a reset"), so
the memcheck evidence is our own: every caller-owned output lives in the middle of a larger
allocation whose head and tail are filled with a bit pattern that no result can produce, and the
bands must be intact after the call — for ragged sizes (Q and N off the tile sizes, k off the
list sizes), the single-launch and three-launch searches, several passes (k > 32), self-exclusion,
the fp32 re-scoring, the gather and the k-way merge."""
import pytest
import torch

import helpers

pytestmark = pytest.mark.gpu

BAND = 4096          # elements on either side


@pytest.fixture(scope="module")
def zs():
    import zsaac_b200
    assert torch.cuda.is_available()
    zsaac_b200.load_library()
    return zsaac_b200


class Banded:
    """A contiguous [shape] view in the middle of a sentinel-filled buffer."""

    def __init__(self, shape, dtype):
        n = 1
        for s in shape:
            n *= s
        self.n = n
        self.buf = torch.empty(n + 2 * BAND, dtype=dtype, device="cuda")
        if dtype == torch.float32:
            self.buf.view(torch.int32).fill_(0x7FC0DEAD)          # a NaN payload no kernel produces
        else:
            self.buf.fill_(-0x5EADBEEF5EADBEE)
        self.pattern = self.buf[:BAND].clone()
        self.view = self.buf[BAND:BAND + n].view(*shape)

    def intact(self):
        a = self.buf[:BAND].view(torch.int32 if self.buf.dtype == torch.float32 else torch.int64)
        b = self.buf[BAND + self.n:].view(a.dtype)
        p = self.pattern.view(a.dtype)
        return bool(torch.equal(a, p) and torch.equal(b, p))


@pytest.mark.parametrize("Q,N,k,solo", [
    (1, 257, 1, None), (17, 3000, 10, None), (129, 3001, 5, "1"), (129, 3001, 5, "0"),
    (300, 9000, 32, "1"), (300, 9000, 33, "1"), (300, 9000, 33, "0"), (975, 49838, 10, None),
    (1045, 19195, 5, None), (257, 511, 100, None), (4097, 1000, 12, None)])
def test_search_stays_inside_its_outputs(zs, monkeypatch, Q, N, k, solo):
    if solo is not None:
        monkeypatch.setenv("ZSAAC_SOLO", solo)
    bank = helpers.seeded((N, 1024), N + k, "cuda")
    q = helpers.seeded((Q, 1024), Q + k, "cuda")
    rb = zs.RelatedBank.from_tensor(bank)
    for self_index in (None, torch.arange(Q, device="cuda") % N):
        if self_index is not None and k > N - 1:
            continue
        s, i = Banded((Q, k), torch.float32), Banded((Q, k), torch.int64)
        rb.search(q, k, self_index=self_index, out=(s.view, i.view))
        torch.cuda.synchronize()
        assert s.intact() and i.intact()
        assert torch.isfinite(s.view).all() and (i.view >= 0).all() and (i.view < N).all()
        if self_index is not None:
            assert not (i.view == self_index[:, None]).any()
    rb.close()


def test_rescore_gather_merge_stay_inside_their_outputs(zs):
    N, Q, k = 5003, 333, 7
    bank = helpers.seeded((N, 1024), 11, "cuda")
    q = helpers.seeded((Q, 1024), 12, "cuda")
    rb = zs.RelatedBank.from_tensor(bank)
    fb = rb.normalize_rows(bank)
    _, cand = rb.search(q, k + 8)
    s, i = Banded((Q, k), torch.float32), Banded((Q, k), torch.int64)
    rb.rescore(q, fb, cand, k, out=(s.view, i.view))
    torch.cuda.synchronize()
    assert s.intact() and i.intact()
    assert (s.view[:, :-1] >= s.view[:, 1:]).all()
    rows = rb.gather_rows(fb, i.view)
    assert torch.equal(rows, fb[i.view])
    # two shard-local results merged: compare with the single-bank result
    parts = []
    for lo, hi in ((0, 2500), (2500, N)):
        shard = zs.RelatedBank.from_tensor(bank[lo:hi], index_offset=lo)
        parts.append(shard.search(q, k))
        torch.cuda.synchronize()
        shard.close()
    ms, mi = rb.merge(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
    ws, wi = rb.search(q, k)
    torch.cuda.synchronize()
    assert torch.equal(ms, ws) and torch.equal(mi, wi)
    rb.close()


@pytest.mark.parametrize("solo", ["0", "1"])
def test_search_window_equals_a_bank_of_those_rows(zs, monkeypatch, solo):
    """zs_bank_window (the multi-GPU path's movable shard boundaries): searching rows [lo, lo + n)
    of a stored bank returns the bits a bank holding exactly those rows returns — indices global."""
    monkeypatch.setenv("ZSAAC_SOLO", solo)
    N, Q = 20011, 300
    bank = helpers.seeded((N, 1024), 21, "cuda")
    bank[9000] = bank[12]                                   # a duplicate inside some windows only
    q = helpers.seeded((Q, 1024), 22, "cuda")
    q[0] = bank[12]
    rb = zs.RelatedBank.from_tensor(bank, index_offset=1000)
    full = rb.search(q, 10)
    for lo, n, k in ((0, N, 10), (0, 5000, 10), (4999, 7001, 10), (8999, 2, 2), (12345, N - 12345, 40), (513, 300, 33)):
        rb.window(lo, n)
        assert rb.plan(Q, k)[0] >= 1
        self_index = (torch.arange(Q, device="cuda") % n) + lo + 1000
        for si in (None, self_index):
            if si is not None and k > n - 1:
                continue
            s, i = rb.search(q, k, self_index=si)
            part = zs.RelatedBank.from_tensor(bank[lo:lo + n], index_offset=1000 + lo)
            ws, wi = part.search(q, k, self_index=si)
            torch.cuda.synchronize()
            part.close()
            assert torch.equal(s, ws) and torch.equal(i, wi), (lo, n, k, si is not None)
            assert (i >= 1000 + lo).all() and (i < 1000 + lo + n).all()
    with pytest.raises(RuntimeError):
        rb.window(N - 5, 6)                                 # outside the stored rows
    rb.window()                                             # lifted: the whole bank again
    s, i = rb.search(q, 10)
    torch.cuda.synchronize()
    assert torch.equal(s, full[0]) and torch.equal(i, full[1])
    # rank_of / debug_scores always see the whole bank
    rb.window(100, 1000)
    ranks, _ = rb.rank_of(q[:5], torch.tensor([12, 9000, 3, 4, 5], device="cuda") + 1000)
    rb.window()
    ranks_full, _ = rb.rank_of(q[:5], torch.tensor([12, 9000, 3, 4, 5], device="cuda") + 1000)
    torch.cuda.synchronize()
    assert torch.equal(ranks, ranks_full)
    rb.close()
