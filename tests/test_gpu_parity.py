"""GPU (-m gpu): the CUDA path, called through the C ABI (ctypes), against the CPU oracle.

Tolerances (BASELINE.json north_star): scores within 1e-3 absolute of the fp32 oracle (the kernel
multiplies bf16-rounded operands with fp32 accumulation); index sets identical except at
near-ties (oracle score gap < 1e-3); shard merge bit-exact with the single-bank result.
"""
import os

import pytest
import torch

import helpers
from oracle import oracle

pytestmark = pytest.mark.gpu

SCORE_TOL = 1e-3     # north_star: |score - fp32 reference| <= 1e-3
BF16_TOL = 1e-4      # against the same product of bf16-rounded operands: accumulation order only


@pytest.fixture(scope="module")
def zs():
    import zsaac_b200
    assert torch.cuda.is_available()
    zsaac_b200.load_library()          # fails loudly if the extension is missing
    return zsaac_b200


@pytest.fixture(params=["1", "2", "auto"])
def cta_group(request, monkeypatch):
    if request.param == "auto":
        monkeypatch.delenv("ZSAAC_CTA_GROUP", raising=False)
    else:
        monkeypatch.setenv("ZSAAC_CTA_GROUP", request.param)
    return request.param


def bf16_scores(q, b, normalize=True):
    """fp32 product of the bf16-rounded operands the kernel sees.  Rows are normalised with the
    library's own fp32 normalise (same arithmetic as its normalise+cast kernel), so the only
    difference left to the fused kernel is the accumulation order."""
    import zsaac_b200
    if normalize:
        helper = zsaac_b200.RelatedBank(1, q.shape[1])
        q = helper.normalize_rows(q.float().cuda()).cpu()
        b = helper.normalize_rows(b.float().cuda()).cpu()
        helper.close()
    return q.bfloat16().float() @ b.bfloat16().float().T


def run_search(zs, q, b, k, **kw):
    rb = zs.RelatedBank.from_tensor(b.cuda(), normalize=kw.pop("normalize", True))
    try:
        s, i = rb.search(q.cuda(), k, **kw)
        torch.cuda.synchronize()
        return s.cpu(), i.cpu()
    finally:
        rb.close()


# ------------------------------------------------------------------------------------------ scores
@pytest.mark.parametrize("Q,N,d", [(128, 256, 64), (1, 257, 1024), (300, 1000, 1024),
                                   (129, 255, 1024), (257, 513, 512), (1045, 19195, 1024)])
def test_score_matrix_matches_oracle(zs, cta_group, Q, N, d):
    q, b = helpers.seeded((Q, d), Q + N), helpers.seeded((N, d), Q * N + 1)
    rb = zs.RelatedBank.from_tensor(b.cuda())
    got = rb.debug_scores(q.cuda()).cpu()
    rb.close()
    assert (got - bf16_scores(q, b)).abs().max().item() < BF16_TOL
    exact = torch.from_numpy(oracle.exact_scores(q, b)).float()
    # the 1e-3 contract is stated at the path's d = 1024; bf16 rounding of unit vectors averages
    # over fewer terms at smaller d (error ~ 1/sqrt(d)), so the bound is scaled for the d < 1024
    # plumbing cases
    tol = SCORE_TOL * (1024 / d) ** 0.5
    assert (got - exact).abs().max().item() < tol


# ------------------------------------------------------------------------------------------ top-k
SHAPES = [
    # Q, N, k
    (1045, 19195, 5),     # BASELINE config 1: Clotho eval
    (975, 49838, 10),     # BASELINE config 2: AudioCaps
    (1, 527, 3),          # one query vs an AudioSet-label-sized bank
    (3, 255, 1),          # single partial tile, k = 1
    (130, 256, 32),       # k = ZS_PASS_K, exact tile multiple, 2 query tiles
    (257, 70001, 17),     # ragged everywhere
    (64, 40, 32),         # k close to N
    (33, 33, 32),         # k = N - 1
]


@pytest.mark.parametrize("Q,N,k", SHAPES)
def test_topk_matches_oracle(zs, cta_group, Q, N, k):
    q, b = helpers.seeded((Q, 1024), 3 * Q + N + k), helpers.seeded((N, 1024), 5 * N + k)
    s, i = run_search(zs, q, b, k)
    rep = oracle.check_topk(s, i, q, b, k, score_tol=SCORE_TOL, tie_tol=1e-3)
    assert rep["ok"], rep
    # against identically rounded operands the result must be exact (no ties in Gaussian data)
    ws, wi = oracle.stable_topk(bf16_scores(q, b), k)
    assert (s - ws).abs().max().item() < BF16_TOL
    assert (i == wi).float().mean().item() > 0.999


@pytest.mark.parametrize("k", [1, 2, 7, 8, 9, 11, 12, 13, 16, 17, 23, 24, 25, 31, 32])
def test_every_list_size(zs, cta_group, k):
    """The register list has 8 / 12 / 16 / 24 / 32 physical slots (unused ones pinned with +inf):
    every k around those boundaries, on a bank with duplicated rows (exact ties)."""
    q, b = helpers.seeded((260, 1024), 70 + k), helpers.seeded((9000, 1024), 71 + k)
    b[4000:4100] = b[100:200]                                  # exact duplicates: ties by index
    s, i = run_search(zs, q, b, k)
    ws, wi = oracle.stable_topk(bf16_scores(q, b), k)
    assert (s - ws).abs().max().item() < BF16_TOL
    assert (i == wi).float().mean().item() > 0.999
    assert oracle.check_topk(s, i, q, b, k)["ok"]


def test_k_equals_n(zs, cta_group):
    q, b = helpers.seeded((5, 1024), 1), helpers.seeded((20, 1024), 2)
    s, i = run_search(zs, q, b, 20)
    assert torch.equal(torch.sort(i, dim=1).values, torch.arange(20).expand(5, 20))
    assert oracle.check_topk(s, i, q, b, 20)["ok"]


def test_clustered_bank_near_ties(zs, cta_group):
    b = helpers.clustered(20000, 1024, 512, 0.05, 11)
    q = b[torch.randperm(20000, generator=torch.Generator().manual_seed(1))[:300]] \
        + 0.01 * helpers.seeded((300, 1024), 12)
    s, i = run_search(zs, q, b, 10)
    rep = oracle.check_topk(s, i, q, b, 10, score_tol=SCORE_TOL, tie_tol=1e-3)
    assert rep["ok"], rep


def test_duplicate_rows_tie_order_is_ascending_index(zs, cta_group):
    b = helpers.seeded((3000, 1024), 21)
    for dst in (17, 700, 2999):
        b[dst] = b[4]                                   # 4 identical rows across tiles / chunks
    q = b[4:5].clone()
    s, i = run_search(zs, q, b, 6)
    assert i[0, :4].tolist() == [4, 17, 700, 2999]
    assert s[0, 0] == s[0, 1] == s[0, 2] == s[0, 3]


def test_zero_norm_rows_and_queries(zs):
    b = helpers.seeded((600, 1024), 31)
    b[100] = 0.0                                        # F.normalize eps: stays the zero vector
    q = helpers.seeded((4, 1024), 32)
    q[2] = 0.0
    s, i = run_search(zs, q, b, 5)
    assert torch.isfinite(s).all()
    assert (s[2] == 0).all()                            # zero query: every similarity is 0
    assert i[2].tolist() == [0, 1, 2, 3, 4]             # all tied -> ascending index
    rep = oracle.check_topk(s[[0, 1, 3]], i[[0, 1, 3]], q[[0, 1, 3]], b, 5)
    assert rep["ok"], rep


def test_bf16_inputs_and_unnormalised_dot_product(zs):
    q = torch.nn.functional.normalize(helpers.seeded((70, 1024), 41), dim=-1)
    b = torch.nn.functional.normalize(helpers.seeded((5000, 1024), 42), dim=-1)
    s32, i32 = run_search(zs, q, b, 8, normalize=False, normalize_queries=False)
    s16, i16 = run_search(zs, q.bfloat16(), b.bfloat16(), 8, normalize=False, normalize_queries=False)
    assert torch.equal(i32, i16) and torch.equal(s32, s16)     # same bf16 operands either way
    assert oracle.check_topk(s32, i32, q, b, 8, normalize=False)["ok"]
    # raw dot product of un-normalised data: ranking by q.b, not by cosine
    q2, b2 = helpers.seeded((9, 64), 43) * 3.0, helpers.seeded((900, 64), 44) * 0.5
    s, i = run_search(zs, q2, b2, 4, normalize=False, normalize_queries=False)
    ref = q2.bfloat16().float() @ b2.bfloat16().float().T
    ws, wi = oracle.stable_topk(ref, 4)
    assert torch.equal(i, wi) and (s - ws).abs().max() < 1e-3 * ws.abs().max()


def test_self_exclusion(zs, cta_group):
    b = helpers.seeded((5000, 1024), 51)
    q = b[:300] + 0.02 * helpers.seeded((300, 1024), 52)
    self_index = torch.arange(300)
    s0, i0 = run_search(zs, q, b, 5)
    assert (i0[:, 0] == self_index).all()                     # reference behaviour: top-1 = itself
    s, i = run_search(zs, q, b, 5, self_index=self_index.cuda())
    assert not (i == self_index[:, None]).any()
    assert oracle.check_topk(s, i, q, b, 5, self_index=self_index)["ok"]
    assert torch.equal(i[:, :4], i0[:, 1:]) and torch.equal(s[:, :4], s0[:, 1:])
    # negative entries mean "no exclusion"
    mixed = self_index.clone()
    mixed[::2] = -1
    s2, i2 = run_search(zs, q, b, 5, self_index=mixed.cuda())
    assert torch.equal(i2[::2], i0[::2]) and torch.equal(i2[1::2], i[1::2])


# ------------------------------------------------------------------------------ shard-merge exactness
@pytest.mark.parametrize("world", [2, 4, 8])
def test_shard_merge_is_bit_exact(zs, cta_group, world):
    """Bank split into `world` row ranges on ONE GPU (shards are just row ranges): merged result
    must be bit-identical to the single-bank result."""
    from zsaac_b200.sharded import shard_bounds
    N, Q, k = 30011, 500, 32
    b = helpers.clustered(N, 1024, 300, 0.08, 61).cuda()
    b[N // 2 + 1] = b[10]                                        # tie across a shard boundary
    q = (b[:Q] + 0.05 * helpers.seeded((Q, 1024), 62).cuda()).contiguous()
    whole = zs.RelatedBank.from_tensor(b)
    s1, i1 = whole.search(q, k)
    parts_s, parts_i = [], []
    for lo, hi in shard_bounds(N, world):
        shard = zs.RelatedBank.from_tensor(b[lo:hi], index_offset=lo)
        s, i = shard.search(q, k)
        parts_s.append(s)
        parts_i.append(i)
        torch.cuda.synchronize()
        shard.close()
    ms, mi = whole.merge(torch.stack(parts_s), torch.stack(parts_i))
    torch.cuda.synchronize()
    assert torch.equal(mi, i1) and torch.equal(ms, s1)
    os_, oi = oracle.merge_lists(torch.stack(parts_s).cpu(), torch.stack(parts_i).cpu())
    assert torch.equal(oi, mi.cpu()) and torch.equal(os_, ms.cpu())
    whole.close()


def test_merge_strided_views_and_validation(zs):
    rb = zs.RelatedBank(256, 64)
    S, Q, k = 3, 11, 4
    raw = torch.sort(helpers.seeded((S, Q, k), 71), dim=2, descending=True).values.cuda()
    idx = torch.stack([torch.arange(k) + 100 * s for s in range(S)]).unsqueeze(1).expand(S, Q, k).contiguous().cuda()
    pad_s = torch.zeros(S, Q * k + 6, device="cuda")
    pad_s[:, :Q * k] = raw.view(S, -1)
    ms, mi = rb.merge(pad_s[:, :Q * k].view(S, Q, k), idx)
    os_, oi = oracle.merge_lists(raw.cpu(), idx.cpu())
    assert torch.equal(ms.cpu(), os_) and torch.equal(mi.cpu(), oi)
    with pytest.raises(ValueError):
        rb.merge(raw, idx[:, :, :2])
    rb.close()


# ------------------------------------------------------------------------------------ error behaviour
def test_argument_errors_raise(zs):
    b = helpers.seeded((40, 1024), 81).cuda()
    rb = zs.RelatedBank.from_tensor(b)
    q = helpers.seeded((2, 1024), 82).cuda()
    with pytest.raises(RuntimeError, match="out of range"):     # torch.topk wording
        rb.search(q, 41)
    with pytest.raises(RuntimeError, match="out of range"):
        rb.search(q, 1025)                                      # > ZS_MAX_K
    with pytest.raises(RuntimeError, match="out of range"):
        rb.search(q, 0)
    with pytest.raises(RuntimeError, match="out of range"):
        rb.search(q, 40, self_index=torch.zeros(2, dtype=torch.int64).cuda())
    with pytest.raises(ValueError):
        rb.search(q[:, :512], 3)
    with pytest.raises(TypeError):
        rb.search(q.half(), 3)
    with pytest.raises(ValueError):
        rb.search(q.cpu(), 3)
    s, i = rb.search(q[:0], 3)                                  # empty batch is a no-op
    assert s.shape == (0, 3) and i.shape == (0, 3)
    rb.close()
    with pytest.raises(RuntimeError):
        zs.RelatedBank(10, 1000)                                # d not a multiple of 64


def test_normalize_and_gather_kernels(zs):
    x = helpers.seeded((1000, 1024), 91) * 7.0
    x[3] = 0.0
    rb = zs.RelatedBank(1, 1024)
    got = rb.normalize_rows(x.cuda()).cpu()
    want = oracle.normalize_rows(x)
    assert (got - want).abs().max().item() < 1e-6
    idx = torch.tensor([[5, 0, 999], [3, 3, 17]])
    g = rb.gather_rows(got.cuda(), idx.cuda()).cpu()
    assert torch.equal(g, got[idx])
    rb.close()


# ------------------------------------------------------------------------ fp32 re-scoring
@pytest.mark.parametrize("k", [1, 5, 10, 24, 32])
def test_rescore_fp32_restores_the_fp32_ranking(zs, k):
    """k + 8 bf16 candidates re-scored in fp32: scores are fp32 cosines (1e-6, not 1e-4), and the
    index lists equal the fp32 oracle's except where the oracle itself is within 1e-6 of a tie.
    (k = 32 leaves no margin: still exact scores, order as good as the candidates allow.)"""
    q, b = helpers.seeded((600, 1024), 500 + k), helpers.clustered(30_000, 1024, 2048, 0.2, 501 + k)
    s0, i0 = zs.related_topk(q.cuda(), b.cuda(), k)
    s1, i1 = zs.related_topk(q.cuda(), b.cuda(), k, rescore_fp32=True)
    torch.cuda.synchronize()
    s0, i0, s1, i1 = s0.cpu(), i0.cpu(), s1.cpu(), i1.cpu()
    full = oracle.exact_scores(q, b)                       # float64 cosines
    ws, wi = oracle.stable_topk(torch.as_tensor(full, dtype=torch.float64), k)
    at = torch.as_tensor(full).gather(1, i1)
    assert (s1.double() - at).abs().max().item() < 2e-6
    assert (s1[:, 1:] <= s1[:, :-1]).all()
    rep = oracle.check_topk(s1, i1, q, b, k)
    assert rep["ok"], rep
    match1 = (i1 == wi).all(dim=1).float().mean().item()
    match0 = (i0 == wi).all(dim=1).float().mean().item()
    assert match1 >= match0
    if k <= 24:
        differs = (i1 != wi)
        gap = (torch.as_tensor(full).gather(1, i1) - ws).abs()
        assert (gap[differs] < 1e-6).all()                  # only fp32-level ties may differ


def test_rescore_f32_standalone_edge_cases(zs):
    """zs_rescore_f32 on hand-made candidate lists: ignored slots (-1, out of range), k == kc,
    raw dot product, shard offset, ties broken by index."""
    g = torch.Generator().manual_seed(77)
    bank = torch.randn(400, 256, generator=g)
    bank[11] = bank[7]                                      # exact tie
    q = torch.randn(3, 256, generator=g)
    rb = zs.RelatedBank(1, 256)                             # any context: the call needs no bf16 bank
    cand = torch.tensor([[7, 11, -1, 399, 5000, 3], [0, 1, 2, 3, 4, 5], [11, 7, -1, -1, -1, 9]])
    s, i = rb.rescore(q.cuda(), bank.cuda(), cand.cuda(), 4, normalize=True)
    torch.cuda.synchronize()
    s, i = s.cpu(), i.cpu()
    cos = torch.nn.functional.normalize(q, dim=-1) @ torch.nn.functional.normalize(bank, dim=-1).T
    for r in range(3):
        valid = sorted({c for c in cand[r].tolist() if 0 <= c < 400}, key=lambda c: (-cos[r, c].item(), c))
        want = (valid + [-1] * 4)[:4]
        assert i[r].tolist() == want, (r, i[r].tolist(), want)
        for slot, c in enumerate(want):
            if c >= 0:
                assert abs(s[r, slot].item() - cos[r, c].item()) < 1e-6
            else:
                assert s[r, slot].item() == float("-inf")
    row0 = i[0].tolist()
    assert row0.index(7) + 1 == row0.index(11)              # the twins tie exactly: lower index first
    # raw dot product, k == kc, bank slice with a global offset
    s2, i2 = rb.rescore(q.cuda(), bank[100:].cuda(), torch.tensor([[100, 101, 250]] * 3).cuda(), 3,
                        normalize=False, index_offset=100)
    torch.cuda.synchronize()
    dots = q @ bank.T
    for r in range(3):
        want = sorted([100, 101, 250], key=lambda c: (-dots[r, c].item(), c))
        assert i2[r].cpu().tolist() == want
        assert torch.allclose(s2[r].cpu(), dots[r, want], atol=1e-4)
    with pytest.raises(RuntimeError):
        rb.rescore(q.cuda(), bank.cuda(), cand.cuda(), 7)   # k > kc
    rb.close()


# --------------------------------------------------------- full-size, size-independent properties
def test_wavcaps_scale_properties(zs):
    """BASELINE config 3 size (8192 x 400k, k=10): checked through properties, not a full oracle."""
    N, Q, k = 400_000, 8192, 10
    g = torch.Generator(device="cuda").manual_seed(103)
    b = torch.randn(N, 1024, device="cuda", generator=g)
    q = torch.randn(Q, 1024, device="cuda", generator=g)
    planted = torch.randperm(N, device="cuda", generator=g)[:Q]
    q[:4096] = b[planted[:4096]] + 0.3 * q[:4096]               # first half: a known nearest row
    rb = zs.RelatedBank.from_tensor(b)
    s, i = rb.search(q, k)
    s2, i2 = rb.search(q, k)
    torch.cuda.synchronize()
    assert torch.equal(s, s2) and torch.equal(i, i2)            # deterministic / idempotent
    assert (i[:4096, 0] == planted[:4096]).all()                # planted neighbour is top-1
    assert (s[:, 1:] <= s[:, :-1]).all() and ((i >= 0) & (i < N)).all()
    assert (torch.sort(i, dim=1).values.diff(dim=1) != 0).all()
    # positive scaling of queries / bank rows cannot change a cosine ranking
    s3, i3 = rb.search(q * 3.5, k)
    assert (i3 == i).float().mean().item() > 0.999 and (s3 - s).abs().max().item() < 1e-3
    # bank permutation: same scores bit for bit, indices mapped through the permutation
    perm = torch.randperm(N, device="cuda", generator=g)
    rbp = zs.RelatedBank.from_tensor(b[perm])
    sp, ip = rbp.search(q[:1024], k)
    torch.cuda.synchronize()
    assert torch.equal(sp, s[:1024])
    mism = perm[ip] != i[:1024]
    tie = torch.zeros_like(mism)
    eq = s[:1024, 1:] == s[:1024, :-1]
    tie[:, 1:] |= eq
    tie[:, :-1] |= eq
    assert not (mism & ~tie).any()            # indices may only differ inside exact score ties
    # a 512-query slice against the fp32 oracle
    rep = oracle.check_topk(s[4000:4512].cpu(), i[4000:4512].cpu(), q[4000:4512].cpu(), b.cpu(), k)
    assert rep["ok"], rep
    rb.close()
    rbp.close()


def _gpu_free_gb():
    free, _ = torch.cuda.mem_get_info()
    return free / 2 ** 30


def test_synthetic_10m_scale_properties(zs):
    """BASELINE config 4 at full size (65,536 queries x 10 M-row bank, k=32) through properties:
    planted rows come back first, lists are sorted / distinct / in range, the search is
    idempotent, and a 48-query sample is re-scored against the whole bank in fp32 blocks."""
    if _gpu_free_gb() < 60:
        pytest.skip("needs ~30 GB of free HBM")
    N, Q, k, blk = 10_000_000, 65_536, 32, 65_536
    dev = torch.device("cuda")

    def block(bi):
        g = torch.Generator(device=dev).manual_seed(204 * 2 ** 32 + bi)
        return torch.randn(min(blk, N - bi * blk), 1024, device=dev, generator=g)

    rb = zs.RelatedBank(N, 1024, device=dev)
    for bi in range(-(-N // blk)):
        rb.upload(block(bi), bi * blk)
    gq = torch.Generator(device=dev).manual_seed(104)
    q = torch.randn(Q, 1024, device=dev, generator=gq)
    planted = torch.randint(0, N, (256,), device=dev, generator=gq)
    for j, r in enumerate(planted.tolist()):            # query j is (a scaled copy of) bank row r
        q[j] = 2.5 * block(r // blk)[r % blk]
    s, i = rb.search(q, k)
    s2, i2 = rb.search(q, k)
    torch.cuda.synchronize()
    assert torch.equal(s, s2) and torch.equal(i, i2)
    assert (i[:256, 0] == planted).all() and (s[:256, 0] > 0.995).all()
    assert (s[:, 1:] <= s[:, :-1]).all() and ((i >= 0) & (i < N)).all()
    assert (torch.sort(i, dim=1).values.diff(dim=1) != 0).all()
    # fp32 re-scoring of a sample against every bank block (the oracle's arithmetic, blocked)
    rows = torch.cat([torch.arange(300, 324), torch.arange(Q - 24, Q)]).to(dev)
    qs = torch.nn.functional.normalize(q[rows], dim=-1)
    best_s = torch.full((rows.numel(), k), -float("inf"), device=dev)
    best_i = torch.zeros((rows.numel(), k), dtype=torch.int64, device=dev)
    for bi in range(-(-N // blk)):
        sc = qs @ torch.nn.functional.normalize(block(bi), dim=-1).T
        cs, ci = sc.topk(k, dim=1)
        ms, sel = torch.cat([best_s, cs], 1).topk(k, dim=1)
        best_i = torch.cat([best_i, ci + bi * blk], 1).gather(1, sel)
        best_s = ms
    got_s, got_i = s[rows], i[rows]
    assert (got_s - best_s).abs().max().item() <= SCORE_TOL
    kth = best_s[:, -1:]
    clear = best_s > kth + SCORE_TOL                     # oracle entries clearly inside the top-k
    present = (best_i.unsqueeze(2) == got_i.unsqueeze(1)).any(dim=2)
    assert (present | ~clear).all()
    rb.close()


def test_allpairs_400k_self_exclusion_properties(zs):
    """BASELINE config 5 at full size: 400 k noise-injected bank rows as queries against the
    same bank, self-exclusion, k=5 (reference utils.py:19-31 for the noise)."""
    if _gpu_free_gb() < 20:
        pytest.skip("needs ~10 GB of free HBM")
    N, k = 400_000, 5
    g = torch.Generator(device="cuda").manual_seed(105)
    b = torch.nn.functional.normalize(torch.randn(N, 1024, device="cuda", generator=g), dim=-1)
    b[7] = b[3]                                          # a twin: each must find the other first
    q = b + torch.randn(N, 1024, device="cuda", generator=g) * (0.001 ** 0.5)
    self_index = torch.arange(N, device="cuda")
    rb = zs.RelatedBank.from_tensor(b)
    s, i = rb.search(q, k, self_index=self_index)
    torch.cuda.synchronize()
    assert not (i == self_index[:, None]).any()
    assert i[3, 0].item() == 7 and i[7, 0].item() == 3
    assert (s[:, 1:] <= s[:, :-1]).all() and ((i >= 0) & (i < N)).all()
    assert (torch.sort(i, dim=1).values.diff(dim=1) != 0).all()
    ex = slice(1000, 5096)                               # (clear of the twins)
    s_in, i_in = rb.search(q[ex], k + 1)                 # without exclusion the row itself is slot 0
    torch.cuda.synchronize()
    assert (i_in[:, 0] == self_index[ex]).all()
    assert torch.equal(i_in[:, 1:], i[ex]) and torch.equal(s_in[:, 1:], s[ex])
    sl = slice(200_000, 200_256)
    rep = oracle.check_topk(s[sl].cpu(), i[sl].cpu(), q[sl].cpu(), b.cpu(), k,
                            self_index=self_index[sl].cpu())
    assert rep["ok"], rep
    rb.close()


# ------------------------------------------------------------------ shared admission thresholds
@pytest.mark.parametrize("Q,N,k,chunks", [(300, 40_000, 10, 150), (700, 60_000, 32, 200),
                                          (129, 30_000, 5, 117), (975, 49_838, 10, 0),
                                          (2100, 21_000, 1, 80)])
def test_threshold_sharing_is_result_neutral(zs, cta_group, monkeypatch, Q, N, k, chunks):
    """Units publish the k-th score of their list per query row and later units start from it
    (SimTopkParams::row_thr).  Many forced chunks make units run in several waves, so most of them
    are seeded by earlier ones; the merged result must be identical to the unshared run."""
    q, b = helpers.seeded((Q, 1024), 400 + k), helpers.clustered(N, 1024, 512, 0.05, 401 + k)
    if chunks:
        monkeypatch.setenv("ZSAAC_CHUNKS", str(chunks))
    monkeypatch.setenv("ZSAAC_SHARE_THR", "0")
    s0, i0 = run_search(zs, q, b, k)
    monkeypatch.setenv("ZSAAC_SHARE_THR", "1")
    s1, i1 = run_search(zs, q, b, k)
    assert torch.equal(s0, s1) and torch.equal(i0, i1)
    rep = oracle.check_topk(s1[:64], i1[:64], q[:64], b, k)
    assert rep["ok"], rep


def test_threshold_sharing_keeps_index_order_on_exact_ties(zs, cta_group, monkeypatch):
    """Exact ties across chunk boundaries: a unit seeded with threshold t still admits scores
    EQUAL to t (the tie-break is the index).  Identical rows -> the first k indices; an all-zero
    query scores exactly 0.0 everywhere (pred(+0) must step below -0 as well)."""
    monkeypatch.setenv("ZSAAC_CHUNKS", "200")
    row = helpers.seeded((1, 1024), 411)
    bank = row.expand(52_000, 1024).contiguous()
    q = torch.cat([helpers.seeded((299, 1024), 412), torch.zeros(1, 1024)])
    for k in (1, 7, 32):
        s, i = run_search(zs, q, bank, k)
        assert torch.equal(i, torch.arange(k).expand(300, k)), k
        assert (s[-1] == 0).all()
    # half the bank identical and best, placed at the END: chunks that run first publish lower
    # thresholds, the tied block must still come out in ascending index order
    bank2 = torch.cat([helpers.seeded((26_000, 1024), 413), row.expand(26_000, 1024)]).contiguous()
    qq = row + 0.01 * helpers.seeded((300, 1024), 414)
    s, i = run_search(zs, qq, bank2, 32)
    assert torch.equal(i, (26_000 + torch.arange(32)).expand(300, 32))


# ---------------------------------------------------------------- adversarial orderings / hypothesis
def test_adversarially_ordered_bank(zs, cta_group):
    """Bank sorted by similarity to the queries: ascending order makes EVERY score a new maximum
    (every element takes the insert path), descending order makes none after the first k."""
    base = helpers.seeded((1, 1024), 301)
    noise = helpers.seeded((6000, 1024), 302)
    w = torch.linspace(0.0, 3.0, 6000).unsqueeze(1)          # row j is closer to `base` as j grows
    asc = noise + w * base
    q = base + 0.05 * helpers.seeded((40, 1024), 303)
    for bank in (asc, asc.flip(0)):
        for k in (1, 10, 32):
            s, i = run_search(zs, q, bank, k)
            rep = oracle.check_topk(s, i, q, bank, k)
            assert rep["ok"], (k, rep)


def test_constant_scores_everywhere(zs):
    """All bank rows identical: every score ties, the answer is the first k indices."""
    row = helpers.seeded((1, 1024), 311)
    bank = row.expand(1000, 1024).contiguous()
    q = helpers.seeded((3, 1024), 312)
    s, i = run_search(zs, q, bank, 7)
    assert torch.equal(i, torch.arange(7).expand(3, 7))
    assert (s == s[:, :1]).all()


def test_random_shapes_property(zs):
    """Random (Q, N, k, offsets) against the oracle; hypothesis drives the shapes."""
    from hypothesis import given, settings, strategies as st, HealthCheck

    @settings(max_examples=12, deadline=None, suppress_health_check=list(HealthCheck))
    @given(q=st.integers(1, 400), n=st.integers(1, 3000), k=st.integers(1, 32),
           off=st.integers(0, 10_000_000), seed=st.integers(0, 10_000))
    def check(q, n, k, off, seed):
        k = min(k, n)
        qs, bank = helpers.seeded((q, 1024), seed), helpers.seeded((n, 1024), seed + 1)
        rb = zs.RelatedBank.from_tensor(bank.cuda(), index_offset=off)
        s, i = rb.search(qs.cuda(), k)
        torch.cuda.synchronize()
        rb.close()
        rep = oracle.check_topk(s.cpu(), i.cpu() - off, qs, bank, k)
        assert rep["ok"], (q, n, k, off, rep)

    check()
