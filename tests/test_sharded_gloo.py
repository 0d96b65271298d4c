"""CPU, world_size 2, gloo: the host logic of the multi-GPU path — shard bounds, global index
offsets, the packed [scores | indices] all-gather layout and the merge call — with the per-rank
CUDA bank replaced by an oracle-backed stand-in (tests may use the oracle; the product cannot)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class OracleBank:
    """Same surface as zsaac_b200.RelatedBank (upload / search / merge), computed by the oracle."""

    def __init__(self, rows, dim, device, index_offset):
        from oracle import oracle
        self.oracle = oracle
        self.rows, self.dim, self.index_offset = rows, dim, index_offset
        self.device = torch.device("cpu")
        self.bank = torch.zeros(rows, dim)
        self.win = (0, rows)

    def window(self, row_lo=0, n_rows=0):
        self.win = (row_lo, n_rows) if n_rows else (0, self.rows)

    def upload(self, rows, dst_row=0, *, normalize=True):
        rows = self.oracle.normalize_rows(rows) if normalize else rows.float()
        self.bank[dst_row:dst_row + rows.shape[0]] = rows

    def search(self, queries, k, *, normalize_queries=True, self_index=None, out=None):
        w_lo, w_n = self.win
        local_self = None
        if self_index is not None:
            local_self = self_index - self.index_offset - w_lo
            local_self = torch.where((local_self >= 0) & (local_self < w_n), local_self,
                                     torch.full_like(local_self, -1))
        q = self.oracle.normalize_rows(queries) if normalize_queries else queries
        s, i = self.oracle.cosine_topk(q, self.bank[w_lo:w_lo + w_n], k, normalize=False, self_index=local_self)
        i = i + self.index_offset + w_lo
        if out is not None:
            out[0].copy_(s)
            out[1].copy_(i)
            return out
        return s, i

    def merge(self, scores, indices):
        return self.oracle.merge_lists(scores, indices)


def _worker(rank, world, port, n_rows, k, exclude_self, result_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import zsaac_b200  # noqa: F401
        from zsaac_b200.sharded import ShardedRelatedBank
        from oracle import oracle
        g = torch.Generator().manual_seed(77)
        bank = torch.randn(n_rows, 128, generator=g)
        bank[n_rows // 2 + 3] = bank[5]                # a duplicate straddling the shard boundary
        queries = torch.randn(37, 128, generator=g)
        queries[0] = bank[5]
        sb = ShardedRelatedBank(n_rows, 128, local_bank_factory=OracleBank)
        assert (sb.lo, sb.hi) == oracle.shard_bounds(n_rows, world)[rank]
        sb.upload_global(bank)
        self_index = torch.arange(37) if exclude_self else None
        s, i = sb.search(queries, k, self_index=self_index)
        ws, wi = oracle.cosine_topk(queries, bank, k, self_index=self_index)
        assert torch.equal(i, wi), (rank, i[0], wi[0])
        assert torch.equal(s, ws)
        if not exclude_self:
            assert i[0, :2].tolist() == [5, n_rows // 2 + 3]      # tie: ascending global index
        # every rank uploads a slice of the host batch, the slices are all-gathered
        for n_q in (37, 2, 1):
            assert torch.equal(sb.replicate_from_host(queries[:n_q]), queries[:n_q])
        with pytest.raises(RuntimeError, match="out of range"):
            sb.search(queries, n_rows)                              # k larger than a shard
        # shards sized by (made-up) GPU speeds: same global answer
        sw = ShardedRelatedBank(n_rows, 128, local_bank_factory=OracleBank, shard_weights=[1.0, 2.5])
        assert sw.bounds[0][1] == sw.bounds[1][0] and sw.bounds[0][1] - sw.bounds[0][0] < n_rows // 2
        sw.upload_global(bank)
        s2, i2 = sw.search(queries, k, self_index=self_index)
        assert torch.equal(i2, wi) and torch.equal(s2, ws)
        # adaptive boundaries: every rank stores half a shard beyond its own; wherever the
        # boundary is moved to, the global answer stays the same
        sa = ShardedRelatedBank(n_rows, 128, local_bank_factory=OracleBank, overlap=0.5)
        assert sa.adaptive and sa.stores[0][1] > sa.base_bounds[0][1] and sa.stores[1][0] < sa.base_bounds[1][0]
        sa.upload_global(bank)
        for cut in (sa.base_bounds[0][1], max(10, sa.stores[1][0]), min(n_rows - 10, sa.stores[0][1]), n_rows // 2 - 7):
            sa.set_bounds([(0, cut), (cut, n_rows)])
            s3, i3 = sa.search(queries, k, self_index=self_index)
            assert torch.equal(i3, wi) and torch.equal(s3, ws), cut
        with pytest.raises(ValueError, match="stored rows|contiguous"):
            sa.set_bounds([(0, n_rows), (n_rows, n_rows)])
        sa.set_bounds(sa.base_bounds)
        new = sa.rebalance(10.0 if rank == 0 else 12.0)            # rank 1 is slower: it gives rows away
        assert new == sa.bounds and new[0][1] > sa.base_bounds[0][1]
        s4, i4 = sa.search(queries, k, self_index=self_index)
        assert torch.equal(i4, wi) and torch.equal(s4, ws)
        torch.save((s, i), os.path.join(result_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("exclude_self", [False, True])
def test_sharded_search_world2_gloo(tmp_path, exclude_self):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, 301, 6, exclude_self, str(tmp_path)), nprocs=world, join=True)
    a = torch.load(tmp_path / "rank0.pt")
    b = torch.load(tmp_path / "rank1.pt")
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])     # every rank holds the result


def _writer_worker(rank, world, port, n_items, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import zsaac_b200  # noqa: F401
        from zsaac_b200 import related_pipeline as rp
        first, last = rp.item_range(n_items, rank, world)
        # what process_data yields on this rank in multi-GPU mode: its contiguous block of items
        mine = ({"text_id": i, "related_embeddings": torch.full((2, 4), float(i))} for i in range(first, last))
        rp.save_data_to_hdf5(mine, out_path, n_items)
        assert not os.path.exists(f"{out_path}.rank{rank:03d}")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_items", [(2, 11), (3, 3), (2, 1)])
def test_multi_rank_writer_emits_the_single_gpu_stream(tmp_path, world, n_items):
    """save_data_to_hdf5 with a process group up (reference :30-34 on N GPUs): every rank writes
    the block of records it processed, the rank files land at their offsets of the output, which
    is appended to like the reference's 'ab' — record order = item order."""
    import pickle
    out = tmp_path / "out.pkl"
    with open(out, "wb") as f:
        pickle.dump({"text_id": -1}, f)                    # the file already holds a record
    mp.spawn(_writer_worker, args=(world, _free_port(), n_items, str(out)), nprocs=world, join=True)
    got = []
    with open(out, "rb") as f:
        while True:
            try:
                got.append(pickle.load(f))
            except EOFError:
                break
    assert [g["text_id"] for g in got] == [-1] + list(range(n_items))
    assert all(torch.equal(g["related_embeddings"], torch.full((2, 4), float(g["text_id"]))) for g in got[1:])
    sys.path.insert(0, ROOT)
    import zsaac_b200  # noqa: F401
    from zsaac_b200.related_pipeline import item_range
    ranges = [item_range(n_items, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n_items
    assert all(ranges[r][1] == ranges[r + 1][0] for r in range(world - 1))


def test_rebalanced_bounds_controller():
    """The boundary controller (sharded.rebalanced_bounds): deterministic, contiguous, inside the
    stored rows, converging to equal times for constant speeds."""
    sys.path.insert(0, ROOT)
    import zsaac_b200  # noqa: F401
    from zsaac_b200.sharded import rebalanced_bounds, shard_bounds
    n, world = 10_000_000, 8
    base = shard_bounds(n, world)
    margin = 157_184
    stores = [(max(0, lo - margin), min(n, hi + margin)) for lo, hi in base]
    ms = [112.2, 109.7, 107.9, 113.4, 109.7, 109.7, 111.5, 107.7]       # measured on 8 B200s of one box
    speed = [(hi - lo) / t for (lo, hi), t in zip(base, ms)]
    cur = base
    for _ in range(6):
        t = [(hi - lo) / v for (lo, hi), v in zip(cur, speed)]
        nxt = rebalanced_bounds(cur, t, stores)
        assert nxt[0][0] == 0 and nxt[-1][1] == n
        assert all(nxt[r][1] == nxt[r + 1][0] for r in range(world - 1))
        assert all(stores[r][0] <= lo < hi <= stores[r][1] for r, (lo, hi) in enumerate(nxt))
        cur = nxt
    t = [(hi - lo) / v for (lo, hi), v in zip(cur, speed)]
    assert max(t) - min(t) < 0.2 and max(t) < 110.5                     # from 113.4: the mean is 110.2
    # equal times: nothing moves (up to the alignment); garbage timings: nothing moves at all
    assert all(abs(a[1] - b[1]) <= 256 for a, b in zip(rebalanced_bounds(base, [5.0] * world, stores), base))
    assert rebalanced_bounds(base, [1.0, float("nan")] + [1.0] * 6, stores) == [tuple(b) for b in base]
    assert rebalanced_bounds(base, [0.0] * world, stores) == [tuple(b) for b in base]
    # a rank that is 3x slower is clamped at what its neighbours store
    slow = rebalanced_bounds(base, [3.0] + [1.0] * 7, stores, damping=1.0)
    assert slow[0][1] == stores[1][0]


def _gate_worker(rank, world, port, result_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import bench
        import zsaac_b200  # noqa: F401
        from zsaac_b200.sharded import shard_bounds
        dev = torch.device("cpu")
        n_rows, n_q, k, seed = 2500, 24, 6, 991
        _, rows = next(bench.bank_rows_fp32(torch, dev, seed, 0, n_rows))
        bank_n = torch.nn.functional.normalize(rows, dim=-1)
        q = bench.gen_queries(torch, n_q, 13)
        s = torch.nn.functional.normalize(q, dim=-1).bfloat16().float() @ bank_n.bfloat16().float().T
        top = torch.sort(s, dim=1, descending=True, stable=True)
        res = (top.values[:, :k].contiguous(), top.indices[:, :k].contiguous())
        lo, hi = shard_bounds(n_rows, world)[rank]
        g = bench.parity_gate(torch, dist, world, dev, lo, hi, seed, q, None, k, res)
        assert g["ok"] and g["sampled_queries"] == n_q, g
        bad = res[1].clone()
        bad[:, 0] = res[1][:, 1]
        g_bad = bench.parity_gate(torch, dist, world, dev, lo, hi, seed, q, None, k, (res[0], bad))
        assert not g_bad["ok"] and g_bad["clear_winners_missing"] > 0
        torch.save(g, os.path.join(result_dir, f"gate{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_bench_parity_gate_is_shard_independent_gloo(tmp_path):
    """bench.parity_gate at N > 1: every rank re-scores ITS shard in fp32, the shard-local fp32
    top-k lists are all-gathered and merged — every rank reaches the verdict one GPU would."""
    world = 3
    mp.spawn(_gate_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    gates = [torch.load(tmp_path / f"gate{r}.pt") for r in range(world)]
    assert all(g == gates[0] for g in gates)
