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

    def upload(self, rows, dst_row=0, *, normalize=True):
        rows = self.oracle.normalize_rows(rows) if normalize else rows.float()
        self.bank[dst_row:dst_row + rows.shape[0]] = rows

    def search(self, queries, k, *, normalize_queries=True, self_index=None, out=None):
        local_self = None
        if self_index is not None:
            local_self = self_index - self.index_offset
            local_self = torch.where((local_self >= 0) & (local_self < self.rows), local_self,
                                     torch.full_like(local_self, -1))
        q = self.oracle.normalize_rows(queries) if normalize_queries else queries
        s, i = self.oracle.cosine_topk(q, self.bank, k, normalize=False, self_index=local_self)
        i = i + self.index_offset
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
