"""CPU: the HOST logic of process_data (embeddings_related_generator.py:19-28 on batches) — batch
cutting, the ragged last batch, the lazy generator, self-exclusion indices, and in multi-rank mode
the item blocks / query slices / step count every rank must agree on — with the CUDA pieces
(bank, SearchPipeline, pinned memory, streams) replaced by oracle-backed stand-ins.  The product
itself has no CPU path; the GPU tests run the same function against the real kernels."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = 64


class OnGpu(torch.Tensor):
    """A CPU tensor that answers is_cuda = True (process_data refuses CPU banks)."""
    is_cuda = True


class FakeLocalBank:
    device = torch.device("cpu")

    def gather_rows(self, src, indices):
        return torch.Tensor(src)[indices]

    def close(self):
        pass


class FakeShardedBank:
    def __init__(self, n_rows, dim, device=None):
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        per = -(-n_rows // self.world)
        self.lo, self.hi = min(self.rank * per, n_rows), min((self.rank + 1) * per, n_rows)
        self.local = FakeLocalBank()
        self.full = None

    def upload_global(self, bank, normalize=True):
        self.full = torch.Tensor(bank)


class FakePipeline:
    """SearchPipeline(result="row_slice", input="slice") on the CPU: rank r's rows of the host batch
    are all-gathered, every rank ranks ITS rows against the whole bank with the oracle."""
    made = []

    def __init__(self, bank, n_queries, k, *, depth, from_host, to_host, result, rescore_from,
                 excludes_self, input):
        assert from_host and not to_host and result == "row_slice" and input == "slice" and depth == 2
        self.sharded = isinstance(bank, FakeShardedBank)
        self.world = bank.world if self.sharded else 1
        self.rank = bank.rank if self.sharded else 0
        self.bank_rows = bank.full if self.sharded else bank.rows
        assert n_queries % self.world == 0
        self.per, self.k = n_queries // self.world, k
        self.results, self.submitted = {}, 0
        self.rescore = rescore_from is not None
        if self.rescore:        # the fp32 rows of this rank's shard
            lo, hi = (bank.lo, bank.hi) if self.sharded else (0, self.bank_rows.shape[0])
            assert torch.equal(torch.Tensor(rescore_from), self.bank_rows[lo:hi])
        FakePipeline.made.append(self)

    def submit(self, buf, self_index=None):
        from oracle import oracle
        assert tuple(buf.shape) == (self.world * self.per, D) and buf.dtype == torch.float32
        mine = buf[self.rank * self.per:(self.rank + 1) * self.per].clone()
        if self.world > 1:
            full = torch.empty(self.world * self.per, D)
            dist.all_gather_into_tensor(full.view(-1), mine.reshape(-1))     # what the NCCL all-gather does
            rows = full[self.rank * self.per:(self.rank + 1) * self.per]
            assert torch.equal(rows, mine)
        si = None
        if self_index is not None:
            assert tuple(self_index.shape) == (self.world * self.per,)
            si = self_index[self.rank * self.per:(self.rank + 1) * self.per]
            # (padding queries of a ragged batch name rows beyond the bank: the kernel ignores those)
            si = torch.where(si < self.bank_rows.shape[0], si, torch.full_like(si, -1))
        pad = mine.abs().sum(dim=1) == 0                                      # ragged tail: zero rows
        q = torch.where(pad[:, None], torch.ones_like(mine), mine)
        s, i = oracle.cosine_topk(q, self.bank_rows, self.k, self_index=si)
        slot = self.submitted % 2
        assert slot not in self.results, "a slot was re-used before its result was taken"
        self.results[slot] = (s, i)
        self.submitted += 1
        return slot

    def wait_stream(self, slot=None):
        pass

    def result_of(self, slot):
        return self.results.pop(slot)


class FakeSingleBank(FakeLocalBank):
    def __init__(self, rows):
        self.rows = torch.Tensor(rows)


def _patch(monkeypatch_or_none, rp, sharded):
    """Swap the CUDA pieces for the stand-ins (monkeypatch inside pytest, plain setattr in a
    spawned worker)."""
    put = monkeypatch_or_none.setattr if monkeypatch_or_none is not None else setattr

    class Stream:
        def synchronize(self):
            pass

    put(rp, "_require_cuda", lambda: None)
    put(rp, "bank_for", lambda bank, normalize: FakeSingleBank(bank))
    put(sharded, "SearchPipeline", FakePipeline)
    put(sharded, "ShardedRelatedBank", FakeShardedBank)
    put(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    put(torch.cuda, "current_stream", lambda device=None: Stream())
    put(torch.cuda, "synchronize", lambda device=None: None)


def _inputs(n, seed=5):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, D, generator=g) * (0.5 + 3 * torch.rand(n, 1, generator=g))
    if n > 4:
        x[n // 2] = x[1] * 2.0                               # a duplicate direction: an exact tie
    recs = [{"caption": f"c{i}", "text_id": i, "text_embedding": x[i:i + 1].clone()} for i in range(n)]
    bank = torch.nn.functional.normalize(x, dim=-1)
    return x, recs, bank


def _expected(x, bank, k, exclude_self):
    from oracle import oracle
    si = torch.arange(x.shape[0]) if exclude_self else None
    _, ids = oracle.cosine_topk(x, bank, k, self_index=si)
    return bank[ids]


@pytest.mark.parametrize("exclude_self", [False, True])
@pytest.mark.parametrize("n,batch", [(23, 7), (7, 7), (8, 7), (4, 7), (5, 16384)])
def test_process_data_host_logic_single_rank(monkeypatch, n, batch, exclude_self):
    sys.path.insert(0, ROOT)
    import zsaac_b200  # noqa: F401
    from zsaac_b200 import related_pipeline as rp, sharded
    _patch(monkeypatch, rp, sharded)
    monkeypatch.setattr(rp, "QUERY_BATCH", batch)
    FakePipeline.made.clear()
    x, recs, bank = _inputs(n)
    gen = rp.process_data(bank.as_subclass(OnGpu), iter(recs) if n == 23 else recs, 3,
                          exclude_self=exclude_self)
    assert iter(gen) is gen and not FakePipeline.made        # lazy: nothing has run yet
    want = _expected(x, bank, 3, exclude_self)
    seen = 0
    for i, item in enumerate(gen):
        assert item is recs[i]                               # the caller's dicts, mutated, in order
        assert item["text_embedding"].device.type == "cpu" and torch.equal(item["text_embedding"], x[i:i + 1])
        rel = item["related_embeddings"]
        assert tuple(rel.shape) == (3, D) and rel.dtype == torch.float32
        assert rel.untyped_storage().nbytes() == 3 * D * 4   # own storage, not a view of the batch
        assert torch.equal(rel, want[i])
        seen += 1
    assert seen == n
    pipe = FakePipeline.made[0]
    assert pipe.rescore and pipe.submitted == -(-n // min(batch, n if n != 23 else batch)) and not pipe.results
    # an empty input yields nothing and launches nothing
    FakePipeline.made.clear()
    assert list(rp.process_data(bank.as_subclass(OnGpu), [], 3)) == []
    assert not FakePipeline.made or FakePipeline.made[0].submitted == 0
    # argument errors come before any work
    with pytest.raises(ValueError, match="CUDA bank"):
        next(rp.process_data(bank, recs, 3))
    with pytest.raises(ValueError, match="dtype"):
        next(rp.process_data(bank.as_subclass(OnGpu), recs, 3, dtype="fp16"))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, batch, exclude_self, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import zsaac_b200  # noqa: F401
        from zsaac_b200 import related_pipeline as rp, sharded
        _patch(None, rp, sharded)
        rp.QUERY_BATCH = batch
        rp.tqdm = lambda it, total=None: it
        x, recs, bank = _inputs(n)
        gen = rp.process_data(bank.as_subclass(OnGpu), recs, 3, exclude_self=exclude_self, rescore_fp32=False)
        first, last = rp.item_range(n, rank, world)
        mine = []

        def tap():                                            # what save_data_to_hdf5 pulls
            for item in gen:
                mine.append(item["text_id"])
                yield item

        rp.save_data_to_hdf5(tap(), out_path, n)
        assert mine == list(range(first, last)), (rank, mine)
        pipe = FakePipeline.made[0]
        per = max(1, min(batch // world, -(-n // world)))
        assert pipe.per == per and pipe.submitted == -(-(-(-n // world)) // per) and not pipe.rescore
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,batch,exclude_self", [(2, 23, 8, False), (3, 23, 9, True), (2, 5, 64, True),
                                                        (4, 10, 8, False), (3, 4, 6, False)])
def test_process_data_host_logic_multi_rank_gloo(tmp_path, world, n, batch, exclude_self):
    """Multi-GPU mode of the generator on the CPU (gloo): every rank processes its contiguous item
    block in the same number of steps (ranks without items still join the collectives), and the
    rank files land in item order — the stream one rank writes."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    out = str(tmp_path / "out.pkl")
    mp.spawn(_worker, args=(world, _free_port(), n, batch, exclude_self, out), nprocs=world, join=True)
    items = helpers.read_related_stream(out)
    x, _, bank = _inputs(n)
    want = _expected(x, bank, 3, exclude_self)
    assert [it["text_id"] for it in items] == list(range(n))
    for i, it in enumerate(items):
        assert torch.equal(it["related_embeddings"], want[i]) and torch.equal(it["text_embedding"], x[i:i + 1])
