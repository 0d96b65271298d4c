"""CPU, gloo, world sizes 2-4: the REAL sharded.SearchPipeline — slot rotation, the padded
[G x per] query / list layouts, the all-gather of query slices, the all-to-all of shard lists for
row-slice results, per-shard fp32 re-scoring, moving shard boundaries — with CUDA streams / events
replaced by no-ops (CPU work is already ordered) and the per-rank bank by an oracle-backed stand-in.
The GPU tests run the same class on 2 and 8 real GPUs; this is what covers world sizes 3 and 4 and
query counts that do not divide by the world size."""
import contextlib
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DIM = 128


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class NoStream:
    def wait_event(self, event):
        pass

    def wait_stream(self, stream):
        pass

    def synchronize(self):
        pass


class NoEvent:
    def __init__(self, enable_timing=False):
        pass

    def record(self, stream=None):
        pass

    def synchronize(self):
        pass

    def elapsed_time(self, other):
        return 1.0


class StandInBank:
    """RelatedBank's surface as SearchPipeline / ShardedRelatedBank use it, computed by the oracle
    (fp32; `rescore` re-ranks the candidates from the fp32 rows it is given, like zs_rescore_f32)."""

    def __init__(self, rows, dim, device, index_offset):
        from oracle import oracle
        self.oracle = oracle
        self.rows, self.dim, self.index_offset = rows, dim, index_offset
        self.device = torch.device("cpu")
        self.bank = torch.zeros(rows, dim)
        self.win = (0, rows)
        self.reserved = None

    def window(self, row_lo=0, n_rows=0):
        self.win = (row_lo, n_rows) if n_rows else (0, self.rows)

    def upload(self, rows, dst_row=0, *, normalize=True):
        rows = self.oracle.normalize_rows(rows) if normalize else rows.float()
        self.bank[dst_row:dst_row + rows.shape[0]] = rows

    def reserve(self, n_queries, k):
        self.reserved = (n_queries, k)

    def search(self, queries, k, *, normalize_queries=True, self_index=None, out=None):
        assert self.reserved is not None and queries.shape[0] <= self.reserved[0] and k <= self.reserved[1]
        w_lo, w_n = self.win
        local_self = None
        if self_index is not None:
            local_self = self_index - self.index_offset - w_lo
            local_self = torch.where((local_self >= 0) & (local_self < w_n), local_self,
                                     torch.full_like(local_self, -1))
        s, i = self.oracle.cosine_topk(queries, self.bank[w_lo:w_lo + w_n], k, normalize=normalize_queries,
                                       self_index=local_self)
        i = i + self.index_offset + w_lo
        if out is not None:
            out[0].copy_(s)
            out[1].copy_(i)
            return out
        return s, i

    def rescore(self, queries, bank_f32, candidates, k, *, normalize=True, index_offset=0, out=None):
        assert bank_f32.shape[0] == self.rows and index_offset == self.index_offset
        q = self.oracle.normalize_rows(queries) if normalize else queries
        rows = self.oracle.normalize_rows(bank_f32) if normalize else bank_f32
        local = candidates - index_offset
        s = torch.einsum("qd,qcd->qc", q, rows[local])
        # (score desc, index asc) among the candidates
        order = torch.sort(candidates, dim=1, stable=True).indices
        s, c = s.gather(1, order), candidates.gather(1, order)
        top = torch.sort(s, dim=1, descending=True, stable=True)
        res = (top.values[:, :k].contiguous(), c.gather(1, top.indices)[:, :k].contiguous())
        if out is not None:
            out[0].copy_(res[0])
            out[1].copy_(res[1])
            return out
        return res

    def merge(self, scores, indices, out=None):
        s, i = self.oracle.merge_lists(scores, indices)
        if out is not None:
            out[0].copy_(s)
            out[1].copy_(i)
            return out
        return s, i


def _worker(rank, world, port, n_rows, n_q, k):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import zsaac_b200  # noqa: F401
        from zsaac_b200.sharded import SearchPipeline, ShardedRelatedBank
        from oracle import oracle
        torch.cuda.Stream = lambda device=None: NoStream()
        torch.cuda.Event = NoEvent
        torch.cuda.stream = lambda s: contextlib.nullcontext()
        torch.cuda.current_stream = lambda device=None: NoStream()
        torch.Tensor.pin_memory = lambda self, *a, **kw: self
        g = torch.Generator().manual_seed(123)
        bank = torch.randn(n_rows, DIM, generator=g)
        bank[n_rows - 2] = bank[3]                             # a duplicate row: an exact tie across shards
        batches = [torch.randn(n_q, DIM, generator=g) for _ in range(3)]
        batches[1][0] = bank[3]
        me = torch.arange(n_q)
        sb = ShardedRelatedBank(n_rows, DIM, local_bank_factory=StandInBank)
        sb.upload_global(bank)
        shard_f32 = bank[sb.lo:sb.hi]
        per = -(-n_q // world)
        lo, hi = min(rank * per, n_q), min(rank * per + per, n_q)

        def run(**kw):
            excl = kw.pop("exclude", False)
            pipe = SearchPipeline(sb, n_q, k, depth=2, self_index=me if excl else None, **kw)
            got = []
            slots = []
            for b, q in enumerate(batches):                    # 3 batches through 2 slots
                host = q.clone()
                if kw.get("input") in ("replicate", "slice"):  # only this rank's rows are meaningful
                    host[:lo] = float("nan")
                    host[hi:] = float("nan")
                slots.append(pipe.submit(host))
                if b >= 1:                                     # take batch b-1 while b is "in flight"
                    pipe.wait_stream(slots[b - 1])
                    s, i = pipe.result_of(slots[b - 1], host=kw.get("to_host", True))
                    got.append((s.clone(), i.clone()))
            pipe.wait_stream()
            s, i = pipe.result_of(slots[2], host=kw.get("to_host", True))
            got.append((s.clone(), i.clone()))
            r_lo, r_hi = pipe.out_rows
            assert (r_lo, r_hi) == ((lo, hi) if kw.get("result") == "row_slice" else (0, n_q))
            for q, (s, i) in zip(batches, got):
                ws, wi = oracle.cosine_topk(q, bank, k, self_index=me if excl else None)
                assert tuple(s.shape) == (r_hi - r_lo, k)
                bad = (i != wi[r_lo:r_hi])
                assert torch.allclose(s, ws[r_lo:r_hi], atol=1e-6), ((s - ws[r_lo:r_hi]).abs().max(), list(kw))
                assert not bad.any(), (bad.sum(), s[bad], ws[r_lo:r_hi][bad], i[bad], wi[r_lo:r_hi][bad], list(kw))
            return pipe

        for result in ("replicated", "row_slice"):
            for inp in ("full", "replicate", "slice"):
                p = run(from_host=True, to_host=True, result=result, input=inp)
                assert p.h2d_bytes == (n_q if inp == "full" else hi - lo) * DIM * 4
                assert p.d2h_bytes == ((per if result == "row_slice" else n_q) * k * 12)
            run(from_host=False, to_host=False, result=result)
            run(from_host=True, to_host=False, result=result, input="slice", exclude=True)
            # per-shard fp32 re-scoring before the exchange
            p = run(from_host=True, to_host=True, result=result, input="slice", rescore_from=shard_f32)
            assert p.kc == min(k + 8, sb.hi - sb.lo)
            run(from_host=True, to_host=False, result=result, input="full", rescore_from=shard_f32,
                excludes_self=True, exclude=True)
        with pytest.raises(ValueError):
            SearchPipeline(sb, n_q, k, result="columns")
        with pytest.raises(ValueError):
            SearchPipeline(sb, n_q, k, rescore_from=bank)        # not this rank's shard
        # moving shard boundaries: every rank stores more than it searches; results do not move
        sa = ShardedRelatedBank(n_rows, DIM, local_bank_factory=StandInBank, overlap=0.5)
        sa.upload_global(bank)
        pipe = SearchPipeline(sa, n_q, k, depth=2, from_host=False, to_host=False, balance_every=2)
        assert pipe.balance_every == 2
        for step in range(7):
            slot = pipe.submit(batches[step % 3])
            pipe.wait_stream(slot)
            s, i = pipe.result_of(slot)
            ws, wi = oracle.cosine_topk(batches[step % 3], bank, k)
            assert torch.equal(i, wi) and torch.allclose(s, ws, atol=1e-6), step
        assert pipe.rebalances >= 2 and len(pipe.rows_log) == 7
        assert sa.bounds[0][0] == 0 and sa.bounds[-1][1] == n_rows
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_rows,n_q,k", [(3, 200, 10, 4), (4, 257, 9, 3), (4, 1000, 64, 10)])
def test_search_pipeline_layouts_gloo(world, n_rows, n_q, k):
    mp.spawn(_worker, args=(world, _free_port(), n_rows, n_q, k), nprocs=world, join=True)
