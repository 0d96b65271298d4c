"""Runs bench.py's OUR arm (run_ours, one rank) on the CPU with stand-ins for everything CUDA — the
bank (a torch emulation of what the kernel ranks: bf16-rounded operands, fp32 sums), SearchPipeline,
events, pinned memory — on shrunken workloads, so that every line of the measurement protocol and
of the JSON assembly executes without a GPU.  Not a measurement of anything: the numbers it prints
are CPU times of the stand-ins.  Run by tests/test_host_cpu.py::test_bench_our_arm_on_standins in a
process of its own (it patches torch.cuda).

    python tests/bench_standins.py [workload] [--world N]     -> the JSON line bench.py would print

--world N > 1 spawns N ranks over gloo: the real ShardedRelatedBank and the real SearchPipeline
(streams and events are no-ops: CPU work is already ordered) around the stand-in bank, i.e. the
sharded protocol of bench.py — per-rank gates, bit-identity across ranks and against one "GPU",
per-rank kernel statistics — exactly as torchrun would drive it.
"""
import argparse
import contextlib
import inspect
import os
import socket
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
import zsaac_b200  # noqa: E402
from zsaac_b200 import sharded  # noqa: E402

D = bench.D


class StandInBank:
    """zsaac_b200.RelatedBank as bench.py uses it."""

    def __init__(self, rows, dim, device=None, index_offset=0):
        self.rows, self.dim, self.index_offset = rows, dim, index_offset
        self.device = torch.device("cpu")
        self.bank = torch.zeros(rows, dim)
        self.launch_count = 0
        self._profiling, self._times = False, []

    def upload(self, rows, dst_row=0, normalize=True):
        r = torch.nn.functional.normalize(rows.float(), dim=-1) if normalize else rows.float()
        self.bank[dst_row:dst_row + r.shape[0]] = r.bfloat16().float()

    def reserve(self, n_queries, k):
        pass

    def window(self, row_lo=0, n_rows=0):
        self.win = (row_lo, n_rows) if n_rows else None

    def search(self, queries, k, self_index=None, normalize_queries=True, out=None):
        q = torch.nn.functional.normalize(queries.float(), dim=-1).bfloat16() if normalize_queries \
            else queries.bfloat16()
        lo, n = self.win if getattr(self, "win", None) else (0, self.rows)
        # float64 sums of bf16 products, rounded once: the bits do not depend on how the bank is cut
        # into shards (the property the kernel has by construction: same K-loop order)
        s = (q.double() @ self.bank[lo:lo + n].double().T).float()
        if self_index is not None:
            col = self_index - self.index_offset - lo
            ok = (col >= 0) & (col < n)
            s[torch.arange(q.shape[0])[ok], col[ok]] = float("-inf")
        top = torch.sort(s, dim=1, descending=True, stable=True)
        self.launch_count += 3
        if self._profiling:
            self._times.append(0.4 + 0.01 * (self.launch_count % 7))
        res = (top.values[:, :k].contiguous(), top.indices[:, :k].contiguous() + self.index_offset + lo)
        if out is not None:
            out[0].copy_(res[0])
            out[1].copy_(res[1])
            return out
        return res

    def merge(self, scores, indices, out=None):
        n_lists, n_q, k = scores.shape
        flat_s = scores.permute(1, 0, 2).reshape(n_q, n_lists * k)
        flat_i = indices.permute(1, 0, 2).reshape(n_q, n_lists * k)
        order = torch.sort(flat_i, dim=1, stable=True).indices                 # (score desc, index asc)
        flat_s, flat_i = flat_s.gather(1, order), flat_i.gather(1, order)
        top = torch.sort(flat_s, dim=1, descending=True, stable=True)
        res = (top.values[:, :k].contiguous(), flat_i.gather(1, top.indices)[:, :k].contiguous())
        self.launch_count += 1
        if out is not None:
            out[0].copy_(res[0])
            out[1].copy_(res[1])
            return out
        return res

    def profile(self, enable):
        self._profiling = enable
        if enable:
            self._times = []

    def kernel_times_ms(self):
        return list(self._times)

    def plan(self, n_queries, k):
        return (1, 2, 3)

    def close(self):
        pass


class NoStream:
    def wait_event(self, event):
        pass

    def wait_stream(self, stream):
        pass

    def synchronize(self):
        pass


class StandInEvent:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def synchronize(self):
        pass

    def elapsed_time(self, other):
        return max((other.t - self.t) * 1e3, 1e-3)


def install_standins():
    torch.cuda.is_available = lambda: True
    torch.cuda.set_device = lambda device: None
    torch.cuda.synchronize = lambda device=None: None
    torch.cuda.Event = StandInEvent
    torch.cuda.Stream = lambda device=None: NoStream()
    torch.cuda.stream = lambda stream: contextlib.nullcontext()
    torch.cuda.current_stream = lambda device=None: NoStream()
    torch.cuda._sleep = lambda cycles: None
    torch.cuda.empty_cache = lambda: None
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    zsaac_b200.RelatedBank = StandInBank
    sharded._default_local_bank = lambda rows, dim, device, index_offset: StandInBank(rows, dim, device, index_offset)
    bench.WORKLOADS = {                      # same keys, shapes a CPU ranks in milliseconds
        "clotho_eval": (200, 2000, 5, False, 101, 201),
        "audiocaps": (50, 3000, 10, False, 102, 202),
        "wavcaps_400k": (256, 5000, 10, False, 103, 203),
        "synthetic_10m": (300, 6000, 32, False, 104, 204),
        "allpairs_400k": (500, 500, 5, True, 105, 203),
    }
    bench.BANK_BLOCK = 2048
    bench.IDLE_BEFORE_S = 0.0
    bench.SHORT_REPS = 4
    # what cannot be patched from outside: the device run_ours picks and the NCCL backend
    src = inspect.getsource(bench.run_ours)
    assert src.count('device = torch.device("cuda", local_rank)') == 1
    assert src.count('dist.init_process_group("nccl", device_id=device)') == 1
    src = src.replace('device = torch.device("cuda", local_rank)', 'device = torch.device("cpu")')
    src = src.replace('dist.init_process_group("nccl", device_id=device)', 'dist.init_process_group("gloo")')
    src = src.replace('device="cuda"', 'device="cpu"')        # the literal loop "on the GPU as written"
    exec(compile(src, os.path.join(ROOT, "bench.py") + ":run_ours(stand-ins)", "exec"), bench.__dict__)


def run_rank(rank, world, port, workload, steps, balance):
    if world > 1:
        os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                          MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    else:
        for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
            os.environ.pop(k, None)
    install_standins()
    args = argparse.Namespace(gpus=world, steps=steps, warmup=3, impl="ours", workload=workload, queries=0,
                              bank_rows=0, no_cpu_baseline=False, balance=balance, headline_only=False)
    rc = bench.run_ours(args)
    if rc:
        raise SystemExit(rc)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", nargs="?", default=bench.DEFAULT_WORKLOAD)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--world", type=int, default=1)
    ap.add_argument("--balance", action="store_true")
    cli = ap.parse_args()
    if cli.world == 1:
        run_rank(0, 1, 0, cli.workload, cli.steps, False)
        return 0
    import torch.multiprocessing as mp
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    mp.spawn(run_rank, args=(cli.world, port, cli.workload, cli.steps, cli.balance), nprocs=cli.world, join=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
