"""Runs bench.py's OUR arm (run_ours, one rank) on the CPU with stand-ins for everything CUDA — the
bank (a torch emulation of what the kernel ranks: bf16-rounded operands, fp32 sums), SearchPipeline,
events, pinned memory — on shrunken workloads, so that every line of the measurement protocol and
of the JSON assembly executes without a GPU.  Not a measurement of anything: the numbers it prints
are CPU times of the stand-ins.  Run by tests/test_host_cpu.py::test_bench_our_arm_on_standins in a
process of its own (it patches torch.cuda).

    python tests/bench_standins.py [workload]      -> the JSON line bench.py would print
"""
import argparse
import inspect
import os
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

    def search(self, queries, k, self_index=None, normalize_queries=True, out=None):
        q = torch.nn.functional.normalize(queries.float(), dim=-1).bfloat16().float()
        s = q @ self.bank.T
        if self_index is not None:
            col = self_index - self.index_offset
            ok = (col >= 0) & (col < self.rows)
            s[torch.arange(q.shape[0])[ok], col[ok]] = float("-inf")
        top = torch.sort(s, dim=1, descending=True, stable=True)
        self.launch_count += 3
        if self._profiling:
            self._times.append(0.4)
        return top.values[:, :k].contiguous(), top.indices[:, :k].contiguous() + self.index_offset

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


class StandInPipeline:
    """zsaac_b200.sharded.SearchPipeline as bench.py uses it on one rank."""

    def __init__(self, bank, n_queries, k, depth=2, from_host=True, to_host=True, result="replicated",
                 self_index=None, input="full", balance_every=0):
        self.bank, self.k, self.self_index = bank, k, self_index
        self.h2d_bytes = n_queries * D * 4 if from_host else 0
        self.d2h_bytes = n_queries * k * 12 if to_host else 0
        self.rebalances, self.rows_log, self.out_rows = 0, [], (0, n_queries)
        self._results, self._submitted = {}, 0

    def submit(self, queries):
        slot = self._submitted % 2
        self._submitted += 1
        self._results[slot] = self.bank.search(queries, self.k, self_index=self.self_index)
        return slot

    def wait_stream(self, idx=None):
        pass

    def result_of(self, idx, host=False):
        return self._results[idx]


class StandInEvent:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max((other.t - self.t) * 1e3, 1e-3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", nargs="?", default=bench.DEFAULT_WORKLOAD)
    ap.add_argument("--steps", type=int, default=3)
    cli = ap.parse_args()
    torch.cuda.is_available = lambda: True
    torch.cuda.set_device = lambda device: None
    torch.cuda.synchronize = lambda device=None: None
    torch.cuda.Event = StandInEvent
    torch.cuda._sleep = lambda cycles: None
    torch.cuda.empty_cache = lambda: None
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    zsaac_b200.RelatedBank = StandInBank
    sharded.SearchPipeline = StandInPipeline
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
    # the one thing that cannot be patched from outside: the device run_ours picks
    src = inspect.getsource(bench.run_ours)
    assert src.count('device = torch.device("cuda", local_rank)') == 1
    src = src.replace('device = torch.device("cuda", local_rank)', 'device = torch.device("cpu")')
    src = src.replace('device="cuda"', 'device="cpu"')        # the literal loop "on the GPU as written"
    exec(compile(src, os.path.join(ROOT, "bench.py") + ":run_ours(stand-ins)", "exec"), bench.__dict__)
    args = argparse.Namespace(gpus=1, steps=cli.steps, warmup=3, impl="ours", workload=cli.workload, queries=0,
                              bank_rows=0, no_cpu_baseline=False, balance=False, headline_only=False)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        os.environ.pop(k, None)
    return bench.run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
