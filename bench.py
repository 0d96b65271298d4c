"""bench.py — related-caption retrieval throughput (queries/s) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A step is ONE pass of the hot path over one batch of synthetic queries: normalise + cast ->
fused cosine-similarity/top-k against the bank (row-sharded over the N ranks) -> all-gather of the
shard-local top-k -> k-way merge.  Default workload: BASELINE.json config 4 (65,536 queries vs a
10 M-row bf16 bank, d=1024, top-32), the configuration its metric ("... at 1/2/4/8 B200") is
quoted on; the bank is fixed as N grows (strong scaling).  Rank 0 prints ONE JSON line.

--impl reference times the reference's CPU formulation of the same path (torch fp32:
F.normalize(q) @ bank.T -> topk, all host threads) on a bounded sample of the workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D = 1024
WORKLOADS = {
    # name: (Q, N, k, exclude_self, q_seed, bank_seed)           BASELINE.json configs[i]
    "clotho_eval": (1045, 19195, 5, False, 101, 201),            # configs[0]
    "audiocaps": (975, 49838, 10, False, 102, 202),              # configs[1]
    "wavcaps_400k": (8192, 400_000, 10, False, 103, 203),        # configs[2]
    "synthetic_10m": (65536, 10_000_000, 32, False, 104, 204),   # configs[3]
    "allpairs_400k": (400_000, 400_000, 5, True, 105, 203),      # configs[4]
}
DEFAULT_WORKLOAD = "synthetic_10m"
BANK_BLOCK = 65536          # rows per generation block; seed = bank_seed * 2**32 + block id
L2_BYTES = 126 << 20        # B200 L2 capacity
REBALANCE_DEFAULT = False
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p["bf16_tflops"]),
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"]))}, "measured"
    except Exception:
        return dict(FALLBACK_PEAKS), "fallback"


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU every 100 ms while the timed region runs."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None

    _REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
                0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def _poll(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
                try:
                    mask = self.nvml.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    mask = self.nvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for bit, name in self._REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nvml is not None:
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)

    def report(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(statistics.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------
def gen_bank_block(torch, device, bank_seed: int, block: int, rows: int):
    g = torch.Generator(device=device).manual_seed(bank_seed * (2 ** 32) + block)
    return torch.randn(rows, D, device=device, generator=g)


def gen_queries(torch, Q: int, seed: int, device="cpu"):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(Q, D, generator=g) if device == "cpu" else torch.randn(Q, D, generator=g).to(device)


def fill_shard(torch, bank, lo: int, hi: int, bank_seed: int, device):
    """Generate global bank rows [lo, hi) block by block (identical for every world size)."""
    row = lo
    while row < hi:
        block = row // BANK_BLOCK
        b_lo = block * BANK_BLOCK
        rows = gen_bank_block(torch, device, bank_seed, block, BANK_BLOCK)
        take_lo = row - b_lo
        take_hi = min(hi - b_lo, BANK_BLOCK)
        bank.upload(rows[take_lo:take_hi], row - lo, normalize=True)
        row = b_lo + take_hi
    torch.cuda.synchronize(device)


class CpuReference:
    """The reference's CPU formulation on a bounded sample of the workload (built once, timed
    per pass).  The sample is a slice of the same synthetic data; queries/s is scaled linearly in
    the bank rows left out (the cost per query is linear in N)."""

    def __init__(self, torch, name: str):
        from oracle import oracle
        self.torch, self.oracle, self.name = torch, oracle, name
        Q, N, k, _, q_seed, bank_seed = WORKLOADS[name]
        self.N, self.k = N, k
        self.n_s = min(N, 1_000_000)
        self.q_s = min(Q, 1024)          # ~0.5 s (400 k rows) to ~1.7 s (1 M rows) of 16-thread work per pass
        bank = torch.empty(self.n_s, D)
        for blk in range(-(-self.n_s // BANK_BLOCK)):
            lo = blk * BANK_BLOCK
            hi = min(lo + BANK_BLOCK, self.n_s)
            g = torch.Generator().manual_seed(bank_seed * (2 ** 32) + blk)
            bank[lo:hi] = torch.randn(hi - lo, D, generator=g)
        self.bank = oracle.normalize_rows(bank)
        self.queries = gen_queries(torch, self.q_s, q_seed)
        oracle.fast_topk(self.queries[:64], self.bank, k)          # thread pool / page warm-up

    def one_pass(self) -> float:
        t0 = time.perf_counter()
        self.oracle.fast_topk(self.queries, self.bank, self.k)
        return time.perf_counter() - t0

    def baseline(self, dt: float) -> dict:
        scale = self.n_s / self.N
        return {
            "value": self.q_s / dt * scale, "unit": "queries/s",
            "cores": int(self.torch.get_num_threads()), "kind": "port",
            "sample": (f"torch CPU fp32 F.normalize(q) @ bank.T -> topk({self.k}) on {self.q_s} "
                       f"queries x {self.n_s} bank rows, {dt * 1e3:.0f} ms per pass"
                       + (f"; queries/s scaled by {scale:.3g} to the {self.N}-row bank (cost is "
                          "linear in N)" if scale != 1 else "")
                       + f"; host has {os.cpu_count()} logical cpus"),
        }


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0                                                  # rank 0 alone runs this arm
    import torch
    # torchrun exports OMP_NUM_THREADS=1 for multi-rank launches; this arm is the CPU reference
    # and is entitled to every host thread
    torch.set_num_threads(os.cpu_count() or 1)
    name = args.workload
    Q, N, k, excl, _, _ = WORKLOADS[name]
    ref = CpuReference(torch, name)
    times = []
    for step in range(args.warmup + args.steps):
        dt = ref.one_pass()
        if step >= args.warmup:
            times.append(dt)
    dt = statistics.mean(times)
    cb = ref.baseline(dt)
    value = cb["value"]
    q_s, n_s = ref.q_s, ref.n_s
    line = {
        "impl": "reference", "metric": "related-caption retrieval queries/s (top-k, d=1024)",
        "value": value, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{name}: {Q} queries vs {N}-row bank, d={D}, top-{k}"
                               + (", self-exclusion" if excl else ""),
                   "device": "host CPU", "sample_queries": q_s, "sample_bank_rows": n_s},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist
    import zsaac_b200
    from zsaac_b200.sharded import ShardedRelatedBank

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs a torchrun launch with {args.gpus} ranks")
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    zsaac_b200.load_library()
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    name = args.workload
    Q, N, k, excl, q_seed, bank_seed = WORKLOADS[name]
    if args.queries:
        Q = args.queries
    if args.bank_rows:
        N = args.bank_rows
    peaks, peaks_src = load_peaks()

    # ---- bank: resident in HBM as bf16 before anything is timed
    if world > 1:
        bank = ShardedRelatedBank(N, D, device=device)
        lo, hi = bank.lo, bank.hi
        local = bank.local
    else:
        local = zsaac_b200.RelatedBank(N, D, device=device)
        bank = local
        lo, hi = 0, N
    fill_shard(torch, local, lo, hi, bank_seed, device)

    # ---- N > 1: size the shards by the measured speed of each GPU (untimed calibration).  Every
    # step ends in an all-gather, so the slowest GPU sets the pace, and the GPUs of one box differ
    # by several per cent under the power cap (kernel_ms_per_rank).  Two calibration searches on
    # equal shards give rows/ms per rank; if the ranks differ by more than 2 % the bank is
    # re-sharded in proportion (ShardedRelatedBank(shard_weights=...)) and refilled.
    shard_weights = None
    if world > 1 and args.rebalance:
        qc = gen_queries(torch, Q, q_seed).to(device)
        local.reserve(Q, k)
        for _ in range(2):
            bank.search(qc, k)
        torch.cuda.synchronize(device)
        local.profile(True)
        for _ in range(2):
            bank.search(qc, k)
        torch.cuda.synchronize(device)
        t_cal = statistics.mean(local.kernel_times_ms())
        local.profile(False)
        speeds = torch.zeros(world, device=device, dtype=torch.float64)
        speeds[rank] = (hi - lo) / t_cal
        dist.all_reduce(speeds, op=dist.ReduceOp.SUM)
        speeds = speeds.tolist()
        if max(speeds) / min(speeds) > 1.02:
            shard_weights = [round(v / max(speeds), 4) for v in speeds]
            local.close()
            del bank, local
            torch.cuda.empty_cache()
            bank = ShardedRelatedBank(N, D, device=device, shard_weights=shard_weights)
            lo, hi = bank.lo, bank.hi
            local = bank.local
            fill_shard(torch, local, lo, hi, bank_seed, device)
        del qc

    if excl:
        # BASELINE config 5: queries are the bank rows themselves after the reference's
        # noise_injection (utils.py:19-31: normalise, add N(0, 0.001 I), renormalise in search)
        q_dev = torch.empty(Q, D, device=device)
        gq = torch.Generator(device=device).manual_seed(q_seed)
        for blk in range(-(-Q // BANK_BLOCK)):
            r0, r1 = blk * BANK_BLOCK, min((blk + 1) * BANK_BLOCK, Q)
            rows = torch.nn.functional.normalize(gen_bank_block(torch, device, bank_seed, blk, BANK_BLOCK)[:r1 - r0], dim=-1)
            q_dev[r0:r1] = rows + torch.randn(r1 - r0, D, device=device, generator=gq) * (0.001 ** 0.5)
        q_host = q_dev.cpu().pin_memory()
    else:
        q_host = gen_queries(torch, Q, q_seed).pin_memory()
        q_dev = q_host.to(device, non_blocking=True)
    self_index = torch.arange(Q, dtype=torch.int64, device=device) if excl else None
    local.reserve(Q, k)
    out_host_s = torch.empty(Q, k, dtype=torch.float32).pin_memory()
    out_host_i = torch.empty(Q, k, dtype=torch.int64).pin_memory()

    def step_device():
        return bank.search(q_dev, k, self_index=self_index)

    def step_e2e():
        # N > 1: every rank uploads 1/N of the host batch and the slices are all-gathered over
        # NVLink (ShardedRelatedBank.replicate_from_host) instead of N full PCIe copies
        qd = bank.replicate_from_host(q_host) if world > 1 else q_host.to(device, non_blocking=True)
        s, i = bank.search(qd, k, self_index=self_index)
        out_host_s.copy_(s, non_blocking=True)
        out_host_i.copy_(i, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    # Timing rule: either the inputs of a step exceed the L2 (126 MB) or the L2 is flushed between
    # timed steps.  Small workloads (configs 1-2: the bf16 bank alone fits in L2) take the second
    # route: every step is bracketed by its own event pair and a 256 MB write runs in between.
    shard_bytes = (hi - lo) * D * 2 + Q * D * 4
    flush_l2 = shard_bytes < 2 * L2_BYTES
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=device) if flush_l2 else None

    def timed(fn, steps):
        barrier()
        if not flush_l2:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            barrier()
            total = e0.elapsed_time(e1)
        else:
            pairs = []
            for _ in range(steps):
                flush_buf.zero_()                     # evicts bank, queries and outputs from L2
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                pairs.append((e0, e1))
            barrier()
            total = sum(a.elapsed_time(b) for a, b in pairs)
        ms = torch.tensor([total], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(max(args.warmup, 3)):
        step_device()
    torch.cuda.synchronize(device)

    # ---- timed region 1: inputs resident in HBM (value) -------------------------------------
    launches0 = local.launch_count
    local.profile(True)
    with ClockSampler(local_rank) as clocks:
        total_ms = timed(step_device, args.steps)
    kernel_ms = local.kernel_times_ms()
    local.profile(False)
    launches = local.launch_count - launches0
    ms_per_step = total_ms / args.steps
    value = Q / (ms_per_step * 1e-3)

    # ---- timed region 2: host buffers in, host buffers out (e2e) -----------------------------
    for _ in range(2):
        step_e2e()
    e2e_ms = timed(step_e2e, args.steps) / args.steps
    e2e_value = Q / (e2e_ms * 1e-3)

    # ---- roofline of the dominant kernel (this rank's shard) ---------------------------------
    shard_rows = hi - lo
    flop = 2.0 * Q * shard_rows * D
    k_ms = statistics.mean(kernel_ms) if kernel_ms else float("nan")
    achieved_tf = flop / (k_ms * 1e-3) / 1e12
    ai = Q                                           # flop per bank byte ~ Q (bank dominates bytes)
    ridge = peaks["bf16_tflops"] * 1e12 / (peaks["hbm_gbs"] * 1e9)
    # which measured peak applies: the sustained one when the timed region ran power-capped (long
    # back-to-back tensor work), the burst one for short kernels at full clocks
    clock_report = clocks.report()
    capped = ("sw_power_cap" in clock_report["reasons"] and clock_report["sm_mhz"] is not None
              and clock_report["sm_max_mhz"] and clock_report["sm_mhz"] < 0.85 * clock_report["sm_max_mhz"])
    long_step = capped or total_ms > 1000.0
    if ai >= ridge:
        peak = peaks["bf16_tflops_sustained"] if long_step else peaks["bf16_tflops"]
        roofline = {"bound": "tensor", "achieved": achieved_tf, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved_tf / peak,
                    "peak_kind": f"{'sustained' if long_step else 'burst'} bf16, {peaks_src}",
                    "frac_of_burst": achieved_tf / peaks["bf16_tflops"]}
    else:
        bytes_alg = 2.0 * shard_rows * D + Q * D * 2.0 + Q * k * 12.0
        gbs = bytes_alg / (k_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": gbs / peaks["hbm_gbs"], "peak_kind": f"copy bandwidth, {peaks_src}"}
    roofline["kernel"] = "zs_simtopk_kernel"
    roofline["kernel_ms"] = k_ms
    if world > 1:
        # every step ends in an all-gather, so the slowest rank's kernel sets the step time
        per_rank = torch.zeros(world, device=device)
        per_rank[rank] = k_ms
        dist.all_reduce(per_rank, op=dist.ReduceOp.SUM)
        roofline["kernel_ms_per_rank"] = [round(v, 3) for v in per_rank.tolist()]
    roofline["kernel_share_of_step"] = k_ms / ms_per_step
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    roofline["traffic"] = None
    try:
        if world == 1:          # captured with ncu on one GPU holding the whole bank
            with open(traffic_path) as f:
                roofline["traffic"] = json.load(f).get(name)
    except Exception:
        pass

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        ref = CpuReference(torch, name)
        cpu_baseline = ref.baseline(min(ref.one_pass() for _ in range(3)))

    if rank == 0:
        plan = local.plan(Q, k)
        line = {
            "metric": "related-caption retrieval queries/s (top-k, d=1024)",
            "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {
                "workload": f"{name}: {Q} queries vs {N}-row bank, d={D}, top-{k}"
                            + (", self-exclusion" if excl else ""),
                "parallelism": f"bank row-sharded over {world} GPU(s), queries replicated, "
                               "all-gather + k-way merge" if world > 1 else "single GPU",
                "bank_rows_per_gpu": shard_rows, "shard_weights": shard_weights,
                "bank_dtype": "bf16", "accumulate": "fp32",
                "l2": ("L2 flushed between timed steps (256 MB write); bank shard %.0f MB + queries "
                       "%.0f MB per step" % (shard_rows * D * 2 / 1e6, Q * D * 4 / 1e6)) if flush_l2
                      else ("inputs larger than L2: bank shard %.1f GB + queries %.0f MB per step"
                            % (shard_rows * D * 2 / 1e9, Q * D * 4 / 1e6)),
                "plan_chunks_tiles_ctas": list(plan),
                "tflops": 2.0 * Q * N * D / (ms_per_step * 1e-3) / 1e12,
            },
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": "queries/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": Q * D * 4, "d2h_bytes_per_step": world * Q * k * 12},
            "gpu_launches": int(launches) * world,
            "clocks": clock_report,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=list(WORKLOADS), default=DEFAULT_WORKLOAD)
    ap.add_argument("--queries", type=int, default=0, help="override the query count (smoke runs)")
    ap.add_argument("--bank-rows", type=int, default=0, help="override the bank rows (smoke runs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--rebalance", action=argparse.BooleanOptionalAction, default=REBALANCE_DEFAULT,
                    help="N > 1: size the bank shards by the measured speed of each GPU")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
