"""bench.py — related-caption retrieval throughput (queries/s) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A step is ONE pass of the hot path over one batch of synthetic queries: normalise + cast ->
fused cosine-similarity/top-k against the bank (row-sharded over the N ranks) -> exchange of the
shard-local top-k -> k-way merge.  Headline workload: BASELINE.json config 4 (65,536 queries vs a
10 M-row bf16 bank, d=1024, top-32), the configuration its metric ("... at 1/2/4/8 B200") is
quoted on; the bank is fixed as N grows (strong scaling).  Rank 0 prints ONE JSON line.

Before anything is timed the run verifies itself (`parity_gate`, SURVEY §8d "correctness gate"):
sampled queries are re-scored in fp32 against the whole (sharded) bank and north_star's rule is
applied (|score - fp32| <= 1e-3; every returned index within 1e-3 of the fp32 k-th score; every
clear fp32 winner returned); at N > 1 the merged result must be bit-identical on every rank and
bit-identical to a single-GPU search of a 4,096-query slice.  A failed gate aborts non-zero.
The other named shapes are timed as well (`workloads`): BASELINE configs 1, 2, 3, 5 and the
HBM-bound batches Q in {1, 32, 128} against the 400 k bank BEFORE the headline bank is built (they
are searches of 50 us - 4 ms; right after the headline's seconds at the power cap they would be
measured at the capped clock), Q in {1, 32, 128} against the resident 10 M bank after it — each
with its own gate and roofline, next to the unfused library strawman (torch.matmul bf16 + topk)
and the reference's literal per-item loop.  Sub-millisecond searches are repeated 50 times with
the stream parked behind a spin kernel first, so the events bracket device time, not the host's
launch latency; shapes of less than 50 ms per search start both of their passes (whole search,
fused kernel alone) after 1 s of idle, i.e. from the clock state a stand-alone run would see.

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
SMALL_Q = (1, 32, 128)      # HBM-bound batches, k = 10
BANK_BLOCK = 65536          # rows per generation block; seed = bank_seed * 2**32 + block id
L2_BYTES = 126 << 20        # B200 L2 capacity
GATE_SAMPLES = 64
SHORT_REPS = 50             # repetitions of a side shape whose search takes < 2 ms
BALANCE_OVERLAP = 0.125     # N > 1: every rank stores this fraction of a shard beyond either end of its own
BALANCE_EVERY = 2           # ... and the boundaries are re-balanced before every 2nd step
IDLE_BEFORE_S = 1.0         # pause before a side shape of < 50 ms per search is timed (a power-capped clock recovers)
GATE_TOL = 1e-3             # north_star: scores within 1e-3 of fp32, index sets equal except near-ties
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

    def now(self):
        """One reading of the SM clock (right after a short timed region, before the GPU idles down)."""
        try:
            return float(self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
        except Exception:
            return None

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


def bank_rows_fp32(torch, device, bank_seed: int, lo: int, hi: int):
    """Yield (first global row, fp32 rows) covering global bank rows [lo, hi) block by block —
    identical for every world size (per-block seeds)."""
    row = lo
    while row < hi:
        block = row // BANK_BLOCK
        b_lo = block * BANK_BLOCK
        rows = gen_bank_block(torch, device, bank_seed, block, BANK_BLOCK)
        take_lo = row - b_lo
        take_hi = min(hi - b_lo, BANK_BLOCK)
        yield row, rows[take_lo:take_hi]
        row = b_lo + take_hi


def fill_shard(torch, bank, lo: int, hi: int, bank_seed: int, device):
    for row, rows in bank_rows_fp32(torch, device, bank_seed, lo, hi):
        bank.upload(rows, row - lo, normalize=True)
    torch.cuda.synchronize(device)


def make_queries(torch, name: str, Q: int, device):
    """(pinned host queries, device queries, self_index) of a workload."""
    _, _, _, excl, q_seed, bank_seed = WORKLOADS[name]
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
        self_index = torch.arange(Q, dtype=torch.int64, device=device)
    else:
        q_host = gen_queries(torch, Q, q_seed).pin_memory()
        q_dev = q_host.to(device, non_blocking=True)
        self_index = None
    return q_host, q_dev, self_index


# ------------------------------------------------------------------------------------------------
def parity_gate(torch, dist, world, device, lo, hi, bank_seed, q_dev, self_index, k, result):
    """fp32 re-scoring of sampled queries against this rank's shard (regenerated block by block
    from the same seeds), reduced over ranks, compared with the merged result by north_star's rule."""
    Q = q_dev.shape[0]
    sel = torch.unique(torch.linspace(0, Q - 1, min(Q, GATE_SAMPLES), device=device).round().long())
    ns = sel.numel()
    got_s, got_i = result[0][sel], result[1][sel]
    qn = torch.nn.functional.normalize(q_dev[sel].float(), dim=-1)
    self_sel = self_index[sel] if self_index is not None else None
    top_s = torch.full((ns, k), float("-inf"), device=device)
    top_i = torch.full((ns, k), -1, dtype=torch.int64, device=device)
    ret_fp32 = torch.full((ns, k), float("-inf"), device=device)
    for row0, rows in bank_rows_fp32(torch, device, bank_seed, lo, hi):
        s = qn @ torch.nn.functional.normalize(rows, dim=-1).T           # fp32 (TF32 off by default)
        n = rows.shape[0]
        inside = (got_i >= row0) & (got_i < row0 + n)
        ret_fp32 = torch.where(inside, s.gather(1, (got_i - row0).clamp(0, n - 1)), ret_fp32)
        if self_sel is not None:
            mine = (self_sel >= row0) & (self_sel < row0 + n)
            s[mine.nonzero().squeeze(1), (self_sel[mine] - row0)] = float("-inf")
        bs, bi = torch.topk(s, min(k, n), dim=1)
        cat_s, cat_i = torch.cat([top_s, bs], dim=1), torch.cat([top_i, bi + row0], dim=1)
        top_s, order = torch.topk(cat_s, k, dim=1)
        top_i = cat_i.gather(1, order)
    if world > 1:
        all_s = [torch.empty_like(top_s) for _ in range(world)]
        all_i = [torch.empty_like(top_i) for _ in range(world)]
        dist.all_gather(all_s, top_s)
        dist.all_gather(all_i, top_i)
        cat_s, cat_i = torch.cat(all_s, dim=1), torch.cat(all_i, dim=1)
        top_s, order = torch.topk(cat_s, k, dim=1)
        top_i = cat_i.gather(1, order)
        dist.all_reduce(ret_fp32, op=dist.ReduceOp.MAX)
    kth = top_s[:, k - 1:k]
    score_err = (got_s - ret_fp32).abs().max().item()
    below = int((ret_fp32 < kth - GATE_TOL).sum().item())               # returned, but not a near-winner
    clear = top_s > kth + GATE_TOL                                      # must be returned
    present = (top_i.unsqueeze(2) == got_i.unsqueeze(1)).any(dim=2)
    missing = int((clear & ~present).sum().item())
    ordered = bool((got_s[:, :-1] >= got_s[:, 1:]).all().item()) if k > 1 else True
    no_self = True if self_sel is None else not bool((got_i == self_sel[:, None]).any().item())
    exact_match = float((got_i == top_i).float().mean().item())
    ok = score_err <= GATE_TOL and below == 0 and missing == 0 and ordered and no_self
    return {"sampled_queries": ns, "max_abs_score_err_vs_fp32": score_err, "returned_below_band": below,
            "clear_winners_missing": missing, "sorted": ordered, "self_excluded": no_self,
            "index_agreement_with_fp32": exact_match, "ok": bool(ok)}


def rank_consistency_gate(torch, dist, zs, world, rank, device, N, bank_seed, q_dev, self_index, k, result,
                          single_gpu=True):
    """N > 1: the merged result must be the same bits on every rank, and the same bits a single
    GPU holding the whole bank returns for a 4,096-query slice."""
    out = {}
    ref_s, ref_i = result[0].clone(), result[1].clone()
    dist.broadcast(ref_s, 0)
    dist.broadcast(ref_i, 0)
    same = torch.tensor([int(torch.equal(ref_s, result[0]) and torch.equal(ref_i, result[1]))], device=device)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    out["identical_on_all_ranks"] = bool(same.item())
    if not single_gpu:
        out["ok"] = out["identical_on_all_ranks"]
        return out
    flag = torch.ones(1, device=device, dtype=torch.int64)
    n_slice = min(4096, q_dev.shape[0])
    if rank == 0:
        whole = zs.RelatedBank(N, D, device=device)
        fill_shard(torch, whole, 0, N, bank_seed, device)
        si = self_index[:n_slice] if self_index is not None else None
        s1, i1 = whole.search(q_dev[:n_slice], k, self_index=si)
        torch.cuda.synchronize(device)
        flag[0] = int(torch.equal(s1, result[0][:n_slice]) and torch.equal(i1, result[1][:n_slice]))
        whole.close()
        del whole
        torch.cuda.empty_cache()
    dist.broadcast(flag, 0)
    out["single_gpu_slice_queries"] = n_slice
    out["bit_identical_to_single_gpu"] = bool(flag.item())
    out["ok"] = out["identical_on_all_ranks"] and out["bit_identical_to_single_gpu"]
    return out


# ------------------------------------------------------------------------------------------------
def roofline_of(peaks, peaks_src, Q, shard_rows, k, k_ms, long_step):
    flop = 2.0 * Q * shard_rows * D
    bytes_alg = 2.0 * shard_rows * D + Q * D * 2.0 + Q * k * 12.0
    t_tensor_burst = flop / (peaks["bf16_tflops"] * 1e12)
    t_hbm = bytes_alg / (peaks["hbm_gbs"] * 1e9)
    if t_tensor_burst >= t_hbm:
        peak = peaks["bf16_tflops_sustained"] if long_step else peaks["bf16_tflops"]
        achieved = flop / (k_ms * 1e-3) / 1e12
        r = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
             "peak_kind": f"{'sustained' if long_step else 'burst'} bf16, {peaks_src}",
             "frac_of_burst": achieved / peaks["bf16_tflops"]}
    else:
        gbs = bytes_alg / (k_ms * 1e-3) / 1e9
        r = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
             "frac": gbs / peaks["hbm_gbs"], "peak_kind": f"copy bandwidth, {peaks_src}"}
    r["kernel"] = "zs_simtopk_kernel"
    r["kernel_ms"] = k_ms
    return r


class Timer:
    """Event-timed steps: back to back when a step's inputs exceed L2, else an event pair per step
    with a 256 MB write in between (L2 flush).  Max over ranks.

    `ahead_ms` > 0 (sub-millisecond shapes) first parks the stream behind a spin kernel of that
    length, so that every launch of the timed region is already queued when the device gets to it:
    the events then bracket device time only, not the host's launch latency (a search of 50-150 us
    is shorter than the Python + ctypes call that enqueues it)."""

    def __init__(self, torch, dist, world, device):
        self.torch, self.dist, self.world, self.device = torch, dist, world, device
        self.flush_buf = None
        self.samples = []          # per-step ms of the last flushed run (this rank)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.device)

    def park(self, ahead_ms):
        if ahead_ms > 0:
            self.torch.cuda._sleep(int(ahead_ms * 1.9e6))      # cycles at <= 1.9 GHz: at least ahead_ms

    def run(self, fn, steps, flush_l2, finish=None, ahead_ms=0.0):
        torch = self.torch
        self.barrier()
        self.samples = []
        if not flush_l2:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.park(ahead_ms)
            e0.record()
            for _ in range(steps):
                fn()
            if finish is not None:
                finish()
            e1.record()
            self.barrier()
            total = e0.elapsed_time(e1)
        else:
            if self.flush_buf is None:
                self.flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=self.device)
            pairs = []
            self.park(ahead_ms)
            for _ in range(steps):
                self.flush_buf.zero_()                 # evicts bank, queries and outputs from L2
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                if finish is not None:
                    finish()
                e1.record()
                pairs.append((e0, e1))
            self.barrier()
            self.samples = [a.elapsed_time(b) for a, b in pairs]
            total = sum(self.samples)
        ms = torch.tensor([total], device=self.device)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return ms.item()


def strawman_topk(torch, q_bf16, bank_bf16, k, q_chunk=1024, n_chunk=1_000_000):
    """The best unfused library formulation on the same GPU: torch.matmul (bf16 in, fp32 out) +
    torch.topk, chunked so that the score block fits."""
    out_s, out_i = [], []
    for q0 in range(0, q_bf16.shape[0], q_chunk):
        qc = q_bf16[q0:q0 + q_chunk]
        best_s = best_i = None
        for n0 in range(0, bank_bf16.shape[0], n_chunk):
            s = torch.matmul(qc, bank_bf16[n0:n0 + n_chunk].T).float()
            ts, ti = torch.topk(s, min(k, s.shape[1]), dim=1)
            ti = ti + n0
            if best_s is None:
                best_s, best_i = ts, ti
            else:
                cs, ci = torch.cat([best_s, ts], 1), torch.cat([best_i, ti], 1)
                best_s, order = torch.topk(cs, k, dim=1)
                best_i = ci.gather(1, order)
        out_s.append(best_s)
        out_i.append(best_i)
    return torch.cat(out_s), torch.cat(out_i)


def workload_config(name: str, Q: int, N: int, k: int, excl: bool, world: int, balance: bool = False) -> dict:
    """`config` of the JSON line: what is measured, as a function of the command line only, so that
    both arms (ours, --impl reference) of the same command print the SAME dict.  Everything that
    depends on the run (launch plan, achieved TFLOP/s, moved shard boundaries) is under `details`."""
    rows_per_gpu = shard_rows_of(N, world)
    shard_bytes = rows_per_gpu * D * 2 + Q * D * 4
    if shard_bytes < 2 * L2_BYTES:
        l2 = ("L2 flushed between timed steps (256 MB write); bank shard %.0f MB + queries %.0f MB per step"
              % (rows_per_gpu * D * 2 / 1e6, Q * D * 4 / 1e6))
    else:
        l2 = ("inputs larger than L2: bank shard %.1f GB + queries %.0f MB per step"
              % (rows_per_gpu * D * 2 / 1e9, Q * D * 4 / 1e6))
    return {
        "workload": f"{name}: {Q} queries vs {N}-row bank, d={D}, top-{k}" + (", self-exclusion" if excl else ""),
        "parallelism": (f"bank row-sharded over {world} GPUs, queries replicated, one NCCL all-gather "
                        "+ k-way merge per step") if world > 1 else "single GPU",
        "bank_rows_per_gpu": rows_per_gpu,
        "shard_balance": ({"stored_overlap_of_a_shard": BALANCE_OVERLAP, "every_steps": BALANCE_EVERY}
                          if (balance and world > 1) else None),
        "bank_dtype": "bf16", "accumulate": "fp32", "l2": l2,
    }


def shard_rows_of(N: int, world: int) -> int:
    """Rows of the largest shard of the equal split (sharded.shard_bounds: ceil(N / world) each)."""
    return -(-N // world)


class CpuReference:
    """The reference's CPU formulation on a bounded sample of the workload (built once, timed
    per pass).  The sample is a slice of the same synthetic data; queries/s is scaled linearly in
    the bank rows left out (the cost per query is linear in N)."""

    def __init__(self, torch, name: str, Q: int = 0, N: int = 0):
        from oracle import oracle
        self.torch, self.oracle, self.name = torch, oracle, name
        wQ, wN, k, _, q_seed, bank_seed = WORKLOADS[name]
        Q, N = Q or wQ, N or wN                                     # (--queries / --bank-rows overrides)
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

    def literal_loop(self, n_items: int) -> dict:
        """The reference's process_data exactly as written (one query per iteration,
        embeddings_related_generator.py:19-28; oracle.process_data_literal) on `n_items` items."""
        items = [{"text_embedding": self.queries[i:i + 1].clone()} for i in range(n_items)]
        t0 = time.perf_counter()
        n = sum(1 for _ in self.oracle.process_data_literal(self.bank, items, self.k))
        dt = time.perf_counter() - t0
        return {"value": n / dt, "unit": "queries/s", "kind": "port", "cores": int(self.torch.get_num_threads()),
                "sample": f"literal per-item loop (normalize -> cosine_similarity -> topk({self.k}) -> gather) "
                          f"on {n} items x {self.n_s} bank rows, {dt:.2f} s"}

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
    Q, N = args.queries or Q, args.bank_rows or N
    ref = CpuReference(torch, name, Q, N)
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
        # the same `config` our arm prints for this command line (the workload both arms measure);
        # what this arm actually ran — a bounded sample of it on the host — is in `reference_sample`
        "config": workload_config(name, Q, N, k, excl, max(1, args.gpus), args.balance),
        "reference_sample": {"device": "host CPU", "threads": cb["cores"], "sample_queries": q_s,
                             "sample_bank_rows": n_s,
                             "scaled_to_bank_rows": N if n_s != N else None},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


class Workload:
    """One named shape, bank resident (row-sharded at N > 1), queries resident and pinned."""

    def __init__(self, env, name, Q=None, N=None, k=None, bank=None, label=None, overlap=0.0):
        torch, zs = env["torch"], env["zs"]
        wQ, wN, wk, excl, _, bank_seed = WORKLOADS[name]
        self.env, self.name = env, name
        self.Q, self.N, self.k = Q or wQ, N or wN, k or wk
        self.excl, self.bank_seed = excl, bank_seed
        self.label = label or name
        world, device = env["world"], env["device"]
        if bank is not None:
            self.bank, self.owns_bank = bank, False
        else:
            self.owns_bank = True
            if world > 1:
                from zsaac_b200.sharded import ShardedRelatedBank
                self.bank = ShardedRelatedBank(self.N, D, device=device, overlap=overlap)
            else:
                self.bank = zs.RelatedBank(self.N, D, device=device)
        self.local = self.bank.local if world > 1 else self.bank
        # [lo, hi): this rank's share of the bank in the equal split (what the gate re-scores and
        # the roofline counts); with overlap the rank STORES more ([store_lo, store_hi)) and the
        # rows it searches move with the measured speed of the GPUs (sharded.py)
        self.lo, self.hi = self.bank.base_bounds[env["rank"]] if world > 1 else (0, self.N)
        if self.owns_bank:
            s_lo, s_hi = (self.bank.store_lo, self.bank.store_hi) if world > 1 else (0, self.N)
            for row, rows in bank_rows_fp32(torch, device, bank_seed, s_lo, s_hi):
                self.local.upload(rows, row - s_lo, normalize=True)
            torch.cuda.synchronize(device)
        self.q_host, self.q_dev, self.self_index = make_queries(torch, name, self.Q, device)
        self.local.reserve(self.Q, self.k)
        shard_bytes = (self.hi - self.lo) * D * 2 + self.Q * D * 4
        self.flush_l2 = shard_bytes < 2 * L2_BYTES

    def search(self):
        return self.bank.search(self.q_dev, self.k, self_index=self.self_index)

    def gate(self):
        env = self.env
        torch, dist = env["torch"], env["dist"]
        res = self.search()
        torch.cuda.synchronize(env["device"])
        self.gate_result = res
        g = parity_gate(torch, dist, env["world"], env["device"], self.lo, self.hi, self.bank_seed,
                        self.q_dev, self.self_index, self.k, res)
        if env["world"] > 1:
            # (the single-GPU comparison needs the whole bank on rank 0: done once per bank, by its owner)
            g["ranks"] = rank_consistency_gate(torch, dist, env["zs"], env["world"], env["rank"], env["device"],
                                               self.N, self.bank_seed, self.q_dev, self.self_index,
                                               self.k, res, single_gpu=self.owns_bank)
            g["ok"] = bool(g["ok"] and g["ranks"]["ok"])
        return g

    def close(self):
        if self.owns_bank:
            self.local.close()


def describe(w: Workload) -> str:
    return (f"{w.label}: {w.Q} queries vs {w.N}-row bank, d={D}, top-{w.k}"
            + (", self-exclusion" if w.excl else ""))


def time_side_workload(env, w: Workload, steps: int, when: str):
    """Gate + device-resident timing of one of the extra shapes -> dict for `workloads`.
    Sub-millisecond searches are repeated SHORT_REPS times behind a parked stream (Timer.run)."""
    torch, timer, peaks, peaks_src = env["torch"], env["timer"], env["peaks"], env["peaks_src"]
    gate = w.gate()
    for _ in range(3):
        w.search()
    probe = timer.run(w.search, 3, w.flush_l2, ahead_ms=1.0) / 3
    short = probe < 2.0
    if short:
        steps = SHORT_REPS
    ahead = 2.0 if short else 0.0
    # The gate (fp32 matmuls over the bank) and whatever ran before may have left the GPU at a
    # power-capped clock; a search of micro- or milliseconds is not what put it there, so both
    # passes below start from the idle state (a shape that runs for seconds caps itself anyway).
    idle = IDLE_BEFORE_S if probe < 50.0 else 0.0

    def rest():
        if idle:
            torch.cuda.synchronize(env["device"])
            time.sleep(idle)

    rest()
    launches0 = w.local.launch_count
    total = timer.run(w.search, steps, w.flush_l2, ahead_ms=ahead)
    per_step = sorted(timer.samples)
    launches = (w.local.launch_count - launches0) / steps
    ms = total / steps
    sm_mhz = env["clock_now"]()
    # the fused kernel alone: a second pass of the same length with the library's own event pair
    # around it (those events would sit between the launches of the timed pass)
    rest()
    w.local.profile(True)
    timer.run(w.search, steps, w.flush_l2, ahead_ms=ahead)
    kernel_ms = w.local.kernel_times_ms()
    w.local.profile(False)
    k_ms = statistics.mean(kernel_ms) if kernel_ms else ms
    roof = roofline_of(peaks, peaks_src, w.Q, w.hi - w.lo, w.k, k_ms, long_step=False)
    roof["kernel_share_of_step"] = k_ms / ms
    # whole search (every launch of it, launch gaps included) against the same roofline
    whole = roofline_of(peaks, peaks_src, w.Q, w.hi - w.lo, w.k, ms, long_step=False)
    out = {"workload": describe(w), "ms_per_step": ms, "value": w.Q / (ms * 1e-3), "unit": "queries/s",
           "steps": steps, "l2": "flushed between steps" if w.flush_l2 else "inputs larger than L2",
           "launches_per_search": launches, "plan_chunks_tiles_ctas": list(w.local.plan(w.Q, w.k)),
           "roofline": roof, "search_frac_of_roofline": whole["frac"], "parity_gate": gate,
           "measured": when + (f", after {idle} s idle" if idle else ""), "sm_mhz_after": sm_mhz}
    if per_step and env["world"] == 1:
        out["ms_min_median_max"] = [per_step[0], statistics.median(per_step), per_step[-1]]
    return out


def side_named(env, head_name: str, steps: int):
    """BASELINE configs other than the headline, each with its own bank, gate and roofline; config 3
    also with the HBM-bound batches and the unfused library strawman.  Run BEFORE the headline bank
    is built: these are searches of 50 us - 4 ms, and right after the headline's seconds at the
    power cap the GPU still runs them at the capped clock (round 2: 0.31 vs 0.45 of roofline for the
    same kernel at config 1), which says nothing about the shape itself."""
    torch, timer, world, device = env["torch"], env["timer"], env["world"], env["device"]
    out = []
    side_steps = max(3, min(steps, 20))
    for name in ("clotho_eval", "audiocaps", "wavcaps_400k", "allpairs_400k"):
        if name == head_name:
            continue
        w = Workload(env, name)
        if name == "wavcaps_400k":      # the HBM-bound batches first, then the 8,192-query batch
            for q_small in SMALL_Q:
                ws = Workload(env, name, Q=q_small, k=10, bank=w.bank, label=f"{name}_q{q_small}")
                out.append(time_side_workload(env, ws, side_steps, "before the headline"))
        out.append(time_side_workload(env, w, 3 if name == "allpairs_400k" else side_steps, "before the headline"))
        if name == "wavcaps_400k":
            if world == 1:      # unfused library strawman on the same GPU, same operands
                bank_bf16 = torch.empty(w.N, D, dtype=torch.bfloat16, device=device)
                for row0, rows in bank_rows_fp32(torch, device, w.bank_seed, 0, w.N):
                    bank_bf16[row0:row0 + rows.shape[0]] = torch.nn.functional.normalize(rows, dim=-1).bfloat16()
                q_bf16 = torch.nn.functional.normalize(w.q_dev, dim=-1).bfloat16()

                def straw():
                    return strawman_topk(torch, q_bf16, bank_bf16, w.k)

                for _ in range(2):
                    straw()
                st_ms = timer.run(straw, 3, False) / 3
                out[-1]["strawman_torch_matmul_bf16_topk_ms"] = st_ms
                del bank_bf16, q_bf16
        w.close()
        del w
        torch.cuda.empty_cache()
    return out


def side_resident(env, head: Workload, steps: int):
    """HBM-bound batches against the resident headline bank, and the strawman on a headline sample."""
    torch, timer, world, device = env["torch"], env["timer"], env["world"], env["device"]
    out = []
    side_steps = max(3, min(steps, 20))
    for q_small in SMALL_Q:
        w = Workload(env, head.name, Q=q_small, k=10, bank=head.bank, label=f"{head.name}_q{q_small}")
        out.append(time_side_workload(env, w, side_steps, "after the headline"))
    if world == 1 and head.name == "synthetic_10m":
        # (the [Q, N] fp32 score block of the whole batch does not fit: 1,024 queries)
        n_q, N, k = 1024, head.N, head.k
        bank_bf16 = torch.empty(N, D, dtype=torch.bfloat16, device=device)
        for row0, rows in bank_rows_fp32(torch, device, head.bank_seed, 0, N):
            bank_bf16[row0:row0 + rows.shape[0]] = torch.nn.functional.normalize(rows, dim=-1).bfloat16()
        q_bf16 = torch.nn.functional.normalize(head.q_dev[:n_q], dim=-1).bfloat16()
        strawman_topk(torch, q_bf16, bank_bf16, k)
        st_ms = timer.run(lambda: strawman_topk(torch, q_bf16, bank_bf16, k), 2, False) / 2
        ours_ms = timer.run(lambda: head.bank.search(head.q_dev[:n_q], k), 5, False) / 5
        out.append({"workload": f"{head.name} sample: {n_q} queries vs {N}-row bank, top-{k}",
                    "ms_per_step": ours_ms, "value": n_q / (ours_ms * 1e-3), "unit": "queries/s",
                    "strawman_torch_matmul_bf16_topk_ms": st_ms})
        del bank_bf16, q_bf16
        torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    import zsaac_b200
    from zsaac_b200.sharded import SearchPipeline

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
    peaks, peaks_src = load_peaks()
    timer = Timer(torch, dist, world, device)
    env = {"torch": torch, "dist": dist, "zs": zsaac_b200, "world": world, "rank": rank, "device": device,
           "peaks": peaks, "peaks_src": peaks_src, "timer": timer, "clock_now": ClockSampler(local_rank).now}
    steps = args.steps
    warmup = max(args.warmup, 3)

    # ---- the other named shapes first (short searches: measured before the headline heats the GPU)
    workloads = []
    workloads_named = [] if args.headline_only else side_named(env, args.workload, steps)

    # ---- headline workload: bank resident in HBM as bf16 before anything is timed
    balance = world > 1 and args.balance
    head = Workload(env, args.workload, Q=args.queries or None, N=args.bank_rows or None,
                    overlap=BALANCE_OVERLAP if balance else 0.0)
    Q, N, k = head.Q, head.N, head.k
    local, bank = head.local, head.bank

    # ---- untimed: the run verifies itself first
    gate = head.gate()
    if not gate["ok"]:
        if rank == 0:
            print(json.dumps({"parity_gate": gate, "error": "parity gate failed; nothing was timed"}), flush=True)
        return 3

    # ---- timed region 1: inputs resident in HBM (value).  N > 1: cast -> fused kernel -> chunk
    # merge -> all-gather -> k-way merge, all on one stream (the fused kernel is persistent and owns
    # every SM: nothing that needs SMs can overlap it, see SearchPipeline), preallocated buffers.
    pipe = SearchPipeline(bank, Q, k, depth=2, from_host=False, to_host=False, result="replicated",
                          self_index=head.self_index,
                          balance_every=BALANCE_EVERY if balance else 0) if world > 1 else None
    dev_slot = [0]

    def step_device():
        if pipe is not None:
            dev_slot[0] = pipe.submit(head.q_dev)
        else:
            head.search()

    finish_device = pipe.wait_stream if pipe is not None else None
    for _ in range(warmup):
        step_device()
    if pipe is not None:
        pipe.wait_stream()
    torch.cuda.synchronize(device)
    launches0 = local.launch_count
    local.profile(True)
    with ClockSampler(local_rank) as clocks:
        total_ms = timer.run(step_device, steps, head.flush_l2, finish=finish_device)
    kernel_ms = local.kernel_times_ms()
    local.profile(False)
    launches = local.launch_count - launches0
    ms_per_step = total_ms / steps
    value = Q / (ms_per_step * 1e-3)
    if pipe is not None:
        # the shard boundaries have moved since the gate: the result must still be the gated bits
        ds, di = pipe.result_of(dev_slot[0])
        flag = torch.tensor([int(torch.equal(ds, head.gate_result[0]) and torch.equal(di, head.gate_result[1]))],
                            device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        gate["timed_result_identical_to_gated"] = bool(flag.item())
        if not gate["timed_result_identical_to_gated"]:
            if rank == 0:
                print(json.dumps({"parity_gate": gate, "error": "result of the timed steps differs from the gated result"}), flush=True)
            return 3

    # ---- timed region 2: host buffers in, host buffers out (e2e).  Every rank uploads the pinned
    # host batch over its own PCIe link and reads back its 1/N of the merged rows; the copy engines
    # run the upload of step i+1 and the read-back of step i-1 under the fused kernel of step i.
    e2e_pipe = SearchPipeline(bank, Q, k, depth=2, from_host=True, to_host=True, result="row_slice",
                              self_index=head.self_index, input="full",
                              balance_every=BALANCE_EVERY if balance else 0)

    last_slot = [0]

    def step_e2e():
        last_slot[0] = e2e_pipe.submit(head.q_host)

    for _ in range(2):
        step_e2e()
    e2e_pipe.wait_stream()
    torch.cuda.synchronize(device)
    # the read-back of the e2e path is checked against the device-resident result (untimed)
    r_lo, r_hi = e2e_pipe.out_rows
    ref_s, ref_i = head.search()
    torch.cuda.synchronize(device)
    hs, hi_ = e2e_pipe.result_of(last_slot[0], host=True)
    e2e_ok = torch.equal(hs, ref_s[r_lo:r_hi].cpu()) and torch.equal(hi_, ref_i[r_lo:r_hi].cpu())
    flag = torch.tensor([int(e2e_ok)], device=device)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    gate["e2e_readback_identical"] = bool(flag.item())
    if not gate["e2e_readback_identical"]:
        if rank == 0:
            print(json.dumps({"parity_gate": gate, "error": "e2e read-back differs from the device result"}), flush=True)
        return 3
    e2e_ms = timer.run(step_e2e, steps, head.flush_l2, finish=e2e_pipe.wait_stream) / steps
    e2e_value = Q / (e2e_ms * 1e-3)
    hs, hi_ = e2e_pipe.result_of(last_slot[0], host=True)      # (timer.run ended with a device synchronise)
    flag = torch.tensor([int(torch.equal(hs, ref_s[r_lo:r_hi].cpu()) and torch.equal(hi_, ref_i[r_lo:r_hi].cpu()))],
                        device=device)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    gate["e2e_timed_readback_identical"] = bool(flag.item())
    if not gate["e2e_timed_readback_identical"]:
        if rank == 0:
            print(json.dumps({"parity_gate": gate, "error": "e2e read-back of the timed steps differs from the device result"}), flush=True)
        return 3

    # ---- roofline of the dominant kernel (this rank's shard) ---------------------------------
    shard_rows = head.hi - head.lo
    if pipe is not None and pipe.rows_log:      # moving boundaries: the rows this rank searched in the timed steps
        shard_rows = statistics.mean(pipe.rows_log[-steps:])
    k_ms = statistics.mean(kernel_ms) if kernel_ms else float("nan")
    clock_report = clocks.report()
    capped = ("sw_power_cap" in clock_report["reasons"] and clock_report["sm_mhz"] is not None
              and clock_report["sm_max_mhz"] and clock_report["sm_mhz"] < 0.85 * clock_report["sm_max_mhz"])
    roofline = roofline_of(peaks, peaks_src, Q, shard_rows, k, k_ms, long_step=capped or total_ms > 1000.0)
    if world > 1:
        # every step ends in an exchange, so the slowest rank's kernel sets the step time
        per_rank = torch.zeros(world, device=device)
        per_rank[rank] = k_ms
        dist.all_reduce(per_rank, op=dist.ReduceOp.SUM)
        roofline["kernel_ms_per_rank"] = [round(v, 3) for v in per_rank.tolist()]
        roofline["step_tail_ms_beyond_slowest_kernel"] = round(ms_per_step - max(per_rank.tolist()), 3)
        # ... of THAT step: the mean over steps of the slowest rank's kernel exceeds the slowest
        # rank's mean when the step-to-step noise of the GPUs is of the order of their differences
        n_k = min(len(kernel_ms), steps)
        per_step = torch.zeros(world, max(n_k, 1), device=device)
        if n_k:
            per_step[rank, :n_k] = torch.tensor(kernel_ms[-n_k:], device=device)
        dist.all_reduce(per_step, op=dist.ReduceOp.SUM)
        slowest = per_step.max(dim=0).values
        roofline["slowest_kernel_per_step_ms"] = {"mean": round(slowest.mean().item(), 3),
                                                  "min": round(slowest.min().item(), 3),
                                                  "max": round(slowest.max().item(), 3),
                                                  "std_of_a_rank": round(per_step.std(dim=1).mean().item(), 3)}
    roofline["kernel_share_of_step"] = k_ms / ms_per_step
    roofline["traffic"] = None
    try:
        if world == 1:          # captured with ncu on one GPU holding the whole bank
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                roofline["traffic"] = json.load(f).get(head.name)
    except Exception:
        pass

    # ---- Q in {1, 32, 128} against the resident headline bank + the strawman on a headline sample
    if not args.headline_only:
        workloads = side_resident(env, head, steps) + workloads_named

    cpu_baseline = None
    literal = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        ref = CpuReference(torch, head.name, Q, N)
        cpu_baseline = ref.baseline(min(ref.one_pass() for _ in range(3)))
        if not args.headline_only:
            # the reference's literal loop on config 1: CPU port, and on this GPU as written
            ref1 = CpuReference(torch, "clotho_eval")
            literal = {"cpu": ref1.literal_loop(128)}
            from oracle import oracle
            bank_cuda = ref1.bank.to(device)
            items = [{"text_embedding": ref1.queries[i:i + 1].clone()} for i in range(ref1.q_s)]
            sum(1 for _ in oracle.process_data_literal(bank_cuda, items[:16], ref1.k, device="cuda"))   # warm-up
            t0 = time.perf_counter()
            n = sum(1 for _ in oracle.process_data_literal(bank_cuda, items, ref1.k, device="cuda"))
            torch.cuda.synchronize(device)
            dt = time.perf_counter() - t0
            literal["torch_cuda_as_written"] = {"value": n / dt, "unit": "queries/s",
                                                "sample": f"{n} items x {ref1.n_s} rows, {dt:.2f} s, torch CUDA fp32"}

    if rank == 0:
        plan = local.plan(Q, k)
        line = {
            "metric": "related-caption retrieval queries/s (top-k, d=1024)",
            "value": value, "unit": "queries/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(head.name, Q, N, k, head.excl, world, balance),
            "details": {
                "bank_rows_this_rank": head.hi - head.lo,
                "bank_rows_searched_per_step_this_rank": shard_rows,
                "shard_balance": ({"rebalances": pipe.rebalances + e2e_pipe.rebalances,
                                   "rows_per_rank_now": [hi - lo for lo, hi in bank.bounds]}
                                  if balance else None),
                "l2_flushed_between_steps": bool(head.flush_l2),
                "plan_chunks_tiles_ctas": list(plan),
                "tflops": 2.0 * Q * N * D / (ms_per_step * 1e-3) / 1e12,
            },
            "parity_gate": gate,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": "queries/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": e2e_pipe.h2d_bytes * world, "d2h_bytes_per_step": e2e_pipe.d2h_bytes * world,
                    "how": "pinned host queries in (every rank uploads the batch over its own PCIe link), merged "
                           "rows out to pinned host memory (1/N of the rows per rank, all-to-all + merge of that "
                           "slice); the copy engines overlap upload and read-back with the neighbouring steps' "
                           "fused kernels, collectives stay on the search stream"},
            "gpu_launches": int(launches) * world,
            "clocks": clock_report,
            "workloads": workloads,
            "literal_reference_loop_clotho_eval": literal,
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
    ap.add_argument("--balance", action="store_true",
                    help="N > 1: shard boundaries follow the measured speed of the GPUs (sharded.py) instead of "
                         "fixed equal shards; measured on 8 B200s: the ranks' mean kernel times become equal, "
                         "the step does not get shorter — the step-to-step noise of the slowest kernel is what "
                         "is left (profiles/r02/SUMMARY.md)")
    ap.add_argument("--headline-only", action="store_true",
                    help="skip the extra shapes / strawman / literal loop (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
