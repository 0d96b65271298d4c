"""All BASELINE configs + small-batch (HBM-bound) points + the unfused GPU strawman
(torch.matmul bf16 -> torch.topk) on one B200.  Prints one line per case; run under gpurun."""
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zsaac_b200

D = 1024
dev = torch.device("cuda", 0)
peaks = {"bf16_tflops": 1648.6, "bf16_tflops_sustained": 1370.0, "hbm_gbs": 6553.3}
try:
    peaks.update(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))))
except Exception:
    pass


def time_ms(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    return statistics.median(times), min(times)


def strawman(q, bank_bf16, k, chunk=4096):
    qn = torch.nn.functional.normalize(q, dim=-1).bfloat16()
    outs = []
    for lo in range(0, qn.shape[0], chunk):
        s = torch.matmul(qn[lo:lo + chunk], bank_bf16.T).float()
        outs.append(s.topk(k, dim=1))
    return outs


cases = [("config1 clotho", 1045, 19195, 5), ("config2 audiocaps", 975, 49838, 10),
         ("config3 wavcaps", 8192, 400_000, 10), ("k=32 on 400k", 16384, 400_000, 32),
         ("Q=1 vs 400k", 1, 400_000, 10), ("Q=32 vs 400k", 32, 400_000, 10),
         ("Q=128 vs 400k", 128, 400_000, 10), ("Q=256 vs 400k", 256, 400_000, 10),
         ("Q=128 vs 4M", 128, 4_000_000, 10), ("Q=1 vs 527 labels", 1, 527, 3)]
results = []
bank_cache = {}
for name, Q, N, k in cases:
    g = torch.Generator(device=dev).manual_seed(N + Q)
    if N not in bank_cache:
        bank_cache.clear()
        rb = zsaac_b200.RelatedBank(N, D, device=dev)
        for lo in range(0, N, 65536):
            rows = torch.randn(min(65536, N - lo), D, device=dev, generator=g)
            rb.upload(rows, lo)
        bank_cache[N] = rb
    rb = bank_cache[N]
    q = torch.randn(Q, D, device=dev, generator=g)
    rb.reserve(Q, k)
    out = (torch.empty(Q, k, device=dev), torch.empty(Q, k, dtype=torch.int64, device=dev))
    reps = 20 if Q * N < 4e9 else 5
    med, best = time_ms(lambda: rb.search(q, k, out=out), reps)
    flop = 2.0 * Q * N * D
    byts = 2.0 * N * D + 2.0 * Q * D + 12.0 * Q * k
    t_ideal = max(flop / (peaks["bf16_tflops"] * 1e12), byts / (peaks["hbm_gbs"] * 1e9))
    line = {"case": name, "Q": Q, "N": N, "k": k, "plan": rb.plan(Q, k), "ms_median": round(med, 4),
            "ms_min": round(best, 4), "queries_per_s": round(Q / (med * 1e-3)),
            "tflops": round(flop / (med * 1e-3) / 1e12, 1), "bank_gbs": round(byts / (med * 1e-3) / 1e9),
            "roofline_frac": round(t_ideal / (med * 1e-3), 3),
            "bound": "tensor" if flop / (peaks["bf16_tflops"] * 1e12) > byts / (peaks["hbm_gbs"] * 1e9) else "hbm"}
    if N <= 400_000 and Q >= 975 and Q <= 8192:
        bank_bf16 = torch.nn.functional.normalize(torch.randn(N, D, device=dev, generator=g), dim=-1).bfloat16()
        sm, _ = time_ms(lambda: strawman(q, bank_bf16, k), 5)
        line["strawman_matmul_topk_ms"] = round(sm, 3)
        del bank_bf16
    print(json.dumps(line), flush=True)
