"""Small / latency-bound shapes, event-timed: the reference's own sizes (configs 1, 2), the
HBM-bound small batches against 400 k rows, the Q = 129..512 range and the label bank.

Banks larger than L2 are timed back to back between one event pair; smaller ones get an event
pair per repetition with a 256 MB write in between (L2 flush).  Variants are selected through the
library's tuning environment (read when a RelatedBank is created).
usage: python tools/bench_small.py [--variants] > gpurun_out/bench_small.jsonl"""
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zsaac_b200

dev = torch.device("cuda", 0)
HBM = 6553.3e9
TF = 1648.6e12
L2 = 126 << 20
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def time_search(rb, q, k, bank_bytes, reps):
    for _ in range(5):
        rb.search(q, k)
    torch.cuda.synchronize()
    if bank_bytes > 2 * L2:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            rb.search(q, k)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, "back-to-back"
    torch.cuda._sleep(4_000_000)        # park the stream: every launch below is queued before the device gets to it
    pairs = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rb.search(q, k)
        e1.record()
        pairs.append((e0, e1))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in pairs), "l2-flushed, parked"


SHAPES = [("clotho_eval", 1045, 19195, 5), ("audiocaps", 975, 49838, 10),
          ("q1_400k", 1, 400_000, 10), ("q32_400k", 32, 400_000, 10), ("q128_400k", 128, 400_000, 10),
          ("q256_400k", 256, 400_000, 10), ("q512_400k", 512, 400_000, 10), ("q1024_400k", 1024, 400_000, 10)]
VARIANTS = [("default", {})]
if "--variants" in sys.argv:
    VARIANTS += [("solo0", {"ZSAAC_SOLO": "0"}), ("boot0", {"ZSAAC_BOOT": "0"}),
                 ("solo0_boot0", {"ZSAAC_SOLO": "0", "ZSAAC_BOOT": "0"}),
                 ("cg1", {"ZSAAC_CTA_GROUP": "1"}), ("cg2", {"ZSAAC_CTA_GROUP": "2"})]

banks = {}
for name, Q, N, k in SHAPES:
    if N not in banks:
        g = torch.Generator(device=dev).manual_seed(N)
        banks[N] = torch.randn(N, 1024, device=dev, generator=g)
    g = torch.Generator(device=dev).manual_seed(Q)
    q = torch.randn(Q, 1024, device=dev, generator=g)
    for vname, env in VARIANTS:
        for key in ("ZSAAC_SOLO", "ZSAAC_BOOT", "ZSAAC_CTA_GROUP"):
            os.environ.pop(key, None)
        os.environ.update(env)
        rb = zsaac_b200.RelatedBank.from_tensor(banks[N])
        n0 = rb.launch_count
        rb.search(q, k)
        launches = rb.launch_count - n0
        ms, how = time_search(rb, q, k, N * 2048, 50)
        flop = 2.0 * Q * N * 1024
        byts = 2.0 * N * 1024 + Q * 1024 * 4.0 + Q * k * 12.0
        ideal = max(flop / TF, byts / HBM)
        print(json.dumps({"case": name, "variant": vname, "Q": Q, "N": N, "k": k, "ms": round(ms, 5),
                          "launches_per_search": launches, "how": how, "plan": list(rb.plan(Q, k)),
                          "tflops": round(flop / ms / 1e9, 1), "bank_gbs": round(byts / ms / 1e6, 0),
                          "roofline_frac": round(ideal * 1e3 / ms, 3),
                          "bound": "tensor" if flop / TF > byts / HBM else "hbm"}), flush=True)
        rb.close()

# label bank: exact fp32 single launch vs the torch formula on the same GPU and on the host CPU
from zsaac_b200.utils import sound_effect_choice
from zsaac_b200.retrieval import exact_topk
import time
labels = torch.nn.functional.normalize(torch.randn(527, 1024), dim=-1)
labels_dev = labels.to(dev)
for B in (1, 32):
    prefix = torch.nn.functional.normalize(torch.randn(B, 1024), dim=-1)
    prefix_dev = prefix.to(dev)
    for _ in range(5):
        exact_topk(prefix_dev, labels_dev, 3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        exact_topk(prefix_dev, labels_dev, 3)
    e1.record()
    torch.cuda.synchronize()
    dev_us = e0.elapsed_time(e1) / 200 * 1e3
    t0 = time.perf_counter()
    for _ in range(200):
        sound_effect_choice(prefix, labels_dev, 3)          # CPU prefix in, CPU indices out (sync)
    wall_us = (time.perf_counter() - t0) / 200 * 1e6
    t0 = time.perf_counter()
    for _ in range(200):
        for b in range(B):                                   # the reference: one call per sample
            sim = prefix[b:b + 1] @ labels.t()
            torch.topk(torch.softmax(sim, dim=-1), 3, dim=-1)
    ref_us = (time.perf_counter() - t0) / 200 * 1e6
    print(json.dumps({"case": f"sound_effect_choice B={B} x 527", "device_us": round(dev_us, 2),
                      "call_wall_us_incl_h2d_d2h": round(wall_us, 1),
                      "reference_cpu_us_for_B_calls": round(ref_us, 1),
                      "cpu_threads": torch.get_num_threads()}), flush=True)
