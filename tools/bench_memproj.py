"""map2memory (softmax-weighted memory projection) on one B200: HBM-bound stream of the fp32 bank."""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zsaac_b200
from zsaac_b200.predict_prompt import map2memory

dev = torch.device("cuda", 0)
for (Q, N) in [(1, 49838), (1, 400_000), (4, 400_000), (1, 2_000_000), (8, 400_000)]:
    g = torch.Generator(device=dev).manual_seed(N + Q)
    bank = torch.nn.functional.normalize(torch.randn(N, 1024, device=dev, generator=g), dim=-1)
    q = torch.nn.functional.normalize(torch.randn(Q, 1024, device=dev, generator=g), dim=-1)
    for _ in range(3):
        map2memory(q, bank)
    torch.cuda.synchronize()
    times = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); map2memory(q, bank); e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = statistics.median(times)
    # passes over the bank: one per query for banks >= 256 MB, one per pair of queries below
    passes = Q if N * 1024 * 4 >= 256e6 else -(-Q // 2)
    gbs = passes * N * 1024 * 4 / (ms * 1e-3) / 1e9
    # torch reference formulation on the same GPU
    def ref():
        sim = (q @ bank.T * 100).softmax(dim=-1); o = sim @ bank; return o / o.norm(dim=-1, keepdim=True)
    for _ in range(3): ref()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ref(); e1.record(); torch.cuda.synchronize()
    print(json.dumps({"Q": Q, "N": N, "ms": round(ms, 4), "bank_stream_GBs": round(gbs),
                      "frac_of_copy_bw_6553": round(gbs / 6553.3, 3), "torch_fp32_same_gpu_ms": round(e0.elapsed_time(e1), 4)}), flush=True)
