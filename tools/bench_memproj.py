"""map2memory (softmax-weighted memory projection, reference predict_prompt.py:23-29) on one B200:
the streaming kernel for up to 7 queries, the tensor-core path (split-bf16 operands, two
contractions on the fused kernel's pipeline) for batches, next to torch's fp32 formulation on the
same GPU and the float64 error of both.  usage: python tools/bench_memproj.py"""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zsaac_b200
from zsaac_b200.predict_prompt import map2memory

dev = torch.device("cuda", 0)
banks = {}
for (Q, N) in [(1, 49838), (1, 400_000), (4, 400_000), (8, 400_000), (32, 400_000), (128, 400_000),
               (256, 400_000), (1045, 400_000), (1045, 19_195), (975, 49_838)]:
    if N not in banks:
        banks.clear()
        g = torch.Generator(device=dev).manual_seed(N)
        banks[N] = torch.nn.functional.normalize(torch.randn(N, 1024, device=dev, generator=g), dim=-1)
    bank = banks[N]
    g = torch.Generator(device=dev).manual_seed(N + Q)
    q = torch.nn.functional.normalize(torch.randn(Q, 1024, device=dev, generator=g), dim=-1)
    q[0] = torch.nn.functional.normalize(bank[7] + 0.02 * q[0], dim=-1)
    for _ in range(3):
        out = map2memory(q, bank)
    torch.cuda.synchronize()
    times = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); map2memory(q, bank); e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = statistics.median(times)

    def ref():
        sim = (q @ bank.T * 100).softmax(dim=-1); o = sim @ bank; return o / o.norm(dim=-1, keepdim=True)
    for _ in range(3):
        want32 = ref()
    torch.cuda.synchronize()
    rt = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ref(); e1.record(); torch.cuda.synchronize()
        rt.append(e0.elapsed_time(e1))
    # float64 reference on a sample of the queries (the whole [Q, N] double matrix is large)
    qs = q[:8].double()
    sim = (qs @ bank.double().T * 100).softmax(dim=-1)
    o = sim @ bank.double()
    want64 = o / o.norm(dim=-1, keepdim=True)
    flop = 4.0 * Q * N * 1024                      # both contractions, algorithmic (fp32-equivalent)
    print(json.dumps({"Q": Q, "N": N, "path": "tensor cores, split bf16" if Q >= 8 else "streaming fp32",
                      "ms": round(ms, 4), "torch_fp32_same_gpu_ms": round(statistics.median(rt), 4),
                      "speedup_vs_torch": round(statistics.median(rt) / ms, 2),
                      "algorithmic_tflops": round(flop / ms / 1e9, 1),
                      "max_err_vs_float64": float((out[:8].double() - want64).abs().max()),
                      "torch_fp32_max_err_vs_float64": float((want32[:8].double() - want64).abs().max())}), flush=True)
