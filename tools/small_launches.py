"""A few searches at the small / latency-bound shapes, for an ncu launch list
(ncu --metrics gpu__time_duration.sum): which kernels does one search launch, and how long do
they take?  usage: python tools/small_launches.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zsaac_b200
from zsaac_b200.retrieval import exact_topk

dev = torch.device("cuda", 0)
for (Q, N, k) in [(975, 49_838, 10), (1045, 19_195, 5), (256, 400_000, 10), (128, 400_000, 10), (1, 400_000, 10),
                  (975, 49_838, 50)]:
    g = torch.Generator(device=dev).manual_seed(Q + N)
    rb = zsaac_b200.RelatedBank.from_tensor(torch.randn(N, 1024, device=dev, generator=g))
    q = torch.randn(Q, 1024, device=dev, generator=g)
    for _ in range(4):
        rb.search(q, k)
    torch.cuda.synchronize()
    rb.close()
# the label bank (utils.sound_effect_choice) and the zero-shot prompts: exact fp32, one launch
labels = torch.nn.functional.normalize(torch.randn(527, 1024, device=dev), dim=-1)
for B in (1, 32):
    prefix = torch.nn.functional.normalize(torch.randn(B, 1024, device=dev), dim=-1)
    for _ in range(4):
        exact_topk(prefix, labels, 3)
torch.cuda.synchronize()
