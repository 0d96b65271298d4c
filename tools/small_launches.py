"""A few searches at the small / HBM-bound shapes, for an ncu launch list
(ncu --metrics gpu__time_duration.sum): how long do the cast and merge kernels around the fused
kernel take?  usage: python tools/small_launches.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zsaac_b200

dev = torch.device("cuda", 0)
for (Q, N, k) in [(128, 400_000, 10), (1, 400_000, 10), (975, 49_838, 10), (1045, 19_195, 5), (1, 527, 3)]:
    g = torch.Generator(device=dev).manual_seed(Q + N)
    rb = zsaac_b200.RelatedBank.from_tensor(torch.randn(N, 1024, device=dev, generator=g))
    q = torch.randn(Q, 1024, device=dev, generator=g)
    for _ in range(4):
        rb.search(q, k)
    torch.cuda.synchronize()
    rb.close()
