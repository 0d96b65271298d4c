"""Experiment (no kernel change): how exact is a 3-term bf16 split product on the tcgen05 path?

x = hi + lo (+ 2^-18 |x| residual), hi = bf16(x), lo = bf16(x - hi).  Concatenating along K,
    [q_hi | q_lo | q_hi] . [b_hi | b_hi | b_lo] = q_hi.b_hi + q_lo.b_hi + q_hi.b_lo  ~  q.b
runs on the existing fused pipeline (debug_scores, d' = 3d) with fp32 accumulation in tensor
memory.  Reports the error against float64 for unit vectors (the memory-projection scores) and,
for the second GEMM of that projection (softmax weights x bank rows: all-positive weights, long K),
how the accumulation error grows with K."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zsaac_b200

dev = torch.device("cuda", 0)


def split3(x):
    hi = x.bfloat16()
    lo = (x - hi.float()).bfloat16()
    return hi, lo


def scores_split(q, b):
    qh, ql = split3(q)
    bh, bl = split3(b)
    qq = torch.cat([qh, ql, qh], dim=1).contiguous()
    bb = torch.cat([bh, bh, bl], dim=1).contiguous()
    rb = zsaac_b200.RelatedBank.from_tensor(bb, normalize=False)
    s = rb.debug_scores(qq, normalize_queries=False)
    torch.cuda.synchronize()
    rb.close()
    return s


g = torch.Generator(device=dev).manual_seed(3)
out = {}
for d in (256, 1024, 1344):                      # 3d <= 4096 (ZS_MAX_DIM) and a multiple of 64
    d3 = 3 * d
    if d3 % 64 or d3 > 4096:
        continue
    q = torch.nn.functional.normalize(torch.randn(300, d, device=dev, generator=g), dim=-1)
    b = torch.nn.functional.normalize(torch.randn(5000, d, device=dev, generator=g), dim=-1)
    exact = (q.double() @ b.double().T)
    s3 = scores_split(q, b)
    s32 = q @ b.T                                 # torch fp32 (cuBLAS, TF32 off)
    rb = zsaac_b200.RelatedBank.from_tensor(b, normalize=False)
    s16 = rb.debug_scores(q, normalize_queries=False)
    torch.cuda.synchronize()
    rb.close()
    out[f"unit_vectors_d{d}"] = {"split3_max_err": (s3.double() - exact).abs().max().item(),
                                 "torch_fp32_max_err": (s32.double() - exact).abs().max().item(),
                                 "bf16_max_err": (s16.double() - exact).abs().max().item()}
# long-K accumulation with all-positive operands (weights x |bank values|): relative error vs K
for K in (256, 1024, 1344):
    p = torch.rand(256, K, device=dev, generator=g) + 0.5
    c = torch.rand(1024, K, device=dev, generator=g) + 0.5
    exact = p.double() @ c.double().T
    s3 = scores_split(p, c)
    s32 = p @ c.T
    rel3 = ((s3.double() - exact) / exact)
    rel32 = ((s32.double() - exact) / exact)
    out[f"positive_K{K}"] = {"split3_max_rel_err": rel3.abs().max().item(), "split3_mean_rel_err": rel3.mean().item(),
                             "torch_fp32_max_rel_err": rel32.abs().max().item(), "torch_fp32_mean_rel_err": rel32.mean().item()}
print(json.dumps(out, indent=1))
