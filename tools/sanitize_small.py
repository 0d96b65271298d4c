"""A handful of small calls through every kernel family, meant to be run under compute-sanitizer
(memcheck / racecheck) — and plainly first, so that a failure under the tool is the tool's finding.
usage: [compute-sanitizer --tool memcheck] python tools/sanitize_small.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zsaac_b200
from zsaac_b200.retrieval import exact_topk, exact_rank, search_rescored

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(7)
bank = torch.randn(3000, 1024, device=dev, generator=g)
rb = zsaac_b200.RelatedBank.from_tensor(bank)


def check(q, k, **kw):
    s, i = rb.search(q, k, **kw)
    ref = torch.nn.functional.normalize(q, dim=-1) @ torch.nn.functional.normalize(bank, dim=-1).T
    ws = ref.topk(k, dim=1).values
    torch.cuda.synchronize()
    assert (s - ws).abs().max().item() < 1e-3, (q.shape, k)
    return s, i


q_small = torch.randn(17, 1024, device=dev, generator=g)      # one 128-row tile: cast -> fused -> merge
q_mid = torch.randn(300, 1024, device=dev, generator=g)       # single-launch mode (pairs)
check(q_small, 10)
check(q_mid, 5)
check(q_mid, 40)                                              # two passes
fb = rb.normalize_rows(bank)
s, i = search_rescored(rb, q_mid, fb, 5)                      # fp32 re-scoring of k + 8 candidates
rows = rb.gather_rows(fb, i)
es, ei = exact_topk(q_small, fb[:527], 3, normalize=True)
er = exact_rank(q_small, fb[:527], torch.arange(17, device=dev).view(17, 1), normalize=True)
from zsaac_b200.predict_prompt import map2memory
m1 = map2memory(torch.nn.functional.normalize(q_small[:1], dim=-1), fb)
m8 = map2memory(torch.nn.functional.normalize(q_small[:16], dim=-1), fb)
torch.cuda.synchronize()
rb.close()
print("sanitize_small: all calls returned", flush=True)
