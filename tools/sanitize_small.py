"""Smallest end-to-end exercise of every kernel, for compute-sanitizer (memcheck) runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zsaac_b200
from zsaac_b200.predict_prompt import map2memory

g = torch.Generator().manual_seed(0)
bank = torch.randn(700, 1024, generator=g).cuda()
q = torch.randn(130, 1024, generator=g).cuda()
for cg in ("1", "2"):
    os.environ["ZSAAC_CTA_GROUP"] = cg
    rb = zsaac_b200.RelatedBank.from_tensor(bank)
    s, i = rb.search(q, 5)
    s2, i2 = rb.search(q, 32, self_index=torch.arange(130).cuda())
    r, ts = rb.rank_of(q, i[:, :3].contiguous())
    rows = rb.gather_rows(rb.normalize_rows(bank), i)
    ms, mi = rb.merge(torch.stack([s, s]), torch.stack([i, i + 1000]))
    torch.cuda.synchronize()
    assert (r == torch.arange(3).cuda()).all()
    rb.close()
out = map2memory(torch.nn.functional.normalize(q[:3], dim=-1), torch.nn.functional.normalize(bank, dim=-1))
torch.cuda.synchronize()
print("sanitize_small ok", float(s.sum()), float(out.sum()))
