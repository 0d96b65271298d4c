"""Where does the time go on small problems?  Per-CTA timeline of the fused kernel from the
zs_debug_trace stamps (ns, relative to the earliest CTA entry).  Run under gpurun."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zsaac_b200

dev = torch.device("cuda", 0)
CASES = [  # Q, N, k, cta_group, forced chunks (0 = planner)
    (975, 49838, 10, "2", 0), (975, 49838, 1, "2", 0), (1045, 19195, 5, "1", 0), (1045, 19195, 5, "2", 0),
    (256, 400000, 10, "2", 0), (128, 400000, 10, "1", 0), (8192, 400000, 10, "2", 0),
]
for (Q, N, k, cg, chunks_forced) in CASES:
    os.environ["ZSAAC_CTA_GROUP"] = cg
    if chunks_forced:
        os.environ["ZSAAC_CHUNKS"] = str(chunks_forced)
    else:
        os.environ.pop("ZSAAC_CHUNKS", None)
    g = torch.Generator(device=dev).manual_seed(5)
    q = torch.randn(Q, 1024, device=dev, generator=g)
    b = torch.randn(N, 1024, device=dev, generator=g)
    rb = zsaac_b200.RelatedBank.from_tensor(b)
    chunks, tpc, ctas = rb.plan(Q, k)
    for _ in range(20):                       # keep the GPU busy so the clocks are up
        rb.search(q, k)
    stamps = torch.zeros(ctas, 16, dtype=torch.int64, device=dev)
    rb.trace(stamps)
    rb.profile(True)
    rb.search(q, k)
    torch.cuda.synchronize()
    kms = rb.kernel_times_ms()[-1]
    rb.trace(None)
    t = stamps.cpu().double()
    t0 = t[:, 0].min()
    rel = (t[:, :11] - t0) / 1e3            # us
    names = ["entry", "setup done", "1st tile MMA done", "last tile scanned", "lists written", "exit"]
    ghz = ((t[:, 7] - t[:, 6]) / (t[:, 5] - t[:, 0])).median()
    units_per_cta = -(-(chunks * -(-Q // (128 * int(cg)))) // (ctas // int(cg)))
    print(f"\nQ={Q} N={N} k={k} cg={cg} plan=(chunks {chunks}, tiles/chunk {tpc}, ctas {ctas}, "
          f"units/cta {units_per_cta}) kernel (events) {kms * 1e3:.1f} us, SM clock {ghz:.3f} GHz")
    for i, n in enumerate(names):
        col = rel[:, i]
        print(f"  {n:20s} min {col.min():8.2f}  median {col.median():8.2f}  max {col.max():8.2f} us")
    if t[:, 8].max() > 0:                  # single-launch mode
        for i, n in ((8, "own rows cast"), (9, "all rows cast"), (10, "all lists written")):
            col = rel[:, i][t[:, i] > 0]
            print(f"  {n:20s} min {col.min():8.2f}  median {col.median():8.2f}  max {col.max():8.2f} us")
    busy = (rel[:, 3] - rel[:, 2]).median()
    tiles = units_per_cta * tpc
    per_tile = busy / max(tiles - 1, 1)
    print(f"  1st tile done -> last tile scanned: {busy:.2f} us over ~{tiles - 1} tiles => "
          f"{per_tile:.2f} us = {per_tile * ghz * 1e3:.0f} cycles per tile (8192 = MMA floor)")
    rb.close()
