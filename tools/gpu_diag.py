"""GPU bring-up diagnostics for the fused similarity/top-k kernel (run under gpurun).

Not a test and not a benchmark: prints how far the tcgen05 pipeline output is from a torch
matmul on the same bf16-rounded operands, then checks top-k and times a few shapes.
Usage: python tools/gpu_diag.py [--cg 1|2] [--stage scores|topk|time|all]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cg", type=int, default=1)
    ap.add_argument("--stage", default="all")
    args = ap.parse_args()
    os.environ["ZSAAC_CTA_GROUP"] = str(args.cg)

    import torch
    import zsaac_b200
    from zsaac_b200.retrieval import RelatedBank

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    print(f"[diag] cg={args.cg} device={torch.cuda.get_device_name(0)}", flush=True)

    def ref_scores(q, b, normalize=True):
        if normalize:
            q = torch.nn.functional.normalize(q, dim=-1)
            b = torch.nn.functional.normalize(b, dim=-1)
        return q.bfloat16().float() @ b.bfloat16().float().T

    ok_all = True
    if args.stage in ("scores", "all"):
        for (Q, N, d) in [(128, 256, 64), (128, 256, 128), (128, 256, 1024), (256, 512, 1024),
                          (300, 1000, 1024), (1, 257, 1024), (1045, 19195, 1024)]:
            g = torch.Generator(device=dev).manual_seed(Q * 7 + N)
            q = torch.randn(Q, d, device=dev, generator=g)
            b = torch.randn(N, d, device=dev, generator=g)
            rb = RelatedBank.from_tensor(b, normalize=True)
            got = rb.debug_scores(q, normalize_queries=True)
            torch.cuda.synchronize()
            want = ref_scores(q, b)
            err = (got - want).abs().max().item()
            ok = err < 2e-3
            ok_all &= ok
            print(f"[scores] Q={Q} N={N} d={d} plan={rb.plan(Q, 1)} max|err|={err:.3e} "
                  f"{'OK' if ok else 'MISMATCH'}", flush=True)
            if not ok:
                bad = (got - want).abs() > 2e-3
                rows_bad = bad.any(dim=1).nonzero().flatten()[:16].tolist()
                cols_bad = bad.any(dim=0).nonzero().flatten()[:32].tolist()
                print(f"   bad rows (first): {rows_bad}\n   bad cols (first): {cols_bad}")
                print(f"   got[0,:8]={got[0,:8].tolist()}\n   want[0,:8]={want[0,:8].tolist()}")
                # is the output a permutation of the expected rows / columns?
                gn = torch.nn.functional.normalize(got[:, :min(N, 256)], dim=1)
                wn = torch.nn.functional.normalize(want[:, :min(N, 256)], dim=1)
                match = (gn @ wn.T).argmax(dim=1)[:16].tolist()
                print(f"   best-matching expected row for got rows 0..15: {match}")
            rb.close()

    if args.stage in ("topk", "all"):
        for (Q, N, d, k) in [(128, 256, 1024, 5), (300, 1000, 1024, 5), (1045, 19195, 1024, 5),
                             (975, 49838, 1024, 10), (64, 100000, 1024, 32), (3, 5000, 1024, 1)]:
            g = torch.Generator(device=dev).manual_seed(Q * 11 + N)
            q = torch.randn(Q, d, device=dev, generator=g)
            b = torch.randn(N, d, device=dev, generator=g)
            rb = RelatedBank.from_tensor(b, normalize=True)
            s, i = rb.search(q, k)
            torch.cuda.synchronize()
            want = ref_scores(q, b)
            ws, wi = want.topk(k, dim=1)
            err = (s - ws).abs().max().item()
            same = (i == wi).float().mean().item()
            # every returned index must be (nearly) as good as the k-th reference score
            got_ref_scores = want.gather(1, i)
            ok = err < 1e-4 and bool((got_ref_scores >= ws[:, -1:] - 1e-5).all())
            ok_all &= ok
            print(f"[topk] Q={Q} N={N} k={k} plan={rb.plan(Q, k)} max|dscore|={err:.3e} "
                  f"index match={same:.4f} {'OK' if ok else 'MISMATCH'}", flush=True)
            rb.close()

    if args.stage in ("time", "all"):
        for (Q, N, d, k) in [(975, 49838, 1024, 10), (8192, 400000, 1024, 10), (128, 400000, 1024, 10),
                             (16384, 400000, 1024, 32)]:
            g = torch.Generator(device=dev).manual_seed(1)
            q = torch.randn(Q, d, device=dev, generator=g)
            b = torch.randn(N, d, device=dev, generator=g)
            rb = RelatedBank.from_tensor(b, normalize=True)
            del b
            rb.reserve(Q, k)
            out = (torch.empty(Q, k, device=dev), torch.empty(Q, k, dtype=torch.int64, device=dev))
            for _ in range(3):
                rb.search(q, k, out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            e0.record()
            for _ in range(reps):
                rb.search(q, k, out=out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            tf = 2.0 * Q * N * d / (ms * 1e-3) / 1e12
            gbs = (2.0 * N * d) / (ms * 1e-3) / 1e9
            print(f"[time] Q={Q} N={N} k={k} plan={rb.plan(Q, k)} {ms:.3f} ms  {tf:.1f} TFLOP/s  "
                  f"bank stream {gbs:.0f} GB/s  {Q / (ms * 1e-3):.0f} q/s", flush=True)
            rb.close()

    print(f"[diag] cg={args.cg} {'ALL OK' if ok_all else 'FAILURES'}", flush=True)
    return 0 if ok_all else 1


if __name__ == "__main__":
    sys.exit(main())
