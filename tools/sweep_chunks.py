"""Sweep the number of bank chunks (ZSAAC_CHUNKS) and k of the fused kernel at fixed shapes:
isolates what a work unit costs beyond its bank tiles (top-k list warm-up, pipeline fill,
partial write-out).  Prints one JSON line per point; run under gpurun.
usage: python tools/sweep_chunks.py [shard|wavcaps|k32|audiocaps ...]"""
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zsaac_b200

D = 1024
dev = torch.device("cuda", 0)

SHAPES = {
    # name: (Q, N, [k ...], [chunks ...]); chunks 0 = the library's own plan
    "shard": (65536, 1_250_000, [32, 10, 1], [0, 2, 4, 9]),      # one rank of config 4 at 8 GPUs
    "wavcaps": (8192, 400_000, [10, 1, 32], [0, 2, 9, 19]),
    "k32": (16384, 400_000, [32, 10, 1], [0, 4, 9]),
    "audiocaps": (975, 49838, [10, 1, 32], [0, 24]),
    "clotho": (1045, 19195, [5, 1], [0]),
    "hbm128": (128, 400_000, [10, 32], [0]),                              # HBM-bound: bank stream
    "hbm1": (1, 400_000, [10], [0]),
}


# A/B variants: ';'-separated sets of ','-separated library tuning variables, e.g.
#   SWEEP_VARIANTS="ZSAAC_SHARE_THR=1;ZSAAC_SHARE_THR=0;ZSAAC_LOCKSTEP=0"
VARIANTS = [dict(kv.split("=") for kv in v.split(",") if kv)
            for v in os.environ.get("SWEEP_VARIANTS", "ZSAAC_SHARE_THR=1;ZSAAC_SHARE_THR=0").split(";")]
TUNABLES = ("ZSAAC_SHARE_THR", "ZSAAC_LOCKSTEP", "ZSAAC_SYNC_WINDOW", "ZSAAC_KCAP_POW2")


def time_kernel(rb, q, k, out, reps):
    for _ in range(3):
        rb.search(q, k, out=out)
    torch.cuda.synchronize()
    rb.profile(True)
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rb.search(q, k, out=out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    kt = rb.kernel_times_ms()
    rb.profile(False)
    return statistics.median(ts), statistics.median(kt) if kt else float("nan")


def main():
    names = sys.argv[1:] or list(SHAPES)
    for name in names:
        Q, N, ks, chunk_list = SHAPES[name]
        if os.environ.get("SWEEP_QUICK"):       # default plan and the first k only
            ks, chunk_list = ks[:1], [0]
        g = torch.Generator(device=dev).manual_seed(N + Q)
        rb = zsaac_b200.RelatedBank(N, D, device=dev)
        for lo in range(0, N, 65536):
            rb.upload(torch.randn(min(65536, N - lo), D, device=dev, generator=g), lo)
        q = torch.randn(Q, D, device=dev, generator=g)
        flop = 2.0 * Q * N * D
        for k in ks:
            out = (torch.empty(Q, k, device=dev), torch.empty(Q, k, dtype=torch.int64, device=dev))
            for chunks in (chunk_list if k == ks[0] else [0]):
                if chunks:
                    os.environ["ZSAAC_CHUNKS"] = str(chunks)
                else:
                    os.environ.pop("ZSAAC_CHUNKS", None)
                rb.reserve(Q, k)
                reps = 5 if flop > 1e14 else 20
                for var in VARIANTS:
                    for t in TUNABLES:
                        os.environ.pop(t, None)
                    os.environ.update(var)
                    med, kern = time_kernel(rb, q, k, out, reps)
                    print(json.dumps({"shape": name, "Q": Q, "N": N, "k": k, "chunks_forced": chunks,
                                      "variant": var, "plan": rb.plan(Q, k),
                                      "search_ms": round(med, 4), "kernel_ms": round(kern, 4),
                                      "kernel_tflops": round(flop / (kern * 1e-3) / 1e12, 1)}), flush=True)
        os.environ.pop("ZSAAC_CHUNKS", None)
        for t in TUNABLES:
            os.environ.pop(t, None)
        rb.close()
        del rb, q
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
