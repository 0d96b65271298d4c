"""Wall-clock split of the generator script as main() runs it — process_data streamed into
save_data_to_hdf5, so the search of batch i+1 overlaps the pickling of batch i — at WavCaps scale,
on 1 and on all visible GPUs, with and without --fast_pickle, next to the reference's literal
loop (oracle restatement, on the same GPU as the reference places its tensors) extrapolated from
a sample.  Run under gpurun.

    python tools/bench_pipeline_stream.py [--records 400000] [--topnumber 5] [--literal-sample 2000]
"""
import argparse
import json
import os
import pickle
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch


def run_cli(src, dst, k, gpus, extra):
    if os.path.exists(dst):
        os.remove(dst)
    cmd = [sys.executable, "-m", "zsaac_b200.data_handing.embeddings_related_generator", "--input_path", src,
           "--output_path", dst, "--topnumber", str(k), "--gpus", str(gpus), *extra]
    env = dict(os.environ, ZSAAC_TIMING="1")
    t0 = time.perf_counter()
    res = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, env=env)
    wall = time.perf_counter() - t0
    if res.returncode != 0:
        return {"error": (res.stdout + res.stderr)[-1500:]}
    timing = {}
    for line in res.stdout.splitlines():
        if line.startswith('{"zsaac_timing"'):
            timing = json.loads(line)["zsaac_timing"]
    timing["wall_s_incl_interpreter_start"] = round(wall, 2)
    timing["output_MB"] = round(os.path.getsize(dst) / 1e6, 1)
    return timing


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=400_000)
    ap.add_argument("--topnumber", type=int, default=5)
    ap.add_argument("--literal-sample", type=int, default=2000)
    args = ap.parse_args()
    n, k = args.records, args.topnumber
    tmp = tempfile.mkdtemp()
    src = os.path.join(tmp, "data.pkl")
    g = torch.Generator().manual_seed(1)
    recs = []
    for lo in range(0, n, 50_000):
        emb = torch.randn(min(50_000, n - lo), 1024, generator=g)
        recs += [{"caption": f"synthetic caption number {lo + i} with a few more words in it", "text_id": lo + i,
                  "text_embedding": emb[i:i + 1].clone()} for i in range(emb.shape[0])]
    with open(src, "wb") as f:
        pickle.dump(recs, f)
    out = {"records": n, "topnumber": k, "input_MB": round(os.path.getsize(src) / 1e6, 1),
           "cpu_count": os.cpu_count()}
    n_gpus = torch.cuda.device_count()
    dst = os.path.join(tmp, "out.pkl")
    out["gpus_1"] = run_cli(src, dst, k, 1, [])
    out["gpus_1_fast_pickle"] = run_cli(src, dst, k, 1, ["--fast_pickle"])
    if n_gpus > 1:
        out[f"gpus_{n_gpus}_exclude_self"] = run_cli(src, dst, k, n_gpus, ["--exclude_self"])
        out[f"gpus_{n_gpus}_exclude_self_fast_pickle"] = run_cli(src, dst, k, n_gpus, ["--exclude_self", "--fast_pickle"])

    # reference: the literal per-item loop on the same GPU, on a sample, extrapolated to n items
    from oracle import oracle
    sample = recs[:args.literal_sample]
    ref_bank = oracle.build_bank(recs).cuda()
    list(oracle.process_data_literal(ref_bank, recs[:50], k, device="cuda"))   # warm-up
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    list(oracle.process_data_literal(ref_bank, sample, k, device="cuda"))
    torch.cuda.synchronize()
    per_item = (time.perf_counter() - t4) / len(sample)
    t5 = time.perf_counter()
    with open(os.path.join(tmp, "ref_out.pkl"), "ab") as f:
        for it in sample:
            pickle.dump(it, f)
    per_item_save = (time.perf_counter() - t5) / len(sample)
    out["reference_literal_loop_on_gpu"] = {
        "sample": len(sample), "process_items_per_s": round(1 / per_item, 1),
        "process_extrapolated_s": round(per_item * n, 1), "save_extrapolated_s": round(per_item_save * n, 1)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
