"""End-to-end timing of the related-caption generator script at a BASELINE shape, phase by phase,
next to the reference's literal per-item loop (restated in oracle/, run on the same GPU exactly as
the reference places its tensors).  Run under gpurun.

    python tools/bench_pipeline.py [--records 49838] [--topnumber 5] [--literal-sample 2000]
"""
import argparse
import json
import os
import pickle
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import zsaac_b200
from zsaac_b200.data_handing import embeddings_related_generator as gen
from oracle import oracle


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=49838)
    ap.add_argument("--topnumber", type=int, default=5)
    ap.add_argument("--literal-sample", type=int, default=2000)
    args = ap.parse_args()
    n, k = args.records, args.topnumber
    tmp = tempfile.mkdtemp()
    src = os.path.join(tmp, "data.pkl")
    g = torch.Generator().manual_seed(1)
    emb = torch.randn(n, 1024, generator=g)
    recs = [{"caption": f"synthetic caption number {i} with a few more words in it", "text_id": i,
             "text_embedding": emb[i:i + 1].clone()} for i in range(n)]
    with open(src, "wb") as f:
        pickle.dump(recs, f)
    del recs
    torch.cuda.init()
    out = {"records": n, "topnumber": k, "input_MB": round(os.path.getsize(src) / 1e6, 1)}

    # warm the library (context creation, first-launch costs) outside the timed run
    zsaac_b200.related_topk(torch.randn(256, 1024).cuda(), torch.randn(4096, 1024).cuda(), k)
    torch.cuda.synchronize()

    t0 = time.perf_counter()
    bank, all_data = gen.load_data(src)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    items = list(gen.process_data(bank, all_data, k))          # search + gather + D2H + per-item clones
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    dst = os.path.join(tmp, "out_related.pkl")
    gen.save_data_to_hdf5(iter(items), dst, len(items))
    t3 = time.perf_counter()
    out.update({"load_data_s": round(t1 - t0, 3), "process_data_s": round(t2 - t1, 3),
                "save_s": round(t3 - t2, 3), "total_s": round(t3 - t0, 3),
                "output_MB": round(os.path.getsize(dst) / 1e6, 1)})
    # opt-in multi-process writer (same bytes)
    procs = min(16, os.cpu_count() or 1)
    dst2 = os.path.join(tmp, "out_related_parallel.pkl")
    t6 = time.perf_counter()
    gen.save_data_to_hdf5(iter(items), dst2, len(items), workers=procs)
    t7 = time.perf_counter()
    out.update({"save_parallel_s": round(t7 - t6, 3), "writer_procs": procs,
                "parallel_output_identical": os.path.getsize(dst) == os.path.getsize(dst2)
                and open(dst, "rb").read(1 << 24) == open(dst2, "rb").read(1 << 24)})
    os.remove(dst2)
    # opt-in numpy-backed tensor pickling (same records after pickle.load, different bytes), and
    # what the reference's reader loop (dataset/dataset.py:64-78) pays for either file
    dst3 = os.path.join(tmp, "out_related_fast.pkl")
    t8 = time.perf_counter()
    gen.save_data_to_hdf5(iter(items), dst3, len(items), fast_pickle=True)
    t9 = time.perf_counter()

    def read_stream(path, limit):
        got = []
        with open(path, "rb") as f:
            while len(got) < limit:
                try:
                    got.append(pickle.load(f))
                except EOFError:
                    break
        return got

    t10 = time.perf_counter()
    back_plain = read_stream(dst, 10000)
    t11 = time.perf_counter()
    back_fast = read_stream(dst3, 10000)
    t12 = time.perf_counter()
    out.update({"save_fast_pickle_s": round(t9 - t8, 3),
                "read_back_10000_records_s": round(t11 - t10, 3),
                "read_back_10000_records_fast_pickle_s": round(t12 - t11, 3),
                "fast_pickle_records_equal": all(
                    torch.equal(a["related_embeddings"], b["related_embeddings"])
                    and torch.equal(a["text_embedding"], b["text_embedding"]) and a["caption"] == b["caption"]
                    for a, b in zip(back_plain, back_fast))})
    del back_plain, back_fast
    os.remove(dst3)

    # the search itself (what the kernel work amounts to inside process_data)
    rb = zsaac_b200.retrieval.bank_for(bank, normalize=True)
    q = torch.cat([it["text_embedding"] for it in items[:16384]]).cuda()
    for _ in range(3):
        rb.search(q, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rb.search(q, k)
    e1.record()
    torch.cuda.synchronize()
    out["search_16384_queries_ms"] = round(e0.elapsed_time(e1), 3)

    # reference: the literal per-item loop on the same GPU, on a sample, extrapolated to n items
    with open(src, "rb") as f:
        ref_data = pickle.load(f)
    sample = ref_data[:args.literal_sample]
    ref_bank = oracle.build_bank(ref_data).cuda()
    list(oracle.process_data_literal(ref_bank, ref_data[:50], k, device="cuda"))   # warm-up
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    ref_items = list(oracle.process_data_literal(ref_bank, sample, k, device="cuda"))
    torch.cuda.synchronize()
    t5 = time.perf_counter()
    per_item = (t5 - t4) / len(sample)
    out.update({"reference_loop_on_gpu_items_per_s": round(1 / per_item, 1),
                "reference_loop_on_gpu_extrapolated_s": round(per_item * n, 1),
                "reference_loop_sample": len(sample)})
    out["process_data_speedup_vs_reference_loop"] = round(per_item * n / (t2 - t1), 1)
    # same related rows as the reference loop?  (the two banks come from two fp32 normalisations
    # with different summation orders, so rows agree to an ulp or two, not bit for bit)
    same = sum(torch.allclose(a["related_embeddings"], b["related_embeddings"], rtol=0, atol=1e-6)
               for a, b in zip(items[:len(sample)], ref_items))
    out["same_related_embeddings_in_sample_atol_1e-6"] = f"{same}/{len(sample)}"
    print(json.dumps(out))


if __name__ == "__main__":
    main()
