"""Host-only timing of the record writers (no GPU): records shaped like the generator's output
(audio_embedding [1,1024] shared by the 5 captions of a clip, text_embedding [1,1024],
related_embeddings [k,1024], caption / audio_id strings) through save_data_to_hdf5.

    python tools/bench_writer.py [--records 20000] [--k 5] [--out /tmp/zsaac_writer_bench.pkl]

Prints one JSON line: microseconds per record for the literal `pickle.dump` loop of the reference
(embeddings_related_generator.py:30-34), the default template writer (same bytes), --fast_pickle,
and the read-back time of the stream with pickle.load and with the library's fast unpickler.
"""
import argparse
import json
import os
import pickle
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zsaac_b200 import related_pipeline as rp  # noqa: E402

rp.tqdm = lambda it, total=None: it      # no progress bar in the timing


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=20000)
    ap.add_argument("--k", type=int, default=5)
    ap.add_argument("--out", default="/tmp/zsaac_writer_bench.pkl")
    args = ap.parse_args()
    g = torch.Generator().manual_seed(1)
    n, k = args.records, args.k
    audio = [torch.randn(1, 1024, generator=g) for _ in range(-(-n // 5))]
    items = [{"audio_embedding": audio[i // 5], "caption": f"synthetic caption number {i} of the benchmark set",
              "text_embedding": torch.randn(1, 1024, generator=g), "audio_id": f"Y{i:08d}.wav",
              "related_embeddings": torch.randn(k, 1024, generator=g)} for i in range(n)]
    out = {"records": n, "k": k, "host_cpus": os.cpu_count()}

    def timed(label, **kw):
        if os.path.exists(args.out):
            os.remove(args.out)
        t0 = time.perf_counter()
        rp.save_data_to_hdf5(iter(items), args.out, n, **kw)
        dt = time.perf_counter() - t0
        out[label + "_us_per_record"] = round(dt / n * 1e6, 2)
        return open(args.out, "rb").read()

    os.environ["ZSAAC_TEMPLATE_PICKLE"] = "0"
    literal = timed("literal_pickle_dump")
    os.environ["ZSAAC_TEMPLATE_PICKLE"] = "1"
    template = timed("template_writer")
    out["template_bytes_identical_to_literal"] = template == literal
    out["bytes_per_record"] = len(literal) // n
    del literal, template
    # read back what the default writer wrote (the reference's reader: dataset/dataset.py:64-78)
    t0 = time.perf_counter()
    with open(args.out, "rb") as f:
        for _ in range(n):
            pickle.load(f)
    out["read_pickle_load_us_per_record"] = round((time.perf_counter() - t0) / n * 1e6, 2)
    t0 = time.perf_counter()
    with open(args.out, "rb") as f:
        for _ in range(n):
            rp._FastTensorUnpickler(f).load()
    out["read_fast_unpickler_us_per_record"] = round((time.perf_counter() - t0) / n * 1e6, 2)
    timed("fast_pickle_writer", fast_pickle=True)
    os.remove(args.out)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
