#!/bin/bash
# round 2, second GPU session: full GPU test suite, default bench (all workloads), small shapes
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
( time python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | tail -3; echo "bench rc=$?"
tail -5 gpurun_out/bench_default.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["e2e"]["value"], d["roofline"]["frac"], d["parity_gate"])
    for w in d.get("workloads", []):
        r = w.get("roofline", {})
        print(w["workload"], round(w["ms_per_step"], 4), "ms", "frac", round(r.get("frac", 0), 3), "search_frac", round(w.get("search_frac_of_roofline", 0), 3),
              "launches", w.get("launches_per_search"), "gate", w.get("parity_gate", {}).get("ok"), "strawman", w.get("strawman_torch_matmul_bf16_topk_ms"))
    print(d.get("literal_reference_loop_clotho_eval"))
    print(d.get("cpu_baseline"))
except Exception as e:
    print("parse failed", e)
PY
python tools/bench_small.py > gpurun_out/bench_small.jsonl 2> gpurun_out/bench_small.err; echo "bench_small rc=$?"
cat gpurun_out/bench_small.jsonl | cut -c1-250
