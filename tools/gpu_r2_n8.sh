#!/bin/bash
# 8-GPU session: real-NCCL tests, bench at N = 8, generator CLI on 1 vs 8 GPUs (100 k records)
mkdir -p gpurun_out
N=$(python -c "import torch; print(torch.cuda.device_count())")
echo "GPUs: $N"
python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_gpu_multi_n$N.log 2>&1; echo "pytest multi rc=$?" | tee -a gpurun_out/pytest_gpu_multi_n$N.log
tail -4 gpurun_out/pytest_gpu_multi_n$N.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_default_n$N.json 2> gpurun_out/bench_default_n$N.err ) 2>&1 | tail -3; echo "bench rc=$?"
tail -3 gpurun_out/bench_default_n$N.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_default_n$N.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"])
    print(d["parity_gate"])
    print(d["roofline"].get("kernel_ms_per_rank"), d["roofline"].get("step_tail_ms_beyond_slowest_kernel"), d["clocks"])
    for w in d.get("workloads", []):
        r = w.get("roofline", {})
        print(w["workload"][:70], round(w["ms_per_step"], 4), "ms", "frac", round(r.get("frac", 0), 3), "search_frac", round(w.get("search_frac_of_roofline", 0), 3),
              "gate", w.get("parity_gate", {}).get("ok"))
except Exception as e:
    print("parse failed", e)
PY
python tools/bench_pipeline_stream.py --records 100000 > gpurun_out/pipeline_stream_100k_n$N.json 2> gpurun_out/pipeline_stream.err; echo "pipeline rc=$?"
cat gpurun_out/pipeline_stream_100k_n$N.json; tail -3 gpurun_out/pipeline_stream.err
