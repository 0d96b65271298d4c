#!/bin/bash
# multi-GPU session: real-NCCL tests + bench at N = all visible GPUs
mkdir -p gpurun_out
N=$(python -c "import torch; print(torch.cuda.device_count())")
echo "GPUs: $N"
python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_gpu_multi_n$N.log 2>&1; echo "pytest multi rc=$?" | tee -a gpurun_out/pytest_gpu_multi_n$N.log
tail -25 gpurun_out/pytest_gpu_multi_n$N.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench_default_n$N.json 2> gpurun_out/bench_default_n$N.err ) 2>&1 | tail -3; echo "bench rc=$?"
tail -5 gpurun_out/bench_default_n$N.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_default_n$N.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"])
    print(d["parity_gate"])
    print(d["roofline"].get("kernel_ms_per_rank"), d["roofline"].get("step_tail_ms_beyond_slowest_kernel"))
    for w in d.get("workloads", []):
        r = w.get("roofline", {})
        print(w["workload"], round(w["ms_per_step"], 4), "ms", "frac", round(r.get("frac", 0), 3), "search_frac", round(w.get("search_frac_of_roofline", 0), 3),
              "launches", w.get("launches_per_search"), "gate", w.get("parity_gate", {}).get("ok"), w.get("parity_gate", {}).get("ranks"))
except Exception as e:
    print("parse failed", e)
PY
