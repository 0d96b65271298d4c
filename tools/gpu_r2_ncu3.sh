#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py tests/test_gpu_metrics.py tests/test_gpu_pipeline.py -x -q > gpurun_out/pytest_gpu_part.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_part.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --headline-only"
$CMD > gpurun_out/ncu_plain1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv \
    --log-file gpurun_out/ncu_launches_synthetic_10m.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches (default) rc=$?"
CMD="python tools/small_launches.py"
$CMD > gpurun_out/ncu_plain3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/ncu_launches_small_shapes.csv $CMD > gpurun_out/ncu_launches_small.log 2>&1
echo "ncu launches (small) rc=$?"
python tools/bench_small.py 2>/dev/null | tail -2
python tools/trace_small.py > gpurun_out/trace_small.log 2>&1; echo "trace rc=$?"
