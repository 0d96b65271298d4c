#!/bin/bash
# evidence session, part 2: GPU tests with the final library, smoke under ncu (as the driver runs
# it), launch list of the small shapes, --set full of the single-launch kernel at config 2
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
CMD="python __graft_entry__.py smoke"
$CMD > gpurun_out/smoke.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv \
    --log-file gpurun_out/ncu_launches_smoke.csv $CMD > gpurun_out/ncu_smoke.log 2>&1
echo "ncu smoke rc=$?"; tail -2 gpurun_out/ncu_smoke.log
CMD="python tools/small_launches.py"
$CMD > gpurun_out/ncu_plain3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/ncu_launches_small_shapes.csv $CMD > gpurun_out/ncu_launches_small.log 2>&1
echo "ncu launches (small) rc=$?"
CMD="python bench.py --workload audiocaps --steps 2 --warmup 3 --no-cpu-baseline --headline-only"
$CMD > gpurun_out/ncu_plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:zs_simtopk -s 4 -c 1 \
    -f -o gpurun_out/prof_simtopk_audiocaps $CMD > gpurun_out/ncu_full_audiocaps.log 2>&1
echo "ncu full (audiocaps) rc=$?"; tail -3 gpurun_out/ncu_full_audiocaps.log
python tools/bench_small.py > gpurun_out/bench_small.jsonl 2> gpurun_out/bench_small.err; echo "bench_small rc=$?"
cut -c1-200 gpurun_out/bench_small.jsonl
