#!/bin/bash
# round 2 evidence session (1 GPU): default bench, then ncu launch lists and --set full captures
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference rc=$?"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/smoke.log
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --headline-only"
$CMD > gpurun_out/ncu_plain1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/ncu_launches_synthetic_10m.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches (default) rc=$?"
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:zs_simtopk -s 4 -c 1 \
    -f -o gpurun_out/prof_simtopk_default $CMD > gpurun_out/ncu_full_default.log 2>&1
echo "ncu full (default) rc=$?"
CMD="python tools/small_launches.py"
$CMD > gpurun_out/ncu_plain3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/ncu_launches_small_shapes.csv $CMD > gpurun_out/ncu_launches_small.log 2>&1
echo "ncu launches (small) rc=$?"
CMD="python bench.py --workload audiocaps --steps 2 --warmup 3 --no-cpu-baseline --headline-only"
$CMD > gpurun_out/ncu_plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:zs_simtopk -s 4 -c 1 \
    -f -o gpurun_out/prof_simtopk_audiocaps $CMD > gpurun_out/ncu_full_audiocaps.log 2>&1
echo "ncu full (audiocaps) rc=$?"
ls -la gpurun_out | tail -20
