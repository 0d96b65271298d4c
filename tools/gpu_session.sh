#!/bin/bash
# One GPU session: parity tests, smoke, benches, then ncu evidence.  Logs under gpurun_out/.
# usage: tools/gpu_session.sh [tests] [bench] [ncu]
set -u
mkdir -p gpurun_out
ARGS=" ${*:-tests bench ncu} "

if [[ "$ARGS" == *" tests "* ]]; then
  timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
  tail -n 3 gpurun_out/pytest_gpu.log
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
  echo "smoke exit $?" >> gpurun_out/smoke.log
  tail -n 2 gpurun_out/smoke.log
fi

if [[ "$ARGS" == *" bench "* ]]; then
  for wl in wavcaps_400k audiocaps clotho_eval allpairs_400k; do
    timeout 900 python bench.py --workload $wl > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
    echo "bench $wl exit $?"; tail -c 700 gpurun_out/bench_$wl.json
  done
  timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
  echo "bench default exit $?"; tail -c 1500 gpurun_out/bench_default.json
  timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
  echo "bench reference exit $?"; tail -c 900 gpurun_out/bench_reference.json
  timeout 600 python tools/bench_all.py > gpurun_out/bench_all.jsonl 2> gpurun_out/bench_all.err
fi

if [[ "$ARGS" == *" ncu "* ]]; then
  CMD="python bench.py --workload wavcaps_400k --steps 2 --warmup 3 --no-cpu-baseline"
  $CMD > gpurun_out/ncu_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches exit $?"
  $CMD > gpurun_out/ncu_plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:zs_simtopk -s 3 -c 2 \
      -f -o gpurun_out/prof_simtopk $CMD > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?"
  CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
  $CMD > gpurun_out/ncu_plain3.log 2>&1 &&
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second \
      --clock-control none -k regex:zs_simtopk -s 3 -c 1 --csv --log-file gpurun_out/ncu_dram_default.csv $CMD > gpurun_out/ncu_dram.log 2>&1
  echo "ncu default dram exit $?"
fi
