#!/bin/bash
# ncu --set full of one fused-kernel launch at config 3 for each value of a tuning variable.
# usage: tools/gpu_ncu_ab.sh VAR v1 v2 ...      (reports land in gpurun_out/ncu_ab_VAR_v.ncu-rep)
set -u
mkdir -p gpurun_out
VAR=$1; shift
CMD="python bench.py --workload wavcaps_400k --steps 2 --warmup 3 --no-cpu-baseline"
for v in "$@"; do
  export $VAR=$v
  $CMD > gpurun_out/ncu_ab_plain_${VAR}_$v.log 2>&1 || { echo "plain run failed for $VAR=$v"; continue; }
  grep -o '"kernel_ms": [0-9.]*' gpurun_out/ncu_ab_plain_${VAR}_$v.log
  ncu --set full --clock-control none --import-source on -k regex:zs_simtopk -s 3 -c 1 -f \
      -o gpurun_out/ncu_ab_${VAR}_$v $CMD > gpurun_out/ncu_ab_${VAR}_$v.log 2>&1
  echo "ncu $VAR=$v exit $?"
done
