#!/bin/bash
# Lock-step on/off comparison on the default workload: throughput, then DRAM bytes per launch.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -n 3 gpurun_out/pytest_gpu.log
for mode in 1 0; do
  ZSAAC_LOCKSTEP=$mode timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_lockstep$mode.json 2> gpurun_out/bench_lockstep$mode.err
  echo "lockstep=$mode: $(python -c "import json;d=json.loads(open('gpurun_out/bench_lockstep$mode.json').read().strip().splitlines()[-1]);print(d['ms_per_step'], d['value'], d['roofline']['achieved'], d['clocks'])")"
done
for mode in 1 0; do
  CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
  ZSAAC_LOCKSTEP=$mode $CMD > gpurun_out/ncu_ls_plain.log 2>&1 &&
  ZSAAC_LOCKSTEP=$mode ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
      --clock-control none -k regex:zs_simtopk -s 3 -c 1 --csv --log-file gpurun_out/ncu_dram_lockstep$mode.csv $CMD > gpurun_out/ncu_ls.log 2>&1
  echo "ncu lockstep=$mode exit $?"; grep -E "dram__|duration|hit_rate|tensor" gpurun_out/ncu_dram_lockstep$mode.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
