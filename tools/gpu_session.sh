#!/bin/bash
# One GPU session under gpurun; logs under gpurun_out/ (copy what should be kept to profiles/rNN/).
#   tools/gpu_session.sh tests        full GPU test suite + smoke
#   tools/gpu_session.sh bench        default bench (all shapes), reference arm, small shapes, map2memory
#   tools/gpu_session.sh ncu          launch lists + --set full captures (each after a plain run of the same command)
#   tools/gpu_session.sh trace        per-CTA timelines of the small shapes
#   tools/gpu_session.sh multi        (gpurun --gpus N) real-NCCL tests + bench at N + generator CLI on 1 vs N GPUs
#   tools/gpu_session.sh pipeline     generator script at 400,000 records, phase by phase
set -u
mkdir -p gpurun_out
ARGS=" ${*:-tests bench} "
N=$(python -c "import torch; print(torch.cuda.device_count())")

if [[ "$ARGS" == *" tests "* ]]; then
  timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log; tail -n 3 gpurun_out/pytest_gpu.log
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
  echo "smoke exit $?" | tee -a gpurun_out/smoke.log; tail -n 2 gpurun_out/smoke.log
fi

if [[ "$ARGS" == *" bench "* ]]; then
  timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
  echo "bench default exit $?"; tail -c 600 gpurun_out/bench_default.json
  timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
  echo "bench reference exit $?"
  timeout 600 python tools/bench_small.py > gpurun_out/bench_small.jsonl 2> gpurun_out/bench_small.err
  echo "bench_small exit $?"; cut -c1-220 gpurun_out/bench_small.jsonl
  timeout 600 python tools/bench_memproj.py > gpurun_out/bench_map2memory.jsonl 2> gpurun_out/bench_map2memory.err
  echo "bench_memproj exit $?"; cut -c1-260 gpurun_out/bench_map2memory.jsonl
  timeout 300 python tools/bench_writer.py > gpurun_out/writer_box_host_bench.json 2> /dev/null   # host only
  echo "bench_writer exit $?"; cat gpurun_out/writer_box_host_bench.json
fi

if [[ "$ARGS" == *" trace "* ]]; then
  python tools/trace_small.py > gpurun_out/trace_small.log 2>&1; echo "trace exit $?"; cat gpurun_out/trace_small.log
fi

if [[ "$ARGS" == *" ncu "* ]]; then
  CMD="python __graft_entry__.py smoke"
  $CMD > gpurun_out/smoke.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv \
      --log-file gpurun_out/ncu_launches_smoke.csv $CMD > gpurun_out/ncu_smoke.log 2>&1
  echo "ncu smoke exit $?"
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --headline-only"
  $CMD > gpurun_out/ncu_plain1.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv \
      --log-file gpurun_out/ncu_launches_synthetic_10m.csv $CMD > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches (default) exit $?"
  CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --headline-only"
  $CMD > gpurun_out/ncu_plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:zs_simtopk -s 4 -c 1 \
      -f -o gpurun_out/prof_simtopk_default $CMD > gpurun_out/ncu_full_default.log 2>&1
  echo "ncu full (default) exit $?"
  CMD="python tools/small_launches.py"
  $CMD > gpurun_out/ncu_plain3.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file gpurun_out/ncu_launches_small_shapes.csv $CMD > gpurun_out/ncu_launches_small.log 2>&1
  echo "ncu launches (small) exit $?"
  CMD="python bench.py --workload audiocaps --steps 2 --warmup 3 --no-cpu-baseline --headline-only"
  $CMD > gpurun_out/ncu_plain4.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:zs_simtopk -s 4 -c 1 \
      -f -o gpurun_out/prof_simtopk_audiocaps $CMD > gpurun_out/ncu_full_audiocaps.log 2>&1
  echo "ncu full (audiocaps) exit $?"
fi

if [[ "$ARGS" == *" multi "* ]]; then
  echo "GPUs: $N"
  python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_gpu_multi_n$N.log 2>&1
  echo "pytest multi exit $?" | tee -a gpurun_out/pytest_gpu_multi_n$N.log; tail -4 gpurun_out/pytest_gpu_multi_n$N.log
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_default_n$N.json 2> gpurun_out/bench_default_n$N.err
  echo "bench N=$N exit $?"; tail -c 400 gpurun_out/bench_default_n$N.json
  python tools/bench_pipeline_stream.py --records 100000 > gpurun_out/pipeline_stream_100k_n$N.json 2> gpurun_out/pipeline_stream.err
  echo "pipeline exit $?"; cat gpurun_out/pipeline_stream_100k_n$N.json
fi

if [[ "$ARGS" == *" pipeline "* ]]; then
  python tools/bench_pipeline_stream.py --records 400000 > gpurun_out/pipeline_stream_400k_n$N.json 2> gpurun_out/pipeline_stream.err
  echo "pipeline exit $?"; cat gpurun_out/pipeline_stream_400k_n$N.json
fi
