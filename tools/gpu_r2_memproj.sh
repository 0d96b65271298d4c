#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_metrics.py tests/test_gpu_parity.py -x -q -k "map2memory or score_matrix or topk_matches" > gpurun_out/pytest_memproj.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_memproj.log
python tools/bench_memproj.py > gpurun_out/bench_map2memory.jsonl 2> gpurun_out/bench_map2memory.err; echo "bench rc=$?"; cat gpurun_out/bench_map2memory.jsonl; tail -3 gpurun_out/bench_map2memory.err
