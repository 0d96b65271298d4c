#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python tools/bench_small.py > gpurun_out/bench_small.jsonl 2> gpurun_out/bench_small.err; echo "bench_small rc=$?"
cat gpurun_out/bench_small.jsonl | cut -c1-250
python tools/trace_small.py 2>&1 | head -40
