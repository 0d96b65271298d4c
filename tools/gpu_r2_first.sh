#!/bin/bash
# round 2, first GPU session: full GPU test suite + small-shape timings with variants
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python tools/bench_small.py --variants > gpurun_out/bench_small.jsonl 2> gpurun_out/bench_small.err; echo "bench_small rc=$?"
tail -3 gpurun_out/bench_small.err
cat gpurun_out/bench_small.jsonl | cut -c1-260
