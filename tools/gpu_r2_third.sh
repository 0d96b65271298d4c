#!/bin/bash
mkdir -p gpurun_out
python tools/trace_small.py > gpurun_out/trace_small.log 2>&1; echo "trace rc=$?"
cat gpurun_out/trace_small.log
python -m pytest tests/test_gpu_round2.py -x -q -k "one_launch or neutral" 2>&1 | tail -3
