#!/bin/bash
mkdir -p gpurun_out
python tools/exp_split_precision.py > gpurun_out/exp_split_precision.json 2> gpurun_out/exp_split.err; echo "exp rc=$?"; cat gpurun_out/exp_split_precision.json; tail -3 gpurun_out/exp_split.err
python tools/bench_pipeline_stream.py --records 400000 > gpurun_out/pipeline_stream_400k_n1.json 2> gpurun_out/pipeline_stream.err; echo "pipeline rc=$?"
cat gpurun_out/pipeline_stream_400k_n1.json; tail -3 gpurun_out/pipeline_stream.err
