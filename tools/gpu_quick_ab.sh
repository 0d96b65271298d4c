#!/bin/bash
# Quick A/B of library tuning variables at the main shapes (default plan, first k).
# usage: tools/gpu_quick_ab.sh "VAR=a;VAR=b" [shapes...]
mkdir -p gpurun_out
export SWEEP_QUICK=1 SWEEP_VARIANTS="$1"; shift
timeout 600 python tools/sweep_chunks.py "$@" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: print(l.strip()); continue
    print(r['shape'], 'k=%d'%r['k'], r['variant'], r['plan'], r['search_ms'], r['kernel_ms'], r['kernel_tflops'])
"
