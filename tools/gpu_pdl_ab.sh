#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do for pdl in 1 0; do for wl in clotho_eval audiocaps; do
  ZSAAC_PDL=$pdl python bench.py --workload $wl --no-cpu-baseline --steps 20 > gpurun_out/b.json 2>gpurun_out/err.log
  python -c "import json;d=json.loads(open('gpurun_out/b.json').read().strip().splitlines()[-1]);print('pdl=$pdl', '$wl', 'ms/step', round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['ms_per_step'],4))"
done; done; done
