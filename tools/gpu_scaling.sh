#!/bin/bash
# bench.py at N ranks (default workload) + NCCL parity check; logs under gpurun_out/
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" -gt 1 ]; then
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/dist_check.py > gpurun_out/dist_check_n$N.log 2>&1
  echo "dist_check n$N exit $?"; grep dist_check gpurun_out/dist_check_n$N.log
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
else
  timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
fi
echo "bench n$N exit $?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_n$N.json') if l.startswith('{')][-1])
print('n=$N', 'ms/step', round(d['ms_per_step'],2), 'q/s', round(d['value']), 'kernel TF', round(d['roofline']['achieved'],1), 'e2e', round(d['e2e']['value']), d['clocks'])
PY
