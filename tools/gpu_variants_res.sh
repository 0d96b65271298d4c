#!/bin/bash
# Resident-query depth A/B: library variants built with -DZS_RES_KBLOCKS=4|6 next to the main
# library (8), on the default bench (config 4) and the quick shapes.
mkdir -p gpurun_out
cp zero-shot-aac_b200/lib/libzsaac_b200.so /tmp/lib_main.so
for v in main:0 res4:8 res6:8 main:8 main:0; do
  n=${v%%:*}; r=${v#*:}
  if [ $n = main ]; then cp /tmp/lib_main.so zero-shot-aac_b200/lib/libzsaac_b200.so; else cp build_variants/lib_$n.so zero-shot-aac_b200/lib/libzsaac_b200.so; fi
  echo "== lib $n ZSAAC_RES=$r"
  ZSAAC_RES=$r timeout 600 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read())
print('default', r['ms_per_step'], round(r['value']), round(r['roofline']['achieved'],1), r['clocks'])
"
  bash tools/gpu_quick_ab.sh "ZSAAC_RES=$r" wavcaps k32 shard
done 2>&1 | tee gpurun_out/variants_res.txt
cp /tmp/lib_main.so zero-shot-aac_b200/lib/libzsaac_b200.so
