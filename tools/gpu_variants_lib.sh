#!/bin/bash
# A/B of library builds: build_variants/lib_<name>.so against the main library, on the default
# bench (config 4) and the quick shapes.  usage: tools/gpu_variants_lib.sh name1 name2 ...
mkdir -p gpurun_out
cp zero-shot-aac_b200/lib/libzsaac_b200.so /tmp/lib_main.so
for n in main "$@" main; do
  if [ $n = main ]; then cp /tmp/lib_main.so zero-shot-aac_b200/lib/libzsaac_b200.so; else cp build_variants/lib_$n.so zero-shot-aac_b200/lib/libzsaac_b200.so; fi
  echo "== lib $n"
  timeout 600 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read())
print('default', r['ms_per_step'], round(r['value']), round(r['roofline']['achieved'],1), r['clocks'])
"
  bash tools/gpu_quick_ab.sh "ZSAAC_SHARE_THR=1" wavcaps k32 shard audiocaps
done 2>&1 | tee gpurun_out/variants_lib.txt
cp /tmp/lib_main.so zero-shot-aac_b200/lib/libzsaac_b200.so
