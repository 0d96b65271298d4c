mkdir -p gpurun_out
./build_variants/cluster_occupancy > gpurun_out/cluster_occupancy.txt 2>&1; cat gpurun_out/cluster_occupancy.txt
export SWEEP_QUICK=1
cp zero-shot-aac_b200/lib/libzsaac_b200.so /tmp/lib_main.so
for v in main:0 main:8 res6:8 res4:8 st4:0 st3:0 main:0; do
  n=${v%%:*}; r=${v#*:}
  if [ $n = main ]; then cp /tmp/lib_main.so zero-shot-aac_b200/lib/libzsaac_b200.so; else cp build_variants/lib_$n.so zero-shot-aac_b200/lib/libzsaac_b200.so; fi
  echo "== lib $n ZSAAC_RES=$r"
  SWEEP_VARIANTS="ZSAAC_RES=$r" timeout 300 python tools/sweep_chunks.py wavcaps shard k32 audiocaps 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: print(l.strip()); continue
    print(r['shape'], 'k=%d'%r['k'], r['plan'], r['search_ms'], r['kernel_ms'], r['kernel_tflops'])
"
done 2>&1 | tee gpurun_out/variants_ring.txt
cp /tmp/lib_main.so zero-shot-aac_b200/lib/libzsaac_b200.so
