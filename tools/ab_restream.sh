#!/bin/bash
# A/B for the DRAM re-streaming of the headline shape: the bank is streamed
# once per wave of 74 query-tile pairs (13 chunks: 92 GB of DRAM traffic per launch vs 20.6 GB
# algorithmic).  Smaller chunks stay in L2 across the waves of their query tiles and cut the
# traffic; do they buy clock under the power cap?  For each chunk count: bench timing + clocks,
# then a metrics-only ncu pass for the DRAM bytes.
mkdir -p gpurun_out
for CH in 13 52 156; do
  CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --headline-only"
  ZSAAC_CHUNKS=$CH $CMD > gpurun_out/ab_restream_chunks$CH.json 2> gpurun_out/ab_restream_chunks$CH.err &&
  ZSAAC_CHUNKS=$CH ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.avg.per_second \
      --clock-control none -k regex:zs_simtopk -s 4 -c 1 --csv --log-file gpurun_out/ab_restream_chunks$CH.csv $CMD > gpurun_out/ab_restream_ncu$CH.log 2>&1
  echo "chunks $CH exit $?"
  python - <<PY
import json, csv
d = json.loads(open("gpurun_out/ab_restream_chunks$CH.json").read().strip().splitlines()[-1])
m = {r[-3]: r[-1] for r in csv.reader(open("gpurun_out/ab_restream_chunks$CH.csv")) if len(r) > 3 and r[0].isdigit()}
print(json.dumps({"chunks": $CH, "plan": d["config"]["plan_chunks_tiles_ctas"], "ms_per_step": d["ms_per_step"],
                  "kernel_ms": d["roofline"]["kernel_ms"], "sm_mhz": d["clocks"]["sm_mhz"], "ncu": m}))
PY
done
