#!/bin/bash
# Second-half session of round 2: tests + smoke, the re-ordered bench, compute-sanitizer probe,
# chunk-count A/B of the headline (DRAM re-streaming).  Arguments select the parts.
set -u
mkdir -p gpurun_out
ARGS=" ${*:-tests bench sanitize restream} "
if [[ "$ARGS" == *" tests "* ]]; then
  timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log; tail -n 3 gpurun_out/pytest_gpu.log
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
  echo "smoke rc=$?" | tee -a gpurun_out/smoke.log; tail -n 2 gpurun_out/smoke.log
fi
if [[ "$ARGS" == *" bench "* ]]; then
  timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
  echo "bench default rc=$? after ${SECONDS}s"
  python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
print("headline", d["ms_per_step"], d["value"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["clocks"])
for w in d["workloads"]:
    r = w.get("roofline", {})
    print(" ", w["workload"][:46].ljust(46), round(w["ms_per_step"], 4), "kernel", round(r.get("kernel_ms", 0), 4),
          "frac", round(r.get("frac", 0), 3), "whole", round(w.get("search_frac_of_roofline", 0), 3),
          w.get("sm_mhz_after"), w.get("ms_min_median_max"), w.get("parity_gate", {}).get("ok"))
PY
fi
if [[ "$ARGS" == *" small "* ]]; then
  timeout 600 python tools/bench_small.py > gpurun_out/bench_small.jsonl 2> gpurun_out/bench_small.err
  echo "bench_small rc=$?"; cut -c1-200 gpurun_out/bench_small.jsonl
  for V in build_variants/*.so; do      # A/B of other builds of the library (one process each)
    [ -e "$V" ] || continue
    B=$(basename $V .so)
    ZSAAC_B200_LIB=$V timeout 600 python tools/bench_small.py > gpurun_out/bench_small_$B.jsonl 2> gpurun_out/bench_small_$B.err
    echo "bench_small $B rc=$?"; cut -c1-200 gpurun_out/bench_small_$B.jsonl
  done
fi
if [[ "$ARGS" == *" sanitize "* ]]; then
  timeout 200 python tools/sanitize_small.py > gpurun_out/sanitize_plain.log 2>&1; echo "plain rc=$?"
  # (compute-sanitizer is closed on this pool: "runs under it have left GPUs needing a reset")
fi
if [[ "$ARGS" == *" restream "* ]]; then
  for CH in 13 39; do
    ZSAAC_CHUNKS=$CH timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --headline-only \
        > gpurun_out/ab_restream_chunks$CH.json 2> gpurun_out/ab_restream_chunks$CH.err
    echo "chunks $CH rc=$?"
    python - <<PY
import json
d = json.loads(open("gpurun_out/ab_restream_chunks$CH.json").read().strip().splitlines()[-1])
print(json.dumps({"chunks": $CH, "plan": d["details"]["plan_chunks_tiles_ctas"], "ms_per_step": d["ms_per_step"],
                  "kernel_ms": d["roofline"]["kernel_ms"], "sm_mhz": d["clocks"]["sm_mhz"]}))
PY
  done
  ZSAAC_CHUNKS=39 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
      --clock-control none -k regex:zs_simtopk -s 4 -c 1 --csv --log-file gpurun_out/ab_restream_chunks39_ncu.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --headline-only > gpurun_out/ab_restream_ncu39.log 2>&1
  echo "ncu chunks 39 rc=$?"; tail -n 6 gpurun_out/ab_restream_chunks39_ncu.csv
fi
