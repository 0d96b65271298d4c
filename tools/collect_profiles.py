"""Copy the artefacts of tools/gpu_session.sh from gpurun_out/ into profiles/<round>/ and reduce
the ncu reports / launch lists to small CSV summaries (run in the build container; `ncu -i` only).
usage: python tools/collect_profiles.py r02"""
import csv
import glob
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out")
DST = os.path.join(ROOT, "profiles", sys.argv[1] if len(sys.argv) > 1 else "r02")
os.makedirs(DST, exist_ok=True)

KEEP = ("gpu__time_duration", "dram__bytes", "gpu__dram_throughput", "sm__pipe_tensor", "sm__throughput",
        "lts__throughput", "lts__t_sector_hit", "lts__t_sectors_srcunit_tex_op_read.sum",
        "l1tex__throughput", "sm__cycles_elapsed", "sm__warps_active", "launch__",
        "smsp__inst_executed.sum", "Kernel Name")


def ncu_csv(report, page, extra=()):
    out = subprocess.run(["ncu", "-i", report, "--page", page, "--csv", *extra],
                         capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def summarise_full(report, name):
    """Metric table + per-instruction hot spots of a --set full capture; returns DRAM bytes per launch."""
    rows = ncu_csv(report, "raw")
    hdr, units = rows[0], rows[1]
    with open(os.path.join(DST, f"ncu_full_simtopk_{name}.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(rows) - 2)])
        for i, h in enumerate(hdr):
            if h.startswith(KEEP):
                w.writerow([h, units[i]] + [r[i] for r in rows[2:]])
    src = ncu_csv(report, "source", ("--print-source", "sass"))
    heads = [i for i, r in enumerate(src) if r and r[0] == "Address"]
    if heads:
        body = src[heads[0] + 1:(heads[1] - 1 if len(heads) > 1 else len(src))]
        h = src[heads[0]]
        ia, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
        total = sum(int(r[isamp]) for r in body)
        with open(os.path.join(DST, f"ncu_source_hotspots_simtopk_{name}.csv"), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["sass_index", "sass", "samples", "warp_instructions_executed", f"(total samples {total})"])
            for i, r in enumerate(body):
                if int(r[isamp]) > max(600, total // 400) or any(t in r[ia] for t in ("LDTM", "UTCHMMA", "UTMALDG", "UTCBAR")):
                    w.writerow([i, r[ia].strip(), r[isamp], r[iex]])
    i_rd, i_wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    mul = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}
    n = len(rows) - 2
    return (sum(float(r[i_rd]) for r in rows[2:]) * mul[units[i_rd]]
            + sum(float(r[i_wr]) for r in rows[2:]) * mul[units[i_wr]]) / n


def slim_launch_list(src, dst):
    """id, kernel, block, grid, unit, duration — drops ncu's constant columns."""
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and (r[0].isdigit() or r[0] == "ID")]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        for r in rows:
            w.writerow([r[0], r[4][:120], r[7], r[8], r[-2], r[-1]])


for pat in ("bench_*.json", "bench_*.jsonl", "pytest_gpu*.log", "smoke.log", "trace_small.log",
            "pipeline_stream_*.json", "exp_split_precision.json"):
    for f in glob.glob(os.path.join(SRC, pat)):
        name = "trace_per_cta_timeline.log" if os.path.basename(f) == "trace_small.log" else os.path.basename(f)
        shutil.copy(f, os.path.join(DST, name))
for f in glob.glob(os.path.join(SRC, "ncu_launches_*.csv")):
    slim_launch_list(f, os.path.join(DST, os.path.basename(f)))

traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
for rep, name in (("prof_simtopk_default.ncu-rep", "synthetic_10m"), ("prof_simtopk_audiocaps.ncu-rep", "audiocaps")):
    path = os.path.join(SRC, rep)
    if os.path.exists(path):
        traffic[name] = summarise_full(path, name)
json.dump(traffic, open(traffic_path, "w"), indent=1)
print(json.dumps({k: v for k, v in traffic.items() if not k.startswith("_")}))
