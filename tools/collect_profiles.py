"""Copy the artefacts of tools/gpu_session.sh from gpurun_out/ into profiles/<round>/ and reduce
the ncu reports to small CSV summaries (run in the build container; needs `ncu` for -i only)."""
import csv
import glob
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out")
DST = os.path.join(ROOT, "profiles", sys.argv[1] if len(sys.argv) > 1 else "r01")
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
        with open(os.path.join(DST, f"ncu_source_hotspots_simtopk_{name}.csv"), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["sass_index", "sass", "samples", "warp_instructions_executed"])
            for i, r in enumerate(body):
                if int(r[isamp]) > 600 or any(t in r[ia] for t in ("TRYWAIT", "LDTM", "UTCHMMA", "UTMALDG", "UTCBAR", "UCGABAR")):
                    w.writerow([i, r[ia].strip(), r[isamp], r[iex]])
    i_rd, i_wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    mul = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}
    n = len(rows) - 2
    return (sum(float(r[i_rd]) for r in rows[2:]) * mul[units[i_rd]]
            + sum(float(r[i_wr]) for r in rows[2:]) * mul[units[i_wr]]) / n


for pat in ("bench_*.json", "bench_all.jsonl", "pytest_gpu.log", "smoke.log", "launches.csv",
            "ncu_dram_default.csv", "trace_small.log", "bench_memproj.jsonl", "pipeline.json",
            "dist_check_n*.log"):
    for f in glob.glob(os.path.join(SRC, pat)):
        shutil.copy(f, os.path.join(DST, os.path.basename(f)))

traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
rep = os.path.join(SRC, "prof_simtopk.ncu-rep")
if os.path.exists(rep):
    traffic["wavcaps_400k"] = summarise_full(rep, "wavcaps_400k")
rep = os.path.join(SRC, "prof_simtopk_default.ncu-rep")      # --set full of the default workload
if os.path.exists(rep):
    traffic["synthetic_10m"] = summarise_full(rep, "synthetic_10m")
dram = os.path.join(SRC, "ncu_dram_default.csv")
if os.path.exists(dram):
    vals = {}
    for r in csv.reader(open(dram)):
        if len(r) > 3 and r[-3].startswith("dram__bytes"):
            vals[r[-3]] = float(r[-1].replace(",", ""))
    if vals and "synthetic_10m" not in traffic:
        traffic["synthetic_10m"] = sum(vals.values())
json.dump(traffic, open(traffic_path, "w"), indent=1)
print(json.dumps({k: v for k, v in traffic.items() if not k.startswith("_")}))
