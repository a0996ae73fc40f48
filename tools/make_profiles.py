#!/usr/bin/env python
"""profiles/ from the ncu captures in gpurun_out/: launch-list summary (share of the step per kernel), full-step summary, DRAM traffic.
usage: python tools/make_profiles.py  (after the two ncu commands quoted in profiles/launches_r01_summary.txt)"""
import collections
import csv
import json
import os
import shutil
import subprocess

import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
rows = list(csv.reader(l for l in open(f"gpurun_out/launches_{TAG}.csv") if l.startswith('"')))
hdr = rows[0]; kn = hdr.index("Kernel Name"); mv = hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[kn].split("(")[0].replace("void ", "").replace("mfn::", "")[:60]
    agg.setdefault(name, [0, 0.0]); agg[name][0] += 1; agg[name][1] += float(r[mv]) / 1e3
tot = sum(v[1] for v in agg.values())
with open(f"profiles/launches_{TAG}_summary.txt", "w") as f:
    f.write(f"ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 600 --csv --log-file gpurun_out/launches_{TAG}.csv \\\n"
            "    python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-render --no-python-layer\n")
    f.write("(first 600 launches: includes warm-up, density-grid updates and torch helper kernels; per-launch times are cold-cache and serialised,\n"
            " so the SHARE per kernel is what compares with bench.py's kernel_us, not the absolute)\n")
    f.write(f"{'kernel':62s} {'launches':>8s} {'total us':>10s} {'avg us':>8s} {'share':>7s}\n")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:62s} {c:8d} {t:10.1f} {t / c:8.1f} {100 * t / tot:6.1f}%\n")
shutil.copy(f"gpurun_out/launches_{TAG}.csv", f"profiles/launches_{TAG}.csv")
open(f"profiles/ncu_{TAG}_step_summary.txt", "w").write(
    "ncu --set full --clock-control none --import-source on --profile-from-start off python tools/prof_step.py   (one eager training step, ~600 k samples)\n" +
    subprocess.run(["python", "tools/ncu_summary.py", f"gpurun_out/prof_{TAG}_step.ncu-rep", "", "--stalls"], capture_output=True, text=True).stdout)
out = subprocess.run(["ncu", "-i", f"gpurun_out/prof_{TAG}_step.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr, units = rows[0], rows[1]
kn = hdr.index("Kernel Name"); rd = hdr.index("dram__bytes_read.sum"); wr = hdr.index("dram__bytes_write.sum"); du = hdr.index("gpu__time_duration.sum")
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
samples = int(open("gpurun_out/ncu_step.log").read().split("samples")[-1].split()[0])
names = {"field_fwd_fused": "field_fwd", "field_bwd_fused": "field_bwd", "grid_scatter_pair": "grid_encode_bwd", "march_count": "march_count", "march_scan_write": "march_write",
         "composite_train_fw": "composite_train_fw", "composite_train_bw": "composite_train_bw", "composite_loss_train": "composite_loss_train", "adam_kernel": "adam", "nerf_loss": "nerf_loss", "reduce_wgrad": "reduce_wgrad",
         "ray_setup": "ray_setup"}
res = {"_source": "ncu --set full --clock-control none, one eager training step of the bench workload (tools/prof_step.py)", "_samples": samples}
for r in rows[2:]:
    for k, v in names.items():
        if k in r[kn] and v not in res:
            b = float(r[rd]) * mult[units[rd]] + float(r[wr]) * mult[units[wr]]
            res[v] = {"dram_bytes": b, "dram_bytes_per_sample": b / samples, "duration_us": float(r[du]) if units[du] == "us" else float(r[du]) / 1e3}
json.dump(res, open("profiles/dram_traffic.json", "w"), indent=1)
json.dump(res, open(f"profiles/dram_traffic_{TAG}.json", "w"), indent=1)
print(open(f"profiles/launches_{TAG}_summary.txt").read())
print({k: (round(v["duration_us"], 1), round(v["dram_bytes_per_sample"], 1)) for k, v in res.items() if isinstance(v, dict)})
