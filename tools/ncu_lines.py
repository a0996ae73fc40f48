#!/usr/bin/env python
"""Per-source-line stall samples of one kernel in an .ncu-rep (needs --import-source on and -lineinfo).
usage: python tools/ncu_lines.py rep kernel-regex [top_n]"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
recs = []
fname = ""
seen_files = set()
for r in rows:
    if r and r[0] in ("File Path", "File Name"):
        fname = r[1].split("/")[-1]
        if fname in seen_files:
            break          # second launch of the same kernel: first launch only
        seen_files.add(fname)
        continue
    if r and r[0] == "Line No":
        hdr = r
        ix0_samples = hdr.index("# Samples")
        continue
    if hdr is None or len(r) < len(hdr) // 2 or not r[0].strip():
        continue          # SASS rows have an empty line number; the source rows carry the roll-up
    try:
        int(r[ix0_samples] or 0)
    except (ValueError, IndexError):
        continue          # a source line whose quoting confused the CSV reader
    r[1] = fname[:14] + ": " + r[1].strip()
    recs.append(r)
ix = {h: i for i, h in enumerate(hdr)}
si = ix["# Samples"]; ii = ix["Instructions Executed"]
tot = sum(int(r[si] or 0) for r in recs)
toti = sum(int(r[ii] or 0) for r in recs)
print(f"total samples {tot}, warp instructions {toti}")
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in sorted(recs, key=lambda r: -int(r[si] or 0))[:top]:
    st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print(f"{int(r[si] or 0):7d} {100.0 * int(r[si] or 0) / max(tot, 1):5.1f}%  inst {int(r[ii] or 0):9d}  L{r[0]:>4s}  {r[1].strip()[:90]:90s} {st}")
