#!/usr/bin/env python
"""profiles/sass_evidence.txt: counts of Blackwell-specific SASS instructions per kernel family of the built library (no GPU needed).
usage: python tools/sass_evidence.py"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mf-nerf_b200", "lib", "libmfnerf_b200.so")
FAMILIES = ["field_fwd_fused_kernel", "field_bwd_fused_kernel", "grid_scatter_pair_kernel", "composite_loss_train_kernel", "adam_kernel"]
PAT = re.compile(r"\b(UTCHMMA|UTCBAR|LDTM\.|UTCATOMSWS\.\w+|UBLKCP\.\w+\.\w+|SYNCS\.[\w.]+|REDG\.[\w.]+|RED\.[\w.]+|REDUX|LDG\.E\.CONSTANT|STG\.E\.\w*EL\w*|ATOMG\.[\w.]+)")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
counts = {f: collections.Counter() for f in FAMILIES}
instances = collections.Counter()
cur = None
for line in sass.splitlines():
    if "Function :" in line:
        cur = next((f for f in FAMILIES if f in line), None)
        if cur:
            instances[cur] += 1
        continue
    if cur:
        m = PAT.search(line)
        if m:
            counts[cur][m.group(1)] += 1
with open(os.path.join(ROOT, "profiles", "sass_evidence.txt"), "w") as f:
    f.write("SASS evidence (cuobjdump -sass mf-nerf_b200/lib/libmfnerf_b200.so, tools/sass_evidence.py), counts of Blackwell-specific instructions per kernel family\n"
            "(summed over the template instances of the family)\n"
            "UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (1-D TMA), SYNCS = mbarrier ops, REDG/RED = red.global, "
            "UTCATOMSWS = tcgen05.alloc / dealloc\n")
    for fam in FAMILIES:
        f.write(f"== {fam} ({instances[fam]} instance(s))\n")
        for k, v in counts[fam].most_common():
            f.write(f"{v:7d} {k}\n")
print(open(os.path.join(ROOT, "profiles", "sass_evidence.txt")).read())
