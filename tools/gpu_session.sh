#!/bin/bash
mkdir -p gpurun_out
T=${1:-s5}
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -W ignore::FutureWarning > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 900 python bench.py --no-python-layer --no-cpu-baseline > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err
echo "bench rc=$?" >> gpurun_out/${T}_bench.err
timeout 300 python tools/fwd_phases.py > gpurun_out/${T}_phases.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/${T}_pytest.log | tail -12
python - <<PY
import json
for f in ("bench",):
    try:
        d=json.loads(open(f"gpurun_out/${T}_{f}.log").read().strip().splitlines()[-1])
        print(f, round(d["value"]/1e6,2),"Mrays/s", round(d["ms_per_step"],4),"ms", round(d["samples_per_sec"]/1e6),"Msamp/s", d["kernel_us"], d["render"]["fps_800x800"])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/${T}_bench.err; tail -8 gpurun_out/${T}_phases.log
timeout 300 python tools/bwd_check.py > gpurun_out/${T}_bwdcheck.log 2>&1; MFN_FIELD_IMPL=v1 timeout 300 python tools/bwd_check.py > gpurun_out/${T}_bwdcheck_v1.log 2>&1
cat gpurun_out/${T}_bwdcheck.log gpurun_out/${T}_bwdcheck_v1.log | cut -c1-220
