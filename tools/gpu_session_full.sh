#!/bin/bash
# 1-GPU box: the GPU suite, then the driver's own bench invocation (all legs), then the phase stamps
mkdir -p gpurun_out
T=${1:-f1}
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout 600 -W ignore::FutureWarning > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 900 python bench.py > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err
echo "bench rc=$?" >> gpurun_out/${T}_bench.err
timeout 300 python tools/fwd_phases.py > gpurun_out/${T}_phases.log 2>&1
grep -E "passed|failed|FAILED|rc=" gpurun_out/${T}_pytest.log | tail -12
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench.log").read().strip().splitlines()[-1])
print(round(d["value"]/1e6,2),"Mrays/s", round(d["ms_per_step"],4),"ms", round(d["samples_per_sec"]/1e6),"Msamp/s", d["kernel_us"], d.get("render"))
print({k:d.get(k) for k in ("e2e","roofline","l2_roofline","cpu_baseline","gpu_reference","frozen_api","amp","clocks")})
PY
tail -3 gpurun_out/${T}_bench.err; tail -8 gpurun_out/${T}_phases.log
