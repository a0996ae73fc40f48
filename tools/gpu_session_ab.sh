#!/bin/bash
# 1-GPU box: engine tests, then A/B of the bench under environment switches (one line per variant)
mkdir -p gpurun_out
T=${1:-ab}
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -x -q -p no:cacheprovider --timeout 600 -W ignore::FutureWarning > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
grep -E "passed|failed|FAILED|Error|rc=" gpurun_out/${T}_pytest.log | tail -8
i=0
while [ $# -gt 1 ]; do
  shift; i=$((i+1))
  env $1 timeout 600 python bench.py --no-python-layer --no-cpu-baseline --no-render > gpurun_out/${T}_bench$i.log 2> gpurun_out/${T}_bench$i.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_bench$i.log").read().strip().splitlines()[-1])
    print("$1", round(d["value"]/1e6,2),"Mrays/s", round(d["ms_per_step"],4),"ms", round(d["samples_per_sec"]/1e6),"Msamp/s", d["kernel_us"], "e2e", round(d["e2e"]["value"]/1e6,2))
except Exception as e:
    print("$1", "ERR", e); print(open("gpurun_out/${T}_bench$i.err").read()[-1500:])
PY
done
