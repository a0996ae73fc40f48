#!/bin/bash
# 1-GPU box, quick loop: field / engine parity tests (or the tests named in $2), then the device-resident bench without the side legs
mkdir -p gpurun_out
T=${1:-q1}
TESTS=${2:-tests/test_field_gpu.py tests/test_engine_gpu.py}
timeout 900 python -m pytest $TESTS -m gpu -x -q -p no:cacheprovider --timeout 600 -W ignore::FutureWarning > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 600 python bench.py --no-python-layer --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err
echo "bench rc=$?" >> gpurun_out/${T}_bench.err
grep -E "passed|failed|FAILED|Error|rc=" gpurun_out/${T}_pytest.log | tail -8
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench.log").read().strip().splitlines()[-1])
print(round(d["value"]/1e6,2),"Mrays/s", round(d["ms_per_step"],4),"ms", round(d["samples_per_sec"]/1e6),"Msamp/s", d["kernel_us"], d["render"]["fps_800x800"] if d.get("render") else None, "e2e", round(d["e2e"]["value"]/1e6,2))
PY
tail -2 gpurun_out/${T}_bench.err
