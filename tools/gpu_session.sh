#!/bin/bash
# one gpurun call: full GPU suite (no -x, every failure listed), default bench, phase stamps, L2 microbenchmark
mkdir -p gpurun_out
T=${1:-s3}
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -W ignore::FutureWarning > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 900 python bench.py > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err
echo "bench rc=$?" >> gpurun_out/${T}_bench.err
timeout 300 python tools/fwd_phases.py > gpurun_out/${T}_phases.log 2>&1
MFN_FIELD_SAVE=full timeout 300 python tools/fwd_phases.py > gpurun_out/${T}_phases_full.log 2>&1
timeout 300 tools/bin/l2_bench > gpurun_out/${T}_l2.json 2> gpurun_out/${T}_l2.err
grep -E "passed|failed" gpurun_out/${T}_pytest.log | tail -3
python - <<PY
import json
for f in ("bench",):
    try:
        d=json.loads(open(f"gpurun_out/${T}_{f}.log").read().strip().splitlines()[-1])
        print(f, round(d["value"]/1e6,2),"Mrays/s", round(d["ms_per_step"],4),"ms", round(d["samples_per_sec"]/1e6),"Msamp/s", d["kernel_us"], {k:(round(v["value"]/1e6,3), round(v.get("samples_per_ray",0),1)) if v and "value" in v else v for k,v in d.items() if k in ("gpu_reference","frozen_api")})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/${T}_bench.err; tail -8 gpurun_out/${T}_phases.log; cat gpurun_out/${T}_l2.json | head -40
