#!/bin/bash
# one gpurun call: full GPU suite (no -x, every failure listed), default bench, A/B of the activation-saving modes
mkdir -p gpurun_out
T=${1:-s2}
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -W ignore::FutureWarning > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 900 python bench.py > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err
echo "bench rc=$?" >> gpurun_out/${T}_bench.err
MFN_FIELD_SAVE=full timeout 600 python bench.py --no-python-layer --no-cpu-baseline --no-render > gpurun_out/${T}_bench_full.log 2> gpurun_out/${T}_bench_full.err
timeout 600 python bench.py --no-python-layer --no-cpu-baseline --no-render > gpurun_out/${T}_bench_min.log 2> gpurun_out/${T}_bench_min.err
grep -E "passed|failed" gpurun_out/${T}_pytest.log | tail -3
python - <<PY
import json
for f in ("bench","bench_full","bench_min"):
    try:
        d=json.loads(open(f"gpurun_out/${T}_{f}.log").read().strip().splitlines()[-1])
        print(f, round(d["value"]/1e6,2),"Mrays/s", round(d["ms_per_step"],4),"ms", round(d["samples_per_sec"]/1e6),"Msamp/s", d["kernel_us"], {k:(round(v["value"]/1e6,3), round(v.get("samples_per_ray",0),1)) if v and "value" in v else v for k,v in d.items() if k in ("gpu_reference","frozen_api")})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/${T}_bench.err
