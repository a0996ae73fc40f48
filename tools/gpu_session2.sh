#!/bin/bash
# 2-GPU box: GPU suite + 1-GPU bench + 2-GPU bench (the driver's launch line) + phases
mkdir -p gpurun_out
T=${1:-m1}
N=${2:-2}
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -W ignore::FutureWarning > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 600 python bench.py --no-python-layer --no-cpu-baseline --steps 500 > gpurun_out/${T}_bench1.log 2> gpurun_out/${T}_bench1.err
echo "bench1 rc=$?" >> gpurun_out/${T}_bench1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 500 --warmup 64 > gpurun_out/${T}_bench${N}.log 2> gpurun_out/${T}_bench${N}.err
echo "bench$N rc=$?" >> gpurun_out/${T}_bench${N}.err
timeout 300 python tools/fwd_phases.py > gpurun_out/${T}_phases.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/${T}_pytest.log | tail -6
python - <<PY
import json
for f in ("bench1","bench$N"):
    try:
        d=json.loads(open(f"gpurun_out/${T}_{f}.log").read().strip().splitlines()[-1])
        print(f, round(d["value"]/1e6,2),"Mrays/s", round(d["ms_per_step"],4),"ms", round(d["samples_per_sec"]/1e6),"Msamp/s", d["kernel_us"], d["render"]["fps_800x800"] if d.get("render") else None, "e2e", round(d["e2e"]["value"]/1e6,2))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -4 gpurun_out/${T}_bench${N}.err; tail -7 gpurun_out/${T}_phases.log
