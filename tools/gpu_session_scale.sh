#!/bin/bash
# N-GPU box: the driver's launch line for bench.py at N GPUs, default workload + the unbounded T=2^21 workload (203 MB gradient exchange)
mkdir -p gpurun_out
T=${1:-sc}
N=${2:-4}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 500 --warmup 64 > gpurun_out/${T}_bench${N}.log 2> gpurun_out/${T}_bench${N}.err
echo "bench$N rc=$?" >> gpurun_out/${T}_bench${N}.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 200 --warmup 32 --config unbounded_T21 --no-render > gpurun_out/${T}_unb${N}.log 2> gpurun_out/${T}_unb${N}.err
echo "unb$N rc=$?" >> gpurun_out/${T}_unb${N}.err
timeout 600 python bench.py --steps 200 --warmup 32 --config unbounded_T21 --no-render > gpurun_out/${T}_unb1.log 2> gpurun_out/${T}_unb1.err
python - <<PY
import json
for f in ("bench$N","unb$N","unb1"):
    try:
        d=json.loads(open(f"gpurun_out/${T}_{f}.log").read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["value"]/1e6,2),"Mrays/s", round(d["ms_per_step"],4),"ms", round(d["samples_per_sec"]/1e6),"Msamp/s", d["kernel_us"], d["render"]["fps_800x800"] if d.get("render") else None, "e2e", round(d["e2e"]["value"]/1e6,2))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/${T}_bench${N}.err gpurun_out/${T}_unb${N}.err gpurun_out/${T}_unb1.err
