#!/bin/bash
# N-GPU box: exchange-kernel check on real symmetric memory, then the driver's own launch line for bench.py at N GPUs (fused NVLink exchange);
# optional third argument "nccl": the same bench with the three NCCL collectives for comparison
mkdir -p gpurun_out
T=${1:-sc}
N=${2:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29519 tools/dp_exchange_check.py > gpurun_out/${T}_check${N}.log 2>&1
echo "check rc=$?" >> gpurun_out/${T}_check${N}.log
timeout 600 $TR --master-port 29517 bench.py --gpus $N --steps 500 --warmup 64 > gpurun_out/${T}_bench${N}.log 2> gpurun_out/${T}_bench${N}.err
echo "bench$N rc=$?" >> gpurun_out/${T}_bench${N}.err
if [ "$3" = "nccl" ]; then
  MFN_DP_EXCHANGE=nccl timeout 600 $TR --master-port 29518 bench.py --gpus $N --steps 500 --warmup 64 --no-render > gpurun_out/${T}_nccl${N}.log 2> gpurun_out/${T}_nccl${N}.err
  echo "nccl$N rc=$?" >> gpurun_out/${T}_nccl${N}.err
fi
grep -v "^\[W\|Warning\|enable_symm\|^$\|^\*\*\*\|OMP_NUM" gpurun_out/${T}_check${N}.log | tail -8
python - <<PY
import json
for f in ("bench$N","nccl$N"):
    try:
        d=json.loads(open(f"gpurun_out/${T}_{f}.log").read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["value"]/1e6,2),"Mrays/s", round(d["ms_per_step"],4),"ms", round(d["samples_per_sec"]/1e6),"Msamp/s", d["samples_per_ray"], d["kernel_us"], "e2e", round(d["e2e"]["value"]/1e6,2), d.get("amp"), d["config"]["parallelism"], d["render"]["fps_800x800"] if d.get("render") else None)
    except Exception as e:
        print(f, "ERR", e)
PY
tail -n 3 gpurun_out/${T}_bench${N}.err
