#!/bin/bash
# N-GPU box: GPU suite (1 GPU), exchange-kernel check on real symmetric memory, bench with the fused exchange vs the NCCL collectives
mkdir -p gpurun_out
T=${1:-dp}
N=${2:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -W ignore::FutureWarning > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 180 $TR --master-port 29519 tools/dp_exchange_check.py > gpurun_out/${T}_check.log 2>&1
echo "check rc=$?" >> gpurun_out/${T}_check.log
timeout 600 $TR --master-port 29517 bench.py --gpus $N --steps 500 --warmup 64 --no-render > gpurun_out/${T}_fused.log 2> gpurun_out/${T}_fused.err
echo "fused rc=$?" >> gpurun_out/${T}_fused.err
MFN_DP_EXCHANGE=nccl timeout 600 $TR --master-port 29518 bench.py --gpus $N --steps 500 --warmup 64 --no-render > gpurun_out/${T}_nccl.log 2> gpurun_out/${T}_nccl.err
echo "nccl rc=$?" >> gpurun_out/${T}_nccl.err
grep -E "passed|failed|FAILED" gpurun_out/${T}_pytest.log | tail -6
grep -v "^\[W\|Warning\|^$" gpurun_out/${T}_check.log | tail -14
python - <<PY
import json
for f in ("fused","nccl"):
    try:
        d=json.loads(open(f"gpurun_out/${T}_{f}.log").read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["value"]/1e6,2),"Mrays/s", round(d["ms_per_step"],4),"ms", round(d["samples_per_sec"]/1e6),"Msamp/s", d["kernel_us"], "e2e", round(d["e2e"]["value"]/1e6,2), d.get("amp"), d["final_loss_terms"])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -n 4 gpurun_out/${T}_fused.err
