#!/bin/bash
mkdir -p gpurun_out
T=${1:-dpt}
N=${2:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29519 tools/dp_exchange_check.py > gpurun_out/${T}_check.log 2>&1
echo "check rc=$?" >> gpurun_out/${T}_check.log
PORT=29530
run() { name=$1; shift; timeout 600 $TR --master-port $PORT bench.py --gpus $N --steps 500 --warmup 64 --no-render "$@" > gpurun_out/${T}_${name}.log 2> gpurun_out/${T}_${name}.err; echo "$name rc=$?" >> gpurun_out/${T}_${name}.err; PORT=$((PORT+1)); }
MFN_DPX_CTAS=148 run c148
MFN_DPX_CTAS=74 run c74
MFN_DPX_CTAS=296 run c296
MFN_DP_EXCHANGE=nccl run nccl
grep -v "^\[W\|Warning\|enable_symm\|^$\|^\*\*\*\|OMP_NUM" gpurun_out/${T}_check.log | grep -E "world|exchange kernel|FAIL|rc="
python - <<PY
import json
for f in ("c148","c74","c296","nccl"):
    try:
        d=json.loads(open(f"gpurun_out/${T}_{f}.log").read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["value"]/1e6,2),"Mrays/s", round(d["ms_per_step"],4),"ms", round(d["samples_per_sec"]/1e6),"Msamp/s", d["kernel_us"].get("adam"), d["kernel_us"].get("march_count"), "e2e", round(d["e2e"]["value"]/1e6,2))
    except Exception as e:
        print(f, "ERR", e)
PY
