#!/bin/bash
# 8-GPU box: exchange-kernel check (both paths), lego bench: fused (multimem), fused (peer), NCCL; unbounded T=2^21: fused vs NCCL
mkdir -p gpurun_out
T=${1:-dp8}
N=${2:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29519 tools/dp_exchange_check.py > gpurun_out/${T}_check.log 2>&1
echo "check rc=$?" >> gpurun_out/${T}_check.log
run() { name=$1; shift; timeout 600 $TR --master-port $PORT bench.py --gpus $N "$@" > gpurun_out/${T}_${name}.log 2> gpurun_out/${T}_${name}.err; echo "$name rc=$?" >> gpurun_out/${T}_${name}.err; PORT=$((PORT+1)); }
PORT=29530
MFN_DP_MULTICAST=1 run fused_mc --steps 500 --warmup 64
MFN_DP_MULTICAST=0 run fused_p2p --steps 500 --warmup 64 --no-render
MFN_DP_EXCHANGE=nccl run nccl --steps 500 --warmup 64 --no-render
run unb_fused --steps 200 --warmup 32 --config unbounded_T21 --no-render
MFN_DP_EXCHANGE=nccl run unb_nccl --steps 200 --warmup 32 --config unbounded_T21 --no-render
grep -v "^\[W\|Warning\|enable_symm\|^$\|^\*\*\*\|OMP_NUM" gpurun_out/${T}_check.log | tail -14
python - <<PY
import json
for f in ("fused_mc","fused_p2p","nccl","unb_fused","unb_nccl"):
    try:
        d=json.loads(open(f"gpurun_out/${T}_{f}.log").read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["value"]/1e6,2),"Mrays/s", round(d["ms_per_step"],4),"ms", round(d["samples_per_sec"]/1e6),"Msamp/s", d["samples_per_ray"], d["kernel_us"].get("adam"), "e2e", round(d["e2e"]["value"]/1e6,2), d.get("amp"), d["final_loss_terms"], d["render"]["fps_800x800"] if d.get("render") else None)
    except Exception as e:
        print(f, "ERR", e)
PY
tail -n 3 gpurun_out/${T}_fused_mc.err
