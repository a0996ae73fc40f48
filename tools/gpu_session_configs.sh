#!/bin/bash
# one B200: every --config workload of bench.py (BASELINE.json configs 3-5 and MF-NeRF's script configurations)
mkdir -p gpurun_out
T=${1:-cfg}
: > gpurun_out/${T}_configs.log
for c in mf_synthetic hash_synthetic unbounded_T21 llff_distortion mf_unbounded_T22; do
  timeout 600 python bench.py --config $c --steps 200 --warmup 32 >> gpurun_out/${T}_configs.log 2> gpurun_out/${T}_$c.err
  echo "$c rc=$?" >> gpurun_out/${T}_rc.log
done
timeout 600 python bench.py --config llff_distortion --rays 131072 --log2-T 21 --steps 50 --warmup 8 --no-render >> gpurun_out/${T}_configs.log 2> gpurun_out/${T}_llff_big.err
echo "llff_big rc=$?" >> gpurun_out/${T}_rc.log
cat gpurun_out/${T}_rc.log
python - <<PY
import json
for l in open("gpurun_out/${T}_configs.log"):
    try:
        d=json.loads(l)
        print(d["config"]["name"], d["config"]["rays_per_step_per_gpu"], d["config"]["engine"].get("log2_T"), round(d["value"]/1e6,3),"Mrays/s", round(d["ms_per_step"],3),"ms", round(d["samples_per_sec"]/1e6),"Msamp/s", round(d["samples_per_ray"],1), d["kernel_us"], d.get("amp"), [round(x,5) for x in d["final_loss_terms"]], round(d["render"]["fps_800x800"],1) if d.get("render") else None)
    except Exception as e:
        print("ERR", e, l[:200])
PY
tail -n 5 gpurun_out/${T}_*.err | tail -30
