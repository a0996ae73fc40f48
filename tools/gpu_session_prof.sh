#!/bin/bash
# profiling call (1 GPU): launch list + one full ncu capture of an eager training step, each only after the same command exited 0 without ncu
mkdir -p gpurun_out
CMD="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-render --no-python-layer"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 600 --csv --log-file gpurun_out/launches_r02.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
python tools/prof_step.py > gpurun_out/prof_step_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_r02_step -f python tools/prof_step.py > gpurun_out/ncu_step.log 2>&1
echo "full capture rc=$?"
tail -2 gpurun_out/ncu_step.log; ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r02.csv
