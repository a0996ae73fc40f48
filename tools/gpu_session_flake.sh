#!/bin/bash
# 1-GPU box: the GPU suite N times in a row (fresh process each time), to catch tolerance flakes before the driver does
mkdir -p gpurun_out
T=${1:-fl}
N=${2:-3}
for i in $(seq 1 $N); do
  timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout 600 -W ignore::FutureWarning > gpurun_out/${T}_pytest$i.log 2>&1
  echo "run $i rc=$? $(grep -E 'passed|failed' gpurun_out/${T}_pytest$i.log | tail -1)"
  grep -E "^FAILED|^E  " gpurun_out/${T}_pytest$i.log | head -12
done
