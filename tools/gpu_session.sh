#!/bin/bash
# one gpurun call: full GPU suite (no -x, every failure listed), three-way step-path diagnosis, short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/s1_smi.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -W ignore::FutureWarning > gpurun_out/s1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s1_pytest.log
timeout 600 python tools/repro_threeway.py 4 > gpurun_out/s1_threeway.log 2>&1
echo "threeway rc=$?" >> gpurun_out/s1_threeway.log
timeout 600 python bench.py --steps 200 --warmup 16 > gpurun_out/s1_bench.log 2> gpurun_out/s1_bench.err
echo "bench rc=$?" >> gpurun_out/s1_bench.err
tail -5 gpurun_out/s1_pytest.log; tail -12 gpurun_out/s1_threeway.log; tail -2 gpurun_out/s1_bench.log
