#!/bin/bash
# Round-2 GPU pass J: 256-wide weight-gradient tiles (parity, step), full suite re-check.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/j_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/j_pytest_all.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/j_bench_n1.json 2> gpurun_out/j_bench_n1.err
timeout 600 python tools/profile_step.py 16 gpurun_out/j_step_profile_b16.txt > /dev/null 2>&1
echo done
