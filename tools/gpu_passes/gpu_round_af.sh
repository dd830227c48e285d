#!/bin/bash
# Round-2 GPU pass AF: validation of the final tree (GPU suite, smoke, default train-step bench without the CPU leg) inside the last 6 GPU-minutes.
mkdir -p gpurun_out
timeout 170 python -m pytest tests -m gpu -q -x > gpurun_out/af_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/af_pytest_all.log
timeout 70 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/af_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/af_smoke.log
timeout 110 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-ginfer > gpurun_out/af_bench_train.json 2> gpurun_out/af_bench_train.err
echo done
