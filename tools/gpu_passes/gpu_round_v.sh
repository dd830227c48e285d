#!/bin/bash
# Round-2 GPU pass V: validation of the final state (full GPU suite, smoke, default bench with the driver's step counts).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/v_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/v_pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/v_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/v_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/v_bench_train.json 2> gpurun_out/v_bench_train.err
echo done
