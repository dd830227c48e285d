#!/bin/bash
# Round-2 GPU pass Z: final tree -- full GPU suite, smoke(), short default bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/z_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/z_pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/z_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/z_bench_train.json 2> gpurun_out/z_bench_train.err
echo done
