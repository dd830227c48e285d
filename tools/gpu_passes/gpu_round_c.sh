#!/bin/bash
# Round-2 GPU pass C: mixed-format MMA diagnostic, precision modes through the wide-network harness, full GPU suite.
mkdir -p gpurun_out
timeout 300 python tools/diag_precision.py > gpurun_out/c_diag.txt 2>&1; echo "rc=$?" >> gpurun_out/c_diag.txt
timeout 900 python -m pytest tests/test_gpu_networks_wide.py -m gpu -q -s > gpurun_out/c_pytest_wide.log 2>&1; echo "rc=$?" >> gpurun_out/c_pytest_wide.log
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_networks_wide.py > gpurun_out/c_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/c_pytest_all.log
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/c_bench_train.json 2> gpurun_out/c_bench_train.err
GP3D_G_TERMS=2 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/c_bench_train_x2w16.json 2> gpurun_out/c_bench_train_x2w16.err
echo done
