#!/bin/bash
# Round-2 GPU pass G: fused filtered_lrelu (parity, ops bench).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "filtered" > gpurun_out/g_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/g_pytest.log
timeout 600 python bench.py --workload ops --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/g_bench_ops.json 2> gpurun_out/g_bench_ops.err
echo done
