#!/bin/bash
# Round-2 GPU pass K: TMA-staged NCHW FIR (parity + ops bench), ray-march factor probe.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q > gpurun_out/k_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/k_pytest.log
timeout 600 python bench.py --workload ops --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/k_bench_ops.json 2> gpurun_out/k_bench_ops.err
timeout 600 python tools/rm_probe.py > gpurun_out/k_rm_probe.txt 2>&1
echo done
