#!/bin/bash
# Round-2 GPU pass AC: coalesced (shared-memory transposed) conv epilogue -- parity, microbench, step.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_networks_wide.py tests/test_gpu_networks.py -m gpu -q -k "not x2w16 and not bf16-" > gpurun_out/ac_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/ac_pytest.log
timeout 300 python tools/conv_microbench.py 16 > gpurun_out/ac_conv_microbench_b16.jsonl 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-ginfer > gpurun_out/ac_bench_n1.json 2> gpurun_out/ac_bench_n1.err
echo done
