#!/bin/bash
# Round-2 GPU pass P: pair-tile three-term conv (parity, microbench, step).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q > gpurun_out/p_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/p_pytest.log
timeout 600 python tools/conv_microbench.py 16 > gpurun_out/p_conv_microbench_b16.jsonl 2>&1
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/p_bench_n1.json 2> gpurun_out/p_bench_n1.err
echo done
