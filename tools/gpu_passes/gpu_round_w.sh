#!/bin/bash
# Round-2 GPU pass W: the four polyphase phases of the up-sampling convolutions in one launch (parity, microbench, step).
mkdir -p gpurun_out
timeout 300 python tools/upconv_microbench.py 16 > gpurun_out/w_upconv_microbench.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_networks_wide.py tests/test_gpu_networks.py -m gpu -q -k "not x2w16 and not bf16-" > gpurun_out/w_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/w_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-ginfer > gpurun_out/w_bench_n1.json 2> gpurun_out/w_bench_n1.err
echo done
