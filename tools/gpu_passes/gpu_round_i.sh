#!/bin/bash
# Round-2 GPU pass I: 256-wide three-term conv tiles (parity + microbench + step), rebalanced filtered_lrelu.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py -m gpu -q -k "wide or filtered or conv" > gpurun_out/i_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/i_pytest.log
timeout 600 python tools/conv_microbench.py 16 > gpurun_out/i_conv_microbench_b16.jsonl 2>&1
timeout 600 python bench.py --workload ops --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/i_bench_ops.json 2> gpurun_out/i_bench_ops.err
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/i_bench_n1.json 2> gpurun_out/i_bench_n1.err
echo done
