#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/debug_graph.py > gpurun_out/u_debug_graph.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_eval.py tests/test_gpu_tc.py -m gpu -q -k "eval or weight_gradient or gradfix" > gpurun_out/u_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/u_pytest.log
timeout 300 python tools/conv_microbench.py 16 > gpurun_out/u_conv_microbench_b16.jsonl 2>&1
echo done
