#!/bin/bash
# Round-2 GPU pass Q: eval-side pipeline tests, default bench (with the ginfer leg), ginfer workload at B=64.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_eval.py -m gpu -q > gpurun_out/q_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/q_pytest.log
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/q_bench_train.json 2> gpurun_out/q_bench_train.err
timeout 600 python bench.py --workload ginfer --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/q_bench_ginfer.json 2> gpurun_out/q_bench_ginfer.err
echo done
