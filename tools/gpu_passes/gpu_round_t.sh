#!/bin/bash
# Round-2 GPU pass T: split-K wave quantisation fix of the weight gradient (parity + step), CUDA-graph generator replay (parity + small-batch ginfer).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_eval.py tests/test_gpu_networks_wide.py -m gpu -q -k "not x2w16 and not bf16-" > gpurun_out/t_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/t_pytest.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-ginfer > gpurun_out/t_bench_n1.json 2> gpurun_out/t_bench_n1.err
timeout 300 python bench.py --workload ginfer --batch-gpu 4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/t_ginfer_b4_eager.json 2> gpurun_out/t_ginfer_b4_eager.err
timeout 300 python bench.py --workload ginfer --batch-gpu 4 --steps 20 --warmup 3 --no-cpu-baseline --graph > gpurun_out/t_ginfer_b4_graph.json 2> gpurun_out/t_ginfer_b4_graph.err
timeout 300 python tools/conv_microbench.py 16 > gpurun_out/t_conv_microbench_b16.jsonl 2>&1
echo done
