cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tc.py tests/test_gpu_networks.py -m gpu -x -q --tb=short 2>&1 | tail -8
timeout 300 python bench.py --workload ops --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ops_v1.json
timeout 300 python tools/conv_microbench.py 2>&1 | tail -8 | tee gpurun_out/conv_microbench_v3.txt
