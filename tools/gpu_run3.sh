mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_networks.py -m gpu -q --tb=short 2>&1 | tail -60) > gpurun_out/tests3.log
(timeout 600 python bench.py --workload train_step --small --batch-gpu 8 --micro-batch 8 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -15) > gpurun_out/bench_small.log
(timeout 1200 python bench.py --workload train_step --batch-gpu 8 --micro-batch 8 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -15) > gpurun_out/bench_b8.log
(timeout 1200 python bench.py --workload train_step --batch-gpu 16 --micro-batch 16 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -15) > gpurun_out/bench_b16.log
nvidia-smi --query-gpu=memory.used,memory.total --format=csv > gpurun_out/mem.txt
tail -40 gpurun_out/tests3.log; cat gpurun_out/bench_small.log gpurun_out/bench_b8.log gpurun_out/bench_b16.log
