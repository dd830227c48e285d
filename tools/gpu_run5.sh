mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q --tb=short 2>&1 | tail -40) > gpurun_out/tests5.log
(timeout 600 python -m pytest tests/test_gpu_networks.py -m gpu -q --tb=short 2>&1 | tail -40) >> gpurun_out/tests5.log
(timeout 1200 python bench.py --workload train_step --batch-gpu 16 --micro-batch 16 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -3) > gpurun_out/bench5_b16.log
cat gpurun_out/tests5.log; cat gpurun_out/bench5_b16.log
