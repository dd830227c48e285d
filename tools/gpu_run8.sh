mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q --tb=short 2>&1 | tail -40) > gpurun_out/tests8.log
(timeout 600 python -m pytest tests/test_gpu_networks.py -m gpu -q --tb=short 2>&1 | tail -40) >> gpurun_out/tests8.log
timeout 900 python tools/profile_step.py 8 gpurun_out/step_profile_b8_v3.txt > gpurun_out/prof8.log 2>&1
(timeout 1200 python bench.py --workload train_step --batch-gpu 16 --micro-batch 16 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -3) > gpurun_out/bench8_b16.log
cat gpurun_out/tests8.log; tail -45 gpurun_out/prof8.log | cut -c1-100,160-250; cat gpurun_out/bench8_b16.log | cut -c1-300
