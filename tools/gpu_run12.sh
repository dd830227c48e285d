mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_networks.py -m gpu -q --tb=short 2>&1 | tail -40) > gpurun_out/tests12.log
timeout 900 python tools/profile_step.py 8 gpurun_out/step_profile_b8_v4.txt > gpurun_out/prof12.log 2>&1
(timeout 1200 python bench.py --workload train_step --batch-gpu 16 --micro-batch 16 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -3) > gpurun_out/bench12_b16.log
cat gpurun_out/tests12.log; tail -40 gpurun_out/prof12.log | cut -c1-100,160-250; cut -c1-330 gpurun_out/bench12_b16.log
