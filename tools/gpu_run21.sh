cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_render.py tests/test_gpu_optim.py -m gpu -x -q --tb=short 2>&1 | tail -25
timeout 900 python -m pytest tests -m gpu -x -q --tb=short --deselect tests/test_gpu_render.py --deselect tests/test_gpu_optim.py 2>&1 | tail -8
timeout 600 python bench.py --batch-gpu 16 --micro-batch 16 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_b16_v7.json
timeout 600 python bench.py --batch-gpu 32 --micro-batch 32 --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_b32_mb32.json
timeout 600 python tools/profile_step.py 16 gpurun_out/step_profile_b16_v7.txt > /dev/null 2>&1
cut -c1-100,190-330 gpurun_out/step_profile_b16_v7.txt | head -30
