mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -15) > gpurun_out/tests17.log
(timeout 1200 python bench.py --workload train_step --batch-gpu 16 --micro-batch 16 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -2) > gpurun_out/bench17_b16.log
cat gpurun_out/tests17.log; cut -c1-330 gpurun_out/bench17_b16.log
