cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -15
timeout 600 python bench.py --batch-gpu 16 --micro-batch 16 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_b16_v6.json
timeout 600 python tools/profile_ops.py 16 > gpurun_out/ops_by_shape_v2.txt 2>&1; head -70 gpurun_out/ops_by_shape_v2.txt
