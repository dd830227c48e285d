cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -8
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_b32_v9.json
timeout 300 python bench.py --workload raymarch --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_raymarch_v9.json
