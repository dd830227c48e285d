cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_b32_v16.json | cut -c1-330
