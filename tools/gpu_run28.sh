cd /root/repo
mkdir -p gpurun_out
timeout 300 python bench.py --workload ops --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ops_v1.json
