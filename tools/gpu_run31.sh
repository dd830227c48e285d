cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -5
timeout 300 python bench.py --workload ops --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ops_v5.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for c in d['config']['cases']: print(c)
print(d['value'])"
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_b32_v13.json | cut -c1-300
timeout 600 python tools/profile_step.py 16 gpurun_out/step_profile_b16_v8.txt > /dev/null 2>&1
cut -c1-100,190-330 gpurun_out/step_profile_b16_v8.txt | head -64
