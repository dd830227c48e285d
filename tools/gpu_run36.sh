cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -6
timeout 300 python bench.py --workload ops --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ops_v10.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for c in d['config']['cases']: print(c)
print(d['value'])"
