cd /root/repo
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -4
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/final_bench_reference.json | cut -c1-700
echo "== bench default"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/final_bench_default.json | cut -c1-2500
