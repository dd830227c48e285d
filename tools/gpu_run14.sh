mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q --tb=short -x 2>&1 | tail -15) > gpurun_out/tests14.log
(timeout 300 python tools/conv_microbench.py 8 2>&1 | tail -8) > gpurun_out/conv_microbench_v2.log
(timeout 900 python -m pytest tests/test_gpu_networks.py -m gpu -q --tb=short 2>&1 | tail -5) >> gpurun_out/tests14.log
(timeout 1200 python bench.py --workload train_step --batch-gpu 16 --micro-batch 16 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -2) > gpurun_out/bench14_b16.log
(timeout 900 python bench.py --workload ginfer --batch-gpu 8 --steps 3 --warmup 3 2>&1 | tail -2) > gpurun_out/bench14_ginfer.log
cat gpurun_out/tests14.log; python - <<'PY'
import json
for l in open('gpurun_out/conv_microbench_v2.log'):
    try:
        d = json.loads(l); print(d['layer'], round(d['ms_bf16x3'],3), round(d['frac_bf16x3'],3), round(d['ms_bf16'],3), round(d['frac_bf16'],3), d.get('ms_wgrad_bf16x3_incl_split'))
    except Exception as e: print(l[:200])
PY
cut -c1-330 gpurun_out/bench14_b16.log; cut -c1-600 gpurun_out/bench14_ginfer.log
