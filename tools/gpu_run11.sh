mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_render.py -m gpu -q --tb=line 2>&1 | tail -25) > gpurun_out/tests11.log
for m in 0 1 2; do (timeout 300 python bench.py --workload raymarch --mlp-mode $m --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench11_rm_mode$m.log; done
(timeout 300 python bench.py --workload raymarch --mlp-mode 2 --planes-fp16 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench11_rm_mode2_fp16.log
(timeout 300 python bench.py --workload raymarch --mlp-mode 1 --planes-fp16 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench11_rm_mode1_fp16.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:raymarch_fwd2 -s 3 -c 1 -o gpurun_out/prof_raymarch_fwd_v2 python bench.py --workload raymarch --mlp-mode 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu11.log 2>&1
cat gpurun_out/tests11.log; for f in gpurun_out/bench11_rm_*.log; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'])"; done
