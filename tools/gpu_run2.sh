mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_render.py -m gpu -q --tb=line 2>&1 | tail -15) > gpurun_out/tests2.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1_raymarch.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:raymarch_fwd -s 3 -c 1 -o gpurun_out/prof_raymarch_fwd_v1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/tests2.log; tail -3 gpurun_out/ncu_full.log
