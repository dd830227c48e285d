mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30) > gpurun_out/tests9.log
(timeout 600 python tools/conv_microbench.py 8 2>&1 | tail -12) > gpurun_out/conv_microbench.log
(timeout 1200 python bench.py --workload train_step --batch-gpu 16 --micro-batch 16 --steps 3 --warmup 3 2>&1 | tail -3) > gpurun_out/bench9_b16.log
(timeout 1200 python bench.py --workload train_step --batch-gpu 32 --micro-batch 16 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -3) > gpurun_out/bench9_b32.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_nhwc_bf16_kernel -s 2 -c 2 -o gpurun_out/prof_conv_tc python tools/conv_microbench.py 4 > gpurun_out/ncu_conv.log 2>&1
cat gpurun_out/tests9.log; cat gpurun_out/conv_microbench.log; cut -c1-330 gpurun_out/bench9_b16.log gpurun_out/bench9_b32.log
