mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_nhwc_bf16_kernel -s 8 -c 2 -o gpurun_out/prof_conv_tc_persistent python tools/conv_microbench.py 4 > gpurun_out/ncu15a.log 2>&1
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 5000 -c 3000 --csv --log-file gpurun_out/launches_train_step.csv python bench.py --batch-gpu 8 --micro-batch 8 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu15b.log 2>&1
timeout 900 python tools/profile_step.py 16 gpurun_out/step_profile_b16_v5.txt > gpurun_out/prof15.log 2>&1
tail -3 gpurun_out/ncu15a.log; tail -2 gpurun_out/ncu15b.log | cut -c1-200; wc -l gpurun_out/launches_train_step.csv
