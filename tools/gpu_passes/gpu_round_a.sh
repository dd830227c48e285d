#!/bin/bash
# Round-2 GPU pass A: parity suite, short bench (train step + arms), conv DRAM-traffic capture.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/a_gpu.txt 2>&1; nproc >> gpurun_out/a_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "wide or ema or baseline_geometry" > gpurun_out/a_pytest_new.log 2>&1; echo "rc=$?" >> gpurun_out/a_pytest_new.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_networks_wide.py > gpurun_out/a_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/a_pytest_all.log
timeout 600 python bench.py --steps 6 --warmup 3 > gpurun_out/a_bench_train.json 2> gpurun_out/a_bench_train.err; echo "rc=$?" >> gpurun_out/a_bench_train.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_nhwc_bf16_kernel -s 150 -c 60 --csv --log-file gpurun_out/a_conv_traffic.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-gpu-baseline > gpurun_out/a_ncu_bench.log 2>&1
echo done
