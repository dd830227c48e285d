#!/bin/bash
# Round-2 GPU pass E: same-format precision codes (fp16-pair x2w16, fp16 x fp16 / bf16 x bf16 D), full suites, train-step bench with both G precisions.
mkdir -p gpurun_out
timeout 600 python tools/probe_formats.py > gpurun_out/e_probe.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_networks_wide.py -m gpu -q -s > gpurun_out/e_pytest_wide.log 2>&1; echo "rc=$?" >> gpurun_out/e_pytest_wide.log
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_networks_wide.py > gpurun_out/e_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/e_pytest_all.log
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/e_bench_train.json 2> gpurun_out/e_bench_train.err
GP3D_G_TERMS=2 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/e_bench_train_x2w16.json 2> gpurun_out/e_bench_train_x2w16.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/e_bench_ref.json 2> gpurun_out/e_bench_ref.err
echo done
