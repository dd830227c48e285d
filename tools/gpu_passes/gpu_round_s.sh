#!/bin/bash
# Round-2 GPU pass S: validation of the shipped state -- full GPU suite, smoke(), default bench + reference arm, ray-march bench,
# ncu launch list (time + DRAM bytes) of one training step, ncu --set full of the wide-tile conv / weight-gradient kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/s_pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/s_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/s_bench_train.json 2> gpurun_out/s_bench_train.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s_bench_ref.json 2> gpurun_out/s_bench_ref.err
timeout 300 python bench.py --workload raymarch --steps 20 --warmup 3 > gpurun_out/s_bench_rm.json 2> gpurun_out/s_bench_rm.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 30000 -c 12000 --csv --log-file gpurun_out/s_launches_step.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-gpu-baseline --no-ginfer > gpurun_out/s_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_nhwc_bf16_kernel|wgrad_kernel" -c 3 -o gpurun_out/s_conv_wide python tools/ncu_kernels.py > gpurun_out/s_ncu_kernels.log 2>&1
echo done
