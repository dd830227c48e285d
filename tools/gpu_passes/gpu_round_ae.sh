#!/bin/bash
# Round-2 GPU pass AE (final tree): ncu launch list (time + DRAM bytes) of one training step after the coalesced conv epilogue,
# ncu --set full of the ray-march backward at configs[2].
mkdir -p gpurun_out
timeout 420 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 30000 -c 12000 --csv --log-file gpurun_out/ae_launches_step.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-gpu-baseline --no-ginfer > gpurun_out/ae_ncu_bench.log 2>&1
echo "launch list rc=$?" >> gpurun_out/ae_ncu_bench.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:raymarch_bwd2 -s 1 -c 1 -o gpurun_out/ae_raymarch python bench.py --workload raymarch --steps 3 --warmup 3 > gpurun_out/ae_ncu_rm.log 2>&1
echo "rm rc=$?" >> gpurun_out/ae_ncu_rm.log
echo done
