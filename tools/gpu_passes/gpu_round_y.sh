#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_nhwc_bf16_kernel -s 1 -c 1 -o gpurun_out/y_conv_b512 python tools/ncu_b512.py > gpurun_out/y_ncu.log 2>&1
echo done
