#!/bin/bash
# Round-2 GPU pass AD: ncu --set full of the conv kernels after the coalesced epilogue (wide three-term, weight gradient, single-term, toRGB).
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"conv_nhwc_bf16_kernel|wgrad_kernel" -c 4 -o gpurun_out/ad_conv_final python tools/ncu_kernels.py > gpurun_out/ad_ncu_kernels.log 2>&1
echo done
