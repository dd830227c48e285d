#!/bin/bash
# Round-2 GPU pass N: TMA box-start probe (innermost coordinate alignment / sign), ray-march MLP arithmetic through the wide generator.
mkdir -p gpurun_out
P=tools/_build/tma_probe
{
for cfg in "3 4 72 18 48 40 0 0" "3 4 72 18 48 40 0 -1" "3 4 72 18 48 40 4 -1" "3 4 72 18 48 40 -4 -1" "3 4 72 18 48 40 -8 -1" "3 4 72 18 48 40 3 -1" "3 4 72 18 48 40 1 0" "3 2 72 18 48 40 -8 -1" "3 2 72 18 48 40 8 -1" "3 2 72 18 48 40 4 -1" "3 2 136 66 264 70 -8 -1" "3 2 72 18 48 40 -1 0" "3 4 32 18 48 40 32 -1" "3 4 72 18 256 128 252 126"; do
  timeout 60 $P $cfg
done
} > gpurun_out/n_tma_probe.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_networks_wide.py -m gpu -q -s -k "mlp_arithmetic" > gpurun_out/n_pytest_mlp.log 2>&1; echo "rc=$?" >> gpurun_out/n_pytest_mlp.log
echo done
