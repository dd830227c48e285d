#!/bin/bash
# Round-2 GPU pass M: TMA box-shape probe (one config per process), ray-march MLP arithmetic through the wide generator.
mkdir -p gpurun_out
P=tools/_build/tma_probe
{
for cfg in "3 4 32 18 48 40" "3 4 64 18 48 40" "3 4 72 18 48 40" "3 4 72 18 256 128" "3 2 72 18 48 40" "3 2 64 18 48 40" "3 2 128 18 256 40" "3 2 136 66 264 70" "4 4 72 18 48 40" "4 2 72 18 48 40" "3 4 72 16 48 40" "3 4 80 18 48 40" "3 4 96 18 48 40" "3 4 128 18 256 40" "3 2 256 8 512 40"; do
  timeout 60 $P $cfg
done
} > gpurun_out/m_tma_probe.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_networks_wide.py -m gpu -q -s -k "mlp_arithmetic" > gpurun_out/m_pytest_mlp.log 2>&1; echo "rc=$?" >> gpurun_out/m_pytest_mlp.log
echo done
