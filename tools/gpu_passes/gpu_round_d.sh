#!/bin/bash
# Round-2 GPU pass D: operand-format probe (one process per combination), the suites without the mixed-format parametrisations.
mkdir -p gpurun_out
timeout 600 python tools/probe_formats.py > gpurun_out/d_probe.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_networks_wide.py -m gpu -q -s -k "not x2w16 and not fp16-class" > gpurun_out/d_pytest_wide.log 2>&1; echo "rc=$?" >> gpurun_out/d_pytest_wide.log
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_networks_wide.py > gpurun_out/d_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/d_pytest_all.log
echo done
