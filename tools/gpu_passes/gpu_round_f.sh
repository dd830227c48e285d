#!/bin/bash
# Round-2 GPU pass F: test re-check + kernel-time breakdown of the B=16 step (torch.profiler) for the round-2 step profile.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_render.py tests/test_gpu_networks_wide.py -m gpu -q > gpurun_out/f_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/f_pytest.log
timeout 600 python tools/profile_step.py 16 gpurun_out/f_step_profile_b16.txt > /dev/null 2>&1
timeout 600 python tools/profile_ops.py 16 > gpurun_out/f_ops_by_shape.txt 2>&1
echo done
