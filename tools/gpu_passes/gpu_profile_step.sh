cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/profile_step.py 16 gpurun_out/step_profile_b16_v9.txt > /dev/null 2>&1
cut -c1-100,190-330 gpurun_out/step_profile_b16_v9.txt | head -75
timeout 600 python tools/profile_ops.py 16 > gpurun_out/ops_by_shape_v3.txt 2>&1; grep -v Warn gpurun_out/ops_by_shape_v3.txt | head -32
