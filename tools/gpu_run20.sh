cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/profile_step.py 16 gpurun_out/step_profile_b16_v6.txt > /dev/null 2>&1
cut -c1-100,190-330 gpurun_out/step_profile_b16_v6.txt | head -60
