mkdir -p gpurun_out
timeout 900 python tools/profile_step.py 8 gpurun_out/step_profile_b8.txt > gpurun_out/prof6.log 2>&1
tail -70 gpurun_out/prof6.log | cut -c1-250
