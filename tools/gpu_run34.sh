cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/profile_dreg.py 16 gpurun_out/dreg_profile_b16.txt 2>&1 | grep -v Warn | tail -30
cut -c1-100,190-330 gpurun_out/dreg_profile_b16.txt | head -40
