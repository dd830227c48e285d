cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/rm_bwd_probe.py 0,1,2 5 2>&1 | tail -4 | tee gpurun_out/rm_bwd_probe.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:raymarch_bwd2 -c 1 -o gpurun_out/rm_bwd2 -f python tools/rm_bwd_probe.py 2 1 > gpurun_out/ncu_bwd2.log 2>&1
tail -3 gpurun_out/ncu_bwd2.log
