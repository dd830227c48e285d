cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_render.py -m gpu -x -q --tb=short 2>&1 | tail -12
timeout 300 python tools/rm_bwd_probe.py 0,1,2 5 2>&1 | tail -3 | tee gpurun_out/rm_bwd_probe_v2.txt
