cd /root/repo
mkdir -p gpurun_out
echo "== ours N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 6 --warmup 3 2>&1 | tail -1 | tee gpurun_out/final_bench_n2.json | cut -c1-420
