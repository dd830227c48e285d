cd /root/repo
mkdir -p gpurun_out
echo "== reference arm N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --impl reference --gpus 2 --steps 2 --warmup 3 2>&1 | tail -2 | cut -c1-400
echo "== ours N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 6 --warmup 3 2>&1 | tail -1 | tee gpurun_out/final_bench_n2.json | cut -c1-500
echo "== ginfer N=1 B=16"; CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --workload ginfer --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/final_bench_ginfer.json | cut -c1-600
