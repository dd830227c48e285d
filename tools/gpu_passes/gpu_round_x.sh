#!/bin/bash
# Round-2 GPU pass X (2 GPUs): N=2 sanity of the shipped schedule (blocking flat all-reduce), driver-style launch.
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/x_bench_n2.json 2> gpurun_out/x_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/x_bench_ref_n2.json 2> gpurun_out/x_bench_ref_n2.err
echo done
