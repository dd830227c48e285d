#!/bin/bash
# Round-2 GPU pass R (4 GPUs): N=4 train step with the overlapped bucketed all-reduce vs the blocking one.
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-ginfer > gpurun_out/r_bench_n4_overlap.json 2> gpurun_out/r_bench_n4_overlap.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-ginfer --no-overlap > gpurun_out/r_bench_n4_blocking.json 2> gpurun_out/r_bench_n4_blocking.err
echo done
