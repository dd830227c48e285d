#!/bin/bash
# Round-2 GPU pass H (2 GPUs): register-blocked filtered_lrelu (parity + ops bench), N=2 train step with / without the overlapped bucketed all-reduce.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "filtered" > gpurun_out/h_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/h_pytest.log
timeout 600 python bench.py --workload ops --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/h_bench_ops.json 2> gpurun_out/h_bench_ops.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/h_bench_n2_overlap.json 2> gpurun_out/h_bench_n2_overlap.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-overlap > gpurun_out/h_bench_n2_blocking.json 2> gpurun_out/h_bench_n2_blocking.err
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/h_bench_n1.json 2> gpurun_out/h_bench_n1.err
echo done
