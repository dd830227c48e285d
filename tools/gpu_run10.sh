mkdir -p gpurun_out
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --batch-gpu 16 2>&1 | tail -4) > gpurun_out/bench10_n2.log
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 3 2>&1 | tail -2) > gpurun_out/bench10_n2_ref.log
cut -c1-500 gpurun_out/bench10_n2.log; cut -c1-300 gpurun_out/bench10_n2_ref.log
