#!/bin/bash
# Round-2 GPU pass O: TMA-staged NCHW FIR with 16-byte-aligned window starts (parity + ops bench).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q > gpurun_out/o_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/o_pytest.log
timeout 600 python bench.py --workload ops --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/o_bench_ops.json 2> gpurun_out/o_bench_ops.err
echo done
