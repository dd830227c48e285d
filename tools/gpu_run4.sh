mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q --tb=short -x 2>&1 | tail -40) > gpurun_out/tests4.log
cat gpurun_out/tests4.log
