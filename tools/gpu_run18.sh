mkdir -p gpurun_out
(timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_ops.py tests/test_gpu_render.py -m gpu -q -x --tb=line -k "not full_size and not large_shape" 2>&1 | tail -25) > gpurun_out/sanitizer_ops_render.log
(timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_tc.py -m gpu -q -x --tb=line -k "fused or conv5x5 or strided or transposed or gemm" 2>&1 | tail -25) > gpurun_out/sanitizer_tc.log
tail -12 gpurun_out/sanitizer_ops_render.log; tail -12 gpurun_out/sanitizer_tc.log
