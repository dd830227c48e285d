mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -15) > gpurun_out/tests13.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) >> gpurun_out/tests13.log
cat gpurun_out/tests13.log
