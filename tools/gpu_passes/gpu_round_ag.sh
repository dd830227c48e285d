#!/bin/bash
# Round-2 GPU pass AG: GPU suite (incl. the training-set stream and training_iterations tests) + smoke on the tree with the dataset / data-path modules.
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -q > gpurun_out/ag_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/ag_pytest_all.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ag_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/ag_smoke.log
echo done
