#!/bin/bash
# Timing experiment: upper bound of activation-window reuse (taps with dy != 0 skip their activation load; results are wrong on purpose).
mkdir -p gpurun_out
echo "# normal" > gpurun_out/aa_skip_a.txt
timeout 300 python tools/conv_microbench.py 16 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print(d['layer'], 'bf16x3 %.3f ms  %.0f TFLOP/s executed' % (d['ms_bf16x3'], 3 * d['gflop'] / d['ms_bf16x3']))" >> gpurun_out/aa_skip_a.txt
echo "# GP3D_DBG_SKIP_A=1 (activation loads of 6 of the 9 taps skipped: 1/3 of the activation traffic)" >> gpurun_out/aa_skip_a.txt
GP3D_DBG_SKIP_A=1 timeout 300 python tools/conv_microbench.py 16 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print(d['layer'], 'bf16x3 %.3f ms  %.0f TFLOP/s executed' % (d['ms_bf16x3'], 3 * d['gflop'] / d['ms_bf16x3']))" >> gpurun_out/aa_skip_a.txt
echo done
