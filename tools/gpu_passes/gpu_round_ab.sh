#!/bin/bash
# Timing experiment: is the conv epilogue (TMEM drain + global stores) the limiter of the low-K layers?  (results are wrong on purpose)
mkdir -p gpurun_out
fmt='import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print(d["layer"], "bf16x3 %.3f ms  %.0f TFLOP/s executed | bf16 %.3f ms" % (d["ms_bf16x3"], 3 * d["gflop"] / d["ms_bf16x3"], d["ms_bf16"]))'
{ echo "# normal"; timeout 300 python tools/conv_microbench.py 16 2>&1 | python -c "$fmt";
  echo "# GP3D_DBG_SKIP_STORE=1 (TMEM drained, nothing stored)"; GP3D_DBG_SKIP_STORE=1 timeout 300 python tools/conv_microbench.py 16 2>&1 | python -c "$fmt";
  echo "# GP3D_DBG_SKIP_STORE=2 (no TMEM reads, no stores)"; GP3D_DBG_SKIP_STORE=2 timeout 300 python tools/conv_microbench.py 16 2>&1 | python -c "$fmt"; } > gpurun_out/ab_skip_store.txt
echo done
