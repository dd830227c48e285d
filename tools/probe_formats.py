"""GPU probe: which kind::f16 operand-format combinations does tcgen05.mma accept on this part?  One combination per PROCESS (a device-side fault is sticky)."""
import importlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODES = {'bf16x3': (3, False, False), 'x2w16': (2, False, False), 'f16xf16': (16, False, False), 'bf16xf16': (16, True, False), 'f16xbf16': (16, False, True), 'bf16xbf16': (1, False, False)}

if len(sys.argv) > 1:
    import torch
    sys.path.insert(0, ROOT)
    tc = importlib.import_module('3dgp_b200.torch_utils.ops.tc')
    importlib.import_module('3dgp_b200.config').set_reference_numerics()
    torch.manual_seed(0)
    terms, xg, wg = MODES[sys.argv[1]]
    x = torch.randn(4, 128, 32, 32, device='cuda'); w = torch.randn(128, 128, 3, 3, device='cuda') / (128 * 9) ** 0.5
    yd = torch.nn.functional.conv2d(x.double(), w.double(), padding=1)
    y = tc.conv2d_forward(x, w, terms, x_is_grad=xg, w_is_grad=wg)
    torch.cuda.synchronize()
    print(f'{sys.argv[1]:10s} OK  l2-rel {((y.double() - yd).norm() / yd.norm()).item():.3e}')
else:
    for m in MODES:
        r = subprocess.run([sys.executable, __file__, m], capture_output=True, text=True, timeout=300)
        out = (r.stdout.strip().splitlines() or [''])[-1]
        err = [l for l in r.stderr.splitlines() if 'error' in l.lower()][:1]
        print(out if r.returncode == 0 else f'{m:10s} FAILED rc={r.returncode} {err}')
