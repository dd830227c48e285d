"""Stride-2 transposed 3x3 conv of the up-sampling layers: four polyphase launches vs the four phases in ONE launch (bf16x3).  python tools/upconv_microbench.py [batch]"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
tc = importlib.import_module('3dgp_b200.torch_utils.ops.tc')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16


def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n


for name, Cin, Cout, H in [('b512.conv0', 256, 128, 256), ('b256.conv0', 512, 256, 128), ('b128.conv0', 1024, 512, 64), ('b64.conv0', 1024, 1024, 32)]:
    x = torch.randn(B, H, H, Cin, device='cuda'); w = torch.randn(Cout, 9, Cin, device='cuda') / (3 * Cin ** 0.5)
    xh, xl = tc.split_bf16(x); wh, wl = tc.split_bf16(w)
    y1 = torch.empty(B, 2 * H + 1, 2 * H + 1, Cout, device='cuda'); y4 = torch.empty_like(y1)

    def four():
        for a in (0, 1):
            kys = [(0, 0), (-1, 2)] if a == 0 else [(0, 1)]
            for b in (0, 1):
                kxs = [(0, 0), (-1, 2)] if b == 0 else [(0, 1)]
                taps = [(dy, dx, ky * 3 + kx) for (dy, ky) in kys for (dx, kx) in kxs]
                tc._taps_launch(xh, xl, wh, wl, y4, B, H, H, Cin, Cout, 9, taps, 1, H + 1 - a, H + 1 - b, 2 * H + 1, 2 * H + 1, 2, 2, a, b)
    t4 = timeit(four)
    t1 = timeit(lambda: tc.conv_transpose_s2_launch(xh, xl, wh, wl, y1, B, H, H, Cin, Cout))
    four(); tc.conv_transpose_s2_launch(xh, xl, wh, wl, y1, B, H, H, Cin, Cout); torch.cuda.synchronize()
    gf = 2.0 * B * Cin * Cout * 9 * H * H / 1e9
    print(f'{name:12s} {Cin:4d}->{Cout:4d} @{H}^2 B={B}: four launches {t4:.3f} ms, one launch {t1:.3f} ms ({(1 - t1 / t4) * 100:+.0f} % time), '
          f'{3 * gf / t1:.0f} executed TFLOP/s, bit-identical: {bool(torch.equal(y1, y4))}', flush=True)
    del x, w, xh, xl, y1, y4
    torch.cuda.empty_cache()
