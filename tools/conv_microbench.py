"""Tensor-core conv microbench: TFLOP/s of the tcgen05 kernels at the BASELINE layer shapes (CUDA events, 3 warm-up + 10 timed).
python tools/conv_microbench.py [batch]"""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
tc = importlib.import_module('3dgp_b200.torch_utils.ops.tc')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else dict(bf16_tflops=1590.0)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


rows = []
for (name, Cin, Cout, H, k) in [('b512.conv1', 128, 128, 512, 3), ('b256.conv1', 256, 256, 256, 3), ('b128.conv1', 512, 512, 128, 3),
                                ('b64.conv1', 1024, 1024, 64, 3), ('D.b128.conv0@64', 1024, 1024, 64, 3), ('b512.torgb', 128, 96, 512, 1)]:
    x = torch.randn(B, H, H, Cin, device='cuda')
    w = torch.randn(Cout, k, k, Cin, device='cuda') / (k * Cin ** 0.5)
    xh, xl = tc.split_bf16(x); wh, wl = tc.split_bf16(w)
    y = torch.empty(B, H, H, Cout, device='cuda')
    L = tc._lib.lib(); st = tc._lib.stream_ptr()
    flops = 2.0 * B * H * H * Cin * Cout * k * k
    f3 = lambda: L.gp3d_conv2d_nhwc_bf16x3(xh.data_ptr(), xl.data_ptr(), wh.data_ptr(), wl.data_ptr(), y.data_ptr(), B, H, H, Cin, Cout, k, 0, st)
    L.gp3d_conv_set_wide3(0); t3n = timeit(f3)
    L.gp3d_conv_set_wide3(1); t3 = timeit(f3)
    t1 = timeit(lambda: L.gp3d_conv2d_nhwc_bf16(xh.data_ptr(), wh.data_ptr(), y.data_ptr(), B, H, H, Cin, Cout, k, 0, st))
    rows.append(dict(layer=name, B=B, Cin=Cin, Cout=Cout, H=H, k=k, gflop=flops / 1e9, ms_bf16x3=t3, ms_bf16x3_bn128=t3n, ms_bf16=t1,
                     tensor_tflops_bf16x3=3 * flops / t3 / 1e9, tensor_tflops_bf16=flops / t1 / 1e9,
                     frac_bf16x3=3 * flops / t3 / 1e9 / peaks['bf16_tflops'], frac_bf16=flops / t1 / 1e9 / peaks['bf16_tflops']))
    if Cin % 128 == 0 and Cout % 128 == 0:
        dy = torch.randn(B, Cout, H, H, device='cuda').contiguous(memory_format=torch.channels_last)
        xx = x.permute(0, 3, 1, 2)
        tw = timeit(lambda: tc.conv_wgrad(dy, xx, k, 'conv', 1, k // 2, 3), n=5)
        rows[-1]['ms_wgrad_bf16x3_incl_split'] = tw
    del x, w, xh, xl, y
    torch.cuda.empty_cache()
for r in rows:
    print(json.dumps(r))
