"""One launch of the 128 -> 128 3x3 bf16x3 convolution at 512^2 (B = 8) for `ncu --set full`: why does this layer run at 1.2 PFLOP/s executed while the
512-channel layers reach 1.6?"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
tc = importlib.import_module('3dgp_b200.torch_utils.ops.tc')
B, H, C = 8, 512, 128
x = torch.randn(B, H, H, C, device='cuda'); w = torch.randn(C, 9, C, device='cuda') / 34
xh, xl = tc.split_bf16(x); wh, wl = tc.split_bf16(w)
y = torch.empty(B, H, H, C, device='cuda')
L = tc._lib.lib()
for _ in range(2):
    L.gp3d_conv2d_nhwc_bf16x3(xh.data_ptr(), xl.data_ptr(), wh.data_ptr(), wl.data_ptr(), y.data_ptr(), B, H, H, C, C, 3, 0, tc._lib.stream_ptr())
torch.cuda.synchronize()
print('done')
