import importlib, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
cg = importlib.import_module('3dgp_b200.torch_utils.ops.conv2d_gradfix')
torch.backends.cudnn.allow_tf32 = False
g = torch.Generator(device='cuda').manual_seed(11)
for (cin, cout, k, H) in [(64, 64, 5, 64), (64, 64, 3, 64), (128, 128, 5, 32), (64, 128, 5, 16)]:
    x = torch.randn(4, cin, H, H, device='cuda', generator=g)
    w = torch.randn(cout, cin, k, k, device='cuda', generator=g) / (k * cin ** 0.5)
    ref = torch.nn.functional.conv2d(x.double().cpu(), w.double().cpu(), padding=k // 2)
    cg.tc_enabled = True
    y = cg.conv2d(x, w, padding=k // 2).double().cpu()
    cg.tc_enabled = False
    y0 = cg.conv2d(x, w, padding=k // 2).double().cpu()
    cg.tc_enabled = True
    rel = lambda a, b: ((a - b).abs().max() / b.abs().max()).item()
    print((cin, cout, k, H), 'tc vs f64', rel(y, ref), ' aten vs f64', rel(y0, ref))
