"""Where does a CUDA-graph replay of the generator diverge from the eager call?  Captures stage by stage and prints mismatch statistics."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import cases
lg = importlib.import_module('3dgp_b200.legacy'); dn = importlib.import_module('3dgp_b200.dnnlib'); inf = importlib.import_module('3dgp_b200.training.inference')
Ge = lg.load_network_pkl(os.path.join(ROOT, 'tests', 'golden', 'snapshot_small.pkl.gz'), device='cuda', names=('G_ema',))['G_ema']
kw = cases.net_kwargs('small')
t = {k: torch.from_numpy(v).cuda() for k, v in cases.net_inputs(kw).items()}
cam = dn.TensorGroup(angles=t['angles'], fov=t['fov'], radius=t['radius'], look_at=t['look_at'])
print({k: tuple(v.shape) for k, v in t.items() if k in ('z', 'c', 'angles', 'fov', 'radius', 'look_at')})
R = Ge.synthesis.renderer


def stat(name, a, b):
    a, b = a.float(), b.float()
    print(f'{name:40s} mismatching {int((a != b).sum())} / {a.numel()}  max abs diff {float((a - b).abs().max()):.3e}  ref max {float(b.abs().max()):.3e}', flush=True)


def graphed(fn, n_warm=2):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(n_warm):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fn()
    return g, out


with torch.no_grad():
    # determinism of the eager path
    R.launch_counter = 10; a = Ge(z=t['z'], c=t['c'], camera_params=cam, camera_angles_cond=cam.angles, noise_mode='const')
    R.launch_counter = 10; b = Ge(z=t['z'], c=t['c'], camera_params=cam, camera_angles_cond=cam.angles, noise_mode='const')
    a = a if torch.is_tensor(a) else a.img; b = b if torch.is_tensor(b) else b.img
    stat('eager vs eager (float img)', a, b)
    # mapping only
    g, ws_g = graphed(lambda: Ge.mapping(t['z'], t['c'], camera_angles=cam.angles))
    g.replay(); stat('mapping: graph vs eager', ws_g, Ge.mapping(t['z'], t['c'], camera_angles=cam.angles))
    ws = Ge.mapping(t['z'], t['c'], camera_angles=cam.angles)
    dec = Ge.synthesis.tri_plane_decoder
    g, pl_g = graphed(lambda: dec(ws[:, :dec.num_ws], noise_mode='const'))
    g.replay(); stat('tri-plane decoder: graph vs eager', pl_g, dec(ws[:, :dec.num_ws], noise_mode='const'))
    def synth():
        return Ge.synthesis(ws, camera_params=cam, noise_mode='const')
    g, o_g = graphed(synth)
    off = R.launch_counter
    g.replay()
    R.launch_counter = off - 1
    o_e = synth()
    o_g = o_g if torch.is_tensor(o_g) else o_g.img; o_e = o_e if torch.is_tensor(o_e) else o_e.img
    stat('synthesis (decoder+render+adaptor)', o_g, o_e)
    for ch in range(o_e.shape[1]):
        stat(f'  channel {ch}', o_g[:, ch], o_e[:, ch])
