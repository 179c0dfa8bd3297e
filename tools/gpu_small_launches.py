"""One fwd+bwd of the opt_shape-like configuration through functional.render (run under ncu --metrics gpu__time_duration.sum)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch, scenes
import gendr_b200 as gd
dev = torch.device('cuda:0')
verts, faces = scenes.icosphere(3)
fvb, ftb = scenes.render_inputs(verts * 0.5, faces, eyes=scenes.orbit_eyes(24), batch=24)
kw = dict(image_size=64, dist_func='logistic', dist_scale=1e-2, dist_eps=100., aggr_alpha_func='probabilistic', double_side=False)
a0, b0 = fvb.to(dev), ftb.to(dev)
g = torch.randn(24, 4, 64, 64, device=dev)
for it in range(4):
    a = a0.clone().requires_grad_(True)
    gd.functional.render(a, b0, **kw).backward(g)
torch.cuda.synchronize()
fv, ft, cfg = scenes.config_c1()
a0, b0 = fv.to(dev), ft.to(dev)
g = torch.randn(1, 4, 32, 32, device=dev)
for it in range(4):
    a = a0.clone().requires_grad_(True)
    gd.functional.render(a, b0, double_side=False, **cfg).backward(g)
torch.cuda.synchronize()
