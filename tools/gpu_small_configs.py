"""Latency of small / realistic configurations through the Python API (host overhead included), ours vs the reference CUDA."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch, scenes, numpy as np
import gendr_b200 as gd
from ref_gpu import load_reference, reference_render
dev = torch.device('cuda:0'); ref = load_reference()
def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
out = {}
cases = {}
fv, ft, cfg = scenes.config_c1(); cases['C1 1 triangle 32x32 B=1'] = (fv, ft, dict(cfg, double_side=False))
verts, faces = scenes.icosphere(3)
B = 256
fv, ft = scenes.render_inputs(verts * 0.5, faces, eyes=scenes.orbit_eyes(B), batch=B)
cases['recon: icosphere 1280 faces 64x64 B=256 uniform tau=10^-1.5 dist_eps=300 hard RGB'] = (fv, ft, dict(image_size=64, dist_func='uniform', dist_scale=10 ** -1.5, dist_eps=300., aggr_alpha_func='probabilistic', aggr_rgb_func='hard', double_side=False))
cases['opt_shape-like: icosphere 1280 faces 64x64 B=24 logistic tau=1e-2 dist_eps=100'] = (fv[:24].contiguous(), ft[:24].contiguous(), dict(image_size=64, dist_func='logistic', dist_scale=1e-2, dist_eps=100., aggr_alpha_func='probabilistic', double_side=False))
for name, (fv, ft, kw) in cases.items():
    a0, b0 = fv.to(dev), ft.to(dev)
    g = torch.randn(fv.shape[0], 4, kw['image_size'], kw['image_size'], device=dev)
    def ours():
        a = a0.clone().requires_grad_(True); gd.functional.render(a, b0, **kw).backward(g)
    def theirs():
        a = a0.clone().requires_grad_(True); reference_render(ref, a, b0, **kw).backward(g)
    r = dict(ours_ms=timeit(ours))
    if ref is not None: r['reference_ms'] = timeit(theirs, n=10); r['speedup'] = r['reference_ms'] / r['ours_ms']
    out[name] = r; print(name, r)
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'small_configs.json'), 'w'), indent=1)
