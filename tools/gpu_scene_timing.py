"""Latency of the whole scene step (Mesh -> Lighting -> LookAt -> GenDR, forward + backward) through the module API:
ours with the deferred/fused scene path, ours with torch glue (FUSE_SCENE off), and the reference package, on the
reference's typical optimisation configurations.  Writes gpurun_out/scene_timing.json."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch  # noqa: E402
import scenes  # noqa: E402
import gendr_b200 as gd  # noqa: E402
from gendr_b200 import mesh as mesh_mod  # noqa: E402
from ref_gpu import load_reference  # noqa: E402

dev = torch.device('cuda:0')
ref = load_reference()


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def case(name, verts, faces, B, S, aa, cfg, out):
    v = (verts * 0.5)[None].repeat(B, 1, 1).to(dev)
    f = faces[None].repeat(B, 1, 1).to(dev)
    eyes = scenes.orbit_eyes(B).to(dev)
    g = torch.randn(B, 4, S, S, device=dev)

    def step(pkg):
        a = v.clone().requires_grad_(True)
        cam = pkg.LookAt(viewing_angle=15)
        cam.set_eyes(eyes)
        img = pkg.GenDR(image_size=S, anti_aliasing=aa, **cfg)(cam(pkg.Lighting()(pkg.Mesh(a, f))))
        img.backward(g)
        return a.grad
    r = {}
    mesh_mod.FUSE_SCENE = True
    r['ours_fused_ms'] = timeit(lambda: step(gd))
    mesh_mod.FUSE_SCENE = False
    r['ours_torch_glue_ms'] = timeit(lambda: step(gd))
    mesh_mod.FUSE_SCENE = True
    if ref is not None:
        r['reference_ms'] = timeit(lambda: step(ref), n=10)
        r['speedup_vs_reference'] = r['reference_ms'] / r['ours_fused_ms']
    r['fused_vs_glue'] = r['ours_torch_glue_ms'] / r['ours_fused_ms']
    out[name] = r
    print(name, r, flush=True)


out = {}
ico_v, ico_f = scenes.icosphere(3)
base = dict(dist_shape=0.0, dist_shift=0.0, aggr_alpha_t_conorm_p=0.0)
case('opt_shape-like: icosphere 1280 faces, 64x64, B=24, logistic+probabilistic, dist_eps=100', ico_v, ico_f, 24, 64, False,
     dict(base, dist_func='logistic', dist_scale=1e-2, dist_eps=100., aggr_alpha_func='probabilistic'), out)
case('recon-like: icosphere 1280 faces, 64x64, B=64, uniform tau=10^-1.5, hard RGB', ico_v, ico_f, 64, 64, False,
     dict(base, dist_func='uniform', dist_scale=10 ** -1.5, dist_eps=300., aggr_alpha_func='probabilistic', aggr_rgb_func='hard'), out)
case('anti-aliased: icosphere 1280 faces, 128x128 (rendered 256x256), B=16, gaussian+einstein', ico_v, ico_f, 16, 128, True,
     dict(base, dist_func='gaussian', aggr_alpha_func='einstein'), out)
gs_v, gs_f = scenes.grid_sphere(64)
case('C3 mesh through the modules: 8192 faces, 256x256, B=64, gaussian+einstein', gs_v * 2, gs_f, 64, 256, False,
     dict(base, dist_func='gaussian', aggr_alpha_func='einstein'), out)
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'scene_timing.json'), 'w'), indent=1)
