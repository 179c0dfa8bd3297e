#!/usr/bin/env python
"""Latency of the small / realistic configurations through the Python API, ours vs the reference's CUDA extension:
  wall_us   per fwd+bwd iteration, back-to-back iterations, one synchronise at the end (what a training loop sees)
  host_us   the same loop timed WITHOUT the final synchronise: CPU time needed to issue one iteration (if ~ wall_us: host bound)
  device_us CUDA-event time of ONE isolated iteration (launch gaps included)
Output: gpurun_out/small_configs_latency.json"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch

import scenes
import gendr_b200 as gd
from ref_gpu import load_reference, reference_render

dev = torch.device('cuda:0')
ref = load_reference()


def measure(fn, n=200):
    for _ in range(20):
        fn()
    wall = host = 1e30
    for _ in range(3):      # best of 3 loops: a fresh box shows +-10 us of host jitter between loops
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        wall, host = min(wall, t2 - t0), min(host, t1 - t0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    devs = []
    for _ in range(10):
        torch.cuda.synchronize()
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        devs.append(e0.elapsed_time(e1) * 1e3)
    devs.sort()
    return {'wall_us': round(wall / n * 1e6, 1), 'host_us': round(host / n * 1e6, 1), 'device_us': round(devs[len(devs) // 2], 1)}


out = {}
cases = {}
fv, ft, cfg = scenes.config_c1()
cases['C1 1 triangle 32x32 B=1 uniform+probabilistic'] = (fv, ft, dict(cfg, double_side=False))
verts, faces = scenes.icosphere(3)
B = 256
fvb, ftb = scenes.render_inputs(verts * 0.5, faces, eyes=scenes.orbit_eyes(B), batch=B)
cases['recon: icosphere 1280 faces 64x64 B=256 uniform tau=10^-1.5 dist_eps=300 hard RGB'] = (fvb, ftb, dict(
    image_size=64, dist_func='uniform', dist_scale=10 ** -1.5, dist_eps=300., aggr_alpha_func='probabilistic', aggr_rgb_func='hard', double_side=False))
cases['opt_shape-like: icosphere 1280 faces 64x64 B=24 logistic tau=1e-2 dist_eps=100'] = (fvb[:24].contiguous(), ftb[:24].contiguous(), dict(
    image_size=64, dist_func='logistic', dist_scale=1e-2, dist_eps=100., aggr_alpha_func='probabilistic', double_side=False))
for name, (fv, ft, kw) in cases.items():
    a0, b0 = fv.to(dev), ft.to(dev)
    g = torch.randn(fv.shape[0], 4, kw['image_size'], kw['image_size'], device=dev)

    def ours():
        a = a0.clone().requires_grad_(True)
        gd.functional.render(a, b0, **kw).backward(g)

    def theirs():
        a = a0.clone().requires_grad_(True)
        reference_render(ref, a, b0, **kw).backward(g)
    r = {'ours': measure(ours)}
    if ref is not None:
        r['reference'] = measure(theirs, n=30)
        r['speedup_wall'] = round(r['reference']['wall_us'] / r['ours']['wall_us'], 2)
    out['functional.render | ' + name] = r
    print(name, r, flush=True)

# module pipeline: Mesh -> Lighting -> LookAt -> GenDR (what experiments/opt_shape.py:257-259 runs per iteration)
for B, S, cfg in ((24, 64, dict(dist_func='logistic', dist_scale=1e-2, dist_eps=100.)),
                  (64, 64, dict(dist_func='uniform', dist_scale=10 ** -1.5, dist_eps=300., aggr_rgb_func='hard'))):
    v = (verts * 0.5)[None].repeat(B, 1, 1).to(dev)
    f = faces[None].repeat(B, 1, 1).to(dev)
    eyes = scenes.orbit_eyes(B).to(dev)
    g = torch.randn(B, 4, S, S, device=dev)
    res = {}
    for label, pkg in (('ours', gd), ('reference', ref)):
        if pkg is None:
            continue
        extra = {} if pkg is gd else dict(dist_shape=0., dist_shift=0., aggr_alpha_t_conorm_p=0.)
        cam = pkg.LookAt(viewing_angle=15)
        cam.set_eyes(eyes)
        light, renderer = pkg.Lighting(), pkg.GenDR(image_size=S, **cfg, **extra)

        def step():
            a = v.clone().requires_grad_(True)
            renderer(cam(light(pkg.Mesh(a, f)))).backward(g)
        res[label] = measure(step, n=200 if pkg is gd else 30)
    if 'reference' in res:
        res['speedup_wall'] = round(res['reference']['wall_us'] / res['ours']['wall_us'], 2)
    out['modules Mesh>Lighting>LookAt>GenDR | icosphere 1280 faces %dx%d B=%d %s' % (S, S, B, cfg['dist_func'])] = res
    print(B, S, cfg, res, flush=True)
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'small_configs_latency.json'), 'w'), indent=1)
