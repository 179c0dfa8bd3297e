"""Where the ~80 us of a C1 iteration (1 triangle, 32x32, fwd+bwd through functional.render) go on the HOST.
Each segment is timed as a back-to-back loop with one synchronise at the end (host-bound regime => wall == host time).
Output: gpurun_out/c1_breakdown.json"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch

import scenes
import gendr_b200 as gd

dev = torch.device('cuda:0')
fv, ft, cfg = scenes.config_c1()
kw = dict(cfg, double_side=False)
a0, b0 = fv.to(dev), ft.to(dev)
g = torch.randn(1, 4, 32, 32, device=dev)


def loop(fn, n=2000):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t0) / n * 1e6)
    return round(best, 2)


class _Id(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x * 1.0

    @staticmethod
    def backward(ctx, gout):
        return gout


a_req = a0.clone().requires_grad_(True)
ones = torch.ones_like(a0)


def seg_clone():
    a0.clone().requires_grad_(True)


def seg_fwd_nograd():
    gd.functional.render(a0, b0, **kw)


def seg_fwd_grad():
    gd.functional.render(a_req, b0, **kw)


def seg_full():
    a = a0.clone().requires_grad_(True)
    gd.functional.render(a, b0, **kw).backward(g)


def seg_full_noclone():
    a_req.grad = None
    gd.functional.render(a_req, b0, **kw).backward(g)


def seg_torch_floor():      # the cheapest possible differentiable op + backward through the engine: torch's own floor
    a_req.grad = None
    (a_req * 1.0).backward(ones)


def seg_pyfunc_floor():
    a_req.grad = None
    _Id.apply(a_req).backward(ones)


out = {name: loop(fn) for name, fn in [
    ('clone+requires_grad', seg_clone), ('render fwd (no grad)', seg_fwd_nograd), ('render fwd (grad graph)', seg_fwd_grad),
    ('fwd+bwd (no clone)', seg_full_noclone), ('fwd+bwd (with clone) = C1 line', seg_full),
    ('torch floor: (a*1).backward', seg_torch_floor), ('python Function floor', seg_pyfunc_floor)]}
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'c1_breakdown.json'), 'w'), indent=1)
