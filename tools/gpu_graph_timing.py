"""CUDA-graph capture of the whole scene step (Lighting -> LookAt -> GenDR, forward + backward) with
torch.cuda.make_graphed_callables: the kernels are launched through the C ABI on torch's current stream, so they are captured
like any torch op.  Latency of the eager fused path vs the graphed one on launch-bound configurations; results must be equal.
Writes gpurun_out/graph_timing.json."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
import scenes  # noqa: E402
import gendr_b200 as gd  # noqa: E402

dev = torch.device('cuda:0')


class SceneStep(nn.Module):
    def __init__(self, faces, eyes, S, cfg):
        super().__init__()
        self.faces, self.lighting, self.renderer = faces, gd.Lighting(), gd.GenDR(image_size=S, **cfg)
        self.camera = gd.LookAt(viewing_angle=15)
        self.camera.set_eyes(eyes)

    def forward(self, vertices):
        return self.renderer(self.camera(self.lighting(gd.Mesh(vertices, self.faces))))


def timeit(fn, n=100):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


out = {}
verts, faces = scenes.icosphere(3)
for name, B, S, cfg in (('opt_shape-like: 1280 faces, 64x64, B=24, logistic+probabilistic', 24, 64, dict(dist_func='logistic', dist_scale=1e-2, dist_eps=100.)),
                        ('recon-like: 1280 faces, 64x64, B=64, uniform, hard RGB', 64, 64, dict(dist_func='uniform', dist_scale=10 ** -1.5, dist_eps=300., aggr_rgb_func='hard')),
                        ('C2-like: 1280 faces, 256x256, B=16, logistic+probabilistic', 16, 256, dict(dist_func='logistic'))):
    v = (verts * 0.5)[None].repeat(B, 1, 1).to(dev)
    f = faces[None].repeat(B, 1, 1).to(dev)
    eyes = scenes.orbit_eyes(B).to(dev)
    g = torch.randn(B, 4, S, S, device=dev)
    step = SceneStep(f, eyes, S, cfg)

    def eager():
        a = v.clone().requires_grad_(True)
        step(a).backward(g)
        return a.grad
    want = eager().clone()
    graphed = torch.cuda.make_graphed_callables(step, (v.clone().requires_grad_(True),))
    static_in = v.clone().requires_grad_(True)

    def replay():
        static_in.grad = None
        graphed(static_in).backward(g)
        return static_in.grad
    got = replay().clone()
    rel = float((got - want).abs().max() / want.abs().max())
    r = {'eager_fused_ms': timeit(eager), 'cuda_graph_ms': timeit(replay), 'grad_rel_diff_vs_eager': rel}
    r['speedup'] = r['eager_fused_ms'] / r['cuda_graph_ms']
    out[name] = r
    print(name, r, flush=True)
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'graph_timing.json'), 'w'), indent=1)
