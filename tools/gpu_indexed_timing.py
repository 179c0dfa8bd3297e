"""Timing of SURVEY 8(f) row 1 on C3: fused indexed path vs torch gather + index backward around render()."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch, scenes, numpy as np
import gendr_b200 as gd
dev = torch.device('cuda:0')
B = 64
verts, faces = scenes.grid_sphere(64, seed=0)
mesh = gd.Mesh(verts[None].repeat(B, 1, 1).to(dev), faces[None].repeat(B, 1, 1).to(dev))
cam = gd.LookAt(viewing_angle=15); cam.set_eyes(scenes.orbit_eyes(B).to(dev))
mesh = cam(gd.Lighting()(mesh))
kw = dict(image_size=256, dist_func='gaussian', aggr_alpha_func='einstein', double_side=False)
g = torch.randn(B, 4, 256, 256, device=dev)
V, T, IDX = mesh.vertices.detach(), mesh.textures.detach(), mesh.faces
def unfused():
    v = V.clone().requires_grad_(True)
    gd.functional.render(gd.functional.face_vertices(v, IDX), T, **kw).backward(g)
    return v.grad
def fused():
    v = V.clone().requires_grad_(True)
    gd.functional.render_indexed(v, IDX, T, **kw).backward(g)
    return v.grad
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
ga, gb = unfused(), fused()
out = dict(config='C3 B=64: vertices [64,4225,3] + faces [64,8192,3] -> images, backward to vertices', unfused_ms=timeit(unfused), fused_ms=timeit(fused),
           max_rel_grad_diff=float((ga - gb).abs().max() / ga.abs().max()))
print(json.dumps(out)); json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'indexed_timing.json'), 'w'))
