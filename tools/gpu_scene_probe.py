#!/usr/bin/env python
"""Dump every intermediate of the reference's camera transform and lighting (torch ops on the GPU) for random inputs, so that
the exact fp32 operation order torch/cuBLAS use can be identified offline (tools/scene_probe_analyze.py) and mirrored by
camera_forward_kernel / lighting_forward_kernel.  Output: gpurun_out/scene_probe.npz"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import torch.nn.functional as F

from ref_gpu import load_reference

dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)
B, V, Fn = 16, 642, 1280
verts = ((torch.rand(B, V, 3, generator=g) - 0.5) * 1.2).to(dev)
eyes = (torch.randn(B, 3, generator=g) * 2 + torch.tensor([0., 0.5, -2.5])).to(dev)
faces = torch.randint(0, V, (B, Fn, 3), generator=g).to(dev)
at = torch.zeros(B, 3, device=dev); up = torch.tensor([0., 1., 0.], device=dev)[None].repeat(B, 1)
out = {'verts': verts, 'eyes': eyes, 'faces': faces}
d = at - eyes
out['d'] = d
out['norm_d'] = torch.norm(d, 2, 1)
z = F.normalize(d, eps=1e-5); out['z'] = z
cx = torch.cross(up, z, dim=1); out['cx'] = cx
x = F.normalize(cx, eps=1e-5); out['x'] = x
cy = torch.cross(z, x, dim=1); out['cy'] = cy
y = F.normalize(cy, eps=1e-5); out['y'] = y
r = torch.cat((x[:, None, :], y[:, None, :], z[:, None, :]), dim=1)
vm = verts - eyes[:, None, :]; out['vm'] = vm
vc = torch.matmul(vm, r.transpose(1, 2)); out['vc'] = vc
angle = torch.tensor(15. / 180 * np.pi, dtype=torch.float32, device=dev)[None]
width = torch.tan(angle)[:, None]; out['width'] = width
zc = vc[:, :, 2]
out['xs'] = vc[:, :, 0] / zc / width; out['ys'] = vc[:, :, 1] / zc / width
# the reference package itself (must equal the above)
ref = load_reference()
if ref is not None:
    m = ref.Mesh(verts, faces.int())
    cam = ref.LookAt(viewing_angle=15); cam.set_eyes(eyes)
    out['ref_screen'] = cam(m).vertices
    lit = ref.Lighting()(m)
    out['ref_lit'] = lit.textures
# lighting pieces (surface)
fv = verts.reshape(B * V, 3)[(faces + (torch.arange(B, device=dev) * V)[:, None, None]).long()]
out['fv'] = fv
a = fv[:, :, 2] - fv[:, :, 1]; e = fv[:, :, 0] - fv[:, :, 1]
cr = torch.cross(a, e, dim=2); out['cr'] = cr
out['norm_cr'] = torch.norm(cr, 2, 2)
n = F.normalize(cr, p=2, dim=2, eps=1e-6); out['n'] = n
ld = torch.tensor([0., 1., 0.], device=dev)[None]
ld2 = torch.tensor([0.3, 0.8, -0.5], device=dev)[None]
out['cos'] = F.relu(torch.sum(n * ld, dim=2))
out['cos2_raw'] = torch.sum(n * ld2, dim=2)
light = torch.zeros(B, Fn, 3, device=dev)
light += 0.5 * torch.ones(1, 3, device=dev)[:, None, :]
light += 0.5 * (torch.ones(1, 3, device=dev)[:, None, :] * out['cos'][:, :, None])
out['light'] = light
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
np.savez(os.path.join(ROOT, 'gpurun_out', 'scene_probe.npz'), **{k: v.detach().cpu().numpy() for k, v in out.items()})
print('saved', {k: tuple(v.shape) for k, v in out.items()})
