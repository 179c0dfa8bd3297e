"""Generates tests/golden/scene_v1.npz: golden vectors for the camera-transform and lighting kernels
from the UNMODIFIED reference Python (gendr.Lighting, gendr.LookAt, gendr.functional.look, transform.orthogonal; pure torch,
run here on the CPU in fp32) and its autograd gradients.  Needs the staged reference package (baseline/_ref, built by
__graft_entry__.build() where /root/reference exists).    Run:  python tests/golden/make_golden_scene.py

Cases (batch 4, icosphere with 42 vertices / 80 faces, jittered):
  a  Lighting() -> LookAt(viewing_angle=15, eyes from angles), perspective, texture_res 1
  b  Lighting(custom colours / direction) -> functional.look(explicit up) + orthogonal(scale 0.8), texture_res 2
     (the reference's Look module crashes on its own default up=None, functional/look.py:36, so the function is called directly)
For each: screen-space vertices, lit textures, and for random cotangents the gradients w.r.t. the world-space vertices
(camera path and lighting path separately) and w.r.t. the unlit textures.  Batch / face counts avoid 3: the reference calls
torch.cross without `dim` (mesh.py:108), which picks the first size-3 dimension."""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import scenes  # noqa: E402
from ref_gpu import load_reference  # noqa: E402

warnings.filterwarnings('ignore')


def main():
    ref = load_reference()
    assert ref is not None, 'baseline/_ref not staged'
    import gendr.transform as ref_transform
    g = torch.Generator().manual_seed(11)
    verts, faces = scenes.icosphere(1)
    B, V, F = 4, verts.shape[0], faces.shape[0]
    world = (verts * 0.5)[None].repeat(B, 1, 1) * (1 + 0.2 * (2 * torch.rand(B, V, 1, generator=g) - 1))
    world = world + 0.05 * torch.randn(B, V, 3, generator=g)
    index = faces[None].repeat(B, 1, 1)
    out = {'vertices': world.numpy(), 'faces': index.numpy()}
    cases = {
        'a': dict(T=1, light=dict(), cam='look_at'),
        'b': dict(T=4, light=dict(intensity_ambient=0.3, color_ambient=[1, 0.9, 0.8], intensity_directionals=0.7,
                                  color_directionals=[0.5, 1, 0.7], directions=[0.3, 0.8, -0.5]), cam='look'),
    }
    for name, c in cases.items():
        tex = torch.rand(B, F, c['T'], 3, generator=g)
        g_screen = torch.randn(B, V, 3, generator=g)
        g_lit = torch.randn(B, F, c['T'], 3, generator=g)
        if c['cam'] == 'look_at':
            eyes = ref.functional.get_points_from_angles(torch.full((B,), 2.732), torch.tensor([30., 10., -20., 45.]),
                                                         torch.tensor([0., 90., 200., -45.]))
        else:
            eyes = torch.tensor([0.1, 0.2, -2.5])

        def pipeline(v, t):
            mesh = ref.Mesh(v, index, t, texture_res=int(c['T'] ** 0.5))
            lit = ref.Lighting(**c['light'])(mesh).textures
            if c['cam'] == 'look_at':
                cam = ref.LookAt(viewing_angle=15)
                cam.set_eyes(eyes)
                screen = cam.transform(v)
            else:
                screen = ref_transform.orthogonal(ref.functional.look(v, eyes, [0.1, -0.2, 1.0], up=torch.tensor([0., 1., 0.])), scale=0.8)
            return screen, lit

        v = world.clone().requires_grad_(True)
        t = tex.clone().requires_grad_(True)
        screen, lit = pipeline(v, t)
        gv_cam, = torch.autograd.grad((screen * g_screen).sum(), v, retain_graph=True)
        gv_light, gt = torch.autograd.grad((lit * g_lit).sum(), (v, t))
        for k, val in dict(textures=tex, eyes=eyes, g_screen=g_screen, g_lit=g_lit, screen=screen, lit=lit, gv_cam=gv_cam,
                           gv_light=gv_light, gt=gt).items():
            out['%s_%s' % (name, k)] = val.detach().numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, 'scene_v1.npz'), **out)
    print('wrote scene_v1.npz:', {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
