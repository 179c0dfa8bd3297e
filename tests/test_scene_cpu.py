"""CPU tests of the scene-step oracle (oracle/scene_oracle.py) against the golden vectors generated from the reference
Python (tests/golden/scene_v1.npz, tests/golden/make_golden_scene.py), and of the torch mirror in gendr_b200/functional."""
import os

import numpy as np
import pytest
import torch

from oracle import scene_oracle as so

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'scene_v1.npz'))
CASES = {
    'a': (dict(mode='look_at', perspective=True, viewing_angle=15.), dict()),
    'b': (dict(mode='look', perspective=False, viewing_scale=0.8, direction=(0.1, -0.2, 1.0)),
          dict(intensity_ambient=0.3, color_ambient=(1, 0.9, 0.8), intensity_directional=0.7, color_directional=(0.5, 1, 0.7),
               direction=(0.3, 0.8, -0.5))),
}


def _close(a, b, rtol, what):
    scale = np.abs(b).max()
    err = np.abs(a - b).max() / scale
    assert err <= rtol, '%s: max|d|/max|ref| = %.3e > %.1e' % (what, err, rtol)


@pytest.mark.parametrize('name', ['a', 'b'])
@pytest.mark.parametrize('dtype,rtol', [(np.float64, 2e-6), (np.float32, 5e-6)])
def test_scene_oracle_matches_reference_python(name, dtype, rtol):
    cam, light = CASES[name]
    v, f = GOLD['vertices'], GOLD['faces']
    g = {k: GOLD['%s_%s' % (name, k)] for k in ('textures', 'eyes', 'g_screen', 'g_lit', 'screen', 'lit', 'gv_cam', 'gv_light', 'gt')}
    _close(so.camera_forward(v, g['eyes'], dtype=dtype, **cam), g['screen'], rtol, 'screen vertices')
    _close(so.lighting_forward(v, f, g['textures'], dtype=dtype, **light), g['lit'], rtol, 'lit textures')
    _close(so.camera_backward(v, g['eyes'], g['g_screen'], dtype=dtype, **cam), g['gv_cam'], rtol, 'camera backward')
    gt, gv = so.lighting_backward(v, f, g['textures'], g['g_lit'], dtype=dtype, **light)
    _close(gt, g['gt'], rtol, 'texture gradient')
    _close(gv, g['gv_light'], 10 * rtol, 'vertex gradient through the normals')


def test_torch_mirror_matches_reference_python():
    """gendr_b200's own Lighting / LookAt (the path taken when the scene steps are not fused) against the golden vectors."""
    import gendr_b200 as gd
    v, f = torch.from_numpy(GOLD['vertices']), torch.from_numpy(GOLD['faces'])
    mesh = gd.Mesh(v, f, torch.from_numpy(GOLD['a_textures']))
    lit = gd.Lighting()(mesh)
    cam = gd.LookAt(viewing_angle=15)
    cam.set_eyes(torch.from_numpy(GOLD['a_eyes']))
    out = cam(lit)
    assert torch.allclose(out.vertices, torch.from_numpy(GOLD['a_screen']), rtol=1e-5, atol=1e-6)
    assert torch.allclose(out.textures, torch.from_numpy(GOLD['a_lit']), rtol=1e-5, atol=1e-6)


def test_deferred_steps_need_cuda():
    """On CPU meshes nothing is deferred (the fused path exists only as CUDA kernels)."""
    import gendr_b200 as gd
    v, f = torch.from_numpy(GOLD['vertices']), torch.from_numpy(GOLD['faces'])
    m = gd.LookAt(viewing_angle=15)(gd.Lighting()(gd.Mesh(v, f)))
    assert m._pending_light is None and m._pending_camera is None


def test_materialising_deferred_steps_keeps_the_reference_order():
    """A mesh carrying deferred lighting + camera steps materialises to exactly what lighting(mesh) then transform(mesh)
    give: the normals are taken from the WORLD-space vertices (lighting first), whichever property is read first."""
    import gendr_b200 as gd
    v, f, t = torch.from_numpy(GOLD['vertices']), torch.from_numpy(GOLD['faces']), torch.from_numpy(GOLD['a_textures'])
    cam = gd.LookAt(viewing_angle=15)
    cam.set_eyes(torch.from_numpy(GOLD['a_eyes']))
    want = cam(gd.Lighting()(gd.Mesh(v, f, t)))
    for first in ('vertices', 'textures', 'face_vertices'):
        m = gd.Mesh(v, f, t, _pending_light=gd.Lighting(), _pending_camera=cam)
        getattr(m, first)
        assert m._pending_light is None and m._pending_camera is None
        assert torch.equal(m.vertices, want.vertices) and torch.equal(m.textures, want.textures)
