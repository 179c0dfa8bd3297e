"""GPU tests (pytest -m gpu) of SURVEY.md 8(f) rows 2 and 3: the camera-transform / lighting kernels and the fused scene
path (gendr_camera_*, gendr_lighting_*, gendr_scene_*), and the fused 2x anti-aliasing (gendr_*_render_aa), all through
the C ABI.

Checkers: the golden vectors generated from the reference Python (tests/golden/scene_v1.npz), the numpy oracle
(oracle/scene_oracle.py) on larger random inputs, torch's own avg_pool2d, and -- when baseline/_ref is staged -- the
reference's full Lighting -> LookAt -> GenDR pipeline on the same GPU.

Image criterion for end-to-end comparisons where the SCREEN-SPACE VERTICES come from different fp32 evaluation orders (our
camera kernel vs torch's normalize + cuBLAS bmm): the reference rasterizer is ill-conditioned on closed meshes.
Measured with the reference algorithm itself on the CPU (oracle, icosphere 1280 faces, orbit cameras): changing only the
summation order of the fp32 camera transform (vertex deltas <= 2.4e-7) moves 1.1 % (gaussian, 96^2) to 2.0 % (logistic,
128^2) of the RGBA values by more than 1e-4*|ref| + 1e-5, 99 % of them by < 2.2e-4 and single silhouette pixels by up to
0.61.  Hence: <= 5 % of values outside 1e-4 and a 99th-percentile |d| <= 1e-3 -- bit-level agreement is asserted where
the inputs are identical (test_fused_scene_equals_staged_kernels, the golden-vector tests).

Tolerances.  Camera / lighting are plain fp32 formulas whose reference implementation (torch reductions + cuBLAS bmm) has
no specified summation order: |d| <= 2e-6 * max|ref| (forward) and 2e-5 * max|ref| (gradients, sums of up to ~12 terms
per vertex).  The anti-aliased image is BIT-identical to avg_pool2d of the full-resolution image; gradients agree to
atomic-order noise (1e-5 * max|ref|)."""
import os

import numpy as np
import pytest
import torch

import scenes
from ref_gpu import load_reference

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'scene_v1.npz'))
CASES = {
    'a': (dict(mode='look_at', perspective=True, viewing_angle=15.), dict()),
    'b': (dict(mode='look', perspective=False, viewing_scale=0.8, direction=(0.1, -0.2, 1.0)),
          dict(intensity_ambient=0.3, color_ambient=(1, 0.9, 0.8), intensity_directional=0.7, color_directional=(0.5, 1, 0.7),
               direction=(0.3, 0.8, -0.5))),
}


def _dev():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda:0')


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _image_stats(new, ref):
    d = (new - ref).abs()
    bad = (d > 1e-4 * ref.abs() + 1e-5).float().mean().item()
    return bad, float(torch.quantile(d.flatten()[:2 ** 24].float(), 0.99)), float(d.max())


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def run_camera(dev, verts, eyes, g_screen, cam_kw):
    from gendr_b200 import _lib
    from gendr_b200.functional import make_camera_params
    lib = _lib.load()
    v = torch.as_tensor(verts, dtype=torch.float32, device=dev).contiguous()
    e = torch.as_tensor(eyes, dtype=torch.float32, device=dev).contiguous()
    g = torch.as_tensor(g_screen, dtype=torch.float32, device=dev).contiguous()
    B, V = v.shape[:2]
    cam = make_camera_params(**cam_kw)
    screen, gv = torch.empty_like(v), torch.full_like(v, float('nan'))
    _lib.check(lib.gendr_camera_forward(v.data_ptr(), e.data_ptr(), int(e.ndimension() == 2), screen.data_ptr(), B, V, cam, _stream(dev)))
    _lib.check(lib.gendr_camera_backward(v.data_ptr(), e.data_ptr(), int(e.ndimension() == 2), g.data_ptr(), gv.data_ptr(), B, V, cam, _stream(dev)))
    torch.cuda.synchronize()
    return screen.cpu().numpy(), gv.cpu().numpy()


def run_lighting(dev, verts, faces, tex, g_lit, light_kw, shared=False):
    from gendr_b200 import _lib
    from gendr_b200.functional import make_light_params
    lib = _lib.load()
    v = torch.as_tensor(verts, dtype=torch.float32, device=dev).contiguous()
    f = torch.as_tensor(faces, dtype=torch.int32, device=dev).contiguous()
    t = torch.as_tensor(tex, dtype=torch.float32, device=dev).contiguous()
    g = torch.as_tensor(g_lit, dtype=torch.float32, device=dev).contiguous()
    B, V = v.shape[:2]
    F, T = t.shape[1], t.shape[2]
    light = make_light_params(**light_kw)
    lit, gt, gv = torch.empty_like(t), torch.full_like(t, float('nan')), torch.zeros_like(v)
    _lib.check(lib.gendr_lighting_forward(v.data_ptr(), f.data_ptr(), int(shared), t.data_ptr(), lit.data_ptr(), B, V, F, T, light, _stream(dev)))
    _lib.check(lib.gendr_lighting_backward(v.data_ptr(), f.data_ptr(), int(shared), t.data_ptr(), g.data_ptr(), gt.data_ptr(), gv.data_ptr(),
                                           B, V, F, T, light, _stream(dev)))
    torch.cuda.synchronize()
    return lit.cpu().numpy(), gt.cpu().numpy(), gv.cpu().numpy()


@pytest.mark.parametrize('name', ['a', 'b'])
def test_camera_and_lighting_kernels_vs_reference_golden(name):
    dev = _dev()
    cam_kw, light_kw = CASES[name]
    g = {k: GOLD['%s_%s' % (name, k)] for k in ('textures', 'eyes', 'g_screen', 'g_lit', 'screen', 'lit', 'gv_cam', 'gv_light', 'gt')}
    screen, gv_cam = run_camera(dev, GOLD['vertices'], g['eyes'], g['g_screen'], cam_kw)
    assert _rel(screen, g['screen']) <= 2e-6 and _rel(gv_cam, g['gv_cam']) <= 2e-5, (_rel(screen, g['screen']), _rel(gv_cam, g['gv_cam']))
    lit, gt, gv = run_lighting(dev, GOLD['vertices'], GOLD['faces'], g['textures'], g['g_lit'], light_kw)
    assert _rel(lit, g['lit']) <= 2e-6 and _rel(gt, g['gt']) <= 2e-6 and _rel(gv, g['gv_light']) <= 2e-5, (
        _rel(lit, g['lit']), _rel(gt, g['gt']), _rel(gv, g['gv_light']))


def test_camera_and_lighting_kernels_vs_oracle_large():
    """C3-sized mesh (4225 vertices, 8192 faces), batch 5, shared index buffer, degenerate + back-lit faces included."""
    from oracle import scene_oracle as so
    dev = _dev()
    rng = np.random.default_rng(3)
    verts, faces = scenes.grid_sphere(64)
    B = 5
    v = (verts[None].repeat(B, 1, 1).numpy() * (1 + 0.1 * rng.standard_normal((B, 1, 1)))).astype(np.float32)
    f = faces.numpy().copy()
    f[7] = [5, 5, 9]                                             # zero-area face: clamped normal (norm <= eps)
    eyes = np.stack([np.array(scenes.gd.functional.get_points_from_angles(2.732, 30., 72. * i), np.float32) for i in range(B)])
    g_screen = rng.standard_normal(v.shape).astype(np.float32)
    tex = rng.random((B, f.shape[0], 1, 3)).astype(np.float32)
    g_lit = rng.standard_normal(tex.shape).astype(np.float32)
    cam_kw, light_kw = CASES['a']
    screen, gv_cam = run_camera(dev, v, eyes, g_screen, cam_kw)
    assert _rel(screen, so.camera_forward(v, eyes, **cam_kw)) <= 2e-6
    assert _rel(gv_cam, so.camera_backward(v, eyes, g_screen, **cam_kw)) <= 2e-5
    light_kw = dict(light_kw, direction=(0.2, 0.9, 0.3))
    lit, gt, gv = run_lighting(dev, v, f, tex, g_lit, light_kw, shared=True)
    e_gt, e_gv = so.lighting_backward(v, f, tex, g_lit, **light_kw)
    assert _rel(lit, so.lighting_forward(v, f, tex, **light_kw)) <= 2e-6
    assert _rel(gt, e_gt) <= 2e-6 and _rel(gv, e_gv) <= 2e-5, (_rel(gt, e_gt), _rel(gv, e_gv))


def _scene_inputs(dev, B=4, sub=2, T=1, seed=0):
    verts, faces = scenes.icosphere(sub)
    g = torch.Generator().manual_seed(seed)
    v = (verts * 0.5)[None].repeat(B, 1, 1) * (1 + 0.1 * torch.rand(B, verts.shape[0], 1, generator=g))
    f = faces[None].repeat(B, 1, 1)
    tex = torch.rand(B, faces.shape[0], T, 3, generator=g)
    eyes = scenes.orbit_eyes(B)
    return v.to(dev), f.to(dev), tex.to(dev), eyes


@pytest.mark.parametrize('aa', [False, True])
def test_fused_scene_equals_staged_kernels(aa):
    """gendr_scene_* == gendr_camera_forward + gendr_lighting_forward + render_indexed + their backward kernels chained by
    hand: images bit-identical, gradients to atomic-order noise."""
    import gendr_b200 as gd
    from gendr_b200 import _lib
    from gendr_b200.functional import make_camera_params, make_light_params
    dev = _dev()
    v, f, tex, eyes = _scene_inputs(dev, T=4)
    S = 64
    kw = dict(image_size=S * (2 if aa else 1), dist_func='logistic', aggr_alpha_func='probabilistic', dist_scale=0.01, double_side=False,
              anti_aliasing=aa)
    cam_kw, light_kw = dict(mode='look_at', viewing_angle=15.), dict(direction=(0.3, 0.8, -0.5))
    g_img = torch.randn(4, 4, S, S, generator=torch.Generator().manual_seed(5)).to(dev)
    a, t = v.clone().requires_grad_(True), tex.clone().requires_grad_(True)
    img = gd.functional.render_scene(a, f, t, eyes, camera=cam_kw, lighting=light_kw, **kw)
    img.backward(g_img)
    # staged
    lib = _lib.load()
    B, V = v.shape[:2]
    F, T = tex.shape[1], tex.shape[2]
    e = torch.as_tensor(eyes, dtype=torch.float32, device=dev).contiguous()
    cam, light = make_camera_params(**cam_kw), make_light_params(**light_kw)
    fi = f.int().contiguous()
    screen, lit = torch.empty_like(v), torch.empty_like(tex)
    _lib.check(lib.gendr_camera_forward(v.data_ptr(), e.data_ptr(), 1, screen.data_ptr(), B, V, cam, _stream(dev)))
    _lib.check(lib.gendr_lighting_forward(v.data_ptr(), fi.data_ptr(), 0, tex.data_ptr(), lit.data_ptr(), B, V, F, T, light, _stream(dev)))
    s2, l2 = screen.clone().requires_grad_(True), lit.clone().requires_grad_(True)
    img2 = gd.functional.render_indexed(s2, fi, l2, **kw)
    img2.backward(g_img)
    assert torch.equal(img, img2)
    gv, gt = torch.empty_like(v), torch.empty_like(tex)
    _lib.check(lib.gendr_camera_backward(v.data_ptr(), e.data_ptr(), 1, s2.grad.contiguous().data_ptr(), gv.data_ptr(), B, V, cam, _stream(dev)))
    _lib.check(lib.gendr_lighting_backward(v.data_ptr(), fi.data_ptr(), 0, tex.data_ptr(), l2.grad.contiguous().data_ptr(), gt.data_ptr(), gv.data_ptr(),
                                           B, V, F, T, light, _stream(dev)))
    torch.cuda.synchronize()
    assert _rel(a.grad.cpu(), gv.cpu()) <= 1e-5 and _rel(t.grad.cpu(), gt.cpu()) <= 1e-5, (_rel(a.grad.cpu(), gv.cpu()), _rel(t.grad.cpu(), gt.cpu()))


@pytest.mark.parametrize('aa', [False, True])
def test_module_pipeline_deferred_vs_torch_glue(aa):
    """lighting(mesh); transform(mesh); renderer(mesh) -- the deferred/fused path against the same modules with the deferral
    switched off (torch glue + indexed render).  Screen-space vertices differ by ulps (cuBLAS bmm vs our kernel), which the
    rasterizer amplifies on silhouette pixels, hence the statistical criterion on the image."""
    import gendr_b200 as gd
    from gendr_b200 import mesh as mesh_mod
    dev = _dev()
    v, f, tex, eyes = _scene_inputs(dev, B=4, sub=3, T=1)
    S = 96
    renderer = gd.GenDR(image_size=S, anti_aliasing=aa, dist_func='gaussian', aggr_alpha_func='einstein', dist_scale=0.01)
    lighting, cam = gd.Lighting(), gd.LookAt(viewing_angle=15)
    cam.set_eyes(eyes)
    g_img = torch.randn(4, 4, S, S, generator=torch.Generator().manual_seed(6)).to(dev)
    res = {}
    for fuse in (True, False):
        mesh_mod.FUSE_SCENE = fuse
        try:
            a, t = v.clone().requires_grad_(True), tex.clone().requires_grad_(True)
            m = cam(lighting(gd.Mesh(a, f, t)))
            assert (m._pending_camera is not None) == fuse and (m._pending_light is not None) == fuse
            img = renderer(m)
            img.backward(g_img)
            res[fuse] = (img.detach().cpu(), a.grad.cpu(), t.grad.cpu())
        finally:
            mesh_mod.FUSE_SCENE = True
    (i1, gv1, gt1), (i0, gv0, gt0) = res[True], res[False]
    assert i1.shape == (4, 4, S, S)
    bad, p99, dmax = _image_stats(i1, i0)
    print('deferred vs torch glue (aa=%s): frac outside %.2e, p99 |d| %.2e, max |d| %.2e; grad rel %.2e / %.2e' % (
        aa, bad, p99, dmax, _rel(gv1, gv0), _rel(gt1, gt0)))
    assert bad <= 5e-2 and p99 <= 1e-3, (bad, p99, dmax)
    assert _rel(gv1, gv0) <= 5e-2 and _rel(gt1, gt0) <= 5e-2, (_rel(gv1, gv0), _rel(gt1, gt0))
    # materialising a deferred mesh gives what the torch path gives
    m = cam(lighting(gd.Mesh(v, f, tex)))
    mesh_mod.FUSE_SCENE = False
    try:
        m0 = cam(lighting(gd.Mesh(v, f, tex)))
    finally:
        mesh_mod.FUSE_SCENE = True
    assert torch.equal(m.vertices, m0.vertices) and torch.equal(m.textures, m0.textures)


@pytest.mark.parametrize('cfg', [dict(dist_func='logistic', aggr_alpha_func='probabilistic'),
                                 dict(dist_func='uniform', aggr_alpha_func='max', aggr_rgb_func='hard'),
                                 dict(dist_func='cauchy', aggr_alpha_func='yager', aggr_alpha_t_conorm_p=2.0, texture_type='vertex')])
def test_fused_anti_aliasing_bit_identical_to_avg_pool(cfg):
    import gendr_b200 as gd
    import torch.nn.functional as F
    dev = _dev()
    fv, ft = scenes.soup(300, batch=3, seed=4, size=0.2)
    if cfg.get('texture_type') == 'vertex':
        ft = torch.rand(3, fv.shape[1], 3, 3, generator=torch.Generator().manual_seed(1))
    S = 50                                                        # output side; rendered at 100 (ragged tiles: 100 = 6*16 + 4)
    kw = dict(image_size=2 * S, dist_scale=0.02, background_color=[0.1, 0.3, 0.5], **cfg)
    g = torch.randn(3, 4, S, S, generator=torch.Generator().manual_seed(2)).to(dev)
    a, b = fv.to(dev).requires_grad_(True), ft.to(dev).requires_grad_(True)
    img = gd.functional.render(a, b, anti_aliasing=True, **kw)
    img.backward(g)
    a2, b2 = fv.to(dev).requires_grad_(True), ft.to(dev).requires_grad_(True)
    img2 = F.avg_pool2d(gd.functional.render(a2, b2, **kw), kernel_size=2, stride=2)
    img2.backward(g)
    assert img.shape == (3, 4, S, S) and torch.equal(img, img2)
    assert _rel(a.grad.cpu(), a2.grad.cpu()) <= 1e-5 and _rel(b.grad.cpu(), b2.grad.cpu()) <= 1e-5


def test_anti_aliasing_needs_even_size():
    import gendr_b200 as gd
    from gendr_b200._lib import GendrCudaError
    dev = _dev()
    fv, ft = scenes.soup(10, batch=1, seed=4, size=0.2)
    with pytest.raises(GendrCudaError):
        gd.functional.render(fv.to(dev), ft.to(dev), image_size=33, anti_aliasing=True)


def test_full_pipeline_vs_reference_cuda():
    """Mesh -> Lighting -> LookAt -> GenDR(anti_aliasing) fwd+bwd: ours (fused scene path) vs the reference package end to end."""
    import gendr_b200 as gd
    dev = _dev()
    ref = load_reference()
    if ref is None:
        pytest.skip('baseline/_ref (reference CUDA build) not staged')
    v, f, tex, eyes = _scene_inputs(dev, B=4, sub=3, T=1)
    S = 64
    g_img = torch.randn(4, 4, S, S, generator=torch.Generator().manual_seed(8)).to(dev)
    cfg = dict(image_size=S, anti_aliasing=True, dist_func='logistic', aggr_alpha_func='probabilistic', dist_scale=0.01, dist_shape=0.0,
               dist_shift=0.0, aggr_alpha_t_conorm_p=0.0)
    out = {}
    for name, pkg in (('ours', gd), ('ref', ref)):
        a, t = v.clone().requires_grad_(True), tex.clone().requires_grad_(True)
        cam = pkg.LookAt(viewing_angle=15)
        cam.set_eyes(eyes.to(dev))
        img = pkg.GenDR(**cfg)(cam(pkg.Lighting()(pkg.Mesh(a, f, t))))
        img.backward(g_img)
        torch.cuda.synchronize()
        out[name] = (img.detach().cpu(), a.grad.cpu(), t.grad.cpu())
    (i1, gv1, gt1), (i0, gv0, gt0) = out['ours'], out['ref']
    bad, p99, dmax = _image_stats(i1, i0)
    print('full pipeline vs reference: frac outside %.2e, p99 |d| %.2e, max |d| %.2e; grad rel %.2e / %.2e' % (
        bad, p99, dmax, _rel(gv1, gv0), _rel(gt1, gt0)))
    assert bad <= 5e-2 and p99 <= 1e-3, (bad, p99, dmax)
    assert _rel(gv1, gv0) <= 5e-2 and _rel(gt1, gt0) <= 5e-2, (_rel(gv1, gv0), _rel(gt1, gt0))


def test_camera_gradient_takes_the_torch_path():
    """An eye that requires grad (experiments/opt_camera.py) is not deferred: look_at runs as torch ops, the rasterizer through the
    indexed path, and d loss / d eye agrees with the reference package."""
    import gendr_b200 as gd
    dev = _dev()
    ref = load_reference()
    v, f, tex, _ = _scene_inputs(dev, B=4, sub=2, T=1)       # not 3: the reference's torch.cross without `dim` (mesh.py:108) would pick the batch axis
    g_img = torch.randn(4, 4, 64, 64, generator=torch.Generator().manual_seed(9)).to(dev)
    grads = {}
    for name, pkg in (('ours', gd), ('ref', ref)):
        if pkg is None:
            continue
        eye = torch.tensor([[0.3, 0.8, -2.6], [1.2, 0.5, -2.3], [-0.9, 0.2, -2.5], [0.0, 1.5, -2.2]], device=dev, requires_grad=True)
        cam = pkg.LookAt(viewing_angle=15)
        cam.set_eyes(eye)
        mesh = cam(pkg.Lighting()(pkg.Mesh(v, f, tex)))
        if pkg is gd:
            assert mesh._pending_camera is None and mesh._pending_light is None      # both steps ran as torch ops
        img = pkg.GenDR(image_size=64, dist_func='logistic', dist_scale=0.03, dist_shape=0.0, dist_shift=0.0, aggr_alpha_t_conorm_p=0.0)(mesh)
        img.backward(g_img)
        grads[name] = eye.grad.cpu()
    assert torch.isfinite(grads['ours']).all() and float(grads['ours'].abs().max()) > 0
    if 'ref' in grads:
        # all but the LAST batch item: its last face samples the texel behind the texture buffer in the reference (quirk Q3:
        # undefined there, 0 here), which changes that item's colours and gradients
        assert _rel(grads['ours'][:-1], grads['ref'][:-1]) <= 1e-3, _rel(grads['ours'][:-1], grads['ref'][:-1])
