"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/gendr_b200.h declares, the
Python mirror keeps the reference's surface (names, defaults, id maps, validation), and the product refuses to run
without CUDA (no fallback).  No compute calls are made here."""
import ctypes as C
import inspect
import os
import re

import pytest
import torch

import gendr_b200 as gd
from gendr_b200 import _lib
from gendr_b200.functional import renderer as fr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, 'include', 'gendr_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(gendr_[a-z_0-9]+)\s*\(', text)))


def test_c_abi_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = C.CDLL(_lib.LIB_PATH)
    names = header_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), 'missing export ' + n
    assert sorted(_lib.SIGNATURES) == names, 'python binding table and header disagree'
    lib.gendr_version.restype = C.c_char_p
    assert b'sm_100a' in lib.gendr_version()
    lib.gendr_workspace_bytes.restype = C.c_size_t
    assert lib.gendr_workspace_bytes(64, 8192) >= 64 * 8192 * (144 + 8)


def test_params_struct_layout_matches_header():
    text = open(os.path.join(ROOT, 'include', 'gendr_b200.h')).read()
    body = re.search(r'typedef struct gendr_render_params \{(.*?)\} gendr_render_params;', text, flags=re.S).group(1)
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    fields = [(t, n) for t, n in re.findall(r'\b(int|float)\s+([a-z_]+)(?:\[3\])?;', body)]
    py = [(('int' if f[1] is C.c_int else 'float'), f[0]) for f in _lib.RenderParams._fields_]
    assert fields == py
    assert C.sizeof(_lib.RenderParams) == 19 * 4


def test_id_maps_are_the_references():
    # gendr/functional/renderer.py:44-83
    assert fr.DIST_FUNC_IDS == {'hard': 0, 'heaviside': 0, 'uniform': 1, 'cubic_hermite': 2, 'wigner_semicircle': 3, 'gaussian': 4,
                                'laplace': 5, 'logistic': 6, 'gudermannian': 7, 'hyperbolic_secant': 7, 'cauchy': 8, 'reciprocal': 9,
                                'gumbel_max': 10, 'gumbel_min': 11, 'exponential': 12, 'exponential_rev': 13, 'gamma': 14,
                                'gamma_rev': 15, 'levy': 16, 'levy_rev': 17}
    assert fr.AGGR_ALPHA_FUNC_IDS == {'hard': 0, 'max': 1, 'probabilistic': 2, 'einstein': 3, 'hamacher': 4, 'frank': 5, 'yager': 6,
                                      'aczel_alsina': 7, 'dombi': 8, 'schweizer_sklar': 9}
    assert fr.AGGR_RGB_FUNC_IDS == {'hard': 0, 'softmax': 1} and fr.TEXTURE_TYPE_IDS == {'surface': 0, 'vertex': 1}


def test_gendr_module_surface():
    sig = inspect.signature(gd.GenDR.__init__)
    want = dict(image_size=256, background_color=[0, 0, 0], anti_aliasing=False, dist_func='uniform', dist_scale=1e-2,
                dist_squared=False, dist_shape=None, dist_shift=None, dist_eps=1e4, aggr_alpha_func='probabilistic',
                aggr_alpha_t_conorm_p=None, aggr_rgb_func='softmax', aggr_rgb_eps=1e-3, aggr_rgb_gamma=1e-3, near=1, far=100,
                double_side=False, texture_type='surface')              # gendr/renderer.py:13-36
    assert {k: v.default for k, v in sig.parameters.items() if k != 'self'} == want
    r = gd.GenDR(dist_func='gaussian', dist_scale=0.03)
    r.dist_scale = 0.5                                                  # attributes stay mutable between calls
    assert r.dist_scale == 0.5 and r.dist_func == 'gaussian'
    with pytest.raises(ValueError):
        gd.GenDR(aggr_rgb_func='mean')
    with pytest.raises(ValueError):
        gd.GenDR(texture_type='volume')
    rsig = inspect.signature(gd.functional.render)
    assert rsig.parameters['double_side'].default is True               # functional/renderer.py:262
    assert list(rsig.parameters)[:4] == ['face_vertices', 'textures', 'image_size', 'background_color']


def test_no_cpu_fallback():
    fv, ft = torch.zeros(1, 2, 3, 3), torch.zeros(1, 2, 1, 3)
    with pytest.raises((TypeError, RuntimeError)):
        gd.functional.render(fv, ft, image_size=8)
    from gendr_b200.cuda import generalized_renderer as ext
    with pytest.raises(RuntimeError):
        ext.forward_render(fv.view(1, 2, 9), ft, torch.zeros(1, 2, 27), torch.zeros(1, 2, 8, 8), torch.ones(1, 4, 8, 8),
                           8, 1, 1e-2, False, 0., 0., 1e4, 2, 0., 1, 1e-3, 1e-3, 1., 100., True, 0)
    with pytest.raises(AssertionError):
        gd.functional.render(fv.cuda() if torch.cuda.is_available() else fv, ft, image_size=8, dist_scale=-1.0)
    for path in ('gendr_b200/_lib.py', 'gendr_b200/functional/renderer.py', 'gendr_b200/cuda/generalized_renderer.py'):
        src = open(os.path.join(ROOT, path)).read()
        assert 'oracle' not in src.replace('# ', ''), path + ' must not reference the oracle'


def test_mesh_camera_lighting_pipeline_on_cpu():
    import scenes
    verts, faces = scenes.icosphere(1)
    assert verts.shape == (42, 3) and faces.shape == (80, 3)
    mesh = gd.Mesh((verts * 0.5)[None].repeat(3, 1, 1), faces[None].repeat(3, 1, 1))
    assert mesh.face_vertices.shape == (3, 80, 3, 3) and mesh.face_textures.shape == (3, 80, 1, 3)
    lit = gd.Lighting()(mesh)
    assert float(lit.textures.min()) >= 0.5 - 1e-6 and float(lit.textures.max()) <= 1.0 + 1e-6
    cam = gd.LookAt(viewing_angle=15)
    cam.set_eyes_from_angles(torch.full((3,), 2.732), torch.full((3,), 30.), torch.tensor([0., 120., 240.]))
    out = cam(lit)
    z = out.vertices[..., 2]
    assert float(z.min()) > 1.0 and float(out.vertices[..., :2].abs().max()) < 1.0
    vm = gd.Mesh(verts, faces, texture_type='vertex')
    assert vm.face_textures.shape == (1, 80, 3, 3) and gd.Lighting()(vm).textures.shape == (1, 42, 3)
    eye = gd.functional.get_points_from_angles(2., 0, 0)
    assert abs(eye[2] + 2.0) < 1e-6 and abs(eye[0]) < 1e-6
    v2 = gd.functional.look_at(torch.zeros(1, 1, 3), eye)
    assert torch.allclose(v2, torch.tensor([[[0., 0., 2.]]]), atol=1e-6)
    fv, ft, cfg = scenes.config_c1()
    want = torch.tensor([[0, 0.2693377435, 2], [-0.3110042512, -0.2693377435, 2], [0.3110042512, -0.2693377435, 2]])
    assert torch.allclose(fv[0, 0], want, atol=1e-6) and torch.allclose(ft, torch.full_like(ft, 0.5))


def test_regularisers_match_the_reference_python():
    """LaplacianLoss / FlattenLoss (gendr/losses.py) against the reference's own modules on a jittered icosphere, and their
    defining properties (flat neighbourhoods cost nothing)."""
    import scenes
    from ref_gpu import load_reference
    verts, faces = scenes.icosphere(2)
    g = torch.Generator().manual_seed(0)
    x = ((verts * 0.5)[None].repeat(3, 1, 1) + 0.02 * torch.randn(3, verts.shape[0], 3, generator=g)).requires_grad_(True)
    lap, flat = gd.LaplacianLoss(verts, faces, average=True), gd.FlattenLoss(faces, average=False)
    a, b = lap(x), flat(x)
    assert a.ndim == 0 and b.shape == (3,) and float(a) > 0 and float(b.min()) > 0
    (a + b.sum()).backward()
    assert torch.isfinite(x.grad).all()
    ref = load_reference()
    if ref is not None:
        ra, rb = ref.LaplacianLoss(verts, faces, average=True)(x), ref.FlattenLoss(faces, average=False)(x)
        assert torch.allclose(a, ra, rtol=1e-5) and torch.allclose(b, rb, rtol=1e-4)
    # a flat square split into two triangles: the shared edge is flat -> cos(dihedral) = -1 -> zero loss
    quad = torch.tensor([[0., 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]])
    loss = gd.FlattenLoss.__new__(gd.FlattenLoss)
    torch.nn.Module.__init__(loss)
    loss.nf, loss.average = 2, False
    for name, idx in (('v0s', [0]), ('v1s', [2]), ('v2s', [1]), ('v3s', [3])):
        loss.register_buffer(name, torch.tensor(idx))
    assert float(loss(quad[None])) < 1e-4


def test_binding_table_matches_header_signatures():
    """Every entry of gendr_b200._lib.SIGNATURES has the argument count and the argument kinds (pointer / int / float / size_t) of
    the C declaration in include/gendr_b200.h -- a mismatch would corrupt the call stack silently under ctypes."""
    text = open(os.path.join(ROOT, 'include', 'gendr_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    decls = dict(re.findall(r'\b(gendr_[a-z_0-9]+)\s*\(([^;{]*?)\)\s*;', text, flags=re.S))
    assert sorted(decls) == sorted(_lib.SIGNATURES)

    def kind_of_c(arg):
        arg = ' '.join(arg.split())
        if arg in ('void', ''):
            return None
        if '*' in arg:
            return 'ptr'
        base = arg.rsplit(' ', 1)[0] if ' ' in arg else arg
        return {'int': 'int', 'float': 'float', 'size_t': 'size_t', 'long long': 'longlong'}[base.replace('const ', '')]

    def kind_of_py(t):
        if t in (C.c_void_p,) or hasattr(t, 'contents') or t is C.c_char_p:
            return 'ptr'
        return {C.c_int: 'int', C.c_float: 'float', C.c_size_t: 'size_t', C.c_longlong: 'longlong'}[t]

    for name, args in decls.items():
        c_kinds = [k for k in (kind_of_c(a) for a in args.split(',')) if k is not None]
        py_kinds = [kind_of_py(t) for t in _lib.SIGNATURES[name][1]]
        assert c_kinds == py_kinds, (name, c_kinds, py_kinds)
    for cls, cname in ((_lib.CameraParams, 'gendr_camera_params'), (_lib.LightParams, 'gendr_light_params')):
        body = re.search(r'typedef struct %s \{(.*?)\} %s;' % (cname, cname), text, flags=re.S).group(1)
        fields = [(t, n) for t, n in re.findall(r'\b(int|float)\s+([a-z_]+)(?:\[3\])?;', body)]
        py = [(('int' if f[1] is C.c_int else 'float'), f[0]) for f in cls._fields_]
        assert fields == py, (cname, fields, py)


def test_bench_reference_arm_contract_on_cpu():
    """`bench.py --impl reference` without a GPU falls back to the CPU shim of the reference kernels and still prints exactly one
    JSON line with the keys of the contract."""
    import json
    import subprocess
    import sys
    if torch.cuda.is_available():
        pytest.skip('a GPU is present: --impl reference times the reference CUDA kernels there (host-to-device bytes are not 0)')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1', '--workload', 'c2'],
                       capture_output=True, text=True, timeout=600)
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert r.returncode == 0 and len(lines) == 1, (r.returncode, r.stdout[-500:], r.stderr[-500:])
    d = json.loads(lines[0])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'dtype', 'data',
                'config', 'cpu_baseline', 'e2e', 'gpu_launches'):
        assert key in d, key
    assert d['impl'] == 'reference' and d['value'] > 0 and d['e2e']['h2d_bytes_per_step'] == 0
    assert d['cpu_baseline']['kind'] in ('reference', 'port') and 'workload' in d['config']


def test_cpp_autograd_node_is_built_and_consistent():
    """gendr_b200/_torchbind (csrc/torch_binding.cpp): the C++ autograd node of functional.render.  Host code over the same C ABI;
    its parameter struct must be the header's, and it must refuse CPU tensors like the Python node (no fallback)."""
    import glob
    if not glob.glob(os.path.join(ROOT, 'gendr_b200', '_torchbind*.so')):
        pytest.skip('gendr_b200/_torchbind not built (make -C gendr_b200/csrc torchbind)')
    from gendr_b200 import _torchbind
    assert _torchbind.params_size() == C.sizeof(_lib.RenderParams)
    p = _lib.RenderParams()
    p.image_size = 8
    with pytest.raises(TypeError):
        _torchbind.render_faces(torch.zeros(1, 1, 3, 3), torch.zeros(1, 1, 1, 3), C.addressof(p), False)


def test_source_fingerprint_ignores_comments_only(tmp_path, monkeypatch):
    """bench.csrc_sha keys the ncu-derived roofline constants on the kernel CODE: comments and whitespace must not change it,
    any token must (a mismatch with the committed constants only warns: bench.py reports traffic = null then)."""
    import json
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    committed = json.load(open(os.path.join(bench.ROOT, 'profiles', 'roofline_traffic.json')))
    if committed['csrc_sha'] != bench.csrc_sha():      # not an error: bench.py then reports roofline.traffic = null with the reason
        import warnings
        warnings.warn('profiles/roofline_traffic.json was captured from other kernel sources: re-run tools/gpu_r2_final.sh + collect_profiles.py')
    d = tmp_path / 'gendr_b200' / 'csrc'
    d.mkdir(parents=True)
    monkeypatch.setattr(bench, 'ROOT', str(tmp_path))
    (d / 'k.cuh').write_text('__global__ void k(float* p) {\n    p[0] = 1.f;   // one\n}\n')
    a = bench.csrc_sha()
    (d / 'k.cuh').write_text('/* header\n   comment */\n__global__ void k(float* p)\n{\n  p[0] = 1.f;      // a different comment\n}\n')
    assert bench.csrc_sha() == a
    (d / 'k.cuh').write_text('__global__ void k(float* p) {\n    p[0] = 2.f;   // one\n}\n')
    assert bench.csrc_sha() != a
