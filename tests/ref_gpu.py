"""Loader for the UNMODIFIED reference package staged under baseline/_ref/ (git-ignored; built for sm_100a by
baseline/build_ref.sh).  Used on the GPU box only, as the strongest parity oracle (the reference's own CUDA kernels)
and as the "reference_cuda" timing in bench.py.  Returns None when the build is not present."""
import importlib
import os
import sys
import types

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(_ROOT, 'baseline', '_ref')


def load_reference():
    if not os.path.isdir(os.path.join(REF_DIR, 'gendr')):
        return None
    import glob
    if not glob.glob(os.path.join(REF_DIR, 'gendr', 'cuda', 'generalized_renderer*.so')):
        return None
    if 'skimage' not in sys.modules:            # only imread/imsave are referenced (load_obj.py:9, save_obj.py:8)
        sk, skio = types.ModuleType('skimage'), types.ModuleType('skimage.io')
        skio.imread = skio.imsave = lambda *a, **k: (_ for _ in ()).throw(RuntimeError('skimage stub'))
        sk.io = skio
        sys.modules['skimage'], sys.modules['skimage.io'] = sk, skio
    # (the three off-path extensions are replaced by empty stand-in modules at staging time, baseline/build_ref.sh)
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    try:
        return importlib.import_module('gendr')
    except Exception as e:      # pragma: no cover
        print('reference import failed:', repr(e))
        return None


def reference_render(gendr_ref, fv, ft, **kw):
    """gendr.functional.render with None -> 0.0 for the three optional parameters."""
    for k in ('dist_shape', 'dist_shift', 'aggr_alpha_t_conorm_p'):
        if kw.get(k) is None:
            kw[k] = 0.0
    return gendr_ref.functional.render(fv, ft, **kw)


# name -> id maps of the reference (gendr/functional/renderer.py:44-83; local dicts there, so restated here)
_DIST = {'hard': 0, 'heaviside': 0, 'uniform': 1, 'cubic_hermite': 2, 'wigner_semicircle': 3, 'gaussian': 4, 'laplace': 5,
         'logistic': 6, 'gudermannian': 7, 'hyperbolic_secant': 7, 'cauchy': 8, 'reciprocal': 9, 'gumbel_max': 10, 'gumbel_min': 11,
         'exponential': 12, 'exponential_rev': 13, 'gamma': 14, 'gamma_rev': 15, 'levy': 16, 'levy_rev': 17}
_TCN = {'hard': 0, 'max': 1, 'probabilistic': 2, 'einstein': 3, 'hamacher': 4, 'frank': 5, 'yager': 6, 'aczel_alsina': 7, 'dombi': 8,
        'schweizer_sklar': 9}


def reference_render_raw(gendr_ref, fv, ft, g, dtype, image_size=256, background_color=(0, 0, 0), dist_func='uniform', dist_scale=1e-2,
                         dist_squared=False, dist_shape=None, dist_shift=None, dist_eps=1e4, aggr_alpha_func='probabilistic',
                         aggr_alpha_t_conorm_p=None, aggr_rgb_func='softmax', aggr_rgb_eps=1e-3, aggr_rgb_gamma=1e-3, near=1, far=100,
                         double_side=True, texture_type='surface'):
    """The reference's pybind functions called directly with buffers of `dtype` (with ALL buffers fp64 the
    AT_DISPATCH_FLOATING_TYPES switch at K.cu:1099 reaches the <double> instantiation -- the accuracy reference for the
    parity context numbers).  Buffer allocation follows gendr/functional/renderer.py:130-151 and :191-197.
    Returns (soft_colors, grad_faces, grad_textures)."""
    import torch
    ext = importlib.import_module('gendr.cuda.generalized_renderer')
    dev = fv.device
    B, F = fv.shape[:2]
    S = int(image_size)
    faces, tex = fv.to(dtype).contiguous().clone(), ft.to(dtype).contiguous().clone()
    faces_info = torch.zeros((B, F, 27), dtype=dtype, device=dev)
    aggrs = torch.zeros((B, 2, S, S), dtype=dtype, device=dev)
    colors = torch.ones((B, 4, S, S), dtype=dtype, device=dev)
    for k in range(3):
        colors[:, k] *= background_color[k]
    scal = (S, _DIST[dist_func] if isinstance(dist_func, str) else dist_func, float(dist_scale), bool(dist_squared), float(dist_shape or 0.0),
            float(dist_shift or 0.0), float(dist_eps), _TCN[aggr_alpha_func] if isinstance(aggr_alpha_func, str) else aggr_alpha_func,
            float(aggr_alpha_t_conorm_p or 0.0), {'hard': 0, 'softmax': 1}[aggr_rgb_func], float(aggr_rgb_eps), float(aggr_rgb_gamma),
            float(near), float(far), bool(double_side), {'surface': 0, 'vertex': 1}[texture_type])
    ext.forward_render(faces, tex, faces_info, aggrs, colors, *scal)
    gf, gt = torch.zeros_like(faces), torch.zeros_like(tex)
    ext.backward_render(faces, tex, colors, faces_info, aggrs, gf, gt, g.to(dtype).contiguous(), *scal)
    return colors, gf, gt
