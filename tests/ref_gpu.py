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
