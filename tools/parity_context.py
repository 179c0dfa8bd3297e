#!/usr/bin/env python
"""Parity context numbers of SURVEY.md 8(d) ("always report three context numbers beside parity"), measured on the B200 box:

  (1) reference vs reference      run-to-run noise of the reference's own CUDA kernels (atomic order in backward)
  (2) reference fp32 vs <double>  the reference's fp32 kernels against its own <double> instantiation (pybind module called
                                  with fp64 buffers; K.cu:1099 AT_DISPATCH_FLOATING_TYPES)
  (3) ours vs <double>            this repo's kernels against the same <double> result
  (+) ours vs reference fp32      the parity figure itself (criterion of tests/test_gpu_parity.py)

    python tools/parity_context.py [--out gpurun_out/parity_r2.json] [--quick]

Each entry: max |d|, 99.9th percentile |d|, fraction of elements outside `1e-4*|ref| + atol` (atol 1e-5 RGBA, 1e-4*max|ref| grads).
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import torch  # noqa: E402

import scenes  # noqa: E402
from ref_gpu import load_reference, reference_render, reference_render_raw  # noqa: E402


def stats(new, ref, atol_abs=None):
    new, ref = new.detach().double().flatten(), ref.detach().double().flatten()
    m = ~(torch.isnan(ref) | torch.isnan(new))
    nan_mismatch = int((torch.isnan(ref) != torch.isnan(new)).sum())
    d = (new[m] - ref[m]).abs()
    scale = float(ref[m].abs().max()) if d.numel() else 0.0
    atol = atol_abs if atol_abs is not None else 1e-4 * scale
    bad = int((d > 1e-4 * ref[m].abs() + atol).sum())
    k = max(1, int(d.numel() * 0.999))
    p999 = float(d.kthvalue(k).values) if d.numel() else 0.0
    return {'max_abs': float(d.max()) if d.numel() else 0.0, 'p99.9_abs': p999, 'max_ref': scale,
            'max_over_max_ref': (float(d.max()) / scale) if scale else 0.0,
            'frac_outside_tol': bad / max(1, d.numel()), 'nan_mask_mismatches': nan_mismatch}


def triple(a, b):
    return {'rgba': stats(a[0], b[0], atol_abs=1e-5), 'grad_faces': stats(a[1], b[1]), 'grad_textures': stats(a[2], b[2])}


def main():
    out = 'gpurun_out/parity_r2.json'
    if '--out' in sys.argv:
        out = sys.argv[sys.argv.index('--out') + 1]
    quick = '--quick' in sys.argv
    only = sys.argv[sys.argv.index('--only') + 1] if '--only' in sys.argv else None
    import gendr_b200 as gd
    dev = torch.device('cuda:0')
    ref = load_reference()
    assert ref is not None, 'baseline/_ref not staged'
    cases = []
    fv, ft, cfg = scenes.config_c2(batch=4 if quick else 16)
    cases.append(('C2 icosphere 1280 faces 256x256 logistic+probabilistic B=%d' % fv.shape[0], fv, ft, dict(double_side=False, **cfg)))
    fv, ft, cfg = scenes.config_c3(batch=2 if quick else 8)
    cases.append(('C3 grid sphere 8192 faces 256x256 gaussian+einstein B=%d' % fv.shape[0], fv, ft, dict(double_side=False, **cfg)))
    fv4, ft4, cfg4 = scenes.config_c4(batch=1 if quick else 2)
    cases.append(('C4 grid sphere 8192 faces 256x256 cauchy+yager(2) B=%d' % fv4.shape[0], fv4, ft4, dict(double_side=False, **cfg4)))
    fv1, ft1, _ = scenes.config_c3(batch=1)
    for dist, dkw, tcn, p in (('uniform', {}, 'probabilistic', None), ('gumbel_min', {}, 'probabilistic', None),
                              ('gamma_rev', dict(dist_shape=2.0), 'probabilistic', None), ('levy_rev', {}, 'probabilistic', None),
                              ('wigner_semicircle', {}, 'max', None), ('logistic', {}, 'dombi', 2.0)):
        cases.append(('C5 8192 faces 256x256 %s+%s B=1' % (dist, tcn), fv1, ft1,
                      dict(image_size=256, dist_func=dist, aggr_alpha_func=tcn, aggr_alpha_t_conorm_p=p, double_side=False, **dkw)))
    result = {'criterion': '|d| <= 1e-4*|ref| + atol; atol = 1e-5 (RGBA), 1e-4*max|ref| (gradients)', 'gpu': torch.cuda.get_device_name(0),
              'cases': {}}
    for name, fv, ft, kw in cases:
        if only and not name.startswith(only):
            continue
        t0 = time.time()
        fv, ft = scenes.with_sentinel(fv, ft)
        B, S = fv.shape[0], kw['image_size']
        g = torch.randn(B, 4, S, S, generator=torch.Generator().manual_seed(2)).to(dev)

        def run_ref32():
            a, b = fv.to(dev).requires_grad_(True), ft.to(dev).requires_grad_(True)
            img = reference_render(ref, a, b, **kw)
            img.backward(g)
            return img.detach(), a.grad.detach(), b.grad.detach()

        def run_ours():
            a, b = fv.to(dev).requires_grad_(True), ft.to(dev).requires_grad_(True)
            img = gd.functional.render(a, b, **kw)
            img.backward(g)
            return img.detach(), a.grad.detach(), b.grad.detach()
        r1, r2 = run_ref32(), run_ref32()
        r64 = reference_render_raw(ref, fv.to(dev), ft.to(dev), g, torch.float64, **kw)
        o1, o2 = run_ours(), run_ours()
        torch.cuda.synchronize()
        result['cases'][name] = {
            'reference_vs_reference_rerun': triple(r2, r1),
            'ours_vs_ours_rerun': triple(o2, o1),
            'reference_fp32_vs_reference_double': triple(r1, r64),
            'ours_vs_reference_double': triple(o1, r64),
            'ours_vs_reference_fp32': triple(o1, r1),
            'rgba_bit_identical_to_reference_fp32': bool(torch.equal(o1[0], r1[0])),
            'seconds': round(time.time() - t0, 2)}
        print(name, json.dumps({k: (v['rgba']['max_abs'], v['grad_faces']['max_over_max_ref']) for k, v in result['cases'][name].items()
                                if isinstance(v, dict)}), flush=True)
    os.makedirs(os.path.dirname(out) or '.', exist_ok=True)
    json.dump(result, open(out, 'w'), indent=1)
    print('wrote', out)


if __name__ == '__main__':
    main()
