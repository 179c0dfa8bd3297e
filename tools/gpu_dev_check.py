"""Development check run on the GPU box: geometry probe (bitwise vs oracle mode 1), small-scene parity vs the CPU
oracle, full-size parity vs the reference's own CUDA kernels (baseline/_ref), and first timings.
Usage: python tools/gpu_dev_check.py [--quick]"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch

import gendr_b200 as gd
import scenes
from gendr_b200 import _lib
from oracle.cpu_oracle import Oracle, make_params
from ref_gpu import load_reference, reference_render

dev = torch.device('cuda:0')
OUT = os.path.join(ROOT, 'gpurun_out'); os.makedirs(OUT, exist_ok=True)
report = {}


def stats(name, new, ref, atol, rtol=1e-4):
    new, ref = new.double(), ref.double()
    nanmask_ok = bool((torch.isnan(new) == torch.isnan(ref)).all())
    m = ~(torch.isnan(new) | torch.isnan(ref))
    d = (new[m] - ref[m]).abs()
    tol = rtol * ref[m].abs() + atol
    frac_bad = float((d > tol).double().mean()) if d.numel() else 0.0
    r = dict(max_abs=float(d.max()) if d.numel() else 0.0, ref_max=float(ref[m].abs().max()) if d.numel() else 0.0,
             frac_bad=frac_bad, n_bad=int((d > tol).sum()), nanmask_ok=nanmask_ok,
             p999=float(torch.quantile(d.flatten()[:4000000], 0.999)) if d.numel() else 0.0)
    print('   %-34s max_abs %.3e (ref_max %.3e) p99.9 %.2e bad %d (%.2e) nan_ok %s' % (name, r['max_abs'], r['ref_max'], r['p999'], r['n_bad'], frac_bad, nanmask_ok))
    return r


def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


# ---- 1. geometry probe: bitwise against the oracle's mode-1 (GPU contraction) arithmetic -------------------------
port = Oracle('port'); port.lib.gendr_oracle_set_mode(1)
rng = np.random.default_rng(0)
n = 200000
c = rng.uniform(-0.9, 0.9, (n, 1, 2)); off = rng.uniform(-1, 1, (n, 3, 2)) * rng.choice([0.3, 0.03, 0.003], (n, 1, 1))
# a third of the faces squashed into slivers
squash = rng.choice([1.0, 1e-2, 1e-4], (n, 1, 1)); off[:, :, 1:2] *= squash
faces = np.concatenate([c + off, rng.uniform(2, 4, (n, 3, 1))], axis=2).astype(np.float32).reshape(n, 9)
xy = np.where(rng.random((n, 1)) < 0.5, c[:, 0, :] + rng.uniform(-1, 1, (n, 2)) * 0.05, rng.uniform(-1, 1, (n, 2))).astype(np.float32)
d_faces, d_xy = torch.from_numpy(faces).to(dev), torch.from_numpy(xy).to(dev)
d_out = torch.empty(n, 10, device=dev)
lib = _lib.load()
_lib.check(lib.gendr_probe_pairs(d_faces.data_ptr(), d_xy.data_ptr(), d_out.data_ptr(), n, None))
got = d_out.cpu().numpy()
exp = np.zeros((n, 10), np.float32)
buf = np.zeros(10, np.float32)
port.lib.gendr_oracle_pair_geometry.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_void_p]
for i in range(n):
    port.lib.gendr_oracle_pair_geometry(faces[i].ctypes.data, float(xy[i, 0]), float(xy[i, 1]), buf.ctypes.data)
    exp[i] = buf
defined = exp[:, 9] == 1
same = (got[:, :9].view(np.uint32) == exp[:, :9].view(np.uint32)) | (np.isnan(got[:, :9]) & np.isnan(exp[:, :9]))
rows_ok = same.all(axis=1)
print('probe: %d pairs, %d defined, bitwise-identical rows: %d (%.4f%%), inside %d' % (n, defined.sum(), (rows_ok & defined).sum(), 100.0 * (rows_ok & defined).sum() / defined.sum(), (exp[:, 8] > 0).sum()))
report['probe'] = dict(n=n, defined=int(defined.sum()), identical=int((rows_ok & defined).sum()))
badrows = np.where(~rows_ok & defined)[0][:5]
for i in badrows: print('   bad', i, got[i], exp[i])

# ---- 2. small scenes vs CPU oracle (mode 1) ---------------------------------------------------------------------
def run_new(fv, ft, g, **kw):
    fv = fv.to(dev).requires_grad_(True); ft = ft.to(dev).requires_grad_(True)
    img = gd.functional.render(fv, ft, **kw)
    img.backward(g.to(dev))
    return img.detach(), fv.grad.detach(), ft.grad.detach()

small = []
fv, ft, cfg = scenes.config_c1(); small.append(('C1', fv, ft, dict(cfg, double_side=False)))
fv, ft = scenes.soup(200, batch=2, seed=5, size=0.08)
for dname, dkw in scenes.DIST_SWEEP[::1]:
    small.append(('soup-' + dname, fv, ft, dict(image_size=40, dist_func=dname, aggr_alpha_func='probabilistic', dist_scale=0.02, **dkw)))
for tname, tp in scenes.TCN_SWEEP:
    small.append(('soup-logistic-' + tname, fv, ft, dict(image_size=40, dist_func='logistic', aggr_alpha_func=tname, aggr_alpha_t_conorm_p=tp, dist_scale=0.02)))
small.append(('soup-hardrgb', fv, ft, dict(image_size=40, dist_func='gaussian', aggr_alpha_func='einstein', aggr_rgb_func='hard', dist_scale=0.02)))
small.append(('soup-squared', fv, ft, dict(image_size=40, dist_func='logistic', aggr_alpha_func='probabilistic', dist_squared=True, dist_scale=4e-4)))
fvi, fti, _ = scenes.config_c2(batch=2, image_size=64); fvi, fti = scenes.with_sentinel(fvi, fti)
small.append(('ico-logistic', fvi, fti, dict(image_size=64, dist_func='logistic', aggr_alpha_func='probabilistic', double_side=False)))
print('== small scenes vs CPU oracle (mode 1)')
worst = 0
for name, fv, ft, kw in small:
    p = make_params(**kw)
    fo = port.forward(fv.numpy(), ft.numpy(), p)
    g = torch.from_numpy(np.random.default_rng(2).standard_normal(fo['soft_colors'].shape).astype(np.float32))
    go = port.backward(fo, g.numpy(), p)
    img, gf, gt = run_new(fv, ft, g, **kw)
    print(name)
    r1 = stats('rgba', img.cpu(), torch.from_numpy(fo['soft_colors']), 1e-5)
    gref = torch.from_numpy(go[0]).reshape(gf.shape)
    r2 = stats('grad_faces', gf.cpu(), gref, 1e-4 * float(gref.abs().max()))
    gtr = torch.from_numpy(go[1]).reshape(gt.shape)
    r3 = stats('grad_textures', gt.cpu(), gtr, 1e-4 * float(gtr.abs().max()) + 1e-12)
    report['small/' + name] = dict(rgba=r1, gf=r2, gt=r3)

# ---- 3. full-size parity + timing vs the reference CUDA kernels -------------------------------------------------
ref = load_reference()
print('reference CUDA build available:', ref is not None)
if ref is not None and '--quick' not in sys.argv:
    big = []
    fv, ft, cfg = scenes.config_c1(); big.append(('C1', fv, ft, dict(cfg, double_side=False), 20))
    fv, ft, cfg = scenes.config_c2(batch=16); fv, ft = scenes.with_sentinel(fv, ft); big.append(('C2', fv, ft, dict(cfg, double_side=False), 3))
    fv, ft, cfg = scenes.config_c3(batch=8); fv, ft = scenes.with_sentinel(fv, ft); big.append(('C3/b8', fv, ft, dict(cfg, double_side=False), 2))
    fv, ft, cfg = scenes.config_c4(batch=2); fv, ft = scenes.with_sentinel(fv, ft); big.append(('C4/b2', fv, ft, dict(cfg, double_side=False), 1))
    for name, fv, ft, kw, nrep in big:
        print('==', name, kw)
        B, F = fv.shape[:2]; S = kw['image_size']
        g = torch.randn(B, 4, S, S, generator=torch.Generator().manual_seed(2))
        img, gf, gt = run_new(fv, ft, g, **kw)
        fr = fv.to(dev).requires_grad_(True); tr = ft.to(dev).requires_grad_(True)
        imr = reference_render(ref, fr, tr, **kw); imr.backward(g.to(dev))
        r1 = stats('rgba vs reference CUDA', img, imr.detach(), 1e-5)
        r1a = stats('alpha only', img[:, 3], imr.detach()[:, 3], 1e-5)
        r2 = stats('grad_faces vs reference CUDA', gf, fr.grad, 1e-4 * float(fr.grad.abs().max()))
        r3 = stats('grad_textures vs reference CUDA', gt, tr.grad, 1e-4 * float(tr.grad.abs().max()) + 1e-12)
        # reference run-to-run noise (atomic order)
        fr2 = fv.to(dev).requires_grad_(True); tr2 = ft.to(dev).requires_grad_(True)
        imr2 = reference_render(ref, fr2, tr2, **kw); imr2.backward(g.to(dev))
        r4 = stats('reference vs reference (noise)', fr2.grad, fr.grad, 1e-4 * float(fr.grad.abs().max()))
        gdev = g.to(dev)
        def new_step():
            a = fv.to(dev).requires_grad_(True); b = ft.to(dev).requires_grad_(True)
            gd.functional.render(a, b, **kw).backward(gdev)
        def ref_step():
            a = fv.to(dev).requires_grad_(True); b = ft.to(dev).requires_grad_(True)
            reference_render(ref, a, b, **kw).backward(gdev)
        t_new = timeit(new_step, n=max(3, nrep * 3))
        t_ref = timeit(ref_step, n=nrep, warm=1)
        pairs = B * S * S * F
        print('   time new %.3f ms  ref %.3f ms  speedup %.1fx   new %.1f Mpix*face/s' % (t_new[0], t_ref[0], t_ref[0] / t_new[0], pairs / t_new[0] / 1e3))
        report['big/' + name] = dict(rgba=r1, alpha=r1a, gf=r2, gt=r3, noise=r4, ms_new=t_new, ms_ref=t_ref, pairs=pairs)
json.dump(report, open(os.path.join(OUT, 'dev_check.json'), 'w'), indent=1)
print('launches', lib.gendr_launch_count())
