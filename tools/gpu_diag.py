"""Diagnostics for one config vs the reference CUDA kernels: where do mismatches sit?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch, scenes
import gendr_b200 as gd
from ref_gpu import load_reference, reference_render
dev = torch.device('cuda:0'); ref = load_reference()
fv, ft, _ = scenes.config_c3(batch=1, n=32); fv, ft = scenes.with_sentinel(fv, ft)
cfgs = [dict(image_size=128, dist_func='gamma_rev', dist_shape=2.0, aggr_alpha_func='probabilistic', aggr_rgb_func='hard', double_side=True),
        dict(image_size=128, dist_func='logistic', aggr_alpha_func='max', double_side=False),
        dict(image_size=128, dist_func='levy_rev', aggr_alpha_func='max', double_side=False)]
for kw in cfgs:
    S = kw['image_size']
    g = torch.randn(1, 4, S, S, generator=torch.Generator().manual_seed(2)).to(dev)
    a = fv.to(dev).requires_grad_(True); b = ft.to(dev).requires_grad_(True)
    img = gd.functional.render(a, b, **kw); img.backward(g)
    a2 = fv.to(dev).requires_grad_(True); b2 = ft.to(dev).requires_grad_(True)
    imr = reference_render(ref, a2, b2, **kw); imr.backward(g)
    d = (img - imr).abs()
    print(kw['dist_func'], kw['aggr_alpha_func'], 'per-channel max', [float(d[0, k].max()) for k in range(4)], 'n>1e-5', [int((d[0, k] > 1e-5).sum()) for k in range(4)])
    idx = (d[0].amax(0) > 1e-5).nonzero()[:6]
    for (y, x) in idx.tolist():
        print('   px', y, x, 'new', img[0, :, y, x].tolist(), 'ref', imr[0, :, y, x].tolist())
    dg = (a.grad - a2.grad).abs().view(-1, 9); scale = float(a2.grad.abs().max())
    badf = (dg.amax(1) > 1e-4 * scale).nonzero().flatten()
    print('   grad_faces: bad faces', badf.numel(), 'max', float(dg.max()), 'scale', scale)
    for f in badf[:4].tolist():
        print('     face', f, 'new', a.grad.view(-1, 9)[f].tolist(), '\n            ref', a2.grad.view(-1, 9)[f].tolist())
