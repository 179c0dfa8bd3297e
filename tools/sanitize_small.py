"""Small forward+backward over every kernel variant class, for compute-sanitizer runs (memcheck / racecheck / initcheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch, scenes
import gendr_b200 as gd
dev = torch.device('cuda:0')
fv, ft = scenes.soup(300, batch=2, seed=5, size=0.08)
g = torch.randn(2, 4, 40, 40)
for kw in (dict(dist_func='gaussian', aggr_alpha_func='einstein'), dict(dist_func='cauchy', aggr_alpha_func='yager', aggr_alpha_t_conorm_p=2.0),
           dict(dist_func='uniform', aggr_alpha_func='probabilistic', aggr_rgb_func='hard', dist_eps=2.0), dict(dist_func='hard', aggr_alpha_func='hard')):
    a, b = fv.to(dev).requires_grad_(True), ft.to(dev).requires_grad_(True)
    img = gd.functional.render(a, b, image_size=40, dist_scale=0.02, **kw)
    img.backward(g.to(dev))
    torch.cuda.synchronize()
    print(kw['dist_func'], float(img.sum()), float(a.grad.abs().sum()))
# a case with more faces than one wave (multi-wave path) and a ragged image size
fv, ft = scenes.soup(700, batch=1, seed=6, size=0.5)
a, b = fv.to(dev).requires_grad_(True), ft.to(dev).requires_grad_(True)
img = gd.functional.render(a, b, image_size=37, dist_func='logistic', dist_scale=0.05)
img.backward(torch.ones_like(img)); torch.cuda.synchronize()
print('multiwave', float(img.sum()))
# fused scene path: camera + lighting kernels, shared-mesh batch-summed gradient, gradient w.r.t. the eye (two launches)
verts, faces = scenes.icosphere(1)
v = (verts * 0.5).to(dev).requires_grad_(True)
eyes = scenes.orbit_eyes(3).to(dev).requires_grad_(True)
tex = torch.rand(3, faces.shape[0], 1, 3).to(dev).requires_grad_(True)
img = gd.functional.render_scene(v, faces.to(dev), tex, eyes, camera=dict(viewing_angle=15.), lighting={}, image_size=48, dist_func='logistic',
                                 dist_scale=0.03, anti_aliasing=True)
img.backward(torch.ones_like(img)); torch.cuda.synchronize()
print('scene', float(img.sum()), float(v.grad.abs().sum()), float(eyes.grad.abs().sum()))
# a grid of more than one wave of CTAs (5 x 256 tiles): tile counters, the counting sort and the sorted CTA schedule; vertex-texture lighting
fv, ft = scenes.soup(60, batch=5, seed=7, size=0.3)
a, b = fv.to(dev).requires_grad_(True), ft.to(dev).requires_grad_(True)
img = gd.functional.render(a, b, image_size=256, dist_func='gaussian', aggr_alpha_func='einstein', dist_scale=0.01, dist_eps=30.)
img.backward(torch.ones_like(img)); torch.cuda.synchronize()
print('sorted schedule', float(img.sum()), float(a.grad.abs().sum()))
vt = torch.rand(3, verts.shape[0], 3).to(dev).requires_grad_(True)
vv = (verts * 0.5)[None].repeat(3, 1, 1).to(dev).requires_grad_(True)
lit = gd.functional.vertex_lighting(vv, faces.to(dev), vt, direction=(0.3, 0.8, -0.5))
lit.backward(torch.ones_like(lit)); torch.cuda.synchronize()
print('vertex lighting', float(lit.sum()), float(vv.grad.abs().sum()))
