import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch, scenes
import gendr_b200 as gd
dev = torch.device('cuda:0')
fv, ft = scenes.soup(300, batch=2, seed=5, size=0.08)
g = torch.randn(2, 4, 40, 40, generator=torch.Generator().manual_seed(1))
for kw in (dict(dist_func='uniform', aggr_alpha_func='probabilistic', aggr_rgb_func='hard', dist_eps=2.0), dict(dist_func='uniform', aggr_alpha_func='probabilistic', aggr_rgb_func='hard'),
           dict(dist_func='uniform', aggr_alpha_func='probabilistic', dist_eps=2.0), dict(dist_func='gaussian', aggr_alpha_func='einstein', dist_eps=2.0)):
    outs = []
    for rep in range(3):
        a, b = fv.to(dev).requires_grad_(True), ft.to(dev).requires_grad_(True)
        img = gd.functional.render(a, b, image_size=40, dist_scale=0.02, **kw)
        img.backward(g.to(dev)); torch.cuda.synchronize()
        outs.append((img.detach().clone(), a.grad.clone(), b.grad.clone()))
    d_img = max(float((o[0] - outs[0][0]).abs().max()) for o in outs)
    d_g = max(float((o[1] - outs[0][1]).abs().max()) for o in outs)
    print(kw, 'img run-to-run', d_img, 'grad run-to-run', d_g, 'grad max', float(outs[0][1].abs().max()), 'abs sum', [float(o[1].abs().sum()) for o in outs])
