"""Per-kernel device times of one fused scene step on the reference scripts' small configurations (run under
ncu --metrics gpu__time_duration.sum)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch, scenes
import gendr_b200 as gd
dev = torch.device('cuda:0')
verts, faces = scenes.icosphere(3)
for B, S, cfg in ((64, 64, dict(dist_func='uniform', dist_scale=10 ** -1.5, dist_eps=300., aggr_rgb_func='hard')),
                  (24, 64, dict(dist_func='logistic', dist_scale=1e-2, dist_eps=100.))):
    v = (verts * 0.5)[None].repeat(B, 1, 1).to(dev); f = faces[None].repeat(B, 1, 1).to(dev)
    eyes = scenes.orbit_eyes(B).to(dev); g = torch.randn(B, 4, S, S, device=dev)
    for it in range(3):
        a = v.clone().requires_grad_(True)
        cam = gd.LookAt(viewing_angle=15); cam.set_eyes(eyes)
        img = gd.GenDR(image_size=S, **cfg)(cam(gd.Lighting()(gd.Mesh(a, f))))
        img.backward(g)
    torch.cuda.synchronize()
