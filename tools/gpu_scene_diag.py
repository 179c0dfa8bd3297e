"""Diagnostic: fused camera/lighting kernels vs the torch glue on the GPU (vertex/texture deltas, image deltas)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch, scenes
import gendr_b200 as gd
from gendr_b200 import _lib, mesh as mesh_mod
from gendr_b200.functional import make_camera_params, make_light_params
dev = torch.device('cuda:0')
lib = _lib.load()
verts, faces = scenes.icosphere(2)
v = (verts * 0.5)[None].to(dev).contiguous(); f = faces[None].to(dev).int().contiguous()
cam_mod = gd.LookAt(viewing_angle=15)
mesh_mod.FUSE_SCENE = False
m = cam_mod(gd.Lighting()(gd.Mesh(v, f)))
mesh_mod.FUSE_SCENE = True
sv_t, lt_t = m.vertices.contiguous(), m.textures.contiguous()
B, V = v.shape[:2]; F = f.shape[1]
eyes = torch.tensor(cam_mod._eye, dtype=torch.float32, device=dev)
cam = make_camera_params(mode='look_at', viewing_angle=15.)
sv_k = torch.empty_like(v)
st = torch.cuda.current_stream().cuda_stream
_lib.check(lib.gendr_camera_forward(v.data_ptr(), eyes.data_ptr(), 0, sv_k.data_ptr(), B, V, cam, st))
tex = torch.ones(B, F, 1, 3, device=dev); lt_k = torch.empty_like(tex)
_lib.check(lib.gendr_lighting_forward(v.data_ptr(), f.data_ptr(), 0, tex.data_ptr(), lt_k.data_ptr(), B, V, F, 1, make_light_params(), st))
torch.cuda.synchronize()
d = (sv_k - sv_t).abs()
print('screen verts: max abs diff', d.max().item(), 'per coord', d.amax((0, 1)).tolist())
ulps = d.cpu().numpy() / np.spacing(np.abs(sv_t.cpu().numpy()))
print('  in ulps: max', ulps.max(), 'mean', ulps.mean(), ' nonzero frac', (ulps > 0).mean())
print('lit tex: max abs diff', (lt_k - lt_t).abs().max().item())
print('width torch', torch.tan(torch.tensor(15 / 180 * np.pi, dtype=torch.float32, device=dev)).item().hex() if False else float(torch.tan(torch.tensor(15 / 180 * np.pi, dtype=torch.float32, device=dev))))
kw = dict(image_size=64, dist_func='logistic', dist_scale=0.02, double_side=False)
img_t = gd.functional.render_indexed(sv_t, f, lt_t, **kw)
img_k = gd.functional.render_indexed(sv_k, f, lt_k, **kw)
img_m = gd.functional.render_indexed(sv_t, f, lt_k, **kw)
img_s = gd.functional.render_scene(v, f, tex, cam_mod._eye, camera=dict(mode='look_at', viewing_angle=15.), lighting={}, **kw)
for name, a in (('kernel verts+tex', img_k), ('torch verts + kernel tex', img_m), ('scene path', img_s)):
    dd = (a - img_t).abs(); tol = 1e-4 * img_t.abs() + 1e-5
    print(name, 'frac outside', (dd > tol).float().mean().item(), 'max', dd.max().item(), 'per channel max', dd.amax((0, 2, 3)).tolist())
print('scene == kernel-staged:', torch.equal(img_s, img_k))
