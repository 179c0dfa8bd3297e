"""One fused scene step at C3 size (through the module API) and one voxelizer call, for an ncu launch list of the kernels around
the rasterizer (camera, lighting, indexed prep, voxel surface / fill)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch, scenes
import gendr_b200 as gd
dev = torch.device('cuda:0')
verts, faces = scenes.grid_sphere(64)
B, S = 64, 256
v = verts[None].repeat(B, 1, 1).to(dev); f = faces[None].repeat(B, 1, 1).to(dev)
eyes = scenes.orbit_eyes(B).to(dev); g = torch.randn(B, 4, S, S, device=dev)
for it in range(2):
    a = v.clone().requires_grad_(True)
    cam = gd.LookAt(viewing_angle=15); cam.set_eyes(eyes)
    img = gd.GenDR(image_size=S, dist_func='gaussian', aggr_alpha_func='einstein')(cam(gd.Lighting()(gd.Mesh(a, f))))
    img.backward(g)
iv, if_ = scenes.icosphere(3)
mesh = gd.Mesh((iv * 0.45)[None].repeat(64, 1, 1).to(dev), if_[None].repeat(64, 1, 1).to(dev))
for it in range(2):
    vox = mesh.voxelize(32)
torch.cuda.synchronize()
print('filled fraction', float(vox.float().mean()))
