"""Voxelizer latency: gendr_b200.functional.voxelization vs the reference (CUDA kernels + Python host loop), B200.
Writes gpurun_out/voxel_timing.json."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch, scenes
import gendr_b200 as gd
from gendr_b200 import _lib
from ref_gpu import load_reference
dev = torch.device('cuda:0'); ref = load_reference()


def timeit(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3


out = {}
verts, faces = scenes.icosphere(3)
for B, size in ((64, 32), (16, 64)):
    v = (verts * 0.45)[None].repeat(B, 1, 1).to(dev); f = faces[None].repeat(B, 1, 1).to(dev)
    fv = gd.functional.face_vertices(v, f) * size / (size - 1) + 0.5
    l0 = _lib.load().gendr_launch_count()
    r = {'ours_ms': timeit(lambda: gd.functional.voxelization(fv, size))}
    r['ours_launches_per_call'] = (_lib.load().gendr_launch_count() - l0) / 33
    if ref is not None and hasattr(ref.functional, 'voxelization'):
        r['reference_ms'] = timeit(lambda: ref.functional.voxelization(fv, size, False), n=10)
        r['speedup'] = r['reference_ms'] / r['ours_ms']
        r['identical'] = bool(torch.equal(gd.functional.voxelization(fv, size), ref.functional.voxelization(fv, size, False)))
    out['icosphere 1280 faces, batch %d, %d^3' % (B, size)] = r
    print(B, size, r, flush=True)
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'voxel_timing.json'), 'w'), indent=1)
