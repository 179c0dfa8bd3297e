"""Synthetic scenes shared by tests, bench.py and smoke(): the configurations C1..C5 of SURVEY.md section 8(d),
built with gendr_b200's own host-side API (Mesh -> Lighting -> LookAt) exactly like the reference's scripts build
their inputs (experiments/opt_shape.py:257-259).  Everything runs on CPU tensors; callers move the results.
"""
import math

import numpy as np
import torch

import gendr_b200 as gd


def one_triangle():
    """The reference's 1-triangle scene (animations/triangles_dist.py:24-46)."""
    verts = torch.tensor([[-0.25 / 1.5, -.2165065 / 1.5, 0.], [0.0, 0.2165065 / 1.5, 0.], [0.25 / 1.5, -.2165065 / 1.5, 0.]],
                         dtype=torch.float32)
    faces = torch.tensor([[1, 0, 2]], dtype=torch.int32)
    return verts, faces


def icosphere(subdivisions=3, radius=1.0):
    """Unit icosphere: 3 subdivisions -> 642 vertices, 1280 faces (same size as experiments/data/sphere_642.obj)."""
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7),
         (9, 8, 1)]
    v = [np.array(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(subdivisions):
        cache, nf = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[key] = len(v) - 1
            return cache[key]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return torch.tensor(np.array(v) * radius, dtype=torch.float32), torch.tensor(f, dtype=torch.int32)


def grid_sphere(n=64, seed=0, jitter=0.15, radius=0.5):
    """Lat-long grid sphere with (n+1)^2 vertices and 2 n^2 triangles, per-vertex radius 0.5*(1 + jitter*U(-1,1)) --
    the "random-vertex mesh" of config C3 (n = 64 -> 8192 faces), SURVEY 8(d)."""
    g = torch.Generator().manual_seed(seed)
    lat = torch.linspace(0, math.pi, n + 1)
    lon = torch.linspace(0, 2 * math.pi, n + 1)
    la, lo = torch.meshgrid(lat, lon, indexing='ij')
    r = radius * (1 + jitter * (2 * torch.rand(n + 1, n + 1, generator=g) - 1))
    verts = torch.stack((r * torch.sin(la) * torch.cos(lo), r * torch.cos(la), r * torch.sin(la) * torch.sin(lo)), dim=-1).reshape(-1, 3)
    idx = torch.arange((n + 1) * (n + 1)).reshape(n + 1, n + 1)
    a, b, c, d = idx[:-1, :-1], idx[:-1, 1:], idx[1:, :-1], idx[1:, 1:]
    faces = torch.cat((torch.stack((a, c, b), -1).reshape(-1, 3), torch.stack((b, c, d), -1).reshape(-1, 3)), 0)
    return verts.float(), faces.int()


def render_inputs(verts, faces, eyes=None, viewing_angle=15, batch=1, textures=None):
    """(verts [V,3], faces [F,3]) -> face_vertices [B,F,3,3], face_textures [B,F,1,3] in screen space: Lighting()
    then LookAt(viewing_angle) with the given eyes ([B,3] tensor or a single 3-tuple)."""
    v = verts[None].repeat(batch, 1, 1)
    f = faces[None].repeat(batch, 1, 1)
    mesh = gd.Mesh(v, f, textures)
    mesh = gd.Lighting()(mesh)
    cam = gd.LookAt(viewing_angle=viewing_angle)
    if eyes is not None:
        cam.set_eyes(eyes)
    mesh = cam(mesh)
    return mesh.face_vertices.contiguous(), mesh.face_textures.contiguous()


def config_c1():
    """C1: 1 triangle, 32x32, uniform + probabilistic, batch 1."""
    verts, faces = one_triangle()
    eye = gd.functional.get_points_from_angles(2., 0, 0)
    fv, ft = render_inputs(verts, faces, eyes=eye, batch=1)
    return fv, ft, dict(image_size=32, dist_func='uniform', aggr_alpha_func='probabilistic')


def orbit_eyes(batch, distance=2.732, elevation=30., step=None):
    az = torch.arange(batch, dtype=torch.float32) * (step if step is not None else 360. / batch)
    return gd.functional.get_points_from_angles(torch.full((batch,), distance), torch.full((batch,), elevation), az)


def config_c2(batch=16, image_size=256):
    """C2: icosphere 1280 faces x0.5, 256x256, logistic + probabilistic, 16 views at -15 deg steps."""
    verts, faces = icosphere(3)
    fv, ft = render_inputs(verts * 0.5, faces, eyes=orbit_eyes(batch, step=-15.), batch=batch)
    return fv, ft, dict(image_size=image_size, dist_func='logistic', aggr_alpha_func='probabilistic')


def config_c3(batch=64, image_size=256, n=64):
    """C3 (headline): jittered grid sphere, 8192 faces, 256x256, gaussian + einstein, 64 views around the object."""
    verts, faces = grid_sphere(n, seed=0)
    fv, ft = render_inputs(verts, faces, eyes=orbit_eyes(batch), batch=batch)
    return fv, ft, dict(image_size=image_size, dist_func='gaussian', aggr_alpha_func='einstein')


def config_c4(batch=64, image_size=256, n=64):
    """C4 (per-GPU slice): C3 mesh, cauchy + yager(p=2) -- dense, nothing can be culled."""
    fv, ft, _ = config_c3(batch, image_size, n)
    return fv, ft, dict(image_size=image_size, dist_func='cauchy', aggr_alpha_func='yager', aggr_alpha_t_conorm_p=2.0)


def soup(num_faces, batch=1, seed=1, size=0.03, sentinel=True):
    """Screen-space triangle soup: centres U(-0.9,0.9)^2, z in U(2,4), vertex offsets size*U(-1,1)^2, random
    per-face colours.  With sentinel=True a last face far off screen is appended so that the reference's
    next-face texel read of the last real face is defined."""
    g = torch.Generator().manual_seed(seed)
    c = (torch.rand(batch, num_faces, 1, 2, generator=g) * 2 - 1) * 0.9
    off = (torch.rand(batch, num_faces, 3, 2, generator=g) * 2 - 1) * size
    z = torch.rand(batch, num_faces, 3, 1, generator=g) * 2 + 2
    fv = torch.cat((c + off, z), dim=-1)
    tex = torch.rand(batch, num_faces, 1, 3, generator=g)
    if sentinel:
        far = torch.tensor([[1e6, 1e6, 3.], [1e6 + 1, 1e6, 3.], [1e6, 1e6 + 1, 3.]])[None, None].repeat(batch, 1, 1, 1)
        fv = torch.cat((fv, far), dim=1)
        tex = torch.cat((tex, torch.full((batch, 1, 1, 3), 0.25)), dim=1)
    return fv.float().contiguous(), tex.float().contiguous()


def with_sentinel(fv, ft):
    """Append the culled-everywhere sentinel face to arbitrary inputs (see soup())."""
    B = fv.shape[0]
    far = torch.tensor([[1e6, 1e6, 3.], [1e6 + 1, 1e6, 3.], [1e6, 1e6 + 1, 3.]], dtype=fv.dtype)[None, None].repeat(B, 1, 1, 1)
    return torch.cat((fv, far), 1).contiguous(), torch.cat((ft, torch.full((B, 1) + tuple(ft.shape[2:]), 0.25, dtype=ft.dtype)), 1).contiguous()


# parameter sweep of config C5: every distribution x every t-conorm with valid shape parameters
DIST_SWEEP = [('hard', {}), ('uniform', {}), ('cubic_hermite', {}), ('wigner_semicircle', {}), ('gaussian', {}),
              ('laplace', {}), ('logistic', {}), ('gudermannian', {}), ('cauchy', {}), ('reciprocal', {}),
              ('gumbel_max', {}), ('gumbel_min', {}), ('exponential', {}), ('exponential_rev', {}),
              ('gamma', dict(dist_shape=2.0)), ('gamma_rev', dict(dist_shape=2.0)), ('levy', {}), ('levy_rev', {})]
TCN_SWEEP = [('hard', None), ('max', None), ('probabilistic', None), ('einstein', None), ('hamacher', 0.5), ('frank', 2.0),
             ('yager', 2.0), ('aczel_alsina', 2.0), ('dombi', 2.0), ('schweizer_sklar', -2.0)]


def voxel_cases():
    """Inputs for the voxelizer tests, as gendr.Mesh.voxelize passes them to
    gendr.functional.voxelization: face_vertices * vs / (vs - 1) + 0.5 (gendr/mesh.py:124-126).
      sphere32   jittered icosphere (320 faces), batch 3, 32^3 -- closed surfaces: the interior must fill
      sphere40   the same mesh at 40^3 (two mask words per row, size not a multiple of 32)
      soup16     open triangle soup partly outside the unit cube, 16^3 (nothing encloses anything; clipping at the borders)
      flat24     axis-aligned and degenerate (zero-area) faces on exact lattice planes, 24^3 (det == 0 / t == 0 edge cases)"""
    import numpy as np
    verts, faces = icosphere(2)
    g = torch.Generator().manual_seed(0)
    B = 3
    v = (verts * 0.4)[None].repeat(B, 1, 1) * (1 + 0.2 * torch.rand(B, verts.shape[0], 1, generator=g))
    fv = gd.functional.face_vertices(v, faces[None].repeat(B, 1, 1))
    cases = {}
    for name, vs in (('sphere32', 32), ('sphere40', 40)):
        cases[name] = ((fv * vs / (vs - 1) + 0.5).numpy().astype(np.float32), vs)
    soup = torch.rand(2, 60, 1, 3, generator=g) * 1.2 - 0.1 + (torch.rand(2, 60, 3, 3, generator=g) - 0.5) * 0.5
    cases['soup16'] = (soup.numpy().astype(np.float32), 16)
    flat = torch.tensor([[[0.25, 0.25, 0.5], [0.75, 0.25, 0.5], [0.25, 0.75, 0.5]],       # in the plane c2 = 12 (exact lattice plane)
                         [[0.5, 0.125, 0.125], [0.5, 0.875, 0.125], [0.5, 0.5, 0.875]],     # in the plane c0 = 12
                         [[0.1, 0.1, 0.1], [0.1, 0.1, 0.1], [0.3, 0.3, 0.3]],               # degenerate: two equal vertices
                         [[0.2, 0.6, 0.7], [0.4, 0.6, 0.7], [0.8, 0.6, 0.7]]])[None]        # degenerate: collinear
    cases['flat24'] = (flat.numpy().astype(np.float32), 24)
    return cases
