"""Pure-PyTorch helpers on either side of the rasterizer: vertex gather, camera transforms, normals, lighting.
Host-side mirror of the reference helpers (O(B*V) glue, out of the hot path):
gendr/functional/{face_vertices,look_at,look,get_points_from_angles,vertex_normals,lighting}.py and
perspective/orthogonal from gendr/transform.py:14-44.  Run on CPU or CUDA tensors alike.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def _vec(v, device, dtype=torch.float32):
    """list / tuple / ndarray / tensor -> float tensor on `device`."""
    if torch.is_tensor(v):
        return v.to(device)
    if isinstance(v, np.ndarray):
        return torch.from_numpy(v).to(device)
    return torch.tensor(v, dtype=dtype, device=device)


def _batch_offsets(faces, num_vertices):
    bs = faces.shape[0]
    return faces + (torch.arange(bs, dtype=faces.dtype, device=faces.device) * num_vertices)[:, None, None]


def face_vertices(vertices, faces):
    """vertices [B,V,3], faces [B,F,3] (int) -> [B,F,3,3]   (functional/face_vertices.py:9-27)."""
    assert vertices.ndimension() == 3 and faces.ndimension() == 3
    assert vertices.shape[0] == faces.shape[0] and vertices.shape[2] == 3 and faces.shape[2] == 3
    bs, nv = vertices.shape[:2]
    return vertices.reshape(bs * nv, 3)[_batch_offsets(faces, nv).long()]


def vertex_normals(vertices, faces):
    """Area-weighted vertex normals [B,V,3]   (functional/vertex_normals.py:10-46)."""
    assert vertices.ndimension() == 3 and faces.ndimension() == 3
    assert vertices.shape[0] == faces.shape[0] and vertices.shape[2] == 3 and faces.shape[2] == 3
    bs, nv = vertices.shape[:2]
    idx = _batch_offsets(faces, nv).reshape(-1, 3).long()
    tri = vertices.reshape(bs * nv, 3)[idx]                       # [B*F, 3, 3]
    normals = torch.zeros(bs * nv, 3, dtype=vertices.dtype, device=vertices.device)
    for k in (1, 2, 0):                                            # same accumulation order as the reference
        a, b = tri[:, (k + 1) % 3] - tri[:, k], tri[:, (k + 2) % 3] - tri[:, k]
        normals.index_add_(0, idx[:, k], torch.cross(a, b, dim=1))
    return F.normalize(normals, eps=1e-6, dim=1).reshape(bs, nv, 3)


def get_points_from_angles(distance, elevation, azimuth, degrees=True):
    """Spherical -> cartesian camera position   (functional/get_points_from_angles.py:12-29)."""
    if isinstance(distance, (float, int)):
        if degrees:
            elevation, azimuth = math.radians(elevation), math.radians(azimuth)
        return (distance * math.cos(elevation) * math.sin(azimuth),
                distance * math.sin(elevation),
                -distance * math.cos(elevation) * math.cos(azimuth))
    if degrees:
        elevation, azimuth = math.pi / 180. * elevation, math.pi / 180. * azimuth
    return torch.stack([distance * torch.cos(elevation) * torch.sin(azimuth),
                        distance * torch.sin(elevation),
                        -distance * torch.cos(elevation) * torch.cos(azimuth)]).transpose(1, 0)


def _camera_rotation(z_axis, up):
    x_axis = F.normalize(torch.cross(up, z_axis, dim=1), eps=1e-5)
    y_axis = F.normalize(torch.cross(z_axis, x_axis, dim=1), eps=1e-5)
    return torch.stack((x_axis, y_axis, z_axis), dim=1)            # [B,3,3], rows = camera axes


def look_at(vertices, eye, at=[0, 0, 0], up=[0, 1, 0], only_rotate=False):
    """World -> camera coordinates for a camera at `eye` looking at `at`   (functional/look_at.py:11-68)."""
    if vertices.ndimension() != 3:
        raise ValueError('vertices Tensor should have 3 dimensions')
    dev, bs = vertices.device, vertices.shape[0]
    eye, at, up = (_vec(v, dev) for v in (eye, at, up))
    eye, at, up = (v[None, :].repeat(bs, 1) if v.ndimension() == 1 else v for v in (eye, at, up))
    r = _camera_rotation(F.normalize(at - eye, eps=1e-5), up)
    if not only_rotate:
        vertices = vertices - (eye[:, None, :] if vertices.shape != eye.shape else eye)
    return torch.matmul(vertices, r.transpose(1, 2))


def look(vertices, eye, direction=[0, 1, 0], up=[0, 1, 0]):
    """Camera at `eye` looking along `direction`   (functional/look.py:11-56; the reference crashes on its own
    default up=None, SURVEY Q-list -- here `up` defaults to +y)."""
    if vertices.ndimension() != 3:
        raise ValueError('vertices Tensor should have 3 dimensions')
    dev = vertices.device
    eye, direction, up = (_vec(v, dev) for v in (eye, direction, up))
    eye, direction, up = (v[None, :] if v.ndimension() == 1 else v for v in (eye, direction, up))
    z_axis = F.normalize(direction, eps=1e-5)
    r = _camera_rotation(z_axis, up.expand_as(z_axis))
    vertices = vertices - (eye[:, None, :] if vertices.shape != eye.shape else eye)
    return torch.matmul(vertices, r.transpose(1, 2))


def perspective(vertices, angle=30.):
    """x,y / z / tan(angle)   (transform.py:14-29)."""
    if vertices.ndimension() != 3:
        raise ValueError('vertices Tensor should have 3 dimensions')
    width = torch.tan(torch.tensor(angle / 180 * math.pi, dtype=torch.float32, device=vertices.device)[None])[:, None]
    z = vertices[:, :, 2]
    return torch.stack((vertices[:, :, 0] / z / width, vertices[:, :, 1] / z / width, z), dim=2)


def orthogonal(vertices, scale=1.):
    """x,y * scale   (transform.py:32-44)."""
    if vertices.ndimension() != 3:
        raise ValueError('vertices Tensor should have 3 dimensions')
    return torch.stack((vertices[:, :, 0] * scale, vertices[:, :, 1] * scale, vertices[:, :, 2]), dim=2)


def ambient_lighting(light, light_intensity=0.5, light_color=(1, 1, 1)):
    """light [B,N,3] += intensity * color   (functional/lighting.py:12-23; in place, like the reference)."""
    color = _vec(light_color, light.device).float()
    if color.ndimension() == 1:
        color = color[None, :]
    light += light_intensity * color[:, None, :]
    return light


def directional_lighting(light, normals, light_intensity=0.5, light_color=(1, 1, 1), light_direction=(0, 1, 0)):
    """light += intensity * color * relu(<n, d>)   (functional/lighting.py:26-48)."""
    color = _vec(light_color, light.device).float()
    direction = _vec(light_direction, light.device).float()
    if color.ndimension() == 1:
        color = color[None, :]
    if direction.ndimension() == 1:
        direction = direction[None, :]
    cosine = F.relu(torch.sum(normals * direction, dim=2))
    light += light_intensity * (color[:, None, :] * cosine[:, :, None])
    return light


def load_obj(filename_obj, normalization=False):
    """Minimal OBJ reader: `v` and `f` records only -> (vertices [V,3] float32, faces [F,3] int32), polygons fan-
    triangulated; optional unit-cube normalisation as functional/load_obj.py:146-151.  (Texture loading is an asset
    pipeline outside the hot path and is not mirrored.)"""
    verts, faces = [], []
    with open(filename_obj) as fh:
        for line in fh:
            parts = line.split()
            if not parts:
                continue
            if parts[0] == 'v':
                verts.append([float(v) for v in parts[1:4]])
            elif parts[0] == 'f':
                ids = [int(p.split('/')[0]) for p in parts[1:]]
                for k in range(1, len(ids) - 1):
                    faces.append([ids[0], ids[k], ids[k + 1]])
    vertices = torch.tensor(verts, dtype=torch.float32)
    faces = torch.tensor(faces, dtype=torch.int32) - 1
    if normalization:
        vertices = vertices - vertices.min(0)[0][None, :]
        vertices = vertices / torch.abs(vertices).max()
        vertices = vertices * 2
        vertices = vertices - vertices.max(0)[0][None, :] / 2
    return vertices, faces


def voxelization(faces, size, normalize=False):
    """faces [B,F,3,3] in unit-cube coordinates -> int32 occupancy grid [B,size,size,size] (surface + inside = 1).
    Drop-in for gendr.functional.voxelization (functional/voxelization.py:45-62); runs as two CUDA launches
    (gendr_voxelize) instead of the reference's 4 kernels + host loop, bit-identical result.  `normalize=True` means, as in
    the reference (:48-49), that the faces are already in voxel units."""
    from .. import _lib
    if not faces.is_cuda:
        raise TypeError('voxelization only supports CUDA Tensors.')
    f = faces.detach().to(torch.float32).contiguous()
    if normalize:
        f = f / size          # the kernels scale by `size` themselves (exact for power-of-two sizes only; the reference path
                              # with normalize=True is dead code: `pass`)
    B, F = int(f.shape[0]), int(f.shape[1])
    lib = _lib.load()
    voxels = torch.empty((B, size, size, size), dtype=torch.int32, device=f.device)
    ws = torch.empty(lib.gendr_voxelize_workspace_bytes(B, int(size)), dtype=torch.uint8, device=f.device)
    with torch.cuda.device(f.device):
        _lib.check(lib.gendr_voxelize(f.data_ptr(), voxels.data_ptr(), B, F, int(size), ws.data_ptr(), ws.numel(),
                                      torch.cuda.current_stream(f.device).cuda_stream))
    return voxels
