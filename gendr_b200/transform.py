"""Camera transforms as nn.Modules mapping Mesh -> Mesh (API mirror of gendr/transform.py:47-168)."""
import copy
import math

import numpy as np
import torch
import torch.nn as nn

from . import functional
from . import mesh as _mesh
from .functional import orthogonal, perspective  # noqa: F401  (re-exported like the reference module)
from .mesh import Mesh


class Transform(nn.Module):
    def transform(self, vertices):
        raise NotImplementedError()

    def forward(self, mesh):
        return Mesh(self.transform(mesh.vertices), mesh.faces, mesh.textures, mesh.texture_res, mesh.texture_type)


class _EyeCamera(Transform):
    def __init__(self, perspective=True, viewing_angle=30, viewing_scale=1.0, eye=None):
        super().__init__()
        self.perspective, self.viewing_angle, self.viewing_scale = perspective, viewing_angle, viewing_scale
        self._eye = eye if eye is not None else [0, 0, -(1. / math.tan(math.radians(viewing_angle)) + 1)]

    def set_eyes(self, eyes):
        self._eye = eyes

    @property
    def eyes(self):
        return self._eye

    def _project(self, vertices):
        if self.perspective:
            return functional.perspective(vertices, angle=self.viewing_angle)
        return functional.orthogonal(vertices, scale=self.viewing_scale)

    def _fusable_eye(self, mesh):
        """The eye as something the fused camera kernel takes ([3] list/tuple or a [B,3] / [3] tensor without grad)."""
        eye = self._eye
        if isinstance(eye, np.ndarray):
            eye = torch.from_numpy(eye)
        if torch.is_tensor(eye):
            if eye.ndimension() not in (1, 2) or eye.shape[-1] != 3 or not eye.is_floating_point():
                return None                                  # anything unusual: torch path
            if eye.ndimension() == 2 and eye.shape[0] != mesh.batch_size:
                return None
            return eye                                       # incl. eyes that require a gradient (experiments/opt_camera.py:236)
        if isinstance(eye, (list, tuple)) and len(eye) == 3 and all(isinstance(x, (int, float)) for x in eye):
            return eye
        return None

    def forward(self, mesh):
        if (_mesh.FUSE_SCENE and mesh._pending_camera is None and mesh._vertices.is_cuda and mesh.texture_type == 'surface'
                and self._fusable_eye(mesh) is not None and self._fused_camera() is not None):
            # deferred: GenDR.forward runs the camera kernel (or .vertices materialises it with transform()).  The reference
            # transforms eagerly, so the camera is SNAPSHOT here: later in-place changes of the eye tensor / attributes of this
            # module must not reach the deferred step.
            cam = copy.copy(self)
            eye = self._eye
            cam._eye = (eye.clone() if eye.requires_grad else eye.detach().clone()) if torch.is_tensor(eye) else copy.deepcopy(eye)
            if hasattr(cam, 'camera_direction'):
                cam.camera_direction = copy.deepcopy(self.camera_direction)
            return Mesh(mesh._vertices, mesh.faces, mesh._textures, mesh.texture_res, mesh.texture_type,
                        _pending_light=mesh._pending_light, _pending_camera=cam)
        return super().forward(mesh)


class LookAt(_EyeCamera):
    """transform.py:109-138"""
    def set_eyes_from_angles(self, distances, elevations, azimuths):
        self._eye = functional.get_points_from_angles(distances, elevations, azimuths)

    def transform(self, vertices):
        return self._project(functional.look_at(vertices, self._eye))

    def _fused_camera(self):
        return dict(mode='look_at', perspective=self.perspective, viewing_angle=self.viewing_angle, viewing_scale=self.viewing_scale)


class Look(_EyeCamera):
    """transform.py:141-168"""
    def __init__(self, camera_direction=[0, 0, 1], perspective=True, viewing_angle=30, viewing_scale=1.0, eye=None):
        super().__init__(perspective, viewing_angle, viewing_scale, eye)
        self.camera_direction = camera_direction

    def transform(self, vertices):
        return self._project(functional.look(vertices, self._eye, self.camera_direction))

    def _fused_camera(self):
        d = self.camera_direction
        if not (isinstance(d, (list, tuple)) and len(d) == 3 and all(isinstance(x, (int, float)) for x in d)):
            return None
        return dict(mode='look', perspective=self.perspective, viewing_angle=self.viewing_angle, viewing_scale=self.viewing_scale,
                    direction=tuple(d))


class Projection(Transform):
    """3x4 projection matrix + OpenCV-style distortion (transform.py:64-106)."""
    def __init__(self, P, dist_coeffs=None, orig_size=512):
        super().__init__()
        if isinstance(P, np.ndarray):
            P = torch.from_numpy(P)
        if P is None or P.ndimension() != 3 or P.shape[1] != 3 or P.shape[2] != 4:
            raise ValueError('You need to provide a valid (batch_size)x3x4 projection matrix')
        self.P, self.orig_size = P, orig_size
        self.dist_coeffs = dist_coeffs if dist_coeffs is not None else torch.zeros(P.shape[0], 5, dtype=P.dtype, device=P.device)

    def transform(self, vertices):
        P, k = self.P.to(vertices.device), self.dist_coeffs.to(vertices.device)
        v = torch.bmm(torch.cat([vertices, torch.ones_like(vertices[:, :, :1])], dim=-1), P.transpose(2, 1))
        x, y, z = v[:, :, 0] / (v[:, :, 2] + 1e-5), v[:, :, 1] / (v[:, :, 2] + 1e-5), v[:, :, 2]
        k1, k2, p1, p2, k3 = (k[:, None, i] for i in range(5))
        r2 = x ** 2 + y ** 2
        radial = 1 + k1 * r2 + k2 * r2 ** 2 + k3 * r2 ** 3
        xd = x * radial + 2 * p1 * x * y + p2 * (r2 + 2 * x ** 2)
        yd = y * radial + p1 * (r2 + 2 * y ** 2) + 2 * p2 * x * y
        half = self.orig_size / 2.
        return torch.stack([2 * (xd - half) / self.orig_size, 2 * (yd - half) / self.orig_size, z], dim=-1)
