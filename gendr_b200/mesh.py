"""`Mesh` container: vertices [B,V,3], faces [B,F,3] int, textures [B,F,R*R,3] (surface) or [B,V,3] (vertex).
API mirror of gendr.Mesh (/root/reference/gendr/mesh.py:14-126) minus OBJ texture I/O (asset pipeline, out of scope).

Deferred scene steps.  The reference's scripts always run
`mesh = lighting(mesh); mesh = transform(mesh); images = renderer(mesh)` (experiments/opt_shape.py:257-259).  On CUDA
meshes `Lighting` and `LookAt`/`Look` do not execute their ~30 small torch kernels right away: they return a Mesh that
remembers the step (`_pending_light`, `_pending_camera`), and `GenDR.forward` then runs lighting + camera + gather +
rasterizer as one fused CUDA path (functional.render_scene).  Anything else that looks at the mesh first -- `.vertices`,
`.textures`, `.face_vertices`, normals -- materialises the pending steps with the plain torch implementation, so the
observable semantics are unchanged.  `gendr_b200.mesh.FUSE_SCENE = False` switches the deferral off.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import functional


FUSE_SCENE = True


def _default_device():
    return torch.device('cuda') if torch.cuda.is_available() else torch.device('cpu')


class Mesh(object):
    def __init__(self, vertices, faces, textures=None, texture_res=1, texture_type='surface', _pending_light=None,
                 _pending_camera=None):
        self._pending_light, self._pending_camera = _pending_light, _pending_camera
        if isinstance(vertices, np.ndarray):
            vertices = torch.from_numpy(vertices).float().to(_default_device())
        if isinstance(faces, np.ndarray):
            faces = torch.from_numpy(faces).int().to(_default_device())
        self._vertices = vertices[None] if vertices.ndimension() == 2 else vertices
        self._faces = faces[None] if faces.ndimension() == 2 else faces
        self.device = self._vertices.device
        self.texture_type = texture_type
        self.batch_size, self.num_vertices = self._vertices.shape[:2]
        self.num_faces = self._faces.shape[1]

        if textures is None:
            if texture_type == 'surface':
                shape, self.texture_res = (self.batch_size, self.num_faces, texture_res ** 2, 3), texture_res
            elif texture_type == 'vertex':
                shape, self.texture_res = (self.batch_size, self.num_vertices, 3), 1
            else:
                raise ValueError('texture type not applicable')
            self._textures = torch.ones(*shape, dtype=torch.float32, device=self.device)
        else:
            if isinstance(textures, np.ndarray):
                textures = torch.from_numpy(textures).float().to(self.device)
            if textures.ndimension() == 3 and texture_type == 'surface':
                textures = textures[None]
            if textures.ndimension() == 2 and texture_type == 'vertex':
                textures = textures[None]
            self._textures = textures
            self.texture_res = int(np.sqrt(self._textures.shape[2]))

    @classmethod
    def from_obj(cls, filename_obj, normalization=False, load_texture=False, texture_res=1, texture_type='surface'):
        """Positional signature of gendr.Mesh.from_obj (gendr/mesh.py:62-81); OBJ/MTL texture loading is the asset pipeline
        (out of scope), so load_texture=True raises instead of silently ignoring the request."""
        if load_texture:
            raise NotImplementedError('gendr_b200.Mesh.from_obj: load_texture=True (OBJ/MTL texture loading) is not part of the hot path')
        vertices, faces = functional.load_obj(filename_obj, normalization=normalization)
        dev = _default_device()
        return cls(vertices.to(dev), faces.to(dev), None, texture_res, texture_type)

    def _materialize(self):
        """Run the deferred lighting / camera steps with their torch implementations (lighting first: it needs the
        world-space normals)."""
        light, camera = self._pending_light, self._pending_camera
        self._pending_light = self._pending_camera = None      # first: lit_textures() reads .vertices (world space) itself
        if light is not None:
            self._textures = light.lit_textures(self)
        if camera is not None:
            self._vertices = camera.transform(self._vertices)

    faces = property(lambda self: self._faces)

    @property
    def vertices(self):
        self._materialize()
        return self._vertices

    @property
    def textures(self):
        self._materialize()
        return self._textures

    @property
    def face_vertices(self):
        return functional.face_vertices(self.vertices, self.faces)

    @property
    def surface_normals(self):
        fv = self.face_vertices
        return F.normalize(torch.cross(fv[:, :, 2] - fv[:, :, 1], fv[:, :, 0] - fv[:, :, 1], dim=2), p=2, dim=2, eps=1e-6)

    @property
    def vertex_normals(self):
        return functional.vertex_normals(self.vertices, self.faces)

    @property
    def face_textures(self):
        if self.texture_type == 'surface':
            return self.textures
        if self.texture_type == 'vertex':
            return functional.face_vertices(self.textures, self.faces)
        raise ValueError('texture type not applicable')

    def voxelize(self, voxel_size=32):
        """gendr/mesh.py:124-126"""
        face_vertices_norm = self.face_vertices * voxel_size / (voxel_size - 1) + 0.5
        return functional.voxelization(face_vertices_norm, voxel_size, False)
