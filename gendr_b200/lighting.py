"""Ambient + directional lighting folded into the textures (API mirror of gendr/lighting.py:11-71)."""
import torch
import torch.nn as nn

from . import functional
from . import mesh as _mesh
from .mesh import Mesh


def _is_vec3(v):
    return isinstance(v, (list, tuple)) and len(v) == 3 and all(isinstance(x, (int, float)) for x in v)


class AmbientLighting(nn.Module):
    def __init__(self, light_intensity=0.5, light_color=(1, 1, 1)):
        super().__init__()
        self.light_intensity, self.light_color = light_intensity, light_color

    def forward(self, light):
        return functional.ambient_lighting(light, self.light_intensity, self.light_color)


class DirectionalLighting(nn.Module):
    def __init__(self, light_intensity=0.5, light_color=(1, 1, 1), light_direction=(0, 1, 0)):
        super().__init__()
        self.light_intensity, self.light_color, self.light_direction = light_intensity, light_color, light_direction

    def forward(self, light, normals):
        return functional.directional_lighting(light, normals, self.light_intensity, self.light_color, self.light_direction)


class _LightSnapshot(object):
    """What a deferred Lighting step needs later: the fused-kernel parameters (a plain dict, snapshot at call time) and the torch
    implementation for meshes that get materialised instead.  (A deepcopy of the nn.Module costs ~100 us per call.)"""
    __slots__ = ('params',)

    def __init__(self, params):
        self.params = params

    def fused_params(self):
        return self.params

    def lit_textures(self, mesh):
        p = self.params
        return Lighting(p['intensity_ambient'], list(p['color_ambient']), p['intensity_directional'], list(p['color_directional']),
                        list(p['direction'])).lit_textures(mesh)


class Lighting(nn.Module):
    def __init__(self, intensity_ambient=0.5, color_ambient=[1, 1, 1], intensity_directionals=0.5,
                 color_directionals=[1, 1, 1], directions=[0, 1, 0]):
        super().__init__()
        self.ambient = AmbientLighting(intensity_ambient, color_ambient)
        self.directionals = nn.ModuleList([DirectionalLighting(intensity_directionals, color_directionals, directions)])

    def lit_textures(self, mesh):
        """textures * (ambient + directional light), the torch implementation (gendr/lighting.py:48-64)."""
        if mesh.texture_type == 'surface':
            like, normals, expand = mesh.faces, mesh.surface_normals, lambda l: l[:, :, None, :]
        elif mesh.texture_type == 'vertex':
            like, normals, expand = mesh.vertices, mesh.vertex_normals, lambda l: l
        else:
            raise ValueError('texture type not applicable')
        light = self.ambient(torch.zeros(like.shape, dtype=torch.float32, device=mesh.device))
        for directional in self.directionals:
            light = directional(light, normals)
        return mesh.textures * expand(light)

    def fused_params(self):
        """dict for functional.make_light_params, or None when this configuration has no fused kernel (tensor-valued
        colours/directions, several directional lights)."""
        if len(self.directionals) != 1:
            return None
        d, a = self.directionals[0], self.ambient
        if not (_is_vec3(a.light_color) and _is_vec3(d.light_color) and _is_vec3(d.light_direction)):
            return None
        if not all(isinstance(x, (int, float)) for x in (a.light_intensity, d.light_intensity)):
            return None
        return dict(intensity_ambient=a.light_intensity, color_ambient=tuple(a.light_color),
                    intensity_directional=d.light_intensity, color_directional=tuple(d.light_color),
                    direction=tuple(d.light_direction))

    def forward(self, mesh):
        if (_mesh.FUSE_SCENE and mesh.texture_type == 'surface' and mesh._pending_light is None and mesh._pending_camera is None
                and mesh._vertices.is_cuda):
            params = self.fused_params()
            if params is not None:
                # deferred: GenDR.forward runs the lighting kernel (or .textures materialises it with lit_textures)
                return Mesh(mesh._vertices, mesh.faces, mesh._textures, mesh.texture_res, mesh.texture_type,
                            _pending_light=_LightSnapshot(params))
        if (_mesh.FUSE_SCENE and mesh.texture_type == 'vertex' and mesh._pending_light is None and mesh._pending_camera is None
                and mesh._vertices.is_cuda):
            params = self.fused_params()
            if params is not None:      # vertex normals + light + multiply as CUDA kernels (2 launches; ~25 torch kernels otherwise)
                lit = functional.vertex_lighting(mesh._vertices, mesh.faces, mesh._textures, **params)
                return Mesh(mesh._vertices, mesh.faces, lit, mesh.texture_res, mesh.texture_type)
        return Mesh(mesh.vertices, mesh.faces, self.lit_textures(mesh), mesh.texture_res, mesh.texture_type)
