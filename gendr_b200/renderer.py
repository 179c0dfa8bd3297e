"""`GenDR` -- the nn.Module front door, constructor-compatible with gendr.GenDR
(/root/reference/gendr/renderer.py:12-125): 18 keyword arguments with the same names and defaults, every one a
plain mutable attribute read at call time (scripts change them between calls, e.g. experiments/opt_camera.py:236),
`forward(mesh)` and `forward_tensors(face_vertices, face_textures)`, 2x supersampling + avg_pool2d anti-aliasing.
"""
import torch.nn as nn
import torch.nn.functional as F

from . import functional

_RENDER_ARGS = ('background_color', 'dist_func', 'dist_scale', 'dist_squared', 'dist_shape', 'dist_shift', 'dist_eps',
                'aggr_alpha_func', 'aggr_alpha_t_conorm_p', 'aggr_rgb_func', 'aggr_rgb_eps', 'aggr_rgb_gamma',
                'near', 'far', 'double_side', 'texture_type')


class GenDR(nn.Module):
    def __init__(self, image_size=256, background_color=[0, 0, 0], anti_aliasing=False,
                 dist_func='uniform', dist_scale=1e-2, dist_squared=False, dist_shape=None, dist_shift=None,
                 dist_eps=1e4, aggr_alpha_func='probabilistic', aggr_alpha_t_conorm_p=None,
                 aggr_rgb_func='softmax', aggr_rgb_eps=1e-3, aggr_rgb_gamma=1e-3,
                 near=1, far=100, double_side=False, texture_type='surface'):
        super().__init__()
        if aggr_rgb_func not in ['hard', 'softmax']:
            raise ValueError('Aggregate function (RGB) currently only supports hard and softmax.')
        if texture_type not in ['surface', 'vertex']:
            raise ValueError('Texture type only support surface and vertex.')
        self.image_size = image_size
        self.anti_aliasing = anti_aliasing
        values = locals()
        for name in _RENDER_ARGS:
            setattr(self, name, values[name])

    def forward_tensors(self, face_vertices, face_textures):
        size = self.image_size * (2 if self.anti_aliasing else 1)
        images = functional.render(face_vertices=face_vertices, textures=face_textures, image_size=size,
                                   **{name: getattr(self, name) for name in _RENDER_ARGS})
        if self.anti_aliasing:
            images = F.avg_pool2d(images, kernel_size=2, stride=2)
        return images

    def forward(self, mesh):
        if mesh.texture_type == 'surface' and getattr(self, 'fused_gather', True):
            # indexed path: vertices[faces] gather and its scatter-add backward run inside the CUDA kernels
            size = self.image_size * (2 if self.anti_aliasing else 1)
            images = functional.render_indexed(mesh.vertices, mesh.faces, mesh.textures, image_size=size,
                                               **{name: getattr(self, name) for name in _RENDER_ARGS})
            return F.avg_pool2d(images, kernel_size=2, stride=2) if self.anti_aliasing else images
        return self.forward_tensors(mesh.face_vertices, mesh.face_textures)
