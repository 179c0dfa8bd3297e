"""`GenDR` -- the nn.Module front door, constructor-compatible with gendr.GenDR
(/root/reference/gendr/renderer.py:12-125): 18 keyword arguments with the same names and defaults, every one a
plain mutable attribute read at call time (scripts change them between calls, e.g. experiments/opt_camera.py:236),
`forward(mesh)` and `forward_tensors(face_vertices, face_textures)`, 2x supersampling + avg_pool2d anti-aliasing.
"""
import torch.nn as nn

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

    def _render_args(self):
        """image_size is doubled under anti-aliasing (renderer.py:68); the 2x2 average-pooling that follows in the
        reference (renderer.py:92-93) is fused into the kernels (anti_aliasing=True; bit-identical)."""
        args = {name: getattr(self, name) for name in _RENDER_ARGS}
        args['image_size'] = self.image_size * (2 if self.anti_aliasing else 1)
        args['anti_aliasing'] = bool(self.anti_aliasing)
        return args

    def forward_tensors(self, face_vertices, face_textures):
        return functional.render(face_vertices=face_vertices, textures=face_textures, **self._render_args())

    def forward(self, mesh):
        if mesh.texture_type == 'surface' and mesh._pending_camera is not None:
            # Lighting -> LookAt/Look were deferred by the mesh (gendr_b200/mesh.py): one fused scene node
            camera = mesh._pending_camera
            lighting = mesh._pending_light.fused_params() if mesh._pending_light is not None else None
            return functional.render_scene(mesh._vertices, mesh.faces, mesh._textures, camera._fusable_eye(mesh),
                                           camera=camera._fused_camera(), lighting=lighting, **self._render_args())
        if mesh.texture_type == 'surface' and getattr(self, 'fused_gather', True):
            # indexed path: vertices[faces] gather and its scatter-add backward run inside the CUDA kernels
            return functional.render_indexed(mesh.vertices, mesh.faces, mesh.textures, **self._render_args())
        return self.forward_tensors(mesh.face_vertices, mesh.face_textures)
