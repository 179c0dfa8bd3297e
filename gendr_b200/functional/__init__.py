from .geometry import (ambient_lighting, directional_lighting, face_vertices, get_points_from_angles, load_obj, look,
                       look_at, orthogonal, perspective, vertex_normals, voxelization)
from .renderer import (GenDRFunction, GenDRIndexedFunction, GenDRSceneFunction, VertexLightingFunction, make_camera_params,
                       make_light_params, render, render_indexed, render_scene, vertex_lighting)

__all__ = ['ambient_lighting', 'directional_lighting', 'face_vertices', 'get_points_from_angles', 'load_obj', 'look',
           'look_at', 'orthogonal', 'perspective', 'vertex_normals', 'voxelization', 'GenDRFunction', 'GenDRIndexedFunction', 'GenDRSceneFunction',
           'VertexLightingFunction', 'make_camera_params', 'make_light_params', 'render', 'render_indexed', 'render_scene', 'vertex_lighting']
