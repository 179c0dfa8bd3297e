from .geometry import (ambient_lighting, directional_lighting, face_vertices, get_points_from_angles, load_obj, look,
                       look_at, orthogonal, perspective, vertex_normals)
from .renderer import GenDRFunction, GenDRIndexedFunction, render, render_indexed

__all__ = ['ambient_lighting', 'directional_lighting', 'face_vertices', 'get_points_from_angles', 'load_obj', 'look',
           'look_at', 'orthogonal', 'perspective', 'vertex_normals', 'GenDRFunction', 'GenDRIndexedFunction', 'render', 'render_indexed']
