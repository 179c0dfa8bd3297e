"""ctypes loader for libgendr_b200.so -- the hand-written sm_100a CUDA library behind the C ABI in
include/gendr_b200.h.  There is deliberately NO fallback: if the library is missing or the tensors are not on a
CUDA device the call raises.  (The reference's own boundary is a pybind11 module,
/root/reference/gendr/cuda/generalized_renderer_cuda.cpp:230-237; here the same functions sit behind a C ABI.)
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GENDR_B200_LIB selects an alternative build of the same library (tuning experiments); default: the in-tree build
LIB_PATH = os.environ.get('GENDR_B200_LIB') or os.path.join(_HERE, 'libgendr_b200.so')


class RenderParams(C.Structure):
    """Mirror of `struct gendr_render_params` (include/gendr_b200.h)."""
    _fields_ = [
        ('image_size', C.c_int),
        ('dist_func', C.c_int), ('dist_scale', C.c_float), ('dist_squared', C.c_int),
        ('dist_shape', C.c_float), ('dist_shift', C.c_float), ('dist_eps', C.c_float),
        ('aggr_alpha_func', C.c_int), ('aggr_alpha_t_conorm_p', C.c_float),
        ('aggr_rgb_func', C.c_int), ('aggr_rgb_eps', C.c_float), ('aggr_rgb_gamma', C.c_float),
        ('near_plane', C.c_float), ('far_plane', C.c_float), ('double_side', C.c_int), ('texture_type', C.c_int),
        ('background', C.c_float * 3),
    ]


class CameraParams(C.Structure):
    """Mirror of `struct gendr_camera_params` (include/gendr_b200.h)."""
    _fields_ = [('mode', C.c_int), ('perspective', C.c_int), ('viewing_angle', C.c_float), ('viewing_scale', C.c_float),
                ('at_or_direction', C.c_float * 3), ('up', C.c_float * 3)]


class LightParams(C.Structure):
    """Mirror of `struct gendr_light_params` (include/gendr_b200.h)."""
    _fields_ = [('intensity_ambient', C.c_float), ('color_ambient', C.c_float * 3),
                ('intensity_directional', C.c_float), ('color_directional', C.c_float * 3), ('direction', C.c_float * 3)]


# every symbol include/gendr_b200.h declares: (restype, argtypes)
_P, _F, _I, _SZ = C.c_void_p, C.c_float, C.c_int, C.c_size_t
_PP = C.POINTER(RenderParams)
_PC, _PL = C.POINTER(CameraParams), C.POINTER(LightParams)
SIGNATURES = {
    'gendr_workspace_bytes': (_SZ, [_I, _I]),
    'gendr_forward_render': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _PP, _I, _P, _SZ, _P]),
    'gendr_backward_render': (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _PP, _I, _I, _P, _SZ, _P]),
    'gendr_backward_render_batchsum': (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _PP, _I, _I, _P, _SZ, _P]),
    'gendr_backward_render_indexed_batchsum': (_I, [_P, _I, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _PP, _I, _P, _SZ, _P]),
    'gendr_forward_render_indexed': (_I, [_P, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _PP, _P, _SZ, _P]),
    'gendr_backward_render_indexed': (_I, [_P, _I, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _PP, _I, _P, _SZ, _P]),
    'gendr_forward_render_aa': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _PP, _P, _SZ, _P]),
    'gendr_backward_render_aa': (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _PP, _I, _I, _P, _SZ, _P]),
    'gendr_camera_forward': (_I, [_P, _P, _I, _P, _I, _I, _PC, _P]),
    'gendr_camera_backward': (_I, [_P, _P, _I, _P, _P, _P, _P, _I, _I, _PC, _P]),
    'gendr_lighting_forward': (_I, [_P, _P, _I, _P, _P, _I, _I, _I, _I, _PL, _P]),
    'gendr_lighting_backward': (_I, [_P, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _PL, _P]),
    'gendr_vertex_lighting_forward': (_I, [_P, _P, _I, _P, _P, _P, _I, _I, _I, _PL, _P]),
    'gendr_vertex_lighting_backward': (_I, [_P, _P, _I, _P, _P, _P, _P, _P, _P, _I, _I, _I, _PL, _P]),
    'gendr_scene_workspace_bytes': (_SZ, [_I, _I, _I, _I]),
    'gendr_scene_forward': (_I, [_P, _I, _P, _I, _P, _P, _I, _PC, _PL, _P, _P, _P, _I, _I, _I, _I, _PP, _P, _SZ, _P]),
    'gendr_scene_backward': (_I, [_P, _I, _P, _I, _P, _P, _I, _PC, _PL, _P, _P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _PP, _P, _SZ, _P]),
    'gendr_voxelize_workspace_bytes': (_SZ, [_I, _I]),
    'gendr_voxelize': (_I, [_P, _P, _I, _I, _I, _P, _SZ, _P]),
    'gendr_render_forward_backward_host': (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _PP]),
    'gendr_release_host_scratch': (None, []),
    'gendr_sigmoid_forward': (_F, [_I, _F, _F, _F, _F, _F]),
    'gendr_sigmoid_backward': (_F, [_I, _F, _F, _F, _F, _F]),
    'gendr_t_conorm_forward': (_F, [_I, _F, _F, _I, _F]),
    'gendr_t_conorm_backward': (_F, [_I, _F, _F, _I, _F]),
    'gendr_last_error': (C.c_char_p, []),
    'gendr_version': (C.c_char_p, []),
    'gendr_launch_count': (C.c_longlong, []),
    'gendr_probe_pairs': (_I, [_P, _P, _P, _I, _P]),
    'gendr_selftest_division': (C.c_longlong, [C.c_longlong]),
    'gendr_selftest_scalar_device': (_F, [_I, _I, _F, _F, _F, _F, _F, _F]),
}

_lib = None


def load():
    """Load the library (once) and bind every exported symbol; raises if the CUDA library is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                'gendr_b200: %s not found -- build it with `python -c "import __graft_entry__ as g; g.build()"` or '
                '`make -C gendr_b200/csrc`.  There is no CPU or PyTorch fallback.' % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


class GendrCudaError(RuntimeError):
    pass


def check(code):
    if code != 0:
        raise GendrCudaError('%s (code %d)' % (load().gendr_last_error().decode(), code))
