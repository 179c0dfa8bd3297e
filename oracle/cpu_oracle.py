"""oracle/cpu_oracle.py -- TEST INFRASTRUCTURE ONLY (imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py; never by the product package gendr_b200).

ctypes front-end to the two CPU checkers built by oracle/Makefile:

  kind="port"       oracle/libgendr_oracle.so          our C restatement of the reference algorithm
  kind="reference"  oracle/_ref/libgendr_ref_cpu.so    the unmodified reference kernels run through ref_shim.h

Both export the same C interface and follow the Python-side allocation/initialisation of the reference's
autograd.Function (/root/reference/gendr/functional/renderer.py:130-151 forward buffers, :191-197 backward
buffers): faces_info zeros [B,F,27], aggrs_info zeros [B,2,S,S], soft_colors ones [B,4,S,S] with the RGB planes
scaled by background_color; grad buffers zero-initialised.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

# name -> id maps of the reference (gendr/functional/renderer.py:44-83)
DIST_FUNCS = {
    'hard': 0, 'heaviside': 0, 'uniform': 1, 'cubic_hermite': 2, 'wigner_semicircle': 3, 'gaussian': 4,
    'laplace': 5, 'logistic': 6, 'gudermannian': 7, 'hyperbolic_secant': 7, 'cauchy': 8, 'reciprocal': 9,
    'gumbel_max': 10, 'gumbel_min': 11, 'exponential': 12, 'exponential_rev': 13, 'gamma': 14,
    'gamma_rev': 15, 'levy': 16, 'levy_rev': 17,
}
AGGR_ALPHA_FUNCS = {
    'hard': 0, 'max': 1, 'probabilistic': 2, 'einstein': 3, 'hamacher': 4, 'frank': 5, 'yager': 6,
    'aczel_alsina': 7, 'dombi': 8, 'schweizer_sklar': 9,
}
AGGR_RGB_FUNCS = {'hard': 0, 'softmax': 1}
TEXTURE_TYPES = {'surface': 0, 'vertex': 1}


class Params(C.Structure):
    _fields_ = [
        ('image_size', C.c_int),
        ('dist_func', C.c_int), ('dist_scale', C.c_float), ('dist_squared', C.c_int),
        ('dist_shape', C.c_float), ('dist_shift', C.c_float), ('dist_eps', C.c_float),
        ('aggr_alpha_func', C.c_int), ('aggr_alpha_t_conorm_p', C.c_float),
        ('aggr_rgb_func', C.c_int), ('aggr_rgb_eps', C.c_float), ('aggr_rgb_gamma', C.c_float),
        ('near', C.c_float), ('far', C.c_float), ('double_side', C.c_int), ('texture_type', C.c_int),
    ]


def make_params(image_size=256, dist_func='uniform', dist_scale=1e-2, dist_squared=False, dist_shape=None,
                dist_shift=None, dist_eps=1e4, aggr_alpha_func='probabilistic', aggr_alpha_t_conorm_p=None,
                aggr_rgb_func='softmax', aggr_rgb_eps=1e-3, aggr_rgb_gamma=1e-3, near=1, far=100,
                double_side=True, texture_type='surface'):
    """Same keyword surface (and defaults) as gendr.functional.render (functional/renderer.py:239-262);
    None -> 0.0 for the three optional shape parameters."""
    def _id(v, m):
        return v if isinstance(v, int) else m[v]
    return Params(int(image_size), _id(dist_func, DIST_FUNCS), float(dist_scale), int(bool(dist_squared)),
                  float(dist_shape or 0.0), float(dist_shift or 0.0), float(dist_eps),
                  _id(aggr_alpha_func, AGGR_ALPHA_FUNCS), float(aggr_alpha_t_conorm_p or 0.0),
                  _id(aggr_rgb_func, AGGR_RGB_FUNCS), float(aggr_rgb_eps), float(aggr_rgb_gamma),
                  float(near), float(far), int(bool(double_side)), _id(texture_type, TEXTURE_TYPES))


def lib_path(kind):
    return os.path.join(_HERE, '_ref', 'libgendr_ref_cpu.so') if kind == 'reference' \
        else os.path.join(_HERE, 'libgendr_oracle.so')


def build(verbose=False):
    """Compile the checkers (the port always; the reference shim only where /root/reference exists)."""
    out = subprocess.run(['make', '-C', _HERE, 'all'], capture_output=True, text=True)
    if verbose or out.returncode:
        print(out.stdout, out.stderr)
    if out.returncode:
        raise RuntimeError('oracle build failed')


def available(kind):
    return os.path.exists(lib_path(kind))


class Oracle:
    def __init__(self, kind='port'):
        path = lib_path(kind)
        if not os.path.exists(path):
            if kind == 'port':
                build()
            else:
                raise FileNotFoundError(path)
        self.kind = kind
        self.lib = C.CDLL(path)
        self.lib.gendr_oracle_kind.restype = C.c_char_p
        for n in ('sigmoid_forward', 'sigmoid_backward'):
            f = getattr(self.lib, 'gendr_oracle_' + n)
            f.restype = C.c_float
            f.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float]
        for n in ('t_conorm_forward', 't_conorm_backward'):
            f = getattr(self.lib, 'gendr_oracle_' + n)
            f.restype = C.c_float
            f.argtypes = [C.c_int, C.c_float, C.c_float, C.c_int, C.c_float]

    # -- scalar functions (reference pybind names: sigmoid_forward, ..., K.cpp:233-236) -------------------
    def sigmoid_forward(self, fid, sign, x, scale, shape=0.0, shift=0.0):
        return float(self.lib.gendr_oracle_sigmoid_forward(fid, sign, x, scale, shape, shift))

    def sigmoid_backward(self, fid, sign, x, scale, shape=0.0, shift=0.0):
        return float(self.lib.gendr_oracle_sigmoid_backward(fid, sign, x, scale, shape, shift))

    def t_conorm_forward(self, tid, a, b, face_id=0, p=0.0):
        return float(self.lib.gendr_oracle_t_conorm_forward(tid, a, b, face_id, p))

    def t_conorm_backward(self, tid, a_all, b, n_faces=0, p=0.0):
        return float(self.lib.gendr_oracle_t_conorm_backward(tid, a_all, b, n_faces, p))

    # -- render ---------------------------------------------------------------------------------------------
    @staticmethod
    def _ptr(a):
        return a.ctypes.data_as(C.c_void_p)

    def forward(self, face_vertices, textures, params, background_color=(0, 0, 0), dtype=np.float32):
        """-> dict(soft_colors [B,4,S,S], aggrs_info [B,2,S,S], faces_info [B,F,27], faces, textures)."""
        faces = np.ascontiguousarray(face_vertices, dtype=dtype).reshape(face_vertices.shape[0], -1, 9)
        B, F = faces.shape[:2]
        tex = np.ascontiguousarray(textures, dtype=dtype)
        tex = tex.reshape(B, F, -1, 3) if tex.size else np.zeros((B, F, 1, 3), dtype=dtype)
        Tn = tex.shape[2]
        # pad one face worth of texels: the reference reads texel R*R of face fn (= texel 0 of face fn+1),
        # SURVEY Q3; for the very last face that read is out of bounds in the reference (UB) -> define as 0.
        tex_padded = np.zeros(B * F * Tn * 3 + Tn * 3 + 3, dtype=dtype)
        tex_padded[:B * F * Tn * 3] = tex.ravel()
        S = params.image_size
        faces_info = np.zeros((B, F, 27), dtype=dtype)
        aggrs_info = np.zeros((B, 2, S, S), dtype=dtype)
        soft_colors = np.ones((B, 4, S, S), dtype=dtype)
        for k in range(3):
            soft_colors[:, k] *= background_color[k]
        fn = self.lib.gendr_oracle_forward_f32 if dtype == np.float32 else self.lib.gendr_oracle_forward_f64
        fn(self._ptr(faces), self._ptr(tex_padded), self._ptr(faces_info), self._ptr(aggrs_info),
           self._ptr(soft_colors), B, F, Tn, C.byref(params))
        return dict(soft_colors=soft_colors, aggrs_info=aggrs_info, faces_info=faces_info, faces=faces,
                    textures=tex, _tex_padded=tex_padded)

    def backward(self, fwd, grad_soft_colors, params):
        """-> (grad_faces [B,F,9], grad_textures [B,F,T,3])."""
        dtype = fwd['faces'].dtype
        faces, tex = fwd['faces'], fwd['textures']
        B, F = faces.shape[:2]
        Tn = tex.shape[2]
        grad_faces = np.zeros_like(faces)
        grad_tex = np.zeros(fwd['_tex_padded'].shape, dtype=dtype)
        g = np.ascontiguousarray(grad_soft_colors, dtype=dtype)
        fn = self.lib.gendr_oracle_backward_f32 if dtype == np.float32 else self.lib.gendr_oracle_backward_f64
        fn(self._ptr(faces), self._ptr(fwd['_tex_padded']), self._ptr(fwd['soft_colors']),
           self._ptr(fwd['faces_info']), self._ptr(fwd['aggrs_info']), self._ptr(grad_faces),
           self._ptr(grad_tex), self._ptr(g), B, F, Tn, C.byref(params))
        return grad_faces, grad_tex[:B * F * Tn * 3].reshape(B, F, Tn, 3)
