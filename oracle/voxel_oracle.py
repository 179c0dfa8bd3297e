"""oracle/voxel_oracle.py -- TEST INFRASTRUCTURE ONLY.  ctypes front-end to the voxelizer checkers built by oracle/Makefile:
  kind="port"       oracle/libgendr_voxel_oracle.so          C restatement (gendr_voxel_oracle.c; mode 0 = C semantics, 1 = GPU FMA pattern)
  kind="reference"  oracle/_ref/libgendr_ref_voxel_cpu.so    the unmodified reference kernels through ref_shim.h (needs /root/reference)
Both compute gendr.functional.voxelization(faces, size, normalize=False) (functional/voxelization.py:45-62)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATHS = {'port': os.path.join(_HERE, 'libgendr_voxel_oracle.so'), 'reference': os.path.join(_HERE, '_ref', 'libgendr_ref_voxel_cpu.so')}


def build():
    subprocess.run(['make', '-C', _HERE, 'all'], check=True, capture_output=True)


def available(kind):
    return os.path.exists(_PATHS[kind])


class VoxelOracle:
    def __init__(self, kind='port'):
        if not available(kind):
            build()
        self.kind = kind
        self.lib = C.CDLL(_PATHS[kind])
        self.fn = getattr(self.lib, 'gendr_voxel_oracle' if kind == 'port' else 'gendr_voxel_ref')
        self.fn.restype = C.c_int
        self.fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]

    def set_mode(self, mode):
        if self.kind == 'port':
            self.lib.gendr_voxel_oracle_set_mode(int(mode))

    def voxelize(self, faces, size):
        faces = np.ascontiguousarray(faces, np.float32)
        B, F = faces.shape[:2]
        out = np.zeros((B, size, size, size), np.int32)
        self.fn(faces.ctypes.data, out.ctypes.data, B, F, int(size))
        return out
