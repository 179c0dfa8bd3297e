"""Generates tests/golden/voxel_v1.npz from the UNMODIFIED reference voxelizer kernels run on the CPU through
oracle/ref_shim.h (oracle/_ref/libgendr_ref_voxel_cpu.so; needs /root/reference, i.e. the build container).
Run:  python tests/golden/make_golden_voxel.py        Outputs are stored bit-packed (np.packbits)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import scenes  # noqa: E402
from oracle.voxel_oracle import VoxelOracle  # noqa: E402


def main():
    ref = VoxelOracle('reference')
    out = {}
    for name, (faces, size) in scenes.voxel_cases().items():
        vox = ref.voxelize(faces, size)
        out[name + '_faces'], out[name + '_size'], out[name + '_packed'] = faces, np.int32(size), np.packbits(vox.astype(np.uint8))
        print(name, faces.shape, size, 'filled fraction %.3f' % vox.mean())
    np.savez_compressed(os.path.join(HERE, 'voxel_v1.npz'), **out)


if __name__ == '__main__':
    main()
