"""CPU tests of the voxelizer oracle (oracle/gendr_voxel_oracle.c) against the golden vectors generated from the unmodified
reference kernels (tests/golden/voxel_v1.npz) and, where /root/reference is present, against that build run live."""
import os

import numpy as np
import pytest

import scenes
from oracle.voxel_oracle import VoxelOracle, available, build

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'voxel_v1.npz'))
NAMES = ['sphere32', 'sphere40', 'soup16', 'flat24']


def _gold(name):
    faces, size = GOLD[name + '_faces'], int(GOLD[name + '_size'])
    n = faces.shape[0] * size ** 3
    return faces, size, np.unpackbits(GOLD[name + '_packed'])[:n].reshape(faces.shape[0], size, size, size).astype(np.int32)


@pytest.fixture(scope='module')
def port():
    build()
    p = VoxelOracle('port')
    p.set_mode(0)
    return p


@pytest.mark.parametrize('name', NAMES)
def test_port_bit_identical_to_golden(port, name):
    faces, size, want = _gold(name)
    assert np.array_equal(faces, scenes.voxel_cases()[name][0])          # the committed inputs are the generator's
    assert np.array_equal(port.voxelize(faces, size), want)


def test_port_bit_identical_to_reference_build(port):
    if not available('reference'):
        pytest.skip('oracle/_ref/libgendr_ref_voxel_cpu.so not built (needs /root/reference)')
    ref = VoxelOracle('reference')
    rng = np.random.default_rng(1)
    for size, nf in ((8, 5), (20, 40), (33, 25)):
        faces = (rng.random((2, nf, 1, 3)) * 1.1 - 0.05 + (rng.random((2, nf, 3, 3)) - 0.5) * 0.6).astype(np.float32)
        assert np.array_equal(port.voxelize(faces, size), ref.voxelize(faces, size)), (size, nf)


def test_known_answers(port):
    """A closed sphere fills; an open soup encloses nothing beyond its own surface voxels; empty input is all outside."""
    faces, size, want = _gold('sphere32')
    assert want[:, 16, 16, 16].tolist() == [1, 1, 1] and want[:, 0, 0, 0].tolist() == [0, 0, 0]
    assert 0.15 < want.mean() < 0.6
    assert port.voxelize(np.zeros((1, 0, 3, 3), np.float32), 8).sum() == 0
    # the GPU contraction pattern (mode 1) changes nothing on these inputs
    port.set_mode(1)
    try:
        for name in NAMES:
            f, s, w = _gold(name)
            assert np.array_equal(port.voxelize(f, s), w), name
    finally:
        port.set_mode(0)
