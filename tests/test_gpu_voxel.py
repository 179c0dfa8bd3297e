"""GPU tests (pytest -m gpu) of the voxelizer (gendr_voxelize through the C ABI): bit-exact against the
golden vectors of the reference kernels, against the C oracle in its GPU-contraction mode on larger inputs, and against the
reference's own CUDA extension + Python driver (baseline/_ref) when staged."""
import os

import numpy as np
import pytest
import torch

import scenes
from ref_gpu import load_reference

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'voxel_v1.npz'))


def _dev():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda:0')


def _ours(faces, size, dev):
    import gendr_b200 as gd
    return gd.functional.voxelization(torch.as_tensor(faces, device=dev), size).cpu().numpy()


@pytest.mark.parametrize('name', ['sphere32', 'sphere40', 'soup16', 'flat24'])
def test_voxelizer_bit_identical_to_reference_golden(name):
    dev = _dev()
    faces, size = GOLD[name + '_faces'], int(GOLD[name + '_size'])
    n = faces.shape[0] * size ** 3
    want = np.unpackbits(GOLD[name + '_packed'])[:n].reshape(faces.shape[0], size, size, size).astype(np.int32)
    got = _ours(faces, size, dev)
    assert got.dtype == np.int32 and np.array_equal(got, want), 'differs in %d voxels' % int((got != want).sum())


@pytest.mark.parametrize('size,sub,batch', [(32, 3, 8), (64, 3, 2), (100, 2, 2), (7, 1, 1)])
def test_voxelizer_vs_oracle(size, sub, batch):
    """icosphere meshes (1280 / 320 / 80 faces) with random scale, sizes covering 1..4 mask words per row, the shared-memory
    and the global-scratch flood fill (100^3 does not fit shared memory)."""
    from oracle.voxel_oracle import VoxelOracle
    dev = _dev()
    verts, faces = scenes.icosphere(sub)
    g = torch.Generator().manual_seed(size)
    import gendr_b200 as gd
    v = (verts * 0.42)[None].repeat(batch, 1, 1) * (1 + 0.15 * torch.rand(batch, verts.shape[0], 1, generator=g))
    fv = gd.functional.face_vertices(v, faces[None].repeat(batch, 1, 1))
    inp = (fv * size / (size - 1) + 0.5).numpy().astype(np.float32)
    oracle = VoxelOracle('port')
    oracle.set_mode(1)
    try:
        want = oracle.voxelize(inp, size)
    finally:
        oracle.set_mode(0)
    got = _ours(inp, size, dev)
    assert np.array_equal(got, want), 'differs in %d voxels' % int((got != want).sum())
    c = size // 2
    assert got[:, c, c, c].min() == 1 and got[:, 0, 0, 0].max() == 0


def test_voxelizer_edge_cases():
    import gendr_b200 as gd
    dev = _dev()
    out = gd.functional.voxelization(torch.zeros(2, 0, 3, 3, device=dev), 16)
    assert out.shape == (2, 16, 16, 16) and int(out.sum()) == 0
    # NaN faces: every comparison of the reference is false, so every ray "hits" at zi = F2I(NaN) = 0 (GPU conversion) --
    # reproduced, not sanitised (oracle mode 1 models the GPU's float -> int conversion)
    from oracle.voxel_oracle import VoxelOracle
    oracle = VoxelOracle('port')
    oracle.set_mode(1)
    try:
        want = oracle.voxelize(np.full((1, 2, 3, 3), np.nan, np.float32), 8)
    finally:
        oracle.set_mode(0)
    out = gd.functional.voxelization(torch.full((1, 2, 3, 3), float('nan'), device=dev), 8)
    assert np.array_equal(out.cpu().numpy(), want) and int(out.sum()) > 1
    with pytest.raises(TypeError):
        gd.functional.voxelization(torch.zeros(1, 1, 3, 3), 8)


def test_voxelizer_vs_reference_cuda_and_mesh_api():
    import gendr_b200 as gd
    dev = _dev()
    ref = load_reference()
    if ref is None or not hasattr(ref.functional, 'voxelization'):
        pytest.skip('baseline/_ref (reference CUDA build with the voxelization extension) not staged')
    verts, faces = scenes.icosphere(3)
    B = 6
    g = torch.Generator().manual_seed(3)
    v = ((verts * 0.45)[None].repeat(B, 1, 1) * (1 + 0.2 * torch.rand(B, verts.shape[0], 1, generator=g))).to(dev)
    f = faces[None].repeat(B, 1, 1).to(dev)
    for size in (32, 48):
        want = ref.Mesh(v, f).voxelize(size)
        got = gd.Mesh(v, f).voxelize(size)
        assert got.dtype == want.dtype and torch.equal(got, want), 'size %d: differs in %d voxels' % (size, int((got != want).sum()))


@pytest.mark.parametrize('size', [16, 32, 50])
def test_voxelizer_box_culling_is_exact_on_hard_faces(size):
    """The surface kernel visits only the lattice rays inside each face's (certified) bounding box, the oracle tests every ray
    against every face like the reference: slivers down to aspect 1e-7, faces larger than the cube, tiny faces, faces on exact
    lattice planes and partly outside must all give the bit-identical grid."""
    from oracle.voxel_oracle import VoxelOracle
    dev = _dev()
    rng = np.random.default_rng(size)
    faces = []
    for k in range(240):
        c = rng.random(3) * 1.2 - 0.1
        kind = k % 6
        if kind == 0:       # ordinary
            tri = c + (rng.random((3, 3)) - 0.5) * 0.3
        elif kind == 1:     # sliver: third vertex almost on the line through the first two
            a, d = c, (rng.random(3) - 0.5) * 0.6
            tri = np.stack([a, a + d, a + d * rng.random() + (rng.random(3) - 0.5) * 10.0 ** rng.uniform(-7, -2)])
        elif kind == 2:     # larger than the cube
            tri = c + (rng.random((3, 3)) - 0.5) * 4.0
        elif kind == 3:     # tiny
            tri = c + (rng.random((3, 3)) - 0.5) * 10.0 ** rng.uniform(-5, -2)
        elif kind == 4:     # vertices on exact lattice positions
            tri = rng.integers(0, size + 1, (3, 3)) / size
        else:               # far outside with one vertex inside
            tri = np.stack([c, c + (rng.random(3) - 0.5) * 50, c + (rng.random(3) - 0.5) * 50])
        faces.append(tri)
    inp = np.asarray(faces, np.float32).reshape(2, 120, 3, 3)
    oracle = VoxelOracle('port')
    oracle.set_mode(1)
    try:
        want = oracle.voxelize(inp, size)
    finally:
        oracle.set_mode(0)
    got = _ours(inp, size, dev)
    assert np.array_equal(got, want), 'differs in %d voxels' % int((got != want).sum())
