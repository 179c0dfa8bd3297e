"""CPU check of the voxelizer's per-face certificate (gendr_b200/csrc/voxel_kernels.cuh, voxel_surface_kernel): whenever a face is
certified, every lattice ray the reference's fp32 test accepts (sub1, voxelization_cuda_kernel.cu:56-71, restated here in numpy
float32) lies inside the box the kernel visits.  Hard faces: slivers down to aspect 1e-7, huge, tiny, lattice-aligned, far outside."""
import numpy as np


def _faces(rng, n, size):
    out = []
    for k in range(n):
        c = rng.random(2) * 1.2 - 0.1
        kind = k % 6
        if kind == 0:
            tri = c + (rng.random((3, 2)) - 0.5) * 0.3
        elif kind == 1:
            a, d = c, (rng.random(2) - 0.5) * 0.6
            tri = np.stack([a, a + d, a + d * rng.random() + (rng.random(2) - 0.5) * 10.0 ** rng.uniform(-7, -2)])
        elif kind == 2:
            tri = c + (rng.random((3, 2)) - 0.5) * 4.0
        elif kind == 3:
            tri = c + (rng.random((3, 2)) - 0.5) * 10.0 ** rng.uniform(-5, -2)
        elif kind == 4:
            tri = rng.integers(0, size + 1, (3, 2)) / size
        else:
            tri = np.stack([c, c + (rng.random(2) - 0.5) * 50, c + (rng.random(2) - 0.5) * 50])
        out.append(tri)
    return (np.asarray(out, np.float32) * np.float32(size)).astype(np.float32)      # faces *= size


def test_certified_box_contains_every_accepted_ray():
    f32 = np.float32
    n_cert = n_total = n_accept = 0
    for size in (16, 32, 50):
        rng = np.random.default_rng(size)
        tri = _faces(rng, 3000, size)                                   # [N,3,2] = (y, x) roles
        f0, f1 = tri[:, 0, 0], tri[:, 0, 1]
        y1d, x1d = tri[:, 1, 0] - f0, tri[:, 1, 1] - f1
        y2d, x2d = tri[:, 2, 0] - f0, tri[:, 2, 1] - f1
        det = (x1d * y2d - x2d * y1d).astype(f32)
        ylo, yhi, xlo, xhi = tri[:, :, 0].min(1), tri[:, :, 0].max(1), tri[:, :, 1].min(1), tri[:, :, 1].max(1)
        w = np.maximum(yhi - ylo, xhi - xlo)
        M = np.maximum(np.maximum(np.maximum(np.abs(ylo), np.abs(yhi)), np.maximum(np.abs(xlo), np.abs(xhi))), f32(size))
        ad = np.abs(det)
        certified = (M < 1e6) & (ad >= f32(7e-5) * w * w * M) & (ad > 1e-30) & (ad < 1e30) & (det != 0)
        y0, y1 = np.maximum(0, np.floor(ylo) - 1), np.minimum(size - 1, np.ceil(yhi) + 1)
        x0, x1 = np.maximum(0, np.floor(xlo) - 1), np.minimum(size - 1, np.ceil(xhi) + 1)
        yy, xx = np.meshgrid(np.arange(size, dtype=f32), np.arange(size, dtype=f32), indexing='ij')
        ypd = yy[None] - f0[:, None, None]
        xpd = xx[None] - f1[:, None, None]
        with np.errstate(all='ignore'):
            t1 = ((y2d[:, None, None] * xpd - x2d[:, None, None] * ypd).astype(f32) / det[:, None, None]).astype(f32)
            t2 = ((-y1d[:, None, None] * xpd + x1d[:, None, None] * ypd).astype(f32) / det[:, None, None]).astype(f32)
            accepted = ~((t1 < 0) | (t2 < 0) | (1 < (t1 + t2).astype(f32))) & (det != 0)[:, None, None]
        inside = (yy[None] >= y0[:, None, None]) & (yy[None] <= y1[:, None, None]) & (xx[None] >= x0[:, None, None]) & (xx[None] <= x1[:, None, None])
        leaked = accepted & ~inside & certified[:, None, None]
        assert not leaked.any(), 'size %d: %d accepted rays outside the certified box' % (size, int(leaked.sum()))
        n_cert += int(certified.sum()); n_total += len(det); n_accept += int((accepted & certified[:, None, None]).sum())
    # the certificate is not vacuous: most faces are certified and they do accept rays
    assert n_cert > 0.6 * n_total and n_accept > 1000, (n_cert, n_total, n_accept)
