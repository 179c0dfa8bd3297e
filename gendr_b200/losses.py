"""Mesh regularisers with the reference's interface (gendr/losses.py:11-120): `LaplacianLoss(vertex, faces, average)` and
`FlattenLoss(faces, average)`.  Plain torch -- O(V) / O(E) glue next to the rasterizer, not on the hot path; kept so that the
reference's optimisation scripts (experiments/opt_shape.py) run unchanged against this package.
"""
import numpy as np
import torch
import torch.nn as nn


class LaplacianLoss(nn.Module):
    """|| L x ||^2 with L the uniform graph Laplacian, rows scaled to a unit diagonal (losses.py:12-44)."""

    def __init__(self, vertex, faces, average=False):
        super().__init__()
        self.nv, self.nf, self.average = vertex.size(0), faces.size(0), average
        f = faces.detach().cpu().numpy().astype(np.int64)
        adjacency = np.zeros((self.nv, self.nv), dtype=np.float32)
        for a, b in ((0, 1), (1, 2), (2, 0)):
            adjacency[f[:, a], f[:, b]] = 1
            adjacency[f[:, b], f[:, a]] = 1
        degree = adjacency.sum(1)
        laplacian = np.diag(degree) - adjacency
        laplacian = laplacian / degree[:, None]          # every row divided by its diagonal entry
        self.register_buffer('laplacian', torch.from_numpy(laplacian.astype(np.float32)))

    def forward(self, x):
        y = torch.matmul(self.laplacian, x)
        per_item = y.pow(2).sum(tuple(range(1, y.ndimension())))
        return per_item.sum() / x.size(0) if self.average else per_item


class FlattenLoss(nn.Module):
    """sum over edges of (cos(dihedral) + 1)^2: zero for a flat neighbourhood (losses.py:47-120)."""

    def __init__(self, faces, average=False):
        super().__init__()
        self.nf, self.average = faces.size(0), average
        f = faces.detach().cpu().numpy().astype(np.int64)
        opposite = {}                                     # undirected edge -> opposite vertices, in face order
        for tri in f:
            for k in range(3):
                a, b, c = int(tri[k]), int(tri[(k + 1) % 3]), int(tri[(k + 2) % 3])
                opposite.setdefault((min(a, b), max(a, b)), []).append(c)
        edges = sorted(opposite)
        v0 = [e[0] for e in edges]
        v1 = [e[1] for e in edges]
        v2 = [opposite[e][0] for e in edges]              # first face on the edge (the reference's v2s) ...
        v3 = [c for e in edges for c in opposite[e][1:]]  # ... every further one (its v3s): one per edge on a closed manifold
        for name, idx in (('v0s', v0), ('v1s', v1), ('v2s', v2), ('v3s', v3)):
            self.register_buffer(name, torch.tensor(idx, dtype=torch.long))

    @staticmethod
    def _rejection(a, b, eps):
        """component of b orthogonal to a, and its length computed as |b| sin(angle) with the reference's eps placement"""
        a2, b2 = a.pow(2).sum(-1), b.pow(2).sum(-1)
        al, bl = (a2 + eps).sqrt(), (b2 + eps).sqrt()
        ab = (a * b).sum(-1)
        cos = ab / (al * bl + eps)
        sin = (1 - cos.pow(2) + eps).sqrt()
        return b - a * (ab / (a2 + eps))[:, :, None], bl * sin

    def forward(self, vertices, eps=1e-6):
        p0, p1 = vertices[:, self.v0s, :], vertices[:, self.v1s, :]
        edge = p1 - p0
        r1, l1 = self._rejection(edge, vertices[:, self.v2s, :] - p0, eps)
        r2, l2 = self._rejection(edge, vertices[:, self.v3s, :] - p0, eps)
        cos = (r1 * r2).sum(-1) / (l1 * l2 + eps)
        per_item = (cos + 1).pow(2).sum(tuple(range(1, cos.ndimension())))
        return per_item.sum() / vertices.size(0) if self.average else per_item
