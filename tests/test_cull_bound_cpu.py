"""The exact-culling argument rests on one inequality: for a face marked cullable,
    d_reference(pixel, face) >= d_true(pixel, face) - E_face
where d_reference is the distance the reference's fp32 arithmetic produces (oracle mode 1 = its GPU contraction) and
E_face is the bound prep_face_record() (gendr_b200/csrc/gendr_device.cuh) derives from the face's own numbers.  This
test restates E_face in numpy and checks the inequality on random faces -- well shaped, thin, and slivers down to
aspect 1e-5 -- against exact (float64) point-triangle distances."""
import ctypes as C

import numpy as np

EPS = np.float32(1.1920929e-7)


def e_face(v):
    """numpy float32 restatement of the E_face / cullable computation in prep_face_record()."""
    f = np.float32
    x0, y0, x1, y1, x2, y2 = (f(v[i]) for i in (0, 1, 3, 4, 6, 7))
    adj = [y1 - y2, x2 - x1, f(np.float64(x1) * y2) - f(x2 * y1), y2 - y0, x0 - x2, f(x2 * y0 - x0 * y2), y0 - y1, x1 - x0, f(x0 * y1 - x1 * y0)]
    adj = [f(a) for a in adj]
    det_raw = f(f(x2 * adj[6]) + f(x0 * adj[0]) + f(x1 * adj[3]))
    det = det_raw if abs(det_raw) > 1e-10 else f(np.copysign(1e-10, det_raw if det_raw != 0 else -1.0))
    inv = [f(a / det) for a in adj]
    pmax = max(abs(x0), abs(x1), abs(x2), abs(y0), abs(y1), abs(y2))
    adet = abs(det_raw)
    terms = abs(x2 * adj[6]) + abs(x0 * adj[0]) + abs(x1 * adj[3])
    if adet == 0:
        return np.inf, False
    rho = 4 * EPS * terms / adet
    wsum = sum(abs(i) for i in inv)
    E = 2 * (rho * (1.5 + pmax) + pmax * 8 * EPS * wsum + 6 * EPS * pmax ** 3 / adet + 4 * EPS * pmax)
    edges2 = [(x0 - x1) ** 2 + (y0 - y1) ** 2, (x1 - x2) ** 2 + (y1 - y2) ** 2, (x2 - x0) ** 2 + (y2 - y0) ** 2]
    cullable = adet > 1e-9 and rho < 0.01 and min(edges2) > 0 and np.isfinite(E) and E < 4
    return float(E), bool(cullable)


def true_distance(v, p):
    P = np.array([[v[0], v[1]], [v[3], v[4]], [v[6], v[7]]], dtype=np.float64)
    p = np.asarray(p, dtype=np.float64)

    def seg(a, b):
        ab = b - a
        t = np.clip(np.dot(p - a, ab) / max(np.dot(ab, ab), 1e-300), 0, 1)
        return np.linalg.norm(a + t * ab - p)
    d = min(seg(P[0], P[1]), seg(P[1], P[2]), seg(P[2], P[0]))
    c = [(P[(k + 1) % 3][0] - P[k][0]) * (p[1] - P[k][1]) - (P[(k + 1) % 3][1] - P[k][1]) * (p[0] - P[k][0]) for k in range(3)]
    inside = all(x >= 0 for x in c) or all(x <= 0 for x in c)
    return 0.0 if inside else d


def test_reference_distance_never_undershoots_by_more_than_e_face(port_oracle):
    lib = port_oracle.lib
    lib.gendr_oracle_set_mode(1)
    fn = lib.gendr_oracle_pair_geometry
    fn.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_void_p]
    rng = np.random.default_rng(11)
    buf = np.zeros(10, np.float32)
    n_cullable = n_checked = 0
    worst = 0.0
    try:
        for i in range(6000):
            c = rng.uniform(-0.9, 0.9, 2)
            size = rng.choice([0.3, 0.05, 0.01])
            off = rng.uniform(-1, 1, (3, 2)) * size
            squash = rng.choice([1.0, 1e-1, 1e-2, 1e-3, 1e-4, 1e-5])
            ang = rng.uniform(0, np.pi)
            R = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
            off = (off * np.array([1.0, squash])) @ R.T
            v = np.zeros(9, np.float32)
            v[[0, 1, 3, 4, 6, 7]] = (c + off).astype(np.float32).ravel()
            v[[2, 5, 8]] = 3.0
            E, ok = e_face(v)
            if not ok:
                continue
            n_cullable += 1
            for _ in range(6):
                p = rng.uniform(-1, 1, 2).astype(np.float32)
                fn(v.ctypes.data, float(p[0]), float(p[1]), buf.ctypes.data)
                if buf[9] != 1 or buf[8] > 0:           # undefined-reference case / pixel classified inside
                    dt = true_distance(v, p)
                    if buf[8] > 0:                      # "inside" must only ever happen within E of the triangle
                        assert dt <= E + 1e-7, (v, p, dt, E)
                    continue
                d_ref = float(np.hypot(np.float64(buf[6]), np.float64(buf[7])))
                d_true = true_distance(v, p)
                n_checked += 1
                worst = max(worst, (d_true - d_ref) / max(E, 1e-30))
                assert d_ref >= d_true - E, ('reference distance undershoots the bound', v, p, d_ref, d_true, E)
    finally:
        lib.gendr_oracle_set_mode(0)
    assert n_cullable > 3000 and n_checked > 15000
    assert worst < 0.5, 'bound should be comfortably conservative (worst undershoot / E = %.3f)' % worst
