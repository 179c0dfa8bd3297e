"""C-ABI robustness on the GPU box: argument validation, the host-buffer entry point and its per-device scratch, non-square
surface textures (T = 2, 3 behave like the reference: texture_res = int(sqrt(T)) = 1), index validation of the fused paths."""
import ctypes as C

import numpy as np
import pytest
import torch

import scenes
from gendr_b200 import _lib
from gendr_b200.cuda import generalized_renderer as ext

pytestmark = pytest.mark.gpu
INVALID = 100001


def _dev():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda:0')


def _call_forward(dev, params, T=1, F=4, B=1, S=16):
    lib = _lib.load()
    faces = torch.rand(B, F, 9, device=dev)
    tex = torch.rand(B, F, T, 3, device=dev)
    colors, aggrs = torch.empty(B, 4, S, S, device=dev), torch.empty(B, 2, S, S, device=dev)
    ws = torch.empty(lib.gendr_workspace_bytes(B, F), dtype=torch.uint8, device=dev)
    return lib.gendr_forward_render(faces.data_ptr(), tex.data_ptr(), None, aggrs.data_ptr(), colors.data_ptr(), B, F, T, C.byref(params), 0,
                                    ws.data_ptr(), ws.numel(), None)


def test_invalid_ids_are_rejected():
    dev = _dev()
    ok = ext.make_params(16, 6, 0.02, False, 0., 0., 1e4, 2, 0., 1, 1e-3, 1e-3, 1., 100., True, 0)
    assert _call_forward(dev, ok) == 0
    for field, value, T in (('aggr_rgb_func', 2, 1), ('aggr_rgb_func', -1, 1), ('texture_type', 2, 1), ('texture_type', -1, 1),
                            ('dist_func', 18, 1), ('aggr_alpha_func', 10, 1), ('image_size', 0, 1)):
        p = ext.make_params(16, 6, 0.02, False, 0., 0., 1e4, 2, 0., 1, 1e-3, 1e-3, 1., 100., True, 0)
        setattr(p, field, value)
        assert _call_forward(dev, p, T=T) == INVALID, (field, value)
        assert b'invalid argument' in _lib.load().gendr_last_error()
    # vertex textures are exactly three colours per face (K.cu:186-190)
    pv = ext.make_params(16, 6, 0.02, False, 0., 0., 1e4, 2, 0., 1, 1e-3, 1e-3, 1., 100., True, 1)
    assert _call_forward(dev, pv, T=3) == 0
    for T in (1, 2, 4, 9):
        assert _call_forward(dev, pv, T=T) == INVALID
    with pytest.raises(_lib.GendrCudaError):
        _lib.check(_call_forward(dev, pv, T=4))


def test_non_square_surface_textures_match_the_oracle(port_oracle):
    """T = 2 and T = 3 surface textures: texture_res = int(sqrt(T)) = 1 (K.cu:1098), texel index 1 is the face's OWN second texel
    and receives gradient (K.cu:201) -- not the single-texel fast path."""
    dev = _dev()
    import gendr_b200 as gd
    from oracle.cpu_oracle import make_params
    port_oracle.lib.gendr_oracle_set_mode(1)
    try:
        fv, _ = scenes.soup(80, batch=2, seed=21, size=0.15)
        gen = torch.Generator().manual_seed(5)
        g = torch.randn(2, 4, 40, 40, generator=gen)
        for T in (2, 3):
            ft = torch.rand(2, fv.shape[1], T, 3, generator=gen)
            for rgb in ('softmax', 'hard'):
                kw = dict(image_size=40, dist_func='logistic', aggr_alpha_func='probabilistic', dist_scale=0.03, aggr_rgb_func=rgb)
                a, b = fv.to(dev).requires_grad_(True), ft.to(dev).requires_grad_(True)
                img = gd.functional.render(a, b, **kw)
                img.backward(g.to(dev))
                p = make_params(**kw)
                f = port_oracle.forward(fv.numpy(), ft.numpy(), p)
                gf, gt = port_oracle.backward(f, g.numpy(), p)
                assert np.abs(img.detach().cpu().numpy() - f['soft_colors']).max() <= 2e-5
                gtn = b.grad.cpu().numpy().reshape(gt.shape)
                assert np.abs(gtn - gt).max() <= 1e-4 * np.abs(gt).max() + 1e-7, (T, rgb)
                assert np.abs(a.grad.cpu().numpy().reshape(gf.shape) - gf).max() <= 1e-4 * np.abs(gf).max()
                assert np.abs(gtn.reshape(2, -1, T, 3)[:, :, 1:]).max() > 0 or rgb == 'hard'      # the second texel does get gradient
    finally:
        port_oracle.lib.gendr_oracle_set_mode(0)


def test_host_buffer_entry_point_and_scratch_release():
    dev = _dev()
    lib = _lib.load()
    fv, ft = scenes.soup(60, batch=2, seed=3, size=0.1)
    B, F = fv.shape[:2]
    S = 32
    params = ext.make_params(S, 4, 0.02, False, 0., 0., 1e4, 3, 0., 1, 1e-3, 1e-3, 1., 100., False, 0)
    g = torch.randn(B, 4, S, S, generator=torch.Generator().manual_seed(1))
    h_col, h_gf, h_gt = torch.empty(B, 4, S, S), torch.empty(B, F, 9), torch.empty(B, F, 1, 3)
    for _ in range(2):      # second round re-creates the scratch after the release
        _lib.check(lib.gendr_render_forward_backward_host(fv.contiguous().data_ptr(), ft.contiguous().data_ptr(), g.data_ptr(), h_col.data_ptr(),
                                                          h_gf.data_ptr(), h_gt.data_ptr(), B, F, 1, C.byref(params)))
        import gendr_b200 as gd
        a, b = fv.to(dev).requires_grad_(True), ft.to(dev).requires_grad_(True)
        img = gd.functional.render(a, b, image_size=S, dist_func='gaussian', aggr_alpha_func='einstein', dist_scale=0.02, double_side=False)
        img.backward(g.to(dev))
        assert torch.equal(img.detach().cpu(), h_col)
        assert torch.allclose(a.grad.cpu().view(B, F, 9), h_gf, rtol=1e-4, atol=1e-5 * float(h_gf.abs().max()))
        lib.gendr_release_host_scratch()


def test_fused_paths_validate_face_indices():
    dev = _dev()
    import gendr_b200 as gd
    verts, faces = scenes.icosphere(1)
    v = (verts * 0.5)[None].to(dev) + torch.tensor([0., 0., 3.], device=dev)
    tex = torch.ones(1, faces.shape[0], 1, 3, device=dev)
    gd.functional.render_indexed(v, faces.to(dev), tex, image_size=16)
    bad = faces.clone(); bad[0, 0] = verts.shape[0]
    with pytest.raises(IndexError):
        gd.functional.render_indexed(v, bad.to(dev), tex, image_size=16)
    neg = faces.clone(); neg[1, 1] = -1
    with pytest.raises(IndexError):
        gd.functional.render_scene(v, neg.to(dev), tex, [0., 0., -3.], image_size=16)


def test_batch_summed_backward_equals_sum_of_per_item_gradients():
    """gendr_backward_render_batchsum / _indexed_batchsum accumulate the gradient of a mesh shared by the batch into ONE [F,9] /
    [V,3] buffer == the batch sum of what gendr_backward_render(_indexed) writes per item (SURVEY 8(e) "fusion with the collective")."""
    dev = _dev()
    lib = _lib.load()
    fv, ft, kw = scenes.config_c2(batch=6, image_size=96)
    B, F = fv.shape[:2]
    S = 96
    for dist_id, tcn_id, p in ((6, 2, 0.0), (8, 6, 2.0)):          # sparse (pixel-stationary) and dense (face-stationary) backward kernels
        params = ext.make_params(S, dist_id, 0.01, False, 0., 0., 1e4, tcn_id, p, 1, 1e-3, 1e-3, 1., 100., False, 0)
        faces, tex = fv.to(dev).view(B, F, 9).contiguous(), ft.to(dev).contiguous()
        g = torch.randn(B, 4, S, S, generator=torch.Generator().manual_seed(4)).to(dev)
        colors, aggrs = torch.empty(B, 4, S, S, device=dev), torch.empty(B, 2, S, S, device=dev)
        ws = ext.workspace_for(faces)
        ext.forward_render_raw(faces, tex, None, aggrs, colors, params, False, ws)
        gf, gt = torch.empty(B, F, 9, device=dev), torch.empty(B, F, 1, 3, device=dev)
        ext.backward_render_raw(faces, tex, colors, aggrs, gf, gt, g, params, ws, True, True)
        gsum, gt2 = torch.full((F, 9), float('nan'), device=dev), torch.empty(B, F, 1, 3, device=dev)
        ext.backward_render_batchsum_raw(faces, tex, colors, aggrs, gsum, gt2, g, params, ws, True, True)
        want = gf.double().sum(0)
        assert float((gsum.double() - want).abs().max()) <= 2e-5 * float(want.abs().max()), (dist_id, tcn_id)
        assert float((gt2 - gt).abs().max()) <= 2e-5 * float(gt.abs().max())
    # indexed: [V,3]
    import gendr_b200 as gd
    verts, index = scenes.icosphere(3)
    v = gd.LookAt(viewing_angle=15).transform((verts * 0.5)[None].repeat(4, 1, 1)).to(dev).contiguous()
    idx = index.to(dev).int().contiguous()
    tex = torch.rand(4, index.shape[0], 1, 3, generator=torch.Generator().manual_seed(1)).to(dev)
    V, F = v.shape[1], idx.shape[0]
    params = ext.make_params(64, 6, 0.02, False, 0., 0., 1e4, 2, 0., 1, 1e-3, 1e-3, 1., 100., False, 0)
    colors, aggrs = torch.empty(4, 4, 64, 64, device=dev), torch.empty(4, 2, 64, 64, device=dev)
    ws = torch.empty(lib.gendr_workspace_bytes(4, F), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(lib.gendr_forward_render_indexed(v.data_ptr(), idx.data_ptr(), 1, tex.data_ptr(), aggrs.data_ptr(), colors.data_ptr(), None, 4, V, F, 1,
                                                C.byref(params), ws.data_ptr(), ws.numel(), st))
    g = torch.randn(4, 4, 64, 64, generator=torch.Generator().manual_seed(5)).to(dev)
    gv, gvs = torch.empty(4, V, 3, device=dev), torch.empty(V, 3, device=dev)
    _lib.check(lib.gendr_backward_render_indexed(idx.data_ptr(), 1, tex.data_ptr(), colors.data_ptr(), aggrs.data_ptr(), gv.data_ptr(), None, g.data_ptr(), 0,
                                                 4, V, F, 1, C.byref(params), 1, ws.data_ptr(), ws.numel(), st))
    _lib.check(lib.gendr_backward_render_indexed_batchsum(idx.data_ptr(), 1, tex.data_ptr(), colors.data_ptr(), aggrs.data_ptr(), gvs.data_ptr(), None,
                                                          g.data_ptr(), 0, 4, V, F, 1, C.byref(params), 1, ws.data_ptr(), ws.numel(), st))
    want = gv.double().sum(0)
    assert float((gvs.double() - want).abs().max()) <= 2e-5 * float(want.abs().max())


def test_sorted_cta_schedule_is_a_permutation_heaviest_first():
    """Grids of more than one wave of CTAs (148 SMs x 4) run longest-first within groups of 16 batch items (tile_order_kernel): the list the render kernels
    index with blockIdx.x must contain every (item, tile) exactly once, grouped by item group, with the candidate counts (exact below
    128 faces, binned by 16 above) non-increasing inside a group; the counts must equal a host recount from the packed rectangles."""
    dev = _dev()
    lib = _lib.load()
    B, S = 20, 256                                        # 20 x 256 tiles = 5120 CTAs; groups of 16 and 4 items
    fv, ft = scenes.soup(300, batch=B, seed=9, size=0.25)
    F, tiles = fv.shape[1], (S // 16) ** 2
    params = ext.make_params(S, 4, 0.01, False, 0., 0., 30., 3, 0., 1, 1e-3, 1e-3, 1., 100., False, 0)
    faces, tex = fv.to(dev).view(B, F, 9).contiguous(), ft.to(dev).contiguous()
    colors, aggrs = torch.empty(B, 4, S, S, device=dev), torch.empty(B, 2, S, S, device=dev)
    ws = torch.empty(lib.gendr_workspace_bytes(B, F), dtype=torch.uint8, device=dev)
    _lib.check(lib.gendr_forward_render(faces.data_ptr(), tex.data_ptr(), None, aggrs.data_ptr(), colors.data_ptr(), B, F, 1, C.byref(params), 0,
                                        ws.data_ptr(), ws.numel(), None))
    torch.cuda.synchronize()
    al = lambda x: (x + 255) & ~255
    off = al(B * F * 176) + al(B * F * 8) + 256
    n = B * tiles
    counts = ws[off:off + 4 * n].view(torch.int32).cpu().numpy()
    order = ws[off + 4 * n:off + 8 * n].view(torch.int32).cpu().numpy()
    assert sorted(order.tolist()) == list(range(n))
    group_of = (order // tiles) // 16
    assert (np.diff(group_of) >= 0).all()                 # item groups in ascending order
    key = np.where(counts < 128, counts, 128 + np.minimum((counts - 128) >> 4, 127))[order]
    for g in np.unique(group_of):
        k = key[group_of == g]
        assert (np.diff(k) <= 0).all(), g                 # heaviest tiles of the group first
    # host recount from the packed rectangles (word30 / word31 of the face records)
    rec = ws[:B * F * 176].view(torch.int32).view(B, F, 44)[:, :, 30:32].cpu().numpy().astype(np.int64)
    want = np.zeros((B, S // 16, S // 16), np.int64)
    for b in range(B):
        for f in range(F):
            wa, wb = rec[b, f, 0] & 0xffffffff, rec[b, f, 1] & 0xffffffff
            ix0, ix1, iy0, iy1 = wa & 0x3fff, (wa >> 16) & 0x3fff, wb & 0x3fff, (wb >> 16) & 0x3fff
            if ix1 < ix0 or iy1 < iy0:
                continue
            tx0, tx1, ty0, ty1 = ix0 // 16, min(ix1 // 16, S // 16 - 1), iy0 // 16, min(iy1 // 16, S // 16 - 1)
            if (tx1 - tx0 + 1) * (ty1 - ty0 + 1) > 128:
                continue
            want[b, ty0:ty1 + 1, tx0:tx1 + 1] += 1
    assert (counts.reshape(B, S // 16, S // 16) == want).all()
    assert counts.max() > 0
    # the image does not depend on the schedule: the same items rendered two at a time (512 CTAs: at most one wave, no list).  Two, because
    # the last face of an item reads its successor's texel (quirk Q3), so item b needs item b + 1 behind it to see the same bytes.
    two, agg2 = torch.empty(2, 4, S, S, device=dev), torch.empty(2, 2, S, S, device=dev)
    ws2 = torch.empty(lib.gendr_workspace_bytes(2, F), dtype=torch.uint8, device=dev)
    for b in (0, 7, B - 2):
        _lib.check(lib.gendr_forward_render(faces[b:b + 2].data_ptr(), tex[b:b + 2].data_ptr(), None, agg2.data_ptr(), two.data_ptr(), 2, F, 1,
                                            C.byref(params), 0, ws2.data_ptr(), ws2.numel(), None))
        assert torch.equal(two[0], colors[b]), b
        if b == B - 2:
            assert torch.equal(two[1], colors[B - 1])
