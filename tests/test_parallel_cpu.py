"""world_size-2 gloo tests (CPU) of the batch-sharding host logic: slices tile the batch, the shared-mesh gradient is
local-sum + one all-reduce, images reassemble in order.  The rasterizer itself is replaced by the CPU oracle here
(test infrastructure) because this container has no GPU; the -m gpu suite covers the CUDA path."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gendr_b200 import parallel


def test_shard_bounds_tile_the_batch():
    for batch in (0, 1, 7, 64, 512, 513):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, 'tests'))
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    os.environ['OMP_NUM_THREADS'] = '2'
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import scenes
    from oracle.cpu_oracle import Oracle, make_params
    oracle = Oracle('port')
    B, S = 5, 24                                     # uneven split: 3 + 2
    verts, faces = scenes.icosphere(1)
    fv, ft = scenes.render_inputs(verts * 0.5, faces, eyes=scenes.orbit_eyes(B), batch=B)    # one mesh, B views
    # sentinel face: otherwise the last face of a batch item samples the NEXT item's first texel (reference quirk Q3),
    # which is the one way batch items can see each other -- and a shard boundary would change that neighbour
    fv, ft = scenes.with_sentinel(fv, ft)
    p = make_params(image_size=S, dist_func='logistic', aggr_alpha_func='probabilistic', dist_scale=0.03)
    g = torch.randn(B, 4, S, S, generator=torch.Generator().manual_seed(0))
    lo, hi = parallel.shard_bounds(B, rank, world)
    f = oracle.forward(parallel.shard_batch(fv).numpy(), parallel.shard_batch(ft).numpy(), p)
    gf, _ = oracle.backward(f, parallel.shard_batch(g).numpy(), p)
    shared = parallel.allreduce_shared_face_grads(torch.from_numpy(gf).view(hi - lo, -1, 3, 3))
    # the batch-summed form (what gendr_backward_render_batchsum accumulates inside the backward kernel) goes to the same all-reduce
    shared_presummed = parallel.allreduce_shared_face_grads(torch.from_numpy(gf).view(hi - lo, -1, 3, 3).sum(0).contiguous())
    assert torch.allclose(shared, shared_presummed, rtol=1e-6, atol=1e-6 * float(shared.abs().max()))
    images = parallel.gather_images(torch.from_numpy(f['soft_colors']), B)
    # indexed / scene form of the same exchange: the gradient w.r.t. the batch-shared WORLD vertices [V,3] (camera backward of
    # the rank's views, numpy scene oracle) is a local sum + one all-reduce
    from oracle import scene_oracle as so
    world_v = (verts * 0.5)[None].repeat(B, 1, 1).numpy()
    eyes = scenes.orbit_eyes(B).numpy()
    g_screen = torch.randn(B, verts.shape[0], 3, generator=torch.Generator().manual_seed(1)).numpy()
    gv_local = so.camera_backward(world_v[lo:hi], eyes[lo:hi], g_screen[lo:hi], viewing_angle=15., dtype=np.float32)
    shared_v = parallel.allreduce_shared_vertex_grads(torch.from_numpy(np.ascontiguousarray(gv_local, np.float32)))
    if rank == 0:
        want_v = so.camera_backward(world_v, eyes, g_screen, viewing_angle=15., dtype=np.float32).sum(0)
        assert np.allclose(shared_v.numpy(), want_v, rtol=1e-5, atol=1e-5 * np.abs(want_v).max()), 'shared vertex gradient'

        full = oracle.forward(fv.numpy(), ft.numpy(), p)
        gfull, _ = oracle.backward(full, g.numpy(), p)
        ok_img = bool(np.array_equal(images.numpy(), full['soft_colors']))
        want = gfull.reshape(B, -1, 3, 3).sum(0)
        ok_grad = bool(np.allclose(shared.numpy(), want, rtol=1e-4, atol=1e-4 * np.abs(want).max()))
        open(os.path.join(out_dir, 'result.txt'), 'w').write('%s %s' % (ok_img, ok_grad))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_render_equals_single_process(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / 'result.txt').read() == 'True True'


def test_bind_host_to_device_helper(tmp_path, monkeypatch):
    """NUMA-local binding of a rank: cpulist parsing, and a silent no-op when there is nothing to bind to (no GPU, no sysfs entry)."""
    import os
    from gendr_b200 import parallel
    assert parallel._parse_cpulist('0-3,8,10-11\n') == {0, 1, 2, 3, 8, 10, 11}
    assert parallel._parse_cpulist('\n') == set()
    before = os.sched_getaffinity(0)
    assert parallel.bind_host_to_device(0, sysfs=str(tmp_path)) is None      # no CUDA device here / no sysfs entry: unchanged
    assert os.sched_getaffinity(0) == before

    class _Props(object):
        pci_domain_id, pci_bus_id, pci_device_id = 0, 0x1b, 0
    monkeypatch.setattr(parallel.torch.cuda, 'get_device_properties', lambda i: _Props())
    d = tmp_path / '0000:1b:00.0'
    d.mkdir()
    one = sorted(before)[0]
    (d / 'local_cpulist').write_text('%d\n' % one)
    try:
        got = parallel.bind_host_to_device(0, sysfs=str(tmp_path))
        if len(before) > 1:
            assert got == {one} and os.sched_getaffinity(0) == {one}
        else:
            assert got is None
    finally:
        os.sched_setaffinity(0, before)
