"""Multi-GPU correctness on real devices (needs >= 2 GPUs; skipped otherwise): the batch-sharded render over NCCL equals
the single-GPU render -- per-sample images bit for bit, the all-reduced shared-mesh gradient to summation noise
(SURVEY.md 8(e) "Verification"), on the dense C4 configuration."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


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
    import torch.distributed as dist
    import scenes
    import gendr_b200 as gd
    from gendr_b200 import parallel
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    B = 6
    fv, ft, cfg = scenes.config_c4(batch=B, n=32)
    fv, ft = scenes.with_sentinel(fv, ft)         # keeps quirk Q3 from coupling neighbouring batch items across the shard boundary
    kw = dict(cfg, image_size=128, double_side=False)
    g = torch.randn(B, 4, 128, 128, generator=torch.Generator().manual_seed(0))
    a = parallel.shard_batch(fv).to(dev).requires_grad_(True)
    img = gd.functional.render(a, parallel.shard_batch(ft).to(dev), **kw)
    img.backward(parallel.shard_batch(g).to(dev))
    shared = parallel.allreduce_shared_face_grads(a.grad)
    images = parallel.gather_images(img.detach(), B)
    if rank == 0:
        full = fv.to(dev).requires_grad_(True)
        ref_img = gd.functional.render(full, ft.to(dev), **kw)
        ref_img.backward(g.to(dev))
        want = full.grad.sum(0)
        ok_img = bool(torch.equal(images, ref_img.detach()))
        ok_grad = bool(((shared - want).abs().max() <= 1e-4 * want.abs().max()).item())
        open(os.path.join(out_dir, 'result.txt'), 'w').write('%s %s' % (ok_img, ok_grad))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_render_over_nccl_equals_single_gpu(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 CUDA devices')
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / 'result.txt').read() == 'True True'
