#!/usr/bin/env python
"""bench.py -- fwd+bwd soft-rasterization throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl ours|reference] [--workload c3|c2|c4]
    (N > 1: launched by torchrun, one rank per GPU; RANK/LOCAL_RANK/WORLD_SIZE/MASTER_* from the env)

A "step" is one forward + backward pass of the hot path over one synthetic batch.  Default workload = the configuration
BASELINE.json quotes the metric on (config C3: 8192-face jittered grid sphere, 256x256, gaussian + einstein, batch 64
per GPU, GenDR defaults otherwise).  Metric: fwd+bwd Mpixel*face/s = B*S^2*F / t / 1e6 with NOMINAL
pairs (culled pairs count, exactly as they do for the reference), whole job over all ranks.

  value         device-resident inputs; CUDA events on the launching stream; max over ranks
  e2e           the same step through the C-ABI host entry (gendr_render_forward_backward_host): pinned HOST buffers in,
                H2D + forward + backward + D2H inside the timed region
  roofline      dominant kernel (backward render) against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  rank 0, N = 1: the unmodified reference kernels run on the host cores through oracle/_ref (kind
                "reference"), or the C port when that build is absent (kind "port"), on a bounded sample
  reference_cuda  (extra) the reference's own CUDA kernels (baseline/_ref, sm_100a build) on the same GPU, same inputs
  --impl reference  rank 0 only: the unmodified reference (baseline/_ref) through gendr.functional.render on one B200 --
                    GenDR's implementation of the path is its CUDA extension; there is no CPU rasterizer -- with the
                    CPU-shim figure as `cpu_baseline`; CPU shim alone when no GPU / no reference build is present
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

METRIC = 'fwd+bwd Mpixel*face/s'
UNIT = 'Mpixel*face/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c3', choices=['c2', 'c3', 'c4'])
    ap.add_argument('--batch', type=int, default=0, help='per-GPU batch (default: the workload\'s)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-reference-cuda', action='store_true')
    return ap.parse_args()


def workload(name, batch, rank=0, world=1):
    import scenes
    import torch
    cfg = {'c2': (scenes.config_c2, 16), 'c3': (scenes.config_c3, 64), 'c4': (scenes.config_c4, 64)}[name]
    b = batch or cfg[1]
    fv, ft, kw = cfg[0](batch=b * world)
    lo = rank * b
    fv, ft = fv[lo:lo + b].contiguous(), ft[lo:lo + b].contiguous()
    kw = dict(kw, double_side=False)            # GenDR module default (renderer.py:34)
    desc = {'c2': 'C2: icosphere 1280 faces, 256x256, logistic+probabilistic',
            'c3': 'C3: jittered grid sphere 8192 faces, 256x256, gaussian+einstein',
            'c4': 'C4: grid sphere 8192 faces, 256x256, cauchy+yager(p=2)'}[name]
    return fv, ft, kw, desc, b


PARAMS_NOTE = 'GenDR defaults (dist_scale 1e-2, dist_eps 1e4, softmax RGB, eps=gamma=1e-3, near 1, far 100, single-sided)'


def common_config(desc, B, world, F, S):
    """The `config` object is IDENTICAL in both arms (the driver compares them); arm-specific remarks go under `notes`."""
    return {'workload': desc, 'per_gpu_batch': B, 'global_batch': B * world, 'faces': F, 'image_size': S, 'texture_res': 1, 'params': PARAMS_NOTE}


def csrc_sha():
    """Fingerprint of the kernel sources: ncu-derived constants in profiles/roofline_traffic.json are only used when they were
    captured from exactly these sources."""
    import hashlib
    import re
    h = hashlib.sha256()
    d = os.path.join(ROOT, 'gendr_b200', 'csrc')
    for name in sorted(os.listdir(d)):
        if name.endswith(('.cu', '.cuh')):
            text = open(os.path.join(d, name), 'r', encoding='utf-8', errors='replace').read()
            # the CODE: comments and whitespace do not change the kernels (the sources hold no string literal with // or /* in it)
            text = re.sub(r'/\*.*?\*/', ' ', text, flags=re.S)
            text = re.sub(r'//[^\n]*', ' ', text)
            h.update(' '.join(text.split()).encode())
    return h.hexdigest()[:16]


def algorithmic_bytes(B, F, S, T):
    """compulsory HBM traffic of the path (faces+textures read twice, grads written once, RGBA+aggrs
    written then read, cotangent read)."""
    fwd = B * F * (36 + 12 * T) + 24 * B * S * S
    bwd = B * F * (72 + 24 * T) + 40 * B * S * S
    return fwd, bwd


class ClockSampler:
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
              'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.FIELDS, '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([time.perf_counter()] + [c.strip() for c in line.split(',')])

    def window(self, t0, t1):
        """restrict the statistics to samples that arrived inside the timed region [t0, t1] (perf_counter seconds)"""
        self.t0, self.t1 = t0, t1

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        t0, t1 = getattr(self, 't0', None), getattr(self, 't1', None)
        inside = [r[1:] for r in self.rows if t0 is not None and t0 <= r[0] <= t1 + 0.03]
        rows = inside if inside else [r[1:] for r in self.rows]
        sm = sorted(float(r[0]) for r in rows if r and r[0].replace('.', '').isdigit())
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith('active') for r in rows)]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons, 'samples': len(sm),
                'sampled': 'inside the timed region' if inside else 'around the timed region'}


def cpu_reference_arm(fv, ft, kw, steps, warmup, sample_size=128):
    """The reference's CPU implementation of the path: its unmodified kernels compiled for the host through the shim
    (oracle/_ref), else the C port.  Sample: 1 batch item of the workload at sample_size^2 pixels (throughput is per
    pixel*face, so the sample is representative: every pixel visits every face in the reference)."""
    import numpy as np
    from oracle.cpu_oracle import Oracle, available, build, make_params
    build()
    kind = 'reference' if available('reference') else 'port'
    oracle = Oracle(kind)
    cores = os.cpu_count() or 1
    os.environ.setdefault('OMP_NUM_THREADS', str(cores))
    B, F = 1, fv.shape[1]
    S = sample_size
    p = make_params(**dict(kw, image_size=S))
    f_in, t_in = fv[:1].numpy(), ft[:1].numpy()
    g = np.random.default_rng(0).standard_normal((B, 4, S, S)).astype(np.float32)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        f = oracle.forward(f_in, t_in, p)
        oracle.backward(f, g, p)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    pairs = B * S * S * F
    total = sum(times)
    return dict(value=pairs * len(times) / total / 1e6, unit=UNIT, cores=cores, kind=kind,
                sample='1 batch item, %d faces, %dx%d pixels, fwd+bwd, %d timed passes' % (F, S, S, len(times)),
                ms_per_step=1e3 * total / len(times))


def reference_arm(args):
    """`--impl reference`: the UNMODIFIED reference through its own public API (gendr.functional.render + backward) on the
    same workload.  GenDR ships no CPU rasterizer -- its implementation of the path IS its CUDA extension -- so the arm runs
    the reference's CUDA kernels (baseline/_ref, stock sources built for sm_100a) on one B200, which is the baseline
    BASELINE.json's north star names.  `value`: inputs resident in HBM; `e2e`: pinned host buffers in, H2D + render +
    backward + D2H of image and gradients inside the timed region.  `cpu_baseline` (always reported beside it): the
    reference's kernels compiled for the host through oracle/ref_shim.h, timed on a bounded sample.  Without a GPU or
    without baseline/_ref the CPU figure becomes the line's value."""
    fv, ft, kw, desc, b = workload(args.workload, args.batch)
    F, S = fv.shape[1], kw['image_size']
    cpu = None
    if not args.no_cpu_baseline:
        cpu = cpu_reference_arm(fv, ft, kw, steps=2, warmup=1)
    line = {'impl': 'reference', 'metric': METRIC, 'unit': UNIT, 'n_gpus': 1, 'steps': args.steps, 'warmup': max(3, args.warmup),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic'}
    ref = None
    try:
        import torch
        if torch.cuda.is_available():
            from ref_gpu import load_reference, reference_render
            ref = load_reference()
    except Exception as e:      # noqa: BLE001
        line['reference_cuda_error'] = repr(e)[:200]
    if ref is None:
        if cpu is None:
            cpu = cpu_reference_arm(fv, ft, kw, args.steps, args.warmup)
            line.update({'steps': args.steps, 'warmup': args.warmup})
        else:
            line.update({'steps': 2, 'warmup': 1})
        line.update({'value': cpu['value'], 'ms_per_step': cpu['ms_per_step'], 'gpu_launches': 0,
                     'config': common_config(desc, b, 1, F, S),
                     'notes': {'timed_on': 'host CPU cores: the reference CUDA kernels compiled for the host (oracle/_ref) -- no GPU or no '
                                           'baseline/_ref build available'},
                     'cpu_baseline': {k: cpu[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
                     'e2e': {'value': cpu['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}})
        return line

    import torch
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    B = b
    gcol = torch.randn(B, 4, S, S, generator=torch.Generator().manual_seed(2))
    d_fv, d_ft, d_g = fv.to(dev), ft.to(dev), gcol.to(dev)

    def step():
        a, t = d_fv.clone().requires_grad_(True), d_ft.clone().requires_grad_(True)
        reference_render(ref, a, t, **kw).backward(d_g)
        return a.grad, t.grad

    warm = max(1, min(args.warmup, 2))          # one reference step is ~1.8 s at B = 64: keep the arm within minutes
    steps = max(1, min(args.steps, 5))
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w_begin = time.perf_counter()
    t0.record()
    for _ in range(steps):
        step()
    t1.record()
    torch.cuda.synchronize()
    sampler.window(w_begin, time.perf_counter())
    clocks = sampler.stop()
    ms = t0.elapsed_time(t1) / steps
    pairs = B * S * S * F
    # end to end: pinned host buffers, copies inside the timed region
    h_fv, h_ft, h_g = fv.pin_memory(), ft.pin_memory(), gcol.pin_memory()
    h_img = torch.empty(B, 4, S, S).pin_memory()
    h_gf, h_gt = torch.empty_like(fv).pin_memory(), torch.empty_like(ft).pin_memory()

    def e2e_step():
        a = h_fv.to(dev, non_blocking=True).requires_grad_(True)
        t = h_ft.to(dev, non_blocking=True).requires_grad_(True)
        g = h_g.to(dev, non_blocking=True)
        img = reference_render(ref, a, t, **kw)
        img.backward(g)
        h_img.copy_(img.detach(), non_blocking=True); h_gf.copy_(a.grad, non_blocking=True); h_gt.copy_(t.grad, non_blocking=True)
        torch.cuda.synchronize()
    e2e_step()
    n_e2e = max(1, min(steps, 3))
    w0 = time.perf_counter()
    for _ in range(n_e2e):
        e2e_step()
    e2e_s = (time.perf_counter() - w0) / n_e2e
    line.update({'value': pairs / (ms * 1e-3) / 1e6, 'ms_per_step': ms, 'steps': steps, 'warmup': warm,
                 'config': common_config(desc, B, 1, F, S),
                 'notes': {'timed_on': "one B200: the reference's own CUDA kernels (unmodified sources, sm_100a build under baseline/_ref) through "
                                       'gendr.functional.render + backward; the reference has no multi-GPU path, so the arm is always 1 GPU',
                           'steps_note': 'steps/warmup clamped (one reference step takes seconds)'},
                 'e2e': {'value': pairs / e2e_s / 1e6, 'unit': UNIT, 'ms_per_step': e2e_s * 1e3,
                         'h2d_bytes_per_step': int((h_fv.numel() + h_ft.numel() + h_g.numel()) * 4),
                         'd2h_bytes_per_step': int((h_img.numel() + h_gf.numel() + h_gt.numel()) * 4)},
                 'gpu_launches': 0, 'clocks': clocks})
    if cpu is not None:
        line['cpu_baseline'] = {k: cpu[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
    return line


_REAL_STDOUT = None


def quiet_stdout():
    """Route everything libraries print to stdout (e.g. NCCL's version banner) to stderr, so that stdout carries exactly ONE
    JSON line (emit())."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + '\n').encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


class Rig:
    """Device-resident buffers of one workload on this rank + the two step flavours (raw C-ABI calls on preallocated buffers)."""

    def __init__(self, name, batch, rank, world, dev):
        import torch
        from gendr_b200.cuda import generalized_renderer as ext
        from gendr_b200.functional import renderer as fr
        self.ext, self.torch = ext, torch
        self.fv, self.ft, self.kw, self.desc, self.B = workload(name, batch, rank, world)
        fv, ft, kw, B = self.fv, self.ft, self.kw, self.B
        self.F, self.S, self.T = fv.shape[1], kw['image_size'], ft.shape[2]
        F, S, T = self.F, self.S, self.T
        self.params = ext.make_params(S, fr.DIST_FUNC_IDS[kw['dist_func']], 1e-2, False, None, None, 1e4, fr.AGGR_ALPHA_FUNC_IDS[kw['aggr_alpha_func']],
                                      kw.get('aggr_alpha_t_conorm_p'), 1, 1e-3, 1e-3, 1, 100, kw['double_side'], 0, (0, 0, 0))
        self.faces = fv.to(dev).view(B, F, 9).contiguous()
        self.tex = ft.to(dev).contiguous()
        self.gcol = torch.randn(B, 4, S, S, generator=torch.Generator().manual_seed(2 + rank)).to(dev)
        self.colors = torch.empty(B, 4, S, S, device=dev)
        self.aggrs = torch.empty(B, 2, S, S, device=dev)
        self.gfaces = torch.empty(B, F, 9, device=dev)
        self.gsum = torch.empty(F, 9, device=dev)          # batch-summed gradient of the shared mesh (N > 1)
        self.gtex = torch.empty(B, F, T, 3, device=dev)
        self.ws = ext.workspace_for(self.faces)

    def step_single(self, ev=None):
        """N = 1: forward + backward with per-item gradients [B,F,9] -- exactly what the reference arm produces."""
        ext = self.ext
        ext.forward_render_raw(self.faces, self.tex, None, self.aggrs, self.colors, self.params, False, self.ws)
        if ev:
            ev[0].record()
        ext.backward_render_raw(self.faces, self.tex, self.colors, self.aggrs, self.gfaces, self.gtex, self.gcol, self.params, self.ws, True, True)
        if ev:
            ev[1].record()

    def step_sharded(self, ev=None):
        """N > 1 (one mesh shared by all views): the backward kernel accumulates the gradient of the shared geometry
        straight into ONE [F,9] buffer (gendr_backward_render_batchsum) and that buffer goes to the single all-reduce -- no
        [B,F,9] intermediate, no reduction kernel."""
        import torch.distributed as dist
        ext = self.ext
        ext.forward_render_raw(self.faces, self.tex, None, self.aggrs, self.colors, self.params, False, self.ws)
        if ev:
            ev[0].record()
        ext.backward_render_batchsum_raw(self.faces, self.tex, self.colors, self.aggrs, self.gsum, self.gtex, self.gcol, self.params, self.ws, True, True)
        if ev:
            ev[1].record()
        dist.all_reduce(self.gsum, op=dist.ReduceOp.SUM)
        if ev:
            ev[2].record()


def timed_run(rig, steps, warmup, world, dev, local_rank, lib):
    """W warm-up steps, then exactly K steps between barrier + synchronize, CUDA events on the launching stream, max over ranks."""
    import torch
    import torch.distributed as dist
    step = rig.step_sharded if world > 1 else rig.step_single
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.gendr_launch_count()
    evs = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    w_begin = time.perf_counter()
    t_begin.record()
    for i in range(steps):
        step(evs[i])
    t_end.record()
    torch.cuda.synchronize()
    sampler.window(w_begin, time.perf_counter())
    if world > 1:
        dist.barrier()
    launches = lib.gendr_launch_count() - launches0
    clocks = sampler.stop()
    elapsed_ms = t_begin.elapsed_time(t_end)
    bwd_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / steps
    allreduce_ms = (sum(e[1].elapsed_time(e[2]) for e in evs) / steps) if world > 1 else 0.0
    if world > 1:
        tmax = torch.tensor([elapsed_ms, bwd_ms, allreduce_ms], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        elapsed_ms, bwd_ms, allreduce_ms = (float(x) for x in tmax.tolist())
    return {'ms_per_step': elapsed_ms / steps, 'bwd_ms': bwd_ms, 'fwd_ms': elapsed_ms / steps - bwd_ms - allreduce_ms, 'allreduce_ms': allreduce_ms,
            'launches': int(launches), 'clocks': clocks}


def sharded_equals_single(rig, rank, world, dev):
    """Hardware proof that the batch-sharded job computes what one GPU would (SURVEY 8e "Verification"): rank 0 re-renders rank 1's
    shard itself and compares the images bit for bit with what rank 1 produced; it also recomputes the gradient sum of the WHOLE
    global batch alone and compares the all-reduced gradient to 1e-4 of its maximum."""
    import torch
    import torch.distributed as dist
    from gendr_b200.cuda import generalized_renderer as ext
    rig.step_sharded()
    torch.cuda.synchronize()
    img_probe = rig.colors[:2].contiguous()                      # first two views of this rank's shard
    got_from_1 = torch.empty_like(img_probe)
    if rank == 1:
        dist.send(img_probe, dst=0)
    if rank == 0:
        dist.recv(got_from_1, src=1)
    out = None
    if rank == 0:
        total = torch.zeros_like(rig.gsum)
        images_equal = True
        for r in range(world):
            other = Rig(rig.name, rig.B, r, world, dev)
            other.step_single()
            total += other.gfaces.sum(0)
            if r == 1:
                images_equal = bool(torch.equal(other.colors[:2], got_from_1))
            del other
        torch.cuda.synchronize()
        scale = float(total.abs().max())
        err = float((rig.gsum - total).abs().max()) / max(scale, 1e-30)
        out = {'images_bit_identical': images_equal, 'allreduced_grad_max_err_over_max': err, 'ok': bool(images_equal and err <= 1e-4)}
    dist.barrier()
    return out


def main():
    args = parse()
    quiet_stdout()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))

    if args.impl == 'reference':
        if rank != 0:
            return
        emit(reference_arm(args))
        return

    import torch
    import torch.distributed as dist
    from gendr_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (the product has no CPU path); use --impl reference for the CPU arm')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    # one process per GPU: keep this rank (and the pinned host buffers of the e2e leg) on the CPUs local to its GPU
    from gendr_b200.parallel import bind_host_to_device
    host_cpus = bind_host_to_device(local_rank) if (world > 1 and not os.environ.get('GENDR_B200_NO_AFFINITY')) else None
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.load()

    rig = Rig(args.workload, args.batch, rank, world, dev)
    rig.name = args.workload
    fv, ft, kw, desc, B, F, S, T = rig.fv, rig.ft, rig.kw, rig.desc, rig.B, rig.F, rig.S, rig.T
    params, faces, tex, gcol, colors = rig.params, rig.faces, rig.tex, rig.gcol, rig.colors
    warm = max(3, args.warmup)
    res = timed_run(rig, args.steps, warm, world, dev, local_rank, lib)
    ms_per_step, bwd_ms, clocks, launches = res['ms_per_step'], res['bwd_ms'], res['clocks'], res['launches']
    pairs_per_step = B * world * S * S * F
    value = pairs_per_step / (ms_per_step * 1e-3) / 1e6

    # ---- the same step through the public Python API (gendr_b200.functional.render + .backward(): autograd node, output allocation,
    #      ctypes) -- what a user of the module calls; device-resident inputs -------------------------------------------------------
    import gendr_b200 as gd
    fv_d, ft_d = faces.view(B, F, 3, 3), tex

    def api_step():
        a, t = fv_d.detach().requires_grad_(True), ft_d.detach().requires_grad_(True)
        gd.functional.render(a, t, **kw).backward(gcol)
        return a.grad
    for _ in range(3):
        api_step()
    torch.cuda.synchronize()
    n_api = max(3, min(args.steps, 10))
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(n_api):
        api_step()
    a1.record()
    torch.cuda.synchronize()
    api_ms = a0.elapsed_time(a1) / n_api
    if world > 1:
        tmax = torch.tensor([api_ms], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        api_ms = float(tmax.item())

    # ---- e2e: host buffers through the C ABI (H2D + fwd + bwd + D2H every step) ---------------------------------
    h_faces, h_tex, h_gcol = faces.cpu().pin_memory(), tex.cpu().pin_memory(), gcol.cpu().pin_memory()
    h_col = torch.empty(B, 4, S, S).pin_memory()
    h_gfaces, h_gtex = torch.empty(B, F, 9).pin_memory(), torch.empty(B, F, T, 3).pin_memory()

    def e2e_step():
        _lib.check(lib.gendr_render_forward_backward_host(h_faces.data_ptr(), h_tex.data_ptr(), h_gcol.data_ptr(), h_col.data_ptr(),
                                                          h_gfaces.data_ptr(), h_gtex.data_ptr(), B, F, T, C.byref(params)))
    for _ in range(2):
        e2e_step()
    if world > 1:
        dist.barrier()
    n_e2e = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        e2e_step()
    e2e_s = (time.perf_counter() - t0) / n_e2e
    if world > 1:
        tmax = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_s = float(tmax.item())
    e2e_ok = bool(torch.allclose(h_col, colors.cpu(), atol=1e-6))
    e2e = {'value': pairs_per_step / e2e_s / 1e6, 'unit': UNIT, 'ms_per_step': e2e_s * 1e3,
           'h2d_bytes_per_step': int((h_faces.numel() + h_tex.numel() + h_gcol.numel()) * 4),
           'd2h_bytes_per_step': int((h_col.numel() + h_gfaces.numel() + h_gtex.numel()) * 4), 'matches_device_path': e2e_ok}
    del h_faces, h_tex, h_gcol, h_col, h_gfaces, h_gtex

    # ---- N > 1: hardware equality proof + the C4 configuration (the one multi-GPU configuration BASELINE.json names) -----------
    equal, c4 = None, None
    if world > 1:
        equal = sharded_equals_single(rig, rank, world, dev)
        if args.workload != 'c4' and not os.environ.get('GENDR_B200_BENCH_SKIP_C4'):      # (the skip is for tuning runs only)
            del rig
            torch.cuda.empty_cache()
            rig4 = Rig('c4', 0, rank, world, dev)
            rig4.name = 'c4'
            r4 = timed_run(rig4, max(2, min(args.steps, 3)), 1, world, dev, local_rank, lib)
            pairs4 = rig4.B * world * rig4.S * rig4.S * rig4.F
            c4 = {'workload': rig4.desc, 'per_gpu_batch': rig4.B, 'global_batch': rig4.B * world, 'metric': METRIC, 'unit': UNIT,
                  'value': pairs4 / (r4['ms_per_step'] * 1e-3) / 1e6, 'ms_per_step': r4['ms_per_step'], 'fwd_ms': r4['fwd_ms'], 'bwd_ms': r4['bwd_ms'],
                  'allreduce_ms': r4['allreduce_ms'], 'steps': max(2, min(args.steps, 3)), 'warmup': 1, 'clocks': r4['clocks'],
                  'note': 'C4 of BASELINE.json: 8192 faces, 256x256, cauchy + yager(p=2), 64 views per GPU (512 views on 8 GPUs), one all-reduce of '
                          'the batch-summed [F,3,3] gradient per step; timed like the headline line (CUDA events, max over ranks)'}
            del rig4

    if rank != 0:
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    fwd_bytes, bwd_bytes = algorithmic_bytes(B, F, S, T)
    traffic, traffic_note, prof = None, None, {}
    try:
        allprof = json.load(open(os.path.join(ROOT, 'profiles', 'roofline_traffic.json')))
        prof = allprof.get(args.workload, {})
        if allprof.get('csrc_sha') != csrc_sha():
            prof, traffic_note = {}, 'null: the committed ncu capture (profiles/roofline_traffic.json) is of other kernel sources than the ones running'
        elif B != prof.get('batch', 64):
            prof, traffic_note = {}, 'null: the committed ncu capture is of per-GPU batch %s' % prof.get('batch', 64)
        else:
            traffic = prof.get('backward_dram_bytes_per_launch')
    except Exception:
        traffic_note = 'null: profiles/roofline_traffic.json not readable'
    achieved = bwd_bytes / (bwd_ms * 1e-3) / 1e9
    roofline = {'bound': 'hbm', 'kernel': 'backward render kernel (render_kernel<..,BWD> pixel-stationary or render_bwd_fs_kernel face-stationary, by density)',
                'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'peak_source': 'MEASURED_PEAKS.json hbm_gbs (of measured)' if peaks else 'fallback 6650 GB/s (of fallback)',
                'traffic': traffic, 'algorithmic_bytes_per_launch': bwd_bytes, 'kernel_ms': bwd_ms,
                'whole_step': {'algorithmic_bytes': fwd_bytes + bwd_bytes, 'achieved_GBps': (fwd_bytes + bwd_bytes) / (ms_per_step * 1e-3) / 1e9,
                               'frac_of_measured': (fwd_bytes + bwd_bytes) / (ms_per_step * 1e-3) / 1e9 / peak,
                               'frac_of_8TBps_nominal': (fwd_bytes + bwd_bytes) / (ms_per_step * 1e-3) / 1e9 / 8000.0},
                'note': 'all-pairs ALU/SFU-bound path: compulsory traffic is ~0.01 B per pixel*face (SURVEY 8d), so the HBM fraction is small by construction'}
    if traffic_note:
        roofline['traffic_note'] = traffic_note
    if prof.get('backward_warp_instructions') and clocks.get('sm_mhz'):
        # the roofline that actually binds: warp-instruction issue rate.  Instructions per launch from the committed
        # ncu capture of this workload and these sources (smsp__inst_executed.sum), duration and SM clock measured live in this run.
        peak_issue = 148 * 4 * clocks['sm_mhz'] * 1e6            # 148 SMs x 4 schedulers x 1 warp-instruction per clock
        ach = prof['backward_warp_instructions'] / (bwd_ms * 1e-3)
        roofline['issue_rate'] = {'bound': 'warp-instruction issue (FP32/SFU pipes)', 'achieved': ach / 1e9, 'peak': peak_issue / 1e9,
                                  'unit': 'G warp-instr/s', 'frac': ach / peak_issue,
                                  'active_lanes_per_instruction': prof.get('backward_active_lanes_per_instruction'),
                                  'source': 'instructions: profiles/roofline_traffic.json (ncu); time and SM clock: this run'}

    cfg = common_config(desc, B, world, F, S)
    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': warm,
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': cfg,
            'notes': {'parallelism': ('batch-sharded dp%d: the backward kernel accumulates the shared-mesh gradient into ONE [F,3,3] buffer per rank, '
                                      'one NCCL all-reduce of it per step' % world) if world > 1 else 'single GPU, per-item gradients [B,F,3,3]',
                      'l2': 'no flush: per-step working set (records+images+grads ~%d MB) exceeds the 126 MB L2' % ((B * F * 176 + B * S * S * 4 * 10 + B * F * 36 * 2) // 1000000),
                      'value_is': 'raw C-ABI calls on preallocated device buffers (gendr_forward_render + gendr_backward_render); value_public_api is the '
                                  'same step through gendr_b200.functional.render(...).backward()'},
            'value_public_api': {'value': pairs_per_step / (api_ms * 1e-3) / 1e6, 'unit': UNIT, 'ms_per_step': api_ms},
            'kernel_ms': {'forward': res['fwd_ms'], 'backward': bwd_ms, 'allreduce': res['allreduce_ms']},
            'e2e': e2e, 'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roofline}
    if world > 1:
        line['notes']['host_affinity'] = ('rank 0 bound to the %d CPUs local to its GPU (gendr_b200.parallel.bind_host_to_device); every rank does the same'
                                          % len(host_cpus)) if host_cpus else 'unchanged (no NUMA-local CPU list found, or it is the whole allowed set)'
    if equal is not None:
        line['sharded_equals_single'] = bool(equal['ok'])
        line['sharded_check'] = equal
    if c4 is not None:
        line['c4'] = c4

    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_arm(fv, ft, kw, steps=2, warmup=1)
        line['cpu_baseline'] = {k: r[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
    if world == 1 and not args.no_reference_cuda:
        try:
            from ref_gpu import load_reference, reference_render
            ref = load_reference()
            if ref is not None:
                nb = min(B, 16)
                f_in, t_in, g_in = fv[:nb].to(dev), ft[:nb].to(dev), gcol[:nb].contiguous()

                def ref_step():
                    a, b_ = f_in.clone().requires_grad_(True), t_in.clone().requires_grad_(True)
                    reference_render(ref, a, b_, **kw).backward(g_in)
                ref_step(); torch.cuda.synchronize()
                ts = []
                for _ in range(2):
                    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a0.record(); ref_step(); a1.record(); torch.cuda.synchronize()
                    ts.append(a0.elapsed_time(a1))
                ref_val = nb * S * S * F / (min(ts) * 1e-3) / 1e6
                line['reference_cuda'] = {'value': ref_val, 'unit': UNIT, 'ms_per_step': min(ts), 'batch': nb,
                                          'what': "the reference's own CUDA kernels (unmodified, built for sm_100a) on this GPU, same inputs, "
                                                  'gendr.functional.render + backward', 'speedup_device_path': value / ref_val}
        except Exception as e:      # the extra must never break the contract line
            line['reference_cuda'] = {'unavailable': repr(e)[:200]}
    emit(line)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == '__main__':
    main()
