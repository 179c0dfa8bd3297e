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
                     'config': {'workload': desc, 'per_gpu_batch': b, 'timed_on': 'host CPU cores: the reference CUDA kernels compiled for the host '
                                '(oracle/_ref) -- no GPU or no baseline/_ref build available'},
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
                 'config': {'workload': desc, 'per_gpu_batch': B, 'global_batch': B, 'faces': F, 'image_size': S,
                            'timed_on': "one B200: the reference's own CUDA kernels (unmodified sources, sm_100a build under baseline/_ref) through "
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

    import numpy as np
    import torch
    import torch.distributed as dist
    from gendr_b200 import _lib, parallel
    from gendr_b200.cuda import generalized_renderer as ext
    from gendr_b200.functional import renderer as fr

    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (the product has no CPU path); use --impl reference for the CPU arm')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.load()

    fv, ft, kw, desc, B = workload(args.workload, args.batch, rank, world)
    F, S, T = fv.shape[1], kw['image_size'], ft.shape[2]
    params = ext.make_params(S, fr.DIST_FUNC_IDS[kw['dist_func']], 1e-2, False, None, None, 1e4, fr.AGGR_ALPHA_FUNC_IDS[kw['aggr_alpha_func']],
                             kw.get('aggr_alpha_t_conorm_p'), 1, 1e-3, 1e-3, 1, 100, kw['double_side'], 0, (0, 0, 0))
    faces = fv.to(dev).view(B, F, 9).contiguous()
    tex = ft.to(dev).contiguous()
    gcol = torch.randn(B, 4, S, S, generator=torch.Generator().manual_seed(2)).to(dev)
    colors = torch.empty(B, 4, S, S, device=dev)
    aggrs = torch.empty(B, 2, S, S, device=dev)
    gfaces = torch.empty(B, F, 9, device=dev)
    gtex = torch.empty(B, F, T, 3, device=dev)
    ws = ext.workspace_for(faces)

    def step(ev=None):
        ext.forward_render_raw(faces, tex, None, aggrs, colors, params, False, ws)
        gfaces.zero_(); gtex.zero_()
        if ev:
            ev[0].record()
        ext.backward_render_raw(faces, tex, colors, aggrs, gfaces, gtex, gcol, params, ws, True, False)
        if ev:
            ev[1].record()
        if world > 1:      # shared mesh: gradient w.r.t. the shared geometry = batch sum + ONE all-reduce
            return parallel.allreduce_shared_face_grads(gfaces.view(B, F, 3, 3))
        return None

    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.gendr_launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    w_begin = time.perf_counter()
    t_begin.record()
    for i in range(args.steps):
        step(evs[i])
    t_end.record()
    torch.cuda.synchronize()
    sampler.window(w_begin, time.perf_counter())
    if world > 1:
        dist.barrier()
    launches = lib.gendr_launch_count() - launches0
    clocks = sampler.stop()
    elapsed_ms = t_begin.elapsed_time(t_end)
    bwd_ms = sum(a.elapsed_time(b) for a, b in evs) / args.steps
    if world > 1:
        tmax = torch.tensor([elapsed_ms], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tmax.item())
    ms_per_step = elapsed_ms / args.steps
    pairs_per_step = B * world * S * S * F
    value = pairs_per_step / (ms_per_step * 1e-3) / 1e6

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
    traffic, prof = None, {}
    try:
        prof = json.load(open(os.path.join(ROOT, 'profiles', 'roofline_traffic.json'))).get(args.workload, {})
        if B == 64:          # the ncu capture is of the default per-GPU batch
            traffic = prof.get('backward_dram_bytes_per_launch')
    except Exception:
        pass
    achieved = bwd_bytes / (bwd_ms * 1e-3) / 1e9
    roofline = {'bound': 'hbm', 'kernel': 'render_kernel<gaussian,simple,BWD> (backward)' if args.workload == 'c3' else 'render_kernel<...,BWD> (backward)',
                'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'peak_source': 'MEASURED_PEAKS.json hbm_gbs (of measured)' if peaks else 'fallback 6650 GB/s (of fallback)',
                'traffic': traffic, 'algorithmic_bytes_per_launch': bwd_bytes, 'kernel_ms': bwd_ms,
                'whole_step': {'algorithmic_bytes': fwd_bytes + bwd_bytes, 'achieved_GBps': (fwd_bytes + bwd_bytes) / (ms_per_step * 1e-3) / 1e9,
                               'frac_of_measured': (fwd_bytes + bwd_bytes) / (ms_per_step * 1e-3) / 1e9 / peak,
                               'frac_of_8TBps_nominal': (fwd_bytes + bwd_bytes) / (ms_per_step * 1e-3) / 1e9 / 8000.0},
                'note': 'all-pairs ALU/SFU-bound path: compulsory traffic is ~0.01 B per pixel*face (SURVEY 8d), so the HBM fraction is small by construction'}
    if prof.get('backward_warp_instructions') and B == 64 and clocks.get('sm_mhz'):
        # the roofline that actually binds: warp-instruction issue rate.  Instructions per launch from the committed
        # ncu capture of this workload (smsp__inst_executed.sum), duration and SM clock measured live in this run.
        peak_issue = 148 * 4 * clocks['sm_mhz'] * 1e6            # 148 SMs x 4 schedulers x 1 warp-instruction per clock
        ach = prof['backward_warp_instructions'] / (bwd_ms * 1e-3)
        roofline['issue_rate'] = {'bound': 'warp-instruction issue (FP32/SFU pipes)', 'achieved': ach / 1e9, 'peak': peak_issue / 1e9,
                                  'unit': 'G warp-instr/s', 'frac': ach / peak_issue,
                                  'active_lanes_per_instruction': prof.get('backward_active_lanes_per_instruction'),
                                  'source': 'instructions: profiles/roofline_traffic.json (ncu); time and SM clock: this run'}

    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup),
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': desc, 'per_gpu_batch': B, 'global_batch': B * world, 'faces': F, 'image_size': S, 'texture_res': 1,
                       'params': 'GenDR defaults (dist_scale 1e-2, dist_eps 1e4, softmax RGB, eps=gamma=1e-3, near 1, far 100, single-sided)',
                       'parallelism': 'batch-sharded dp%d, one all-reduce of the shared-mesh face gradient [F,3,3] per step' % world if world > 1 else 'single GPU',
                       'l2': 'no flush: per-step working set (records+images+grads ~%d MB) exceeds the 126 MB L2' % ((B * F * 144 + B * S * S * 4 * 10 + B * F * 36 * 2) // 1000000)},
            'e2e': e2e, 'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roofline}

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
