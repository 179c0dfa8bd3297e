"""A/B equality of two builds of libgendr_b200.so: forward images must be bit-identical, gradients equal to atomic-order noise.
    python tools/gpu_ab_equal.py <libA.so> <libB.so>         (spawns one process per library; workloads: C2, C3 slice, C4 slice,
    small dist_eps, logistic wide, squared distances, hard RGB, soup with slivers)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker(out_path):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch
    import scenes
    import gendr_b200 as gd
    dev = torch.device('cuda:0')
    cases = []
    fv, ft, cfg = scenes.config_c3(batch=8); cases.append(('c3', fv, ft, dict(cfg, double_side=False)))
    fv, ft, cfg = scenes.config_c2(batch=8); cases.append(('c2', fv, ft, dict(cfg, double_side=False)))
    fv, ft, cfg = scenes.config_c4(batch=2); cases.append(('c4', fv, ft, dict(cfg, double_side=True)))
    fv, ft, cfg = scenes.config_c3(batch=4)
    cases.append(('c3 small dist_eps', fv, ft, dict(cfg, dist_eps=4.0, dist_scale=3e-3)))
    cases.append(('c3 logistic wide', fv, ft, dict(cfg, dist_func='logistic', aggr_alpha_func='probabilistic', dist_scale=3e-2)))
    cases.append(('c3 squared uniform hard', fv, ft, dict(cfg, dist_func='uniform', aggr_alpha_func='max', dist_squared=True, dist_scale=1e-3, aggr_rgb_func='hard')))
    cases.append(('c3 128px', fv, ft, dict(cfg, image_size=100)))
    fv, ft = scenes.soup(3000, batch=3, seed=9, size=0.05)
    fv[0, 5, :, 1] = fv[0, 5, 0, 1] + (fv[0, 5, :, 0] - fv[0, 5, 0, 0]) * 1e-3
    cases.append(('soup', fv, ft, dict(image_size=200, dist_func='gaussian', aggr_alpha_func='einstein', dist_scale=5e-3)))
    # randomised configurations: every distribution, random scales / eps / image sizes / t-conorms, soups with slivers,
    # big and off-screen faces
    import random
    rnd = random.Random(7)
    dists = [d for d, _ in scenes.DIST_SWEEP]
    dkws = dict(scenes.DIST_SWEEP)
    for i in range(int(os.environ.get('AB_RANDOM_CASES', '24'))):
        nf = rnd.choice([50, 400, 2000])
        fv, ft = scenes.soup(nf, batch=2, seed=100 + i, size=rnd.choice([0.02, 0.08, 0.3, 1.5]))
        fv[0, 3, :, 1] = fv[0, 3, 0, 1] + (fv[0, 3, :, 0] - fv[0, 3, 0, 0]) * 10 ** rnd.uniform(-5, -2)      # sliver
        fv[1, 7, :, :2] += 1.5                                                                              # partly off screen
        dist = rnd.choice(dists)
        tname, tp = rnd.choice(scenes.TCN_SWEEP)
        kw = dict(image_size=rnd.choice([33, 64, 100, 129, 256]), dist_func=dist, aggr_alpha_func=tname, aggr_alpha_t_conorm_p=tp,
                  dist_scale=10 ** rnd.uniform(-3, -1.3), dist_eps=10 ** rnd.uniform(0.3, 4), dist_squared=rnd.random() < 0.2,
                  aggr_rgb_func=rnd.choice(['softmax', 'softmax', 'hard']), double_side=rnd.random() < 0.5, **dkws[dist])
        cases.append(('random %d %s/%s S=%d' % (i, dist, tname, kw['image_size']), fv, ft, kw))
    out = {}
    for name, fv, ft, kw in cases:
        a, b = fv.to(dev).requires_grad_(True), ft.to(dev).requires_grad_(True)
        img = gd.functional.render(a, b, **kw)
        g = torch.randn(img.shape, generator=torch.Generator().manual_seed(1)).to(dev)
        img.backward(g)
        out[name] = (img.detach().cpu(), a.grad.cpu(), b.grad.cpu())
    torch.save(out, out_path)


if __name__ == '__main__':
    if sys.argv[1] == '--worker':
        worker(sys.argv[2])
        sys.exit(0)
    import torch
    res = []
    for i, lib in enumerate(sys.argv[1:3]):
        path = '/tmp/ab_%d.pt' % i
        subprocess.run([sys.executable, __file__, '--worker', path], check=True, env=dict(os.environ, GENDR_B200_LIB=os.path.abspath(lib)))
        res.append(torch.load(path))
    ok = True
    for name in res[0]:
        (i0, g0, t0), (i1, g1, t1) = res[0][name], res[1][name]
        same = bool(torch.equal(i0, i1)) or bool(((i0 == i1) | (torch.isnan(i0) & torch.isnan(i1))).all())

        def rel(a, b):
            if not bool((torch.isnan(a) == torch.isnan(b)).all()):
                return float('inf')
            m = torch.isfinite(a) & torch.isfinite(b)
            if not bool(m.any()):
                return 0.0
            return float((a[m] - b[m]).abs().max() / a[m].abs().max().clamp_min(1e-30))
        rg, rt = rel(g0, g1), rel(t0, t1)
        print('%-44s images bit-identical: %s   grad_faces rel diff %.2e   grad_textures rel diff %.2e' % (name, same, rg, rt))
        ok = ok and same and rg < 1e-3 and rt < 1e-3
    print('A/B EQUAL' if ok else 'A/B MISMATCH')
    sys.exit(0 if ok else 1)
