"""Generates tests/golden/golden_v1.npz from the UNMODIFIED reference kernels run on the CPU through
oracle/ref_shim.h (oracle/_ref/libgendr_ref_cpu.so; needs /root/reference, i.e. the build container).
Run:  python tests/golden/make_golden.py
The vectors pin oracle/gendr_oracle.c (tests/test_oracle_cpu.py) wherever the reference itself is unavailable."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import scenes  # noqa: E402
from oracle.cpu_oracle import Oracle, build, make_params  # noqa: E402

CASES = []
for dist, dkw in scenes.DIST_SWEEP:
    CASES.append(dict(dist_func=dist, aggr_alpha_func='probabilistic', **dkw))
for tname, tp in scenes.TCN_SWEEP:
    CASES.append(dict(dist_func='logistic', aggr_alpha_func=tname, aggr_alpha_t_conorm_p=tp))
CASES += [dict(dist_func='gaussian', aggr_alpha_func='einstein', aggr_rgb_func='hard'),
          dict(dist_func='cauchy', aggr_alpha_func='yager', aggr_alpha_t_conorm_p=2.0, dist_squared=True, dist_scale=9e-4),
          dict(dist_func='uniform', aggr_alpha_func='probabilistic', double_side=False, dist_eps=4.0),
          dict(dist_func='logistic', aggr_alpha_func='probabilistic', texture_type='vertex')]


def main():
    build()
    ref = Oracle('reference')
    out = {}
    # (1) the reference's own 1-triangle scene + a culled sentinel face (SURVEY 8c known answers)
    fv1, ft1, _ = scenes.config_c1()
    fv1, ft1 = scenes.with_sentinel(fv1, ft1)
    ft1[:, 1] = 0.5
    # (2) a 60-face soup, 2 batch items, with slivers
    fv2, ft2 = scenes.soup(60, batch=2, seed=21, size=0.15)
    fv2[0, 5, :, 1] = fv2[0, 5, 0, 1] + (fv2[0, 5, :, 0] - fv2[0, 5, 0, 0]) * 1e-3      # near-degenerate sliver
    ftv = np.random.default_rng(5).random((2, fv2.shape[1], 3, 3)).astype(np.float32)
    out['c1_faces'], out['c1_textures'] = fv1.numpy(), ft1.numpy()
    out['soup_faces'], out['soup_textures'], out['soup_vertex_textures'] = fv2.numpy(), ft2.numpy(), ftv
    rng = np.random.default_rng(7)
    g1 = np.zeros((1, 4, 32, 32), np.float32); g1[:, 3] = 1
    g2 = rng.standard_normal((2, 4, 24, 24)).astype(np.float32)
    out['c1_grad'], out['soup_grad'] = g1, g2
    c1_cfgs = [dict(dist_func='uniform', aggr_alpha_func='probabilistic', dist_scale=.01, double_side=False),
               dict(dist_func='logistic', aggr_alpha_func='probabilistic', dist_scale=.01),
               dict(dist_func='gaussian', aggr_alpha_func='einstein', dist_scale=.03),
               dict(dist_func='cauchy', aggr_alpha_func='yager', aggr_alpha_t_conorm_p=2., dist_scale=.01),
               dict(dist_func='hard', aggr_alpha_func='hard', aggr_rgb_func='hard')]
    for i, kw in enumerate(c1_cfgs):
        p = make_params(image_size=32, **kw)
        f = ref.forward(out['c1_faces'], out['c1_textures'], p)
        gf, gt = ref.backward(f, g1, p)
        out['c1_%d_colors' % i], out['c1_%d_aggrs' % i], out['c1_%d_gfaces' % i], out['c1_%d_gtex' % i] = f['soft_colors'], f['aggrs_info'], gf, gt
    for i, kw in enumerate(CASES):
        kw = dict(dict(image_size=24, dist_scale=0.03), **kw)
        tex = out['soup_vertex_textures'] if kw.get('texture_type') == 'vertex' else out['soup_textures']
        p = make_params(**kw)
        f = ref.forward(out['soup_faces'], tex, p, background_color=(0.2, 0.4, 0.6))
        gf, gt = ref.backward(f, g2, p)
        out['soup_%d_colors' % i], out['soup_%d_aggrs' % i], out['soup_%d_gfaces' % i], out['soup_%d_gtex' % i] = f['soft_colors'], f['aggrs_info'], gf, gt
        if i == 0:
            out['soup_faces_info'] = f['faces_info']
    np.savez_compressed(os.path.join(HERE, 'golden_v1.npz'), **out)
    print('wrote golden_v1.npz:', len(out), 'arrays,', os.path.getsize(os.path.join(HERE, 'golden_v1.npz')) // 1024, 'KiB')


if __name__ == '__main__':
    main()
