#!/usr/bin/env python
"""Forward / backward kernel times of the hot path on named workloads (device-resident inputs, CUDA events, median of n).
    python tools/gpu_time_kernels.py [c3:64 c4:16 c2:16 ...] [--n 5] [--tag name]
Environment: GENDR_B200_LIB (alternative build of the library), GENDR_B200_BWD=ps (pixel-stationary backward)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch  # noqa: E402

import scenes  # noqa: E402
from gendr_b200.cuda import generalized_renderer as ext  # noqa: E402
from gendr_b200.functional import renderer as fr  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if ':' in a and not a.startswith('--') and a.split(':')[0] in ('c2', 'c3', 'c4')]
    n = int(sys.argv[sys.argv.index('--n') + 1]) if '--n' in sys.argv else 5
    tag = sys.argv[sys.argv.index('--tag') + 1] if '--tag' in sys.argv else os.environ.get('GENDR_B200_LIB', 'default').split('/')[-1]
    dev = torch.device('cuda:0')
    out = {'tag': tag, 'bwd': os.environ.get('GENDR_B200_BWD', 'fs')}
    for spec in (args or ['c3:64', 'c4:16']):
        name, B = spec.split(':'); B = int(B)
        fv, ft, kw = {'c2': scenes.config_c2, 'c3': scenes.config_c3, 'c4': scenes.config_c4}[name](batch=B)
        F, S, T = fv.shape[1], kw['image_size'], ft.shape[2]
        params = ext.make_params(S, fr.DIST_FUNC_IDS[kw['dist_func']], 1e-2, False, None, None, 1e4, fr.AGGR_ALPHA_FUNC_IDS[kw['aggr_alpha_func']],
                                 kw.get('aggr_alpha_t_conorm_p'), 1, 1e-3, 1e-3, 1, 100, False, 0, (0, 0, 0))
        faces, tex = fv.to(dev).view(B, F, 9).contiguous(), ft.to(dev).contiguous()
        gcol = torch.randn(B, 4, S, S, generator=torch.Generator().manual_seed(2)).to(dev)
        colors, aggrs = torch.empty(B, 4, S, S, device=dev), torch.empty(B, 2, S, S, device=dev)
        gfaces, gtex = torch.empty(B, F, 9, device=dev), torch.empty(B, F, T, 3, device=dev)
        ws = ext.workspace_for(faces)
        tf, tb = [], []
        for i in range(n + 2):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            ext.forward_render_raw(faces, tex, None, aggrs, colors, params, False, ws)
            e[1].record()
            ext.backward_render_raw(faces, tex, colors, aggrs, gfaces, gtex, gcol, params, ws, True, True)
            e[2].record()
            torch.cuda.synchronize()
            if i >= 2:
                tf.append(e[0].elapsed_time(e[1])); tb.append(e[1].elapsed_time(e[2]))
        tf.sort(); tb.sort()
        out[spec] = {'fwd_ms': round(tf[len(tf) // 2], 3), 'bwd_ms': round(tb[len(tb) // 2], 3), 'sum_ms': round(tf[len(tf) // 2] + tb[len(tb) // 2], 3),
                     'gf_sum': float(gfaces.double().sum()), 'gf_abs': float(gfaces.double().abs().sum()), 'img_sum': float(colors.double().sum())}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
