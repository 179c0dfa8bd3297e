"""How many faces carry the certified-division flag (FLAG_FASTDIV, word31 bit 14 of the face record) in C2/C3/C4."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch, scenes
from gendr_b200.cuda import generalized_renderer as ext
from gendr_b200.functional import renderer as fr
dev = torch.device('cuda:0')
for name, B in (('c2', 16), ('c3', 8), ('c4', 2)):
    fv, ft, kw = {'c2': scenes.config_c2, 'c3': scenes.config_c3, 'c4': scenes.config_c4}[name](batch=B)
    F, S, T = fv.shape[1], kw['image_size'], ft.shape[2]
    params = ext.make_params(S, fr.DIST_FUNC_IDS[kw['dist_func']], 1e-2, False, None, None, 1e4, fr.AGGR_ALPHA_FUNC_IDS[kw['aggr_alpha_func']],
                             kw.get('aggr_alpha_t_conorm_p'), 1, 1e-3, 1e-3, 1, 100, False, 0, (0, 0, 0))
    faces, tex = fv.to(dev).view(B, F, 9).contiguous(), ft.to(dev).contiguous()
    colors, aggrs = torch.empty(B, 4, S, S, device=dev), torch.empty(B, 2, S, S, device=dev)
    ws = ext.workspace_for(faces)
    ext.forward_render_raw(faces, tex, None, aggrs, colors, params, False, ws)
    torch.cuda.synchronize()
    rec = ws[:B * F * 176].view(torch.int32).view(B * F, 44)
    flag = (rec[:, 31] >> 14) & 1
    print(name, 'faces', B * F, 'certified', int(flag.sum()), 'fraction %.4f' % float(flag.float().mean()), flush=True)
    rcull = ws[:B * F * 176].view(torch.float32).view(B * F, 44)[:, 35]
    unc = torch.isinf(rcull)
    pk = rec[:, 30:32]
    ix0, ix1 = pk[:, 0] & 0x3fff, (pk[:, 0] >> 16) & 0x3fff
    iy0, iy1 = pk[:, 1] & 0x3fff, (pk[:, 1] >> 16) & 0x3fff
    area = ((ix1 - ix0 + 1).clamp(min=0) * (iy1 - iy0 + 1).clamp(min=0)).double()
    print('   uncullable faces', int(unc.sum()), 'rect pixels: all %.3e  uncullable %.3e (%.1f%%)' % (float(area.sum()), float(area[unc].sum()),
          100 * float(area[unc].sum()) / float(area.sum())), 'mean rcull (finite) %.4f' % float(rcull[~unc].mean()), flush=True)
