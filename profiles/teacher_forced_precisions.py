"""Teacher-forced hierarchical FaceNeRF frame (the oracle's z_samples injected) per tensor-core precision:
max-abs error of the rendered maps against the reference's golden chunk.  python profiles/teacher_forced_precisions.py"""
import sys, torch
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import dfa_nerf_b200 as dfn
from conftest import load_golden
from oracle import nerf_oracle as O
import synth
DEV='cuda'
g = load_golden('render_rays')
def mk(s):
    m = dfn.FaceNeRF(D=8, W=256, input_ch=63, input_ch_views=27, dim_aud=64, output_ch=4, skips=[4], use_viewdirs=True)
    m.load_state_dict(synth.facenerf_state_dict(s)); return m.to(DEV)
nc, nf = mk(g['coarse_seed']), mk(g['fine_seed'])
ro, rd = O.get_rays(g['H'], g['W'], g['focal'], g['c2w'], g['cx'], g['cy'])
ro, rd = ro.reshape(-1,3).contiguous(), rd.reshape(-1,3)
vd = rd / torch.norm(rd, dim=-1, keepdim=True)
n = ro.shape[0]
near, far = torch.full((n,), g['near']), torch.full((n,), g['far'])
for name in ('PREC_BF16X3','PREC_FP16','PREC_BF16'):
    eng = dfn.RenderEngine(nc, nf, 64, 128, precision=getattr(dfn, name))
    out = eng.render_rays(ro.to(DEV), rd.to(DEV), vd.to(DEV), near.to(DEV), far.to(DEV), g['bc_rgb'].to(DEV), g['aud'].to(DEV),
                          z_samples=g['z_samples'].to(DEV), want=('rgb_map','acc_map','last_weight','rgb0'))
    e = lambda a,b: (a.cpu().double()-b.double()).abs().max().item()
    print(name, 'teacher-forced: rgb0 %.2e rgb_map %.2e acc %.2e last_weight %.2e' % (e(out['rgb0'],g['rgb0']), e(out['rgb_map'],g['rgb_map']), e(out['acc_map'],g['acc_map']), e(out['last_weight'],g['weights'][:,-1])))
