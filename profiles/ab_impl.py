"""A/B of the tcgen05 kernel generations on the same query: max |diff| against impl 1 and timing.
    python profiles/ab_impl.py [R] [impls, e.g. 1,2,3]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dfa_nerf_b200 as dfn  # noqa: E402
import synth  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
impls = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [1, 2, 3]
S = 192
NAMES = {0: 'ts', 1: 'tc', 2: 'pp', 3: 'pair-ew8', 7: 'tc2', 8: 'pair-ew4', 4: 'tc-unicast', 5: 'tc-unicast-cd', 6: 'tc-cd'}
dev = torch.device('cuda', 0)
net = dfn.FaceNeRF(D=8, W=256, input_ch=63, input_ch_views=27, dim_aud=64, output_ch=4, skips=[4], use_viewdirs=True)
net.load_state_dict(synth.facenerf_state_dict(1))
net = net.to(dev)
fr = synth.frame_inputs(H=450, W=450, seed=0)
ro, rd, vd = dfn.get_rays(450, 450, fr['focal'], fr['c2w'], fr['cx'], fr['cy'], device=dev, return_viewdirs=True)
ro, rd, vd = [t.reshape(-1, 3)[:R].contiguous() for t in (ro, rd, vd)]
z, _ = torch.sort(torch.rand(R, S, device=dev) * 0.6 + 0.4, -1)
aud = fr['aud'].to(dev)
for mode, prec in (('bf16', dfn.PREC_BF16), ('bf16x3', dfn.PREC_BF16X3)):
    eng = dfn.RenderEngine(net, None, S, 0, precision=prec)
    outs = {}
    for impl in impls:
        if (impl & 15) >= 3 and mode != 'bf16':
            continue
        dfn.lib.dfn_debug_set_impl(impl)
        for _ in range(2):
            raw = eng.query_points(net, ro, rd, vd, z, aud)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        raw = eng.query_points(net, ro, rd, vd, z, aud)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        outs[impl] = raw.clone()
        msg = ''
        if impls[0] in outs and impl != impls[0]:
            d = (outs[impl] - outs[impls[0]]).abs()
            msg = '  max|diff vs %s| rgb %.3e sigma %.3e finite=%s' % (NAMES.get(impls[0], str(impls[0])), d[..., :3].max().item(), d[..., 3].max().item(),
                                                                      bool(torch.isfinite(outs[impl]).all()))
        print('%s impl=%s: %.3f ms -> %.1f TFLOP/s%s' % (mode, NAMES.get(impl, 'impl%d+flags%d' % (impl & 15, impl >> 4)), ms, 2 * 557184 * R * S / ms / 1e9, msg), flush=True)
dfn.lib.dfn_debug_set_impl(-1)
