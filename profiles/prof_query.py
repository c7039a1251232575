"""Profiling driver: one fine-pass sized network query (R rays x 192 samples) per precision, for ncu.
    ncu --set full --clock-control none --import-source on -k regex:mlp_tc -c 2 -o gpurun_out/prof python profiles/prof_query.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dfa_nerf_b200 as dfn  # noqa: E402
import synth  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
S = 192
modes = sys.argv[2].split(',') if len(sys.argv) > 2 else ['bf16', 'bf16x3']
dev = torch.device('cuda', 0)
net = dfn.FaceNeRF(D=8, W=256, input_ch=63, input_ch_views=27, dim_aud=64, output_ch=4, skips=[4], use_viewdirs=True)
net.load_state_dict(synth.facenerf_state_dict(1))
net = net.to(dev)
fr = synth.frame_inputs(H=450, W=450, seed=0)
ro, rd, vd = dfn.get_rays(450, 450, fr['focal'], fr['c2w'], fr['cx'], fr['cy'], device=dev, return_viewdirs=True)
ro, rd, vd = [t.reshape(-1, 3)[:R].contiguous() for t in (ro, rd, vd)]
z, _ = torch.sort(torch.rand(R, S, device=dev) * 0.6 + 0.4, -1)
aud = fr['aud'].to(dev)
for mode in modes:
    eng = dfn.RenderEngine(net, None, S, 0, precision={'bf16': dfn.PREC_BF16, 'bf16x3': dfn.PREC_BF16X3}[mode])
    for _ in range(2):
        raw = eng.query_points(net, ro, rd, vd, z, aud)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    raw = eng.query_points(net, ro, rd, vd, z, aud)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    print('%s: %d rays x %d: %.3f ms  -> %.1f TFLOP/s algorithmic' % (mode, R, S, ms, 2 * 557184 * R * S / ms / 1e9))
