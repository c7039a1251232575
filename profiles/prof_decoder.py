"""Profiling driver for the live model: one head and one torso field query (R rays x 64 samples) per precision, for ncu.
    ncu --set full --clock-control none --import-source on -k regex:mlp_pp -c 2 -o gpurun_out/prof python profiles/prof_decoder.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dfa_nerf_b200 as dfn  # noqa: E402
import synth  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
mode = sys.argv[2] if len(sys.argv) > 2 else 'bf16'
S = 64
dev = torch.device('cuda', 0)
dec = dfn.Decoder(z_dim=256, hidden_size=256, dim_signal=96, use_deformation_field=True)
dec.load_state_dict(synth.decoder_state_dict(0))
dec = dec.to(dev)
fr = synth.frame_inputs(H=450, W=450, seed=0)
ro, rd = dfn.get_rays(450, 450, fr['focal'], fr['c2w'], fr['cx'], fr['cy'], device=dev)
ro, rd = [t.reshape(-1, 3)[:R].contiguous() for t in (ro, rd)]
z = dfn.z_vals_uniform(torch.full((R,), 0.4, device=dev), torch.full((R,), 1.0, device=dev), S)
g = torch.Generator().manual_seed(0)
zs, za = torch.randn(1, 256, generator=g).to(dev), torch.randn(1, 256, generator=g).to(dev)
sig = {'head': torch.randn(1, 96, generator=g).to(dev), 'torso': torch.randn(1, 42, generator=g).to(dev)}
prec = {'bf16': dfn.PREC_BF16, 'bf16x3': dfn.PREC_BF16X3}[mode]
h = dec.dfn_handle(dev)
for which in ('head', 'torso'):
    for _ in range(2):
        dec.query_rays(ro, rd, z, zs, za, sig[which], which, precision=prec)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    dec.query_rays(ro, rd, z, zs, za, sig[which], which, precision=prec)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    macs = dfn.lib.dfn_decoder_macs_per_sample(h, 0 if which == 'head' else 1)
    print('%s %s: %d rays x %d: %.3f ms -> %.1f TFLOP/s algorithmic' % (mode, which, R, S, ms, 2 * macs * R * S / ms / 1e9))
