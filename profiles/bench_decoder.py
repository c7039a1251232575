"""Timing of the reference's live two-field frame (config 3 of BASELINE.json: head + torso Decoder, 450x450 x 64
samples, MAIN:633-708) through dfn_render_head_torso, per precision; algorithmic TFLOP/s with the folded MAC counts.
    python profiles/bench_decoder.py [H] [steps]
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dfa_nerf_b200 as dfn  # noqa: E402
import synth  # noqa: E402

H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 450
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
S = 64
dev = torch.device('cuda', 0)
dec = dfn.Decoder(z_dim=256, hidden_size=256, dim_signal=96, use_deformation_field=True)
dec.load_state_dict(synth.decoder_state_dict(0))
dec = dec.to(dev)
fr = synth.frame_inputs(H=H, W=W, seed=0)
fr_t = synth.frame_inputs(H=H, W=W, seed=7)
g = torch.Generator().manual_seed(0)
zs, za = torch.randn(1, 2, 256, generator=g).to(dev), torch.randn(1, 2, 256, generator=g).to(dev)
sig, sig_t = torch.randn(1, 96, generator=g).to(dev), torch.randn(1, 42, generator=g).to(dev)
bc = fr['bc_rgb'].to(dev)
for name, prec in (('bf16', dfn.PREC_BF16), ('bf16x3', dfn.PREC_BF16X3)):
    def step():
        return dfn.render_head_torso(dec, H, W, fr['focal'], fr['c2w'], fr_t['c2w'], bc, zs, za, sig, sig_t, fr['near'], fr['far'],
                                     fr['cx'], fr['cy'], N_samples=S, precision=prec)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    dfn.lib.dfn_profile_enable(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        rh, rp = step()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    k_ms, k_n, k_macs = C.c_double(), C.c_int64(), C.c_double()
    dfn.lib.dfn_profile_collect(C.byref(k_ms), C.byref(k_n), C.byref(k_macs))
    dfn.lib.dfn_profile_enable(0)
    h = dec.dfn_handle(dev)
    print('%s: %dx%d x %d samples, head+torso: %.2f ms/frame = %.3f M rays/s; MLP kernels %.2f ms (%d launches), '
          '%.1f TFLOP/s algorithmic (folded MACs/sample head %.0f torso %.0f); finite=%s' % (
              name, H, W, S, ms, H * W / ms / 1e3, k_ms.value / steps, k_n.value, 2 * k_macs.value / (k_ms.value * 1e-3) / 1e12,
              dfn.lib.dfn_decoder_macs_per_sample(h, 0), dfn.lib.dfn_decoder_macs_per_sample(h, 1),
              bool(torch.isfinite(rp).all())), flush=True)
