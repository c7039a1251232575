"""Head / torso field queries of the live Decoder (202,500 rays x 64 samples each, bf16) on the CTA-pair kernel (mlp_pair.cu, default) and on
the layer-program interpreter (mlp_pp.cu, debug flags): ms per launch.
    python profiles/bench_decoder_fields.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dfa_nerf_b200 as dfn  # noqa: E402
import synth  # noqa: E402

dev = torch.device('cuda', 0)
H = W = 450
S = 64
dec = dfn.Decoder(z_dim=256, hidden_size=256, dim_signal=96, use_deformation_field=True)
dec.load_state_dict(synth.decoder_state_dict(0))
dec = dec.to(dev)
fr = synth.frame_inputs(H=H, W=W, seed=0)
ro, rd = dfn.get_rays(H, W, fr['focal'], fr['c2w'], fr['cx'], fr['cy'], device=dev)
ro, rd = ro.reshape(-1, 3).contiguous(), rd.reshape(-1, 3).contiguous()
R = ro.shape[0]
z = (torch.linspace(0.4, 1.0, S, device=dev)[None, :].expand(R, S)).contiguous()
g = torch.Generator().manual_seed(0)
zs, za = torch.randn(1, 256, generator=g).to(dev), torch.randn(1, 256, generator=g).to(dev)
sig = {'head': torch.randn(1, 96, generator=g).to(dev), 'torso': torch.randn(1, 42, generator=g).to(dev)}
net = dfn.FaceNeRF(D=8, W=256, input_ch=63, input_ch_views=27, dim_aud=64, output_ch=4, skips=[4], use_viewdirs=True)
net.load_state_dict(synth.facenerf_state_dict(1))
net = net.to(dev)
eng = dfn.RenderEngine(net, None, 64, 0, precision=dfn.PREC_BF16)
for name, impl in (('mlp_pair.cu', -1), ('mlp_pp.cu', 3 + 16 * (3 + 8 + 16 + 1))):
    dfn.lib.dfn_debug_set_impl(impl)
    eng.query_points(net, ro[:64], rd[:64], rd[:64].contiguous(), z[:64].contiguous(), fr['aud'].to(dev))   # the switch takes effect on a query
    for which in ('head', 'torso'):
        for _ in range(2):
            dec.query_rays(ro, rd, z, zs, za, sig[which], which, precision=dfn.PREC_BF16)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            dec.query_rays(ro, rd, z, zs, za, sig[which], which, precision=dfn.PREC_BF16)
        e1.record()
        torch.cuda.synchronize()
        print('%-12s %-5s %.2f ms per 202,500 x 64 query' % (name, which, e0.elapsed_time(e1) / 5), flush=True)
dfn.lib.dfn_debug_set_impl(-1)
