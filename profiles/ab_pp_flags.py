"""A/B of the split-precision kernel's schedule switches (dfn_debug_set_pp_flags) on one FaceNeRF query and on the Decoder's head /
torso fields: bit-level agreement with flags = 0 and timing.
    python profiles/ab_pp_flags.py [R] [flag sets, e.g. 0,1]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dfa_nerf_b200 as dfn  # noqa: E402
import synth  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
flag_sets = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [0, 3, 7, 15]
S = 192
dev = torch.device('cuda', 0)
net = dfn.FaceNeRF(D=8, W=256, input_ch=63, input_ch_views=27, dim_aud=64, output_ch=4, skips=[4], use_viewdirs=True)
net.load_state_dict(synth.facenerf_state_dict(1))
net = net.to(dev)
fr = synth.frame_inputs(H=450, W=450, seed=0)
ro, rd, vd = dfn.get_rays(450, 450, fr['focal'], fr['c2w'], fr['cx'], fr['cy'], device=dev, return_viewdirs=True)
ro, rd, vd = [t.reshape(-1, 3)[:R].contiguous() for t in (ro, rd, vd)]
z, _ = torch.sort(torch.rand(R, S, device=dev) * 0.6 + 0.4, -1)
aud = fr['aud'].to(dev)


N_TIMED = int(os.environ.get('AB_REPEATS', 3))
CLOCKS = os.environ.get('AB_CLOCKS') == '1'      # sample nvidia-smi SM clocks / power during the timed loop (bench.ClockSampler)


def timed(fn, n=None):
    n = n or N_TIMED
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    smp = None
    if CLOCKS:
        import bench
        smp = bench.ClockSampler(0)
        smp.start()
        import time
        time.sleep(0.2)
        smp.mark()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(n):
        out = fn()
    ev1.record()
    torch.cuda.synchronize()
    if smp is not None:
        c = smp.stop()
        pw = sorted(float(r[2]) for t, r in smp.rows if t >= smp.t0 and len(r) > 2 and r[2].replace('.', '').isdigit())
        print('   clocks: %s MHz median, power %.0f W median, %s' % (c['sm_mhz'], pw[len(pw) // 2] if pw else -1, c['reasons']), flush=True)
    return out, ev0.elapsed_time(ev1) / n


try:
    for mode, prec in (('bf16x3', dfn.PREC_BF16X3), ('fp16x3m', dfn.PREC_FP16X3M)):
        eng = dfn.RenderEngine(net, None, S, 0, precision=prec)
        outs = {}
        for f in flag_sets:
            dfn.lib.dfn_debug_set_pp_flags(f)
            raw, ms = timed(lambda: eng.query_points(net, ro, rd, vd, z, aud))
            outs[f] = raw.clone()
            msg = ''
            if f != flag_sets[0]:
                d = (outs[f] - outs[flag_sets[0]]).abs()
                msg = '  bit-identical to flags %d: %s  (max |d rgb_raw| %.2e, |d sigma| %.2e)  finite=%s' % (
                    flag_sets[0], bool(torch.equal(outs[f], outs[flag_sets[0]])), d[..., :3].max().item(), d[..., 3].max().item(),
                    bool(torch.isfinite(outs[f]).all()))
            print('FaceNeRF %s flags=%d: %.3f ms -> %.1f TFLOP/s%s' % (mode, f, ms, 2 * 557184 * R * S / ms / 1e9, msg), flush=True)
    # Decoder fields (bf16x3 on mlp_pp.cu)
    dec = dfn.Decoder(z_dim=256, hidden_size=256, dim_signal=96, use_deformation_field=True)
    dec.load_state_dict(synth.decoder_state_dict(0))
    dec = dec.to(dev)
    g = torch.Generator().manual_seed(0)
    zs, za = torch.randn(1, 2, 256, generator=g).to(dev), torch.randn(1, 2, 256, generator=g).to(dev)
    sig, sig_t = torch.randn(1, 96, generator=g).to(dev), torch.randn(1, 42, generator=g).to(dev)
    c2w_t = synth.frame_inputs(H=450, W=450, seed=7)['c2w']
    outs = {}
    for f in flag_sets:
        dfn.lib.dfn_debug_set_pp_flags(f)
        (head, person), ms = timed(lambda: dfn.render_head_torso(dec, 450, 450, fr['focal'], fr['c2w'], c2w_t, fr['bc_rgb'].to(dev), zs, za, sig, sig_t,
                                                                 fr['near'], fr['far'], fr['cx'], fr['cy'], N_samples=64, ray_range=(0, min(R * 3, 450 * 450)),
                                                                 precision=dfn.PREC_BF16X3))
        outs[f] = (head.clone(), person.clone())
        msg = ''
        if f != flag_sets[0]:
            msg = '  bit-identical to flags %d: %s (max |d person| %.2e)' % (
                flag_sets[0], bool(torch.equal(outs[f][0], outs[flag_sets[0]][0]) and torch.equal(outs[f][1], outs[flag_sets[0]][1])),
                (outs[f][1] - outs[flag_sets[0]][1]).abs().max().item())
        print('Decoder head+torso bf16x3 (%d rays x 64) flags=%d: %.3f ms%s' % (min(R * 3, 450 * 450), f, ms, msg), flush=True)
finally:
    dfn.lib.dfn_debug_set_pp_flags(15)
