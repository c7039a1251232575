"""Output-side overlap check: the head+torso frame loop at 450x450 x 64 with and without the JPEG writer.
python profiles/bench_render_person.py [frames]   -> one line per variant (ms/frame, wall clock around the whole loop)"""
import os
import sys
import tempfile
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dfa_nerf_b200 as dfn  # noqa: E402
import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = torch.device('cuda', 0)
H = W = 450
fr = synth.frame_inputs(H=H, W=W, seed=0)
dec = dfn.Decoder(z_dim=256, hidden_size=256, dim_signal=96, use_deformation_field=True)
dec.load_state_dict(synth.decoder_state_dict(0))
dec = dec.to(dev)
a, e = dfn.AudioNet_W2L().to(dev), dfn.ExpressionEnc().to(dev)
a.load_state_dict(synth.mlp_encoder_state_dict(1))
e.load_state_dict(synth.mlp_encoder_state_dict(2, (64, 32, 32)))
g = torch.Generator().manual_seed(0)
data = {'poses': synth.pose_sequence(n, 3).to(dev), 'auds': torch.randn(n, 512, generator=g).to(dev),
        'exp': torch.randn(n, 64, generator=g).to(dev), 'bc_img': fr['bc_rgb'].reshape(H, W, 3).to(dev),
        'hwfcxy': [H, W, fr['focal'], fr['cx'], fr['cy']]}
body = synth.camera_pose(31)
zs, za = torch.randn(1, 2, 256, generator=g).to(dev), torch.randn(1, 2, 256, generator=g).to(dev)


def loop(write, out_dir):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if write:
        dfn.render_person(dec, data, body, zs, za, a, e, out_dir, fr['near'], fr['far'], precision=dfn.PREC_BF16)
    else:
        sig = dfn.encode_signal_sequence(data['auds'], data['exp'], a, e)
        sig_t = dfn.encode_signal_torso_sequence(data['poses'])
        dfn.render_sequence_head_torso(dec, H, W, fr['focal'], data['poses'], body.to(dev), data['bc_img'], zs, za, sig, sig_t,
                                       fr['near'], fr['far'], fr['cx'], fr['cy'], precision=dfn.PREC_BF16, with_head=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / n


with tempfile.TemporaryDirectory() as d:
    loop(True, d)                                      # warm-up (workspaces, thread pool, Pillow import)
    for name, w in (('frames to pinned host memory only', False), ('+ JPEG files (render_com + render_head)', True)):
        ms = min(loop(w, d) for _ in range(2))
        print('%-44s %7.2f ms/frame  (%d frames, bf16, 450x450 x 64, head + torso)' % (name, ms, n))
    print('files written:', len(os.listdir(os.path.join(d, 'render_com'))), '+', len(os.listdir(os.path.join(d, 'render_head'))))
