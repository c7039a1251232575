"""Golden vectors for the Decoder options the live script (MAIN:518) leaves at their defaults: listener layers (head, signal
None), use_expression, several skips, no view directions, no final sigmoid, other widths, torso without the deformation field.

Run in the build container only (needs /root/reference):
    python oracle/make_golden_decoder_options.py

Builds the reference's own decoder.Decoder for each case, runs its forward on seeded inputs, asserts the restatement
(oracle/nerf_oracle.decoder_forward) gives the same tensors bit-for-bit, and stores state_dict + inputs + the REFERENCE's
outputs in tests/golden/decoder_options.npz.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference/NeRFs/DFANeRF')
import decoder as DEC  # noqa: E402
from oracle import nerf_oracle as O  # noqa: E402

torch.autograd.set_detect_anomaly(False)

from oracle.decoder_option_cases import CASES  # noqa: E402


def main():
    out = {}
    P = 53
    for ci, (name, (kw, calls)) in enumerate(CASES.items()):
        torch.manual_seed(100 + ci)
        m = DEC.Decoder(**kw).eval()
        with torch.no_grad():            # default init leaves the outputs nearly constant; widen the weights
            for prm in m.parameters():
                prm.mul_(2.5)
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        for k, v in sd.items():
            out['%s/sd/%s' % (name, k)] = v.numpy()
        g = torch.Generator().manual_seed(7 + ci)
        p = (torch.rand(1, P, 3, generator=g) * 2 - 1) * 0.8
        rd = torch.randn(1, P, 3, generator=g)
        zs = torch.randn(1, kw['z_dim'], generator=g)
        za = torch.randn(1, kw['z_dim'], generator=g)
        sig_h = torch.randn(1, kw['dim_signal'], generator=g)
        sig_t = torch.randn(1, kw['dim_et_embed'], generator=g)
        ex = torch.randn(1, kw.get('dim_exp', 256), generator=g)
        for k, v in dict(p=p, ray_d=rd, z_shape=zs, z_app=za, signal=sig_h, signal_torso=sig_t, expression=ex).items():
            out['%s/in/%s' % (name, k)] = v.numpy()
        for call, which, has_sig, has_ex, has_rd in calls:
            sig = (sig_h if which == 'head' else sig_t) if has_sig else None
            arg = [sig, ex if has_ex else None] if which == 'head' else sig
            with torch.no_grad():
                feat, sigma = m(p, rd if has_rd else None, zs, za, arg, which)
                of, osg = O.decoder_forward(sd, p, rd if has_rd else None, zs, za, sig, which, n_freq=kw.get('n_freq_posenc', 10),
                                            n_freq_views=kw.get('n_freq_posenc_views', 4), skips=tuple(kw['skips']),
                                            n_blocks=kw['n_blocks'], expression=ex if has_ex and kw.get('use_expression') else None,
                                            use_deformation_field=kw['use_deformation_field'],
                                            final_sigmoid=kw.get('final_sigmoid_activation', True))
            assert torch.equal(of, feat) and torch.equal(osg, sigma), (name, call)
            print('  ok  %-6s %-16s feat %s sigma %s  (feat std %.3f, sigma std %.3f)' %
                  (name, call, tuple(feat.shape), tuple(sigma.shape), feat.std().item(), sigma.std().item()))
            out['%s/out/%s/feat' % (name, call)] = feat.numpy()
            out['%s/out/%s/sigma' % (name, call)] = sigma.numpy()
    path = os.path.join(ROOT, 'tests', 'golden', 'decoder_options.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
