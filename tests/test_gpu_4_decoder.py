"""GPU parity of the reference's LIVE model path (what scripts/test_obama.sh runs): Decoder + DeformationField_ori
(DEC:77-349) built from dfn_embed / dfn_linear, and the two-field compositing of MAIN:669-708
(dfn_composite_head_torso), against the vectors the reference itself produced (tests/golden)."""
import pytest
import torch

from oracle import nerf_oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def maxerr(a, b):
    return (a.detach().cpu().double() - b.detach().cpu().double()).abs().max().item()


@pytest.fixture(scope='module')
def dfn():
    import dfa_nerf_b200
    return dfa_nerf_b200


def make_decoder(dfn, seed):
    m = dfn.Decoder(z_dim=256, hidden_size=256, dim_signal=96, use_deformation_field=True)
    m.load_state_dict(synth.decoder_state_dict(seed))
    return m.to(DEV)


def test_linear_building_block(dfn):
    from dfa_nerf_b200.decoder import _linear
    g = torch.Generator().manual_seed(0)
    P, N, K1, K2 = 77, 50, 33, 9
    x1, x2 = torch.randn(P, K1, generator=g), torch.randn(1, K2, generator=g)
    w, b, a = torch.randn(N, K1 + K2, generator=g) * 0.2, torch.randn(N, generator=g), torch.randn(P, N, generator=g)
    ref = torch.relu(torch.cat([x1, x2.expand(P, -1)], -1) @ w.t() + b) + a
    y = _linear(P, N, x1.to(DEV), K1, K1, w.to(DEV), b.to(DEV), 1, X2=x2.to(DEV), ld2=0, K2=K2, addend=a.to(DEV), ld_add=N)
    assert maxerr(y, ref) < 1e-5
    ref2 = torch.sigmoid(x1 @ w[:, :K1].t() + a)           # no bias, addend before the activation
    y2 = _linear(P, N, x1.to(DEV), K1, K1, w[:, :K1].contiguous().to(DEV), None, 2 | 4, addend=a.to(DEV), ld_add=N)
    assert maxerr(y2, ref2) < 1e-6


def test_make_points_bit_exact(dfn):
    ro, rd, z = torch.randn(9, 3), torch.randn(9, 3), torch.rand(9, 64)
    p, r = dfn.make_points(ro.to(DEV), rd.to(DEV), z.to(DEV))
    assert torch.equal(p.cpu(), ro[:, None, :] + rd[:, None, :] * z[:, :, None])
    assert torch.equal(r.cpu(), rd[:, None, :].expand(9, 64, 3))


def test_decoder_forward_golden(dfn, golden):
    g = golden('decoder')
    dec = make_decoder(dfn, g['seed'])
    zs, za = g['z_shape'].to(DEV), g['z_app'].to(DEV)
    fh, sh = dec(g['p'].to(DEV), g['ray_d'].to(DEV), zs[:, 0], za[:, 0], [g['signal'].to(DEV), None], 'head')
    ft, st = dec(g['p'].to(DEV), g['ray_d'].to(DEV), zs[:, 1], za[:, 1], g['signal_torso'].to(DEV), 'torso')
    assert fh.shape == (1, 40, 3) and sh.shape == (1, 40)
    assert maxerr(fh, g['feat_head']) < 1e-5 and maxerr(ft, g['feat_torso']) < 1e-5
    # sigma carries the x400 density gain of the synthetic head: 2e-4 absolute is ~1e-6 relative
    assert maxerr(sh, g['sigma_head']) < 2e-4 and maxerr(st, g['sigma_torso']) < 2e-4


def test_decoder_forward_vs_oracle_ragged(dfn):
    sd = synth.decoder_state_dict(3)
    dec = make_decoder(dfn, 3)
    g = torch.Generator().manual_seed(3)
    for P in (1, 129, 2000):
        p = (torch.rand(1, P, 3, generator=g) * 2 - 1) * 0.7
        rd = torch.randn(1, P, 3, generator=g)
        zs, za = torch.randn(1, 256, generator=g), torch.randn(1, 256, generator=g)
        sig = torch.randn(1, 42, generator=g)
        with torch.no_grad():
            rf, rs = O.decoder_forward(sd, p, rd, zs, za, sig, 'torso')
        f, s = dec(p.to(DEV), rd.to(DEV), zs.to(DEV), za.to(DEV), sig.to(DEV), 'torso')
        assert maxerr(f, rf) < 1e-5 and maxerr(s, rs) < 5e-4


def test_head_torso_frame_golden(dfn, golden):
    """The whole live chunk, MAIN:633-708: rays, points, both fields, background splice, mix, weights, colour."""
    g = golden('head_torso')
    dec = make_decoder(dfn, g['seed'])
    rgb_head, rgb_person = dfn.render_head_torso(
        dec, g['H'], g['W'], g['focal'], g['c2w'], g['c2w_torso'], g['bc_rgb'].to(DEV), g['z_shape'].to(DEV),
        g['z_app'].to(DEV), g['signal'].to(DEV), g['signal_torso'].to(DEV), g['near'], g['far'], g['cx'], g['cy'],
        precision=dfn.PREC_FP32)
    assert maxerr(rgb_head, g['rgb_head']) < 1e-5
    assert maxerr(rgb_person, g['rgb_person']) < 1e-5


def test_composite_head_torso_edge_cases(dfn):
    """Empty head (the last-sample colour then comes from the torso field), opaque head, random."""
    R, S = 64, 64
    g = torch.Generator().manual_seed(1)
    z = O.z_vals_uniform(torch.full((R, 1), 0.4), torch.ones(R, 1), S).expand(R, S).contiguous()
    rd, rdt = torch.randn(R, 3, generator=g), torch.randn(R, 3, generator=g)
    fh, ft = torch.rand(R, S, 3, generator=g), torch.rand(R, S, 3, generator=g)
    sh, st = torch.randn(R, S, generator=g) * 8, torch.randn(R, S, generator=g) * 8
    sh[0], st[0] = -1., -1.                  # both fields empty
    sh[1] = 80.                              # opaque head
    sh[2] = -5.                              # empty head, live torso
    bc = torch.rand(R, 3, generator=g)
    # oracle: the reference's op sequence on given field outputs (MAIN:669-708)
    import torch.nn.functional as F
    feat_h = torch.cat((fh[None, :, :-1, :], bc.reshape(1, R, 1, 3)), dim=-2)
    sig_t = st[None].clone()
    sig_t[:, :, -1] = 0
    s1, f1 = F.relu(torch.stack([sh[None]], 0)), torch.stack([feat_h], 0)
    s2, f2 = F.relu(torch.stack([sh[None], sig_t], 0)), torch.stack([feat_h, ft[None]], 0)
    s1[-1, :, :, -1] += 1e-6
    s2[-1, :, :, -1] += 1e-6
    ss1, fw1 = O.composite_function(s1, f1)
    ss2, fw2 = O.composite_function(s2, f2)
    w1 = O.calc_volume_weights(z[None], rd[None], ss1)
    w2 = O.calc_volume_weights(z[None], rdt[None], ss2)
    ref_h = torch.sum(w1.unsqueeze(-1) * fw1, dim=-2)[0]
    ref_p = torch.sum(w2.unsqueeze(-1) * fw2, dim=-2)[0]
    from dfa_nerf_b200._lib import lib, ptr, stream_ptr
    t = [x.contiguous().to(DEV) for x in (fh, sh, ft, st, bc, z, rd, rdt)]
    oh, op = torch.empty(R, 3, device=DEV), torch.empty(R, 3, device=DEV)
    assert lib.dfn_composite_head_torso(R, S, *[ptr(x) for x in t], 1e10, ptr(oh), ptr(op), stream_ptr()) == 0
    assert maxerr(oh, ref_h) < 2e-6 and maxerr(op, ref_p) < 2e-6


def test_decoder_options_golden(dfn, golden):
    """Decoder options outside MAIN:518's configuration (listener layers DEC:305-308,322-325; use_expression DEC:279-281,
    333-334; several skips; ray_d None; no final sigmoid; no deformation field; other widths) through the fp32 building blocks,
    against the reference's own outputs; the state_dict loads strictly (same parameter names, incl. expnet / w2lnet)."""
    from oracle.decoder_option_cases import CASES
    g = golden('decoder_options')
    for name, (kw, calls) in CASES.items():
        sd = {k[len(name) + 4:]: v for k, v in g.items() if k.startswith(name + '/sd/')}
        inp = {k[len(name) + 4:]: v.to(DEV) for k, v in g.items() if k.startswith(name + '/in/')}
        dec = dfn.Decoder(**kw)
        assert list(dec.state_dict().keys()) == list(sd.keys())
        dec.load_state_dict(sd, strict=True)
        dec = dec.to(DEV)
        for call, which, has_sig, has_ex, has_rd in calls:
            sig = (inp['signal'] if which == 'head' else inp['signal_torso']) if has_sig else None
            arg = [sig, inp['expression'] if has_ex else None] if which == 'head' else sig
            f, s = dec(inp['p'], inp['ray_d'] if has_rd else None, inp['z_shape'], inp['z_app'], arg, which)
            rf, rs = g['%s/out/%s/feat' % (name, call)], g['%s/out/%s/sigma' % (name, call)]
            assert f.shape == rf.shape and s.shape == rs.shape
            assert maxerr(f, rf) < 2e-5 and maxerr(s, rs) < 2e-5, (name, call, maxerr(f, rf), maxerr(s, rs))
    # what the reference cannot run either is refused with a reason, and the fused path says what it is built for
    with pytest.raises(dfn.DfnError):
        dfn.Decoder(positional_encoding='gauss')
    d2 = dfn.Decoder(n_blocks_view=2).to(DEV)
    with pytest.raises(dfn.DfnError):
        d2(inp['p'], inp['ray_d'], torch.zeros(1, 64, device=DEV), torch.zeros(1, 64, device=DEV), torch.zeros(1, 64, device=DEV), 'head')
    with pytest.raises(dfn.DfnError):
        dec.query_rays(inp['p'][0], inp['ray_d'][0], torch.zeros(53, 1, device=DEV), inp['z_shape'], inp['z_app'], inp['signal'], 'head')
