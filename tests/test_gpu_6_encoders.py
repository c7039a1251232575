"""GPU parity of the latent encoders and the per-frame signal assembly (HELP:109-240, MAIN:28-111): the
sequence-level pre-pass against the vectors the reference produced frame by frame (tests/golden/encoders.npz)."""
import types

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


def load(mod, sd):
    mod.load_state_dict(sd)
    return mod.to(DEV)


def test_audionet_golden(dfn, golden):
    g = golden('encoders')
    net = load(dfn.AudioNet(dim_aud=76, win_size=16), synth.audionet_state_dict(0))
    y = net(g['x_audionet'].to(DEV))
    assert y.shape == (5, 76) and maxerr(y, g['y_audionet']) < 1e-5


def test_mlp_encoders_vs_oracle(dfn):
    g = torch.Generator().manual_seed(1)
    sd_a, sd_e = synth.mlp_encoder_state_dict(1), synth.mlp_encoder_state_dict(2, (64, 32, 32))
    a, e = load(dfn.AudioNet_W2L(), sd_a), load(dfn.ExpressionEnc(), sd_e)
    for n in (1, 37):
        xa, xe = torch.randn(n, 512, generator=g), torch.randn(n, 64, generator=g)
        with torch.no_grad():
            assert maxerr(a(xa.to(DEV)), O.mlp_encoder_forward(sd_a, xa)) < 1e-5
            assert maxerr(e(xe.to(DEV)), O.mlp_encoder_forward(sd_e, xe)) < 1e-5


def test_encode_signal_sequence_golden(dfn, golden):
    """Whole-sequence pre-pass == the reference's per-frame encode_signal, both branches (window edges included)."""
    g = golden('encoders')
    a, e = load(dfn.AudioNet_W2L(), synth.mlp_encoder_state_dict(1)), load(dfn.ExpressionEnc(), synth.mlp_encoder_state_dict(2, (64, 32, 32)))
    att = load(dfn.AudioAttNet(dim_aud=96, seq_len=4), synth.audio_att_state_dict(3, 96, 4))
    auds, exps = g['auds'].to(DEV), g['exps'].to(DEV)
    plain = dfn.encode_signal_sequence(auds, exps, a, e)
    smooth = dfn.encode_signal_sequence(auds, exps, a, e, att)
    assert plain.shape == (20, 96) and maxerr(plain, g['sig_plain']) < 1e-5
    assert maxerr(smooth, g['sig_smooth']) < 1e-5
    # the reference's per-frame signature on top of it
    args = types.SimpleNamespace(nosmo_iters=10, smo_size=4, smo_torse_size=8)
    ds = [{'auds': auds, 'exp': exps}]
    for i in (0, 1, 10, 19):
        assert maxerr(dfn.encode_signal(ds, 0, i, 96, a, e, att, 0, args, 20)[0], g['sig_plain'][i:i + 1]) < 1e-5
        assert maxerr(dfn.encode_signal(ds, 0, i, 96, a, e, att, 20, args, 20)[0], g['sig_smooth'][i:i + 1]) < 1e-5


def test_encode_signal_torso_sequence_golden(dfn, golden):
    g = golden('encoders')
    patt = load(dfn.AudioAttNet(dim_aud=42, seq_len=8), synth.audio_att_state_dict(4, 42, 8))
    poses = g['poses'].to(DEV)
    plain = dfn.encode_signal_torso_sequence(poses)
    smooth = dfn.encode_signal_torso_sequence(poses, patt)
    assert plain.shape == (12, 42) and maxerr(plain, g['torso_plain']) < 2e-6
    assert maxerr(smooth, g['torso_smooth']) < 1e-5
    with torch.no_grad():
        assert maxerr(dfn.pose_to_euler_trans(poses), O.pose_to_euler_trans(g['poses'])) < 1e-6
        assert maxerr(dfn.encode_signal_torso_sequence(poses[:, :3].contiguous()), g['torso_plain']) < 2e-6   # [N,3,4] poses
    args = types.SimpleNamespace(nosmo_iters=10, smo_size=4, smo_torse_size=8)
    ds = [{'poses': poses}]
    for i in (0, 3, 11):
        assert maxerr(dfn.encode_signal_torso(ds, 0, i, patt, 0, args, 12), g['torso_plain'][i:i + 1]) < 2e-6
        assert maxerr(dfn.encode_signal_torso(ds, 0, i, patt, 20, args, 12).reshape(1, -1), g['torso_smooth'][i:i + 1]) < 1e-5


def test_audio_att_single_window(dfn):
    sd = synth.audio_att_state_dict(5, 64, 8)
    att = load(dfn.AudioAttNet(dim_aud=64, seq_len=8), sd)
    x = torch.randn(8, 96, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        assert maxerr(att(x.to(DEV)), O.audio_att_forward(sd, x, 64, 8)) < 1e-5


def test_live_loop_sequence_end_to_end(dfn):
    """The reference's driven-sequence loop MAIN:624-733 end to end: latents (AudioNet_W2L + ExpressionEnc, pose signal)
    -> head + torso Decoder fields -> two-field compositing -> to8b, against the oracle chain frame by frame."""
    H, W, n = 10, 12, 3
    fr = synth.frame_inputs(H=H, W=W, seed=3)
    sd_a, sd_e, sd_d = synth.mlp_encoder_state_dict(1), synth.mlp_encoder_state_dict(2, (64, 32, 32)), synth.decoder_state_dict(2)
    a, e = load(dfn.AudioNet_W2L(), sd_a), load(dfn.ExpressionEnc(), sd_e)
    dec = load(dfn.Decoder(z_dim=256, hidden_size=256, dim_signal=96, use_deformation_field=True), sd_d)
    g = torch.Generator().manual_seed(8)
    auds, exps = torch.randn(n, 512, generator=g), torch.randn(n, 64, generator=g)
    poses = synth.pose_sequence(n, 20)
    body = synth.camera_pose(31)
    zs, za = torch.randn(1, 2, 256, generator=g), torch.randn(1, 2, 256, generator=g)
    signals = dfn.encode_signal_sequence(auds.to(DEV), exps.to(DEV), a, e)
    signals_t = dfn.encode_signal_torso_sequence(poses.to(DEV))
    frames = dfn.render_sequence_head_torso(dec, H, W, fr['focal'], poses, body, fr['bc_rgb'].to(DEV), zs.to(DEV), za.to(DEV),
                                            signals, signals_t, fr['near'], fr['far'], fr['cx'], fr['cy'], N_samples=64,
                                            precision=dfn.PREC_BF16X3)
    assert frames.shape == (n, H, W, 3) and frames.dtype == torch.uint8
    rot, rdt = [t.reshape(-1, 3) for t in O.get_rays(H, W, fr['focal'], body, fr['cx'], fr['cy'])]
    z = O.z_vals_uniform(torch.full((H * W, 1), fr['near']), torch.full((H * W, 1), fr['far']), 64)
    with torch.no_grad():
        for i in range(n):
            sig = O.encode_signal(auds, exps, i, sd_a, sd_e)
            sig_t = O.encode_signal_torso(poses, i)
            ro, rd = [t.reshape(-1, 3) for t in O.get_rays(H, W, fr['focal'], poses[i, :3, :4], fr['cx'], fr['cy'])]
            _, person = O.render_head_torso_chunk(sd_d, ro, rd, rot, rdt, z, fr['bc_rgb'], zs, za, sig, sig_t)
            ref8 = torch.from_numpy(O.to8b(person.numpy())).reshape(H, W, 3)
            diff = (frames[i].int() - ref8.int()).abs()
            assert diff.max().item() <= 1 and (diff > 0).float().mean().item() < 0.01   # 1e-6 errors can flip a truncation
