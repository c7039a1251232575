"""CPU: the in-repo oracle reproduces the vectors the REFERENCE produced (oracle/make_golden.py).

make_golden.py asserted bit-equality oracle == reference in the build container; here the
same comparison is repeated from the committed fixtures (tolerance 2e-6 absolute on floats
to allow for a different host BLAS; indices exact)."""
import numpy as np
import torch

from oracle import nerf_oracle as O
from oracle import synth

TOL = 2e-6


def close(a, b, tol=TOL):
    assert a.shape == b.shape
    assert (a.double() - b.double()).abs().max().item() <= tol


def test_get_rays(golden):
    g = golden('get_rays')
    for tag in 'abc':
        H, W, f, cx, cy, stride = [float(v) for v in g[tag + '_args']]
        o, d = O.get_rays(int(H), int(W), f, g[tag + '_c2w'], None if cx < 0 else cx, None if cy < 0 else cy, int(stride))
        assert torch.equal(o, g[tag + '_o']) and torch.equal(d, g[tag + '_d'])
    _, d = O.get_rays(450, 450, 1200., g['d_c2w'], 225., 225.)
    assert torch.equal(d.reshape(-1, 3)[g['d_idx']], g['d_d'])


def test_z_vals(golden):
    g = golden('z_vals')
    assert torch.equal(O.linspace_table(64), g['t64']) and torch.equal(O.linspace_table(128), g['t128'])
    assert torch.equal(O.z_vals_uniform(torch.full((1, 1), 0.4), torch.ones(1, 1), 64)[0], g['z64'])


def test_embed(golden):
    g = golden('embed')
    close(O.embed(g['x'], 10), g['pe10'])
    close(O.embed(g['x'], 4), g['pe4'])
    close(O.embed(g['x'], 3), g['pe3'])
    close(O.decoder_transform_points(g['x'][None], 10)[0], g['tp10'])
    close(O.decoder_transform_points(g['x'][None], 4)[0], g['tp4'])
    assert O.embed_dim(10) == 63 and O.embed_dim(4) == 27


def test_mlps(golden):
    g = golden('mlp')
    close(O.facenerf_forward(synth.facenerf_state_dict(g['face_seed']), g['x_face']), g['y_face'], 2e-4)
    close(O.nerf_forward(synth.nerf_state_dict(g['nerf_seed']), g['x_nerf']), g['y_nerf'], 2e-4)


def test_decoder(golden):
    g = golden('decoder')
    sd = synth.decoder_state_dict(g['seed'])
    fh, sh = O.decoder_forward(sd, g['p'], g['ray_d'], g['z_shape'][:, 0], g['z_app'][:, 0], g['signal'], 'head')
    ft, st = O.decoder_forward(sd, g['p'], g['ray_d'], g['z_shape'][:, 1], g['z_app'][:, 1], g['signal_torso'], 'torso')
    close(fh, g['feat_head'], 1e-5)
    close(sh, g['sigma_head'], 2e-4)
    close(ft, g['feat_torso'], 1e-5)
    close(st, g['sigma_torso'], 2e-4)


def test_composite(golden):
    g = golden('composite')
    close(O.calc_volume_weights(g['z'][None], g['rays_d'][None], g['sigma'][None])[0], g['weights'])
    ss, fw = O.composite_function(g['sigma2'][:, None].clone(), g['feat2'][:, None])
    close(ss[0], g['sigma_sum'])
    close(fw[0], g['feat_w'])


def test_raw2outputs(golden):
    g = golden('raw2outputs')
    rgb, disp, acc, w, depth = O.raw2outputs(g['raw'], g['z'], g['rays_d'], g['bc_rgb'])
    close(rgb, g['rgb_map'])
    close(acc, g['acc_map'])
    close(w, g['weights'])
    close(depth, g['depth_map'])
    assert torch.allclose(disp, g['disp_map'], rtol=1e-5)


def test_sample_pdf(golden):
    g = golden('sample_pdf')
    s, i = O.sample_pdf(g['bins'], g['weights'], 128, det=True, return_inds=True)
    assert torch.equal(i, g['det_inds'])
    close(s, g['det'])
    s, i = O.sample_pdf(g['bins'], g['weights'], 128, u=g['u_py'], return_inds=True)
    assert torch.equal(i, g['py_inds'])
    close(s, g['py'])
    s, i = O.sample_pdf(g['bins'][:3], g['edge_weights'], 128, det=True, return_inds=True)
    assert torch.equal(i, g['edge_inds'])
    close(s, g['edge'])
    # u = 1.0 lands past the last knot: below = above = 62 -> sample == bins[62] (SURVEY A, probe C.3)
    assert torch.equal(s[:, -1], g['bins'][:3, -1])


def test_render_rays_stagewise(golden):
    """Full hierarchical pass; gated stage-wise because coarse->fine is ill-conditioned."""
    g = golden('render_rays')
    sd_c, sd_f = synth.facenerf_state_dict(g['coarse_seed']), synth.facenerf_state_dict(g['fine_seed'])
    _, rd = O.get_rays(g['H'], g['W'], g['focal'], g['c2w'], g['cx'], g['cy'])
    ro = g['c2w'][:3, -1].expand(rd.shape).reshape(-1, 3)
    rd = rd.reshape(-1, 3)
    assert torch.equal(rd, g['rays_d'])
    n = rd.shape[0]
    vd = rd / torch.norm(rd, dim=-1, keepdim=True)
    rays = torch.cat([ro, rd, g['near'] * torch.ones(n, 1), g['far'] * torch.ones(n, 1), vd], -1)
    with torch.no_grad():
        o = O.render_rays(rays, g['bc_rgb'], g['aud'], sd_c, sd_f, 64, 128, z_samples_override=g['z_samples'], retraw=True)
    close(o['raw0'], g['raw0'], 5e-4)
    close(o['rgb0'], g['rgb0'], 1e-5)
    assert torch.equal(o['z_vals'], g['z_vals'])        # teacher-forced: identical sample depths
    close(o['raw'], g['raw'], 5e-4)
    close(o['rgb_map'], g['rgb_map'], 1e-5)


def test_head_torso(golden):
    g = golden('head_torso')
    sd = synth.decoder_state_dict(g['seed'])
    ro, rd = O.get_rays(g['H'], g['W'], g['focal'], g['c2w'], g['cx'], g['cy'])
    rot, rdt = O.get_rays(g['H'], g['W'], g['focal'], g['c2w_torso'], g['cx'], g['cy'])
    ro, rd, rot, rdt = [v.reshape(-1, 3) for v in (ro, rd, rot, rdt)]
    z = O.z_vals_uniform(g['near'] * torch.ones(ro.shape[0], 1), g['far'] * torch.ones(ro.shape[0], 1), 64)
    with torch.no_grad():
        h, p = O.render_head_torso_chunk(sd, ro, rd, rot, rdt, z, g['bc_rgb'], g['z_shape'], g['z_app'], g['signal'], g['signal_torso'])
    close(h, g['rgb_head'], 1e-5)
    close(p, g['rgb_person'], 1e-5)


def test_bf16_restatement_is_close_to_fp32(golden):
    """The bf16-operand restatement differs from fp32 by O(1e-2) relative on sigma at most -- sanity only."""
    g = golden('mlp')
    sd = synth.facenerf_state_dict(g['face_seed'])
    y = O.facenerf_forward_bf16(sd, g['x_face'])
    assert (y[:, :3] - g['y_face'][:, :3]).abs().max() < 5e-2


def test_encoders(golden):
    """Latent encoders and the per-frame signal assembly (HELP:109-240, MAIN:28-111) against the reference's outputs."""
    g = golden('encoders')
    with torch.no_grad():
        close(O.audionet_forward(synth.audionet_state_dict(0), g['x_audionet']), g['y_audionet'], 1e-6)
        sd_a, sd_e = synth.mlp_encoder_state_dict(1), synth.mlp_encoder_state_dict(2, (64, 32, 32))
        n = g['auds'].shape[0]
        plain = torch.cat([O.encode_signal(g['auds'], g['exps'], i, sd_a, sd_e) for i in range(n)], 0)
        smooth = torch.cat([O.encode_signal(g['auds'], g['exps'], i, sd_a, sd_e, synth.audio_att_state_dict(3, 96, 4), 4, 96)
                            for i in range(n)], 0)
        close(plain, g['sig_plain'], 1e-6)
        close(smooth, g['sig_smooth'], 1e-6)
        m = g['poses'].shape[0]
        close(torch.cat([O.encode_signal_torso(g['poses'], i) for i in range(m)], 0), g['torso_plain'], 1e-6)
        close(torch.cat([O.encode_signal_torso(g['poses'], i, synth.audio_att_state_dict(4, 42, 8), 8, 3) for i in range(m)], 0),
              g['torso_smooth'], 1e-6)


def test_to8b():
    assert np.array_equal(O.to8b(np.array([-1., 0., 0.5, 1., 2.])), np.array([0, 0, 127, 255, 255], np.uint8))


def test_decoder_options(golden):
    """Listener layers, expression term, several skips, no view directions / sigmoid / deformation field: the restatement against
    the reference's own outputs (oracle/make_golden_decoder_options.py)."""
    from oracle.decoder_option_cases import CASES, oracle_kwargs
    g = golden('decoder_options')
    for name, (kw, calls) in CASES.items():
        sd = {k[len(name) + 4:]: v for k, v in g.items() if k.startswith(name + '/sd/')}
        inp = {k[len(name) + 4:]: v for k, v in g.items() if k.startswith(name + '/in/')}
        for call, which, has_sig, has_ex, has_rd in calls:
            sig = (inp['signal'] if which == 'head' else inp['signal_torso']) if has_sig else None
            f, s = O.decoder_forward(sd, inp['p'], inp['ray_d'] if has_rd else None, inp['z_shape'], inp['z_app'], sig, which,
                                     expression=inp['expression'] if has_ex else None, **oracle_kwargs(kw, has_ex))
            close(f, g['%s/out/%s/feat' % (name, call)], 1e-6)
            close(s, g['%s/out/%s/sigma' % (name, call)], 1e-5)
