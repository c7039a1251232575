"""Pin the oracle against the real reference and write tests/golden/*.npz.

Run in the build container only (needs /root/reference):
    python oracle/make_golden.py

1. imports the reference's own modules (run_nerf_helpers, decoder,
   run_nerf_com_trainExpLater with imageio/configargparse stubbed -- neither is touched
   by the hot path);
2. asserts every oracle function == the reference function bit-for-bit on seeded inputs;
3. stores small input/output vectors produced BY THE REFERENCE under tests/golden/.

The reference has no tests or golden vectors of its own; these
vectors are what pins parity.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = '/root/reference/NeRFs/DFANeRF'
sys.path.insert(0, REF)
for m in ('imageio', 'configargparse'):
    sys.modules.setdefault(m, types.ModuleType(m))

import run_nerf_helpers as HELP  # noqa: E402
import decoder as DEC  # noqa: E402
import run_nerf_com_trainExpLater as MAIN  # noqa: E402

from oracle import nerf_oracle as O  # noqa: E402
from oracle import synth  # noqa: E402

torch.autograd.set_detect_anomaly(False)
OUT = os.path.join(ROOT, 'tests', 'golden')
os.makedirs(OUT, exist_ok=True)


def same(a, b, what):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if not torch.equal(a, b):
        d = (a.double() - b.double()).abs().max().item()
        raise AssertionError('%s: oracle != reference (max abs %.3e)' % (what, d))
    print('  ok  %-34s %s' % (what, tuple(a.shape)))


def save(name, **arrs):
    np.savez_compressed(os.path.join(OUT, name + '.npz'),
                        **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()})


def ref_module(cls, sd, **kw):
    m = cls(**kw)
    m.load_state_dict(sd)
    return m.eval()


@torch.no_grad()
def main():
    g = torch.Generator().manual_seed(1234)

    # ---- a1 get_rays ------------------------------------------------------------
    print('get_rays (HELP:449)')
    cases = {}
    for tag, (H, W, f, cx, cy, stride, seed) in {
            'a': (4, 6, 10., 3., 2., 1, 0), 'b': (18, 14, 37.5, 6.6, 9.1, 1, 1),
            'c': (16, 12, 40., None, None, 2, 2), 'd': (450, 450, 1200., 225., 225., 1, 3)}.items():
        c2w = synth.camera_pose(seed)
        ro, rd = HELP.get_rays(H, W, f, c2w, cx, cy, stride)
        oo, od = O.get_rays(H, W, f, c2w, cx, cy, stride)
        same(oo, ro, 'rays_o ' + tag)
        same(od, rd, 'rays_d ' + tag)
        if tag != 'd':
            cases.update({'%s_args' % tag: np.array([H, W, f, -1 if cx is None else cx, -1 if cy is None else cy, stride], np.float64),
                          '%s_c2w' % tag: c2w, '%s_o' % tag: ro.contiguous(), '%s_d' % tag: rd})
        else:
            idx = torch.tensor([0, 1, 449, 450, 101250, 202499])
            cases.update(d_c2w=c2w, d_idx=idx, d_d=rd.reshape(-1, 3)[idx])
    save('get_rays', **cases)

    # ---- a2 z sampling -------------------------------------------------------------
    print('z_vals (MAIN:617-619)')
    near = 0.4 * torch.ones((5, 1))
    far = 1.0 * torch.ones((5, 1))
    t = torch.linspace(0., 1., steps=64)
    z_ref = near * (1. - t) + far * t
    same(O.z_vals_uniform(near, far, 64), z_ref, 'z_vals 64')
    save('z_vals', t64=t, t128=torch.linspace(0., 1., steps=128), z64=z_ref[0])

    # ---- a4 / a4' positional encodings ------------------------------------------------
    print('Embedder (HELP:21-70), Decoder.transform_points (DEC:257)')
    x = (torch.rand((96, 3), generator=g) * 2 - 1) * 1.2
    e10, d10 = HELP.get_embedder(10, 0)
    e4, d4 = HELP.get_embedder(4, 0)
    e3, d3 = HELP.get_embedder(3, 0)
    assert (d10, d4, d3) == (63, 27, 21)
    same(O.embed(x, 10), e10(x), 'embed L=10')
    same(O.embed(x, 4), e4(x), 'embed L=4')
    same(O.embed(x, 3), e3(x), 'embed L=3')
    dsd = synth.decoder_state_dict(seed=5)
    dec = ref_module(DEC.Decoder, dsd, z_dim=256, hidden_size=256, dim_signal=96, use_deformation_field=True)
    tp10 = dec.transform_points(x[None])
    tp4 = dec.transform_points(x[None], views=True)
    same(O.decoder_transform_points(x[None], 10), tp10, 'transform_points L=10')
    same(O.decoder_transform_points(x[None], 4), tp4, 'transform_points L=4')
    save('embed', x=x, pe10=e10(x), pe4=e4(x), pe3=e3(x), tp10=tp10[0], tp4=tp4[0])

    # ---- a5 / a5' MLPs --------------------------------------------------------------
    print('FaceNeRF.forward (HELP:275), NeRF.forward (HELP:372)')
    sd_f = synth.facenerf_state_dict(seed=0)
    face = ref_module(HELP.FaceNeRF, sd_f, D=8, W=256, input_ch=63, input_ch_views=27, dim_aud=64,
                      output_ch=4, skips=[4], use_viewdirs=True)
    pts = (torch.rand((48, 3), generator=g) * 2 - 1) * 0.8
    vd = torch.nn.functional.normalize(torch.randn((48, 3), generator=g), dim=-1)
    aud = torch.randn((64,), generator=g)
    xin = torch.cat([e10(pts), aud[None].expand(48, -1), e4(vd)], -1)
    y_ref = face(xin)
    same(O.facenerf_forward(sd_f, xin), y_ref, 'FaceNeRF forward')
    sd_n = synth.nerf_state_dict(seed=2)
    nerf = ref_module(HELP.NeRF, sd_n, D=8, W=256, input_ch=63, input_ch_views=27, output_ch=4, skips=[4], use_viewdirs=True)
    xn = torch.cat([e10(pts), e4(vd)], -1)
    yn_ref = nerf(xn)
    same(O.nerf_forward(sd_n, xn), yn_ref, 'NeRF forward')
    save('mlp', pts=pts, viewdirs=vd, aud=aud, x_face=xin, y_face=y_ref, x_nerf=xn, y_nerf=yn_ref,
         face_seed=0, nerf_seed=2)

    # ---- a5'' Decoder (+ DeformationField_ori) -------------------------------------------
    print('Decoder.forward head/torso (DEC:277), DeformationField_ori (DEC:109)')
    P = 40
    p_in = ((torch.rand((1, P, 3), generator=g) * 2 - 1) * 0.7)
    rd = torch.randn((1, P, 3), generator=g)
    z_shape = torch.randn((1, 2, 256), generator=g)
    z_app = torch.randn((1, 2, 256), generator=g)
    sig_h = torch.randn((1, 96), generator=g)
    sig_t = torch.randn((1, 42), generator=g)
    fh, sh = dec(p_in, rd, z_shape[:, 0], z_app[:, 0], [sig_h, None], 'head')
    ft, st = dec(p_in, rd, z_shape[:, 1], z_app[:, 1], sig_t, 'torso')
    ofh, osh = O.decoder_forward(dsd, p_in, rd, z_shape[:, 0], z_app[:, 0], sig_h, 'head')
    oft, ost = O.decoder_forward(dsd, p_in, rd, z_shape[:, 1], z_app[:, 1], sig_t, 'torso')
    same(ofh, fh, 'Decoder head feat')
    same(osh, sh, 'Decoder head sigma')
    same(oft, ft, 'Decoder torso feat')
    same(ost, st, 'Decoder torso sigma')
    save('decoder', p=p_in, ray_d=rd, z_shape=z_shape, z_app=z_app, signal=sig_h, signal_torso=sig_t,
         feat_head=fh, sigma_head=sh, feat_torso=ft, sigma_torso=st, seed=5)

    # ---- a7 / a8 compositing ------------------------------------------------------------
    print('composite_function (MAIN:146), calc_volume_weights (MAIN:169)')
    R, S = 24, 64
    z = O.z_vals_uniform(0.4 * torch.ones(R, 1), torch.ones(R, 1), S).expand(R, S).contiguous()
    rdir = torch.randn((R, 3), generator=g)
    sigma = torch.randn((1, R, S), generator=g) * 6 + 1
    sigma[0, 3] = 0.                      # an empty ray
    sigma[0, 4] = 50.                     # an opaque ray
    w_ref = MAIN.calc_volume_weights(z[None], rdir[None], sigma, last_dist=1e10)
    same(O.calc_volume_weights(z[None], rdir[None], sigma), w_ref, 'calc_volume_weights')
    sig2 = torch.relu(torch.randn((2, 1, R, S), generator=g) * 4)
    sig2[:, :, 5, :7] = 0.                # both fields empty -> denominator patch
    feat2 = torch.rand((2, 1, R, S, 3), generator=g)
    ss_ref, fw_ref = MAIN.composite_function(sig2.clone(), feat2)
    oss, ofw = O.composite_function(sig2.clone(), feat2)
    same(oss, ss_ref, 'composite sigma_sum')
    same(ofw, fw_ref, 'composite feat')
    ss1_ref, fw1_ref = MAIN.composite_function(sig2[:1].clone(), feat2[:1])
    same(O.composite_function(sig2[:1].clone(), feat2[:1])[0], ss1_ref, 'composite n_box=1')
    save('composite', z=z, rays_d=rdir, sigma=sigma[0], weights=w_ref[0],
         sigma2=sig2[:, 0], feat2=feat2[:, 0], sigma_sum=ss_ref[0], feat_w=fw_ref[0])

    # ---- raw2outputs (upstream glue over calc_volume_weights) ---------------------------
    print('raw2outputs (Appendix B; core == MAIN:169-179)')
    raw = torch.randn((R, S, 4), generator=g)
    raw[..., 3] = raw[..., 3] * 8 + 1
    bc = torch.rand((R, 3), generator=g)
    rgb_map, disp, acc, wts, depth = O.raw2outputs(raw, z, rdir, bc)
    w_chk = MAIN.calc_volume_weights(z[None], rdir[None], raw[None, ..., 3])[0]
    same(wts, w_chk, 'raw2outputs weights')
    rgbs = torch.sigmoid(raw[..., :3])
    rgbs = torch.cat((rgbs[:, :-1, :], bc.unsqueeze(1)), dim=1)
    same(rgb_map, torch.sum(w_chk.unsqueeze(-1) * rgbs, dim=-2), 'raw2outputs rgb (MAIN:706)')
    save('raw2outputs', raw=raw, z=z, rays_d=rdir, bc_rgb=bc, rgb_map=rgb_map, disp_map=disp,
         acc_map=acc, weights=wts, depth_map=depth)

    # ---- a10 sample_pdf -----------------------------------------------------------------
    print('sample_pdf (HELP:537)')
    z_mid = .5 * (z[..., 1:] + z[..., :-1])
    wmid = wts[..., 1:-1].contiguous()
    s_det = HELP.sample_pdf(z_mid, wmid, 128, det=True)
    o_det, o_inds = O.sample_pdf(z_mid, wmid, 128, det=True, return_inds=True)
    same(o_det, s_det, 'sample_pdf det')
    s_py = HELP.sample_pdf(z_mid, wmid, 128, det=False, pytest=True)
    np.random.seed(0)
    u_py = torch.Tensor(np.random.rand(R, 128))
    o_py, o_py_inds = O.sample_pdf(z_mid, wmid, 128, u=u_py, return_inds=True)
    same(o_py, s_py, 'sample_pdf pytest=True (np seed 0)')
    s_pyd = HELP.sample_pdf(z_mid, wmid, 128, det=True, pytest=True)
    o_pyd = O.sample_pdf(z_mid, wmid, 128, u=torch.Tensor(np.linspace(0., 1., 128)))
    same(o_pyd, s_pyd, 'sample_pdf det+pytest (np.linspace u)')
    wz = torch.zeros((3, 62))
    wz[1, 10] = 1.
    wz[2, :] = 1e-7
    s_edge = HELP.sample_pdf(z_mid[:3], wz, 128, det=True)
    o_edge, o_edge_inds = O.sample_pdf(z_mid[:3], wz, 128, det=True, return_inds=True)
    same(o_edge, s_edge, 'sample_pdf edge (zero / delta weights)')
    save('sample_pdf', bins=z_mid, weights=wmid, det=s_det, det_inds=o_inds, u_py=u_py, py=s_py, py_inds=o_py_inds,
         edge_weights=wz, edge=s_edge, edge_inds=o_edge_inds)

    # ---- full hierarchical render_rays (composition of the reference's own functions) ----
    print('render_rays 64+128 (composition: HELP.get_rays/Embedder/FaceNeRF/sample_pdf + MAIN.calc_volume_weights)')
    fr = synth.frame_inputs(H=20, W=16, seed=7)
    sd_c, sd_fine = synth.facenerf_state_dict(seed=0), synth.facenerf_state_dict(seed=1)
    face_c = face
    face_f = ref_module(HELP.FaceNeRF, sd_fine, D=8, W=256, input_ch=63, input_ch_views=27, dim_aud=64,
                        output_ch=4, skips=[4], use_viewdirs=True)
    ro, rdd = HELP.get_rays(fr['H'], fr['W'], fr['focal'], fr['c2w'], fr['cx'], fr['cy'])
    ro, rdd = ro.reshape(-1, 3), rdd.reshape(-1, 3)
    vdd = rdd / torch.norm(rdd, dim=-1, keepdim=True)
    n = ro.shape[0]
    near, far = fr['near'] * torch.ones(n, 1), fr['far'] * torch.ones(n, 1)

    def ref_query(net, pts_, vd_):
        flat = pts_.reshape(-1, 3)
        dirs = vd_[:, None].expand(pts_.shape).reshape(-1, 3)
        xx = torch.cat([e10(flat), fr['aud'][None].expand(flat.shape[0], -1), e4(dirs)], -1)
        return net(xx).reshape(list(pts_.shape[:-1]) + [4])

    def ref_raw2outputs(raw_, z_, rd_, bc_):
        rgb_ = torch.sigmoid(raw_[..., :3])
        rgb_ = torch.cat((rgb_[:, :-1, :], bc_.unsqueeze(1)), dim=1)
        w_ = MAIN.calc_volume_weights(z_[None], rd_[None], raw_[None, ..., 3])[0]
        return torch.sum(w_.unsqueeze(-1) * rgb_, dim=-2), w_

    tt = torch.linspace(0., 1., steps=64)
    zc = (near * (1. - tt) + far * tt).expand(n, 64)
    ptc = ro[..., None, :] + rdd[..., None, :] * zc[..., :, None]
    raw_c = ref_query(face_c, ptc, vdd)
    rgb0, w0 = ref_raw2outputs(raw_c, zc, rdd, fr['bc_rgb'])
    zm = .5 * (zc[..., 1:] + zc[..., :-1])
    zs = HELP.sample_pdf(zm, w0[..., 1:-1], 128, det=True)
    zf, _ = torch.sort(torch.cat([zc, zs], -1), -1)
    ptf = ro[..., None, :] + rdd[..., None, :] * zf[..., :, None]
    raw_f = ref_query(face_f, ptf, vdd)
    rgb1, w1 = ref_raw2outputs(raw_f, zf, rdd, fr['bc_rgb'])
    rays = torch.cat([ro, rdd, near, far, vdd], -1)
    o = O.render_rays(rays, fr['bc_rgb'], fr['aud'], sd_c, sd_fine, 64, 128, retraw=True)
    same(o['raw0'], raw_c, 'render_rays coarse raw')
    same(o['rgb0'], rgb0, 'render_rays coarse rgb')
    same(o['z_samples'], zs, 'render_rays z_samples')
    same(o['z_vals'], zf, 'render_rays merged z')
    same(o['raw'], raw_f, 'render_rays fine raw')
    same(o['rgb_map'], rgb1, 'render_rays fine rgb')
    r2 = O.render(fr['H'], fr['W'], fr['focal'], fr['cx'], fr['cy'], fr['c2w'], fr['bc_rgb'], fr['aud'],
                  sd_c, sd_fine, fr['near'], fr['far'], chunk=128)
    same(r2['rgb_map'], rgb1, 'render() chunked == unchunked')
    fg = 1. - o['last_weight']
    print('       foreground opacity mean %.3f, max %.3f' % (fg.mean().item(), fg.max().item()))
    save('render_rays', H=fr['H'], W=fr['W'], focal=fr['focal'], cx=fr['cx'], cy=fr['cy'], near=fr['near'], far=fr['far'],
         c2w=fr['c2w'], aud=fr['aud'], bc_rgb=fr['bc_rgb'], coarse_seed=0, fine_seed=1,
         rays_d=rdd, raw0=raw_c, rgb0=rgb0, weights0=w0, z_samples=zs, z_vals=zf, raw=raw_f, rgb_map=rgb1,
         weights=w1, disp_map=o['disp_map'], acc_map=o['acc_map'])

    # ---- live two-field chunk (MAIN:661-708) ----------------------------------------------
    print('head+torso chunk (MAIN:661-708)')
    fr2 = synth.frame_inputs(H=8, W=12, seed=9)
    ro, rdd = HELP.get_rays(fr2['H'], fr2['W'], fr2['focal'], fr2['c2w'], fr2['cx'], fr2['cy'])
    rot, rdt = HELP.get_rays(fr2['H'], fr2['W'], fr2['focal'], synth.camera_pose(10), fr2['cx'], fr2['cy'])
    ro, rdd, rot, rdt = [v.reshape(-1, 3) for v in (ro, rdd, rot, rdt)]
    n = ro.shape[0]
    zc = (fr2['near'] * (1. - tt) + fr2['far'] * tt).expand(n, 64)
    # reference ops, MAIN:638-708, batch_size == 1
    p_i = (ro[..., None, :] + rdd[..., None, :] * zc[..., :, None]).reshape(1, -1, 3)
    r_i = rdd.unsqueeze(1).expand([n, 64, 3]).reshape(1, -1, 3)
    p_t = (rot[..., None, :] + rdt[..., None, :] * zc[..., :, None]).reshape(1, -1, 3)
    r_t = rdt.unsqueeze(1).expand([n, 64, 3]).reshape(1, -1, 3)
    feat_i, sigma_i = dec(p_i, r_i, z_shape[:, 0], z_app[:, 0], [sig_h, None], 'head')
    sigma_i = sigma_i.reshape(1, -1, 64)
    feat_i = feat_i.reshape(1, -1, 64, 3)
    feat_i = torch.cat((feat_i[..., :-1, :], fr2['bc_rgb'].reshape(1, n, 1, 3)), dim=-2)
    feat_it, sigma_it = dec(p_t, r_t, z_shape[:, 1], z_app[:, 1], sig_t, 'torso')
    sigma_it = sigma_it.reshape(1, -1, 64)
    feat_it = feat_it.reshape(1, -1, 64, 3)
    sigma_it[:, :, -1] = 0
    sg = torch.relu(torch.stack([sigma_i], dim=0))
    ft_ = torch.stack([feat_i], dim=0)
    sgt = torch.relu(torch.stack([sigma_i, sigma_it], dim=0))
    ftt = torch.stack([feat_i, feat_it], dim=0)
    sg[-1, :, :, -1] = sg[-1, :, :, -1] + 1e-6
    sgt[-1, :, :, -1] = sgt[-1, :, :, -1] + 1e-6
    ssum, fwt = MAIN.composite_function(sg, ft_)
    ssum_t, fwt_t = MAIN.composite_function(sgt, ftt)
    wh = MAIN.calc_volume_weights(zc[None], rdd[None], ssum, last_dist=1e10)
    wt = MAIN.calc_volume_weights(zc[None], rdt[None], ssum_t, last_dist=1e10)
    rgb_head = torch.sum(wh.unsqueeze(-1) * fwt, dim=-2).squeeze(0)
    rgb_person = torch.sum(wt.unsqueeze(-1) * fwt_t, dim=-2).squeeze(0)
    oh, op = O.render_head_torso_chunk(dsd, ro, rdd, rot, rdt, zc, fr2['bc_rgb'], z_shape, z_app, sig_h, sig_t)
    same(oh, rgb_head, 'two-field rgb_head')
    same(op, rgb_person, 'two-field rgb_person')
    save('head_torso', H=fr2['H'], W=fr2['W'], focal=fr2['focal'], cx=fr2['cx'], cy=fr2['cy'], near=fr2['near'], far=fr2['far'],
         c2w=fr2['c2w'], c2w_torso=synth.camera_pose(10), bc_rgb=fr2['bc_rgb'], z_shape=z_shape, z_app=z_app,
         signal=sig_h, signal_torso=sig_t, rgb_head=rgb_head, rgb_person=rgb_person, seed=5)
    # ---- latent encoders, the callers of the path -----------------------------
    print('encoders (HELP:109-240) + encode_signal / encode_signal_torso (MAIN:28-111)')
    ge = torch.Generator().manual_seed(77)
    sd_an = synth.audionet_state_dict(0)
    x_an = torch.randn(5, 16, 29, generator=ge)
    y_an = ref_module(HELP.AudioNet, sd_an, dim_aud=76, win_size=16)(x_an)
    same(O.audionet_forward(sd_an, x_an), y_an, 'AudioNet')
    sd_w2l, sd_exp = synth.mlp_encoder_state_dict(1), synth.mlp_encoder_state_dict(2, (64, 32, 32))
    m_w2l, m_exp = ref_module(HELP.AudioNet_W2L, sd_w2l), ref_module(HELP.ExpressionEnc, sd_exp)
    n_fr = 20
    auds, exps = torch.randn(n_fr, 512, generator=ge), torch.randn(n_fr, 64, generator=ge)
    same(O.mlp_encoder_forward(sd_w2l, auds), m_w2l(auds), 'AudioNet_W2L')
    same(O.mlp_encoder_forward(sd_exp, exps), m_exp(exps), 'ExpressionEnc')
    sd_att = synth.audio_att_state_dict(3, 96, 4)          # scripts/test_obama.sh: --dim_aud=96 --smo_size=4
    m_att = ref_module(HELP.AudioAttNet, sd_att, dim_aud=96, seq_len=4)
    sd_patt = synth.audio_att_state_dict(4, 42, 8)         # --smo_torse_size 8, dim_torso_signal = 42 (MAIN:515-516)
    m_patt = ref_module(HELP.AudioAttNet, sd_patt, dim_aud=42, seq_len=8)
    args = types.SimpleNamespace(nosmo_iters=10, smo_size=4, smo_torse_size=8)
    ds = [{'auds': auds, 'exp': exps}]
    sig_plain = torch.cat([MAIN.encode_signal(ds, 0, i, 96, m_w2l, m_exp, m_att, 0, args, n_fr)[0] for i in range(n_fr)], 0)
    sig_smooth = torch.cat([MAIN.encode_signal(ds, 0, i, 96, m_w2l, m_exp, m_att, 20, args, n_fr)[0] for i in range(n_fr)], 0)
    same(torch.cat([O.encode_signal(auds, exps, i, sd_w2l, sd_exp) for i in range(n_fr)], 0), sig_plain, 'encode_signal (no smoothing)')
    same(torch.cat([O.encode_signal(auds, exps, i, sd_w2l, sd_exp, sd_att, 4, 96) for i in range(n_fr)], 0), sig_smooth,
         'encode_signal (AudioAttNet window)')
    # rot_to_euler hard-codes .cuda() (MAIN:184): run it on the CPU by making .cuda() the identity for this call
    poses = synth.pose_sequence(12, 0)
    embed_fn, _ = HELP.get_embedder(3, 0)
    dsp = [{'poses': poses}]
    cuda_attr = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        t_plain = torch.cat([MAIN.encode_signal_torso(dsp, 0, i, m_patt, 0, args, 12, embed_fn=embed_fn).reshape(1, -1) for i in range(12)], 0)
        t_smooth = torch.cat([MAIN.encode_signal_torso(dsp, 0, i, m_patt, 20, args, 12, embed_fn=embed_fn).reshape(1, -1) for i in range(12)], 0)
    finally:
        torch.Tensor.cuda = cuda_attr
    same(torch.cat([O.encode_signal_torso(poses, i) for i in range(12)], 0), t_plain, 'encode_signal_torso (no smoothing)')
    same(torch.cat([O.encode_signal_torso(poses, i, sd_patt, 8, 3) for i in range(12)], 0), t_smooth, 'encode_signal_torso (pose AudioAttNet)')
    save('encoders', x_audionet=x_an, y_audionet=y_an, auds=auds, exps=exps, sig_plain=sig_plain, sig_smooth=sig_smooth,
         poses=poses, torso_plain=t_plain, torso_smooth=t_smooth)
    print('golden vectors written to', OUT)


if __name__ == '__main__':
    main()
