"""Pin oracle/train_oracle.py against the real reference and write tests/golden/train_step.npz.

Run in the build container only (needs /root/reference):
    python oracle/make_golden_train.py

The reference's training step is inline code of train() (MAIN:738-931), not a callable; it is re-assembled here from the
reference's OWN modules and functions in the order MAIN uses them, with the JPEG reads replaced by synthetic target
tensors.  Test infrastructure, like everything under oracle/.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = '/root/reference/NeRFs/DFANeRF'
sys.path.insert(0, REF)
for m in ('imageio', 'configargparse'):
    sys.modules.setdefault(m, types.ModuleType(m))

import run_nerf_helpers as HELP  # noqa: E402
import decoder as DEC  # noqa: E402
import run_nerf_com_trainExpLater as MAIN  # noqa: E402

from oracle import train_oracle as TO  # noqa: E402
from oracle import synth  # noqa: E402

torch.autograd.set_detect_anomaly(False)
H, W, N_RAND, NS, LRATE = 24, 20, 96, 16, 5e-4


def make_batch():
    g = torch.Generator().manual_seed(11)
    n = 6
    fr = synth.frame_inputs(H=H, W=W, seed=2)
    poses = torch.cat([synth.pose_sequence(n, 7), torch.tensor([0., 0., 0., 1.]).expand(n, 1, 4)], 1)     # [n,4,4]
    np.random.seed(3)
    coords = TO.select_coords(H, W, [5, 4, 9, 8], N_RAND, 0.95)
    return dict(H=H, W=W, focal=fr['focal'], cx=fr['cx'], cy=fr['cy'], near=fr['near'], far=fr['far'], poses=poses, img_i=2,
                pose=poses[2, :3, :4], pose_torso=poses[0, :3, :4], auds=torch.randn(n, 512, generator=g),
                exps=torch.randn(n, 64, generator=g), coords=coords, target_com=torch.rand(H, W, 3, generator=g),
                target_head_neck=torch.rand(H, W, 3, generator=g), bc_img=torch.rand(H, W, 3, generator=g),
                z_shape=torch.randn(1, 2, 256, generator=g), z_app=torch.randn(1, 2, 256, generator=g))


def reference_select_coords(Hh, Ww, rect, N_rand, sample_rate):
    """MAIN:787-820 verbatim in behaviour (the reference's meshgrid call carries no `indexing`, i.e. 'ij')."""
    coords = torch.stack(torch.meshgrid(torch.linspace(0, Hh - 1, Hh), torch.linspace(0, Ww - 1, Ww)), -1)
    coords = torch.reshape(coords, [-1, 2])
    rect_inds = (coords[:, 0] >= rect[0]) & (coords[:, 0] <= rect[0] + rect[2]) & (coords[:, 1] >= rect[1]) & (coords[:, 1] <= rect[1] + rect[3])
    rect_torso = [1 * Hh / 2, 0, Hh / 2, Ww]
    rect_inds_torso = (coords[:, 0] >= rect_torso[0]) & (coords[:, 0] <= rect_torso[0] + rect_torso[2]) & \
        (coords[:, 1] >= rect_torso[1]) & (coords[:, 1] <= rect_torso[1] + rect_torso[3])
    rect_inds = rect_inds | rect_inds_torso
    coords_rect, coords_norect = coords[rect_inds], coords[~rect_inds]
    rect_num = int(N_rand * sample_rate)
    a = np.random.choice(coords_rect.shape[0], size=[rect_num], replace=False)
    b = np.random.choice(coords_norect.shape[0], size=[N_rand - rect_num], replace=False)
    return torch.cat((coords_rect[a].long(), coords_norect[b].long()), dim=0)


def reference_step(mods, opts, b):
    """One iteration of MAIN:764-931 through the reference's modules (global_step = 0 < nosmo_iters, noexp_iters = 0)."""
    decoder, AudNet, ExpNet = mods
    args = types.SimpleNamespace(nosmo_iters=10 ** 9, smo_size=8, smo_torse_size=4, dim_aud=64)
    dataset = [{'auds': b['auds'], 'exp': b['exps'], 'poses': b['poses']}]
    c = b['coords']
    near, far = b['near'] * torch.ones((N_RAND, 1)), b['far'] * torch.ones((N_RAND, 1))
    t_vals = torch.linspace(0., 1., steps=NS)
    z_vals = (near * (1. - t_vals) + far * t_vals).expand([N_RAND, NS])
    signal = MAIN.encode_signal(dataset, 0, b['img_i'], 64, AudNet, ExpNet, None, 0, args, 6)
    from oracle import nerf_oracle as O
    signal_torso = O.encode_signal_torso(b['poses'], b['img_i'])        # MAIN.rot_to_euler hard-codes .cuda()
    target_s_com = b['target_com'][c[:, 0], c[:, 1]]
    target_s_head_neck = b['target_head_neck'][c[:, 0], c[:, 1]]
    bc_rgb = b['bc_img'][c[:, 0], c[:, 1]]
    rays_o, rays_d = HELP.get_rays(b['H'], b['W'], b['focal'], b['pose'], b['cx'], b['cy'])
    rays_o, rays_d = rays_o[c[:, 0], c[:, 1]], rays_d[c[:, 0], c[:, 1]]
    p_i = (rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]).reshape(1, -1, 3)
    r_i = rays_d.unsqueeze(1).expand([N_RAND, NS, 3]).reshape(1, -1, 3)
    rays_o_t, rays_d_t = HELP.get_rays(b['H'], b['W'], b['focal'], b['pose_torso'], b['cx'], b['cy'])
    rays_o_t, rays_d_t = rays_o_t[c[:, 0], c[:, 1]], rays_d_t[c[:, 0], c[:, 1]]
    p_t = (rays_o_t[..., None, :] + rays_d_t[..., None, :] * z_vals[..., :, None]).reshape(1, -1, 3)
    r_t = rays_d_t.unsqueeze(1).expand([N_RAND, NS, 3]).reshape(1, -1, 3)
    feat_i, sigma_i = decoder(p_i, r_i, b['z_shape'][:, 0], b['z_app'][:, 0], signal, 'head')
    sigma_i = sigma_i.reshape(1, N_RAND, NS)
    feat_i = feat_i.reshape(1, N_RAND, NS, -1)
    feat_i = torch.cat((feat_i[..., :-1, :], bc_rgb.reshape(1, N_RAND, 1, 3)), dim=-2)
    feat_t, sigma_t = decoder(p_t, r_t, b['z_shape'][:, 1], b['z_app'][:, 1], signal_torso, 'torso')
    sigma_t = sigma_t.reshape(1, N_RAND, NS)
    feat_t = feat_t.reshape(1, N_RAND, NS, -1)
    sigma_t[:, :, -1] = 0
    # .clone(): MAIN:884-886 writes the +1e-6 into the ReLU outputs in place; torch 2.x autograd refuses that at
    # backward time ("modified by an inplace operation"), so the step as written does not run here -- same arithmetic
    sigma = F.relu(torch.stack([sigma_i], dim=0)).clone()
    sigma_torso = F.relu(torch.stack([sigma_i, sigma_t], dim=0)).clone()
    sigma[-1, :, :, -1] = sigma[-1, :, :, -1] + 1e-6
    sigma_torso[-1, :, :, -1] = sigma_torso[-1, :, :, -1] + 1e-6
    feat, feat_torso = torch.stack([feat_i], dim=0), torch.stack([feat_i, feat_t], dim=0)
    sigma_sum, feat_weighted = MAIN.composite_function(sigma, feat)
    sigma_torso_sum, feat_torso_weighted = MAIN.composite_function(sigma_torso, feat_torso)
    weights = MAIN.calc_volume_weights(z_vals.unsqueeze(0), rays_d.unsqueeze(0), sigma_sum, last_dist=1e10)
    weights_torso = MAIN.calc_volume_weights(z_vals.unsqueeze(0), rays_d_t.unsqueeze(0), sigma_torso_sum, last_dist=1e10)
    rgb_com = torch.sum(weights.unsqueeze(-1) * feat_weighted, dim=-2).squeeze(0)
    rgb_com_torso = torch.sum(weights_torso.unsqueeze(-1) * feat_torso_weighted, dim=-2).squeeze(0)
    loss = 0
    loss += HELP.img2mse(rgb_com_torso, target_s_com)
    loss += HELP.img2mse(rgb_com, target_s_head_neck)
    for o in opts:
        o.zero_grad()
    loss.backward()
    for o in opts:           # decoder, AudNet, ExpNet (noexp_iters = 0); the attention nets do not step before nosmo_iters
        o.step()
    return loss.detach()


def main():
    b = make_batch()
    np.random.seed(3)
    assert torch.equal(reference_select_coords(H, W, [5, 4, 9, 8], N_RAND, 0.95), b['coords']), 'select_coords != reference'
    sds = {'dec': synth.decoder_state_dict(6), 'aud': synth.mlp_encoder_state_dict(7), 'exp': synth.mlp_encoder_state_dict(8, (64, 32, 32))}
    # reference side
    dec = DEC.Decoder(z_dim=256, hidden_size=256, dim_signal=96, use_deformation_field=True)
    aud, exp = HELP.AudioNet_W2L(), HELP.ExpressionEnc()
    for m, k in ((dec, 'dec'), (aud, 'aud'), (exp, 'exp')):
        m.load_state_dict(sds[k])
    opts = [torch.optim.Adam(params=list(m.parameters()), lr=LRATE, betas=(0.9, 0.999)) for m in (dec, aud, exp)]
    # oracle side
    params = {k: {n: v.clone().requires_grad_(True) for n, v in sd.items()} for k, sd in sds.items()}
    order = {'dec': [n for n, _ in dec.named_parameters()], 'aud': [n for n, _ in aud.named_parameters()],
             'exp': [n for n, _ in exp.named_parameters()]}
    oopt = {k: torch.optim.Adam(params=[params[k][n] for n in order[k]], lr=LRATE, betas=(0.9, 0.999)) for k in params}
    out = {}
    for step in range(2):                      # two steps: the second one exercises Adam's moment state
        l_ref = reference_step((dec, aud, exp), opts, b)
        l_or = TO.train_step(params, b, oopt, global_step=step, noexp_iters=0, N_samples=NS)
        assert torch.equal(l_ref, l_or), ('loss', step, float(l_ref), float(l_or))
        n_used = 0
        for k, m in (('dec', dec), ('aud', aud), ('exp', exp)):
            for n, p in m.named_parameters():
                q = params[k][n]
                if p.grad is None:
                    assert q.grad is None or not q.grad.any(), (k, n)
                    continue
                n_used += 1
                assert torch.equal(p.grad, q.grad), ('grad', step, k, n, float((p.grad - q.grad).abs().max()))
                assert torch.equal(p.detach(), q.detach()), ('param', step, k, n)
        out['loss%d' % step] = l_ref.numpy()
        print('  ok  step %d  loss %.9f  (%d parameter tensors with gradients, all bit-equal)' % (step, float(l_ref), n_used))
        for k, n in (('dec', 'sigma_out.weight'), ('dec', 'blocks.6.weight'), ('dec', 'fc_in_torso.weight'), ('dec', 'feat_out.bias'),
                     ('dec', 'deform_net.out_embed.weight'), ('dec', 'fc_view.weight'), ('aud', 'encoder.0.weight'),
                     ('exp', 'encoder.2.bias')):
            g_, p_ = params[k][n].grad, params[k][n].detach()
            out['grad%d/%s/%s' % (step, k, n)] = g_.reshape(-1)[:64].numpy().copy()
            out['gradnorm%d/%s/%s' % (step, k, n)] = np.float64(g_.double().norm())
            out['param%d/%s/%s' % (step, k, n)] = p_.reshape(-1)[:64].numpy().copy()
    out['coords'] = b['coords'].numpy()
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'train_step.npz'), **out)
    print('wrote tests/golden/train_step.npz')


if __name__ == '__main__':
    main()
