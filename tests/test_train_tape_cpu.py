"""CPU: the LOGIC of the CUDA training step's tape (dfa-nerf_b200/train.py: which GEMM reads what with which strides, masks and
accumulation flags; per-frame latents as bias vectors; the deformation field's joined output; aliasing of additive skips)
with every libdfn kernel wrapper replaced by a few lines of torch doing what the kernel's contract says (include/dfn.h:
dfn_gemm, dfn_colsum, dfn_head_torso_loss_bwd, dfn_adam_step).  Loss, all gradients and two Adam steps must match the
training-step oracle (oracle/train_oracle.py, itself bit-equal to the reference's modules).  The kernels themselves are
checked on the GPU (tests/test_gpu_8_train.py)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import nerf_oracle as O
from oracle import synth
from oracle import train_oracle as TO

from test_train_oracle_golden import make_batch, H, W, N_RAND, NS, LRATE


def act_grad(mode, y):
    if mode == 1:
        return (y > 0).double()
    if mode == 2:
        return torch.where(y > 0, torch.ones_like(y), torch.full_like(y, 0.02)).double()
    if mode == 3:
        return (y * (1 - y)).double()
    return torch.ones_like(y).double()


def ref_mm(A, B, Cm, bias=None, addend=None, pre_add=False, act=0, mask=None, mask_mode=0, beta=0, k_splits=1, precision=None):
    a = A.double()
    if mask is not None and mask_mode:
        assert mask.shape == A.shape and mask.stride() == A.stride()
        a = a * act_grad(mask_mode, mask)
    x = a @ B.double().t()
    if bias is not None:
        x = x + bias.double()
    if addend is not None and pre_add:
        x = x + addend.double()
    if act == 1:
        x = F.relu(x)
    elif act == 2:
        x = torch.sigmoid(x)
    elif act == 3:
        x = F.leaky_relu(x, 0.02)
    if addend is not None and not pre_add:
        x = x + addend.double()
    assert k_splits == 1 or (act == 0 and beta == 1)
    if beta:
        x = x + Cm.double()
    Cm.copy_(x.float())
    return 1


def ref_colsum(X, Y, mode, out):
    x = X.double()
    if Y is not None and mode:
        x = x * act_grad(mode, Y)
    out.add_(x.sum(0).float())


def composite(feat_h, sig_h, feat_t, sig_t, bc, z, rd_h, rd_t, last_dist=1e10):
    """The second half of O.render_head_torso_chunk (MAIN:669-708) on the fields' outputs."""
    R, S = z.shape
    sig_h, sig_t = sig_h.reshape(1, R, S), sig_t.reshape(1, R, S).clone()
    feat_h, feat_t = feat_h.reshape(1, R, S, 3), feat_t.reshape(1, R, S, 3)
    feat_h = torch.cat((feat_h[..., :-1, :], bc.reshape(1, R, 1, 3)), dim=-2)
    sig_t[:, :, -1] = 0
    s1 = F.relu(torch.stack([sig_h], 0)).clone()
    s2 = F.relu(torch.stack([sig_h, sig_t], 0)).clone()
    s1[-1, :, :, -1] = s1[-1, :, :, -1] + 1e-6
    s2[-1, :, :, -1] = s2[-1, :, :, -1] + 1e-6
    ss1, fw1 = O.composite_function(s1, torch.stack([feat_h], 0))
    ss2, fw2 = O.composite_function(s2, torch.stack([feat_h, feat_t], 0))
    w1 = O.calc_volume_weights(z[None], rd_h[None], ss1, last_dist)
    w2 = O.calc_volume_weights(z[None], rd_t[None], ss2, last_dist)
    return torch.sum(w1.unsqueeze(-1) * fw1, dim=-2).squeeze(0), torch.sum(w2.unsqueeze(-1) * fw2, dim=-2).squeeze(0)


def ref_loss_bwd(R, S, feat_h, sig_h, feat_t, sig_t, bc, z, rd_h, rd_t, tgt_head, tgt_person, loss2, rgb_head, rgb_person, g_feat_h,
                 g_sig_h, g_feat_t, g_sig_t, last_dist=1e10):
    with torch.enable_grad():
        leaves = [t.detach().double().requires_grad_(True) for t in (feat_h, sig_h, feat_t, sig_t)]
        rh, rp = composite(leaves[0], leaves[1], leaves[2], leaves[3], bc.double(), z.double(), rd_h.double(), rd_t.double(), last_dist)
        l0, l1 = torch.mean((rh - tgt_head.double()) ** 2), torch.mean((rp - tgt_person.double()) ** 2)
        (l0 + l1).backward()
    loss2.add_(torch.stack([l0, l1]).detach().float())
    rgb_head.copy_(rh.detach().float())
    rgb_person.copy_(rp.detach().float())
    fh, ft = feat_h.double(), feat_t.double()
    g_feat_h.copy_((leaves[0].grad.reshape(fh.shape) * fh * (1 - fh)).float())
    g_feat_t.copy_((leaves[2].grad.reshape(ft.shape) * ft * (1 - ft)).float())
    g_sig_h.copy_(leaves[1].grad.reshape(g_sig_h.shape).float())
    g_sig_t.copy_(leaves[3].grad.reshape(g_sig_t.shape).float())


def ref_adam(flat, grad, m, v, lr, betas, eps, step):
    b1, b2 = betas
    m.lerp_(grad, 1 - b1)
    v.mul_(b2).addcmul_(grad, grad, value=1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = (v.sqrt() / np.sqrt(bc2)).add_(eps)
    flat.addcdiv_(m, denom, value=-lr / bc1)


@pytest.fixture
def cpu_train(monkeypatch):
    from dfa_nerf_b200 import train
    monkeypatch.setattr(train, 'mm', ref_mm)
    monkeypatch.setattr(train, 'colsum', ref_colsum)
    monkeypatch.setattr(train, 'loss_bwd', ref_loss_bwd)
    monkeypatch.setattr(train, 'adam_step', ref_adam)
    monkeypatch.setattr(train, 'get_rays', lambda H, W, focal, c2w, cx, cy, device=None: O.get_rays(H, W, focal, c2w, cx, cy))
    monkeypatch.setattr(train, 'z_vals_uniform', lambda near, far, n: O.z_vals_uniform(near[:, None], far[:, None], n))
    monkeypatch.setattr(train, 'make_points', lambda o, d, z: (o[:, None, :] + d[:, None, :] * z[:, :, None],
                                                               d[:, None, :].expand(z.shape[0], z.shape[1], 3)))

    def tp(p, n_freq, views=False, normalize=False):
        if normalize:
            p = p / torch.norm(p, dim=-1, keepdim=True)
        return O.decoder_transform_points(p, n_freq).contiguous()
    monkeypatch.setattr(train, 'decoder_transform_points', tp)
    monkeypatch.setattr(train, 'encode_signal_torso_sequence', lambda poses: O.encode_signal_torso(poses, 0))     # poses [1,3,4]
    return train


def test_tape_reproduces_the_reference_training_step(cpu_train):
    import dfa_nerf_b200 as dfn
    gold = np.load('tests/golden/train_step.npz') if False else None
    b = make_batch()
    sds = {'dec': synth.decoder_state_dict(6), 'aud': synth.mlp_encoder_state_dict(7), 'exp': synth.mlp_encoder_state_dict(8, (64, 32, 32))}
    # oracle side (autograd)
    params = {k: {n: v.clone().requires_grad_(True) for n, v in sd.items()} for k, sd in sds.items()}
    opt = {k: torch.optim.Adam(params=list(params[k].values()), lr=LRATE, betas=(0.9, 0.999)) for k in params}
    # product side (explicit tape, kernels emulated in torch)
    dec = dfn.Decoder(z_dim=256, hidden_size=256, dim_signal=96, use_deformation_field=True)
    dec.load_state_dict(sds['dec'])
    aud, exp = dfn.AudioNet_W2L(), dfn.ExpressionEnc()
    aud.load_state_dict(sds['aud'])
    exp.load_state_dict(sds['exp'])
    tr = cpu_train.Trainer(dec, aud, exp, lrate=LRATE, N_samples=NS, device='cpu')
    for step in range(2):
        loss_ref = TO.train_step(params, b, opt, global_step=step, noexp_iters=0, N_samples=NS)
        loss = tr.step(b, global_step=step, noexp_iters=0)
        assert abs(float(loss) - float(loss_ref)) < 2e-6 * max(1., abs(float(loss_ref))), (step, float(loss), float(loss_ref))
        n_checked = 0
        for k in params:
            for n, q in params[k].items():
                g = tr.grads[k][n]
                if q.grad is None:
                    assert not g.any(), (k, n)
                    continue
                scale = q.grad.abs().max().item()
                err = (g - q.grad).abs().max().item()
                assert err <= 2e-5 * scale + 1e-10, (step, k, n, err, scale)
                n_checked += 1
                # parameters after the Adam step: in units of the learning rate
                assert (tr.params[k][n] - q.detach()).abs().max().item() <= 0.02 * LRATE, (step, k, n)
        assert n_checked == 74
    assert tr.last_launches > 100


def test_select_coords_matches_the_oracle():
    from dfa_nerf_b200.train import select_coords
    for rate in (0.95, 0.):
        np.random.seed(3)
        a = select_coords(H, W, [5, 4, 9, 8], N_RAND, rate)
        np.random.seed(3)
        bq = TO.select_coords(H, W, [5, 4, 9, 8], N_RAND, rate)
        assert torch.equal(a, bq)
