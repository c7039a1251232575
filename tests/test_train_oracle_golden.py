"""CPU: the training-step oracle (oracle/train_oracle.py; SURVEY 8f-3, MAIN:738-931) reproduces what the REFERENCE's own
modules produced (oracle/make_golden_train.py asserted loss, all 74 gradients and all updated parameters bit-equal over
two Adam steps in the build container): loss, gradient norms / slices and updated-parameter slices of a few tensors from
the committed fixture, with a small tolerance for a different host BLAS.  The CUDA training step is not built yet
(DESIGN.md section 8); this is the oracle it will be checked against."""
import os

import numpy as np
import torch

from oracle import synth
from oracle import train_oracle as TO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'train_step.npz')
H, W, N_RAND, NS, LRATE = 24, 20, 96, 16, 5e-4          # as oracle/make_golden_train.py


def make_batch():
    g = torch.Generator().manual_seed(11)
    n = 6
    fr = synth.frame_inputs(H=H, W=W, seed=2)
    poses = torch.cat([synth.pose_sequence(n, 7), torch.tensor([0., 0., 0., 1.]).expand(n, 1, 4)], 1)
    np.random.seed(3)
    coords = TO.select_coords(H, W, [5, 4, 9, 8], N_RAND, 0.95)
    return dict(H=H, W=W, focal=fr['focal'], cx=fr['cx'], cy=fr['cy'], near=fr['near'], far=fr['far'], poses=poses, img_i=2,
                pose=poses[2, :3, :4], pose_torso=poses[0, :3, :4], auds=torch.randn(n, 512, generator=g),
                exps=torch.randn(n, 64, generator=g), coords=coords, target_com=torch.rand(H, W, 3, generator=g),
                target_head_neck=torch.rand(H, W, 3, generator=g), bc_img=torch.rand(H, W, 3, generator=g),
                z_shape=torch.randn(1, 2, 256, generator=g), z_app=torch.randn(1, 2, 256, generator=g))


def test_select_coords_matches_reference_choice():
    gold = np.load(GOLD)
    b = make_batch()
    assert np.array_equal(b['coords'].numpy(), gold['coords'])            # the reference's two np.random.choice calls, seed 3
    c = b['coords']
    inside = ((c[:, 0] >= 5) & (c[:, 0] <= 14) & (c[:, 1] >= 4) & (c[:, 1] <= 12)) | (c[:, 0] >= H / 2)
    assert int(inside.sum()) == int(N_RAND * 0.95)                        # 95 % of the rays inside face rect | lower half
    assert len({(int(r), int(q)) for r, q in c}) == N_RAND                # without replacement
    np.random.seed(3)
    u = TO.select_coords(H, W, None, N_RAND, 0)                           # train_obama.sh: --sample_rate=0, uniform
    assert u.shape == (N_RAND, 2) and len({(int(r), int(q)) for r, q in u}) == N_RAND


def test_two_adam_steps_match_the_reference():
    gold = np.load(GOLD)
    b = make_batch()
    sds = {'dec': synth.decoder_state_dict(6), 'aud': synth.mlp_encoder_state_dict(7), 'exp': synth.mlp_encoder_state_dict(8, (64, 32, 32))}
    params = {k: {n: v.clone().requires_grad_(True) for n, v in sd.items()} for k, sd in sds.items()}
    opt = {k: torch.optim.Adam(params=list(params[k].values()), lr=LRATE, betas=(0.9, 0.999)) for k in params}
    picks = [k for k in gold.files if k.startswith('grad0/')]
    assert len(picks) == 8
    for step in range(2):
        loss = TO.train_step(params, b, opt, global_step=step, noexp_iters=0, N_samples=NS)
        assert abs(float(loss) - float(gold['loss%d' % step])) < 1e-6
        for key in picks:
            _, k, n = key.split('/')
            g = params[k][n].grad
            gn = float(gold['gradnorm%d/%s/%s' % (step, k, n)])
            assert abs(float(g.double().norm()) - gn) <= 1e-4 * gn + 1e-9, (step, k, n)
            ref = gold['grad%d/%s/%s' % (step, k, n)]
            assert np.abs(g.reshape(-1)[:64].numpy() - ref).max() <= 1e-4 * np.abs(ref).max() + 1e-9, (step, k, n)
            # Adam's first steps move every touched weight by ~lr whatever the gradient's size: compare loosely in lr units
            pref = gold['param%d/%s/%s' % (step, k, n)]
            assert np.abs(params[k][n].detach().reshape(-1)[:64].numpy() - pref).max() <= 0.05 * LRATE, (step, k, n)
    # untouched heads stay untouched: the listener / unused inputs get no gradient (the reference leaves them at None)
    assert params['dec']['fc_in_listener.weight'].grad is None


def test_loss_pieces():
    b = make_batch()
    sds = {'dec': synth.decoder_state_dict(6), 'aud': synth.mlp_encoder_state_dict(7), 'exp': synth.mlp_encoder_state_dict(8, (64, 32, 32))}
    with torch.no_grad():
        loss, l_com, l_head = TO.train_losses(sds['dec'], sds['aud'], sds['exp'], b, NS)
    assert float(loss) == float(l_com + l_head) and float(l_com) > 0 and float(l_head) > 0
    assert abs(float(TO.mse2psnr(torch.tensor(0.01))) - 20.0) < 1e-5       # HELP:14
