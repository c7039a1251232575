"""CPU restatement of the reference's TRAINING step (MAIN:738-931) -- test infrastructure only.

The backward pass is not built yet; this file and tests/golden/train_step.npz are the oracle the
CUDA training step will be checked against.  Everything is plain differentiable torch on top of oracle/nerf_oracle.py, so
`torch.autograd` gives the reference gradients.  Pinned: oracle/make_golden_train.py runs the same step through the
reference's own modules and functions (decoder.Decoder, AudioNet_W2L, ExpressionEnc, encode_signal, get_rays,
composite_function, calc_volume_weights, img2mse, torch.optim.Adam as configured at MAIN:522-535) and demands the loss,
every gradient and every updated parameter to be bit-equal.

  select_coords      MAIN:787-820   pixel choice: sample_rate of the N_rand rays inside face rect | lower image half
  train_losses       MAIN:822-907   head + torso decoder on the chosen rays, two-field compositing, the two MSE losses
  train_step         MAIN:909-931   zero_grad / backward / Adam steps (global_step < nosmo_iters: decoder + AudNet;
                                    ExpNet steps once global_step >= noexp_iters)
"""
import numpy as np
import torch

from . import nerf_oracle as O


def select_coords(H, W, rect, N_rand, sample_rate, rng=np.random):
    """MAIN:787-820.  rect = (x, y, w, h) of the face in ROW/COLUMN order as the reference indexes it (coords[:, 0] is the
    row).  Consumes the global numpy RNG exactly like the reference (two choice() calls, or one when sample_rate == 0)."""
    coords = torch.stack(torch.meshgrid(torch.linspace(0, H - 1, H), torch.linspace(0, W - 1, W), indexing='ij'), -1).reshape(-1, 2)
    if sample_rate > 0:
        def inside(r):
            return (coords[:, 0] >= r[0]) & (coords[:, 0] <= r[0] + r[2]) & (coords[:, 1] >= r[1]) & (coords[:, 1] <= r[1] + r[3])
        rect_inds = inside(rect) | inside([1 * H / 2, 0, H / 2, W])
        coords_rect, coords_norect = coords[rect_inds], coords[~rect_inds]
        rect_num = int(N_rand * sample_rate)
        sel_rect = rng.choice(coords_rect.shape[0], size=[rect_num], replace=False)
        sel_norect = rng.choice(coords_norect.shape[0], size=[N_rand - rect_num], replace=False)
        return torch.cat((coords_rect[sel_rect].long(), coords_norect[sel_norect].long()), dim=0)
    return coords[rng.choice(coords.shape[0], size=[N_rand], replace=False)].long()


def img2mse(x, y):
    """HELP:11."""
    return torch.mean((x - y) ** 2)


def mse2psnr(x):
    """HELP:14."""
    return -10. * torch.log(x) / torch.log(torch.Tensor([10.]))


def train_losses(sd_dec, sd_aud, sd_exp, batch, N_samples=64, last_dist=1e10):
    """MAIN:764-907 for one object, global_step < nosmo_iters.  batch: H, W, focal, cx, cy, near, far, pose [3,4], pose_torso
    [3,4], poses [N,4,4] (for the torso signal), img_i, auds [N,512], exps [N,64], coords [N_rand,2] (long), target_com /
    target_head_neck / bc_img [H,W,3], z_shape / z_app [1,2,z_dim].  Returns (loss, img_loss_com, img_loss_head_neck)."""
    H, W, c = batch['H'], batch['W'], batch['coords']
    n = c.shape[0]
    z_vals = O.z_vals_uniform(torch.full((n, 1), float(batch['near'])), torch.full((n, 1), float(batch['far'])), N_samples)
    signal = O.encode_signal(batch['auds'], batch['exps'], batch['img_i'], sd_aud, sd_exp)
    signal_torso = O.encode_signal_torso(batch['poses'], batch['img_i'])
    pick = lambda img: img[c[:, 0], c[:, 1]]                                                  # noqa: E731
    ro, rd = [pick(t) for t in O.get_rays(H, W, batch['focal'], batch['pose'], batch['cx'], batch['cy'])]
    rot, rdt = [pick(t) for t in O.get_rays(H, W, batch['focal'], batch['pose_torso'], batch['cx'], batch['cy'])]
    rgb_com, rgb_com_torso = O.render_head_torso_chunk(sd_dec, ro, rd, rot, rdt, z_vals, pick(batch['bc_img']), batch['z_shape'],
                                                       batch['z_app'], signal, signal_torso, last_dist)
    img_loss_head_neck = img2mse(rgb_com, pick(batch['target_head_neck']))
    img_loss_com = img2mse(rgb_com_torso, pick(batch['target_com']))
    loss = 0
    loss += img_loss_com
    loss += img_loss_head_neck
    return loss, img_loss_com, img_loss_head_neck


def train_step(params, batch, opt, global_step=0, noexp_iters=0, N_samples=64):
    """MAIN:909-931.  params = {'dec': sd, 'aud': sd, 'exp': sd} of leaf tensors with requires_grad; opt = the matching dict
    of torch.optim.Adam (lr = lrate, betas (0.9, 0.999)).  Returns the loss; gradients stay in the .grad fields."""
    for o in opt.values():
        o.zero_grad()
    loss, _, _ = train_losses(params['dec'], params['aud'], params['exp'], batch, N_samples)
    loss.backward()
    opt['dec'].step()
    opt['aud'].step()
    if global_step >= noexp_iters:
        opt['exp'].step()
    return loss.detach()
