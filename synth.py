"""Seeded synthetic inputs of the real shapes.  Shared by the
tests, bench.py, profiles/ and smoke(); it generates *data* only and imports neither the oracle
arithmetic (oracle/) nor the product, so both sides can use it.

No trained checkpoint or processed dataset ships with the reference, so
every "Obama" configuration uses these tensors.
"""
import math

import numpy as np
import torch


def _uniform(gen, shape, bound):
    return (torch.rand(shape, generator=gen) * 2 - 1) * bound


def _linear(gen, sd, name, fan_in, fan_out):
    # nn.Linear default: kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for both
    b = 1. / math.sqrt(fan_in)
    sd[name + '.weight'] = _uniform(gen, (fan_out, fan_in), b)
    sd[name + '.bias'] = _uniform(gen, (fan_out,), b)


# Density-head calibration per (model, seed): gain and bias that give sigma ~ mean 2, std 12
# over the synthetic camera frustum (half-occupied field with sharp surfaces), measured once
# with the oracle and frozen here so that every machine builds bit-identical weights.
# Default-init heads give sigma ~ 1e-2 and an all-background image (SURVEY probe C.5).
_SIGMA_CAL = {
    ('face', 0): (1431., -38.52), ('face', 1): (1370., 17.16), ('face', 2): (1641., 35.00), ('face', 3): (1660., -50.78),
    ('nerf', 0): (1820., 18.58), ('nerf', 1): (1614., 25.38), ('nerf', 2): (1988., -44.84), ('nerf', 3): (1672., -33.81),
}


def _sigma_cal(kind, seed, gain, bias):
    g, b = _SIGMA_CAL.get((kind, seed), (1500., 2.0))
    return (g if gain is None else gain), (b if bias is None else bias)


def facenerf_state_dict(seed=0, D=8, W=256, input_ch=63, input_ch_views=27, dim_aud=64,
                        skips=(4,), sigma_gain=None, sigma_bias=None):
    """FaceNeRF weights with the reference's parameter names (HELP:257-273).  The density
    head is scaled (SURVEY section 8d / probe C.5) so compositing and sample_pdf see a real field."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    n_in = input_ch + dim_aud
    for i in range(D):
        k = n_in if i == 0 else (W + n_in if (i - 1) in skips else W)
        _linear(g, sd, 'pts_linears.%d' % i, k, W)
    _linear(g, sd, 'views_linears.0', input_ch_views + W, W // 2)
    for i in range(D // 4):
        _linear(g, sd, 'views_linears.%d' % (i + 1), W // 2, W // 2)
    _linear(g, sd, 'feature_linear', W, W)
    _linear(g, sd, 'alpha_linear', W, 1)
    _linear(g, sd, 'rgb_linear', W // 2, 3)
    sigma_gain, sigma_bias = _sigma_cal('face', seed, sigma_gain, sigma_bias)
    sd['alpha_linear.weight'] = sd['alpha_linear.weight'] * sigma_gain
    sd['alpha_linear.bias'] = torch.full((1,), float(sigma_bias))
    return sd


def nerf_state_dict(seed=0, D=8, W=256, input_ch=63, input_ch_views=27, skips=(4,),
                    sigma_gain=None, sigma_bias=None):
    """NeRF weights (HELP:354-370)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for i in range(D):
        k = input_ch if i == 0 else (W + input_ch if (i - 1) in skips else W)
        _linear(g, sd, 'pts_linears.%d' % i, k, W)
    _linear(g, sd, 'views_linears.0', input_ch_views + W, W // 2)
    _linear(g, sd, 'feature_linear', W, W)
    _linear(g, sd, 'alpha_linear', W, 1)
    _linear(g, sd, 'rgb_linear', W // 2, 3)
    sigma_gain, sigma_bias = _sigma_cal('nerf', seed, sigma_gain, sigma_bias)
    sd['alpha_linear.weight'] = sd['alpha_linear.weight'] * sigma_gain
    sd['alpha_linear.bias'] = torch.full((1,), float(sigma_bias))
    return sd


def decoder_state_dict(seed=0, hidden=256, z_dim=256, dim_signal=96, dim_et=42, n_freq=10,
                       n_freq_views=4, n_blocks=8, sigma_gain=400., sigma_bias=2.0):
    """Decoder + DeformationField_ori weights with the reference's names (DEC:166-255, DEC:78-105)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    de, dv = 6 * n_freq, 6 * n_freq_views
    # deform_net (dim_embed=de, dim_signal=dim_et, hidden 64, n_blocks 7)
    for br, od in (('embed', de), ('signal', dim_et)):
        _linear(g, sd, 'deform_net.blocks_%s.0' % br, de + dim_et, 64)
        for i in range(1, 5):
            _linear(g, sd, 'deform_net.blocks_%s.%d' % (br, i), 64, 64)
        _linear(g, sd, 'deform_net.out_%s' % br, 64, od)
    _linear(g, sd, 'deform_net.fc_embed_skips.0', de, 64)
    _linear(g, sd, 'deform_net.fc_signal_skips.0', dim_et, 64)
    _linear(g, sd, 'fc_in', de + dim_signal, hidden)
    _linear(g, sd, 'fc_in_listener', de, hidden)
    _linear(g, sd, 'fc_in_torso', de + dim_et, hidden)
    _linear(g, sd, 'fc_z', z_dim, hidden)
    for i in range(n_blocks - 1):
        _linear(g, sd, 'blocks.%d' % i, hidden, hidden)
    _linear(g, sd, 'fc_z_skips.0', z_dim, hidden)
    _linear(g, sd, 'fc_p_skips.0', de + dim_signal, hidden)
    _linear(g, sd, 'fc_p_skips_listener.0', de, hidden)
    _linear(g, sd, 'fc_p_skips_torso.0', de + dim_et, hidden)
    _linear(g, sd, 'sigma_out', hidden, 1)
    _linear(g, sd, 'fc_z_view', z_dim, hidden)
    _linear(g, sd, 'feat_view', hidden, hidden)
    _linear(g, sd, 'fc_view', dv, hidden)
    _linear(g, sd, 'feat_out', hidden, 3)
    sd['sigma_out.weight'] = sd['sigma_out.weight'] * sigma_gain
    sd['sigma_out.bias'] = torch.full((1,), float(sigma_bias))
    return sd


def _conv1d(gen, sd, name, c_in, c_out, k=3):
    b = 1. / math.sqrt(c_in * k)       # nn.Conv1d default init
    sd[name + '.weight'] = _uniform(gen, (c_out, c_in, k), b)
    sd[name + '.bias'] = _uniform(gen, (c_out,), b)


def audionet_state_dict(seed=0, dim_aud=76):
    """HELP:109-132 AudioNet."""
    g, sd = torch.Generator().manual_seed(seed), {}
    for i, (ci, co) in zip((0, 2, 4, 6), ((29, 32), (32, 32), (32, 64), (64, 64))):
        _conv1d(g, sd, 'encoder_conv.%d' % i, ci, co)
    _linear(g, sd, 'encoder_fc1.0', 64, 64)
    _linear(g, sd, 'encoder_fc1.2', 64, dim_aud)
    return sd


def mlp_encoder_state_dict(seed=0, dims=(512, 256, 128, 64)):
    """HELP:165-178 AudioNet_W2L (default dims) / HELP:182-193 ExpressionEnc (dims=(64, 32, 32))."""
    g, sd = torch.Generator().manual_seed(seed), {}
    for j in range(len(dims) - 1):
        _linear(g, sd, 'encoder.%d' % (2 * j), dims[j], dims[j + 1])
    return sd


def audio_att_state_dict(seed=0, dim_aud=64, seq_len=8):
    """HELP:210-231 AudioAttNet."""
    g, sd = torch.Generator().manual_seed(seed), {}
    ch = (dim_aud, 16, 8, 4, 2, 1)
    for j in range(5):
        _conv1d(g, sd, 'attentionConvNet.%d' % (2 * j), ch[j], ch[j + 1])
    _linear(g, sd, 'attentionNet.0', seq_len, seq_len)
    return sd


def pose_sequence(n, seed=0):
    """[n,4,4] head poses with the synthetic camera statistics (camera_pose per frame)."""
    out = torch.zeros(n, 4, 4)
    for i in range(n):
        out[i, :3] = camera_pose(seed + i)
        out[i, 3, 3] = 1.
    return out


def euler_to_rot(e):
    """Rx(theta) Ry(phi) Rz(psi) rotation for synthetic head poses."""
    t, p, s = [float(v) for v in e]
    rx = np.array([[1, 0, 0], [0, math.cos(t), -math.sin(t)], [0, math.sin(t), math.cos(t)]])
    ry = np.array([[math.cos(p), 0, math.sin(p)], [0, 1, 0], [-math.sin(p), 0, math.cos(p)]])
    rz = np.array([[math.cos(s), -math.sin(s), 0], [math.sin(s), math.cos(s), 0], [0, 0, 1]])
    return rx @ ry @ rz


def camera_pose(seed=0):
    """c2w [3,4] = [R^T | -R^T t] (data_util/process_data_ba.py:402-423) with small head
    rotation and t ~ (0,0,-0.6)."""
    rng = np.random.RandomState(seed)
    e = rng.uniform(-0.15, 0.15, size=3)
    t = np.array([rng.uniform(-.03, .03), rng.uniform(-.03, .03), -0.6])
    R = euler_to_rot(e)
    c2w = np.concatenate([R.T, (-R.T @ t)[:, None]], axis=1)
    return torch.from_numpy(c2w.astype(np.float32))


def frame_inputs(H=450, W=450, seed=0, dim_aud=64, n_frames=1):
    """Camera intrinsics + pose + latent + background for one frame (or a sequence)."""
    focal = 1200.0 * W / 450.0
    g = torch.Generator().manual_seed(1000 + seed)
    out = dict(H=H, W=W, focal=focal, cx=W / 2.0, cy=H / 2.0, near=0.4, far=1.0,
               c2w=camera_pose(seed),
               aud=torch.randn((n_frames, dim_aud), generator=g),
               bc_rgb=torch.rand((H * W, 3), generator=g))
    if n_frames == 1:
        out['aud'] = out['aud'][0]
    else:
        out['c2w_seq'] = torch.stack([camera_pose(seed + i) for i in range(n_frames)], 0)
    return out
