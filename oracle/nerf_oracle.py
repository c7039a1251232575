"""CPU oracle for the DFA-NeRF volume-rendering hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain PyTorch CPU fp32, the arithmetic of the reference's
per-ray path so that the CUDA kernels can be checked against it on machines where
``/root/reference`` is not mounted (the GPU box).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it; the product package never does.

Pinning status: the reference ships no tests, golden vectors or fixtures for this
path, so parity is *unpinned upstream*.  It is pinned here by
``oracle/make_golden.py``, which imports the real reference modules from
``/root/reference/NeRFs/DFANeRF`` in the build container, checks every function
below against them bit-for-bit on seeded inputs, and writes the vectors committed
under ``tests/golden/``.

Reference citations (relative to /root/reference/NeRFs/DFANeRF):
  HELP = run_nerf_helpers.py, DEC = decoder.py, MAIN = run_nerf_com_trainExpLater.py
The upstream-convention glue (render / batchify_rays / render_rays / raw2outputs /
run_network) is absent from the reference; it is
assembled here from the reference's own pieces in the AD-NeRF order.

Models are evaluated functionally from a ``state_dict`` whose keys equal the
reference's nn.Module keys, so one set of weights feeds reference, oracle and CUDA.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------- rays


def get_rays(H, W, focal, c2w, cx=None, cy=None, stride=1):
    """Pinhole rays.  Follows HELP:449-465.

    Returns (rays_o, rays_d), each [H//stride, W//stride, 3]; rays_d = R @ dir,
    NOT normalised; pixel (row y, col x) has dir ((x-cx)/f, -(y-cy)/f, -1).
    """
    xs = torch.linspace(0, W - 1, W // stride)
    ys = torch.linspace(0, H - 1, H // stride)
    if cx is None:
        cx = W * .5
    if cy is None:
        cy = H * .5
    col = xs[None, :].expand(ys.numel(), xs.numel())
    row = ys[:, None].expand(ys.numel(), xs.numel())
    cam = torch.stack([(col - cx) / focal, -(row - cy) / focal, -torch.ones_like(col)], -1)
    rot = c2w[:3, :3]
    rays_d = (cam[..., None, :] * rot).sum(-1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def linspace_table(n):
    """torch.linspace(0, 1, n) fp32 -- the t / u table (MAIN:617, HELP:546)."""
    return torch.linspace(0., 1., steps=n)


def z_vals_uniform(near, far, n_samples):
    """MAIN:617-619: z = near*(1-t) + far*t with t = linspace(0,1,S).  near/far: [R,1]."""
    t = linspace_table(n_samples)
    return near * (1. - t) + far * t


def z_vals_stratified(z_vals, rand):
    """Upstream render_rays jitter (north-star a2): lower + (upper-lower)*rand."""
    mids = .5 * (z_vals[..., 1:] + z_vals[..., :-1])
    upper = torch.cat([mids, z_vals[..., -1:]], -1)
    lower = torch.cat([z_vals[..., :1], mids], -1)
    return lower + (upper - lower) * rand


# ----------------------------------------------------------------- positional enc


def embed(x, multires, include_input=True):
    """HELP:21-70: [x, sin(2^0 x), cos(2^0 x), sin(2^1 x), ...]; no pi factor."""
    parts = [x] if include_input else []
    for k in range(multires):
        f = 2. ** torch.linspace(0., multires - 1, steps=multires)[k]
        parts.append(torch.sin(x * f))
        parts.append(torch.cos(x * f))
    return torch.cat(parts, -1)


def embed_dim(multires, include_input=True):
    return 3 * (2 * multires + (1 if include_input else 0))


def decoder_transform_points(p, n_freq, downscale_p_by=2.):
    """DEC:257-275 ('normal' encoding): p/=2; cat_k [sin(2^k pi p), cos(2^k pi p)]."""
    p = p / downscale_p_by
    return torch.cat([torch.cat([torch.sin((2 ** k) * math.pi * p),
                                 torch.cos((2 ** k) * math.pi * p)], dim=-1)
                      for k in range(n_freq)], dim=-1)


# ------------------------------------------------------------------------- models


def _lin(sd, name, x):
    return F.linear(x, sd[name + '.weight'], sd[name + '.bias'])


def facenerf_forward(sd, x, input_ch=63, input_ch_views=27, dim_aud=64, D=8, skips=(4,)):
    """HELP:275-299 with use_viewdirs=True.  x: [P, input_ch+dim_aud+input_ch_views] -> [P,4].

    feature_linear is constructed but bypassed (HELP:287); three view layers
    (1 + D//4, HELP:265-266); output [rgb(3), alpha(1)], no output activation.
    """
    pts, views = torch.split(x, [input_ch + dim_aud, input_ch_views], dim=-1)
    h = pts
    for i in range(D):
        h = F.relu(_lin(sd, 'pts_linears.%d' % i, h))
        if i in skips:
            h = torch.cat([pts, h], -1)
    alpha = _lin(sd, 'alpha_linear', h)
    h = torch.cat([h, views], -1)
    n_view = 1 + D // 4
    for i in range(n_view):
        h = F.relu(_lin(sd, 'views_linears.%d' % i, h))
    rgb = _lin(sd, 'rgb_linear', h)
    return torch.cat([rgb, alpha], -1)


def nerf_forward(sd, x, input_ch=63, input_ch_views=27, D=8, skips=(4,)):
    """HELP:372-396 with use_viewdirs=True: feature_linear applied, one view layer."""
    pts, views = torch.split(x, [input_ch, input_ch_views], dim=-1)
    h = pts
    for i in range(D):
        h = F.relu(_lin(sd, 'pts_linears.%d' % i, h))
        if i in skips:
            h = torch.cat([pts, h], -1)
    alpha = _lin(sd, 'alpha_linear', h)
    h = torch.cat([_lin(sd, 'feature_linear', h), views], -1)
    h = F.relu(_lin(sd, 'views_linears.0', h))
    rgb = _lin(sd, 'rgb_linear', h)
    return torch.cat([rgb, alpha], -1)


def deformation_forward(sd, x, dim_embed=60, dim_signal=42, skips=(4,), prefix='deform_net.'):
    """DEC:109-134 (DeformationField_ori): two 5-layer branches on [PE | signal]."""
    emb, sig = x[..., :dim_embed], x[..., -dim_signal:]
    outs = []
    for branch, skip_name, skip_in, out_name in (
            ('blocks_embed', 'fc_embed_skips', emb, 'out_embed'),
            ('blocks_signal', 'fc_signal_skips', sig, 'out_signal')):
        n = 0
        while prefix + '%s.%d.weight' % (branch, n) in sd:
            n += 1
        net, s = x, 0
        for idx in range(n):
            net = F.relu(_lin(sd, prefix + '%s.%d' % (branch, idx), net))
            if (idx + 1) in skips and idx < n - 1:
                net = net + _lin(sd, prefix + '%s.%d' % (skip_name, s), skip_in)
                s += 1
        outs.append(_lin(sd, prefix + out_name, net))
    return torch.cat(outs, -1)


def decoder_forward(sd, p_in, ray_d, z_shape, z_app, signal, head_or_torso,
                    n_freq=10, n_freq_views=4, skips=(4,), n_blocks=8, expression=None, use_deformation_field=True,
                    final_sigmoid=True):
    """DEC:277-349 (live reference model).  p_in, ray_d: [1,P,3]; z_*: [1,z_dim];
    signal: [1,dim_signal] (head) or [1,dim_et_embed] (torso).  Returns feat [1,P,3] (sigmoid inside),
    sigma [1,P] (no relu inside).
    The options the live script leaves at their defaults: signal None on the head selects the listener layers (DEC:305-308,
    322-325); expression [1,dim_exp] adds expnet(expression) to the view layer's input (DEC:279-281, 333-334; the caller
    passes it only when the module was built with use_expression); ray_d None skips the view term AND the relu (DEC:336-340);
    several skips use fc_*_skips.{0,1,..} in turn (DEC:314-327)."""
    p = decoder_transform_points(p_in, n_freq)
    if signal is not None:
        p = torch.cat((p, signal.expand(p.shape[1], -1).unsqueeze(0)), -1)
    if head_or_torso == 'torso':
        if use_deformation_field:
            p = deformation_forward(sd, p, dim_embed=6 * n_freq, dim_signal=signal.shape[-1]) + p
        net = _lin(sd, 'fc_in_torso', p)
        pskip = 'fc_p_skips_torso.%d'
    elif head_or_torso == 'head':
        net = _lin(sd, 'fc_in' if signal is not None else 'fc_in_listener', p)
        pskip = 'fc_p_skips.%d' if signal is not None else 'fc_p_skips_listener.%d'
    else:
        raise ValueError('head_or_torso')
    net = F.relu(net + _lin(sd, 'fc_z', z_shape).unsqueeze(1))
    skip_idx = 0
    for idx in range(n_blocks - 1):
        net = F.relu(_lin(sd, 'blocks.%d' % idx, net))
        if (idx + 1) in skips and idx < n_blocks - 2:
            net = net + _lin(sd, 'fc_z_skips.%d' % skip_idx, z_shape).unsqueeze(1)
            net = net + _lin(sd, pskip % skip_idx, p)
            skip_idx += 1
    sigma = _lin(sd, 'sigma_out', net).squeeze(-1)
    net = _lin(sd, 'feat_view', net) + _lin(sd, 'fc_z_view', z_app).unsqueeze(1)
    if expression is not None:
        net = net + _lin(sd, 'expnet', expression)
    if ray_d is not None:
        d = ray_d / torch.norm(ray_d, dim=-1, keepdim=True)
        net = F.relu(net + _lin(sd, 'fc_view', decoder_transform_points(d, n_freq_views)))
    feat = _lin(sd, 'feat_out', net)
    if final_sigmoid:
        feat = torch.sigmoid(feat)
    return feat, sigma


# -------------------------------------------------------------------- compositing


def composite_function(sigma, feat):
    """MAIN:146-166: density-weighted mix of n_box fields."""
    if sigma.shape[0] == 1:
        return sigma.squeeze(0), feat.squeeze(0)
    den = torch.sum(sigma, dim=0, keepdim=True)
    den[den == 0] = 1e-4
    w = sigma / den
    return torch.sum(sigma, dim=0), (feat * w.unsqueeze(-1)).sum(0)


def calc_volume_weights(z_vals, ray_vector, sigma, last_dist=1e10):
    """MAIN:169-179: alpha = 1-exp(-(relu(sigma)+1e-6) delta |d|); w = alpha * excl-cumprod(1-alpha+1e-10).
    Works for [b,R,S] (reference) and [R,S] (upstream raw2outputs rank)."""
    d = z_vals[..., 1:] - z_vals[..., :-1]
    d = torch.cat([d, torch.full_like(d[..., :1], last_dist)], dim=-1)
    d = d * torch.norm(ray_vector, dim=-1, keepdim=True)
    alpha = 1. - torch.exp(-(F.relu(sigma) + 1e-6) * d)
    trans = torch.cumprod(torch.cat([torch.ones_like(alpha[..., :1]), 1. - alpha + 1e-10], dim=-1), dim=-1)
    return alpha * trans[..., :-1]


def raw2outputs(raw, z_vals, rays_d, bc_rgb, raw_noise_std=0, white_bkgd=False, noise=None):
    """Upstream raw2outputs built on calc_volume_weights.
    raw [R,S,4] (rgb pre-sigmoid, sigma), bc_rgb [R,3] replaces the last sample colour.
    Returns rgb_map, disp_map, acc_map, weights, depth_map."""
    rgb = torch.sigmoid(raw[..., :3])
    rgb = torch.cat((rgb[:, :-1, :], bc_rgb.unsqueeze(1)), dim=1)
    sigma = raw[..., 3]
    if raw_noise_std > 0.:
        sigma = sigma + noise * raw_noise_std
    weights = calc_volume_weights(z_vals, rays_d, sigma)
    rgb_map = torch.sum(weights[..., None] * rgb, -2)
    depth_map = torch.sum(weights * z_vals, -1)
    acc_map = torch.sum(weights, -1)
    disp_map = 1. / torch.max(1e-10 * torch.ones_like(depth_map), depth_map / acc_map)
    if white_bkgd:
        rgb_map = rgb_map + (1. - acc_map[..., None])
    return rgb_map, disp_map, acc_map, weights, depth_map


# ------------------------------------------------------------------- resampling


def sample_pdf(bins, weights, N_samples, det=False, u=None, return_inds=False):
    """HELP:537-581.  `u` may be injected ([R,N] or [N]); det -> linspace(0,1,N)."""
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    if u is None:
        assert det, 'oracle needs det=True or an injected u'
        u = linspace_table(N_samples)
    u = u.expand(list(cdf.shape[:-1]) + [N_samples]).contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_b, cdf_a = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    bin_b, bin_a = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    den = cdf_a - cdf_b
    den = torch.where(den < 1e-5, torch.ones_like(den), den)
    samples = bin_b + (u - cdf_b) / den * (bin_a - bin_b)
    if return_inds:
        return samples, inds
    return samples


# ------------------------------------------------------------- upstream call glue


def run_network(sd, model, pts, viewdirs, aud, multires=10, multires_views=4, netchunk=1024 * 64):
    """network_query_fn (Appendix B): PE(xyz) | aud | PE(viewdir) -> model."""
    flat = pts.reshape(-1, 3)
    parts = [embed(flat, multires)]
    if model == 'facenerf':
        parts.append(aud.reshape(1, -1).expand(flat.shape[0], -1))
    dirs = viewdirs[:, None].expand(pts.shape).reshape(-1, 3)
    parts.append(embed(dirs, multires_views))
    x = torch.cat(parts, -1)
    fwd = facenerf_forward if model == 'facenerf' else nerf_forward
    kw = dict(dim_aud=aud.numel()) if model == 'facenerf' else {}
    out = torch.cat([fwd(sd, x[i:i + netchunk], **kw) for i in range(0, x.shape[0], netchunk)], 0)
    return out.reshape(list(pts.shape[:-1]) + [4])


def render_rays(ray_batch, bc_rgb, aud, sd_coarse, sd_fine, N_samples, N_importance=0,
                model='facenerf', perturb_rand=None, z_samples_override=None, retraw=False):
    """Upstream render_rays (Appendix B).  ray_batch [R,11] = o,d,near,far,viewdir."""
    rays_o, rays_d = ray_batch[:, 0:3], ray_batch[:, 3:6]
    near, far = ray_batch[:, 6:7], ray_batch[:, 7:8]
    viewdirs = ray_batch[:, 8:11]
    z_vals = z_vals_uniform(near, far, N_samples).expand(ray_batch.shape[0], N_samples)
    if perturb_rand is not None:
        z_vals = z_vals_stratified(z_vals, perturb_rand)
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z_vals[:, :, None]
    raw = run_network(sd_coarse, model, pts, viewdirs, aud)
    rgb_map, disp_map, acc_map, weights, depth_map = raw2outputs(raw, z_vals, rays_d, bc_rgb)
    ret = {}
    if N_importance > 0:
        ret.update(rgb0=rgb_map, disp0=disp_map, acc0=acc_map, weights0=weights, z_vals0=z_vals, raw0=raw)
        z_mid = .5 * (z_vals[..., 1:] + z_vals[..., :-1])
        z_samples = sample_pdf(z_mid, weights[..., 1:-1], N_importance, det=True)
        if z_samples_override is not None:
            z_samples = z_samples_override
        ret['z_samples'] = z_samples
        z_vals, _ = torch.sort(torch.cat([z_vals, z_samples], -1), -1)
        pts = rays_o[:, None, :] + rays_d[:, None, :] * z_vals[:, :, None]
        raw = run_network(sd_fine if sd_fine is not None else sd_coarse, model, pts, viewdirs, aud)
        rgb_map, disp_map, acc_map, weights, depth_map = raw2outputs(raw, z_vals, rays_d, bc_rgb)
    ret.update(rgb_map=rgb_map, disp_map=disp_map, acc_map=acc_map, last_weight=weights[..., -1],
               weights=weights, z_vals=z_vals, depth_map=depth_map)
    if retraw:
        ret['raw'] = raw
    return ret


def render(H, W, focal, cx, cy, c2w, bc_rgb, aud, sd_coarse, sd_fine, near, far,
           N_samples=64, N_importance=128, chunk=2048, model='facenerf', ray_slice=None,
           keys=('rgb_map', 'disp_map', 'acc_map', 'last_weight')):
    """Upstream render(): get_rays -> viewdirs -> batchify(render_rays).  bc_rgb [H*W,3].
    ray_slice=(begin,end) renders a contiguous ray range only (bounded CPU samples)."""
    rays_o, rays_d = get_rays(H, W, focal, c2w, cx, cy)
    rays_o, rays_d = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
    viewdirs = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
    n = rays_o.shape[0]
    rays = torch.cat([rays_o, rays_d, near * torch.ones(n, 1), far * torch.ones(n, 1), viewdirs], -1)
    b, e = ray_slice if ray_slice is not None else (0, n)
    outs = []
    # chunk count follows MAIN:655: ceil(n/chunk), the last slice short
    for i in range(b, e, chunk):
        j = min(i + chunk, e)
        outs.append(render_rays(rays[i:j], bc_rgb[i:j], aud, sd_coarse, sd_fine,
                                N_samples, N_importance, model=model))
    return {k: torch.cat([o[k] for o in outs], 0) for k in keys}


# ---------------------------------------------- live two-field path (MAIN:633-708)


def render_head_torso_chunk(sd, rays_o, rays_d, rays_o_t, rays_d_t, z_vals, bc_rgb,
                            z_shape, z_app, signal, signal_torso, last_dist=1e10):
    """One chunk of the reference's live render loop (MAIN:661-708), concate_bg=True.
    z_shape/z_app: [1,2,256]; returns (rgb_head [R,3], rgb_person [R,3])."""
    R, S = z_vals.shape
    p = (rays_o[:, None, :] + rays_d[:, None, :] * z_vals[:, :, None]).reshape(1, -1, 3)
    r = rays_d[:, None, :].expand(R, S, 3).reshape(1, -1, 3)
    pt = (rays_o_t[:, None, :] + rays_d_t[:, None, :] * z_vals[:, :, None]).reshape(1, -1, 3)
    rt = rays_d_t[:, None, :].expand(R, S, 3).reshape(1, -1, 3)
    feat_h, sig_h = decoder_forward(sd, p, r, z_shape[:, 0], z_app[:, 0], signal, 'head')
    sig_h = sig_h.reshape(1, R, S)
    feat_h = feat_h.reshape(1, R, S, 3)
    feat_h = torch.cat((feat_h[..., :-1, :], bc_rgb.reshape(1, R, 1, 3)), dim=-2)
    feat_t, sig_t = decoder_forward(sd, pt, rt, z_shape[:, 1], z_app[:, 1], signal_torso, 'torso')
    sig_t = sig_t.reshape(1, R, S).clone()
    feat_t = feat_t.reshape(1, R, S, 3)
    sig_t[:, :, -1] = 0
    # .clone(): the reference writes the +1e-6 into the ReLU output in place (MAIN:692-694, 884-886), which autograd of
    # current PyTorch rejects when this chunk is differentiated (oracle/train_oracle.py); same arithmetic
    s1 = F.relu(torch.stack([sig_h], 0)).clone()
    f1 = torch.stack([feat_h], 0)
    s2 = F.relu(torch.stack([sig_h, sig_t], 0)).clone()
    f2 = torch.stack([feat_h, feat_t], 0)
    s1[-1, :, :, -1] = s1[-1, :, :, -1] + 1e-6
    s2[-1, :, :, -1] = s2[-1, :, :, -1] + 1e-6
    ss1, fw1 = composite_function(s1, f1)
    ss2, fw2 = composite_function(s2, f2)
    w1 = calc_volume_weights(z_vals[None], rays_d[None], ss1, last_dist)
    w2 = calc_volume_weights(z_vals[None], rays_d_t[None], ss2, last_dist)
    rgb_head = torch.sum(w1.unsqueeze(-1) * fw1, dim=-2).squeeze(0)
    rgb_person = torch.sum(w2.unsqueeze(-1) * fw2, dim=-2).squeeze(0)
    return rgb_head, rgb_person


# ------------------------------------------------- reduced-precision restatement


def bf16_round(x):
    return x.to(torch.bfloat16).to(torch.float32)


def facenerf_forward_bf16(sd, x, **kw):
    """facenerf_forward with every GEMM operand (activations and weights) rounded to
    bf16 and fp32 accumulation -- what a 1-pass bf16 tensor-core kernel computes,
    up to accumulation order.  The latent and view-direction columns stay fp32
    because the CUDA path folds them into fp32 biases."""
    input_ch, dim_aud, input_ch_views = kw.get('input_ch', 63), kw.get('dim_aud', 64), kw.get('input_ch_views', 27)
    D, skips = kw.get('D', 8), kw.get('skips', (4,))
    pe, aud, views = torch.split(x, [input_ch, dim_aud, input_ch_views], dim=-1)

    def lin_split(name, parts):
        # parts: list of (tensor, col_begin, col_end, rounded?)
        W, b = sd[name + '.weight'], sd[name + '.bias']
        acc = b.expand(parts[0][0].shape[0], -1).clone()
        for t, c0, c1, rnd in parts:
            Wp = W[:, c0:c1]
            if rnd:
                acc = acc + bf16_round(t).double().matmul(bf16_round(Wp).double().t()).float()
            else:
                acc = acc + t.matmul(Wp.t())
        return acc

    n_in = input_ch + dim_aud
    h = None
    for i in range(D):
        name = 'pts_linears.%d' % i
        if i == 0:
            a = lin_split(name, [(pe, 0, input_ch, True), (aud, input_ch, n_in, False)])
        elif (i - 1) in skips:
            a = lin_split(name, [(pe, 0, input_ch, True), (aud, input_ch, n_in, False), (h, n_in, n_in + h.shape[1], True)])
        else:
            a = lin_split(name, [(h, 0, h.shape[1], True)])
        h = F.relu(a)
    alpha = lin_split('alpha_linear', [(h, 0, h.shape[1], True)])
    a = lin_split('views_linears.0', [(h, 0, h.shape[1], True), (views, h.shape[1], h.shape[1] + input_ch_views, False)])
    h = F.relu(a)
    for i in range(1, 1 + D // 4):
        h = F.relu(lin_split('views_linears.%d' % i, [(h, 0, h.shape[1], True)]))
    rgb = lin_split('rgb_linear', [(h, 0, h.shape[1], True)])
    return torch.cat([rgb, alpha], -1)


# --------------------------------------------- latent encoders (callers of the path)


def _leaky(x):
    return F.leaky_relu(x, 0.02)


def audionet_forward(sd, x, win_size=16):
    """HELP:109-141 AudioNet: x [N,16,29] -> [N,dim_aud] (four stride-2 Conv1d + two Linear, LeakyReLU(0.02))."""
    half_w = int(win_size / 2)
    x = x[:, 8 - half_w:8 + half_w, :].permute(0, 2, 1)
    for i in (0, 2, 4, 6):
        x = _leaky(F.conv1d(x, sd['encoder_conv.%d.weight' % i], sd['encoder_conv.%d.bias' % i], stride=2, padding=1))
    x = x.squeeze(-1)
    x = _leaky(F.linear(x, sd['encoder_fc1.0.weight'], sd['encoder_fc1.0.bias']))
    return F.linear(x, sd['encoder_fc1.2.weight'], sd['encoder_fc1.2.bias'])


def mlp_encoder_forward(sd, x):
    """HELP:165-178 AudioNet_W2L (512->256->128->64) / HELP:182-193 ExpressionEnc (64->32->32): Linear + LeakyReLU(0.02)
    chain over `encoder.{0,2,..}`, no activation after the last layer."""
    idx = sorted(int(k.split('.')[1]) for k in sd if k.endswith('.weight'))
    for j, i in enumerate(idx):
        x = F.linear(x, sd['encoder.%d.weight' % i], sd['encoder.%d.bias' % i])
        if j + 1 < len(idx):
            x = _leaky(x)
    return x


def audio_att_forward(sd, x, dim_aud, seq_len):
    """HELP:210-240 AudioAttNet: x [seq_len, D] -> [D]: attention weights from the first dim_aud columns (five Conv1d k=3
    + LeakyReLU, Linear(seq,seq), softmax) applied to ALL D columns."""
    y = x[..., :dim_aud].permute(1, 0).unsqueeze(0)
    for i in (0, 2, 4, 6, 8):
        y = _leaky(F.conv1d(y, sd['attentionConvNet.%d.weight' % i], sd['attentionConvNet.%d.bias' % i], stride=1, padding=1))
    y = F.linear(y.view(1, seq_len), sd['attentionNet.0.weight'], sd['attentionNet.0.bias'])
    y = torch.softmax(y, dim=1).view(seq_len, 1)
    return torch.sum(y * x, dim=0)


def _window(x, img_i, half, n):
    """MAIN:36-61 / MAIN:86-101: rows [img_i-half, img_i+half) of x, zero rows where the window leaves [0, n)."""
    left_i, right_i = img_i - half, img_i + half
    pad_left, pad_right = 0, 0
    if left_i < 0:
        pad_left, left_i = -left_i, 0
    if right_i > n:
        pad_right, right_i = right_i - n, n
    win = x[left_i:right_i]
    if pad_left > 0:
        win = torch.cat((torch.zeros_like(win)[:pad_left], win), dim=0)
    if pad_right > 0:
        win = torch.cat((win, torch.zeros_like(win)[:pad_right]), dim=0)
    return win


def encode_signal(auds, exps, img_i, sd_aud, sd_exp, sd_att=None, smo_size=8, dim_aud=64, n=None):
    """MAIN:28-68, itr_obj == 0: the head's per-frame signal [1, 64+32].  sd_att=None is the global_step < nosmo_iters
    branch (MAIN:63-66); otherwise the smoothing window + AudioAttNet of MAIN:35-61."""
    n = auds.shape[0] if n is None else n
    if sd_att is None:
        return torch.cat([mlp_encoder_forward(sd_aud, auds[img_i:img_i + 1]), mlp_encoder_forward(sd_exp, exps[img_i:img_i + 1])], 1)
    half = int(smo_size / 2)
    win = torch.cat([mlp_encoder_forward(sd_aud, _window(auds, img_i, half, n)),
                     mlp_encoder_forward(sd_exp, _window(exps, img_i, half, n))], 1)
    return audio_att_forward(sd_att, win, dim_aud, smo_size).unsqueeze(0)


def rot_to_euler(R):
    """MAIN:182-199."""
    e = torch.ones((R.shape[0], 3))
    e[:, 2] = torch.atan2(R[:, 0, 0], -R[:, 0, 1])
    e[:, 1] = torch.asin(-R[:, 0, 2])
    e[:, 0] = torch.atan2(R[:, 2, 2], R[:, 1, 2])
    return e


def pose_to_euler_trans(poses):
    """MAIN:202-205."""
    return torch.cat((rot_to_euler(poses), poses[:, :3, 3]), dim=1)


def encode_signal_torso(poses, img_i, sd_att=None, smo_size=4, multires=3, n=None):
    """MAIN:78-111: the torso signal [1, 42] = Embedder_3(euler) | Embedder_3(trans) of the head pose (optionally
    smoothed by the pose AudioAttNet over a window of euler/trans rows, zero rows outside the sequence)."""
    n = poses.shape[0] if n is None else n
    if sd_att is None:
        et = pose_to_euler_trans(poses[img_i].unsqueeze(0))
        return torch.cat((embed(et[:, :3], multires), embed(et[:, 3:], multires)), dim=1)
    half = int(smo_size / 2)
    left_i, right_i = max(img_i - half, 0), min(img_i + half, n)
    et = pose_to_euler_trans(poses[left_i:right_i])
    pad_left, pad_right = max(half - img_i, 0), max(img_i + half - n, 0)
    if pad_left > 0:
        et = torch.cat((torch.zeros_like(et)[:pad_left], et), dim=0)
    if pad_right > 0:
        et = torch.cat((et, torch.zeros_like(et)[:pad_right]), dim=0)
    et_embed = torch.cat((embed(et[:, :3], multires), embed(et[:, 3:], multires)), dim=1)
    return audio_att_forward(sd_att, et_embed, et_embed.shape[1], smo_size).unsqueeze(0)


def to8b(x):
    """HELP:17."""
    return (255 * np.clip(x, 0, 1)).astype(np.uint8)
