"""Reference-named functions of the hot path, backed by the CUDA stage kernels.

Names, argument order and return shapes follow NeRFs/DFANeRF/run_nerf_helpers.py (HELP) and
run_nerf_com_trainExpLater.py (MAIN) of the reference so host code can switch imports:
    get_rays          HELP:449     get_embedder / Embedder   HELP:21-70
    sample_pdf        HELP:537     calc_volume_weights       MAIN:169
    composite_function MAIN:146    raw2outputs               upstream name
All tensors are CUDA fp32; outputs are fresh tensors on the input's device.
"""
import ctypes as C

import numpy as np
import torch

from ._lib import lib, check, dev, ptr, stream_ptr, DfnError

_tables = {}


def linspace_table(n, device, lo=0., hi=1.):
    """torch.linspace evaluated on the CPU (the arithmetic the oracle pins) and cached on `device`."""
    key = (int(n), str(device), float(lo), float(hi))
    t = _tables.get(key)
    if t is None:
        t = torch.linspace(lo, hi, steps=int(n)).to(device)
        _tables[key] = t
    return t


def _device_of(*ts):
    for t in ts:
        if torch.is_tensor(t) and t.is_cuda:
            return t.device
    if not torch.cuda.is_available():
        raise DfnError('dfa_nerf_b200 has no CPU path: CUDA is not available')
    return torch.device('cuda', torch.cuda.current_device())


def get_rays(H, W, focal, c2w, cx=None, cy=None, stride=1, device=None, return_viewdirs=False):
    """HELP:449-465.  c2w: [3,4] or [4,4] tensor / array (host or device).  Returns rays_o, rays_d
    of shape [H//stride, W//stride, 3]; rays_d is R @ dir, not normalised.
    The pose reaches the kernel as launch arguments, so a DEVICE c2w costs one blocking device->host copy per call: frame loops
    keep their pose table on the host (render_sequence / render_sequence_head_torso copy it once per sequence)."""
    device = device or _device_of(c2w)
    n_cols, n_rows = int(W) // stride, int(H) // stride
    xs = linspace_table(n_cols, device, 0., float(W - 1))
    ys = linspace_table(n_rows, device, 0., float(H - 1))
    if cx is None:
        cx = W * .5
    if cy is None:
        cy = H * .5
    c2w_host = torch.as_tensor(c2w, dtype=torch.float32).detach().cpu()[:3, :4].contiguous()
    arr = (C.c_float * 12)(*c2w_host.reshape(-1).tolist())
    rays_o = torch.empty((n_rows, n_cols, 3), dtype=torch.float32, device=device)
    rays_d = torch.empty_like(rays_o)
    vd = torch.empty_like(rays_o) if return_viewdirs else None
    with torch.cuda.device(device):
        check(lib.dfn_get_rays(n_rows, n_cols, ptr(xs), ptr(ys), float(focal), float(cx), float(cy), arr,
                               ptr(rays_o), ptr(rays_d), ptr(vd), stream_ptr()), 'dfn_get_rays')
    if return_viewdirs:
        return rays_o, rays_d, vd
    return rays_o, rays_d


def z_vals_uniform(near, far, N_samples, perturb_rand=None):
    """MAIN:617-619.  near, far: [R] or [R,1] CUDA tensors -> z_vals [R, N_samples]."""
    near, pn = dev(near.reshape(-1), 'near')
    far, pf = dev(far.reshape(-1), 'far')
    R = near.numel()
    t = linspace_table(N_samples, near.device)
    z = torch.empty((R, N_samples), dtype=torch.float32, device=near.device)
    pr = C.c_void_p(0)
    if perturb_rand is not None:
        perturb_rand, pr = dev(perturb_rand, 'perturb_rand')
    with torch.cuda.device(near.device):
        check(lib.dfn_z_vals(R, N_samples, ptr(t), pn, pf, pr, ptr(z), stream_ptr()), 'dfn_z_vals')
    return z


def _embed(x, L, kind):
    x, px = dev(x, 'inputs')
    if x.shape[-1] != 3:
        raise DfnError('positional encoding expects [...,3] inputs')
    D = 3 + 6 * L if kind == 0 else 6 * L
    out = torch.empty(x.shape[:-1] + (D,), dtype=torch.float32, device=x.device)
    P = x.numel() // 3
    if P:
        with torch.cuda.device(x.device):
            check(lib.dfn_embed(P, px, L, kind, ptr(out), stream_ptr()), 'dfn_embed')
    return out


class Embedder:
    """HELP:21-52 with the one configuration get_embedder builds (include_input, log sampling, sin/cos)."""

    def __init__(self, **kwargs):
        self.kwargs = kwargs
        if kwargs.get('input_dims', 3) != 3 or not kwargs.get('include_input', True) or not kwargs.get('log_sampling', True):
            raise DfnError('Embedder: only input_dims=3, include_input=True, log_sampling=True are built')
        self.num_freqs = int(kwargs['num_freqs'])
        if int(kwargs.get('max_freq_log2', self.num_freqs - 1)) != self.num_freqs - 1:
            raise DfnError('Embedder: max_freq_log2 must equal num_freqs-1')
        self.out_dim = 3 + 6 * self.num_freqs

    def embed(self, inputs):
        return _embed(inputs, self.num_freqs, 0)


def get_embedder(multires, i=0):
    """HELP:55-70: returns (embed_fn, out_dim)."""
    if i == -1:
        return torch.nn.Identity(), 3
    eo = Embedder(include_input=True, input_dims=3, max_freq_log2=multires - 1, num_freqs=multires,
                  log_sampling=True, periodic_fns=[torch.sin, torch.cos])
    return (lambda x, eo=eo: eo.embed(x)), eo.out_dim


def decoder_transform_points(p, n_freq, views=False, normalize=False):
    """DEC:257-275 ('normal' encoding, downscale_p_by=2); normalize=True encodes p/||p|| (DEC:337-338)."""
    return _embed(p, n_freq, 2 if normalize else 1)


def make_points(rays_o, rays_d, z_vals):
    """MAIN:638-641: (p [R,S,3] = o + d*z, r [R,S,3] = d repeated per sample)."""
    rays_o, po = dev(rays_o, 'rays_o')
    rays_d, pd = dev(rays_d, 'rays_d')
    z_vals, pz = dev(z_vals, 'z_vals')
    R, S = z_vals.shape
    pts = torch.empty((R, S, 3), dtype=torch.float32, device=z_vals.device)
    dirs = torch.empty((R, S, 3), dtype=torch.float32, device=z_vals.device)
    with torch.cuda.device(z_vals.device):
        check(lib.dfn_make_points(R, S, po, pd, pz, ptr(pts), ptr(dirs), stream_ptr()), 'dfn_make_points')
    return pts, dirs


def calc_volume_weights(z_vals, ray_vector, sigma, last_dist=1e10):
    """MAIN:169-179.  z_vals/sigma [b,R,S] (or [R,S]), ray_vector [b,R,3] -> weights like sigma."""
    shape = sigma.shape
    S = shape[-1]
    sigma, ps = dev(sigma, 'sigma')
    z_vals, pz = dev(z_vals.expand(shape), 'z_vals')
    ray_vector, pr = dev(ray_vector.expand(shape[:-1] + (3,)), 'ray_vector')
    R = sigma.numel() // S
    w = torch.empty(shape, dtype=torch.float32, device=sigma.device)
    with torch.cuda.device(sigma.device):
        check(lib.dfn_calc_volume_weights(R, S, pz, pr, ps, float(last_dist), ptr(w), stream_ptr()),
              'dfn_calc_volume_weights')
    return w


def composite_function(sigma, feat):
    """MAIN:146-166.  sigma [n_box, ...], feat [n_box, ..., 3] -> (sigma_sum [...], feat_weighted [..., 3])."""
    n_box = sigma.shape[0]
    sigma, ps = dev(sigma, 'sigma')
    feat, pf = dev(feat, 'feat')
    n = sigma.numel() // n_box
    ssum = torch.empty(sigma.shape[1:], dtype=torch.float32, device=sigma.device)
    fw = torch.empty(feat.shape[1:], dtype=torch.float32, device=sigma.device)
    with torch.cuda.device(sigma.device):
        check(lib.dfn_composite_fields(n_box, n, ps, pf, ptr(ssum), ptr(fw), stream_ptr()), 'dfn_composite_fields')
    return ssum, fw


def raw2outputs(raw, z_vals, rays_d, bc_rgb=None, raw_noise_std=0, white_bkgd=False, pytest=False,
                raw_is_feat=False, last_dist=1e10):
    """Upstream raw2outputs: raw [R,S,4] -> rgb_map, disp_map, acc_map, weights, depth_map."""
    R, S = raw.shape[0], raw.shape[1]
    raw, _ = dev(raw, 'raw')
    if raw_noise_std > 0.:
        if pytest:
            np.random.seed(0)
            noise = torch.Tensor(np.random.rand(R, S) * raw_noise_std).to(raw.device)
        else:
            noise = torch.randn((R, S), device=raw.device) * raw_noise_std
        raw = raw.clone()
        raw[..., 3] += noise
    z_vals, pz = dev(z_vals, 'z_vals')
    rays_d, pd = dev(rays_d, 'rays_d')
    pb = C.c_void_p(0)
    if bc_rgb is not None:
        bc_rgb, pb = dev(bc_rgb, 'bc_rgb')
    d = raw.device
    rgb_map = torch.empty((R, 3), dtype=torch.float32, device=d)
    disp = torch.empty((R,), dtype=torch.float32, device=d)
    acc = torch.empty((R,), dtype=torch.float32, device=d)
    w = torch.empty((R, S), dtype=torch.float32, device=d)
    depth = torch.empty((R,), dtype=torch.float32, device=d)
    with torch.cuda.device(d):
        check(lib.dfn_raw2outputs(R, S, ptr(raw), pz, pd, pb, int(bool(raw_is_feat)), int(bool(white_bkgd)),
                                  float(last_dist), ptr(rgb_map), ptr(disp), ptr(acc), ptr(w), ptr(depth),
                                  stream_ptr()), 'dfn_raw2outputs')
    return rgb_map, disp, acc, w, depth


def sample_pdf(bins, weights, N_samples, det=False, pytest=False, u=None, return_inds=False):
    """HELP:537-581.  bins [R,nb], weights [R,nb-1] -> samples [R,N_samples].
    `u` ([N] or [R,N]) may be injected; otherwise linspace (det) / uniform random."""
    bins, pb = dev(bins, 'bins')
    weights, pw = dev(weights, 'weights')
    R, nb = bins.shape
    if weights.shape != (R, nb - 1):
        raise DfnError('sample_pdf: weights must be [R, nb-1]')
    d = bins.device
    if u is None:
        if pytest:
            np.random.seed(0)
            u = torch.Tensor(np.linspace(0., 1., N_samples) if det else np.random.rand(R, N_samples)).to(d)
        elif det:
            u = linspace_table(N_samples, d)
        else:
            u = torch.rand((R, N_samples), device=d)
    u, pu = dev(u, 'u')
    per_ray = 1 if u.dim() == 2 else 0
    samples = torch.empty((R, N_samples), dtype=torch.float32, device=d)
    inds = torch.empty((R, N_samples), dtype=torch.int64, device=d) if return_inds else None
    with torch.cuda.device(d):
        check(lib.dfn_sample_pdf(R, nb, pb, pw, nb - 1, N_samples, pu, per_ray, ptr(samples), ptr(inds),
                                 stream_ptr()), 'dfn_sample_pdf')
    return (samples, inds) if return_inds else samples


def invert_cdf(bins, cdf, u):
    """The inversion half of sample_pdf (HELP:563-579) on a given cdf: returns (samples, inds)."""
    bins, pb = dev(bins, 'bins')
    cdf, pc = dev(cdf, 'cdf')
    u, pu = dev(u, 'u')
    R, nb = bins.shape
    N = u.shape[-1]
    samples = torch.empty((R, N), dtype=torch.float32, device=bins.device)
    inds = torch.empty((R, N), dtype=torch.int64, device=bins.device)
    with torch.cuda.device(bins.device):
        check(lib.dfn_invert_cdf(R, nb, pb, pc, N, pu, 1 if u.dim() == 2 else 0, ptr(samples), ptr(inds),
                                 stream_ptr()), 'dfn_invert_cdf')
    return samples, inds


def sort_merge(a, b):
    """z_vals, _ = torch.sort(torch.cat([a, b], -1), -1) (upstream render_rays)."""
    a, pa = dev(a, 'a')
    b, pb_ = dev(b, 'b')
    R = a.shape[0]
    out = torch.empty((R, a.shape[1] + b.shape[1]), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        check(lib.dfn_sort_merge(R, a.shape[1], pa, b.shape[1], pb_, ptr(out), stream_ptr()), 'dfn_sort_merge')
    return out


def coarse_to_fine(raw0, z_vals, rays_d, N_importance, bc_rgb=None, u=None, z_samples=None, white_bkgd=False, want_rgb0=True,
                   last_dist=1e10):
    """The coarse half of upstream render_rays after the network query, fused (dfn_coarse_to_fine): raw2outputs(raw0) ->
    z_mid -> sample_pdf(z_mid, weights[..., 1:-1], N_importance, det) -> sort(cat(z_vals, z_samples)).
    Returns (z_all [R, S+N_importance], z_samples [R, N_importance], rgb0 [R,3] or None); `u` as sample_pdf, `z_samples` injects
    the new depths (teacher forcing)."""
    raw0, pr = dev(raw0, 'raw0')
    z_vals, pz = dev(z_vals, 'z_vals')
    rays_d, pd = dev(rays_d, 'rays_d')
    R, S = z_vals.shape
    d = raw0.device
    pb = C.c_void_p(0)
    if bc_rgb is not None:
        bc_rgb, pb = dev(bc_rgb, 'bc_rgb')
    if u is None and z_samples is None:
        u = linspace_table(N_importance, d)
    pu, per_ray = C.c_void_p(0), 0
    if u is not None:
        u, pu = dev(u, 'u')
        per_ray = 1 if u.dim() == 2 else 0
    pzs = C.c_void_p(0)
    if z_samples is not None:
        z_samples, pzs = dev(z_samples, 'z_samples')
    z_all = torch.empty((R, S + N_importance), dtype=torch.float32, device=d)
    zs = torch.empty((R, N_importance), dtype=torch.float32, device=d)
    rgb0 = torch.empty((R, 3), dtype=torch.float32, device=d) if want_rgb0 else None
    with torch.cuda.device(d):
        check(lib.dfn_coarse_to_fine(R, S, N_importance, pr, pz, pd, pb, int(bool(white_bkgd)), float(last_dist), pu, per_ray, pzs,
                                     ptr(rgb0), ptr(zs), ptr(z_all), stream_ptr()), 'dfn_coarse_to_fine')
    return z_all, zs, rgb0


def to8b(x):
    """HELP:17 on the device: uint8(255 * clip(x, 0, 1)), same shape (numpy's fp32 product and truncation)."""
    x, px = dev(x, 'x')
    out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    if x.numel():
        with torch.cuda.device(x.device):
            check(lib.dfn_to8b(x.numel(), px, ptr(out), stream_ptr()), 'dfn_to8b')
    return out
