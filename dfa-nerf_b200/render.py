"""Upstream-convention render surface (AD-NeRF run_nerf.py names) over
dfn_render_rays / dfn_query_points.  The reference inlines this logic four times in
run_nerf_com_trainExpLater.py (MAIN:590-734 etc.) and its own render_rays (MAIN:114) is dead code;
the names and argument order here are the ones run_nerf.py-style host code calls.

The whole per-ray pipeline (z sampling -> encoding -> MLP -> compositing -> sample_pdf -> fine
pass) runs inside the native library; ``network_query_fn`` / ``embed_fn`` arguments are accepted
for signature compatibility and ignored, since encoding is fused into the MLP kernel.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import lib, check, dev, ptr, stream_ptr, DfnError, RenderIO, Workspace
from .functional import get_rays, linspace_table


def _model_handle(net, device):
    if net is None:
        return None
    if not hasattr(net, 'dfn_handle'):
        raise DfnError('network must be a dfa_nerf_b200.NeRF / FaceNeRF module')
    return net.dfn_handle(device)


class RenderEngine:
    """Holds the coarse/fine networks and sampling configuration; renders ray batches or frames."""

    def __init__(self, network_fn, network_fine=None, N_samples=64, N_importance=128, precision=_lib.PREC_BF16X3,
                 white_bkgd=False, perturb=0., raw_noise_std=0., lindisp=False):
        if lindisp:
            raise DfnError('lindisp sampling is not built (the reference never enables it)')
        if raw_noise_std:
            raise DfnError('raw_noise_std is a training-time option; the render path is inference only')
        self.network_fn, self.network_fine = network_fn, network_fine
        self.N_samples, self.N_importance = int(N_samples), int(N_importance)
        self.precision, self.white_bkgd, self.perturb = int(precision), bool(white_bkgd), float(perturb)
        self.last_launches = 0

    def render_rays(self, rays_o, rays_d, viewdirs, near, far, bc_rgb=None, aud=None, z_samples=None,
                    want=('rgb_map', 'disp_map', 'acc_map', 'last_weight'), out=None, pytest=False):
        """rays_* [R,3], near/far [R] -> dict of CUDA tensors.  `z_samples` ([R,N_importance]) teacher-forces
        the fine depths.  `out` may hold preallocated output tensors by name."""
        rays_o, p_o = dev(rays_o, 'rays_o')
        rays_d, p_d = dev(rays_d, 'rays_d')
        viewdirs, p_v = dev(viewdirs, 'viewdirs')
        d = rays_o.device
        R = rays_o.shape[0]
        near, p_n = dev(near.reshape(-1).expand(R), 'near')
        far, p_f = dev(far.reshape(-1).expand(R), 'far')
        Nc, Nf = self.N_samples, self.N_importance
        hc = _model_handle(self.network_fn, d)
        hf = _model_handle(self.network_fine, d)
        io = RenderIO()
        keep = [rays_o, rays_d, viewdirs, near, far]
        io.rays_o, io.rays_d, io.viewdirs, io.near, io.far = p_o, p_d, p_v, p_n, p_f
        if bc_rgb is not None:
            bc_rgb, io.bc_rgb = dev(bc_rgb, 'bc_rgb')
            keep.append(bc_rgb)
        if aud is not None:
            aud, io.latent = dev(aud.reshape(-1), 'aud_para')
            keep.append(aud)
        t_vals = linspace_table(Nc, d)
        io.t_vals = ptr(t_vals)
        if self.perturb > 0.:
            if pytest:
                import numpy as np
                np.random.seed(0)
                rnd = torch.Tensor(np.random.rand(R, Nc)).to(d)
            else:
                rnd = torch.rand((R, Nc), device=d)
            keep.append(rnd)
            io.perturb_rand = ptr(rnd)
        if Nf > 0:
            if self.perturb > 0.:
                if pytest:      # upstream sample_pdf(pytest=True): np.random.seed(0); u = np.random.rand(R, N) (HELP:553-561)
                    import numpy as np
                    np.random.seed(0)
                    u = torch.Tensor(np.random.rand(R, Nf)).to(d)
                else:
                    u = torch.rand((R, Nf), device=d)
                io.u_per_ray = 1
            else:
                u = linspace_table(Nf, d)
            keep.append(u)
            io.u_vals = ptr(u)
            if z_samples is not None:
                z_samples, io.z_samples_in = dev(z_samples, 'z_samples')
                keep.append(z_samples)
        shapes = {'rgb_map': (R, 3), 'disp_map': (R,), 'acc_map': (R,), 'last_weight': (R,), 'rgb0': (R, 3),
                  'z_samples_out': (R, max(Nf, 1)), 'z_vals_out': (R, Nc + Nf)}
        res = {}
        for name in want:
            key = {'z_samples': 'z_samples_out', 'z_vals': 'z_vals_out'}.get(name, name)
            if key not in shapes:
                raise DfnError('unknown output %r' % name)
            if key in ('rgb0', 'z_samples_out') and Nf == 0:
                continue
            t = out[name] if out is not None and name in out else torch.empty(shapes[key], dtype=torch.float32, device=d)
            if tuple(t.shape) != shapes[key] or not t.is_contiguous() or t.dtype != torch.float32:
                raise DfnError('output %r must be contiguous fp32 of shape %s' % (name, shapes[key]))
            setattr(io, key, ptr(t))
            res[name] = t
        nbytes = lib.dfn_render_workspace_bytes2(hc, hf, R, Nc, Nf, self.precision)    # sized for the larger of the two networks
        ws = Workspace.get(nbytes, d, 'render')
        with torch.cuda.device(d):
            check(lib.dfn_render_rays(hc, hf, R, Nc, Nf, C.byref(io), int(self.white_bkgd), self.precision, ptr(ws),
                                      nbytes, stream_ptr()), 'dfn_render_rays')
        self.last_launches = lib.dfn_last_launch_count()
        return res

    def render_frame(self, H, W, focal, c2w, bc_rgb, aud, near, far, cx=None, cy=None, ray_range=None, device=None,
                     want=('rgb_map',), out=None):
        """get_rays + render_rays for a frame (or the contiguous ray range [begin,end) of it)."""
        device = device or (bc_rgb.device if torch.is_tensor(bc_rgb) and bc_rgb.is_cuda else None)
        rays_o, rays_d, vd = get_rays(H, W, focal, c2w, cx, cy, device=device, return_viewdirs=True)
        rays_o, rays_d, vd = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3), vd.reshape(-1, 3)
        b, e = ray_range if ray_range is not None else (0, rays_o.shape[0])
        n = e - b
        dv = rays_o.device
        nr = torch.full((n,), float(near), device=dv)
        fr = torch.full((n,), float(far), device=dv)
        bc = bc_rgb.reshape(-1, 3)[b:e] if bc_rgb is not None else None
        return self.render_rays(rays_o[b:e], rays_d[b:e], vd[b:e], nr, fr, bc, aud, want=want, out=out)

    def query_points(self, net, rays_o, rays_d, viewdirs, z_vals, aud=None, precision=None):
        """network_query_fn: raw [R,S,4] for pts = o + d*z (fused encode + MLP)."""
        rays_o, p_o = dev(rays_o, 'rays_o')
        rays_d, p_d = dev(rays_d, 'rays_d')
        viewdirs, p_v = dev(viewdirs, 'viewdirs')
        z_vals, p_z = dev(z_vals, 'z_vals')
        d = rays_o.device
        R, S = z_vals.shape
        h = _model_handle(net, d)
        p_a = C.c_void_p(0)
        if aud is not None:
            aud, p_a = dev(aud.reshape(-1), 'aud_para')
        prec = self.precision if precision is None else int(precision)
        raw = torch.empty((R, S, 4), dtype=torch.float32, device=d)
        nbytes = lib.dfn_query_workspace_bytes(h, R, S, prec)
        ws = Workspace.get(nbytes, d, 'query')
        with torch.cuda.device(d):
            check(lib.dfn_query_points(h, R, S, p_o, p_d, p_v, p_z, p_a, ptr(raw), prec, ptr(ws), nbytes, stream_ptr()),
                  'dfn_query_points')
        self.last_launches = lib.dfn_last_launch_count()
        return raw


# ------------------------------------------------------------------ upstream-named functions

def run_network(inputs, viewdirs, aud_para, fn, embed_fn=None, embeddirs_fn=None, netchunk=1024 * 64):
    """Upstream run_network: embed(inputs) | aud | embed(viewdirs) -> fn.  Unfused module path (fp32)."""
    from .functional import get_embedder
    flat = inputs.reshape(-1, inputs.shape[-1])
    embed_fn = embed_fn or get_embedder((fn.input_ch - 3) // 6)[0]
    parts = [embed_fn(flat)]
    if aud_para is not None and getattr(fn, 'dim_aud', 0) > 0:
        parts.append(aud_para.reshape(1, -1).expand(flat.shape[0], -1))
    if viewdirs is not None:
        embeddirs_fn = embeddirs_fn or get_embedder((fn.input_ch_views - 3) // 6)[0]
        dirs = viewdirs[:, None].expand(inputs.shape).reshape(-1, inputs.shape[-1])
        parts.append(embeddirs_fn(dirs))
    x = torch.cat(parts, -1)
    out = torch.cat([fn(x[i:i + netchunk]) for i in range(0, x.shape[0], netchunk)], 0)
    return out.reshape(list(inputs.shape[:-1]) + [out.shape[-1]])


def render_rays(ray_batch, bc_rgb, aud_para, network_fn, network_query_fn=None, N_samples=64, retraw=False,
                lindisp=False, perturb=0., N_importance=0, network_fine=None, white_bkgd=False, raw_noise_std=0.,
                verbose=False, pytest=False, precision=_lib.PREC_BF16X3, z_samples=None):
    """Upstream render_rays.  ray_batch [R, 8|11] = rays_o, rays_d, near, far(, viewdirs)."""
    if retraw:
        raise DfnError('retraw: the fused path keeps raw on chip; use RenderEngine.query_points')
    rays_o, rays_d = ray_batch[:, 0:3], ray_batch[:, 3:6]
    near, far = ray_batch[:, 6], ray_batch[:, 7]
    if ray_batch.shape[-1] > 8:
        viewdirs = ray_batch[:, 8:11]
    else:
        viewdirs = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
    eng = RenderEngine(network_fn, network_fine, N_samples, N_importance, precision, white_bkgd, perturb,
                       raw_noise_std, lindisp)
    want = ['rgb_map', 'disp_map', 'acc_map', 'last_weight']
    if N_importance > 0:
        want.append('rgb0')
    return eng.render_rays(rays_o, rays_d, viewdirs, near, far, bc_rgb, aud_para, z_samples=z_samples, want=want,
                           pytest=pytest)


def batchify_rays(rays_flat, bc_rgb, aud_para, chunk=1024 * 32, **kwargs):
    """Upstream batchify_rays: render in chunks of `chunk` rays (the last one short, MAIN:655) and concatenate."""
    outs = {}
    for i in range(0, rays_flat.shape[0], chunk):
        ret = render_rays(rays_flat[i:i + chunk], bc_rgb[i:i + chunk] if bc_rgb is not None else None, aud_para, **kwargs)
        for k, v in ret.items():
            outs.setdefault(k, []).append(v)
    return {k: torch.cat(v, 0) for k, v in outs.items()}


def render(H, W, focal, cx, cy, chunk=1024 * 32, rays=None, bc_rgb=None, aud_para=None, c2w=None, ndc=False,
           near=0., far=1., use_viewdirs=False, c2w_staticcam=None, **kwargs):
    """Upstream render(): returns [rgb_map, disp_map, acc_map, last_weight, extras]."""
    if ndc:
        raise DfnError('ndc rays are not built (dead code in the reference, HELP:484)')
    if c2w is not None:
        rays_o, rays_d = get_rays(H, W, focal, c2w, cx, cy, device=bc_rgb.device if torch.is_tensor(bc_rgb) else None)
    else:
        rays_o, rays_d = rays
    if c2w_staticcam is not None:
        viewsrc = rays_d
        rays_o, rays_d = get_rays(H, W, focal, c2w_staticcam, cx, cy, device=rays_d.device)
    else:
        viewsrc = rays_d
    sh = rays_d.shape
    viewdirs = viewsrc / torch.norm(viewsrc, dim=-1, keepdim=True)
    rays_o, rays_d, viewdirs = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3), viewdirs.reshape(-1, 3)
    n = rays_o.shape[0]
    nr = near * torch.ones((n, 1), device=rays_d.device)
    fr = far * torch.ones((n, 1), device=rays_d.device)
    rays_flat = torch.cat([rays_o, rays_d, nr, fr, viewdirs], -1)
    if bc_rgb is not None:
        bc_rgb = bc_rgb.reshape(-1, 3)
    all_ret = batchify_rays(rays_flat, bc_rgb, aud_para, chunk, **kwargs)
    for k in all_ret:
        all_ret[k] = all_ret[k].reshape(list(sh[:-1]) + list(all_ret[k].shape[1:]))
    k_extract = ['rgb_map', 'disp_map', 'acc_map', 'last_weight']
    ret_list = [all_ret[k] for k in k_extract]
    ret_dict = {k: all_ret[k] for k in all_ret if k not in k_extract}
    return ret_list + [ret_dict]
