"""The reference's live radiance model: ``Decoder`` + ``DeformationField_ori`` (DEC:77-134, DEC:137-349), with the
reference's parameter names so ``load_state_dict`` of its checkpoints works.  Inference only.  Two paths:

* ``forward`` (explicit points, the reference's signature) composes the CUDA fp32 building blocks of libdfn
  (dfn_embed for DEC:257-275, dfn_linear for every nn.Linear with the skip / latent / view terms fused as biases
  and addends); no torch arithmetic touches the per-point tensors.
* ``query_rays`` / ``render_head_torso`` (rays) run the fused tcgen05 kernel (dfn_decoder_query,
  dfn_render_head_torso): encoding, deformation field, trunk, density and colour heads of a whole chunk in one
  launch per field, nothing but raw [R,S,4] leaving the SMs.

``render_head_torso`` is one chunk of the live render loop (MAIN:633-708): head and torso fields, background
splice, two-field density mix, weights and colour sums.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from ._lib import lib, check, dev, ptr, stream_ptr, DfnError, DecoderDesc, HeadTorsoIO, Workspace
from .functional import decoder_transform_points, get_rays, z_vals_uniform, make_points, linspace_table


def _linear(P, N, X1, ld1, K1, W, bias, act=0, X2=None, ld2=0, K2=0, addend=None, ld_add=0, out=None, ldy=None):
    """One dfn_linear launch; returns the [P, N] output tensor."""
    if out is None:
        out = torch.empty((P, N), dtype=torch.float32, device=X1.device)
    ldy = N if ldy is None else ldy
    with torch.cuda.device(X1.device):
        check(lib.dfn_linear(P, N, K1, ptr(X1), ld1, K2, ptr(X2), ld2, ptr(W), ptr(bias), act, ptr(addend), ld_add,
                             ptr(out), ldy, stream_ptr()), 'dfn_linear')
    return out


RELU, SIGMOID, PRE_ADD = 1, 2, 4


class DeformationField_ori(nn.Module):
    """DEC:77-134 with n_blocks=7, skips=[4], hidden 64."""

    def __init__(self, dim_embed, dim_signal, hidden_size=64, n_blocks=7, skips=[4]):
        super().__init__()
        if hidden_size != 64 or n_blocks != 7 or list(skips) != [4]:
            raise DfnError('DeformationField_ori: only hidden_size=64, n_blocks=7, skips=[4] are built (what DEC:208 constructs)')
        self.dim_embed, self.dim_signal, self.skips = dim_embed, dim_signal, skips
        self.blocks_embed = nn.ModuleList([nn.Linear(dim_embed + dim_signal, hidden_size)] +
                                          [nn.Linear(hidden_size, hidden_size) for _ in range(n_blocks - 3)])
        self.out_embed = nn.Linear(hidden_size, dim_embed)
        self.blocks_signal = nn.ModuleList([nn.Linear(dim_embed + dim_signal, hidden_size)] +
                                           [nn.Linear(hidden_size, hidden_size) for _ in range(n_blocks - 3)])
        self.out_signal = nn.Linear(hidden_size, dim_signal)
        n_skips = sum([i in skips for i in range(n_blocks - 1)])
        self.fc_embed_skips = nn.ModuleList([nn.Linear(dim_embed, hidden_size) for _ in range(n_skips)])
        self.fc_signal_skips = nn.ModuleList([nn.Linear(dim_signal, hidden_size) for _ in range(n_skips)])

    def deform(self, pe, sig):
        """pe [P, dim_embed], sig [1, dim_signal] (per-frame) -> (pe + d_embed [P, dim_embed], sig + d_signal [P, dim_signal])."""
        P, de, ds = pe.shape[0], self.dim_embed, self.dim_signal
        outs = []
        for blocks, skip, out_lin, is_embed in ((self.blocks_embed, self.fc_embed_skips[0], self.out_embed, True),
                                                (self.blocks_signal, self.fc_signal_skips[0], self.out_signal, False)):
            if is_embed:       # skip term from the (per-point) encoding
                s_term, ld_s = _linear(P, 64, pe, de, de, skip.weight, skip.bias), 64
            else:              # skip term from the per-frame signal: one broadcast row
                s_term, ld_s = _linear(1, 64, sig, ds, ds, skip.weight, skip.bias), 0
            net = _linear(P, 64, pe, de, de, blocks[0].weight, blocks[0].bias, RELU, X2=sig, ld2=0, K2=ds)
            n = len(blocks)
            for idx in range(1, n):
                add = (idx + 1) in self.skips and idx < n - 1
                net = _linear(P, 64, net, 64, 64, blocks[idx].weight, blocks[idx].bias, RELU,
                              addend=s_term if add else None, ld_add=ld_s)
            if is_embed:
                outs.append(_linear(P, de, net, 64, 64, out_lin.weight, out_lin.bias, 0, addend=pe, ld_add=de))
            else:
                outs.append(_linear(P, ds, net, 64, 64, out_lin.weight, out_lin.bias, 0, addend=sig, ld_add=0))
        return outs[0], outs[1]


class Decoder(nn.Module):
    """DEC:137-349.  The fp32 ``forward`` covers the constructor's options ('normal' positional encoding, downscale_p_by=2): any
    n_blocks / skips / hidden_size / rgb_out_dim, the listener branch (head, signal None: fc_in_listener, fc_p_skips_listener,
    DEC:305-308,322-325), use_expression (DEC:279-281,333-334), use_viewdirs off or ray_d None (DEC:336), final sigmoid off, no
    deformation field.  The fused tcgen05 path (``query_rays`` / ``render_head_torso``) is built for what MAIN:518 constructs
    (z_dim=256, hidden_size=256, dim_signal=96, use_deformation_field=True, n_blocks=8, skips=[4], viewdirs, final sigmoid).
    Not built: 'gauss' encoding (DEC:188-199, needs CUDA at construction in the reference), z_dim <= 0 and n_blocks_view > 1 (both
    fail inside the reference's own forward: DEC:319 / DEC:341-343 feed hidden_size features to a dim_embed_view+hidden_size layer)."""

    def __init__(self, hidden_size=128, n_blocks=8, n_blocks_view=1, dim_signal=64, skips=[4], use_viewdirs=True,
                 n_freq_posenc=10, dim_exp=256, dim_et_embed=42, n_freq_posenc_views=4, use_aud_net=False, dim_aud=64,
                 z_dim=64, rgb_out_dim=3, final_sigmoid_activation=True, downscale_p_by=2., positional_encoding="normal",
                 use_wav2lip=False, dim_w2lfeature=512, gauss_dim_pos=10, gauss_dim_view=4, gauss_std=4.,
                 use_deformation_field=False, use_expression=False, **kwargs):
        super().__init__()
        if positional_encoding != 'normal' or downscale_p_by != 2. or z_dim <= 0:
            raise DfnError("Decoder: only the 'normal' positional encoding with downscale_p_by=2 and z_dim > 0 is built")
        self.use_viewdirs, self.n_blocks_view, self.final_sigmoid_activation = use_viewdirs, n_blocks_view, final_sigmoid_activation
        self.use_expression, self.use_wav2lip, self.rgb_out_dim = use_expression, use_wav2lip, rgb_out_dim
        self.n_freq_posenc, self.n_freq_posenc_views, self.skips = n_freq_posenc, n_freq_posenc_views, skips
        self.z_dim, self.hidden_size, self.n_blocks, self.dim_signal = z_dim, hidden_size, n_blocks, dim_signal
        self.dim_et_embed, self.use_deformation_field, self.dim_exp = dim_et_embed, use_deformation_field, dim_exp
        de, dv = 6 * n_freq_posenc, 6 * n_freq_posenc_views
        # registration order = the reference's (DEC:206-255), so state_dict() lists the same keys in the same order
        if use_deformation_field:
            self.deform_net = DeformationField_ori(de, dim_et_embed)
        if use_expression:
            self.expnet = nn.Linear(dim_exp, hidden_size)
        if use_wav2lip:
            self.w2lnet = nn.Linear(dim_w2lfeature, hidden_size)      # constructed, never used by forward (DEC:221-222)
        self.fc_in = nn.Linear(de + dim_signal, hidden_size)
        self.fc_in_listener = nn.Linear(de, hidden_size)
        self.fc_in_torso = nn.Linear(de + dim_et_embed, hidden_size)
        self.fc_z = nn.Linear(z_dim, hidden_size)
        self.blocks = nn.ModuleList([nn.Linear(hidden_size, hidden_size) for _ in range(n_blocks - 1)])
        n_skips = sum([i in skips for i in range(n_blocks - 1)])
        if n_skips > 0:
            self.fc_z_skips = nn.ModuleList([nn.Linear(z_dim, hidden_size) for _ in range(n_skips)])
            self.fc_p_skips = nn.ModuleList([nn.Linear(de + dim_signal, hidden_size) for _ in range(n_skips)])
            self.fc_p_skips_listener = nn.ModuleList([nn.Linear(de, hidden_size) for _ in range(n_skips)])
            self.fc_p_skips_torso = nn.ModuleList([nn.Linear(de + dim_et_embed, hidden_size) for _ in range(n_skips)])
        self.sigma_out = nn.Linear(hidden_size, 1)
        self.fc_z_view = nn.Linear(z_dim, hidden_size)
        self.feat_view = nn.Linear(hidden_size, hidden_size)
        self.fc_view = nn.Linear(dv, hidden_size)
        self.feat_out = nn.Linear(hidden_size, rgb_out_dim)
        if use_viewdirs and n_blocks_view > 1:
            self.blocks_view = nn.ModuleList([nn.Linear(dv + hidden_size, hidden_size) for _ in range(n_blocks_view - 1)])

    def _fused_config_error(self):
        """None when the module is what the fused tcgen05 programs are compiled for (MAIN:518), else the reason."""
        if not self.use_deformation_field:
            return 'use_deformation_field=True'
        if self.n_blocks != 8 or list(self.skips) != [4]:
            return 'n_blocks=8, skips=[4]'
        if not self.use_viewdirs or self.n_blocks_view != 1 or not self.final_sigmoid_activation or self.rgb_out_dim != 3:
            return 'use_viewdirs, one view block, final sigmoid, rgb_out_dim=3'
        return None

    def transform_points(self, p, views=False):
        return decoder_transform_points(p, self.n_freq_posenc_views if views else self.n_freq_posenc)

    # -- libdfn handle of the fused tcgen05 path ------------------------------------------------------------
    def _tensor_list(self):
        """{weight, bias} in the load order of dfn_decoder_load (include/dfn.h)."""
        why = self._fused_config_error()
        if why is not None:
            raise DfnError('Decoder: the fused path is built for the configuration of MAIN:518 (%s); use forward() in fp32' % why)
        dn = self.deform_net
        mods = list(dn.blocks_embed) + [dn.out_embed] + list(dn.blocks_signal) + [dn.out_signal, dn.fc_embed_skips[0],
                                                                                  dn.fc_signal_skips[0]]
        mods += [self.fc_in, self.fc_in_torso, self.fc_z] + list(self.blocks)
        mods += [self.fc_z_skips[0], self.fc_p_skips[0], self.fc_p_skips_torso[0], self.sigma_out, self.fc_z_view,
                 self.feat_view, self.fc_view, self.feat_out]
        out = []
        for m in mods:
            out += [m.weight, m.bias]
        return out

    def dfn_handle(self, device=None):
        """Creates the native decoder on first use and re-uploads when parameters changed."""
        params = self._tensor_list()
        device = device or params[0].device
        if torch.device(device).type != 'cuda':
            raise DfnError('dfa_nerf_b200 has no CPU path: move the module to CUDA')
        sig = (str(device),) + tuple((p.data_ptr(), p._version) for p in params)
        if getattr(self, '_handle', None) is not None and sig == self._loaded_sig:
            return self._handle
        if getattr(self, '_handle', None) is None:
            desc = DecoderDesc(self.hidden_size, self.z_dim, self.dim_signal, self.dim_et_embed, self.n_freq_posenc,
                               self.n_freq_posenc_views, self.n_blocks, 4)
            h = C.c_void_p()
            check(lib.dfn_decoder_create(C.byref(desc), C.byref(h)), 'dfn_decoder_create')
            self._handle = h
        host = [p.detach().to('cpu', torch.float32).contiguous() for p in params]
        arr = (C.c_void_p * len(host))(*[t.data_ptr() for t in host])
        with torch.cuda.device(device):
            check(lib.dfn_decoder_load(self._handle, arr, len(host), stream_ptr()), 'dfn_decoder_load')
        self._loaded_sig = sig
        return self._handle

    def __del__(self):
        h = getattr(self, '_handle', None)
        if h is not None:
            try:
                lib.dfn_decoder_destroy(h)
            except Exception:
                pass

    @torch.no_grad()
    def query_rays(self, rays_o, rays_d, z_vals, z_shape, z_app, signal, head_or_torso, precision=_lib.PREC_BF16X3):
        """Fused field query for pts = rays_o + rays_d*z (MAIN:638-641 + DEC:277-349): rays_o, rays_d [R,3] (rays_d
        un-normalised), z_vals [R,S] -> (feat [R,S,3] after the sigmoid, sigma [R,S] before MAIN:688's relu)."""
        if head_or_torso not in ('head', 'torso'):
            raise Exception('Do not give head or torso!!')
        expression = None
        if isinstance(signal, (list, tuple)):
            if self.use_expression and head_or_torso == 'head' and len(signal) > 1 and signal[1] is not None:
                expression = signal[1]
            signal = signal[0]
        if signal is None:
            raise DfnError('Decoder.query_rays: the listener branch (signal None, DEC:307) runs through forward() in fp32')
        rays_o, p_o = dev(rays_o, 'rays_o')
        rays_d, p_d = dev(rays_d, 'rays_d')
        z_vals, p_z = dev(z_vals, 'z_vals')
        z_shape, p_zs = dev(z_shape.reshape(-1), 'z_shape')
        z_app, p_za = dev(z_app.reshape(-1), 'z_app')
        signal, p_s = dev(signal.reshape(-1), 'signal')
        field = 0 if head_or_torso == 'head' else 1
        if z_shape.numel() != self.z_dim or z_app.numel() != self.z_dim or \
                signal.numel() != (self.dim_signal if field == 0 else self.dim_et_embed):
            raise DfnError('Decoder.query_rays: latent sizes do not match the module')
        d = rays_o.device
        R, S = z_vals.shape
        h = self.dfn_handle(d)
        raw = torch.empty((R, S, 4), dtype=torch.float32, device=d)
        nbytes = lib.dfn_decoder_query_workspace_bytes(h, R, S)
        ws = Workspace.get(nbytes, d, 'decoder')
        term = self._expression_term(expression, d)
        with torch.cuda.device(d):
            check(lib.dfn_decoder_query_ex(h, field, R, S, p_o, p_d, p_z, p_zs, p_za, p_s, ptr(term), ptr(raw), int(precision),
                                           ptr(ws), nbytes, stream_ptr()), 'dfn_decoder_query')
        return raw[..., :3], raw[..., 3]

    def _expression_term(self, expression, device):
        """expnet(expression) [1, hidden] (DEC:279-281), the per-frame row the fused programs add at the view layer; None without."""
        if expression is None:
            return None
        ex, _ = dev(expression.reshape(1, -1).to(device), 'expression')
        return _linear(1, self.hidden_size, ex, ex.shape[1], ex.shape[1], self.expnet.weight, self.expnet.bias, 0)

    @torch.no_grad()
    def forward(self, p_in, ray_d, z_shape=None, z_app=None, signal=None, head_or_torso=None, precision=_lib.PREC_FP32):
        """DEC:277-349.  p_in, ray_d [1,P,3]; z_shape, z_app [1,z_dim] (None: drawn from N(0,1) as DEC:287-290); signal: head --
        [1,dim_signal], or the reference's [signal, expression] pair (signal None selects the listener layers, expression
        [1,dim_exp] feeds expnet when use_expression); torso -- [1,dim_et_embed].  -> (feat [1,P,rgb_out_dim], sigma [1,P]).
        precision PREC_FP32 (default): the fp32 FFMA building blocks, every constructor option above; PREC_BF16 / PREC_FP16 /
        PREC_BF16X3: the fused tcgen05 kernel on the explicit points (every point is a one-sample ray: origin p, depth 0, its own
        view direction), MAIN:518's configuration only."""
        if head_or_torso not in ('head', 'torso'):
            raise Exception('Do not give head or torso!!')
        head = head_or_torso == 'head'
        expression = None
        if isinstance(signal, (list, tuple)):
            if head and self.use_expression and len(signal) > 1 and signal[1] is not None:
                expression = signal[1]
            signal = signal[0]
        if signal is None and not head:
            raise DfnError('Decoder.forward: the torso field needs its signal (fc_in_torso takes dim_embed + dim_et_embed, DEC:309)')
        if self.use_viewdirs and ray_d is not None and self.n_blocks_view > 1:
            raise DfnError('Decoder.forward: n_blocks_view > 1 feeds hidden_size features to layers of dim_embed_view + hidden_size '
                           'inputs (DEC:253,341-343); the reference fails there too')
        device = p_in.device
        if z_shape is None:
            z_shape = torch.randn(1, self.z_dim, device=device)
        if z_app is None:
            z_app = torch.randn(1, self.z_dim, device=device)
        if precision != _lib.PREC_FP32:
            if ray_d is None:
                raise DfnError('Decoder.forward: the fused path needs view directions; use PREC_FP32')
            pts = p_in.reshape(-1, 3)
            n = pts.shape[0]
            feat, sigma = self.query_rays(pts, ray_d.reshape(-1, 3), torch.zeros((n, 1), dtype=torch.float32, device=pts.device),
                                          z_shape, z_app, [signal, expression], head_or_torso, precision=precision)
            return feat.reshape(1, n, 3), sigma.reshape(1, n)
        p_in, _ = dev(p_in.reshape(-1, 3), 'p_in')
        z_shape, _ = dev(z_shape.reshape(1, -1), 'z_shape')
        z_app, _ = dev(z_app.reshape(1, -1), 'z_app')
        P, H, zd = p_in.shape[0], self.hidden_size, self.z_dim
        de = 6 * self.n_freq_posenc
        pe = self.transform_points(p_in)                                   # [P, de]
        sig_pts, ld_sig, ds = None, 0, 0
        if signal is not None:
            sig, _ = dev(signal.reshape(1, -1), 'signal')
            sig_pts, ds = sig, sig.shape[1]                                # per-frame signal: one broadcast row (DEC:293-295)
        n_skips = sum([i in self.skips for i in range(self.n_blocks - 1)])
        if not head:
            if self.use_deformation_field:
                pe, sig_pts = self.deform_net.deform(pe, sig)              # DEC:297-299: p = deform_net(p) + p
                ld_sig = sig_pts.shape[1]
            fc_in, fc_skips = self.fc_in_torso, (self.fc_p_skips_torso if n_skips else [])
        elif signal is None:
            fc_in, fc_skips = self.fc_in_listener, (self.fc_p_skips_listener if n_skips else [])
        else:
            fc_in, fc_skips = self.fc_in, (self.fc_p_skips if n_skips else [])
        if fc_in.in_features != de + ds:
            raise DfnError('Decoder.forward: signal has %d values, the %s layers take %d' % (ds, head_or_torso, fc_in.in_features - de))
        row = lambda b: b.reshape(1, -1)
        # per-frame latent terms become biases (DEC:311, DEC:319, DEC:332-334)
        b_in = _linear(1, H, z_shape, zd, zd, self.fc_z.weight, self.fc_z.bias, 0, addend=row(fc_in.bias), ld_add=H)
        net = _linear(P, H, pe, de, de, fc_in.weight, b_in, RELU, X2=sig_pts, ld2=ld_sig, K2=ds)
        n = len(self.blocks)
        skip_idx = 0
        for idx, layer in enumerate(self.blocks):
            s_term = None
            if (idx + 1) in self.skips and idx < n - 1:                    # DEC:316-325: added after the relu
                fc_skip = fc_skips[skip_idx]
                b_skip = _linear(1, H, z_shape, zd, zd, self.fc_z_skips[skip_idx].weight, self.fc_z_skips[skip_idx].bias, 0,
                                 addend=row(fc_skip.bias), ld_add=H)
                s_term = _linear(P, H, pe, de, de, fc_skip.weight, b_skip, 0, X2=sig_pts, ld2=ld_sig, K2=ds)
                skip_idx += 1
            net = _linear(P, H, net, H, H, layer.weight, layer.bias, RELU, addend=s_term, ld_add=H)
        sigma = _linear(P, 1, net, H, H, self.sigma_out.weight, self.sigma_out.bias, 0)
        b_view = _linear(1, H, z_app, zd, zd, self.fc_z_view.weight, self.fc_z_view.bias, 0,
                         addend=row(self.feat_view.bias), ld_add=H)
        if expression is not None:                                         # DEC:279-281, 333-334
            ex, _ = dev(expression.reshape(1, -1), 'expression')
            b_view = _linear(1, H, ex, ex.shape[1], ex.shape[1], self.expnet.weight, self.expnet.bias, 0, addend=b_view, ld_add=H)
        net = _linear(P, H, net, H, H, self.feat_view.weight, b_view, 0)
        if self.use_viewdirs and ray_d is not None:                        # DEC:336-340; without it there is no relu either
            ray_d, _ = dev(ray_d.reshape(-1, 3), 'ray_d')
            pev = decoder_transform_points(ray_d, self.n_freq_posenc_views, normalize=True)   # [P, dv], DEC:337-338
            dv = pev.shape[1]
            net = _linear(P, H, pev, dv, dv, self.fc_view.weight, self.fc_view.bias, RELU | PRE_ADD, addend=net, ld_add=H)
        feat = _linear(P, self.rgb_out_dim, net, H, H, self.feat_out.weight, self.feat_out.bias,
                       SIGMOID if self.final_sigmoid_activation else 0)
        return feat.reshape(1, P, self.rgb_out_dim), sigma.reshape(1, P)


@torch.no_grad()
def render_head_torso(decoder, H, W, focal, c2w_head, c2w_torso, bc_rgb, z_shape, z_app, signal, signal_torso, near, far,
                      cx=None, cy=None, N_samples=64, ray_range=None, last_dist=1e10, precision=_lib.PREC_BF16X3,
                      rays_torso=None):
    """One frame (or ray range) of the reference's live loop MAIN:633-708: returns (rgb_head, rgb_person) [R,3].
    z_shape / z_app: [1,2,z_dim] (index 0 head, 1 torso, MAIN:664-674).
    precision PREC_BF16 / PREC_BF16X3: the fused tcgen05 path (dfn_render_head_torso, 5 launches per call);
    PREC_FP32: the explicit-points Decoder.forward built from the fp32 FFMA blocks.
    rays_torso = (rays_o, rays_d) [H*W,3] of the body pose, which is fixed over a sequence (MAIN:644): a frame loop computes
    them once and passes them instead of c2w_torso."""
    device = bc_rgb.device
    ro, rd = get_rays(H, W, focal, c2w_head, cx, cy, device=device)
    rot, rdt = rays_torso if rays_torso is not None else get_rays(H, W, focal, c2w_torso, cx, cy, device=device)
    b, e = ray_range if ray_range is not None else (0, H * W)
    ro, rd, rot, rdt = [t.reshape(-1, 3)[b:e].contiguous() for t in (ro, rd, rot, rdt)]
    R = e - b
    if precision != _lib.PREC_FP32:
        expression = None
        if isinstance(signal, (list, tuple)):
            if decoder.use_expression and len(signal) > 1 and signal[1] is not None:
                expression = signal[1]
            signal = signal[0]
        io = HeadTorsoIO()
        nr = torch.full((R,), float(near), device=device)
        fr = torch.full((R,), float(far), device=device)
        t_vals = linspace_table(N_samples, device)
        bc, io.bc_rgb = dev(bc_rgb.reshape(-1, 3)[b:e], 'bc_rgb')
        zs, io.z_shape = dev(z_shape.reshape(2, -1), 'z_shape')
        za, io.z_app = dev(z_app.reshape(2, -1), 'z_app')
        sg, io.signal = dev(signal.reshape(-1), 'signal')
        sgt, io.signal_torso = dev(signal_torso.reshape(-1), 'signal_torso')
        io.rays_o_head, io.rays_d_head, io.rays_o_torso, io.rays_d_torso = ptr(ro), ptr(rd), ptr(rot), ptr(rdt)
        io.near, io.far, io.t_vals = ptr(nr), ptr(fr), ptr(t_vals)
        rgb_head = torch.empty((R, 3), dtype=torch.float32, device=device)
        rgb_person = torch.empty((R, 3), dtype=torch.float32, device=device)
        io.rgb_head, io.rgb_person, io.last_dist = ptr(rgb_head), ptr(rgb_person), float(last_dist)
        term = decoder._expression_term(expression, device)
        io.expression_term = ptr(term)
        h = decoder.dfn_handle(device)
        nbytes = lib.dfn_render_head_torso_workspace_bytes(h, R, N_samples)
        ws = Workspace.get(nbytes, device, 'head_torso')
        with torch.cuda.device(device):
            check(lib.dfn_render_head_torso(h, R, N_samples, C.byref(io), int(precision), ptr(ws), nbytes, stream_ptr()),
                  'dfn_render_head_torso')
        render_head_torso.last_launches = lib.dfn_last_launch_count() + 2       # + the two get_rays
        return rgb_head, rgb_person
    z = z_vals_uniform(torch.full((R,), float(near), device=device), torch.full((R,), float(far), device=device), N_samples)
    # points: o + d*z (MAIN:638-651); the Decoder takes explicit points, so they are materialised here
    p, r = [t.reshape(1, -1, 3) for t in make_points(ro, rd, z)]
    pt, rt = [t.reshape(1, -1, 3) for t in make_points(rot, rdt, z)]
    feat_h, sig_h = decoder(p, r, z_shape[:, 0], z_app[:, 0], signal, 'head')
    feat_t, sig_t = decoder(pt, rt, z_shape[:, 1], z_app[:, 1], signal_torso, 'torso')
    rgb_head = torch.empty((R, 3), dtype=torch.float32, device=device)
    rgb_person = torch.empty((R, 3), dtype=torch.float32, device=device)
    bc, pb = dev(bc_rgb.reshape(-1, 3)[b:e], 'bc_rgb')
    with torch.cuda.device(device):
        check(lib.dfn_composite_head_torso(R, N_samples, ptr(feat_h), ptr(sig_h), ptr(feat_t), ptr(sig_t), pb, ptr(z),
                                           ptr(rd), ptr(rdt), float(last_dist), ptr(rgb_head), ptr(rgb_person),
                                           stream_ptr()), 'dfn_composite_head_torso')
    return rgb_head, rgb_person
