"""The reference's live radiance model: ``Decoder`` + ``DeformationField_ori`` (DEC:77-134, DEC:137-349), with the
reference's parameter names so ``load_state_dict`` of its checkpoints works.  ``forward`` composes the CUDA fp32
building blocks of libdfn (dfn_embed for DEC:257-275, dfn_linear for every nn.Linear with the skip / latent / view
terms fused as biases and addends); no torch arithmetic touches the per-point tensors.  Inference only.

``render_head_torso`` is one chunk of the live render loop (MAIN:633-708): head and torso fields, background
splice, two-field density mix, weights and colour sums (dfn_composite_head_torso).
"""
import ctypes as C

import torch
import torch.nn as nn

from ._lib import lib, check, dev, ptr, stream_ptr, DfnError
from .functional import decoder_transform_points, get_rays, z_vals_uniform, make_points


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
    """DEC:137-349 in the configuration MAIN:518 builds (z_dim=256, hidden_size=256, dim_signal=96,
    use_deformation_field=True, 'normal' positional encoding, use_viewdirs, final sigmoid)."""

    def __init__(self, hidden_size=128, n_blocks=8, n_blocks_view=1, dim_signal=64, skips=[4], use_viewdirs=True,
                 n_freq_posenc=10, dim_exp=256, dim_et_embed=42, n_freq_posenc_views=4, use_aud_net=False, dim_aud=64,
                 z_dim=64, rgb_out_dim=3, final_sigmoid_activation=True, downscale_p_by=2., positional_encoding="normal",
                 use_wav2lip=False, dim_w2lfeature=512, gauss_dim_pos=10, gauss_dim_view=4, gauss_std=4.,
                 use_deformation_field=False, use_expression=False, **kwargs):
        super().__init__()
        if positional_encoding != 'normal' or not use_viewdirs or n_blocks_view != 1 or downscale_p_by != 2. or \
                use_expression or use_wav2lip or not final_sigmoid_activation or z_dim <= 0 or rgb_out_dim != 3:
            raise DfnError('Decoder: only the configuration of MAIN:518 is built')
        self.n_freq_posenc, self.n_freq_posenc_views, self.skips = n_freq_posenc, n_freq_posenc_views, skips
        self.z_dim, self.hidden_size, self.n_blocks, self.dim_signal = z_dim, hidden_size, n_blocks, dim_signal
        self.dim_et_embed, self.use_deformation_field = dim_et_embed, use_deformation_field
        de, dv = 6 * n_freq_posenc, 6 * n_freq_posenc_views
        if use_deformation_field:
            self.deform_net = DeformationField_ori(de, dim_et_embed)
        self.fc_in = nn.Linear(de + dim_signal, hidden_size)
        self.fc_in_listener = nn.Linear(de, hidden_size)
        self.fc_in_torso = nn.Linear(de + dim_et_embed, hidden_size)
        self.fc_z = nn.Linear(z_dim, hidden_size)
        self.blocks = nn.ModuleList([nn.Linear(hidden_size, hidden_size) for _ in range(n_blocks - 1)])
        n_skips = sum([i in skips for i in range(n_blocks - 1)])
        self.fc_z_skips = nn.ModuleList([nn.Linear(z_dim, hidden_size) for _ in range(n_skips)])
        self.fc_p_skips = nn.ModuleList([nn.Linear(de + dim_signal, hidden_size) for _ in range(n_skips)])
        self.fc_p_skips_listener = nn.ModuleList([nn.Linear(de, hidden_size) for _ in range(n_skips)])
        self.fc_p_skips_torso = nn.ModuleList([nn.Linear(de + dim_et_embed, hidden_size) for _ in range(n_skips)])
        self.sigma_out = nn.Linear(hidden_size, 1)
        self.fc_z_view = nn.Linear(z_dim, hidden_size)
        self.feat_view = nn.Linear(hidden_size, hidden_size)
        self.fc_view = nn.Linear(dv, hidden_size)
        self.feat_out = nn.Linear(hidden_size, rgb_out_dim)

    def transform_points(self, p, views=False):
        return decoder_transform_points(p, self.n_freq_posenc_views if views else self.n_freq_posenc)

    @torch.no_grad()
    def forward(self, p_in, ray_d, z_shape=None, z_app=None, signal=None, head_or_torso=None):
        """p_in, ray_d [1,P,3]; z_shape, z_app [1,z_dim]; signal [1,dim_signal] (head; a [signal, None] list as the
        reference passes is accepted) or [1,dim_et_embed] (torso) -> (feat [1,P,3], sigma [1,P])."""
        if head_or_torso not in ('head', 'torso'):
            raise Exception('Do not give head or torso!!')
        if isinstance(signal, (list, tuple)):
            signal = signal[0]
        if z_shape is None or z_app is None or signal is None:
            raise DfnError('Decoder.forward: z_shape, z_app and signal are required (as in MAIN:666,675)')
        p_in, _ = dev(p_in.reshape(-1, 3), 'p_in')
        ray_d, _ = dev(ray_d.reshape(-1, 3), 'ray_d')
        z_shape, _ = dev(z_shape.reshape(1, -1), 'z_shape')
        z_app, _ = dev(z_app.reshape(1, -1), 'z_app')
        sig, _ = dev(signal.reshape(1, -1), 'signal')
        P, H, zd = p_in.shape[0], self.hidden_size, self.z_dim
        de = 6 * self.n_freq_posenc
        pe = self.transform_points(p_in)                                   # [P, 60]
        if head_or_torso == 'torso':
            if self.use_deformation_field:
                pe, sig_pts = self.deform_net.deform(pe, sig)              # DEC:297-299: p = deform_net(p) + p
                ld_sig = sig_pts.shape[1]
            else:
                sig_pts, ld_sig = sig, 0
            fc_in, fc_skip = self.fc_in_torso, self.fc_p_skips_torso[0]
        else:
            sig_pts, ld_sig = sig, 0                                       # per-frame signal: one broadcast row
            fc_in, fc_skip = self.fc_in, self.fc_p_skips[0]
        ds = sig.shape[1]
        # per-frame latent terms become biases (DEC:311, DEC:319, DEC:332)
        b_in = _linear(1, H, z_shape, zd, zd, self.fc_z.weight, self.fc_z.bias, 0, addend=fc_in.bias.reshape(1, -1), ld_add=H)
        b_skip = _linear(1, H, z_shape, zd, zd, self.fc_z_skips[0].weight, self.fc_z_skips[0].bias, 0,
                         addend=fc_skip.bias.reshape(1, -1), ld_add=H)
        b_view = _linear(1, H, z_app, zd, zd, self.fc_z_view.weight, self.fc_z_view.bias, 0,
                         addend=self.feat_view.bias.reshape(1, -1), ld_add=H)
        net = _linear(P, H, pe, de, de, fc_in.weight, b_in, RELU, X2=sig_pts, ld2=ld_sig, K2=ds)
        s_term = _linear(P, H, pe, de, de, fc_skip.weight, b_skip, 0, X2=sig_pts, ld2=ld_sig, K2=ds)
        n = len(self.blocks)
        for idx, layer in enumerate(self.blocks):
            add = (idx + 1) in self.skips and idx < n - 1
            net = _linear(P, H, net, H, H, layer.weight, layer.bias, RELU, addend=s_term if add else None, ld_add=H)
        sigma = _linear(P, 1, net, H, H, self.sigma_out.weight, self.sigma_out.bias, 0)
        t = _linear(P, H, net, H, H, self.feat_view.weight, b_view, 0)
        pev = decoder_transform_points(ray_d, self.n_freq_posenc_views, normalize=True)   # [P, 24], DEC:337-338
        dv = pev.shape[1]
        net = _linear(P, H, pev, dv, dv, self.fc_view.weight, self.fc_view.bias, RELU | PRE_ADD, addend=t, ld_add=H)
        feat = _linear(P, 3, net, H, H, self.feat_out.weight, self.feat_out.bias, SIGMOID)
        return feat.reshape(1, P, 3), sigma.reshape(1, P)


@torch.no_grad()
def render_head_torso(decoder, H, W, focal, c2w_head, c2w_torso, bc_rgb, z_shape, z_app, signal, signal_torso, near, far,
                      cx=None, cy=None, N_samples=64, ray_range=None, last_dist=1e10):
    """One frame (or ray range) of the reference's live loop MAIN:633-708: returns (rgb_head, rgb_person) [R,3].
    z_shape / z_app: [1,2,z_dim] (index 0 head, 1 torso, MAIN:664-674)."""
    device = bc_rgb.device
    ro, rd = get_rays(H, W, focal, c2w_head, cx, cy, device=device)
    rot, rdt = get_rays(H, W, focal, c2w_torso, cx, cy, device=device)
    b, e = ray_range if ray_range is not None else (0, H * W)
    ro, rd, rot, rdt = [t.reshape(-1, 3)[b:e].contiguous() for t in (ro, rd, rot, rdt)]
    R = e - b
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
