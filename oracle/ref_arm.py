"""The reference's OWN code on the host cores -- the CPU arm of bench.py when oracle/_ref/ exists (oracle/build_ref.py).
TEST INFRASTRUCTURE ONLY.

The reference has no callable for a whole frame (its render loops are inlined in train(), MAIN:590-734, and its
render_rays, MAIN:114, is dead code that raises), so the per-chunk glue below is the upstream order of SURVEY.md
Appendix B -- but every arithmetic step is a call into the reference's modules as vendored from /root/reference:
  HELP.get_rays (HELP:449), HELP.get_embedder / Embedder (HELP:21-70), HELP.FaceNeRF (HELP:242-299), HELP.sample_pdf
  (HELP:537), MAIN.calc_volume_weights (MAIN:169), MAIN.composite_function (MAIN:146), DEC.Decoder (DEC:137-349).
oracle/make_golden.py has shown the restatement in oracle/nerf_oracle.py bit-equal to exactly these calls.
"""
import os
import sys
import types

import torch

_REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')


def available():
    return all(os.path.exists(os.path.join(_REF, f)) for f in
               ('run_nerf_helpers.py', 'decoder.py', 'run_nerf_com_trainExpLater.py', 'load_audface.py'))


_mods = None


def modules():
    """(HELP, DEC, MAIN) imported from oracle/_ref with the two I/O-only packages the image lacks stubbed."""
    global _mods
    if _mods is None:
        if not available():
            raise ImportError('oracle/_ref is empty: run python oracle/build_ref.py where /root/reference exists')
        for m in ('imageio', 'configargparse'):
            if m not in sys.modules:
                try:
                    __import__(m)
                except Exception:
                    sys.modules[m] = types.ModuleType(m)
        sys.path.insert(0, _REF)
        try:
            import run_nerf_helpers as HELP
            import decoder as DEC
            import run_nerf_com_trainExpLater as MAIN
        finally:
            sys.path.remove(_REF)
        torch.autograd.set_detect_anomaly(False)       # HELP:5 switches it on globally at import
        _mods = (HELP, DEC, MAIN)
    return _mods


class FaceNeRFFrame:
    """Hierarchical FaceNeRF render of a ray range with the reference's modules (chunk = 2048 as scripts/test_obama.sh)."""

    def __init__(self, sd_coarse, sd_fine, N_samples=64, N_importance=128):
        HELP, _, MAIN = modules()
        self.HELP, self.MAIN = HELP, MAIN

        def mk(sd):
            m = HELP.FaceNeRF(D=8, W=256, input_ch=63, input_ch_views=27, dim_aud=64, output_ch=4, skips=[4], use_viewdirs=True)
            m.load_state_dict(sd)
            return m.eval()
        self.nets = (mk(sd_coarse), mk(sd_fine) if sd_fine is not None else None)
        self.embed_fn, _ = HELP.get_embedder(10, 0)
        self.embeddirs_fn, _ = HELP.get_embedder(4, 0)
        self.Nc, self.Nf = N_samples, N_importance

    def _query(self, net, pts, viewdirs, aud):
        flat = pts.reshape(-1, 3)
        dirs = viewdirs[:, None].expand(pts.shape).reshape(-1, 3)
        x = torch.cat([self.embed_fn(flat), aud.reshape(1, -1).expand(flat.shape[0], -1), self.embeddirs_fn(dirs)], -1)
        return net(x).reshape(list(pts.shape[:-1]) + [4])

    def _composite(self, raw, z, rays_d, bc):
        rgb = torch.sigmoid(raw[..., :3])
        rgb = torch.cat((rgb[:, :-1, :], bc.unsqueeze(1)), dim=1)
        w = self.MAIN.calc_volume_weights(z[None], rays_d[None], raw[..., 3][None])[0]
        return torch.sum(w[..., None] * rgb, -2), w

    @torch.no_grad()
    def render(self, H, W, focal, cx, cy, c2w, bc_rgb, aud, near, far, ray_slice, chunk=2048):
        rays_o, rays_d = self.HELP.get_rays(H, W, focal, c2w, cx, cy)
        rays_o, rays_d = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
        b, e = ray_slice
        out = {'rgb_map': [], 'rgb0': [], 'z_samples': []}
        for i in range(b, e, chunk):
            j = min(i + chunk, e)
            ro, rd, bc = rays_o[i:j], rays_d[i:j], bc_rgb[i:j]
            vd = rd / torch.norm(rd, dim=-1, keepdim=True)
            t = torch.linspace(0., 1., steps=self.Nc)                                     # MAIN:617-619
            z = (near * (1. - t) + far * t).expand(j - i, self.Nc)
            raw = self._query(self.nets[0], ro[:, None] + rd[:, None] * z[:, :, None], vd, aud)
            rgb0, w = self._composite(raw, z, rd, bc)
            out['rgb0'].append(rgb0)
            if self.Nf > 0:
                z_mid = .5 * (z[..., 1:] + z[..., :-1])
                zs = self.HELP.sample_pdf(z_mid, w[..., 1:-1], self.Nf, det=True).detach()
                out['z_samples'].append(zs)
                z, _ = torch.sort(torch.cat([z, zs], -1), -1)
                raw = self._query(self.nets[1] or self.nets[0], ro[:, None] + rd[:, None] * z[:, :, None], vd, aud)
                rgb, w = self._composite(raw, z, rd, bc)
                out['rgb_map'].append(rgb)
            else:
                out['rgb_map'].append(rgb0)
        return {k: torch.cat(v, 0) for k, v in out.items() if v}


class HeadTorsoFrame:
    """The reference's live chunk loop MAIN:655-708 (Decoder head + torso, two-field compositing) on a ray range."""

    def __init__(self, sd_decoder):
        HELP, DEC, MAIN = modules()
        self.HELP, self.MAIN = HELP, MAIN
        self.dec = DEC.Decoder(z_dim=256, hidden_size=256, dim_signal=96, use_deformation_field=True)
        self.dec.load_state_dict(sd_decoder)
        self.dec.eval()

    @torch.no_grad()
    def render(self, H, W, focal, cx, cy, c2w, c2w_torso, bc_rgb, z_shape, z_app, signal, signal_torso, near, far, ray_slice,
               N_samples=64, chunk=2048, last_dist=1e10):
        import torch.nn.functional as F
        MAIN = self.MAIN
        ro_a, rd_a = [t.reshape(-1, 3) for t in self.HELP.get_rays(H, W, focal, c2w, cx, cy)]
        rot_a, rdt_a = [t.reshape(-1, 3) for t in self.HELP.get_rays(H, W, focal, c2w_torso, cx, cy)]
        b, e = ray_slice
        heads, persons = [], []
        for i in range(b, e, chunk):
            j = min(i + chunk, e)
            n = j - i
            t = torch.linspace(0., 1., steps=N_samples)
            z = (near * (1. - t) + far * t).expand(n, N_samples)
            ro, rd, rot, rdt = ro_a[i:j], rd_a[i:j], rot_a[i:j], rdt_a[i:j]
            p_i = (ro[..., None, :] + rd[..., None, :] * z[..., :, None]).reshape(1, -1, 3)
            r_i = rd.unsqueeze(1).expand(n, N_samples, 3).reshape(1, -1, 3)
            p_t = (rot[..., None, :] + rdt[..., None, :] * z[..., :, None]).reshape(1, -1, 3)
            r_t = rdt.unsqueeze(1).expand(n, N_samples, 3).reshape(1, -1, 3)
            feat_i, sigma_i = self.dec(p_i, r_i, z_shape[:, 0], z_app[:, 0], signal, 'head')
            sigma_i = sigma_i.reshape(1, n, N_samples)
            feat_i = feat_i.reshape(1, n, N_samples, -1)
            feat_i = torch.cat((feat_i[..., :-1, :], bc_rgb[i:j].reshape(1, n, 1, 3)), dim=-2)
            feat_t, sigma_t = self.dec(p_t, r_t, z_shape[:, 1], z_app[:, 1], signal_torso, 'torso')
            sigma_t = sigma_t.reshape(1, n, N_samples)
            feat_t = feat_t.reshape(1, n, N_samples, -1)
            sigma_t[:, :, -1] = 0
            sigma = F.relu(torch.stack([sigma_i], dim=0))
            sigma_torso = F.relu(torch.stack([sigma_i, sigma_t], dim=0))
            sigma[-1, :, :, -1] += 1e-6
            sigma_torso[-1, :, :, -1] += 1e-6
            s1, f1 = MAIN.composite_function(sigma, torch.stack([feat_i], dim=0))
            s2, f2 = MAIN.composite_function(sigma_torso, torch.stack([feat_i, feat_t], dim=0))
            w1 = MAIN.calc_volume_weights(z.unsqueeze(0), rd.unsqueeze(0), s1, last_dist=last_dist)
            w2 = MAIN.calc_volume_weights(z.unsqueeze(0), rdt.unsqueeze(0), s2, last_dist=last_dist)
            heads.append(torch.sum(w1.unsqueeze(-1) * f1, dim=-2).squeeze(0))
            persons.append(torch.sum(w2.unsqueeze(-1) * f2, dim=-2).squeeze(0))
        return torch.cat(heads, 0), torch.cat(persons, 0)
