"""CPU emulation of tensor-core operand-precision schemes for the FaceNeRF query (no GPU needed).

For every scheme the 12 GEMM layers of HELP:275-299 are evaluated with their operands (activations a, weights w) replaced
by what the scheme feeds the tensor cores, products accumulated exactly (fp64), bias / ReLU in fp32 -- i.e. the scheme's
arithmetic without accumulation-order noise.  The latent and view-direction columns stay fp32 (the CUDA path folds them
into fp32 biases).  Output: teacher-forced (golden z_samples injected) max-abs error of rgb0 / rgb_map / weights against
the reference's golden chunk (tests/golden/render_rays.npz, 320 rays x (64 + 192) samples).

    python profiles/precision_emulation.py
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_golden  # noqa: E402
from oracle import nerf_oracle as O  # noqa: E402
import synth  # noqa: E402


def q(x, dt):
    return x.to(dt).to(torch.float32)


BF, H16 = torch.bfloat16, torch.float16
E4, E5 = torch.float8_e4m3fn, torch.float8_e5m2


def sat(x, dt):
    m = torch.finfo(dt).max
    return q(x.clamp(-m, m), dt)


# a scheme maps (a, w) -> list of (a_term, w_term, scale): the layer computes sum_t scale_t * a_t @ w_t^T
def s_fp32(a, w):
    return [(a, w, 1.)]


def s_single(dt):
    def f(a, w):
        return [(sat(a, dt), q(w, dt), 1.)]
    return f


def s_x3(dt):
    def f(a, w):
        ah, wh = sat(a, dt), q(w, dt)
        al, wl = q(a - ah, dt), q(w - wh, dt)
        return [(ah, wh, 1.), (al, wh, 1.), (ah, wl, 1.)]
    return f


def s_x2a(dt):   # activations split, weights single
    def f(a, w):
        ah, wh = sat(a, dt), q(w, dt)
        al = q(a - ah, dt)
        return [(ah, wh, 1.), (al, wh, 1.)]
    return f


def s_x2w(dt):   # weights split, activations single
    def f(a, w):
        ah, wh = sat(a, dt), q(w, dt)
        wl = q(w - wh, dt)
        return [(ah, wh, 1.), (ah, wl, 1.)]
    return f


def s_f16_f8(a_lo_dt=E5, w8_dt=E4, a8_dt=E4, wl_dt=E5, a8_scale=2. ** -4, wl_scale=2. ** 4, hi=H16):
    """fp16 main pass + two fp8 correction passes (kind::f8f6f4 runs K=32 per instruction: the two corrections cost one
    fp16 pass): a_lo (fp8) x w (fp8)  +  a (fp8, scaled) x w_lo (fp8, scaled)."""
    def f(a, w):
        ah, wh = sat(a, hi), q(w, hi)
        al = sat(a - ah, a_lo_dt)
        w8 = sat(w, w8_dt)
        a8 = sat(a * a8_scale, a8_dt)
        wl = sat((w - wh) * wl_scale, wl_dt)
        return [(ah, wh, 1.), (al, w8, 1.), (a8, wl, 1. / (a8_scale * wl_scale))]
    return f


def lin(scheme, a, w):
    acc = torch.zeros(a.shape[0], w.shape[0], dtype=torch.float64)
    for at, wt, sc in scheme(a, w):
        acc += sc * (at.double() @ wt.double().t())
    return acc


def facenerf_q(sd, pe, aud, views, scheme_for_layer, D=8, skips=(4,)):
    """scheme_for_layer(name) -> scheme.  Folded latent / view columns in fp32."""
    n_pe, n_aud = pe.shape[1], aud.shape[1]
    n_in = n_pe + n_aud

    def layer(name, parts, fold):
        W, b = sd[name + '.weight'], sd[name + '.bias']
        sch = scheme_for_layer(name)
        acc = torch.zeros(parts[0][0].shape[0], W.shape[0], dtype=torch.float64)
        for t, c0 in parts:
            acc += lin(sch, t, W[:, c0:c0 + t.shape[1]])
        bias = b.double().expand_as(acc).clone()
        for t, c0 in fold:
            bias += t.double() @ W[:, c0:c0 + t.shape[1]].double().t()
        return (acc + bias).float()

    h = None
    for i in range(D):
        name = 'pts_linears.%d' % i
        if i == 0:
            a = layer(name, [(pe, 0)], [(aud, n_pe)])
        elif (i - 1) in skips:
            a = layer(name, [(pe, 0), (h, n_in)], [(aud, n_pe)])
        else:
            a = layer(name, [(h, 0)], [])
        h = F.relu(a)
    alpha = layer('alpha_linear', [(h, 0)], [])
    h = F.relu(layer('views_linears.0', [(h, 0)], [(views, h.shape[1])]))
    for i in range(1, 1 + D // 4):
        h = F.relu(layer('views_linears.%d' % i, [(h, 0)], []))
    rgb = layer('rgb_linear', [(h, 0)], [])
    return torch.cat([rgb, alpha], -1)


def main():
    g = load_golden('render_rays')
    sd_c, sd_f = synth.facenerf_state_dict(g['coarse_seed']), synth.facenerf_state_dict(g['fine_seed'])
    ro, rd = O.get_rays(g['H'], g['W'], g['focal'], g['c2w'], g['cx'], g['cy'])
    ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
    vd = rd / torch.norm(rd, dim=-1, keepdim=True)
    n = ro.shape[0]
    near, far = torch.full((n, 1), g['near']), torch.full((n, 1), g['far'])
    z0 = O.z_vals_uniform(near, far, 64).expand(n, 64)
    z1 = g['z_vals']

    def query(sd, z, pick):
        pts = (ro[:, None] + rd[:, None] * z[:, :, None]).reshape(-1, 3)
        pe = O.embed(pts, 10)
        aud = g['aud'].reshape(1, -1).expand(pts.shape[0], -1)
        views = O.embed(vd[:, None].expand(n, z.shape[1], 3).reshape(-1, 3), 4)
        return facenerf_q(sd, pe, aud, views, pick).reshape(n, z.shape[1], 4)

    trunk = lambda name: name.startswith('pts_linears') or name == 'alpha_linear'  # noqa: E731
    schemes = {
        'fp32 (sanity)': lambda name: s_fp32,
        'bf16': lambda name: s_single(BF),
        'fp16': lambda name: s_single(H16),
        'bf16x3': lambda name: s_x3(BF),
        'fp16x3': lambda name: s_x3(H16),
        'fp16 a-split (2 passes)': lambda name: s_x2a(H16),
        'fp16 w-split (2 passes)': lambda name: s_x2w(H16),
        'fp16 + fp8 corrections (e5m2 a_lo x e4m3 w, e4m3 a x e5m2 w_lo)': lambda name: s_f16_f8(),
        'fp16 + fp8 corrections, all e5m2': lambda name: s_f16_f8(E5, E5, E5, E5, 2. ** -6, 2. ** 6),
        'fp16 + fp8 corr. on trunk, fp16 single on view layers': lambda name: s_f16_f8() if trunk(name) else s_single(H16),
        'bf16x3 on trunk, fp16 single on view layers': lambda name: s_x3(BF) if trunk(name) else s_single(H16),
        'bf16 + fp8 corrections (bf16 hi)': lambda name: s_f16_f8(hi=BF, a8_scale=1., wl_scale=2. ** 6),
    }
    if sys.argv[1:2] == ['--sweep']:
        # per-layer sensitivity: ONE layer in a cheaper scheme, all others split in three; then cumulative prefixes
        names = ['pts_linears.%d' % i for i in range(8)] + ['alpha_linear', 'views_linears.0', 'views_linears.1', 'views_linears.2', 'rgb_linear']
        base, cheap = s_x3(H16), {'fp16': s_single(H16), 'fp16 w-split': s_x2w(H16), 'fp16 a-split': s_x2a(H16)}
        schemes = {'fp16x3 everywhere': lambda name: base}
        for cn, cs in cheap.items():
            for nm in names:
                schemes['%s on %s only' % (cn, nm)] = (lambda name, nm=nm, cs=cs: cs if name == nm else base)
        view = ('views_linears.0', 'views_linears.1', 'views_linears.2', 'rgb_linear')
        for k in range(0, 9):
            single = set(names[:k]) | set(view)
            schemes['fp16 single on the view layers + pts_linears.0..%d, x3 on the rest' % (k - 1)] = (lambda name, single=single: s_single(H16) if name in single else base)
        for k in range(0, 9):
            single = set(names[:k]) | set(view)
            schemes['fp16 w-split on the view layers + pts_linears.0..%d, x3 on the rest' % (k - 1)] = (lambda name, single=single: s_x2w(H16) if name in single else base)
        sys.argv = sys.argv[:1]
    only = sys.argv[1:]
    for label, pick in schemes.items():
        if only and not any(o in label for o in only):
            continue
        with torch.no_grad():
            raw0 = query(sd_c, z0, pick)
            rgb0, _, _, w0, _ = O.raw2outputs(raw0, z0, rd, g['bc_rgb'])
            raw1 = query(sd_f, z1, pick)
            rgb1, _, acc1, w1, _ = O.raw2outputs(raw1, z1, rd, g['bc_rgb'])
        e = lambda a, b: (a.double() - b.double()).abs().max().item()  # noqa: E731
        print('%-70s rgb0 %.2e rgb_map %.2e weights %.2e sigma %.2e (|sigma|max %.0f) colours(raw) %.2e' % (
            label, e(rgb0, g['rgb0']), e(rgb1, g['rgb_map']), e(w1, g['weights']),
            e(raw1[..., 3], g['raw'][..., 3]), g['raw'][..., 3].abs().max().item(), e(raw1[..., :3], g['raw'][..., :3])), flush=True)


if __name__ == '__main__':
    main()
