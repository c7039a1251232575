"""CPU: the host-side "layer program compiler" of the live model (csrc/mlp_dec.cu: folded per-frame latents, skips
composed into the following layer, block-diagonal deformation field with an identity residual, accumulate-chained
layers for the torso's two staged inputs) -- dumped by dfn_decoder_program_host (no CUDA calls) and interpreted here in
numpy fp64 -- computes the reference's Decoder.forward (DEC:277-349, DEC:109-134) for both fields."""
import ctypes as C

import numpy as np
import torch

from oracle import nerf_oracle as O
from oracle import synth

from program_dump import (EPI_RELU, EPI_VIEW0, EPI_RGB, EPI_SIGMA, EPI_CONT, EPI_STAGE, KB_PE, KB_IN1, KB_DIR, F_ACCUM,
                          F_DOT_SIGMA, ORDER, dump_program, dump_model_program)


def run_program(prog, pe, latent, pe_dir):
    """pe [P,6*n_freq] fp64, latent [dimL], pe_dir [P,24] = PE(dir/|dir|) (staged block TC_KB_DIR) -> (feat [P,3], sigma [P])."""
    layers, weights, bias, folds, dimL, view_layer, dot_w = prog
    P = pe.shape[0]
    blocks = {k: np.zeros((P, 64)) for k in range(7)}
    blocks[KB_PE][:, :pe.shape[1]] = pe
    blocks[KB_DIR][:, :pe_dir.shape[1]] = pe_dir
    acc, sigma, feat = None, None, None
    for l, L in enumerate(layers):
        n, kbs = L.n, [L.kb[i] for i in range(L.nkb)]
        x = np.concatenate([blocks[k] for k in kbs], 1)
        W = weights[l, :n, :L.nkb].reshape(n, -1).astype(np.float64)
        b = bias[l, :n].astype(np.float64)
        if l in folds:
            b = b + folds[l][:, :n].astype(np.float64).T @ latent
        out = x @ W.T
        acc = acc + out if (L.flags & F_ACCUM) else out
        if L.epi == EPI_CONT:
            continue
        if L.epi == EPI_RELU:
            h = np.maximum(acc + b, 0.)
            if L.flags & F_DOT_SIGMA:        # folded head: density from this layer's activations
                sigma = h @ dot_w[:256] + dot_w[256]
        elif L.epi == EPI_VIEW0:
            raise AssertionError('the Decoder programs carry the view term as a K-block, not as a per-ray bias')
        elif L.epi == EPI_SIGMA:
            sigma = acc[:, 0] + b[0]
            continue
        elif L.epi == EPI_STAGE:
            o = acc + b
            blocks[KB_PE], blocks[KB_IN1] = o[:, :64].copy(), o[:, 64:128].copy()
            continue
        elif L.epi == EPI_RGB:
            feat = 1. / (1. + np.exp(-(acc[:, :3] + b[:3])))
            continue
        for i in range(n // 64):
            blocks[i] = h[:, 64 * i:64 * (i + 1)]
    return feat, sigma


def test_decoder_layer_programs_compute_the_reference_forward():
    sd = synth.decoder_state_dict(3)
    g = torch.Generator().manual_seed(0)
    P = 97
    p = (torch.rand(1, P, 3, generator=g) * 2 - 1) * 0.7
    rd = torch.randn(1, P, 3, generator=g)
    zs, za = torch.randn(1, 256, generator=g), torch.randn(1, 256, generator=g)
    sig = {0: torch.randn(1, 96, generator=g), 1: torch.randn(1, 42, generator=g)}
    sd64 = {k: v.double() for k, v in sd.items()}
    with torch.no_grad():
        pe = O.decoder_transform_points(p.double(), 10)[0].numpy()
        d = rd.double() / torch.norm(rd.double(), dim=-1, keepdim=True)
        pe_dir = O.decoder_transform_points(d, 4)[0].numpy()
        for field, which, folded in ((0, 'head', 0), (1, 'torso', 0), (0, 'head', 1), (1, 'torso', 1)):
            prog = dump_program(sd, field, folded)
            n_layers = len(prog[0])
            assert n_layers == (11 if field == 0 else 19) - folded
            assert [L.flags & F_DOT_SIGMA for L in prog[0]].count(F_DOT_SIGMA) == folded
            assert all(L.epi != EPI_SIGMA for L in prog[0]) == bool(folded)
            latent = torch.cat([sig[field].reshape(-1), zs.reshape(-1), za.reshape(-1)]).double().numpy()
            assert prog[4] == latent.shape[0]
            assert prog[0][prog[5]].kb[0] == KB_DIR          # the view layer stages the direction encoding
            feat, sigma = run_program(prog, pe, latent, pe_dir)
            rf, rs = O.decoder_forward(sd64, p.double(), rd.double(), zs.double(), za.double(), sig[field].double(), which)
            ef = np.abs(feat - rf[0].numpy()).max()
            es = np.abs(sigma - rs[0].numpy()).max() / np.abs(rs[0].numpy()).max()
            # the compiled program stores fp32 weights (products composed in fp64, rounded once): 1e-6 relative
            assert ef < 2e-6 and es < 2e-6, (which, ef, es)


def test_decoder_layer_programs_other_shapes():
    """The program compiler for non-default module sizes the C ABI accepts (fewer PE frequencies, other latent widths):
    (hidden, z_dim, dim_signal, dim_et_embed, n_freq, n_freq_views)."""
    for shape in ((256, 64, 64, 30, 6, 2), (256, 128, 32, 18, 10, 1)):
        _, zd, ds, dt, nf, nfv = shape
        sd = synth.decoder_state_dict(9, z_dim=zd, dim_signal=ds, dim_et=dt, n_freq=nf, n_freq_views=nfv)
        g = torch.Generator().manual_seed(2)
        P = 33
        p = (torch.rand(1, P, 3, generator=g) * 2 - 1) * 0.7
        rd = torch.randn(1, P, 3, generator=g)
        zs, za = torch.randn(1, zd, generator=g), torch.randn(1, zd, generator=g)
        sig = {0: torch.randn(1, ds, generator=g), 1: torch.randn(1, dt, generator=g)}
        sd64 = {k: v.double() for k, v in sd.items()}
        with torch.no_grad():
            pe = O.decoder_transform_points(p.double(), nf)[0].numpy()
            d = rd.double() / torch.norm(rd.double(), dim=-1, keepdim=True)
            pe_dir = O.decoder_transform_points(d, nfv)[0].numpy()
            for field, which in ((0, 'head'), (1, 'torso')):
                for folded in (0, 1):
                    prog = dump_program(sd, field, folded, shape)
                    latent = torch.cat([sig[field].reshape(-1), zs.reshape(-1), za.reshape(-1)]).double().numpy()
                    assert prog[4] == latent.shape[0]
                    feat, sigma = run_program(prog, pe, latent, pe_dir)
                    rf, rs = O.decoder_forward(sd64, p.double(), rd.double(), zs.double(), za.double(), sig[field].double(), which,
                                               n_freq=nf, n_freq_views=nfv)
                    ef = np.abs(feat - rf[0].numpy()).max()
                    es = np.abs(sigma - rs[0].numpy()).max() / np.abs(rs[0].numpy()).max()
                    assert ef < 2e-6 and es < 2e-6, (shape, which, folded, ef, es)


def test_program_host_argument_checks():
    from dfa_nerf_b200 import _lib
    # the folded-head program needs somewhere to put the head row
    sd = synth.decoder_state_dict(3)
    host = []
    for name in ORDER:
        host += [sd[name + '.weight'].contiguous().float(), sd[name + '.bias'].contiguous().float()]
    arr = (C.c_void_p * len(host))(*[t.data_ptr() for t in host])
    d0 = _lib.DecoderDesc(256, 256, 96, 42, 10, 4, 8, 4)
    layers = (_lib.LayerInfo * 20)()
    n = C.c_int()
    w, b, fw = np.zeros((20, 256, 6, 64), np.float32), np.zeros((20, 256), np.float32), np.zeros((8, 1024 * 256), np.float32)
    fl = (C.c_int * 8)()
    args = [C.byref(d0), arr, len(host), 0, 1, 20, layers, C.byref(n), w.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p),
            C.byref(C.c_int()), fl, fw.ctypes.data_as(C.c_void_p), C.byref(C.c_int()), C.byref(C.c_int())]
    assert _lib.lib.dfn_decoder_program_host(*args, None) == -1 and b'dot_w' in _lib.lib.dfn_last_error()
    assert _lib.lib.dfn_decoder_program_host(*args[:4], 0, *args[5:], None) == 0       # plain program: dot_w not needed
    assert n.value == 11
    desc = _lib.DecoderDesc(128, 256, 96, 42, 10, 4, 8, 4)      # hidden 128: outside the tcgen05 coverage
    h = C.c_void_p()
    assert _lib.lib.dfn_decoder_create(C.byref(desc), C.byref(h)) == -1
    desc = _lib.DecoderDesc(256, 256, 96, 42, 10, 4, 8, 4)
    assert _lib.lib.dfn_decoder_create(C.byref(desc), C.byref(h)) == 0
    assert _lib.lib.dfn_decoder_num_tensors(h) == 64
    assert _lib.lib.dfn_decoder_macs_per_sample(h, 0) == 0.0          # nothing loaded yet
    _lib.lib.dfn_decoder_destroy(h)


def _run_model_program(prog, pe, latent, pe_view):
    layers, weights, bias, fold_layer, fold_w, view_w, view_b = prog
    P = pe.shape[0]
    blocks = {k: np.zeros((P, 64)) for k in range(5)}
    blocks[KB_PE][:, :63] = pe
    alpha, rgb = None, None
    for l, L in enumerate(layers):
        n = L.n
        x = np.concatenate([blocks[L.kb[i]] for i in range(L.nkb)], 1)
        acc = x @ weights[l, :n, :L.nkb].reshape(n, -1).astype(np.float64).T
        b = bias[l, :n].astype(np.float64)
        if latent is not None and l in fold_layer:
            b = b + np.pad(fold_w[fold_layer.index(l)].astype(np.float64) @ latent, (0, 0))[:n]
        if L.epi == EPI_RGB:
            rgb = acc[:, :3] + b[:3]
            continue
        if L.epi == EPI_VIEW0:
            wh = view_w.shape[0]
            alpha = acc[:, wh] + b[wh]
            h = np.maximum(acc[:, :wh] + view_b.astype(np.float64) + pe_view @ view_w.astype(np.float64).T, 0.)
            n = wh
        else:
            h = np.maximum(acc + b, 0.)
        for i in range(n // 64):
            blocks[i] = h[:, 64 * i:64 * (i + 1)]
    return np.concatenate([rgb, alpha[:, None]], 1)


def test_facenerf_and_nerf_layer_programs_compute_the_reference_forward():
    """FaceNeRF (HELP:275-299: latent columns folded into biases, view columns into a per-ray term, alpha riding the first
    view layer) and NeRF (HELP:372-396: feature_linear composed into views_linears.0) as compiled by tc_pack_model."""
    import dfa_nerf_b200 as dfn
    from dfa_nerf_b200 import _lib
    g = torch.Generator().manual_seed(1)
    P = 61
    pts = (torch.rand(P, 3, generator=g) * 2 - 1).double()
    vd = torch.randn(P, 3, generator=g).double()
    vd = vd / torch.norm(vd, dim=-1, keepdim=True)
    aud = torch.randn(64, generator=g).double()
    pe, pev = O.embed(pts, 10), O.embed(vd, 4)
    trunk = ['pts_linears.%d' % i for i in range(8)]
    cases = ((_lib.MODEL_FACENERF, synth.facenerf_state_dict(2), trunk + ['views_linears.%d' % i for i in range(3)] +
              ['feature_linear', 'alpha_linear', 'rgb_linear'], 12),
             (_lib.MODEL_NERF, synth.nerf_state_dict(2), trunk + ['views_linears.0', 'feature_linear', 'alpha_linear', 'rgb_linear'], 10))
    with torch.no_grad():
        for kind, sd, names, nl in cases:
            prog = dump_model_program(kind, sd, names)
            assert len(prog[0]) == nl
            sd64 = {k: v.double() for k, v in sd.items()}
            if kind == _lib.MODEL_FACENERF:
                ref = O.facenerf_forward(sd64, torch.cat([pe, aud[None].expand(P, -1), pev], -1))
                out = _run_model_program(prog, pe.numpy(), aud.numpy(), pev.numpy())
            else:
                ref = O.nerf_forward(sd64, torch.cat([pe, pev], -1))
                out = _run_model_program(prog, pe.numpy(), None, pev.numpy())
            scale = np.abs(ref.numpy()).max(0)
            err = (np.abs(out - ref.numpy()) / scale).max()
            assert err < 2e-6, (kind, err)
