"""CPU: the host-side "layer program compiler" of the live model (csrc/mlp_dec.cu: folded per-frame latents, skips
composed into the following layer, block-diagonal deformation field with an identity residual, accumulate-chained
layers for the torso's two staged inputs) -- dumped by dfn_decoder_program_host (no CUDA calls) and interpreted here in
numpy fp64 -- computes the reference's Decoder.forward (DEC:277-349, DEC:109-134) for both fields."""
import ctypes as C

import numpy as np
import torch

from oracle import nerf_oracle as O
from oracle import synth

EPI_RELU, EPI_VIEW0, EPI_RGB, EPI_SIGMA, EPI_CONT, EPI_STAGE = range(6)
KB_PE, KB_IN1, KB_DIR = 4, 5, 6
ORDER = (['deform_net.blocks_embed.%d' % i for i in range(5)] + ['deform_net.out_embed'] +
         ['deform_net.blocks_signal.%d' % i for i in range(5)] + ['deform_net.out_signal', 'deform_net.fc_embed_skips.0',
                                                                  'deform_net.fc_signal_skips.0', 'fc_in', 'fc_in_torso', 'fc_z'] +
         ['blocks.%d' % i for i in range(7)] + ['fc_z_skips.0', 'fc_p_skips.0', 'fc_p_skips_torso.0', 'sigma_out', 'fc_z_view',
                                                'feat_view', 'fc_view', 'feat_out'])


def dump_program(sd, field):
    from dfa_nerf_b200 import _lib
    host = []
    for name in ORDER:
        host += [sd[name + '.weight'].contiguous().float(), sd[name + '.bias'].contiguous().float()]
    arr = (C.c_void_p * len(host))(*[t.data_ptr() for t in host])
    desc = _lib.DecoderDesc(256, 256, 96, 42, 10, 4, 8, 4)
    ML = 20
    layers = (_lib.LayerInfo * ML)()
    n_layers, n_fold, dimL, view_layer = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    weights = np.zeros((ML, 256, 6, 64), np.float32)
    bias = np.zeros((ML, 256), np.float32)
    fold_layer = (C.c_int * 8)()
    fold_w = np.zeros((8, 1024 * 256), np.float32)
    rc = _lib.lib.dfn_decoder_program_host(C.byref(desc), arr, len(host), field, ML, layers, C.byref(n_layers),
                                           weights.ctypes.data_as(C.c_void_p), bias.ctypes.data_as(C.c_void_p), C.byref(n_fold),
                                           fold_layer, fold_w.ctypes.data_as(C.c_void_p), C.byref(dimL), C.byref(view_layer))
    assert rc == 0, _lib.lib.dfn_last_error()
    dl = dimL.value
    folds = {fold_layer[i]: fold_w.reshape(-1)[i * dl * 256:(i + 1) * dl * 256].reshape(dl, 256) for i in range(n_fold.value)}
    return [layers[i] for i in range(n_layers.value)], weights, bias, folds, dl, view_layer.value


def run_program(prog, pe, latent, pe_dir):
    """pe [P,60] fp64, latent [dimL], pe_dir [P,24] = PE(dir/|dir|) (staged block TC_KB_DIR) -> (feat [P,3], sigma [P])."""
    layers, weights, bias, folds, dimL, view_layer = prog
    P = pe.shape[0]
    blocks = {k: np.zeros((P, 64)) for k in range(7)}
    blocks[KB_PE][:, :60] = pe
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
        acc = acc + out if (L.flags & 1) else out
        if L.epi == EPI_CONT:
            continue
        if L.epi == EPI_RELU:
            h = np.maximum(acc + b, 0.)
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
        for field, which in ((0, 'head'), (1, 'torso')):
            prog = dump_program(sd, field)
            n_layers = len(prog[0])
            assert n_layers == (11 if field == 0 else 19)
            latent = torch.cat([sig[field].reshape(-1), zs.reshape(-1), za.reshape(-1)]).double().numpy()
            assert prog[4] == latent.shape[0]
            assert prog[0][prog[5]].kb[0] == KB_DIR          # the view layer stages the direction encoding
            feat, sigma = run_program(prog, pe, latent, pe_dir)
            rf, rs = O.decoder_forward(sd64, p.double(), rd.double(), zs.double(), za.double(), sig[field].double(), which)
            ef = np.abs(feat - rf[0].numpy()).max()
            es = np.abs(sigma - rs[0].numpy()).max() / np.abs(rs[0].numpy()).max()
            # the compiled program stores fp32 weights (products composed in fp64, rounded once): 1e-6 relative
            assert ef < 2e-6 and es < 2e-6, (which, ef, es)


def test_program_host_argument_checks():
    from dfa_nerf_b200 import _lib
    desc = _lib.DecoderDesc(128, 256, 96, 42, 10, 4, 8, 4)      # hidden 128: outside the tcgen05 coverage
    h = C.c_void_p()
    assert _lib.lib.dfn_decoder_create(C.byref(desc), C.byref(h)) == -1
    desc = _lib.DecoderDesc(256, 256, 96, 42, 10, 4, 8, 4)
    assert _lib.lib.dfn_decoder_create(C.byref(desc), C.byref(h)) == 0
    assert _lib.lib.dfn_decoder_num_tensors(h) == 64
    assert _lib.lib.dfn_decoder_macs_per_sample(h, 0) == 0.0          # nothing loaded yet
    _lib.lib.dfn_decoder_destroy(h)
