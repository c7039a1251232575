"""Host-only dumps of the tcgen05 layer programs (dfn_decoder_program_host / dfn_model_program_host: no CUDA calls), shared by
the CPU tests that interpret them (tests/test_layer_programs_cpu.py in fp64, tests/test_quantized_cpu.py with the kernels'
operand rounding) and by the GPU parity tests."""
import ctypes as C

import numpy as np

EPI_RELU, EPI_VIEW0, EPI_RGB, EPI_SIGMA, EPI_CONT, EPI_STAGE = range(6)
KB_PE, KB_IN1, KB_DIR = 4, 5, 6
F_ACCUM, F_DOT_SIGMA = 1, 2
ORDER = (['deform_net.blocks_embed.%d' % i for i in range(5)] + ['deform_net.out_embed'] +
         ['deform_net.blocks_signal.%d' % i for i in range(5)] + ['deform_net.out_signal', 'deform_net.fc_embed_skips.0',
                                                                  'deform_net.fc_signal_skips.0', 'fc_in', 'fc_in_torso', 'fc_z'] +
         ['blocks.%d' % i for i in range(7)] + ['fc_z_skips.0', 'fc_p_skips.0', 'fc_p_skips_torso.0', 'sigma_out', 'fc_z_view',
                                                'feat_view', 'fc_view', 'feat_out'])


def dump_program(sd, field, folded=0, shape=(256, 256, 96, 42, 10, 4)):
    from dfa_nerf_b200 import _lib
    host = []
    for name in ORDER:
        host += [sd[name + '.weight'].contiguous().float(), sd[name + '.bias'].contiguous().float()]
    arr = (C.c_void_p * len(host))(*[t.data_ptr() for t in host])
    desc = _lib.DecoderDesc(*shape, 8, 4)
    ML = 20
    layers = (_lib.LayerInfo * ML)()
    n_layers, n_fold, dimL, view_layer = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    weights = np.zeros((ML, 256, 6, 64), np.float32)
    bias = np.zeros((ML, 256), np.float32)
    fold_layer = (C.c_int * 8)()
    fold_w = np.zeros((8, 1024 * 256), np.float32)
    dot_w = np.zeros(256 + 4, np.float32)
    rc = _lib.lib.dfn_decoder_program_host(C.byref(desc), arr, len(host), field, folded, ML, layers, C.byref(n_layers),
                                           weights.ctypes.data_as(C.c_void_p), bias.ctypes.data_as(C.c_void_p), C.byref(n_fold),
                                           fold_layer, fold_w.ctypes.data_as(C.c_void_p), C.byref(dimL), C.byref(view_layer),
                                           dot_w.ctypes.data_as(C.c_void_p))
    assert rc == 0, _lib.lib.dfn_last_error()
    dl = dimL.value
    folds = {fold_layer[i]: fold_w.reshape(-1)[i * dl * 256:(i + 1) * dl * 256].reshape(dl, 256) for i in range(n_fold.value)}
    return [layers[i] for i in range(n_layers.value)], weights, bias, folds, dl, view_layer.value, dot_w.astype(np.float64)


def dump_model_program(kind, sd, names):
    from dfa_nerf_b200 import _lib
    dim_aud = 64 if kind == _lib.MODEL_FACENERF else 0
    desc = _lib.ModelDesc(kind, 8, 256, 63, 27, dim_aud, 4, 10, 4)
    h = C.c_void_p()
    assert _lib.lib.dfn_model_create(C.byref(desc), C.byref(h)) == 0
    host = []
    for n in names:
        host += [sd[n + '.weight'].contiguous().float(), sd[n + '.bias'].contiguous().float()]
    assert _lib.lib.dfn_model_num_tensors(h) == len(host)
    arr = (C.c_void_p * len(host))(*[t.data_ptr() for t in host])
    ML = 20
    layers = (_lib.LayerInfo * ML)()
    n_layers = C.c_int()
    weights = np.zeros((ML, 256, 6, 64), np.float32)
    bias = np.zeros((ML, 256), np.float32)
    fold_layer = (C.c_int * 2)()
    fold_w = np.zeros((2, 256, max(dim_aud, 1)), np.float32)
    view_w, view_b = np.zeros((128, 27), np.float32), np.zeros(128, np.float32)
    rc = _lib.lib.dfn_model_program_host(h, arr, len(host), ML, layers, C.byref(n_layers), weights.ctypes.data_as(C.c_void_p),
                                         bias.ctypes.data_as(C.c_void_p), fold_layer, fold_w.ctypes.data_as(C.c_void_p),
                                         view_w.ctypes.data_as(C.c_void_p), view_b.ctypes.data_as(C.c_void_p))
    assert rc == 0, _lib.lib.dfn_last_error()
    _lib.lib.dfn_model_destroy(h)
    return [layers[i] for i in range(n_layers.value)], weights, bias, [fold_layer[0], fold_layer[1]], fold_w, view_w, view_b


