"""ctypes binding of libdfn.so -- the only way into the CUDA kernels (include/dfn.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C dfa-nerf_b200/csrc``.
Importing this module without it raises: there is no CPU or PyTorch fallback.
"""
import ctypes as C
import os

import torch

PREC_FP32, PREC_BF16, PREC_FP16, PREC_BF16X3, PREC_FP16X3M = 0, 1, 2, 3, 4
MODEL_FACENERF, MODEL_NERF = 0, 1

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libdfn.so')


class DfnError(RuntimeError):
    pass


class ModelDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in ('kind', 'D', 'W', 'input_ch', 'input_ch_views', 'dim_aud', 'skip',
                                       'multires', 'multires_views')]


class RenderIO(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        'rays_o', 'rays_d', 'viewdirs', 'near', 'far', 'bc_rgb', 'latent', 't_vals', 'u_vals', 'perturb_rand',
        'z_samples_in', 'rgb_map', 'disp_map', 'acc_map', 'last_weight', 'rgb0', 'z_samples_out', 'z_vals_out')] + \
        [('u_per_ray', C.c_int64)]


class DecoderDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in ('hidden', 'z_dim', 'dim_signal', 'dim_et_embed', 'n_freq', 'n_freq_views',
                                       'n_blocks', 'skip')]


class ConvStack(C.Structure):
    _fields_ = [('n', C.c_int), ('ch', C.c_int * 7), ('stride', C.c_int), ('w', C.c_void_p * 6), ('b', C.c_void_p * 6)]


class LayerInfo(C.Structure):
    _fields_ = [('n', C.c_int), ('nkb', C.c_int), ('epi', C.c_int), ('flags', C.c_int), ('kb', C.c_int * 6)]


class HeadTorsoIO(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ('rays_o_head', 'rays_d_head', 'rays_o_torso', 'rays_d_torso', 'near', 'far',
                                          't_vals', 'bc_rgb', 'z_shape', 'z_app', 'signal', 'signal_torso', 'rgb_head',
                                          'rgb_person')] + [('last_dist', C.c_float), ('expression_term', C.c_void_p)]


class GemmDesc(C.Structure):
    _fields_ = [('A', C.c_void_p), ('a_ld_r', C.c_int64), ('a_ld_k', C.c_int64), ('A_mask', C.c_void_p), ('a_mask_mode', C.c_int),
                ('B', C.c_void_p), ('b_ld_r', C.c_int64), ('b_ld_k', C.c_int64),
                ('C', C.c_void_p), ('c_ld_r', C.c_int64), ('c_ld_c', C.c_int64),
                ('bias', C.c_void_p), ('addend', C.c_void_p), ('add_ld_r', C.c_int64), ('add_ld_c', C.c_int64),
                ('act', C.c_int), ('M', C.c_int), ('N', C.c_int), ('K', C.c_int), ('beta', C.c_int), ('k_splits', C.c_int),
                ('precision', C.c_int)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError('dfa_nerf_b200: %s is missing -- run `python -c "import __graft_entry__ as g; g.build()"` '
                          '(there is no CPU fallback)' % LIB_PATH)
    lib_ = C.CDLL(LIB_PATH)
    vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float
    sig = {
        'dfn_abi_version': (i32, []),
        'dfn_last_error': (C.c_char_p, []),
        'dfn_last_launch_count': (i32, []),
        'dfn_profile_enable': (i32, [i32]),
        'dfn_debug_trace': (i32, [vp, i32]),
        'dfn_debug_set_impl': (i32, [i32]),
        'dfn_debug_set_pp_flags': (i32, [i32]),
        'dfn_profile_collect': (i32, [C.POINTER(C.c_double), C.POINTER(i64), C.POINTER(C.c_double)]),
        'dfn_get_rays': (i32, [i32, i32, vp, vp, f32, f32, f32, C.POINTER(f32), vp, vp, vp, vp]),
        'dfn_z_vals': (i32, [i32, i32, vp, vp, vp, vp, vp, vp]),
        'dfn_make_points': (i32, [i32, i32, vp, vp, vp, vp, vp, vp]),
        'dfn_embed': (i32, [i64, vp, i32, i32, vp, vp]),
        'dfn_composite_fields': (i32, [i32, i64, vp, vp, vp, vp, vp]),
        'dfn_calc_volume_weights': (i32, [i32, i32, vp, vp, vp, f32, vp, vp]),
        'dfn_raw2outputs': (i32, [i32, i32, vp, vp, vp, vp, i32, i32, f32, vp, vp, vp, vp, vp, vp]),
        'dfn_composite_head_torso': (i32, [i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, f32, vp, vp, vp]),
        'dfn_linear': (i32, [i64, i32, i32, vp, i64, i32, vp, i64, vp, vp, i32, vp, i64, vp, i64, vp]),
        'dfn_sample_pdf': (i32, [i32, i32, vp, vp, i64, i32, vp, i32, vp, vp, vp]),
        'dfn_invert_cdf': (i32, [i32, i32, vp, vp, i32, vp, i32, vp, vp, vp]),
        'dfn_sort_merge': (i32, [i32, i32, vp, i32, vp, vp, vp]),
        'dfn_coarse_to_fine': (i32, [i32, i32, i32, vp, vp, vp, vp, i32, f32, vp, i32, vp, vp, vp, vp, vp]),
        'dfn_to8b': (i32, [i64, vp, vp, vp]),
        'dfn_audionet_forward': (i32, [i32, i32, vp, C.POINTER(ConvStack), vp, vp, vp, vp, vp, vp]),
        'dfn_att_smooth': (i32, [i32, i32, i32, i32, vp, vp, C.POINTER(ConvStack), vp, vp, vp, vp]),
        'dfn_pose_signal': (i32, [i32, vp, i32, i32, vp, vp, vp]),
        'dfn_model_create': (i32, [C.POINTER(ModelDesc), C.POINTER(vp)]),
        'dfn_model_destroy': (None, [vp]),
        'dfn_model_num_tensors': (i32, [vp]),
        'dfn_model_load': (i32, [vp, C.POINTER(vp), i32, vp]),
        'dfn_mlp_workspace_bytes': (i64, [vp, i64]),
        'dfn_mlp_forward': (i32, [vp, i64, vp, vp, vp, i64, vp]),
        'dfn_query_workspace_bytes': (i64, [vp, i64, i32, i32]),
        'dfn_query_points': (i32, [vp, i64, i32, vp, vp, vp, vp, vp, vp, i32, vp, i64, vp]),
        'dfn_render_workspace_bytes': (i64, [vp, i64, i32, i32, i32]),
        'dfn_render_workspace_bytes2': (i64, [vp, vp, i64, i32, i32, i32]),
        'dfn_render_rays': (i32, [vp, vp, i64, i32, i32, C.POINTER(RenderIO), i32, i32, vp, i64, vp]),
        'dfn_decoder_create': (i32, [C.POINTER(DecoderDesc), C.POINTER(vp)]),
        'dfn_decoder_destroy': (None, [vp]),
        'dfn_decoder_num_tensors': (i32, [vp]),
        'dfn_decoder_load': (i32, [vp, C.POINTER(vp), i32, vp]),
        'dfn_decoder_query_workspace_bytes': (i64, [vp, i64, i32]),
        'dfn_decoder_query': (i32, [vp, i32, i64, i32, vp, vp, vp, vp, vp, vp, vp, i32, vp, i64, vp]),
        'dfn_decoder_query_ex': (i32, [vp, i32, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, i64, vp]),
        'dfn_decoder_macs_per_sample': (C.c_double, [vp, i32]),
        'dfn_model_program_host': (i32, [vp, C.POINTER(vp), i32, i32, C.POINTER(LayerInfo), C.POINTER(i32), vp, vp,
                                         C.POINTER(i32), vp, vp, vp]),
        'dfn_decoder_program_host': (i32, [C.POINTER(DecoderDesc), C.POINTER(vp), i32, i32, i32, i32, C.POINTER(LayerInfo),
                                           C.POINTER(i32), vp, vp, C.POINTER(i32), C.POINTER(i32), vp, C.POINTER(i32),
                                           C.POINTER(i32), vp]),
        'dfn_gemm': (i32, [C.POINTER(GemmDesc), vp]),
        'dfn_colsum': (i32, [i64, i32, vp, i64, vp, i32, vp, vp]),
        'dfn_head_torso_loss_bwd': (i32, [i32, i32] + [vp] * 8 + [f32] + [vp] * 9 + [vp]),
        'dfn_adam_step': (i32, [i64, vp, vp, vp, vp, f32, f32, f32, f32, i32, vp]),
        'dfn_render_head_torso_workspace_bytes': (i64, [vp, i64, i32]),
        'dfn_render_head_torso': (i32, [vp, i64, i32, C.POINTER(HeadTorsoIO), i32, vp, i64, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib_, name)
        fn.restype = res
        fn.argtypes = args
    if lib_.dfn_abi_version() != 1:
        raise ImportError('dfa_nerf_b200: libdfn.so ABI mismatch')
    return lib_


lib = _load()
EXPORTS = ['dfn_abi_version', 'dfn_last_error', 'dfn_last_launch_count', 'dfn_profile_enable', 'dfn_profile_collect', 'dfn_debug_trace', 'dfn_debug_set_impl', 'dfn_debug_set_pp_flags', 'dfn_get_rays', 'dfn_z_vals', 'dfn_make_points', 'dfn_embed',
           'dfn_composite_fields', 'dfn_composite_head_torso', 'dfn_linear', 'dfn_calc_volume_weights', 'dfn_raw2outputs', 'dfn_sample_pdf', 'dfn_invert_cdf',
           'dfn_sort_merge', 'dfn_to8b', 'dfn_audionet_forward', 'dfn_att_smooth', 'dfn_pose_signal', 'dfn_model_create', 'dfn_model_destroy', 'dfn_model_num_tensors', 'dfn_model_load',
           'dfn_mlp_workspace_bytes', 'dfn_mlp_forward', 'dfn_query_workspace_bytes', 'dfn_query_points',
           'dfn_render_workspace_bytes', 'dfn_render_workspace_bytes2', 'dfn_render_rays', 'dfn_decoder_create', 'dfn_decoder_destroy',
           'dfn_decoder_num_tensors', 'dfn_decoder_load', 'dfn_decoder_query_workspace_bytes', 'dfn_decoder_query', 'dfn_decoder_query_ex',
           'dfn_decoder_macs_per_sample', 'dfn_decoder_program_host', 'dfn_model_program_host', 'dfn_render_head_torso_workspace_bytes', 'dfn_render_head_torso',
           'dfn_coarse_to_fine', 'dfn_gemm', 'dfn_colsum', 'dfn_head_torso_loss_bwd', 'dfn_adam_step']


def check(rc, what=''):
    if rc != 0:
        msg = lib.dfn_last_error()
        raise DfnError('%s failed (%d): %s' % (what or 'libdfn call', rc, msg.decode() if msg else ''))


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dev(t, name='tensor', dtype=torch.float32):
    """Validates a CUDA tensor for the ABI and returns (contiguous tensor, pointer)."""
    if not torch.is_tensor(t):
        raise TypeError('%s must be a torch tensor' % name)
    if not t.is_cuda:
        raise DfnError('dfa_nerf_b200 has no CPU path: %s must be a CUDA tensor' % name)
    if t.dtype != dtype:
        t = t.to(dtype)
    t = t.contiguous()
    return t, C.c_void_p(t.data_ptr())


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Workspace:
    """Grow-only device scratch buffer per (device, stream, tag); avoids allocator traffic per call.  Keyed by the
    current stream as well, so calls issued on different streams never share scratch memory."""
    _bufs = {}

    @classmethod
    def get(cls, nbytes, device, tag='default'):
        key = (str(device), torch.cuda.current_stream(device).cuda_stream, tag)
        buf = cls._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=device)
            cls._bufs[key] = buf
        return buf
