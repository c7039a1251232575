"""NeRF / FaceNeRF modules with the reference's constructor signatures and parameter names
(HELP:242-273, HELP:342-370) so ``load_state_dict`` of a reference checkpoint works unchanged.
``forward`` runs the CUDA fp32 path (dfn_mlp_forward); the render pipeline uses the same handle
for the fused tcgen05 path.  Inference only: outputs carry no autograd graph.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from ._lib import lib, check, dev, ptr, stream_ptr, DfnError, ModelDesc, Workspace


class _DfnMLP(nn.Module):
    _kind = None

    def _build(self, D, W, input_ch, input_ch_views, dim_aud, skips, use_viewdirs):
        if not use_viewdirs:
            raise DfnError('dfa_nerf_b200 builds the use_viewdirs=True head only (the one the render path uses)')
        if len(skips) != 1:
            raise DfnError('exactly one skip connection is supported (reference default skips=[4])')
        self.D, self.W = D, W
        self.input_ch, self.input_ch_views, self.dim_aud = input_ch, input_ch_views, dim_aud
        self.skips, self.use_viewdirs = list(skips), use_viewdirs
        n_in = input_ch + dim_aud
        self.pts_linears = nn.ModuleList(
            [nn.Linear(n_in, W)] + [nn.Linear(W + n_in, W) if i in self.skips else nn.Linear(W, W) for i in range(D - 1)])
        n_view = 1 + D // 4 if self._kind == _lib.MODEL_FACENERF else 1
        self.views_linears = nn.ModuleList(
            [nn.Linear(input_ch_views + W, W // 2)] + [nn.Linear(W // 2, W // 2) for _ in range(n_view - 1)])
        self.feature_linear = nn.Linear(W, W)
        self.alpha_linear = nn.Linear(W, 1)
        self.rgb_linear = nn.Linear(W // 2, 3)
        self._handle = None
        self._loaded_sig = None

    # -- libdfn handle -------------------------------------------------------------------
    def _tensor_list(self):
        mods = list(self.pts_linears) + list(self.views_linears) + [self.feature_linear, self.alpha_linear, self.rgb_linear]
        out = []
        for m in mods:
            out += [m.weight, m.bias]
        return out

    def dfn_handle(self, device=None):
        """Creates the native model on first use and re-uploads when parameters changed."""
        params = self._tensor_list()
        device = device or params[0].device
        if torch.device(device).type != 'cuda':
            raise DfnError('dfa_nerf_b200 has no CPU path: move the module to CUDA')
        sig = (str(device),) + tuple((p.data_ptr(), p._version) for p in params)
        if self._handle is not None and sig == self._loaded_sig:
            return self._handle
        if self._handle is None:
            if (self.input_ch - 3) % 6 or (self.input_ch_views - 3) % 6:
                raise DfnError('input_ch / input_ch_views must be 3+6*multires')
            desc = ModelDesc(self._kind, self.D, self.W, self.input_ch, self.input_ch_views, self.dim_aud,
                             self.skips[0], (self.input_ch - 3) // 6, (self.input_ch_views - 3) // 6)
            h = C.c_void_p()
            check(lib.dfn_model_create(C.byref(desc), C.byref(h)), 'dfn_model_create')
            self._handle = h
        host = [p.detach().to('cpu', torch.float32).contiguous() for p in params]
        arr = (C.c_void_p * len(host))(*[t.data_ptr() for t in host])
        with torch.cuda.device(device):
            check(lib.dfn_model_load(self._handle, arr, len(host), stream_ptr()), 'dfn_model_load')
        self._loaded_sig = sig
        return self._handle

    def __del__(self):
        h = getattr(self, '_handle', None)
        if h is not None:
            try:
                lib.dfn_model_destroy(h)
            except Exception:
                pass

    def forward(self, x):
        """x [..., input_ch+dim_aud+input_ch_views] -> [..., 4] = (rgb pre-sigmoid, sigma)."""
        in_dim = self.input_ch + self.dim_aud + self.input_ch_views
        if x.shape[-1] != in_dim:
            raise DfnError('expected last dim %d, got %d' % (in_dim, x.shape[-1]))
        xc, px = dev(x.reshape(-1, in_dim), 'x')
        h = self.dfn_handle(xc.device)
        P = xc.shape[0]
        out = torch.empty((P, 4), dtype=torch.float32, device=xc.device)
        if P:
            nbytes = lib.dfn_mlp_workspace_bytes(h, P)
            ws = Workspace.get(nbytes, xc.device, 'mlp')
            with torch.cuda.device(xc.device):
                check(lib.dfn_mlp_forward(h, P, px, ptr(out), ptr(ws), nbytes, stream_ptr()), 'dfn_mlp_forward')
        return out.reshape(x.shape[:-1] + (4,))


class FaceNeRF(_DfnMLP):
    """HELP:242-299: audio-conditioned 8x256 skip-MLP; feature_linear exists but is bypassed; 3 view layers."""
    _kind = _lib.MODEL_FACENERF

    def __init__(self, D=8, W=256, input_ch=3, input_ch_views=3, dim_aud=76, output_ch=4, skips=[4], use_viewdirs=False):
        super().__init__()
        self._build(D, W, input_ch, input_ch_views, dim_aud, skips, use_viewdirs)


class NeRF(_DfnMLP):
    """HELP:342-396: 8x256 skip-MLP, feature_linear applied, one view layer."""
    _kind = _lib.MODEL_NERF

    def __init__(self, D=8, W=256, input_ch=3, input_ch_views=3, output_ch=4, skips=[4], use_viewdirs=False):
        super().__init__()
        self._build(D, W, input_ch, input_ch_views, 0, skips, use_viewdirs)
