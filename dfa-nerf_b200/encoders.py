"""Latent encoders -- the callers immediately before the hot path.

Modules with the reference's constructor arguments and parameter names (so its checkpoints load unchanged):
``AudioNet`` (HELP:109-141), ``AudioNet_W2L`` (HELP:165-178), ``ExpressionEnc`` (HELP:182-193), ``AudioAttNet``
(HELP:210-240); ``forward`` runs on the CUDA kernels of libdfn (encoders.cu, dfn_linear).  Inference only.

The reference encodes one frame per Python iteration (MAIN:627-630): window slicing, zero padding, three small networks,
a dozen micro-kernels per frame.  ``encode_signal_sequence`` / ``encode_signal_torso_sequence`` do the same arithmetic
for a whole driven sequence in a handful of launches and return the ``[N, 96]`` / ``[N, 42]`` signal tables that
``render_sequence_head_torso`` consumes; ``encode_signal`` / ``encode_signal_torso`` keep the reference's per-frame
signatures on top of them.
"""
import ctypes as C

import torch
import torch.nn as nn

from ._lib import lib, check, dev, ptr, stream_ptr, DfnError, ConvStack
from .functional import get_embedder

LEAKY = 3   # dfn_linear activation code of nn.LeakyReLU(0.02)


def _linear(x, lin, act=0):
    P, K = x.shape
    out = torch.empty((P, lin.out_features), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.dfn_linear(P, lin.out_features, K, ptr(x), K, 0, None, 0, ptr(lin.weight), ptr(lin.bias), act, None, 0,
                             ptr(out), lin.out_features, stream_ptr()), 'dfn_linear')
    return out


def _conv_stack(convs, stride):
    cs = ConvStack()
    cs.n, cs.stride = len(convs), stride
    keep = []
    for i, c in enumerate(convs):
        if c.kernel_size != (3,) or c.padding != (1,) or c.stride != (stride,):
            raise DfnError('conv stack: kernel 3, padding 1, stride %d expected' % stride)
        cs.ch[i], cs.ch[i + 1] = c.in_channels, c.out_channels
        w, b = c.weight.detach().contiguous(), c.bias.detach().contiguous()
        keep += [w, b]
        cs.w[i], cs.b[i] = w.data_ptr(), b.data_ptr()
    return cs, keep


class _MlpEncoder(nn.Module):
    """Linear + LeakyReLU(0.02) chain under ``encoder.{0,2,4}`` (no activation after the last layer)."""

    def __init__(self, dims):
        super().__init__()
        layers = []
        for i in range(len(dims) - 1):
            layers.append(nn.Linear(dims[i], dims[i + 1]))
            if i + 2 < len(dims):
                layers.append(nn.LeakyReLU(0.02, True))
        self.encoder = nn.Sequential(*layers)

    @torch.no_grad()
    def forward(self, x):
        x, _ = dev(x.reshape(-1, x.shape[-1]), 'x')
        lins = [m for m in self.encoder if isinstance(m, nn.Linear)]
        for i, lin in enumerate(lins):
            x = _linear(x, lin, LEAKY if i + 1 < len(lins) else 0)
        return x


class AudioNet_W2L(_MlpEncoder):
    """HELP:165-178: Wav2Lip audio feature [N,512] -> [N,64]."""

    def __init__(self):
        super().__init__((512, 256, 128, 64))


class ExpressionEnc(_MlpEncoder):
    """HELP:182-193: expression code [N,64] -> [N,32]."""

    def __init__(self):
        super().__init__((64, 32, 32))


class AudioNet(nn.Module):
    """HELP:109-141: DeepSpeech windows [N,16,29] -> [N,dim_aud]."""

    def __init__(self, dim_aud=76, win_size=16):
        super().__init__()
        if win_size != 16:
            raise DfnError('AudioNet: win_size 16 only (the four stride-2 convs must reduce the window to one step)')
        self.win_size, self.dim_aud = win_size, dim_aud
        act = lambda: nn.LeakyReLU(0.02, True)  # noqa: E731
        self.encoder_conv = nn.Sequential(
            nn.Conv1d(29, 32, kernel_size=3, stride=2, padding=1, bias=True), act(),
            nn.Conv1d(32, 32, kernel_size=3, stride=2, padding=1, bias=True), act(),
            nn.Conv1d(32, 64, kernel_size=3, stride=2, padding=1, bias=True), act(),
            nn.Conv1d(64, 64, kernel_size=3, stride=2, padding=1, bias=True), act())
        self.encoder_fc1 = nn.Sequential(nn.Linear(64, 64), act(), nn.Linear(64, dim_aud))

    @torch.no_grad()
    def forward(self, x):
        x, px = dev(x, 'x')
        if x.dim() != 3 or tuple(x.shape[1:]) != (16, 29):
            raise DfnError('AudioNet.forward: x must be [N,16,29]')
        N = x.shape[0]
        cs, keep = _conv_stack([m for m in self.encoder_conv if isinstance(m, nn.Conv1d)], 2)
        fc1, fc2 = self.encoder_fc1[0], self.encoder_fc1[2]
        out = torch.empty((N, self.dim_aud), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            check(lib.dfn_audionet_forward(N, self.dim_aud, px, C.byref(cs), ptr(fc1.weight), ptr(fc1.bias), ptr(fc2.weight),
                                           ptr(fc2.bias), ptr(out), stream_ptr()), 'dfn_audionet_forward')
        return out


class AudioAttNet(nn.Module):
    """HELP:210-240: attention over a window of seq_len frames; weights from the first dim_aud columns."""

    def __init__(self, dim_aud=32, seq_len=8):
        super().__init__()
        self.seq_len, self.dim_aud = seq_len, dim_aud
        act = lambda: nn.LeakyReLU(0.02, True)  # noqa: E731
        ch = (dim_aud, 16, 8, 4, 2, 1)
        mods = []
        for i in range(5):
            mods += [nn.Conv1d(ch[i], ch[i + 1], kernel_size=3, stride=1, padding=1, bias=True), act()]
        self.attentionConvNet = nn.Sequential(*mods)
        self.attentionNet = nn.Sequential(nn.Linear(in_features=seq_len, out_features=seq_len, bias=True), nn.Softmax(dim=1))

    @torch.no_grad()
    def smooth_sequence(self, feats, pad_row):
        """feats [N,D] -> [N,D]: row i = forward(window of rows [i-seq_len/2, i+seq_len/2), pad_row outside [0,N))."""
        feats, pf = dev(feats, 'feats')
        pad_row, pp = dev(pad_row.reshape(-1), 'pad_row')
        N, D = feats.shape
        cs, keep = _conv_stack([m for m in self.attentionConvNet if isinstance(m, nn.Conv1d)], 1)
        lin = self.attentionNet[0]
        out = torch.empty((N, D), dtype=torch.float32, device=feats.device)
        with torch.cuda.device(feats.device):
            check(lib.dfn_att_smooth(N, D, self.dim_aud, self.seq_len, pf, pp, C.byref(cs), ptr(lin.weight), ptr(lin.bias),
                                     ptr(out), stream_ptr()), 'dfn_att_smooth')
        return out

    @torch.no_grad()
    def forward(self, x):
        """x [seq_len, D] -> [D] (the reference's signature): one window, centred in a padded 2*seq_len sequence."""
        x, _ = dev(x, 'x')
        if x.shape[0] != self.seq_len:
            raise DfnError('AudioAttNet.forward: x must have seq_len rows')
        half = self.seq_len // 2
        return self.smooth_sequence(x, torch.zeros(x.shape[1], device=x.device))[half]


# ------------------------------------------------------------------ sequence-level pre-pass


@torch.no_grad()
def encode_signal_sequence(auds, exps, AudNet, ExpNet, AudAttNet=None):
    """MAIN:28-68 (itr_obj == 0) for every frame at once: auds [N,512], exps [N,64] -> signals [N, 64+32].
    AudAttNet=None: the global_step < nosmo_iters branch (what scripts/test_obama.sh runs); otherwise the smoothing
    window + attention of MAIN:35-61 (zero INPUT rows outside the sequence, as the reference pads before encoding)."""
    auds, _ = dev(auds, 'auds')
    exps, _ = dev(exps, 'exps')
    n = auds.shape[0]
    if AudAttNet is None:
        return torch.cat([AudNet(auds), ExpNet(exps)], 1)
    a = AudNet(torch.cat([auds, torch.zeros_like(auds[:1])], 0))       # last row: the encoding of a zero input
    e = ExpNet(torch.cat([exps, torch.zeros_like(exps[:1])], 0))
    feats = torch.cat([a, e], 1)
    return AudAttNet.smooth_sequence(feats[:n].contiguous(), feats[n].contiguous())


@torch.no_grad()
def pose_to_euler_trans(poses):
    """MAIN:202-205 on the device: poses [N,3|4,4] -> [N,6] = euler | translation."""
    return _pose_signal(poses, 0)[1]


def _pose_signal(poses, L):
    poses, pp = dev(poses, 'poses')
    if poses.dim() != 3 or poses.shape[2] != 4 or poses.shape[1] not in (3, 4):
        raise DfnError('poses must be [N,3,4] or [N,4,4]')
    N = poses.shape[0]
    out = torch.empty((N, 2 * (3 + 6 * L)), dtype=torch.float32, device=poses.device)
    et = torch.empty((N, 6), dtype=torch.float32, device=poses.device)
    with torch.cuda.device(poses.device):
        check(lib.dfn_pose_signal(N, pp, poses.shape[1] * 4, L, ptr(out), ptr(et), stream_ptr()), 'dfn_pose_signal')
    return out, et


@torch.no_grad()
def encode_signal_torso_sequence(poses, PoseAttNet=None, multires=3):
    """MAIN:78-111 for every frame at once: head poses [N,3|4,4] -> torso signals [N, 2*(3+6*multires)] (42)."""
    sig, _ = _pose_signal(poses, multires)
    if PoseAttNet is None:
        return sig
    embed_fn, _ = get_embedder(multires, 0)
    pad = embed_fn(torch.zeros((1, 3), device=sig.device))              # the reference pads euler/trans rows with zeros
    return PoseAttNet.smooth_sequence(sig, torch.cat([pad, pad], 1).reshape(-1))


# ------------------------------------------------------------------ the reference's per-frame signatures


def encode_signal(dataset, itr_obj, img_i, dim_aud, AudNet, ExpNet, AudAttNet, global_step, args, len_auds, embed_fn=None):
    """MAIN:28-75.  Returns [aud [1,96], None] for the talking head (itr_obj == 0), [None, exp] otherwise."""
    if itr_obj != 0:
        return [None, dataset[itr_obj]['exp'][img_i:img_i + 1]]
    auds, exps = dataset[itr_obj]['auds'], dataset[itr_obj]['exp']
    if global_step < args.nosmo_iters:
        return [encode_signal_sequence(auds[img_i:img_i + 1], exps[img_i:img_i + 1], AudNet, ExpNet), None]
    half = int(args.smo_size / 2)
    lo, hi = max(img_i - half, 0), min(img_i + half, len_auds)
    sig = encode_signal_sequence(auds[lo:hi], exps[lo:hi], AudNet, ExpNet, None)
    a0 = torch.cat([AudNet(torch.zeros_like(auds[:1])), ExpNet(torch.zeros_like(exps[:1]))], 1)
    win = torch.cat([a0.expand(max(half - img_i, 0), -1), sig, a0.expand(max(img_i + half - len_auds, 0), -1)], 0)
    return [AudAttNet(win.contiguous()).unsqueeze(0), None]


def encode_signal_torso(dataset, itr_obj, img_i, PoseAttNet, global_step, args, len_poses, embed_fn=None):
    """MAIN:78-111."""
    poses = dataset[itr_obj]['poses']
    if global_step < args.nosmo_iters:
        return encode_signal_torso_sequence(poses[img_i:img_i + 1])
    half = int(args.smo_torse_size / 2)
    lo, hi = max(img_i - half, 0), min(img_i + half, len_poses)
    sig = encode_signal_torso_sequence(poses[lo:hi])
    emb, _ = get_embedder(3, 0)
    z = emb(torch.zeros((1, 3), device=sig.device))
    pad = torch.cat([z, z], 1)
    win = torch.cat([pad.expand(max(half - img_i, 0), -1), sig, pad.expand(max(img_i + half - len_poses, 0), -1)], 0)
    return PoseAttNet(win.contiguous())
