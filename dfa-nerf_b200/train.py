"""The reference's training step on the B200 (MAIN:764-931): ray selection, the two-field
Decoder forward on N_rand rays x N_samples, two-field compositing, the two image losses, the backward pass and the Adam
updates of the decoder and the latent encoders.  No torch.autograd: the step is an explicit tape over libdfn kernels --

  * every nn.Linear, forward and backward, is `dfn_gemm` (csrc/gemm_tc.cu): a strided tcgen05 GEMM that converts its fp32
    operands to split bf16 on the way into shared memory and applies the activation derivative while it loads the
    gradient operand, so  Y = act(X W^T + b),  dX = (dH * act'(Y)) W  and  dW += (dH * act'(Y))^T X  need no transposed or
    masked copies;
  * per-frame inputs (the audio / expression signal, z_shape, z_app, the torso signal) never become [P, .] tensors: they
    enter as bias vectors (VNode) and their gradients are column sums (`dfn_colsum`) pushed through tiny GEMMs;
  * compositing + losses + their gradient are one kernel (`dfn_head_torso_loss_bwd`, a reverse per-ray scan);
  * each optimiser group lives in one flat buffer and is stepped by one `dfn_adam_step` launch.

Parameter names follow the reference's modules, so `Trainer.grads['dec']['blocks.6.weight']` is what
`decoder.blocks[6].weight.grad` holds after MAIN:923.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, ptr, stream_ptr, DfnError, GemmDesc as _GemmDesc
from .functional import get_rays, z_vals_uniform, make_points, decoder_transform_points
from .encoders import encode_signal_torso_sequence

ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_LEAKY = 0, 1, 2, 3
_MASK_OF_ACT = {ACT_NONE: 0, ACT_RELU: 1, ACT_LEAKY: 2, ACT_SIGMOID: 3}


def select_coords(H, W, rect, N_rand, sample_rate, rng=np.random):
    """MAIN:787-820: `sample_rate` of the rays inside the face rectangle | lower image half, the rest outside; consumes the
    numpy RNG exactly like the reference (two choice() calls, or one when sample_rate == 0).  Host side, returns int64 [N_rand,2]
    = (row, column)."""
    rows, cols = np.meshgrid(np.arange(H), np.arange(W), indexing='ij')
    coords = np.stack([rows, cols], -1).reshape(-1, 2)
    if sample_rate > 0:
        def inside(r):
            return (coords[:, 0] >= r[0]) & (coords[:, 0] <= r[0] + r[2]) & (coords[:, 1] >= r[1]) & (coords[:, 1] <= r[1] + r[3])
        sel = inside(rect) | inside([1 * H / 2, 0, H / 2, W])
        c_in, c_out = coords[sel], coords[~sel]
        n_in = int(N_rand * sample_rate)
        a = rng.choice(c_in.shape[0], size=[n_in], replace=False)
        b = rng.choice(c_out.shape[0], size=[N_rand - n_in], replace=False)
        return torch.from_numpy(np.concatenate([c_in[a], c_out[b]], 0).astype(np.int64))
    return torch.from_numpy(coords[rng.choice(coords.shape[0], size=[N_rand], replace=False)].astype(np.int64))


N_SM = 148


def mm(A, B, Cm, bias=None, addend=None, pre_add=False, act=ACT_NONE, mask=None, mask_mode=0, beta=0, k_splits=1,
       precision=_lib.PREC_BF16X3):
    """Cm[m,n] = act(sum_k A[m,k] B[n,k] + bias[n] (+ addend)) (+ addend) (+ Cm).  A [M,K], B [N,K], Cm [M,N]: 2-D CUDA fp32 views
    with arbitrary strides (pass .t() views for the transposed products); mask: a view shaped and strided like A."""
    M, K = A.shape
    N = B.shape[0]
    if B.shape[1] != K or tuple(Cm.shape) != (M, N):
        raise DfnError('mm: shapes %s x %s^T -> %s' % (tuple(A.shape), tuple(B.shape), tuple(Cm.shape)))
    if mask is not None and (tuple(mask.shape) != (M, K) or mask.stride() != A.stride()):
        raise DfnError('mm: the mask must be shaped and strided like A')
    launches = 0
    for n0 in range(0, N, 256):
        n1 = min(N, n0 + 256)
        d = _GemmDesc()
        d.A, d.a_ld_r, d.a_ld_k = A.data_ptr(), A.stride(0), A.stride(1)
        d.A_mask, d.a_mask_mode = (mask.data_ptr(), mask_mode) if mask is not None and mask_mode else (None, 0)
        Bv, Cv = B[n0:n1], Cm[:, n0:n1]
        d.B, d.b_ld_r, d.b_ld_k = Bv.data_ptr(), Bv.stride(0), Bv.stride(1)
        d.C, d.c_ld_r, d.c_ld_c = Cv.data_ptr(), Cv.stride(0), Cv.stride(1)
        d.bias = bias[n0:n1].data_ptr() if bias is not None else None
        if addend is not None:
            av = addend[:, n0:n1]
            d.addend, d.add_ld_r, d.add_ld_c = av.data_ptr(), av.stride(0), av.stride(1)
        d.act = act | (4 if pre_add else 0)
        d.M, d.N, d.K, d.beta, d.k_splits, d.precision = M, n1 - n0, K, beta, k_splits, precision
        check(lib.dfn_gemm(C.byref(d), stream_ptr()), 'dfn_gemm')
        launches += 1
    return launches


def colsum(X, Y, mask_mode, out):
    """out[n] += sum_m X[m,n] * act'(Y[m,n]); X, Y [M,N] row-major with the same leading dimension."""
    M, N = X.shape
    if X.stride(1) != 1 or (Y is not None and Y.stride() != X.stride()):
        raise DfnError('colsum: row-major operands with equal strides expected')
    check(lib.dfn_colsum(M, N, X.data_ptr(), X.stride(0), Y.data_ptr() if (Y is not None and mask_mode) else None, mask_mode,
                         out.data_ptr(), stream_ptr()), 'dfn_colsum')


def loss_bwd(R, S, feat_h, sig_h, feat_t, sig_t, bc, z, rd_h, rd_t, tgt_head, tgt_person, loss2, rgb_head, rgb_person, g_feat_h,
             g_sig_h, g_feat_t, g_sig_t, last_dist=1e10):
    """dfn_head_torso_loss_bwd: compositing of MAIN:884-905, loss2 += (img2mse head, img2mse person), and the gradient of their sum
    with respect to the fields' pre-sigmoid colours and raw densities."""
    check(lib.dfn_head_torso_loss_bwd(R, S, ptr(feat_h), ptr(sig_h), ptr(feat_t), ptr(sig_t), ptr(bc), ptr(z), ptr(rd_h), ptr(rd_t),
                                      float(last_dist), ptr(tgt_head), ptr(tgt_person), ptr(loss2), ptr(rgb_head), ptr(rgb_person),
                                      ptr(g_feat_h), ptr(g_sig_h), ptr(g_feat_t), ptr(g_sig_t), stream_ptr()), 'dfn_head_torso_loss_bwd')


def adam_step(flat, grad, m, v, lr, betas, eps, step):
    check(lib.dfn_adam_step(flat.numel(), flat.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), float(lr), float(betas[0]),
                            float(betas[1]), float(eps), int(step), stream_ptr()), 'dfn_adam_step')


class _Group:
    """One optimiser group in one flat buffer: parameters (the modules' tensors are re-pointed at views of it), gradients
    and Adam moments.  MAIN:522-535 builds one torch.optim.Adam per network; a step is one dfn_adam_step launch."""

    def __init__(self, module, device):
        named = list(module.named_parameters())
        n = sum(p.numel() for _, p in named)
        self.flat = torch.empty(n, dtype=torch.float32, device=device)
        self.grad = torch.zeros(n, dtype=torch.float32, device=device)
        self.m = torch.zeros(n, dtype=torch.float32, device=device)
        self.v = torch.zeros(n, dtype=torch.float32, device=device)
        self.p, self.g, o = {}, {}, 0
        for name, p in named:
            k = p.numel()
            view = self.flat[o:o + k].view(p.shape)
            view.copy_(p.data)
            p.data = view                      # the module now reads (and the render path re-packs) the trained weights
            self.p[name], self.g[name] = view, self.grad[o:o + k].view(p.shape)
            o += k
        self.steps, self.module = 0, module

    def step(self, lr, betas=(0.9, 0.999), eps=1e-8):
        self.steps += 1
        adam_step(self.flat, self.grad, self.m, self.v, lr, betas, eps, self.steps)
        for m in self.module.modules():         # the native inference handles re-pack on their next use
            if hasattr(m, '_loaded_sig'):
                m._loaded_sig = None


class VNode:
    """Per-frame vector value = act(sum of terms), terms: a bias parameter or W[:, cols] @ (another vector)."""

    def __init__(self, tape, n, terms, act=ACT_NONE):
        self.tape, self.n, self.terms, self.act = tape, n, terms, act
        dev = tape.device
        self.value = torch.empty(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        out = self.value[None, :]
        bias = [t for t in terms if t[0] == 'b']
        mvs = [t for t in terms if t[0] == 'mv']
        bvec = None
        if len(bias) == 1:
            bvec = bias[0][1]
        elif len(bias) > 1:
            bvec = torch.stack([b[1] for b in bias], 0).sum(0)           # (a few hundred floats)
        if not mvs:
            self.value.copy_(bvec)
            assert act == ACT_NONE
        for i, (_, W, gW, src) in enumerate(mvs):
            x = (src.value if isinstance(src, VNode) else src).reshape(1, -1)
            last = i == len(mvs) - 1
            if len(mvs) == 1:
                tape.count(mm(x, W, out, bias=bvec, act=act, precision=tape.precision))
            elif i == 0:
                tape.count(mm(x, W, out, bias=bvec, precision=tape.precision))
            elif not last or act == ACT_NONE:
                tape.count(mm(x, W, out, beta=1, precision=tape.precision))
            else:
                tmp = out.clone()
                tape.count(mm(x, W, out, addend=tmp, pre_add=True, act=act, precision=tape.precision))
        tape.vnodes.append(self)

    def backward(self):
        tape = self.tape
        g = self.grad[None, :]
        mode = _MASK_OF_ACT[self.act]
        y = self.value[None, :] if mode else None
        for t in self.terms:
            if t[0] == 'b':
                if t[2] is not None:
                    colsum(g, y, mode, t[2])
                    tape.count(1)
            else:
                _, W, gW, src = t
                x = (src.value if isinstance(src, VNode) else src).reshape(1, -1)
                if gW is not None:       # gW [n, kv] += outer(da, x):  A = da^T [n,1], B = x^T [kv,1]
                    tape.count(mm(g.t(), x.t(), gW, mask=y.t() if mode else None, mask_mode=mode, beta=1, precision=tape.precision))
                if isinstance(src, VNode):   # src.grad [1,kv] += da [1,n] @ W [n,kv]
                    tape.count(mm(g, W.t(), src.grad[None, :], mask=y, mask_mode=mode, beta=1, precision=tape.precision))


class _PT:
    """A per-point tensor [P, n] that receives a gradient."""
    Y = None
    gbuf = None
    g_written = False
    bwd_act = ACT_NONE

    def grad_buffer(self):
        if self.gbuf is None:
            self.gbuf = torch.empty_like(self.Y)
        return self.gbuf

    def backward(self):
        pass


class _Joined(_PT):
    """[P, n1 + n2] buffer whose column blocks are written by two layers (the deformation field's two branches, DEC:132-134):
    its gradient buffer is handed to the producers as column-slice views."""

    def __init__(self, Y, producers):
        self.Y, self.producers = Y, producers       # [(node, col0, col1)]

    def grad_buffer(self):
        if self.gbuf is None:
            self.gbuf = torch.empty_like(self.Y)
            for node, c0, c1 in self.producers:
                if node.gbuf is not None:
                    raise DfnError('tape: a joined producer already has a gradient')
                node.gbuf, node.g_written = self.gbuf[:, c0:c1], True
        return self.gbuf


class PNode(_PT):
    """Per-point layer  Y = act(X W^T + bias (+ addend before the activation)) (+ addend after it)."""

    def __init__(self, tape, X, W, gW, bias, act=ACT_NONE, addend=None, pre_add=False, out=None, needs_input_grad=True, bwd_act=None):
        self.tape, self.X, self.W, self.gW, self.bias, self.act = tape, X, W, gW, bias, act
        self.addend, self.pre_add = addend, pre_add
        self.bwd_act = act if bwd_act is None else bwd_act
        x = X.Y if isinstance(X, _PT) else X
        P = x.shape[0]
        self.Y = out if out is not None else torch.empty((P, W.shape[0]), dtype=torch.float32, device=tape.device)
        add = addend.Y if isinstance(addend, _PT) else addend
        bvec = bias.value if isinstance(bias, VNode) else bias
        tape.count(mm(x, W, self.Y, bias=bvec, addend=add, pre_add=pre_add, act=act, precision=tape.precision))
        self.gbuf, self.gmask, self.gmode, self.g_written = None, None, 0, False
        self.needs_input_grad = needs_input_grad and isinstance(X, _PT)
        tape.pnodes.append(self)

    def backward(self):
        tape = self.tape
        if self.gbuf is None:
            return
        A = self.gbuf
        if self.bwd_act != ACT_NONE:
            M, mode = self.Y, _MASK_OF_ACT[self.bwd_act]
        else:
            M, mode = self.gmask, self.gmode
        x = self.X.Y if isinstance(self.X, _PT) else self.X
        P = x.shape[0]
        if self.gW is not None:
            m_tiles = (self.W.shape[0] + 127) // 128
            # split-K partial tiles are reduced with fp32 atomics (M x N per split): one wave of CTAs, not two -- the atomics, not the
            # K loop, were most of this GEMM's time
            splits = max(1, min((P + 63) // 64, N_SM // m_tiles))
            tape.count(mm(A.t(), x.t(), self.gW, mask=M.t() if M is not None else None, mask_mode=mode, beta=1, k_splits=splits,
                          precision=tape.precision))
        if isinstance(self.bias, VNode):
            colsum(A, M, mode, self.bias.grad)
            tape.count(1)
        if self.needs_input_grad:
            X = self.X
            tape.count(mm(A, self.W.t(), X.grad_buffer(), mask=M, mask_mode=mode, beta=1 if X.g_written else 0, precision=tape.precision))
            X.g_written = True
        if isinstance(self.addend, _PT):
            a = self.addend
            if a.gbuf is not None:
                raise DfnError('tape: an addend must have this layer as its only consumer')
            if self.pre_add:        # d addend = dH * act'(Y): share the buffer, carry the mask
                if a.bwd_act != ACT_NONE:
                    raise DfnError('tape: a pre-activation addend must itself be linear')
                a.gbuf, a.gmask, a.gmode, a.g_written = A, M, mode, True
            else:                   # added after the activation: d addend = dH
                if self.bwd_act != ACT_NONE:
                    raise DfnError('tape: a post-activation addend needs a linear layer')
                a.gbuf, a.gmask, a.gmode, a.g_written = A, self.gmask, self.gmode, True


class _Tape:
    def __init__(self, device, precision):
        self.device, self.precision = device, precision
        self.pnodes, self.vnodes, self.launches = [], [], 0

    def count(self, n):
        self.launches += n

    def backward(self):
        for n in reversed(self.pnodes):
            n.backward()
        for v in reversed(self.vnodes):
            v.backward()


class Trainer:
    """trainer = Trainer(decoder, AudNet, ExpNet, lrate=5e-4); loss = trainer.step(batch, global_step)

    decoder: dfa_nerf_b200.Decoder (use_deformation_field=True); AudNet / ExpNet: AudioNet_W2L / ExpressionEnc (MAIN:497-535).
    batch (as oracle/train_oracle.train_losses): H, W, focal, cx, cy, near, far, pose [3,4], pose_torso [3,4], poses [N,4,4],
    img_i, auds [N,512], exps [N,64], coords [N_rand,2] (row, column), target_com / target_head_neck / bc_img [H,W,3] (CUDA),
    z_shape / z_app [1,2,z_dim].  The step is the `global_step < nosmo_iters` branch (no attention nets, MAIN:31, 81)."""

    def __init__(self, decoder, AudNet, ExpNet, lrate=5e-4, betas=(0.9, 0.999), N_samples=64, precision=_lib.PREC_BF16X3,
                 lrate_decay=0, decay_rate=0.1, device='cuda'):
        self.device = torch.device(device)
        if not decoder.use_deformation_field:
            raise DfnError('Trainer: the reference trains Decoder(use_deformation_field=True) (MAIN:518)')
        self.decoder, self.AudNet, self.ExpNet = decoder.to(self.device), AudNet.to(self.device), ExpNet.to(self.device)
        self.groups = {'dec': _Group(self.decoder, self.device), 'aud': _Group(self.AudNet, self.device),
                       'exp': _Group(self.ExpNet, self.device)}
        self.params = {k: g.p for k, g in self.groups.items()}
        self.grads = {k: g.g for k, g in self.groups.items()}
        self.lrate, self.betas, self.N_samples, self.precision = lrate, betas, N_samples, precision
        self.lrate_decay, self.decay_rate = lrate_decay, decay_rate
        self.loss2 = torch.zeros(2, dtype=torch.float32, device=self.device)
        self.last_launches = 0
        self.last = {}

    # ------------------------------------------------------------------------------------------------ forward pieces
    def _encoder(self, tape, key, x):
        """AudioNet_W2L / ExpressionEnc (HELP:165-193): Linear + LeakyReLU(0.02) chain, no activation after the last layer."""
        p, g = self.params[key], self.grads[key]
        idx = sorted({int(k.split('.')[1]) for k in p})
        node = x
        for j, i in enumerate(idx):
            W, b = p['encoder.%d.weight' % i], p['encoder.%d.bias' % i]
            node = VNode(tape, W.shape[0], [('b', b, g['encoder.%d.bias' % i]), ('mv', W, g['encoder.%d.weight' % i], node)],
                         ACT_LEAKY if j + 1 < len(idx) else ACT_NONE)
        return node

    def _field(self, tape, which, pe, pev, sig_terms, z_shape, z_app):
        """Decoder.forward (DEC:277-349) of one field on P points.  pe [P,60], pev [P,24]; sig_terms(name) -> the 'mv' terms that
        put the field's per-frame signal columns of weight `name` into a bias vector (head), or None (torso: the deformed
        signal is per point).  Returns (feat node, sigma node)."""
        p, g = self.params['dec'], self.grads['dec']
        de = pe.shape[1]
        H = self.decoder.hidden_size

        def par(name):
            return p[name + '.weight'], g[name + '.weight'], p[name + '.bias'], g[name + '.bias']

        torso = which == 'torso'
        if torso:
            dt = self.decoder.dim_et_embed
            sig_t = sig_terms          # the constant torso signal [dt]
            xin = self._deform(tape, pe, sig_t)                         # PNode-like: Y [P, de+dt], requires grad
            fcin, pskip = 'fc_in_torso', 'fc_p_skips_torso.0'
        else:
            fcin, pskip = 'fc_in', 'fc_p_skips.0'
        Wi, gWi, bi, gbi = par(fcin)
        Wz, gWz, bz, gbz = par('fc_z')
        terms = [('b', bi, gbi), ('b', bz, gbz), ('mv', Wz, gWz, z_shape)]
        if torso:
            net = PNode(tape, xin, Wi, gWi, VNode(tape, H, terms), ACT_RELU)
        else:
            terms += sig_terms(Wi, gWi, de)
            net = PNode(tape, pe, Wi[:, :de], gWi[:, :de], VNode(tape, H, terms), ACT_RELU)
        n_blocks = len(self.decoder.blocks)
        for idx in range(n_blocks):
            W, gW, b, gb = par('blocks.%d' % idx)
            net = PNode(tape, net, W, gW, VNode(tape, H, [('b', b, gb)]), ACT_RELU)
            if (idx + 1) in self.decoder.skips and idx < n_blocks - 1:
                Wp, gWp, bp, gbp = par(pskip)
                Ws, gWs, bs, gbs = par('fc_z_skips.0')
                terms = [('b', bp, gbp), ('b', bs, gbs), ('mv', Ws, gWs, z_shape)]
                if torso:
                    net = PNode(tape, xin, Wp, gWp, VNode(tape, H, terms), ACT_NONE, addend=net)
                else:
                    terms += sig_terms(Wp, gWp, de)
                    net = PNode(tape, pe, Wp[:, :de], gWp[:, :de], VNode(tape, H, terms), ACT_NONE, addend=net)
        W, gW, b, gb = par('sigma_out')
        sigma = PNode(tape, net, W, gW, VNode(tape, 1, [('b', b, gb)]), ACT_NONE)
        Wf, gWf, bf, gbf = par('feat_view')
        Wzv, gWzv, bzv, gbzv = par('fc_z_view')
        Wv, gWv, bv, gbv = par('fc_view')
        t = PNode(tape, net, Wf, gWf, VNode(tape, H, [('b', bf, gbf), ('b', bzv, gbzv), ('b', bv, gbv), ('mv', Wzv, gWzv, z_app)]), ACT_NONE)
        hv = PNode(tape, pev, Wv, gWv, None, ACT_RELU, addend=t, pre_add=True)
        Wo, gWo, bo, gbo = par('feat_out')
        feat = PNode(tape, hv, Wo, gWo, VNode(tape, 3, [('b', bo, gbo)]), ACT_SIGMOID, bwd_act=ACT_NONE)   # the loss kernel returns
        return feat, sigma                                                                             # d(pre-sigmoid)

    def _deform(self, tape, pe, sig_t):
        """p <- deform_net(p) + p on p = [PE | torso signal] (DEC:297-299, DEC:109-134).  The per-frame signal enters the first
        layers as a bias vector; the signal branch's additive skip is a vector, composed into the next layer's bias."""
        p, g = self.params['dec'], self.grads['dec']
        de, dt = pe.shape[1], sig_t.shape[0]
        P = pe.shape[0]
        out = torch.empty((P, de + dt), dtype=torch.float32, device=self.device)

        def par(name):
            return p['deform_net.%s.weight' % name], g['deform_net.%s.weight' % name], p['deform_net.%s.bias' % name], \
                g['deform_net.%s.bias' % name]

        producers = []
        for br, skipname, odim, col0 in (('embed', 'fc_embed_skips.0', de, 0), ('signal', 'fc_signal_skips.0', dt, de)):
            W, gW, b, gb = par('blocks_%s.0' % br)
            net = PNode(tape, pe, W[:, :de], gW[:, :de], VNode(tape, 64, [('b', b, gb), ('mv', W[:, de:], gW[:, de:], sig_t)]), ACT_RELU)
            n = 5
            extra = None
            for idx in range(1, n):
                W, gW, b, gb = par('blocks_%s.%d' % (br, idx))
                terms = [('b', b, gb)]
                if extra is not None:                     # (net + v) W^T = net W^T + W v: the vector skip of the signal branch
                    terms.append(('mv', W, gW, extra))
                    extra = None
                net = PNode(tape, net, W, gW, VNode(tape, 64, terms), ACT_RELU)
                if (idx + 1) in (4,) and idx < n - 1:     # DEC:118-121, DEC:128-131: skips=[4]
                    Ws, gWs, bs, gbs = par(skipname)
                    if br == 'embed':
                        net = PNode(tape, pe, Ws, gWs, VNode(tape, 64, [('b', bs, gbs)]), ACT_NONE, addend=net)
                    else:
                        extra = VNode(tape, 64, [('b', bs, gbs), ('mv', Ws, gWs, sig_t)])
            W, gW, b, gb = par('out_%s' % br)
            res = pe if br == 'embed' else sig_t[None, :].expand(P, dt)            # + p (DEC:299); stride-0 rows for the signal
            node = PNode(tape, net, W, gW, VNode(tape, odim, [('b', b, gb)]), ACT_NONE, addend=res, out=out[:, col0:col0 + odim])
            producers.append((node, col0, col0 + odim))
        return _Joined(out, producers)

    # ------------------------------------------------------------------------------------------------------- the step
    def losses_and_grads(self, batch, N_samples=None):
        """Forward + backward of MAIN:822-923 (global_step < nosmo_iters).  Returns the loss tensor [1] on the device; the
        gradients are left in self.grads (zeroed at the start, like optimizer.zero_grad() at MAIN:909-921)."""
        S = N_samples or self.N_samples
        dev = self.device
        for grp in self.groups.values():
            grp.grad.zero_()
        self.loss2.zero_()
        tape = _Tape(dev, self.precision)
        c = batch['coords'].to(dev)
        R = c.shape[0]
        H, W = int(batch['H']), int(batch['W'])
        pick = lambda img: img.to(dev)[c[:, 0], c[:, 1]].contiguous()                               # noqa: E731
        ro, rd = [pick(t) for t in get_rays(H, W, batch['focal'], batch['pose'], batch['cx'], batch['cy'], device=dev)]
        rot, rdt = [pick(t) for t in get_rays(H, W, batch['focal'], batch['pose_torso'], batch['cx'], batch['cy'], device=dev)]
        z = z_vals_uniform(torch.full((R,), float(batch['near']), device=dev), torch.full((R,), float(batch['far']), device=dev), S)
        bc, tgt_com, tgt_head = pick(batch['bc_img']), pick(batch['target_com']), pick(batch['target_head_neck'])
        # per-frame latents
        i = int(batch['img_i'])
        aud = self._encoder(tape, 'aud', batch['auds'][i].to(dev).float().contiguous())
        exp = self._encoder(tape, 'exp', batch['exps'][i].to(dev).float().contiguous())
        n_aud = aud.n
        sig_torso = encode_signal_torso_sequence(batch['poses'][i:i + 1, :3, :4].to(dev).float().contiguous())[0].contiguous()   # MAIN:78-84, no parameters
        zs, za = batch['z_shape'].to(dev).float(), batch['z_app'].to(dev).float()

        def head_signal(Wfull, gWfull, de):      # [aud (64) | exp (32)] columns of a weight that reads [PE | signal] (DEC:293-295)
            return [('mv', Wfull[:, de:de + n_aud], gWfull[:, de:de + n_aud], aud),
                    ('mv', Wfull[:, de + n_aud:], gWfull[:, de + n_aud:], exp)]

        fields = {}
        for which, o, d_, k in (('head', ro, rd, 0), ('torso', rot, rdt, 1)):
            pts, dirs = make_points(o, d_, z)
            pe = decoder_transform_points(pts.reshape(-1, 3), self.decoder.n_freq_posenc)
            pev = decoder_transform_points(dirs.reshape(-1, 3), self.decoder.n_freq_posenc_views, normalize=True)
            fields[which] = self._field(tape, which, pe, pev, head_signal if which == 'head' else sig_torso,
                                        zs[0, k].contiguous(), za[0, k].contiguous())
        (fh, sh), (ft, st) = fields['head'], fields['torso']
        P = R * S
        rgb_head = torch.empty((R, 3), dtype=torch.float32, device=dev)
        rgb_person = torch.empty((R, 3), dtype=torch.float32, device=dev)
        loss_bwd(R, S, fh.Y, sh.Y, ft.Y, st.Y, bc, z, rd, rdt, tgt_head, tgt_com, self.loss2, rgb_head, rgb_person,
                 fh.grad_buffer(), sh.grad_buffer(), ft.grad_buffer(), st.grad_buffer())
        tape.count(1)
        for n in (fh, sh, ft, st):
            n.g_written = True
        tape.backward()
        self.last_launches = tape.launches
        self.last = {'rgb_head': rgb_head, 'rgb_person': rgb_person, 'img_loss_head_neck': self.loss2[0], 'img_loss_com': self.loss2[1],
                     'points': 2 * P}
        return self.loss2.sum()

    def step(self, batch, global_step=0, noexp_iters=0, N_samples=None):
        """MAIN:909-931 + the learning-rate schedule of MAIN:1081-1094.  Returns the loss (device tensor, no sync)."""
        loss = self.losses_and_grads(batch, N_samples)
        lr = self.lrate
        if self.lrate_decay:
            lr = self.lrate * (self.decay_rate ** (global_step / (self.lrate_decay * 1000)))
        self.groups['dec'].step(lr, self.betas)
        self.groups['aud'].step(lr, self.betas)
        if global_step >= noexp_iters:
            self.groups['exp'].step(lr, self.betas)
        self.last_launches += 3
        return loss
