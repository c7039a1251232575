"""Reduced-precision restatements of the reference's radiance networks.  TEST INFRASTRUCTURE ONLY.

What a single-pass bf16 / fp16 tensor-core evaluation of HELP:275-299 (FaceNeRF.forward), HELP:372-396 (NeRF.forward) and
DEC:277-349 / DEC:109-134 (Decoder.forward + DeformationField_ori) computes when the roundings sit where the CUDA kernels
put them (csrc/mlp_tc.cu, csrc/mlp_pp.cu):

  * every GEMM operand is rounded ONCE to the 16-bit type, round-to-nearest-even: weights when the model is packed,
    activations after bias + ReLU (fp16: saturating), the positional encoding when it is staged;
  * products are exact and accumulated in fp32 (emulated here with fp64 sums rounded to fp32 once: the difference is
    accumulation-order noise, ~1e-7 relative);
  * biases are fp32 and are added to the fp32 accumulator; the per-frame latent columns (HELP:276) and per-ray
    view-direction columns (HELP:288) never pass through the tensor cores -- they are fp32 bias terms;
  * sigma is read from the fp32 accumulator (no rounding of the output), the colours likewise.

`facenerf_forward_q` / `nerf_forward_q` are written from the reference's state_dict (independent of the product's
packer); `run_program_q` interprets a layer program as dumped by dfn_*_program_host (host-only) with the same rounding
points -- tests/test_quantized_cpu.py checks the two against each other and (dtype=None) against the fp64 reference
forward, then the GPU tests gate the kernels against them.
"""
import numpy as np
import torch
import torch.nn.functional as F

EPI_RELU, EPI_VIEW0, EPI_RGB, EPI_SIGMA, EPI_CONT, EPI_STAGE = range(6)
KB_PE, KB_IN1, KB_DIR = 4, 5, 6
F_ACCUM, F_DOT_SIGMA = 1, 2


def rnd(x, dtype):
    """fp32 tensor -> nearest `dtype` value (ties to even; fp16 saturates like cvt.rn.satfinite) -> fp32.  None: identity."""
    if dtype is None:
        return x
    x = x.float()
    if dtype == torch.float16:
        x = x.clamp(-65504., 65504.)
    return x.to(dtype).float()


def _mm(a, w):
    """a [P,K], w [N,K] (both already rounded): exact products, fp64 sums -> fp32."""
    return (a.double() @ w.double().t()).float()


def _fold(w, v):
    """Per-frame / per-ray fp32 bias term sum_j w[n][j] v[..., j], accumulated in fp32 like fold_latent_kernel /
    view_bias_kernel (sequential fmaf; torch's fp32 matmul differs by summation order only)."""
    return (v.double() @ w.double().t()).float()


def facenerf_forward_q(sd, x, dtype, input_ch=63, input_ch_views=27, dim_aud=64, D=8, skips=(4,), nerf=False):
    """HELP:275-299 (nerf=False) / HELP:372-396 (nerf=True) with 16-bit tensor-core operands.  x [P, input_ch + dim_aud +
    input_ch_views] fp32 -> [P,4] = (rgb pre-sigmoid, sigma)."""
    if nerf:
        dim_aud = 0
    pe, aud, views = torch.split(x.float(), [input_ch, dim_aud, input_ch_views], dim=-1)
    n_in = input_ch + dim_aud
    pe_q = rnd(pe, dtype)

    def W(name):
        return sd[name + '.weight'].float(), sd[name + '.bias'].float()

    h = None
    for i in range(D):
        w, b = W('pts_linears.%d' % i)
        has_in = i == 0 or (i - 1) in skips
        parts_a, parts_w = [], []          # one accumulator per layer: [PE | h] against [W_pe | W_h]
        if has_in:
            parts_a.append(pe_q)
            parts_w.append(rnd(w[:, :input_ch], dtype))
            bias = b + (_fold(w[:, input_ch:n_in], aud[:1])[0] if dim_aud else 0.)     # the latent is the same for every point
        else:
            bias = b
        if i > 0:
            parts_a.append(h)
            parts_w.append(rnd(w[:, n_in if has_in else 0:], dtype))
        acc = _mm(torch.cat(parts_a, 1), torch.cat(parts_w, 1))
        h = rnd(F.relu(acc + bias), dtype)
    wa, ba = W('alpha_linear')
    sigma = _mm(h, rnd(wa, dtype)) + ba
    wv, bv = W('views_linears.0')
    Wd = h.shape[1]
    if nerf:
        # feature_linear has no activation (HELP:384): composed into views_linears.0 in fp64, rounded once (tc_pack_model)
        wf, bf = W('feature_linear')
        wc = (wv[:, :Wd].double() @ wf.double()).float()
        bc = (bv.double() + wv[:, :Wd].double() @ bf.double()).float()
        n_view = 1
    else:
        wc, bc = wv[:, :Wd], bv
        n_view = 1 + D // 4
    h = rnd(F.relu(_mm(h, rnd(wc, dtype)) + (bc + _fold(wv[:, Wd:], views))), dtype)
    for i in range(1, n_view):
        w, b = W('views_linears.%d' % i)
        h = rnd(F.relu(_mm(h, rnd(w, dtype)) + b), dtype)
    wr, br = W('rgb_linear')
    rgb = _mm(h, rnd(wr, dtype)) + br
    return torch.cat([rgb, sigma], -1)


def nerf_forward_q(sd, x, dtype, **kw):
    return facenerf_forward_q(sd, x, dtype, nerf=True, **kw)


# ------------------------------------------------------------------------------------------ layer programs


def run_program_q(layers, weights, bias, blocks_in, dtype, fold_bias=None, dot_w=None, view_bias=None, sigmoid=True):
    """Interprets a tcgen05 layer program (csrc/model.h TcLayer semantics) with the kernels' rounding points.

    layers     list of objects with n, nkb, kb[], epi, flags            (dfn_*_program_host)
    weights    [n_layers][256][6][64] fp32 dense weights per (output row, input-block slot, position)
    bias       [n_layers][256] fp32, per-frame latents ALREADY folded in (or pass fold_bias {layer: [256] extra})
    blocks_in  {block id: [P,64] fp32}: staged input blocks (KB_PE, KB_DIR, ...), rounded here as the PE warps do
    view_bias  [P, view_w] per-ray bias of the EPI_VIEW0 layer (FaceNeRF / NeRF programs) or None
    dot_w      [260]: folded density head (F_DOT_SIGMA): row [256] + bias at [256]
    Returns (colours [P,3] -- sigmoid applied when `sigmoid` (Decoder programs) --, sigma [P])."""
    P = next(iter(blocks_in.values())).shape[0]
    blocks = {k: torch.zeros(P, 64) for k in range(7)}
    for k, v in blocks_in.items():
        blocks[k] = rnd(torch.as_tensor(v).float(), dtype)
    weights = torch.as_tensor(weights)
    bias = torch.as_tensor(bias).float()
    acc, sigma, col = None, None, None
    for l, L in enumerate(layers):
        n = int(L.n)
        kbs = [int(L.kb[i]) for i in range(L.nkb)]
        x = torch.cat([blocks[k] for k in kbs], 1)
        Wl = rnd(weights[l, :n, :L.nkb].reshape(n, -1).float(), dtype)
        out = _mm(x, Wl)
        acc = acc + out if (L.flags & F_ACCUM) else out
        b = bias[l, :n]
        if fold_bias is not None and l in fold_bias:
            b = b + torch.as_tensor(fold_bias[l]).float()[:n]
        if L.epi == EPI_CONT:
            continue
        if L.epi == EPI_SIGMA:
            sigma = acc[:, 0] + b[0]
            continue
        if L.epi == EPI_STAGE:
            o = rnd(acc + b, dtype)
            blocks[KB_PE], blocks[KB_IN1] = o[:, :64].clone(), o[:, 64:128].clone()
            blocks[3] = o[:, 64:128].clone()      # mlp_pair.cu also leaves signal' in hidden block 3 (read by the one-layer fc_in_torso)
            continue
        if L.epi == EPI_RGB:
            col = acc[:, :3] + b[:3]
            if sigmoid:
                col = torch.sigmoid(col)
            continue
        if L.epi == EPI_VIEW0:
            wh = view_bias.shape[1]
            sigma = acc[:, wh] + b[wh]
            hf = F.relu(acc[:, :wh] + torch.as_tensor(view_bias).float())
            n = wh
        else:
            hf = F.relu(acc + b)
            if L.flags & F_DOT_SIGMA:
                # density from this layer's fp32 activations on the CUDA cores; the head row is held as 16-bit pairs
                dw = torch.as_tensor(dot_w).float()
                sigma = (hf.double() @ rnd(dw[:256], dtype).double()).float() + dw[256]
        h = rnd(hf, dtype)
        for i in range(n // 64):
            blocks[i] = h[:, 64 * i:64 * (i + 1)]
    return col, sigma


def stats(err):
    """(max, p99, median) of an absolute-error tensor."""
    e = np.abs(np.asarray(torch.as_tensor(err).detach().cpu().double().reshape(-1)))
    if e.size == 0:
        return 0., 0., 0.
    return float(e.max()), float(np.quantile(e, 0.99)), float(np.median(e))
