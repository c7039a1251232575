"""GPU parity of the skip-MLP: fp32 FFMA module forward, then the fused tcgen05 query path
(bf16x3 against the fp32 oracle at 1e-4 on what compositing consumes; bf16 against the
bf16-operand restatement of the oracle)."""
import pytest
import torch

from oracle import nerf_oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def maxerr(a, b):
    return (a.detach().cpu().double() - b.detach().cpu().double()).abs().max().item()


@pytest.fixture(scope='module')
def dfn():
    import dfa_nerf_b200
    return dfa_nerf_b200


def face(dfn, seed):
    m = dfn.FaceNeRF(D=8, W=256, input_ch=63, input_ch_views=27, dim_aud=64, output_ch=4, skips=[4], use_viewdirs=True)
    m.load_state_dict(synth.facenerf_state_dict(seed))
    return m.to(DEV)


def nerf(dfn, seed):
    m = dfn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=4, skips=[4], use_viewdirs=True)
    m.load_state_dict(synth.nerf_state_dict(seed))
    return m.to(DEV)


def test_module_forward_fp32_golden(dfn, golden):
    g = golden('mlp')
    y = face(dfn, g['face_seed'])(g['x_face'].to(DEV))
    assert y.shape == (48, 4)
    # sigma has magnitude ~40 (gain 1431): 2e-4 absolute is ~5e-6 relative
    assert maxerr(y[:, :3], g['y_face'][:, :3]) < 2e-5
    assert maxerr(y[:, 3], g['y_face'][:, 3]) < 5e-4
    yn = nerf(dfn, g['nerf_seed'])(g['x_nerf'].to(DEV))
    assert maxerr(yn[:, :3], g['y_nerf'][:, :3]) < 2e-5
    assert maxerr(yn[:, 3], g['y_nerf'][:, 3]) < 5e-4


def test_module_forward_fp32_ragged(dfn):
    sd = synth.facenerf_state_dict(1)
    m = face(dfn, 1)
    for P in (1, 63, 1000):
        x = torch.randn(P, 154)
        with torch.no_grad():
            ref = O.facenerf_forward(sd, x)
        y = m(x.to(DEV))
        assert maxerr(y[:, :3], ref[:, :3]) < 5e-5 and maxerr(y[:, 3], ref[:, 3]) < 2e-3
    assert m(torch.zeros(0, 154, device=DEV)).shape == (0, 4)
    assert m(torch.zeros(2, 5, 154, device=DEV)).shape == (2, 5, 4)


def _query_case(R, S, seed=11):
    fr = synth.frame_inputs(H=R, W=1, seed=seed)
    g = torch.Generator().manual_seed(seed)
    ro = fr['c2w'][:3, -1].expand(R, 3).contiguous()
    rd = torch.randn(R, 3, generator=g) * 0.2 + torch.tensor([0., 0., -1.])
    vd = rd / torch.norm(rd, dim=-1, keepdim=True)
    z, _ = torch.sort(torch.rand(R, S, generator=g) * 0.6 + 0.4, -1)
    return ro, rd, vd, z, fr['aud']


def _alpha_err(raw, ref, z, rd):
    """Error of what compositing consumes: weights and sigmoid colours."""
    w = O.calc_volume_weights(z, rd, raw[..., 3].cpu())
    wr = O.calc_volume_weights(z, rd, ref[..., 3])
    return maxerr(w, wr), maxerr(torch.sigmoid(raw[..., :3]), torch.sigmoid(ref[..., :3]))


@pytest.mark.parametrize('R,S', [(2, 64), (37, 64), (21, 192), (300, 64)])
def test_query_points_fp32(dfn, R, S):
    ro, rd, vd, z, aud = _query_case(R, S)
    sd = synth.facenerf_state_dict(0)
    eng = dfn.RenderEngine(face(dfn, 0), None, S, 0, precision=dfn.PREC_FP32)
    raw = eng.query_points(eng.network_fn, ro.to(DEV), rd.to(DEV), vd.to(DEV), z.to(DEV), aud.to(DEV))
    with torch.no_grad():
        ref = O.run_network(sd, 'facenerf', ro[:, None] + rd[:, None] * z[:, :, None], vd, aud)
    ew, ec = _alpha_err(raw, ref, z, rd)
    assert ew < 1e-5 and ec < 1e-5, (ew, ec)


@pytest.mark.parametrize('prec_name', ['PREC_BF16X3', 'PREC_FP16X3M'])
@pytest.mark.parametrize('R,S', [(2, 64), (37, 64), (21, 192), (600, 64)])
def test_query_points_parity_modes_vs_fp32_oracle(dfn, R, S, prec_name):
    """The two fp32-parity modes against the fp32 oracle at the north star's 1e-4: split bf16 on every layer, and fp16 operands with the
    three split products only on the layers that form the density after the skip connection (DFN_PREC_FP16X3M)."""
    ro, rd, vd, z, aud = _query_case(R, S)
    sd = synth.facenerf_state_dict(0)
    eng = dfn.RenderEngine(face(dfn, 0), None, S, 0, precision=getattr(dfn, prec_name))
    raw = eng.query_points(eng.network_fn, ro.to(DEV), rd.to(DEV), vd.to(DEV), z.to(DEV), aud.to(DEV))
    with torch.no_grad():
        ref = O.run_network(sd, 'facenerf', ro[:, None] + rd[:, None] * z[:, :, None], vd, aud)
    assert torch.isfinite(raw).all()
    ew, ec = _alpha_err(raw, ref, z, rd)
    print('%s R=%d S=%d: weights err %.2e, colour err %.2e, raw sigma err %.2e' % (prec_name, R, S, ew, ec, maxerr(raw[..., 3], ref[..., 3])))
    assert ew < 1e-4 and ec < 1e-4, (ew, ec)      # north-star tolerance: 1e-4 max-abs


def test_query_points_mixed_mode_at_scale_and_nerf(dfn):
    """DFN_PREC_FP16X3M on 20,000 rays x 192 samples against the FFMA kernels (compositing weights and colours <= 1e-4), and on the NeRF
    class (no latent, feature_linear composed into the view layer)."""
    R, S = 20000, 192
    ro, rd, vd, z, aud = _query_case(R, S, seed=11)
    net = face(dfn, 1)
    args = [t.to(DEV) for t in (ro, rd, vd, z, aud)]
    a = dfn.RenderEngine(net, None, S, 0, precision=dfn.PREC_FP32).query_points(net, *args)
    b = dfn.RenderEngine(net, None, S, 0, precision=dfn.PREC_FP16X3M).query_points(net, *args)
    wa = dfn.calc_volume_weights(args[3], args[1], a[..., 3].contiguous())
    wb = dfn.calc_volume_weights(args[3], args[1], b[..., 3].contiguous())
    print('mixed mode, 20,000 x 192: weights %.2e colours %.2e' % (maxerr(wa, wb), maxerr(torch.sigmoid(a[..., :3]), torch.sigmoid(b[..., :3]))))
    assert maxerr(wa, wb) < 1e-4 and maxerr(torch.sigmoid(a[..., :3]), torch.sigmoid(b[..., :3])) < 1e-4
    nn_ = dfn.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=4, skips=[4], use_viewdirs=True)
    nn_.load_state_dict(synth.nerf_state_dict(0))
    nn_ = nn_.to(DEV)
    ro, rd, vd, z, _ = _query_case(300, 64, seed=5)
    args = [t.to(DEV) for t in (ro, rd, vd, z)]
    a = dfn.RenderEngine(nn_, None, 64, 0, precision=dfn.PREC_FP32).query_points(nn_, *args)
    b = dfn.RenderEngine(nn_, None, 64, 0, precision=dfn.PREC_FP16X3M).query_points(nn_, *args)
    wa = dfn.calc_volume_weights(args[3], args[1], a[..., 3].contiguous())
    wb = dfn.calc_volume_weights(args[3], args[1], b[..., 3].contiguous())
    assert maxerr(wa, wb) < 1e-4 and maxerr(torch.sigmoid(a[..., :3]), torch.sigmoid(b[..., :3])) < 1e-4


# Gates of the single-pass kernels against the reduced-precision restatement (oracle/quantized.py: same operands rounded
# at the same places, exact products, one fp32 rounding per accumulator).  What is left between the two is the tensor
# cores' own accumulation (order, truncated alignment: ~1e-6 relative on a pre-activation) and the MUFU sin/cos of the
# fused encoding -- either can flip the 16-bit rounding of an activation that sits on a rounding boundary, and ONE flip
# in the last trunk layer moves sigma by gain x |w| x ulp ~ 3e-3 |sigma|max in bf16 (tests/test_quantized_cpu.py).  So the
# max is gated at a few flips, p99 and the median far below, and everything is printed.  Measured on the B200 (round 2):
# bf16 colours <= 6.5e-5, sigma max <= 7.5e-3, p99 <= 1.2e-3, median ~1e-7 of |sigma|max; fp16 9e-6 / 1.1e-3 / 4.6e-4 / 4e-7 --
# i.e. at the median the restatement reproduces the kernel to fp32 rounding, 10^4 times closer than the fp32 forward is.
#                 colours(sigmoid) max, sigma max / p99 / median as fractions of |sigma|max
Q_GATE = {'bf16': (1.3e-4, 1.5e-2, 2.5e-3, 1e-6), 'fp16': (2e-5, 2.5e-3, 1e-3, 2e-6)}


def _gate_quantized(tag, prec_name, raw, refq, ref32):
    from oracle import quantized as Q
    raw = raw.detach().cpu()
    smax = ref32[..., 3].abs().max().item()
    ec = maxerr(torch.sigmoid(raw[..., :3]), torch.sigmoid(refq[..., :3]))
    mx, p99, med = Q.stats(raw[..., 3] - refq[..., 3])
    mx32, p9932, med32 = Q.stats(raw[..., 3] - ref32[..., 3])
    print('%s %s vs %s-operand restatement: colours %.2e | sigma max %.2e p99 %.2e median %.2e of |sigma|max=%.1f  '
          '(vs fp32 oracle: max %.2e p99 %.2e median %.2e)'
          % (tag, prec_name, prec_name, ec, mx / smax, p99 / smax, med / smax, smax, mx32 / smax, p9932 / smax, med32 / smax))
    gc, gm, g99, gmed = Q_GATE[prec_name]
    assert torch.isfinite(raw).all()
    assert ec <= gc and mx <= gm * smax and p99 <= g99 * smax and med <= gmed * smax, (tag, ec, mx / smax, p99 / smax, med / smax)
    # the restatement explains the kernel: far closer to it than to the fp32 forward
    assert med < 0.02 * med32 and p99 < p9932, (tag, med, med32, p99, p9932)


def _face_x(ro, rd, vd, z, aud):
    R, S = z.shape
    pts = (ro[:, None] + rd[:, None] * z[:, :, None]).reshape(-1, 3)
    return torch.cat([O.embed(pts, 10), aud[None].expand(pts.shape[0], -1), O.embed(vd[:, None].expand(R, S, 3).reshape(-1, 3), 4)], -1)


@pytest.mark.parametrize('prec_name', ['bf16', 'fp16'])
@pytest.mark.parametrize('R,S', [(37, 64), (600, 64), (20000, 8)])
def test_query_points_single_pass_vs_quantized_oracle(dfn, prec_name, R, S):
    """mlp_tc_kernel<bf16> / <fp16> (the benchmarked kernels) against the bf16- / fp16-operand restatement of HELP:275-299."""
    from oracle import quantized as Q
    ro, rd, vd, z, aud = _query_case(R, S)
    sd = synth.facenerf_state_dict(0)
    prec, dt = {'bf16': (dfn.PREC_BF16, torch.bfloat16), 'fp16': (dfn.PREC_FP16, torch.float16)}[prec_name]
    eng = dfn.RenderEngine(face(dfn, 0), None, S, 0, precision=prec)
    raw = eng.query_points(eng.network_fn, ro.to(DEV), rd.to(DEV), vd.to(DEV), z.to(DEV), aud.to(DEV))
    x = _face_x(ro, rd, vd, z, aud)
    with torch.no_grad():
        refq = Q.facenerf_forward_q(sd, x, dt).reshape(R, S, 4)
        ref32 = O.facenerf_forward(sd, x).reshape(R, S, 4)
    _gate_quantized('FaceNeRF R=%d S=%d' % (R, S), prec_name, raw, refq, ref32)


@pytest.mark.parametrize('prec_name', ['bf16', 'fp16'])
def test_query_points_nerf_single_pass_vs_quantized_oracle(dfn, prec_name):
    """NeRF (HELP:372-396; feature_linear composed into views_linears.0 before rounding) on the same kernels."""
    from oracle import quantized as Q
    R, S = 300, 64
    ro, rd, vd, z, aud = _query_case(R, S, seed=5)
    sd = synth.nerf_state_dict(2)
    prec, dt = {'bf16': (dfn.PREC_BF16, torch.bfloat16), 'fp16': (dfn.PREC_FP16, torch.float16)}[prec_name]
    net = nerf(dfn, 2)
    eng = dfn.RenderEngine(net, None, S, 0, precision=prec)
    raw = eng.query_points(net, ro.to(DEV), rd.to(DEV), vd.to(DEV), z.to(DEV), None)
    pts = (ro[:, None] + rd[:, None] * z[:, :, None]).reshape(-1, 3)
    x = torch.cat([O.embed(pts, 10), O.embed(vd[:, None].expand(R, S, 3).reshape(-1, 3), 4)], -1)
    with torch.no_grad():
        refq = Q.nerf_forward_q(sd, x, dt).reshape(R, S, 4)
        ref32 = O.nerf_forward(sd, x).reshape(R, S, 4)
    _gate_quantized('NeRF R=%d S=%d' % (R, S), prec_name, raw, refq, ref32)


def test_query_points_pp_kernel_bf16_vs_quantized_oracle(dfn):
    """The ping-pong kernel (mlp_pp_kernel<bf16>, dfn_debug_set_impl(2)) in its single-pass instantiation: same gate."""
    from oracle import quantized as Q
    R, S = 600, 64
    ro, rd, vd, z, aud = _query_case(R, S)
    sd = synth.facenerf_state_dict(0)
    eng = dfn.RenderEngine(face(dfn, 0), None, S, 0, precision=dfn.PREC_BF16)
    try:
        dfn.lib.dfn_debug_set_impl(2)
        raw = eng.query_points(eng.network_fn, ro.to(DEV), rd.to(DEV), vd.to(DEV), z.to(DEV), aud.to(DEV))
    finally:
        dfn.lib.dfn_debug_set_impl(-1)
    x = _face_x(ro, rd, vd, z, aud)
    with torch.no_grad():
        refq = Q.facenerf_forward_q(sd, x, torch.bfloat16).reshape(R, S, 4)
        ref32 = O.facenerf_forward(sd, x).reshape(R, S, 4)
    _gate_quantized('FaceNeRF (mlp_pp) R=%d S=%d' % (R, S), 'bf16', raw, refq, ref32)


@pytest.mark.parametrize('R,S', [(37, 64), (300, 64), (21, 192)])
def test_query_points_fp16_between_bf16_and_fp32(dfn, R, S):
    """DFN_PREC_FP16: the bf16 kernel and schedule with fp16 operands (11-bit significands).  Same speed, several times
    closer to the fp32 oracle than bf16 on everything compositing consumes; reported, and gated on that ordering."""
    ro, rd, vd, z, aud = _query_case(R, S)
    sd = synth.facenerf_state_dict(0)
    net = face(dfn, 0)
    with torch.no_grad():
        ref = O.run_network(sd, 'facenerf', ro[:, None] + rd[:, None] * z[:, :, None], vd, aud)
    err = {}
    for name, prec in (('bf16', dfn.PREC_BF16), ('fp16', dfn.PREC_FP16)):
        eng = dfn.RenderEngine(net, None, S, 0, precision=prec)
        raw = eng.query_points(net, ro.to(DEV), rd.to(DEV), vd.to(DEV), z.to(DEV), aud.to(DEV))
        assert torch.isfinite(raw).all()
        err[name] = _alpha_err(raw, ref, z, rd) + (maxerr(raw[..., 3], ref[..., 3]),)
    print('R=%d S=%d  weights / colour / raw sigma error: bf16 %.2e %.2e %.2e | fp16 %.2e %.2e %.2e' % ((R, S) + err['bf16'] + err['fp16']))
    assert err['fp16'][0] < 0.5 * err['bf16'][0] and err['fp16'][2] < 0.5 * err['bf16'][2]
    assert err['fp16'][0] < 5e-3 and err['fp16'][1] < 5e-5


def test_query_points_nerf_model(dfn):
    R, S = 19, 64
    ro, rd, vd, z, aud = _query_case(R, S, seed=5)
    sd = synth.nerf_state_dict(2)
    net = nerf(dfn, 2)
    with torch.no_grad():
        ref = O.run_network(sd, 'nerf', ro[:, None] + rd[:, None] * z[:, :, None], vd, aud)
    for prec, tol in ((dfn.PREC_FP32, 1e-5), (dfn.PREC_BF16X3, 1e-4)):
        eng = dfn.RenderEngine(net, None, S, 0, precision=prec)
        raw = eng.query_points(net, ro.to(DEV), rd.to(DEV), vd.to(DEV), z.to(DEV), None)
        ew, ec = _alpha_err(raw, ref, z, rd)
        assert ew < tol and ec < tol, (prec, ew, ec)


def test_tc_matches_fp32_kernel_at_scale(dfn):
    """Full-size cross-check on the device: 20k rays x 192 samples, bf16x3 vs the fp32 FFMA kernels."""
    R, S = 20000, 192
    ro, rd, vd, z, aud = _query_case(R, S, seed=3)
    net = face(dfn, 1)
    args = [t.to(DEV) for t in (ro, rd, vd, z, aud)]
    e32 = dfn.RenderEngine(net, None, S, 0, precision=dfn.PREC_FP32)
    ex3 = dfn.RenderEngine(net, None, S, 0, precision=dfn.PREC_BF16X3)
    a = e32.query_points(net, *args)
    b = ex3.query_points(net, *args)
    wa = dfn.calc_volume_weights(args[3], args[1], a[..., 3].contiguous())
    wb = dfn.calc_volume_weights(args[3], args[1], b[..., 3].contiguous())
    assert maxerr(wa, wb) < 1e-4
    assert maxerr(torch.sigmoid(a[..., :3]), torch.sigmoid(b[..., :3])) < 1e-4


def test_kernel_variants_agree(dfn):
    """The tcgen05 kernel generations (dfn_debug_set_impl: 1 = mlp_tc.cu, one CTA per tile pair of slots; 2 = mlp_pp.cu, the layer-program
    interpreter; 3 = mlp_pair.cu, cta_group::2 CTA pairs -- the default of the single-pass precisions; 8 = the same with four epilogue warps
    per slot; 3 + 16 (f + 1): pair kernel with flag set f, 0 = every slot streams its own weights and cluster-scope releases) compute the
    same function: bit-identical in bf16 and fp16 (same operands, same fp32 bias add and rounding); in bf16x3 the hi/lo products are
    accumulated in a different order, so the two generations agree to the operand precision.  Sizes: a ragged tail, fewer tiles than
    SMs, an odd tile count, and several tiles per CTA."""
    net = face(dfn, 1)
    try:
        for R, S in ((700, 192), (1, 64), (37, 64), (5000, 64)):
            ro, rd, vd, z, aud = _query_case(R, S, seed=9)
            args = [t.to(DEV) for t in (ro, rd, vd, z, aud)]
            for prec in (dfn.PREC_BF16, dfn.PREC_FP16, dfn.PREC_BF16X3):
                eng = dfn.RenderEngine(net, None, S, 0, precision=prec)
                outs = {}
                impls = (1, 2) if prec == dfn.PREC_BF16X3 else ((1, 2, 3, 8, 3 + 16, 3 + 32, 3 + 48) if prec == dfn.PREC_BF16 else (1, 3, 8))
                for impl in impls:
                    dfn.lib.dfn_debug_set_impl(impl)
                    outs[impl] = eng.query_points(net, *args).clone()
                    assert torch.isfinite(outs[impl]).all()
                dfn.lib.dfn_debug_set_impl(-1)
                outs[-1] = eng.query_points(net, *args).clone()
                if prec != dfn.PREC_BF16X3:
                    for impl in outs:
                        assert torch.equal(outs[impl], outs[1]), (R, S, prec, impl)
                else:
                    for impl in outs:
                        assert maxerr(outs[impl], outs[1]) < 2e-3, impl      # x3: K-half order differs
                    wa = dfn.calc_volume_weights(args[3], args[1], outs[1][..., 3].contiguous())
                    wb = dfn.calc_volume_weights(args[3], args[1], outs[2][..., 3].contiguous())
                    assert maxerr(wa, wb) < 1e-4
    finally:
        dfn.lib.dfn_debug_set_impl(-1)


def test_pair_kernel_stress_bit_identical(dfn):
    """The CTA-pair kernel against the 1-CTA generation on a full 450 x 450 x 192 query (405,000 tiles, 4.9 M tile-layers), bit for bit:
    the peer CTA's `aready` arrivals carry a CTA-scope release (mlp_pair.cu), so a mis-ordered handshake would show up here as a stale
    activation block."""
    R, S = 202500, 192
    net = face(dfn, 2)
    fr = synth.frame_inputs(H=450, W=450, seed=3)
    ro, rd, vd = dfn.get_rays(450, 450, fr['focal'], fr['c2w'], fr['cx'], fr['cy'], device=DEV, return_viewdirs=True)
    ro, rd, vd = [t.reshape(-1, 3).contiguous() for t in (ro, rd, vd)]
    z, _ = torch.sort(torch.rand(R, S, device=DEV) * 0.6 + 0.4, -1)
    aud = fr['aud'].to(DEV)
    eng = dfn.RenderEngine(net, None, S, 0, precision=dfn.PREC_BF16)
    try:
        dfn.lib.dfn_debug_set_impl(1)
        ref = eng.query_points(net, ro, rd, vd, z, aud).clone()
        dfn.lib.dfn_debug_set_impl(-1)
        for rep in range(3):
            out = eng.query_points(net, ro, rd, vd, z, aud)
            assert torch.equal(out, ref), rep
    finally:
        dfn.lib.dfn_debug_set_impl(-1)


def test_split_schedule_switches(dfn):
    """dfn_debug_set_pp_flags: early staging of the skip layer's input and the weight-barrier order (bits 0, 1) change WHEN things
    happen, never what is computed -- bit-identical to the round-1 schedule (flags 0) in both parity modes, ragged and multi-tile sizes;
    bit 2 (fp16x3m: alpha_linear in fp32 inside the last trunk layer's epilogue, views_linears.0 single-pass) changes the arithmetic and is
    gated against the fp32 oracle like the mode itself; bit 3 (the view layer's per-ray bias rows staged in shared memory) does not."""
    net = face(dfn, 1)
    sd = synth.facenerf_state_dict(1)
    try:
        for R, S in ((37, 64), (700, 192), (5000, 64)):
            ro, rd, vd, z, aud = _query_case(R, S, seed=13)
            args = [t.to(DEV) for t in (ro, rd, vd, z, aud)]
            for prec in (dfn.PREC_BF16X3, dfn.PREC_FP16X3M):
                eng = dfn.RenderEngine(net, None, S, 0, precision=prec)
                outs = {}
                for f in (0, 1, 2, 3, 7, 8, 15):
                    dfn.lib.dfn_debug_set_pp_flags(f)
                    outs[f] = eng.query_points(net, *args).clone()
                for f in (1, 2, 3, 8):                            # bit 3: the view layer's bias rows from shared memory, same arithmetic
                    assert torch.equal(outs[f], outs[0]), (R, S, prec, f)
                assert torch.equal(outs[15], outs[7]), (R, S, prec)
                if prec == dfn.PREC_BF16X3:
                    assert torch.equal(outs[7], outs[0])          # bit 2 is an fp16x3m switch
                elif R <= 700:
                    with torch.no_grad():
                        ref = O.run_network(sd, 'facenerf', ro[:, None] + rd[:, None] * z[:, :, None], vd, aud)
                    for f in (3, 7):
                        ew, ec = _alpha_err(outs[f], ref, z, rd)
                        print('fp16x3m flags %d R=%d: weights %.2e colours %.2e sigma %.2e' % (f, R, ew, ec, maxerr(outs[f][..., 3], ref[..., 3])))
                        assert ew < 1e-4 and ec < 1e-4, (f, ew, ec)
    finally:
        dfn.lib.dfn_debug_set_pp_flags(15)
