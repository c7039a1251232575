"""CPU: the reduced-precision restatements (oracle/quantized.py) are consistent
  * with the reference forward when no rounding is requested (dtype=None),
  * with each other: the state_dict restatement of FaceNeRF / NeRF equals the interpretation of the layer program the
    product's packer compiled (same weights rounded at the same place),
  * with expectations: bf16 operands cost ~1e-2 relative on sigma, fp16 ~8x less,
and the Decoder programs (plain and folded-head) interpreted without rounding give the reference Decoder.forward.
The GPU tests (test_gpu_2_mlp.py, test_gpu_5_decoder_tc.py) gate the tcgen05 kernels against these restatements."""
import numpy as np
import torch

from oracle import nerf_oracle as O
from oracle import quantized as Q
from oracle import synth

from program_dump import KB_PE, KB_DIR, dump_program, dump_model_program

TRUNK = ['pts_linears.%d' % i for i in range(8)]
FACE_NAMES = TRUNK + ['views_linears.%d' % i for i in range(3)] + ['feature_linear', 'alpha_linear', 'rgb_linear']
NERF_NAMES = TRUNK + ['views_linears.0', 'feature_linear', 'alpha_linear', 'rgb_linear']


def _inputs(P, seed=1):
    g = torch.Generator().manual_seed(seed)
    pts = (torch.rand(P, 3, generator=g) * 2 - 1)
    vd = torch.randn(P, 3, generator=g)
    vd = vd / torch.norm(vd, dim=-1, keepdim=True)
    aud = torch.randn(64, generator=g)
    return O.embed(pts, 10), aud, O.embed(vd, 4)


def model_program_q(prog, pe, aud, pev, dtype):
    """FaceNeRF / NeRF program of tc_pack_model with the per-call folds formed as fold_latent_kernel / view_bias_kernel do."""
    layers, weights, bias, fold_layer, fold_w, view_w, view_b = prog
    fold = {}
    if aud is not None:
        for i, l in enumerate(fold_layer):
            if l >= 0:
                fold[l] = (torch.as_tensor(fold_w[i]).double() @ aud.double()).float()
    vb = torch.as_tensor(view_b).float() + (pev.double() @ torch.as_tensor(view_w).double().t()).float()
    pe64 = torch.zeros(pe.shape[0], 64)
    pe64[:, :63] = pe
    col, sigma = Q.run_program_q(layers, weights, bias, {KB_PE: pe64}, dtype, fold_bias=fold, view_bias=vb, sigmoid=False)
    return torch.cat([col, sigma[:, None]], 1)


def test_no_rounding_is_the_reference_forward():
    pe, aud, pev = _inputs(41)
    with torch.no_grad():
        for sd, nerf in ((synth.facenerf_state_dict(2), False), (synth.nerf_state_dict(2), True)):
            if nerf:
                x = torch.cat([pe, pev], -1)
                ref = O.nerf_forward({k: v.double() for k, v in sd.items()}, x.double())
                out = Q.nerf_forward_q(sd, x, None)
            else:
                x = torch.cat([pe, aud[None].expand(41, -1), pev], -1)
                ref = O.facenerf_forward({k: v.double() for k, v in sd.items()}, x.double())
                out = Q.facenerf_forward_q(sd, x, None)
            scale = ref.abs().max(0).values
            assert ((out.double() - ref).abs() / scale).max() < 2e-6


def test_state_dict_restatement_equals_the_packed_program():
    from dfa_nerf_b200 import _lib
    pe, aud, pev = _inputs(53, seed=4)
    with torch.no_grad():
        for kind, sd, names, nerf in ((_lib.MODEL_FACENERF, synth.facenerf_state_dict(3), FACE_NAMES, False),
                                      (_lib.MODEL_NERF, synth.nerf_state_dict(3), NERF_NAMES, True)):
            prog = dump_model_program(kind, sd, names)
            x = torch.cat([pe, pev], -1) if nerf else torch.cat([pe, aud[None].expand(53, -1), pev], -1)
            for dtype in (torch.bfloat16, torch.float16):
                a = Q.facenerf_forward_q(sd, x, dtype, nerf=nerf)
                b = model_program_q(prog, pe, None if nerf else aud, pev, dtype)
                scale = a.abs().max(0).values
                # same operands rounded at the same places; only the fp32 bias folds are summed in another order (a rare
                # 1-ulp flip of an activation's 16-bit rounding is possible: 2^-9 of one activation, far below this bound)
                assert ((a - b).abs() / scale).max() < 2e-4, (kind, dtype)


def test_operand_precision_ordering():
    pe, aud, pev = _inputs(400, seed=6)
    sd = synth.facenerf_state_dict(0)
    x = torch.cat([pe, aud[None].expand(400, -1), pev], -1)
    with torch.no_grad():
        ref = O.facenerf_forward(sd, x)
        e = {}
        for name, dt in (('bf16', torch.bfloat16), ('fp16', torch.float16)):
            out = Q.facenerf_forward_q(sd, x, dt)
            e[name] = ((out[:, 3] - ref[:, 3]).abs().max().item(), (out[:, :3] - ref[:, :3]).abs().max().item())
        old = O.facenerf_forward_bf16(sd, x)         # round 1's restatement: same rounding points
    smax = ref[:, 3].abs().max().item()
    # same rounding points, another accumulation split: identical except where a pre-activation sat within an fp32 ulp of a
    # bf16 rounding boundary -- ONE such flip moves sigma by up to ~3e-3 |sigma|max (gain 1431 on a 2^-9 step of one activation);
    # that, not 1e-6, is the resolution any bf16 gate against this restatement can have
    d = (old - Q.facenerf_forward_q(sd, x, torch.bfloat16)).abs()
    assert (d[:, 3] == 0).float().mean() > 0.9 and d[:, 3].max() < 1e-2 * smax and d[:, :3].max() < 1e-4
    assert 1e-3 * smax < e['bf16'][0] < 5e-2 * smax
    assert e['fp16'][0] < 0.25 * e['bf16'][0] and e['fp16'][1] < 0.25 * e['bf16'][1]


def test_decoder_programs_without_rounding_are_the_reference():
    sd = synth.decoder_state_dict(3)
    g = torch.Generator().manual_seed(0)
    P = 64
    p = (torch.rand(1, P, 3, generator=g) * 2 - 1) * 0.7
    rd = torch.randn(1, P, 3, generator=g)
    zs, za = torch.randn(1, 256, generator=g), torch.randn(1, 256, generator=g)
    sig = {0: torch.randn(1, 96, generator=g), 1: torch.randn(1, 42, generator=g)}
    with torch.no_grad():
        pe = torch.zeros(P, 64)
        pe[:, :60] = O.decoder_transform_points(p, 10)[0]
        ped = torch.zeros(P, 64)
        ped[:, :24] = O.decoder_transform_points(rd / torch.norm(rd, dim=-1, keepdim=True), 4)[0]
        for field, which in ((0, 'head'), (1, 'torso')):
            rf, rs = O.decoder_forward(sd, p, rd, zs, za, sig[field], which)
            # 0 / 1: the plain / folded-head program (mlp_pp.cu); 2 / 3: the same in the layout the CTA-pair kernel runs (mlp_pair.cu: the
            # torso's fc_in_torso as one layer over [PE' | signal' in hidden block 3])
            for folded in (0, 1, 2, 3):
                layers, weights, bias, folds, dimL, view_layer, dot_w = dump_program(sd, field, folded)
                if folded >= 2 and field == 1:
                    assert len(layers) == len(dump_program(sd, field, folded - 2)[0]) - 1
                latent = torch.cat([sig[field].reshape(-1), zs.reshape(-1), za.reshape(-1)])
                fold = {l: (torch.as_tensor(fw).double().t() @ latent.double()).float() for l, fw in folds.items()}
                feat, sigma = Q.run_program_q(layers, weights, bias, {KB_PE: pe, KB_DIR: ped}, None, fold_bias=fold, dot_w=dot_w)
                assert (feat - rf[0]).abs().max() < 5e-6
                assert (sigma - rs[0]).abs().max() < 5e-6 * rs.abs().max()
                # and with 16-bit operands: the expected size of the operand-rounding error
                f16, s16 = Q.run_program_q(layers, weights, bias, {KB_PE: pe, KB_DIR: ped}, torch.bfloat16, fold_bias=fold, dot_w=dot_w)
                assert 1e-5 < (f16 - rf[0]).abs().max() < 2e-2
                assert (s16 - rs[0]).abs().max() < 5e-2 * rs.abs().max()


def test_stats():
    mx, p99, med = Q.stats(np.arange(101.))
    assert mx == 100. and abs(p99 - 99.) < 1e-9 and med == 50.
