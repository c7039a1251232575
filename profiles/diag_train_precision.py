"""Diagnostic (GPU): where does the training step's gradient error against the fp32 autograd oracle come from?  Runs one step with the
product kernels, then with dfn_gemm replaced by an fp64 torch matmul (tests/test_train_tape_cpu.py: ref_mm), then with the bias-gradient
and loss-backward kernels replaced too, and prints the largest relative gradient errors of each configuration."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import synth                      # noqa: E402
from oracle import train_oracle as TO         # noqa: E402
from test_train_oracle_golden import make_batch, NS, LRATE   # noqa: E402
import test_train_tape_cpu as T               # noqa: E402
import dfa_nerf_b200 as dfn                   # noqa: E402
from dfa_nerf_b200 import train               # noqa: E402

DEV = 'cuda'
PREC = {'fp32': dfn.PREC_FP32, 'bf16x3': dfn.PREC_BF16X3, 'bf16': dfn.PREC_BF16}[sys.argv[1] if len(sys.argv) > 1 else 'bf16x3']
b = make_batch()
sds = {'dec': synth.decoder_state_dict(6), 'aud': synth.mlp_encoder_state_dict(7), 'exp': synth.mlp_encoder_state_dict(8, (64, 32, 32))}
params = {k: {n: v.clone().requires_grad_(True) for n, v in sd.items()} for k, sd in sds.items()}
loss_ref, _, _ = TO.train_losses(params['dec'], params['aud'], params['exp'], b, NS)
loss_ref.backward()
bd = {k: (v.to(DEV) if torch.is_tensor(v) and k in ('target_com', 'target_head_neck', 'bc_img') else v) for k, v in b.items()}


def run(label, patches):
    saved = {k: getattr(train, k) for k in patches}
    for k, v in patches.items():
        setattr(train, k, v)
    try:
        dec = dfn.Decoder(z_dim=256, hidden_size=256, dim_signal=96, use_deformation_field=True)
        dec.load_state_dict(sds['dec'])
        aud, exp = dfn.AudioNet_W2L(), dfn.ExpressionEnc()
        aud.load_state_dict(sds['aud'])
        exp.load_state_dict(sds['exp'])
        tr = train.Trainer(dec, aud, exp, lrate=LRATE, N_samples=NS, precision=PREC)
        loss = tr.losses_and_grads(bd)
        torch.cuda.synchronize()
        errs = []
        for k in params:
            for n, q in params[k].items():
                if q.grad is None:
                    continue
                g = tr.grads[k][n].cpu().double()
                errs.append((((g - q.grad.double()).abs().max() / (q.grad.double().abs().max() + 1e-30)).item(), k + '/' + n))
        errs.sort(reverse=True)
        print('%-44s loss rel err %.1e | %s' % (label, abs(float(loss) - float(loss_ref)) / float(loss_ref),
                                                  ', '.join('%s %.1e' % (n, e) for e, n in errs[:8])), flush=True)
    finally:
        for k, v in saved.items():
            setattr(train, k, v)


run('product kernels (%s)' % (sys.argv[1] if len(sys.argv) > 1 else 'bf16x3'), {})
run('gemm -> fp64 torch', {'mm': T.ref_mm})
run('gemm, colsum -> fp64 torch', {'mm': T.ref_mm, 'colsum': T.ref_colsum})
run('gemm, colsum, loss_bwd -> torch', {'mm': T.ref_mm, 'colsum': T.ref_colsum, 'loss_bwd': T.ref_loss_bwd})
run('colsum -> fp64 torch only', {'colsum': T.ref_colsum})

# ---- per-call check: every dfn_gemm of one step against the fp64 product of ITS OWN inputs ----
real_mm = train.mm
worst = []


def checked_mm(A, B, Cm, **kw):
    ref = Cm.clone()
    T.ref_mm(A, B, ref, **{k: v for k, v in kw.items() if k != 'precision'})
    n = real_mm(A, B, Cm, **kw)
    e = ((Cm.double() - ref.double()).abs().max() / (ref.double().abs().max() + 1e-30)).item()
    worst.append((e, tuple(A.shape), tuple(A.stride()), tuple(B.shape), tuple(B.stride()), {k: (v if not torch.is_tensor(v) else 'T') for k, v in kw.items() if k != 'precision'},
                  float(ref.abs().max()), float(A.abs().max()), float(B.abs().max())))
    return n


run('per-call check', {'mm': checked_mm})
worst.sort(key=lambda t: -t[0])
print('%d gemm calls; the ten largest relative errors against fp64 on the same inputs:' % len(worst))
for w in worst[:10]:
    print('  %.2e  A %s%s  B %s%s  %s  |C|max %.2e |A|max %.2e |B|max %.2e' % w)


def is_backward(kw):
    return 'mask' in kw or kw.get('beta', 0) == 1


def fwd_exact(A, B, Cm, **kw):
    return real_mm(A, B, Cm, **kw) if is_backward(kw) else T.ref_mm(A, B, Cm, **{k: v for k, v in kw.items() if k != 'precision'})


def bwd_exact(A, B, Cm, **kw):
    return real_mm(A, B, Cm, **kw) if not is_backward(kw) else T.ref_mm(A, B, Cm, **{k: v for k, v in kw.items() if k != 'precision'})


run('forward GEMMs exact, backward kernel', {'mm': fwd_exact})
run('forward GEMMs kernel, backward exact', {'mm': bwd_exact})
for name in ('N1', 'vec'):
    def pick(A, B, Cm, _n=name, **kw):
        small = (B.shape[0] == 1) if _n == 'N1' else (A.shape[0] == 1 or A.shape[1] == 1)
        return T.ref_mm(A, B, Cm, **{k: v for k, v in kw.items() if k != 'precision'}) if small else real_mm(A, B, Cm, **kw)
    run('only the %s GEMMs exact' % ('N = 1 (density head)' if name == 'N1' else 'per-frame vector'), {'mm': pick})
