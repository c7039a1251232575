"""Debug timeline of CTA 0 of the Decoder programs on mlp_pp_kernel (dfn_debug_trace): per layer of the torso (or head)
program, the MMA issuer's waits / issue time and the epilogue's wait / work.  python profiles/trace_decoder.py [head|torso] [bf16|bf16x3]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dfa_nerf_b200 as dfn  # noqa: E402
import synth  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else 'torso'
mode = sys.argv[2] if len(sys.argv) > 2 else 'bf16'
R, S = 60000, 64
dev = torch.device('cuda', 0)
dec = dfn.Decoder(z_dim=256, hidden_size=256, dim_signal=96, use_deformation_field=True)
dec.load_state_dict(synth.decoder_state_dict(0))
dec = dec.to(dev)
fr = synth.frame_inputs(H=450, W=450, seed=0)
ro, rd = dfn.get_rays(450, 450, fr['focal'], fr['c2w'], fr['cx'], fr['cy'], device=dev)
ro, rd = [t.reshape(-1, 3)[:R].contiguous() for t in (ro, rd)]
z = dfn.z_vals_uniform(torch.full((R,), 0.4, device=dev), torch.full((R,), 1.0, device=dev), S)
g = torch.Generator().manual_seed(0)
zs, za = torch.randn(1, 256, generator=g).to(dev), torch.randn(1, 256, generator=g).to(dev)
sig = (torch.randn(1, 96, generator=g) if which == 'head' else torch.randn(1, 42, generator=g)).to(dev)
prec = {'bf16': dfn.PREC_BF16, 'bf16x3': dfn.PREC_BF16X3, 'fp16': dfn.PREC_FP16}[mode]
dec.query_rays(ro, rd, z, zs, za, sig, which, precision=prec)
T, NL = 5, (11 if which == "head" else 19) - (0 if mode == "bf16x3" else 1)   # single-pass programs: sigma_out folded
buf = torch.zeros(2 * T * NL * 8, dtype=torch.int64, device=dev)
dfn.lib.dfn_debug_trace(C.c_void_p(buf.data_ptr()), T)
dec.query_rays(ro, rd, z, zs, za, sig, which, precision=prec)
torch.cuda.synchronize()
dfn.lib.dfn_debug_trace(None, 0)
b = buf.cpu().reshape(2, T, NL, 2, 4)
t0 = int(b[0, 0, 0, 0, 0])
nslot = 1 if mode == 'bf16x3' else 2
j = 2
print('%s %s: layer slot | MMA start, wait_aready, issue (full-wait) | epilogue wait_acc, work' % (which, mode))
for l in range(NL):
    for s in range(nslot):
        w0, w1, e, fw = [int(x) for x in b[0, j, l, s]]
        e0, e1, e2, _ = [int(x) for x in b[1, j, l, s]]
        print('  l=%2d s=%d  start %8d  wait_aready %6d  issue %6d (full-wait %6d) | wait_acc %6d  work %6d'
              % (l, s, w0 - t0, w1 - w0, e - w1, fw & 0xFFFFFFFF, e1 - e0, e2 - e1))
nxt = int(b[0, j + 1, 0, 0, 0])
print('tile period (slot 0, l=0 to next l=0): %d cycles' % (nxt - int(b[0, j, 0, 0, 0])))
