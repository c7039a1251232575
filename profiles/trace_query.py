"""Debug timeline of CTA 0 of the tcgen05 MLP kernel (dfn_debug_trace): per layer, how long the MMA issuer waits
for activations / weight stages and how long the epilogue waits for the accumulator and works."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dfa_nerf_b200 as dfn  # noqa: E402
import synth  # noqa: E402

R, S = 40000, 192
mode = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
dev = torch.device('cuda', 0)
net = dfn.FaceNeRF(D=8, W=256, input_ch=63, input_ch_views=27, dim_aud=64, output_ch=4, skips=[4], use_viewdirs=True)
net.load_state_dict(synth.facenerf_state_dict(1))
net = net.to(dev)
fr = synth.frame_inputs(H=450, W=450, seed=0)
ro, rd, vd = dfn.get_rays(450, 450, fr['focal'], fr['c2w'], fr['cx'], fr['cy'], device=dev, return_viewdirs=True)
ro, rd, vd = [t.reshape(-1, 3)[:R].contiguous() for t in (ro, rd, vd)]
z, _ = torch.sort(torch.rand(R, S, device=dev) * 0.6 + 0.4, -1)
aud = fr['aud'].to(dev)
eng = dfn.RenderEngine(net, None, S, 0, precision={'bf16': dfn.PREC_BF16, 'bf16x3': dfn.PREC_BF16X3, 'fp16x3m': dfn.PREC_FP16X3M}[mode])
if os.environ.get('DFN_PP_FLAGS'):
    dfn.lib.dfn_debug_set_pp_flags(int(os.environ['DFN_PP_FLAGS']))      # schedule switches of mlp_pp.cu
eng.query_points(net, ro, rd, vd, z, aud)
T, NL = 6, 12
buf = torch.zeros(2 * T * NL * 8 + T * NL * 2 * 16, dtype=torch.int64, device=dev)
dfn.lib.dfn_debug_trace(C.c_void_p(buf.data_ptr()), T)
eng.query_points(net, ro, rd, vd, z, aud)
torch.cuda.synchronize()
dfn.lib.dfn_debug_trace(None, 0)
detail = buf.cpu()[2 * T * NL * 8:].reshape(T, NL, 2, 4, 4)
b = buf.cpu()[:2 * T * NL * 8].reshape(2, T, NL, 2, 4)
t0 = int(b[0, 0, 0, 0, 0])
nslot = 2 if mode == 'bf16' else 1
if len(sys.argv) > 2:
    nslot = int(sys.argv[2])
print('MMA issuer (cycles rel. to start): tile layer slot | wait_aready  issue(incl. full waits)  full_wait')
for j in range(2, 4):
    for l in range(NL):
        for s in range(nslot):
            w0, w1, e, fw = [int(x) for x in b[0, j, l, s]]
            print('  j=%d l=%2d s=%d  start %8d  wait_aready %6d  issue %6d  (full-wait %6d, peer-full-wait %6d)' % (j, l, s, w0 - t0, w1 - w0, e - w1, fw & 0xFFFFFFFF, fw >> 32))
print('EPILOGUE: tile layer slot | wait_acc  work')
for j in range(2, 4):
    for l in range(NL):
        for s in range(nslot):
            w0, w1, e, tt = [int(x) for x in b[1, j, l, s]]
            extra = ' (tile start->first wait: %d = PE)' % (w0 - tt) if l == 0 else ''
            print('  j=%d l=%2d s=%d  start %8d  wait_acc %6d  work %6d%s' % (j, l, s, w0 - t0, w1 - w0, e - w1, extra))

if int(detail.abs().sum()) != 0:
    print('PAIR kernel, per K-block of (j=3; l=2,3): producer saw the entry empty | issuer began waiting | own half landed | peer half landed')
    for l in (2, 3):
        for s in range(nslot):
            for k in range(4):
                pe, f0, f1, f2 = [int(x) - t0 for x in detail[3, l, s, :, k]]
                print('  l=%d s=%d kb=%d  empty %8d  wait-begin %8d  full %8d  pfull %8d' % (l, s, k, pe, f0, f1, f2))
