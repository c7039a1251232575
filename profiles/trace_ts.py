"""Timeline of CTA 0 of the TMEM-activation kernel (dfn_debug_trace): per (layer, half) the MMA issuer's
dependency / weight waits and the epilogue's accumulator wait and work time, in SM cycles."""
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
eng = dfn.RenderEngine(net, None, S, 0, precision={'bf16': dfn.PREC_BF16, 'bf16x3': dfn.PREC_BF16X3}[mode])
eng.query_points(net, ro, rd, vd, z, aud)
T, NL = 6, 12
buf = torch.zeros(2 * T * NL * 8, dtype=torch.int64, device=dev)
dfn.lib.dfn_debug_trace(C.c_void_p(buf.data_ptr()), T)
eng.query_points(net, ro, rd, vd, z, aud)
torch.cuda.synchronize()
dfn.lib.dfn_debug_trace(None, 0)
b = buf.cpu().reshape(2, T, NL, 2, 4)
t0 = int(b[0, 2, 0, 0, 0])
print('tile 2..3 of CTA 0; cycles relative to tile 2 start')
for i in range(2, 4):
    for l in range(NL):
        for h in range(2):
            m0, need, m1, full = [int(x) for x in b[0, i, l, h]]
            e0, e1, e2, _ = [int(x) for x in b[1, i, l, h]]
            if m0 == 0:
                continue
            print('i=%d l=%2d h=%d | MMA start %7d dur %5d (dep-wait %5d, weight-wait %5d) | EPI wait-from %7d waited %5d work %5d done %7d'
                  % (i, l, h, m0 - t0, m1 - m0, need, full, e0 - t0, e1 - e0, e2 - e1, e2 - t0))
