"""Profiling driver for the standalone encode kernel (dfn_embed, HELP:21-52: 12 B in + 252 B out per point):
    ncu --set full --clock-control none -k regex:embed_kernel -c 1 -o gpurun_out/prof_embed python profiles/prof_embed.py
Prints achieved GB/s from CUDA events (algorithmic bytes 264 B/point)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dfa_nerf_b200 as dfn  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 202500 * 64
dev = torch.device('cuda', 0)
x = torch.rand(P, 3, device=dev) * 2 - 1
embed, dim = dfn.get_embedder(10)
for _ in range(3):
    y = embed(x)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(5):
    y = embed(x)
ev1.record()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / 5
print('dfn_embed: %d points -> [%d, %d]: %.3f ms = %.0f GB/s algorithmic (264 B/point)' % (P, P, dim, ms, 264.0 * P / ms / 1e6))
