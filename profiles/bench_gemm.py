"""The training step's three GEMM shapes in isolation (dfn_gemm, bf16x3): forward Y = X W^T, data gradient dX = dH W, weight gradient
dW = dH^T X with split-K, at the step's size (131,072 points, 256 x 256 layer): ms and TFLOP/s per launch.
    python profiles/bench_gemm.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dfa_nerf_b200 as dfn  # noqa: E402
from dfa_nerf_b200 import train  # noqa: E402

dev = 'cuda'
P, N, K = 131072, 256, 256
X = torch.randn(P, K, device=dev)
W = torch.randn(N, K, device=dev) * 0.05
b = torch.randn(N, device=dev)
Y = torch.empty(P, N, device=dev)
dH = torch.randn(P, N, device=dev)
dX = torch.empty(P, K, device=dev)
dW = torch.zeros(N, K, device=dev)


def timeit(name, fn, flops, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print('%-46s %8.3f ms  %7.1f TFLOP/s' % (name, ms, flops / ms / 1e9), flush=True)


for pname, prec in (('bf16x3', dfn.PREC_BF16X3), ('bf16', dfn.PREC_BF16)):
    fl = 2.0 * P * N * K
    timeit('%s forward  Y = relu(X W^T + b)' % pname, lambda: train.mm(X, W, Y, bias=b, act=1, precision=prec), fl)
    timeit('%s data grad dX = (dH * relu\'(Y)) W' % pname, lambda: train.mm(dH, W.t(), dX, mask=Y, mask_mode=1, precision=prec), fl)
    for splits in (74, 148, 296):
        timeit('%s weight grad dW += dH^T X, %d splits' % (pname, splits),
               lambda: train.mm(dH.t(), X.t(), dW, mask=Y.t(), mask_mode=1, beta=1, k_splits=splits, precision=prec), fl)
