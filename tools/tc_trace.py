#!/usr/bin/env python3
"""Pipeline trace of the integer tensor-core rows (bnpc_ll_matrix_i8): clock64 stamps of CTA 0
(producer warp 0, MMA thread, epilogue warp 8) and timings against the number of columns."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from bnpc_b200 import _lib  # noqa: E402

L = _lib.lib()
dev = 'cuda'


def sp():
    return torch.cuda.current_stream().cuda_stream


N, M = 100000, 1000
rng = np.random.default_rng(1)
W = 4 * ((M + 127) // 128)
x1 = torch.zeros((N, W), dtype=torch.int32, device=dev)
x0 = torch.zeros_like(x1)
n1 = torch.zeros(N, dtype=torch.int32, device=dev)
n0 = torch.zeros_like(n1)
data = rng.integers(0, 2, (N, M)).astype(np.int8)
data[rng.random((N, M)) < 0.1] = -1
d = torch.as_tensor(data, device=dev)
L.pack_planes(None, d.data_ptr(), N, M, W, x1.data_ptr(), x0.data_ptr(), n1.data_ptr(), n0.data_ptr(), sp())
cells = torch.as_tensor(rng.permutation(N).astype(np.int32), device=dev)
ident = torch.arange(N, dtype=torch.int32, device=dev)
trace = torch.zeros(4096, dtype=torch.int64, device=dev)
for K in (8, 24, 64):
    theta = torch.as_tensor(np.clip(rng.random((K, M)), 1e-5, 1 - 1e-5).astype(np.float32), device=dev)
    lp = torch.zeros(2 * K * M, dtype=torch.float64, device=dev)
    L.logprob_tables(theta.data_ptr(), None, K, M, 0.2, 0.01, lp.data_ptr(), sp())
    vmax = float(lp.abs().max().item()) * 1.0001
    kp = (K + 7) & ~7
    llf = torch.zeros((N, kp), dtype=torch.float32, device=dev)
    bs = torch.zeros(W * 2 * kp * 64, dtype=torch.int16, device=dev)
    for order, cc in (('permuted', cells), ('identity', ident)):
        def run():
            L.ll_matrix_i8(x1.data_ptr(), x0.data_ptr(), W, M, cc.data_ptr(), 1, N, lp.data_ptr(), bs.data_ptr(), K,
                           vmax, llf.data_ptr(), kp, sp())
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            run()
        a.record()
        for _ in range(10):
            run()
        b.record()
        torch.cuda.synchronize()
        print(f'K={K} ({order} visiting order): {a.elapsed_time(b) / 10 * 1e3:.1f} us per launch', flush=True)
    if K == 24:
        L.debug_set_trace(trace.data_ptr())
        L.ll_matrix_i8(x1.data_ptr(), x0.data_ptr(), W, M, cells.data_ptr(), 1, N, lp.data_ptr(), bs.data_ptr(), K,
                       vmax, llf.data_ptr(), kp, sp())
        torch.cuda.synchronize()
        L.debug_set_trace(None)
        t = trace.cpu().numpy()
        n_st = W // 4
        prod = t[:3 * 48].reshape(48, 3)
        mma = t[1024:1024 + 4 * 48].reshape(48, 4)
        epi = t[2048:2048 + 12].reshape(6, 2)
        t0 = prod[0, 0]
        print(f'stages per tile {n_st}; cycles relative to the first producer stamp')
        print('stage | producer: data expanded, previous store done + slot free, store issued | MMA: stage full, first MMA issued, '
              'all 8 issued, committed')
        for i in range(48):
            print(f'{i:3d} | {prod[i, 0] - t0:7d} {prod[i, 1] - t0:7d} {prod[i, 2] - t0:7d} | '
                  f'{mma[i, 0] - t0:7d} {mma[i, 1] - t0:7d} {mma[i, 2] - t0:7d} {mma[i, 3] - t0:7d}')
        print('epilogue (accumulator full seen, row written):', [(int(x - t0), int(y - t0)) for x, y in epi])
