#!/usr/bin/env python3
"""Tensor-core rows (bnpc_ll_matrix_tc: bf16-split, bnpc_ll_matrix_i8: integer digits) against the
FP64 matrix (bnpc_ll_matrix) on random data, with timings at the benchmark shapes."""
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


ok_all = True
SHAPES = [(300, 128, 5), (1000, 200, 16), (5000, 1000, 24), (4096, 640, 64), (777, 50, 40),
          (100000, 1000, 24), (10000, 500, 24), (50000, 5000, 24), (1000000, 50, 16)]
TIMED = {(100000, 1000, 24), (10000, 500, 24), (50000, 5000, 24), (1000000, 50, 16)}
for (N, M, K) in SHAPES:
    rng = np.random.default_rng(N + M + K)
    W = 4 * ((M + 127) // 128)
    x1 = torch.zeros((N, W), dtype=torch.int32, device=dev)
    x0 = torch.zeros_like(x1)
    n1 = torch.zeros(N, dtype=torch.int32, device=dev)
    n0 = torch.zeros_like(n1)
    step = max(1, (64 << 20) // M)
    for r0 in range(0, N, step):
        n = min(step, N - r0)
        data = rng.integers(0, 2, (n, M)).astype(np.int8)
        data[rng.random((n, M)) < 0.1] = -1
        d = torch.as_tensor(data, device=dev)
        L.pack_planes(None, d.data_ptr(), n, M, W, x1[r0:].data_ptr(), x0[r0:].data_ptr(), n1[r0:].data_ptr(),
                      n0[r0:].data_ptr(), sp())
        torch.cuda.synchronize()
    theta = torch.as_tensor(np.clip(rng.random((K, M)), 1e-5, 1 - 1e-5).astype(np.float32), device=dev)
    lp = torch.zeros(2 * K * M, dtype=torch.float64, device=dev)
    L.logprob_tables(theta.data_ptr(), None, K, M, 0.2, 0.01, lp.data_ptr(), sp())
    vmax = float(lp.abs().max().item()) * 1.0001
    cells = torch.as_tensor(rng.permutation(N).astype(np.int32), device=dev)
    ldk = K | 1
    ll = torch.zeros(N * ldk, dtype=torch.float64, device=dev)
    kp = (K + 7) & ~7
    llf = torch.full((N, kp), float('nan'), dtype=torch.float32, device=dev)
    bs = torch.zeros(W * 2 * kp * 64, dtype=torch.int16, device=dev)

    def fp64():
        L.ll_matrix(x1.data_ptr(), x0.data_ptr(), W, M, cells.data_ptr(), 1, N, lp.data_ptr(), K, ll.data_ptr(),
                    ldk, sp())

    def bf16():
        L.ll_matrix_tc(x1.data_ptr(), x0.data_ptr(), W, M, cells.data_ptr(), 1, N, lp.data_ptr(), bs.data_ptr(), K,
                       llf.data_ptr(), kp, sp())

    def i8():
        L.ll_matrix_i8(x1.data_ptr(), x0.data_ptr(), W, M, cells.data_ptr(), 1, N, lp.data_ptr(), bs.data_ptr(), K,
                       vmax, llf.data_ptr(), kp, sp())

    fp64()
    torch.cuda.synchronize()
    want = ll.cpu().numpy().reshape(N, ldk)[:, :K]
    for name, fn, tol in (('bf16', bf16, 0.02 + 2e-4 * np.abs(want)),
                          ('i8', i8, M * vmax / 65535 / 2 + 2.0 ** -22 * np.abs(want) + 1e-6)):
        llf.fill_(float('nan'))
        fn()
        torch.cuda.synchronize()
        got = llf.cpu().numpy()[:, :K].astype(np.float64)
        err = np.abs(got - want)
        bad = int((~(err <= tol)).sum())
        print(f'N={N} M={M} K={K} {name}: max abs err {np.nanmax(err):.4g} (|ll| up to {np.abs(want).max():.1f}, '
              f'bound {np.max(tol):.3g}), bad={bad}, nan={int(np.isnan(got).sum())}', flush=True)
        if bad:
            ok_all = False
            i, j = np.argwhere(~(err <= tol))[0]
            print('  first bad', i, j, got[i, j], want[i, j], 'row got', got[i, :6], 'want', want[i, :6])
    if (N, M, K) in TIMED:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for fn, name in ((bf16, 'bf16'), (i8, 'i8'), (fp64, 'fp64')):
            for _ in range(3):
                fn()
            a.record()
            for _ in range(10):
                fn()
            b.record()
            torch.cuda.synchronize()
            us = a.elapsed_time(b) / 10 * 1e3
            print(f'  {name}: {us:.1f} us per launch ({4.0 * N * M * K / us / 1e6:.1f} algorithmic TFLOP/s)', flush=True)
print('OK' if ok_all else 'FAILED')
