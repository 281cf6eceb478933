#!/usr/bin/env python3
"""tcgen05 rows (bnpc_ll_matrix_tc) against the FP64 matrix (bnpc_ll_matrix) on random data."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bnpc_b200 import _lib

L = _lib.lib()
dev = 'cuda'
sp = lambda: torch.cuda.current_stream().cuda_stream
ok_all = True
for (N, M, K) in [(300, 128, 5), (1000, 200, 16), (5000, 1000, 24), (100000, 1000, 24), (4096, 640, 64), (777, 50, 40)]:
    rng = np.random.default_rng(N + M + K)
    data = rng.integers(0, 2, (N, M)).astype(np.int8)
    data[rng.random((N, M)) < 0.1] = -1
    W = 4 * ((M + 127) // 128)
    x1 = torch.zeros((N, W), dtype=torch.int32, device=dev); x0 = torch.zeros_like(x1)
    n1 = torch.zeros(N, dtype=torch.int32, device=dev); n0 = torch.zeros_like(n1)
    d = torch.as_tensor(data, device=dev)
    L.pack_planes(None, d.data_ptr(), N, M, W, x1.data_ptr(), x0.data_ptr(), n1.data_ptr(), n0.data_ptr(), sp())
    theta = torch.as_tensor(np.clip(rng.random((K, M)), 1e-5, 1 - 1e-5).astype(np.float32), device=dev)
    lp = torch.zeros(2 * K * M, dtype=torch.float64, device=dev)
    L.logprob_tables(theta.data_ptr(), None, K, M, 0.2, 0.01, lp.data_ptr(), sp())
    cells = torch.as_tensor(rng.permutation(N).astype(np.int32), device=dev)
    ldk = K | 1
    ll = torch.zeros(N * ldk, dtype=torch.float64, device=dev)
    L.ll_matrix(x1.data_ptr(), x0.data_ptr(), W, M, cells.data_ptr(), 1, N, lp.data_ptr(), K, ll.data_ptr(), ldk, sp())
    kp = (K + 7) & ~7
    llf = torch.full((N, kp), float('nan'), dtype=torch.float32, device=dev)
    bs = torch.zeros(W * 2 * kp * 64, dtype=torch.int16, device=dev)
    torch.cuda.synchronize()
    L.ll_matrix_tc(x1.data_ptr(), x0.data_ptr(), W, M, cells.data_ptr(), 1, N, lp.data_ptr(), bs.data_ptr(), K,
                   llf.data_ptr(), kp, sp())
    torch.cuda.synchronize()
    want = ll.cpu().numpy().reshape(N, ldk)[:, :K]
    got = llf.cpu().numpy()[:, :K].astype(np.float64)
    err = np.abs(got - want)
    tol = 0.02 + 2e-4 * np.abs(want)
    bad = int((~(err <= tol)).sum())
    print(f'N={N} M={M} K={K}: max abs err {np.nanmax(err):.4g} (|ll| up to {np.abs(want).max():.1f}), bad={bad}, nan={int(np.isnan(got).sum())}')
    if bad:
        ok_all = False
        i, j = np.argwhere(~(err <= tol))[0]
        print('  first bad', i, j, got[i, j], want[i, j], 'row got', got[i, :6], 'want', want[i, :6])
    if N == 100000 and not bad:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for fn, name in ((lambda: L.ll_matrix_tc(x1.data_ptr(), x0.data_ptr(), W, M, cells.data_ptr(), 1, N, lp.data_ptr(), bs.data_ptr(), K, llf.data_ptr(), kp, sp()), 'tc'),
                         (lambda: L.ll_matrix(x1.data_ptr(), x0.data_ptr(), W, M, cells.data_ptr(), 1, N, lp.data_ptr(), K, ll.data_ptr(), ldk, sp()), 'fp64')):
            for _ in range(3): fn()
            a.record()
            for _ in range(10): fn()
            b.record(); torch.cuda.synchronize()
            print(f'  {name}: {a.elapsed_time(b) / 10 * 1e3:.1f} us per launch')
print('OK' if ok_all else 'FAILED')
