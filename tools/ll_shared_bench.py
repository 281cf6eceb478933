#!/usr/bin/env python
"""Integer tensor-core rows for several chains per tile (bnpc_ll_matrix_i8_shared) at a benchmark
shape: values against the one-chain call (bit-identical) and the FP64 matrix (quantisation bound),
and the launch time per chain with CUDA events for 1..8 chains per call.

    python tools/ll_shared_bench.py [N M K]          (default 100000 1000 24: C3)
    BNPC_LL_I8_SINGLE=1 python tools/ll_shared_bench.py    times the one-chain kernel of bnpc_tc_i8.cuh
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bnpc_b200 import _lib  # noqa: E402


def main():
    N, M, K0 = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (100000, 1000, 24)
    L = _lib.lib()
    dev = torch.device('cuda:0')
    rng = np.random.default_rng(1)
    W = 4 * ((M + 127) // 128)
    sp = lambda: torch.cuda.current_stream().cuda_stream
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
    single = bool(os.environ.get('BNPC_LL_I8_SINGLE'))
    NC = 8
    Ks = [K0, K0 + 4, K0 - 2, K0 + 1, K0, K0 + 3, K0 - 1, K0 + 2][:NC]
    lps, bss, llfs, vmaxs, kps = [], [], [], [], []
    for c, K in enumerate(Ks):
        theta = torch.as_tensor(np.clip(rng.random((K, M)), 1e-5, 1 - 1e-5).astype(np.float32), device=dev)
        lp = torch.zeros(2 * K * M, dtype=torch.float64, device=dev)
        L.logprob_tables(theta.data_ptr(), None, K, M, 0.2, 0.01, lp.data_ptr(), sp())
        kp = (K + 7) & ~7
        lps.append(lp); kps.append(kp)
        vmaxs.append(float(lp.abs().max().item()) * 1.0001)
        bss.append(torch.zeros(W * kp * 128, dtype=torch.uint8, device=dev))
        llfs.append(torch.full((N, kp), float('nan'), dtype=torch.float32, device=dev))

    def one(c, out=None):
        o = llfs[c] if out is None else out
        L.ll_matrix_i8(x1.data_ptr(), x0.data_ptr(), W, M, None, 1, N, lps[c].data_ptr(), bss[c].data_ptr(), Ks[c],
                       vmaxs[c], o.data_ptr(), kps[c], sp())

    def shared(nc):
        P = C.c_void_p * nc
        L.ll_matrix_i8_shared(x1.data_ptr(), x0.data_ptr(), W, M, N, nc,
                              P(*[lps[c].data_ptr() for c in range(nc)]), P(*[bss[c].data_ptr() for c in range(nc)]),
                              (C.c_int * nc)(*Ks[:nc]), (C.c_double * nc)(*vmaxs[:nc]),
                              P(*[llfs[c].data_ptr() for c in range(nc)]), (C.c_int * nc)(*kps[:nc]), sp())

    def timed(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps * 1e3

    ok = True
    # values: one chain per call
    refs = []
    for c in range(NC):
        ref = torch.full_like(llfs[c], float('nan'))
        one(c, ref)
        refs.append(ref)
    torch.cuda.synchronize()
    # FP64 check of chain 0 on a slice
    ns = min(N, 4096)
    ldk = Ks[0] | 1
    ll = torch.zeros(ns * ldk, dtype=torch.float64, device=dev)
    cells = torch.arange(ns, dtype=torch.int32, device=dev)
    L.ll_matrix(x1.data_ptr(), x0.data_ptr(), W, M, cells.data_ptr(), 1, ns, lps[0].data_ptr(), Ks[0], ll.data_ptr(),
                ldk, sp())
    torch.cuda.synchronize()
    want = ll.cpu().numpy().reshape(ns, ldk)[:, :Ks[0]]
    got = refs[0][:ns].cpu().numpy()[:, :Ks[0]].astype(np.float64)
    tol = M * vmaxs[0] / 65535 / 2 + 2.0 ** -22 * np.abs(want) + 1e-6
    bad = int((~(np.abs(got - want) <= tol)).sum())
    print(f'one chain vs FP64: max abs err {np.nanmax(np.abs(got - want)):.4g}, bad={bad}', flush=True)
    ok &= bad == 0
    if not single:
        for nc in (2, 3, 4, 5, 8):
            for c in range(nc):
                llfs[c].fill_(float('nan'))
            shared(nc)
            torch.cuda.synchronize()
            for c in range(nc):
                same = torch.equal(llfs[c][:, :Ks[c]], refs[c][:, :Ks[c]])
                if not same:
                    d = (llfs[c][:, :Ks[c]] != refs[c][:, :Ks[c]])
                    print(f'  nc={nc} chain {c}: {int(d.sum())} entries differ, nan={int(torch.isnan(llfs[c][:, :Ks[c]]).sum())}')
                ok &= same
        print('shared == single-chain values:', ok, flush=True)
    flops = lambda nc: 4.0 * N * M * sum(Ks[:nc])
    t1 = timed(lambda: one(0))
    print(f'{"single-chain kernel" if single else "shared kernel, nc=1"}: {t1:.1f} us per launch (incl. table split), '
          f'{flops(1) / t1 / 1e6:.0f} algorithmic TFLOP/s', flush=True)
    if not single:
        for nc in (2, 3, 4, 5, 8):
            t = timed(lambda: shared(nc))
            print(f'shared nc={nc} (K={Ks[:nc]}): {t:.1f} us per launch, {t / nc:.1f} us per chain, '
                  f'{flops(nc) / t / 1e6:.0f} algorithmic TFLOP/s', flush=True)
    return 0 if ok else 1


if __name__ == '__main__':
    sys.exit(main())
