#!/usr/bin/env python3
"""GPU-idle fraction of the step loop, before and after the native group driver.

nsys is not part of this image; tools/timeline/cupti_timeline.cpp collects the same thing from
CUPTI activity records: start/end of every kernel (and memcpy) of a region of the process.  For 8
chains at C3 on one GPU the timed region is K steps of every chain under
  mirror-threads   the round-1 architecture: one Python thread per chain over the per-method mirror
                   (`Chain.do_step` + `Chain.update_results`)
  lockstep         native driver, phase-synchronous group of 8 (BNPC_LOCKSTEP=1)
  async            native driver, asynchronous wave scheduler (default: groups of 4)
and prints span, GPU-busy time (union of the kernel intervals), idle fraction, kernels, the sum of
the kernel durations and the largest number of kernels in flight.

    python tools/gpu_timeline.py [--steps 20] [--warmup 5] [--chains 8] > profiles/r2_gpu_timeline.txt
"""
import argparse
import ctypes as C
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
HERE = os.path.join(ROOT, 'tools', 'timeline')


def load_tracer():
    so = os.path.join(HERE, 'libcupti_timeline.so')
    if not os.path.exists(so):
        lib = '/usr/local/cuda/targets/x86_64-linux/lib'
        subprocess.run(['g++', '-O2', '-shared', '-fPIC', '-std=c++17', '-I/usr/local/cuda/include',
                        '-I/usr/local/cuda/targets/x86_64-linux/include', os.path.join(HERE, 'cupti_timeline.cpp'),
                        '-o', so, f'-L{lib}', '-lcupti', f'-Wl,-rpath,{lib}'], check=True)
    return C.CDLL(so)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--chains', type=int, default=8)
    ap.add_argument('--config', default='C3')
    args = ap.parse_args()
    import torch
    import bench
    import libs.MCMC as mcmc
    from bnpc_b200.group import ChainGroup
    tracer = load_tracer()
    cfg = dict(bench.CONFIGS[args.config])
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    b = bench.Bench(cfg, dev, 0, 1, args.chains)
    W, K, n = args.warmup, args.steps, args.chains

    def traced(run_warm, run_timed):
        run_warm()
        torch.cuda.synchronize()
        assert tracer.tl_start() == 0, 'CUPTI activity tracing could not be enabled'
        t0 = time.perf_counter()
        run_timed()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        out = (C.c_double * 8)()
        tracer.tl_stop(out)
        return wall, list(out)

    rows = []

    # ---- round-1 architecture: one interpreter thread per chain over the per-method mirror
    chains = b.chains(n, W + K + 1)

    def mirror(first, count):
        def work(ch):
            torch.cuda.set_device(dev)
            for s in range(first, first + count):
                ch.do_step()
                ch.update_results(s, False)
        ths = [threading.Thread(target=work, args=(ch,)) for ch in chains]
        [t.start() for t in ths]
        [t.join() for t in ths]
    rows.append(('mirror-threads (round-1 host)', ) + traced(lambda: mirror(1, W), lambda: mirror(1 + W, K)))

    # ---- native driver, both schedulers
    for label, env, gs in (('lockstep group of 8', {'BNPC_LOCKSTEP': '1'}, n), ('async groups of 4 (default)', {}, mcmc.GROUP_SIZE),
                           ('async group of 8 (one host thread)', {}, n)):
        os.environ.update(env)
        chains = b.chains(n, W + K + 1)
        parts = [chains[i:i + gs] for i in range(0, n, gs)]
        groups = [ChainGroup(p, b.moves, False) for p in parts]
        for ch in chains:
            ch._prepare_params(0)

        def run_all(first, count):
            ths = [threading.Thread(target=g.run, args=(first, count)) for g in groups]
            [t.start() for t in ths]
            [t.join() for t in ths]
        rows.append((label, ) + traced(lambda: run_all(1, W), lambda: run_all(1 + W, K)))
        for g in groups:
            g.close()
        for k in env:
            del os.environ[k]

    print(f'# GPU timeline (CUPTI activity records), {args.config}: {cfg["cells"]} cells x {cfg["muts"]} mutations, {n} chains '
          f'on one {torch.cuda.get_device_name(0)}, {K} steps per chain after {W} warm-up')
    print('# (tracing costs a few us per launch: compare the columns, not with bench.py)')
    print(f'{"host":38s} {"wall ms":>9s} {"span ms":>9s} {"busy ms":>9s} {"idle %":>7s} {"kernels":>8s} {"per c-step":>10s} '
          f'{"sum ms":>8s} {"in flight":>9s} {"chain-steps/s":>13s}')
    for label, wall, o in rows:
        span, busy, kernels, copies, ksum, depth = o[0], o[1], o[2], o[3], o[4], o[5]
        print(f'{label:38s} {wall:9.1f} {span:9.1f} {busy:9.1f} {100 * (1 - busy / span):7.1f} {int(kernels):8d} '
              f'{kernels / (n * K):10.1f} {ksum:8.1f} {int(depth):9d} {n * K / (wall / 1e3):13.0f}')


if __name__ == '__main__':
    main()
