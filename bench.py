#!/usr/bin/env python3
"""Benchmark of the BnpC MCMC hot path on B200 (contract: see the repo brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C3]

One "step" = one Chain.do_step() + Chain.update_results() (reference libs/MCMC.py:381-386)
of ONE chain.  Workload (default): BASELINE.json configs[2] = synthetic 100k cells x 1k
mutations, learned error rates, 10 % missing, 8 chains per GPU started from the true
assignment (steady state, K=20), default move probabilities.  Chains are independent:
with N GPUs every rank runs its own 8 chains (weak scaling, no data-path collective).

Printed JSON (rank 0, one line):
  value   box chain-steps/s with traces kept on the device (no per-step host copy of the
          assignment vector);   e2e = the same through the public driver API
          (libs.MCMC.Chain_steps: host-side traces, device->host copies every step).
  roofline  the dominant kernel (by CUDA-event time inside the timed region)
  cpu_baseline  the CPU restatement of the reference (oracle/) timed on the host cores on
          a bounded cell sample, extrapolated linearly in N (every per-step cost of the
          reference is linear in the number of cells at fixed K and M).
--impl reference times that CPU path with one process per chain on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from bnpc_b200.synth import CONFIGS, make_matrix  # noqa: E402

METRIC = 'MCMC steps/sec/chain and box chain-steps/sec at 100k cells x 1k mutations'
MOVES = dict(sm_prob=0.33, dpa_prob=0.25, error_prob=0.25, sm_ratios=[0.75, 0.25], sm_steps=3)
LEARN_KW = dict(FP_mean=0.01, FP_sd=0.01, FN_mean=0.2, FN_sd=0.1)       # run_BnpC.py defaults
REF_US_PER_CELL_STEP = 550.0     # planning anchor to size the CPU sample


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='C3', choices=sorted(CONFIGS))
    ap.add_argument('--chains-per-gpu', type=int, default=8)
    ap.add_argument('--cells', type=int, default=0, help='override the number of cells (debug)')
    ap.add_argument('--muts', type=int, default=0, help='override the number of mutations (debug)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-sample-cells', type=int, default=0)
    return ap.parse_args()


def workload(args):
    cfg = dict(CONFIGS[args.config])
    if args.cells:
        cfg['cells'] = args.cells
    if args.muts:
        cfg['muts'] = args.muts
    return cfg


def moves_of(cfg):
    m = dict(MOVES)
    if 'sm_prob' in cfg:
        m['sm_prob'] = cfg['sm_prob']
    if not cfg['learning']:
        m['error_prob'] = 0.0
    return m


def model_kwargs(cfg):
    if cfg['learning']:
        return dict(DP_alpha=[-1, -1], param_beta=list(cfg['pp']), **LEARN_KW)
    return dict(DP_alpha=[-1, -1], param_beta=list(cfg['pp']), FN_error=cfg['FN'], FP_error=cfg['FP'])


# ---------------------------------------------------------------------------------------
# CPU arm: the oracle restatement of the reference (the only place bench.py touches oracle/)
# ---------------------------------------------------------------------------------------
def _cpu_chain(job):
    """One chain of the CPU restatement on a cell sample; returns seconds for the timed steps."""
    cfg, sample_cells, seed, warm, steps = job
    from oracle.crp_oracle import OracleCRP, OracleCRPLearnErrors, do_step
    from oracle.rng_tape import LegacyRandom
    data, z = make_matrix(sample_cells, cfg['muts'], cfg['k_true'], cfg['fn'], cfg['fp'], cfg['miss'], seed=0)
    np.random.seed(seed)
    rnd = LegacyRandom()
    cls = OracleCRPLearnErrors if cfg['learning'] else OracleCRP
    model = cls(data, rnd=rnd, **model_kwargs(cfg))
    model.init(assign=[int(v) for v in z])
    moves = moves_of(cfg)

    def one():
        do_step(model, rnd, moves, cfg['learning'])
        ll = model.get_ll_full()                       # Chain.update_results, libs/MCMC.py:252-258
        _ = ll + model.get_lprior_full()
        _ = model.assignment.copy()

    for _ in range(warm):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    return time.perf_counter() - t0


def cpu_sample_cells(cfg, steps, warm, budget_s, forced=0):
    if forced:
        return min(forced, cfg['cells'])
    per_cell = REF_US_PER_CELL_STEP * 1e-6 * cfg['muts'] / 1000.0
    n = int(budget_s / max(1, steps + warm) / per_cell)
    return int(min(cfg['cells'], max(300, min(n, 5000))))


def run_cpu(cfg, n_chains, steps, warm, budget_s, forced=0):
    import multiprocessing as mp
    cores = max(1, min(n_chains, os.cpu_count() or 1))
    sample = cpu_sample_cells(cfg, steps, warm, budget_s, forced)
    jobs = [(cfg, sample, 1000 + c, warm, steps) for c in range(cores)]
    if cores == 1:
        secs = [_cpu_chain(jobs[0])]
    else:
        with mp.get_context('spawn').Pool(cores) as pool:
            secs = pool.map(_cpu_chain, jobs)
    scale = sample / cfg['cells']
    # chain-steps/s on the sample, scaled to the full number of cells (costs linear in N)
    box = sum(steps / s for s in secs) * scale
    return dict(value=box, unit='chain-steps/s', cores=cores, kind='port',
                sample=f'{sample} of {cfg["cells"]} cells x {cfg["muts"]} mutations, {steps} steps after '
                       f'{warm} warm-up per chain, {cores} chain(s) = {cores} process(es); rate scaled by '
                       f'{sample}/{cfg["cells"]} (per-step cost of the reference is linear in cells)',
                per_chain=box / cores, seconds=max(secs))


# ---------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(index), f'--query-gpu={self.Q}',
                                       '--format=csv,noheader,nounits', '-lms', '200'],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.p.terminate()
        self.p.wait()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        self.f.close()
        os.unlink(self.f.name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None,
                    sm_max_mhz=float(np.max(mx)) if mx else None, reasons=sorted(reasons),
                    samples=len(sm))


def measured_peaks():
    """(HBM GB/s, dense bf16 TFLOP/s sustained, source): the driver-written measurements of this
    pool's B200s, else the fallback of B200_PROFILING.md."""
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            d = json.load(f)
        return float(d['hbm_gbs']), float(d.get('bf16_tflops_sustained', d['bf16_tflops'])), \
            'measured (MEASURED_PEAKS.json; bf16 = sustained figure, the kernel runs inside a long step)'
    except Exception:
        return 6650.0, 1400.0, 'fallback (B200_PROFILING.md)'


def ncu_traffic():
    """DRAM bytes per launch of the profiled kernels (ncu --set full captures summarised under
    profiles/): {kernel: bytes} or {}."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'r1_traffic.json')) as f:
            return json.load(f)
    except Exception:
        return {}


def run_gpu(args, cfg):
    import torch
    import torch.distributed as dist
    from bnpc_b200 import _lib
    from bnpc_b200.rng import PhiloxRandom
    import libs.CRP as crp
    import libs.CRP_learning_errors as crple
    from libs.MCMC import Chain_steps

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    sys.setswitchinterval(2e-4)
    n_gpus = world
    cpg = args.chains_per_gpu
    K, W = args.steps, max(args.warmup, 3)
    N, M = cfg['cells'], cfg['muts']

    data, z = make_matrix(N, M, cfg['k_true'], cfg['fn'], cfg['fp'], cfg['miss'], seed=0)
    z_list = [int(v) for v in z]
    cls = crple.CRP_errors_learning if cfg['learning'] else crp.CRP
    proto = cls(data, **model_kwargs(cfg))
    moves = dict(moves_of(cfg), param_proposal_sd=np.array([0.1, 0.25, 0.5]))

    from copy import deepcopy
    chains = []

    def make_chain(c):
        m = deepcopy(proto)
        m.device = dev
        m.rnd = PhiloxRandom(4242 + rank * cpg + c)       # keyed by the global chain id
        m.init(assign=z_list)
        chains.append(Chain_steps(m, rank * cpg + c + 1, W + K + 2, 0, moves, 0, False))

    def build_chains():
        """(re)start all chains of this rank from the true assignment with their own seeds: the
        two legs then walk the SAME trajectory (the cost of a step depends on the chain state,
        which drifts while the chain runs)"""
        chains.clear()
        make_chain(0)                                      # packs the matrix once per device
        ths = [threading.Thread(target=make_chain, args=(c,)) for c in range(1, cpg)]
        [t.start() for t in ths]
        [t.join() for t in ths]
        chains.sort(key=lambda ch: ch.no)
        torch.cuda.synchronize()

    build_chains()
    dev_trace = [torch.zeros((4, N), dtype=torch.int32, device=dev) for _ in range(cpg)]

    def step_device(ch, i, tr):
        """do_step + the per-step trace with the assignment row kept on the device."""
        ch.do_step()
        ll = ch.model.get_ll_full()
        ch.results['ML'][i] = ll
        ch.results['MAP'][i] = ll + ch.model.get_lprior_full()
        ch.model.copy_assignment_to(tr[i % 4])

    def step_e2e(ch, i, tr):
        """the public driver path: Chain.do_step + Chain.update_results (host traces)."""
        ch.do_step()
        ch.update_results(i, False)

    def leg(step_fn, n_warm, n_timed, profile):
        bar = threading.Barrier(len(chains) + 1)
        errs = []

        def worker(ch, tr):
            try:
                torch.cuda.set_device(dev)                 # the current device is per host thread
                for i in range(n_warm):
                    step_fn(ch, 1 + i, tr)
                ch.model.stream.synchronize()
                ch.model.profile = profile
                ch.model.kernel_times_ms()
                ch.model.h2d_bytes = ch.model.d2h_bytes = 0
                bar.wait()
                bar.wait()
                for i in range(n_timed):
                    step_fn(ch, 1 + n_warm + i, tr)
                ch.model.stream.synchronize()
            except BaseException as e:                     # noqa: BLE001
                errs.append(e)
                bar.abort()

        ths = [threading.Thread(target=worker, args=(ch, tr)) for ch, tr in zip(chains, dev_trace)]
        [t.start() for t in ths]
        bar.wait()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local) if rank == 0 else None
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = _lib.launch_count()
        start.record()
        bar.wait()
        [t.join() for t in ths]
        torch.cuda.synchronize()
        end.record()
        end.synchronize()
        if errs:
            raise errs[0]
        ms = torch.tensor([start.elapsed_time(end)], device=dev, dtype=torch.float64)
        launches = torch.tensor([_lib.launch_count() - launches0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(launches, op=dist.ReduceOp.SUM)
        clocks = sampler.stop() if sampler else None
        return float(ms.item()), int(launches.item()), clocks

    # ---- leg 1: device-resident traces (value) ------------------------------------------
    ms_dev, launches, clocks = leg(step_device, W, K, True)
    ktimes, kwork = {}, {}
    for ch in chains:
        for name, v in ch.model.kernel_times_ms().items():
            ktimes.setdefault(name, []).extend(v)
        for name, v in ch.model.kernel_work().items():
            kwork.setdefault(name, []).extend(v)
        ch.model.profile = False
    sweep_stats = chains[0].model.sweep_stats
    k_live = len(chains[0].model.cells_per_cluster)
    # ---- leg 2: public driver API with host traces (e2e) --------------------------------
    build_chains()                                         # same seeds: the same K + W steps again
    ms_e2e, _, _ = leg(step_e2e, W, K, False)
    h2d = sum(ch.model.h2d_bytes for ch in chains) / (len(chains) * K)
    d2h = sum(ch.model.d2h_bytes for ch in chains) / (len(chains) * K)
    # the host-side int64 trace row is read back as int32 [N] + the theta snapshot + scalars
    total_chains = cpg * n_gpus

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return None

    value = total_chains * K / (ms_dev / 1e3)
    e2e = total_chains * K / (ms_e2e / 1e3)
    hbm_peak, bf16_peak, peak_src = measured_peaks()
    traffic = ncu_traffic()
    tot = {k: float(np.sum(v)) for k, v in ktimes.items()}
    n_unc = float(sweep_stats.get('uncertain', N))
    # algorithmic work per launch:
    #   likelihood rows  4*N*M*K flop (2 planes x multiply-add), N*M/4 + 16*K*M + 4*N*K bytes
    #   sweep            the records of the uncertain visits (160 B each) + 4 B per visit of output
    # per launch, from the live clusters / visits / uncertain visits of THAT launch (the list of
    # clusters grows while the chain runs), averaged over the launches of the timed region
    def mean_work(name, fn, fallback):
        w = kwork.get(name) or []
        return float(np.mean([fn(x) for x in w])) if w and len(w) == len(ktimes.get(name, [])) else fallback

    alg = {
        'll_matrix': dict(
            flops=mean_work('ll_matrix', lambda x: 4.0 * x['rows'] * M * x['K'], 4.0 * N * M * k_live),
            bytes=mean_work('ll_matrix', lambda x: x['rows'] * M / 4 + 16.0 * x['K'] * M + 4.0 * x['rows'] * x['K'],
                            N * M / 4 + 16.0 * k_live * M + 4.0 * N * k_live)),
        'gibbs_sweep': dict(
            flops=0.0,
            bytes=mean_work('gibbs_sweep', lambda x: 160.0 * x.get('n_unc', x['rows']) + 4.0 * x['rows'],
                            160.0 * n_unc + 4.0 * N)),
    }

    def roof_of(name):
        if name not in ktimes:
            return None
        avg_ms = float(np.mean(ktimes[name]))
        if name == 'll_matrix':
            ach = alg[name]['flops'] / (avg_ms * 1e-3) / 1e12
            return dict(kernel='ll_matrix_i8_kernel (tcgen05 kind::i8 likelihood rows) + its digit-table kernel',
                        bound='tensor', achieved=ach,
                        peak=bf16_peak, unit='TFLOP/s', frac=ach / bf16_peak,
                        traffic=traffic.get('ll_matrix_i8_kernel'), peak_source=peak_src, avg_launch_ms=avg_ms,
                        algorithmic_flops_per_launch=alg[name]['flops'],
                        algorithmic_bytes_per_launch=alg[name]['bytes'],
                        hbm_frac=alg[name]['bytes'] / (avg_ms * 1e-3) / 1e9 / hbm_peak,
                        note='algorithmic flops 4*N*M*K; the kernel executes two base-256 digits per '
                             'log-probability on the int8 tensor pipe (nominal peak 2 x bf16) and pads K to a '
                             'multiple of 8: algorithmic flops / measured bf16 peak = executed int8 ops / '
                             '(2 x measured bf16 peak)')
        ach = alg[name]['bytes'] / (avg_ms * 1e-3) / 1e9
        return dict(kernel='gibbs_sweep_kernel', bound='hbm', achieved=ach, peak=hbm_peak, unit='GB/s',
                    frac=ach / hbm_peak, traffic=traffic.get('gibbs_sweep_kernel'), peak_source=peak_src,
                    avg_launch_ms=avg_ms, algorithmic_bytes_per_launch=alg[name]['bytes'],
                    note='the sweep is a sequential chain of dependent categorical draws, one CTA per '
                         'chain: latency-bound by construction; the HBM fraction is reported for information')

    roof = roof_of(max(tot, key=tot.get)) if tot else None
    roof_ll = roof_of('ll_matrix')
    kernels = {k: dict(launches=len(v), avg_ms=float(np.mean(v)),
                       share_of_step=float(np.sum(v)) / (len(chains) * ms_dev))
               for k, v in ktimes.items()}
    out = dict(
        metric=METRIC, value=value, unit='chain-steps/s', n_gpus=n_gpus, steps=K, warmup=W,
        ms_per_step=ms_dev / K, higher_is_better=True, scaling='weak', vs_baseline=None,
        dtype='f64', data='synthetic',
        config=dict(workload=f'{args.config}: synthetic {N} cells x {M} mutations, K_true={cfg["k_true"]}, '
                             f'{int(cfg["miss"] * 100)}% missing, '
                             f'{"learned" if cfg["learning"] else "fixed"} error rates, start = true assignment',
                    chains_per_gpu=cpg, chains_total=total_chains, moves=moves_of(cfg),
                    live_clusters=k_live,
                    l2='no explicit flush: every step streams its own visit records, approximate rows, '
                       'option records and draws (about 130 B per cell per chain) besides the bit-planes: '
                       f'{cpg * 130 * N / 1e6 + N * M / 4e6:.0f} MB for the {cpg} concurrent chains vs 126 MB of L2'),
        steps_per_sec_per_chain=K / (ms_dev / 1e3),
        e2e=dict(value=e2e, unit='chain-steps/s', h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                 steps_per_sec_per_chain=K / (ms_e2e / 1e3), ms_per_step=ms_e2e / K,
                 api='libs.MCMC.Chain_steps.do_step + update_results (host traces)'),
        gpu_launches=launches, roofline=roof, roofline_likelihood=roof_ll, kernels=kernels,
        sweep=sweep_stats, clocks=clocks)
    if not args.no_cpu_baseline and n_gpus == 1:
        out['cpu_baseline'] = run_cpu(cfg, 1, 2, 1, 25.0, args.cpu_sample_cells)
    if world > 1:
        dist.destroy_process_group()
    return out


def _emit(line, fd):
    os.write(fd, (line + '\n').encode())


def main():
    args = parse()
    cfg = workload(args)
    rank = int(os.environ.get('RANK', 0))
    # stdout carries exactly ONE JSON line: everything libraries print there while the bench runs
    # (NCCL's version banner, warnings) is sent to stderr instead
    sys.stdout.flush()
    out_fd = os.dup(1)
    os.dup2(2, 1)
    if args.impl == 'reference':
        if rank != 0:
            return
        chains = args.chains_per_gpu * max(1, args.gpus)
        res = run_cpu(cfg, chains, args.steps, min(args.warmup, 1), 100.0, args.cpu_sample_cells)
        out = dict(impl='reference', metric=METRIC, value=res['value'], unit='chain-steps/s',
                   n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                   ms_per_step=1e3 * res['cores'] / res['value'] if res['value'] else None,
                   higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f64', data='synthetic',
                   config=dict(workload=f'{args.config}: synthetic {cfg["cells"]} cells x {cfg["muts"]} mutations '
                                        '(CPU restatement of the reference, bounded cell sample)',
                               chains_total=res['cores'], moves=moves_of(cfg)),
                   steps_per_sec_per_chain=res['per_chain'], cpu_baseline=res,
                   e2e=dict(value=res['value'], unit='chain-steps/s', h2d_bytes_per_step=0,
                            d2h_bytes_per_step=0), gpu_launches=0)
        _emit(json.dumps(out), out_fd)
        return
    out = run_gpu(args, cfg)
    sys.stdout.flush()
    if out is not None:
        _emit(json.dumps(out), out_fd)


if __name__ == '__main__':
    main()
