#!/usr/bin/env python3
"""Benchmark of the BnpC MCMC hot path on B200 (contract: see the repo brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C3]

One "step" = one Chain.do_step() + Chain.update_results() (reference libs/MCMC.py:381-386) of ONE
chain.  Workload (default): BASELINE.json configs[2] = synthetic 100k cells x 1k mutations, learned
error rates, 10 % missing, 8 chains per GPU started from the true assignment (K = 20), default move
probabilities.  Chains are independent: with N GPUs every rank runs its own 8 chains (weak scaling,
no data-path collective).

Both legs go through the public driver (libs.MCMC.Chain_steps objects stepped by
libs.MCMC / bnpc_b200.group.ChainGroup, what MCMC.run does):
  value   box chain-steps/s with the assignment trace kept in the device ring (scalar traces and
          theta rows still reach the host);
  e2e     the same with the full host-side traces of the reference (assignment vector, theta rows,
          ML/MAP/alpha/FN/FP copied to pinned host arrays every step inside the timed region).
A timed window is W warm-up + K timed steps from freshly initialised chains (same seeds every
window, so every window walks the same trajectory); it is repeated --windows times and the MEDIAN
is reported with the spread.  Extra figures: one chain alone, the evolved state (250 more warm-up
steps), the other BASELINE shapes (C2, C4, C5; one short window each).
  roofline      the kernel with the largest share of the device time of the step (CUDA events
                around every launch, a separate untimed pass), with its algorithmic bytes / flops
  cpu_baseline  the UNMODIFIED reference classes (baseline/_ref, shipped by oracle/fetch_ref.py;
                else the oracle port) on the host cores on a bounded cell sample
--impl reference times the reference's own Chain_steps.do_step + update_results on the FULL
workload, one process per chain (its mp.Pool layout), for as many steps as fit its time budget.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
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
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='C3', choices=sorted(CONFIGS))
    ap.add_argument('--chains-per-gpu', type=int, default=8)
    ap.add_argument('--windows', type=int, default=5, help='repetitions of the timed window (median reported)')
    ap.add_argument('--cells', type=int, default=0, help='override the number of cells (debug)')
    ap.add_argument('--muts', type=int, default=0, help='override the number of mutations (debug)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip single-chain / evolved-state / C2,C4,C5 figures')
    ap.add_argument('--cpu-sample-cells', type=int, default=0)
    ap.add_argument('--group-size', type=int, default=0,
                    help='chains per lockstep group (0: libs.MCMC.GROUP_SIZE); experiment switch')
    ap.add_argument('--ref-budget-s', type=float, default=240.0, help="--impl reference: seconds of stepping")
    return ap.parse_args()


def workload(args, name=None):
    cfg = dict(CONFIGS[name or args.config])
    if args.cells and name is None:
        cfg['cells'] = args.cells
    if args.muts and name is None:
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


def describe(name, cfg):
    return (f'{name}: synthetic {cfg["cells"]} cells x {cfg["muts"]} mutations, K_true={cfg["k_true"]}, '
            f'{int(cfg["miss"] * 100)}% missing, {"learned" if cfg["learning"] else "fixed"} error rates, '
            'start = true assignment')


# ---------------------------------------------------------------------------------------
# CPU arm: the unmodified reference (baseline/_ref) or, without it, the oracle port -- the only
# place bench.py touches oracle/
# ---------------------------------------------------------------------------------------
_CPU_DATA = {}


def _cpu_chain(job):
    """One chain on the host: the reference's own Chain_steps (do_step + update_results), timed
    step by step until `budget_s` of stepping or `steps` steps; returns the per-step seconds."""
    kind, cfg, cells, seed, warm, steps, budget_s = job
    data, z = _CPU_DATA.get('d') or make_matrix(cells, cfg['muts'], cfg['k_true'], cfg['fn'], cfg['fp'],
                                                cfg['miss'], seed=0)
    moves = moves_of(cfg)
    np.random.seed(seed)
    if kind == 'reference':
        from oracle.ref_shim import load_reference, ref_errstate
        ref = load_reference(with_mcmc=True)
        cls = ref.CRP_learning_errors.CRP_errors_learning if cfg['learning'] else ref.CRP.CRP
        with ref_errstate():
            model = cls(data, **model_kwargs(cfg))
            model.init(assign=[int(v) for v in z])
            params = dict(moves, param_proposal_sd=np.array([0.1, 0.25, 0.5]))
            chain = ref.MCMC.Chain_steps(model, 1, warm + steps, 0, params, 0, False)

            def one(i):
                chain.do_step()                       # libs/MCMC.py:320-342
                chain.update_results(i, False)        # libs/MCMC.py:242-282
            return _timed_steps(one, warm, steps, budget_s)
    from oracle.crp_oracle import OracleCRP, OracleCRPLearnErrors, do_step
    from oracle.rng_tape import LegacyRandom
    rnd = LegacyRandom()
    model = (OracleCRPLearnErrors if cfg['learning'] else OracleCRP)(data, rnd=rnd, **model_kwargs(cfg))
    model.init(assign=[int(v) for v in z])

    def one(i):
        do_step(model, rnd, moves, cfg['learning'])
        ll = model.get_ll_full()                       # Chain.update_results, libs/MCMC.py:252-258
        _ = ll + model.get_lprior_full()
        _ = model.assignment.copy()
    return _timed_steps(one, warm, steps, budget_s)


def _timed_steps(one, warm, steps, budget_s):
    for i in range(warm):
        one(1 + i)
    secs, t_begin = [], time.perf_counter()
    for i in range(steps):
        t0 = time.perf_counter()
        one(1 + warm + i)
        secs.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin > budget_s:
            break
    return secs


def cpu_kind():
    from oracle.ref_shim import reference_available
    return 'reference' if reference_available() else 'port'


def run_cpu(cfg, n_chains, steps, warm, budget_s, cells):
    """n_chains chains, one process each on the host cores (the reference's mp.Pool layout,
    libs/MCMC.py:100,113); chains beyond the core count queue, as in the reference."""
    import multiprocessing as mp
    kind = cpu_kind()
    if kind == 'reference':
        from oracle.ref_shim import load_reference
        load_reference(with_mcmc=True)                      # imported once, inherited by the workers
    cores = max(1, min(n_chains, os.cpu_count() or 1))
    _CPU_DATA['d'] = make_matrix(cells, cfg['muts'], cfg['k_true'], cfg['fn'], cfg['fp'], cfg['miss'], seed=0)
    jobs = [(kind, cfg, cells, 1000 + c, warm, steps, budget_s) for c in range(cores)]
    if cores == 1:
        per_chain = [_cpu_chain(jobs[0])]
    else:
        with mp.get_context('fork').Pool(cores) as pool:     # the matrix is shared copy-on-write
            per_chain = pool.map(_cpu_chain, jobs)
    _CPU_DATA.clear()
    done = min(len(s) for s in per_chain)                   # steps every chain completed
    wall = max(sum(s[:done]) for s in per_chain)            # the slowest chain bounds the run
    return dict(kind=kind, cores=cores, chains=cores, steps_done=done, seconds=wall,
                box_rate=cores * done / wall, per_chain_rate=done / wall,
                step_seconds_mean=float(np.mean([np.mean(s[:done]) for s in per_chain])))


def cpu_baseline_sample(cfg, cells_forced=0):
    """the bounded CPU leg of the default run: one chain on a cell sample (10-30 s of work)"""
    per_cell = REF_US_PER_CELL_STEP * 1e-6 * cfg['muts'] / 1000.0
    cells = cells_forced or int(min(cfg['cells'], max(300, min(int(25.0 / 3 / per_cell), 5000))))
    r = run_cpu(cfg, 1, 2, 1, 60.0, cells)
    scale = cells / cfg['cells']
    return dict(value=r['box_rate'] * scale, unit='chain-steps/s', cores=1, kind=r['kind'],
                sample=f'{cells} of {cfg["cells"]} cells x {cfg["muts"]} mutations, {r["steps_done"]} steps after 1 '
                       f'warm-up, 1 chain = 1 process; rate scaled by {cells}/{cfg["cells"]} (every per-step cost '
                       'of the reference is linear in cells); the full-size figure is `--impl reference`',
                seconds=r['seconds'],
                implementation='unmodified reference classes (baseline/_ref), numpy stand-in for bottleneck'
                if r['kind'] == 'reference' else 'oracle/crp_oracle.py (CPU restatement of the reference)')


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
                                       '--format=csv,noheader,nounits', '-lms', '100'],
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
    """(HBM GB/s, dense bf16 TFLOP/s burst, sustained, source): the driver-written measurements of
    this pool's B200s, else the fallback of B200_PROFILING.md."""
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            d = json.load(f)
        return (float(d['hbm_gbs']), float(d['bf16_tflops']), float(d.get('bf16_tflops_sustained', d['bf16_tflops'])),
                'measured (MEASURED_PEAKS.json)')
    except Exception:
        return 6650.0, 1590.0, 1400.0, 'fallback (B200_PROFILING.md)'


def ncu_traffic():
    """DRAM bytes per launch of the profiled kernels (ncu --set full captures summarised under
    profiles/): {kernel: {bytes, chains}} or {}."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'r2_traffic.json')) as f:
            return {k: v for k, v in json.load(f).items() if isinstance(v, dict)}
    except Exception:
        return {}


class Bench:
    """chains of one rank on one device + the timing protocol"""

    def __init__(self, cfg, dev, rank, world, cpg):
        import torch
        import libs.CRP as crp
        import libs.CRP_learning_errors as crple
        self.torch = torch
        self.cfg, self.dev, self.rank, self.world, self.cpg = cfg, dev, rank, world, cpg
        self.N, self.M = cfg['cells'], cfg['muts']
        data, z = make_matrix(self.N, self.M, cfg['k_true'], cfg['fn'], cfg['fp'], cfg['miss'], seed=0)
        self.z_list = [int(v) for v in z]
        cls = crple.CRP_errors_learning if cfg['learning'] else crp.CRP
        self.proto = cls(data, **model_kwargs(cfg))
        self.moves = dict(moves_of(cfg), param_proposal_sd=np.array([0.1, 0.25, 0.5]))

    def chains(self, n, total_steps):
        """n freshly initialised chains (seeds keyed by the global chain id) with trace room"""
        from copy import deepcopy
        from bnpc_b200.rng import PhiloxRandom
        from libs.MCMC import Chain_steps
        out = []
        for c in range(n):
            m = deepcopy(self.proto)
            m.device = self.dev
            m.rnd = PhiloxRandom(4242 + self.rank * self.cpg + c)
            m.init(assign=self.z_list)
            out.append(Chain_steps(m, self.rank * self.cpg + c + 1, total_steps, 0, self.moves, 0, False))
        self.torch.cuda.synchronize()
        return out

    def window(self, n, warm, steps, host_assign, profile=False, group_size=0):
        """one timed window: fresh chains, `warm` untimed steps, `steps` timed steps between
        barriers + device synchronisation; returns (ms [max over ranks], launches, extra).
        group_size: chains per lockstep group (0: libs.MCMC.GROUP_SIZE, the driver's default); several
        groups of one device are stepped by one host thread each, as libs.MCMC.run_chains does."""
        import threading
        import torch.distributed as dist
        from bnpc_b200 import _lib
        from bnpc_b200.group import ChainGroup
        import libs.MCMC as mcmc
        torch = self.torch
        chains = self.chains(n, warm + steps + 1)
        gs = group_size or mcmc.GROUP_SIZE
        parts = [chains[i:i + gs] for i in range(0, n, gs)]
        groups = [ChainGroup(p, self.moves, False, host_assign=host_assign) for p in parts]
        for ch in chains:
            ch._prepare_params(0)

        def run_all(step0, count):
            if len(groups) == 1:
                groups[0].run(step0, count)
                return
            errs = []

            def work(g):
                try:
                    g.run(step0, count)
                except BaseException as exc:          # noqa: BLE001
                    errs.append(exc)
            ths = [threading.Thread(target=work, args=(g,)) for g in groups]
            [t.start() for t in ths]
            [t.join() for t in ths]
            if errs:
                raise errs[0]
        try:
            run_all(1, warm)
            if self.world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            launches0 = _lib.launch_count()
            if profile:
                _lib.lib().prof_enable(1)
            a.record()
            run_all(1 + warm, steps)
            torch.cuda.synchronize()
            b.record()
            b.synchronize()
            prof = None
            if profile:
                _lib.lib().prof_enable(0)
                prof = _lib.lib().prof_report_text()
            ms = torch.tensor([a.elapsed_time(b)], device=self.dev, dtype=torch.float64)
            launches = torch.tensor([_lib.launch_count() - launches0], device=self.dev, dtype=torch.float64)
            if self.world > 1:
                dist.barrier()
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                dist.all_reduce(launches, op=dist.ReduceOp.SUM)
            m0 = chains[0].model
            extra = dict(k_live=len(m0.cells_per_cluster), sweep=dict(m0.sweep_stats), prof=prof,
                         ml_last=float(chains[0].results['ML'][warm + steps]),
                         k_all=[len(ch.model.cells_per_cluster) for ch in chains], groups=len(groups))
            return float(ms.item()), int(launches.item()), extra
        finally:
            for g in groups:
                g.close()

    def trace_bytes_per_step(self, k_live):
        """device->host bytes of one chain-step of the e2e leg, counted from the copies the driver
        makes: the assignment row, the theta rows of the live clusters, five trace scalars, the
        status words / live list / decision scalars of the phases (<= 1 KB)"""
        return 4 * self.N + 4 * k_live * self.M + 5 * 8 + 1024

    def repeat(self, n, warm, steps, windows, host_assign, group_size=0):
        ms, launches, extra = [], [], None
        for _ in range(windows):
            t, l, extra = self.window(n, warm, steps, host_assign, group_size=group_size)
            ms.append(t)
            launches.append(l)
        return ms, launches, extra


def _rate(n_chains, steps, ms):
    return n_chains * steps / (ms / 1e3)


def _spread(vals):
    v = np.asarray(vals, dtype=np.float64)
    med = float(np.median(v))
    return dict(median=med, min=float(v.min()), max=float(v.max()),
                rel_spread=float((v.max() - v.min()) / med) if med else None, windows=len(vals))


def kernel_table(prof, steps, n_chains, step_ms):
    """per kernel: launches and device time per chain-step, share of the summed device time"""
    total = sum(v['total_ms'] for v in prof.values()) or 1.0
    rows = {}
    for name, v in sorted(prof.items(), key=lambda kv: -kv[1]['total_ms']):
        rows[name] = dict(launches_per_step=v['launches'] / steps, chains_per_launch=v['chains'] / max(1, v['launches']),
                          avg_launch_us=1e3 * v['total_ms'] / max(1, v['launches']),
                          us_per_chain_step=1e3 * v['total_ms'] / (steps * n_chains),
                          share_of_device_time=v['total_ms'] / total)
    return rows, total / steps


def roofline_of(name, row, work, peaks, traffic):
    """roofline entry of one kernel from its measured launch time and its algorithmic work per
    launch; work = dict(N, M, K, n_unc, chains)"""
    hbm, bf16_burst, bf16_sus, src = peaks
    N, M, K, n_unc = work['N'], work['M'], work['K'], work['n_unc']
    chains = row['chains_per_launch']
    sec = row['avg_launch_us'] * 1e-6
    base = name.split('<')[0]
    out = dict(kernel=name, avg_launch_us=row['avg_launch_us'], chains_per_launch=chains, peak_source=src,
               share_of_device_time=row['share_of_device_time'])
    cap = traffic.get(base)
    # DRAM bytes of an ncu capture of this kernel, scaled from the chains of the captured launch to
    # the chains per launch measured here (chains of one launch share the bit-planes in L2)
    out['traffic'] = cap['bytes'] * chains / cap['chains'] if cap else None
    if base in ('ll_matrix_i8_kernel', 'll_matrix_i8s_kernel'):
        flops = 4.0 * N * M * K * chains
        ach = flops / sec / 1e12
        out.update(bound='tensor', achieved=ach, peak=bf16_sus, unit='TFLOP/s', frac=ach / bf16_sus,
                   algorithmic_flops_per_launch=flops,
                   note='4*N*M*K algorithmic flops per chain; the kernel executes two base-256 digits per '
                        'log-probability on the int8 tensor pipe (nominal 2 x bf16); peak = sustained bf16'
                        + ('; ll_matrix_i8s: the chains of a launch whose digit tables fit one MMA side by side '
                           '(sum of 2*Kp <= 256) share the expanded data operand in tensor memory'
                           if base == 'll_matrix_i8s_kernel' else ''))
        return out
    per_chain = {
        'gibbs_sweep_kernel': 160.0 * n_unc + 4.0 * N,                 # records of the uncertain visits + assignments
        'gibbs_exact_kernel': 160.0 * n_unc + 16.0 * K * M,            # records written + table read
        'suffstat_kernel': N * M / 4.0 + 4.0 * N + 8.0 * K * M,        # planes + members + S1/S0
        'mh_theta_kernel': 36.0 * K * M,                               # theta, S1, S0, three draws
        'gibbs_options_kernel': 4.0 * N * ((K + 7) // 8 * 8) + 16.0 * N,
        'rg_serial_kernel': 8.0 * N / max(K, 1),
        'll_few_kernel': (N / max(K, 1)) * (M / 4.0 + 16.0),
    }.get(base)
    if per_chain is None:
        per_chain = 4.0 * N
    byts = per_chain * chains
    ach = byts / sec / 1e9
    out.update(bound='hbm', achieved=ach, peak=hbm, unit='GB/s', frac=ach / hbm, algorithmic_bytes_per_launch=byts)
    if base in ('gibbs_sweep_kernel', 'rg_serial_kernel'):
        out['note'] = ('sequential chain of dependent categorical draws, one CTA per chain: latency-bound by '
                       'construction; the HBM fraction is reported for information')
    return out


def run_gpu(args, cfg):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    cpg, K, W, R = args.chains_per_gpu, args.steps, max(args.warmup, 3), max(1, args.windows)
    bench = Bench(cfg, dev, rank, world, cpg)
    N, M = bench.N, bench.M

    bench.window(cpg, 2, 3, True)                      # library, allocator and clocks warm
    sampler = ClockSampler(local) if rank == 0 else None
    ms_dev, launches, ex_dev = bench.repeat(cpg, W, K, R, host_assign=False, group_size=args.group_size)
    ms_e2e, _, ex_e2e = bench.repeat(cpg, W, K, R, host_assign=True, group_size=args.group_size)
    clocks = sampler.stop() if sampler else None
    total_chains = cpg * world
    med_dev, med_e2e = float(np.median(ms_dev)), float(np.median(ms_e2e))
    value, e2e = _rate(total_chains, K, med_dev), _rate(total_chains, K, med_e2e)

    # per-kernel device times: CUDA events around EVERY launch, a separate untimed pass
    # (phase-synchronous driver on one stream for this pass: launches of different chains merge and
    # nothing overlaps, so an event pair brackets exactly one launch)
    os.environ['BNPC_LOCKSTEP'] = os.environ['BNPC_ONE_STREAM'] = '1'
    _, _, ex_prof = bench.window(cpg, W, K, False, profile=True, group_size=cpg)
    del os.environ['BNPC_LOCKSTEP'], os.environ['BNPC_ONE_STREAM']
    extras = {}
    if not args.no_extras:
        s_ms, s_l, _ = bench.repeat(1, W, K, min(R, 3), host_assign=True)
        extras['single_chain'] = dict(value=_rate(world, K, float(np.median(s_ms))), unit='chain-steps/s',
                                      steps_per_sec_per_chain=_rate(1, K, float(np.median(s_ms))),
                                      ms_per_step=float(np.median(s_ms)) / K, launches_per_step=float(np.median(s_l)) / K / world,
                                      windows_ms=_spread(s_ms), api='e2e (host traces), 1 chain on 1 GPU')
        ev_ms, _, ev_ex = bench.window(cpg, 250 + W, K, True, group_size=args.group_size)
        extras['evolved_state'] = dict(value=_rate(total_chains, K, ev_ms), unit='chain-steps/s', warmup_steps=250 + W,
                                       ms_per_step=ev_ms / K, live_clusters=ev_ex['k_all'], sweep=ev_ex['sweep'],
                                       api='e2e (host traces)')
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return None

    peaks = measured_peaks()
    traffic = ncu_traffic()
    k_live = ex_prof['k_live']
    n_unc = float(ex_prof['sweep'].get('uncertain', N) or N)
    ktab, dev_ms_per_step = kernel_table(ex_prof['prof'], K, cpg, med_dev / K)
    work = dict(N=N, M=M, K=k_live, n_unc=n_unc)
    top = next(iter(ktab))
    roof = roofline_of(top, ktab[top], work, peaks, traffic)
    # the rows of a sweep's first epoch (cell order, tiles shared by the chains of a launch); the
    # one-chain kernel only serves the gathered rows of later epochs
    ll_name = next((n for n in ktab if n.startswith('ll_matrix_i8s_kernel')), None) or \
        next((n for n in ktab if n.startswith('ll_matrix_i8_kernel')), None)
    roof_ll = roofline_of(ll_name, ktab[ll_name], work, peaks, traffic) if ll_name else None
    d2h = bench.trace_bytes_per_step(ex_e2e['k_live'])
    out = dict(
        metric=METRIC, value=value, unit='chain-steps/s', n_gpus=world, steps=K, warmup=W,
        ms_per_step=med_dev / K, higher_is_better=True, scaling='weak', vs_baseline=None,
        dtype='f64', data='synthetic',
        config=dict(workload=describe(args.config, cfg), chains_per_gpu=cpg, chains_total=total_chains,
                    moves=moves_of(cfg), live_clusters=ex_dev['k_all'], lockstep_groups_per_gpu=ex_dev['groups'],
                    windows=f'{R} timed windows of {W} warm-up + {K} timed steps, fresh chains with the same seeds '
                            'per window; median reported',
                    l2='no explicit flush: per step every chain streams its own visit records, approximate rows, '
                       'option records and draws (about 130 B per cell per chain) besides the shared bit-planes: '
                       f'{cpg * 130 * N / 1e6 + N * M / 4e6:.0f} MB for the {cpg} concurrent chains vs 126 MB of L2'),
        steps_per_sec_per_chain=K / (med_dev / 1e3),
        windows_ms=_spread(ms_dev),
        e2e=dict(value=e2e, unit='chain-steps/s', h2d_bytes_per_step=4.0 * 4 * k_live + 64, d2h_bytes_per_step=d2h,
                 steps_per_sec_per_chain=K / (med_e2e / 1e3), ms_per_step=med_e2e / K, windows_ms=_spread(ms_e2e),
                 api='libs.MCMC.Chain_steps chains stepped by libs.MCMC.run_chains / bnpc_b200.group.ChainGroup (what '
                     'MCMC.run does): full host traces, copied from the device rings by a copy stream'),
        gpu_launches=int(np.median(launches)), launches_per_chain_step=float(np.median(launches)) / (K * total_chains),
        roofline=roof, roofline_likelihood=roof_ll, kernels=ktab,
        device_time=dict(sum_of_kernel_ms_per_step=dev_ms_per_step, wall_ms_per_step=med_dev / K,
                         note='sum of the per-launch device times of a step of all chains of the GPU (profiling pass: '
                              'phase-synchronous driver on ONE stream, every launch between its own pair of events) '
                              'against the wall time of such a step in the timed windows, where the waves of the '
                              'asynchronous scheduler overlap on the GPU (sum > wall)'),
        sweep=ex_prof['sweep'], clocks=clocks, **extras)
    if not args.no_extras and world == 1:
        out['other_configs'] = other_configs(args, dev)
    if not args.no_cpu_baseline and world == 1:
        out['cpu_baseline'] = cpu_baseline_sample(cfg, args.cpu_sample_cells)
    if world > 1:
        dist.destroy_process_group()
    return out


def other_configs(args, dev):
    """one short e2e window of the other BASELINE shapes (also C2, C4, C5)"""
    import torch
    out = {}
    for name, n_chains in (('C2', 1), ('C4', 1), ('C5', 1)):
        if name == args.config:
            continue
        try:
            cfg = workload(args, name)
            b = Bench(cfg, dev, 0, 1, n_chains)
            steps = 20 if name != 'C5' else 6
            ms, launches, ex = b.window(n_chains, 3, steps, True)
            out[name] = dict(workload=describe(name, cfg), chains=n_chains, steps=steps, warmup=3,
                             value=_rate(n_chains, steps, ms), unit='chain-steps/s', ms_per_step=ms / steps,
                             live_clusters=ex['k_all'], sweep=ex['sweep'], api='e2e (host traces)')
            del b
            torch.cuda.empty_cache()
        except Exception as exc:                              # noqa: BLE001  (reported, not hidden)
            out[name] = dict(error=repr(exc))
    return out


def run_reference(args, cfg):
    chains = args.chains_per_gpu * max(1, args.gpus)
    warm = min(args.warmup, 1)
    cells = args.cpu_sample_cells or cfg['cells']
    r = run_cpu(cfg, chains, args.steps, warm, args.ref_budget_s, cells)
    scale = cells / cfg['cells']
    value = r['box_rate'] * scale
    sample = (f'{cells} of {cfg["cells"]} cells x {cfg["muts"]} mutations (FULL workload), ' if scale == 1.0 else
              f'{cells} of {cfg["cells"]} cells (rate scaled by {scale:g}), ')
    sample += (f'{r["steps_done"]} timed steps of the {args.steps} requested (a step of this sample takes '
               f'{r["step_seconds_mean"]:.1f} s per chain; stepping is bounded to {args.ref_budget_s:.0f} s) after '
               f'{warm} warm-up, {r["chains"]} chains = {r["cores"]} processes on {os.cpu_count()} host cores')
    return dict(impl='reference', metric=METRIC, value=value, unit='chain-steps/s', n_gpus=args.gpus,
                steps=r['steps_done'], steps_requested=args.steps, warmup=warm,
                ms_per_step=1e3 * r['seconds'] / max(1, r['steps_done']),
                higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f64', data='synthetic',
                config=dict(workload=describe(args.config, cfg), chains_total=r['chains'], moves=moves_of(cfg),
                            chains_requested=chains),
                steps_per_sec_per_chain=r['per_chain_rate'] * scale,
                cpu_baseline=dict(value=value, unit='chain-steps/s', cores=r['cores'], kind=r['kind'], sample=sample,
                                  seconds=r['seconds'],
                                  implementation='unmodified reference classes (baseline/_ref): Chain_steps.do_step + '
                                                 'update_results, numpy stand-in for bottleneck'
                                  if r['kind'] == 'reference' else 'oracle/crp_oracle.py (CPU restatement)'),
                e2e=dict(value=value, unit='chain-steps/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)


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
        _emit(json.dumps(run_reference(args, cfg)), out_fd)
        return
    out = run_gpu(args, cfg)
    sys.stdout.flush()
    if out is not None:
        _emit(json.dumps(out), out_fd)


if __name__ == '__main__':
    main()
