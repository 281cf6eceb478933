"""Chain driver with the reference's class API (`MCMC`, `Chain`, `Chain_steps`,
`Chain_time`; cbg-ethz/BnpC libs/MCMC.py:26-440), adapted to device-resident chains.

What changed against the reference driver:
  * chains are not forked into worker processes (libs/MCMC.py:113-120) -- a CUDA
    context does not survive fork().  Each chain is one host thread driving its own
    CUDA stream; chain c runs on GPU `c mod G`.  Under torchrun (WORLD_SIZE > 1)
    rank r owns the chains {c : c mod WORLD_SIZE == r} on its LOCAL_RANK device and
    the finished traces are gathered on rank 0 (`gather_results`); there is no
    inter-GPU traffic inside the step loop.
  * every chain draws from its own counter-based stream keyed by the chain seed, so
    a chain's trace does not depend on where it runs.
  * the number of chains is not capped at the CPU count (libs/MCMC.py:100).
The move schedule (`Chain.do_step`, libs/MCMC.py:320-342) and the trace layout
(`Chain.update_results`, libs/MCMC.py:242-282) are the reference's.
"""
import os
import threading
from copy import deepcopy
from datetime import datetime

import numpy as np

try:
    import torch
except ImportError:                                    # pragma: no cover
    torch = None

from bnpc_b200.rng import PhiloxRandom


def _visible_gpus():
    if torch is None or not torch.cuda.is_available():
        return 0
    return torch.cuda.device_count()


def dist_info():
    """(rank, world_size, local_rank) from the torchrun environment (1 process: 0,1,0)."""
    return (int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)),
            int(os.environ.get('LOCAL_RANK', 0)))


def chains_of_rank(n_chains, rank, world):
    """Chain c belongs to rank c mod world."""
    return [c for c in range(n_chains) if c % world == rank]


def psrf_lugsail(ml_traces, burn_in):
    """Lugsail batch-means potential scale reduction factor over the ML traces of several
    chains with cube-root batches (Vats & Flegal 2018).  The lugsail run mode uses the reference's
    own estimator, libs.utils.get_lugsail_batch_means_est (square-root batches); this variant is
    kept for callers that want one common burn-in."""
    x = np.stack([np.asarray(t[burn_in:], dtype=np.float64) for t in ml_traces])
    m, n = x.shape
    b = max(1, int(np.floor(n ** (1 / 3))))
    if n < 3 * b or n // b < 2:
        return np.inf

    def tau(bs):
        a = n // bs
        means = x[:, :a * bs].reshape(m, a, bs).mean(axis=2)
        mu = x.mean(axis=1, keepdims=True)
        return bs * np.sum((means - mu) ** 2, axis=1) / (a - 1)

    t2 = np.mean(2 * tau(b) - tau(max(1, b // 3)))
    s2 = np.mean(np.var(x, axis=1, ddof=1))
    sigma2 = ((n - 1) * s2 + t2) / n
    return float(np.sqrt(sigma2 / s2)) if s2 > 0 else 1.0


class MCMC:
    def __init__(self, model, sm_prob=0.33, dpa_prob=0.5, error_prob=0.1,
                 sm_ratios=(0.75, 0.25), sm_steps=5):
        self.model = model
        self.chains = []
        self.seeds = []
        self.params = {
            'sm_prob': sm_prob, 'dpa_prob': dpa_prob, 'error_prob': error_prob,
            'param_proposal_sd': np.array([0.1, 0.25, 0.5]),
            'sm_ratios': list(sm_ratios), 'sm_steps': sm_steps,
        }

    def __str__(self):
        return ('Move probabilitites:\n'
                '\tSplit/merge:\t{sm_prob}\n\t\tsplit/merge ratio:\t{sm_ratios}\n'
                '\t\tintermediate Gibbs:\t{sm_steps}\n'
                '\tCRP a_0 update:\t{dpa_prob}\n'
                '\tErrors update:\t{error_prob}\n').format(**self.params)

    def get_results(self):
        results = [chain.get_result() for chain in self.chains]
        if not results or 'burn_in' not in results[0]:
            raise RuntimeError('Error in sampling from MCMC')
        return results

    def get_seeds(self):
        return self.seeds

    # ------------------------------------------------------------------ running
    def run(self, run_var, seed, n=1, verbosity=1, assign_file='', debug=False, assign=None):
        """run_var: (steps:int, burn_in:int) | (cutoff:float, 0) | (end:datetime, burn:datetime),
        as produced by the reference's dpmmIO._get_mcmc_termination."""
        cutoff = None
        if isinstance(run_var[0], (int, np.integer)):
            chain_type = Chain_steps
        elif isinstance(run_var[0], float):
            chain_type = Chain_steps
            cutoff = run_var[0]
            run_var = (max(10, int(1 / (cutoff ** 2 - 1))), 0)
            verbosity_ls, verbosity = verbosity, 0
        else:
            chain_type = Chain_time
        # -fa: the loaded assignment is used and never updated (libs/MCMC.py:95-98,132);
        # `assign=` (extension) only chooses the starting state
        self.fix_assign = bool(assign_file)
        if assign_file:
            assign = [int(v) for v in np.loadtxt(assign_file, dtype=int).ravel()]
        # chain seeds as the reference derives them (libs/MCMC.py:102-104)
        if seed > 0:
            np.random.seed(seed)
        self.seeds = np.random.randint(0, 2 ** 32 - 1, n)
        if debug:
            print(f'\nSeed set to: {self.seeds[0]}\n')
            n = 1

        rank, world, local = dist_info()
        if world > 1 and torch is not None and not _dist_ready():
            # launched by torchrun without a process group: NCCL over the GPUs (gloo on CPU-only
            # hosts), rendezvous from the MASTER_ADDR / MASTER_PORT of the environment
            if torch.cuda.is_available():
                torch.cuda.set_device(local)
                torch.distributed.init_process_group('nccl', device_id=torch.device('cuda', local))
            else:
                torch.distributed.init_process_group('gloo')
        mine = chains_of_rank(n, rank, world)
        gpus = max(1, _visible_gpus())
        self.chains = [None] * n
        errors = []

        def work(c):
            try:
                dev = f'cuda:{local}' if world > 1 else f'cuda:{c % gpus}'
                self.chains[c] = self.run_chain(chain_type, run_var, assign, c, verbosity, dev)
            except BaseException as exc:              # surfaced below, never swallowed
                errors.append((c, exc))

        if len(mine) == 1 or debug:
            for c in mine:
                work(c)
        else:
            threads = [threading.Thread(target=work, args=(c,), name=f'chain{c}') for c in mine]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
        if errors:
            raise RuntimeError(f'chain {errors[0][0]} failed: {errors[0][1]!r}') from errors[0][1]
        if cutoff:
            self.run_lugsail_chains(cutoff, mine, verbosity_ls)
        if world > 1:
            self.chains = gather_chains(self.chains, n, rank, world)
        self.chains = [c for c in self.chains if c is not None]

    def run_chain(self, chain_type, run_var, assign, i, verbosity, device=None):
        model = deepcopy(self.model)
        if hasattr(model, 'device'):
            model.device = device
            model.rnd = PhiloxRandom(int(self.seeds[i]))
        model.init(assign=assign)
        chain = chain_type(model, i + 1, *run_var, self.params, verbosity,
                           getattr(self, 'fix_assign', False))
        chain.run()
        return chain

    def run_lugsail_chains(self, cutoff, mine, verbosity, n=200):
        """libs/MCMC.py:138-193: extend all chains by n steps until the PSRF of the ML
        traces undercuts the cutoff."""
        local = [self.chains[c] for c in mine]
        while True:
            steps_run = local[0].results['ML'].size
            traces = all_gather_objects([c.results['ML'] for c in local])
            # the reference's estimator (libs/utils.py:427-461), restated in libs/utils.py
            from libs.utils import get_lugsail_batch_means_est
            psrf = get_lugsail_batch_means_est([(t, steps_run // 2) for part in traces for t in part])
            if verbosity > 1:
                print(f'\tPSRF at {steps_run}:\t{psrf:.5f}')
            for c in local:
                c.results.setdefault('PSRF', []).append((steps_run, psrf))
            if psrf <= cutoff:
                break
            for c in local:
                old = c.get_steps()
                c._extend_results(n, False)
                c.set_steps(n)
                c.run(init_steps=old - 1)
        burn_in = (steps_run // 2) + 1
        for c in local:
            c.results['burn_in'] = burn_in
            c.results['params'] = c.results['params'][burn_in:]
            c.results['PSRF_cutoff'] = cutoff


# ------------------------------------------------------------------------------
# cross-rank plumbing (only at the end of a run / every lugsail round)
# ------------------------------------------------------------------------------
def _dist_ready():
    return torch is not None and torch.distributed.is_available() and torch.distributed.is_initialized()


def all_gather_objects(obj):
    if not _dist_ready():
        return [obj]
    out = [None] * torch.distributed.get_world_size()
    torch.distributed.all_gather_object(out, obj)
    return out


def gather_trace_tensors(results, device):
    """Gather the big per-chain trace arrays of all ranks on rank 0 with tensor collectives
    (NCCL on GPUs, gloo on CPU): assignments [S,N] int32, params [S,Kmax,M] float32 zero-padded
    to the global Kmax (as libs/utils.py:206-223 pads), scalar traces [S] float64.
    `results` is this rank's list of chain result dicts (equal count on every rank).
    Returns the list for all chains in rank-major order on rank 0, None elsewhere."""
    dist = torch.distributed
    world, rank = dist.get_world_size(), dist.get_rank()
    kmax = torch.tensor([max([r['params'].shape[1] for r in results] + [1])], device=device)
    dist.all_reduce(kmax, op=dist.ReduceOp.MAX)
    kmax = int(kmax.item())
    gathered = []
    for r in results:
        S, k, M = r['params'].shape
        par = np.zeros((S, kmax, M), dtype=np.float32)
        par[:, :k] = r['params']
        packs = {
            'assignments': torch.as_tensor(r['assignments'].astype(np.int32), device=device),
            'params': torch.as_tensor(par, device=device),
            'scalars': torch.as_tensor(np.stack([r['ML'], r['MAP'], r['DP_alpha'], r['FN'], r['FP']]),
                                       device=device),
        }
        parts = {}
        for key, t in packs.items():
            bufs = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
            dist.gather(t, bufs, dst=0)
            parts[key] = bufs
        gathered.append(parts)
    if rank != 0:
        return None
    out = []
    for w in range(world):
        for li, r in enumerate(results):
            sc = gathered[li]['scalars'][w].cpu().numpy()
            out.append(dict(assignments=gathered[li]['assignments'][w].cpu().numpy().astype(np.int64),
                            params=gathered[li]['params'][w].cpu().numpy(), ML=sc[0], MAP=sc[1],
                            DP_alpha=sc[2], FN=sc[3], FP=sc[4], burn_in=r['burn_in']))
    return out


def gather_chains(chains, n, rank, world):
    """End-of-run gather replacing the pickle-through-pipe of libs/MCMC.py:114-118.  Falls
    back to object gather when ranks hold different chain counts."""
    if not _dist_ready():
        return chains
    mine = [c for c in chains if c is not None]
    counts = all_gather_objects(len(mine))
    if len(set(counts)) == 1 and mine and 'params' in mine[0].results:
        dev = mine[0].model.device if hasattr(mine[0].model, 'device') else 'cpu'
        res = gather_trace_tensors([c.results for c in mine], dev)
        if rank != 0:
            return []
        return [_ResultOnly(r) for r in res]
    parts = all_gather_objects([c.results for c in mine])
    if rank != 0:
        return []
    return [_ResultOnly(r) for part in parts for r in part]


class _ResultOnly:
    """A finished chain gathered from another rank: traces without the device model."""

    def __init__(self, results):
        self.results = results

    def get_result(self):
        return self.results


# ------------------------------------------------------------------------------
# chains
# ------------------------------------------------------------------------------
class Chain:
    def __init__(self, model, mcmc, no, verbosity=1, fix_assign=False):
        self.model = model
        self.mcmc = mcmc
        self.no = no
        self.learning_errors = model.__module__ == 'libs.CRP_learning_errors' \
            or getattr(model, 'learning', False)
        self.results = {}
        self.MH_counter = np.zeros((5, 2))
        self.verbosity = verbosity
        self.fix_assign = fix_assign

    def __str__(self):
        return f'Chain: {self.no:0>2d}'

    def get_result(self):
        return self.results

    def init_results(self, steps):
        """libs/MCMC.py:231-239.  The assignment trace is int32 (the reference's `int` is 64-bit:
        twice the host memory, 4 GB per chain at 100k cells x 5000 steps) and is touched once here so
        that recording a step never page-faults."""
        n = self.model.cells_total
        self.results = dict(ML=np.zeros(steps), MAP=np.zeros(steps), DP_alpha=np.zeros(steps),
                            FN=np.empty(steps), FP=np.empty(steps),
                            assignments=np.empty((steps, n), dtype=np.int32))
        self.results['assignments'].fill(0)
        self._par_buf = None          # [kept steps, cluster capacity, M]; results['params'] views it
        self._par_k = 0

    def _params_row(self, step, room, k):
        """Row of the theta trace for this step with room for k clusters.  The reference pads the
        whole [steps, K, M] array by one cluster whenever K grows (libs/MCMC.py:271-276); here the
        cluster axis has spare capacity and results['params'] is a view of the columns in use."""
        r = self.results
        if self._par_buf is None:
            cap = max(8, 2 * k)
            self._par_buf = np.zeros((room, cap, self.model.muts_total), dtype=np.float32)
        if k > self._par_buf.shape[1]:
            grown = np.zeros((self._par_buf.shape[0], max(k, 2 * self._par_buf.shape[1]),
                              self.model.muts_total), dtype=np.float32)
            grown[:, :self._par_buf.shape[1]] = self._par_buf
            self._par_buf = grown
        if k > self._par_k or 'params' not in r or r['params'].base is not self._par_buf:
            self._par_k = max(self._par_k, k)
            r['params'] = self._par_buf[:, :self._par_k]
        first_kept = r['ML'].size - self._par_buf.shape[0]
        return self._par_buf[step - first_kept]

    def update_results(self, step, burn_in=True):
        """libs/MCMC.py:242-282: one trace row per step; theta rows of the SORTED live
        cluster ids are kept after burn-in."""
        r = self.results
        room = r['ML'].size - step
        if room == 0:
            self._extend_results(burn_in=burn_in)
        ll = self.model.get_ll_full()
        r['ML'][step] = ll
        r['MAP'][step] = ll + self.model.get_lprior_full()
        r['DP_alpha'][step] = self.model.DP_a
        r['FN'][step] = self.model.FN
        r['FP'][step] = self.model.FP
        if hasattr(self.model, 'assignment_into'):
            self.model.assignment_into(r['assignments'][step])
        else:
            r['assignments'][step] = self.model.assignment
        if burn_in:
            return
        clusters = np.sort(np.fromiter(self.model.cells_per_cluster.keys(), dtype=int))
        row = self._params_row(step, room, clusters.size)
        if hasattr(self.model, 'parameters_into'):
            self.model.parameters_into(clusters, row)
        else:
            row[:clusters.size] = self.model.parameters[clusters]

    def _extend_results(self, add_size=None, burn_in=True):
        r = self.results
        if not add_size:
            add_size = min(200, r['ML'].size)
        if not burn_in and self._par_buf is not None:
            self._par_buf = np.concatenate(
                [self._par_buf, np.zeros((add_size,) + self._par_buf.shape[1:], dtype=np.float32)])
            r['params'] = self._par_buf[:, :self._par_k]
        for key in ('ML', 'MAP', 'DP_alpha', 'FN', 'FP'):
            r[key] = np.append(r[key], np.zeros(add_size))
        r['assignments'] = np.concatenate(
            [r['assignments'], np.zeros((add_size, self.model.cells_total), dtype=np.int32)])

    def stdout_progress(self):
        def show(counter, name, tabs=2):
            total = counter.sum()
            ratio = counter[0] / total if total else np.nan
            print('{}{}:\t{:.2f}'.format('\t' * tabs, name, ratio))
        show(self.MH_counter[0], 'parameters', 1)
        if not self.fix_assign:
            show(self.MH_counter[1], 'splits')
            show(self.MH_counter[2], 'merges')
        if self.learning_errors:
            show(self.MH_counter[3], 'FP')
            show(self.MH_counter[4], 'FN')
        self.MH_counter = np.zeros((5, 2))

    def do_step(self):
        """libs/MCMC.py:320-342.  Move-selection uniforms come from the chain's own stream."""
        rnd, mc, model = self.model.rnd, self.mcmc, self.model
        if not self.fix_assign:
            if rnd.random() < mc['sm_prob']:
                res, move = model.update_assignments_split_merge(mc['sm_ratios'], mc['sm_steps'])
                self.MH_counter[1 if move == 0 else 2] += res
            else:
                model.update_assignments_Gibbs()
            if rnd.random() < mc['dpa_prob']:
                model.update_DP_alpha()
        declined, accepted = model.update_parameters()
        self.MH_counter[0] += (accepted, declined)
        if self.learning_errors and rnd.random() < mc['error_prob']:
            fp, fn = model.update_error_rates()
            self.MH_counter[3] += fp
            self.MH_counter[4] += fn


class Chain_steps(Chain):
    def __init__(self, model, no, steps, burn_in, mcmc, verbosity=1, fix_assign=False):
        super().__init__(model, mcmc, no, verbosity, fix_assign)
        self.steps = steps + 1
        self.burn_in = burn_in
        self.init_results(steps + 1)
        self.update_results(0, burn_in != 0)

    def set_steps(self, n):
        self.steps = n + 1

    def get_steps(self):
        return self.results['ML'].size

    def stdout_progress(self, step_no, total):
        print(f'\t{self}\tstep:\t{step_no: >3} / {total - 1}\n\t\tmean MH accept. ratio:')
        super().stdout_progress()

    def run(self, init_steps=0):
        every = max(1, self.steps // 10)
        for step in range(1, self.steps):
            if self.verbosity > 1 and step % every == 0:
                self.stdout_progress(step + init_steps, self.steps + init_steps)
            self.do_step()
            try:
                burn_in = step < self.burn_in
            except TypeError:
                burn_in = False
            self.update_results(step + init_steps, burn_in)
        self.results['burn_in'] = self.burn_in


class Chain_time(Chain):
    def __init__(self, model, no, end_time, burn_in, mcmc, verbosity=1, fix_assign=False):
        super().__init__(model, mcmc, no, verbosity, fix_assign)
        self.end_time = end_time
        self.burn_in = burn_in
        self.init_results(500)
        self.update_results(0)

    def stdout_progress(self, step_no, total):
        print(f'\t{self}\tstep:\t{step_no: >3}\t(remaining: {total:.1f} mins.)\n'
              '\t\tmean MH accept. ratio:')
        super().stdout_progress()

    def run(self):
        step = 0
        while True:
            now = datetime.now()
            if now > self.end_time:
                break
            if self.verbosity > 1 and step % 1000 == 0:
                self.stdout_progress(step, (self.end_time - now).seconds / 60)
            step += 1
            self.do_step()
            try:
                burn_in = now < self.burn_in
            except TypeError:
                burn_in = False
            self.update_results(step, burn_in)
        used = step + 1
        if 'params' in self.results:
            first_kept = self.results['ML'].size - self.results['params'].shape[0]
            self.results['params'] = self.results['params'][:max(0, used - first_kept)]
        for key in list(self.results):
            if key == 'params':
                continue
            self.results[key] = self.results[key][:used]
        n_par = self.results['params'].shape[0] if 'params' in self.results else 0
        self.results['burn_in'] = self.results['ML'].size - n_par
