"""Chain driver with the reference's class API (`MCMC`, `Chain`, `Chain_steps`,
`Chain_time`; cbg-ethz/BnpC libs/MCMC.py:26-440), adapted to device-resident chains.

What changed against the reference driver:
  * chains are not forked into worker processes (libs/MCMC.py:113-120) -- a CUDA
    context does not survive fork().  The chains that share a GPU form one
    `bnpc_b200.group.ChainGroup`: ONE host thread steps them in lockstep through the
    library's native driver (every kernel launched once for all chains, one stream
    synchronisation per phase for the whole group); chain c runs on GPU `c mod G`.
    Under torchrun (WORLD_SIZE > 1) rank r owns the chains {c : c mod WORLD_SIZE == r}
    on its LOCAL_RANK device and the finished traces are gathered on rank 0
    (`gather_chains`); there is no inter-GPU traffic inside the step loop.
    `Chain.do_step` / `Chain.update_results` remain the per-chain public API (the
    per-method Python mirror of the model, used by the parity tests); `Chain.run`,
    `MCMC.run` and `run_chains` take the native driver whenever the models allow.
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
from bnpc_b200.group import ChainGroup, native_ok, pinned_zeros


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
            from libs.dpmmIO import load_txt
            assign = load_txt(assign_file)
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
        # the chains of one device are initialised one after another (the packed matrix is built
        # once per device) and then stepped together by one host thread per device
        by_device = {}
        for c in mine:
            dev = f'cuda:{local}' if world > 1 else f'cuda:{c % gpus}'
            self.chains[c] = self.make_chain(chain_type, run_var, assign, c, verbosity, dev)
            by_device.setdefault(dev, []).append(self.chains[c])
        errors = []

        def work(group):
            try:
                run_chains(group)
            except BaseException as exc:              # surfaced below, never swallowed
                errors.append(exc)

        groups = list(by_device.values())
        if len(groups) <= 1:
            for group in groups:
                work(group)
        else:
            threads = [threading.Thread(target=work, args=(group,), name=f'gpu{i}')
                       for i, group in enumerate(groups)]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
        if errors:
            raise RuntimeError(f'a chain group failed: {errors[0]!r}') from errors[0]
        if cutoff:
            self.run_lugsail_chains(cutoff, mine, verbosity_ls)
        if world > 1:
            self.chains = gather_chains(self.chains, n, rank, world)
        self.chains = [c for c in self.chains if c is not None]

    def make_chain(self, chain_type, run_var, assign, i, verbosity, device=None):
        """libs/MCMC.py:123-135 without the run: copy of the model, its own random stream keyed
        by the chain seed, initial state, trace row 0."""
        model = deepcopy(self.model)
        if hasattr(model, 'device'):
            model.device = device
            model.rnd = PhiloxRandom(int(self.seeds[i]))
        model.init(assign=assign)
        return chain_type(model, i + 1, *run_var, self.params, verbosity, getattr(self, 'fix_assign', False))

    def run_chain(self, chain_type, run_var, assign, i, verbosity, device=None):
        chain = self.make_chain(chain_type, run_var, assign, i, verbosity, device)
        chain.run()
        return chain

    def run_lugsail_chains(self, cutoff, mine, verbosity, n=200):
        """libs/MCMC.py:138-193: extend all chains by n steps until the PSRF of the ML
        traces undercuts the cutoff.  Every rank takes part in every round (also one that owns
        no chain: fewer chains than ranks)."""
        from libs.utils import get_lugsail_batch_means_est
        local = [self.chains[c] for c in mine]
        while True:
            traces = all_gather_objects([c.results['ML'] for c in local])
            flat = [t for part in traces for t in part]
            steps_run = flat[0].size
            # the reference's estimator (libs/utils.py:427-461), restated in libs/utils.py
            psrf = get_lugsail_batch_means_est([(t, steps_run // 2) for t in flat])
            if verbosity > 1:
                print(f'\tPSRF at {steps_run}:\t{psrf:.5f}')
            for c in local:
                c.results.setdefault('PSRF', []).append((steps_run, psrf))
            if psrf <= cutoff:
                break
            olds = [c.get_steps() for c in local]
            for c in local:
                c._extend_results(n, False)
                c.set_steps(n)
            if local:
                run_chains(local, init_steps=olds[0] - 1)
        burn_in = (steps_run // 2) + 1
        for c in local:
            c.results['burn_in'] = burn_in
            c.results['params'] = c.results['params'][burn_in:]
            c.results['PSRF_cutoff'] = cutoff


# chains per lockstep group: the chains of one device are stepped in groups of at most this many,
# one host thread per group (env BNPC_GROUP_SIZE; measured on B200 at 100k x 1k)
GROUP_SIZE = int(os.environ.get('BNPC_GROUP_SIZE', 4))


def run_chains(chains, init_steps=0):
    """Run chains that share a device to completion: in lockstep through the native group driver
    when their models allow it (CUDA model, counter-based random streams), else one after another
    through the per-method Python mirror (parity tapes, foreign models)."""
    if not chains:
        return
    native = torch is not None and all(native_ok(ch.model) for ch in chains)
    if not native:
        for ch in chains:
            ch.run_python(init_steps) if isinstance(ch, Chain_steps) else ch.run_python()
        return
    if len(chains) > GROUP_SIZE:
        # several lockstep groups side by side on the device, one host thread each
        parts = [chains[i:i + GROUP_SIZE] for i in range(0, len(chains), GROUP_SIZE)]
        errors = []

        def work(part):
            try:
                run_chains(part, init_steps)
            except BaseException as exc:              # noqa: BLE001  (surfaced below)
                errors.append(exc)
        threads = [threading.Thread(target=work, args=(part,)) for part in parts]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return
    group = ChainGroup(chains, chains[0].mcmc, chains[0].fix_assign)
    try:
        if isinstance(chains[0], Chain_steps):
            _run_steps_native(group, chains, init_steps)
        else:
            _run_time_native(group, chains)
    finally:
        group.close()


def _run_steps_native(group, chains, init_steps):
    """Chain_steps.run (libs/MCMC.py:375-388) for a group: steps 1 .. steps-1, trace rows
    init_steps + step; progress every tenth of the run."""
    lead = chains[0]
    steps = lead.steps
    for ch in chains:
        if ch.steps != steps:
            raise RuntimeError('the chains of a group run the same number of steps')
        ch._prepare_params(init_steps)
    every = max(1, steps // 10) if lead.verbosity > 1 else steps
    step = 1
    while step < steps:
        stop = min(steps, (step // every + 1) * every)
        if lead.verbosity > 1 and step % every == 0:
            for ch in chains:
                ch.stdout_progress(step + init_steps, steps + init_steps)
        group.run(step + init_steps, stop - step)
        step = stop
    for ch in chains:
        ch.results['burn_in'] = ch.burn_in


def _run_time_native(group, chains):
    """Chain_time.run (libs/MCMC.py:423-440) for a group: step until the wall clock passes
    end_time; theta rows are kept once it has passed burn_in."""
    lead = chains[0]
    step = 0
    while True:
        now = datetime.now()
        if now > lead.end_time:
            break
        if lead.verbosity > 1 and step % 1000 == 0:
            for ch in chains:
                ch.stdout_progress(step, (lead.end_time - now).seconds / 60)
        step += 1
        try:
            burn_in = now < lead.burn_in
        except TypeError:
            burn_in = False
        for ch in chains:
            if ch.results['ML'].size - step == 0:
                ch._extend_results(burn_in=burn_in)
            if not burn_in and ch._par_buf is None:
                ch._par_alloc(step, len(ch.model.cells_per_cluster))
        group.run(step, 1)
    for ch in chains:
        ch._trim(step + 1)


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


_SCALAR_KEYS = ('ML', 'MAP', 'DP_alpha', 'FN', 'FP')


def gather_trace_tensors(results, device, chain_ids=None):
    """Gather the per-chain trace arrays of all ranks on rank 0 with tensor collectives (NCCL on
    GPUs, gloo on CPU): assignments [S,N] int32, params [Sp,Kmax,M] float32 zero-padded to the
    global Kmax (as libs/utils.py:206-223 pads), scalar traces [S] float64.  Chains may differ in
    length (run-time mode, libs/MCMC.py:395-440): every chain's (S, Sp, burn_in, id) header is
    all-gathered first and the buffers are padded to the longest chain, then trimmed.
    `results` is this rank's list of chain result dicts (equal count on every rank).
    Returns [(chain id, result dict)] for all chains on rank 0, None elsewhere."""
    dist = torch.distributed
    world, rank = dist.get_world_size(), dist.get_rank()
    ids = list(chain_ids) if chain_ids is not None else [rank * len(results) + i for i in range(len(results))]
    head = torch.tensor([[r['ML'].size, r['params'].shape[0], r['params'].shape[1], int(r['burn_in']), ids[i]]
                         for i, r in enumerate(results)], dtype=torch.int64, device=device)
    heads = [torch.empty_like(head) for _ in range(world)]
    dist.all_gather(heads, head)
    heads = torch.stack(heads).cpu().numpy()                    # [world][local][5]
    s_max, sp_max, kmax = (int(heads[..., j].max()) for j in range(3))
    kmax = max(kmax, 1)
    gathered = []
    for r in results:
        S, (Sp, k, M) = r['ML'].size, r['params'].shape
        N = r['assignments'].shape[1]
        par = np.zeros((sp_max, kmax, M), dtype=np.float32)
        par[:Sp, :k] = r['params']
        asg = np.zeros((s_max, N), dtype=np.int32)
        asg[:S] = r['assignments']
        sc = np.zeros((len(_SCALAR_KEYS), s_max))
        for j, key in enumerate(_SCALAR_KEYS):
            sc[j, :S] = r[key]
        packs = {'assignments': torch.as_tensor(asg, device=device), 'params': torch.as_tensor(par, device=device),
                 'scalars': torch.as_tensor(sc, device=device)}
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
        for li in range(len(results)):
            S, Sp, _, burn_in, cid = (int(v) for v in heads[w, li])
            sc = gathered[li]['scalars'][w].cpu().numpy()
            res = dict(assignments=gathered[li]['assignments'][w].cpu().numpy()[:S],
                       params=gathered[li]['params'][w].cpu().numpy()[:Sp], burn_in=burn_in)
            for j, key in enumerate(_SCALAR_KEYS):
                res[key] = sc[j, :S]
            out.append((cid, res))
    return out


def gather_chains(chains, n, rank, world):
    """End-of-run gather replacing the pickle-through-pipe of libs/MCMC.py:114-118; the chains
    come back in chain (= seed) order.  Scalar extras of a result (PSRF values of the lugsail
    mode) travel as objects; falls back to an object gather when ranks hold different chain
    counts."""
    if not _dist_ready():
        return chains
    mine = [(c, ch) for c, ch in enumerate(chains) if ch is not None]
    counts = all_gather_objects(len(mine))
    extras = all_gather_objects([(c, {k: v for k, v in ch.results.items()
                                      if k not in _SCALAR_KEYS + ('assignments', 'params', 'burn_in')})
                                 for c, ch in mine])
    if len(set(counts)) == 1 and mine and 'params' in mine[0][1].results:
        dev = mine[0][1].model.device if hasattr(mine[0][1].model, 'device') else 'cpu'
        res = gather_trace_tensors([ch.results for _, ch in mine], dev, [c for c, _ in mine])
        if rank != 0:
            return []
    else:
        parts = all_gather_objects([(c, ch.results) for c, ch in mine])
        if rank != 0:
            return []
        res = [item for part in parts for item in part]
    extra_of = {c: e for part in extras for c, e in part}
    out = []
    for c, r in sorted(res, key=lambda item: item[0]):
        r.update(extra_of.get(c, {}))
        out.append(_ResultOnly(r))
    return out


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
        twice the host memory, 4 GB per chain at 100k cells x 5000 steps) in page-locked memory:
        the native driver's copy stream writes the rows while the next step runs."""
        n = self.model.cells_total
        self._owners = {}
        asg, self._owners['assignments'] = pinned_zeros((steps, n), np.int32)
        self.results = dict(ML=np.zeros(steps), MAP=np.zeros(steps), DP_alpha=np.zeros(steps),
                            FN=np.zeros(steps), FP=np.zeros(steps), assignments=asg)
        self._par_buf = None          # [kept steps, cluster capacity, M]; results['params'] views it
        self._par_first = 0           # trace row of _par_buf[0]
        self._par_k = 0

    # -- theta trace: the reference pads the whole [steps, K, M] array by one cluster whenever K
    # grows (libs/MCMC.py:271-276); here the cluster axis has spare capacity and
    # results['params'] is a view of the columns in use
    def _par_alloc(self, first_step, k):
        rows = self.results['ML'].size - first_step
        self._par_first = first_step
        self._par_buf, self._owners['params'] = pinned_zeros((rows, max(8, 2 * k), self.model.muts_total),
                                                             np.float32)

    def _par_grow(self, k):
        old = self._par_buf
        if k <= old.shape[1]:
            return
        self._par_buf, self._owners['params'] = pinned_zeros((old.shape[0], max(k, 2 * old.shape[1]),
                                                              old.shape[2]), np.float32)
        self._par_buf[:, :old.shape[1]] = old

    def _params_row(self, step, k):
        r = self.results
        if self._par_buf is None:
            self._par_alloc(step, k)
        self._par_grow(k)
        if k > self._par_k or 'params' not in r or r['params'].base is not self._par_buf:
            self._par_k = max(self._par_k, k)
            r['params'] = self._par_buf[:, :self._par_k]
        return self._par_buf[step - self._par_first]

    def update_results(self, step, burn_in=True):
        """libs/MCMC.py:242-282: one trace row per step; theta rows of the SORTED live
        cluster ids are kept after burn-in."""
        r = self.results
        if r['ML'].size - step == 0:
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
        row = self._params_row(step, clusters.size)
        if hasattr(self.model, 'parameters_into'):
            self.model.parameters_into(clusters, row)
        else:
            row[:clusters.size] = self.model.parameters[clusters]

    def _extend_results(self, add_size=None, burn_in=True):
        r = self.results
        if not add_size:
            add_size = min(200, r['ML'].size)
        if not burn_in and self._par_buf is not None:
            old = self._par_buf
            self._par_buf, self._owners['params'] = pinned_zeros((old.shape[0] + add_size,) + old.shape[1:],
                                                                 np.float32)
            self._par_buf[:old.shape[0]] = old
            r['params'] = self._par_buf[:, :self._par_k]
        for key in ('ML', 'MAP', 'DP_alpha', 'FN', 'FP'):
            r[key] = np.append(r[key], np.zeros(add_size))
        old = r['assignments']
        r['assignments'], self._owners['assignments'] = pinned_zeros((old.shape[0] + add_size, old.shape[1]),
                                                                     np.int32)
        r['assignments'][:old.shape[0]] = old

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

    def run(self, *args):
        """libs/MCMC.py:375-388 / 423-440: this chain alone (a group of one)."""
        run_chains([self], *args)


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

    def _is_burn_in(self, step):
        try:
            return step < self.burn_in
        except TypeError:
            return False

    def _prepare_params(self, init_steps):
        """theta trace of the rows this run will keep (native driver: the buffer exists before
        the run; its first row is the first step past burn-in)"""
        first = next((s for s in range(1, self.steps) if not self._is_burn_in(s)), None)
        if first is not None and self._par_buf is None:
            self._par_alloc(first + init_steps, len(self.model.cells_per_cluster))

    def run_python(self, init_steps=0):
        """the reference's loop over the per-method API (do_step + update_results)"""
        every = max(1, self.steps // 10)
        for step in range(1, self.steps):
            if self.verbosity > 1 and step % every == 0:
                self.stdout_progress(step + init_steps, self.steps + init_steps)
            self.do_step()
            self.update_results(step + init_steps, self._is_burn_in(step))
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

    def _trim(self, used):
        """keep the rows that were written (libs/MCMC.py:434-440)"""
        if self._par_buf is not None:
            kept = max(0, used - self._par_first)
            self._par_buf = self._par_buf[:kept]
            self.results['params'] = self._par_buf[:, :self._par_k]
        for key in list(self.results):
            if key == 'params':
                continue
            self.results[key] = self.results[key][:used]
        n_par = self.results['params'].shape[0] if 'params' in self.results else 0
        self.results['burn_in'] = self.results['ML'].size - n_par

    def run_python(self):
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
        self._trim(step + 1)
