"""CUDA-backed Dirichlet-process mixture model with the reference's class contract.

`DeviceCRP` / `DeviceCRPLearnErrors` expose exactly what libs/MCMC.py of the
reference calls on a model: `init`,
`update_assignments_Gibbs`, `update_assignments_split_merge`, `update_DP_alpha`,
`update_parameters`, `update_error_rates`, `get_ll_full`, `get_lprior_full` and
the attributes `assignment`, `cells_per_cluster`, `parameters`, `DP_a`, `FN`,
`FP`, `cells_total`, `muts_total`.  All per-cell and per-(cluster, mutation)
arithmetic runs in the sm_100a kernels of libbnpc_b200.so; this file only
sequences launches, keeps the host mirror of the cluster list and does the
O(1)/O(K) scalar algebra of the Metropolis-Hastings ratios.  There is no CPU
fallback: without the library or a CUDA device every method raises.

Method docstrings cite the reference lines (cbg-ethz/BnpC v0.2.1) they replace.
"""
import ctypes as C
import threading
from collections import OrderedDict

import numpy as np
import torch
from scipy.special import gamma as _gamma_fn
from scipy.special import gammaln, xlogy
from scipy.stats import beta as _beta_dist
from scipy.stats import gamma as _gamma_dist
from scipy.stats import truncnorm as _truncnorm

from . import _lib
from .rng import PhiloxRandom

EPS = np.finfo(np.float64).resolution            # libs/CRP.py:11
THETA_LO = 1e-5                                   # libs/CRP.py:12-13
THETA_HI = 1 - THETA_LO
LL_BUDGET_BYTES = 1 << 30                         # largest ll matrix built per epoch
_PACK_LOCK = threading.Lock()                     # chains of one model pack the input once
_NORM_LOGC = np.log(np.sqrt(2 * np.pi))           # scipy _norm_pdf_logC


def _tn_logpdf(x, lo, hi, loc, scale):
    """scipy.stats.truncnorm.logpdf(x, lo, hi, loc, scale) for scalars, without the frozen /
    argument-checking machinery (same private formulas: _norm_logpdf - _log_gauss_mass)."""
    y = (x - loc) / scale
    if not (lo <= y <= hi):
        return -np.inf
    return float(_truncnorm._logpdf(np.float64(y), np.float64(lo), np.float64(hi))) - np.log(scale)


class _TruncNormPrior:
    """Frozen truncated normal on [0,1] (libs/CRP_learning_errors.py:24,30): logpdf with the
    log mass of the interval computed once."""

    def __init__(self, mean, sd):
        self.mean, self.sd = mean, sd
        self.lo, self.hi = (0 - mean) / sd, (1 - mean) / sd
        self.args = (self.lo, self.hi, mean, sd)
        # _logpdf(0) = -log(sqrt(2 pi)) - log(mass of [lo, hi])
        self._const = float(_truncnorm._logpdf(np.float64(0.0), np.float64(self.lo), np.float64(self.hi)))

    def logpdf(self, x):
        y = (x - self.mean) / self.sd
        if not (self.lo <= y <= self.hi):
            return -np.inf
        return (-y ** 2 / 2.0 + self._const) - np.log(self.sd)


class _Shared:
    """Read-only device data shared by all chains of one model on one device:
    the two bit-planes, per-cell popcounts and the log(n) table."""

    def __init__(self, data, device):
        L = _lib.lib()
        N, M = data.shape
        self.W = 4 * ((M + 127) // 128)
        self.x1 = torch.empty((N, self.W), dtype=torch.int32, device=device)
        self.x0 = torch.empty((N, self.W), dtype=torch.int32, device=device)
        self.n1 = torch.empty(N, dtype=torch.int32, device=device)
        self.n0 = torch.empty(N, dtype=torch.int32, device=device)
        sp = torch.cuda.current_stream(device).cuda_stream
        rows = max(1, (256 << 20) // max(M, 1))            # upload in <=256 MB slabs
        for r0 in range(0, N, rows):
            blk = data[r0:r0 + rows]
            code = np.full(blk.shape, -1, dtype=np.int8)
            code[blk == 1] = 1
            code[blk == 0] = 0
            d = torch.as_tensor(code, device=device)
            n = blk.shape[0]
            L.pack_planes(None, d.data_ptr(), n, M, self.W,
                          self.x1[r0:].data_ptr(), self.x0[r0:].data_ptr(),
                          self.n1[r0:].data_ptr(), self.n0[r0:].data_ptr(), sp)
            torch.cuda.current_stream(device).synchronize()
        with np.errstate(divide='ignore'):
            logn = np.log(np.arange(N + 1, dtype=np.float64))
        self.logn = torch.as_tensor(logn, device=device)


class _ThetaView:
    """`model.parameters[ids]` -> float32 numpy rows (libs/MCMC.py:281-282)."""

    def __init__(self, owner):
        self._o = owner

    def __getitem__(self, ids):
        o = self._o
        scalar = np.ndim(ids) == 0
        idx = torch.as_tensor(np.atleast_1d(np.asarray(ids, dtype=np.int64)), device=o.device)
        with torch.cuda.stream(o.stream):
            rows = o._down(o.theta.index_select(0, idx))
        return rows[0] if scalar else rows


class DeviceCRP:
    """Fixed error rates (reference class `CRP`, libs/CRP.py:17-66)."""

    learning = False

    def __init__(self, data, DP_alpha=(-1, -1), param_beta=(1, 1), FN_error=EPS, FP_error=EPS,
                 device=None, rnd=None):
        self.data = data
        self.cells_total, self.muts_total = data.shape
        self.p, self.q = param_beta
        self.param_prior = _beta_dist(self.p, self.q)
        self.beta_prior_uniform = bool(self.p == self.q == 1)
        b0 = _gamma_fn(self.p) * _gamma_fn(self.q + 1) / _gamma_fn(self.p + self.q + 1)
        b1 = _gamma_fn(self.p + 1) * _gamma_fn(self.q) / _gamma_fn(self.p + 1 + self.q)
        self._beta_mix_const = np.array([b0, b1]) / (b0 + b1)
        self.FP = FP_error
        self.FN = FN_error
        if DP_alpha[0] < 0 or DP_alpha[1] < 0:
            self.DP_a_gamma = (np.sqrt(self.cells_total), 1)
        else:
            self.DP_a_gamma = tuple(DP_alpha)
        self.DP_a_prior = _gamma_dist(*self.DP_a_gamma)
        self.DP_a = self.DP_a_prior.mean()
        self.param_proposal_sd = np.array([0.1, 0.25, 0.5])
        self.cells_per_cluster = None
        self.device = device
        self.rnd = rnd
        self._shared = {}            # device -> _Shared, shared between deep copies
        self._dev_ready = False
        self.sweep_stats = {}
        self.profile = False
        self._events = []
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def __str__(self):
        return ('\nDPMM with:\n'
                f'\t{self.cells_total} cells\n\t{self.muts_total} mutations\n'
                f'\tFixed FN rate: {self.FN}\n\tFixed FP rate: {self.FP}\n'
                '\n\tPriors:\n'
                f'\tParams.:\tBeta({self.p},{self.q})\n'
                f'\tCRP a_0:\tGamma({self.DP_a_gamma[0]:.1f},{self.DP_a_gamma[1]})\n')

    def __deepcopy__(self, memo):
        # libs/MCMC.py:128 deep-copies the model once per chain BEFORE init(); the
        # input matrix and its packed device copy are immutable, so share them.
        if self._dev_ready:
            raise RuntimeError('a model cannot be copied after init()')
        new = self.__class__.__new__(self.__class__)
        new.__dict__.update(self.__dict__)
        new.sweep_stats = {}
        new._events = []
        return new

    # ------------------------------------------------------------------ plumbing
    def _sp(self):
        return self.stream.cuda_stream

    def _buf(self, name, numel, dtype):
        t = self._bufs.get(name)
        if t is None or t.numel() < numel or t.dtype != dtype:
            t = torch.empty(max(int(numel), 1), dtype=dtype, device=self.device)
            self._bufs[name] = t
        return t

    def _up(self, arr, dtype):
        t = torch.as_tensor(np.ascontiguousarray(arr), dtype=dtype, device=self.device)
        self.h2d_bytes += t.numel() * t.element_size()
        return t

    def _down(self, t):
        """device -> host read (synchronises this chain's stream)"""
        self.d2h_bytes += t.numel() * t.element_size()
        return t.cpu().numpy()

    def _down_small(self, *tensors):
        """One synchronisation for several small device -> host reads (pinned staging)."""
        out, off = [], 0
        for t in tensors:
            nb = t.numel() * t.element_size()
            if off + nb > self._pin.numel():
                raise RuntimeError('pinned staging buffer too small')
            view = self._pin[off:off + nb].view(t.dtype)
            view.copy_(t.reshape(-1), non_blocking=True)
            out.append(view)
            off += (nb + 15) & ~15
            self.d2h_bytes += nb
        self.stream.synchronize()
        return [v.numpy().copy() for v in out]

    class _Timed:
        """CUDA-event bracket around one launch on the chain's stream (bench.py roofline)."""

        def __init__(self, owner, name):
            self.o, self.name = owner, name

        def __enter__(self):
            if self.o.profile:
                self.a = torch.cuda.Event(enable_timing=True)
                self.b = torch.cuda.Event(enable_timing=True)
                self.a.record(self.o.stream)

        def __exit__(self, *exc):
            if self.o.profile:
                self.b.record(self.o.stream)
                self.o._events.append((self.name, self.a, self.b))
            return False

    def kernel_times_ms(self):
        """name -> list of durations of the bracketed launches since the last call."""
        self.stream.synchronize()
        out = {}
        for name, a, b in self._events:
            out.setdefault(name, []).append(a.elapsed_time(b))
        self._events = []
        return out

    def _setup_device(self):
        if not torch.cuda.is_available():
            raise RuntimeError('bnpc_b200 needs a CUDA device (sm_100a); there is no CPU path')
        self.L = _lib.lib()
        if self.device is None:
            self.device = torch.device('cuda', torch.cuda.current_device())
        self.device = torch.device(self.device)
        torch.cuda.set_device(self.device)
        self.stream = torch.cuda.Stream(self.device)
        if self.rnd is None:
            self.rnd = PhiloxRandom(np.random.SeedSequence().entropy & 0xFFFFFFFFFFFFFFFF)
        self.rnd.bind(self.device)
        with torch.cuda.stream(self.stream):
            key = str(self.device)
            with _PACK_LOCK:
                if key not in self._shared:
                    self._shared[key] = _Shared(self.data, self.device)
            self.sh = self._shared[key]
            self._bufs = {}
            N = self.cells_total
            self.assign_d = torch.zeros(N, dtype=torch.int32, device=self.device)
            self.visit = torch.empty(N * _lib.VISIT_BYTES, dtype=torch.uint8, device=self.device)
            self.cand = torch.empty(N * _lib.CAND_BYTES, dtype=torch.uint8, device=self.device)
            self.visit_c = torch.empty(N * _lib.VISIT_BYTES, dtype=torch.uint8, device=self.device)
            self.cand_c = torch.empty(N * _lib.CAND_BYTES, dtype=torch.uint8, device=self.device)
            self.cblk = torch.empty((N + 127) // 128 + 2, dtype=torch.int32, device=self.device)
            self.members = torch.empty(N, dtype=torch.int32, device=self.device)
            self.cells_d = torch.empty(N + 8, dtype=torch.int32, device=self.device)
            self.half = torch.zeros(N + 8, dtype=torch.int32, device=self.device)
            self.gblk = torch.empty(2 * ((N + 1023) // 1024) + 2, dtype=torch.int32, device=self.device)
            self.seg3 = torch.zeros(8, dtype=torch.int32, device=self.device)
            self.rg_work = torch.zeros(2 * N + 16, dtype=torch.int32, device=self.device)
            self.st = torch.zeros(_lib.ST_WORDS, dtype=torch.int32, device=self.device)
            self.rg_theta = torch.zeros((3, self.muts_total), dtype=torch.float32, device=self.device)
            self.rg_S1 = torch.zeros((3, self.muts_total), dtype=torch.int32, device=self.device)
            self.rg_S0 = torch.zeros((3, self.muts_total), dtype=torch.int32, device=self.device)
            self.rg_dec = torch.zeros(4, dtype=torch.int32, device=self.device)
            self.rg_scal = torch.zeros(32, dtype=torch.float64, device=self.device)
            self._pin = torch.empty(1 << 16, dtype=torch.uint8).pin_memory()
            self.idcap = 0
        self._dev_ready = True
        self._version = 0
        self._stats_version = -1
        self._trace_cache = None

    def _grow_ids(self, need):
        """Make room for cluster ids < need (theta rows, counters, maps)."""
        if need <= self.idcap:
            return
        cap = max(need, 2 * self.idcap, 64)
        M = self.muts_total
        theta = torch.zeros((cap, M), dtype=torch.float32, device=self.device)
        if self.idcap:
            theta[:self.idcap] = self.theta
        self.theta = theta
        self.cnt = torch.zeros(cap, dtype=torch.int32, device=self.device)
        self.lst = torch.zeros(cap, dtype=torch.int32, device=self.device)
        self.col_of_id = torch.full((cap,), -1, dtype=torch.int32, device=self.device)
        self.rank_of_id = torch.zeros(cap, dtype=torch.int32, device=self.device)
        self.live_io = torch.zeros(2 * cap, dtype=torch.int32, device=self.device)
        self.idcap = cap

    def _touch(self):
        self._version += 1
        self._trace_cache = None

    @property
    def assignment(self):
        with torch.cuda.stream(self.stream):
            return self._down(self.assign_d).astype(np.int64)

    def copy_assignment_to(self, row):
        """device-side trace: row (int32 [N] device tensor) <- current assignment."""
        with torch.cuda.stream(self.stream):
            row.copy_(self.assign_d, non_blocking=True)

    @property
    def parameters(self):
        return _ThetaView(self)

    def _ids_sizes(self):
        ids = np.fromiter(self.cells_per_cluster.keys(), dtype=np.int64)
        sizes = np.fromiter(self.cells_per_cluster.values(), dtype=np.int64)
        return ids, sizes

    def get_empty_cluster(self):
        """libs/CRP.py:297-299."""
        i = 0
        while i in self.cells_per_cluster:
            i += 1
        return i

    # --------------------------------------------------------------------- init
    def init(self, mode='random', assign=False):
        """libs/CRP.py:119-152, 155-180."""
        if not self._dev_ready:
            self._setup_device()
        N, M = self.cells_total, self.muts_total
        with torch.cuda.stream(self.stream):
            given = assign is not None and assign is not False and len(assign) > 0
            if given:
                raw = np.array(assign)
            elif mode == 'random':
                raw = self.rnd.init_labels(N)
            else:
                raise TypeError(f'Unsupported Initialization: {mode}')
            labels, sizes = np.unique(raw, return_counts=True)
            a = np.searchsorted(labels, raw)             # relabel 0..K-1 in label order
            K = labels.size
            self.cells_per_cluster = OrderedDict((i, int(sizes[i])) for i in range(K))
            self._grow_ids(K + _lib.MAX_EXTRA + 2)
            self.assign_d.copy_(self._up(a, torch.int32))
            self._touch()
            ids_d = self._up(np.arange(K), torch.int32)
            if given:
                self._refresh_stats()
                tape = self.rnd.beta_rows(K, M)
                self.L.beta_rows(self.S1.data_ptr(), self.S0.data_ptr(), K, M, float(self.p),
                                 float(self.q), tape.data_ptr() if tape is not None else None,
                                 self.rnd.device_seed, self.rnd.next_stream(),
                                 self.theta.data_ptr(), ids_d.data_ptr(), self._sp())
            else:
                u = self.rnd.uniform_rows(K, M)
                self.L.theta_from_uniform(u.data_ptr(), K, M, self.theta.data_ptr(),
                                          ids_d.data_ptr(), self._sp())
            self.stream.synchronize()
        self._touch()

    # ------------------------------------------------------- sufficient statistics
    def _refresh_stats(self):
        """S1/S0 [K][M] for the live clusters in list order (replaces the
        data[cells] gathers of libs/CRP.py:308,360-367)."""
        if self._stats_version == self._version:
            return
        ids, sizes = self._ids_sizes()
        K, M, N = ids.size, self.muts_total, self.cells_total
        seg = np.concatenate(([0], np.cumsum(sizes))).astype(np.int32)
        self.ids_d = self._up(ids, torch.int32)
        seg_d = self._up(seg, torch.int32)
        cur = self._buf('cursor', K, torch.int32)
        self.S1 = self._buf('S1', K * M, torch.int32)
        self.S0 = self._buf('S0', K * M, torch.int32)
        sp = self._sp()
        self.L.set_ranks(self.ids_d.data_ptr(), K, self.rank_of_id.data_ptr(), sp)
        self.L.group_members(self.assign_d.data_ptr(), N, self.rank_of_id.data_ptr(), seg_d.data_ptr(),
                             cur.data_ptr(), K, self.members.data_ptr(), sp)
        self.L.suffstat(self.sh.x1.data_ptr(), self.sh.x0.data_ptr(), self.sh.W, M,
                        self.members.data_ptr(), seg_d.data_ptr(), K, int(sizes.max()),
                        self.S1.data_ptr(), self.S0.data_ptr(), sp)
        self._stats_version = self._version

    # ----------------------------------------------------------------- traces
    def _row_loglik(self, theta, ids_ptr, R, S1, S0, fn, fp, want_prior):
        E = len(fn)
        out = self._buf('rl_out', 4 * R + R, torch.float64)
        tot = self._buf('rl_tot', 8, torch.float64)
        fn_a = (C.c_double * E)(*fn) if E else None
        fp_a = (C.c_double * E)(*fp) if E else None
        pr_ptr = out.data_ptr() + 8 * E * R if want_prior else None
        sp = self._sp()
        self.L.row_loglik(theta.data_ptr(), ids_ptr, R, self.muts_total, S1.data_ptr(), S0.data_ptr(),
                          fn_a, fp_a, E, float(self.p), float(self.q), out.data_ptr(), pr_ptr, sp)
        rows = E + (1 if want_prior else 0)
        self.L.row_sum(out.data_ptr(), rows, R, tot.data_ptr(), sp)
        return self._down(tot[:rows])

    def _trace_scalars(self):
        if self._trace_cache is None:
            with torch.cuda.stream(self.stream):
                self._refresh_stats()
                K = len(self.cells_per_cluster)
                r = self._row_loglik(self.theta, self.ids_d.data_ptr(), K, self.S1, self.S0,
                                     [float(self.FN)], [float(self.FP)], not self.beta_prior_uniform)
            self._trace_cache = (float(r[0]), float(r[1]) if not self.beta_prior_uniform else 0.0)
        return self._trace_cache

    def get_ll_full(self):
        """libs/CRP.py:237-238, from sufficient statistics instead of an [N,M] pass."""
        return self._trace_scalars()[0]

    def _crp_table_at(self, n):
        # entries of init_DP_prior's table (libs/CRP.py:191-194, 83-85)
        return np.log(np.asarray(n, dtype=np.float64)) - np.log(self.cells_total - 1 + self.DP_a)

    def get_lprior_full(self):
        """libs/CRP.py:241-251."""
        _, sizes = self._ids_sizes()
        # scipy gamma(a, loc=b).logpdf: xlogy(a-1, y) - y - gammaln(a) at y = x - loc (scale 1)
        a0, loc = self.DP_a_gamma
        y = self.DP_a - loc
        lp_alpha = float(xlogy(a0 - 1.0, y) - y - gammaln(a0)) if y > 0 else -np.inf
        lp = lp_alpha + np.nansum(self._crp_table_at(sizes))
        if not self.beta_prior_uniform:
            lp += self._trace_scalars()[1]
        return lp

    # ------------------------------------------------------------- Gibbs sweep
    def update_assignments_Gibbs(self):
        """libs/CRP.py:254-299.  The sweep runs in epochs: for the clusters alive
        at the start of an epoch the cells x clusters log-likelihood matrix is
        built by one dense kernel, then one persistent CTA walks the permutation;
        clusters born inside an epoch get their column computed on the spot."""
        N, M = self.cells_total, self.muts_total
        L, sh = self.L, self.sh
        with torch.cuda.stream(self.stream):
            sp = self._sp()
            perm, u, beta_tape, n_tape = self.rnd.gibbs_draws(N, M)
            mix0, mix1 = self._beta_mix_const
            FN, FP = float(self.FN), float(self.FP)
            # popcount form of get_lpost_single_new_cluster (libs/CRP.py:230-234)
            c1 = float(np.log(mix1 * (1 - FN) + mix0 * FP))
            c0 = float(np.log(mix1 * FN + mix0 * (1 - FP)))
            c_norm = float(np.log(N - 1 + self.DP_a))
            lnew_prior = float(np.log(self.DP_a) - np.log(N - 1 + self.DP_a))
            L.gibbs_prepare(perm.data_ptr(), u.data_ptr(), self.assign_d.data_ptr(), sh.n1.data_ptr(),
                            sh.n0.data_ptr(), N, c1, c0, lnew_prior, self.visit.data_ptr(), sp)
            seed, stream_id = self.rnd.device_seed, self.rnd.next_stream()
            t, first, epochs = 0, 1, 0
            stall = 0
            while t < N:
                ids, sizes = self._ids_sizes()
                K = ids.size
                self._grow_ids(K + _lib.MAX_EXTRA + 2)
                live = np.empty(2 * K, dtype=np.int32)
                live[0::2] = ids
                live[1::2] = sizes
                self.live_io[:2 * K].copy_(self._up(live, torch.int32))
                L.gibbs_epoch_begin(self.live_io.data_ptr(), K, self.lst.data_ptr(), self.cnt.data_ptr(),
                                    self.col_of_id.data_ptr(), self.idcap, self.st.data_ptr(), first, sp)
                # odd row stride (bank-conflict-free per-lane row reads in the warp regime)
                ldk = max(3, K | 1)
                rows = int(min(N - t, max(1, LL_BUDGET_BYTES // (8 * ldk))))
                lp = self._buf('lp', 2 * K * M, torch.float64)
                ll = self._buf('ll', rows * ldk + 2, torch.float64)
                lpx = self._buf('lpx', 2 * _lib.MAX_EXTRA * M, torch.float64)
                llx = self._buf('llx', _lib.MAX_EXTRA * rows, torch.float64)
                scratch = self._buf('scratch', self.idcap + 1, torch.float64)
                L.logprob_tables(self.theta.data_ptr(), self.lst.data_ptr(), K, M, FN, FP, lp.data_ptr(), sp)
                # cell indices are read straight out of the visit records
                with self._Timed(self, 'll_matrix'):
                    L.ll_matrix(sh.x1.data_ptr(), sh.x0.data_ptr(), sh.W, M,
                                self.visit.data_ptr() + t * _lib.VISIT_BYTES + _lib.VISIT_CELL_OFFSET,
                                _lib.VISIT_BYTES // 4, rows, lp.data_ptr(), K, ll.data_ptr(), ldk, sp)
                compacted = ldk <= _lib.MAX_LIST
                if compacted:
                    L.gibbs_candidates(ll.data_ptr(), ldk, K, self.col_of_id.data_ptr(),
                                       self.visit.data_ptr() + t * _lib.VISIT_BYTES,
                                       self.cand.data_ptr() + t * _lib.CAND_BYTES, rows,
                                       float(np.log(N)), c_norm, self.cblk.data_ptr(), sp)
                    L.gibbs_compact(self.visit.data_ptr() + t * _lib.VISIT_BYTES,
                                    self.cand.data_ptr() + t * _lib.CAND_BYTES, rows,
                                    self.cblk.data_ptr(), self.visit_c.data_ptr(),
                                    self.cand_c.data_ptr(), self.st.data_ptr(), sp)
                a = _lib.SweepArgs(
                    x1=sh.x1.data_ptr(), x0=sh.x0.data_ptr(), W=sh.W, N=N, M=M,
                    assign=self.assign_d.data_ptr(), cnt=self.cnt.data_ptr(), lst=self.lst.data_ptr(),
                    col_of_id=self.col_of_id.data_ptr(), theta=self.theta.data_ptr(), idcap=self.idcap,
                    st=self.st.data_ptr(), live_out=self.live_io.data_ptr(),
                    ll=ll.data_ptr(), ldk=ldk, t_epoch0=t,
                    lpx=lpx.data_ptr(), llx=llx.data_ptr(), ldx=rows, scratch=scratch.data_ptr(),
                    visit=self.visit.data_ptr(), cand=self.cand.data_ptr(), t_begin=t, t_end=t + rows,
                    visit_c=self.visit_c.data_ptr() if compacted else None,
                    cand_c=self.cand_c.data_ptr() if compacted else None,
                    beta_rows=beta_tape.data_ptr() if beta_tape is not None else None,
                    n_beta_rows=n_tape, seed=seed, stream_id=stream_id,
                    logn=sh.logn.data_ptr(), c_norm=c_norm, FN=FN, FP=FP,
                    p=float(self.p), q=float(self.q))
                with self._Timed(self, 'gibbs_sweep'):
                    L.gibbs_sweep(C.byref(a), 256 if K < 1000 else 1024, sp)
                # status block and live list in one read (at most MAX_EXTRA births per epoch)
                k_cap = K + _lib.MAX_EXTRA + 2
                if 8 * k_cap + 256 <= self._pin.numel():
                    st, pairs = self._down_small(self.st, self.live_io[:2 * k_cap])   # synchronises
                else:
                    st, pairs = self._down(self.st), self._down(self.live_io[:2 * k_cap])
                flags = int(st[_lib.ST_FLAGS])
                if flags & (_lib.STOP_TAPE_EMPTY | _lib.STOP_HANG):
                    raise RuntimeError(f'gibbs_sweep stopped with flags {flags:#x} at t={st[_lib.ST_TDONE]}')
                K = int(st[_lib.ST_K])
                self.cells_per_cluster = OrderedDict(
                    (int(pairs[2 * j]), int(pairs[2 * j + 1])) for j in range(K))
                t_new = int(st[_lib.ST_TDONE])
                stall = stall + 1 if t_new == t else 0
                if stall > 2:
                    raise RuntimeError(f'gibbs_sweep made no progress at t={t} (flags {flags:#x})')
                if flags & _lib.STOP_CAPACITY:
                    self._grow_ids(2 * self.idcap)
                t, first = t_new, 0
                epochs += 1
            if getattr(self.rnd, 'is_tape', False):
                if int(st[_lib.ST_BIRTHS]) != n_tape:
                    raise RuntimeError(f'parity tape held {n_tape} cluster births, sweep made '
                                       f'{int(st[_lib.ST_BIRTHS])}')
            self.sweep_stats = dict(epochs=epochs, births=int(st[_lib.ST_BIRTHS]),
                                    moved=int(st[_lib.ST_MOVED]), slow=int(st[_lib.ST_SLOW]),
                                    uncertain=int(st[_lib.ST_NUNC]),
                                    kcycles=int(st[8]), us=int(st[9]) * 1.024,
                                    sm_mhz=(int(st[8]) / max(1, int(st[9]))) * 1e3)
        self._touch()

    # ----------------------------------------------------------------- MH theta
    def update_parameters(self, step_no=None):
        """libs/CRP.py:302-311, 314-344 for all live clusters in one launch.
        Returns (declined, accepted) summed over clusters x mutations."""
        with torch.cuda.stream(self.stream):
            self._refresh_stats()
            K, M = len(self.cells_per_cluster), self.muts_total
            rnd = self.rnd.mh_theta_draws(K, M)
            dec = self._buf('declined', K, torch.int32)
            dec[:K].zero_()
            self.L.mh_theta(self.theta.data_ptr(), self.ids_d.data_ptr(), K, M, self.S1.data_ptr(),
                            self.S0.data_ptr(), rnd.data_ptr(), float(self.FN), float(self.FP),
                            float(self.p), float(self.q), 0, None, dec.data_ptr(), self._sp())
            declined = int(dec[:K].sum().item())
        self._trace_cache = None
        return declined, K * M - declined

    # ------------------------------------------------------------------ DP alpha
    def update_DP_alpha(self):
        """libs/CRP.py:386-410 (host scalar algebra; the rate is passed as numpy's
        scale, as the reference does)."""
        k = len(self.cells_per_cluster)
        eta = self.rnd.beta(self.DP_a + 1, self.cells_total)
        w = (self.DP_a_gamma[0] + k - 1) / (self.cells_total * (self.DP_a_gamma[1] - np.log(eta)))
        pi_eta = w / (1 + w)
        if self.rnd.random() < pi_eta:
            draw = self.rnd.gamma(self.DP_a_gamma[0] + k, self.DP_a_gamma[1] - np.log(eta))
        else:
            draw = self.rnd.gamma(self.DP_a_gamma[0] + k - 1, self.DP_a_gamma[1] - np.log(eta))
        self.DP_a = max(1 + EPS, draw)
        self._trace_cache = None

    # -------------------------------------------------------------- split-merge
    def update_assignments_split_merge(self, ratios=(.75, .25), step_no=5):
        """libs/CRP.py:417-431.  Returns ([accepted, declined], move)."""
        k = len(self.cells_per_cluster)
        with torch.cuda.stream(self.stream):
            if k == 1:
                return (self._try_split(step_no), 0)
            if k == self.cells_total:
                return (self._try_merge(step_no), 1)
            move = self.rnd.pick_weighted(ratios)
            if move == 0:
                return (self._try_split(step_no), move)
            return (self._try_merge(step_no), move)

    def _try_split(self, scans):
        """libs/CRP.py:434-481."""
        L, sp, N = self.L, self._sp(), self.cells_total
        ids, sizes = self._ids_sizes()
        weight = sizes / sizes.sum()
        while True:
            where = self.rnd.pick_weighted(weight)
            target = int(ids[where])
            n = int(sizes[where])
            if n != 1:
                break
        a_i, a_j = self.rnd.first_two_of_permutation(n)
        L.gather_members(self.assign_d.data_ptr(), N, target, -1, self.cells_d.data_ptr(),
                         self.gblk.data_ptr(), sp)
        L.anchor_swaps(self.cells_d.data_ptr(), n, n, a_i, a_j, 0, sp)
        lq_pick = np.log(weight[where]) - np.log(n) - np.log(n - 1)
        others = np.delete(sizes, where)
        ok, ones = self._restricted_gibbs('split', n, n, (lq_pick, others), scans, target, -1)
        if not ok:
            return [0, 1]
        new_id = self.get_empty_cluster()
        self._grow_ids(new_id + _lib.MAX_EXTRA + 2)
        self.theta[target].copy_(self.rg_theta[0])
        self.theta[new_id].copy_(self.rg_theta[1])
        L.apply_split(self.cells_d.data_ptr(), n, self.half.data_ptr(), new_id,
                      self.assign_d.data_ptr(), sp)
        moved = ones + 1
        self.cells_per_cluster[target] -= moved
        self.cells_per_cluster[new_id] = moved
        self._touch()
        return [1, 0]

    def _try_merge(self, scans):
        """libs/CRP.py:484-524."""
        L, sp, N = self.L, self._sp(), self.cells_total
        ids, sizes = self._ids_sizes()
        inv = 1 / sizes
        weight = inv / inv.sum()
        w_i, w_j = self.rnd.pick_two_weighted(weight)
        cl_i, cl_j = int(ids[w_i]), int(ids[w_j])
        n_a, n_b = int(sizes[w_i]), int(sizes[w_j])
        a_i = self.rnd.randint(n_a)
        a_j = self.rnd.randint(n_b)
        n = n_a + n_b
        L.gather_members(self.assign_d.data_ptr(), N, cl_i, cl_j, self.cells_d.data_ptr(),
                         self.gblk.data_ptr(), sp)
        L.anchor_swaps(self.cells_d.data_ptr(), n, n_a, a_i, a_j, 1, sp)
        both = np.argwhere((ids == cl_j) | (ids == cl_i)).flatten()
        lq_pick = np.nansum(np.log(weight[both])) - np.nansum(np.log(sizes[both]))
        ok, _ = self._restricted_gibbs('merge', n, n_a, lq_pick, scans, cl_i, cl_j)
        if not ok:
            return [0, 1]
        self.theta[cl_i].copy_(self.rg_theta[2])
        L.apply_merge(self.cells_d.data_ptr(), n_a, n, cl_i, self.assign_d.data_ptr(), sp)
        self.cells_per_cluster[cl_i] += n_b
        del self.cells_per_cluster[cl_j]
        self._touch()
        return [1, 0]

    # -- restricted Gibbs machinery (libs/CRP.py:527-638) -----------------------
    # Everything a split-merge move computes on the device is enqueued without waiting; the
    # scalars the Metropolis-Hastings decision needs are reduced into slots of `rg_scal` and
    # read back ONCE (one stream synchronisation per move).
    SL_FWD_ASSIGN, SL_FWD_THETA, SL_BACK, SL_BACK_LQ, SL_PRIOR_NEW, SL_PRIOR_OLD, SL_LL3 = 0, 1, 2, 3, 4, 6, 8

    def _rg_side_stats(self, n):
        """rg_S[0], rg_S[1] = sufficient statistics of the two halves (anchors
        included) for the current `half`; returns nothing (device only)."""
        L, sp, M = self.L, self._sp(), self.muts_total
        L.rg_sides(self.cells_d.data_ptr(), n, self.half.data_ptr(), self.members.data_ptr(),
                   self.seg3.data_ptr(), sp)
        L.suffstat(self.sh.x1.data_ptr(), self.sh.x0.data_ptr(), self.sh.W, M, self.members.data_ptr(),
                   self.seg3.data_ptr(), 2, n, self.rg_S1.data_ptr(), self.rg_S0.data_ptr(), sp)
        self._stats_version = -1          # `members` was reused

    def _rg_sum_to(self, v, rows, m, slot):
        """rg_scal[slot + r] = sum of row r of v[rows][m] (fixed order)."""
        self.L.row_sum(v.data_ptr(), rows, m, self.rg_scal.data_ptr() + 8 * slot, self._sp())

    def _rg_mh(self, row0, rows, slot=None):
        """MH_cluster_params on rg_theta[row0:row0+rows] (libs/CRP.py:583,601); with a slot the
        transition log-probability (trans_prob=True) is summed into rg_scal[slot]."""
        M = self.muts_total
        rnd = self.rnd.mh_theta_draws(rows, M)
        want = slot is not None
        logq = self._buf('rg_logq', 3 * M, torch.float64)
        self.L.mh_theta(self.rg_theta[row0:].data_ptr(), None, rows, M, self.rg_S1[row0:].data_ptr(),
                        self.rg_S0[row0:].data_ptr(), rnd.data_ptr(), float(self.FN), float(self.FP),
                        float(self.p), float(self.q), 1 if want else 0,
                        logq.data_ptr() if want else None, self.rg_dec.data_ptr(), self._sp())
        if want:
            self._rg_sum_to(logq, 1, rows * M, slot)

    def _rg_pair_ll(self, theta, ids_ptr, n):
        """ll of the n-2 free cells under two theta rows (libs/CRP.py:635-638)."""
        L, sp, M, nf = self.L, self._sp(), self.muts_total, n - 2
        lp = self._buf('rg_lp', 4 * M, torch.float64)
        ll2 = self._buf('rg_ll2', 2 * nf, torch.float64)
        L.logprob_tables(theta.data_ptr(), ids_ptr, 2, M, float(self.FN), float(self.FP), lp.data_ptr(), sp)
        L.ll_matrix(self.sh.x1.data_ptr(), self.sh.x0.data_ptr(), self.sh.W, M,
                    self.cells_d.data_ptr() + 4, 1, nf, lp.data_ptr(), 2, ll2.data_ptr(), 2, sp)
        return ll2

    def _scan_split(self, n, want_logq=False):
        """libs/CRP.py:570-578, 590-632."""
        L, sp, nf = self.L, self._sp(), n - 2
        if n > 2:
            ll2 = self._rg_pair_ll(self.rg_theta, None, n)
            perm, u = self.rnd.scan_draws(nf)
            lq = self._buf('rg_lq', nf, torch.float64)
            L.rg_scan(ll2.data_ptr(), 2, n, perm.data_ptr(), u.data_ptr(), self.half.data_ptr(),
                      float(self.DP_a), 0, None, None, -1, lq.data_ptr() if want_logq else None,
                      self.rg_work.data_ptr(), sp)
            if want_logq:
                self._rg_sum_to(lq, 1, nf, self.SL_FWD_ASSIGN)
        self._rg_side_stats(n)
        # both halves in one launch; draws are taken side 0 first, as the reference does
        self._rg_mh(0, 2, self.SL_FWD_THETA if want_logq else None)

    def _scan_merged(self, want_logq=False):
        """libs/CRP.py:581-587."""
        self._rg_mh(2, 1, self.SL_FWD_THETA if want_logq else None)

    def _restricted_gibbs(self, move, n, n_a, size_term, scans, cl_i, cl_j):
        """libs/CRP.py:527-567.  Returns (accepted, number of free cells on side j)."""
        L, sp, M = self.L, self._sp(), self.muts_total
        FN, FP = float(self.FN), float(self.FP)
        mix0 = self._beta_mix_const[0]
        self.rg_scal.zero_()
        # launch state: free cells go to the anchor whose raw row explains them better
        if n > 2:
            k6 = (C.c_double * 6)(
                np.log(1.0 * (1 - FN) + 0.0 * FP), np.log(1.0 * FN + 0.0 * (1 - FP)),
                np.log(0.0 * (1 - FN) + 1.0 * FP), np.log(0.0 * FN + 1.0 * (1 - FP)),
                np.log(mix0 * (1 - FN) + (1 - mix0) * FP), np.log(mix0 * FN + (1 - mix0) * (1 - FP)))
            L.rg_launch_halves(self.sh.x1.data_ptr(), self.sh.x0.data_ptr(), self.sh.W,
                               self.cells_d.data_ptr(), n, k6, self.half.data_ptr(), sp)
        self._rg_side_stats(n)
        tape = self.rnd.beta_rows(2, M)
        L.beta_rows(self.rg_S1.data_ptr(), self.rg_S0.data_ptr(), 2, M, float(self.p), float(self.q),
                    tape.data_ptr() if tape is not None else None, self.rnd.device_seed,
                    self.rnd.next_stream(), self.rg_theta.data_ptr(), None, sp)
        # statistics of all cells of the move never change: S_all = S_i + S_j
        torch.add(self.rg_S1[0], self.rg_S1[1], out=self.rg_S1[2])
        torch.add(self.rg_S0[0], self.rg_S0[1], out=self.rg_S0[2])
        tape = self.rnd.beta_rows(1, M)
        L.beta_rows(self.rg_S1[2].data_ptr(), self.rg_S0[2].data_ptr(), 1, M, float(self.p), float(self.q),
                    tape.data_ptr() if tape is not None else None, self.rnd.device_seed,
                    self.rnd.next_stream(), self.rg_theta[2].data_ptr(), None, sp)
        for _ in range(scans):
            self._scan_split(n)
            self._scan_merged()
        if move == 'split':
            return self._decide_split(n, size_term, cl_i)
        return self._decide_merge(n, n_a, size_term, cl_i, cl_j)

    def _prior_sum_to(self, theta, rows, slot):
        """rg_scal[slot + r] = sum_m Beta(p,q).logpdf(theta[r][m]) (device reduction)."""
        self.L.row_loglik(theta.data_ptr(), None, rows, self.muts_total, self.rg_S1.data_ptr(),
                          self.rg_S0.data_ptr(), None, None, 0, float(self.p), float(self.q),
                          self.rg_scal.data_ptr(), self.rg_scal.data_ptr() + 8 * slot, self._sp())

    def _ll_three_to(self, slot):
        """flat ll of side i, side j (under rg_theta[0:2]) and of all cells (under
        rg_theta[2]) from the three statistics rows (libs/CRP.py:726-728)."""
        fn = (C.c_double * 1)(float(self.FN))
        fp = (C.c_double * 1)(float(self.FP))
        self.L.row_loglik(self.rg_theta.data_ptr(), None, 3, self.muts_total, self.rg_S1.data_ptr(),
                          self.rg_S0.data_ptr(), fn, fp, 1, float(self.p), float(self.q),
                          self.rg_scal.data_ptr() + 8 * slot, None, self._sp())

    def _decide_split(self, n, size_term, cl_i):
        """libs/CRP.py:641-653 with :668-682, :695-733, :757-764."""
        L, sp, M = self.L, self._sp(), self.muts_total
        self._scan_split(n, want_logq=True)
        sd = self.rnd.step_sd_index(1, M)
        A = self._buf('rg_A', 2 * M, torch.float64)
        L.theta_log_ratio(self.theta[cl_i].data_ptr(), self.rg_theta[2].data_ptr(), 1, M,
                          self.rg_S1[2].data_ptr(), self.rg_S0[2].data_ptr(), sd.data_ptr(),
                          THETA_LO, THETA_HI, float(self.FN), float(self.FP),
                          float(self.p), float(self.q), A.data_ptr(), sp)
        self._rg_sum_to(A, 1, M, self.SL_BACK)
        if not self.beta_prior_uniform:
            self._prior_sum_to(self.rg_theta, 2, self.SL_PRIOR_NEW)
            self._prior_sum_to(self.theta[cl_i:cl_i + 1], 1, self.SL_PRIOR_OLD)
        self._ll_three_to(self.SL_LL3)
        sc, seg = self._down_small(self.rg_scal[:16], self.seg3)
        ones = int(seg[3]) if n > 2 else 0                  # free cells on side j (rg_sides)
        logq_ratio = sc[self.SL_BACK] - (sc[self.SL_FWD_ASSIGN] + sc[self.SL_FWD_THETA])
        # eq. 7: prior ratio (libs/CRP.py:695-713)
        n_j = ones + 1
        n_i = n - n_j
        r = np.log(self.DP_a) - gammaln(n)
        if n_i > 0:
            r += gammaln(n_j)
        if n_j > 0:
            r += gammaln(n_i)
        if not self.beta_prior_uniform:
            r += (sc[self.SL_PRIOR_NEW] + sc[self.SL_PRIOR_NEW + 1]) - sc[self.SL_PRIOR_OLD]
        ll_ratio = sc[self.SL_LL3] + sc[self.SL_LL3 + 1] - sc[self.SL_LL3 + 2]
        lq_pick, others = size_term
        norm = np.nansum(1 / np.append(others, [n_i, n_j]))
        size_ratio = (np.log(1 / n_i / norm) + np.log(1 / n_j / norm)) - lq_pick
        total = logq_ratio + r + ll_ratio + size_ratio
        if n > 2 and (ones == 0 or ones == n - 2):
            return False, ones              # np.unique(rg_assignment).size == 1
        return bool(np.log(self.rnd.random()) < total), ones

    def _decide_merge(self, n, n_a, size_term, cl_i, cl_j):
        """libs/CRP.py:656-665 with :685-692, :736-754, :767-820."""
        L, sp, M, nf = self.L, self._sp(), self.muts_total, n - 2
        self._scan_merged(want_logq=True)
        # probability of walking from the launch split back to the original split
        sd = self.rnd.step_sd_index(2, M)
        orig = self._buf('rg_orig', 2 * M, torch.float32)
        orig[:M].copy_(self.theta[cl_i])
        orig[M:2 * M].copy_(self.theta[cl_j])
        A = self._buf('rg_A', 2 * M, torch.float64)
        L.theta_log_ratio(orig.data_ptr(), self.rg_theta.data_ptr(), 2, M, self.rg_S1.data_ptr(),
                          self.rg_S0.data_ptr(), sd.data_ptr(), 0.0, 1.0,
                          float(self.FN), float(self.FP), float(self.p), float(self.q), A.data_ptr(), sp)
        self._rg_sum_to(A, 1, 2 * M, self.SL_BACK)
        if n > 2:
            ll2 = self._rg_pair_ll(orig, None, n)
            lq = self._buf('rg_lq', nf, torch.float64)
            L.rg_scan(ll2.data_ptr(), 2, n, None, None, self.half.data_ptr(), float(self.DP_a), 1,
                      self.cells_d.data_ptr(), self.assign_d.data_ptr(), cl_i, lq.data_ptr(),
                      self.rg_work.data_ptr(), sp)
            self._rg_sum_to(lq, 1, nf, self.SL_BACK_LQ)
        if not self.beta_prior_uniform:
            self._prior_sum_to(self.rg_theta[2:3], 1, self.SL_PRIOR_NEW)
            self._prior_sum_to(orig, 2, self.SL_PRIOR_OLD)
        # `half` now equals the original split (reference quirk)
        self._rg_side_stats(n)
        self._ll_three_to(self.SL_LL3)
        sc, = self._down_small(self.rg_scal[:16])
        logq_ratio = (sc[self.SL_BACK] + sc[self.SL_BACK_LQ]) - sc[self.SL_FWD_THETA]
        n_j = (n - n_a - 1) + 1
        n_i = n - n_j
        r = gammaln(n) - np.log(self.DP_a)
        if n_i > 0:
            r -= gammaln(n_i)
        if n_j > 0:
            r -= gammaln(n_j)
        if not self.beta_prior_uniform:
            r += sc[self.SL_PRIOR_NEW] - (sc[self.SL_PRIOR_OLD] + sc[self.SL_PRIOR_OLD + 1])
        ll_ratio = sc[self.SL_LL3 + 2] - sc[self.SL_LL3] - sc[self.SL_LL3 + 1]
        if nf - 1 > 0:
            back_size = -np.log(self.cells_total) - np.log(nf - 1)
        else:
            back_size = -np.log(self.cells_total)      # the reference's FloatingPointError branch
        size_ratio = back_size - size_term
        total = logq_ratio + r + ll_ratio + size_ratio
        return bool(np.log(self.rnd.random()) < total), 0


class DeviceCRPLearnErrors(DeviceCRP):
    """Learned error rates (reference class `CRP_errors_learning`,
    libs/CRP_learning_errors.py:17-111)."""

    learning = True

    def __init__(self, data, DP_alpha=(-1, -1), param_beta=(1, 1), FP_mean=0.001, FP_sd=0.0005,
                 FN_mean=0.25, FN_sd=0.05, device=None, rnd=None):
        super().__init__(data, DP_alpha, param_beta, FN_mean, FP_mean, device=device, rnd=rnd)
        self.FP_prior = _TruncNormPrior(FP_mean, FP_sd)
        self.FP_sd = np.array([FP_sd * 0.5, FP_sd, FP_sd * 1.5])
        self.FN_prior = _TruncNormPrior(FN_mean, FN_sd)
        self.FN_sd = np.array([FN_sd * 0.5, FN_sd, FN_sd * 1.5])

    def __str__(self):
        return ('\nDPMM with:\n'
                f'\t{self.cells_total} cells\n\t{self.muts_total} mutations\n'
                '\tlearning errors\n\n\tPriors:\n'
                f'\tparams.:\tBeta({self.p},{self.q})\n'
                f'\tCRP a_0:\tGamma({self.DP_a_gamma[0]:.2f},{self.DP_a_gamma[1]})\n'
                f'\tFP:\t\ttrunc norm({self.FP_prior.args[2]},{self.FP_prior.args[3]})\n'
                f'\tFN:\t\ttrunc norm({self.FN_prior.args[2]},{self.FN_prior.args[3]})\n')

    def get_lprior_full(self):
        """libs/CRP_learning_errors.py:47-49."""
        return super().get_lprior_full() + self.FP_prior.logpdf(self.FP) + self.FN_prior.logpdf(self.FN)

    def update_error_rates(self):
        """libs/CRP_learning_errors.py:52-55.  Returns ([acc,dec]_FP, [acc,dec]_FN)."""
        self.FP, fp_count = self._mh_error('FP')
        self.FN, fn_count = self._mh_error('FN')
        self._trace_cache = None
        return fp_count, fn_count

    def _ll_at(self, pairs):
        """full-data log-likelihood at each (FP, FN) pair, from the [K][M] sufficient
        statistics (libs/CRP_learning_errors.py:58-63 without the [N,M] pass)."""
        with torch.cuda.stream(self.stream):
            self._refresh_stats()
            K = len(self.cells_per_cluster)
            r = self._row_loglik(self.theta, self.ids_d.data_ptr(), K, self.S1, self.S0,
                                 [float(fn) for _, fn in pairs], [float(fp) for fp, _ in pairs], False)
        return [float(x) for x in r]

    def _mh_error(self, which):
        """libs/CRP_learning_errors.py:66-111; scalar algebra on the host with the
        same scipy calls as the reference, likelihoods from the device."""
        if which == 'FP':
            cur, prior, steps = self.FP, self.FP_prior, self.FP_sd
        else:
            cur, prior, steps = self.FN, self.FN_prior, self.FN_sd
        sd = steps[self.rnd.randint(3)]
        lo = (0 - cur) / sd
        hi = (1 - cur) / sd
        u = self.rnd.random()
        with np.errstate(divide='raise', invalid='raise', over='ignore', under='ignore'):
            try:
                prop = float(_truncnorm._ppf(np.float64(u), np.float64(lo), np.float64(hi)) * sd + cur)
            except FloatingPointError:
                prop = float(_truncnorm._ppf(np.float64(u), np.float64(lo), np.float64(np.inf)) * sd + cur)
        fwd = _tn_logpdf(prop, lo, hi, cur, sd)
        lo_r, hi_r = (0 - prop) / sd, (1 - prop) / sd
        rev = _tn_logpdf(cur, lo_r, hi_r, prop, sd)
        if which == 'FP':
            ll_new, ll_old = self._ll_at([(prop, self.FN), (cur, self.FN)])
        else:
            ll_new, ll_old = self._ll_at([(self.FP, prop), (self.FP, cur)])
        A = ll_new + prior.logpdf(prop) - ll_old - prior.logpdf(cur) + rev - fwd
        if np.log(self.rnd.random()) < A:
            return prop, [1, 0]
        return cur, [0, 1]
