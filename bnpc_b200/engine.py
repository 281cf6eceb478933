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
from scipy.special import log1p as _sc_log1p
from scipy.special import log_ndtr as _sc_log_ndtr
from scipy.special import ndtr as _sc_ndtr
from scipy.special import ndtri_exp as _sc_ndtri_exp
from scipy.stats._continuous_distns import _log_gauss_mass as _scipy_log_gauss_mass
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


def _log_gauss_mass(lo, hi):
    """scipy.stats._continuous_distns._log_gauss_mass for scalars.  The interval of an error-rate
    move straddles 0 (scipy's central case: log1p(-ndtr(a) - ndtr(-b))), evaluated here with the
    same scalar special functions; the tail cases go through scipy itself."""
    if lo <= 0 < hi:
        return _sc_log1p(-_sc_ndtr(np.float64(lo)) - _sc_ndtr(np.float64(-hi)))
    return np.float64(_scipy_log_gauss_mass(np.float64(lo), np.float64(hi)))


def _tn_ppf(q, lo, hi):
    """scipy.stats.truncnorm._ppf(q, lo, hi) for scalars (ppf_left, the case lo < 0 of an
    error-rate move): ndtri_exp(logsumexp([log_ndtr(lo), log(q) + log_gauss_mass])) with scipy's
    two-element logsumexp written out (log1p(exp(min - max)) + log(1) + max)."""
    if not lo < 0:
        return np.float64(_truncnorm._ppf(np.float64(q), np.float64(lo), np.float64(hi)))
    t1 = _sc_log_ndtr(np.float64(lo))
    t2 = np.log(np.float64(q)) + _log_gauss_mass(lo, hi)
    if t1 == t2:
        lphi = np.log1p(np.float64(0.0)) + np.log(np.float64(2.0)) + t1
    else:
        top, low = (t1, t2) if t1 > t2 else (t2, t1)
        lphi = np.log1p(np.exp(low - top)) + np.float64(0.0) + top
    return _sc_ndtri_exp(lphi)


def _tn_logpdf(x, lo, hi, loc, scale):
    """scipy.stats.truncnorm.logpdf(x, lo, hi, loc, scale) for scalars, without the frozen /
    argument-checking machinery (same private formulas: _norm_logpdf - _log_gauss_mass)."""
    y = (x - loc) / scale
    if not (lo <= y <= hi):
        return -np.inf
    y = np.float64(y)
    return float((-y ** 2 / 2.0 - _NORM_LOGC) - _log_gauss_mass(lo, hi)) - np.log(scale)


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

    def _fetch(self, ids):
        o = self._o
        idx = np.atleast_1d(np.asarray(ids, dtype=np.int32))
        n, M = idx.size, o.muts_total
        pin = o._pinned('theta_rows', n * M, torch.float32)
        o.h_in[:n] = idx
        o.L.chain_theta_rows(o.ws, n, pin.data_ptr(), o._sp())
        o._sync()
        o.h2d_bytes += 4 * n
        o.d2h_bytes += 4 * n * M
        return pin[:n * M].numpy().reshape(n, M)

    def __getitem__(self, ids):
        rows = self._fetch(ids).copy()
        return rows[0] if np.ndim(ids) == 0 else rows


class DeviceCRP:
    """Fixed error rates (reference class `CRP`, libs/CRP.py:17-66)."""

    learning = False
    lean_enabled = True           # class-wide switch (tests force the dense FP64 matrix with False)
    lean_rows = 3                 # approximate rows of lean epochs: 3 tcgen05 integer digits, 2 tcgen05 bf16-split, 1 FP32 FMA
    serial_sweep = False          # lean epochs: one sequencer warp (True) or one per component group
    wide_enabled = True           # many-rival data: dense epochs of <= 63 clusters walk option weights (sweep_wide)
    force_wide = False            # tests: take the wide route whenever K <= 63

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
        self._event_meta = {}
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
        new._event_meta = {}
        return new

    # ------------------------------------------------------------------ plumbing
    def _sp(self):
        # the CUDA "current device" is per host thread and kernel launches go to it: a chain that is
        # stepped from another thread than the one that initialised it (bench legs, thread pools)
        # must select its device there first (new threads start on device 0)
        ident = threading.get_ident()
        if ident != self._dev_thread:
            torch.cuda.set_device(self.device)
            self._dev_thread = ident
        return self._stream_ptr

    def _dev(self, name, shape, dtype, zero=False):
        """(Re)allocate the named workspace buffer and publish its address in the C workspace."""
        with torch.cuda.stream(self.stream):      # the fill of torch.zeros runs on the chain's stream
            t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=self.device)
        self._t[name] = t
        if hasattr(self.ws, name):
            setattr(self.ws, name, t.data_ptr())
        return t

    def _put(self, dst, arr):
        """host array -> the head of a device workspace buffer (parity tapes, init)"""
        src = torch.from_numpy(np.ascontiguousarray(arr)).to(dst.dtype).reshape(-1)
        with torch.cuda.stream(self.stream):
            dst.view(-1)[:src.numel()].copy_(src)
        self.h2d_bytes += src.numel() * src.element_size()

    def _down(self, t):
        """device -> host read (synchronises this chain's stream)"""
        self.d2h_bytes += t.numel() * t.element_size()
        return t.cpu().numpy()

    def _sync(self):
        self.L.stream_sync(self._sp())

    def _pinned(self, name, numel, dtype):
        """a pinned host staging buffer of at least numel elements (grown on demand)"""
        t = self._pins.get(name)
        if t is None or t.numel() < numel or t.dtype != dtype:
            t = torch.empty(max(int(numel), 1), dtype=dtype).pin_memory()
            self._pins[name] = t
        return t

    class _Timed:
        """CUDA-event bracket around launches on the chain's stream (bench.py roofline)."""

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

    def kernel_work(self):
        """name -> per-launch work descriptors (live clusters / rows / uncertain visits) of the
        library-bracketed launches since the last call, in the order of kernel_times_ms()."""
        out, self._event_meta = self._event_meta, {}
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
        self._stream_ptr = self.stream.cuda_stream
        self._dev_thread = threading.get_ident()
        if self.rnd is None:
            self.rnd = PhiloxRandom(np.random.SeedSequence().entropy & 0xFFFFFFFFFFFFFFFF)
        self.rnd.bind(self.device)
        N, M = self.cells_total, self.muts_total
        i32, f64, f32, u8 = torch.int32, torch.float64, torch.float32, torch.uint8
        with torch.cuda.stream(self.stream):
            key = str(self.device)
            with _PACK_LOCK:
                if key not in self._shared:
                    self._shared[key] = _Shared(self.data, self.device)
            self.sh = sh = self._shared[key]
            self.ws = ws = _lib.ChainWs()
            self.ep = _lib.Epoch()
            self.rg = _lib.RgMove()
            self._t = {}
            self._pins = {}
            ws.x1, ws.x0, ws.n1, ws.n0 = (sh.x1.data_ptr(), sh.x0.data_ptr(), sh.n1.data_ptr(),
                                          sh.n0.data_ptr())
            ws.logn = sh.logn.data_ptr()
            ws.W, ws.N, ws.M = sh.W, N, M
            self.assign_d = self._dev('assign', N, i32, zero=True)
            self.st = self._dev('st', _lib.ST_WORDS, i32, zero=True)
            for name, nbytes in (('visit', _lib.VISIT_BYTES), ('cand', _lib.CAND_BYTES),
                                 ('visit_c', _lib.VISIT_BYTES), ('cand_c', _lib.CAND_BYTES)):
                self._dev(name, N * nbytes, u8)
            self._dev('cblk', (N + 127) // 128 + 2, i32)
            self._dev('perm', N, i32)
            self._dev('u', N, f64)
            self._dev('lpx', 2 * _lib.MAX_EXTRA * M, f64)
            self._dev('lpf', 2 * _lib.LEAN_MAXK * M, f32)
            self._dev('llf', N * _lib.LEAN_MAXK, f32)
            self._dev('opt', N * _lib.OPT_BYTES, u8)
            self._dev('n_cert', _lib.LEAN_MAXK, i32, zero=True)
            self._dev('idx_c', N, i32)
            self._dev('bsplit', sh.W * 2 * _lib.LEAN_MAXK * 64, torch.int16)
            self._dev('comp', 512, i32, zero=True)
            self._lean_ok = True
            self._lean_cooldown = 0
            self.members = self._dev('members', N, i32)
            self._dev('rl_tot', 8, f64)
            # split-merge
            self._dev('cells', N + 8, i32)
            self._dev('half', N + 8, i32, zero=True)
            self._dev('gblk', 2 * ((N + 1023) // 1024) + 2, i32)
            self.seg3 = self._dev('seg3', 8, i32, zero=True)
            self._dev('rg_work', 2 * N + 16, i32, zero=True)
            self.rg_theta = self._dev('rg_theta', (3, M), f32, zero=True)
            self.rg_S1 = self._dev('rg_S1', (3, M), i32, zero=True)
            self.rg_S0 = self._dev('rg_S0', (3, M), i32, zero=True)
            self._dev('rg_dec', 4, i32, zero=True)
            self._dev('rg_scal', 32, f64, zero=True)
            self._dev('rg_lp', 4 * M, f64)
            self._dev('rg_ll2', 2 * N, f64)
            self._dev('rg_lq', N, f64)
            self._dev('rg_logq', 3 * M, f64)
            self._dev('rg_A', 2 * M, f64)
            self._dev('rg_orig', 2 * M, f32)
            self._dev('rg_perm', N, i32)
            self._dev('rg_u', N, f64)
            self._dev('rg_rnd', 6 * M, f64)
            self._dev('rg_sd', 2 * M, f64)
            self._dev('rg_beta', 3 * M, f64)
            self.idcap = 0
            self._ll_cap = 0
            self._llx_cap = 0
        self._dev_ready = True
        self._version = 0
        self._stats_version = -1
        self._trace_cache = None

    def _grow_ids(self, need):
        """Make room for cluster ids < need: theta rows, counters, maps and every buffer whose
        size follows the number of live clusters (K <= idcap)."""
        if need <= self.idcap:
            return
        cap = max(need, 2 * self.idcap, 256)
        M = self.muts_total
        i32, f64 = torch.int32, torch.float64
        old = self._t.get('theta')
        self.theta = self._dev('theta', (cap, M), torch.float32, zero=True)
        if old is not None:
            with torch.cuda.stream(self.stream):
                self.theta[:self.idcap] = old
        for name in ('cnt', 'lst', 'rank_of_id', 'ids', 'cursor', 'declined'):
            self._dev(name, cap + 2, i32, zero=True)
        self._dev('seg', cap + 2, i32, zero=True)
        t = self._dev('col_of_id', cap, i32)
        with torch.cuda.stream(self.stream):
            t.fill_(-1)
        self._dev('live_io', 2 * cap, i32, zero=True)
        self._dev('scratch', cap + 1, f64)
        self._dev('lp', 2 * cap * M, f64)
        self.S1 = self._dev('S1', cap * M, i32)
        self.S0 = self._dev('S0', cap * M, i32)
        self._dev('rnd', 3 * cap * M, f64)
        self._dev('rl_out', 5 * cap, f64)
        # pinned host staging
        self._h_in = torch.empty(4 * cap + 16, dtype=i32).pin_memory()
        self._h_out = torch.empty(_lib.ST_WORDS + 2 * cap + 16, dtype=i32).pin_memory()
        self._h_scal = torch.empty(32, dtype=f64).pin_memory()
        # trace staging is sized once per capacity: pinned allocations synchronise the device
        self._pinned('theta_rows', cap * M, torch.float32)
        self._pinned('assign', self.cells_total, i32)
        self.h_in, self.h_out, self.h_scal = self._h_in.numpy(), self._h_out.numpy(), self._h_scal.numpy()
        self.ws.h_in, self.ws.h_out, self.ws.h_scal = (self._h_in.data_ptr(), self._h_out.data_ptr(),
                                                       self._h_scal.data_ptr())
        self.idcap = cap
        self.ws.idcap = cap
        self._stats_version = -1

    def _touch(self):
        self._version += 1
        self._trace_cache = None

    def _assignment_pinned(self):
        N = self.cells_total
        pin = self._pinned('assign', N, torch.int32)
        self.L.copy_async(pin.data_ptr(), self.assign_d.data_ptr(), 4 * N, 2, self._sp())
        self._sync()
        self.d2h_bytes += 4 * N
        return pin[:N].numpy()

    @property
    def assignment(self):
        return self._assignment_pinned().astype(np.int64)

    def assignment_into(self, row):
        """trace row (any integer numpy array of length N) <- current assignment, one pass"""
        np.copyto(row, self._assignment_pinned(), casting='unsafe')

    def copy_assignment_to(self, row):
        """device-side trace: row (int32 [N] device tensor) <- current assignment."""
        self.L.copy_async(row.data_ptr(), self.assign_d.data_ptr(), 4 * self.cells_total, 3, self._sp())

    @property
    def parameters(self):
        return _ThetaView(self)

    def parameters_into(self, ids, out):
        """out[:len(ids)] <- theta rows of the given cluster ids (trace recording, one copy)"""
        out[:len(ids)] = _ThetaView(self)._fetch(ids)

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

    def _streams(self, n):
        """n fresh random stream ids of this chain (production mode)"""
        return self.rnd.reserve(n)

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
            self._put(self.assign_d, a.astype(np.int32))
            self._touch()
            ids_d = torch.arange(K, dtype=torch.int32, device=self.device)
            if given:
                self._refresh_stats()
                tape = self.rnd.beta_rows(K, M)
                tape_d = None
                if tape is not None:
                    tape_d = torch.as_tensor(tape, dtype=torch.float64, device=self.device)
                self.L.beta_rows(self.S1.data_ptr(), self.S0.data_ptr(), K, M, float(self.p),
                                 float(self.q), tape_d.data_ptr() if tape_d is not None else None,
                                 self.rnd.device_seed, self._streams(1) + 1,
                                 self.theta.data_ptr(), ids_d.data_ptr(), self._sp())
            else:
                u = self.rnd.uniform_rows(K, M)
                u = torch.as_tensor(u, dtype=torch.float64, device=self.device)
                self.L.theta_from_uniform(u.data_ptr(), K, M, self.theta.data_ptr(),
                                          ids_d.data_ptr(), self._sp())
            self._sync()
        self._touch()

    # ------------------------------------------------------- sufficient statistics
    def _refresh_stats(self):
        """S1/S0 [K][M] for the live clusters in list order (replaces the
        data[cells] gathers of libs/CRP.py:308,360-367)."""
        if self._stats_version == self._version:
            return
        K = len(self.cells_per_cluster)
        h = self.h_in
        h[:K] = np.fromiter(self.cells_per_cluster.keys(), dtype=np.int32, count=K)
        sizes = np.fromiter(self.cells_per_cluster.values(), dtype=np.int32, count=K)
        h[K] = 0
        np.cumsum(sizes, out=h[K + 1:2 * K + 1])
        # (h_in is consumed before it is written again: every method synchronises the stream
        # after its last call)
        self.L.chain_stats(self.ws, K, int(sizes.max()), self._sp())
        self.h2d_bytes += 4 * (2 * K + 1)
        self._stats_version = self._version

    # ----------------------------------------------------------------- traces
    def _loglik(self, fn, fp, want_prior):
        """full-data log-likelihood at each (FN, FP) pair [+ the Beta prior of theta], from the
        [K][M] sufficient statistics"""
        self._refresh_stats()
        E = len(fn)
        K = len(self.cells_per_cluster)
        fn_a = (C.c_double * max(E, 1))(*fn)
        fp_a = (C.c_double * max(E, 1))(*fp)
        self.L.chain_loglik(self.ws, K, fn_a, fp_a, E, 1 if want_prior else 0, float(self.p),
                            float(self.q), self._sp())
        self._sync()
        rows = E + (1 if want_prior else 0)
        self.d2h_bytes += 8 * rows
        return self.h_scal[:rows].copy()

    def _trace_scalars(self):
        if self._trace_cache is None:
            r = self._loglik([float(self.FN)], [float(self.FP)], not self.beta_prior_uniform)
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
        clusters born inside an epoch get their column computed on the spot.
        One C call and one stream synchronisation per epoch."""
        N, M = self.cells_total, self.muts_total
        L, ep = self.L, self.ep
        sp = self._sp()
        mix0, mix1 = self._beta_mix_const
        FN, FP = float(self.FN), float(self.FP)
        # popcount form of get_lpost_single_new_cluster (libs/CRP.py:230-234)
        ep.c1 = float(np.log(mix1 * (1 - FN) + mix0 * FP))
        ep.c0 = float(np.log(mix1 * FN + mix0 * (1 - FP)))
        ep.c_norm = float(np.log(N - 1 + self.DP_a))
        ep.lnew_prior = float(np.log(self.DP_a) - np.log(N - 1 + self.DP_a))
        ep.log_n = float(np.log(N))
        ep.FN, ep.FP, ep.p, ep.q = FN, FP, float(self.p), float(self.q)
        n_tape = 0
        if self.rnd.is_tape:
            perm, u, beta, n_tape = self.rnd.gibbs_draws(N, M)
            self._put(self._t['perm'], perm)
            self._put(self._t['u'], u)
            with torch.cuda.stream(self.stream):
                beta_d = torch.as_tensor(beta, dtype=torch.float64, device=self.device)
            ep.rand_ready, ep.beta_rows, ep.n_beta_rows = 1, beta_d.data_ptr(), n_tape
            ep.seed, ep.stream_id = 0, 0
        else:
            ep.rand_ready, ep.beta_rows, ep.n_beta_rows = 0, None, 0
            ep.seed, ep.stream_id = self.rnd.device_seed, self._streams(3)
        t, first, epochs, stall = 0, 1, 0, 0
        if not self._lean_ok:
            # the dense route was chosen because of many-rival visits: try lean rows again after a
            # few sweeps (the attempt costs the approximate rows only, see BNPC_STOP_MANY)
            self._lean_cooldown -= 1
            if self._lean_cooldown <= 0:
                self._lean_ok = True
        while t < N:
            K = len(self.cells_per_cluster)
            self._grow_ids(K + _lib.MAX_EXTRA + 2)
            h = self.h_in
            h[0:2 * K:2] = np.fromiter(self.cells_per_cluster.keys(), dtype=np.int32, count=K)
            h[1:2 * K:2] = np.fromiter(self.cells_per_cluster.values(), dtype=np.int32, count=K)
            # odd row stride (bank-conflict-free per-lane row reads in the warp regime)
            ldk = max(3, K | 1)
            # lean epoch: approximate rows select the options, FP64 only where a decision
            # needs it; dense FP64 matrix for longer lists (or when many cells have > 8 rivals)
            lean = (K <= _lib.LEAN_MAXK and self._lean_ok and self.lean_enabled
                    and not (self.force_wide and K <= 63))
            # wide: the dense matrix becomes option weights and one warp walks all visits with
            # lanes <-> clusters (data whose visits mostly have more rivals than an option record)
            wide = (not lean) and self.wide_enabled and K <= 63 and (self.force_wide or not self._lean_ok)
            rows = N - t if lean else int(min(N - t, max(1, LL_BUDGET_BYTES // (8 * ldk))))
            ep.lean = self.lean_rows if lean else (-1 if wide else 0)
            if not lean and rows * ldk + 2 > self._ll_cap:
                self._ll_cap = rows * ldk + 2
                self._dev('ll', self._ll_cap, torch.float64)
            if not lean and not wide and _lib.MAX_EXTRA * rows > self._llx_cap:
                self._llx_cap = _lib.MAX_EXTRA * rows
                self._dev('llx', self._llx_cap, torch.float64)
            ep.first, ep.K, ep.t, ep.rows, ep.ldk = first, K, t, rows, ldk
            if self.profile:
                # CUDA events recorded by the library around the two dominant launches
                evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                for e in evs:
                    e.record(self.stream)            # creates the underlying event
                ep.ev_ll0, ep.ev_ll1, ep.ev_sw0, ep.ev_sw1 = [e.cuda_event for e in evs]
                self._events += [('ll_matrix', evs[0], evs[1]), ('gibbs_sweep', evs[2], evs[3])]
                meta = dict(K=K, rows=rows, lean=bool(lean))
                self._event_meta.setdefault('ll_matrix', []).append(meta)
                self._event_meta.setdefault('gibbs_sweep', []).append(meta)
            else:
                ep.ev_ll0 = ep.ev_ll1 = ep.ev_sw0 = ep.ev_sw1 = None
            L.chain_gibbs_epoch(self.ws, ep, sp)
            self._sync()
            self.h2d_bytes += 8 * K
            st = self.h_out[:_lib.ST_WORDS]
            flags = int(st[_lib.ST_FLAGS])
            if flags & (_lib.STOP_TAPE_EMPTY | _lib.STOP_HANG):
                raise RuntimeError(f'gibbs_sweep stopped with flags {flags:#x} at t={st[_lib.ST_TDONE]}')
            if flags & _lib.STOP_MANY:
                # most visits have more rivals than an option record holds: nothing was done, run
                # this epoch (and the next sweeps) on the dense FP64 matrix
                self._lean_ok = False
                self._lean_cooldown = 8
                continue
            if self.profile:
                meta['n_unc'] = int(st[_lib.ST_NUNC])
            K = int(st[_lib.ST_K])
            self.d2h_bytes += 4 * _lib.ST_WORDS + 8 * K
            pairs = self.h_out[_lib.ST_WORDS:_lib.ST_WORDS + 2 * K].tolist()
            self.cells_per_cluster = OrderedDict(zip(pairs[0::2], pairs[1::2]))
            t_new = int(st[_lib.ST_TDONE])
            if lean and int(st[_lib.ST_NMANY]) > max(64, rows // 50):
                self._lean_ok = False          # too many cells with > 8 rivals: dense rows pay
                self._lean_cooldown = 8
            stall = stall + 1 if t_new == t else 0
            if stall > 2:
                raise RuntimeError(f'gibbs_sweep made no progress at t={t} (flags {flags:#x})')
            if flags & _lib.STOP_CAPACITY:
                self._grow_ids(2 * self.idcap)
            t, first = t_new, 0
            epochs += 1
        if self.rnd.is_tape and int(st[_lib.ST_BIRTHS]) != n_tape:
            raise RuntimeError(f'parity tape held {n_tape} cluster births, sweep made '
                               f'{int(st[_lib.ST_BIRTHS])}')
        self.sweep_stats = dict(epochs=epochs, births=int(st[_lib.ST_BIRTHS]),
                                moved=int(st[_lib.ST_MOVED]), slow=int(st[_lib.ST_SLOW]),
                                uncertain=int(st[_lib.ST_NUNC]),
                                kcycles=int(st[8]), us=int(st[9]) * 1.024, phases_kcyc=[int(st[12 + k]) for k in range(4)],
                                sm_mhz=(int(st[8]) / max(1, int(st[9]))) * 1e3)
        self._touch()

    # ----------------------------------------------------------------- MH theta
    def update_parameters(self, step_no=None):
        """libs/CRP.py:302-311, 314-344 for all live clusters in one launch.
        Returns (declined, accepted) summed over clusters x mutations."""
        self._refresh_stats()
        K, M = len(self.cells_per_cluster), self.muts_total
        if self.rnd.is_tape:
            self._put(self._t['rnd'], self.rnd.mh_theta_draws(K, M))
            ready, seed, sid = 1, 0, 0
        else:
            ready, seed, sid = 0, self.rnd.device_seed, self._streams(2)
        self.L.chain_mh_theta(self.ws, K, ready, seed, sid, float(self.FN), float(self.FP),
                              float(self.p), float(self.q), self._sp())
        self._sync()
        declined = int(self.h_out[0])
        self.d2h_bytes += 4
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
        if k == 1:
            return (self._try_split(step_no), 0)
        if k == self.cells_total:
            return (self._try_merge(step_no), 1)
        move = self.rnd.pick_weighted(ratios)
        if move == 0:
            return (self._try_split(step_no), move)
        return (self._try_merge(step_no), move)

    def _rg_begin(self, n, n_a, cl_i, cl_j, a_i, a_j, is_merge):
        g = self.rg
        g.n, g.n_a, g.cl_i, g.cl_j, g.a_i, g.a_j, g.is_merge = n, n_a, cl_i, cl_j, a_i, a_j, is_merge
        g.rand_ready = 1 if self.rnd.is_tape else 0
        g.seed = self.rnd.device_seed
        FN, FP = float(self.FN), float(self.FP)
        g.alpha, g.FN, g.FP, g.p, g.q = float(self.DP_a), FN, FP, float(self.p), float(self.q)
        mix0 = self._beta_mix_const[0]
        with np.errstate(divide='ignore'):
            k6 = (np.log(1.0 * (1 - FN) + 0.0 * FP), np.log(1.0 * FN + 0.0 * (1 - FP)),
                  np.log(0.0 * (1 - FN) + 1.0 * FP), np.log(0.0 * FN + 1.0 * (1 - FP)),
                  np.log(mix0 * (1 - FN) + (1 - mix0) * FP), np.log(mix0 * FN + (1 - mix0) * (1 - FP)))
        for i in range(6):
            g.k6[i] = float(k6[i])
        self._stats_version = -1          # `members` is reused by the move
        return g

    def _rg_rand(self, n_streams):
        """production: reserve the streams of the next composite call"""
        if not self.rnd.is_tape:
            self.rg.stream_id = self._streams(n_streams)

    def _try_split(self, scans):
        """libs/CRP.py:434-481."""
        ids, sizes = self._ids_sizes()
        weight = sizes / sizes.sum()
        while True:
            where = self.rnd.pick_weighted(weight)
            target = int(ids[where])
            n = int(sizes[where])
            if n != 1:
                break
        a_i, a_j = self.rnd.first_two_of_permutation(n)
        g = self._rg_begin(n, n, target, -1, a_i, a_j, 0)
        lq_pick = np.log(weight[where]) - np.log(n) - np.log(n - 1)
        others = np.delete(sizes, where)
        ok, ones = self._restricted_gibbs('split', n, n, (lq_pick, others), scans, target, -1)
        if not ok:
            return [0, 1]
        new_id = self.get_empty_cluster()
        self._grow_ids(new_id + _lib.MAX_EXTRA + 2)
        self.L.chain_rg_apply(self.ws, g, new_id, self._sp())
        moved = ones + 1
        self.cells_per_cluster[target] -= moved
        self.cells_per_cluster[new_id] = moved
        self._touch()
        return [1, 0]

    def _try_merge(self, scans):
        """libs/CRP.py:484-524."""
        ids, sizes = self._ids_sizes()
        inv = 1 / sizes
        weight = inv / inv.sum()
        w_i, w_j = self.rnd.pick_two_weighted(weight)
        cl_i, cl_j = int(ids[w_i]), int(ids[w_j])
        n_a, n_b = int(sizes[w_i]), int(sizes[w_j])
        a_i = self.rnd.randint(n_a)
        a_j = self.rnd.randint(n_b)
        n = n_a + n_b
        g = self._rg_begin(n, n_a, cl_i, cl_j, a_i, a_j, 1)
        both = np.argwhere((ids == cl_j) | (ids == cl_i)).flatten()
        lq_pick = np.nansum(np.log(weight[both])) - np.nansum(np.log(sizes[both]))
        ok, _ = self._restricted_gibbs('merge', n, n_a, lq_pick, scans, cl_i, cl_j)
        if not ok:
            return [0, 1]
        self.L.chain_rg_apply(self.ws, g, -1, self._sp())
        self.cells_per_cluster[cl_i] += n_b
        del self.cells_per_cluster[cl_j]
        self._touch()
        return [1, 0]

    # -- restricted Gibbs machinery (libs/CRP.py:527-638) -----------------------
    # Everything a split-merge move computes on the device is enqueued without waiting (one C call
    # per scan); the scalars the Metropolis-Hastings decision needs are reduced into slots of
    # `rg_scal` and read back ONCE (one stream synchronisation per move).
    SL_FWD_ASSIGN, SL_FWD_THETA, SL_BACK, SL_BACK_LQ, SL_PRIOR_NEW, SL_PRIOR_OLD, SL_LL3 = 0, 1, 2, 3, 4, 6, 8

    def _scan_split(self, n, want_logq=False):
        """libs/CRP.py:570-578, 590-632."""
        M, nf = self.muts_total, n - 2
        if self.rnd.is_tape:
            if n > 2:
                perm, u = self.rnd.scan_draws(nf)
                self._put(self._t['rg_perm'], perm)
                self._put(self._t['rg_u'], u)
            self._put(self._t['rg_rnd'], self.rnd.mh_theta_draws(2, M))
        else:
            self._rg_rand(4)
        self.L.chain_rg_scan_split(self.ws, self.rg, 1 if want_logq else 0, self._sp())

    def _scan_merged(self, want_logq=False):
        """libs/CRP.py:581-587."""
        if self.rnd.is_tape:
            self._put(self._t['rg_rnd'], self.rnd.mh_theta_draws(1, self.muts_total))
        else:
            self._rg_rand(2)
        self.L.chain_rg_scan_merged(self.ws, self.rg, 1 if want_logq else 0, self._sp())

    def _restricted_gibbs(self, move, n, n_a, size_term, scans, cl_i, cl_j):
        """libs/CRP.py:527-567.  Returns (accepted, number of free cells on side j)."""
        M = self.muts_total
        # launch state: free cells go to the anchor whose raw row explains them better
        if self.rnd.is_tape:
            self._put(self._t['rg_beta'], np.concatenate([self.rnd.beta_rows(2, M), self.rnd.beta_rows(1, M)]))
        else:
            self._rg_rand(2)
        self.L.chain_rg_setup(self.ws, self.rg, self._sp())
        for _ in range(scans):
            self._scan_split(n)
            self._scan_merged()
        if move == 'split':
            return self._decide_split(n, size_term, cl_i)
        return self._decide_merge(n, n_a, size_term, cl_i, cl_j)

    def _decide_split(self, n, size_term, cl_i):
        """libs/CRP.py:641-653 with :668-682, :695-733, :757-764."""
        M = self.muts_total
        self._scan_split(n, want_logq=True)
        if self.rnd.is_tape:
            self._put(self._t['rg_sd'], self.rnd.step_sd_index(1, M))
        else:
            self._rg_rand(1)
        self.L.chain_rg_decide_split(self.ws, self.rg, 1 if self.beta_prior_uniform else 0, self._sp())
        self._sync()
        self.d2h_bytes += 8 * 16 + 32
        sc = self.h_scal[:16].copy()
        ones = int(self.h_out[3]) if n > 2 else 0           # free cells on side j (rg_sides)
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
        M, nf = self.muts_total, n - 2
        self._scan_merged(want_logq=True)
        if self.rnd.is_tape:
            self._put(self._t['rg_sd'], self.rnd.step_sd_index(2, M))
        else:
            self._rg_rand(1)
        self.L.chain_rg_decide_merge(self.ws, self.rg, 1 if self.beta_prior_uniform else 0, self._sp())
        self._sync()
        self.d2h_bytes += 8 * 16
        sc = self.h_scal[:16].copy()
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
        r = self._loglik([float(fn) for _, fn in pairs], [float(fp) for fp, _ in pairs], False)
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
                prop = float(_tn_ppf(u, lo, hi) * sd + cur)
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
