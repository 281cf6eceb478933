"""All chains of one GPU stepped by the library's native lockstep driver (bnpc_group_run).

The reference runs one OS process per chain (libs/MCMC.py:113-120).  Here the chains that share
a GPU form a `ChainGroup`: one host thread, one C call per batch of steps; inside, every kernel
is launched once for all chains (blockIdx.z = chain), the host waits once per phase for the whole
group, and the per-step trace rows (libs/MCMC.py:242-282) travel through device rings that a copy
stream drains into pinned host arrays while the next step runs.

`ChainGroup` works on chain objects with the reference's `Chain` interface (`libs.MCMC.Chain`):
`.model` (an initialised `DeviceCRP` with a `PhiloxRandom` source), `.results` (trace arrays) and
`.MH_counter`.  The per-method Python mirror of the model (`bnpc_b200.engine`) stays the parity
path (random tape against the oracle); a chain walks the same trajectory under either host.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib

RING_SLOTS = 8


def pinned_zeros(shape, dtype):
    """(numpy view, owner tensor) of page-locked host memory; falls back to pageable memory when
    the allocation fails (the copies then run synchronously)"""
    tdtype = {np.dtype(np.int32): torch.int32, np.dtype(np.float32): torch.float32,
              np.dtype(np.float64): torch.float64}[np.dtype(dtype)]
    try:
        t = torch.zeros(tuple(int(s) for s in shape), dtype=tdtype)
        if torch.cuda.is_available() and t.numel() > 0:
            t = t.pin_memory()
    except RuntimeError:
        t = torch.zeros(tuple(int(s) for s in shape), dtype=tdtype)
    return t.numpy(), t


def native_ok(model):
    """True for a CUDA model whose draws come from counter-based streams (not a parity tape)"""
    rnd = getattr(model, 'rnd', None)
    return hasattr(model, 'ws') and getattr(model, '_dev_ready', False) and rnd is not None \
        and not getattr(rnd, 'is_tape', True)


class ChainGroup:
    def __init__(self, chains, moves, fix_assign=False, host_assign=True):
        """chains: objects with .model/.results/.MH_counter on ONE device; moves: the dict of
        libs.MCMC.MCMC.params; host_assign=False keeps the assignment trace in the device ring
        (no per-step device->host copy of the assignment vector)."""
        if not chains:
            raise ValueError('empty chain group')
        self.chains = list(chains)
        self.models = [ch.model for ch in self.chains]
        for m in self.models:
            if not native_ok(m):
                raise RuntimeError('ChainGroup needs initialised CUDA models with PhiloxRandom sources')
        dev = {str(m.device) for m in self.models}
        if len(dev) != 1:
            raise RuntimeError(f'the chains of a group share one device, got {sorted(dev)}')
        self.device = self.models[0].device
        self.L = _lib.lib()
        self.host_assign = host_assign
        n = len(self.chains)
        torch.cuda.set_device(self.device)
        self.sA, self.sB, self.sC = (torch.cuda.Stream(self.device) for _ in range(3))
        if os.environ.get('BNPC_ONE_STREAM'):          # debugging: split-merge moves on the main stream
            self.sB = self.sA
        self.states = [_lib.ChainState() for _ in range(n)]
        self.traces = [_lib.Trace() for _ in range(n)]
        self._live = [np.zeros(2 * max(m.idcap, 64), dtype=np.int32) for m in self.models]
        self._nclu = [None] * n
        mv = _lib.Moves()
        mv.sm_prob, mv.dpa_prob, mv.error_prob = (float(moves['sm_prob']), float(moves['dpa_prob']),
                                                  float(moves['error_prob']))
        mv.sm_ratios[0], mv.sm_ratios[1] = float(moves['sm_ratios'][0]), float(moves['sm_ratios'][1])
        mv.sm_steps, mv.fix_assign = int(moves['sm_steps']), int(bool(fix_assign))
        self._moves = mv
        self._ws_arr = (C.POINTER(_lib.ChainWs) * n)(*[C.pointer(m.ws) for m in self.models])
        self._st_arr = (C.POINTER(_lib.ChainState) * n)(*[C.pointer(s) for s in self.states])
        self._tr_arr = (C.POINTER(_lib.Trace) * n)(*[C.pointer(t) for t in self.traces])
        self._cb_error = None
        self._cb = _lib.GROW_FN(self._grow)
        self.handle = self.L.group_create(n, self._ws_arr, self._st_arr, self._tr_arr, C.byref(mv), self._cb, None,
                                          self.sA.cuda_stream, self.sB.cuda_stream, self.sC.cuda_stream)
        N, M = self.models[0].cells_total, self.models[0].muts_total
        self.ring_assign = torch.zeros((RING_SLOTS, n, N), dtype=torch.int32, device=self.device)
        self.ring_kcap = max(128, 2 * max(len(m.cells_per_cluster) for m in self.models))
        self.ring_theta = torch.zeros((RING_SLOTS, n, self.ring_kcap, M), dtype=torch.float32, device=self.device)
        torch.cuda.synchronize(self.device)
        self.L.group_set_ring(self.handle, RING_SLOTS, self.ring_assign.data_ptr(), self.ring_theta.data_ptr(),
                              self.ring_kcap)
        for i, m in enumerate(self.models):
            self._constants(i, m)

    def close(self):
        if self.handle:
            self.L.group_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:                        # noqa: BLE001  (interpreter shutdown)
            pass

    # ------------------------------------------------------------------ state in / out
    def _constants(self, i, m):
        s = self.states[i]
        s.p, s.q = float(m.p), float(m.q)
        s.mix0, s.mix1 = float(m._beta_mix_const[0]), float(m._beta_mix_const[1])
        s.dp_a0, s.dp_b0 = float(m.DP_a_gamma[0]), float(m.DP_a_gamma[1])
        s.learning = int(bool(m.learning))
        s.beta_prior_uniform = int(bool(m.beta_prior_uniform))
        if m.learning:
            s.fp_mean, s.fp_sd = float(m.FP_prior.mean), float(m.FP_prior.sd)
            s.fn_mean, s.fn_sd = float(m.FN_prior.mean), float(m.FN_prior.sd)
            s.fp_prior_const, s.fn_prior_const = float(m.FP_prior._const), float(m.FN_prior._const)
        s.lean_enabled, s.lean_rows = int(bool(m.lean_enabled)), int(m.lean_rows)
        s.serial_sweep, s.wide_enabled, s.force_wide = (int(bool(m.serial_sweep)), int(bool(m.wide_enabled)),
                                                        int(bool(m.force_wide)))

    def _import(self):
        for i, m in enumerate(self.models):
            s = self.states[i]
            m.stream.synchronize()
            s.lean_ok, s.lean_cooldown = int(bool(m._lean_ok)), int(m._lean_cooldown)
            s.stats_fresh = int(m._stats_version == m._version)
            s.DP_a, s.FN, s.FP = float(m.DP_a), float(m.FN), float(m.FP)
            s.seed, s.host_ctr, s.dev_calls = m.rnd.seed, m.rnd.host_ctr.value, m.rnd.calls
            K = len(m.cells_per_cluster)
            if 2 * K > self._live[i].size:
                self._live[i] = np.zeros(4 * K + 64, dtype=np.int32)
            live = self._live[i]
            live[0:2 * K:2] = np.fromiter(m.cells_per_cluster.keys(), dtype=np.int32, count=K)
            live[1:2 * K:2] = np.fromiter(m.cells_per_cluster.values(), dtype=np.int32, count=K)
            s.K, s.live, s.live_cap = K, live.ctypes.data, live.size
            s.ll_cap, s.llx_cap = int(m._ll_cap), int(m._llx_cap)
            for j in range(10):
                s.mh_counter[j] = 0.0

    def _export(self):
        from collections import OrderedDict
        for i, (ch, m) in enumerate(zip(self.chains, self.models)):
            s = self.states[i]
            live = self._live[i]
            m.cells_per_cluster = OrderedDict(zip(live[0:2 * s.K:2].tolist(), live[1:2 * s.K:2].tolist()))
            m.DP_a, m.FN, m.FP = float(s.DP_a), float(s.FN), float(s.FP)
            m.rnd.host_ctr.value, m.rnd.calls = s.host_ctr, int(s.dev_calls)
            m._lean_ok, m._lean_cooldown = bool(s.lean_ok), int(s.lean_cooldown)
            m._touch()
            m._stats_version = m._version if s.stats_fresh else -1
            m.sweep_stats = dict(epochs=int(s.last_epochs), births=int(s.last_births), moved=int(s.last_moved),
                                 uncertain=int(s.last_nunc))
            ch.MH_counter += np.array(list(s.mh_counter)).reshape(5, 2)

    # ------------------------------------------------------------------ growth (called from C)
    def _grow(self, _ctx, ci, kind, need):
        try:
            m, s = self.models[ci], self.states[ci]
            torch.cuda.set_device(self.device)
            self.sA.synchronize()
            self.sB.synchronize()
            if kind == _lib.GROW_IDS:
                m._grow_ids(int(need))
            elif kind == _lib.GROW_LL:
                m._ll_cap = int(need)
                m._dev('ll', m._ll_cap, torch.float64)
                s.ll_cap = m._ll_cap
            elif kind == _lib.GROW_LLX:
                m._llx_cap = int(need)
                m._dev('llx', m._llx_cap, torch.float64)
                s.llx_cap = m._llx_cap
            elif kind == _lib.GROW_LIVE:
                live = np.zeros(2 * int(need), dtype=np.int32)
                live[:self._live[ci].size] = self._live[ci]
                self._live[ci] = live
                s.live, s.live_cap = live.ctypes.data, live.size
            elif kind == _lib.GROW_PARAMS:
                self.sC.synchronize()
                self.chains[ci]._par_grow(int(need))
                self._trace_params(ci)
            elif kind == _lib.GROW_RING_K:
                self.sC.synchronize()
                n, M = len(self.models), m.muts_total
                self.ring_kcap = max(int(need), 2 * self.ring_kcap)
                self.ring_theta = torch.zeros((RING_SLOTS, n, self.ring_kcap, M), dtype=torch.float32,
                                              device=self.device)
                torch.cuda.synchronize(self.device)
                self.L.group_set_ring(self.handle, RING_SLOTS, self.ring_assign.data_ptr(),
                                      self.ring_theta.data_ptr(), self.ring_kcap)
            else:
                return 1
            m.stream.synchronize()
            return 0
        except BaseException as exc:             # noqa: BLE001  (must not unwind through C)
            self._cb_error = exc
            return 1

    # ------------------------------------------------------------------ traces
    def _trace_params(self, ci):
        ch, tr = self.chains[ci], self.traces[ci]
        buf = getattr(ch, '_par_buf', None)
        if buf is None:
            tr.params_h, tr.params_kcap, tr.params_first = None, 0, 2 ** 31 - 1
        else:
            tr.params_h, tr.params_kcap, tr.params_first = buf.ctypes.data, buf.shape[1], int(ch._par_first)

    def _bind_traces(self):
        for i, ch in enumerate(self.chains):
            r, tr = ch.results, self.traces[i]
            for key, field in (('ML', 'ml'), ('MAP', 'map'), ('DP_alpha', 'alpha'), ('FN', 'fn'), ('FP', 'fp')):
                a = r[key]
                if a.dtype != np.float64 or not a.flags.c_contiguous:
                    raise RuntimeError(f'trace {key} must be a contiguous float64 array')
                setattr(tr, field, a.ctypes.data)
            steps = r['ML'].size
            if self._nclu[i] is None or self._nclu[i].size < steps:
                grown = np.zeros(steps, dtype=np.int32)
                if self._nclu[i] is not None:
                    grown[:self._nclu[i].size] = self._nclu[i]
                self._nclu[i] = grown
            tr.n_clusters = self._nclu[i].ctypes.data
            a = r['assignments']
            if self.host_assign:
                if a.dtype != np.int32 or not a.flags.c_contiguous:
                    raise RuntimeError('the assignment trace must be a contiguous int32 array')
                tr.assign_h, tr.assign_stride = a.ctypes.data, a.shape[1]
            else:
                tr.assign_h, tr.assign_stride = None, 0
            self._trace_params(i)

    def _finish_params(self, step0, n_steps):
        for i, ch in enumerate(self.chains):
            if getattr(ch, '_par_buf', None) is None:
                continue
            lo = max(step0, int(ch._par_first))
            if lo < step0 + n_steps:
                ch._par_k = max(ch._par_k, int(self._nclu[i][lo:step0 + n_steps].max()))
            ch.results['params'] = ch._par_buf[:, :ch._par_k]

    # ------------------------------------------------------------------ running
    def _call(self, fn, *args):
        torch.cuda.set_device(self.device)
        self._import()
        self._bind_traces()
        self._cb_error = None
        try:
            fn(self.handle, *args)
        except RuntimeError as exc:
            if self._cb_error is not None:
                raise RuntimeError(f'growing a chain buffer failed: {self._cb_error!r}') from self._cb_error
            raise exc
        finally:
            self._export()

    def run(self, step0, n_steps):
        """steps step0 .. step0+n_steps-1 of every chain (trace rows of those indices)"""
        if n_steps <= 0:
            return
        self._call(self.L.group_run, int(step0), int(n_steps))
        self._finish_params(step0, n_steps)

    def record(self, step):
        """the trace row of the current state (libs/MCMC.py:218: update_results(0))"""
        self._call(self.L.group_record, int(step))
        self._finish_params(step, 1)

    def assignment_ring(self):
        """device ring of the most recent assignment rows: int32 [slots][chains][N]"""
        return self.ring_assign
