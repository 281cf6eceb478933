"""ctypes binding of libbnpc_b200.so (C ABI declared in include/bnpc_b200.h).

The shared library is built in-tree by `build()` (plain nvcc, sm_100a only) and
loaded lazily.  There is no fallback: if the library is missing or a call fails
a RuntimeError is raised.
"""
import ctypes as C
import glob
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
SO_PATH = os.path.join(_HERE, 'libbnpc_b200.so')
SOURCES = [os.path.join(_HERE, 'csrc', 'bnpc_kernels.cu')]
# every header of the translation unit: editing any of them makes the library stale
HEADERS = sorted(glob.glob(os.path.join(_HERE, 'csrc', '*.cuh'))) + sorted(glob.glob(os.path.join(_ROOT, 'include', '*.h')))

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3',
              '-fmad=false', '-std=c++17', '-shared', '-Xcompiler', '-fPIC']

ABI_VERSION = 12
MAX_EXTRA = 32
ST_K, ST_TDONE, ST_FLAGS, ST_NEXTRA, ST_BIRTHS, ST_MOVED, ST_SLOW = range(7)
ST_NUNC = 10
ST_NMANY = 11
LEAN_MAXK = 64
OPT_BYTES = 16
ST_WORDS = 16
STOP_EXTRA_FULL, STOP_REPACK, STOP_TAPE_EMPTY, STOP_CAPACITY, STOP_MANY, STOP_HANG = 1, 2, 4, 8, 16, 0x100
VISIT_BYTES = 64
VISIT_CELL_OFFSET = 32
CAND_BYTES = 96
MAX_LIST = 1024          # longest cluster list the sequencer regime of the sweep handles


def _stale():
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile the CUDA sources for sm_100a into bnpc_b200/libbnpc_b200.so."""
    if not force and not _stale():
        return SO_PATH
    nvcc = os.environ.get('NVCC', 'nvcc')
    cmd = [nvcc] + NVCC_FLAGS + ['-o', SO_PATH] + SOURCES
    if verbose:
        print(' '.join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f'nvcc failed:\n{res.stdout}\n{res.stderr}')
    return SO_PATH


class SweepArgs(C.Structure):
    _fields_ = [
        ('x1', C.c_void_p), ('x0', C.c_void_p), ('W', C.c_int32), ('N', C.c_int32), ('M', C.c_int32),
        ('assign', C.c_void_p), ('cnt', C.c_void_p), ('lst', C.c_void_p), ('col_of_id', C.c_void_p),
        ('theta', C.c_void_p), ('idcap', C.c_int32), ('st', C.c_void_p), ('live_out', C.c_void_p),
        ('ll', C.c_void_p), ('ldk', C.c_int32), ('t_epoch0', C.c_int32), ('lp', C.c_void_p), ('comp', C.c_void_p),
        ('lpx', C.c_void_p), ('llx', C.c_void_p), ('ldx', C.c_int32),
        ('scratch', C.c_void_p),
        ('visit', C.c_void_p), ('cand', C.c_void_p), ('t_begin', C.c_int32), ('t_end', C.c_int32),
        ('visit_c', C.c_void_p), ('cand_c', C.c_void_p),
        ('beta_rows', C.c_void_p), ('n_beta_rows', C.c_int32),
        ('seed', C.c_uint64), ('stream_id', C.c_uint64),
        ('logn', C.c_void_p),
        ('c_norm', C.c_double), ('FN', C.c_double), ('FP', C.c_double), ('p', C.c_double), ('q', C.c_double),
        ('owner_c', C.c_void_p), ('wide', C.c_int32),
    ]


_P, _I, _D, _U64, _I64, _F = C.c_void_p, C.c_int, C.c_double, C.c_uint64, C.c_int64, C.c_float


class ChainWs(C.Structure):
    """bnpc_chain_t: device (and pinned host) buffers of one chain."""
    _fields_ = [(n, C.c_void_p) for n in ('x1', 'x0', 'n1', 'n0', 'logn')] + \
        [(n, C.c_int32) for n in ('W', 'N', 'M', 'idcap')] + \
        [(n, C.c_void_p) for n in (
            'assign', 'theta', 'cnt', 'lst', 'col_of_id', 'rank_of_id', 'live_io', 'st',
            'visit', 'cand', 'visit_c', 'cand_c', 'cblk', 'perm', 'u',
            'lp', 'll', 'lpx', 'llx', 'scratch', 'lpf', 'llf', 'opt', 'n_cert', 'idx_c', 'bsplit', 'comp',
            'ids', 'seg', 'cursor', 'members', 'S1', 'S0', 'rnd', 'declined', 'rl_out', 'rl_tot',
            'cells', 'half', 'gblk', 'seg3', 'rg_work', 'rg_theta', 'rg_S1', 'rg_S0', 'rg_dec', 'rg_scal',
            'rg_lp', 'rg_ll2', 'rg_lq', 'rg_logq', 'rg_A', 'rg_orig', 'rg_perm', 'rg_u', 'rg_rnd', 'rg_sd',
            'rg_beta', 'h_in', 'h_out', 'h_scal')]


class Epoch(C.Structure):
    """bnpc_epoch_t"""
    _fields_ = [(n, C.c_int32) for n in ('first', 'K', 't', 'rows', 'ldk', 'rand_ready', 'lean', 'serial_sweep')] + \
        [(n, C.c_double) for n in ('c1', 'c0', 'lnew_prior', 'c_norm', 'log_n', 'FN', 'FP', 'p', 'q')] + \
        [('seed', C.c_uint64), ('stream_id', C.c_uint64), ('beta_rows', C.c_void_p),
         ('n_beta_rows', C.c_int32)] + \
        [(n, C.c_void_p) for n in ('ev_ll0', 'ev_ll1', 'ev_sw0', 'ev_sw1')]


class RgMove(C.Structure):
    """bnpc_rg_t"""
    _fields_ = [(n, C.c_int32) for n in ('n', 'n_a', 'cl_i', 'cl_j', 'a_i', 'a_j', 'is_merge', 'rand_ready')] + \
        [(n, C.c_double) for n in ('alpha', 'FN', 'FP', 'p', 'q')] + [('k6', C.c_double * 6)] + \
        [('seed', C.c_uint64), ('stream_id', C.c_uint64)]


class ChainState(C.Structure):
    """bnpc_chain_state_t: host-side state of one chain of a group"""
    _fields_ = [(n, C.c_double) for n in ('p', 'q', 'mix0', 'mix1', 'dp_a0', 'dp_b0', 'fp_mean', 'fp_sd', 'fn_mean',
                                          'fn_sd', 'fp_prior_const', 'fn_prior_const')] + \
        [(n, C.c_int32) for n in ('learning', 'beta_prior_uniform', 'lean_enabled', 'lean_rows', 'serial_sweep',
                                  'wide_enabled', 'force_wide', 'lean_ok', 'lean_cooldown', 'stats_fresh')] + \
        [(n, C.c_double) for n in ('DP_a', 'FN', 'FP')] + \
        [(n, C.c_uint64) for n in ('seed', 'host_ctr', 'dev_calls')] + \
        [('K', C.c_int32), ('live_cap', C.c_int32), ('live', C.c_void_p), ('ll_cap', C.c_int64),
         ('llx_cap', C.c_int64), ('mh_counter', C.c_double * 10)] + \
        [(n, C.c_int32) for n in ('last_epochs', 'last_births', 'last_moved', 'last_nunc')]


class Moves(C.Structure):
    """bnpc_moves_t"""
    _fields_ = [('sm_prob', C.c_double), ('dpa_prob', C.c_double), ('error_prob', C.c_double),
                ('sm_ratios', C.c_double * 2), ('sm_steps', C.c_int32), ('fix_assign', C.c_int32)]


class Trace(C.Structure):
    """bnpc_trace_t"""
    _fields_ = [(n, C.c_void_p) for n in ('ml', 'map', 'alpha', 'fn', 'fp', 'n_clusters', 'assign_h')] + \
        [('assign_stride', C.c_int64), ('params_h', C.c_void_p), ('params_kcap', C.c_int32),
         ('params_first', C.c_int32)]


GROW_IDS, GROW_LL, GROW_LLX, GROW_LIVE, GROW_PARAMS, GROW_RING_K = 1, 2, 3, 4, 5, 6
GROW_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int64)

_WS, _EP, _RG = C.POINTER(ChainWs), C.POINTER(Epoch), C.POINTER(RgMove)

# name -> argument ctypes, exactly as declared in include/bnpc_b200.h
SIGNATURES = {
    'bnpc_pack_planes': [_P, _P, _I, _I, _I, _P, _P, _P, _P, _P],
    'bnpc_fill_uniform': [_P, _I64, _U64, _U64, _I, _P],
    'bnpc_fill_permutation': [_P, _I, _U64, _U64, _P],
    'bnpc_logprob_tables': [_P, _P, _I, _I, _D, _D, _P, _P],
    'bnpc_ll_matrix': [_P, _P, _I, _I, _P, _I, _I, _P, _I, _P, _I, _P],
    'bnpc_gibbs_prepare': [_P, _P, _P, _P, _P, _I, _D, _D, _D, _P, _P],
    'bnpc_gibbs_candidates': [_P, _I, _I, _P, _P, _P, _I, _D, _D, _P, _P],
    'bnpc_gibbs_compact': [_P, _P, _I, _P, _P, _P, _P, _P],
    'bnpc_ll_matrix_f32': [_P, _P, _I, _I, _P, _I, _I, _P, _P, _I, _P, _I, _P],
    'bnpc_ll_matrix_tc': [_P, _P, _I, _I, _P, _I, _I, _P, _P, _I, _P, _I, _P],
    'bnpc_cocluster_counts': [_P, _I, _I, _P, _P],
    'bnpc_mpear_sums': [_P, _I, _P, _I, _P, _P],
    'bnpc_mpear_sums_weighted': [_P, _I, _P, _I, _P, _P, _P],
    'bnpc_ll_matrix_i8': [_P, _P, _I, _I, _P, _I, _I, _P, _P, _I, _D, _P, _I, _P],
    'bnpc_ll_matrix_i8_shared': [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P],
    'bnpc_ll_shared_plan': [_I, _P, _I, _I, _P, _P, _P, _P, _P, _P],
    'bnpc_gibbs_options': [_P, _I, _I, _P, _P, _P, _P, _I, _D, _D, _I, _D, _P],
    'bnpc_gibbs_exact': [_P, _P, _I, _I, _P, _I, _P, _P, _P, _I, _P, _P, _P, _P, _P, _D, _D, _P, _P, _P],
    'bnpc_gibbs_epoch_begin': [_P, _I, _P, _P, _P, _I, _P, _I, _P],
    'bnpc_gibbs_sweep': [C.POINTER(SweepArgs), _I, _P],
    'bnpc_group_members': [_P, _I, _P, _P, _P, _I, _P, _P],
    'bnpc_set_ranks': [_P, _I, _P, _P],
    'bnpc_suffstat': [_P, _P, _I, _I, _P, _P, _I, _I, _P, _P, _P],
    'bnpc_beta_rows': [_P, _P, _I, _I, _D, _D, _P, _U64, _U64, _P, _P, _P],
    'bnpc_theta_from_uniform': [_P, _I, _I, _P, _P, _P],
    'bnpc_mh_theta': [_P, _P, _I, _I, _P, _P, _P, _D, _D, _D, _D, _I, _P, _P, _P],
    'bnpc_theta_log_ratio': [_P, _P, _I, _I, _P, _P, _P, _F, _F, _D, _D, _D, _D, _P, _P],
    'bnpc_row_loglik': [_P, _P, _I, _I, _P, _P, C.POINTER(_D), C.POINTER(_D), _I, _D, _D, _P, _P, _P],
    'bnpc_row_sum': [_P, _I, _I, _P, _P],
    'bnpc_gather_members': [_P, _I, _I, _I, _P, _P, _P],
    'bnpc_anchor_swaps': [_P, _I, _I, _I, _I, _I, _P],
    'bnpc_rg_launch_halves': [_P, _P, _I, _P, _I, C.POINTER(_D), _P, _P],
    'bnpc_rg_sides': [_P, _I, _P, _P, _P, _P],
    'bnpc_rg_scan': [_P, _I, _I, _P, _P, _P, _D, _I, _P, _P, _I, _P, _P, _P],
    'bnpc_apply_split': [_P, _I, _P, _I, _P, _P],
    'bnpc_apply_merge': [_P, _I, _I, _I, _P, _P],
    'bnpc_chain_gibbs_epoch': [_WS, _EP, _P],
    'bnpc_chain_stats': [_WS, _I, _I, _P],
    'bnpc_chain_mh_theta': [_WS, _I, _I, _U64, _U64, _D, _D, _D, _D, _P],
    'bnpc_chain_loglik': [_WS, _I, C.POINTER(_D), C.POINTER(_D), _I, _I, _D, _D, _P],
    'bnpc_chain_theta_rows': [_WS, _I, _P, _P],
    'bnpc_copy_async': [_P, _P, _I64, _I, _P],
    'bnpc_stream_sync': [_P],
    'bnpc_chain_rg_setup': [_WS, _RG, _P],
    'bnpc_chain_rg_scan_split': [_WS, _RG, _I, _P],
    'bnpc_chain_rg_scan_merged': [_WS, _RG, _I, _P],
    'bnpc_chain_rg_decide_split': [_WS, _RG, _I, _P],
    'bnpc_chain_rg_decide_merge': [_WS, _RG, _I, _P],
    'bnpc_chain_rg_apply': [_WS, _RG, _I, _P],
    'bnpc_group_set_ring': [_P, _I, _P, _P, _I],
    'bnpc_group_run': [_P, _I, _I],
    'bnpc_group_record': [_P, _I],
    'bnpc_batch_begin': [],
    'bnpc_batch_slot': [_I],
    'bnpc_batch_flush': [_P],
    'bnpc_prof_enable': [_I],
    'bnpc_prof_report': [C.c_char_p, _I],
}

# double-valued host helpers and the two entry points that do not return a status
HOST_SCALAR_SIGNATURES = {
    'bnpc_host_random': [_U64, C.POINTER(_U64)],
    'bnpc_host_gamma': [_U64, C.POINTER(_U64), _D],
    'bnpc_host_beta': [_U64, C.POINTER(_U64), _D, _D],
    'bnpc_host_truncnorm_ppf': [_D, _D, _D],
    'bnpc_host_truncnorm_logpdf': [_D, _D, _D, _D, _D],
}
OTHER_SYMBOLS = ('bnpc_abi_version', 'bnpc_last_error', 'bnpc_launch_count', 'bnpc_group_create',
                 'bnpc_group_destroy')

_lock = threading.Lock()
_lib = None


def launch_count():
    """kernels launched through the library since it was loaded (bench.py reports it)"""
    return int(lib()._dll.bnpc_launch_count())


class _Lib:
    def __init__(self, path):
        self._dll = C.CDLL(path)
        self._dll.bnpc_last_error.restype = C.c_char_p
        self._dll.bnpc_abi_version.restype = C.c_int
        self._dll.bnpc_launch_count.restype = C.c_int64
        if self._dll.bnpc_abi_version() != ABI_VERSION:
            raise RuntimeError(f'{path} has ABI version {self._dll.bnpc_abi_version()}, this package expects '
                               f'{ABI_VERSION}: rebuild it (python -c "import __graft_entry__ as g; g.build()")')
        d = self._dll
        d.bnpc_group_create.restype = C.c_void_p
        d.bnpc_group_create.argtypes = [_I, C.POINTER(_WS), C.POINTER(C.POINTER(ChainState)),
                                        C.POINTER(C.POINTER(Trace)), C.POINTER(Moves), GROW_FN, _P, _P, _P, _P]
        d.bnpc_group_destroy.restype = None
        d.bnpc_group_destroy.argtypes = [_P]
        # debugging hook (include/bnpc_b200_debug.h; tools/tc_trace.py), not part of the drop-in ABI
        d.bnpc_debug_set_trace.argtypes = [_P]
        d.bnpc_debug_set_trace.restype = C.c_int
        self.debug_set_trace = d.bnpc_debug_set_trace
        for name, args in HOST_SCALAR_SIGNATURES.items():
            fn = getattr(d, name)
            fn.restype = _D
            fn.argtypes = args
            setattr(self, name[len('bnpc_'):], fn)
        for name, args in SIGNATURES.items():
            fn = getattr(self._dll, name)
            fn.argtypes = args
            fn.restype = C.c_int
            setattr(self, name[len('bnpc_'):], self._wrap(name, fn))

    def _wrap(self, name, fn):
        def call(*args):
            rc = fn(*args)
            if rc != 0:
                raise RuntimeError(f'{name} failed ({rc}): '
                                   f'{self._dll.bnpc_last_error().decode()}')
        call.__name__ = name
        return call

    def abi_version(self):
        return self._dll.bnpc_abi_version()

    def group_create(self, *args):
        h = self._dll.bnpc_group_create(*args)
        if not h:
            raise RuntimeError(f'bnpc_group_create failed: {self._dll.bnpc_last_error().decode()}')
        return h

    def group_destroy(self, h):
        self._dll.bnpc_group_destroy(h)

    def prof_report_text(self):
        buf = C.create_string_buffer(1 << 16)
        self.prof_report(buf, len(buf))
        out = {}
        for line in buf.value.decode().splitlines():
            name, launches, chains, ms = line.rsplit(' ', 3)
            out[name] = dict(launches=int(launches), chains=int(chains), total_ms=float(ms))
        return out


def lib():
    """The loaded library; raises RuntimeError if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(SO_PATH):
                    raise RuntimeError(
                        f'{SO_PATH} is missing: run `python -c "import __graft_entry__ as g; '
                        f'g.build()"` (needs nvcc); there is no CPU fallback')
                _lib = _Lib(SO_PATH)
    return _lib
