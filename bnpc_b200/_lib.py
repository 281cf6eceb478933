"""ctypes binding of libbnpc_b200.so (C ABI declared in include/bnpc_b200.h).

The shared library is built in-tree by `build()` (plain nvcc, sm_100a only) and
loaded lazily.  There is no fallback: if the library is missing or a call fails
a RuntimeError is raised.
"""
import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
SO_PATH = os.path.join(_HERE, 'libbnpc_b200.so')
SOURCES = [os.path.join(_HERE, 'csrc', 'bnpc_kernels.cu')]
HEADERS = [os.path.join(_HERE, 'csrc', 'bnpc_math.cuh'),
           os.path.join(_ROOT, 'include', 'bnpc_b200.h')]

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3',
              '-fmad=false', '-std=c++17', '-shared', '-Xcompiler', '-fPIC']

MAX_EXTRA = 32
ST_K, ST_TDONE, ST_FLAGS, ST_NEXTRA, ST_BIRTHS, ST_MOVED, ST_SLOW = range(7)
ST_NUNC = 10
ST_WORDS = 16
STOP_EXTRA_FULL, STOP_REPACK, STOP_TAPE_EMPTY, STOP_CAPACITY, STOP_HANG = 1, 2, 4, 8, 0x100
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
        ('ll', C.c_void_p), ('ldk', C.c_int32), ('t_epoch0', C.c_int32),
        ('lpx', C.c_void_p), ('llx', C.c_void_p), ('ldx', C.c_int32),
        ('scratch', C.c_void_p),
        ('visit', C.c_void_p), ('cand', C.c_void_p), ('t_begin', C.c_int32), ('t_end', C.c_int32),
        ('visit_c', C.c_void_p), ('cand_c', C.c_void_p),
        ('beta_rows', C.c_void_p), ('n_beta_rows', C.c_int32),
        ('seed', C.c_uint64), ('stream_id', C.c_uint64),
        ('logn', C.c_void_p),
        ('c_norm', C.c_double), ('FN', C.c_double), ('FP', C.c_double), ('p', C.c_double), ('q', C.c_double),
    ]


_P, _I, _D, _U64, _I64, _F = C.c_void_p, C.c_int, C.c_double, C.c_uint64, C.c_int64, C.c_float

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
}

_lock = threading.Lock()
_lib = None
launch_count = 0          # kernels launched through this binding (bench.py reports it)

# kernels launched per entry point (memset nodes are not counted)
_KERNELS_PER_CALL = {'bnpc_gather_members': 3, 'bnpc_rg_sides': 2, 'bnpc_rg_scan': 3, 'bnpc_gibbs_compact': 2}


class _Lib:
    def __init__(self, path):
        self._dll = C.CDLL(path)
        self._dll.bnpc_last_error.restype = C.c_char_p
        self._dll.bnpc_abi_version.restype = C.c_int
        for name, args in SIGNATURES.items():
            fn = getattr(self._dll, name)
            fn.argtypes = args
            fn.restype = C.c_int
            setattr(self, name[len('bnpc_'):], self._wrap(name, fn))

    def _wrap(self, name, fn):
        n_k = _KERNELS_PER_CALL.get(name, 1)

        def call(*args):
            global launch_count
            rc = fn(*args)
            if rc != 0:
                raise RuntimeError(f'{name} failed ({rc}): '
                                   f'{self._dll.bnpc_last_error().decode()}')
            launch_count += n_k
        call.__name__ = name
        return call

    def abi_version(self):
        return self._dll.bnpc_abi_version()


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
