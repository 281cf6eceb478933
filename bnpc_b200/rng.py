"""Random sources of the CUDA-backed model.

Two interchangeable objects provide every draw the MCMC step needs, named by the
reference call site they stand in for:

* `PhiloxRandom(seed)`  -- production.  Bulk draws are generated on the device by
  the library's Philox4x32-10 kernels (counter = element index, stream id =
  running call counter), host scalars by the library's host functions on the same
  key (bnpc_host_random / _gamma / _beta: draw i of a chain is a function of
  (seed, i)).  A chain's stream depends only on its seed, never on which GPU or
  rank runs it, nor on whether this Python mirror or the native group driver
  (bnpc_group_run) steps the chain.
* `TapeRandom(tape)`    -- parity mode.  Replays a recorded sequence of primitive
  numpy-legacy draws ('u', 'int', 'perm', 'beta', 'gamma' records; see
  tests/golden/README.md) in the order the reference consumes them, uploading
  the values the kernels need.

Both return DEVICE tensors for bulk draws and Python scalars for host decisions.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

KINDS = ('u', 'int', 'perm', 'beta', 'gamma')


class TapeError(RuntimeError):
    pass


class Tape:
    """Sequence of (kind, float64 values) records; stored as three flat arrays."""

    def __init__(self, kinds, sizes, values):
        self.kinds = np.asarray(kinds, dtype=np.int8)
        self.sizes = np.asarray(sizes, dtype=np.int64)
        self.values = np.asarray(values, dtype=np.float64)
        self.offsets = np.concatenate(([0], np.cumsum(self.sizes)))
        self.pos = 0

    def take(self, kind, count):
        if self.pos >= self.kinds.size:
            raise TapeError(f'tape exhausted: wanted {kind}[{count}]')
        k = KINDS[self.kinds[self.pos]]
        n = int(self.sizes[self.pos])
        if k != kind or n != count:
            raise TapeError(f'tape misaligned at record {self.pos}: have {k}[{n}], '
                            f'wanted {kind}[{count}]')
        o = self.offsets[self.pos]
        self.pos += 1
        return self.values[o:o + n]

    def next_is(self, kind):
        return self.pos < self.kinds.size and KINDS[self.kinds[self.pos]] == kind

    def exhausted(self):
        return self.pos >= self.kinds.size


def _searchsorted_right(p, u):
    # numpy legacy choice(p=...): cdf = p.cumsum(); cdf /= cdf[-1]; searchsorted(u, 'right')
    cdf = np.cumsum(np.asarray(p, dtype=np.float64))
    cdf /= cdf[-1]
    return cdf.searchsorted(u, side='right')


class _Base:
    device = None

    def bind(self, device):
        self.device = device
        return self

    def _up(self, arr, dtype):
        return torch.as_tensor(np.ascontiguousarray(arr), dtype=dtype, device=self.device)

    # ---- host-side decisions built on the scalar primitives -----------------
    def pick_weighted(self, p):
        """np.random.choice(a, p=p) -> index (libs/CRP.py:427,442,625)."""
        return int(_searchsorted_right(p, self.random()))

    def pick_two_weighted(self, p):
        """np.random.choice(a, p=p, size=2, replace=False) -> two indices
        (libs/CRP.py:490): numpy's draw-and-dedupe loop."""
        p = np.array(p, dtype=np.float64)
        found = []
        while len(found) < 2:
            x = self.uniform_host(2 - len(found))
            if found:
                p[found] = 0
            new = _searchsorted_right(p, x)
            _, first = np.unique(new, return_index=True)
            first.sort()
            found.extend(int(i) for i in new.take(first))
        return found[0], found[1]


class TapeRandom(_Base):
    is_tape = True

    def __init__(self, tape):
        self.tape = tape

    # scalars
    def random(self):
        return float(self.tape.take('u', 1)[0])

    def uniform_host(self, n):
        return self.tape.take('u', n).copy()

    def randint(self, high):
        return int(self.tape.take('int', 1)[0])

    def beta(self, a, b):
        return float(self.tape.take('beta', 1)[0])

    def gamma(self, shape, scale):
        return float(self.tape.take('gamma', 1)[0])

    def first_two_of_permutation(self, n):
        v = self.tape.take('perm', n)
        return int(v[0]), int(v[1])

    def init_labels(self, n):
        return self.tape.take('int', n).astype(np.int64)

    # bulk: host arrays, uploaded by the model into its workspace buffers
    def uniform_rows(self, rows, m):
        return self.tape.take('u', rows * m).copy()

    def beta_rows(self, rows, m):
        """`rows` consecutive Beta draws of length m (libs/CRP.py:172,184)."""
        return np.concatenate([self.tape.take('beta', m) for _ in range(rows)])

    def step_sd_index(self, rows, m):
        return self.tape.take('int', rows * m).copy()

    def mh_theta_draws(self, rows, m):
        """Per row: proposal-sd indices, truncnorm uniforms, acceptance uniforms
        (libs/CRP.py:328,331,335) -> [3][rows][m]."""
        out = np.empty((3, rows, m))
        for r in range(rows):
            out[0, r] = self.tape.take('int', m)
            out[1, r] = self.tape.take('u', m)
            out[2, r] = self.tape.take('u', m)
        return out

    def gibbs_draws(self, n, m):
        """permutation(N), then per visited cell one uniform and -- when that cell
        opened a cluster -- a Beta row (libs/CRP.py:260,277,293)."""
        perm = self.tape.take('perm', n).astype(np.int32)
        u = np.empty(n)
        rows = []
        for t in range(n):
            u[t] = self.tape.take('u', 1)[0]
            if self.tape.next_is('beta'):
                rows.append(self.tape.take('beta', m))
        # a non-NULL (dummy) pointer keeps the kernel in tape mode when no cluster was born
        beta = np.stack(rows) if rows else np.zeros(1)
        return perm, u, beta, len(rows)

    def scan_draws(self, nf):
        """permutation(n-2) and one uniform per free cell (libs/CRP.py:616,625)."""
        perm = self.tape.take('perm', nf).astype(np.int32)
        u = np.concatenate([self.tape.take('u', 1) for _ in range(nf)]) if nf else np.zeros(0)
        return perm, u

    device_seed = 0

    def reserve(self, n):
        return 0


class PhiloxRandom(_Base):
    is_tape = False

    def __init__(self, seed):
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.host_ctr = C.c_uint64(0)        # host scalars drawn so far
        self.calls = 0                       # device streams reserved so far
        self.device_seed = self.seed

    def next_stream(self):
        self.calls += 1
        return self.calls

    def reserve(self, n):
        """n consecutive device stream ids; returns the id BEFORE the first (ids base+1..base+n)"""
        base = self.calls
        self.calls += n
        return base

    def _stream_ptr(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    # scalars
    def random(self):
        return _lib.lib().host_random(self.seed, C.byref(self.host_ctr))

    def uniform_host(self, n):
        return np.array([self.random() for _ in range(n)])

    def randint(self, high):
        return int(np.floor(self.random() * high))

    def beta(self, a, b):
        return _lib.lib().host_beta(self.seed, C.byref(self.host_ctr), float(a), float(b))

    def gamma(self, shape, scale):
        return _lib.lib().host_gamma(self.seed, C.byref(self.host_ctr), float(shape)) * scale

    def first_two_of_permutation(self, n):
        i = self.randint(n)
        j = self.randint(n - 1)
        if j >= i:
            j += 1
        return i, j

    def init_labels(self, n):
        # initial labels are drawn once per chain, from a generator keyed like the chain
        return np.random.Generator(np.random.Philox(key=self.seed)).integers(0, n, size=n)

    # bulk, device
    def _fill(self, count, levels=0):
        out = torch.empty(count, dtype=torch.float64, device=self.device)
        _lib.lib().fill_uniform(out.data_ptr(), count, self.seed, self.next_stream(), levels,
                                self._stream_ptr())
        return out

    def uniform_rows(self, rows, m):
        return self._fill(rows * m)

    def beta_rows(self, rows, m):
        return None                      # the kernel samples Beta variates itself

    def step_sd_index(self, rows, m):
        return self._fill(rows * m, 3)

    def mh_theta_draws(self, rows, m):
        out = torch.empty(3 * rows * m, dtype=torch.float64, device=self.device)
        L = _lib.lib()
        sp = self._stream_ptr()
        L.fill_uniform(out.data_ptr(), rows * m, self.seed, self.next_stream(), 3, sp)
        L.fill_uniform(out.data_ptr() + 8 * rows * m, 2 * rows * m, self.seed, self.next_stream(),
                       0, sp)
        return out

    def gibbs_draws(self, n, m):
        perm = torch.empty(n, dtype=torch.int32, device=self.device)
        _lib.lib().fill_permutation(perm.data_ptr(), n, self.seed, self.next_stream(),
                                    self._stream_ptr())
        return perm, self._fill(n), None, 0

    def scan_draws(self, nf):
        perm = torch.empty(max(nf, 1), dtype=torch.int32, device=self.device)
        _lib.lib().fill_permutation(perm.data_ptr(), nf, self.seed, self.next_stream(),
                                    self._stream_ptr())
        return perm, self._fill(max(nf, 1))
