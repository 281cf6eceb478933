"""TEST INFRASTRUCTURE -- not part of the product path.

Random "tape" for the BnpC hot path: a numpy-legacy-compatible random source
whose primitive draws can be recorded and replayed.

Why: every random draw of the reference comes from numpy's global legacy
RandomState (MT19937) -- directly (`np.random.permutation/choice/beta/random/
randint/uniform/gamma`, reference libs/CRP.py:140,178,184,260,277,328,335,394,
400-406,427,442,448,490,495,499,616,625,650,662,674,778 and
libs/CRP_learning_errors.py:78,108, libs/MCMC.py:322,332,339) or through
`scipy.stats.truncnorm.rvs(random_state=None)` (libs/CRP.py:331,
libs/CRP_learning_errors.py:82).  Parity of the CUDA path is defined on "the same
chain state and the same injected random stream", so the stream has to be
something all three implementations (reference, oracle restatement, CUDA path)
can consume.  `choice` is decomposed into the primitive draws numpy's legacy
implementation makes internally (one `random_sample` for p-weighted draws,
`randint` for uniform ones, `permutation` for replace=False, the iterative
draw-and-dedupe loop for weighted replace=False), so the tape holds primitives
only:

    kind   payload
    'u'    uniform(0,1) doubles            (random/random_sample/rand/uniform)
    'int'  integers                        (randint)
    'perm' a permutation of range(n)       (permutation)
    'beta' Beta variates                   (beta)
    'gamma' Gamma variates                 (gamma)

Beta/Gamma/permutation consume a data-dependent number of MT19937 words, so they
are taped as VALUES.
"""
import numpy as np

KINDS = ('u', 'int', 'perm', 'beta', 'gamma')
_KIND_ID = {k: i for i, k in enumerate(KINDS)}


class Tape:
    """An ordered list of (kind, float64 array) records with a read cursor."""

    def __init__(self, records=None):
        self.records = list(records) if records is not None else []
        self.pos = 0

    def append(self, kind, values):
        self.records.append((kind, np.array(values, dtype=np.float64).ravel()))

    def take(self, kind, count):
        if self.pos >= len(self.records):
            raise IndexError(f'tape exhausted: wanted {kind}[{count}]')
        k, v = self.records[self.pos]
        if k != kind or v.size != count:
            raise ValueError(
                f'tape misaligned at record {self.pos}: have {k}[{v.size}], '
                f'wanted {kind}[{count}]')
        self.pos += 1
        return v

    def peek_kind(self):
        if self.pos >= len(self.records):
            return None
        return self.records[self.pos][0]

    def exhausted(self):
        return self.pos >= len(self.records)

    def rewind(self):
        self.pos = 0

    def slice_from(self, start):
        return Tape(self.records[start:])

    # -- (de)serialisation: three flat arrays, usable inside an .npz ----------
    def to_arrays(self):
        kinds = np.array([_KIND_ID[k] for k, _ in self.records], dtype=np.int8)
        sizes = np.array([v.size for _, v in self.records], dtype=np.int64)
        if self.records:
            values = np.concatenate([v for _, v in self.records])
        else:
            values = np.zeros(0)
        return kinds, sizes, values

    @classmethod
    def from_arrays(cls, kinds, sizes, values):
        recs = []
        off = 0
        for k, n in zip(kinds, sizes):
            recs.append((KINDS[int(k)], np.array(values[off:off + n], dtype=np.float64)))
            off += int(n)
        return cls(recs)


class GlobalLegacySource:
    """Primitive draws from numpy's global legacy RandomState."""

    def uniform01(self, size):
        return np.random.random_sample(size)

    def randint(self, low, high, size):
        return np.random.randint(low, high, size)

    def permutation(self, n):
        return np.random.permutation(n)

    def beta(self, a, b):
        return np.random.beta(a, b)

    def gamma(self, shape, scale):
        return np.random.gamma(shape, scale)


class TapeSource:
    """Primitive draws replayed from a Tape."""

    def __init__(self, tape):
        self.tape = tape

    @staticmethod
    def _count(size):
        if size is None:
            return 1
        return int(np.prod(size))

    def _shape(self, v, size):
        if size is None:
            return v[0]
        return v.reshape(size)

    def uniform01(self, size):
        return self._shape(self.tape.take('u', self._count(size)), size)

    def randint(self, low, high, size):
        v = self.tape.take('int', self._count(size)).astype(np.int64)
        return self._shape(v, size)

    def permutation(self, n):
        return self.tape.take('perm', int(n)).astype(np.int64)

    def beta(self, a, b):
        shape = np.broadcast(np.asarray(a), np.asarray(b)).shape
        v = self.tape.take('beta', int(np.prod(shape)) if shape else 1)
        return v.reshape(shape) if shape else v[0]

    def gamma(self, shape, scale):
        return self.tape.take('gamma', 1)[0]


class LegacyRandom:
    """Drop-in for the subset of `numpy.random` the hot path uses.

    `source` provides primitives; if `record` is a Tape every primitive draw is
    appended to it.  `choice` follows numpy's legacy RandomState.choice
    algorithm draw for draw (checked against the real thing in
    tests/test_rng_tape.py).
    """

    def __init__(self, source=None, record=None):
        self.source = source if source is not None else GlobalLegacySource()
        self.record = record

    def _rec(self, kind, v):
        if self.record is not None:
            self.record.append(kind, v)
        return v

    # -- primitives ----------------------------------------------------------
    def random_sample(self, size=None):
        return self._rec('u', self.source.uniform01(size))

    random = random_sample

    def rand(self, *shape):
        return self.random_sample(shape if shape else None)

    def uniform(self, low=0.0, high=1.0, size=None):
        if low != 0.0 or high != 1.0:
            raise NotImplementedError('only uniform(0,1) is on the hot path')
        return self.random_sample(size)

    def randint(self, low, high=None, size=None):
        if high is None:
            low, high = 0, low
        return self._rec('int', self.source.randint(low, high, size))

    def permutation(self, n):
        return self._rec('perm', self.source.permutation(int(n)))

    def beta(self, a, b, size=None):
        if size is not None:
            raise NotImplementedError
        return self._rec('beta', self.source.beta(a, b))

    def gamma(self, shape, scale=1.0, size=None):
        if size is not None:
            raise NotImplementedError
        return self._rec('gamma', self.source.gamma(shape, scale))

    def seed(self, s=None):
        np.random.seed(s)

    # -- numpy legacy `choice`, decomposed into primitives --------------------
    def choice(self, a, size=None, replace=True, p=None):
        a = np.asarray(a)
        if a.ndim == 0:
            pop_size = int(a)
            population = None
        else:
            pop_size = a.shape[0]
            population = a
        if p is not None:
            p = np.array(p, dtype=np.float64)
        shape = size
        if shape is not None:
            total = int(np.prod(shape))
        else:
            total = 1

        if replace:
            if p is not None:
                cdf = p.cumsum()
                cdf /= cdf[-1]
                u = self.random_sample(shape)
                idx = cdf.searchsorted(u, side='right')
                idx = np.asarray(idx)
            else:
                idx = np.asarray(self.randint(0, pop_size, size=shape))
        else:
            if total > pop_size:
                raise ValueError('Cannot take a larger sample than population')
            if p is not None:
                n_uniq = 0
                p = p.copy()
                found = np.zeros(shape, dtype=np.int64)
                flat_found = found.ravel()
                while n_uniq < total:
                    x = self.rand(total - n_uniq)
                    if n_uniq > 0:
                        p[flat_found[0:n_uniq]] = 0
                    cdf = np.cumsum(p)
                    cdf /= cdf[-1]
                    new = cdf.searchsorted(x, side='right')
                    _, unique_indices = np.unique(new, return_index=True)
                    unique_indices.sort()
                    new = new.take(unique_indices)
                    flat_found[n_uniq:n_uniq + new.size] = new
                    n_uniq += new.size
                idx = found
            else:
                idx = self.permutation(pop_size)[:total]
                if shape is not None:
                    idx = idx.reshape(shape)

        if shape is None and isinstance(idx, np.ndarray):
            idx = idx.item() if idx.ndim == 0 else idx
        if population is None:
            return idx
        return population[idx]

    # -- scipy truncnorm.rvs(random_state=None), decomposed -------------------
    def truncnorm_rvs(self, a, b, loc=0.0, scale=1.0, size=None):
        """scipy rv_continuous.rvs for truncnorm: `_ppf(uniform(size)) * scale
        + loc` (scipy/stats/_distn_infrastructure.py rvs/_rvs; truncnorm has no
        `_rvs` override), one uniform per variate."""
        from scipy.stats import truncnorm
        if size is None:
            shape = np.broadcast(np.asarray(a), np.asarray(b), np.asarray(loc),
                                 np.asarray(scale)).shape
        else:
            shape = (size,) if np.isscalar(size) else tuple(size)
        u = self.random_sample(shape if shape else None)
        vals = truncnorm._ppf(np.asarray(u, dtype=np.float64),
                              np.asarray(a, dtype=np.float64),
                              np.asarray(b, dtype=np.float64))
        vals = vals * scale + loc
        if shape == ():
            vals = np.asarray(vals)[()]
        return vals
