"""TEST INFRASTRUCTURE -- not part of the product path.

Imports the UNMODIFIED reference (cbg-ethz/BnpC, read-only at /root/reference)
in this container so that (a) the CPU restatement in oracle/crp_oracle.py can be
pinned against it and (b) golden vectors can be generated from it
(tests/golden/make_golden.py).  /root/reference does not exist on the GPU box;
an unmodified copy travels there under baseline/_ref (git-ignored, made by
oracle/fetch_ref.py), which `bench.py --impl reference` times on the host cores.
Nothing in the product path imports this module.

The reference needs `bottleneck` (not installed; requirements.txt:1) -- a
6-function numpy stand-in is injected for the duration of the import only
(SURVEY.md Appendix A).  The reference's package is called `libs`, the same name
as this repo's drop-in package, so it is imported with this repo's `libs`
temporarily hidden from sys.modules.
"""
import importlib
import os
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIPPED = os.path.join(os.path.dirname(_HERE), 'baseline', '_ref', 'BnpC')      # oracle/fetch_ref.py


def _find_root():
    """the reference checkout: $BNPC_REFERENCE_ROOT, the build container's /root/reference, or the
    unmodified copy shipped to the GPU box under baseline/_ref (git-ignored)"""
    for cand in (os.environ.get('BNPC_REFERENCE_ROOT'), '/root/reference', _SHIPPED):
        if cand and os.path.isfile(os.path.join(cand, 'libs', 'CRP.py')):
            return cand
    return '/root/reference'


REF_ROOT = _find_root()


def reference_available():
    return os.path.isfile(os.path.join(REF_ROOT, 'libs', 'CRP.py'))


def _bottleneck_standin():
    bn = types.ModuleType('bottleneck')
    bn.__version__ = '0.0.0'
    bn.nansum = np.nansum
    bn.nanargmax = np.nanargmax
    bn.nanmean = np.nanmean

    def nanvar(a, axis=None, ddof=0):
        return np.nanvar(a, axis=axis, ddof=ddof)

    def replace(a, old, new):
        if old != old:
            a[np.isnan(a)] = new
        else:
            a[a == old] = new

    def move_std(a, window, axis=-1, ddof=0):
        a = np.asarray(a, dtype=np.float64)
        out = np.full(a.shape, np.nan)
        a_m = np.moveaxis(a, axis, -1)
        out_m = np.moveaxis(out, axis, -1)
        for i in range(window - 1, a_m.shape[-1]):
            out_m[..., i] = a_m[..., i - window + 1:i + 1].std(axis=-1, ddof=ddof)
        return out

    bn.nanvar = nanvar
    bn.replace = replace
    bn.move_std = move_std
    return bn


def _stub_plot_modules():
    mods = {}
    mpl = types.ModuleType('matplotlib')
    mpl.use = lambda *a, **k: None
    mpl.__path__ = []
    plt = types.ModuleType('matplotlib.pyplot')
    gs = types.ModuleType('matplotlib.gridspec')
    gs.GridSpec = object
    tick = types.ModuleType('matplotlib.ticker')
    tick.MaxNLocator = object
    mpl.pyplot = plt
    mpl.gridspec = gs
    mpl.ticker = tick
    sns = types.ModuleType('seaborn')
    for name, m in (('matplotlib', mpl), ('matplotlib.pyplot', plt),
                    ('matplotlib.gridspec', gs), ('matplotlib.ticker', tick),
                    ('seaborn', sns)):
        mods[name] = m
    return mods


_CACHE = {}


def load_reference(with_mcmc=False):
    """Return a namespace with the reference modules: .CRP, .CRP_learning_errors
    (and .MCMC when with_mcmc).  Modules are private copies: they are removed
    from sys.modules again so that this repo's own `libs` package stays
    importable afterwards."""
    key = bool(with_mcmc)
    if key in _CACHE:
        return _CACHE[key]
    if not reference_available():
        raise FileNotFoundError(f'reference not found under {REF_ROOT}')

    if with_mcmc:
        import pandas  # noqa: F401  (must be imported BEFORE the bottleneck stand-in exists)
        import sklearn.metrics  # noqa: F401
        import scipy.cluster.hierarchy  # noqa: F401

    hidden = {k: sys.modules.pop(k) for k in list(sys.modules)
              if k == 'libs' or k.startswith('libs.')}
    injected = {}
    if 'bottleneck' not in sys.modules:
        injected['bottleneck'] = _bottleneck_standin()
    if with_mcmc:
        for name, m in _stub_plot_modules().items():
            if name not in sys.modules:
                injected[name] = m
    sys.modules.update(injected)
    sys.path.insert(0, REF_ROOT)
    old_err = np.geterr()
    try:
        ns = types.SimpleNamespace()
        ns.CRP = importlib.import_module('libs.CRP')
        ns.CRP_learning_errors = importlib.import_module('libs.CRP_learning_errors')
        if with_mcmc:
            ns.MCMC = importlib.import_module('libs.MCMC')
            ns.utils = importlib.import_module('libs.utils')
    finally:
        sys.path.remove(REF_ROOT)
        for k in list(sys.modules):
            if k == 'libs' or k.startswith('libs.'):
                del sys.modules[k]
        for k in injected:
            sys.modules.pop(k, None)
        sys.modules.update(hidden)
        # the reference calls np.seterr(divide='raise', invalid='raise') at import
        # (libs/CRP.py:10); do not leak that into the test process.
        np.seterr(**old_err)
    _CACHE[key] = ns
    return ns


class _NumpyWithRandom:
    """`np` as seen by the reference modules, with `.random` swapped."""

    def __init__(self, rnd):
        self.random = rnd

    def __getattr__(self, name):
        return getattr(np, name)


class _TruncnormWithRvs:
    """scipy.stats.truncnorm with `.rvs` routed through the tape source."""

    def __init__(self, rnd):
        from scipy.stats import truncnorm
        self._tn = truncnorm
        self._rnd = rnd

    def rvs(self, a, b, loc=0.0, scale=1.0, size=None):
        return self._rnd.truncnorm_rvs(a, b, loc=loc, scale=scale, size=size)

    def __call__(self, *args, **kw):
        return self._tn(*args, **kw)

    def __getattr__(self, name):
        return getattr(self._tn, name)


class patched_random:
    """Context manager: route every random draw of the reference's model modules
    through `rnd` (an oracle.rng_tape.LegacyRandom)."""

    def __init__(self, ref, rnd):
        self.ref = ref
        self.rnd = rnd
        self.saved = []

    def __enter__(self):
        mods = [self.ref.CRP, self.ref.CRP_learning_errors]
        if hasattr(self.ref, 'MCMC'):
            mods.append(self.ref.MCMC)
        for m in mods:
            self.saved.append((m, 'np', m.np))
            m.np = _NumpyWithRandom(self.rnd)
            if hasattr(m, 'truncnorm'):
                self.saved.append((m, 'truncnorm', m.truncnorm))
                m.truncnorm = _TruncnormWithRvs(self.rnd)
        return self

    def __exit__(self, *exc):
        for m, name, val in self.saved:
            setattr(m, name, val)
        self.saved = []
        return False


class ref_errstate:
    """The floating-point error mode the reference runs under (libs/CRP.py:10)."""

    def __enter__(self):
        self.ctx = np.errstate(divide='raise', over='ignore', under='ignore',
                               invalid='raise')
        return self.ctx.__enter__()

    def __exit__(self, *exc):
        return self.ctx.__exit__(*exc)
