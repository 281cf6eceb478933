"""Shared test utilities: golden fixtures, tape conversion, state comparison."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def golden_names():
    """chain fixtures (tests/golden/make_golden.py)"""
    return sorted(n for n in (os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz')))
                  if not n.startswith('estimators_'))


def estimator_golden_names():
    """estimator fixtures (tests/golden/make_golden_estimators.py)"""
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, 'estimators_*.npz')))


class EstimatorGolden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'), allow_pickle=False)
        self.z = z
        code = z['data']
        self.data = np.where(code < 0, np.nan, code.astype(np.float64))
        self.results = []
        for c in range(int(z['n_chains'])):
            r = {k: z[f'chain{c}_{k}'] for k in ('assignments', 'params', 'ML', 'MAP', 'DP_alpha', 'FN', 'FP')}
            r['burn_in'] = int(z[f'chain{c}_burn_in'])
            self.results.append(r)


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'), allow_pickle=False)
        self.name = name
        self.meta = json.loads(str(z['meta']))
        self.steplog = json.loads(str(z['steplog']))
        code = z['code']
        self.data = np.where(code == 1, 1.0, np.where(code == 0, 0.0, np.nan))
        self.tape_arrays = (z['tape_kinds'], z['tape_sizes'], z['tape_values'])
        self.tape_pos = z['tape_pos']
        self.n_states = z['assignment'].shape[0]
        self._z = z

    def state(self, i):
        z = self._z
        k = int(z['n_clusters'][i])
        return dict(assignment=z['assignment'][i], ids=z['ids'][i, :k], sizes=z['sizes'][i, :k],
                    theta=z['theta'][i, :k], alpha=float(z['alpha'][i]), FN=float(z['FN'][i]),
                    FP=float(z['FP'][i]), ll=float(z['ll'][i]), lpost=float(z['lpost'][i]))

    def oracle_tape(self):
        from oracle.rng_tape import Tape
        return Tape.from_arrays(*self.tape_arrays)

    def product_tape(self):
        from bnpc_b200.rng import Tape
        return Tape(*self.tape_arrays)


def assert_state(got, want, where, exact_float=True, rtol=1e-9):
    """Decisions (assignment, cluster list, sizes) must be identical; float32 theta identical;
    float64 scalars identical (oracle) or within rtol (CUDA path)."""
    np.testing.assert_array_equal(got['ids'], want['ids'], err_msg=f'{where}: cluster ids/order')
    np.testing.assert_array_equal(got['sizes'], want['sizes'], err_msg=f'{where}: cluster sizes')
    np.testing.assert_array_equal(got['assignment'], want['assignment'], err_msg=f'{where}: assignment')
    np.testing.assert_array_equal(got['theta'], want['theta'], err_msg=f'{where}: theta')
    for k in ('alpha', 'FN', 'FP', 'll', 'lpost'):
        if exact_float:
            assert got[k] == want[k], f'{where}: {k} {got[k]!r} != {want[k]!r}'
        else:
            np.testing.assert_allclose(got[k], want[k], rtol=rtol, atol=0, err_msg=f'{where}: {k}')
