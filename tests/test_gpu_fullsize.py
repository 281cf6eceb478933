"""Full-size checks at the BASELINE.json shapes (C2 10k x 500, C3 100k x 1k, C4 50k x 5k fixed
error rates / split-merge heavy, C5 1M x 50 panel with 30 % missing), where the CPU oracle would
take hours per step.  Size-independent properties instead:

* the six independent CUDA routes of the Gibbs sweep -- integer tcgen05 rows + one sequencer
  warp per component group (production), FP32-FMA rows, bf16-split tcgen05 rows, the dense FP64
  matrix, one sequencer warp, option weights walked with lanes <-> clusters -- must
  produce the SAME chain from the same Philox seed (bit-identical assignments, cluster lists and
  float32 parameters; the dense FP64 route is the arithmetic pinned against the oracle);
* bookkeeping invariants after every step (sizes = bincount of the assignment, ordered ids);
* `get_ll_full` (device, from the [K,M] sufficient statistics) equals a host float64 evaluation
  of the reference formula (libs/CRP.py:197-212, 237-238) on the full matrix, 1e-9 relative;
* the chain started at the simulated truth stays there (ARI > 0.9).
"""
import zlib

import numpy as np
import pytest

from bnpc_b200.synth import CONFIGS, make_matrix
from oracle.crp_oracle import DEFAULT_MOVES, do_step

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')

STEPS = 5
MODES = {
    'production': {},
    'fma_rows': dict(lean_rows=1),
    'bf16_rows': dict(lean_rows=2),
    'dense_fp64': dict(lean_enabled=False),
    'serial_sweep': dict(serial_sweep=True),
    'wide': dict(force_wide=True),
}


def _model(cfg, data, seed):
    from bnpc_b200.rng import PhiloxRandom
    import libs.CRP as crp
    import libs.CRP_learning_errors as crple
    c = CONFIGS[cfg]
    rnd = PhiloxRandom(seed)
    if c['learning']:
        m = crple.CRP_errors_learning(data, DP_alpha=[-1, -1], param_beta=list(c['pp']), FP_mean=0.01,
                                      FP_sd=0.01, FN_mean=0.2, FN_sd=0.1, rnd=rnd)
    else:
        m = crp.CRP(data, DP_alpha=[-1, -1], param_beta=list(c['pp']), FN_error=c['FN'], FP_error=c['FP'],
                    rnd=rnd)
    return m, rnd


def _host_ll(data, assign, ids, theta, FN, FP):
    """sum_n ll[n, z_n] in float64 on the host: per cluster column sums of ones / zeros, then the
    reference's per-entry terms with its float32 (1 - theta) (libs/CRP.py:197-212)."""
    total = 0.0
    for row, cid in enumerate(ids):
        x = data[assign == cid]
        s1 = np.nansum(x, axis=0)
        s0 = np.sum(x == 0, axis=0).astype(np.float64)
        t = theta[row]                                 # float32
        one_minus = (1 - t).astype(np.float64)         # rounded to float32 first
        t64 = t.astype(np.float64)
        lp1 = np.log(t64 * (1 - FN) + one_minus * FP)
        lp0 = np.log(t64 * FN + one_minus * (1 - FP))
        total += float(np.sum(s1 * lp1) + np.sum(s0 * lp0))
    return total


def _run(cfg, data, z, mode, monkeypatch):
    from bnpc_b200.engine import DeviceCRP
    for k, v in MODES[mode].items():
        monkeypatch.setattr(DeviceCRP, k, v)
    c = CONFIGS[cfg]
    moves = dict(DEFAULT_MOVES, sm_prob=c.get('sm_prob', 0.33))
    m, rnd = _model(cfg, data, seed=77)
    m.init(assign=[int(v) for v in z])
    N = data.shape[0]
    trace = []
    for s in range(STEPS):
        log = do_step(m, rnd, moves, c['learning'])
        a = m.assignment
        ids = np.fromiter(m.cells_per_cluster.keys(), dtype=np.int64)
        sizes = np.fromiter(m.cells_per_cluster.values(), dtype=np.int64)
        # bookkeeping invariants
        assert sizes.sum() == N and (sizes > 0).all(), f'{cfg} {mode} step {s + 1}: sizes'
        cnt = np.bincount(a, minlength=int(ids.max()) + 1)
        np.testing.assert_array_equal(cnt[ids], sizes, err_msg=f'{cfg} {mode} step {s + 1}: counts')
        assert cnt.sum() == cnt[ids].sum(), f'{cfg} {mode} step {s + 1}: a cell sits in a dead cluster'
        theta = m.parameters[ids]
        trace.append(dict(log=log, crc=zlib.crc32(np.ascontiguousarray(a).tobytes()), ids=ids.copy(),
                          sizes=sizes.copy(), theta_crc=zlib.crc32(np.ascontiguousarray(theta).tobytes()),
                          ll=m.get_ll_full(), lprior=m.get_lprior_full(), alpha=m.DP_a, FN=m.FN, FP=m.FP,
                          stats=dict(m.sweep_stats)))
    final = dict(assign=a, ids=ids, theta=theta, FN=float(m.FN), FP=float(m.FP), ll=trace[-1]['ll'])
    monkeypatch.undo()
    return trace, final


@pytest.mark.parametrize('cfg', ['C2', 'C3', 'C4', 'C5'])
def test_full_size_routes_agree(cfg, monkeypatch):
    from sklearn.metrics import adjusted_rand_score
    c = CONFIGS[cfg]
    data, z = make_matrix(c['cells'], c['muts'], c['k_true'], c['fn'], c['fp'], c['miss'], seed=0)
    ref_trace, ref_final = _run(cfg, data, z, 'production', monkeypatch)
    # host float64 evaluation of the full-data log-likelihood
    want = _host_ll(data, ref_final['assign'], ref_final['ids'], ref_final['theta'], ref_final['FN'],
                    ref_final['FP'])
    np.testing.assert_allclose(ref_final['ll'], want, rtol=1e-9, err_msg=f'{cfg}: get_ll_full vs host float64')
    assert adjusted_rand_score(z, ref_final['assign']) > 0.9
    assert ref_trace[-1]['stats'].get('uncertain', 0) <= c['cells']
    for mode in ('fma_rows', 'bf16_rows', 'dense_fp64', 'serial_sweep', 'wide'):
        trace, final = _run(cfg, data, z, mode, monkeypatch)
        for s, (g, w) in enumerate(zip(trace, ref_trace)):
            where = f'{cfg} {mode} vs production, step {s + 1}'
            assert g['log'] == w['log'], where
            np.testing.assert_array_equal(g['ids'], w['ids'], err_msg=where)
            np.testing.assert_array_equal(g['sizes'], w['sizes'], err_msg=where)
            assert g['crc'] == w['crc'], f'{where}: assignment'
            assert g['theta_crc'] == w['theta_crc'], f'{where}: theta'
            for k in ('ll', 'lprior', 'alpha', 'FN', 'FP'):
                np.testing.assert_allclose(g[k], w[k], rtol=1e-12, err_msg=f'{where}: {k}')
        np.testing.assert_array_equal(final['assign'], ref_final['assign'])
