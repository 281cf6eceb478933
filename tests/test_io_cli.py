"""CPU: the input loader and the command line (libs/dpmmIO.py, run_BnpC.py) against the reference's
own loader / argument parser where the reference is mounted, and on generated files."""
import os
import sys

import numpy as np
import pytest

from oracle import ref_shim
import libs.dpmmIO as io
import run_BnpC


def _write(path, mat, sep, header, index):
    rows, cols = mat.shape
    with open(path, 'w') as f:
        if header:
            f.write(sep.join(([''] if index else []) + [f'cell{j}' for j in range(cols)]) + '\n')
        for i in range(rows):
            vals = [str(int(v)) for v in mat[i]]
            f.write(sep.join(([f'mut{i}'] if index else []) + vals) + '\n')


@pytest.mark.parametrize('sep', ['\t', ',', ' '])
@pytest.mark.parametrize('header,index', [(False, False), (True, True), (True, False)])
def test_load_data_formats(tmp_path, sep, header, index):
    rng = np.random.default_rng(1)
    mat = rng.choice([0, 1, 2, 3], size=(7, 9), p=[0.5, 0.3, 0.05, 0.15])     # mutations x cells
    path = str(tmp_path / 'm.csv')
    _write(path, mat, sep, header, index)
    got, names = io.load_data(path, transpose=True, get_names=True)
    want = mat.T.astype(float)
    want[want == 3] = np.nan
    want[want == 2] = 1
    np.testing.assert_array_equal(got, want)
    assert got.shape == (9, 7) and names[0].size == 9 and names[1].size == 7
    np.testing.assert_array_equal(io.load_data(path, transpose=False), want.T)


@pytest.mark.skipif(not ref_shim.reference_available(), reason='reference checkout not mounted')
def test_load_data_matches_reference_on_its_example():
    ref = ref_shim.load_reference(with_mcmc=True)
    ref_io = sys.modules.get('libs.dpmmIO')
    path = os.path.join(ref_shim.REF_ROOT, 'example_data', 'data.csv')
    got, names = io.load_data(path, transpose=True, get_names=True)
    # the reference's loader lives in its own libs.dpmmIO; import it the way the shim does
    import importlib
    hidden = {k: sys.modules.pop(k) for k in list(sys.modules) if k == 'libs' or k.startswith('libs.')}
    sys.modules.update({'bottleneck': ref_shim._bottleneck_standin(), **ref_shim._stub_plot_modules()})
    sys.path.insert(0, ref_shim.REF_ROOT)
    old = np.geterr()
    try:
        rio = importlib.import_module('libs.dpmmIO')
        want, want_names = rio.load_data(path, transpose=True, get_names=True)
    finally:
        sys.path.remove(ref_shim.REF_ROOT)
        for k in list(sys.modules):
            if k == 'libs' or k.startswith('libs.') or k in ('bottleneck', 'seaborn') or k.startswith('matplotlib'):
                del sys.modules[k]
        sys.modules.update(hidden)
        np.seterr(**old)
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(names[0], want_names[0])
    np.testing.assert_array_equal(names[1], want_names[1])
    del ref, ref_io


def test_cli_defaults_are_the_references():
    a = run_BnpC.parse_args(['x.csv'])
    assert (a.chains, a.steps, a.runtime, a.lugsail, a.burn_in) == (1, 5000, -1, -1, 0.33)
    assert (a.conc_update_prob, a.error_update_prob, a.split_merge_prob, a.split_merge_steps) == (0.25, 0.25, 0.33, 3)
    assert a.split_merge_ratios == [0.75, 0.25] and a.param_prior == [.25, .25] and a.DPa_prior == [-1, -1]
    assert (a.falseNegative, a.falsePositive) == (-1, -1)
    assert (a.falseNegative_mean, a.falseNegative_std, a.falsePositive_mean, a.falsePositive_std) == (0.2, 0.1, 0.01, 0.01)
    assert a.estimator == 'posterior' and a.transpose is True and a.seed == -1
    b = run_BnpC.parse_args(['x.csv', '-n', '8', '-s', '100', '-e', 'posterior', 'MAP', '-smp', '0.75', '-pp', '1', '1',
                             '-FN', '0.3', '-FP', '0.0001', '-np'])
    assert b.chains == 8 and b.estimator == ['posterior', 'MAP'] and b.param_prior == [1.0, 1.0] and b.no_plots
    with pytest.raises(SystemExit):
        run_BnpC.parse_args(['x.csv', '-ls', '2.0'])


def test_termination_and_writers(tmp_path):
    import argparse
    from datetime import datetime
    import pandas as pd
    args = argparse.Namespace(runtime=-1, lugsail=-1, steps=900, burn_in=0.33, time=[datetime(2026, 1, 1)])
    assert io._get_mcmc_termination(args) == ((900, 297), 'for 900 steps')
    args.lugsail = 1.05
    assert io._get_mcmc_termination(args)[0] == (1.05, 0)
    geno = pd.DataFrame(np.array([[0.9, 0.1], [0.2, 0.8], [1.0, 0.0]]))
    inferred = {'mean': {'MAP': dict(step=5, a=3.0, assignment=[0, 1], genotypes=geno, FN=np.float64(0.2),
                                     FP=np.float64(0.01), FN_geno=np.float64(0.1), FP_geno=np.float64(0.02))}}
    ns = argparse.Namespace(estimator=['MAP'], chains=1, time=[datetime(2026, 1, 1)], falseNegative=-1,
                            falsePositive=-1, falseNegative_mean=.2, falseNegative_std=.1, falsePositive_mean=.01,
                            falsePositive_std=.01)
    io.save_run(inferred, ns, str(tmp_path), (np.array(['c0', 'c1']), np.array(['m0', 'm1', 'm2'])))
    assert io.load_txt(str(tmp_path / 'assignment.txt')) == [0, 1]
    assert sorted(os.listdir(tmp_path)) == ['args.txt', 'assignment.txt', 'errors.txt', 'genotypes_MAP_mean.tsv',
                                            'genotypes_cont_MAP_mean.tsv']
    io.save_ARI(inferred, [0, 1], str(tmp_path))
    assert float(pd.read_csv(tmp_path / 'ARI.txt', sep='\t')['ARI'][0]) == 1.0


def test_packed_cache_replaces_the_text_parse(tmp_path, monkeypatch):
    """the loader keeps the parsed matrix next to the input as two packed bit-planes and reads it
    back while the text file is unchanged"""
    rng = np.random.default_rng(2)
    mat = rng.choice([0, 1, 2, 3], size=(11, 37), p=[0.5, 0.3, 0.05, 0.15])
    path = str(tmp_path / 'm.csv')
    _write(path, mat, '\t', True, True)
    first, names = io.load_data(path, transpose=True, get_names=True)
    assert os.path.exists(path + io.CACHE_SUFFIX)
    monkeypatch.setattr(io.pd, 'read_csv', lambda *a, **k: (_ for _ in ()).throw(AssertionError('text parse')))
    again, names2 = io.load_data(path, transpose=True, get_names=True)
    np.testing.assert_array_equal(first, again)
    assert list(names[0]) == list(names2[0]) and list(names[1]) == list(names2[1])
    np.testing.assert_array_equal(io.load_data(path, transpose=False), first.T)
    monkeypatch.undo()
    # a changed input invalidates the cache
    mat2 = mat.copy()
    mat2[0, 0] = 1 - min(mat2[0, 0], 1)
    _write(path, mat2, '\t', True, True)
    os.utime(path, ns=(os.stat(path).st_atime_ns, os.stat(path).st_mtime_ns + 10 ** 9))
    changed = io.load_data(path, transpose=True)
    want = mat2.T.astype(float)
    want[want == 3] = np.nan
    want[want == 2] = 1
    np.testing.assert_array_equal(changed, want)
    # and it can be switched off
    monkeypatch.setenv('BNPC_NO_CACHE', '1')
    os.remove(path + io.CACHE_SUFFIX)
    io.load_data(path)
    assert not os.path.exists(path + io.CACHE_SUFFIX)
