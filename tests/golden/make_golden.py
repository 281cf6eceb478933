#!/usr/bin/env python3
"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container, where the unmodified reference is mounted read-only at
/root/reference:

    python tests/golden/make_golden.py

For every scenario the reference classes (libs/CRP.py `CRP`, libs/CRP_learning_errors.py
`CRP_errors_learning`) are constructed on a small seeded matrix, initialised and stepped
in the order of libs/MCMC.py:320-342 while every random draw is recorded
(oracle/rng_tape.py, oracle/ref_shim.py).  Each fixture holds

    code            int8 [N,M]   the input matrix (1, 0, -1 = missing)
    meta            JSON         constructor kwargs, move probabilities, init mode
    tape_kinds/sizes/values      the recorded primitive draws, in order
    tape_pos        int [S+2]    tape record index at: start, after init, after each step
    assignment      int [S+1,N]  state after init and after every step
    n_clusters      int [S+1]
    ids, sizes      int [S+1,Kmax]     live cluster ids / sizes in dict order (-1 padded)
    theta           f32 [S+1,Kmax,M]   parameters of the live clusters, same order
    alpha, FN, FP, ll, lpost     f64 [S+1]
    steplog         JSON         what each step did (move type, accept counters)

The fixtures pin three things: the oracle restatement (tests/test_oracle_golden.py,
CPU), the CUDA path (tests/test_gpu_parity.py, GPU) and the tape format.
The reference needs `bottleneck`, which is not installed: a numpy stand-in is used
(oracle/ref_shim.py); sums are pairwise (numpy) instead of sequential (bottleneck).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_shim, rng_tape  # noqa: E402
from oracle.crp_oracle import do_step, simulate, snapshot  # noqa: E402

SCENARIOS = [
    dict(name='learn_pp025_random', n=40, m=24, k=4, miss=0.10, seed=1, steps=12, learning=True,
         pp=[0.25, 0.25], init='random', np_seed=101,
         moves=dict(sm_prob=0.33, dpa_prob=0.25, error_prob=0.25, sm_ratios=[0.75, 0.25], sm_steps=3)),
    dict(name='fixed_pp11_assign_smheavy', n=48, m=40, k=3, miss=0.10, seed=2, steps=12,
         learning=False, pp=[1, 1], init='assign', np_seed=102, FN=0.3, FP=0.0001,
         moves=dict(sm_prob=0.75, dpa_prob=0.25, error_prob=0.0, sm_ratios=[0.75, 0.25], sm_steps=3)),
    dict(name='learn_pp11_panel_missing30', n=64, m=16, k=3, miss=0.30, seed=3, steps=10,
         learning=True, pp=[1, 1], init='random', np_seed=103,
         moves=dict(sm_prob=0.5, dpa_prob=0.5, error_prob=0.5, sm_ratios=[0.5, 0.5], sm_steps=2)),
    dict(name='learn_pp025_wide150', n=36, m=150, k=3, miss=0.10, seed=4, steps=6, learning=True,
         pp=[0.25, 0.25], init='assign', np_seed=104,
         moves=dict(sm_prob=0.5, dpa_prob=0.25, error_prob=0.25, sm_ratios=[0.75, 0.25], sm_steps=3)),
    dict(name='fixed_pp025_ragged70', n=33, m=70, k=5, miss=0.15, seed=5, steps=8, learning=False,
         pp=[0.25, 0.25], init='random', np_seed=105, FN=0.2, FP=0.01,
         moves=dict(sm_prob=0.4, dpa_prob=0.3, error_prob=0.0, sm_ratios=[0.75, 0.25], sm_steps=3)),
]


def model_kwargs(sc):
    if sc['learning']:
        return dict(DP_alpha=[-1, -1], param_beta=sc['pp'], FP_mean=0.01, FP_sd=0.01,
                    FN_mean=0.2, FN_sd=0.1)
    return dict(DP_alpha=[-1, -1], param_beta=sc['pp'], FN_error=sc['FN'], FP_error=sc['FP'])


def initial_assignment(sc, z_true):
    """A deliberately imperfect start (labels 10*z, some cells scattered) so that the
    relabelling of libs/CRP.py:124-127 and both split and merge moves are exercised."""
    rng = np.random.default_rng(sc['seed'] + 1000)
    a = (10 * z_true + 3).astype(int)
    flip = rng.random(a.size) < 0.15
    a[flip] = rng.integers(0, 4, flip.sum()) * 10 + 3
    return [int(x) for x in a]


def run_scenario(ref, sc):
    data, z = simulate(sc['n'], sc['m'], k_true=sc['k'], miss=sc['miss'], seed=sc['seed'])
    kw = model_kwargs(sc)
    tape = rng_tape.Tape()
    rnd = rng_tape.LegacyRandom(record=tape)
    np.random.seed(sc['np_seed'])
    if sc['learning']:
        model = ref.CRP_learning_errors.CRP_errors_learning(data.copy(), **kw)
    else:
        model = ref.CRP.CRP(data.copy(), **kw)
    assign = initial_assignment(sc, z) if sc['init'] == 'assign' else None
    snaps, logs, pos = [], [], [0]
    with ref_shim.ref_errstate(), ref_shim.patched_random(ref, rnd):
        model.init(assign=assign)
        pos.append(len(tape.records))
        snaps.append(snapshot(model))
        for _ in range(sc['steps']):
            logs.append(do_step(model, rnd, sc['moves'], sc['learning']))
            pos.append(len(tape.records))
            snaps.append(snapshot(model))
    return data, assign, tape, pos, snaps, logs, kw


def save_fixture(path, sc, data, assign, tape, pos, snaps, logs, kw):
    S1 = len(snaps)
    N, M = data.shape
    kmax = max(s['ids'].size for s in snaps)
    ids = np.full((S1, kmax), -1, dtype=np.int64)
    sizes = np.full((S1, kmax), -1, dtype=np.int64)
    theta = np.zeros((S1, kmax, M), dtype=np.float32)
    for i, s in enumerate(snaps):
        k = s['ids'].size
        ids[i, :k] = s['ids']
        sizes[i, :k] = s['sizes']
        theta[i, :k] = s['theta']
    code = np.full(data.shape, -1, dtype=np.int8)
    code[data == 1] = 1
    code[data == 0] = 0
    kinds, tsizes, values = tape.to_arrays()
    meta = dict(name=sc['name'], learning=sc['learning'], kwargs=kw, moves=sc['moves'],
                init=sc['init'], init_assign=assign, steps=sc['steps'],
                generator='tests/golden/make_golden.py', reference='cbg-ethz/BnpC v0.2.1',
                numpy=np.__version__)
    np.savez_compressed(
        path, code=code, meta=json.dumps(meta), tape_kinds=kinds, tape_sizes=tsizes,
        tape_values=values, tape_pos=np.array(pos, dtype=np.int64),
        assignment=np.stack([s['assignment'] for s in snaps]),
        n_clusters=np.array([s['ids'].size for s in snaps]), ids=ids, sizes=sizes, theta=theta,
        alpha=np.array([s['alpha'] for s in snaps]), FN=np.array([s['FN'] for s in snaps]),
        FP=np.array([s['FP'] for s in snaps]), ll=np.array([s['ll'] for s in snaps]),
        lpost=np.array([s['lpost'] for s in snaps]), steplog=json.dumps(logs))


def main():
    ref = ref_shim.load_reference()
    for sc in SCENARIOS:
        out = run_scenario(ref, sc)
        path = os.path.join(HERE, sc['name'] + '.npz')
        save_fixture(path, sc, *out)
        logs = out[5]
        kinds = [l['move'] + ('+' if l['sm'] and l['sm'][0] else '') for l in logs]
        print(f"{sc['name']}: {os.path.getsize(path) / 1024:.0f} KB, {len(out[2].records)} records, "
              f"K {out[4][0]['ids'].size}->{out[4][-1]['ids'].size}, moves {' '.join(kinds)}")


if __name__ == '__main__':
    main()
