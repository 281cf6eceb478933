#!/usr/bin/env python3
"""Wall time of the pieces of Chain.update_results for one C3 chain."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bnpc_b200.synth import CONFIGS, make_matrix
import torch
import libs.CRP_learning_errors as crple
from bnpc_b200.rng import PhiloxRandom
from libs.MCMC import Chain_steps

cfg = CONFIGS['C3']
data, z = make_matrix(cfg['cells'], cfg['muts'], cfg['k_true'], cfg['fn'], cfg['fp'], cfg['miss'], seed=0)
m = crple.CRP_errors_learning(data, DP_alpha=[-1, -1], param_beta=list(cfg['pp']), FP_mean=0.01, FP_sd=0.01,
                              FN_mean=0.2, FN_sd=0.1, rnd=PhiloxRandom(4242), device='cuda:0')
m.init(assign=[int(v) for v in z])
moves = dict(sm_prob=0.33, dpa_prob=0.25, error_prob=0.25, sm_ratios=[0.75, 0.25], sm_steps=3,
             param_proposal_sd=np.array([0.1, 0.25, 0.5]))
ch = Chain_steps(m, 1, 60, 0, moves, 0, False)
T = {}


def tic(name, fn):
    t0 = time.perf_counter()
    r = fn()
    T.setdefault(name, []).append(time.perf_counter() - t0)
    return r


row = np.zeros((64, cfg['cells']), dtype=int)
for i in range(40):
    tic('do_step', ch.do_step)
    tic('get_ll_full', m.get_ll_full)
    tic('get_lprior_full', m.get_lprior_full)
    tic('assignment_into', lambda: m.assignment_into(row[i]))
    cl = tic('sort_keys', lambda: np.sort(np.fromiter(m.cells_per_cluster.keys(), dtype=int)))
    tic('parameters[]', lambda: m.parameters[cl])
    tic('update_results(all)', lambda: ch.update_results(1 + i, False))
for k, v in T.items():
    print(f'{k:22s} mean {1e3 * np.mean(v[5:]):7.3f} ms   max {1e3 * np.max(v[5:]):7.3f} ms')
