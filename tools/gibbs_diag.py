#!/usr/bin/env python3
"""Per-call wall time and sweep statistics of update_assignments_Gibbs for one chain."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bnpc_b200.synth import CONFIGS, make_matrix  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--config', default='C3')
ap.add_argument('--steps', type=int, default=60)
ap.add_argument('--seed', type=int, default=7)
args = ap.parse_args()

import torch  # noqa: E402
import libs.CRP_learning_errors as crple  # noqa: E402
import libs.CRP as crp  # noqa: E402
from bnpc_b200.rng import PhiloxRandom  # noqa: E402
from libs.MCMC import Chain_steps  # noqa: E402

cfg = CONFIGS[args.config]
data, z = make_matrix(cfg['cells'], cfg['muts'], cfg['k_true'], cfg['fn'], cfg['fp'], cfg['miss'], seed=0)
if cfg['learning']:
    m = crple.CRP_errors_learning(data, DP_alpha=[-1, -1], param_beta=list(cfg['pp']), FP_mean=0.01, FP_sd=0.01,
                                  FN_mean=0.2, FN_sd=0.1, rnd=PhiloxRandom(args.seed), device='cuda:0')
else:
    m = crp.CRP(data, DP_alpha=[-1, -1], param_beta=list(cfg['pp']), FN_error=cfg['FN'], FP_error=cfg['FP'],
                rnd=PhiloxRandom(args.seed), device='cuda:0')
m.init(assign=[int(v) for v in z])
moves = dict(sm_prob=cfg.get('sm_prob', 0.33), dpa_prob=0.25, error_prob=0.25 if cfg['learning'] else 0.0,
             sm_ratios=[0.75, 0.25], sm_steps=3, param_proposal_sd=np.array([0.1, 0.25, 0.5]))
ch = Chain_steps(m, 1, args.steps + 8, 0, moves, 0, False)
fn = m.update_assignments_Gibbs
rows = []


def timed():
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    rows.append((time.perf_counter() - t0, dict(m.sweep_stats), len(m.cells_per_cluster)))


m.update_assignments_Gibbs = timed
for i in range(args.steps):
    ch.do_step()
    ch.update_results(1 + i, False)
for dt, st, k in rows:
    print(f'{1e3 * dt:8.3f} ms K={k:3d} epochs={st["epochs"]} births={st["births"]} moved={st["moved"]} '
          f'slow={st["slow"]} unc={st.get("uncertain")} kernel_us={st["us"]:.0f} kcycles={st["kcycles"]} '
          f'phases(wait|own|walk|barrier)={st.get("phases_kcyc")}')
