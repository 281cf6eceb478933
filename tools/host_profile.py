#!/usr/bin/env python3
"""cProfile of the host side of one chain's steps (C3), excluding set-up."""
import cProfile, pstats, sys, os, io
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
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
mode = sys.argv[2] if len(sys.argv) > 2 else 'dev'
ch = Chain_steps(m, 1, steps + 8, 0, moves, 0, False)
tr = torch.zeros((4, cfg['cells']), dtype=torch.int32, device='cuda:0')


def run(n, off):
    for i in range(n):
        ch.do_step()
        if mode == 'dev':
            ll = ch.model.get_ll_full()
            ch.results['ML'][off + i] = ll
            ch.results['MAP'][off + i] = ll + ch.model.get_lprior_full()
            ch.model.copy_assignment_to(tr[i % 4])
        else:
            ch.update_results(off + i, False)


run(5, 1)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
run(steps, 6)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
st = pstats.Stats(pr, stream=s)
st.sort_stats('tottime').print_stats(28)
print(s.getvalue())
s = io.StringIO()
st = pstats.Stats(pr, stream=s)
st.sort_stats('cumulative').print_stats(22)
print(s.getvalue())
