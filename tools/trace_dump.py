#!/usr/bin/env python3
"""Run a few chains through libs.MCMC.MCMC.run (the public driver) on a seeded simulated matrix and
save every trace of every chain (rank 0 only under torchrun).  Used to check that a chain's trace
does not depend on the GPU / rank it ran on (the distributed test):

    python tools/trace_dump.py --out single.npz
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \\
        tools/trace_dump.py --out two_ranks.npz
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bnpc_b200.synth import make_matrix  # noqa: E402
import libs.CRP_learning_errors as crple  # noqa: E402
from libs.MCMC import MCMC, dist_info  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', required=True)
    ap.add_argument('--chains', type=int, default=4)
    ap.add_argument('--steps', type=int, default=40)
    ap.add_argument('--cells', type=int, default=3000)
    ap.add_argument('--muts', type=int, default=200)
    args = ap.parse_args()
    rank, world, local = dist_info()
    data, z = make_matrix(args.cells, args.muts, k_true=6, seed=3)
    model = crple.CRP_errors_learning(data, DP_alpha=[-1, -1], param_beta=[0.25, 0.25], FP_mean=0.01, FP_sd=0.01,
                                      FN_mean=0.2, FN_sd=0.1)
    mcmc = MCMC(model, sm_prob=0.33, dpa_prob=0.5, error_prob=0.1, sm_ratios=[0.75, 0.25], sm_steps=3)
    mcmc.run((args.steps, args.steps // 4), 42, n=args.chains, verbosity=0, assign=[int(v) for v in z])
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    if rank != 0:
        _shutdown(world)
        return
    out = {}
    for c, r in enumerate(mcmc.get_results()):
        for key in ('assignments', 'params', 'ML', 'MAP', 'DP_alpha', 'FN', 'FP'):
            out[f'chain{c}_{key}'] = np.asarray(r[key])
        out[f'chain{c}_burn_in'] = np.asarray(r['burn_in'])
    np.savez(args.out, n_chains=len(mcmc.get_results()), world=world, **out)
    print(f'{len(mcmc.get_results())} chains from {world} rank(s) -> {args.out}')
    _shutdown(world)


def _shutdown(world):
    if world > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
