#!/usr/bin/env python3
"""Multi-GPU end-to-end check, to be launched with torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/multi_gpu_check.py

Every rank writes the same generated matrix file, runs run_BnpC.py's main() with 4 chains (chain c
on rank c mod world), rank 0 gathers the traces over NCCL (libs/MCMC.py::gather_chains), runs the
estimators and checks the outputs (all four traces arrived, posterior ARI against the simulation).
"""
import os
import sys
import tempfile

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import run_BnpC  # noqa: E402
from libs.MCMC import dist_info  # noqa: E402
from oracle.crp_oracle import simulate  # noqa: E402  (test infrastructure: data generator only)


def main():
    rank, world, local = dist_info()
    import torch
    torch.cuda.set_device(local)
    data, z = simulate(600, 80, k_true=5, miss=0.1, seed=33)
    tmp = tempfile.mkdtemp(prefix=f'bnpc_mg_{rank}_')
    path = os.path.join(tmp, 'data.csv')
    pd.DataFrame(np.where(np.isnan(data), 3, data).astype(int).T).to_csv(path, sep='\t', header=False, index=False)
    out = os.path.join(tmp, 'out')
    args = run_BnpC.parse_args([path, '-n', '4', '-s', '60', '-e', 'posterior', 'MAP', '-o', out, '--seed', '11',
                                '-v', '0', '-np'])
    res_dir = run_BnpC.main(args)
    if rank != 0:
        return
    assign = pd.read_csv(os.path.join(res_dir, 'assignment.txt'), sep='\t')
    assert list(assign['estimator']) == ['posterior', 'MAP'], assign
    from sklearn.metrics import adjusted_rand_score
    post = [int(v) for v in assign['Assignment'][0].split(' ')]
    ari = adjusted_rand_score(z, post)
    cfg = open(os.path.join(res_dir, 'args.txt')).read()
    steps_line = [l for l in cfg.splitlines() if l.startswith('steps:')][0]
    assert steps_line.count(',') == 3, steps_line          # traces of all 4 chains reached rank 0
    print(f'multi-GPU check ok on {world} ranks: posterior ARI {ari:.3f}, outputs in {res_dir}')
    assert ari > 0.8


if __name__ == '__main__':
    main()
