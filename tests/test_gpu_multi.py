"""GPU, two devices: a chain's trace must not depend on the GPU / rank that ran it (SURVEY.md
section 8e "this is the distributed test").  The same four chains run once in one process on one
GPU and once under torchrun on two ranks (chain c on rank c mod 2, traces gathered over NCCL by
libs.MCMC.gather_chains); every trace of every chain must be bit-identical, in chain (= seed) order.
Skipped on a single-GPU box (the 2-rank gloo test of the gather runs on CPU)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_traces_do_not_depend_on_the_rank(tmp_path):
    single, double = str(tmp_path / 'single.npz'), str(tmp_path / 'double.npz')
    env = {k: v for k, v in os.environ.items() if k not in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK')}
    tool = os.path.join(ROOT, 'tools', 'trace_dump.py')
    subprocess.run([sys.executable, tool, '--out', single], check=True, env=dict(env, CUDA_VISIBLE_DEVICES='0'),
                   timeout=600)
    subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                    '--master-addr', '127.0.0.1', '--master-port', '29611', tool, '--out', double],
                   check=True, env=env, timeout=900)
    a, b = np.load(single), np.load(double)
    assert int(a['n_chains']) == int(b['n_chains']) == 4 and int(b['world']) == 2
    for key in a.files:
        if key in ('n_chains', 'world'):
            continue
        x, y = a[key], b[key]
        if key.endswith('_params'):
            # the gather pads the cluster axis to the largest K of ALL chains (libs/utils.py:206-223 does
            # the same when it concatenates chains)
            k = min(x.shape[1], y.shape[1])
            assert not x[:, k:].any() and not y[:, k:].any(), key
            x, y = x[:, :k], y[:, :k]
        np.testing.assert_array_equal(x, y, err_msg=key)
