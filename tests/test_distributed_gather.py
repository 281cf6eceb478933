"""CPU, world_size 2, gloo: the only cross-rank step of the path -- the
end-of-run gather of the chains' traces on rank 0 (libs/MCMC.py::gather_chains, replacing the
pickle-through-pipe of the reference's mp.Pool, libs/MCMC.py:114-118) -- and the chain -> rank
partition.  The step loop itself has no collective."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip('torch')
import torch.multiprocessing as mp  # noqa: E402

from libs.MCMC import chains_of_rank  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _fake_results(chain, steps=7, cells=50, muts=12, ragged=False):
    """trace dict of one finished chain, as Chain.update_results leaves it; K differs by chain and
    -- ragged: the run-time mode, libs/MCMC.py:395-440 -- so do the length and the burn-in"""
    rng = np.random.default_rng(100 + chain)
    k = 2 + chain
    burn = 0
    if ragged:
        steps, burn = steps + 3 * chain, 1 + chain
    return dict(ML=rng.random(steps), MAP=rng.random(steps), DP_alpha=rng.random(steps),
                FN=rng.random(steps), FP=rng.random(steps),
                assignments=rng.integers(0, k, (steps, cells)).astype(np.int32),
                params=rng.random((steps - burn, k, muts)).astype(np.float32), burn_in=burn)


class _Chain:
    def __init__(self, results):
        self.results = results
        self.model = type('M', (), {'device': 'cpu'})()

    def get_result(self):
        return self.results


def _worker(rank, world, port, n_chains, out_path, ragged=False):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from libs.MCMC import gather_chains
    dist.init_process_group('gloo', rank=rank, world_size=world)
    mine = chains_of_rank(n_chains, rank, world)
    chains = [None] * n_chains
    for c in mine:
        chains[c] = _Chain(_fake_results(c, ragged=ragged))
    got = gather_chains(chains, n_chains, rank, world)
    if rank == 0:
        np.save(out_path, np.array([len(got)] + [g.get_result()['params'].shape[1] for g in got]))
        # chain (= seed) order, whatever rank ran the chain
        kmax = max(2 + c for c in range(n_chains))
        for g, c in zip(got, range(n_chains)):
            want, res = _fake_results(c, ragged=ragged), g.get_result()
            assert res['burn_in'] == want['burn_in']
            np.testing.assert_array_equal(res['assignments'], want['assignments'])
            k = want['params'].shape[1]
            assert res['params'].shape[1] == kmax and res['params'].shape[0] == want['params'].shape[0]
            np.testing.assert_array_equal(res['params'][:, :k], want['params'])
            assert not res['params'][:, k:].any()            # zero padding to the global Kmax
            for key in ('ML', 'MAP', 'DP_alpha', 'FN', 'FP'):
                np.testing.assert_array_equal(res[key], want[key])
    else:
        assert got == []
    dist.barrier()
    dist.destroy_process_group()


def test_chain_partition():
    assert chains_of_rank(8, 0, 1) == list(range(8))
    assert chains_of_rank(8, 1, 2) == [1, 3, 5, 7]
    parts = [chains_of_rank(11, r, 4) for r in range(4)]
    assert sorted(c for p in parts for c in p) == list(range(11))


@pytest.mark.timeout(180)
def test_gather_traces_world_size_2(tmp_path):
    out = str(tmp_path / 'n.npy')
    mp.spawn(_worker, args=(2, _free_port(), 4, out), nprocs=2, join=True)
    n = np.load(out)
    assert n[0] == 4 and (n[1:] == 5).all()


@pytest.mark.timeout(180)
def test_gather_traces_of_unequal_length(tmp_path):
    # chains of the run-time mode end after different numbers of steps with different burn-ins
    out = str(tmp_path / 'n.npy')
    mp.spawn(_worker, args=(2, _free_port(), 4, out, True), nprocs=2, join=True)
    n = np.load(out)
    assert n[0] == 4 and (n[1:] == 5).all()


def _lugsail_worker(rank, world, port):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from libs.MCMC import MCMC
    dist.init_process_group('gloo', rank=rank, world_size=world)

    class _Done:                       # a chain whose ML trace has converged: no extension round
        def __init__(self, seed):
            self.results = dict(ML=np.random.default_rng(seed).normal(-100, 1e-6, 400),
                                params=np.zeros((400, 2, 3), dtype=np.float32))
    mcmc = MCMC(model=None)
    n = 1                              # fewer chains than ranks: rank 1 owns none
    mine = chains_of_rank(n, rank, world)
    mcmc.chains = [_Done(c) if c in mine else None for c in range(n)]
    mcmc.run_lugsail_chains(1.5, mine, verbosity=0)
    if rank == 0:
        r = mcmc.chains[0].results
        assert r['burn_in'] == 201 and r['PSRF_cutoff'] == 1.5 and r['PSRF'][-1][1] <= 1.5
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_lugsail_round_with_a_rank_that_owns_no_chain():
    mp.spawn(_lugsail_worker, args=(2, _free_port()), nprocs=2, join=True)
