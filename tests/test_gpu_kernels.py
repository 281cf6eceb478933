"""GPU: each kernel of libbnpc_b200.so, called through the C ABI, against numpy / the oracle."""
import ctypes as C

import numpy as np
import pytest

from oracle.crp_oracle import OracleCRP, simulate
from oracle.rng_tape import LegacyRandom, Tape

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')


@pytest.fixture(scope='module')
def L():
    from bnpc_b200 import _lib
    return _lib.lib()


def dev(a, dtype):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device='cuda')


def sp():
    return torch.cuda.current_stream().cuda_stream


def pack(L, data):
    N, M = data.shape
    W = 4 * ((M + 127) // 128)
    code = np.where(data == 1, 1, np.where(data == 0, 0, -1)).astype(np.int8)
    x1 = torch.zeros((N, W), dtype=torch.int32, device='cuda')
    x0 = torch.zeros((N, W), dtype=torch.int32, device='cuda')
    n1 = torch.zeros(N, dtype=torch.int32, device='cuda')
    n0 = torch.zeros(N, dtype=torch.int32, device='cuda')
    code_d = dev(code, torch.int8)
    L.pack_planes(None, code_d.data_ptr(), N, M, W, x1.data_ptr(), x0.data_ptr(),
                  n1.data_ptr(), n0.data_ptr(), sp())
    torch.cuda.synchronize()
    return W, x1, x0, n1, n0


def planes_numpy(data, W):
    N, M = data.shape
    out = []
    for val in (1, 0):
        bits = np.zeros((N, W * 32), dtype=np.uint8)
        bits[:, :M] = (data == val)
        out.append(np.packbits(bits, axis=1, bitorder='little').view(np.uint32).reshape(N, W))
    return out


@pytest.mark.parametrize('shape', [(1, 1), (7, 31), (33, 32), (50, 129), (300, 1000), (2000, 50)])
def test_pack_planes(L, shape):
    data, _ = simulate(*shape, k_true=3, miss=0.2, seed=1)
    W, x1, x0, n1, n0 = pack(L, data)
    w1, w0 = planes_numpy(data, W)
    np.testing.assert_array_equal(x1.cpu().numpy().view(np.uint32), w1)
    np.testing.assert_array_equal(x0.cpu().numpy().view(np.uint32), w0)
    np.testing.assert_array_equal(n1.cpu().numpy(), (data == 1).sum(1))
    np.testing.assert_array_equal(n0.cpu().numpy(), (data == 0).sum(1))
    # float64 input path gives the same planes
    N, M = shape
    y1 = torch.zeros_like(x1)
    y0 = torch.zeros_like(x0)
    data_d = dev(data, torch.float64)
    L.pack_planes(data_d.data_ptr(), None, N, M, W, y1.data_ptr(), y0.data_ptr(),
                  n1.data_ptr(), n0.data_ptr(), sp())
    assert torch.equal(x1, y1) and torch.equal(x0, y0)


def test_philox_fills(L):
    n = 100001
    a = torch.empty(n, dtype=torch.float64, device='cuda')
    b = torch.empty(n, dtype=torch.float64, device='cuda')
    L.fill_uniform(a.data_ptr(), n, 1234, 7, 0, sp())
    L.fill_uniform(b.data_ptr(), n, 1234, 7, 0, sp())
    assert torch.equal(a, b)
    x = a.cpu().numpy()
    assert x.min() >= 0 and x.max() < 1 and abs(x.mean() - 0.5) < 0.01 and abs(x.var() - 1 / 12) < 0.01
    L.fill_uniform(b.data_ptr(), n, 1234, 8, 0, sp())
    assert not torch.equal(a, b)
    L.fill_uniform(b.data_ptr(), n, 1234, 9, 3, sp())
    y = b.cpu().numpy()
    assert set(np.unique(y)) == {0.0, 1.0, 2.0}
    assert np.all(np.abs(np.bincount(y.astype(int)) / n - 1 / 3) < 0.01)
    for m in (1, 2, 3, 17, 1000, 100000):
        p = torch.empty(m, dtype=torch.int32, device='cuda')
        L.fill_permutation(p.data_ptr(), m, 99, 3, sp())
        v = p.cpu().numpy()
        np.testing.assert_array_equal(np.sort(v), np.arange(m))
        if m >= 1000:
            q = torch.empty(m, dtype=torch.int32, device='cuda')
            L.fill_permutation(q.data_ptr(), m, 99, 4, sp())
            assert (q.cpu().numpy() != v).mean() > 0.9
            assert abs(np.corrcoef(v, np.arange(m))[0, 1]) < 0.05


@pytest.mark.parametrize('shape,K', [((64, 40), 5), ((257, 333), 11), ((1500, 1000), 20), ((300, 50), 70)])
def test_ll_matrix_matches_reference_arithmetic(L, shape, K):
    data, _ = simulate(*shape, k_true=4, miss=0.15, seed=2)
    N, M = shape
    orc = OracleCRP(data, param_beta=[0.25, 0.25], FN_error=0.17, FP_error=0.013)
    rng = np.random.default_rng(3)
    theta = np.clip(rng.random((K + 3, M)), 1e-5, 1 - 1e-5).astype(np.float32)
    ids = rng.permutation(K + 3)[:K].astype(np.int32)
    W, x1, x0, _, _ = pack(L, data)
    lp = torch.empty(K * M * 2, dtype=torch.float64, device='cuda')
    theta_d, ids_d = dev(theta, torch.float32), dev(ids, torch.int32)
    L.logprob_tables(theta_d.data_ptr(), ids_d.data_ptr(), K, M, 0.17, 0.013, lp.data_ptr(), sp())
    th = theta[ids]
    t64 = th.astype(np.float64)
    omt = (np.float32(1) - th).astype(np.float64)
    want1 = np.log(t64 * (1 - 0.17) + omt * 0.013)
    want0 = np.log(t64 * 0.17 + omt * (1 - 0.013))
    got = lp.cpu().numpy().reshape(K, M, 2)
    np.testing.assert_allclose(got[..., 0], want1, rtol=1e-15, atol=1e-15)
    np.testing.assert_allclose(got[..., 1], want0, rtol=1e-15, atol=1e-15)
    cells = rng.permutation(N).astype(np.int32)
    ldk = K + (K & 1)
    ll = torch.zeros(N * ldk, dtype=torch.float64, device='cuda')
    cells_d = dev(cells, torch.int32)
    L.ll_matrix(x1.data_ptr(), x0.data_ptr(), W, M, cells_d.data_ptr(), 1, N,
                lp.data_ptr(), K, ll.data_ptr(), ldk, sp())
    got = ll.cpu().numpy().reshape(N, ldk)[:, :K]
    want = np.stack([orc.loglik(data[cells], th[k]) for k in range(K)], axis=1)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-10)


def test_suffstat_and_row_loglik(L):
    from bnpc_b200.engine import DeviceCRP
    from bnpc_b200.rng import PhiloxRandom
    data, z = simulate(700, 130, k_true=6, miss=0.2, seed=4)
    m = DeviceCRP(data, param_beta=[0.25, 0.25], FN_error=0.2, FP_error=0.01, rnd=PhiloxRandom(5))
    m.init(assign=[int(v) * 3 + 1 for v in z])
    with torch.cuda.stream(m.stream):
        m._refresh_stats()
        m.stream.synchronize()
    K = len(m.cells_per_cluster)
    S1 = m.S1[:K * 130].cpu().numpy().reshape(K, 130)
    S0 = m.S0[:K * 130].cpu().numpy().reshape(K, 130)
    a = m.assignment
    for k in range(K):
        np.testing.assert_array_equal(S1[k], (data[a == k] == 1).sum(0))
        np.testing.assert_array_equal(S0[k], (data[a == k] == 0).sum(0))
    orc = OracleCRP(data, param_beta=[0.25, 0.25], FN_error=0.2, FP_error=0.01)
    orc.assignment = a
    orc.cells_per_cluster = dict(m.cells_per_cluster)
    orc.parameters = np.zeros((700, 130), dtype=np.float32)
    orc.parameters[:K] = m.parameters[np.arange(K)]
    orc.refresh_crp_table()
    np.testing.assert_allclose(m.get_ll_full(), orc.get_ll_full(), rtol=1e-12)
    np.testing.assert_allclose(m.get_lprior_full(), orc.get_lprior_full(), rtol=1e-12)


@pytest.mark.parametrize('pp', [(0.25, 0.25), (1, 1)])
@pytest.mark.parametrize('want_logq', [False, True])
def test_mh_theta_matches_scipy_path(L, pp, want_logq):
    """The device truncated-normal proposal / acceptance ratio against the oracle's scipy
    calls on the same draws, including the clipped boundary values of theta."""
    data, z = simulate(400, 257, k_true=3, miss=0.1, seed=6)
    M = 257
    orc = OracleCRP(data, param_beta=list(pp), FN_error=0.2, FP_error=0.01)
    rng = np.random.default_rng(7)
    rows = 6
    theta = np.clip(rng.random((rows, M)), 1e-5, 1 - 1e-5)
    theta[0, :40] = 1e-5
    theta[1, :40] = 1 - 1e-5
    theta[2, :60] = rng.random(60) * 1e-3
    theta[3, :60] = 1 - rng.random(60) * 1e-3
    theta = np.clip(theta, 1e-5, 1 - 1e-5).astype(np.float32)
    members = [np.flatnonzero(z == (r % 3))[: 20 + 50 * r] for r in range(rows)]
    S1 = np.stack([(data[c] == 1).sum(0) for c in members]).astype(np.int32)
    S0 = np.stack([(data[c] == 0).sum(0) for c in members]).astype(np.int32)
    np.random.seed(11)
    want_theta, want_lq, want_dec, draws = [], [], [], np.empty((3, rows, M))
    for r in range(rows):
        t = Tape()
        orc.rnd = LegacyRandom(record=t)
        with np.errstate(divide='raise', invalid='raise', under='ignore', over='ignore'):
            new, lq, dec = orc.mh_theta_row(theta[r].copy(), members[r], want_logq)
        want_theta.append(new)
        want_lq.append(lq)
        want_dec.append(dec)
        draws[0, r], draws[1, r], draws[2, r] = t.records[0][1], t.records[1][1], t.records[2][1]
    th_d = dev(theta, torch.float32)
    logq = torch.zeros(rows * M, dtype=torch.float64, device='cuda')
    dec = torch.zeros(rows, dtype=torch.int32, device='cuda')
    S1_d, S0_d, draws_d = dev(S1, torch.int32), dev(S0, torch.int32), dev(draws, torch.float64)
    L.mh_theta(th_d.data_ptr(), None, rows, M, S1_d.data_ptr(), S0_d.data_ptr(), draws_d.data_ptr(), 0.2, 0.01,
               float(pp[0]), float(pp[1]), 1 if want_logq else 0,
               logq.data_ptr() if want_logq else None, dec.data_ptr(), sp())
    got = th_d.cpu().numpy()
    want = np.stack(want_theta)
    mism = np.argwhere(got != want)
    assert mism.shape[0] == 0, f'{mism.shape[0]} theta mismatches, first {mism[:5]}: ' \
                               f'{got[tuple(mism[0])]!r} vs {want[tuple(mism[0])]!r}'
    np.testing.assert_array_equal(dec.cpu().numpy(), np.array(want_dec))
    if want_logq:
        got_lq = logq.cpu().numpy().reshape(rows, M).sum(1)
        np.testing.assert_allclose(got_lq, np.array(want_lq), rtol=1e-9)


def test_gather_members_and_anchor_swaps(L):
    rng = np.random.default_rng(8)
    N = 5000
    a = rng.integers(0, 7, N).astype(np.int32)
    a_d = dev(a, torch.int32)
    out = torch.full((N + 8,), -1, dtype=torch.int32, device='cuda')
    blk = torch.zeros(2 * ((N + 1023) // 1024) + 2, dtype=torch.int32, device='cuda')
    L.gather_members(a_d.data_ptr(), N, 3, -1, out.data_ptr(), blk.data_ptr(), sp())
    c3 = np.flatnonzero(a == 3)
    np.testing.assert_array_equal(out.cpu().numpy()[:c3.size], c3)
    L.gather_members(a_d.data_ptr(), N, 5, 2, out.data_ptr(), blk.data_ptr(), sp())
    c5, c2 = np.flatnonzero(a == 5), np.flatnonzero(a == 2)
    cells = np.concatenate([c5, c2])
    np.testing.assert_array_equal(out.cpu().numpy()[:cells.size], cells)
    # merge swaps (libs/CRP.py:496,500)
    ci, cj = c5.copy(), c2.copy()
    ci[0], ci[7] = ci[7], ci[0]
    cj[-1], cj[4] = cj[4], cj[-1]
    L.anchor_swaps(out.data_ptr(), cells.size, c5.size, 7, 4, 1, sp())
    np.testing.assert_array_equal(out.cpu().numpy()[:cells.size], np.concatenate([ci, cj]))
    # split swaps on a fresh list, including the aliasing case idx_j == 0 (libs/CRP.py:449-450)
    for i, j in ((5, 9), (3, 0), (c3.size - 1, 2)):
        L.gather_members(a_d.data_ptr(), N, 3, -1, out.data_ptr(), blk.data_ptr(), sp())
        w = c3.copy()
        w[0], w[i] = w[i], w[0]
        w[-1], w[j] = w[j], w[-1]
        L.anchor_swaps(out.data_ptr(), c3.size, c3.size, i, j, 0, sp())
        np.testing.assert_array_equal(out.cpu().numpy()[:c3.size], w)


@pytest.mark.parametrize('shape', [(300, 128, 5), (5000, 1000, 24), (4096, 640, 64), (777, 50, 40)])
@pytest.mark.parametrize('rows', ['tcgen05', 'tcgen05_i8', 'fp32_fma'])
def test_approximate_rows_within_their_error_bound(L, shape, rows):
    """The approximate log-likelihood rows of lean epochs (tcgen05 tensor cores, bf16-split
    operands, FP32 accumulation; FP32-FMA reference kernel) against the FP64 matrix.  The bound
    the option selection relies on is terms * 2^-22 * (|ll| + 64) + 0.05 (bnpc_gibbs_options);
    measured errors are ~1e-6 relative."""
    N, M, K = shape
    rng = np.random.default_rng(N + M + K)
    data = rng.integers(0, 2, (N, M)).astype(np.float64)
    data[rng.random((N, M)) < 0.1] = np.nan
    W, x1, x0, n1, n0 = pack(L, data)
    theta = dev(np.clip(rng.random((K, M)), 1e-5, 1 - 1e-5).astype(np.float32), torch.float32)
    lp = torch.zeros(2 * K * M, dtype=torch.float64, device='cuda')
    L.logprob_tables(theta.data_ptr(), None, K, M, 0.2, 0.01, lp.data_ptr(), sp())
    cells = dev(rng.permutation(N).astype(np.int32), torch.int32)
    ldk = K | 1
    ll = torch.zeros(N * ldk, dtype=torch.float64, device='cuda')
    L.ll_matrix(x1.data_ptr(), x0.data_ptr(), W, M, cells.data_ptr(), 1, N, lp.data_ptr(), K, ll.data_ptr(), ldk, sp())
    kp = (K + 7) & ~7
    llf = torch.full((N, kp), float('nan'), dtype=torch.float32, device='cuda')
    if rows == 'tcgen05':
        scratch = torch.zeros(W * 2 * kp * 64, dtype=torch.int16, device='cuda')
        L.ll_matrix_tc(x1.data_ptr(), x0.data_ptr(), W, M, cells.data_ptr(), 1, N, lp.data_ptr(),
                       scratch.data_ptr(), K, llf.data_ptr(), kp, sp())
    elif rows == 'tcgen05_i8':
        # integer digits: exact accumulation, the error is the 16-bit quantisation of the table
        vmax = float(lp.abs().max().item()) * 1.0001
        scratch = torch.zeros(W * kp * 128, dtype=torch.uint8, device='cuda')
        L.ll_matrix_i8(x1.data_ptr(), x0.data_ptr(), W, M, cells.data_ptr(), 1, N, lp.data_ptr(),
                       scratch.data_ptr(), K, vmax, llf.data_ptr(), kp, sp())
    else:
        scratch = torch.zeros(2 * K * M, dtype=torch.float32, device='cuda')
        L.ll_matrix_f32(x1.data_ptr(), x0.data_ptr(), W, M, cells.data_ptr(), 1, N, lp.data_ptr(),
                        scratch.data_ptr(), K, llf.data_ptr(), kp, sp())
    torch.cuda.synchronize()
    want = ll.cpu().numpy().reshape(N, ldk)[:, :K]
    got = llf.cpu().numpy()[:, :K].astype(np.float64)
    assert not np.isnan(got).any()
    bound = 2 * M * 2.0 ** -22 * (np.abs(want) + 64) + 0.05
    if rows == 'tcgen05_i8':
        # what bnpc_chain_gibbs_epoch passes as err_abs, and the sharper statement: at most half a
        # quantisation step per observed entry plus the float rounding of the result
        bound = bound + M * (vmax / 65535) / 2
        observed = (~np.isnan(data)).sum(axis=1)[cells.cpu().numpy()][:, None]
        sharp = observed * (vmax / 65535) / 2 + 2.0 ** -22 * np.abs(want) + 1e-9
        assert (np.abs(got - want) <= sharp).all()
    assert (np.abs(got - want) <= bound).all()
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max() + 1e-3 + (0.02 if rows == 'tcgen05_i8' else 0)


@pytest.mark.parametrize('shape', [(5000, 1000, (24, 28, 22, 25)), (777, 50, (40, 3)), (40000, 300, (9, 17)),
                                   (90000, 200, (30, 31, 32, 8, 24, 16, 20, 12)), (300, 128, (64, 64, 5)),
                                   (80000, 130, (20,))])
def test_shared_integer_rows_equal_the_one_chain_rows(L, shape):
    """bnpc_ll_matrix_i8_shared (csrc/bnpc_tc_i8s.cuh): the integer tensor-core rows of several chains
    over the same cells in cell order, sharing the expanded data operand in tensor memory (and two
    tiles per table chunk at the larger shapes), against one bnpc_ll_matrix_i8 call per chain with an
    explicit cell list (the one-chain kernel of csrc/bnpc_tc_i8.cuh).  Integer accumulation is exact,
    so the floats are identical.  The shapes cover mixed cluster counts (different paddings in one
    MMA), more chains than fit one group, a last tile that is not full, a single chunk pair per row
    (M = 50), resident and streamed tables."""
    N, M, Ks = shape
    rng = np.random.default_rng(N + M + len(Ks))
    data = rng.integers(0, 2, (N, M)).astype(np.float64)
    data[rng.random((N, M)) < 0.1] = np.nan
    W, x1, x0, n1, n0 = pack(L, data)
    cells = dev(np.arange(N, dtype=np.int32), torch.int32)
    lps, bss, outs, refs, vmaxs, kps = [], [], [], [], [], []
    for K in Ks:
        theta = dev(np.clip(rng.random((K, M)), 1e-5, 1 - 1e-5).astype(np.float32), torch.float32)
        lp = torch.zeros(2 * K * M, dtype=torch.float64, device='cuda')
        L.logprob_tables(theta.data_ptr(), None, K, M, 0.2, 0.01, lp.data_ptr(), sp())
        kp = (K + 7) & ~7
        vmax = float(lp.abs().max().item()) * 1.0001
        bs = torch.zeros(W * kp * 128, dtype=torch.uint8, device='cuda')
        ref = torch.full((N, kp), float('nan'), dtype=torch.float32, device='cuda')
        L.ll_matrix_i8(x1.data_ptr(), x0.data_ptr(), W, M, cells.data_ptr(), 1, N, lp.data_ptr(), bs.data_ptr(), K,
                       vmax, ref.data_ptr(), kp, sp())
        torch.cuda.synchronize()
        lps.append(lp); bss.append(bs); refs.append(ref); vmaxs.append(vmax); kps.append(kp)
        outs.append(torch.full((N, kp), float('nan'), dtype=torch.float32, device='cuda'))
    nc = len(Ks)
    P = C.c_void_p * nc
    L.ll_matrix_i8_shared(x1.data_ptr(), x0.data_ptr(), W, M, N, nc, P(*[t.data_ptr() for t in lps]),
                          P(*[t.data_ptr() for t in bss]), (C.c_int * nc)(*Ks), (C.c_double * nc)(*vmaxs),
                          P(*[t.data_ptr() for t in outs]), (C.c_int * nc)(*kps), sp())
    torch.cuda.synchronize()
    for c, K in enumerate(Ks):
        assert not torch.isnan(outs[c][:, :K]).any()
        assert torch.equal(outs[c][:, :K], refs[c][:, :K]), f'chain {c} (K = {K})'
    # the one-chain call without a cell list takes the same kernel (a group of one)
    one = torch.full((N, kps[0]), float('nan'), dtype=torch.float32, device='cuda')
    L.ll_matrix_i8(x1.data_ptr(), x0.data_ptr(), W, M, None, 1, N, lps[0].data_ptr(), bss[0].data_ptr(), Ks[0],
                   vmaxs[0], one.data_ptr(), kps[0], sp())
    torch.cuda.synchronize()
    assert torch.equal(one[:, :Ks[0]], refs[0][:, :Ks[0]])


@pytest.mark.parametrize('nf', [1, 37, 1024, 1025, 5000, 20000])
@pytest.mark.parametrize('spread', [0.02, 3.0])
def test_rg_scan_matches_the_serial_scan(L, nf, spread):
    """One restricted Gibbs scan (libs/CRP.py:609-632) on random two-column log-likelihoods:
    the device scan (parallel thresholds + the speculative-segment integer pass + write-back)
    against the serial scan in the reference's arithmetic.  spread = 0.02: nearly every cell is
    ambiguous (thresholds all over the range, the speculation is corrected often); 3.0: most
    cells are decisive."""
    from oracle.crp_oracle import log_crp_weight, log_normalise_pair
    rng = np.random.default_rng(nf + int(100 * spread))
    n = nf + 2
    alpha = 7.5
    ll2 = np.stack([rng.normal(-300, 20, nf)] * 2, axis=1) + rng.normal(0, spread, (nf, 2))
    perm = rng.permutation(nf).astype(np.int32)
    u = rng.random(nf)
    half0 = (rng.random(nf) < 0.4).astype(np.int32)
    # serial model
    half = half0.astype(np.float64).copy()
    lq_want = np.zeros(nf)
    with np.errstate(divide='raise', invalid='raise', over='ignore', under='ignore'):
        for s in range(nf):
            c = perm[s]
            half[c] = -1
            n_j = np.nansum(half) + 2
            n_i = n - n_j - 1
            lprob = log_normalise_pair(ll2[c] + log_crp_weight(np.array([n_i, n_j]), n, alpha))
            p = np.exp(lprob)
            side = 0 if u[s] < p[0] / (p[0] + p[1]) else 1
            half[c] = side
            lq_want[c] = lprob[side]
    ll2_d = dev(ll2, torch.float64)
    perm_d, u_d, half_d = dev(perm, torch.int32), dev(u, torch.float64), dev(half0, torch.int32)
    lq_d = torch.zeros(nf, dtype=torch.float64, device='cuda')
    work = torch.zeros(2 * nf + 16, dtype=torch.int32, device='cuda')
    L.rg_scan(ll2_d.data_ptr(), 2, n, perm_d.data_ptr(), u_d.data_ptr(), half_d.data_ptr(), alpha, 0, None,
              None, -1, lq_d.data_ptr(), work.data_ptr(), sp())
    torch.cuda.synchronize()
    np.testing.assert_array_equal(half_d.cpu().numpy(), half.astype(np.int32))
    np.testing.assert_allclose(lq_d.cpu().numpy(), lq_want, rtol=1e-12, atol=1e-300)


def test_device_beta_variates_follow_scipy(L):
    """production mode: bnpc_beta_rows samples theta ~ Beta(p + S1, q + S0) on the device (Philox +
    Marsaglia-Tsang, libs/CRP.py:172-175,183-188 draw them from numpy); Kolmogorov-Smirnov against
    scipy for the shapes the model uses (sparse prior, flat prior, well-populated clusters)."""
    from scipy.stats import beta, kstest
    n = 200_000
    for p, q, s1, s0 in ((0.25, 0.25, 0, 0), (1.0, 1.0, 0, 0), (0.25, 0.25, 3, 1), (1.0, 1.0, 400, 4000)):
        S1 = torch.full((1, n), s1, dtype=torch.int32, device='cuda')
        S0 = torch.full((1, n), s0, dtype=torch.int32, device='cuda')
        out = torch.zeros((1, n), dtype=torch.float32, device='cuda')
        L.beta_rows(S1.data_ptr(), S0.data_ptr(), 1, n, p, q, None, 1234, 77, out.data_ptr(), None,
                    torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        x = out.cpu().numpy().ravel().astype(np.float64)
        assert x.min() >= 1e-5 * 0.999 and x.max() <= 1 - 1e-5 * 0.999
        inner = x[(x > 2e-5) & (x < 1 - 2e-5)]                       # the clipped mass sits on the two bounds
        dist = beta(p + s1, q + s0)
        lo, hi = dist.cdf(2e-5), dist.cdf(1 - 2e-5)
        stat = kstest(inner, lambda v: (dist.cdf(v) - lo) / (hi - lo))
        assert stat.pvalue > 1e-4, (p, q, s1, s0, stat)
        assert abs((x <= 1.5e-5).mean() - dist.cdf(1e-5)) < 5e-3
