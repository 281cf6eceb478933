"""Estimators and convergence diagnostics of the chain results -- the callers on the output side
of the MCMC hot path (reference cbg-ethz/BnpC v0.2.1 `libs/utils.py`).

Same function names, arguments and return values as the reference (used by `libs/dpmmIO.py::
_infer_results`, :199-225, and `libs/MCMC.py`, :138-171).  The two O(S N^2) pieces of the posterior
estimator run on the GPU through the C ABI (`bnpc_cocluster_counts`, `bnpc_mpear_sums`):

* `get_dist`      libs/utils.py:90-97    pairwise co-clustering distance of the posterior samples
* `_get_MPEAR`    libs/utils.py:100-130  cut of the ward dendrogram with the best MPEAR score; the
                                         three pair sums of `_calc_MPEAR` (:133-145) for ALL
                                         candidate cuts come from one kernel launch, as exact
                                         integers
The ward linkage itself (scipy, O(N^2) memory on the host) and the O(S N) genotype averaging stay
on the host.  There is no CPU path for the two kernels: without the library or a CUDA device the
posterior estimator raises.

Large matrices (BASELINE config 3: 100k cells): the reference's N(N-1)/2 pair vector would be 20 GB
per sample.  Cells with the SAME assignment profile over all posterior samples are at distance 0
from each other and at equal distances from everybody else, and ward linkage merges them first:
the dendrogram above those merges is the weighted ward dendrogram of the DISTINCT profiles
(`_ward_linkage_weighted`: scipy's nearest-neighbour-chain algorithm with its ward update, started
from clusters of the profiles' multiplicities), and the three pair sums of the MPEAR score follow
from the profiles' pairs weighted by the multiplicities (`bnpc_mpear_sums_weighted`).  The result
is the reference's whenever the dendrogram has no tie at a candidate cut (ties are broken by
position in both, but the positions differ); the tests compare it with scipy on all cells.
"""
import numpy as np
import pandas as pd
from scipy.cluster.hierarchy import cut_tree, linkage
from scipy.special import binom, gamma
from scipy.stats import chi2
from sklearn.metrics import adjusted_rand_score
from sklearn.metrics.cluster import v_measure_score

EPSILON = np.finfo(np.float64).resolution
log_EPSILON = np.log(EPSILON)
MAX_LINKAGE_CELLS = 30_000       # condensed float64 distances of the host linkage: 3.6 GB at 30k cells
MAX_PROFILES = 24_000            # square float64 distances of the weighted linkage: 4.6 GB at 24k profiles


# ------------------------------------------------------------------------------ evaluation
def get_v_measure(pred_clusters, true_clusters, out_file=''):
    score = v_measure_score(true_clusters, pred_clusters)
    if out_file:
        _write_to_file(out_file, score)
    return score


def get_ARI(pred_clusters, true_clusters, out_file=''):
    score = adjusted_rand_score(true_clusters, pred_clusters)
    if out_file:
        _write_to_file(out_file, score)
    return score


def get_hamming_dist(df_pred, df_true):
    """libs/utils.py:63-72."""
    if df_true.shape != df_pred.shape:
        return np.count_nonzero(df_pred.round() != df_true.T)
    score = np.count_nonzero(df_pred.round() != df_true)
    score_t = np.count_nonzero(df_pred.round() != df_true.T)   # N x N frames that got transposed
    return min(score, score_t)


def _write_to_file(file, content, attach=False):
    with open(file, 'a' if attach else 'w') as f:
        f.write(str(content))


# ------------------------------------------------------------------- device side of MPEAR
class _PairCounts:
    """int32 counts[pair] on the device: samples in which the two cells are in different clusters."""

    def __init__(self, assignments, device=None):
        import torch
        from bnpc_b200 import _lib
        if not torch.cuda.is_available():
            raise RuntimeError('the posterior estimator needs a CUDA device (sm_100a); there is no CPU path')
        self.L = _lib.lib()
        self.torch = torch
        self.device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        a = np.ascontiguousarray(assignments, dtype=np.int32)
        self.S, self.N = a.shape
        if self.N < 2:
            raise ValueError('need at least two cells')
        with torch.cuda.device(self.device):
            self.stream = torch.cuda.current_stream(self.device).cuda_stream
            a_d = torch.from_numpy(a).to(self.device)
            self.counts = torch.empty(self.N * (self.N - 1) // 2, dtype=torch.int32, device=self.device)
            self.L.cocluster_counts(a_d.data_ptr(), self.S, self.N, self.counts.data_ptr(), self.stream)
            torch.cuda.current_stream(self.device).synchronize()

    def dist(self):
        """counts / S in float64 on the host (exactly the reference's `dist / steps`)."""
        return self.counts.cpu().numpy() / self.S

    def dist_square(self):
        """the same distances as a square float64 matrix (zero diagonal)"""
        from scipy.spatial.distance import squareform
        return squareform(self.dist(), checks=False)

    def sums(self, labels, weight=None):
        """labels int [n_cand, N] -> (T, A[n_cand], B[n_cand]) exact integers (bnpc_mpear_sums);
        weight [N]: multiplicity of every point (bnpc_mpear_sums_weighted)."""
        torch = self.torch
        lab = np.ascontiguousarray(labels, dtype=np.int32)
        n_cand = lab.shape[0]
        with torch.cuda.device(self.device):
            lab_d = torch.from_numpy(lab).to(self.device)
            out = torch.zeros(1 + 2 * n_cand, dtype=torch.int64, device=self.device)
            if weight is None:
                self.L.mpear_sums(self.counts.data_ptr(), self.N, lab_d.data_ptr(), n_cand, out.data_ptr(), self.stream)
            else:
                w_d = torch.from_numpy(np.ascontiguousarray(weight, dtype=np.int32)).to(self.device)
                self.L.mpear_sums_weighted(self.counts.data_ptr(), self.N, lab_d.data_ptr(), n_cand, w_d.data_ptr(),
                                           out.data_ptr(), self.stream)
            torch.cuda.current_stream(self.device).synchronize()
        o = out.cpu().numpy()
        return int(o[0]), o[1::2].copy(), o[2::2].copy()


def get_dist(assignments):
    """libs/utils.py:90-97: mean posterior cell-wise Hamming distance, condensed (pdist order)."""
    return _PairCounts(assignments).dist()


def _mpear_scores(total, same_pairs, same_counts, steps, cells):
    """Fritsch & Ickstadt (2009) eq. 13 (libs/utils.py:133-145) from the integer pair sums:
    I_sum = A, pi_sum = P - T/S, sum(I * pi) = A - B/S with pi = 1 - counts/S."""
    pairs = binom(cells, 2)
    i_sum = same_pairs.astype(np.float64)
    pi_sum = pairs - total / steps
    index = i_sum - same_counts.astype(np.float64) / steps
    expected = (i_sum * pi_sum) / pairs
    max_index = .5 * (i_sum + pi_sum)
    return (index - expected) / (max_index - expected)


def _candidate_cluster_numbers(assignments):
    """libs/utils.py:106-114."""
    cl_no = [int((np.unique(a, return_counts=True)[1] > 2).sum()) for a in assignments]
    avg_cl_no = np.mean(cl_no)
    return np.arange(max(2, avg_cl_no * 0.2), min(avg_cl_no * 2.5, assignments.shape[1]), dtype=int)


def _calc_MPEAR(pi, c):
    """libs/utils.py:133-145 for one labelling, from a condensed similarity `pi` (host float64;
    kept for callers of the reference's helper -- the estimator itself scores all cuts at once)."""
    iu = np.triu_indices(c.size, k=1)
    same = (c[iu[0]] == c[iu[1]]).astype(np.float64)
    i_sum, pi_sum, index = same.sum(), pi.sum(), (same * pi).sum()
    expected = (i_sum * pi_sum) / binom(c.size, 2)
    return (index - expected) / (.5 * (i_sum + pi_sum) - expected)


def _unique_profiles(assignments):
    """Group the cells by their assignment profile (column of `assignments`).  Returns
    (rep [U]: the LAST cell of every group, groups ordered by it -- the slot scipy's linkage leaves
    a merged cluster in --, inverse [N]: group of every cell, weight [U]).  Columns are hashed with
    two independent 64-bit multiplicative hashes and every cell is then compared with its group's
    representative, so the grouping is exact."""
    a = np.ascontiguousarray(assignments)
    steps, cells = a.shape
    rng = np.random.default_rng(0x5EED)
    h = np.zeros((2, cells), dtype=np.uint64)
    mult = rng.integers(1, 2 ** 63, size=(2, steps), dtype=np.uint64) | np.uint64(1)
    with np.errstate(over='ignore'):
        for s0 in range(0, steps, 256):                      # bounded temporaries
            blk = a[s0:s0 + 256].astype(np.uint64) + np.uint64(1)
            for k in range(2):
                h[k] += (blk * mult[k, s0:s0 + 256, None]).sum(axis=0, dtype=np.uint64)
    key = h[0] ^ (h[1] * np.uint64(0x9E3779B97F4A7C15))
    order = np.lexsort((np.arange(cells), h[1], key))
    sk, s1 = key[order], h[1][order]
    new = np.ones(cells, dtype=bool)
    new[1:] = (sk[1:] != sk[:-1]) | (s1[1:] != s1[:-1])
    gid_sorted = np.cumsum(new) - 1
    gid = np.empty(cells, dtype=np.int64)
    gid[order] = gid_sorted
    last = np.zeros(gid_sorted[-1] + 1, dtype=np.int64)
    np.maximum.at(last, gid, np.arange(cells))
    if not (a == a[:, last[gid]]).all():                     # a hash collision (never seen): split exactly
        _, gid = np.unique(a.T, axis=0, return_inverse=True)
        gid = gid.ravel()
        last = np.zeros(gid.max() + 1, dtype=np.int64)
        np.maximum.at(last, gid, np.arange(cells))
    rank = np.argsort(np.argsort(last))                      # groups ordered by their last cell
    inverse = rank[gid]
    rep = np.sort(last)
    weight = np.bincount(inverse, minlength=rep.size)
    return rep, inverse, weight


def _ward_linkage_weighted(dist_sq, weight):
    """Ward linkage of points with multiplicities: scipy.cluster.hierarchy.linkage(method='ward')
    (nearest-neighbour chain, `_ward` distance update, stable sort by height) started from clusters
    of `weight[i]` coincident points instead of singletons.  dist_sq: square float64 matrix of the
    plain distances (destroyed).  Returns a linkage matrix over the U points (column 3 counts points, as scipy
    requires; the heights are those of the weighted problem)."""
    D = dist_sq
    n = D.shape[0]
    size = np.asarray(weight, dtype=np.float64).copy()
    active = np.ones(n, dtype=bool)
    # ward distance of two clusters of coincident points (the state scipy's update reaches once the
    # zero-distance merges are done): d * sqrt(2 n_u n_v / (n_u + n_v)); = d for single points
    D *= np.sqrt(2.0 * np.outer(size, size) / np.add.outer(size, size))
    np.fill_diagonal(D, np.inf)
    merges = np.empty((n - 1, 3))
    chain = []
    lowest = 0
    for k in range(n - 1):
        if not chain:
            while not active[lowest]:
                lowest += 1
            chain.append(lowest)
        while True:
            x = chain[-1]
            row = D[x]
            y = int(np.argmin(row))                          # first minimum, as the scan `dist < current_min`
            if len(chain) > 1 and row[chain[-2]] <= row[y]:
                y = chain[-2]                                # the previous element is preferred on ties
            if len(chain) > 1 and y == chain[-2]:
                break
            chain.append(y)
        d_xy = D[x, y]
        chain.pop()
        chain.pop()
        if x > y:
            x, y = y, x
        nx, ny = size[x], size[y]
        merges[k] = (x, y, d_xy)
        # scipy _ward: sqrt((ni+nx) t dxi^2 + (ni+ny) t dyi^2 - ni t dxy^2), t = 1 / (nx + ny + ni)
        t = 1.0 / (nx + ny + size)
        with np.errstate(invalid='ignore'):
            new = np.sqrt((size + nx) * t * D[x] * D[x] + (size + ny) * t * D[y] * D[y] - size * t * d_xy * d_xy)
        new[~active] = np.inf
        new[x] = new[y] = np.inf
        D[y, :] = new
        D[:, y] = new
        D[x, :] = np.inf
        D[:, x] = np.inf
        active[x] = False
        size[y] = nx + ny
        size[x] = 0.0
    order = np.argsort(merges[:, 2], kind='mergesort')
    merges = merges[order]
    # scipy `label`: union-find over the sorted merges -> ids of the merged clusters, point counts
    parent = np.arange(2 * n - 1)
    count = np.ones(2 * n - 1)
    Z = np.empty((n - 1, 4))

    def find(i):
        root = i
        while parent[root] != root:
            root = parent[root]
        while parent[i] != root:
            parent[i], i = root, parent[i]
        return root
    for k in range(n - 1):
        a, b = find(int(merges[k, 0])), find(int(merges[k, 1]))
        if a > b:
            a, b = b, a
        Z[k] = (a, b, merges[k, 2], count[a] + count[b])
        parent[a] = parent[b] = n + k
        count[n + k] = count[a] + count[b]
    return Z


def _canonical_labels(labels):
    """cluster numbers as scipy's cut_tree assigns them: clusters ranked by their first cell"""
    _, first, inv = np.unique(labels, return_index=True, return_inverse=True)
    return np.argsort(np.argsort(first))[inv.ravel()]


def _get_MPEAR(assignments):
    """libs/utils.py:100-130."""
    assignments = np.asarray(assignments)
    steps, cells = assignments.shape
    n_range = _candidate_cluster_numbers(assignments)
    if n_range.size == 0:
        return None
    rep, inverse, weight = _unique_profiles(assignments)
    if rep.size == cells and cells <= MAX_LINKAGE_CELLS:
        # no two cells share a profile: the reference's own route over all pairs of cells
        pc = _PairCounts(assignments)
        Z = linkage(pc.dist(), method='ward')
        cuts = cut_tree(Z, n_clusters=n_range)                   # [cells, candidates]
        total, same_pairs, same_counts = pc.sums(cuts.T)
    else:
        if rep.size > MAX_PROFILES:
            raise NotImplementedError(
                f'{rep.size} distinct assignment profiles among {cells} cells: the weighted ward linkage holds a '
                f'square float64 distance matrix on the host and is limited to {MAX_PROFILES} profiles '
                '(use fewer posterior samples or the MAP estimator)')
        n_range = n_range[n_range < rep.size]                    # (cut_tree mislabels n_clusters == points)
        if n_range.size == 0:
            return None
        pc = _PairCounts(assignments[:, rep])                    # pairs of distinct profiles
        Z = _ward_linkage_weighted(pc.dist_square(), weight)
        cuts_u = cut_tree(Z, n_clusters=n_range)                 # [profiles, candidates]
        total, same_pairs, same_counts = pc.sums(cuts_u.T, weight)
        same_pairs = same_pairs + int((weight.astype(np.int64) * (weight.astype(np.int64) - 1) // 2).sum())
        cuts = cuts_u[inverse]
    with np.errstate(divide='ignore', invalid='ignore'):
        scores = _mpear_scores(total, same_pairs, same_counts, steps, cells)
    best = int(np.argmax(np.where(np.isnan(scores), -np.inf, scores)))   # first of equal scores, as `>`
    if rep.size == cells and cells <= MAX_LINKAGE_CELLS:
        return cuts[:, best].copy()
    return _canonical_labels(cuts[:, best])


# --------------------------------------------------------------------- posterior estimator
def get_mean_hierarchy_assignment(assignments, params_full):
    """libs/utils.py:148-192.  Returns (assignment, genotypes DataFrame [M, N])."""
    assignments = np.asarray(assignments)
    steps = assignments.shape[0]
    assign = _get_MPEAR(assignments)
    clusters = np.unique(assign)
    params = np.zeros((clusters.size, params_full.shape[2]))
    for i, cluster in enumerate(clusters):
        member = assign == cluster
        own = assignments[:, member]
        rest = assignments[:, ~member]
        together = (own == own[:, :1]).all(axis=1)            # criterion 1: one cluster in the sample
        major = np.array([np.bincount(row).argmax() for row in own])
        alone = ~(rest == major[:, None]).any(axis=1)         # criterion 2: nobody else in it
        if together.any():
            keep = np.flatnonzero(together & alone) if (together & alone).any() else np.flatnonzero(together)
            for s in keep:
                # row of the cluster in the sample's parameter block (rows follow the sorted ids)
                params[i] += params_full[s][np.searchsorted(np.unique(rest[s]), major[s])]
            params[i] /= keep.size
        else:
            for s in range(steps):
                ids, cnt = np.unique(own[s], return_counts=True)
                rows = np.searchsorted(np.unique(assignments[s]), ids)
                params[i] += cnt @ params_full[s][rows]
            params[i] /= steps * own.shape[1]
    params_df = pd.DataFrame(params).T[assign]
    return assign, params_df


def get_latents_posterior(results, data, single_chains=False):
    if single_chains:
        return [_get_latents_posterior_chain(result, data) for result in results]
    return [_get_latents_posterior_chain(_concat_chain_results(results), data)]


def _concat_chain_results(results):
    """libs/utils.py:206-221."""
    out = {key: np.concatenate([r[key][r['burn_in']:] for r in results])
           for key in ('assignments', 'DP_alpha', 'ML', 'MAP', 'FN', 'FP')}
    blocks = [r['params'] for r in results]
    widest = max(b.shape[1] for b in blocks)
    out['params'] = np.concatenate([np.pad(b, [(0, 0), (0, widest - b.shape[1]), (0, 0)]) for b in blocks])
    out['burn_in'] = 0
    return out


def _geno_error_rates(geno, data):
    """FN / FP rates implied by rounded genotypes (DataFrame [M, N]) -- libs/utils.py:233-236."""
    g = geno.T.values.round()
    fn = (((g == 1) & (data == 0)).sum() + EPSILON) / (g.sum() + EPSILON)
    fp = (((g == 0) & (data == 1)).sum() + EPSILON) / ((1 - g).sum() + EPSILON)
    return fn, fp


def _get_latents_posterior_chain(result, data):
    """libs/utils.py:224-241.  The theta trace of a chain starts at the first step after burn-in
    (libs/MCMC.py:261-282) while the other traces keep their burn-in rows; the reference slices BOTH
    by burn_in here (`--single_chains`), which pairs assignment rows with the theta rows of burn_in
    steps later and runs off the end once burn_in > steps / 2.  This version pairs them row by row."""
    burn_in = result['burn_in']
    params = result['params']
    if params.shape[0] != result['assignments'].shape[0] - burn_in:
        params = params[burn_in:]                            # a trace that does keep its burn-in rows
    assign, geno = get_mean_hierarchy_assignment(result['assignments'][burn_in:], params)
    fn_geno, fp_geno = _geno_error_rates(geno, data)
    return {'a': _get_posterior_avg(result['DP_alpha'][burn_in:]), 'assignment': assign, 'genotypes': geno,
            'FN': _get_posterior_avg(result['FN'][burn_in:]), 'FP': _get_posterior_avg(result['FP'][burn_in:]),
            'FN_geno': fn_geno, 'FP_geno': fp_geno}


def _get_posterior_avg(data):
    return np.mean(data), np.std(data)


# ------------------------------------------------------------------------- point estimators
def get_latents_point(results, est, data, single_chains=False):
    if single_chains:
        return [_get_latents_point_chain(result, est, data) for result in results]
    scores = [np.max(r[est][r['burn_in']:]) for r in results]
    return [_get_latents_point_chain(results[int(np.argmax(scores))], est, data)]


def _get_latents_point_chain(result, est, data):
    """libs/utils.py:261-283: the sample with the highest ML / MAP trace after burn-in."""
    burn_in = result['burn_in']
    kept = int(np.argmax(result[est][burn_in:]))
    step = kept + burn_in
    assignment = np.asarray(result['assignments'][step]).tolist()
    names = np.unique(assignment)
    geno = pd.DataFrame(result['params'][kept][np.arange(names.size)], index=names).T[assignment]
    fn_geno, fp_geno = _geno_error_rates(geno, data)
    return {'step': step, 'a': result['DP_alpha'][step], 'assignment': assignment, 'genotypes': geno,
            'FN': result['FN'][step], 'FP': result['FP'][step], 'FN_geno': fn_geno, 'FP_geno': fp_geno}


# ------------------------------------------------------------------------------ convergence
def get_lugsail_batch_means_est(data_in, steps=None):
    """libs/utils.py:427-461 (Vats & Knudson 2018): PSRF from lugsail batch-means variances of the
    chains' ML traces; data_in = [(trace, burn_in), ...]."""
    tau, var, size = [], [], []
    for trace, burn_in in data_in:
        x = np.asarray(trace[burn_in:steps], dtype=np.float64)
        if x.size < 9:
            return np.inf
        b = int(x.size ** (1 / 2))
        mean = np.nanmean(x)
        tau.append(2 * get_tau_lugsail(b, x, mean) - get_tau_lugsail(b // 3, x, mean))
        var.append(np.nanvar(x, ddof=1))
        size.append(x.size)
    t_l, s, n = np.mean(tau), np.mean(var), np.round(np.mean(size))
    sigma_l = ((n - 1) * s + t_l) / n
    with np.errstate(divide='ignore', invalid='ignore'):
        r = np.sqrt(sigma_l / s)
    return np.inf if not np.isfinite(r) else r


def get_tau_lugsail(b, data, chain_mean):
    a = data.size // b
    batch_mean = np.nanmean(np.reshape(data[:a * b], (a, b)), axis=1)
    return (b / (a - 1)) * np.nansum(np.square(batch_mean - chain_mean))


def get_cutoff_lugsail(e, a=0.05):
    M = (4 * np.pi * chi2.ppf(1 - a, 1)) / (gamma(1 / 2) ** 2 * e ** 2)
    return np.sqrt(1 + 1 / M)
