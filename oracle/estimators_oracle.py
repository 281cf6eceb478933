"""TEST INFRASTRUCTURE -- not part of the product path.

CPU restatement (numpy / scipy) of the reference's point and posterior estimators,
cbg-ethz/BnpC v0.2.1 `libs/utils.py`: the co-clustering distance,
the MPEAR search over cuts of a ward dendrogram, the genotype averaging of the posterior
estimator and the MAP / ML point estimators.  Pinned against the unmodified reference by
tests/test_oracle_estimators.py (run here, where /root/reference is mounted) and by the golden
fixtures tests/golden/estimators_*.npz generated from the reference
(tests/golden/make_golden_estimators.py).  Only tests/ may import this module.
"""
import numpy as np
from scipy.cluster.hierarchy import cut_tree, linkage
from scipy.special import binom

EPSILON = np.finfo(np.float64).resolution          # libs/utils.py:16


def pair_counts(assignments):
    """Number of samples in which cells i < j sit in DIFFERENT clusters, condensed in pdist
    order (the integer the reference accumulates in libs/utils.py:90-96 before dividing)."""
    steps, cells = assignments.shape
    iu = np.triu_indices(cells, k=1)
    cnt = np.zeros(iu[0].size, dtype=np.int32)
    for a in assignments:
        cnt += (a[iu[0]] != a[iu[1]]).astype(np.int32)
    return cnt


def get_dist(assignments):
    """libs/utils.py:90-97: mean posterior cell-wise Hamming distance, condensed."""
    return pair_counts(assignments) / assignments.shape[0]


def calc_mpear(sim, c):
    """libs/utils.py:133-145 (Fritsch & Ickstadt 2009, eq. 13) on a condensed similarity."""
    iu = np.triu_indices(c.size, k=1)
    same = (c[iu[0]] == c[iu[1]]).astype(np.float64)
    i_sum = same.sum()
    pi_sum = sim.sum()
    index = (same * sim).sum()
    expected = (i_sum * pi_sum) / binom(c.size, 2)
    max_index = .5 * (i_sum + pi_sum)
    return (index - expected) / (max_index - expected)


def cluster_number_range(assignments):
    """libs/utils.py:106-114: candidate cluster numbers around the mean number of clusters
    with more than two cells."""
    cl_no = [int((np.unique(a, return_counts=True)[1] > 2).sum()) for a in assignments]
    avg = np.mean(cl_no)
    return np.arange(max(2, avg * 0.2), min(avg * 2.5, assignments.shape[1]), dtype=int)


def get_mpear_assignment(assignments):
    """libs/utils.py:100-130: ward linkage on the distances, cut at every candidate number of
    clusters, keep the first cut with the highest MPEAR score."""
    dist = get_dist(assignments)
    sim = 1 - dist
    Z = linkage(dist, method='ward')
    best, best_score = None, -np.inf
    for n in cluster_number_range(assignments):
        clusters = cut_tree(Z, n_clusters=n).flatten()
        score = calc_mpear(sim, clusters)
        if score > best_score:
            best, best_score = clusters, score
    return best


def mean_hierarchy_genotypes(assignments, params_full, assign):
    """libs/utils.py:148-192 after the MPEAR call: per estimated cluster the parameters of the
    samples in which its cells form one cluster of their own (criteria 1 and 2 of the paper,
    section 2.3), else a cell-weighted mean over all samples.  Returns [clusters, M]."""
    steps = assignments.shape[0]
    clusters = np.unique(assign)
    params = np.zeros((clusters.size, params_full.shape[2]))
    for i, cluster in enumerate(clusters):
        in_cl = assign == cluster
        cells = np.nonzero(in_cl)[0]
        other = np.nonzero(~in_cl)[0]
        sub = assignments[:, cells]
        if cells.size == 1:
            same_cluster = np.ones(steps, dtype=bool)
        else:
            same_cluster = (sub == sub[:, :1]).all(axis=1)       # zero moving std of window 2
        cl_ids = np.array([np.argmax(np.bincount(r)) for r in sub])
        other_ids = assignments[:, other]
        no_others = np.array([cl_ids[j] not in other_ids[j] for j in range(steps)], dtype=bool)
        if same_cluster.any():
            both = same_cluster & no_others
            step_idx = np.argwhere(both if both.any() else same_cluster).flatten()
            for step in step_idx:
                all_ids = np.append(np.unique(other_ids[step]), cl_ids[step])
                rel = np.argwhere(np.sort(all_ids) == cl_ids[step])[0][0]
                params[i] += params_full[step][rel]
            params[i] /= step_idx.size
        else:
            for step, step_assign in enumerate(assignments):
                all_ids = np.unique(step_assign)
                ids, cnt = np.unique(step_assign[cells], return_counts=True)
                rows = np.argwhere(np.isin(all_ids, ids)).flatten()
                params[i] += np.dot(cnt, params_full[step][rows])
            params[i] /= steps * cells.size
    return params


def error_rates_from_genotypes(geno_cells, data):
    """libs/utils.py:233-236: FN / FP rates implied by rounded genotypes [N, M] and the data."""
    g = np.round(geno_cells)
    fn = (((g == 1) & (data == 0)).sum() + EPSILON) / (g.sum() + EPSILON)
    fp = (((g == 0) & (data == 1)).sum() + EPSILON) / ((1 - g).sum() + EPSILON)
    return fn, fp


def latents_posterior_chain(result, data):
    """libs/utils.py:224-241 for one (possibly concatenated) chain; genotypes as [N, M]."""
    b = result['burn_in']
    assignments = result['assignments'][b:]
    assign = get_mpear_assignment(assignments)
    params = mean_hierarchy_genotypes(assignments, result['params'][b:], assign)
    geno = params[np.searchsorted(np.unique(assign), assign)]
    fn_g, fp_g = error_rates_from_genotypes(geno, data)
    return dict(a=(np.mean(result['DP_alpha'][b:]), np.std(result['DP_alpha'][b:])), assignment=assign,
                genotypes=geno, FN=(np.mean(result['FN'][b:]), np.std(result['FN'][b:])),
                FP=(np.mean(result['FP'][b:]), np.std(result['FP'][b:])), FN_geno=fn_g, FP_geno=fp_g)


def latents_point_chain(result, est, data):
    """libs/utils.py:261-283: the sample with the highest `est` ('MAP' or 'ML') trace."""
    b = result['burn_in']
    step_no_bi = int(np.argmax(result[est][b:]))
    step = step_no_bi + b
    assignment = np.asarray(result['assignments'][step])
    names = np.unique(assignment)
    geno_all = result['params'][step_no_bi][np.arange(names.size)]
    geno = geno_all[np.searchsorted(names, assignment)]
    fn_g, fp_g = error_rates_from_genotypes(geno, data)
    return dict(step=step, a=result['DP_alpha'][step], assignment=assignment, genotypes=geno,
                FN=result['FN'][step], FP=result['FP'][step], FN_geno=fn_g, FP_geno=fp_g)


def concat_chain_results(results):
    """libs/utils.py:206-221."""
    cat = {k: np.concatenate([r[k][r['burn_in']:] for r in results])
           for k in ('assignments', 'DP_alpha', 'ML', 'MAP', 'FN', 'FP')}
    params = [r['params'] for r in results]
    kmax = max(p.shape[1] for p in params)
    cat['params'] = np.concatenate([np.pad(p, [(0, 0), (0, kmax - p.shape[1]), (0, 0)]) for p in params])
    cat['burn_in'] = 0
    return cat
