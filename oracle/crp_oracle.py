"""TEST INFRASTRUCTURE -- not part of the product path.

CPU restatement (numpy, float64) of the BnpC per-step MCMC hot path: the
Bernoulli-with-errors likelihood, the Gibbs reassignment sweep, the
non-conjugate split-merge move, the Metropolis-Hastings updates of the cluster
parameters / error rates and the concentration-parameter update, plus the step
schedule that strings them together.  It is the CHECKER for the CUDA path
(tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference
legs) and must never be imported by the product package.

Parity status: PINNED.  tests/test_oracle_vs_reference.py (runs wherever
/root/reference is mounted) drives the unmodified reference and this
restatement from the same numpy seed and requires identical traces; the
committed fixtures under tests/golden/ were produced from the reference itself
by tests/golden/make_golden.py and are replayed against this file on every CPU
test run.

Every function cites the reference lines it follows (paths relative to the
reference checkout, cbg-ethz/BnpC v0.2.1).  The arithmetic is kept in the
reference's own order and dtypes (float32 cluster parameters, `1 - theta`
rounded in float32, float64 everywhere else) so agreement is bit-level given
the same random draws; all draws go through `self.rnd`, a numpy-legacy-like
object (oracle/rng_tape.py) so they can be recorded or injected.
"""
from collections import OrderedDict

import numpy as np
from scipy.special import gamma as _gamma_fn
from scipy.special import gammaln
from scipy.stats import beta as _beta_dist
from scipy.stats import gamma as _gamma_dist
from scipy.stats import truncnorm as _truncnorm

from oracle.rng_tape import LegacyRandom

# libs/CRP.py:11-14
EPS = np.finfo(np.float64).resolution
LOG_EPS = np.log(EPS)
THETA_LO = 1e-5
THETA_HI = 1 - THETA_LO
THETA_STEP_SD = np.array([0.1, 0.25, 0.5])     # libs/CRP.py:65


def _fp_mode():
    # libs/CRP.py:10
    return np.errstate(divide='raise', over='ignore', under='ignore',
                       invalid='raise')


def softmax_floor(lp):
    """libs/CRP.py:88-100 (_normalize_log_probs): probabilities of a vector of
    log weights, each floored at 1e-15."""
    top = np.nanargmax(lp)
    rest = np.arange(lp.size) != top
    try:
        tail = np.exp(lp[rest] - lp[top])
    except FloatingPointError:
        tail = np.exp(np.clip(lp[rest] - lp[top], LOG_EPS, 0))
    z = lp - lp[top] - np.log1p(np.nansum(tail))
    return np.exp(np.clip(z, LOG_EPS, 0))


def log_normalise_pair(lp):
    """libs/CRP.py:103-116 (_normalize_log)."""
    top = np.nanargmax(lp, axis=0)
    try:
        out = lp - lp[top] - np.log1p(np.nansum(
            np.exp(lp[np.arange(lp.size) != top] - lp[top])))
    except FloatingPointError:
        if lp[0] > lp[1]:
            return np.array([0, LOG_EPS])
        return np.array([LOG_EPS, 0])
    return out


def log_crp_weight(n_i, n, a):
    """libs/CRP.py:83-85 (log_CRP_prior)."""
    return np.log(n_i, dtype=np.float64) - np.log(n - 1 + a, dtype=np.float64)


class OracleCRP:
    """Fixed-error-rate model (reference class libs/CRP.py:17 `CRP`)."""

    learning = False

    def __init__(self, data, DP_alpha=(-1, -1), param_beta=(1, 1),
                 FN_error=EPS, FP_error=EPS, rnd=None):
        # libs/CRP.py:27-65
        self.data = data
        self.cells_total, self.muts_total = data.shape
        self.p, self.q = param_beta
        self.theta_prior = _beta_dist(self.p, self.q)
        self.flat_prior = bool(self.p == self.q == 1)
        b0 = _gamma_fn(self.p) * _gamma_fn(self.q + 1) / _gamma_fn(self.p + self.q + 1)
        b1 = _gamma_fn(self.p + 1) * _gamma_fn(self.q) / _gamma_fn(self.p + 1 + self.q)
        self.mix = np.array([b0, b1]) / (b0 + b1)          # _beta_mix_const
        self.FP = FP_error
        self.FN = FN_error
        if DP_alpha[0] < 0 or DP_alpha[1] < 0:
            self.DP_a_gamma = (np.sqrt(self.cells_total), 1)
        else:
            self.DP_a_gamma = DP_alpha
        self.alpha_prior = _gamma_dist(*self.DP_a_gamma)    # shape a, LOC b (sic)
        self.DP_a = self.alpha_prior.mean()
        self.crp_table = None
        self.assignment = None
        self.parameters = None
        self.cells_per_cluster = None
        self.rnd = rnd if rnd is not None else LegacyRandom()

    # ------------------------------------------------------------------ init
    def init(self, mode='random', assign=False):
        """libs/CRP.py:119-152 (only the modes a caller can select: a given
        assignment, or 'random')."""
        if assign:
            self.assignment = np.array(assign)
            labels, sizes = np.unique(assign, return_counts=True)
            self._relabel(labels, sizes)
            self.parameters = self._initial_theta('assign')
        elif mode == 'random':
            self.assignment = self.rnd.randint(0, self.cells_total,
                                               size=self.cells_total)
            labels, sizes = np.unique(self.assignment, return_counts=True)
            self._relabel(labels, sizes)
            self.parameters = self._initial_theta('random')
        else:
            raise TypeError(f'Unsupported Initialization: {mode}')
        self.refresh_crp_table()

    def _relabel(self, labels, sizes):
        # libs/CRP.py:123-127 / 143-147: bn.replace(assignment, cl[i], i) in
        # ascending label order (in place, so later labels see earlier rewrites)
        self.cells_per_cluster = OrderedDict()
        for i in range(labels.size):
            self.assignment[self.assignment == labels[i]] = i
            self.cells_per_cluster[i] = sizes[i]

    def _initial_theta(self, mode):
        # libs/CRP.py:155-180
        theta = np.zeros(self.data.shape)
        if mode == 'assign':
            for cl in self.cells_per_cluster:
                rows = self.data[np.where(self.assignment == cl)]
                theta[cl] = self.rnd.beta(
                    self.p + np.nansum(rows * 1, axis=0),
                    self.q + np.nansum((1 - rows) * 1, axis=0))
        else:
            k = np.unique(self.assignment)
            theta[k] = self.rnd.uniform(size=(k.size, self.muts_total))
        return np.clip(theta, THETA_LO, THETA_HI).astype(np.float32)

    def draw_theta_for(self, cells):
        """libs/CRP.py:183-188 (_init_cl_params_new)."""
        rows = self.data[cells]
        draw = self.rnd.beta(self.p + np.nansum(rows * 1, axis=0),
                             self.q + np.nansum((1 - rows) * 1, axis=0))
        return np.clip(draw, THETA_LO, THETA_HI).astype(np.float32)

    def refresh_crp_table(self):
        """libs/CRP.py:191-194 (init_DP_prior): [0, log 1..log N, log alpha] -
        log(N-1+alpha), with a leading 0."""
        n = np.append(np.arange(1, self.cells_total + 1), self.DP_a)
        self.crp_table = np.append(0, log_crp_weight(n, self.cells_total, self.DP_a))

    # ------------------------------------------------------------ likelihood
    def _obs_fn(self, x):
        return (1 - self.FN) ** x * self.FN ** (1 - x)     # libs/CRP.py:207-208

    def _obs_fp(self, x):
        return (1 - self.FP) ** (1 - x) * self.FP ** x     # libs/CRP.py:211-212

    def loglik(self, x, theta, flat=False):
        """libs/CRP.py:197-204 (_calc_ll)."""
        mut = theta * self._obs_fn(x)
        wt = (1 - theta) * self._obs_fp(x)
        cell_ll = np.log(mut + wt)
        if flat:
            return np.nansum(cell_ll)
        return np.nansum(cell_ll, axis=1)

    def score_existing(self, cell, ids):
        """libs/CRP.py:223-227 (get_lpost_single)."""
        ll = self.loglik(self.data[[cell]], self.parameters[ids])
        sizes = np.fromiter(self.cells_per_cluster.values(), dtype=int)
        return ll + self.crp_table[sizes]

    def score_new_cluster(self):
        """libs/CRP.py:230-234 (get_lpost_single_new_cluster)."""
        wt = self.mix[0] * self._obs_fp(self.data)
        mut = self.mix[1] * self._obs_fn(self.data)
        return np.nansum(np.log(mut + wt), axis=1) + self.crp_table[-1]

    def get_ll_full(self):
        """libs/CRP.py:237-238."""
        with _fp_mode():
            return self.loglik(self.data, self.parameters[self.assignment], True)

    def get_lprior_full(self):
        """libs/CRP.py:241-251."""
        with _fp_mode():
            sizes = np.fromiter(self.cells_per_cluster.values(), dtype=int)
            lp = self.alpha_prior.logpdf(self.DP_a) + np.nansum(self.crp_table[sizes])
            if not self.flat_prior:
                ids = np.fromiter(self.cells_per_cluster.keys(), dtype=int)
                lp += np.nansum(self.theta_prior.logpdf(self.parameters[ids]))
            return lp

    # ------------------------------------------------------------ Gibbs sweep
    def update_assignments_Gibbs(self):
        """libs/CRP.py:254-288."""
        with _fp_mode():
            fresh = self.score_new_cluster()
            for cell in self.rnd.permutation(self.cells_total):
                was = self.assignment[cell]
                if self.cells_per_cluster[was] == 1:
                    del self.cells_per_cluster[was]
                else:
                    self.cells_per_cluster[was] -= 1
                ids = np.fromiter(self.cells_per_cluster.keys(), dtype=int)
                weights = softmax_floor(
                    np.append(self.score_existing(cell, ids), fresh[cell]))
                pick = self.rnd.choice(np.append(ids, -1), p=weights)
                if pick == -1:
                    pick = self.open_cluster(cell)
                self.assignment[cell] = pick
                if pick in self.cells_per_cluster:
                    self.cells_per_cluster[pick] += 1
                else:
                    self.cells_per_cluster[pick] = 1

    def open_cluster(self, cell):
        """libs/CRP.py:291-294 (init_new_cluster)."""
        new_id = self.lowest_free_id()
        self.parameters[new_id] = self.draw_theta_for([cell])
        return new_id

    def lowest_free_id(self):
        """libs/CRP.py:297-299 (get_empty_cluster)."""
        i = 0
        while i in self.cells_per_cluster:
            i += 1
        return i

    # ---------------------------------------------------------------- MH theta
    def update_parameters(self, step_no=None):
        """libs/CRP.py:302-311.  Returns (declined, accepted)."""
        with _fp_mode():
            declined = np.zeros(len(self.cells_per_cluster), dtype=int)
            for i, cl in enumerate(self.cells_per_cluster):
                members = np.argwhere(self.assignment == cl).flatten()
                self.parameters[cl], _, declined[i] = self.mh_theta_row(
                    self.parameters[cl], members)
            return np.nansum(declined), np.nansum(self.muts_total - declined)

    def mh_theta_row(self, old, cells, want_logq=False):
        """libs/CRP.py:314-344 (MH_cluster_params)."""
        sd = self.rnd.choice(THETA_STEP_SD, size=self.muts_total)
        lo = (THETA_LO - old) / sd
        hi = (THETA_HI - old) / sd
        prop = self.rnd.truncnorm_rvs(lo, hi, loc=old, scale=sd,
                                      size=self.muts_total).astype(np.float32)
        A = self.mh_theta_log_ratio(prop, old, cells, lo, hi, sd, want_logq)
        u = np.log(self.rnd.random(self.muts_total))
        rejected = u >= A
        prop[rejected] = old[rejected]
        if want_logq:
            A[rejected] = np.log(-1 * np.expm1(A[rejected]))
            return prop, np.nansum(A), np.nansum(rejected)
        return prop, np.nan, np.nansum(rejected)

    def mh_theta_log_ratio(self, new, old, cells, lo, hi, sd, clip=False):
        """libs/CRP.py:347-383 (_get_log_A)."""
        fwd = _truncnorm.logpdf(new, lo, hi, loc=old, scale=sd)
        lo_r = (THETA_LO - new) / sd
        hi_r = (THETA_HI - new) / sd
        rev = _truncnorm.logpdf(old, lo_r, hi_r, loc=new, scale=sd)
        x = self.data[cells]
        o_fn = self._obs_fn(x)
        o_fp = self._obs_fp(x)
        ll_new = np.nansum(np.log(new * o_fn + (1 - new) * o_fp), axis=0)
        ll_old = np.nansum(np.log(old * o_fn + (1 - old) * o_fp), axis=0)
        if self.flat_prior:
            pr_new = 0
            pr_old = 0
        else:
            pr_new = self.theta_prior.logpdf(new)
            pr_old = self.theta_prior.logpdf(old)
        A = ll_new + pr_new - ll_old - pr_old + rev - fwd
        if clip:
            return np.clip(A, a_min=None, a_max=0)
        return A

    # --------------------------------------------------------------- DP alpha
    def update_DP_alpha(self):
        """libs/CRP.py:386-410 (Escobar & West 1995).  The rate is handed to
        numpy's gamma as its SCALE argument, as in the reference."""
        with _fp_mode():
            k = len(self.cells_per_cluster)
            eta = self.rnd.beta(self.DP_a + 1, self.cells_total)
            w = (self.DP_a_gamma[0] + k - 1) \
                / (self.cells_total * (self.DP_a_gamma[1] - np.log(eta)))
            pi_eta = w / (1 + w)
            if self.rnd.random() < pi_eta:
                draw = self.rnd.gamma(self.DP_a_gamma[0] + k,
                                      self.DP_a_gamma[1] - np.log(eta))
            else:
                draw = self.rnd.gamma(self.DP_a_gamma[0] + k - 1,
                                      self.DP_a_gamma[1] - np.log(eta))
            self.DP_a = max(1 + EPS, draw)
            self.refresh_crp_table()

    # ------------------------------------------------------------ split-merge
    def update_assignments_split_merge(self, ratios=(.75, .25), step_no=5):
        """libs/CRP.py:417-431.  Returns ([accepted, declined], move)."""
        with _fp_mode():
            k = len(self.cells_per_cluster)
            if k == 1:
                return (self.try_split(step_no), 0)
            if k == self.cells_total:
                return (self.try_merge(step_no), 1)
            move = self.rnd.choice([0, 1], p=ratios)
            if move == 0:
                return (self.try_split(step_no), move)
            return (self.try_merge(step_no), move)

    def try_split(self, scans):
        """libs/CRP.py:434-481 (do_split_move)."""
        ids = np.fromiter(self.cells_per_cluster.keys(), dtype=int)
        sizes = np.fromiter(self.cells_per_cluster.values(), dtype=int)
        weight = sizes / sizes.sum()
        while True:
            target = self.rnd.choice(ids, p=weight)
            cells = np.argwhere(self.assignment == target).flatten()
            if cells.size != 1:
                break
        a_i, a_j = self.rnd.choice(cells.size, size=2, replace=False)
        cells[0], cells[a_i] = cells[a_i], cells[0]
        cells[-1], cells[a_j] = cells[a_j], cells[-1]

        where = np.argwhere(ids == target).flatten()
        lq_pick = np.log(weight[where]) \
            - np.log(self.cells_per_cluster[target]) \
            - np.log(self.cells_per_cluster[target] - 1)
        others = np.delete(sizes, where)

        ok, halves, theta2 = self.restricted_gibbs('split', cells,
                                                   (lq_pick, others), scans)
        if not ok:
            return [0, 1]
        new_id = self.lowest_free_id()
        self.parameters[target] = theta2[0]
        self.parameters[new_id] = theta2[1]
        moved = np.append(cells[1:-1][np.where(halves == 1)], cells[-1])
        self.assignment[moved] = new_id
        self.cells_per_cluster[target] -= moved.size
        self.cells_per_cluster[new_id] = moved.size
        return [1, 0]

    def try_merge(self, scans):
        """libs/CRP.py:484-524 (do_merge_move)."""
        ids = np.fromiter(self.cells_per_cluster.keys(), dtype=int)
        sizes = np.fromiter(self.cells_per_cluster.values(), dtype=int)
        inv = 1 / sizes
        weight = inv / inv.sum()
        cl_i, cl_j = self.rnd.choice(ids, p=weight, size=2, replace=False)

        cells_i = np.argwhere(self.assignment == cl_i).flatten()
        a_i = self.rnd.choice(cells_i.size)
        cells_i[0], cells_i[a_i] = cells_i[a_i], cells_i[0]
        cells_j = np.argwhere(self.assignment == cl_j).flatten()
        a_j = self.rnd.choice(cells_j.size)
        cells_j[-1], cells_j[a_j] = cells_j[a_j], cells_j[-1]
        cells = np.concatenate((cells_i, cells_j)).flatten()

        both = np.argwhere((ids == cl_j) | (ids == cl_i)).flatten()
        lq_pick = np.nansum(np.log(weight[both])) - np.nansum(np.log(sizes[both]))

        ok, theta1 = self.restricted_gibbs('merge', cells, lq_pick, scans)
        if not ok:
            return [0, 1]
        self.parameters[cl_i] = theta1
        self.assignment[cells_j] = cl_i
        self.cells_per_cluster[cl_i] += cells_j.size
        del self.cells_per_cluster[cl_j]
        return [1, 0]

    def restricted_gibbs(self, move, cells, size_term, scans):
        """libs/CRP.py:527-544 (run_rg_nc)."""
        self.launch_split(cells)
        self.rg_theta_merged = self.draw_theta_for(cells)
        for _ in range(scans):
            self.scan_split(cells)
            self.scan_merged(cells)
        if move == 'split':
            return self.decide_split(cells, size_term)
        return self.decide_merge(cells, size_term)

    def launch_split(self, cells):
        """libs/CRP.py:547-567 (_rg_init_split, random=False): each free cell
        goes to the anchor whose RAW data row (NaN -> mix[0]) explains it
        better; then theta of both halves is drawn."""
        i, j, free = cells[0], cells[-1], cells[1:-1]
        if free.size == 0:
            self.rg_half = np.array([])
        else:
            ll_i = self.loglik(self.data[free],
                               np.nan_to_num(self.data[i], nan=self.mix[0]))
            ll_j = self.loglik(self.data[free],
                               np.nan_to_num(self.data[j], nan=self.mix[0]))
            self.rg_half = np.where(ll_j > ll_i, 1, 0)
        side_i = np.append(free[np.argwhere(self.rg_half == 0)], i)
        side_j = np.append(free[np.argwhere(self.rg_half == 1)], j)
        th_i = self.draw_theta_for(side_i)
        th_j = self.draw_theta_for(side_j)
        self.rg_theta_split = np.stack([th_i, th_j])

    def scan_split(self, cells, want_logq=False):
        """libs/CRP.py:570-578 (_rg_scan_split)."""
        if cells.size == 2:
            lq_assign = 0
        else:
            lq_assign = self.scan_split_assign(cells, want_logq)
        lq_theta = self.scan_split_theta(cells, want_logq)
        if want_logq:
            return lq_assign + lq_theta

    def scan_merged(self, cells, want_logq=False):
        """libs/CRP.py:581-587 (_rg_scan_merge)."""
        self.rg_theta_merged, lq, _ = self.mh_theta_row(
            self.rg_theta_merged, cells, want_logq)
        if want_logq:
            return lq

    def scan_split_theta(self, cells, want_logq=False):
        """libs/CRP.py:590-606 (_rg_scan_params)."""
        i, j, free = cells[0], cells[-1], cells[1:-1]
        lq = np.zeros(2)
        for side in range(2):
            if side == 0:
                members = np.append(free[np.argwhere(self.rg_half == 0)], i)
            else:
                members = np.append(free[np.argwhere(self.rg_half == 1)], j)
            self.rg_theta_split[side], lq[side], _ = self.mh_theta_row(
                self.rg_theta_split[side], members, want_logq)
        if want_logq:
            return lq.sum()

    def scan_split_assign(self, cells, want_logq=False):
        """libs/CRP.py:609-632 (_rg_scan_assign)."""
        ll = self.pair_loglik(cells[1:-1], self.rg_theta_split)
        n = cells.size
        if want_logq:
            lq = np.zeros(n - 2)
        for c in self.rnd.permutation(n - 2):
            self.rg_half[c] = -1
            n_j = np.nansum(self.rg_half) + 2
            n_i = n - n_j - 1
            lpost = ll[c] + log_crp_weight([n_i, n_j], n, self.DP_a)
            lprob = log_normalise_pair(lpost)
            side = self.rnd.choice([0, 1], p=np.exp(lprob))
            self.rg_half[c] = side
            if want_logq:
                lq[c] = lprob[side]
        if want_logq:
            return np.nansum(lq)

    def pair_loglik(self, cells, theta2):
        """libs/CRP.py:635-638 (_rg_get_ll)."""
        return np.stack([self.loglik(self.data[cells], theta2[0]),
                         self.loglik(self.data[cells], theta2[1])], axis=1)

    def decide_split(self, cells, size_term):
        """libs/CRP.py:641-653 (_do_rg_split_MH)."""
        A = self.logq_ratio_split(cells) \
            + self.lprior_ratio_split(cells) \
            + self.ll_ratio(cells, 'split') \
            + self.lq_size_ratio_split(*size_term)
        if np.unique(self.rg_half).size == 1:
            return (False, [], [])
        if np.log(self.rnd.random()) < A:
            return (True, self.rg_half, self.rg_theta_split)
        return (False, [], [])

    def decide_merge(self, cells, size_term):
        """libs/CRP.py:656-665 (_do_rg_merge_MH)."""
        A = self.logq_ratio_merge(cells) \
            + self.lprior_ratio_merge(cells) \
            + self.ll_ratio(cells, 'merge') \
            + self.lq_size_ratio_merge(size_term)
        if np.log(self.rnd.random()) < A:
            return (True, self.rg_theta_merged)
        return (False, [])

    def logq_ratio_split(self, cells):
        """libs/CRP.py:668-682 (_get_trans_prob_ratio_split)."""
        fwd = self.scan_split(cells, want_logq=True)
        sd = self.rnd.choice(THETA_STEP_SD, size=self.muts_total)
        lo = (THETA_LO - self.rg_theta_merged) / sd
        hi = (THETA_HI - self.rg_theta_merged) / sd
        back = np.nansum(self.mh_theta_log_ratio(
            self.parameters[self.assignment[cells[0]]], self.rg_theta_merged,
            cells, lo, hi, sd, True))
        return back - fwd

    def logq_ratio_merge(self, cells):
        """libs/CRP.py:685-692 (_get_trans_prob_ratio_merge)."""
        fwd = self.scan_merged(cells, want_logq=True)
        back = self.logq_back_to_split(cells)
        return back - fwd

    def lprior_ratio_split(self, cells):
        """libs/CRP.py:695-713 (_get_lprior_ratio_split)."""
        n = self.rg_half.size + 2
        n_j = np.nansum(self.rg_half) + 1
        n_i = n - n_j
        r = np.log(self.DP_a) - gammaln(n)
        if n_i > 0:
            r += gammaln(n_j)
        if n_j > 0:
            r += gammaln(n_i)
        if not self.flat_prior:
            cl = self.assignment[cells[0]]
            r += np.nansum(self.theta_prior.logpdf(self.rg_theta_split)) \
                - np.nansum(self.theta_prior.logpdf(self.parameters[cl]))
        return r

    def ll_ratio(self, cells, move):
        """libs/CRP.py:716-733 (_get_ll_ratio)."""
        side_i = np.append(cells[1:-1][np.argwhere(self.rg_half == 0)], cells[0])
        side_j = np.append(cells[1:-1][np.nonzero(self.rg_half)], cells[-1])
        ll_i = self.loglik(self.data[side_i], self.rg_theta_split[0], True)
        ll_j = self.loglik(self.data[side_j], self.rg_theta_split[1], True)
        ll_all = self.loglik(self.data[cells], self.rg_theta_merged, True)
        if move == 'split':
            return ll_i + ll_j - ll_all
        return ll_all - ll_i - ll_j

    def lprior_ratio_merge(self, cells):
        """libs/CRP.py:736-754 (_get_lprior_ratio_merge)."""
        n = cells.size
        n_j = np.nansum(self.rg_half) + 1
        n_i = n - n_j
        r = gammaln(n) - np.log(self.DP_a)
        if n_i > 0:
            r -= gammaln(n_i)
        if n_j > 0:
            r -= gammaln(n_j)
        if not self.flat_prior:
            cl = self.assignment[[cells[0], cells[-1]]]
            r += np.nansum(self.theta_prior.logpdf(self.rg_theta_merged)) \
                - np.nansum(self.theta_prior.logpdf(self.parameters[cl]))
        return r

    def lq_size_ratio_split(self, lq_pick, others):
        """libs/CRP.py:757-764 (_get_ltrans_prob_size_ratio_split)."""
        n_j = np.nansum(self.rg_half) + 1
        n_i = self.rg_half.size + 2 - n_j
        norm = np.nansum(1 / np.append(others, [n_i, n_j]))
        back = np.log(1 / n_i / norm) + np.log(1 / n_j / norm)
        return back - lq_pick[0]

    def lq_size_ratio_merge(self, lq_pick):
        """libs/CRP.py:767-774 (_get_ltrans_prob_size_ratio_merge)."""
        try:
            back = -np.log(self.cells_total) - np.log(self.rg_half.size - 1)
        except FloatingPointError:
            back = -np.log(self.cells_total)
        return back - lq_pick

    def logq_back_to_split(self, cells):
        """libs/CRP.py:777-820 (_rg_get_split_prob): probability of moving from
        the launch split state to the ORIGINAL two clusters.  Proposal
        truncation here is [0,1], not [1e-5,1-1e-5] (reference quirk); on exit
        rg_half equals the original split."""
        sd = self.rnd.choice(THETA_STEP_SD, size=(2, self.muts_total))
        lo = (0 - self.rg_theta_split) / sd
        hi = (1 - self.rg_theta_split) / sd
        i, j, free = cells[0], cells[-1], cells[1:-1]
        cl_i = self.assignment[i]
        cl_j = self.assignment[j]
        lq_i = np.nansum(self.mh_theta_log_ratio(
            self.parameters[cl_i], self.rg_theta_split[0],
            np.append(free[np.argwhere(self.rg_half == 0)], i),
            lo[0], hi[0], sd[0], True))
        lq_j = np.nansum(self.mh_theta_log_ratio(
            self.parameters[cl_j], self.rg_theta_split[1],
            np.append(free[np.argwhere(self.rg_half == 1)], j),
            lo[1], hi[1], sd[1], True))
        ll = self.pair_loglik(free, (self.parameters[cl_i], self.parameters[cl_j]))
        n = cells.size
        lq = np.zeros(free.size)
        orig = np.where(self.assignment[free] == cl_i, 0, 1)
        for c in range(free.size):
            self.rg_half[c] = -1
            n_j = np.nansum(self.rg_half) + 2
            n_i = n - n_j - 1
            lpost = ll[c] + log_crp_weight([n_i, n_j], n, self.DP_a)
            lprob = log_normalise_pair(lpost)
            self.rg_half[c] = orig[c]
            lq[c] = lprob[orig[c]]
        return lq_i + lq_j + np.nansum(lq)


class OracleCRPLearnErrors(OracleCRP):
    """Learned FN/FP rates (reference libs/CRP_learning_errors.py:17
    `CRP_errors_learning`)."""

    learning = True

    def __init__(self, data, DP_alpha=(-1, -1), param_beta=(1, 1), FP_mean=0.001,
                 FP_sd=0.0005, FN_mean=0.25, FN_sd=0.05, rnd=None):
        # libs/CRP_learning_errors.py:18-32
        super().__init__(data, DP_alpha, param_beta, FN_mean, FP_mean, rnd=rnd)
        self.FP_prior = _truncnorm((0 - FP_mean) / FP_sd, (1 - FP_mean) / FP_sd,
                                   FP_mean, FP_sd)
        self.FP_steps = np.array([FP_sd * 0.5, FP_sd, FP_sd * 1.5])
        self.FN_prior = _truncnorm((0 - FN_mean) / FN_sd, (1 - FN_mean) / FN_sd,
                                   FN_mean, FN_sd)
        self.FN_steps = np.array([FN_sd * 0.5, FN_sd, FN_sd * 1.5])

    def get_lprior_full(self):
        # libs/CRP_learning_errors.py:47-49
        return super().get_lprior_full() \
            + self.FP_prior.logpdf(self.FP) + self.FN_prior.logpdf(self.FN)

    def update_error_rates(self):
        # libs/CRP_learning_errors.py:52-55.  Returns ([acc,dec]_FP, [acc,dec]_FN)
        with _fp_mode():
            self.FP, fp_count = self.mh_error_rate('FP')
            self.FN, fn_count = self.mh_error_rate('FN')
            return fp_count, fn_count

    def loglik_at(self, FP, FN):
        # libs/CRP_learning_errors.py:58-63 (get_ll_full_error)
        th = self.parameters[self.assignment]
        mut = th * (1 - FN) ** self.data * FN ** (1 - self.data)
        wt = (1 - th) * (1 - FP) ** (1 - self.data) * FP ** self.data
        return np.nansum(np.log(mut + wt))

    def mh_error_rate(self, which):
        # libs/CRP_learning_errors.py:66-111 (MH_error_rates)
        if which == 'FP':
            cur, prior, steps = self.FP, self.FP_prior, self.FP_steps
        else:
            cur, prior, steps = self.FN, self.FN_prior, self.FN_steps
        sd = self.rnd.choice(steps)
        lo = (0 - cur) / sd
        hi = (1 - cur) / sd
        try:
            prop = self.rnd.truncnorm_rvs(lo, hi, loc=cur, scale=sd)
        except FloatingPointError:
            prop = self.rnd.truncnorm_rvs(lo, np.inf, loc=cur, scale=sd)
        fwd = _truncnorm.logpdf(prop, lo, hi, loc=cur, scale=sd)
        lo_r, hi_r = (0 - prop) / sd, (1 - prop) / sd
        rev = _truncnorm.logpdf(cur, lo_r, hi_r, loc=prop, scale=sd)
        if which == 'FP':
            ll_new = self.loglik_at(prop, self.FN)
            ll_old = self.loglik_at(cur, self.FN)
        else:
            ll_new = self.loglik_at(self.FP, prop)
            ll_old = self.loglik_at(self.FP, cur)
        A = ll_new + prior.logpdf(prop) - ll_old - prior.logpdf(cur) + rev - fwd
        if np.log(self.rnd.random()) < A:
            return prop, [1, 0]
        return cur, [0, 1]


# -----------------------------------------------------------------------------
# step schedule (reference libs/MCMC.py:320-342 `Chain.do_step` and :242-258
# `Chain.update_results`) -- used to drive reference, oracle and CUDA model alike
# -----------------------------------------------------------------------------
DEFAULT_MOVES = dict(sm_prob=0.33, dpa_prob=0.25, error_prob=0.25,
                     sm_ratios=[0.75, 0.25], sm_steps=3)   # run_BnpC.py CLI defaults


def do_step(model, rnd, moves, learning, fix_assign=False):
    """One MCMC step in the reference's order; returns a dict of what happened.
    `model` is any object with the CRP method contract (reference class, oracle,
    CUDA-backed class); `rnd` supplies the move-selection uniforms."""
    log = {'move': 'gibbs', 'sm': None, 'alpha': False, 'errors': None}
    if not fix_assign:
        if rnd.random() < moves['sm_prob']:
            res, kind = model.update_assignments_split_merge(
                moves['sm_ratios'], moves['sm_steps'])
            log['move'] = 'split' if kind == 0 else 'merge'
            log['sm'] = [int(res[0]), int(res[1])]
        else:
            model.update_assignments_Gibbs()
        if rnd.random() < moves['dpa_prob']:
            model.update_DP_alpha()
            log['alpha'] = True
    dec, acc = model.update_parameters()
    log['theta'] = [int(dec), int(acc)]
    if learning and rnd.random() < moves['error_prob']:
        fp, fn = model.update_error_rates()
        log['errors'] = [list(map(int, fp)), list(map(int, fn))]
    return log


def snapshot(model):
    """State that the chain driver records every step (libs/MCMC.py:252-258)
    plus the full live-cluster bookkeeping, as plain numpy."""
    ids = np.fromiter(model.cells_per_cluster.keys(), dtype=np.int64)
    sizes = np.fromiter(model.cells_per_cluster.values(), dtype=np.int64)
    ll = float(model.get_ll_full())
    return dict(
        assignment=np.array(model.assignment, dtype=np.int64),
        ids=ids, sizes=sizes,
        theta=np.array(model.parameters[ids], dtype=np.float32),
        alpha=float(model.DP_a), FN=float(model.FN), FP=float(model.FP),
        ll=ll, lpost=ll + float(model.get_lprior_full()))


def simulate(n_cells, n_muts, k_true=20, fn=0.2, fp=0.01, miss=0.10, seed=0,
             geno_p=0.3):
    """Synthetic benchmark matrix: K_true Bernoulli(0.3)
    genotypes, uniform cluster membership, FN/FP flips, then missing entries.
    Returns (data float64 [N,M] with NaN, z_true)."""
    rng = np.random.default_rng(seed)
    geno = (rng.random((k_true, n_muts)) < geno_p)
    z = rng.integers(0, k_true, size=n_cells)
    truth = geno[z]
    flip = rng.random((n_cells, n_muts))
    obs = np.where(truth, flip >= fn, flip < fp).astype(np.float64)
    obs[rng.random((n_cells, n_muts)) < miss] = np.nan
    return obs, z
