// Native lockstep driver: all chains of one GPU stepped by ONE host thread (declared in
// include/bnpc_b200.h, "chain group" section).
//
// The reference forks one process per chain (libs/MCMC.py:113-120) and every process runs
// Chain.do_step / Chain.update_results (libs/MCMC.py:320-342, 242-282).  Here the chains of a GPU
// advance in lockstep: per step the host draws each chain's move, records the launches of all
// chains through the composite entry points (bnpc_chain_*), flushes them as chain-batched launches
// (bnpc_batch.cuh) and synchronises ONCE per phase for all chains:
//   phase 1  assignment move: Gibbs epochs (stream A) and split-merge moves (stream B) side by side
//   phase 2  DP alpha (host), sufficient statistics, MH on theta, the likelihoods of the error-rate
//            moves and of the trace row in one pass
//   trace    assignment vector and theta rows snapshot into device rings; a copy stream drains the
//            rings into the caller's pinned trace arrays while the next step runs
// The host algebra mirrors bnpc_b200/engine.py (the per-method Python mirror of the reference
// classes, pinned against the oracle with a random tape) line by line and draws from the same
// counter-based streams: a chain stepped here and a chain stepped through engine.py from the same
// seed produce the same trace (tests/test_gpu_group.py).
#pragma once
#include <math.h>

#include <algorithm>
#include <vector>

namespace bnpc {

// ---------------------------------------------------------------- host random numbers
static inline void philox4x32_10(uint64_t key, uint64_t c0, uint64_t c1, uint32_t out[4]) {
    uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
    uint32_t c[4] = {(uint32_t)c0, (uint32_t)(c0 >> 32), (uint32_t)c1, (uint32_t)(c1 >> 32)};
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}
// host scalars of a chain: draw i = Philox(seed)(i, HOST_STREAM); the device kernels use small
// stream ids (running call counters), never this one
constexpr uint64_t HOST_STREAM = 0xB057000000000000ull;
static inline double host_random(uint64_t seed, uint64_t* ctr) {
    uint32_t r[4];
    philox4x32_10(seed, (*ctr)++, HOST_STREAM, r);
    const uint64_t v = ((uint64_t)r[0] << 32) | r[1];
    return (double)(v >> 11) * (1.0 / 9007199254740992.0);
}
static inline double host_random_open(uint64_t seed, uint64_t* ctr) {
    double u;
    do { u = host_random(seed, ctr); } while (u <= 0.0);
    return u;
}
static double host_gamma(uint64_t seed, uint64_t* ctr, double shape) {
    // Marsaglia & Tsang (2000), as bnpc_math.cuh gamma_sample
    double boost = 1.0;
    if (shape < 1.0) { boost = pow(host_random_open(seed, ctr), 1.0 / shape); shape += 1.0; }
    const double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (;;) {
        double x, v;
        do {
            const double u1 = host_random_open(seed, ctr), u2 = host_random(seed, ctr);
            x = sqrt(-2.0 * log(u1)) * cos(2.0 * M_PI * u2);
            v = 1.0 + c * x;
        } while (v <= 0.0);
        v = v * v * v;
        const double u = host_random_open(seed, ctr);
        if (u < 1.0 - 0.0331 * (x * x) * (x * x)) return boost * d * v;
        if (log(u) < 0.5 * x * x + d * (1.0 - v + log(v))) return boost * d * v;
    }
}
static double host_beta(uint64_t seed, uint64_t* ctr, double a, double b) {
    const double x = host_gamma(seed, ctr, a), y = host_gamma(seed, ctr, b);
    const double t = x + y;
    return (t > 0.0) ? x / t : 0.5;
}

// ---------------------------------------------------------------- host truncated normal
// scipy.stats.truncnorm as libs/CRP_learning_errors.py:82-91 uses it (scalar error-rate moves):
// the formulas of bnpc_math.cuh with libm's erf/erfc
static inline double h_ndtr(double x) {
    const double z = x * M_SQRT1_2, az = fabs(z);
    if (az < M_SQRT1_2) return 0.5 + 0.5 * erf(z);
    const double y = 0.5 * erfc(az);
    return (z > 0) ? 1.0 - y : y;
}
static inline double h_log_ndtr(double x) {
    const double t = x * M_SQRT1_2;
    if (x < -1.0) return log(0.5 * erfc(-t));
    return log1p(-0.5 * erfc(t));
}
static inline double h_log_diff_exp(double lp, double lq) { return log(1.0 - exp(lq - lp)) + lp; }
static inline double h_log_gauss_mass(double a, double b) {
    if (b <= 0) return h_log_diff_exp(h_log_ndtr(b), h_log_ndtr(a));
    if (a > 0) return h_log_diff_exp(h_log_ndtr(-a), h_log_ndtr(-b));
    return log1p(-h_ndtr(a) - h_ndtr(-b));
}
// x with log(Phi(x)) = y for y <= log(1/2): Newton on the concave increasing log Phi from a
// tail/central first guess, to the last bit
static double h_ndtri_exp_lower(double y) {
    double x;
    if (y < -3.0) {
        x = -sqrt(-2.0 * y);                                  // Phi(x) ~ exp(-x^2/2)
        x = -sqrt(fmax(0.0, -2.0 * (y + log(-x) + 0.91893853320467274178)));
    } else {
        const double p = exp(y);
        x = -sqrt(-2.0 * log(2.0 * p)) * 0.8 - 0.0;         // crude, x = 0 at p = 1/2
        if (!(x == x)) x = 0.0;
    }
    for (int it = 0; it < 60; ++it) {
        const double f = h_log_ndtr(x) - y;
        const double dlog = exp(-0.5 * x * x - 0.91893853320467274178 - h_log_ndtr(x));   // phi / Phi
        double dx = f / dlog;
        if (dx > 2.0) dx = 2.0;
        if (dx < -2.0) dx = -2.0;
        x -= dx;
        if (fabs(dx) <= 4e-16 * fmax(1.0, fabs(x))) break;
    }
    return x;
}
static double h_ndtri_exp(double y) {
    if (y > -0.6931471805599453) return -h_ndtri_exp_lower(log(-expm1(y)));
    return h_ndtri_exp_lower(y);
}
static inline double h_log_sum_exp2(double a, double b) {
    if (a == b) return log(2.0) + a;
    const double top = a > b ? a : b, low = a > b ? b : a;
    return log1p(exp(low - top)) + top;
}
static double h_tn_ppf(double q, double lo, double hi) {
    if (lo < 0) return h_ndtri_exp(h_log_sum_exp2(h_log_ndtr(lo), log(q) + h_log_gauss_mass(lo, hi)));
    return -h_ndtri_exp(h_log_sum_exp2(h_log_ndtr(-hi), log1p(-q) + h_log_gauss_mass(lo, hi)));
}
static double h_tn_logpdf(double x, double lo, double hi, double loc, double scale) {
    const double y = (x - loc) / scale;
    if (!(lo <= y && y <= hi)) return -INFINITY;
    return ((-(y * y) / 2.0 - 0.91893853320467274178) - h_log_gauss_mass(lo, hi)) - log(scale);
}

// ---------------------------------------------------------------- one chain on the host
struct GroupChain {
    bnpc_chain_t* w;
    bnpc_chain_state_t* s;
    bnpc_epoch_t ep;
    bnpc_rg_t rg;
    std::vector<int> ids, sizes;           // live clusters in list order (libs/CRP.py:268)
    // this step
    int move;                              // -1 none, 0 Gibbs, 1 split, 2 merge
    bool move_done;
    // Gibbs sweep
    int t, first, epochs, stall, rows;
    bool lean, wide;
    // split-merge
    int sm_n, sm_na, sm_cl_i, sm_cl_j, sm_where;
    double sm_lq_pick;
    std::vector<double> sm_others;
    int sm_result[2];
    // parameters / errors / trace
    bool err_move;
    double fp_prop, fn_prop, fp_fwd, fp_rev, fn_fwd, fn_rev, fp_acc_u, fn_acc_u;
    int loglik_rows;
    bool stats_fresh;
    // asynchronous scheduler
    int state, step, sub;                  // CS_*, trace row of the step in hand, wave the chain travels in
    long long snaps;                       // trace rows snapshot so far (ring slot = snaps % ring)
    bool theta_direct;                     // more live clusters than a ring row holds: theta rows copied directly
    cudaEvent_t prewait;                   // event the chain's next wave must wait for (ring slot drained) or NULL

    uint64_t seed() const { return s->seed; }
    double random() { return host_random(s->seed, &s->host_ctr); }
    int randint(int high) { return (int)floor(random() * high); }
    uint64_t reserve(int n) { const uint64_t b = s->dev_calls; s->dev_calls += n; return b; }
    int K() const { return (int)ids.size(); }
};

struct Group {
    int n;
    std::vector<GroupChain> ch;
    bnpc_moves_t mv;
    bnpc_grow_fn grow;
    void* grow_ctx;
    cudaStream_t sA, sB, sC;
    cudaEvent_t evA, evB;                  // cross-stream ordering of the two phases
    // trace rings
    int ring;                              // slots
    int32_t* ring_assign;                  // [ring][n][N]
    float* ring_theta;                     // [ring][n][ring_kcap][M]
    int ring_kcap;
    std::vector<cudaEvent_t> ev_snap, ev_drained;
    long long steps_done, snaps;
    std::vector<bnpc_trace_t*> tr;
    // asynchronous scheduler: waves (sets of chains that became ready together and travel through
    // one phase on one stream) and per-(chain, ring slot) drain events
    struct Wave { cudaStream_t s; cudaEvent_t ev; unsigned long long mask; bool busy; };
    std::vector<Wave> waves;
    std::vector<cudaEvent_t> ev_slot;      // [n][ring]: ring slot drained to the host
    bool lockstep;
};

static int group_fail(Group* g, const char* what) {
    snprintf(g_err, sizeof(g_err), "group: %s", what);
    return 3;
}

static int pick_weighted(const std::vector<double>& p, double u) {
    // np.random.choice(a, p=p): cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(u, 'right')
    std::vector<double> cdf(p.size());
    double acc = 0.0;
    for (size_t i = 0; i < p.size(); ++i) { acc += p[i]; cdf[i] = acc; }
    const double last = cdf.back();
    int idx = 0;
    for (size_t i = 0; i < cdf.size(); ++i) {
        if (cdf[i] / last <= u) idx = (int)i + 1; else break;
    }
    return idx;
}

static void load_list(GroupChain& c) {
    c.ids.resize(c.s->K);
    c.sizes.resize(c.s->K);
    for (int j = 0; j < c.s->K; ++j) { c.ids[j] = c.s->live[2 * j]; c.sizes[j] = c.s->live[2 * j + 1]; }
}
static int store_list(Group* g, int ci) {
    GroupChain& c = g->ch[ci];
    if (2 * c.K() > c.s->live_cap) {
        if (g->grow(g->grow_ctx, ci, BNPC_GROW_LIVE, 2 * c.K() + 64)) return group_fail(g, "grow(live) failed");
        if (2 * c.K() > c.s->live_cap) return group_fail(g, "live list capacity");
    }
    c.s->K = c.K();
    for (int j = 0; j < c.K(); ++j) { c.s->live[2 * j] = c.ids[j]; c.s->live[2 * j + 1] = c.sizes[j]; }
    return 0;
}

static int ensure_ids(Group* g, int ci, int need) {
    GroupChain& c = g->ch[ci];
    if (need <= c.w->idcap) return 0;
    // (the chain itself is idle here; the caller's allocator may reuse the old buffers at once, so
    // nothing of this chain may still be in flight: recorded launches of OTHER chains are not touched)
    cudaDeviceSynchronize();
    if (g->grow(g->grow_ctx, ci, BNPC_GROW_IDS, need)) return group_fail(g, "grow(ids) failed");
    if (need > c.w->idcap) return group_fail(g, "cluster id capacity");
    c.stats_fresh = false;                 // S1/S0 were reallocated
    return 0;
}

static int get_empty_cluster(const GroupChain& c) {
    // libs/CRP.py:297-299
    int i = 0;
    for (;;) {
        if (std::find(c.ids.begin(), c.ids.end(), i) == c.ids.end()) return i;
        ++i;
    }
}

// ------------------------------------------------------------------ Gibbs sweep (engine.py
// update_assignments_Gibbs, libs/CRP.py:254-299) as an epoch state machine
static void gibbs_begin(Group* g, GroupChain& c) {
    const int N = c.w->N;
    bnpc_epoch_t& ep = c.ep;
    memset(&ep, 0, sizeof(ep));
    const double FN = c.s->FN, FP = c.s->FP, mix0 = c.s->mix0, mix1 = c.s->mix1;
    ep.c1 = log(mix1 * (1 - FN) + mix0 * FP);
    ep.c0 = log(mix1 * FN + mix0 * (1 - FP));
    ep.c_norm = log((double)N - 1 + c.s->DP_a);
    ep.lnew_prior = log(c.s->DP_a) - log((double)N - 1 + c.s->DP_a);
    ep.log_n = log((double)N);
    ep.FN = FN; ep.FP = FP; ep.p = c.s->p; ep.q = c.s->q;
    ep.rand_ready = 0; ep.beta_rows = nullptr; ep.n_beta_rows = 0;
    ep.seed = c.s->seed; ep.stream_id = c.reserve(3);
    ep.serial_sweep = c.s->serial_sweep;
    c.t = 0; c.first = 1; c.epochs = 0; c.stall = 0;
    if (!c.s->lean_ok) {
        c.s->lean_cooldown -= 1;
        if (c.s->lean_cooldown <= 0) c.s->lean_ok = 1;
    }
    (void)g;
}

static int gibbs_enqueue(Group* g, int ci) {
    GroupChain& c = g->ch[ci];
    const int N = c.w->N, K = c.K();
    if (int rc = ensure_ids(g, ci, K + BNPC_MAX_EXTRA + 2)) return rc;
    int32_t* h = c.w->h_in;
    for (int j = 0; j < K; ++j) { h[2 * j] = c.ids[j]; h[2 * j + 1] = c.sizes[j]; }
    const int ldk = std::max(3, K | 1);
    c.lean = K <= BNPC_LEAN_MAXK && c.s->lean_ok && c.s->lean_enabled && !(c.s->force_wide && K <= 63);
    c.wide = !c.lean && c.s->wide_enabled && K <= 63 && (c.s->force_wide || !c.s->lean_ok);
    const long long budget = (1ll << 30) / (8ll * ldk);
    c.rows = c.lean ? N - c.t : (int)std::min<long long>(N - c.t, std::max<long long>(1, budget));
    c.ep.lean = c.lean ? c.s->lean_rows : (c.wide ? -1 : 0);
    if (!c.lean) {
        const long long need = (long long)c.rows * ldk + 2, needx = (long long)BNPC_MAX_EXTRA * c.rows;
        if (need > c.s->ll_cap && (g->grow(g->grow_ctx, ci, BNPC_GROW_LL, need) || need > c.s->ll_cap))
            return group_fail(g, "grow(ll) failed");
        if (!c.wide && needx > c.s->llx_cap &&
            (g->grow(g->grow_ctx, ci, BNPC_GROW_LLX, needx) || needx > c.s->llx_cap))
            return group_fail(g, "grow(llx) failed");
    }
    c.ep.first = c.first; c.ep.K = K; c.ep.t = c.t; c.ep.rows = c.rows; c.ep.ldk = ldk;
    return bnpc_chain_gibbs_epoch(c.w, &c.ep, nullptr);
}

// after the synchronisation: returns 1 when the sweep is complete, 0 when another epoch is due
static int gibbs_after(Group* g, int ci, int* done) {
    GroupChain& c = g->ch[ci];
    const int N = c.w->N;
    const int32_t* st = c.w->h_out;
    const int flags = st[BNPC_ST_FLAGS];
    *done = 0;
    if (flags & (BNPC_STOP_TAPE_EMPTY | 0x100)) return group_fail(g, "gibbs_sweep stopped (tape/hang flag)");
    if (flags & BNPC_STOP_MANY) {
        c.s->lean_ok = 0;
        c.s->lean_cooldown = 8;
        return 0;                          // nothing was done: the same epoch again, dense
    }
    const int K = st[BNPC_ST_K];
    c.ids.resize(K);
    c.sizes.resize(K);
    const int32_t* pairs = c.w->h_out + BNPC_ST_WORDS;
    for (int j = 0; j < K; ++j) { c.ids[j] = pairs[2 * j]; c.sizes[j] = pairs[2 * j + 1]; }
    const int t_new = st[BNPC_ST_TDONE];
    if (c.lean && st[BNPC_ST_NMANY] > std::max(64, c.rows / 50)) {
        c.s->lean_ok = 0;
        c.s->lean_cooldown = 8;
    }
    c.stall = (t_new == c.t) ? c.stall + 1 : 0;
    if (c.stall > 2) return group_fail(g, "gibbs_sweep made no progress");
    if (flags & BNPC_STOP_CAPACITY) {
        if (int rc = ensure_ids(g, ci, 2 * c.w->idcap)) return rc;
    }
    c.t = t_new;
    c.first = 0;
    c.epochs += 1;
    if (c.t >= N) {
        *done = 1;
        c.s->last_epochs = c.epochs;
        c.s->last_births = st[BNPC_ST_BIRTHS];
        c.s->last_moved = st[BNPC_ST_MOVED];
        c.s->last_nunc = st[BNPC_ST_NUNC];
        c.stats_fresh = false;
    }
    return 0;
}

// ------------------------------------------------------------------ split-merge (engine.py
// update_assignments_split_merge, _try_split/_try_merge, _restricted_gibbs; libs/CRP.py:417-567)
static void rg_begin(GroupChain& c, int n, int n_a, int cl_i, int cl_j, int a_i, int a_j, int is_merge) {
    bnpc_rg_t& r = c.rg;
    memset(&r, 0, sizeof(r));
    r.n = n; r.n_a = n_a; r.cl_i = cl_i; r.cl_j = cl_j; r.a_i = a_i; r.a_j = a_j; r.is_merge = is_merge;
    r.rand_ready = 0;
    r.seed = c.s->seed;
    const double FN = c.s->FN, FP = c.s->FP, mix0 = c.s->mix0;
    r.alpha = c.s->DP_a; r.FN = FN; r.FP = FP; r.p = c.s->p; r.q = c.s->q;
    r.k6[0] = log(1.0 * (1 - FN) + 0.0 * FP);
    r.k6[1] = log(1.0 * FN + 0.0 * (1 - FP));
    r.k6[2] = log(0.0 * (1 - FN) + 1.0 * FP);
    r.k6[3] = log(0.0 * FN + 1.0 * (1 - FP));
    r.k6[4] = log(mix0 * (1 - FN) + (1 - mix0) * FP);
    r.k6[5] = log(mix0 * FN + (1 - mix0) * (1 - FP));
    c.stats_fresh = false;                 // `members` is reused by the move
}

static int sm_enqueue(Group* g, int ci) {
    GroupChain& c = g->ch[ci];
    const int k = c.K(), scans = g->mv.sm_steps;
    int move;
    if (k == 1) move = 0;
    else if (k == c.w->N) move = 1;
    else {
        std::vector<double> ratios = {g->mv.sm_ratios[0], g->mv.sm_ratios[1]};
        move = pick_weighted(ratios, c.random());
    }
    c.move = 1 + move;
    c.sm_others.clear();
    if (move == 0) {
        // libs/CRP.py:434-450
        double tot = 0.0;
        for (int v : c.sizes) tot += (double)v;
        std::vector<double> weight(k);
        for (int j = 0; j < k; ++j) weight[j] = (double)c.sizes[j] / tot;
        int where, n;
        for (;;) {
            where = pick_weighted(weight, c.random());
            n = c.sizes[where];
            if (n != 1) break;
        }
        const int target = c.ids[where];
        int a_i = c.randint(n);
        int a_j = c.randint(n - 1);
        if (a_j >= a_i) a_j += 1;
        rg_begin(c, n, n, target, -1, a_i, a_j, 0);
        c.sm_lq_pick = log(weight[where]) - log((double)n) - log((double)n - 1);
        for (int j = 0; j < k; ++j) if (j != where) c.sm_others.push_back((double)c.sizes[j]);
        c.sm_n = n; c.sm_na = n; c.sm_cl_i = target; c.sm_cl_j = -1; c.sm_where = where;
    } else {
        // libs/CRP.py:484-507
        std::vector<double> weight(k);
        double tot = 0.0;
        for (int j = 0; j < k; ++j) { weight[j] = 1.0 / (double)c.sizes[j]; tot += weight[j]; }
        for (int j = 0; j < k; ++j) weight[j] /= tot;
        // np.random.choice(p=weight, size=2, replace=False): draw-and-dedupe
        std::vector<double> p = weight;
        int found[2], nf = 0;
        while (nf < 2) {
            const int want = 2 - nf;
            double x[2];
            for (int i = 0; i < want; ++i) x[i] = c.random();
            if (nf) p[found[0]] = 0.0;
            int cand[2];
            for (int i = 0; i < want; ++i) cand[i] = pick_weighted(p, x[i]);
            for (int i = 0; i < want && nf < 2; ++i) {
                bool dup = false;
                for (int j = 0; j < i; ++j) dup |= cand[j] == cand[i];
                if (!dup) found[nf++] = cand[i];
            }
        }
        const int w_i = found[0], w_j = found[1];
        const int cl_i = c.ids[w_i], cl_j = c.ids[w_j];
        const int n_a = c.sizes[w_i], n_b = c.sizes[w_j];
        const int a_i = c.randint(n_a);
        const int a_j = c.randint(n_b);
        const int n = n_a + n_b;
        rg_begin(c, n, n_a, cl_i, cl_j, a_i, a_j, 1);
        double lw = 0.0, ls = 0.0;
        for (int j = 0; j < k; ++j)
            if (c.ids[j] == cl_i || c.ids[j] == cl_j) { lw += log(weight[j]); ls += log((double)c.sizes[j]); }
        c.sm_lq_pick = lw - ls;
        c.sm_n = n; c.sm_na = n_a; c.sm_cl_i = cl_i; c.sm_cl_j = cl_j; c.sm_where = w_i;
    }
    // _restricted_gibbs: launch state, `scans` intermediate scans, the final scan with
    // transition probabilities, the decision scalars
    c.rg.stream_id = c.reserve(2);
    if (int rc = bnpc_chain_rg_setup(c.w, &c.rg, nullptr)) return rc;
    for (int sc = 0; sc < scans; ++sc) {
        c.rg.stream_id = c.reserve(4);
        if (int rc = bnpc_chain_rg_scan_split(c.w, &c.rg, 0, nullptr)) return rc;
        c.rg.stream_id = c.reserve(2);
        if (int rc = bnpc_chain_rg_scan_merged(c.w, &c.rg, 0, nullptr)) return rc;
    }
    if (move == 0) {
        c.rg.stream_id = c.reserve(4);
        if (int rc = bnpc_chain_rg_scan_split(c.w, &c.rg, 1, nullptr)) return rc;
        c.rg.stream_id = c.reserve(1);
        return bnpc_chain_rg_decide_split(c.w, &c.rg, c.s->beta_prior_uniform, nullptr);
    }
    c.rg.stream_id = c.reserve(2);
    if (int rc = bnpc_chain_rg_scan_merged(c.w, &c.rg, 1, nullptr)) return rc;
    c.rg.stream_id = c.reserve(1);
    return bnpc_chain_rg_decide_merge(c.w, &c.rg, c.s->beta_prior_uniform, nullptr);
}

// decision after the synchronisation (engine.py _decide_split / _decide_merge); an accepted move
// records its write-back
static int sm_after(Group* g, int ci) {
    GroupChain& c = g->ch[ci];
    const double* sc = c.w->h_scal;
    enum { FWD_ASSIGN = 0, FWD_THETA = 1, BACK = 2, BACK_LQ = 3, PRIOR_NEW = 4, PRIOR_OLD = 6, LL3 = 8 };
    const int n = c.sm_n;
    const bool flat = c.s->beta_prior_uniform != 0;
    bool accept;
    int ones = 0;
    if (c.move == 1) {
        ones = (n > 2) ? c.w->h_out[3] : 0;
        const double logq_ratio = sc[BACK] - (sc[FWD_ASSIGN] + sc[FWD_THETA]);
        const int n_j = ones + 1, n_i = n - n_j;
        double r = log(c.s->DP_a) - lgamma((double)n);
        if (n_i > 0) r += lgamma((double)n_j);
        if (n_j > 0) r += lgamma((double)n_i);
        if (!flat) r += (sc[PRIOR_NEW] + sc[PRIOR_NEW + 1]) - sc[PRIOR_OLD];
        const double ll_ratio = sc[LL3] + sc[LL3 + 1] - sc[LL3 + 2];
        double norm = 0.0;
        for (double v : c.sm_others) norm += 1.0 / v;
        norm += 1.0 / (double)n_i;
        norm += 1.0 / (double)n_j;
        const double size_ratio = (log(1.0 / n_i / norm) + log(1.0 / n_j / norm)) - c.sm_lq_pick;
        const double total = logq_ratio + r + ll_ratio + size_ratio;
        if (n > 2 && (ones == 0 || ones == n - 2)) accept = false;
        else accept = log(c.random()) < total;
    } else {
        const int n_a = c.sm_na, nf = n - 2;
        const double logq_ratio = (sc[BACK] + sc[BACK_LQ]) - sc[FWD_THETA];
        const int n_j = (n - n_a - 1) + 1, n_i = n - n_j;
        double r = lgamma((double)n) - log(c.s->DP_a);
        if (n_i > 0) r -= lgamma((double)n_i);
        if (n_j > 0) r -= lgamma((double)n_j);
        if (!flat) r += sc[PRIOR_NEW] - (sc[PRIOR_OLD] + sc[PRIOR_OLD + 1]);
        const double ll_ratio = sc[LL3 + 2] - sc[LL3] - sc[LL3 + 1];
        const double N = (double)c.w->N;
        const double back_size = (nf - 1 > 0) ? -log(N) - log((double)nf - 1) : -log(N);
        const double size_ratio = back_size - c.sm_lq_pick;
        const double total = logq_ratio + r + ll_ratio + size_ratio;
        accept = log(c.random()) < total;
    }
    const int slot = (c.move == 1) ? 1 : 2;
    c.s->mh_counter[2 * slot + (accept ? 0 : 1)] += 1.0;
    if (!accept) return 0;
    if (c.move == 1) {
        const int new_id = get_empty_cluster(c);
        if (int rc = ensure_ids(g, ci, new_id + BNPC_MAX_EXTRA + 2)) return rc;
        if (int rc = bnpc_chain_rg_apply(c.w, &c.rg, new_id, nullptr)) return rc;
        const int moved = ones + 1;
        c.sizes[c.sm_where] -= moved;
        c.ids.push_back(new_id);
        c.sizes.push_back(moved);
    } else {
        if (int rc = bnpc_chain_rg_apply(c.w, &c.rg, -1, nullptr)) return rc;
        const int n_b = n - c.sm_na;
        int at_j = -1;
        for (int j = 0; j < c.K(); ++j) {
            if (c.ids[j] == c.sm_cl_i) c.sizes[j] += n_b;
            if (c.ids[j] == c.sm_cl_j) at_j = j;
        }
        c.ids.erase(c.ids.begin() + at_j);
        c.sizes.erase(c.sizes.begin() + at_j);
    }
    c.stats_fresh = false;
    return 0;
}

// ------------------------------------------------------------------ DP alpha (engine.py
// update_DP_alpha, libs/CRP.py:386-410; host scalars)
static void update_dp_alpha(GroupChain& c) {
    const int k = c.K();
    const double N = (double)c.w->N, a0 = c.s->dp_a0, b0 = c.s->dp_b0;
    const double eta = host_beta(c.s->seed, &c.s->host_ctr, c.s->DP_a + 1, N);
    const double w = (a0 + k - 1) / (N * (b0 - log(eta)));
    const double pi_eta = w / (1 + w);
    const double scale = b0 - log(eta);
    double draw;
    if (c.random() < pi_eta) draw = host_gamma(c.s->seed, &c.s->host_ctr, a0 + k) * scale;
    else draw = host_gamma(c.s->seed, &c.s->host_ctr, a0 + k - 1) * scale;
    c.s->DP_a = std::max(1 + 1e-15, draw);
}

// ------------------------------------------------------------------ phase 2: statistics, MH on
// theta, error-rate proposals, likelihoods (engine.py _refresh_stats / update_parameters /
// update_error_rates / _trace_scalars)
static double prior_logpdf(double x, double mean, double sd, double cst) {
    const double lo = (0 - mean) / sd, hi = (1 - mean) / sd;
    const double y = (x - mean) / sd;
    if (!(lo <= y && y <= hi)) return -INFINITY;
    return (-(y * y) / 2.0 + cst) - log(sd);
}

static void propose_error(GroupChain& c, double cur, double sd0, double* prop, double* fwd, double* rev, double* acc_u) {
    // libs/CRP_learning_errors.py:66-92
    const double steps[3] = {sd0 * 0.5, sd0, sd0 * 1.5};
    const double sd = steps[c.randint(3)];
    const double lo = (0 - cur) / sd, hi = (1 - cur) / sd;
    const double u = c.random();
    *prop = h_tn_ppf(u, lo, hi) * sd + cur;
    *fwd = h_tn_logpdf(*prop, lo, hi, cur, sd);
    const double lo_r = (0 - *prop) / sd, hi_r = (1 - *prop) / sd;
    *rev = h_tn_logpdf(cur, lo_r, hi_r, *prop, sd);
    *acc_u = c.random();
}

static int params_enqueue(Group* g, int ci, bool do_errors) {
    GroupChain& c = g->ch[ci];
    const int K = c.K();
    if (!c.stats_fresh) {
        int32_t* h = c.w->h_in;
        int run = 0, mx = 0;
        for (int j = 0; j < K; ++j) h[j] = c.ids[j];
        h[K] = 0;
        for (int j = 0; j < K; ++j) { run += c.sizes[j]; h[K + 1 + j] = run; mx = std::max(mx, c.sizes[j]); }
        if (int rc = bnpc_chain_stats(c.w, K, mx, nullptr)) return rc;
        c.stats_fresh = true;
    }
    const uint64_t sid = c.reserve(2);
    if (int rc = bnpc_chain_mh_theta(c.w, K, 0, c.s->seed, sid, c.s->FN, c.s->FP, c.s->p, c.s->q, nullptr)) return rc;
    // error-rate moves: both proposals depend only on the current rates and on host draws, so the
    // four likelihoods the two decisions (and the trace row) can need are evaluated in one pass
    c.err_move = do_errors;
    double fn[4], fp[4];
    int E = 1;
    fn[0] = c.s->FN; fp[0] = c.s->FP;
    if (do_errors) {
        propose_error(c, c.s->FP, c.s->fp_sd, &c.fp_prop, &c.fp_fwd, &c.fp_rev, &c.fp_acc_u);
        propose_error(c, c.s->FN, c.s->fn_sd, &c.fn_prop, &c.fn_fwd, &c.fn_rev, &c.fn_acc_u);
        fp[0] = c.fp_prop; fn[0] = c.s->FN;
        fp[1] = c.s->FP;   fn[1] = c.s->FN;
        fp[2] = c.fp_prop; fn[2] = c.fn_prop;
        fp[3] = c.s->FP;   fn[3] = c.fn_prop;
        E = 4;
    }
    const int want_prior = c.s->beta_prior_uniform ? 0 : 1;
    c.loglik_rows = E + want_prior;
    return bnpc_chain_loglik(c.w, K, fn, fp, E, want_prior, c.s->p, c.s->q, nullptr);
}

static void params_after(Group* g, GroupChain& c, double* ml, double* lprior, bool with_moves) {
    const int K = c.K(), M = c.w->M;
    if (with_moves) {
        const int declined = c.w->h_out[0];
        c.s->mh_counter[0] += (double)((long long)K * M - declined);
        c.s->mh_counter[1] += (double)declined;
    }
    const double* r = c.w->h_scal;
    double ll = r[0];
    if (c.err_move) {
        // FP then FN (libs/CRP_learning_errors.py:52-55, 93-111)
        const double A_fp = r[0] + prior_logpdf(c.fp_prop, c.s->fp_mean, c.s->fp_sd, c.s->fp_prior_const) - r[1] -
                            prior_logpdf(c.s->FP, c.s->fp_mean, c.s->fp_sd, c.s->fp_prior_const) + c.fp_rev - c.fp_fwd;
        const bool acc_fp = log(c.fp_acc_u) < A_fp;
        c.s->mh_counter[6 + (acc_fp ? 0 : 1)] += 1.0;
        if (acc_fp) c.s->FP = c.fp_prop;
        const double ll_new = acc_fp ? r[2] : r[3], ll_old = acc_fp ? r[0] : r[1];
        const double A_fn = ll_new + prior_logpdf(c.fn_prop, c.s->fn_mean, c.s->fn_sd, c.s->fn_prior_const) - ll_old -
                            prior_logpdf(c.s->FN, c.s->fn_mean, c.s->fn_sd, c.s->fn_prior_const) + c.fn_rev - c.fn_fwd;
        const bool acc_fn = log(c.fn_acc_u) < A_fn;
        c.s->mh_counter[8 + (acc_fn ? 0 : 1)] += 1.0;
        if (acc_fn) c.s->FN = c.fn_prop;
        ll = acc_fn ? ll_new : ll_old;
    }
    *ml = ll;
    // engine.py get_lprior_full (libs/CRP.py:241-251, libs/CRP_learning_errors.py:47-49)
    const double a0 = c.s->dp_a0, loc = c.s->dp_b0, y = c.s->DP_a - loc;
    double lp = (y > 0) ? (((a0 - 1.0 == 0.0) ? 0.0 : (a0 - 1.0) * log(y)) - y - lgamma(a0)) : -INFINITY;
    const double cn = log((double)c.w->N - 1 + c.s->DP_a);
    double crp = 0.0;
    for (int v : c.sizes) crp += log((double)v) - cn;
    lp += crp;
    if (!c.s->beta_prior_uniform) lp += r[c.loglik_rows - 1];
    if (c.s->learning)
        lp += prior_logpdf(c.s->FP, c.s->fp_mean, c.s->fp_sd, c.s->fp_prior_const) +
              prior_logpdf(c.s->FN, c.s->fn_mean, c.s->fn_sd, c.s->fn_prior_const);
    *lprior = lp;
    (void)g;
}

// the trace row of the current state (Chain.update_results without a move)
static int trace_only_enqueue(Group* g, int ci) {
    GroupChain& c = g->ch[ci];
    const int K = c.K();
    if (!c.stats_fresh) {
        int32_t* h = c.w->h_in;
        int run = 0, mx = 0;
        for (int j = 0; j < K; ++j) h[j] = c.ids[j];
        h[K] = 0;
        for (int j = 0; j < K; ++j) { run += c.sizes[j]; h[K + 1 + j] = run; mx = std::max(mx, c.sizes[j]); }
        if (int rc = bnpc_chain_stats(c.w, K, mx, nullptr)) return rc;
        c.stats_fresh = true;
    }
    c.err_move = false;
    const double fn[1] = {c.s->FN}, fp[1] = {c.s->FP};
    const int want_prior = c.s->beta_prior_uniform ? 0 : 1;
    c.loglik_rows = 1 + want_prior;
    (void)g;
    return bnpc_chain_loglik(c.w, K, fn, fp, 1, want_prior, c.s->p, c.s->q, nullptr);
}

// ------------------------------------------------------------------ trace rows
static int sync_stream(cudaStream_t s) {
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return fail("cudaStreamSynchronize", e);
    return 0;
}
#define CU(call, what)                                   \
    do {                                                 \
        cudaError_t e__ = (call);                        \
        if (e__ != cudaSuccess) return fail(what, e__);  \
    } while (0)

// snapshot of the assignment vector and of the theta rows of the sorted live ids into ring slot
// `slot` (recorded: part of the phase-2 flush)
static int snapshot_enqueue(Group* g, int ci, int slot, int step) {
    GroupChain& c = g->ch[ci];
    const bnpc_trace_t* tr = g->tr[ci];
    const int N = c.w->N, M = c.w->M, K = c.K();
    int32_t* dst = g->ring_assign + ((size_t)slot * g->n + ci) * N;
    if (int rc = copy_async(dst, c.w->assign, sizeof(int32_t) * (size_t)N, cudaMemcpyDeviceToDevice, nullptr)) return rc;
    c.theta_direct = false;
    if (tr->params_h && step >= tr->params_first) {
        if (K > g->ring_kcap) {
            // more live clusters than a ring row holds: the rows are copied straight from theta once
            // the phase is done (theta_rows_direct); the ring is not regrown under recorded launches
            c.theta_direct = true;
            return 0;
        }
        // sorted ids (libs/MCMC.py:261) behind the statistics' staging area of h_in
        int32_t* h = c.w->h_in + 2 * K + 8;
        std::vector<int> sorted(c.ids);
        std::sort(sorted.begin(), sorted.end());
        for (int j = 0; j < K; ++j) h[j] = sorted[j];
        if (int rc = copy_async(c.w->cursor, h, sizeof(int32_t) * (size_t)K, cudaMemcpyHostToDevice, nullptr)) return rc;
        float* rows = g->ring_theta + ((size_t)slot * g->n + ci) * (size_t)g->ring_kcap * M;
        BNPC_LAUNCH(gather_rows_kernel, 0, 0, cdiv((long long)K * M, 256), 256, 0, nullptr, c.w->theta, c.w->cursor, K, M, rows);
    }
    return 0;
}

// theta rows of the sorted live ids straight from the chain's theta (synchronous; only for chains
// with more live clusters than a ring row holds)
static int theta_rows_direct(Group* g, int ci, int step) {
    GroupChain& c = g->ch[ci];
    bnpc_trace_t* tr = g->tr[ci];
    const int M = c.w->M, K = c.K();
    if (K > tr->params_kcap) {
        CU(cudaStreamSynchronize(g->sC), "trace sync");
        if (g->grow(g->grow_ctx, ci, BNPC_GROW_PARAMS, K) || K > tr->params_kcap) return group_fail(g, "grow(params) failed");
    }
    std::vector<int> sorted(c.ids);
    std::sort(sorted.begin(), sorted.end());
    float* dst = tr->params_h + (size_t)(step - tr->params_first) * tr->params_kcap * M;
    for (int j = 0; j < K; ++j)
        CU(cudaMemcpyAsync(dst + (size_t)j * M, c.w->theta + (size_t)sorted[j] * M, sizeof(float) * M,
                           cudaMemcpyDeviceToHost, g->sC), "trace copy (theta rows)");
    CU(cudaStreamSynchronize(g->sC), "trace sync");
    return 0;
}

// one chain's ring slot -> its pinned trace rows, on the copy stream
static int drain_chain(Group* g, int ci, int slot, int step) {
    GroupChain& c = g->ch[ci];
    bnpc_trace_t* tr = g->tr[ci];
    const int N = c.w->N, M = c.w->M, K = c.K();
    if (tr->assign_h) {
        const int32_t* src = g->ring_assign + ((size_t)slot * g->n + ci) * N;
        CU(cudaMemcpyAsync(tr->assign_h + (size_t)step * tr->assign_stride, src, sizeof(int32_t) * (size_t)N,
                           cudaMemcpyDeviceToHost, g->sC), "trace copy (assignment)");
    }
    if (tr->params_h && step >= tr->params_first) {
        if (c.theta_direct) return theta_rows_direct(g, ci, step);
        if (K > tr->params_kcap) {
            CU(cudaStreamSynchronize(g->sC), "trace sync");
            if (g->grow(g->grow_ctx, ci, BNPC_GROW_PARAMS, K) || K > tr->params_kcap)
                return group_fail(g, "grow(params) failed");
        }
        const float* src = g->ring_theta + ((size_t)slot * g->n + ci) * (size_t)g->ring_kcap * M;
        float* dst = tr->params_h + (size_t)(step - tr->params_first) * tr->params_kcap * M;
        CU(cudaMemcpyAsync(dst, src, sizeof(float) * (size_t)K * M, cudaMemcpyDeviceToHost, g->sC),
           "trace copy (theta)");
    }
    return 0;
}

// ------------------------------------------------------------------ asynchronous scheduler
// Chains do not wait for each other: whenever the GPU work a chain waits for is done, the host
// advances THAT chain (decisions of the finished phase, recording of the next one).  The chains that
// became ready in the same polling round and enter the same kind of phase form a WAVE: their
// launches are merged (bnpc_batch.cuh) and go out on one stream of a pool, followed by one event the
// members wait for.  Waves of different kinds (Gibbs epochs, split-merge moves, parameter phases) and
// of different rounds run side by side on the GPU.
enum { CS_START = 0, CS_WAIT_SM, CS_WAIT_GIBBS, CS_WAIT_P2, CS_FINISHED };
enum { KIND_GIBBS = 0, KIND_SM = 1, KIND_P2 = 2, KIND_NONE = 3 };

static int enqueue_phase2(Group* g, int ci) {
    GroupChain& c = g->ch[ci];
    if (!g->mv.fix_assign && c.random() < g->mv.dpa_prob) update_dp_alpha(c);
    const bool do_err = c.s->learning && c.random() < g->mv.error_prob;
    if (int rc = params_enqueue(g, ci, do_err)) return rc;
    const int slot = (int)(c.snaps % g->ring);
    c.prewait = g->ev_slot[(size_t)ci * g->ring + slot];       // the slot must have been drained
    return snapshot_enqueue(g, ci, slot, c.step);
}

// the chain's pending GPU work is done: decide, record what comes next; *kind = phase recorded
static int advance(Group* g, int ci, int end_step, int* kind) {
    GroupChain& c = g->ch[ci];
    *kind = KIND_NONE;
    g_rec.cur = ci;
    for (;;) {
        switch (c.state) {
            case CS_START:
                if (c.step >= end_step) { c.state = CS_FINISHED; return 0; }
                if (g->mv.fix_assign) {
                    if (int rc = enqueue_phase2(g, ci)) return rc;
                    c.state = CS_WAIT_P2; *kind = KIND_P2;
                    return 0;
                }
                if (c.random() < g->mv.sm_prob) {
                    if (int rc = sm_enqueue(g, ci)) return rc;
                    c.state = CS_WAIT_SM; *kind = KIND_SM;
                } else {
                    gibbs_begin(g, c);
                    if (int rc = gibbs_enqueue(g, ci)) return rc;
                    c.state = CS_WAIT_GIBBS; *kind = KIND_GIBBS;
                }
                return 0;
            case CS_WAIT_SM:
                if (int rc = sm_after(g, ci)) return rc;            // an accepted move records its write-back
                if (int rc = enqueue_phase2(g, ci)) return rc;
                c.state = CS_WAIT_P2; *kind = KIND_P2;
                return 0;
            case CS_WAIT_GIBBS: {
                int done = 0;
                if (int rc = gibbs_after(g, ci, &done)) return rc;
                if (!done) {
                    if (int rc = gibbs_enqueue(g, ci)) return rc;
                    *kind = KIND_GIBBS;
                    return 0;
                }
                if (int rc = enqueue_phase2(g, ci)) return rc;
                c.state = CS_WAIT_P2; *kind = KIND_P2;
                return 0;
            }
            case CS_WAIT_P2: {
                bnpc_trace_t* tr = g->tr[ci];
                double ml, lprior;
                params_after(g, c, &ml, &lprior, true);
                tr->ml[c.step] = ml;
                tr->map[c.step] = ml + lprior;
                tr->alpha[c.step] = c.s->DP_a;
                tr->fn[c.step] = c.s->FN;
                tr->fp[c.step] = c.s->FP;
                if (tr->n_clusters) tr->n_clusters[c.step] = c.K();
                const int slot = (int)(c.snaps % g->ring);
                // (the wave's event has completed: the copy stream may read the slot at once)
                if (int rc = drain_chain(g, ci, slot, c.step)) return rc;
                CU(cudaEventRecord(g->ev_slot[(size_t)ci * g->ring + slot], g->sC), "record(drained)");
                c.snaps += 1;
                c.step += 1;
                c.state = CS_START;
                break;                                               // straight on to the next step
            }
            default:
                return 0;
        }
    }
}

static int group_run_async(Group* g, int step0, int n_steps) {
    const int end_step = step0 + n_steps;
    const int n = g->n;
    for (int ci = 0; ci < n; ++ci) {
        GroupChain& c = g->ch[ci];
        c.state = CS_START; c.step = step0; c.sub = -1; c.prewait = nullptr; c.theta_direct = false;
    }
    int finished = 0;
    unsigned long long ready = (n >= 64) ? ~0ull : ((1ull << n) - 1ull);
    long long idle_spins = 0;
    while (finished < n) {
        // waves whose event has completed release their members
        for (Group::Wave& wv : g->waves) {
            if (!wv.busy) continue;
            const cudaError_t q = cudaEventQuery(wv.ev);
            if (q == cudaErrorNotReady) continue;
            if (q != cudaSuccess) return fail("cudaEventQuery", q);
            ready |= wv.mask;
            wv.busy = false;
        }
        if (!ready) {
            if (++idle_spins > (1ll << 34)) return group_fail(g, "scheduler made no progress");
            continue;
        }
        idle_spins = 0;
        unsigned long long by_kind[3] = {0ull, 0ull, 0ull};
        g_rec.on = true;
        int rc = 0;
        for (int ci = 0; ci < n && !rc; ++ci) {
            if (!((ready >> ci) & 1ull)) continue;
            int kind = KIND_NONE;
            rc = advance(g, ci, end_step, &kind);
            if (kind != KIND_NONE) by_kind[kind] |= 1ull << ci;
            else if (!rc && g->ch[ci].state == CS_FINISHED) ++finished;
        }
        g_rec.on = false;
        ready = 0ull;
        if (rc) { for (int ci = 0; ci < GROUP_MAX; ++ci) g_rec.q[ci].clear(); return rc; }
        for (int k = 0; k < 3; ++k) {
            if (!by_kind[k]) continue;
            Group::Wave* wv = nullptr;
            for (Group::Wave& cand : g->waves) if (!cand.busy) { wv = &cand; break; }
            if (!wv) return group_fail(g, "no free wave");
            for (int ci = 0; ci < n; ++ci) {
                GroupChain& c = g->ch[ci];
                if (((by_kind[k] >> ci) & 1ull) && c.prewait) {
                    CU(cudaStreamWaitEvent(wv->s, c.prewait, 0), "wait(drained)");
                    c.prewait = nullptr;
                }
            }
            if ((rc = recorder_flush(wv->s, by_kind[k]))) return rc;
            CU(cudaEventRecord(wv->ev, wv->s), "record(wave)");
            wv->mask = by_kind[k];
            wv->busy = true;
        }
    }
    g->steps_done += n_steps;
    return 0;
}

// phase 2 + trace row for all chains; with_moves = false records the current state only
static int phase2_and_trace(Group* g, int step, bool with_moves) {
    const int slot = (int)(g->snaps % g->ring);
    // the ring slot must have been drained
    CU(cudaStreamWaitEvent(g->sA, g->ev_drained[slot], 0), "wait(drained)");
    g_rec.on = true;
    int rc = 0;
    for (int ci = 0; ci < g->n && !rc; ++ci) {
        GroupChain& c = g->ch[ci];
        g_rec.cur = ci;
        if (with_moves) {
            const bool do_err = c.s->learning && c.random() < g->mv.error_prob;
            rc = params_enqueue(g, ci, do_err);
        } else {
            rc = trace_only_enqueue(g, ci);
        }
        if (!rc) rc = snapshot_enqueue(g, ci, slot, step);
    }
    g_rec.on = false;
    if (rc) { for (int ci = 0; ci < GROUP_MAX; ++ci) g_rec.q[ci].clear(); return rc; }
    if ((rc = recorder_flush(g->sA))) return rc;
    CU(cudaEventRecord(g->ev_snap[slot], g->sA), "record(snapshot)");
    if ((rc = sync_stream(g->sA))) return rc;
    for (int ci = 0; ci < g->n; ++ci) {
        GroupChain& c = g->ch[ci];
        bnpc_trace_t* tr = g->tr[ci];
        double ml, lprior;
        params_after(g, c, &ml, &lprior, with_moves);
        tr->ml[step] = ml;
        tr->map[step] = ml + lprior;
        tr->alpha[step] = c.s->DP_a;
        tr->fn[step] = c.s->FN;
        tr->fp[step] = c.s->FP;
        if (tr->n_clusters) tr->n_clusters[step] = c.K();
    }
    CU(cudaStreamWaitEvent(g->sC, g->ev_snap[slot], 0), "wait(snapshot)");
    for (int ci = 0; ci < g->n; ++ci)
        if ((rc = drain_chain(g, ci, slot, step))) return rc;
    CU(cudaEventRecord(g->ev_drained[slot], g->sC), "record(drained)");
    g->snaps += 1;
    return 0;
}

// one MCMC step of all chains (libs/MCMC.py:320-342 + :242-282)
static int group_step(Group* g, int step) {
    int rc = 0;
    unsigned long long maskG = 0, maskS = 0;
    if (!g->mv.fix_assign) {
        for (int ci = 0; ci < g->n; ++ci) {
            GroupChain& c = g->ch[ci];
            c.move = (c.random() < g->mv.sm_prob) ? 1 : 0;
            if (c.move) maskS |= 1ull << ci; else maskG |= 1ull << ci;
        }
        // split-merge chains wait for the parameter phase of the previous step (stream A)
        if (maskS) CU(cudaStreamWaitEvent(g->sB, g->evA, 0), "wait(phase 2)");
        g_rec.on = true;
        for (int ci = 0; ci < g->n && !rc; ++ci) {
            GroupChain& c = g->ch[ci];
            g_rec.cur = ci;
            if (c.move) rc = sm_enqueue(g, ci);
            else { gibbs_begin(g, c); rc = gibbs_enqueue(g, ci); }
        }
        g_rec.on = false;
        if (rc) { for (int ci = 0; ci < GROUP_MAX; ++ci) g_rec.q[ci].clear(); return rc; }
        if (maskS && (rc = recorder_flush(g->sB, maskS))) return rc;
        if (maskG && (rc = recorder_flush(g->sA, maskG))) return rc;
        // split-merge decisions
        if (maskS) {
            if ((rc = sync_stream(g->sB))) return rc;
            g_rec.on = true;
            for (int ci = 0; ci < g->n && !rc; ++ci)
                if ((maskS >> ci) & 1ull) { g_rec.cur = ci; rc = sm_after(g, ci); }
            g_rec.on = false;
            if (rc) { for (int ci = 0; ci < GROUP_MAX; ++ci) g_rec.q[ci].clear(); return rc; }
            if ((rc = recorder_flush(g->sB, maskS))) return rc;
            CU(cudaEventRecord(g->evB, g->sB), "record(split-merge)");
            CU(cudaStreamWaitEvent(g->sA, g->evB, 0), "wait(split-merge)");
        }
        // Gibbs epochs until every sweep is complete
        unsigned long long open = maskG;
        while (open) {
            if ((rc = sync_stream(g->sA))) return rc;
            g_rec.on = true;
            for (int ci = 0; ci < g->n && !rc; ++ci) {
                if (!((open >> ci) & 1ull)) continue;
                int done = 0;
                g_rec.cur = ci;
                rc = gibbs_after(g, ci, &done);
                if (!rc) {
                    if (done) open &= ~(1ull << ci);
                    else rc = gibbs_enqueue(g, ci);
                }
            }
            g_rec.on = false;
            if (rc) { for (int ci = 0; ci < GROUP_MAX; ++ci) g_rec.q[ci].clear(); return rc; }
            if (open && (rc = recorder_flush(g->sA, open))) return rc;
        }
        for (int ci = 0; ci < g->n; ++ci) {
            GroupChain& c = g->ch[ci];
            if (c.random() < g->mv.dpa_prob) update_dp_alpha(c);
        }
    }
    if ((rc = phase2_and_trace(g, step, true))) return rc;
    CU(cudaEventRecord(g->evA, g->sA), "record(phase 2)");
    g->steps_done += 1;
    return 0;
}

}  // namespace bnpc

extern "C" {

double bnpc_host_random(uint64_t seed, uint64_t* ctr) { return bnpc::host_random(seed, ctr); }
double bnpc_host_gamma(uint64_t seed, uint64_t* ctr, double shape) { return bnpc::host_gamma(seed, ctr, shape); }
double bnpc_host_beta(uint64_t seed, uint64_t* ctr, double a, double b) { return bnpc::host_beta(seed, ctr, a, b); }
double bnpc_host_truncnorm_ppf(double q, double lo, double hi) { return bnpc::h_tn_ppf(q, lo, hi); }
double bnpc_host_truncnorm_logpdf(double x, double lo, double hi, double loc, double scale) {
    return bnpc::h_tn_logpdf(x, lo, hi, loc, scale);
}

int bnpc_batch_begin(void) {
    for (int c = 0; c < bnpc::GROUP_MAX; ++c) bnpc::g_rec.q[c].clear();
    bnpc::g_rec.on = true;
    bnpc::g_rec.cur = 0;
    return 0;
}
int bnpc_batch_slot(int slot) {
    if (slot < 0 || slot >= bnpc::GROUP_MAX) return bad_arg("slot");
    bnpc::g_rec.cur = slot;
    return 0;
}
int bnpc_batch_flush(void* stream) {
    bnpc::g_rec.on = false;
    return recorder_flush((cudaStream_t)stream);
}

int bnpc_prof_enable(int on) {
    bnpc::g_prof.on = on != 0;
    return 0;
}
int bnpc_prof_report(char* buf, int cap) {
    using namespace bnpc;
    struct Acc { const char* name; long long launches, chains; double ms; };
    std::vector<Acc> acc;
    for (ProfSlot& sl : g_prof.slots) {
        float ms = 0.f;
        cudaEventSynchronize(sl.b);
        cudaEventElapsedTime(&ms, sl.a, sl.b);
        Acc* a = nullptr;
        for (Acc& x : acc) if (x.name == sl.name || strcmp(x.name, sl.name) == 0) { a = &x; break; }
        if (!a) { acc.push_back(Acc{sl.name, 0, 0, 0.0}); a = &acc.back(); }
        a->launches += 1; a->chains += sl.chains; a->ms += ms;
        g_prof.pool.push_back(sl.a);
        g_prof.pool.push_back(sl.b);
    }
    g_prof.slots.clear();
    int off = 0;
    for (Acc& x : acc) {
        const int n = snprintf(buf + off, cap > off ? cap - off : 0, "%s %lld %lld %.6f\n", x.name, x.launches,
                               x.chains, x.ms);
        if (n < 0 || off + n >= cap) break;
        off += n;
    }
    if (cap > 0) buf[off < cap ? off : cap - 1] = 0;
    return 0;
}

bnpc_group_t* bnpc_group_create(int n, bnpc_chain_t* const* ws, bnpc_chain_state_t* const* st,
                                bnpc_trace_t* const* tr, const bnpc_moves_t* moves, bnpc_grow_fn grow,
                                void* grow_ctx, void* stream_main, void* stream_side, void* stream_copy) {
    using namespace bnpc;
    if (n <= 0 || n > GROUP_MAX || !ws || !st || !tr || !moves || !grow) { bad_arg("group arguments"); return nullptr; }
    Group* g = new Group();
    g->n = n;
    g->ch.resize(n);
    g->tr.assign(tr, tr + n);
    for (int i = 0; i < n; ++i) { g->ch[i].w = ws[i]; g->ch[i].s = st[i]; }
    g->mv = *moves;
    g->grow = grow; g->grow_ctx = grow_ctx;
    g->sA = (cudaStream_t)stream_main; g->sB = (cudaStream_t)stream_side; g->sC = (cudaStream_t)stream_copy;
    g->ring = 0; g->ring_assign = nullptr; g->ring_theta = nullptr; g->ring_kcap = 0;
    g->snaps = 0; g->steps_done = 0;
    if (cudaEventCreateWithFlags(&g->evA, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&g->evB, cudaEventDisableTiming) != cudaSuccess) {
        delete g;
        bad_arg("cudaEventCreate");
        return nullptr;
    }
    cudaEventRecord(g->evA, g->sA);
    g->lockstep = getenv("BNPC_LOCKSTEP") != nullptr;        // the phase-synchronous driver, for comparison
    // waves: every chain is in at most one, and a round opens at most three
    g->waves.resize((size_t)n + 3);
    for (Group::Wave& wv : g->waves) {
        wv.mask = 0ull; wv.busy = false;
        if (cudaStreamCreateWithFlags(&wv.s, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&wv.ev, cudaEventDisableTiming) != cudaSuccess) {
            bad_arg("wave stream / event");
            return nullptr;
        }
    }
    return reinterpret_cast<bnpc_group_t*>(g);
}

int bnpc_group_set_ring(bnpc_group_t* gp, int slots, int32_t* ring_assign, float* ring_theta, int kcap) {
    using namespace bnpc;
    Group* g = reinterpret_cast<Group*>(gp);
    if (!g || slots <= 0 || !ring_assign) return bad_arg("ring");
    // pending drains read the old ring
    if (cudaStreamSynchronize(g->sC) != cudaSuccess || cudaStreamSynchronize(g->sA) != cudaSuccess)
        return bad_arg("ring sync");
    while ((int)g->ev_snap.size() < slots) {
        cudaEvent_t a, b;
        if (cudaEventCreateWithFlags(&a, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&b, cudaEventDisableTiming) != cudaSuccess) return bad_arg("cudaEventCreate");
        cudaEventRecord(b, g->sC);
        g->ev_snap.push_back(a);
        g->ev_drained.push_back(b);
    }
    g->ring = slots; g->ring_assign = ring_assign; g->ring_theta = ring_theta; g->ring_kcap = kcap;
    while ((int)g->ev_slot.size() < g->n * slots) {
        cudaEvent_t e;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return bad_arg("cudaEventCreate");
        cudaEventRecord(e, g->sC);
        g->ev_slot.push_back(e);
    }
    return 0;
}

static int group_enter(bnpc::Group* g) {
    using namespace bnpc;
    if (!g || g->ring <= 0) return bad_arg("group without a trace ring");
    for (int ci = 0; ci < g->n; ++ci) {
        GroupChain& c = g->ch[ci];
        load_list(c);
        c.stats_fresh = c.s->stats_fresh != 0;
        c.move = -1;
    }
    for (int c = 0; c < GROUP_MAX; ++c) g_rec.q[c].clear();
    return 0;
}
static int group_leave(bnpc::Group* g, int rc) {
    using namespace bnpc;
    g_rec.on = false;
    // the traces have reached the host when this returns
    cudaError_t e = cudaStreamSynchronize(g->sC);
    if (!rc && e != cudaSuccess) rc = fail("trace drain", e);
    cudaStreamSynchronize(g->sB);
    cudaStreamSynchronize(g->sA);
    for (Group::Wave& wv : g->waves) { cudaStreamSynchronize(wv.s); wv.busy = false; }
    for (int ci = 0; ci < g->n; ++ci) {
        GroupChain& c = g->ch[ci];
        if (int r2 = store_list(g, ci)) { if (!rc) rc = r2; }
        c.s->stats_fresh = c.stats_fresh ? 1 : 0;
    }
    return rc;
}

int bnpc_group_run(bnpc_group_t* gp, int step0, int n_steps) {
    using namespace bnpc;
    Group* g = reinterpret_cast<Group*>(gp);
    if (int rc = group_enter(g)) return rc;
    int rc = 0;
    if (g->lockstep) {
        for (int s = 0; s < n_steps && !rc; ++s) rc = group_step(g, step0 + s);
    } else {
        rc = group_run_async(g, step0, n_steps);
    }
    return group_leave(g, rc);
}

int bnpc_group_record(bnpc_group_t* gp, int step) {
    using namespace bnpc;
    Group* g = reinterpret_cast<Group*>(gp);
    if (int rc = group_enter(g)) return rc;
    return group_leave(g, phase2_and_trace(g, step, false));
}

void bnpc_group_destroy(bnpc_group_t* gp) {
    using namespace bnpc;
    Group* g = reinterpret_cast<Group*>(gp);
    if (!g) return;
    cudaStreamSynchronize(g->sC);
    cudaEventDestroy(g->evA);
    cudaEventDestroy(g->evB);
    for (cudaEvent_t e : g->ev_snap) cudaEventDestroy(e);
    for (cudaEvent_t e : g->ev_drained) cudaEventDestroy(e);
    for (cudaEvent_t e : g->ev_slot) cudaEventDestroy(e);
    for (Group::Wave& wv : g->waves) { cudaStreamSynchronize(wv.s); cudaEventDestroy(wv.ev); cudaStreamDestroy(wv.s); }
    delete g;
}

}  // extern "C"
