/*
 * bnpc_b200.h -- C ABI of the B200-native BnpC MCMC hot path.
 *
 * The reference (cbg-ethz/BnpC v0.2.1) is pure Python: it has no FFI of its own.
 * Its boundary for this path is the duck-typed class contract consumed by
 * libs/MCMC.py:320-342 (Chain.do_step) and libs/MCMC.py:242-282
 * (Chain.update_results).  The entry points below are the device-side pieces
 * of that contract; each one cites the reference function(s) whose arithmetic
 * it replaces.  Host orchestration (which entry point is called when) mirrors
 * the reference methods one-to-one and lives in bnpc_b200/engine.py behind
 * libs/CRP.py / libs/CRP_learning_errors.py.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure;
 *     bnpc_last_error() then returns a static, thread-local message;
 *   - all pointers are DEVICE pointers owned by the caller (allocated by
 *     PyTorch on the current device) unless the name ends in _h; the library
 *     allocates nothing and keeps no state between calls;
 *   - `stream` is a cudaStream_t passed as void*; every launch is asynchronous
 *     on that stream; no entry point synchronises;
 *   - data layout: two bit-planes per matrix, cell-major.  x1[n*W + w] bit b is
 *     set iff data[n][32*w+b] == 1, x0 likewise for == 0; a missing entry has
 *     neither bit; W is a multiple of 4 (rows are 16-byte aligned) and bits
 *     beyond M are zero;
 *   - cluster parameters theta are float32 [idcap][M], indexed by the
 *     reference's cluster id (smallest unused non-negative integer,
 *     libs/CRP.py:297-299);
 *   - "list order" is the insertion order of the reference's
 *     cells_per_cluster dict (libs/CRP.py:268).
 */
#ifndef BNPC_B200_H
#define BNPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BNPC_ABI_VERSION 12

/* capacity of clusters born since the ll matrix of the current epoch was built */
#define BNPC_MAX_EXTRA 32

/* ints in the sweep status block `st` */
#define BNPC_ST_K        0   /* live clusters (length of the list)            */
#define BNPC_ST_TDONE    1   /* sweep position reached (cells [0,t) processed) */
#define BNPC_ST_FLAGS    2   /* BNPC_STOP_* bits                               */
#define BNPC_ST_NEXTRA   3   /* clusters born in this epoch                    */
#define BNPC_ST_BIRTHS   4   /* clusters born in this sweep (tape rows used)   */
#define BNPC_ST_MOVED    5   /* cells whose cluster changed in this sweep      */
#define BNPC_ST_SLOW     6   /* cells that took the exact (slow) path          */
#define BNPC_ST_CYCLES   8   /* SM clock cycles of the last sweep launch, / 1024 */
#define BNPC_ST_NANOS    9   /* globaltimer ns of the last sweep launch, / 1024  */
#define BNPC_ST_NUNC     10  /* visits of the epoch that are not statically certain (bnpc_gibbs_compact) */
#define BNPC_ST_NMANY    11  /* lean epochs: uncertain visits with more than BNPC_MAX_CAND rivals         */
#define BNPC_ST_WORDS    16

#define BNPC_STOP_EXTRA_FULL  1  /* BNPC_MAX_EXTRA births: start a new epoch      */
#define BNPC_STOP_REPACK      2  /* list shrank below a warp but ll rows are wide */
#define BNPC_STOP_TAPE_EMPTY  4  /* parity mode: more births than taped rows      */
#define BNPC_STOP_CAPACITY    8  /* list/id capacity reached: grow and relaunch   */
#define BNPC_STOP_MANY       16  /* lean epoch: too many visits with more than BNPC_MAX_CAND rivals
                                    (st[BNPC_ST_NMANY]); nothing was done, run the epoch dense */

/* per-cell record of a sweep, in visiting order (64 bytes) */
typedef struct {
    double  u;      /* uniform for the categorical draw (libs/CRP.py:277)             */
    double  lnew;   /* new-cluster log posterior of this cell (libs/CRP.py:230-234)   */
    double  e_new;  /* weight of the new-cluster option: alpha * exp(ll_new - ref)    */
    double  ref;    /* reference level of the option weights (largest ll among them)  */
    int32_t cell;   /* cell index = permutation[t] (libs/CRP.py:260)                  */
    int32_t old;    /* its cluster id before the sweep                                 */
    int32_t c_old;  /* ll column of that cluster in the current epoch                  */
    int32_t n_opt;  /* options (own cluster + rivals); BNPC_MAX_OPT+1 = too many/unknown */
    int32_t i_old;  /* index of the own cluster among the options                      */
    int32_t t;      /* position of the visit in the sweep                              */
    int32_t flags;  /* BNPC_VISIT_CERTAIN                                              */
    float   e_max;  /* largest option weight e[i], rounded up                          */
} bnpc_visit_t;
/* The cell stays where it is whatever the cluster sizes are (as long as its cluster keeps a
 * second member and no cluster was born since the options were listed): its own cluster is its
 * only option, the new-cluster option is more than 40 nats below it, and u is farther than
 * 3e-10 from 0 and 1.  The sequential part of the sweep may skip such visits.               */
#define BNPC_VISIT_CERTAIN 1

/* options of one visited cell (96 bytes): its own cluster and the rivals that can come
 * within 40 nats of it for ANY cluster sizes, in ll-column order (= list order).  The
 * draw of libs/CRP.py:274-277 restricted to these options is linear in the cluster
 * sizes: P(option i) = n_i * e[i] / (sum_j n_j * e[j] + e_new), n_own counted without
 * the cell itself.                                                                    */
#define BNPC_MAX_CAND 8
#define BNPC_MAX_OPT  (BNPC_MAX_CAND + 1)
typedef struct {
    double   e[BNPC_MAX_OPT];     /* exp(ll - ref) of the option                        */
    uint16_t col[BNPC_MAX_OPT];   /* its ll column in the current epoch                 */
    uint16_t pad[3];
} bnpc_cand_t;

/* Lean epochs (lists of at most BNPC_LEAN_MAXK clusters): options of a visit as selected from
 * the APPROXIMATE log-likelihood row (16 bytes): a superset of the clusters that can come within
 * 40 + log N nats of the cell's own cluster.                                                    */
#define BNPC_LEAN_MAXK 64
#define BNPC_OPT_MANY  2     /* more than BNPC_MAX_CAND rivals (or unknown own column)            */
typedef struct {
    uint8_t col[BNPC_MAX_OPT];  /* ll columns in list order                                       */
    uint8_t n_opt;              /* own cluster + rivals; BNPC_MAX_OPT+1: too many                  */
    uint8_t i_old;              /* index of the own cluster among them                             */
    uint8_t flags;              /* BNPC_VISIT_CERTAIN | BNPC_OPT_MANY                              */
    int32_t pad;
} bnpc_opt_t;

int         bnpc_abi_version(void);
const char* bnpc_last_error(void);
/* kernels launched through this library since it was loaded (all threads) */
int64_t     bnpc_launch_count(void);

/* ---- input path: libs/dpmmIO.py:27-98 produces float64 {0,1,NaN}; this packs it.
 * x_f64 [N][M] row-major (NaN or any value other than 0/1 = missing), or
 * x_i8 [N][M] (0, 1, anything else = missing).  Exactly one of the two is non-NULL.
 * Outputs: x1,x0 [N][W]; n1,n0 [N] = per-cell counts of ones / zeros.            */
int bnpc_pack_planes(const double* x_f64, const int8_t* x_i8, int N, int M, int W,
                     uint32_t* x1, uint32_t* x0, int32_t* n1, int32_t* n0, void* stream);

/* ---- counter-based random numbers (production mode; parity mode injects a tape).
 * out[i] = U[0,1) from Philox4x32-10(key=seed, counter=(i, stream_id)); if
 * n_levels > 0 the value is floor(u * n_levels) instead (numpy randint stand-in,
 * libs/CRP.py:328).                                                              */
int bnpc_fill_uniform(double* out, int64_t n, uint64_t seed, uint64_t stream_id,
                      int n_levels, void* stream);
/* out = a pseudo-random permutation of 0..n-1 (keyed Feistel network with cycle
 * walking); stands in for np.random.permutation (libs/CRP.py:260,616).           */
int bnpc_fill_permutation(int32_t* out, int n, uint64_t seed, uint64_t stream_id,
                          void* stream);

/* ---- likelihood, libs/CRP.py:197-212 (_calc_ll, _Bernoulli_FN, _Bernoulli_FP).
 * lp[r][m] = { log(th*(1-FN) + (1-th)*FP), log(th*FN + (1-th)*(1-FP)) } for
 * th = theta[ids[r]][m] (ids==NULL: row r), with (1-th) rounded in float32 as
 * the reference does.                                                            */
int bnpc_logprob_tables(const float* theta, const int32_t* ids, int R, int M,
                        double FN, double FP, double* lp /* [R][M][2] */, void* stream);
/* ll[r][k] = sum_m x1[c][m]*lp[k][m][0] + x0[c][m]*lp[k][m][1], c = cells[r]
 * (cells==NULL: c = r; cells may point into bnpc_visit_t records: pass
 * cell_stride = 8 ints, else 1).  FP64 accumulation in mutation order.          */
int bnpc_ll_matrix(const uint32_t* x1, const uint32_t* x0, int W, int M,
                   const int32_t* cells, int cell_stride, int C,
                   const double* lp, int K, double* ll, int ldk, void* stream);

/* ---- Gibbs sweep, libs/CRP.py:254-299 (+ :88-100, :183-188, :223-234).
 * bnpc_gibbs_prepare builds the visit records for a whole sweep:
 *   visit[t] = {u[t], n1[c]*c1 + n0[c]*c0 + lnew_prior, c = perm[t], assign[c]}.  */
int bnpc_gibbs_prepare(const int32_t* perm, const double* u, const int32_t* assign,
                       const int32_t* n1, const int32_t* n0, int N,
                       double c1, double c0, double lnew_prior,
                       bnpc_visit_t* visit, void* stream);
/* After bnpc_ll_matrix of an epoch with at most 1024 columns: fill ref, e_new, c_old, n_opt,
 * i_old of the visit records [t0, t0+C) and their option records.  A cluster k can rival the
 * current cluster o of a cell (come within 40 nats of it once the CRP weights log n_k are
 * added) only if ll_k > ll_o - 40 - slack with slack = log N, whatever the sizes are; the
 * sequential sweep then only looks at these options.  c_norm = log(N-1+alpha).         */
int bnpc_gibbs_candidates(const double* ll, int ldk, int K, const int32_t* col_of_id,
                          bnpc_visit_t* visit_t0, bnpc_cand_t* cand_t0, int C, double slack,
                          double c_norm, int32_t* blk, void* stream);
/* blk (scratch, [ceil(C/128)+1] ints) receives per-block counts of the visits that are not
 * BNPC_VISIT_CERTAIN.  bnpc_gibbs_compact then copies those visit and option records, in
 * visiting order, to visit_c / cand_c and writes their number to st[BNPC_ST_NUNC]: the
 * sequential sweep walks only them while no cluster is born and every cluster keeps two cells. */
int bnpc_gibbs_compact(const bnpc_visit_t* visit_t0, const bnpc_cand_t* cand_t0, int C,
                       int32_t* blk, bnpc_visit_t* visit_c, bnpc_cand_t* cand_c, int32_t* st,
                       void* stream);
/* ---- lean epoch (see bnpc_opt_t): approximate rows -> options -> FP64 only for the options of
 * the uncertain visits.  bnpc_ll_matrix_f32: float copy of lp (lpf [K][M][2] is written first)
 * and FP32-FMA rows llf[r][k], r < C.  bnpc_gibbs_options: opt[r], certain-visit counts per
 * column in n_cert[K] (zeroed inside); terms = number of summed terms of a row (2*M) for the
 * error bound of the approximate rows.  bnpc_gibbs_exact: finalises the certain flags (a
 * cluster never keeps a single certain visit), compacts the uncertain visits in visiting order
 * (idx_c, st[BNPC_ST_NUNC]) and writes their visit / option records with FP64 log-likelihoods in
 * the arithmetic of bnpc_ll_matrix (order: scratch of C ints, the processing order of the uncertain
 * visits grouped by own cluster; on return its first st[BNPC_ST_NUNC] BYTES hold the owner warp of
 * every compacted record, bnpc_sweep_args_t.owner_c).  comp (int32[512]) receives the option graph on the columns
 * (which clusters share a visit), its connected components and the sequencer warp that owns
 * each column: visits of different components never interact, the sweep walks them in parallel. */
int bnpc_ll_matrix_f32(const uint32_t* x1, const uint32_t* x0, int W, int M, const int32_t* cells,
                       int cell_stride, int C, const double* lp, float* lpf, int K, float* llf,
                       int ldf, void* stream);
/* The same rows on the tcgen05 tensor cores: lp split into two bf16 terms (bsplit: scratch of
 * W * 2*Kp * 64 bf16, Kp = K rounded up to 8), 0/1 data expanded from the bit-planes into
 * tensor memory, FP32 accumulation; llf[r][0..Kp) is written, ldf >= Kp, ldf % 4 == 0.        */
int bnpc_ll_matrix_tc(const uint32_t* x1, const uint32_t* x0, int W, int M, const int32_t* cells,
                      int cell_stride, int C, const double* lp, uint16_t* bsplit, int K, float* llf,
                      int ldf, void* stream);
/* The same rows on the tensor cores in exact integer arithmetic (tcgen05.mma kind::i8): -lp is
 * quantised to 16-bit fixed point with step q = vmax / 65535 (the caller guarantees |lp| <= vmax;
 * larger entries are clamped), the two base-256 digits are separate columns (bdigits: scratch of
 * W * Kp * 128 bytes), int32 accumulators.  A row is off by at most (observed entries) * q / 2
 * (<= M * q / 2: pass it to bnpc_gibbs_options as err_abs) plus one float rounding.  M < 32768. */
/* ---- posterior (MPEAR) estimator, reference libs/utils.py:90-145 -------------------------
 * bnpc_cocluster_counts: assign int32 [S][N] (posterior samples of the assignment vector) ->
 * counts int32 [N(N-1)/2], the number of samples in which cells i < j sit in different clusters,
 * condensed in scipy pdist order; the reference's get_dist is counts / S (libs/utils.py:90-97).
 * bnpc_mpear_sums: labels int32 [n_cand][N] (candidate cuts of the dendrogram) -> out uint64
 * [1 + 2 n_cand]: out[0] = sum of all counts, out[1+2c] = pairs with equal labels under c,
 * out[2+2c] = sum of counts over those pairs (the three sums of libs/utils.py:133-145 as exact
 * integers).  out is zeroed inside. */
int bnpc_cocluster_counts(const int32_t* assign, int S, int N, int32_t* counts, void* stream);
int bnpc_mpear_sums(const int32_t* counts, int N, const int32_t* labels, int n_cand,
                    unsigned long long* out, void* stream);
/* The same sums when the N points are the DISTINCT assignment profiles of the cells and point i
 * stands for weight[i] cells with that profile: pair (i, j) counts weight[i] * weight[j] times
 * (pairs inside one profile never differ: the caller adds their number to out[1+2c]).  This is
 * how the estimator reaches 100k cells: the N(N-1)/2 pair vector of the reference (20 GB there)
 * shrinks to the profiles' pairs.                                                             */
int bnpc_mpear_sums_weighted(const int32_t* counts, int N, const int32_t* labels, int n_cand,
                             const int32_t* weight, unsigned long long* out, void* stream);
int bnpc_ll_matrix_i8(const uint32_t* x1, const uint32_t* x0, int W, int M, const int32_t* cells,
                      int cell_stride, int C, const double* lp, uint8_t* bdigits, int K, double vmax,
                      float* llf, int ldf, void* stream);
/* The integer rows of n_chains <= 8 chains over the SAME cells in cell order (row r = cell r, the
 * first epoch of a sweep): chains whose digit tables fit one MMA side by side (sum of 2*Kp <= 256)
 * share the expanded data operand in tensor memory -- one pass over the bit-planes and one TMEM
 * read per MMA for all of them (csrc/bnpc_tc_i8s.cuh).  Per-chain arguments are host arrays; the
 * values equal bnpc_ll_matrix_i8's bit for bit (exact integer accumulation).  This is what the
 * recorded bnpc_ll_matrix_i8 calls (cells = NULL) of the chains of one wave are merged into. */
/* Host logic of that merge, without a launch (no device needed): for n <= 8 chains with paddings
 * kpad[i] (multiples of 8, <= 64) over C cells with W plane words per row -> the group of every
 * chain (first-fit, sum of 2*kpad <= 256 columns per group), its first column in the group, the
 * tiles per supertile of the launch (2 only when every group has <= 128 columns and there are at
 * least 4 tiles per SM), and per group the persistent CTAs and the table chunk slots (= W/2: the
 * tables stay resident in shared memory). */
int bnpc_ll_shared_plan(int n, const int* kpad, int C, int W, int* group_of, int* column_of,
                        int* tiles_per_supertile, int* ctas_of_group, int* slots_of_group, int* n_groups);
int bnpc_ll_matrix_i8_shared(const uint32_t* x1, const uint32_t* x0, int W, int M, int C, int n_chains,
                             const double* const* lp, uint8_t* const* bdigits, const int* K,
                             const double* vmax, float* const* llf, const int* ldf, void* stream);
/* err_abs: absolute error of the approximate rows on top of the FP32 accumulation bound (0 for the
 * float rows). */
int bnpc_gibbs_options(const float* llf, int ldf, int K, const int32_t* col_of_id,
                       const bnpc_visit_t* visit_t0, bnpc_opt_t* opt_t0, int32_t* n_cert, int C,
                       double log_n, double c_norm, int terms, double err_abs, void* stream);
int bnpc_gibbs_exact(const uint32_t* x1, const uint32_t* x0, int W, int M, const double* lp, int K,
                     const bnpc_visit_t* visit_t0, bnpc_opt_t* opt_t0, const int32_t* n_cert, int C,
                     int32_t* blk, int32_t* idx_c, int32_t* st, bnpc_visit_t* visit_c,
                     bnpc_cand_t* cand_c, double log_n, double c_norm, int32_t* comp, int32_t* order,
                     void* stream);
/* Start of an epoch: rebuild cnt[] from the host-authoritative list
 * live[2*j] = id, live[2*j+1] = size (list order), set col_of_id[id] = j and
 * clear the epoch's extra-cluster bookkeeping.  first != 0 also resets the
 * per-sweep counters in st.                                                       */
int bnpc_gibbs_epoch_begin(const int32_t* live, int K, int32_t* lst, int32_t* cnt,
                           int32_t* col_of_id, int idcap, int32_t* st, int first,
                           void* stream);
typedef struct {
    /* data */
    const uint32_t* x1; const uint32_t* x0; int32_t W; int32_t N; int32_t M;
    /* chain state */
    int32_t* assign; int32_t* cnt; int32_t* lst; int32_t* col_of_id; float* theta;
    int32_t idcap; int32_t* st; int32_t* live_out /* [2*idcap] (id,size) at exit */;
    /* epoch */
    const double* ll /* NULL: lean epoch, exact rows are computed on demand from lp */; int32_t ldk;
    int32_t t_epoch0; const double* lp /* [K][M][2] of the epoch (lean epochs) */;
    const int32_t* comp /* lean epochs: option-graph components (bnpc_gibbs_exact) or NULL */;
    double* lpx /* [MAX_EXTRA][M][2] */; double* llx /* [MAX_EXTRA][ldx] */; int32_t ldx;
    double* scratch /* [idcap+1] */;
    /* sweep inputs */
    const bnpc_visit_t* visit; const bnpc_cand_t* cand; int32_t t_begin; int32_t t_end;
    /* compacted records of the epoch's uncertain visits (bnpc_gibbs_compact) or NULL */
    const bnpc_visit_t* visit_c; const bnpc_cand_t* cand_c;
    const double* beta_rows /* parity tape [n_beta_rows][M] or NULL */; int32_t n_beta_rows;
    uint64_t seed; uint64_t stream_id;
    /* constants */
    const double* logn /* [N+1], logn[n] = log(n) as numpy computes it */;
    double c_norm /* log(N-1+alpha) */; double FN; double FP; double p; double q;
    const uint8_t* owner_c /* lean epochs: owner warp per compacted record (bnpc_gibbs_exact) or NULL */;
    int32_t wide /* 1: ll holds option weights exp(ll - ref) and visit[].ref / e_new / c_old are set
                    (dense epoch of at most 63 clusters walked by sweep_wide); 0 otherwise */;
} bnpc_sweep_args_t;
int bnpc_gibbs_sweep(const bnpc_sweep_args_t* args_h, int block_threads, void* stream);

/* ---- sufficient statistics: S1[r][m], S0[r][m] = number of cells of segment r
 * with a 1 / a 0 at mutation m.  Segment r = members[seg_off[r] .. seg_off[r+1]).
 * Replaces the data[cells] gathers of libs/CRP.py:308,360-367 and the [N,M]
 * passes of libs/CRP.py:237-238, libs/CRP_learning_errors.py:58-63.              */
int bnpc_group_members(const int32_t* assign, int N, const int32_t* rank_of_id,
                       const int32_t* seg_off, int32_t* cursor, int K,
                       int32_t* members, void* stream);
int bnpc_set_ranks(const int32_t* ids, int K, int32_t* rank_of_id, void* stream);
int bnpc_suffstat(const uint32_t* x1, const uint32_t* x0, int W, int M,
                  const int32_t* members, const int32_t* seg_off, int R, int max_len,
                  int32_t* S1, int32_t* S0, void* stream);

/* ---- Beta draws for cluster parameters, libs/CRP.py:172-175,183-188:
 * theta_out[r][m] = float32(clip(Beta(p + S1[r][m], q + S0[r][m]), 1e-5, 1-1e-5)).
 * tape != NULL: tape[r][m] holds the Beta variates (parity mode).                */
int bnpc_beta_rows(const int32_t* S1, const int32_t* S0, int R, int M, double p, double q,
                   const double* tape, uint64_t seed, uint64_t stream_id,
                   float* theta_out, const int32_t* out_ids, void* stream);
/* theta_out[ids[r]][m] = float32(clip(u[r][m], 1e-5, 1-1e-5)), libs/CRP.py:176-180 */
int bnpc_theta_from_uniform(const double* u, int R, int M, float* theta_out,
                            const int32_t* out_ids, void* stream);

/* ---- Metropolis-Hastings on cluster parameters, libs/CRP.py:314-383
 * (MH_cluster_params, _get_log_A).  Row r updates theta[ids[r]] (ids==NULL: row
 * r) in place from S1/S0 row r.  rnd = [3][R][M]: proposal-sd index, truncnorm
 * uniform, acceptance uniform.  flags: bit0 = want transition log-probabilities
 * (trans_prob=True): logq[r][m] is written and A is clipped at 0.
 * declined[r] is incremented by the number of rejected proposals of row r.       */
int bnpc_mh_theta(float* theta, const int32_t* ids, int R, int M,
                  const int32_t* S1, const int32_t* S0, const double* rnd,
                  double FN, double FP, double p, double q, int flags,
                  double* logq, int32_t* declined, void* stream);
/* A[r][m] of libs/CRP.py:347-383 for given (new, old) rows without a draw (used
 * by libs/CRP.py:674-681 and :777-797): bounds are lo=(blo-old)/sd, hi=(bhi-old)/sd
 * with the subtraction rounded in float32; A clipped at 0.                        */
int bnpc_theta_log_ratio(const float* th_new, const float* th_old, int R, int M,
                         const int32_t* S1, const int32_t* S0, const double* sd_idx,
                         float blo, float bhi, double FN, double FP, double p, double q,
                         double* A, void* stream);

/* ---- reductions over [R][M]: libs/CRP.py:237-251, 716-733,
 * libs/CRP_learning_errors.py:58-63.  For each of the E (FN,FP) pairs
 * out[e*R + r] = sum_m S1*log p1 + S0*log p0 for row r (theta[ids[r]]);
 * if prior_out != NULL, prior_out[r] = sum_m Beta(p,q).logpdf(theta[ids[r]][m]).
 * Deterministic (fixed summation order).                                          */
int bnpc_row_loglik(const float* theta, const int32_t* ids, int R, int M,
                    const int32_t* S1, const int32_t* S0,
                    const double* fn_h, const double* fp_h, int E,
                    double p, double q, double* out, double* prior_out, void* stream);
/* out[r] = sum_m v[r][m] (fixed order) */
int bnpc_row_sum(const double* v, int R, int M, double* out, void* stream);

/* ---- split-merge support, libs/CRP.py:434-567,609-638,777-820 ---------------- */
/* ordered member lists: cells_out = [cells of id_a ascending..., cells of id_b
 * ascending...] (id_b < 0: only id_a).  blk is scratch [ceil(N/1024)*2+2].        */
int bnpc_gather_members(const int32_t* assign, int N, int id_a, int id_b,
                        int32_t* cells_out, int32_t* blk, void* stream);
/* the anchor swaps of libs/CRP.py:449-450 (split: n_a = n) and :496,:500 (merge) */
int bnpc_anchor_swaps(int32_t* cells, int n, int n_a, int idx_i, int idx_j, int is_merge,
                      void* stream);
/* libs/CRP.py:547-561: half[s] = 1 if the j anchor's raw row explains free cell s
 * better than the i anchor's.  k6 = log-likelihood constants for (anchor,cell) =
 * (1,1),(1,0),(0,1),(0,0),(missing,1),(missing,0).                                */
int bnpc_rg_launch_halves(const uint32_t* x1, const uint32_t* x0, int W,
                          const int32_t* cells, int n, const double* k6_h,
                          int32_t* half, void* stream);
/* split `cells` by half into members = [side i..., side j...]; seg_off[0..2]      */
int bnpc_rg_sides(const int32_t* cells, int n, const int32_t* half,
                  int32_t* members, int32_t* seg_off, void* stream);
/* one restricted Gibbs scan over the n-2 free cells, libs/CRP.py:609-632
 * (mode 0: sampled, order = perm, uniforms u) or the forced replay of
 * libs/CRP.py:806-818 (mode 1: side = (assign[cell] != id_i), natural order).
 * ll2 [n-2][ldk>=2] from bnpc_ll_matrix.  lq[c] = log-probability of the side
 * taken (NULL: not wanted).  half is updated in place.  work is scratch of
 * 2*(n-2)+8 int32.  Three launches: per-cell count thresholds in parallel (the
 * two-way draw is monotone in the running side-j count), a serial integer
 * pass, and a parallel write-back.                                                */
int bnpc_rg_scan(const double* ll2, int ldk, int n, const int32_t* perm, const double* u,
                 int32_t* half, double alpha, int mode, const int32_t* cells,
                 const int32_t* assign, int id_i, double* lq, int32_t* work, void* stream);
/* accepted split (libs/CRP.py:471-474): assign[cell] = new_id for side-j cells;
 * accepted merge (libs/CRP.py:517): assign[cells[n_a..n)] = id.                   */
int bnpc_apply_split(const int32_t* cells, int n, const int32_t* half, int new_id,
                     int32_t* assign, void* stream);
int bnpc_apply_merge(const int32_t* cells, int n_a, int n, int id, int32_t* assign,
                     void* stream);

/* ==== chain workspace: one C call per model method ================================
 * The entry points above launch one kernel each.  The ones below enqueue, in one call, every
 * launch a method of the reference's model classes needs (libs/MCMC.py:320-342 calls them once
 * per step), reading their buffers from a workspace the caller fills once.  All device buffers
 * are still allocated and owned by the caller; h_* are PINNED HOST buffers used for the small
 * per-call inputs/outputs (copied asynchronously on `stream`; the caller synchronises the
 * stream before reading h_out / h_scal).  Nothing is allocated, nothing synchronises.        */
typedef struct {
    /* read-only data shared by the chains of one device */
    const uint32_t* x1; const uint32_t* x0; const int32_t* n1; const int32_t* n0; const double* logn;
    int32_t W; int32_t N; int32_t M; int32_t idcap;
    /* chain state */
    int32_t* assign; float* theta /* [idcap][M] */; int32_t* cnt; int32_t* lst; int32_t* col_of_id;
    int32_t* rank_of_id; int32_t* live_io /* [2*idcap] */; int32_t* st /* [BNPC_ST_WORDS] */;
    /* Gibbs sweep */
    bnpc_visit_t* visit; bnpc_cand_t* cand; bnpc_visit_t* visit_c; bnpc_cand_t* cand_c /* [N] each */;
    int32_t* cblk /* [N/128+2] */; int32_t* perm /* [N] */; double* u /* [N] */;
    double* lp /* [K][M][2] */; double* ll /* [rows][ldk] */; double* lpx /* [MAX_EXTRA][M][2] */;
    double* llx /* [MAX_EXTRA][rows] */; double* scratch /* [idcap+1] */;
    /* lean epochs */
    float* lpf /* [K][M][2] */; float* llf /* [N][ldf] */; bnpc_opt_t* opt /* [N] */;
    int32_t* n_cert /* [BNPC_LEAN_MAXK] */; int32_t* idx_c /* [N] */;
    uint16_t* bsplit /* [W][2*BNPC_LEAN_MAXK][64] bf16 */; int32_t* comp /* [512] */;
    /* sufficient statistics of the live clusters, list order */
    int32_t* ids; int32_t* seg; int32_t* cursor /* [K+1] each */; int32_t* members /* [N] */;
    int32_t* S1; int32_t* S0 /* [K][M] */; double* rnd /* [3][K][M] */; int32_t* declined /* [K+1] */;
    double* rl_out /* [5][K] */; double* rl_tot /* [8] */;
    /* split-merge */
    int32_t* cells /* [N+8] */; int32_t* half /* [N+8] */; int32_t* gblk; int32_t* seg3 /* [8] */;
    int32_t* rg_work /* [2N+16] */; float* rg_theta /* [3][M] */; int32_t* rg_S1; int32_t* rg_S0 /* [3][M] */;
    int32_t* rg_dec /* [4] */; double* rg_scal /* [32] */; double* rg_lp /* [2][M][2] */;
    double* rg_ll2 /* [N][2] */; double* rg_lq /* [N] */; double* rg_logq /* [3][M] */; double* rg_A /* [2][M] */;
    float* rg_orig /* [2][M] */; int32_t* rg_perm /* [N] */; double* rg_u /* [N] */; double* rg_rnd /* [3][2][M] */;
    double* rg_sd /* [2][M] */; double* rg_beta /* [3][M], parity tape */;
    /* pinned host staging */
    int32_t* h_in; int32_t* h_out; double* h_scal;
} bnpc_chain_t;

/* One epoch of a Gibbs sweep (libs/CRP.py:254-299) over visits [t, t+rows): the live list
 * (id, size pairs in list order) is read from h_in[0..2K); on return h_out holds the status
 * block st[BNPC_ST_WORDS] followed by the live list after the epoch.  first != 0 starts a sweep:
 * visiting order and uniforms are drawn (streams stream_id+1, +2 of `seed`) unless rand_ready
 * says perm/u already hold them (parity tape), and the visit records are built.  Cluster births
 * use stream_id + 2^24*(b+1) or the rows of beta_rows.                                       */
typedef struct {
    int32_t first; int32_t K; int32_t t; int32_t rows; int32_t ldk; int32_t rand_ready;
    int32_t lean /* 0: dense FP64 matrix; lean epoch (K <= BNPC_LEAN_MAXK) with approximate rows
                    from 1: FP32 FMA, 2: tcgen05 bf16-split, 3: tcgen05 integer digits;
                    -1: dense FP64 matrix turned into option weights, all visits walked by one warp
                    with lanes <-> clusters (K <= 63; data whose visits mostly have > 8 rivals) */;
    int32_t serial_sweep /* lean epochs: 1 = one sequencer warp instead of one per component group */;
    double c1; double c0; double lnew_prior; double c_norm; double log_n;
    double FN; double FP; double p; double q;
    uint64_t seed; uint64_t stream_id;
    const double* beta_rows; int32_t n_beta_rows;
    /* optional cudaEvent_t handles recorded around the ll matrix and the sweep launch (NULL: none) */
    void* ev_ll0; void* ev_ll1; void* ev_sw0; void* ev_sw1;
} bnpc_epoch_t;
int bnpc_chain_gibbs_epoch(const bnpc_chain_t* w, const bnpc_epoch_t* e, void* stream);

/* S1/S0 of the K live clusters (libs/CRP.py:308,360-367): h_in = ids[K], seg[K+1] (segment
 * offsets = running sum of the sizes), max_len = largest size.                              */
int bnpc_chain_stats(const bnpc_chain_t* w, int K, int max_len, void* stream);
/* update_parameters (libs/CRP.py:302-344) for the K live clusters; draws from streams
 * stream_id+1, +2 unless rand_ready (rnd already filled); h_out[0] = proposals declined.    */
int bnpc_chain_mh_theta(const bnpc_chain_t* w, int K, int rand_ready, uint64_t seed, uint64_t stream_id,
                        double FN, double FP, double p, double q, void* stream);
/* h_scal[e] = full-data log-likelihood at (fn_h[e], fp_h[e]), e < E <= 4 (libs/CRP.py:237-238,
 * libs/CRP_learning_errors.py:58-63); want_prior: h_scal[E] = sum of Beta(p,q).logpdf(theta)  */
int bnpc_chain_loglik(const bnpc_chain_t* w, int K, const double* fn_h, const double* fp_h, int E,
                      int want_prior, double p, double q, void* stream);

/* theta rows of the n cluster ids in h_in[0..n) -> dst_h (pinned host [n][M] float): the trace
 * read of libs/MCMC.py:281-282 (`model.parameters[clusters]`)                                 */
int bnpc_chain_theta_rows(const bnpc_chain_t* w, int n, float* dst_h, void* stream);
/* plumbing for hosts without their own CUDA runtime binding: asynchronous copy on the stream
 * (kind 1 host->device, 2 device->host, 3 device->device) and stream synchronisation           */
int bnpc_copy_async(void* dst, const void* src, int64_t bytes, int kind, void* stream);
int bnpc_stream_sync(void* stream);

/* A split-merge move (libs/CRP.py:434-567).  n cells of cluster cl_i (split: cl_j = -1) or of
 * cl_i then cl_j (merge, n_a cells in cl_i); a_i, a_j = anchor positions.  Random streams
 * stream_id+1.. of `seed` per call (at most 4) unless rand_ready (rg_perm, rg_u, rg_rnd, rg_sd,
 * rg_beta filled from the parity tape).  rg_scal slots: 0 forward assignment log-prob, 1 forward
 * theta log-prob, 2 backward theta, 3 backward assignment, 4-5 prior of the new rows, 6-7 of the
 * old rows, 8-10 log-likelihood of side i, side j, all cells.                                */
typedef struct {
    int32_t n; int32_t n_a; int32_t cl_i; int32_t cl_j; int32_t a_i; int32_t a_j; int32_t is_merge;
    int32_t rand_ready;
    double alpha; double FN; double FP; double p; double q; double k6[6];
    uint64_t seed; uint64_t stream_id;
} bnpc_rg_t;
int bnpc_chain_rg_setup(const bnpc_chain_t* w, const bnpc_rg_t* g, void* stream);
int bnpc_chain_rg_scan_split(const bnpc_chain_t* w, const bnpc_rg_t* g, int want_logq, void* stream);
int bnpc_chain_rg_scan_merged(const bnpc_chain_t* w, const bnpc_rg_t* g, int want_logq, void* stream);
/* decision scalars -> h_scal[0..16) (= rg_scal), split also seg3 -> h_out[0..8)             */
int bnpc_chain_rg_decide_split(const bnpc_chain_t* w, const bnpc_rg_t* g, int flat_prior, void* stream);
int bnpc_chain_rg_decide_merge(const bnpc_chain_t* w, const bnpc_rg_t* g, int flat_prior, void* stream);
int bnpc_chain_rg_apply(const bnpc_chain_t* w, const bnpc_rg_t* g, int new_id, void* stream);

/* ==== chain group: all chains of one GPU stepped in lockstep by one host thread ==============
 * The reference runs one process per chain (libs/MCMC.py:113-120), each looping over
 * Chain.do_step + Chain.update_results (libs/MCMC.py:320-342, 242-282).  bnpc_group_run does the
 * same for n chains of one device at once: per step the launches of all chains are recorded
 * through the bnpc_chain_* entry points above and issued as chain-batched launches (one launch
 * per kernel for up to 8 chains, blockIdx.z = chain), with one stream synchronisation per phase
 * for the whole group.  Every chain draws from its own counter-based streams, so its trace does
 * not depend on the group it runs in.
 *
 * bnpc_chain_state_t is the host-side state of one chain: constants of the model, then the
 * mutable part (imported when a run starts, written back when it returns).  `live` is a HOST
 * array of (id, size) pairs in list order.                                                    */
typedef struct {
    /* model constants (libs/CRP.py:27-66, libs/CRP_learning_errors.py:18-32) */
    double p; double q; double mix0; double mix1;       /* Beta prior, _beta_mix_const            */
    double dp_a0; double dp_b0;                         /* Gamma prior of alpha: shape, loc        */
    double fp_mean; double fp_sd; double fn_mean; double fn_sd;
    double fp_prior_const; double fn_prior_const;       /* truncnorm._logpdf(0, lo, hi) of the priors */
    int32_t learning; int32_t beta_prior_uniform;
    /* route switches of the sweep (bnpc_epoch_t.lean / serial_sweep) */
    int32_t lean_enabled; int32_t lean_rows; int32_t serial_sweep; int32_t wide_enabled; int32_t force_wide;
    /* mutable */
    int32_t lean_ok; int32_t lean_cooldown; int32_t stats_fresh;
    double DP_a; double FN; double FP;
    uint64_t seed; uint64_t host_ctr /* host scalars drawn */; uint64_t dev_calls /* device streams reserved */;
    int32_t K; int32_t live_cap /* pairs' ints available in live */; int32_t* live;
    int64_t ll_cap; int64_t llx_cap /* doubles in bnpc_chain_t.ll / .llx */;
    double mh_counter[10] /* [params, splits, merges, FP, FN] x [accepted, declined] */;
    int32_t last_epochs; int32_t last_births; int32_t last_moved; int32_t last_nunc;
} bnpc_chain_state_t;

/* move probabilities of libs/MCMC.py:26-60 */
typedef struct {
    double sm_prob; double dpa_prob; double error_prob; double sm_ratios[2];
    int32_t sm_steps; int32_t fix_assign;
} bnpc_moves_t;

/* per-chain trace destinations (libs/MCMC.py:231-282).  Scalar traces are host arrays indexed by
 * step.  assign_h: PINNED host rows [.][assign_stride] int32 (NULL: the assignment trace stays in
 * the device ring).  params_h: PINNED host [.][params_kcap][M] float, row (step - params_first)
 * for steps >= params_first (NULL: not recorded): theta rows of the SORTED live cluster ids.     */
typedef struct {
    double* ml; double* map; double* alpha; double* fn; double* fp; int32_t* n_clusters;
    int32_t* assign_h; int64_t assign_stride;
    float* params_h; int32_t params_kcap; int32_t params_first;
} bnpc_trace_t;

/* the library never allocates device memory: when a chain needs more room it calls back.
 * kind: BNPC_GROW_IDS (cluster id capacity >= need; the caller regrows the K-sized buffers and
 * updates the bnpc_chain_t in place), _LL / _LLX (doubles in .ll / .llx; update ll_cap / llx_cap),
 * _LIVE (ints in state.live), _PARAMS (trace.params_kcap >= need), _RING_K (theta ring clusters
 * >= need: call bnpc_group_set_ring).  Returns 0 on success.                                  */
#define BNPC_GROW_IDS 1
#define BNPC_GROW_LL 2
#define BNPC_GROW_LLX 3
#define BNPC_GROW_LIVE 4
#define BNPC_GROW_PARAMS 5
#define BNPC_GROW_RING_K 6
typedef int (*bnpc_grow_fn)(void* ctx, int chain, int kind, int64_t need);

typedef struct bnpc_group bnpc_group_t;
/* ws / st / tr: arrays of n pointers, kept by the group (the caller owns the structs and may
 * update their fields from the grow callback).  streams: main (Gibbs + parameters), side
 * (split-merge moves), copy (trace drain), cudaStream_t as void*.                             */
bnpc_group_t* bnpc_group_create(int n, bnpc_chain_t* const* ws, bnpc_chain_state_t* const* st,
                                bnpc_trace_t* const* tr, const bnpc_moves_t* moves,
                                bnpc_grow_fn grow, void* grow_ctx,
                                void* stream_main, void* stream_side, void* stream_copy);
/* device trace rings: assign [slots][n][N] int32, theta [slots][n][kcap][M] float */
int bnpc_group_set_ring(bnpc_group_t* g, int slots, int32_t* ring_assign, float* ring_theta, int kcap);
/* steps step0 .. step0+n_steps-1 (trace rows of those indices) for all chains; returns after the
 * traces have reached the host                                                                */
int bnpc_group_run(bnpc_group_t* g, int step0, int n_steps);
/* the trace row of the CURRENT state without a move (Chain.update_results(0), libs/MCMC.py:218) */
int bnpc_group_record(bnpc_group_t* g, int step);
void bnpc_group_destroy(bnpc_group_t* g);

/* host scalars of a chain's random stream (draw *ctr, advanced): the Python mirror of the model
 * draws from the same functions so that both hosts walk the same trajectory                    */
double bnpc_host_random(uint64_t seed, uint64_t* ctr);
double bnpc_host_gamma(uint64_t seed, uint64_t* ctr, double shape);
double bnpc_host_beta(uint64_t seed, uint64_t* ctr, double a, double b);
/* scipy.stats.truncnorm ppf / logpdf for scalars as the error-rate move uses them
 * (libs/CRP_learning_errors.py:82-91)                                                          */
double bnpc_host_truncnorm_ppf(double q, double lo, double hi);
double bnpc_host_truncnorm_logpdf(double x, double lo, double hi, double loc, double scale);

/* recorded launches: between begin and flush the bnpc_chain_* calls of the calling thread are
 * queued under `slot` (bnpc_batch_slot) instead of launched; flush issues them on `stream`,
 * merging the same launch of different slots into one chain-batched launch.                   */
int bnpc_batch_begin(void);
int bnpc_batch_slot(int slot);
int bnpc_batch_flush(void* stream);
/* per-kernel device times of the calling thread's launches (CUDA events around every launch; for
 * profiling passes, not for timed runs).  report: text lines "name launches chains total_ms".   */
int bnpc_prof_enable(int on);
int bnpc_prof_report(char* buf, int cap);

#ifdef __cplusplus
}
#endif
#endif /* BNPC_B200_H */
