// Composite entry points: one C call enqueues everything one method of the model needs on the
// chain's stream (declared in include/bnpc_b200.h, "chain workspace" section).  They only
// sequence the single-kernel entry points above; with several chains per GPU driven by host
// threads this keeps the per-step host work (and the time the interpreter lock is held) small.
// Included at the end of bnpc_kernels.cu, inside extern "C".

#define TRY(call)                  \
    do {                           \
        const int rc__ = (call);   \
        if (rc__ != 0) return rc__; \
    } while (0)

__device__ __forceinline__ void add_rows_kernel(const int32_t* __restrict__ a, const int32_t* __restrict__ b,
                                int32_t* __restrict__ c, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) c[i] = a[i] + b[i];
}
__device__ __forceinline__ void copy_rows_kernel(const float* __restrict__ src0, const float* __restrict__ src1,
                                 float* __restrict__ dst, int M) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M) dst[i] = src0[i];
    else if (i < 2 * M && src1) dst[i] = src1[i - M];
}
__device__ __forceinline__ void sum_int_kernel(const int32_t* __restrict__ v, int n, int32_t* out) {
    __shared__ int red[32];
    int s = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
        if (threadIdx.x == 0) *out = s;
    }
}

__device__ __forceinline__ void gather_rows_kernel(const float* __restrict__ theta, const int32_t* __restrict__ ids, int n, int M,
                                   float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n * M) return;
    const int r = (int)(i / M), m = (int)(i % M);
    out[i] = theta[(long long)ids[r] * M + m];
}

int bnpc_copy_async(void* dst, const void* src, int64_t bytes, int kind, void* stream) {
    if (bytes <= 0) return 0;
    const cudaMemcpyKind k = kind == 1 ? cudaMemcpyHostToDevice
                             : kind == 2 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    return copy_async(dst, src, (size_t)bytes, k, stream);
}

int bnpc_stream_sync(void* stream) {
    cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess) return fail("cudaStreamSynchronize", e);
    return 0;
}

// theta rows of the n cluster ids in h_in[0..n) -> dst_h (pinned host, [n][M] float); uses the
// `cursor` and `rnd` scratch buffers of the workspace
int bnpc_chain_theta_rows(const bnpc_chain_t* w, int n, float* dst_h, void* stream) {
    if (!w || n <= 0 || n > w->idcap) return bad_arg("workspace/n");
    TRY(copy_async(w->cursor, w->h_in, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, stream));
    float* stage = reinterpret_cast<float*>(w->rnd);
    BNPC_LAUNCH(gather_rows_kernel, 0, 0, cdiv((long long)n * w->M, 256), 256, 0, (cudaStream_t)stream, w->theta, w->cursor, n, w->M, stage);
    TRY(copy_async(dst_h, stage, sizeof(float) * (size_t)n * w->M, cudaMemcpyDeviceToHost, stream));
    return 0;
}

int bnpc_chain_gibbs_epoch(const bnpc_chain_t* w, const bnpc_epoch_t* e, void* stream) {
    if (!w || !e) return bad_arg("workspace/epoch");
    const int N = w->N, M = w->M, K = e->K, t = e->t, rows = e->rows, ldk = e->ldk;
    if (K <= 0 || K > w->idcap || rows <= 0 || t < 0 || t + rows > N) return bad_arg("epoch range");
    if (e->first) {
        // production: the records are built from the chain's streams (stream_id+1 visiting order,
        // +2 uniforms); parity mode: from the taped perm / u
        TRY(gibbs_prepare_impl(e->rand_ready ? w->perm : nullptr, e->rand_ready ? w->u : nullptr, w->assign, w->n1,
                               w->n0, N, e->c1, e->c0, e->lnew_prior, w->visit, e->seed, e->stream_id, stream));
    }
    const bool lean_epoch = e->lean > 0;
    if (lean_epoch) {
        // scratch of the option / exact passes, cleared in one launch
        TRY(zero_async(w->n_cert, sizeof(int32_t) * BNPC_LEAN_MAXK, stream, "epoch memset"));
        TRY(zero_async(w->comp, sizeof(int32_t) * 512, stream, "epoch memset"));
    }
    // live list (id, size) pairs in list order: host staging -> device
    TRY(copy_async(w->live_io, w->h_in, sizeof(int32_t) * 2 * (size_t)K, cudaMemcpyHostToDevice, stream));
    TRY(bnpc_gibbs_epoch_begin(w->live_io, K, w->lst, w->cnt, w->col_of_id, w->idcap, w->st, e->first, stream));
    TRY(bnpc_logprob_tables(w->theta, w->lst, K, M, e->FN, e->FP, w->lp, stream));
    // cell indices are read straight out of the visit records
    const int32_t* cells = &w->visit[t].cell;
    const int cstride = (int)(sizeof(bnpc_visit_t) / 4);
    const bool lean = e->lean > 0;
    const bool wide = e->lean < 0;               // dense FP64 matrix turned into option weights (sweep_wide)
    if (wide && K > 63) return bad_arg("wide epoch with more than 63 clusters");
    const bool compacted = !wide && (lean || ldk <= SW_MAXL);
    if (lean) {
        // approximate rows pick the options; FP64 only for the options of the uncertain visits
        if (K > BNPC_LEAN_MAXK) return bad_arg("lean epoch with K > BNPC_LEAN_MAXK");
        const int ldf = (K + 7) & ~7;
        TRY(record_event(e->ev_ll0, stream));
        double err_abs = 0.0;
        int by_cell = 0;
        if (e->lean == 3) {
            // integer rows: log-probabilities are linear in theta in [1e-5, 1 - 1e-5] before the
            // log, so the most negative one sits at an end of that interval
            const double tl = 1e-5 * 0.999, th = 1.0 - tl;
            const double pmin = fmin(fmin(tl * (1 - e->FN) + th * e->FP, th * (1 - e->FN) + tl * e->FP),
                                     fmin(tl * e->FN + th * (1 - e->FP), th * e->FN + tl * (1 - e->FP)));
            const double vmax = -log(fmax(pmin, 1e-300)) * 1.0001 + 1e-6;
            err_abs = (double)M * (vmax / 65535.0) * 0.5;
            // an epoch over all cells writes its rows in CELL order: every chain of the GPU then
            // reads the same tiles and chains in the same wave share them (bnpc_tc_i8s.cuh)
            by_cell = (t == 0 && rows == N) ? 1 : 0;
            TRY(bnpc_ll_matrix_i8(w->x1, w->x0, w->W, M, by_cell ? nullptr : cells, cstride, rows, w->lp,
                                  reinterpret_cast<uint8_t*>(w->bsplit), K, vmax, w->llf, ldf, stream));
        } else if (e->lean == 2)
            TRY(bnpc_ll_matrix_tc(w->x1, w->x0, w->W, M, cells, cstride, rows, w->lp, w->bsplit, K, w->llf, ldf,
                                  stream));
        else
            TRY(bnpc_ll_matrix_f32(w->x1, w->x0, w->W, M, cells, cstride, rows, w->lp, w->lpf, K, w->llf, ldf,
                                   stream));
        TRY(record_event(e->ev_ll1, stream));
        TRY(gibbs_options_impl(w->llf, ldf, K, w->col_of_id, w->visit + t, w->opt + t, w->n_cert, rows, e->log_n,
                               e->c_norm, 2 * M, err_abs, false, stream, by_cell));
        TRY(gibbs_exact_impl(w->x1, w->x0, w->W, M, w->lp, K, w->visit + t, w->opt + t, w->n_cert, rows, w->cblk,
                             w->idx_c, w->st, w->visit_c, w->cand_c, e->log_n, e->c_norm, w->comp, w->rg_perm, false,
                             stream));
    } else {
        TRY(record_event(e->ev_ll0, stream));
        TRY(bnpc_ll_matrix(w->x1, w->x0, w->W, M, cells, cstride, rows, w->lp, K, w->ll, ldk, stream));
        if (wide) {
            BNPC_LAUNCH(gibbs_weights_kernel, 256, 0, cdiv(rows, 8), 256, 0, (cudaStream_t)stream, w->ll, ldk, K, w->col_of_id, w->visit + t, rows, e->c_norm);
        }
        TRY(record_event(e->ev_ll1, stream));
        if (compacted) {
            TRY(bnpc_gibbs_candidates(w->ll, ldk, K, w->col_of_id, w->visit + t, w->cand + t, rows, e->log_n,
                                      e->c_norm, w->cblk, stream));
            TRY(bnpc_gibbs_compact(w->visit + t, w->cand + t, rows, w->cblk, w->visit_c, w->cand_c, w->st, stream));
        }
    }
    bnpc_sweep_args_t a;
    memset(&a, 0, sizeof(a));
    a.x1 = w->x1; a.x0 = w->x0; a.W = w->W; a.N = N; a.M = M;
    a.assign = w->assign; a.cnt = w->cnt; a.lst = w->lst; a.col_of_id = w->col_of_id; a.theta = w->theta;
    a.idcap = w->idcap; a.st = w->st; a.live_out = w->live_io;
    a.ll = lean ? nullptr : w->ll; a.ldk = ldk; a.t_epoch0 = t; a.lp = w->lp;
    a.comp = (lean && !e->serial_sweep) ? w->comp : nullptr;
    a.owner_c = (lean && !e->serial_sweep) ? reinterpret_cast<const uint8_t*>(w->rg_perm) : nullptr;
    a.wide = wide ? 1 : 0;
    a.lpx = w->lpx; a.llx = w->llx; a.ldx = rows; a.scratch = w->scratch;
    a.visit = w->visit; a.cand = lean ? nullptr : w->cand; a.t_begin = t; a.t_end = t + rows;
    a.visit_c = compacted ? w->visit_c : nullptr;
    a.cand_c = compacted ? w->cand_c : nullptr;
    a.beta_rows = e->beta_rows; a.n_beta_rows = e->n_beta_rows;
    a.seed = e->seed; a.stream_id = e->stream_id;
    a.logn = w->logn; a.c_norm = e->c_norm; a.FN = e->FN; a.FP = e->FP; a.p = e->p; a.q = e->q;
    TRY(record_event(e->ev_sw0, stream));
    TRY(bnpc_gibbs_sweep(&a, (lean && !e->serial_sweep) ? 512 : (K < 1000 ? 256 : 1024), stream));
    TRY(record_event(e->ev_sw1, stream));
    // status block + live list back to the host staging area (the caller synchronises)
    const int k_cap = (K + BNPC_MAX_EXTRA + 2 < w->idcap) ? K + BNPC_MAX_EXTRA + 2 : w->idcap;
    TRY(copy_async(w->h_out, w->st, sizeof(int32_t) * BNPC_ST_WORDS, cudaMemcpyDeviceToHost, stream));
    TRY(copy_async(w->h_out + BNPC_ST_WORDS, w->live_io, sizeof(int32_t) * 2 * (size_t)k_cap,
                   cudaMemcpyDeviceToHost, stream));
    return 0;
}

int bnpc_chain_stats(const bnpc_chain_t* w, int K, int max_len, void* stream) {
    if (!w || K <= 0) return bad_arg("workspace/K");
    // cursors and statistics cleared in one launch
    TRY(zero_async(w->cursor, sizeof(int32_t) * (size_t)K, stream, "stats memset"));
    TRY(zero_async(w->S1, sizeof(int32_t) * (size_t)K * w->M, stream, "stats memset"));
    TRY(zero_async(w->S0, sizeof(int32_t) * (size_t)K * w->M, stream, "stats memset"));
    // h_in = ids[K] then seg[K+1]
    TRY(copy_async(w->ids, w->h_in, sizeof(int32_t) * (size_t)K, cudaMemcpyHostToDevice, stream));
    TRY(copy_async(w->seg, w->h_in + K, sizeof(int32_t) * ((size_t)K + 1), cudaMemcpyHostToDevice, stream));
    TRY(bnpc_set_ranks(w->ids, K, w->rank_of_id, stream));
    TRY(group_members_impl(w->assign, w->N, w->rank_of_id, w->seg, w->cursor, K, w->members, false, stream));
    TRY(suffstat_impl(w->x1, w->x0, w->W, w->M, w->members, w->seg, K, max_len, w->S1, w->S0, false, stream));
    return 0;
}

int bnpc_chain_mh_theta(const bnpc_chain_t* w, int K, int rand_ready, uint64_t seed, uint64_t stream_id,
                        double FN, double FP, double p, double q, void* stream) {
    if (!w || K <= 0) return bad_arg("workspace/K");
    TRY(zero_async(w->declined, sizeof(int32_t) * ((size_t)K + 1), stream, "mh_theta memset"));
    TRY(mh_theta_impl(w->theta, w->ids, K, w->M, w->S1, w->S0, rand_ready ? w->rnd : nullptr, seed, stream_id, FN, FP,
                      p, q, 0, nullptr, w->declined, stream));
    BNPC_LAUNCH(sum_int_kernel, 0, 0, 1, 256, 0, (cudaStream_t)stream, w->declined, K, w->declined + K);
    TRY(copy_async(w->h_out, w->declined + K, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    return 0;
}

int bnpc_chain_loglik(const bnpc_chain_t* w, int K, const double* fn_h, const double* fp_h, int E,
                      int want_prior, double p, double q, void* stream) {
    if (!w || K <= 0) return bad_arg("workspace/K");
    const int rows = E + (want_prior ? 1 : 0);
    if (rows <= 0) return 0;
    TRY(bnpc_row_loglik(w->theta, w->ids, K, w->M, w->S1, w->S0, fn_h, fp_h, E, p, q, w->rl_out,
                        want_prior ? w->rl_out + (size_t)E * K : nullptr, stream));
    TRY(bnpc_row_sum(w->rl_out, rows, K, w->rl_tot, stream));
    TRY(copy_async(w->h_scal, w->rl_tot, sizeof(double) * rows, cudaMemcpyDeviceToHost, stream));
    return 0;
}

// ---- split-merge (libs/CRP.py:527-567): launch state of the restricted Gibbs sampler ----------
static int rg_side_stats(const bnpc_chain_t* w, int n, void* stream) {
    TRY(zero_async(w->seg3, sizeof(int32_t) * 8, stream, "rg stats memset"));
    TRY(zero_async(w->rg_S1, sizeof(int32_t) * 2 * (size_t)w->M, stream, "rg stats memset"));
    TRY(zero_async(w->rg_S0, sizeof(int32_t) * 2 * (size_t)w->M, stream, "rg stats memset"));
    TRY(rg_sides_impl(w->cells, n, w->half, w->members, w->seg3, false, stream));
    TRY(suffstat_impl(w->x1, w->x0, w->W, w->M, w->members, w->seg3, 2, n, w->rg_S1, w->rg_S0, false, stream));
    return 0;
}

static int rg_mh(const bnpc_chain_t* w, const bnpc_rg_t* g, int row0, int rows, int slot, int* streams_used,
                 void* stream) {
    const int M = w->M;
    const long long RM = (long long)rows * M;
    const uint64_t sid = g->stream_id + *streams_used;      // draws of streams sid+1, sid+2
    if (!g->rand_ready) *streams_used += 2;
    const bool want = slot >= 0;
    TRY(mh_theta_impl(w->rg_theta + (size_t)row0 * M, nullptr, rows, M, w->rg_S1 + (size_t)row0 * M,
                      w->rg_S0 + (size_t)row0 * M, g->rand_ready ? w->rg_rnd : nullptr, g->seed, sid, g->FN, g->FP,
                      g->p, g->q, want ? 1 : 0, want ? w->rg_logq : nullptr, w->rg_dec, stream,
                      row0 == 0 ? w->rg_lp : nullptr));         // rows 0, 1: table of the next scan
    if (want) TRY(bnpc_row_sum(w->rg_logq, 1, (int)RM, w->rg_scal + slot, stream));
    return 0;
}

int bnpc_chain_rg_setup(const bnpc_chain_t* w, const bnpc_rg_t* g, void* stream) {
    if (!w || !g || g->n < 2) return bad_arg("workspace/move");
    const int n = g->n, M = w->M;
    TRY(zero_async(w->rg_scal, sizeof(double) * 32, stream, "rg_setup memset"));
    TRY(bnpc_gather_members(w->assign, w->N, g->cl_i, g->cl_j, w->cells, w->gblk, stream));
    TRY(bnpc_anchor_swaps(w->cells, n, g->n_a, g->a_i, g->a_j, g->is_merge, stream));
    if (n > 2) TRY(bnpc_rg_launch_halves(w->x1, w->x0, w->W, w->cells, n, g->k6, w->half, stream));
    TRY(rg_side_stats(w, n, stream));
    // theta of the two halves, then of all cells of the move (libs/CRP.py:562-566)
    // (the kernels that write the theta rows of the two sides -- this draw and the Metropolis-Hastings
    // update at the end of every scan -- also write their log-probability table rg_lp for the next scan)
    TRY(beta_rows_impl(w->rg_S1, w->rg_S0, 2, M, g->p, g->q, g->rand_ready ? w->rg_beta : nullptr, g->seed,
                       g->stream_id + 1, w->rg_theta, nullptr, w->rg_lp, g->FN, g->FP, stream));
    BNPC_LAUNCH(add_rows_kernel, 0, 0, cdiv(M, 256), 256, 0, (cudaStream_t)stream, w->rg_S1, w->rg_S1 + M, w->rg_S1 + 2 * M, M);
    BNPC_LAUNCH(add_rows_kernel, 0, 0, cdiv(M, 256), 256, 0, (cudaStream_t)stream, w->rg_S0, w->rg_S0 + M, w->rg_S0 + 2 * M, M);
    TRY(bnpc_beta_rows(w->rg_S1 + 2 * M, w->rg_S0 + 2 * M, 1, M, g->p, g->q,
                       g->rand_ready ? w->rg_beta + 2 * (size_t)M : nullptr, g->seed, g->stream_id + 2,
                       w->rg_theta + 2 * (size_t)M, nullptr, stream));
    return 0;
}

// one restricted Gibbs scan over the split state (libs/CRP.py:570-578, 590-632); want_logq: the
// final scan whose transition probabilities enter the acceptance ratio (slots 0 and 1)
int bnpc_chain_rg_scan_split(const bnpc_chain_t* w, const bnpc_rg_t* g, int want_logq, void* stream) {
    if (!w || !g || g->n < 2) return bad_arg("workspace/move");
    const int n = g->n, nf = n - 2, M = w->M;
    int used = 0;
    if (n > 2) {
        // rg_lp = log-probability table of rg_theta rows 0, 1 (written with them)
        TRY(bnpc_ll_matrix(w->x1, w->x0, w->W, M, w->cells + 1, 1, nf, w->rg_lp, 2, w->rg_ll2, 2, stream));
        if (!g->rand_ready) used = 2;                    // streams stream_id+1 (order), +2 (uniforms)
        TRY(rg_scan_impl(w->rg_ll2, 2, n, g->rand_ready ? w->rg_perm : nullptr, g->rand_ready ? w->rg_u : nullptr,
                         g->seed, g->stream_id, w->half, g->alpha, 0, nullptr, nullptr, -1,
                         want_logq ? w->rg_lq : nullptr, w->rg_work, stream));
        if (want_logq) TRY(bnpc_row_sum(w->rg_lq, 1, nf, w->rg_scal + 0, stream));
    }
    TRY(rg_side_stats(w, n, stream));
    // both halves in one launch; draws are taken side 0 first, as the reference does
    TRY(rg_mh(w, g, 0, 2, want_logq ? 1 : -1, &used, stream));
    return 0;
}

// libs/CRP.py:581-587
int bnpc_chain_rg_scan_merged(const bnpc_chain_t* w, const bnpc_rg_t* g, int want_logq, void* stream) {
    if (!w || !g) return bad_arg("workspace/move");
    int used = 0;
    TRY(rg_mh(w, g, 2, 1, want_logq ? 1 : -1, &used, stream));
    return 0;
}

// scalars of the split decision (libs/CRP.py:641-653 with :668-682, :695-733) -> h_scal[0..16), seg3 -> h_out
int bnpc_chain_rg_decide_split(const bnpc_chain_t* w, const bnpc_rg_t* g, int flat_prior, void* stream) {
    if (!w || !g) return bad_arg("workspace/move");
    const int M = w->M;
    const float* th_old = w->theta + (size_t)g->cl_i * M;
    TRY(theta_log_ratio_impl(th_old, w->rg_theta + 2 * (size_t)M, 1, M, w->rg_S1 + 2 * M, w->rg_S0 + 2 * M,
                             g->rand_ready ? w->rg_sd : nullptr, g->seed, g->stream_id + 1, (float)kThetaLo,
                             (float)kThetaHi, g->FN, g->FP, g->p, g->q, w->rg_A, stream));
    TRY(bnpc_row_sum(w->rg_A, 1, M, w->rg_scal + 2, stream));
    if (!flat_prior) {
        TRY(bnpc_row_loglik(w->rg_theta, nullptr, 2, M, w->rg_S1, w->rg_S0, nullptr, nullptr, 0, g->p, g->q,
                            w->rg_scal, w->rg_scal + 4, stream));
        TRY(bnpc_row_loglik(th_old, nullptr, 1, M, w->rg_S1, w->rg_S0, nullptr, nullptr, 0, g->p, g->q,
                            w->rg_scal, w->rg_scal + 6, stream));
    }
    const double fn[1] = {g->FN}, fp[1] = {g->FP};
    TRY(bnpc_row_loglik(w->rg_theta, nullptr, 3, M, w->rg_S1, w->rg_S0, fn, fp, 1, g->p, g->q, w->rg_scal + 8,
                        nullptr, stream));
    TRY(copy_async(w->h_scal, w->rg_scal, sizeof(double) * 16, cudaMemcpyDeviceToHost, stream));
    TRY(copy_async(w->h_out, w->seg3, sizeof(int32_t) * 8, cudaMemcpyDeviceToHost, stream));
    return 0;
}

// scalars of the merge decision (libs/CRP.py:656-665 with :685-692, :736-754, :767-820) -> h_scal[0..16)
int bnpc_chain_rg_decide_merge(const bnpc_chain_t* w, const bnpc_rg_t* g, int flat_prior, void* stream) {
    if (!w || !g) return bad_arg("workspace/move");
    const int M = w->M, n = g->n, nf = n - 2;
    BNPC_LAUNCH(copy_rows_kernel, 0, 0, cdiv(2 * M, 256), 256, 0, (cudaStream_t)stream,  w->theta + (size_t)g->cl_i * M, w->theta + (size_t)g->cl_j * M, w->rg_orig, M);
    // probability of walking from the launch split back to the original split
    TRY(theta_log_ratio_impl(w->rg_orig, w->rg_theta, 2, M, w->rg_S1, w->rg_S0, g->rand_ready ? w->rg_sd : nullptr,
                             g->seed, g->stream_id + 1, 0.0f, 1.0f, g->FN, g->FP, g->p, g->q, w->rg_A, stream));
    TRY(bnpc_row_sum(w->rg_A, 1, 2 * M, w->rg_scal + 2, stream));
    if (n > 2) {
        if (ll_few_fits(2, w->W)) {
            TRY(ll_few_from_theta(w->x1, w->x0, w->W, M, w->cells + 1, 1, nf, w->rg_orig, 2, g->FN, g->FP, w->rg_ll2, 2,
                                  stream));
        } else {
            TRY(bnpc_logprob_tables(w->rg_orig, nullptr, 2, M, g->FN, g->FP, w->rg_lp, stream));
            TRY(bnpc_ll_matrix(w->x1, w->x0, w->W, M, w->cells + 1, 1, nf, w->rg_lp, 2, w->rg_ll2, 2, stream));
        }
        TRY(rg_scan_impl(w->rg_ll2, 2, n, nullptr, nullptr, 0, 0, w->half, g->alpha, 1, w->cells, w->assign, g->cl_i,
                         w->rg_lq, w->rg_work, stream));
        TRY(bnpc_row_sum(w->rg_lq, 1, nf, w->rg_scal + 3, stream));
    }
    if (!flat_prior) {
        TRY(bnpc_row_loglik(w->rg_theta + 2 * (size_t)M, nullptr, 1, M, w->rg_S1, w->rg_S0, nullptr, nullptr, 0,
                            g->p, g->q, w->rg_scal, w->rg_scal + 4, stream));
        TRY(bnpc_row_loglik(w->rg_orig, nullptr, 2, M, w->rg_S1, w->rg_S0, nullptr, nullptr, 0, g->p, g->q,
                            w->rg_scal, w->rg_scal + 6, stream));
    }
    // `half` now equals the original split (reference quirk)
    TRY(rg_side_stats(w, n, stream));
    const double fn[1] = {g->FN}, fp[1] = {g->FP};
    TRY(bnpc_row_loglik(w->rg_theta, nullptr, 3, M, w->rg_S1, w->rg_S0, fn, fp, 1, g->p, g->q, w->rg_scal + 8,
                        nullptr, stream));
    TRY(copy_async(w->h_scal, w->rg_scal, sizeof(double) * 16, cudaMemcpyDeviceToHost, stream));
    return 0;
}

// accepted move: write theta rows and assignments (libs/CRP.py:471-474, 514-517)
int bnpc_chain_rg_apply(const bnpc_chain_t* w, const bnpc_rg_t* g, int new_id, void* stream) {
    if (!w || !g) return bad_arg("workspace/move");
    const int M = w->M;
    if (!g->is_merge) {
        if (new_id < 0 || new_id >= w->idcap) return bad_arg("new_id");
        BNPC_LAUNCH(copy_rows_kernel, 0, 0, cdiv(M, 256), 256, 0, (cudaStream_t)stream, w->rg_theta, nullptr, w->theta + (size_t)g->cl_i * M, M);
        BNPC_LAUNCH(copy_rows_kernel, 0, 0, cdiv(M, 256), 256, 0, (cudaStream_t)stream, w->rg_theta + M, nullptr, w->theta + (size_t)new_id * M, M);
        TRY(bnpc_apply_split(w->cells, g->n, w->half, new_id, w->assign, stream));
    } else {
        BNPC_LAUNCH(copy_rows_kernel, 0, 0, cdiv(M, 256), 256, 0, (cudaStream_t)stream, w->rg_theta + 2 * (size_t)M, nullptr, w->theta + (size_t)g->cl_i * M, M);
        TRY(bnpc_apply_merge(w->cells, g->n_a, g->n, g->cl_i, w->assign, stream));
    }
    return 0;
}

#undef TRY
