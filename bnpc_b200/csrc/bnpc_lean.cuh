// "Lean" epoch of the Gibbs sweep for lists of at most BNPC_LEAN_MAXK clusters.
//
// The sweep needs exact (FP64) log-likelihoods only where a decision depends on them.  An
// APPROXIMATE cells x clusters matrix (FP32 accumulation: FP32 FMA here, bf16-split operands on
// the tcgen05 tensor cores in bnpc_tc.cuh) is enough to prove, with its rounding error bounded,
// that a cluster cannot come within 40 + log N nats of the cell's own cluster for ANY cluster
// sizes -- such clusters sit on the reference's 1e-15 probability floor (libs/CRP.py:88-100).
// Only the surviving (cell, cluster) pairs of the visits that are not statically certain are
// then evaluated in FP64, in exactly the arithmetic of ll_matrix_kernel.
//
//   lp_to_f32_kernel        (log p1, log p0) table in float
//   ll_matrix_f32_kernel    approximate ll rows, FP32 FMA (reference for / fallback of the TC kernel)
//   gibbs_options_kernel    per visit: option columns from the approximate row, provisional
//                           "certain" flag, per-cluster counts of certain visits
//   gibbs_finalize_kernel   a cluster with a single certain visit loses it (the cluster could
//                           shrink to that one cell); per-block counts of uncertain visits
//   compact_index_kernel    visit indices of the uncertain visits, in visiting order
//   gibbs_exact_kernel      FP64 log-likelihoods of their options -> compacted visit/option records

__device__ __forceinline__ void lp_to_f32_kernel(const double2* __restrict__ lp, long long n, float2* __restrict__ lpf) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const double2 v = lp[i]; lpf[i] = make_float2((float)v.x, (float)v.y); }
}

#define LLF_KT 16
__device__ __forceinline__ void ll_matrix_f32_kernel(const uint32_t* __restrict__ x1, const uint32_t* __restrict__ x0, int W, int M,
                     const int32_t* __restrict__ cells, int cell_stride, int C,
                     const float2* __restrict__ lp, int K, float* __restrict__ ll, int ldk) {
    __shared__ float2 tile[LL_MT][LLF_KT];
    const int r = blockIdx.x * LL_THREADS + threadIdx.x;
    const int k0 = blockIdx.y * LLF_KT;
    const bool live = r < C;
    const long long cell = live ? (cells ? cells[(long long)r * cell_stride] : r) : 0;
    const uint4* p1 = reinterpret_cast<const uint4*>(x1 + cell * W);
    const uint4* p0 = reinterpret_cast<const uint4*>(x0 + cell * W);
    float acc[LLF_KT];
#pragma unroll
    for (int kk = 0; kk < LLF_KT; ++kk) acc[kk] = 0.0f;
    for (int m0 = 0; m0 < M; m0 += LL_MT) {
        __syncthreads();
        for (int i = threadIdx.x; i < LL_MT * LLF_KT; i += LL_THREADS) {
            const int kk = i / LL_MT, mm = i % LL_MT;
            float2 v = make_float2(0.0f, 0.0f);
            if (k0 + kk < K && m0 + mm < M) v = lp[(long long)(k0 + kk) * M + m0 + mm];
            tile[mm][kk] = v;
        }
        __syncthreads();
        if (live) {
            const uint4 a = p1[m0 >> 7], b = p0[m0 >> 7];
            const uint32_t w1[4] = {a.x, a.y, a.z, a.w}, w0[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t u1 = w1[q], u0 = w0[q];
                if ((u1 | u0) == 0u) continue;
#pragma unroll 4
                for (int bit = 0; bit < 32; ++bit) {
                    const float f1 = (float)((u1 >> bit) & 1u), f0 = (float)((u0 >> bit) & 1u);
                    const float2* t = tile[q * 32 + bit];
#pragma unroll
                    for (int kk = 0; kk < LLF_KT; ++kk) {
                        const float2 v = t[kk];
                        acc[kk] = fmaf(f1, v.x, fmaf(f0, v.y, acc[kk]));
                    }
                }
            }
        }
    }
    if (live) {
#pragma unroll
        for (int kk = 0; kk < LLF_KT; ++kk)
            if (k0 + kk < K) ll[(long long)r * ldk + k0 + kk] = acc[kk];
    }
}

// err_rel = (number of summed terms) * 2^-22: every term of a row sum is <= 0, so partial sums
// never exceed the total in magnitude and FP32 accumulation (rounded or truncated, in any order)
// is off by at most terms * 2^-23 * |sum|; the factor 2 and the constant are head-room.  err_abs
// = 0.05 + the absolute error of the rows (the quantisation of the integer rows, M * q / 2).
__device__ __forceinline__ void gibbs_options_kernel(const float* __restrict__ llf, int ldf, int K, const int32_t* __restrict__ col_of_id,
                     const bnpc_visit_t* __restrict__ visit, bnpc_opt_t* __restrict__ opt,
                     int32_t* __restrict__ n_cert, int C, float slack, double c_norm, float err_rel,
                     float err_abs, int by_cell) {
    __shared__ int s_cert[BNPC_LEAN_MAXK];
    if (threadIdx.x < BNPC_LEAN_MAXK) s_cert[threadIdx.x] = 0;
    __syncthreads();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < C) {
        // by_cell: the rows were written in cell order (first epoch of a sweep, the tiles shared by
        // the chains of the GPU); otherwise row r belongs to visit r of the epoch
        const float* row = llf + (long long)(by_cell ? visit[r].cell : r) * ldf;
        const int c_old = col_of_id[visit[r].old];
        bnpc_opt_t o;
#pragma unroll
        for (int i = 0; i < BNPC_MAX_OPT; ++i) o.col[i] = 0;
        o.n_opt = BNPC_MAX_OPT + 1; o.i_old = 0; o.flags = BNPC_OPT_MANY; o.pad = 0;
        if (c_old >= 0 && c_old < K) {
            const float v_old = row[c_old];
            const float err = err_rel * (fabsf(v_old) + 64.0f) + err_abs;
            const float thr = v_old - 40.0f - slack - 2.0f * err;
            // options as a bit mask over the columns (rows are read four floats at a time: ldf is a
            // multiple of 4 and the columns beyond K are masked), then the first BNPC_MAX_OPT set
            // bits in increasing order
            unsigned long long mask = 1ull << c_old;
            const float4* row4 = reinterpret_cast<const float4*>(row);
            for (int k4 = 0; k4 < (K + 3) / 4; ++k4) {
                const float4 v = row4[k4];
                const unsigned m4 = (v.x > thr ? 1u : 0u) | (v.y > thr ? 2u : 0u) | (v.z > thr ? 4u : 0u) | (v.w > thr ? 8u : 0u);
                mask |= (unsigned long long)m4 << (4 * k4);
            }
            if (K < 64) mask &= (1ull << K) - 1ull;
            const int n = __popcll(mask);
            const int i_old = __popcll(mask & ((1ull << c_old) - 1ull));
#pragma unroll
            for (int i = 0; i < BNPC_MAX_OPT; ++i) {
                if (mask) {
                    o.col[i] = (uint8_t)(__ffsll((long long)mask) - 1);
                    mask &= mask - 1ull;
                }
            }
            if (n <= BNPC_MAX_OPT) {
                o.n_opt = (uint8_t)n; o.i_old = (uint8_t)i_old; o.flags = 0;
                const double lnew_ll = visit[r].lnew + c_norm;
                const double u = visit[r].u;
                if (n == 1 && lnew_ll < (double)(v_old - 40.0f - err) && u > 3e-10 && u < 1.0 - 3e-10) {
                    o.flags = BNPC_VISIT_CERTAIN;
                    atomicAdd(&s_cert[c_old], 1);
                }
            }
        }
        opt[r] = o;
    }
    __syncthreads();
    if (threadIdx.x < K && s_cert[threadIdx.x]) atomicAdd(&n_cert[threadIdx.x], s_cert[threadIdx.x]);
}

__device__ __forceinline__ void gibbs_finalize_kernel(bnpc_opt_t* __restrict__ opt, const int32_t* __restrict__ n_cert, int C,
                      int32_t* __restrict__ blk) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    int uncertain = 0;
    if (r < C) {
        const uint8_t flags = opt[r].flags;
        bool certain = (flags & BNPC_VISIT_CERTAIN) != 0;
        if (certain && n_cert[opt[r].col[0]] < 2) {
            certain = false;
            opt[r].flags = flags & ~BNPC_VISIT_CERTAIN;
        }
        uncertain = !certain;
    }
    const int cnt = __syncthreads_count(uncertain);
    if (threadIdx.x == 0) blk[blockIdx.x] = cnt;
}

__device__ __forceinline__ void compact_index_kernel(const bnpc_opt_t* __restrict__ opt, int C, const int32_t* __restrict__ blk,
                     int32_t* __restrict__ idx_c) {
    __shared__ int wcnt[CAND_THREADS / 32];
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const bool take = r < C && !(opt[r].flags & BNPC_VISIT_CERTAIN);
    const unsigned m = __ballot_sync(FULL, take);
    if (lane == 0) wcnt[w] = __popc(m);
    __syncthreads();
    if (!take) return;
    int pos = blk[blockIdx.x] + __popc(m & ((1u << lane) - 1u));
    for (int i = 0; i < w; ++i) pos += wcnt[i];
    idx_c[pos] = r;
}

// Processing order of the uncertain visits for gibbs_exact_kernel: grouped by the column of the
// visit's own cluster (counting sort; comp[256..320) histogram, comp[320..384) bases/cursors).
// Visits of one cluster mostly share their option set, so the lanes of a warp read the same table
// entries (shared-memory broadcasts instead of 4-8-way bank conflicts).  The order inside a group
// is arbitrary -- every visit is computed independently and written to its own slot.
__device__ __forceinline__ void exact_hist_kernel(const bnpc_opt_t* __restrict__ opt, const int32_t* __restrict__ idx_c,
                  const int32_t* __restrict__ st, int32_t* __restrict__ comp) {
    __shared__ int h[BNPC_LEAN_MAXK];
    if (threadIdx.x < BNPC_LEAN_MAXK) h[threadIdx.x] = 0;
    __syncthreads();
    const int n_unc = st[BNPC_ST_NUNC];
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_unc) {
        const bnpc_opt_t o = opt[idx_c[j]];
        const int c = (o.n_opt <= BNPC_MAX_OPT) ? (o.col[o.i_old] & (BNPC_LEAN_MAXK - 1)) : 0;
        atomicAdd(&h[c], 1);
    }
    __syncthreads();
    if (threadIdx.x < BNPC_LEAN_MAXK && h[threadIdx.x]) atomicAdd(&comp[256 + threadIdx.x], h[threadIdx.x]);
}

__device__ __forceinline__ void exact_scan_kernel(int32_t* __restrict__ comp) {
    __shared__ int v[BNPC_LEAN_MAXK];
    const int c = threadIdx.x;
    v[c] = comp[256 + c];
    __syncthreads();
    int base = 0;
    for (int i = 0; i < c; ++i) base += v[i];
    comp[320 + c] = base;
}

__device__ __forceinline__ void exact_scatter_kernel(const bnpc_opt_t* __restrict__ opt, const int32_t* __restrict__ idx_c,
                     const int32_t* __restrict__ st, int32_t* __restrict__ comp, int32_t* __restrict__ order) {
    const int n_unc = st[BNPC_ST_NUNC];
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_unc) return;
    const bnpc_opt_t o = opt[idx_c[j]];
    const int c = (o.n_opt <= BNPC_MAX_OPT) ? (o.col[o.i_old] & (BNPC_LEAN_MAXK - 1)) : 0;
    // warp-aggregated cursor bump
    const unsigned active = __activemask();
    const unsigned peers = __match_any_sync(active, c);
    const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&comp[320 + c], __popc(peers));
    base = __shfl_sync(peers, base, leader);
    order[base + __popc(peers & ((1u << lane) - 1u))] = j;
}

#define EX_THREADS 128
#define EX_WORDS 4            /* words of a row (128 mutations) staged per round */
#define EX_COLS 16            /* columns staged per pass */
#define EX_ROWS (32 * EX_WORDS)        /* mutations of a round; row EX_ROWS of the tile is all zeros */
#define EX_STRIDE (2 * (EX_COLS + 1))  /* doubles per tile row: EX_COLS (log p1, log p0) pairs + a zero pair */
// The additions of one round (EX_WORDS words of the visit's row) for NM options, branch-free: a
// missing entry reads the zero row, an option whose column is not staged in this pass (or that the
// visit does not have) reads the zero pair of the row -- adding 0.0 leaves a partial sum unchanged,
// so the sums are those of the guarded loop.  Per entry: one address, then a load and an addition
// per option (the guarded form cost ~22 instructions per option and entry: the compiler re-derived
// the bit tests and the address inside every option's branch).
template <int NM>
__device__ __forceinline__ void ex_accumulate(const double* __restrict__ tile_d, const uint32_t (&u1)[EX_WORDS],
                                              const uint32_t (&u0)[EX_WORDS], const int (&off)[BNPC_MAX_OPT],
                                              double (&part)[EX_WORDS][BNPC_MAX_OPT]) {
#pragma unroll 4
    for (int bit = 0; bit < 32; ++bit) {
#pragma unroll
        for (int q = 0; q < EX_WORDS; ++q) {
            const uint32_t b1 = (u1[q] >> bit) & 1u, b0 = (u0[q] >> bit) & 1u;
            const int row = (b1 | b0) ? q * 32 + bit : EX_ROWS;
            const double* t = tile_d + row * EX_STRIDE + (b1 ? 0 : 1);      // log p1 (.x) or log p0 (.y)
#pragma unroll
            for (int i = 0; i < NM; ++i) part[q][i] += t[off[i]];
        }
    }
}

// One thread per uncertain visit: FP64 log-likelihood of each of its options, then the option
// weights exactly as gibbs_candidates_kernel derives them from the FP64 matrix.  The row is summed
// in four interleaved partial sums per option (word w goes to partial w % 4; the chain of dependent
// additions is the latency bound of this kernel) that are combined as (s0 + s1) + (s2 + s3); a term
// is the log-probability selected by the entry (adding it is what fma(1, lp, acc) does in
// ll_matrix_kernel).  The summation order differs from ll_matrix_kernel's, i.e. values agree to
// ~1e-13 relative, far inside the guard band of the sweep's fast draws (SW_GUARD).
__device__ __forceinline__ void gibbs_exact_kernel(const uint32_t* __restrict__ x1, const uint32_t* __restrict__ x0, int W, int M,
                   const double2* __restrict__ lp, int K, const bnpc_visit_t* __restrict__ visit,
                   const bnpc_opt_t* __restrict__ opt, const int32_t* __restrict__ idx_c,
                   int32_t* __restrict__ st, bnpc_visit_t* __restrict__ visit_c,
                   bnpc_cand_t* __restrict__ cand_c, double slack, double c_norm, int32_t* __restrict__ comp,
                   const int32_t* __restrict__ order) {
    extern __shared__ __align__(16) unsigned char ex_smem[];
    __shared__ unsigned long long s_adj[BNPC_LEAN_MAXK];
    __shared__ int s_num[BNPC_LEAN_MAXK];
    __shared__ unsigned long long s_used;
    __shared__ int s_cols[BNPC_LEAN_MAXK];
    double2* tile = reinterpret_cast<double2*>(ex_smem);          // [EX_ROWS + 1][EX_COLS + 1]
    const double* tile_d = reinterpret_cast<const double*>(ex_smem);
    // (the block may be narrower than EX_THREADS: few uncertain visits are spread over more, smaller
    // CTAs -- one visit per thread either way, so the result does not depend on the block size)
    const int n_unc = st[BNPC_ST_NUNC];
    const int nthr = blockDim.x;
    if (blockIdx.x * nthr >= n_unc) return;
    for (int i = threadIdx.x; i < BNPC_LEAN_MAXK; i += nthr) { s_adj[i] = 0ull; s_num[i] = 0; }
    // the zero pair of every row and the zero row (never overwritten by the staging below)
    for (int i = threadIdx.x; i < EX_ROWS + EX_COLS + 1; i += nthr)
        tile[i < EX_ROWS ? i * (EX_COLS + 1) + EX_COLS : EX_ROWS * (EX_COLS + 1) + (i - EX_ROWS)] = make_double2(0.0, 0.0);
    const int q = blockIdx.x * nthr + threadIdx.x;
    const bool live = q < n_unc;
    const int j = live ? order[q] : order[0];        // slot of the visit among the compacted records
    const int r = idx_c[j];
    bnpc_visit_t v = visit[r];
    const bnpc_opt_t o = opt[r];
    const int nn = (live && o.n_opt <= BNPC_MAX_OPT) ? o.n_opt : 0;
    int col[BNPC_MAX_OPT];
    double part[EX_WORDS][BNPC_MAX_OPT];
#pragma unroll
    for (int i = 0; i < BNPC_MAX_OPT; ++i) {
        col[i] = (i < nn) ? o.col[i] : 0;
#pragma unroll
        for (int q = 0; q < EX_WORDS; ++q) part[q][i] = 0.0;
    }
    int n_max = nn;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) n_max = max(n_max, __shfl_xor_sync(FULL, n_max, s));
    // Only the columns that occur among the options of THIS CTA's visits are staged (the visits
    // arrive grouped by own cluster: a handful of columns instead of all K), EX_COLS per pass:
    // 32 KB of shared memory whatever K is (three CTAs per SM), K/5 of the staging traffic.
    if (threadIdx.x == 0) s_used = 0ull;
    __syncthreads();
    {
        unsigned long long mask = 0ull;
#pragma unroll
        for (int i = 0; i < BNPC_MAX_OPT; ++i)
            if (i < nn) mask |= 1ull << col[i];
        if (mask) atomicOr(&s_used, mask);
    }
    __syncthreads();
    const unsigned long long used = s_used;
    const int n_used = __popcll(used);
    for (int k = threadIdx.x; k < BNPC_LEAN_MAXK; k += nthr)
        if ((used >> k) & 1ull) s_cols[__popcll(used & ((1ull << k) - 1ull))] = k;
    int lcol[BNPC_MAX_OPT];                                    // index of the option's column among the used ones
#pragma unroll
    for (int i = 0; i < BNPC_MAX_OPT; ++i) lcol[i] = (i < nn) ? __popcll(used & ((1ull << col[i]) - 1ull)) : -1;
    const uint4* p1 = reinterpret_cast<const uint4*>(x1 + (long long)v.cell * W);
    const uint4* p0 = reinterpret_cast<const uint4*>(x0 + (long long)v.cell * W);
    const int words = (M + 31) >> 5;
    for (int c0 = 0; c0 < n_used; c0 += EX_COLS) {
        const int nc = min(EX_COLS, n_used - c0);
        int off[BNPC_MAX_OPT];                                 // 2 * column slot in this pass, or the zero pair
#pragma unroll
        for (int i = 0; i < BNPC_MAX_OPT; ++i)
            off[i] = (lcol[i] >= c0 && lcol[i] < c0 + nc) ? 2 * (lcol[i] - c0) : 2 * EX_COLS;
        for (int w0 = 0; w0 < words; w0 += EX_WORDS) {
            __syncthreads();
            for (int i = threadIdx.x; i < 32 * EX_WORDS * nc; i += nthr) {
                const int kk = i / (32 * EX_WORDS), mm = i % (32 * EX_WORDS), m = w0 * 32 + mm;   // coalesced along mutations
                tile[mm * (EX_COLS + 1) + kk] = (m < M) ? lp[(long long)s_cols[c0 + kk] * M + m] : make_double2(0.0, 0.0);
            }
            __syncthreads();
            if (n_max == 0 || nn == 0) continue;
            const uint4 a = p1[w0 >> 2], b = p0[w0 >> 2];         // rows are padded with zeros to W words
            const uint32_t u1[EX_WORDS] = {a.x, a.y, a.z, a.w}, u0[EX_WORDS] = {b.x, b.y, b.z, b.w};
            switch (n_max) {                                      // warp-uniform
                case 1: ex_accumulate<1>(tile_d, u1, u0, off, part); break;
                case 2: ex_accumulate<2>(tile_d, u1, u0, off, part); break;
                case 3: ex_accumulate<3>(tile_d, u1, u0, off, part); break;
                case 4: ex_accumulate<4>(tile_d, u1, u0, off, part); break;
                case 5: case 6: ex_accumulate<6>(tile_d, u1, u0, off, part); break;
                default: ex_accumulate<BNPC_MAX_OPT>(tile_d, u1, u0, off, part); break;
            }
        }
    }
    double acc[BNPC_MAX_OPT];
#pragma unroll
    for (int i = 0; i < BNPC_MAX_OPT; ++i) acc[i] = (part[0][i] + part[1][i]) + (part[2][i] + part[3][i]);
    // (every thread stays for the block-wide publication of the option graph below)
    // same selection and weights as gibbs_candidates_kernel, on the exact values
    const double lnew_ll = v.lnew + c_norm;
    bnpc_cand_t out;
    double val[BNPC_MAX_OPT];
#pragma unroll
    for (int i = 0; i < BNPC_MAX_OPT; ++i) { out.e[i] = 0.0; out.col[i] = 0; val[i] = -BNPC_INF; }
    out.pad[0] = out.pad[1] = out.pad[2] = 0;
    int n = BNPC_MAX_OPT + 1, i_old = 0;
    double ref = 0.0, e_new = 0.0, e_max = 1.0;
    int c_old = -1;
    if (nn > 0) {
        double v_old = 0.0;
#pragma unroll
        for (int i = 0; i < BNPC_MAX_OPT; ++i)
            if (i == o.i_old) { v_old = acc[i]; c_old = col[i]; }
        const double thr = v_old - 40.0 - slack;
        n = 0;
        ref = fmax(v_old, lnew_ll);
#pragma unroll
        for (int i = 0; i < BNPC_MAX_OPT; ++i) {
            if (i < nn) {
                const bool own = (i == o.i_old);
                if (own || acc[i] > thr) {
                    if (own) i_old = n;
#pragma unroll
                    for (int s = 0; s < BNPC_MAX_OPT; ++s)
                        if (s == n) { val[s] = acc[i]; out.col[s] = (uint16_t)col[i]; }
                    ref = fmax(ref, acc[i]);
                    ++n;
                }
            }
        }
        e_max = 0.0;
#pragma unroll
        for (int i = 0; i < BNPC_MAX_OPT; ++i)
            if (i < n) { out.e[i] = exp(val[i] - ref); e_max = fmax(e_max, out.e[i]); }
        e_new = exp(lnew_ll - ref);
    } else if (live) {
        atomicAdd(&st[BNPC_ST_NMANY], 1);
    }
    if (live) {
        v.e_new = e_new;
        v.ref = ref;
        v.c_old = c_old;
        v.n_opt = n;
        v.i_old = i_old;
        v.flags = 0;
        v.e_max = __double2float_ru(e_max);
        visit_c[j] = v;
        cand_c[j] = out;
        // option graph on the columns: the visit links its own cluster with every rival
        if (nn > 0 && c_old >= 0) {
            unsigned long long mask = 0ull;
#pragma unroll
            for (int i = 0; i < BNPC_MAX_OPT; ++i)
                if (i < n) mask |= 1ull << out.col[i];
            atomicOr(&s_adj[c_old], mask);
            atomicAdd(&s_num[c_old], 1);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += nthr) {
        unsigned long long* adj = reinterpret_cast<unsigned long long*>(comp);
        if (s_adj[k]) atomicOr(&adj[k], s_adj[k]);
        if (s_num[k]) atomicAdd(&comp[128 + k], s_num[k]);
    }
}

// The same kernel staging ALL K columns per round (shared memory 2 KB x K per CTA): the variant
// of most round-1 measurements, kept selectable (BNPC_EXACT_STAGING=all) for comparison.
__device__ __forceinline__ void gibbs_exact_allcols_kernel(const uint32_t* __restrict__ x1, const uint32_t* __restrict__ x0, int W, int M,
                   const double2* __restrict__ lp, int K, const bnpc_visit_t* __restrict__ visit,
                   const bnpc_opt_t* __restrict__ opt, const int32_t* __restrict__ idx_c,
                   int32_t* __restrict__ st, bnpc_visit_t* __restrict__ visit_c,
                   bnpc_cand_t* __restrict__ cand_c, double slack, double c_norm, int32_t* __restrict__ comp,
                   const int32_t* __restrict__ order) {
    extern __shared__ __align__(16) unsigned char ex_smem[];
    __shared__ unsigned long long s_adj[BNPC_LEAN_MAXK];
    __shared__ int s_num[BNPC_LEAN_MAXK];
    double2* tile = reinterpret_cast<double2*>(ex_smem);          // [32 * EX_WORDS][K]
    const double* tile_d = reinterpret_cast<const double*>(ex_smem);
    const int n_unc = st[BNPC_ST_NUNC];
    if (blockIdx.x * EX_THREADS >= n_unc) return;
    if (threadIdx.x < BNPC_LEAN_MAXK) { s_adj[threadIdx.x] = 0ull; s_num[threadIdx.x] = 0; }
    const int q = blockIdx.x * EX_THREADS + threadIdx.x;
    const bool live = q < n_unc;
    const int j = live ? order[q] : order[0];        // slot of the visit among the compacted records
    const int r = idx_c[j];
    bnpc_visit_t v = visit[r];
    const bnpc_opt_t o = opt[r];
    const int nn = (live && o.n_opt <= BNPC_MAX_OPT) ? o.n_opt : 0;
    int col[BNPC_MAX_OPT];
    double part[EX_WORDS][BNPC_MAX_OPT];
#pragma unroll
    for (int i = 0; i < BNPC_MAX_OPT; ++i) {
        col[i] = (i < nn) ? o.col[i] : 0;
#pragma unroll
        for (int q = 0; q < EX_WORDS; ++q) part[q][i] = 0.0;
    }
    int n_max = nn;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) n_max = max(n_max, __shfl_xor_sync(FULL, n_max, s));
    const uint4* p1 = reinterpret_cast<const uint4*>(x1 + (long long)v.cell * W);
    const uint4* p0 = reinterpret_cast<const uint4*>(x0 + (long long)v.cell * W);
    const int words = (M + 31) >> 5;
    for (int w0 = 0; w0 < words; w0 += EX_WORDS) {
        __syncthreads();
        for (int i = threadIdx.x; i < 32 * EX_WORDS * K; i += EX_THREADS) {
            const int kk = i / (32 * EX_WORDS), mm = i % (32 * EX_WORDS), m = w0 * 32 + mm;   // coalesced along mutations
            tile[mm * K + kk] = (m < M) ? lp[(long long)kk * M + m] : make_double2(0.0, 0.0);
        }
        __syncthreads();
        if (n_max == 0 || nn == 0) continue;
        const uint4 a = p1[w0 >> 2], b = p0[w0 >> 2];         // rows are padded with zeros to W words
        const uint32_t u1[EX_WORDS] = {a.x, a.y, a.z, a.w}, u0[EX_WORDS] = {b.x, b.y, b.z, b.w};
#pragma unroll 2
        for (int bit = 0; bit < 32; ++bit) {
#pragma unroll
            for (int q = 0; q < EX_WORDS; ++q) {
                const uint32_t b1 = (u1[q] >> bit) & 1u, b0 = (u0[q] >> bit) & 1u;
                // the entry selects log p1 (.x), log p0 (.y) or nothing
                const double* t = tile_d + (q * 32 + bit) * 2 * K + (b1 ? 0 : 1);
                const bool any = (b1 | b0) != 0u;
#pragma unroll
                for (int i = 0; i < BNPC_MAX_OPT; ++i) {
                    if (i >= n_max) break;                               // warp-uniform
                    if (i < nn) {
                        const double term = t[2 * col[i]];
                        part[q][i] += any ? term : 0.0;
                    }
                }
            }
        }
    }
    double acc[BNPC_MAX_OPT];
#pragma unroll
    for (int i = 0; i < BNPC_MAX_OPT; ++i) acc[i] = (part[0][i] + part[1][i]) + (part[2][i] + part[3][i]);
    // (every thread stays for the block-wide publication of the option graph below)
    // same selection and weights as gibbs_candidates_kernel, on the exact values
    const double lnew_ll = v.lnew + c_norm;
    bnpc_cand_t out;
    double val[BNPC_MAX_OPT];
#pragma unroll
    for (int i = 0; i < BNPC_MAX_OPT; ++i) { out.e[i] = 0.0; out.col[i] = 0; val[i] = -BNPC_INF; }
    out.pad[0] = out.pad[1] = out.pad[2] = 0;
    int n = BNPC_MAX_OPT + 1, i_old = 0;
    double ref = 0.0, e_new = 0.0, e_max = 1.0;
    int c_old = -1;
    if (nn > 0) {
        double v_old = 0.0;
#pragma unroll
        for (int i = 0; i < BNPC_MAX_OPT; ++i)
            if (i == o.i_old) { v_old = acc[i]; c_old = col[i]; }
        const double thr = v_old - 40.0 - slack;
        n = 0;
        ref = fmax(v_old, lnew_ll);
#pragma unroll
        for (int i = 0; i < BNPC_MAX_OPT; ++i) {
            if (i < nn) {
                const bool own = (i == o.i_old);
                if (own || acc[i] > thr) {
                    if (own) i_old = n;
#pragma unroll
                    for (int s = 0; s < BNPC_MAX_OPT; ++s)
                        if (s == n) { val[s] = acc[i]; out.col[s] = (uint16_t)col[i]; }
                    ref = fmax(ref, acc[i]);
                    ++n;
                }
            }
        }
        e_max = 0.0;
#pragma unroll
        for (int i = 0; i < BNPC_MAX_OPT; ++i)
            if (i < n) { out.e[i] = exp(val[i] - ref); e_max = fmax(e_max, out.e[i]); }
        e_new = exp(lnew_ll - ref);
    } else if (live) {
        atomicAdd(&st[BNPC_ST_NMANY], 1);
    }
    if (live) {
        v.e_new = e_new;
        v.ref = ref;
        v.c_old = c_old;
        v.n_opt = n;
        v.i_old = i_old;
        v.flags = 0;
        v.e_max = __double2float_ru(e_max);
        visit_c[j] = v;
        cand_c[j] = out;
        // option graph on the columns: the visit links its own cluster with every rival
        if (nn > 0 && c_old >= 0) {
            unsigned long long mask = 0ull;
#pragma unroll
            for (int i = 0; i < BNPC_MAX_OPT; ++i)
                if (i < n) mask |= 1ull << out.col[i];
            atomicOr(&s_adj[c_old], mask);
            atomicAdd(&s_num[c_old], 1);
        }
    }
    __syncthreads();
    if (threadIdx.x < K) {
        unsigned long long* adj = reinterpret_cast<unsigned long long*>(comp);
        if (s_adj[threadIdx.x]) atomicOr(&adj[threadIdx.x], s_adj[threadIdx.x]);
        if (s_num[threadIdx.x]) atomicAdd(&comp[128 + threadIdx.x], s_num[threadIdx.x]);
    }
}


// comp layout (int32[256]): [0,128) adjacency masks (64 x uint64), [128,192) uncertain visits per
// column, [192,256) owner warp per column.  One block of 64 threads: connected components of the
// option graph, then the components are dealt to `n_warps` warps, heaviest first, each to the
// warp with the least records so far.
__device__ __forceinline__ void components_kernel(int32_t* __restrict__ comp, int K, int n_warps) {
    __shared__ unsigned long long m[BNPC_LEAN_MAXK];
    __shared__ int weight[BNPC_LEAN_MAXK], owner[BNPC_LEAN_MAXK], load[32];
    const int c = threadIdx.x;
    const unsigned long long* adj = reinterpret_cast<const unsigned long long*>(comp);
    m[c] = (c < K) ? (adj[c] | (1ull << c)) : (1ull << c);
    __syncthreads();
    // make the relation symmetric, then close it
    {
        unsigned long long x = m[c];
        while (x) { const int d = __ffsll((long long)x) - 1; x &= x - 1; atomicOr(&m[d], 1ull << c); }
    }
    __syncthreads();
    for (int it = 0; it < BNPC_LEAN_MAXK; ++it) {
        unsigned long long x = m[c], acc = m[c];
        while (x) { const int d = __ffsll((long long)x) - 1; x &= x - 1; acc |= m[d]; }
        const int changed = __syncthreads_or(acc != m[c]);
        m[c] = acc;
        __syncthreads();
        if (!changed) break;
    }
    const int root = __ffsll((long long)m[c]) - 1;
    int wsum = 0;
    if (root == c) {
        unsigned long long x = m[c];
        while (x) { const int d = __ffsll((long long)x) - 1; x &= x - 1; if (d < K) wsum += comp[128 + d]; }
    }
    weight[c] = (root == c) ? wsum : -1;
    owner[c] = 0;
    if (c < 32) load[c] = 0;
    __syncthreads();
    if (c == 0) {
        for (;;) {
            int best = -1, bw = 0;
            for (int r = 0; r < BNPC_LEAN_MAXK; ++r)
                if (weight[r] > bw) { bw = weight[r]; best = r; }
            if (best < 0) break;
            int wmin = 0;
            for (int q = 1; q < n_warps; ++q)
                if (load[q] < load[wmin]) wmin = q;
            owner[best] = wmin;
            load[wmin] += bw;
            weight[best] = -1;
        }
    }
    __syncthreads();
    comp[192 + c] = owner[root];
}

// Owner warp of every compacted record (0xff: more options than a record holds or unknown own
// column -- warp 0 posts it for the exact path), read by the parallel sequencer one byte per
// record instead of two fields of the 64-byte visit records.
__device__ __forceinline__ void owner_bytes_kernel(const bnpc_visit_t* __restrict__ visit_c, const int32_t* __restrict__ st,
                   const int32_t* __restrict__ comp, uint8_t* __restrict__ owner) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= st[BNPC_ST_NUNC]) return;
    const int n_opt = visit_c[j].n_opt, c_old = visit_c[j].c_old;
    const bool global = n_opt > BNPC_MAX_OPT || c_old < 0 || c_old >= BNPC_LEAN_MAXK;
    owner[j] = global ? (uint8_t)0xff : (uint8_t)comp[192 + c_old];
}

// Wide epochs: the dense FP64 matrix becomes option WEIGHTS in place, ll[t][k] <- exp(ll[t][k] - ref_t)
// with ref_t = max(max_k ll[t][k], new-cluster score), and the visit record gets ref, the weight
// of the new-cluster option and the column of the cell's own cluster.  One warp per visit.
__device__ __forceinline__ void gibbs_weights_kernel(double* __restrict__ ll, int ldk, int K, const int32_t* __restrict__ col_of_id,
                     bnpc_visit_t* __restrict__ visit, int C, double c_norm) {
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= C) return;
    double* row = ll + (long long)r * ldk;
    const double lnew_ll = visit[r].lnew + c_norm;
    double m = lnew_ll;
    for (int k = lane; k < K; k += 32) m = fmax(m, row[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    for (int k = lane; k < K; k += 32) row[k] = exp(row[k] - m);
    if (lane == 0) {
        visit[r].ref = m;
        visit[r].e_new = exp(lnew_ll - m);
        visit[r].c_old = col_of_id[visit[r].old];
        visit[r].n_opt = BNPC_MAX_OPT + 1;
        visit[r].flags = 0;
    }
}
