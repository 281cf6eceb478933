// Integer tensor-core rows (see bnpc_tc_i8.cuh) for SEVERAL CHAINS in one tile: the chains of a
// GPU sample the same data matrix, so when their epochs cover the same cells in the same order the
// expanded A operand (the 0/1 data in tensor memory) is the same for all of them and only the digit
// tables differ.  The tables are concatenated along the N dimension of the MMA,
//
//   D[128 cells x n_tot] += A[128 x 32] * [B_0; B_1; ...; B_{nc-1}]^T,   n_tot = sum_c 2*KPAD_c <= 256,
//
// so one pass of the producers and one TMEM read of A per MMA serve nc chains.  With N = 48..64
// columns (one chain) an MMA is bound by the 64 B/clk read of its 4 KB A operand (64 cycles for 24 to
// 32 cycles of tensor-pipe work); at N = 256 the tensor pipe (128 cycles) is the bound.
//   A   producers as in ll_matrix_i8_kernel (16 warps, 512 reduction indices per stage, two stages
//       in TMEM columns [256, 512)), for rows in CELL order only (row r = cell r: the first epoch of
//       a sweep writes its rows by cell and gibbs_options_kernel reads them by cell; the gathered
//       rows of later epochs stay with ll_matrix_i8_kernel).
//   B   a ring of CHUNK slots (128 reduction indices x n_tot rows = n_tot * 128 bytes), decoupled
//       from the A stages: a 256-column stage would not fit shared memory twice.  Chain c's rows sit
//       at row offset off_c (a multiple of 16, so its 8-row swizzle groups keep their 1024-byte
//       stride): one bulk copy per chain and chunk out of that chain's own table, issued by lane c
//       of the loader warp.  When all chunks of a row fit (n_tot * W * 64 bytes <= 200 KB) every
//       chunk has its own slot and the tables are loaded ONCE per CTA (resident): streaming them per
//       tile costs about 10 us per chain and launch at 100k x 1k (100 MB through L2 -> SM).
//   T   tiles per supertile (template parameter, 1 or 2): with T = 2 a stage of A holds two chunks of
//       each of two tiles (producer group g: tile g % 2, chunk g / 2), so a table chunk is fetched
//       and read from shared memory once for 256 cells; needs 2 * n_tot accumulator columns <= 256.
//   D   columns [0, T * n_tot) (+ a second set behind them when T * n_tot <= 128: the epilogue of a
//       supertile then overlaps the MMAs of the next one; with one set the MMAs wait for it).
// Warps: 0-15 producers, 16-19 epilogue, 20 MMA issuer, 21 B loader (704 threads: the register
// budget of the producers' loop, 80 per thread, is what fixes the warp count).
// Measured on B200 (tools/ll_shared_bench.py, 100k x 1k, K = 22..28): 27 us for one chain (the
// one-chain kernel: 28), 48 / 69 / 79 / 136 us for 2 / 3 / 4 / 8 chains, 38 us for two chains with
// resident tables; in the benchmark 73 us per launch of 5.15 chains = 0.51 of the sustained bf16 peak
// in algorithmic flops.

#define T8S_MAXC 8
#define T8S_EPI_WARPS 4
#define T8S_THREADS (32 * (T8_PWARPS + T8S_EPI_WARPS + 2))
#define T8S_MAX_BSLOTS 32
#define T8S_MAX_NTOT 256

struct ll_shared_chain_t {
    const uint8_t* Bg;      // digit table of the chain (lp_split_u8_kernel with its own KPAD)
    float* llf;             // rows of the chain
    float neg_q;
    int32_t ldf, kpad, off; // off: first column of the chain in the concatenated N dimension
};
struct ll_shared_t {
    const uint32_t* x1;
    const uint32_t* x0;
    int32_t W, C, nc, n_tot, n_ctas, b_slots, T, pad;   // T: tiles per supertile (1 or 2)
    ll_shared_chain_t ch[T8S_MAXC];
};

template <int T>
__device__ __forceinline__ void ll_matrix_i8s_kernel(const ll_shared_t& a) {
    const uint32_t* __restrict__ x1 = a.x1;
    const uint32_t* __restrict__ x0 = a.x0;
    const int W = a.W, C = a.C, n_ctas = a.n_ctas;
    const int n_tot = a.n_tot, b_slots = a.b_slots;
    const uint32_t chunk_bytes = (uint32_t)n_tot * 128u;
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    // barriers in the first 1024 bytes (compile-time addresses), the chunk ring behind them
    uint64_t* a_full = reinterpret_cast<uint64_t*>(tc_smem);
    unsigned char* const ring_smem = tc_smem + 1024;
    uint64_t* a_empty = a_full + T8_NST;
    uint64_t* b_full = a_empty + T8_NST;                    // [T8S_MAX_BSLOTS]
    uint64_t* b_empty = b_full + T8S_MAX_BSLOTS;            // [T8S_MAX_BSLOTS]
    uint64_t* acc_full = b_empty + T8S_MAX_BSLOTS;          // [2]
    uint64_t* acc_empty = acc_full + 2;                     // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half = W / 2;                                 // 64-bit pieces per plane = chunks per row
    // a CTA works on a SUPERTILE of T tiles at a time: a stage of A holds cps = 4 / T chunks of each of
    // the T tiles, so a chunk of the tables is read from shared memory (and fetched from L2) once
    // for T tiles
    constexpr int cps = (T8_PIECES / 2) / T;
    const int n_stages = (half + cps - 1) / cps;
    const int n_tiles = ((C + 127) / 128 + T - 1) / T;      // supertiles
    const int my_tiles = (n_tiles > (int)blockIdx.x) ? (n_tiles - 1 - (int)blockIdx.x) / n_ctas + 1 : 0;
    const int acc_cols = T * n_tot;                         // accumulator columns of a supertile
    const int n_sets = (acc_cols <= 128) ? 2 : 1;
    const bool resident = b_slots >= half;                  // the digit tables of all chunks fit shared memory

    if (threadIdx.x == 0) {
        for (int s = 0; s < T8_NST; ++s) { mbar_init(&a_full[s], T8_PWARPS); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < T8S_MAX_BSLOTS; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], T8S_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == T8_PWARPS + T8S_EPI_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)TC_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < T8_PWARPS) {
        // ---- A producers: the pipeline of ll_matrix_i8_kernel for rows in cell order (row r = cell r,
        // no index gather: the address of a piece follows from the tile and the stage).  Group g =
        // warp / 4 expands chunk g / T of every stage for tile g % T of the supertile ----
        const int g = warp >> 2;
        const int tg = g & (T - 1), cg = g / T;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const long long total = (long long)my_tiles * n_stages;
        const int row_in_super = tg * 128 + (warp & 3) * 32 + lane;
        long long pf_q = 0;
        int pf_sidx = 0;
        long long pf_row = (long long)blockIdx.x * (T * 128) + row_in_super;
        auto pf_next = [&](bool& ok) -> const uint4* {
            const int c = 2 * (pf_sidx * cps + cg);          // first of this thread's two pieces (= its chunk)
            ok = pf_q < total && pf_row < C && c < W;
            const int cc = (c < W) ? c : 0;
            const long long base = (pf_row < C ? pf_row : 0) * W;
            const uint32_t* src = (cc < half) ? x1 + base + 2 * cc : x0 + base + 2 * (cc - half);
            if (pf_q < total) {
                ++pf_q;
                if (++pf_sidx == n_stages) {
                    pf_sidx = 0;
                    pf_row += (long long)n_ctas * (T * 128);
                }
            }
            return reinterpret_cast<const uint4*>(src);
        };
        uint4 ring[T8_PF];
        bool ring_ok[T8_PF];
#pragma unroll
        for (int j = 0; j < T8_PF; ++j) ring[j] = __ldg(pf_next(ring_ok[j]));
        uint32_t it = 0;
        int pend = -1;
        for (long long q0 = 0; q0 < total; q0 += T8_PF) {
#pragma unroll
            for (int j = 0; j < T8_PF; ++j) {
                const uint4 raw = ring[j];
                const bool ok = ring_ok[j];
                ring[j] = __ldg(pf_next(ring_ok[j]));
                if (q0 + j < total) {
                    const uint4 w = ok ? raw : make_uint4(0u, 0u, 0u, 0u);
                    const int slot = it % T8_NST;
                    uint32_t regs[32];
#pragma unroll
                    for (int p = 0; p < 8; ++p) {
                        regs[p] = (w.x >> p) & 0x01010101u;
                        regs[8 + p] = (w.y >> p) & 0x01010101u;
                        regs[16 + p] = (w.z >> p) & 0x01010101u;
                        regs[24 + p] = (w.w >> p) & 0x01010101u;
                    }
                    if (pend >= 0) {
                        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&a_full[pend]);
                    }
                    if (it >= T8_NST) mbar_wait(&a_empty[slot], ((it / T8_NST) - 1) & 1);
                    tc_fence_after();
                    const uint32_t dst = tmem + T8_A_COL0 + slot * T8_A_STAGE_COLS + g * 32 + lane_base;
                    tc_st32(dst, regs);
                    pend = slot;
                    ++it;
                }
            }
        }
        if (pend >= 0) {
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[pend]);
        }
    } else if (warp < T8_PWARPS + T8S_EPI_WARPS) {
        // ---- epilogue: warp e reads TMEM lanes of quarter e % 4 and every second 8-column block ----
        const int e = warp - T8_PWARPS;
        const int row = (e & 3) * 32 + lane;
        const uint32_t lane_base = (uint32_t)((e & 3) * 32) << 16;
        uint32_t tile_count = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += n_ctas, ++tile_count) {
            const uint32_t set = tile_count % n_sets, use = tile_count / n_sets;
            mbar_wait(&acc_full[set], use & 1);
            tc_fence_after();
#pragma unroll
            for (int t = 0; t < T; ++t) {
            const long long r = ((long long)tile * T + t) * 128 + row;
            const bool live = r < C;
            const uint32_t d0 = tmem + lane_base + set * acc_cols + t * n_tot;
            for (int c = 0; c < a.nc; ++c) {
                const int kpad = a.ch[c].kpad, off = a.ch[c].off;
                const float neg_q = a.ch[c].neg_q;
                float4* dst = reinterpret_cast<float4*>(a.ch[c].llf + (live ? r : 0) * a.ch[c].ldf);
                // two 8-column blocks per round: four TMEM loads in flight behind one wait
                const int nb = kpad / 8;
                for (int j = 0; j < nb; j += 2) {
                    uint32_t hi[16], lo[16];
                    const bool two = j + 1 < nb;
                    tc_ld8(d0 + off + j * 8, hi);
                    tc_ld8(d0 + off + kpad + j * 8, lo);
                    if (two) {
                        tc_ld8(d0 + off + j * 8 + 8, hi + 8);
                        tc_ld8(d0 + off + kpad + j * 8 + 8, lo + 8);
                    }
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (live) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            if (h == 1 && !two) break;
                            float v[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] = neg_q * (float)(int)((hi[8 * h + i] << 8) + lo[8 * h + i]);
                            dst[2 * (j + h)] = make_float4(v[0], v[1], v[2], v[3]);
                            dst[2 * (j + h) + 1] = make_float4(v[4], v[5], v[6], v[7]);
                        }
                    }
                }
            }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[set]);
        }
    } else if (warp == T8_PWARPS + T8S_EPI_WARPS) {
        // ---- MMA issuer (warp-uniform control flow, one elected lane issues) ----
        const uint32_t tmem_u = __shfl_sync(FULL, tmem, 0);
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
        // instruction descriptor: D s32, A/B u8, both K-major, N = n_tot, M = 128
        const uint32_t idesc = (2u << 4) | ((uint32_t)(n_tot >> 3) << 17) | (8u << 24);
        const uint32_t smem_base = smem_u32(ring_smem);
        uint32_t it = 0, bi = 0, tile_count = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += n_ctas, ++tile_count) {
            const uint32_t set = tile_count % n_sets, use = tile_count / n_sets;
            if (use >= 1) mbar_wait(&acc_empty[set], (use - 1) & 1);
            tc_fence_after();
            for (int sidx = 0; sidx < n_stages; ++sidx, ++it) {
                const uint32_t slot = it % T8_NST;
                mbar_wait(&a_full[slot], (it / T8_NST) & 1);
                tc_fence_after();
                const uint32_t a_addr = tmem_u + T8_A_COL0 + slot * T8_A_STAGE_COLS;
                const int chunks = min(cps, half - sidx * cps);
                for (int cc = 0; cc < chunks; ++cc, ++bi) {
                    // resident tables: slot = chunk of the row, loaded once (phase 0 stays complete)
                    const uint32_t bslot = resident ? (uint32_t)(sidx * cps + cc) : bi % (uint32_t)b_slots;
                    mbar_wait(&b_full[bslot], resident ? 0u : (bi / (uint32_t)b_slots) & 1);
                    tc_fence_after();
                    if (leader) {
                        // K-major, 128B swizzle: LBO 1, SBO 1024 B, version 1, layout type 2
                        const uint32_t b_addr = smem_base + bslot * chunk_bytes;
                        const uint64_t desc0 = (uint64_t)((b_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) |
                                               (1ull << 46) | (2ull << 61);
#pragma unroll
                        for (int t = 0; t < T; ++t) {
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                tc_mma_ts_i8(tmem_u + set * acc_cols + t * n_tot, a_addr + (cc * T + t) * 32 + j * 8,
                                             desc0 + (uint64_t)(2 * j), idesc, (sidx | cc | j) != 0 ? 1u : 0u);
                        }
                        if (!resident) tc_commit(&b_empty[bslot]);
                    }
                    __syncwarp();
                }
                if (leader) tc_commit(&a_empty[slot]);
                __syncwarp();
            }
            if (leader) tc_commit(&acc_full[set]);
            __syncwarp();
        }
    } else {
        // ---- B loader: chunk ring; lane c copies chain c's rows of the chunk (source, size and
        // offset live in its registers), lane 0 waits for the slot and announces the bytes ----
        const int c = lane < a.nc ? lane : 0;
        const uint32_t my_bytes = (uint32_t)(2 * a.ch[c].kpad) * 128u;
        const unsigned char* my_src = a.ch[c].Bg;
        const uint32_t my_off = (uint32_t)a.ch[c].off * 128u;
        const bool mine = lane < a.nc;
        uint32_t bi = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += n_ctas) {
            if (resident && tile != (int)blockIdx.x) break;          // the whole table stays in its slots
            for (int kc = 0; kc < half; ++kc, ++bi) {
                const uint32_t bslot = bi % (uint32_t)b_slots;
                if (bi >= (uint32_t)b_slots) mbar_wait(&b_empty[bslot], ((bi / (uint32_t)b_slots) - 1) & 1);
                if (lane == 0) mbar_expect_tx(&b_full[bslot], chunk_bytes);
                __syncwarp();
                if (mine)
                    bulk_g2s(ring_smem + (size_t)bslot * chunk_bytes + my_off, my_src + (size_t)kc * my_bytes, my_bytes,
                             &b_full[bslot]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == T8_PWARPS + T8S_EPI_WARPS) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)TC_TMEM_COLS)
                     : "memory");
    }
}

// ---- host side ----------------------------------------------------------------------------------
// tiles per supertile of a group: two when their accumulators fit (half the table traffic per tile)
// and there are enough tiles to keep every SM busy; BNPC_LL_T=1|2 overrides (experiments)
static inline int ll_shared_sms() {
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}
static inline int ll_shared_T(const ll_shared_t& g) {
    static const int force_T = getenv("BNPC_LL_T") ? atoi(getenv("BNPC_LL_T")) : 0;
    int T = (2 * g.n_tot <= T8S_MAX_NTOT && cdiv(g.C, 128) >= 4 * ll_shared_sms()) ? 2 : 1;
    if (force_T == 1 || (force_T == 2 && 2 * g.n_tot <= T8S_MAX_NTOT)) T = force_T;
    return T;
}
static inline void ll_shared_finish(ll_shared_t& g, int T) {
    const int sms = ll_shared_sms();
    g.T = T;
    const int tiles = cdiv(cdiv(g.C, 128), g.T);
    g.n_ctas = tiles < sms ? tiles : sms;
    const size_t chunk = (size_t)g.n_tot * 128;
    int slots = (int)((200 * 1024) / chunk);
    slots = slots > T8S_MAX_BSLOTS ? T8S_MAX_BSLOTS : slots;
    // all chunks of a row fit: the tables stay resident, one slot per chunk
    g.b_slots = slots >= g.W / 2 ? g.W / 2 : slots;
}
static inline size_t ll_shared_smem(const ll_shared_t& g) { return (size_t)g.b_slots * g.n_tot * 128 + 1024; }

// true when chain `one` (a single-chain argument block) can share the tiles of group g
static inline bool ll_shared_fits(const ll_shared_t& g, const ll_shared_t& one) {
    return g.x1 == one.x1 && g.x0 == one.x0 && g.W == one.W &&
           g.C == one.C && g.nc < T8S_MAXC && g.n_tot + one.n_tot <= T8S_MAX_NTOT;
}
static inline void ll_shared_append(ll_shared_t& g, const ll_shared_t& one) {
    ll_shared_chain_t c = one.ch[0];
    c.off = g.n_tot;
    g.ch[g.nc++] = c;
    g.n_tot += one.n_tot;
}

// merged launch of the recorder: ops of chains that can share their tiles become one group each,
// the groups of a launch are its blockIdx.z
// Plan of one merged launch: the single-chain argument blocks `ones[0..n)` packed first-fit into
// groups (chains with the same planes and cell count, at most T8S_MAXC chains and T8S_MAX_NTOT
// columns per group), the tiles per supertile of the launch (two only when every group allows it:
// one kernel instance per launch), and per group the persistent CTAs and the chunk slots.  Host
// logic only (bnpc_ll_shared_plan exposes it to the CPU tests).
struct ll_plan_t {
    ll_shared_t groups[bnpc::BATCH_MAX];
    int ng, T;
    unsigned gx;
    size_t smem;
};
static void ll_shared_plan(const ll_shared_t* ones, int n, bool no_share, ll_plan_t& plan) {
    ll_shared_t* groups = plan.groups;
    int ng = 0;
    for (int i = 0; i < n; ++i) {
        const ll_shared_t& one = ones[i];
        int g = -1;
        for (int j = 0; j < ng && g < 0 && !no_share; ++j)
            if (ll_shared_fits(groups[j], one)) g = j;
        if (g < 0) groups[ng++] = one;
        else ll_shared_append(groups[g], one);
    }
    unsigned gx = 1;
    size_t smem = 0;
    int T = 2;
    for (int j = 0; j < ng; ++j) T = ll_shared_T(groups[j]) < T ? ll_shared_T(groups[j]) : T;
    for (int j = 0; j < ng; ++j) {
        ll_shared_finish(groups[j], T);
        gx = (unsigned)groups[j].n_ctas > gx ? (unsigned)groups[j].n_ctas : gx;
        const size_t sm = ll_shared_smem(groups[j]);
        smem = sm > smem ? sm : smem;
    }
    plan.ng = ng; plan.T = T; plan.gx = gx; plan.smem = smem;
}

static int launch_ll_shared_merged(bnpc::Op* const* ops, int n, cudaStream_t s) {
    using namespace bnpc;
    using P = typename FnTraits<decltype(&ll_matrix_i8s_kernel<1>)>::pack_t;
    static_assert(sizeof(P) <= ARG_BYTES, "ll_shared_t larger than ARG_BYTES");
    static const int no_share = getenv("BNPC_LL_NO_SHARE") ? 1 : 0;      // debugging: one chain per group
    if (n > BATCH_MAX) return bad_arg("more than BATCH_MAX chains in one merged launch");
    ll_shared_t ones[BATCH_MAX];
    for (int i = 0; i < n; ++i) memcpy(&ones[i], ops[i]->args, sizeof(ll_shared_t));
    ll_plan_t plan;
    ll_shared_plan(ones, n, no_share != 0, plan);
    const ll_shared_t* groups = plan.groups;
    const int ng = plan.ng, T = plan.T;
    const unsigned gx = plan.gx;
    const size_t smem = plan.smem;
    if (ng == 1) {
        Batch<1, P> B;
        memset(&B, 0, sizeof(B));
        B.gx[0] = gx; B.gy[0] = 1;
        memcpy(&B.a[0], &groups[0], sizeof(ll_shared_t));
        if (T == 2) return launch_wrapper<&ll_matrix_i8s_kernel<2>, 1, T8S_THREADS, 1>(ops[0]->name, B, dim3(gx, 1, 1), T8S_THREADS, smem, s, n);
        return launch_wrapper<&ll_matrix_i8s_kernel<1>, 1, T8S_THREADS, 1>(ops[0]->name, B, dim3(gx, 1, 1), T8S_THREADS, smem, s, n);
    }
    Batch<BATCH_MAX, P> B;
    memset(&B, 0, sizeof(B));
    for (int j = 0; j < ng; ++j) {
        B.gx[j] = (unsigned)groups[j].n_ctas; B.gy[j] = 1;
        memcpy(&B.a[j], &groups[j], sizeof(ll_shared_t));
    }
    if (T == 2) return launch_wrapper<&ll_matrix_i8s_kernel<2>, BATCH_MAX, T8S_THREADS, 1>(ops[0]->name, B, dim3(gx, 1, ng), T8S_THREADS, smem, s, n);
    return launch_wrapper<&ll_matrix_i8s_kernel<1>, BATCH_MAX, T8S_THREADS, 1>(ops[0]->name, B, dim3(gx, 1, ng), T8S_THREADS, smem, s, n);
}

// one chain's rows: launched at once, or recorded under the chain's slot so that the flush can
// put chains with equal tiles into one group
static int launch_ll_i8s(const uint32_t* x1, const uint32_t* x0, int W, int C,
                         const uint8_t* Bg, int kpad, float neg_q, float* llf, int ldf, cudaStream_t s) {
    using namespace bnpc;
    ll_shared_t one;
    memset(&one, 0, sizeof(one));
    one.x1 = x1; one.x0 = x0; one.W = W; one.C = C;
    one.nc = 1; one.n_tot = 2 * kpad;
    one.ch[0].Bg = Bg; one.ch[0].llf = llf; one.ch[0].neg_q = neg_q; one.ch[0].ldf = ldf; one.ch[0].kpad = kpad;
    one.ch[0].off = 0;
    if (g_rec.on) {
        g_rec.q[g_rec.cur].emplace_back();
        Op& op = g_rec.q[g_rec.cur].back();
        op.kind = 0; op.name = "ll_matrix_i8s_kernel"; op.merged = &launch_ll_shared_merged;
        op.gx = 1; op.gy = 1; op.block = T8S_THREADS; op.smem = 0;
        memcpy(op.args, &one, sizeof(one));
        return 0;
    }
    Op op;
    op.kind = 0; op.name = "ll_matrix_i8s_kernel"; op.block = T8S_THREADS;
    memcpy(op.args, &one, sizeof(one));
    Op* ops[1] = {&op};
    return launch_ll_shared_merged(ops, 1, s);
}
