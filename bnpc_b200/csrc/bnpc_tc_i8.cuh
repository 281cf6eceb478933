// Approximate cells x clusters log-likelihood rows on the tensor cores in EXACT INTEGER
// arithmetic (tcgen05.mma kind::i8, sm_100a): the production rows of lean Gibbs epochs.
//
//   llf[r][k] = sum_m x1[c][m] * LP1[k][m] + x0[c][m] * LP0[k][m],   c = cell of visit r
//
// The log-probabilities are quantised once per epoch to 16-bit fixed point,
//   -LP = q * (256 * hi + lo),   q = vmax / 65535,   hi, lo in [0, 255],
// and the two base-256 digits are separate output columns n = digit * KPAD + k, so that
//   D[128 visits x 2*KPAD] += A[128 x 32] * B[2*KPAD x 32]^T      (u8 x u8 -> s32, K = 32)
// accumulates integers with no rounding at all: the only error of a row is the quantisation,
// at most (number of observed entries of the cell) * q / 2  (<= M * q / 2; bnpc_gibbs_options
// gets it as err_abs), plus one float rounding of the result.
//   A  the 0/1 data, one byte per entry, expanded by the producer warps from the bit-planes
//      straight into TENSOR MEMORY (tcgen05.st; one TMEM lane per visit, four entries per 32-bit
//      column): (word >> p) & 0x01010101 makes four matrix elements.
//   B  the digit table in the K-major 128-byte-swizzled UMMA layout, one bulk async copy
//      (cp.async.bulk) per stage; a stage is 512 reduction indices = 16 MMAs.
//   D  int32 accumulators in TMEM, two sets: the epilogue of a tile overlaps the next tile.
// Warp roles (22 warps): 0-15 produce A (TMEM lane quarter = warp % 4; group warp / 4 expands two
// 64-bit pieces = 128 indices of every stage), 16-19 epilogue (TMEM -> registers -> one float row
// per visit), 20 issues the MMAs (warp-uniform, one elected lane), 21 streams B.  The 16-byte piece of a
// visit's row that a producer thread expands per stage is prefetched T8_PF stages ahead (the
// row gather is the long-latency part: visiting order is a random permutation of the cells),
// across tile boundaries, with the cell indices two tiles ahead of that.
// The order of the 32 reduction indices inside a word is a fixed permutation of the bit order
// (element 4p+b <-> bit p+8b), the same for A and B.

#define T8_NST 2                 /* pipeline stages (A in TMEM, B in shared memory) */
#define T8_GROUPS 4              /* producer groups: each expands two 64-bit pieces of every stage */
#define T8_PIECES 8              /* 64-bit pieces of a row per stage (512 reduction indices, 16 MMAs):
                                    tcgen05.st + wait::st + arrive cost ~550 cycles per warp and stage
                                    whatever the size, so a stage is as large as tensor memory allows */
#define T8_PWARPS (4 * T8_GROUPS)
#define T8_THREADS (32 * (T8_PWARPS + 6))   /* producers, 4 epilogue warps, MMA warp, B loader */
#define T8_A_COL0 256            /* first TMEM column of the A stages (accumulators use [0, 256)) */
#define T8_A_STAGE_COLS 128      /* 512 one-byte entries per visit and stage */
#define T8_PF 4                  /* row-piece prefetch depth, in stages */

// D[tmem_d] (+)= A[tmem_a] * B[smem desc]: kind::i8 (u8 x u8 -> s32), A from tensor memory
__device__ __forceinline__ void tc_mma_ts_i8(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
        "%15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

// Digit table: B chunk kc (128 reduction indices = two 64-bit pieces of one plane) x N rows x 128
// bytes, each chunk stored exactly as its shared-memory image (K-major, 128-byte swizzle: 8-row
// groups of 1024 bytes, the 16-byte piece c of row n at piece position c ^ (n & 7)).
__device__ __forceinline__ void lp_split_u8_kernel(const double2* __restrict__ lp, int K, int M, int W, int KPAD, double inv_q,
                                   uint8_t* __restrict__ Bg) {
    const int N = 2 * KPAD;
    const long long total = (long long)(W / 2) * N * 128;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int e = (int)(idx & 127);
    const int n = (int)((idx >> 7) % N);
    const int kc = (int)(idx / (128ll * N));
    const int digit = n / KPAD, k = n % KPAD;
    const int half = W / 2;                                  // 64-bit pieces per plane
    const int c0 = 2 * kc;                                   // first 64-bit piece of the chunk
    const int plane = c0 >= half;
    const int j = plane ? c0 - half : c0;
    const int q = e >> 5, p = (e & 31) >> 2, b = e & 3;      // element 32q + 4p + b <-> word q, bit p + 8b
    const int m = (2 * j + q) * 32 + p + 8 * b;
    int iv = 0;
    if (k < K && m < M) {
        const double2 t = lp[(long long)k * M + m];
        const double v = plane ? t.y : t.x;                  // <= 0
        iv = (int)rint(-v * inv_q);
        iv = iv < 0 ? 0 : (iv > 65535 ? 65535 : iv);
    }
    const int out = digit == 0 ? (iv >> 8) : (iv & 255);
    const long long off = (long long)kc * N * 128 + (n >> 3) * 1024 + (n & 7) * 128 + (((e >> 4) ^ (n & 7)) << 4) + (e & 15);
    Bg[off] = (uint8_t)out;
}

static long long* g_t8_trace = nullptr;      // debug hook (bnpc_debug_set_trace)

template <int KPAD>
__device__ __forceinline__ void ll_matrix_i8_kernel(const uint32_t* __restrict__ x1, const uint32_t* __restrict__ x0, int W,
                    const int32_t* __restrict__ cells, int cell_stride, int C,
                    const uint8_t* __restrict__ Bg, float neg_q, float* __restrict__ llf, int ldf,
                    long long* __restrict__ trace, int n_ctas) {
    // n_ctas: persistent CTAs of THIS chain's launch (a batched launch may hold more blocks)
    // trace (debug, normally NULL): clock64 stamps of CTA 0 -- [0,1024) producer warp 0 (3 per
    // stage: data expanded, previous store done + slot free, store issued), [1024,2048) MMA thread (4 per stage: stage full, first MMA issued, all issued, committed), [2048,..) epilogue warp 8 (2 per tile)
    const bool tr = trace != nullptr && blockIdx.x == 0 && (threadIdx.x & 31) == 0;
    constexpr int N = 2 * KPAD;
    constexpr uint32_t B_CHUNK_BYTES = (uint32_t)N * 128u;          // 128 reduction indices
    constexpr uint32_t B_STAGE_BYTES = (T8_PIECES / 2) * B_CHUNK_BYTES;
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(tc_smem + T8_NST * B_STAGE_BYTES);
    uint64_t* empty = full + T8_NST;
    uint64_t* acc_full = empty + T8_NST;                    // [2]
    uint64_t* acc_empty = acc_full + 2;                     // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // the row of a visit is 2 planes x W words = W pieces of 64 bits; a stage is 4 pieces
    const int half = W / 2;
    const int n_stages = (W + T8_PIECES - 1) / T8_PIECES;           // (W is a multiple of 4: the last stage may be half)
    const int n_tiles = (C + 127) / 128;
    const int my_tiles = (n_tiles > (int)blockIdx.x) ? (n_tiles - 1 - (int)blockIdx.x) / n_ctas + 1 : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < T8_NST; ++s) { mbar_init(&full[s], T8_PWARPS + 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == T8_PWARPS + 4) {
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
        // ---- A producers: one TMEM lane = one visit; group g = warp / 4 expands pieces 2g, 2g+1 of
        // every stage (one aligned 16-byte load, 32 TMEM columns) ----
        const int g = warp >> 2;
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const long long total = (long long)my_tiles * n_stages;
        // prefetch cursor: runs T8_PF stages ahead of the stage being expanded.  Every load is
        // issued unconditionally from a valid address (row 0 stands in for visits that do not
        // exist) straight into its ring register, and masked when it is consumed: a select or a
        // branch at the load would make the load's latency part of the iteration.
        long long pf_q = 0;
        int pf_sidx = 0, pf_tile = blockIdx.x;
        auto row_of_tile = [&](int tile, bool& ok) -> long long {
            const long long r = (long long)tile * 128 + row;
            ok = tile < n_tiles && r < C;
            return ok ? r : 0;
        };
        bool ok_cur, ok_n1, ok_n2;
        long long r_cur = row_of_tile(pf_tile, ok_cur);
        long long r_n1 = row_of_tile(pf_tile + n_ctas, ok_n1);
        long long r_n2 = row_of_tile(pf_tile + 2 * n_ctas, ok_n2);
        // cell of a row: loaded (or the row itself) -- always from a valid address
        int cell_cur = cells ? __ldg(cells + r_cur * cell_stride) : (int)r_cur;
        int cell_n1 = cells ? __ldg(cells + r_n1 * cell_stride) : (int)r_n1;
        int cell_n2 = cells ? __ldg(cells + r_n2 * cell_stride) : (int)r_n2;
        auto pf_next = [&](bool& ok) -> const uint4* {
            ok = ok_cur && pf_q < total && (pf_sidx * T8_PIECES + 2 * g) < W;    // (a half last stage: zeros)
            const int c = pf_sidx * T8_PIECES + 2 * g;       // first of this thread's two pieces of the stage
            const long long base = (long long)cell_cur * W;
            const int cc = (c < W) ? c : 0;
            const uint32_t* src = (cc < half) ? x1 + base + 2 * cc : x0 + base + 2 * (cc - half);
            if (pf_q < total) {
                ++pf_q;
                if (++pf_sidx == n_stages) {
                    pf_sidx = 0;
                    pf_tile += n_ctas;
                    cell_cur = cell_n1; ok_cur = ok_n1;
                    cell_n1 = cell_n2; ok_n1 = ok_n2;
                    r_n2 = row_of_tile(pf_tile + 2 * n_ctas, ok_n2);
                    cell_n2 = cells ? __ldg(cells + r_n2 * cell_stride) : (int)r_n2;
                }
            }
            return reinterpret_cast<const uint4*>(src);
        };
        uint4 ring[T8_PF];
        bool ring_ok[T8_PF];
#pragma unroll
        for (int j = 0; j < T8_PF; ++j) ring[j] = __ldg(pf_next(ring_ok[j]));
        // The store of stage s (tcgen05.st -> wait::st -> arrive) completes while the data of
        // stage s + 1 is expanded: `pend` is the slot whose store is still in flight.
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
                    if (tr && warp == 0 && it < 300) trace[it * 3] = clock64();
                    if (pend >= 0) {
                        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&full[pend]);
                    }
                    if (it >= T8_NST) mbar_wait(&empty[slot], ((it / T8_NST) - 1) & 1);
                    tc_fence_after();
                    if (tr && warp == 0 && it < 300) trace[it * 3 + 1] = clock64();
                    const uint32_t dst = tmem + T8_A_COL0 + slot * T8_A_STAGE_COLS + g * 32 + lane_base;
                    tc_st32(dst, regs);
                    pend = slot;
                    if (tr && warp == 0 && it < 300) trace[it * 3 + 2] = clock64();
                    ++it;
                }
            }
        }
        if (pend >= 0) {
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[pend]);
        }
    } else if (warp < T8_PWARPS + 4) {
        // ---- epilogue: D lane `row`, columns [0, N) of the tile's accumulator set ----
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        uint32_t tile_count = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += n_ctas, ++tile_count) {
            const uint32_t set = tile_count & 1u;
            mbar_wait(&acc_full[set], (tile_count >> 1) & 1);
            tc_fence_after();
            if (tr && warp == T8_PWARPS && tile_count < 100) trace[2048 + tile_count * 2] = clock64();
            const long long r = (long long)tile * 128 + row;
            float4* dst = reinterpret_cast<float4*>(llf + (r < C ? r : 0) * ldf);
#pragma unroll
            for (int c = 0; c < KPAD / 8; ++c) {
                uint32_t hi[8], lo[8];
                tc_ld8(tmem + lane_base + set * N + c * 8, hi);
                tc_ld8(tmem + lane_base + set * N + KPAD + c * 8, lo);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = neg_q * (float)(int)((hi[i] << 8) + lo[i]);
                if (r < C) {
                    dst[2 * c] = make_float4(v[0], v[1], v[2], v[3]);
                    dst[2 * c + 1] = make_float4(v[4], v[5], v[6], v[7]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[set]);
            if (tr && warp == T8_PWARPS && tile_count < 100) trace[2048 + tile_count * 2 + 1] = clock64();
        }
    } else if (warp == T8_PWARPS + 4) {
        // ---- MMA issuer: the whole warp walks the loops (warp-uniform control flow and operands,
        // so the descriptors live in uniform registers and an MMA is one instruction instead of a
        // register-to-uniform waterfall); one elected lane issues ----
        const uint32_t tmem_u = __shfl_sync(FULL, tmem, 0);
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
        // instruction descriptor: D s32, A/B u8, both K-major, N, M = 128
        const uint32_t idesc = (2u << 4) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        const uint32_t smem_base = smem_u32(tc_smem);
        uint32_t it = 0, tile_count = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += n_ctas, ++tile_count) {
            const uint32_t set = tile_count & 1u;
            if (tile_count >= 2) mbar_wait(&acc_empty[set], ((tile_count >> 1) - 1) & 1);
            tc_fence_after();
            for (int sidx = 0; sidx < n_stages; ++sidx, ++it) {
                const uint32_t slot = it % T8_NST;
                mbar_wait(&full[slot], (it / T8_NST) & 1);
                tc_fence_after();
                if (tr && it < 250) trace[1024 + it * 4] = clock64();
                const uint32_t b_addr = smem_base + slot * B_STAGE_BYTES;
                const uint32_t a_addr = tmem_u + T8_A_COL0 + slot * T8_A_STAGE_COLS;
                const int chunks = min(T8_PIECES / 2, W / 2 - sidx * (T8_PIECES / 2));
                if (leader) {
#pragma unroll
                    for (int cc = 0; cc < T8_PIECES / 2; ++cc) {
                        if (cc >= chunks) break;
                        // K-major, 128B swizzle: LBO 1, SBO 1024 B, version 1, layout type 2
                        const uint64_t desc0 = (uint64_t)(((b_addr + cc * B_CHUNK_BYTES) >> 4) & 0x3FFFu) |
                                               (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            tc_mma_ts_i8(tmem_u + set * N, a_addr + cc * 32 + j * 8, desc0 + (uint64_t)(2 * j), idesc,
                                         (sidx | cc | j) != 0 ? 1u : 0u);
                            if (tr && it < 250 && cc == 0 && j == 0) trace[1024 + it * 4 + 1] = clock64();
                        }
                    }
                    if (tr && it < 250) trace[1024 + it * 4 + 2] = clock64();
                    tc_commit(&empty[slot]);
                    if (tr && it < 250) trace[1024 + it * 4 + 3] = clock64();
                }
                __syncwarp();
            }
            if (leader) tc_commit(&acc_full[set]);
            __syncwarp();
        }
    } else {
        // ---- B loader ----
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += n_ctas) {
                for (int sidx = 0; sidx < n_stages; ++sidx, ++it) {
                    const int slot = it % T8_NST;
                    if (it >= T8_NST) mbar_wait(&empty[slot], ((it / T8_NST) - 1) & 1);
                    const uint32_t bytes = (uint32_t)min(T8_PIECES / 2, W / 2 - sidx * (T8_PIECES / 2)) * B_CHUNK_BYTES;
                    mbar_expect_tx(&full[slot], bytes);
                    bulk_g2s(tc_smem + slot * B_STAGE_BYTES, Bg + (long long)sidx * B_STAGE_BYTES, bytes, &full[slot]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == T8_PWARPS + 4) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)TC_TMEM_COLS)
                     : "memory");
    }
}

template <int KPAD>
static int launch_ll_i8(const uint32_t* x1, const uint32_t* x0, int W, const int32_t* cells, int cell_stride,
                        int C, const uint8_t* Bg, float neg_q, float* llf, int ldf, cudaStream_t s) {
    const size_t smem = (size_t)T8_NST * (T8_PIECES / 2) * (2 * KPAD) * 128 + 256;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int tiles = cdiv(C, 128);
    BNPC_LAUNCH(ll_matrix_i8_kernel<KPAD>, T8_THREADS, 1, tiles < sms ? tiles : sms, T8_THREADS, smem, s, x1, x0, W, cells, cell_stride, C, Bg, neg_q, llf, ldf, g_t8_trace, tiles < sms ? tiles : sms);
    return 0;
}
