// Approximate cells x clusters log-likelihood rows on the 5th-generation tensor cores
// (tcgen05.mma, sm_100a), used by lean Gibbs epochs to select each visit's options.
//
//   llf[r][k] = sum_m x1[c][m] * LP1[k][m] + x0[c][m] * LP0[k][m],   c = cell of visit r
//
// as one GEMM  D[128 visits x N] += A[128 x 64] * B[N x 64]^T  per 64-bit slice of a bit-plane:
//   A  the 0/1 data of the slice, expanded by the producer warps from the bit-planes straight
//      into TENSOR MEMORY (tcgen05.st; one TMEM lane per visit, two bf16 per 32-bit column).
//      A set bit becomes bf16 2.0 = 0x4000 -- a single bit, so a rotate and an AND expand two
//      matrix elements -- and B carries the factor 0.5 (exact).
//   B  the (log p1 | log p0) table split into S = 2 bf16 terms (hi, lo: 16 significant bits) as
//      separate output columns n = s*KPAD + k, prepared once per epoch in the K-major
//      128-byte-swizzled UMMA layout so that a stage is one bulk async copy (cp.async.bulk).
//   D  FP32 accumulators in TMEM; the epilogue adds the S column groups and writes floats.
// Warp roles: warps 0-7 produce A (TMEM lane quarter = warp % 4; warps 0-3 expand the first half
// of every stage and run the epilogue, warps 4-7 the second half), warp 8 issues the MMAs (one
// thread), warp 9 streams B.  A stage is 256 reduction indices (16 MMAs) so that the barrier round
// trip of a stage is amortised; 2-stage pipeline on mbarriers; tcgen05.commit frees a stage.
// Measured (B200, 100k x 1k x 24): 56 us; ~90 cycles per MMA whatever N is -- with N = 2*Kp <= 128
// output columns the instruction is bound by reading its 4 KB A operand from tensor memory, not
// by the tensor pipe (62 cycles per MMA at N = 128).  The production rows are the integer kernel of
// bnpc_tc_i8.cuh (8-bit A: half the instructions per tile, exact accumulation); this one stays as an
// independent route (lean = 2) for the parity tests.  The order of the 64 reduction indices inside
// a stage is a fixed permutation of the bit order (element 2p+h <-> bit p+16h), the same for A and B.
#include <cuda_bf16.h>

#define TC_NST 2
#define TC_CPS 4              /* 64-bit chunks of the row per pipeline stage */
#define TC_THREADS 320
#define TC_A_COL0 256            /* first TMEM column of the A stages (accumulators use [0, 256)) */
#define TC_TMEM_COLS 512

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    long long spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1ll << 26)) asm volatile("trap;");      // a protocol bug must not hang the GPU
    }
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem_d] (+)= A[tmem_a] * B[smem desc]: kind::f16 (bf16 x bf16 -> f32), A from tensor memory
__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
        "%15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
        "%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}

// B table: chunk kc (64 reduction indices of one plane) x N rows x 64 bf16, each chunk stored
// exactly as its shared-memory stage (K-major, 128-byte swizzle: 8-row groups of 1024 bytes, the
// 16-byte piece c of row n at piece position c ^ (n & 7)).
__device__ __forceinline__ void lp_split_bf16_kernel(const double2* __restrict__ lp, int K, int M, int W, int KPAD,
                                     uint16_t* __restrict__ Bg) {
    const int N = 2 * KPAD;
    const long long total = (long long)W * N * 64;          // W chunks: W/2 per plane
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int e = (int)(idx & 63);
    const int n = (int)((idx >> 6) % N);
    const int kc = (int)(idx / (64ll * N));
    const int s = n / KPAD, k = n % KPAD;
    const int plane = kc >= W / 2;
    const int j = plane ? kc - W / 2 : kc;
    const int word = e >> 5, q = e & 31, p = q >> 1, h = q & 1;
    const int m = (2 * j + word) * 32 + p + 16 * h;
    double v = 0.0;
    if (k < K && m < M) {
        const double2 t = lp[(long long)k * M + m];
        v = 0.5 * (plane ? t.y : t.x);
    }
    const __nv_bfloat16 hi = __float2bfloat16_rn((float)v);
    const double rest = v - (double)__bfloat162float(hi);
    const __nv_bfloat16 lo = __float2bfloat16_rn((float)rest);
    const __nv_bfloat16 out = s == 0 ? hi : lo;
    const long long off = (long long)kc * N * 64 + (n >> 3) * 512 + (n & 7) * 64 + (((e >> 3) ^ (n & 7)) << 3) + (e & 7);
    Bg[off] = __bfloat16_as_ushort(out);
}

template <int KPAD>
__device__ __forceinline__ void ll_matrix_tc_kernel(const uint32_t* __restrict__ x1, const uint32_t* __restrict__ x0, int W,
                    const int32_t* __restrict__ cells, int cell_stride, int C,
                    const uint16_t* __restrict__ Bg, float* __restrict__ llf, int ldf, int n_ctas) {
    constexpr int N = 2 * KPAD;
    // consecutive MMAs go to NACC independent accumulator tiles, summed in the epilogue (no
    // read-after-write chain on one TMEM tile between back-to-back instructions)
    constexpr int NACC = (N <= 64) ? 4 : 2;
    constexpr uint32_t B_CHUNK_BYTES = (uint32_t)N * 128u;          // 64 reduction indices
    constexpr uint32_t B_STAGE_BYTES = TC_CPS * B_CHUNK_BYTES;
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(tc_smem + TC_NST * B_STAGE_BYTES);
    uint64_t* empty = full + TC_NST;
    uint64_t* acc_full = empty + TC_NST;
    uint64_t* acc_empty = acc_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // the row of a visit is 2 planes x W words = W chunks of 64 bits; a stage is TC_CPS chunks
    const int half = W / 2;                                 // chunks per plane
    const int n_stages = (W + TC_CPS - 1) / TC_CPS;
    const int n_tiles = (C + 127) / 128;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_NST; ++s) { mbar_init(&full[s], 9); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)TC_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < 8) {
        // ---- A producers: one TMEM lane = one visit; group g = warp / 4 expands chunks
        // [2g, 2g+2) of every stage; group 0 also runs the epilogue of the tile ----
        const int g = warp >> 2;
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        uint32_t it = 0, tile_count = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += n_ctas, ++tile_count) {
            const int r = tile * 128 + row;
            const bool live = r < C;
            const long long cell = cells ? cells[(long long)(live ? r : 0) * cell_stride] : (live ? r : 0);
            const uint32_t* p1 = x1 + cell * W;
            const uint32_t* p0 = x0 + cell * W;
            // words of chunk c (64 bits of plane 1, then of plane 0); zero beyond the row
            auto fetch = [&](int c) -> uint2 {
                if (!live || c >= W) return make_uint2(0u, 0u);
                const uint32_t* src = (c < half) ? p1 + 2 * c : p0 + 2 * (c - half);
                return __ldg(reinterpret_cast<const uint2*>(src));
            };
            // this group's chunks of the next stage are in flight while the current ones are expanded
            uint2 nxt0 = fetch(2 * g), nxt1 = fetch(2 * g + 1);
            for (int sidx = 0; sidx < n_stages; ++sidx, ++it) {
                const uint2 w0 = nxt0, w1 = nxt1;
                nxt0 = fetch((sidx + 1) * TC_CPS + 2 * g);
                nxt1 = fetch((sidx + 1) * TC_CPS + 2 * g + 1);
                const int slot = it % TC_NST;
                uint32_t regs[64];
#pragma unroll
                for (int p = 0; p < 16; ++p) {
                    regs[p] = __funnelshift_l(w0.x, w0.x, (14 - p) & 31) & 0x40004000u;
                    regs[16 + p] = __funnelshift_l(w0.y, w0.y, (14 - p) & 31) & 0x40004000u;
                    regs[32 + p] = __funnelshift_l(w1.x, w1.x, (14 - p) & 31) & 0x40004000u;
                    regs[48 + p] = __funnelshift_l(w1.y, w1.y, (14 - p) & 31) & 0x40004000u;
                }
                if (it >= TC_NST) mbar_wait(&empty[slot], ((it / TC_NST) - 1) & 1);
                tc_fence_after();
                const uint32_t dst = tmem + TC_A_COL0 + slot * (TC_CPS * 32) + g * 64 + lane_base;
                tc_st32(dst, regs);
                tc_st32(dst + 32, regs + 32);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[slot]);
            }
            if (g != 0) continue;
            // epilogue: D lane `row`, columns [0, N): add the split groups, write the row
            mbar_wait(acc_full, tile_count & 1);
            tc_fence_after();
            float acc[KPAD];
#pragma unroll
            for (int c = 0; c < KPAD / 8; ++c) {
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[c * 8 + i] = 0.0f;
#pragma unroll
                for (int a = 0; a < NACC; ++a) {
                    uint32_t v0[8], v1[8];
                    tc_ld8(tmem + lane_base + a * N + c * 8, v0);
                    tc_ld8(tmem + lane_base + a * N + KPAD + c * 8, v1);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[c * 8 + i] += __uint_as_float(v0[i]) + __uint_as_float(v1[i]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
            if (live) {
                float4* dst = reinterpret_cast<float4*>(llf + (long long)r * ldf);
#pragma unroll
                for (int i = 0; i < KPAD / 4; ++i)
                    dst[i] = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
            }
        }
    } else if (warp == 8) {
        // ---- MMA issuer ----
        if (lane == 0) {
            // instruction descriptor: D f32, A/B bf16, both K-major, N, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
            uint32_t it = 0, tile_count = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += n_ctas, ++tile_count) {
                if (tile_count > 0) mbar_wait(acc_empty, (tile_count - 1) & 1);
                tc_fence_after();
                for (int sidx = 0; sidx < n_stages; ++sidx, ++it) {
                    const int slot = it % TC_NST;
                    mbar_wait(&full[slot], (it / TC_NST) & 1);
                    tc_fence_after();
                    const uint32_t b_addr = smem_u32(tc_smem + slot * B_STAGE_BYTES);
                    const uint32_t a_addr = tmem + TC_A_COL0 + slot * (TC_CPS * 32);
                    const int chunks = min(TC_CPS, W - sidx * TC_CPS);
                    for (int cc = 0; cc < chunks; ++cc) {
                        // K-major, 128B swizzle: LBO 1, SBO 1024 B, version 1, layout type 2
                        const uint64_t desc0 = (uint64_t)(((b_addr + cc * B_CHUNK_BYTES) >> 4) & 0x3FFFu) |
                                               (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            tc_mma_ts(tmem + (j % NACC) * N, a_addr + cc * 32 + j * 8, desc0 + (uint64_t)(2 * j), idesc,
                                      ((sidx | cc) != 0 || j >= NACC) ? 1u : 0u);
                    }
                    tc_commit(&empty[slot]);
                }
                tc_commit(acc_full);
            }
        }
    } else {
        // ---- B loader ----
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += n_ctas) {
                for (int sidx = 0; sidx < n_stages; ++sidx, ++it) {
                    const int slot = it % TC_NST;
                    if (it >= TC_NST) mbar_wait(&empty[slot], ((it / TC_NST) - 1) & 1);
                    const uint32_t bytes = (uint32_t)min(TC_CPS, W - sidx * TC_CPS) * B_CHUNK_BYTES;
                    mbar_expect_tx(&full[slot], bytes);
                    bulk_g2s(tc_smem + slot * B_STAGE_BYTES, Bg + (long long)sidx * TC_CPS * N * 64, bytes,
                             &full[slot]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)TC_TMEM_COLS)
                     : "memory");
    }
}

template <int KPAD>
static int launch_ll_tc(const uint32_t* x1, const uint32_t* x0, int W, const int32_t* cells, int cell_stride,
                        int C, const uint16_t* Bg, float* llf, int ldf, cudaStream_t s) {
    const size_t smem = (size_t)TC_NST * TC_CPS * (2 * KPAD) * 128 + 256;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int tiles = cdiv(C, 128);
    BNPC_LAUNCH(ll_matrix_tc_kernel<KPAD>, TC_THREADS, 1, tiles < sms ? tiles : sms, TC_THREADS, smem, s, x1, x0, W, cells, cell_stride, C, Bg, llf, ldf, tiles < sms ? tiles : sms);
    return 0;
}

