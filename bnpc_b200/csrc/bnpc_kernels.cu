// BnpC MCMC hot path -- sm_100a kernels and their C ABI (include/bnpc_b200.h).
//
// Kernel inventory (reference function each one stands in for is cited in the header):
//   pack_planes_kernel      float64/int8 matrix -> two cell-major bit-planes + popcounts
//   fill_uniform_kernel     Philox4x32-10 uniforms / small integers
//   fill_permutation_kernel keyed Feistel permutation with cycle walking
//   logprob_tables_kernel   theta -> (log p1, log p0) per (cluster, mutation)
//   ll_matrix_kernel        cells x clusters log-likelihood, FP64 FMA over bit-planes
//   gibbs_prepare_kernel    per-visit records of a sweep
//   gibbs_epoch_begin_kernel
//   gibbs_sweep_kernel      the sequential Gibbs sweep: one persistent CTA per chain,
//                           warp-level categorical sampling for K<=31 (ll rows and visit
//                           records staged into shared memory by bulk async copies),
//                           CTA-wide sampling for larger K, cluster births/deaths on device
//   group_members_kernel / suffstat_kernel   per-cluster sufficient statistics
//   beta_rows_kernel / theta_from_uniform_kernel
//   mh_theta_kernel / theta_log_ratio_kernel   Metropolis-Hastings on theta
//   row_loglik_kernel / row_sum_kernel         deterministic [R][M] reductions
//   gather_* / anchor_swaps / rg_*             split-merge restricted Gibbs support
#include <cuda_runtime.h>
#include <atomic>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/bnpc_b200.h"
#include "../../include/bnpc_b200_debug.h"
#include "bnpc_math.cuh"

using namespace bnpc;

static thread_local char g_err[512] = "";

static int fail(const char* what, cudaError_t e) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return 1;
}
static int bad_arg(const char* what) {
    snprintf(g_err, sizeof(g_err), "invalid argument: %s", what);
    return 2;
}
#include "bnpc_batch.cuh"

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// =============================================================================
// recordable stream operations (see bnpc_batch.cuh)
// =============================================================================
// Zero-fills and small copies as kernel bodies, so that the many clears / staging copies of a step
// merge across chains like every other launch (cudaMemsetAsync / cudaMemcpyAsync nodes cannot be
// batched).  Up to SEG_MAX segments per launch (blockIdx.y = segment): consecutive recorded
// clears (or copies) of one chain become ONE operation.
#define SEG_MAX 4
struct SegArgs {
    void* dst[SEG_MAX];
    const void* src[SEG_MAX];
    long long n_words[SEG_MAX];
    int vec[SEG_MAX];
    int count;
};
__device__ __forceinline__ void seg_zero_kernel(SegArgs a) {
    const int sg = blockIdx.y;
    if (sg >= a.count) return;
    uint32_t* p = reinterpret_cast<uint32_t*>(a.dst[sg]);
    const long long n = a.n_words[sg];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (!a.vec[sg]) {
        if (i < n) p[i] = 0u;
        return;
    }
    const long long n4 = n >> 2;
    if (i < n4) reinterpret_cast<uint4*>(p)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (i < (n & 3)) p[4 * n4 + i] = 0u;
}
// copies between device memory and PINNED host memory (reachable from the device under unified
// addressing): the live list in, the status block out; device-to-device snapshots
__device__ __forceinline__ void seg_copy_kernel(SegArgs a) {
    const int sg = blockIdx.y;
    if (sg >= a.count) return;
    uint32_t* d = reinterpret_cast<uint32_t*>(a.dst[sg]);
    const uint32_t* s = reinterpret_cast<const uint32_t*>(a.src[sg]);
    const long long n = a.n_words[sg];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (!a.vec[sg]) {
        if (i < n) d[i] = s[i];
        return;
    }
    const long long n4 = n >> 2;
    if (i < n4) reinterpret_cast<uint4*>(d)[i] = reinterpret_cast<const uint4*>(s)[i];
    if (i < (n & 3)) d[4 * n4 + i] = s[4 * n4 + i];
}

template <auto Body>
static int seg_launch(const char* name, void* dst, const void* src, long long words, int vec, void* stream) {
    using namespace bnpc;
    const unsigned blocks = (unsigned)cdiv(vec ? (words + 3) / 4 : words, 256);
    if (g_rec.on && !g_rec.q[g_rec.cur].empty()) {
        Op& last = g_rec.q[g_rec.cur].back();
        if (last.kind == 0 && last.merged == &launch_merged<Body, 0, 0>) {
            SegArgs* a = reinterpret_cast<SegArgs*>(last.args);     // Pack<SegArgs>: head at offset 0
            if (a->count < SEG_MAX) {
                const int k = a->count++;
                a->dst[k] = dst; a->src[k] = src; a->n_words[k] = words; a->vec[k] = vec;
                last.gx = blocks > last.gx ? blocks : last.gx;
                last.gy = (unsigned)a->count;
                return 0;
            }
        }
    }
    SegArgs a;
    memset(&a, 0, sizeof(a));
    a.dst[0] = dst; a.src[0] = src; a.n_words[0] = words; a.vec[0] = vec; a.count = 1;
    return bnpc::launch<Body, 0, 0>(name, dim3(blocks, 1), 256, 0, (cudaStream_t)stream, a);
}

static int zero_async(void* ptr, size_t bytes, void* stream, const char* what) {
    if (bytes == 0) return 0;
    if ((bytes & 3) || ((uintptr_t)ptr & 3)) {
        if (bnpc::g_rec.on) return bad_arg("recorded clears must be whole aligned words");
        cudaError_t e = cudaMemsetAsync(ptr, 0, bytes, (cudaStream_t)stream);
        if (e != cudaSuccess) return fail(what, e);
        return 0;
    }
    return seg_launch<seg_zero_kernel>("seg_zero_kernel", ptr, nullptr, (long long)(bytes >> 2), ((uintptr_t)ptr & 15) ? 0 : 1, stream);
}

#define BNPC_SMALL_COPY_BYTES 8192
static int g_uva_copies = -1;      // small host<->device copies as kernels over mapped pinned memory
static int copy_async(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, void* stream) {
    if (bytes == 0) return 0;
    if (g_uva_copies < 0) {
        const char* env = getenv("BNPC_UVA_COPIES");
        g_uva_copies = (env && env[0] == '0') ? 0 : 1;
    }
    // recorded: device-to-device copies of any size and small copies from / to pinned host memory
    // become kernel launches (they merge across chains); the rest stays a copy-engine operation
    if (bnpc::g_rec.on && !(bytes & 3) && !(((uintptr_t)dst | (uintptr_t)src) & 3) &&
        (kind == cudaMemcpyDeviceToDevice || (g_uva_copies && bytes <= BNPC_SMALL_COPY_BYTES))) {
        const int vec = (((uintptr_t)dst | (uintptr_t)src) & 15) ? 0 : 1;
        return seg_launch<seg_copy_kernel>("seg_copy_kernel", dst, src, (long long)(bytes >> 2), vec, stream);
    }
    if (bnpc::g_rec.on) {
        bnpc::g_rec.q[bnpc::g_rec.cur].emplace_back();
        bnpc::Op& op = bnpc::g_rec.q[bnpc::g_rec.cur].back();
        op.kind = 1; op.name = "memcpy"; op.merged = nullptr;
        op.dst = dst; op.src = src; op.bytes = bytes; op.mk = kind;
        return 0;
    }
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, kind, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail("cudaMemcpyAsync", e);
    return 0;
}

static int record_event(void* ev, void* stream) {
    if (!ev) return 0;
    if (bnpc::g_rec.on) {
        bnpc::g_rec.q[bnpc::g_rec.cur].emplace_back();
        bnpc::Op& op = bnpc::g_rec.q[bnpc::g_rec.cur].back();
        op.kind = 2; op.name = "event"; op.merged = nullptr; op.dst = ev;
        return 0;
    }
    cudaError_t e = cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail("cudaEventRecord", e);
    return 0;
}

// Issue the recorded operations of all chain slots on `s`: per chain in recorded order, across
// chains merged -- the slot with the most operations left leads (a chain that recorded an extra
// operation catches up alone), every slot whose next operation is the same kernel with the same
// block size joins its launch.
static int recorder_flush(cudaStream_t s, unsigned long long slots = ~0ull) {
    using namespace bnpc;
    Recorder& R = g_rec;
    size_t cur[GROUP_MAX];
    // slots outside the mask keep their queues (they go to another stream)
    for (int c = 0; c < GROUP_MAX; ++c) cur[c] = ((slots >> c) & 1ull) ? 0 : R.q[c].size();
    int rc = 0;
    for (;;) {
        int lead = -1;
        size_t most = 0;
        for (int c = 0; c < GROUP_MAX; ++c) {
            const size_t left = R.q[c].size() - cur[c];
            if (left > most) { most = left; lead = c; }
        }
        if (lead < 0) break;
        Op& key = R.q[lead][cur[lead]];
        if (key.kind == 1) {
            cudaError_t e = cudaMemcpyAsync(key.dst, key.src, key.bytes, key.mk, s);
            if (e != cudaSuccess) { rc = fail("cudaMemcpyAsync", e); break; }
            ++cur[lead];
            continue;
        }
        if (key.kind == 2) {
            cudaError_t e = cudaEventRecord((cudaEvent_t)key.dst, s);
            if (e != cudaSuccess) { rc = fail("cudaEventRecord", e); break; }
            ++cur[lead];
            continue;
        }
        Op* ops[BATCH_MAX];
        int who[BATCH_MAX];
        int n = 0;
        ops[n] = &key; who[n++] = lead;
        // debugging: BNPC_NO_MERGE=1 one chain per launch; BNPC_MERGE_ONLY=a,b,... merges only the
        // kernels whose name contains one of the tokens
        static const int no_merge = getenv("BNPC_NO_MERGE") ? 1 : 0;
        static const char* only = getenv("BNPC_MERGE_ONLY");
        bool mergeable = !no_merge;
        if (mergeable && only) {
            mergeable = false;
            char buf[512];
            snprintf(buf, sizeof(buf), "%s", only);
            for (char* tok = strtok(buf, ","); tok; tok = strtok(nullptr, ","))
                if (strstr(key.name, tok)) mergeable = true;
        }
        for (int c = 0; c < GROUP_MAX && n < BATCH_MAX && mergeable; ++c) {
            if (c == lead || cur[c] >= R.q[c].size()) continue;
            Op& o = R.q[c][cur[c]];
            if (o.kind == 0 && o.merged == key.merged && o.block == key.block) { ops[n] = &o; who[n++] = c; }
        }
        rc = key.merged(ops, n, s);
        if (rc) break;
        for (int i = 0; i < n; ++i) ++cur[who[i]];
    }
    for (int c = 0; c < GROUP_MAX; ++c)
        if ((slots >> c) & 1ull) R.q[c].clear();
    return rc;
}

// =============================================================================
// warp / block helpers
// =============================================================================
#define FULL 0xffffffffu

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ double warp_scan_incl(double v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double t = __shfl_up_sync(FULL, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// CTA-wide reductions over up to 1024 threads; `red` is 33 doubles of shared memory.
// Fixed combination order => deterministic.  All threads get the result.
__device__ __forceinline__ double block_max(double v, double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        double x = (lane < nw) ? red[lane] : -BNPC_INF;
        x = warp_max(x);
        if (lane == 0) red[32] = x;
    }
    __syncthreads();
    return red[32];
}
__device__ __forceinline__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        double x = (lane < nw) ? red[lane] : 0.0;
        x = warp_sum(x);
        if (lane == 0) red[32] = x;
    }
    __syncthreads();
    return red[32];
}
// inclusive scan across threads; returns this thread's inclusive value, *total = sum
__device__ __forceinline__ double block_scan_incl(double v, double* red, double* total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    double inc = warp_scan_incl(v, lane);
    __syncthreads();
    if (lane == 31) red[w] = inc;
    __syncthreads();
    if (w == 0) {
        double x = (lane < nw) ? red[lane] : 0.0;
        double s = warp_scan_incl(x, lane);
        red[lane] = s - x;                       // exclusive prefix of warp sums
        if (lane == 31) red[32] = s;
    }
    __syncthreads();
    inc += red[w];
    *total = red[32];
    return inc;
}

// =============================================================================
// input path: bit-plane packing
// =============================================================================
__device__ __forceinline__ void pack_planes_kernel(const double* __restrict__ xf, const int8_t* __restrict__ xi,
                                   int N, int M, int W, uint32_t* __restrict__ x1,
                                   uint32_t* __restrict__ x0, int32_t* __restrict__ n1,
                                   int32_t* __restrict__ n0, int n_ctas) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)n_ctas * blockDim.x) >> 5;
    for (long long cell = warp; cell < N; cell += nwarps) {
        int c1 = 0, c0 = 0;
        for (int w = 0; w < W; ++w) {
            const int m = w * 32 + lane;
            int code = 0;
            if (m < M) {
                if (xf) {
                    const double v = xf[cell * M + m];
                    code = (v == 1.0) ? 1 : ((v == 0.0) ? 2 : 0);
                } else {
                    const int v = xi[cell * M + m];
                    code = (v == 1) ? 1 : ((v == 0) ? 2 : 0);
                }
            }
            const uint32_t b1 = __ballot_sync(FULL, code == 1);
            const uint32_t b0 = __ballot_sync(FULL, code == 2);
            if (lane == 0) { x1[cell * W + w] = b1; x0[cell * W + w] = b0; }
            c1 += __popc(b1); c0 += __popc(b0);
        }
        if (lane == 0) { n1[cell] = c1; n0[cell] = c0; }
    }
}

// =============================================================================
// random numbers (production mode)
// =============================================================================
__device__ __forceinline__ void fill_uniform_kernel(double* __restrict__ out, long long n, uint64_t seed,
                                    uint64_t stream_id, int n_levels) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // pair index
    if (2 * i >= n) return;
    Philox g(seed);
    const uint4 r = g((uint64_t)i, stream_id);
    double a = u01(r.x, r.y), b = u01(r.z, r.w);
    if (n_levels > 0) { a = floor(a * n_levels); b = floor(b * n_levels); }
    out[2 * i] = a;
    if (2 * i + 1 < n) out[2 * i + 1] = b;
}

__device__ __forceinline__ uint32_t feistel_round(uint32_t r, uint32_t k) {
    uint32_t h = r * 0x9E3779B1u + k;
    h ^= h >> 15; h *= 0x85EBCA77u; h ^= h >> 13; h *= 0xC2B2AE3Du; h ^= h >> 16;
    return h;
}
// Element i of the streams the fill kernels write, computed where it is consumed (production
// mode: the consumers take a NULL buffer and the stream instead; one launch less per draw).
__device__ __forceinline__ double uniform_at(uint64_t seed, uint64_t stream_id, long long i, int n_levels) {
    Philox g(seed);
    const uint4 r = g((uint64_t)(i >> 1), stream_id);
    double v = (i & 1) ? u01(r.z, r.w) : u01(r.x, r.y);
    if (n_levels > 0) v = floor(v * n_levels);
    return v;
}
__device__ __forceinline__ int permutation_at(uint64_t seed, uint64_t stream_id, int i, int n, int half_bits) {
    Philox g(seed);
    const uint4 k0 = g(0xFE157E1ull, stream_id), k1 = g(0xFE157E2ull, stream_id);
    const uint32_t keys[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
    const uint32_t mask = (1u << half_bits) - 1u;
    uint32_t x = (uint32_t)i;
    do {                                      // cycle walking keeps the image inside [0,n)
        uint32_t l = x >> half_bits, r = x & mask;
#pragma unroll
        for (int rd = 0; rd < 8; ++rd) {
            const uint32_t t = l ^ (feistel_round(r, keys[rd]) & mask);
            l = r; r = t;
        }
        x = (l << half_bits) | r;
    } while (x >= (uint32_t)n);
    return (int)x;
}
static inline int feistel_half_bits(int n) {
    int bits = 2;
    while ((1ll << bits) < n) ++bits;
    if (bits & 1) ++bits;
    return bits / 2;
}
__device__ __forceinline__ void fill_permutation_kernel(int32_t* __restrict__ out, int n, uint64_t seed,
                                        uint64_t stream_id, int half_bits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = permutation_at(seed, stream_id, i, n, half_bits);
}

// =============================================================================
// likelihood
// =============================================================================
__device__ __forceinline__ void logprob_tables_kernel(const float* __restrict__ theta, const int32_t* __restrict__ ids,
                                      int R, int M, double FN, double FP, double2* __restrict__ lp) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)R * M) return;
    const int r = (int)(i / M), m = (int)(i % M);
    const long long row = ids ? ids[r] : r;
    double a, b;
    log_p1_p0(theta[row * M + m], FN, FP, a, b);
    lp[i] = make_double2(a, b);
}

// ll[r][k0+kk] for a tile of 128 cells x LL_KT clusters.  One thread per cell keeps LL_KT
// FP64 accumulators; the (log p1, log p0) pairs of 128 mutations x LL_KT clusters are staged
// in shared memory and read as 16-byte broadcasts.  Algorithmic work: 2 FMA per
// (cell, mutation, cluster); bits select exact 0.0/1.0 multipliers so the sum equals the
// reference's masked sum.
#define LL_KT 8
#define LL_MT 128
#define LL_THREADS 128
__device__ __forceinline__ void ll_matrix_kernel(const uint32_t* __restrict__ x1, const uint32_t* __restrict__ x0, int W, int M,
                 const int32_t* __restrict__ cells, int cell_stride, int C,
                 const double2* __restrict__ lp, int K, double* __restrict__ ll, int ldk) {
    __shared__ double2 tile[LL_MT][LL_KT];
    const int r = blockIdx.x * LL_THREADS + threadIdx.x;
    const int k0 = blockIdx.y * LL_KT;
    const bool live = r < C;
    const long long cell = live ? (cells ? cells[(long long)r * cell_stride] : r) : 0;
    const uint4* p1 = reinterpret_cast<const uint4*>(x1 + cell * W);
    const uint4* p0 = reinterpret_cast<const uint4*>(x0 + cell * W);
    double acc[LL_KT];
#pragma unroll
    for (int kk = 0; kk < LL_KT; ++kk) acc[kk] = 0.0;

    for (int m0 = 0; m0 < M; m0 += LL_MT) {
        __syncthreads();
        for (int i = threadIdx.x; i < LL_MT * LL_KT; i += LL_THREADS) {
            const int kk = i / LL_MT, mm = i % LL_MT;      // coalesced along mutations
            double2 v = make_double2(0.0, 0.0);
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
                    // an entry selects log p1 (.x), log p0 (.y) or nothing: one addition per
                    // entry, bit-identical to fma(f1, lp1, fma(f0, lp0, acc)) with 0/1 factors
                    // and half its chain of dependent operations
                    const bool b1 = (u1 >> bit) & 1u, b0 = (u0 >> bit) & 1u;
                    const double2* t = tile[q * 32 + bit];
#pragma unroll
                    for (int kk = 0; kk < LL_KT; ++kk) {
                        const double2 v = t[kk];
                        acc[kk] += b1 ? v.x : (b0 ? v.y : 0.0);
                    }
                }
            }
        }
    }
    if (live) {
#pragma unroll
        for (int kk = 0; kk < LL_KT; ++kk)
            if (k0 + kk < K) ll[(long long)r * ldk + k0 + kk] = acc[kk];
    }
}

// ll[r][0..K) for K <= 2 rows of log-probabilities (the 2-column matrices of the restricted Gibbs
// scans, libs/CRP.py:635-638): the cells of a split-merge move are few (one or two clusters), so
// one thread per cell leaves the GPU idle.  Eight threads share a cell (thread s takes the words
// w = s, s + 8, ...), the whole table sits in shared memory, partial sums are combined by a fixed
// butterfly.  Lane s walks the bits of a word rotated by s: a table entry is 16 bytes (log p1, log p0),
// so the eight lanes of a cell read eight different 16-byte bank groups (a rotation by 4 s left
// only two distinct groups: 4-way conflicts, ncu: short-scoreboard stalls dominated).
#define LLP_CELLS 32
__device__ __forceinline__ void ll_few_kernel(const uint32_t* __restrict__ x1, const uint32_t* __restrict__ x0, int W, int M,
              const int32_t* __restrict__ cells, int cell_stride, int C,
              const double2* __restrict__ lp, int K, double* __restrict__ ll, int ldk,
              const float* __restrict__ theta, double FN, double FP) {
    extern __shared__ __align__(16) unsigned char llp_smem[];
    double2* tab = reinterpret_cast<double2*>(llp_smem);          // [K][W * 32 + 1], zero beyond M
    const int Mp = W * 32;
    // lp == NULL: the table is built here from the K rows of theta (what bnpc_logprob_tables
    // would have written; one launch less per restricted Gibbs scan).  Entry Mp of every column
    // is a zero pair: a missing entry of the data reads it, so the inner loop has no branch and
    // no select (adding 0.0 leaves the sum as it is).
    for (int i = threadIdx.x; i < K * (Mp + 1); i += blockDim.x) {
        const int k = i / (Mp + 1), m = i % (Mp + 1);
        double2 v = make_double2(0.0, 0.0);
        if (m < M) {
            if (lp) v = lp[(long long)k * M + m];
            else log_p1_p0(theta[(long long)k * M + m], FN, FP, v.x, v.y);
        }
        tab[i] = v;
    }
    __syncthreads();
    const int sub = threadIdx.x & 7;
    const int r = blockIdx.x * LLP_CELLS + (threadIdx.x >> 3);
    const bool live = r < C;
    const long long cell = live ? (cells ? cells[(long long)r * cell_stride] : r) : 0;
    const uint32_t* r1 = x1 + cell * W;
    const uint32_t* r0 = x0 + cell * W;
    const double* t0 = reinterpret_cast<const double*>(tab);
    const double* t1 = t0 + 2 * (Mp + 1);
    double a0 = 0.0, a1 = 0.0;
    if (live) {
        for (int w = sub; w < W; w += 8) {
            const uint32_t u1 = r1[w], u0 = r0[w];
            if ((u1 | u0) == 0u) continue;
#pragma unroll 4
            for (int i = 0; i < 32; ++i) {
                const int bit = (i + sub) & 31;
                const uint32_t b1 = (u1 >> bit) & 1u, b0 = (u0 >> bit) & 1u;
                const int at = (b1 | b0) ? 2 * (w * 32 + bit) + (b1 ? 0 : 1) : 2 * Mp;
                a0 += t0[at];
                if (K > 1) a1 += t1[at];
            }
        }
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        a0 += __shfl_xor_sync(FULL, a0, o);
        a1 += __shfl_xor_sync(FULL, a1, o);
    }
    if (live && sub == 0) {
        ll[(long long)r * ldk] = a0;
        if (K > 1) ll[(long long)r * ldk + 1] = a1;
    }
}

// log-likelihood of one cell under one (log p1, log p0) row; same arithmetic as above
__device__ __forceinline__ double cell_row_ll(const uint32_t* __restrict__ r1,
                                              const uint32_t* __restrict__ r0, int W,
                                              const double2* __restrict__ lp) {
    double acc = 0.0;
    for (int w = 0; w < W; ++w) {
        const uint32_t u1 = r1[w], u0 = r0[w];
        if ((u1 | u0) == 0u) continue;
        const double2* t = lp + w * 32;
#pragma unroll 4
        for (int bit = 0; bit < 32; ++bit) {
            if ((u1 | u0) >> bit & 1u) {
                const double2 v = t[bit];
                acc += ((u1 >> bit) & 1u) ? v.x : v.y;
            }
        }
    }
    return acc;
}

// =============================================================================
// Gibbs sweep
// =============================================================================
__device__ __forceinline__ void gibbs_prepare_kernel(const int32_t* __restrict__ perm, const double* __restrict__ u,
                                     const int32_t* __restrict__ assign, const int32_t* __restrict__ n1,
                                     const int32_t* __restrict__ n0, int N, double c1, double c0,
                                     double lnew_prior, bnpc_visit_t* __restrict__ visit, uint64_t seed,
                                     uint64_t stream_id, int half_bits) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N) return;
    // perm / u == NULL: the visiting order and the uniforms of streams stream_id+1, +2 (what
    // bnpc_fill_permutation / bnpc_fill_uniform would have written)
    const int c = perm ? perm[t] : permutation_at(seed, stream_id + 1, t, N, half_bits);
    bnpc_visit_t v;
    v.u = u ? u[t] : uniform_at(seed, stream_id + 2, t, 0);
    // popcount form of libs/CRP.py:230-234: every observed 1 contributes c1, every 0 c0
    v.lnew = ((double)n1[c] * c1 + (double)n0[c] * c0) + lnew_prior;
    v.e_new = 0.0;
    v.ref = 0.0;
    v.cell = c;
    v.old = assign[c];
    v.c_old = -1;
    v.n_opt = BNPC_MAX_OPT + 1;       // "unknown rivals" until bnpc_gibbs_candidates has run
    v.i_old = 0;
    v.t = t;
    v.flags = 0;
    v.e_max = 1.0f;
    visit[t] = v;
}

// static options of every visited cell (see include/bnpc_b200.h), kept in column order: columns
// follow the list order at the start of the epoch and deaths preserve relative order, so a walk
// over the options is a walk in list order.
#define CAND_THREADS 128
__device__ __forceinline__ void gibbs_candidates_kernel(const double* __restrict__ ll, int ldk, int K,
                        const int32_t* __restrict__ col_of_id,
                        bnpc_visit_t* __restrict__ visit, bnpc_cand_t* __restrict__ cand,
                        int C, double slack, double c_norm, int32_t* __restrict__ blk) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    int uncertain = 0;
    if (r < C) {
        const double* row = ll + (long long)r * ldk;
        const int c_old = col_of_id[visit[r].old];
        const double lnew_ll = visit[r].lnew + c_norm;          // ll_new + log(alpha)
        const double u = visit[r].u;
        bnpc_cand_t out;
        double val[BNPC_MAX_OPT];
#pragma unroll
        for (int i = 0; i < BNPC_MAX_OPT; ++i) { out.e[i] = 0.0; out.col[i] = 0; val[i] = -BNPC_INF; }
        out.pad[0] = out.pad[1] = out.pad[2] = 0;
        int n = BNPC_MAX_OPT + 1, i_old = 0, flags = 0;
        double ref = 0.0, e_new = 0.0, e_max = 1.0;
        if (c_old >= 0 && c_old < K) {
            const double v_old = row[c_old];
            const double thr = v_old - 40.0 - slack;
            n = 0;
            ref = fmax(v_old, lnew_ll);
            for (int k = 0; k < K; ++k) {
                const double v = row[k];
                if (k != c_old && !(v > thr)) continue;
                if (k == c_old) i_old = n;
#pragma unroll
                for (int i = 0; i < BNPC_MAX_OPT; ++i)
                    if (i == n) { val[i] = v; out.col[i] = (uint16_t)k; }
                ref = fmax(ref, v);
                ++n;
            }
            if (n > BNPC_MAX_OPT) {
                n = BNPC_MAX_OPT + 1;
            } else {
                e_max = 0.0;
#pragma unroll
                for (int i = 0; i < BNPC_MAX_OPT; ++i)
                    if (i < n) { out.e[i] = exp(val[i] - ref); e_max = fmax(e_max, out.e[i]); }
                e_new = exp(lnew_ll - ref);
                if (n == 1 && lnew_ll < v_old - 40.0 && u > 3e-10 && u < 1.0 - 3e-10)
                    flags = BNPC_VISIT_CERTAIN;
            }
        }
        visit[r].e_new = e_new;
        visit[r].ref = ref;
        visit[r].c_old = c_old;
        visit[r].n_opt = n;
        visit[r].i_old = i_old;
        visit[r].flags = flags;
        visit[r].e_max = __double2float_ru(e_max);
        cand[r] = out;
        uncertain = !(flags & BNPC_VISIT_CERTAIN);
    }
    const int cnt = __syncthreads_count(uncertain);
    if (threadIdx.x == 0) blk[blockIdx.x] = cnt;
}

// exclusive scan of the per-block counts (one CTA); blk[nb] and st[BNPC_ST_NUNC] = total
__device__ __forceinline__ void compact_scan_kernel(int32_t* blk, int nb, int32_t* st) {
    __shared__ int wsum[33];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int per = (nb + 1023) / 1024;
    const int j0 = tid * per, j1 = min(j0 + per, nb);
    int s = 0;
    for (int j = j0; j < j1; ++j) s += blk[j];
    int inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        const int x = wsum[lane];
        int si = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL, si, o);
            if (lane >= o) si += t;
        }
        wsum[lane] = si - x;
        if (lane == 31) wsum[32] = si;
    }
    __syncthreads();
    int run = wsum[w] + inc - s;
    for (int j = j0; j < j1; ++j) { const int c = blk[j]; blk[j] = run; run += c; }
    if (tid == 0) { blk[nb] = wsum[32]; st[BNPC_ST_NUNC] = wsum[32]; }
}

__device__ __forceinline__ void compact_scatter_kernel(const bnpc_visit_t* __restrict__ visit, const bnpc_cand_t* __restrict__ cand,
                       int C, const int32_t* __restrict__ blk, bnpc_visit_t* __restrict__ visit_c,
                       bnpc_cand_t* __restrict__ cand_c) {
    __shared__ int wcnt[CAND_THREADS / 32];
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const bool take = r < C && !(visit[r].flags & BNPC_VISIT_CERTAIN);
    const unsigned m = __ballot_sync(FULL, take);
    if (lane == 0) wcnt[w] = __popc(m);
    __syncthreads();
    if (!take) return;
    int pos = blk[blockIdx.x] + __popc(m & ((1u << lane) - 1u));
    for (int i = 0; i < w; ++i) pos += wcnt[i];
    const uint4* sv = reinterpret_cast<const uint4*>(visit + r);
    uint4* dv = reinterpret_cast<uint4*>(visit_c + pos);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(bnpc_visit_t) / 16); ++i) dv[i] = sv[i];
    const uint4* sc = reinterpret_cast<const uint4*>(cand + r);
    uint4* dc = reinterpret_cast<uint4*>(cand_c + pos);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(bnpc_cand_t) / 16); ++i) dc[i] = sc[i];
}

__device__ __forceinline__ void gibbs_epoch_begin_kernel(const int32_t* __restrict__ live, int K, int32_t* lst,
                                         int32_t* cnt, int32_t* col_of_id, int idcap, int32_t* st,
                                         int first) {
    for (int i = threadIdx.x; i < idcap; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    for (int j = threadIdx.x; j < K; j += blockDim.x) {
        const int id = live[2 * j];
        lst[j] = id;
        cnt[id] = live[2 * j + 1];
        col_of_id[id] = j;
    }
    if (threadIdx.x == 0) {
        st[BNPC_ST_K] = K;
        st[BNPC_ST_NEXTRA] = 0;
        st[BNPC_ST_FLAGS] = 0;
        st[BNPC_ST_NMANY] = 0;
        if (first) {
            st[BNPC_ST_TDONE] = 0; st[BNPC_ST_BIRTHS] = 0; st[BNPC_ST_MOVED] = 0; st[BNPC_ST_SLOW] = 0;
            st[12] = 0; st[13] = 0; st[14] = 0; st[15] = 0;       // phase clocks of the parallel sequencer
        }
    }
}

// ---- bulk async copy (TMA engine, 1-D) + mbarrier -----------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

#include "bnpc_lean.cuh"
#include "bnpc_tc.cuh"
#include "bnpc_tc_i8.cuh"
#include "bnpc_tc_i8s.cuh"
#include "bnpc_estimators.cuh"

#define SW_STAGE_CELLS 32
#define SW_NSTAGE 16
#define SW_MAXL 1024          /* longest list / widest ll matrix the sequencer regime handles */

#define SW_PAR_WARPS 16
#define SW_WINDOW 384         /* records per window of the parallel sequencer */
#define SW_BUF_STAGES 24      /* staging for two windows (768 records) */

struct SweepShared {
    alignas(128) bnpc_visit_t vis_stage[SW_BUF_STAGES][SW_STAGE_CELLS];
    alignas(128) bnpc_cand_t cand_stage[SW_BUF_STAGES][SW_STAGE_CELLS];
    alignas(8) uint64_t bar[SW_NSTAGE];
    double red[40];
    // the live list (position j <-> insertion order) while it has at most SW_MAXL entries
    int s_id[SW_MAXL], s_cnt[SW_MAXL], s_src[SW_MAXL];   // id, size, ll column (>=0) or -(extra+2)
    int s_pos_of_col[SW_MAXL];                  // ll column -> list position, -1 once the cluster died
    int s_xpos[BNPC_MAX_EXTRA];                 // cluster born in this epoch -> list position or -1
    int L, t, pending, stop;
    int birth_cell, birth_t;
    int n_extra, births, moved, slow;
    int tmp_i, pick;
    int hang;
    int compact, crec;                          // compact mode on / next compacted record
    double s_row[BNPC_LEAN_MAXK];               // lean epochs: exact ll row of the cell in hand, by column
    // parallel sequencer (lean epochs): one warp per group of option-graph components
    int owner_of_col[BNPC_LEAN_MAXK];
    int abort_idx;                              // window-relative index of the first record that needs the exact path
    unsigned short own_idx[SW_PAR_WARPS][SW_WINDOW];     // a warp's records of the window, in order
    unsigned int mlog[SW_PAR_WARPS][SW_WINDOW];           // its moves of the window: idx | from << 9 | to << 17
    unsigned char own_b[2][SW_WINDOW];                    // owner warp of each record of the window (0xff: warp 0, exact path)
};

// One exact categorical draw for a list that fits a warp (libs/CRP.py:88-100 + numpy choice,
// libs/CRP.py:274-277).  Lanes [0,L) hold live clusters, lane L the new-cluster option.
__device__ __forceinline__ int warp_categorical(double l, int L, double u, int lane) {
    const bool in = lane <= L;
    const double lmax = warp_max(in ? l : -BNPC_INF);
    const double d = l - lmax;
    const double e = in ? exp(d) : 0.0;
    double S = warp_sum(e) - 1.0;                  // sum over all but (one copy of) the max
    if (S < 0.0) S = 0.0;
    double z = d - log1p(S);
    z = fmin(fmax(z, kLogEps), 0.0);
    const double p = in ? exp(z) : 0.0;
    const double cdf = warp_scan_incl(p, lane);
    const double total = __shfl_sync(FULL, cdf, L);
    const unsigned gt = __ballot_sync(FULL, in && (cdf / total > u));
    return gt ? (__ffs(gt) - 1) : L;
}

// column / extra -> list position maps, rebuilt after every structural change (rare)
__device__ __forceinline__ void sweep_rebuild_maps(SweepShared& sh, int L) {
    const int lane = threadIdx.x;
    for (int c = lane; c < SW_MAXL; c += 32) sh.s_pos_of_col[c] = -1;
    for (int e = lane; e < BNPC_MAX_EXTRA; e += 32) sh.s_xpos[e] = -1;
    __syncwarp();
    for (int j = lane; j < L; j += 32) {
        const int src = sh.s_src[j];
        if (src >= 0) { if (src < SW_MAXL) sh.s_pos_of_col[src] = j; }
        else if (src <= -2) sh.s_xpos[-src - 2] = j;
    }
    __syncwarp();
}

// Exact treatment of one cell by one warp, lanes <-> clusters, for lists of at most 31 clusters
// (libs/CRP.py:262-288).  `row` is the cell's ll row in GLOBAL memory.  Returns 0: the cell
// stayed, 1: list/sizes changed, 2: it opens a new cluster (CTA-wide work follows).
__device__ int sweep_exact_cell(const bnpc_sweep_args_t& a, SweepShared& sh, const bnpc_visit_t& v,
                                const double* row, int t, int& L) {
    const int lane = threadIdx.x;
    int id = -1, cnt = 0, src = -1;
    if (lane < L) { id = sh.s_id[lane]; cnt = sh.s_cnt[lane]; src = sh.s_src[lane]; }
    const int old = v.old;
    const unsigned om = __ballot_sync(FULL, lane < L && id == old);
    int lo = __ffs(om) - 1;
    const int ocnt = __shfl_sync(FULL, cnt, lo < 0 ? 0 : lo);
    if (lo >= 0 && ocnt == 1) {
        // the cluster dies with its last cell: close the gap, list order stays insertion order
        const int id2 = __shfl_down_sync(FULL, id, 1), cnt2 = __shfl_down_sync(FULL, cnt, 1),
                  src2 = __shfl_down_sync(FULL, src, 1);
        if (lane == lo) a.cnt[old] = 0;
        if (lane >= lo && lane < L - 1) {
            id = id2; cnt = cnt2; src = src2;
            a.lst[lane] = id;
        } else if (lane == L - 1) {
            id = -1; cnt = 0; src = -1;
        }
        --L;
        lo = -1;
        __syncwarp();
        if (lane <= L) { sh.s_id[lane] = id; sh.s_cnt[lane] = cnt; sh.s_src[lane] = src; }
        __syncwarp();
        sweep_rebuild_maps(sh, L);
    }
    double l = -BNPC_INF;
    if (lane < L) {
        const double val = (src >= 0) ? row[src] : a.llx[(long long)(-src - 2) * a.ldx + (t - a.t_epoch0)];
        // CRP weight log n - log(N-1+alpha) with the cell itself taken out (libs/CRP.py:83-85,262-270)
        l = val + (a.logn[(lane == lo) ? cnt - 1 : cnt] - a.c_norm);
    } else if (lane == L) {
        l = v.lnew;
    }
    const int pick = warp_categorical(l, L, v.u, lane);
    if (pick == lo) return 0;
    if (lo >= 0 && lane == lo) {                    // leave the old cluster
        --cnt;
        a.cnt[id] = cnt;
        sh.s_cnt[lane] = cnt;
    }
    if (pick == L) {
        if (lane == 0) { sh.pending = 1; sh.birth_cell = v.cell; sh.birth_t = t; }
        __syncwarp();
        return 2;
    }
    if (lane == pick) {
        ++cnt;
        a.cnt[id] = cnt;
        a.assign[v.cell] = id;
        sh.s_cnt[lane] = cnt;
    }
    __syncwarp();
    return 1;
}

// The same for lists of 32..63 clusters: every lane holds list positions `lane` and `lane + 32`
// (the live list itself stays in shared memory).  Panel-like data (short rows, many plausible
// clusters per cell) sends most visits here.
__device__ int sweep_exact_cell2(const bnpc_sweep_args_t& a, SweepShared& sh, const bnpc_visit_t& v,
                                 const double* row, int t, int& L) {
    const int lane = threadIdx.x;
    const int old = v.old;
    int lo = -1;
    {
        const unsigned m0 = __ballot_sync(FULL, lane < L && sh.s_id[lane] == old);
        const unsigned m1 = __ballot_sync(FULL, lane + 32 < L && sh.s_id[lane + 32] == old);
        lo = m0 ? (__ffs(m0) - 1) : (m1 ? 32 + __ffs(m1) - 1 : -1);
    }
    if (lo >= 0 && sh.s_cnt[lo] == 1) {
        // the cluster dies with its last cell: close the gap, list order stays insertion order
        int id_[2], cnt_[2], src_[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int p = lane + 32 * h;
            const bool mv = p >= lo && p < L - 1;
            id_[h] = mv ? sh.s_id[p + 1] : 0; cnt_[h] = mv ? sh.s_cnt[p + 1] : 0; src_[h] = mv ? sh.s_src[p + 1] : 0;
        }
        __syncwarp();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int p = lane + 32 * h;
            if (p >= lo && p < L - 1) {
                sh.s_id[p] = id_[h]; sh.s_cnt[p] = cnt_[h]; sh.s_src[p] = src_[h];
                a.lst[p] = id_[h];
            } else if (p == L - 1) {
                sh.s_id[p] = -1; sh.s_cnt[p] = 0; sh.s_src[p] = -1;
            }
        }
        if (lane == 0) a.cnt[old] = 0;
        --L;
        lo = -1;
        __syncwarp();
        sweep_rebuild_maps(sh, L);
    }
    double l[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int p = lane + 32 * h;
        l[h] = -BNPC_INF;
        if (p < L) {
            const int src = sh.s_src[p], cnt = sh.s_cnt[p];
            const double val = (src >= 0) ? row[src] : a.llx[(long long)(-src - 2) * a.ldx + (t - a.t_epoch0)];
            l[h] = val + (a.logn[(p == lo) ? cnt - 1 : cnt] - a.c_norm);
        } else if (p == L) {
            l[h] = v.lnew;
        }
    }
    // libs/CRP.py:88-100 + numpy choice over L + 1 entries
    const bool in0 = lane <= L, in1 = lane + 32 <= L;
    const double lmax = warp_max(fmax(in0 ? l[0] : -BNPC_INF, in1 ? l[1] : -BNPC_INF));
    const double d0 = l[0] - lmax, d1 = l[1] - lmax;
    double S = warp_sum((in0 ? exp(d0) : 0.0) + (in1 ? exp(d1) : 0.0)) - 1.0;
    if (S < 0.0) S = 0.0;
    const double lse = log1p(S);
    const double p0 = in0 ? exp(fmin(fmax(d0 - lse, kLogEps), 0.0)) : 0.0;
    const double p1 = in1 ? exp(fmin(fmax(d1 - lse, kLogEps), 0.0)) : 0.0;
    const double cdf0 = warp_scan_incl(p0, lane);
    const double cdf1 = warp_scan_incl(p1, lane) + __shfl_sync(FULL, cdf0, 31);
    const double total = (L < 32) ? __shfl_sync(FULL, cdf0, L & 31) : __shfl_sync(FULL, cdf1, (L - 32) & 31);
    const unsigned g0 = __ballot_sync(FULL, in0 && (cdf0 / total > v.u));
    const unsigned g1 = __ballot_sync(FULL, in1 && (cdf1 / total > v.u));
    const int pick = g0 ? (__ffs(g0) - 1) : (g1 ? 32 + __ffs(g1) - 1 : L);
    if (pick == lo) return 0;
    if (lane == 0) {
        if (lo >= 0) {                                  // leave the old cluster
            const int c = sh.s_cnt[lo] - 1;
            sh.s_cnt[lo] = c;
            a.cnt[old] = c;
        }
        if (pick == L) {
            sh.pending = 1; sh.birth_cell = v.cell; sh.birth_t = t;
        } else {
            const int c = sh.s_cnt[pick] + 1, idp = sh.s_id[pick];
            sh.s_cnt[pick] = c;
            a.cnt[idp] = c;
            a.assign[v.cell] = idp;
        }
    }
    __syncwarp();
    return pick == L ? 2 : 1;
}

#define OUT_STAY 0
#define OUT_MOVE 1
#define OUT_COMPLEX 2
#define SW_GUARD 1e-10        /* draws closer than this (in probability) to an interval edge go exact */

// Sequencer regime (lists of at most SW_MAXL clusters), run by warp 0.  The sweep is
// sequential, but only through the cluster sizes.  32 consecutive records are scored in
// parallel (lane <-> visit) against the current sizes, looking only at the cell's static
// options (bnpc_gibbs_candidates) plus clusters born in this epoch; every other cluster sits
// on the reference's 1e-15 probability floor.  Restricted to the options the draw of
// _normalize_log_probs + numpy choice (libs/CRP.py:88-100, 277) is LINEAR in the cluster sizes:
// weight_i = n_i * exp(ll_i - ref), the new-cluster option weighs alpha * exp(ll_new - ref), and
// the pick is the first option whose running sum exceeds u * total -- integer sizes times
// per-cell constants, no transcendental in the sequential part.
//
// Each lane keeps the margin of its decision in weight units.  One move shifts every interval
// edge and u*total of another cell by at most 2*e_lim, so a lane's decision survives nb earlier
// moves while margin > 4*e_lim*nb: all movers of the stage up to the first lane whose margin (or
// cluster size) does not allow that are applied AT ONCE; the rest is scored again.  The linear
// form and the reference's log-space arithmetic differ by rounding (~1e-13) and by the floored
// entries (<= 1024e-15): whenever u is within SW_GUARD of an interval edge, and for cluster
// death, a new cluster, or more than BNPC_MAX_CAND rivals, the cell goes through the exact draw
// over the whole list in the reference's own arithmetic (warp-cooperative for lists of up to 31
// clusters, CTA-wide otherwise).
//
// Compact mode: while no cluster was born in the epoch and every cluster has at least two cells,
// visits flagged BNPC_VISIT_CERTAIN stay whatever the sizes are; the sequencer then walks only
// the compacted records of the other visits (bnpc_gibbs_compact).  The first event that could
// invalidate this (a birth, a cluster down to one cell) sends the rest of the epoch through the
// dense records.
__device__ void sweep_sequencer(const bnpc_sweep_args_t& a, SweepShared& sh) {
    const int lane = threadIdx.x;
    const unsigned lt_mask = (1u << lane) - 1u;
    int L = sh.L;
    const int ldk = a.ldk;
    int moved = 0, slow = 0;

    int cmin = 0x7fffffff;
    for (int j = lane; j < L; j += 32) {
        const int id = a.lst[j];
        const int c = a.cnt[id];
        sh.s_id[j] = id; sh.s_cnt[j] = c; sh.s_src[j] = a.col_of_id[id];
        cmin = min(cmin, c);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cmin = min(cmin, __shfl_xor_sync(FULL, cmin, o));
    if (lane == 0) {
        for (int s = 0; s < SW_NSTAGE; ++s) mbar_init(&sh.bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    sweep_rebuild_maps(sh, L);
    const int n_extra = sh.n_extra;
    // lean epochs: a cluster never holds a single certain visit (gibbs_finalize_kernel), so the
    // compacted records stay valid whatever the sizes do; dense epochs check the sizes instead
    const bool lean = a.ll == nullptr;
    const bool compact = lean || (sh.compact && cmin >= 2 && n_extra == 0);
    const int need_cnt = (compact && !lean) ? 2 : 1;   // cells a cluster keeps after a batched move

    const bnpc_visit_t* vbase = compact ? a.visit_c : a.visit + a.t_epoch0;
    const bnpc_cand_t* cbase = compact ? a.cand_c : a.cand + a.t_epoch0;
    const int n_rec = compact ? a.st[BNPC_ST_NUNC] : a.t_end - a.t_epoch0;
    const int first_rec = compact ? sh.crec : sh.t - a.t_epoch0;

    // stages = 32 consecutive visit + option records, streamed into shared memory by bulk
    // async copies (TMA engine) NSTAGE ahead of the sequencer
    const int s_first = first_rec / SW_STAGE_CELLS;
    const int n_stages = (first_rec < n_rec) ? (n_rec - 1) / SW_STAGE_CELLS - s_first + 1 : 0;
    auto issue = [&](int g) {                      // lane 0 only
        const int slot = g % SW_NSTAGE;
        const int sp = (s_first + g) * SW_STAGE_CELLS;
        const int nc = min(SW_STAGE_CELLS, n_rec - sp);
        const uint32_t b_v = (uint32_t)(nc * sizeof(bnpc_visit_t));
        const uint32_t b_c = (uint32_t)(nc * sizeof(bnpc_cand_t));
        mbar_expect_tx(&sh.bar[slot], b_v + b_c);
        bulk_g2s(sh.vis_stage[slot], vbase + sp, b_v, &sh.bar[slot]);
        bulk_g2s(sh.cand_stage[slot], cbase + sp, b_c, &sh.bar[slot]);
    };
    int issued = 0;
    if (lane == 0)
        for (; issued < n_stages && issued < SW_NSTAGE; ++issued) issue(issued);
    issued = __shfl_sync(FULL, issued, 0);

    int waited = 0;
    bool leave = false;
    int next_t = a.t_end, next_rec = n_rec;
    for (int g = 0; g < n_stages && !leave; ++g) {
        const int slot = g % SW_NSTAGE;
        const uint32_t parity = (uint32_t)((g / SW_NSTAGE) & 1);
        {
            long long spins = 0;
            while (!mbar_try_wait(&sh.bar[slot], parity)) {
                if (++spins > (1ll << 24)) { sh.hang = 1; break; }
            }
        }
        waited = g + 1;
        const int rec0 = (s_first + g) * SW_STAGE_CELLS;
        const int nc = min(SW_STAGE_CELLS, n_rec - rec0);
        const bnpc_visit_t* vis = sh.vis_stage[slot];
        const int my = lane < nc ? lane : 0;
        const bnpc_visit_t v = vis[my];
        const long long tx = v.t - a.t_epoch0;
        // the lane's options live in registers for the whole stage
        double e[BNPC_MAX_OPT];
        int col[BNPC_MAX_OPT];
        {
            const bnpc_cand_t* cd = &sh.cand_stage[slot][my];
#pragma unroll
            for (int i = 0; i < BNPC_MAX_OPT; ++i) { e[i] = cd->e[i]; col[i] = cd->col[i]; }
        }
        const int n_opt = v.n_opt, i_old = v.i_old;
        // warp-uniform bound of the option loops
        int n_max = (lane < nc && n_opt <= BNPC_MAX_OPT) ? n_opt : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n_max = max(n_max, __shfl_xor_sync(FULL, n_max, o));

        int lo_lane = max(0, first_rec - rec0);
        bool need_eval = lane >= lo_lane && lane < nc;      // per lane
        int outcome = OUT_STAY, to_pos = -1, p_old = -1, c_own = 0;
        double slack = 0.0;                                 // validity margin of `outcome` (weight units)
        double e_lim = (double)v.e_max;                     // largest weight among all its options
        while (lo_lane < nc) {
            if (need_eval) {
                need_eval = false;
                outcome = OUT_COMPLEX;
                to_pos = -1;
                slack = 0.0;
                c_own = 0;
                p_old = (n_opt <= BNPC_MAX_OPT) ? sh.s_pos_of_col[v.c_old] : -1;
                if (p_old >= 0) c_own = sh.s_cnt[p_old];
                if (p_old >= 0 && c_own > 1 && sh.s_id[p_old] == v.old) {
                    double cum[BNPC_MAX_OPT];
                    int pos[BNPC_MAX_OPT];
                    double S = 0.0;
#pragma unroll
                    for (int i = 0; i < BNPC_MAX_OPT; ++i) {
                        cum[i] = 0.0; pos[i] = -1;
                    }
#pragma unroll
                    for (int i = 0; i < BNPC_MAX_OPT; ++i) {
                        if (i >= n_max) break;                // warp-uniform
                        if (i < n_opt) {
                            const int p = sh.s_pos_of_col[col[i]];
                            int c = 0;
                            if (p >= 0) c = sh.s_cnt[p];
                            if (i == i_old) --c;
                            pos[i] = p;
                            S += (double)c * e[i];
                            cum[i] = S;
                        }
                    }
                    double Sx = 0.0;
                    for (int x = 0; x < n_extra; ++x) {       // clusters born in this epoch
                        const int pe = sh.s_xpos[x];
                        if (pe >= 0) {
                            const double ex = exp(a.llx[(long long)x * a.ldx + tx] - v.ref);
                            e_lim = fmax(e_lim, ex);
                            Sx += (double)sh.s_cnt[pe] * ex;
                        }
                    }
                    const double total = (S + Sx) + v.e_new;
                    const double target = v.u * total;
                    double lo = 0.0, hi = 0.0;
                    int pick = -1;
#pragma unroll
                    for (int i = 0; i < BNPC_MAX_OPT; ++i) {
                        if (i >= n_max) break;                // warp-uniform
                        if (i < n_opt && pick < 0) {
                            if (cum[i] > target) { pick = i; hi = cum[i]; to_pos = pos[i]; }
                            else lo = cum[i];
                        }
                    }
                    if (pick < 0) {
                        double run = S;
                        for (int x = 0; x < n_extra && pick < 0; ++x) {
                            const int pe = sh.s_xpos[x];
                            if (pe >= 0) {
                                lo = run;
                                run += (double)sh.s_cnt[pe] * exp(a.llx[(long long)x * a.ldx + tx] - v.ref);
                                if (run > target) { pick = BNPC_MAX_OPT + x; hi = run; to_pos = pe; }
                            }
                        }
                    }
                    // pick < 0: the new-cluster option (or rounding at the top edge) -> exact draw
                    if (pick >= 0 && to_pos >= 0) {
                        slack = fmin(target - lo, hi - target) - 2.0 * SW_GUARD * total;
                        if (slack > 0.0) outcome = (pick == i_old) ? OUT_STAY : OUT_MOVE;
                    }
                }
            }
            const bool act = lane >= lo_lane && lane < nc;
            const unsigned pend = __ballot_sync(FULL, act && outcome != OUT_STAY);
            if (!pend) break;
            // would this lane's decision survive if every earlier pending lane moved a cell?
            const int nb = __popc(pend & lt_mask);
            bool ok = true;
            if (act) {
                if (outcome == OUT_COMPLEX) ok = false;
                else if (nb > 0) ok = (slack - 4.0 * e_lim * (double)nb > 0.0) &&
                                      (c_own - nb > ((outcome == OUT_MOVE) ? need_cnt : 1));
                else ok = (outcome == OUT_STAY) || (c_own > need_cnt);
            }
            const unsigned bad = __ballot_sync(FULL, !ok);
            const int fb = bad ? (__ffs(bad) - 1) : 32;
            const unsigned batch = pend & ((fb >= 32) ? FULL : ((1u << fb) - 1u));
            if (batch) {
                // all pending lanes before the first doubtful one are movers: apply them at once
                if ((batch >> lane) & 1u) {
                    atomicSub(&sh.s_cnt[p_old], 1);
                    atomicAdd(&sh.s_cnt[to_pos], 1);
                    a.assign[v.cell] = sh.s_id[to_pos];
                }
                moved += __popc(batch);
                __syncwarp();
                if (fb >= nc) break;
                lo_lane = fb;                               // score the rest against the new sizes
                need_eval = lane >= lo_lane && lane < nc;
                continue;
            }
            // the first pending lane cannot take the linear form: exact draw
            const int f = fb;
            const int t_f = __shfl_sync(FULL, v.t, f);
            ++slow;
            if (lean || L > 63) {                   // CTA-wide work (exact row / exact draw), then come back
                if (lane == 0) sh.pending = lean ? 3 : 2;
                next_t = t_f; next_rec = rec0 + f + 1;
                leave = true;
                break;
            }
            const double* row_f = a.ll + (long long)(t_f - a.t_epoch0) * ldk;
            const int status = (L > 31) ? sweep_exact_cell2(a, sh, vis[f], row_f, t_f, L)
                                        : sweep_exact_cell(a, sh, vis[f], row_f, t_f, L);
            if (status) ++moved;
            if (status == 2 || (compact && status != 0)) {
                // a birth, or (compact mode) sizes changed outside the batched bookkeeping:
                // re-enter and decide again whether the compacted records may be used
                next_t = t_f + 1; next_rec = rec0 + f + 1;
                leave = true;
                break;
            }
            lo_lane = f + 1;
            if (status != 0) need_eval = lane >= lo_lane && lane < nc;   // structure may have changed
        }
        __syncwarp();
        if (!leave && lane == 0 && issued < n_stages) issue(issued);
        if (!leave && issued < n_stages) ++issued;
    }
    // drain copies still in flight before shared memory is reused or the kernel ends
    for (int g = waited; g < issued; ++g) {
        long long spins = 0;
        while (!mbar_try_wait(&sh.bar[g % SW_NSTAGE], (uint32_t)((g / SW_NSTAGE) & 1))) {
            if (++spins > (1ll << 24)) { sh.hang = 1; break; }
        }
    }
    __syncwarp();
    // sizes changed by batched moves live in shared memory only: publish them
    for (int j = lane; j < L; j += 32) a.cnt[sh.s_id[j]] = sh.s_cnt[j];
    if (lane == 0) {
        sh.L = L;
        sh.t = next_t;
        if (compact) sh.crec = next_rec; else sh.compact = 0;
        sh.moved += moved;
        sh.slow += slow;
    }
}

// Parallel sequencer of lean epochs (all 8 warps of the CTA).  Two visits interact only through
// the sizes of clusters both have among their options, so the records split into the connected
// components of the "option graph" on the clusters (components_kernel); a warp owns a group of
// components and walks ITS records in visiting order with the batched scoring of
// sweep_sequencer, while the other warps do the same for theirs.  Records are taken in windows of
// 256 (bulk-copied into shared memory, double-buffered); all warps meet at the end of a window.
// A record that needs the exact path (guard band, death, birth, > 8 rivals) must see the state of
// ALL clusters exactly at its position: the warp that meets it posts its index, every warp undoes
// the moves it made past that index in this window (they are logged; assignments are only
// written at the end of a window), and the CTA handles the record before the walk resumes behind it.
__device__ void sweep_sequencer_par(const bnpc_sweep_args_t& a, SweepShared& sh) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    int L = sh.L;
    if (w == 0) {
        for (int j = lane; j < L; j += 32) {
            const int id = a.lst[j];
            sh.s_id[j] = id; sh.s_cnt[j] = a.cnt[id]; sh.s_src[j] = a.col_of_id[id];
        }
        __syncwarp();
        sweep_rebuild_maps(sh, L);
    }
    if (tid < BNPC_LEAN_MAXK) sh.owner_of_col[tid] = a.comp[192 + tid];
    if (tid == 0) {
        mbar_init(&sh.bar[0], 1);
        mbar_init(&sh.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        sh.abort_idx = 0x7fffffff;
    }
    __syncthreads();
    const int n_rec = a.st[BNPC_ST_NUNC];
    const int first_rec = sh.crec;
    const int n_win = (first_rec < n_rec) ? (n_rec - first_rec + SW_WINDOW - 1) / SW_WINDOW : 0;
    bnpc_visit_t* vbuf = &sh.vis_stage[0][0];           // two windows of 256 records
    bnpc_cand_t* cbuf = &sh.cand_stage[0][0];
    auto issue = [&](int g) {                            // thread 0 only
        const int r0 = first_rec + g * SW_WINDOW;
        const int nw = min(SW_WINDOW, n_rec - r0);
        const uint32_t b_v = (uint32_t)(nw * sizeof(bnpc_visit_t)), b_c = (uint32_t)(nw * sizeof(bnpc_cand_t));
        mbar_expect_tx(&sh.bar[g & 1], b_v + b_c);
        bulk_g2s(vbuf + (g & 1) * SW_WINDOW, a.visit_c + r0, b_v, &sh.bar[g & 1]);
        bulk_g2s(cbuf + (g & 1) * SW_WINDOW, a.cand_c + r0, b_c, &sh.bar[g & 1]);
    };
    int issued = 0;
    if (tid == 0)
        for (; issued < n_win && issued < 2; ++issued) issue(issued);
    // owner bytes of a window (bnpc_gibbs_exact) travel one window ahead through a register
    auto owner_of = [&](int g) -> unsigned char {
        const int r = first_rec + g * SW_WINDOW + tid;
        return (g < n_win && tid < SW_WINDOW && r < n_rec) ? a.owner_c[r] : (unsigned char)0xfe;
    };
    if (tid < SW_WINDOW) sh.own_b[0][tid] = owner_of(0);
    unsigned char nxt_owner = owner_of(1);
    __syncthreads();
    int moved = 0, slow = 0;
    int next_t = a.t_end, next_rec = n_rec;
    bool leave = false;
    int waited = 0;
    // phase clocks of warp 0 (kilo-cycles into st[12..15]: window data wait | ownership pass |
    // batched walk | window barrier + commit)
    long long ph[4] = {0, 0, 0, 0}, tk = clock64();
    for (int g = 0; g < n_win && !leave; ++g) {
        {
            long long spins = 0;
            while (!mbar_try_wait(&sh.bar[g & 1], (uint32_t)((g >> 1) & 1))) {
                if (++spins > (1ll << 24)) { sh.hang = 1; break; }
            }
        }
        { const long long now = clock64(); ph[0] += now - tk; tk = now; }
        waited = g + 1;
        const int r0 = first_rec + g * SW_WINDOW;
        const int nw = min(SW_WINDOW, n_rec - r0);
        const bnpc_visit_t* vis = vbuf + (g & 1) * SW_WINDOW;
        const bnpc_cand_t* cnd = cbuf + (g & 1) * SW_WINDOW;
        // ---- this warp's records of the window, in order (all owner bytes first, then the
        // ballots, then the writes: no chain through the running count) ----
        unsigned mine_mask[SW_WINDOW / 32];
#pragma unroll
        for (int c = 0; c < SW_WINDOW / 32; ++c) {
            const int rec = c * 32 + lane;
            const int o = (rec < nw) ? sh.own_b[g & 1][rec] : 0xfe;
            mine_mask[c] = __ballot_sync(FULL, (o == 0xff) ? (w == 0) : (o == w));
        }
        int own = 0;
#pragma unroll
        for (int c = 0; c < SW_WINDOW / 32; ++c) {
            const unsigned m = mine_mask[c];
            if ((m >> lane) & 1u) sh.own_idx[w][own + __popc(m & lt_mask)] = (unsigned short)(c * 32 + lane);
            own += __popc(m);
        }
        __syncwarp();
        { const long long now = clock64(); ph[1] += now - tk; tk = now; }
        int n_log = 0;
        bool stopped = false;
        for (int base = 0; base < own && !stopped; base += 32) {
            int nc = min(32, own - base);
            const int nc_batch = nc;
            const int rec = sh.own_idx[w][base + (lane < nc ? lane : 0)];
            const bnpc_visit_t v = vis[rec];
            double e[BNPC_MAX_OPT];
            int pos[BNPC_MAX_OPT];
            {
                const bnpc_cand_t* cd = &cnd[rec];
#pragma unroll
                for (int i = 0; i < BNPC_MAX_OPT; ++i) { e[i] = cd->e[i]; pos[i] = cd->col[i]; }
            }
            const int n_opt = v.n_opt, i_old = v.i_old;
            int n_max = (lane < nc && n_opt <= BNPC_MAX_OPT) ? n_opt : 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) n_max = max(n_max, __shfl_xor_sync(FULL, n_max, o));
            // list positions of the options (columns -> positions change only on the exact path)
#pragma unroll
            for (int i = 0; i < BNPC_MAX_OPT; ++i) {
                if (i >= n_max) break;
                pos[i] = (i < n_opt && n_opt <= BNPC_MAX_OPT) ? sh.s_pos_of_col[pos[i]] : -1;
            }
            const int p_old = (n_opt <= BNPC_MAX_OPT && v.c_old >= 0) ? sh.s_pos_of_col[v.c_old] : -1;
            const bool own_ok = p_old >= 0 && sh.s_id[p_old] == v.old;
            const double e_lim = (double)v.e_max;
            int lo_lane = 0;
            bool need_eval = lane < nc;
            int outcome = OUT_STAY, to_pos = -1, c_own = 0;
            double slack = 0.0;
            while (lo_lane < nc) {
                // records at or behind a posted exact-path record are left for the next walk
                const int ab = *(volatile int*)&sh.abort_idx;
                if (ab != 0x7fffffff) {
                    const unsigned keep = __ballot_sync(FULL, lane < nc && rec < ab);
                    nc = __popc(keep);                   // own_idx is ascending: a prefix survives
                    // the walk of this warp ends here only if the posted record cuts THIS batch (then
                    // all later batches lie behind it too); a batch that lies wholly in front of it is
                    // followed by the next one -- every record in front of the posted one must be walked
                    if (nc < nc_batch) stopped = true;
                    if (lo_lane >= nc) break;
                }
                if (need_eval) {
                    need_eval = false;
                    outcome = OUT_COMPLEX;
                    to_pos = -1;
                    slack = 0.0;
                    c_own = own_ok ? sh.s_cnt[p_old] : 0;
                    if (own_ok && c_own > 1) {
                        double cum[BNPC_MAX_OPT];
                        double S = 0.0;
#pragma unroll
                        for (int i = 0; i < BNPC_MAX_OPT; ++i) cum[i] = 0.0;
#pragma unroll
                        for (int i = 0; i < BNPC_MAX_OPT; ++i) {
                            if (i >= n_max) break;            // warp-uniform
                            int c = 0;
                            if (i < n_opt && pos[i] >= 0) c = sh.s_cnt[pos[i]];
                            if (i == i_old) --c;
                            S = fma((double)c, e[i], S);      // e[i] = 0 beyond the options
                            cum[i] = S;
                        }
                        const double total = S + v.e_new;
                        const double target = v.u * total;
                        int pick = 0;
                        double dmin = target;                 // distance to the nearest interval edge
#pragma unroll
                        for (int i = 0; i < BNPC_MAX_OPT; ++i) {
                            if (i >= n_max) break;
                            if (i < n_opt) {
                                const double d = cum[i] - target;
                                pick += (d <= 0.0) ? 1 : 0;
                                dmin = fmin(dmin, fabs(d));
                            }
                        }
#pragma unroll
                        for (int i = 0; i < BNPC_MAX_OPT; ++i)
                            if (i == pick) to_pos = pos[i];
                        // pick == n_opt: the new-cluster option (or rounding at the top edge) -> exact draw
                        if (pick < n_opt && to_pos >= 0) {
                            slack = dmin - 2.0 * SW_GUARD * total;
                            if (slack > 0.0) outcome = (pick == i_old) ? OUT_STAY : OUT_MOVE;
                        }
                    }
                }
                const bool act = lane >= lo_lane && lane < nc;
                const unsigned pend = __ballot_sync(FULL, act && outcome != OUT_STAY);
                if (!pend) break;
                const int nb = __popc(pend & lt_mask);
                bool ok = true;
                if (act) {
                    if (outcome == OUT_COMPLEX) ok = false;
                    else if (nb > 0) ok = (slack - 4.0 * e_lim * (double)nb > 0.0) && (c_own - nb > 1);
                }
                const unsigned bad = __ballot_sync(FULL, !ok);
                const int fb = bad ? (__ffs(bad) - 1) : 32;
                const unsigned batch = pend & ((fb >= 32) ? FULL : ((1u << fb) - 1u));
                if (batch) {
                    if ((batch >> lane) & 1u) {
                        atomicSub(&sh.s_cnt[p_old], 1);
                        atomicAdd(&sh.s_cnt[to_pos], 1);
                        sh.mlog[w][n_log + __popc(batch & lt_mask)] =
                            (unsigned)rec | ((unsigned)p_old << 9) | ((unsigned)to_pos << 17);
                    }
                    n_log += __popc(batch);
                    __syncwarp();
                    if (fb >= nc) break;
                    lo_lane = fb;
                    need_eval = lane >= lo_lane && lane < nc;
                    continue;
                }
                // the first pending record cannot take the linear form: post it, stop this warp here
                const int rec_f = __shfl_sync(FULL, rec, fb);
                if (lane == 0) atomicMin(&sh.abort_idx, rec_f);
                stopped = true;
                break;
            }
        }
        { const long long now = clock64(); ph[2] += now - tk; tk = now; }
        if (tid < SW_WINDOW) sh.own_b[(g + 1) & 1][tid] = nxt_owner;
        nxt_owner = owner_of(g + 2);
        __syncthreads();
        // ---- end of the window: commit the moves in front of a posted record, undo the rest ----
        const int ab = sh.abort_idx;
        for (int i = lane; i < n_log; i += 32) {
            const unsigned en = sh.mlog[w][i];
            const int rec = en & 0x1ff, from = (en >> 9) & 0xff, to = (en >> 17) & 0xff;
            if (rec < ab) {
                a.assign[vis[rec].cell] = sh.s_id[to];
                ++moved;
            } else {
                atomicAdd(&sh.s_cnt[from], 1);
                atomicSub(&sh.s_cnt[to], 1);
            }
        }
        __syncthreads();
        if (ab != 0x7fffffff) {
            // the CTA handles record `ab` exactly, then the walk resumes behind it
            next_t = vis[ab].t;
            next_rec = r0 + ab + 1;
            if (tid == 0) { sh.pending = 3; ++slow; }
            leave = true;
        } else if (tid == 0 && issued < n_win) {
            issue(issued);
        }
        if (!leave && issued < n_win) ++issued;      // (uniform: every thread tracks the count)
        { const long long now = clock64(); ph[3] += now - tk; tk = now; }
    }
    if (tid == 0)
        for (int k = 0; k < 4; ++k) a.st[12 + k] += (int)(ph[k] >> 10);
    // drain copies still in flight before shared memory is reused or the kernel ends
    for (int g = waited; g < ((tid == 0) ? issued : 0); ++g) {
        long long spins = 0;
        while (!mbar_try_wait(&sh.bar[g & 1], (uint32_t)((g >> 1) & 1))) {
            if (++spins > (1ll << 24)) { sh.hang = 1; break; }
        }
    }
    // moves counted per lane: fold into the CTA totals
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) moved += __shfl_xor_sync(FULL, moved, o);
    if (lane == 0 && moved) atomicAdd(&sh.moved, moved);
    __syncthreads();
    for (int j = tid; j < L; j += blockDim.x) a.cnt[sh.s_id[j]] = sh.s_cnt[j];
    if (tid == 0) {
        sh.t = next_t;
        sh.crec = next_rec;
        sh.slow += slow;
        sh.abort_idx = 0x7fffffff;
    }
    __syncthreads();
}

// Wide regime (dense epochs of at most 63 clusters in which most visits have more rivals than an
// option record holds: short rows, e.g. the panel shape): a.ll holds the WEIGHTS
// e[t][k] = exp(ll[t][k] - ref_t) (gibbs_weights_kernel), and one warp walks ALL visits with lanes
// <-> list positions (2 l, 2 l + 1).  Restricted to nothing, the draw of _normalize_log_probs +
// numpy choice (libs/CRP.py:88-100, 277) is linear in the cluster sizes, as in sweep_sequencer:
// weight_k = n_k e_k, the new-cluster option weighs e_new; one warp scan of the lane pair sums gives
// the running sums in list order, the pick is the first position whose running sum exceeds
// u * total.  Within SW_GUARD of an interval edge, for a cluster that would die, and for the
// new-cluster option the visit goes to the exact draw (pending = 3: exact FP64 row on demand, then
// sweep_exact_cell / sweep_exact_cell2); a birth ends the epoch.  Visit fields and the two weights
// of a lane are prefetched four visits ahead.
#define WIDE_PF 4
__device__ void sweep_wide(const bnpc_sweep_args_t& a, SweepShared& sh) {
    const int lane = threadIdx.x;
    int L = sh.L;
    for (int j = lane; j < L; j += 32) {
        const int id = a.lst[j];
        sh.s_id[j] = id; sh.s_cnt[j] = a.cnt[id]; sh.s_src[j] = a.col_of_id[id];
    }
    __syncwarp();
    sweep_rebuild_maps(sh, L);
    const int pa = 2 * lane, pb = 2 * lane + 1;
    const int src_a = (pa < L) ? sh.s_src[pa] : -1, src_b = (pb < L) ? sh.s_src[pb] : -1;
    const int ldk = a.ldk, t_end = a.t_end;
    int t = sh.t, moved = 0;
    struct Pre { double u, e_new, ea, eb; int cell, old, c_old; };
    auto fetch = [&](int tt) -> Pre {
        Pre p;
        const int tc = (tt < t_end) ? tt : t_end - 1;              // always a valid record
        const bnpc_visit_t* v = a.visit + tc;
        p.u = v->u; p.e_new = v->e_new; p.cell = v->cell; p.old = v->old; p.c_old = v->c_old;
        const double* row = a.ll + (long long)(tc - a.t_epoch0) * ldk;
        p.ea = (src_a >= 0) ? row[src_a] : 0.0;
        p.eb = (src_b >= 0) ? row[src_b] : 0.0;
        return p;
    };
    Pre ring[WIDE_PF];
#pragma unroll
    for (int k = 0; k < WIDE_PF; ++k) ring[k] = fetch(t + k);
    bool leave = false;
    while (t < t_end && !leave) {
#pragma unroll
        for (int k = 0; k < WIDE_PF; ++k) {
            // (the refill is unconditional: a load under a branch would be waited for at once)
            const Pre c = ring[k];
            ring[k] = fetch(t + WIDE_PF);
            if (t < t_end && !leave) {
                const int po = (c.c_old >= 0 && c.c_old < SW_MAXL) ? sh.s_pos_of_col[c.c_old] : -1;
                bool exact = po < 0;
                int c_own = 0;
                if (!exact) {
                    c_own = sh.s_cnt[po];
                    exact = sh.s_id[po] != c.old || c_own <= 1;
                }
                if (!exact) {
                    const int na = (pa < L) ? sh.s_cnt[pa] - (pa == po ? 1 : 0) : 0;
                    const int nb = (pb < L) ? sh.s_cnt[pb] - (pb == po ? 1 : 0) : 0;
                    const double wa = (double)na * c.ea, wb = (double)nb * c.eb;
                    const double cum_b = warp_scan_incl(wa + wb, lane);
                    const double cum_a = cum_b - wb;
                    const double total = __shfl_sync(FULL, cum_b, 31) + c.e_new;
                    const double target = c.u * total;
                    const unsigned ga = __ballot_sync(FULL, pa < L && cum_a > target);
                    const unsigned gb = __ballot_sync(FULL, pb < L && cum_b > target);
                    int pick = 0x7fffffff;
                    if (ga) pick = 2 * (__ffs(ga) - 1);
                    if (gb) pick = min(pick, 2 * (__ffs(gb) - 1) + 1);
                    if (pick >= L) {
                        exact = true;                              // the new-cluster option
                    } else {
                        const double hi = __shfl_sync(FULL, (pick & 1) ? cum_b : cum_a, pick >> 1);
                        const int q = (pick > 0) ? pick - 1 : 0;
                        double lo = __shfl_sync(FULL, (q & 1) ? cum_b : cum_a, q >> 1);
                        if (pick == 0) lo = 0.0;
                        const double slack = fmin(target - lo, hi - target) - 2.0 * SW_GUARD * total;
                        if (!(slack > 0.0)) {
                            exact = true;
                        } else if (pick != po) {
                            if (lane == 0) {
                                sh.s_cnt[po] = c_own - 1;
                                sh.s_cnt[pick] += 1;
                                a.assign[c.cell] = sh.s_id[pick];
                            }
                            ++moved;
                            __syncwarp();
                        }
                    }
                }
                if (exact) leave = true; else ++t;
            }
        }
    }
    __syncwarp();
    for (int j = lane; j < L; j += 32) a.cnt[sh.s_id[j]] = sh.s_cnt[j];
    if (lane == 0) {
        sh.L = L;
        sh.t = t;
        sh.moved += moved;
        if (leave) { sh.pending = 3; sh.slow += 1; }
    }
}

// one cell with the whole CTA (list longer than a warp)
__device__ void sweep_block_cell(const bnpc_sweep_args_t& a, SweepShared& sh, int t, int L,
                                 const double* rowp) {
    const int tid = threadIdx.x, B = blockDim.x;
    const bnpc_visit_t v = a.visit[t];
    int old = v.old;
    const bool died = (a.cnt[old] == 1);
    __syncthreads();
    if (died) {
        if (tid == 0) sh.tmp_i = -1;
        __syncthreads();
        for (int j = tid; j < L; j += B)
            if (a.lst[j] == old) sh.tmp_i = j;
        __syncthreads();
        const int pos = sh.tmp_i;
        for (int base = pos; base < L - 1; base += B) {
            const int j = base + tid;
            int nv = 0;
            if (j < L - 1) nv = a.lst[j + 1];
            __syncthreads();
            if (j < L - 1) a.lst[j] = nv;
            __syncthreads();
        }
        if (tid == 0) a.cnt[old] = 0;
        --L;
        old = -1;
        __syncthreads();
    }
    const long long tx = t - a.t_epoch0;
    double m_all = -BNPC_INF, m_riv = -BNPC_INF, l_old = -BNPC_INF;
    for (int j = tid; j <= L; j += B) {
        double l;
        if (j < L) {
            const int id = a.lst[j];
            int n = a.cnt[id];
            if (id == old) --n;
            const int col = a.col_of_id[id];
            const double val = (col >= 0) ? rowp[col] : a.llx[(long long)(-col - 2) * a.ldx + tx];
            l = val + (a.logn[n] - a.c_norm);
            if (id == old) l_old = l; else m_riv = fmax(m_riv, l);
        } else {
            l = v.lnew;
            m_riv = fmax(m_riv, l);
        }
        a.scratch[j] = l;
        m_all = fmax(m_all, l);
    }
    const double lmax = block_max(m_all, sh.red);
    const double lriv = block_max(m_riv, sh.red);
    const double lold = block_max(l_old, sh.red);
    if (!died && lriv <= lold - 40.0 && v.u > 1e-8 && v.u < 1.0 - 1e-8) {
        if (tid == 0) sh.t = t + 1;
        __syncthreads();
        return;
    }
    double s = 0.0;
    for (int j = tid; j <= L; j += B) s += exp(a.scratch[j] - lmax);
    double S = block_sum(s, sh.red) - 1.0;
    if (S < 0.0) S = 0.0;
    const double lse = log1p(S);
    const int seg = (L + 1 + B - 1) / B;
    const int j0 = tid * seg, j1 = min(j0 + seg, L + 1);
    double T = 0.0;
    for (int j = j0; j < j1; ++j) {
        const double z = fmin(fmax(a.scratch[j] - lmax - lse, kLogEps), 0.0);
        const double p = exp(z);
        a.scratch[j] = p;
        T += p;
    }
    double total;
    const double inc = block_scan_incl(T, sh.red, &total);
    if (tid == 0) sh.tmp_i = 0x7fffffff;
    __syncthreads();
    if (j0 < j1 && inc / total > v.u) atomicMin(&sh.tmp_i, tid);
    __syncthreads();
    if (tid == sh.tmp_i) {
        double cdf = inc - T;
        int pick = j1 - 1;
        for (int j = j0; j < j1; ++j) {
            cdf += a.scratch[j];
            if (cdf / total > v.u) { pick = j; break; }
        }
        sh.pick = pick;
    }
    if (sh.tmp_i == 0x7fffffff && tid == 0) sh.pick = L;
    __syncthreads();
    if (tid == 0) {
        const int pick = sh.pick;
        sh.slow += 1;
        if (pick == L) {
            if (!died) a.cnt[old] -= 1;
            sh.pending = 1; sh.birth_cell = v.cell; sh.birth_t = t;
            sh.moved += 1;
        } else {
            const int idp = a.lst[pick];
            if (idp != old) {
                if (!died) a.cnt[old] -= 1;
                a.cnt[idp] += 1;
                a.assign[v.cell] = idp;
                sh.moved += 1;
            }
        }
        sh.t = t + 1;
        sh.L = L;
    }
    __syncthreads();
}

// a cell opened a new cluster (libs/CRP.py:281-282,291-299,183-188): pick the smallest free
// id, draw its theta row, build its log-prob row and its ll column for the rest of the epoch
__device__ void sweep_birth(const bnpc_sweep_args_t& a, SweepShared& sh) {
    const int tid = threadIdx.x, B = blockDim.x;
    const int cell = sh.birth_cell, t = sh.birth_t, L = sh.L, e = sh.n_extra, b = sh.births;
    const int M = a.M, W = a.W;
    if (a.beta_rows && b >= a.n_beta_rows) {
        if (tid == 0) { sh.stop |= BNPC_STOP_TAPE_EMPTY; sh.pending = 0; }
        __syncthreads();
        return;
    }
    if (tid == 0) sh.tmp_i = 0x7fffffff;
    __syncthreads();
    for (int i = tid; i <= L && i < a.idcap; i += B)
        if (a.cnt[i] == 0) atomicMin(&sh.tmp_i, i);
    __syncthreads();
    const int nid = sh.tmp_i;
    if (nid >= a.idcap) {
        if (tid == 0) { sh.stop |= BNPC_STOP_CAPACITY; sh.pending = 0; }
        __syncthreads();
        return;
    }
    const bool lean = a.ll == nullptr || a.wide;   // lean and wide epochs end at a birth: no ll column is built
    double2* lpx = lean ? nullptr : reinterpret_cast<double2*>(a.lpx) + (long long)e * M;
    const uint32_t* r1 = a.x1 + (long long)cell * W;
    const uint32_t* r0 = a.x0 + (long long)cell * W;
    for (int m = tid; m < M; m += B) {
        const int b1 = (r1[m >> 5] >> (m & 31)) & 1, b0 = (r0[m >> 5] >> (m & 31)) & 1;
        double val;
        if (a.beta_rows) {
            val = a.beta_rows[(long long)b * M + m];
        } else {
            PhiloxStream rs(a.seed, a.stream_id + 0x1000000ull * (uint64_t)(b + 1), (uint64_t)m);
            val = beta_sample(a.p + b1, a.q + b0, rs);
        }
        const float th = clip_theta(val);
        a.theta[(long long)nid * M + m] = th;
        if (!lean) {
            double p1, p0;
            log_p1_p0(th, a.FN, a.FP, p1, p0);
            lpx[m] = make_double2(p1, p0);
        }
    }
    __syncthreads();
    if (!lean) {
        double* col = a.llx + (long long)e * a.ldx;
        for (int tt = t + 1 + tid; tt < a.t_end; tt += B) {
            const long long c2 = a.visit[tt].cell;
            col[tt - a.t_epoch0] = cell_row_ll(a.x1 + c2 * W, a.x0 + c2 * W, (M + 31) >> 5, lpx);
        }
    }
    if (tid == 0) {
        a.lst[L] = nid;
        a.cnt[nid] = 1;
        a.col_of_id[nid] = -(e + 2);
        a.assign[cell] = nid;
        sh.L = L + 1;
        sh.n_extra = e + 1;
        sh.births = b + 1;
        sh.pending = 0;
        if (lean || e + 1 >= BNPC_MAX_EXTRA) sh.stop |= BNPC_STOP_EXTRA_FULL;
    }
    __syncthreads();
}

template <int MAXT>
__device__ __forceinline__ void gibbs_sweep_kernel(const bnpc_sweep_args_t& a) {
    extern __shared__ __align__(128) unsigned char sweep_smem[];
    SweepShared& sh = *reinterpret_cast<SweepShared*>(sweep_smem);
    const int tid = threadIdx.x;
    long long clk0 = 0;
    unsigned long long ns0 = 0;
    if (tid == 0) {
        clk0 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0));
        sh.L = a.st[BNPC_ST_K];
        sh.t = a.t_begin;
        sh.pending = 0;
        sh.stop = a.st[BNPC_ST_FLAGS];
        sh.n_extra = a.st[BNPC_ST_NEXTRA];
        sh.births = a.st[BNPC_ST_BIRTHS];
        sh.moved = a.st[BNPC_ST_MOVED];
        sh.slow = a.st[BNPC_ST_SLOW];
        sh.hang = 0;
        sh.compact = (a.visit_c != nullptr && a.cand_c != nullptr && a.t_begin == a.t_epoch0 &&
                      sh.n_extra == 0) ? 1 : 0;
        sh.crec = 0;
        // lean epoch whose visits mostly have more rivals than an option record holds (short rows,
        // e.g. panel data): every such visit would take the exact path with an on-demand FP64 row;
        // the dense FP64 matrix is the better route -- tell the host before anything is done
        if (a.ll == nullptr && a.t_begin == a.t_epoch0 && !sh.stop &&
            a.st[BNPC_ST_NMANY] > max(64, (a.t_end - a.t_begin) / 50)) {
            sh.stop = BNPC_STOP_MANY;
            a.st[BNPC_ST_FLAGS] = BNPC_STOP_MANY;
        }
    }
    __syncthreads();
    if (sh.stop) return;                            // an earlier launch of this sweep stopped

    for (;;) {
        const int t = sh.t, L = sh.L, stop = sh.stop | sh.hang, pending = sh.pending;
        __syncthreads();
        if (t >= a.t_end || stop) break;
        if (L + 1 + BNPC_MAX_EXTRA >= a.idcap) {
            if (tid == 0) sh.stop |= BNPC_STOP_CAPACITY;
            __syncthreads();
            continue;
        }
        if (pending == 2) {
            // the sequencer handed this cell to the CTA-wide exact draw
            if (tid == 0) sh.pending = 0;
            __syncthreads();
            sweep_block_cell(a, sh, t, L, a.ll + (long long)(t - a.t_epoch0) * a.ldk);
        } else if (pending == 3) {
            // lean epoch: the exact FP64 row of this cell, one thread per live cluster in the
            // arithmetic of ll_matrix_kernel, then the exact draw
            if (tid == 0) sh.pending = 0;
            const long long cell = a.visit[t].cell;
            if (tid < L) {
                const int col = a.col_of_id[a.lst[tid]];
                sh.s_row[col] = cell_row_ll(a.x1 + cell * a.W, a.x0 + cell * a.W, (a.M + 31) >> 5,
                                            reinterpret_cast<const double2*>(a.lp) + (long long)col * a.M);
            }
            __syncthreads();
            if (L <= 63) {
                if (tid < 32) {
                    int L2 = L;
                    const int status = (L > 31) ? sweep_exact_cell2(a, sh, a.visit[t], sh.s_row, t, L2)
                                                : sweep_exact_cell(a, sh, a.visit[t], sh.s_row, t, L2);
                    if (tid == 0) {
                        sh.L = L2;
                        sh.t = t + 1;
                        if (status) sh.moved += 1;
                    }
                }
                __syncthreads();
            } else {
                sweep_block_cell(a, sh, t, L, sh.s_row);
            }
        } else if (a.wide) {
            if (L > 63) {                            // (cannot happen: births end a wide epoch)
                if (tid == 0) sh.stop |= BNPC_STOP_REPACK;
                __syncthreads();
                continue;
            }
            if (tid < 32) sweep_wide(a, sh);
            __syncthreads();
        } else if (a.ll == nullptr && a.comp != nullptr && a.owner_c != nullptr && blockDim.x == 32 * SW_PAR_WARPS) {
            sweep_sequencer_par(a, sh);
        } else if (L < SW_MAXL && a.ldk <= SW_MAXL) {
            if (tid < 32) sweep_sequencer(a, sh);
            __syncthreads();
        } else if (L < SW_MAXL / 2) {
            // the list shrank far below the width of this epoch's ll matrix: let the host
            // start a narrower epoch so that the sequencer regime can take over
            if (tid == 0) sh.stop |= BNPC_STOP_REPACK;
            __syncthreads();
            continue;
        } else {
            sweep_block_cell(a, sh, t, L, a.ll + (long long)(t - a.t_epoch0) * a.ldk);
        }
        if (sh.pending == 1) sweep_birth(a, sh);
    }
    __syncthreads();
    const int L = sh.L;
    for (int j = tid; j < L; j += blockDim.x) {
        const int id = a.lst[j];
        a.live_out[2 * j] = id;
        a.live_out[2 * j + 1] = a.cnt[id];
    }
    if (tid == 0) {
        a.st[BNPC_ST_K] = L;
        a.st[BNPC_ST_TDONE] = sh.t;
        a.st[BNPC_ST_FLAGS] = sh.stop | (sh.hang ? 0x100 : 0);
        a.st[BNPC_ST_NEXTRA] = sh.n_extra;
        a.st[BNPC_ST_BIRTHS] = sh.births;
        a.st[BNPC_ST_MOVED] = sh.moved;
        a.st[BNPC_ST_SLOW] = sh.slow;
        unsigned long long ns1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
        const long long dc = clock64() - clk0;
        a.st[BNPC_ST_CYCLES] = (int)(dc >> 10);          // kilo-cycles of this launch
        a.st[BNPC_ST_NANOS] = (int)((ns1 - ns0) >> 10);   // ~microseconds of this launch
    }
}

// =============================================================================
// sufficient statistics
// =============================================================================
__device__ __forceinline__ void set_ranks_kernel(const int32_t* __restrict__ ids, int K, int32_t* rank_of_id) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < K) rank_of_id[ids[j]] = j;
}

__device__ __forceinline__ void group_members_kernel(const int32_t* __restrict__ assign, int N,
                                     const int32_t* __restrict__ rank_of_id,
                                     const int32_t* __restrict__ seg_off, int32_t* cursor,
                                     int32_t* __restrict__ members) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int lane = threadIdx.x & 31;
    const int r = rank_of_id[assign[n]];
    const unsigned active = __activemask();
    const unsigned peers = __match_any_sync(active, r);
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&cursor[r], __popc(peers));
    base = __shfl_sync(peers, base, leader);
    members[seg_off[r] + base + __popc(peers & ((1u << lane) - 1u))] = n;
}

#define SS_CHUNK 256          /* rows per CTA: 32 per thread at the widest rows, 4 batches of 8 rows in flight (measured:
                                 33 us at C3 against 47 us with 512-row chunks and 74 us with 1024-thread CTAs) */
#define SS_THREADS 256        /* 8 row groups at the widest rows: 1024 threads over 2040-row chunks were
                                 2x slower (32-way contention on the shared-memory counters) */
// grid (chunks, R, word blocks): a CTA counts ones/zeros per mutation over up to SS_CHUNK members
// of one segment for a block of 32 mutation words.  A thread owns one word column and strides
// over the rows, so that a warp reads whole 128-byte row pieces (coalesced); the 32 per-bit
// counters of a word are bit-sliced (7 planes, carry-save adder tree over eight rows at a time) and
// spread once per chunk into 8 registers of four byte-wide counters each (byte b of a[j] counts
// bit j + 8 b) for the shared-memory accumulation; at most 127 rows per thread.
__device__ __forceinline__ void suffstat_kernel(const uint32_t* __restrict__ x1, const uint32_t* __restrict__ x0, int W, int M,
                const int32_t* __restrict__ members, const int32_t* __restrict__ seg_off,
                int32_t* __restrict__ S1, int32_t* __restrict__ S0, int wc_log2, int n_wblk, int chunk) {
    // [plane][word column][bit], rows padded to 33: the 32 lanes of a warp hold 32 different word
    // columns and add to the SAME bit index at the same time -- unpadded that is one bank for all of
    // them (ncu: 18 M shared-memory bank conflicts per launch of 4 chains)
    __shared__ int cnt[2][32][33];
    // blockIdx.y = segment * n_wblk + word block (blockIdx.z is the chain of a batched launch)
    const int r = blockIdx.y / n_wblk, wblk = blockIdx.y % n_wblk;
    const int beg = seg_off[r] + blockIdx.x * chunk;
    const int end = min(seg_off[r + 1], beg + chunk);
    if (beg >= end) return;
    const int wc = 1 << wc_log2;
    const int col = threadIdx.x & (wc - 1), rsub = threadIdx.x >> wc_log2, rstep = SS_THREADS >> wc_log2;
    const int w = wblk * 32 + col;
    for (int i = threadIdx.x; i < 2 * 32 * 33; i += SS_THREADS) (&cnt[0][0][0])[i] = 0;
    __syncthreads();
    if (w < W) {
        // Bit-sliced counters: plane k of c1 / c0 holds bit k of the 32 per-bit counts of this
        // thread's word column.  Eight rows go through a carry-save adder tree (Harley-Seal):
        // 7 full adders = 14 LOP3 for eight rows and 32 bit positions, against 8 x 24 shift/mask/add
        // operations when every row is spread into byte-wide counters; the planes are spread into
        // the byte counters once per chunk.  A thread sees at most chunk / rstep <= 127 rows.
        uint32_t c1[7], c0[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) { c1[k] = 0u; c0[k] = 0u; }
        // (h, l) = carry and sum of a + b + c, bit position by bit position
#define SS_CSA(h, l, a, b, c) { const uint32_t u__ = (a) ^ (b); const uint32_t h__ = ((a) & (b)) | (u__ & (c)); l = u__ ^ (c); h = h__; }
        // add a number with the single plane `x` of weight 2^k0 into the counters
#define SS_RIPPLE(cn, k0, x) { uint32_t cy__ = (x); _Pragma("unroll") for (int k = (k0); k < 7; ++k) { const uint32_t t__ = cn[k] & cy__; cn[k] ^= cy__; cy__ = t__; } }
        int i = beg + rsub;
        // eight rows in flight per thread: the row gather (member index, then its words) is the
        // latency of this kernel
        for (; i + 7 * rstep < end; i += 8 * rstep) {
            long long c[8];
            uint32_t p[8], q[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) c[k] = members[i + k * rstep];
#pragma unroll
            for (int k = 0; k < 8; ++k) { p[k] = x1[c[k] * W + w]; q[k] = x0[c[k] * W + w]; }
            uint32_t tA, tB, fA, fB, eC;
            SS_CSA(tA, c1[0], c1[0], p[0], p[1]); SS_CSA(tB, c1[0], c1[0], p[2], p[3]); SS_CSA(fA, c1[1], c1[1], tA, tB);
            SS_CSA(tA, c1[0], c1[0], p[4], p[5]); SS_CSA(tB, c1[0], c1[0], p[6], p[7]); SS_CSA(fB, c1[1], c1[1], tA, tB);
            SS_CSA(eC, c1[2], c1[2], fA, fB);
            SS_RIPPLE(c1, 3, eC);
            SS_CSA(tA, c0[0], c0[0], q[0], q[1]); SS_CSA(tB, c0[0], c0[0], q[2], q[3]); SS_CSA(fA, c0[1], c0[1], tA, tB);
            SS_CSA(tA, c0[0], c0[0], q[4], q[5]); SS_CSA(tB, c0[0], c0[0], q[6], q[7]); SS_CSA(fB, c0[1], c0[1], tA, tB);
            SS_CSA(eC, c0[2], c0[2], fA, fB);
            SS_RIPPLE(c0, 3, eC);
        }
        for (; i < end; i += rstep) {
            const long long cc = members[i];
            const uint32_t p0 = x1[cc * W + w], q0 = x0[cc * W + w];
            SS_RIPPLE(c1, 0, p0);
            SS_RIPPLE(c0, 0, q0);
        }
#undef SS_CSA
#undef SS_RIPPLE
        // planes -> 8 registers of four byte-wide counters each (byte b of a[j] counts bit j + 8 b)
        uint32_t a1[8], a0[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            a1[j] = 0u; a0[j] = 0u;
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                a1[j] += ((c1[k] >> j) & 0x01010101u) << k;
                a0[j] += ((c0[k] >> j) & 0x01010101u) << k;
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int v1 = (a1[j] >> (8 * b)) & 0xff, v0 = (a0[j] >> (8 * b)) & 0xff;
                if (v1) atomicAdd(&cnt[0][col][j + 8 * b], v1);
                if (v0) atomicAdd(&cnt[1][col][j + 8 * b], v0);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * 32; i += SS_THREADS) {
        const int c = i >> 5, bit = i & 31;
        const int m = (wblk * 32 + c) * 32 + bit;
        if (c < wc && m < M) {
            const int v1 = cnt[0][c][bit], v0 = cnt[1][c][bit];
            if (v1) atomicAdd(&S1[(long long)r * M + m], v1);
            if (v0) atomicAdd(&S0[(long long)r * M + m], v0);
        }
    }
}

// =============================================================================
// theta draws and Metropolis-Hastings
// =============================================================================
__device__ __forceinline__ void beta_rows_kernel(const int32_t* __restrict__ S1, const int32_t* __restrict__ S0, int R,
                                 int M, double p, double q, const double* __restrict__ tape,
                                 uint64_t seed, uint64_t stream_id, float* __restrict__ theta_out,
                                 const int32_t* __restrict__ out_ids, double2* __restrict__ lp_out, double FN,
                                 double FP) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)R * M) return;
    const int r = (int)(i / M), m = (int)(i % M);
    double val;
    if (tape) {
        val = tape[i];
    } else {
        PhiloxStream rs(seed, stream_id, (uint64_t)i);
        val = beta_sample(p + (double)S1[i], q + (double)S0[i], rs);
    }
    const long long row = out_ids ? out_ids[r] : r;
    const float th = clip_theta(val);
    theta_out[row * M + m] = th;
    // lp_out (rows in order, [R][M]): the (log p1, log p0) table of the rows just drawn, for the
    // restricted Gibbs scan that follows (what bnpc_logprob_tables would write)
    if (lp_out) {
        double2 v;
        log_p1_p0(th, FN, FP, v.x, v.y);
        lp_out[i] = v;
    }
}

__device__ __forceinline__ void theta_from_uniform_kernel(const double* __restrict__ u, int R, int M,
                                          float* __restrict__ theta_out,
                                          const int32_t* __restrict__ out_ids) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)R * M) return;
    const int r = (int)(i / M), m = (int)(i % M);
    const long long row = out_ids ? out_ids[r] : r;
    theta_out[row * M + m] = clip_theta(u[i]);
}

struct MhConst {
    double FN, FP, p, q, betaln;
    int flat;
};

// log acceptance ratio of libs/CRP.py:347-383 for one (row, mutation)
__device__ __forceinline__ double theta_log_A(float th_new, float th_old, int s1, int s0, double lo,
                                              double hi, double sd, const MhConst& c, bool clip,
                                              double2* lp_new = nullptr, double2* lp_old = nullptr,
                                              const double* fwd_mass = nullptr) {
    const float lo_f = (float)kThetaLo, hi_f = (float)kThetaHi;
    const double lsd = log(sd);
    const double y = (double)(th_new - th_old) / sd;
    // fwd_mass: log mass of [lo, hi] when the caller has it already (the proposal's ppf needs the
    // same value: one of the three Gaussian-mass evaluations of an element saved)
    const double fwd = (fwd_mass ? truncnorm_logpdf_mass(y, lo, hi, *fwd_mass) : truncnorm_logpdf_std(y, lo, hi)) - lsd;
    const double rlo = (double)(lo_f - th_new) / sd, rhi = (double)(hi_f - th_new) / sd;
    const double yr = (double)(th_old - th_new) / sd;
    const double rev = truncnorm_logpdf_std(yr, rlo, rhi) - lsd;
    double n1, n0, o1, o0;
    log_p1_p0(th_new, c.FN, c.FP, n1, n0);
    log_p1_p0(th_old, c.FN, c.FP, o1, o0);
    if (lp_new) { *lp_new = make_double2(n1, n0); *lp_old = make_double2(o1, o0); }
    const double ll_new = (double)s1 * n1 + (double)s0 * n0;
    const double ll_old = (double)s1 * o1 + (double)s0 * o0;
    double pr_new = 0.0, pr_old = 0.0;
    if (!c.flat) {
        pr_new = beta_logpdf((double)th_new, c.p, c.q, c.betaln);
        pr_old = beta_logpdf((double)th_old, c.p, c.q, c.betaln);
    }
    double A = ll_new + pr_new - ll_old - pr_old + rev - fwd;
    if (clip) A = fmin(A, 0.0);
    return A;
}

__device__ __constant__ const double kStepSd[3] = {0.1, 0.25, 0.5};   // libs/CRP.py:65

__device__ __forceinline__ void mh_theta_kernel(float* theta, const int32_t* __restrict__ ids, int R, int M,
                                const int32_t* __restrict__ S1, const int32_t* __restrict__ S0,
                                const double* __restrict__ rnd, MhConst c, int flags,
                                double* __restrict__ logq, int32_t* declined, uint64_t seed, uint64_t stream_id,
                                double2* __restrict__ lp_out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long RM = (long long)R * M;
    if (i >= RM) return;
    const int r = (int)(i / M), m = (int)(i % M);
    const long long row = ids ? ids[r] : r;
    const float old = theta[row * M + m];
    // rnd == NULL: streams stream_id+1 (proposal-sd indices) and +2 (2 RM uniforms)
    const double sd = kStepSd[rnd ? (int)rnd[i] : (int)uniform_at(seed, stream_id + 1, i, 3)];
    const double ut = rnd ? rnd[RM + i] : uniform_at(seed, stream_id + 2, i, 0);
    const double ua = rnd ? rnd[2 * RM + i] : uniform_at(seed, stream_id + 2, RM + i, 0);
    const float lo_f = (float)kThetaLo, hi_f = (float)kThetaHi;
    const double lo = (double)(lo_f - old) / sd, hi = (double)(hi_f - old) / sd;
    const double mass = log_gauss_mass(lo, hi);
    const double x = truncnorm_ppf_mass(ut, lo, hi, mass);
    const float prop = (float)(x * sd + (double)old);
    const bool want_logq = flags & 1;
    double2 lp_new, lp_old;
    const double A = theta_log_A(prop, old, S1[i], S0[i], lo, hi, sd, c, want_logq, &lp_new, &lp_old, &mass);
    const bool rej = log(ua) >= A;
    if (!rej) theta[row * M + m] = prop;
    else atomicAdd(&declined[r], 1);
    // lp_out ([R][M]): the (log p1, log p0) pair of the value the element keeps -- computed above
    // for the acceptance ratio -- so the restricted Gibbs scan that follows reads a finished table
    if (lp_out) lp_out[i] = rej ? lp_old : lp_new;
    if (want_logq) logq[i] = rej ? log(-1.0 * expm1(A)) : A;
}

__device__ __forceinline__ void theta_log_ratio_kernel(const float* __restrict__ th_new, const float* __restrict__ th_old,
                                       int R, int M, const int32_t* __restrict__ S1,
                                       const int32_t* __restrict__ S0, const double* __restrict__ sd_idx,
                                       float blo, float bhi, MhConst c, double* __restrict__ A, uint64_t seed,
                                       uint64_t stream_id) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)R * M) return;
    const float old = th_old[i];
    const double sd = kStepSd[sd_idx ? (int)sd_idx[i] : (int)uniform_at(seed, stream_id, i, 3)];
    const double lo = (double)(blo - old) / sd, hi = (double)(bhi - old) / sd;
    A[i] = theta_log_A(th_new[i], old, S1[i], S0[i], lo, hi, sd, c, true);
}

// =============================================================================
// deterministic reductions over [R][M]
// =============================================================================
#define RL_MAXE 4
struct RlArgs {
    double fn[RL_MAXE], fp[RL_MAXE];
    int E;
    double p, q, betaln;
};
__device__ __forceinline__ void row_loglik_kernel(const float* __restrict__ theta, const int32_t* __restrict__ ids, int R, int M,
                  const int32_t* __restrict__ S1, const int32_t* __restrict__ S0, RlArgs g,
                  double* __restrict__ out, double* __restrict__ prior_out) {
    __shared__ double red[40];
    const int r = blockIdx.x;
    const long long row = ids ? ids[r] : r;
    double acc[RL_MAXE] = {0.0, 0.0, 0.0, 0.0};
    double pr = 0.0;
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
        const float th = theta[row * M + m];
        const double s1 = (double)S1[(long long)r * M + m], s0 = (double)S0[(long long)r * M + m];
#pragma unroll
        for (int e = 0; e < RL_MAXE; ++e) {
            if (e < g.E) {
                double a, b;
                log_p1_p0(th, g.fn[e], g.fp[e], a, b);
                acc[e] += s1 * a + s0 * b;
            }
        }
        if (prior_out) pr += beta_logpdf((double)th, g.p, g.q, g.betaln);
    }
    for (int e = 0; e < g.E; ++e) {
        const double s = block_sum(acc[e], red);
        if (threadIdx.x == 0) out[(long long)e * R + r] = s;
    }
    if (prior_out) {
        const double s = block_sum(pr, red);
        if (threadIdx.x == 0) prior_out[r] = s;
    }
}

__device__ __forceinline__ void row_sum_kernel(const double* __restrict__ v, int R, int M, double* __restrict__ out) {
    __shared__ double red[40];
    const int r = blockIdx.x;
    double acc = 0.0;
    for (int m = threadIdx.x; m < M; m += blockDim.x) acc += v[(long long)r * M + m];
    const double s = block_sum(acc, red);
    if (threadIdx.x == 0) out[r] = s;
}

// =============================================================================
// split-merge support
// =============================================================================
__device__ __forceinline__ void gather_count_kernel(const int32_t* __restrict__ assign, int N, int id_a, int id_b, int32_t* blk) {
    __shared__ int ca, cb;
    if (threadIdx.x == 0) { ca = 0; cb = 0; }
    __syncthreads();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = (n < N) ? assign[n] : -1;
    const unsigned ma = __ballot_sync(FULL, v == id_a && n < N);
    const unsigned mb = __ballot_sync(FULL, id_b >= 0 && v == id_b && n < N);
    if ((threadIdx.x & 31) == 0) {
        if (ma) atomicAdd(&ca, __popc(ma));
        if (mb) atomicAdd(&cb, __popc(mb));
    }
    __syncthreads();
    if (threadIdx.x == 0) { blk[2 * blockIdx.x] = ca; blk[2 * blockIdx.x + 1] = cb; }
}
__device__ __forceinline__ void gather_scan_kernel(int32_t* blk, int nb) {
    // single thread: exclusive scans; the id_b region starts after all of id_a
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int ta = 0;
    for (int b = 0; b < nb; ++b) { const int c = blk[2 * b]; blk[2 * b] = ta; ta += c; }
    int tb = ta;
    for (int b = 0; b < nb; ++b) { const int c = blk[2 * b + 1]; blk[2 * b + 1] = tb; tb += c; }
    blk[2 * nb] = ta;
    blk[2 * nb + 1] = tb;
}
__device__ __forceinline__ void gather_scatter_kernel(const int32_t* __restrict__ assign, int N, int id_a, int id_b,
                      const int32_t* __restrict__ blk, int32_t* __restrict__ cells_out) {
    __shared__ int wa[32], wb[32];
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int v = (n < N) ? assign[n] : -1;
    const bool ia = (n < N) && v == id_a, ib = (n < N) && id_b >= 0 && v == id_b;
    const unsigned ma = __ballot_sync(FULL, ia), mb = __ballot_sync(FULL, ib);
    if (lane == 0) { wa[w] = __popc(ma); wb[w] = __popc(mb); }
    __syncthreads();
    if (w == 0) {
        int a = wa[lane], b = wb[lane];
        int sa = a, sb = b;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int ta = __shfl_up_sync(FULL, sa, o), tb = __shfl_up_sync(FULL, sb, o);
            if (lane >= o) { sa += ta; sb += tb; }
        }
        wa[lane] = sa - a; wb[lane] = sb - b;
    }
    __syncthreads();
    const unsigned lt = (1u << lane) - 1u;
    if (ia) cells_out[blk[2 * blockIdx.x] + wa[w] + __popc(ma & lt)] = n;
    if (ib) cells_out[blk[2 * blockIdx.x + 1] + wb[w] + __popc(mb & lt)] = n;
}

__device__ __forceinline__ void anchor_swaps_kernel(int32_t* cells, int n, int n_a, int idx_i, int idx_j, int is_merge) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (!is_merge) {                       // libs/CRP.py:449-450
        int t = cells[0]; cells[0] = cells[idx_i]; cells[idx_i] = t;
        t = cells[n - 1]; cells[n - 1] = cells[idx_j]; cells[idx_j] = t;
    } else {                               // libs/CRP.py:496,500 (each half swapped on its own)
        int t = cells[0]; cells[0] = cells[idx_i]; cells[idx_i] = t;
        t = cells[n - 1]; cells[n - 1] = cells[n_a + idx_j]; cells[n_a + idx_j] = t;
    }
}

struct K6 { double k[6]; };
__device__ __forceinline__ void rg_launch_halves_kernel(const uint32_t* __restrict__ x1, const uint32_t* __restrict__ x0,
                                        int W, const int32_t* __restrict__ cells, int n, K6 k6,
                                        int32_t* __restrict__ half) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n - 2) return;
    const long long c = cells[s + 1], ci = cells[0], cj = cells[n - 1];
    int ci_[6] = {0, 0, 0, 0, 0, 0}, cj_[6] = {0, 0, 0, 0, 0, 0};
    for (int w = 0; w < W; ++w) {
        const uint32_t s1 = x1[c * W + w], s0 = x0[c * W + w];
        uint32_t a1 = x1[ci * W + w], a0 = x0[ci * W + w], am = ~(a1 | a0);
        ci_[0] += __popc(a1 & s1); ci_[1] += __popc(a1 & s0);
        ci_[2] += __popc(a0 & s1); ci_[3] += __popc(a0 & s0);
        ci_[4] += __popc(am & s1); ci_[5] += __popc(am & s0);
        a1 = x1[cj * W + w]; a0 = x0[cj * W + w]; am = ~(a1 | a0);
        cj_[0] += __popc(a1 & s1); cj_[1] += __popc(a1 & s0);
        cj_[2] += __popc(a0 & s1); cj_[3] += __popc(a0 & s0);
        cj_[4] += __popc(am & s1); cj_[5] += __popc(am & s0);
    }
    double li = 0.0, lj = 0.0;
#pragma unroll
    for (int q = 0; q < 6; ++q) { li += (double)ci_[q] * k6.k[q]; lj += (double)cj_[q] * k6.k[q]; }
    half[s] = (lj > li) ? 1 : 0;
}

__device__ __forceinline__ void rg_count_kernel(const int32_t* __restrict__ half, int nfree, int32_t* seg_off) {
    // seg_off[3] accumulates the number of side-1 free cells; seg_off[4..5] are cursors
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = (s < nfree) ? half[s] : 0;
    const unsigned m = __ballot_sync(FULL, v == 1);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&seg_off[3], __popc(m));
}
__device__ __forceinline__ void rg_sides_kernel(const int32_t* __restrict__ cells, int n, const int32_t* __restrict__ half,
                                int32_t* __restrict__ members, int32_t* seg_off) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int n_i = (n - 2 - seg_off[3]) + 1;
    const int side = (s == 0) ? 0 : ((s == n - 1) ? 1 : half[s - 1]);
    const int pos = atomicAdd(&seg_off[4 + side], 1);
    members[(side ? n_i : 0) + pos] = cells[s];
    if (s == 0) { seg_off[0] = 0; seg_off[1] = n_i; seg_off[2] = n; }
}

// ---- restricted Gibbs scan (libs/CRP.py:609-632 / :806-818) -------------------------------
// The two-way draw of a free cell depends on the scan history only through n_j, the current size
// of side j, and the probability of side j grows with n_j.  So for a given uniform the draw is
// "side j iff n_j - 1 >= tau" for an integer threshold tau that can be found for all cells IN
// PARALLEL (closed form, then corrected with the reference's own arithmetic, rg_side below).
// The sequential part of the scan is then integer-only: ex = ones - own; side = ex >= tau.
__device__ __forceinline__ void rg_pair_logprob(double a0, double a1, int n_i, int n_j, double cn,
                                                double& lp0, double& lp1) {
    // libs/CRP.py:622-623 with _normalize_log (libs/CRP.py:103-116)
    const double p0 = a0 + (log((double)n_i) - cn);
    const double p1 = a1 + (log((double)n_j) - cn);
    const int top = (p1 > p0) ? 1 : 0;
    const double other = (top ? p0 : p1) - (top ? p1 : p0);
    const double lse = log1p(exp(other));
    const double lpt = 0.0 - lse, lpo = other - lse;
    lp0 = top ? lpo : lpt;
    lp1 = top ? lpt : lpo;
}
__device__ __forceinline__ int rg_side(double a0, double a1, int ex, int n, double cn, double u) {
    // np.random.choice([0,1], p=exp(log_probs)) with `ex` other free cells on side j
    const int n_j = ex + 1, n_i = n - n_j - 1;
    double lp0, lp1;
    rg_pair_logprob(a0, a1, n_i, n_j, cn, lp0, lp1);
    const double e0 = exp(lp0), e1 = exp(lp1);
    return (u < e0 / (e0 + e1)) ? 0 : 1;
}

#define RG_FORCE_1 0
#define RG_FORCE_0 (1 << 29)

__device__ __forceinline__ void rg_prepare_kernel(const double* __restrict__ ll2, int ldk, int n,
                                  const int32_t* __restrict__ perm, const double* __restrict__ u,
                                  const int32_t* __restrict__ half, double alpha, int mode,
                                  const int32_t* __restrict__ cells, const int32_t* __restrict__ assign,
                                  int id_i, int32_t* __restrict__ work, uint64_t seed, uint64_t stream_id,
                                  int half_bits) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int nf = n - 2;
    if (s >= nf) return;
    // mode 0 with perm / u == NULL: streams stream_id+1 (order) and +2 (uniforms)
    const int c = (mode == 0) ? (perm ? perm[s] : permutation_at(seed, stream_id + 1, s, nf, half_bits)) : s;
    int tau;
    if (mode != 0) {
        tau = (assign[cells[c + 1]] == id_i) ? RG_FORCE_0 : RG_FORCE_1;
    } else {
        const double a0 = ll2[(long long)c * ldk], a1 = ll2[(long long)c * ldk + 1];
        const double uu = u ? u[s] : uniform_at(seed, stream_id + 2, s, 0);
        const double cn = log((double)n - 1.0 + alpha);
        // side j iff n_j/(n-1) >= sigmoid(logit(1-u) - (a1-a0)); n_j = ex + 1
        const double x = (log1p(-uu) - log(uu)) - (a1 - a0);
        double guess = ceil(((double)n - 1.0) / (1.0 + exp(-x)) - 1.0);
        if (!(guess >= 0.0)) guess = 0.0;
        if (guess > (double)nf) guess = (double)nf;
        tau = (int)guess;
        for (int it = 0; it < 64 && tau > 0 && rg_side(a0, a1, tau - 1, n, cn, uu) == 1; ++it) --tau;
        for (int it = 0; it < 64 && tau < nf && rg_side(a0, a1, tau, n, cn, uu) == 0; ++it) ++tau;
    }
    work[s] = tau * 2 + half[c];
}

#define RG_CHUNK 1024
#define RG_NEUTRAL (RG_FORCE_0 * 2)     /* padding record: the cell is on side 0 and stays there */
// The sequential integer pass x -> x - b + [x - b >= tau] over the cells of a scan, by one warp
// in chunks of 1024 cells.  Lane l walks cells [32 l, 32 l + 32) of the chunk from a GUESS of
// the running count at its first cell and records how far the count could have been off without
// changing any of its 32 decisions (a step is monotone in x, so a segment is a pure shift on that
// interval).  A prefix sum of the segments' net changes then gives every lane its true start as
// long as all lanes before it were inside their intervals; the lanes up to the first one that
// was not are final, the others walk again from the corrected starts.  Every round finalises at
// least one lane (the first open lane starts from an exact value), typically all 32.
__device__ __forceinline__ void rg_serial_kernel(const int32_t* __restrict__ half, int nf, int32_t* __restrict__ work, int out_off) {
    __shared__ int32_t pin[32 * 33], pout[32 * 33];
    const int lane = threadIdx.x;
    int ones = 0;
    for (int s = lane; s < nf; s += 32) ones += half[s];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ones += __shfl_xor_sync(FULL, ones, o);
    int32_t* out = work + out_off;
    int x0 = ones;                                 // the count before the first cell of the chunk
    for (int base = 0; base < nf; base += RG_CHUNK) {
        const int cnt = min(RG_CHUNK, nf - base);
#pragma unroll 8
        for (int i = 0; i < 32; ++i) {             // segment i, position lane (bank-conflict-free both ways)
            const int e = i * 32 + lane;
            pin[i * 33 + lane] = (e < cnt) ? work[base + e] : RG_NEUTRAL;
        }
        __syncwarp();
        int first = 0, guess = x0, start0 = x0, shift = 0;
        int x_end = x0;
        for (int round = 0; round < 33 && first < 32; ++round) {
            int x = guess, up = 0x3fffffff, dn = 0x3fffffff;
            if (lane >= first) {
#pragma unroll 8
                for (int i = 0; i < 32; ++i) {
                    const int v = pin[lane * 33 + i];
                    const int ex = x - (v & 1), tau = v >> 1;
                    const int side = (ex >= tau) ? 1 : 0;
                    if (side) dn = min(dn, ex - tau); else up = min(up, tau - ex - 1);
                    x = ex + side;
                    pout[lane * 33 + i] = ex * 2 + side;
                }
            }
            const int delta = (lane >= first) ? x - guess : 0;
            int incl = delta;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += t;
            }
            const int start = start0 + incl - delta;          // true start if all open lanes before are valid
            const int d = start - guess;
            const bool valid = lane < first || (d >= -dn && d <= up);
            const unsigned bad = __ballot_sync(FULL, !valid);
            const int f = bad ? (__ffs(bad) - 1) : 32;
            if (lane >= first && lane < f) shift = d;          // final: decisions stand, counts move by d
            // the value after the last final lane = the exact start of lane f
            const int end_true = start + delta;                // valid for lanes < f
            const int nxt = __shfl_sync(FULL, end_true, f > 0 ? f - 1 : 0);
            if (f == 32) { x_end = nxt; }
            else {
                if (lane >= f) guess = start;                  // corrected starts (exact for lane f)
                start0 = __shfl_sync(FULL, start, f);
            }
            first = f;
        }
        if (shift != 0) {
#pragma unroll 8
            for (int i = 0; i < 32; ++i) pout[lane * 33 + i] += 2 * shift;
        }
        __syncwarp();
#pragma unroll 8
        for (int i = 0; i < 32; ++i) {
            const int e = i * 32 + lane;
            if (e < cnt) out[base + e] = pout[i * 33 + lane];
        }
        __syncwarp();
        x0 = x_end;
    }
}

__device__ __forceinline__ void rg_finish_kernel(const double* __restrict__ ll2, int ldk, int n,
                                 const int32_t* __restrict__ perm, int32_t* __restrict__ half,
                                 double alpha, int mode, const int32_t* __restrict__ work, int out_off,
                                 double* __restrict__ lq, uint64_t seed, uint64_t stream_id, int half_bits) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int nf = n - 2;
    if (s >= nf) return;
    const int c = (mode == 0) ? (perm ? perm[s] : permutation_at(seed, stream_id + 1, s, nf, half_bits)) : s;
    const int o = work[out_off + s];
    const int side = o & 1, ex = o >> 1;
    half[c] = side;
    if (lq) {
        const double cn = log((double)n - 1.0 + alpha);
        double lp0, lp1;
        rg_pair_logprob(ll2[(long long)c * ldk], ll2[(long long)c * ldk + 1], n - (ex + 1) - 1, ex + 1, cn,
                        lp0, lp1);
        lq[c] = side ? lp1 : lp0;
    }
}

__device__ __forceinline__ void apply_split_kernel(const int32_t* __restrict__ cells, int n, const int32_t* __restrict__ half,
                                   int new_id, int32_t* assign) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const bool mv = (s == n - 1) || (s > 0 && half[s - 1] == 1);
    if (mv) assign[cells[s]] = new_id;
}
__device__ __forceinline__ void apply_merge_kernel(const int32_t* __restrict__ cells, int n_a, int n, int id, int32_t* assign) {
    const int s = n_a + blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) assign[cells[s]] = id;
}

// =============================================================================
// C ABI
// =============================================================================
extern "C" {

int bnpc_abi_version(void) { return BNPC_ABI_VERSION; }
const char* bnpc_last_error(void) { return g_err; }
int64_t bnpc_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

int bnpc_pack_planes(const double* x_f64, const int8_t* x_i8, int N, int M, int W, uint32_t* x1,
                     uint32_t* x0, int32_t* n1, int32_t* n0, void* stream) {
    if ((x_f64 == nullptr) == (x_i8 == nullptr)) return bad_arg("exactly one of x_f64 / x_i8");
    if (W % 4 != 0 || W * 32 < M) return bad_arg("W must be a multiple of 4 with 32*W >= M");
    if (N <= 0) return 0;
    const int blocks = (int)min((long long)cdiv((long long)N * 32, 256), 148ll * 64);
    BNPC_LAUNCH(pack_planes_kernel, 0, 0, blocks, 256, 0, (cudaStream_t)stream, x_f64, x_i8, N, M, W, x1, x0, n1, n0, blocks);
    return 0;
}

int bnpc_fill_uniform(double* out, int64_t n, uint64_t seed, uint64_t stream_id, int n_levels,
                      void* stream) {
    if (n <= 0) return 0;
    const long long pairs = (n + 1) / 2;
    BNPC_LAUNCH(fill_uniform_kernel, 0, 0, cdiv(pairs, 256), 256, 0, (cudaStream_t)stream, out, n, seed, stream_id, n_levels);
    return 0;
}

int bnpc_fill_permutation(int32_t* out, int n, uint64_t seed, uint64_t stream_id, void* stream) {
    if (n <= 0) return 0;
    BNPC_LAUNCH(fill_permutation_kernel, 0, 0, cdiv(n, 256), 256, 0, (cudaStream_t)stream, out, n, seed, stream_id, feistel_half_bits(n));
    return 0;
}

int bnpc_logprob_tables(const float* theta, const int32_t* ids, int R, int M, double FN, double FP,
                        double* lp, void* stream) {
    if (R <= 0) return 0;
    BNPC_LAUNCH(logprob_tables_kernel, 0, 0, cdiv((long long)R * M, 256), 256, 0, (cudaStream_t)stream,  theta, ids, R, M, FN, FP, reinterpret_cast<double2*>(lp));
    return 0;
}

static bool ll_few_fits(int K, int W) { return K <= 2 && sizeof(double2) * (size_t)K * W * 32 <= 200 * 1024; }

// the K <= 2 rows of a restricted Gibbs scan straight from theta (libs/CRP.py:635-638): the
// log-probability table is built inside the kernel
static int ll_few_from_theta(const uint32_t* x1, const uint32_t* x0, int W, int M, const int32_t* cells,
                             int cell_stride, int C, const float* theta, int K, double FN, double FP, double* ll,
                             int ldk, void* stream) {
    if (C <= 0 || K <= 0) return 0;
    const size_t smem = sizeof(double2) * (size_t)K * (W * 32 + 1);
    BNPC_LAUNCH(ll_few_kernel, 8 * LLP_CELLS, 0, cdiv(C, LLP_CELLS), 8 * LLP_CELLS, smem, (cudaStream_t)stream,  x1, x0, W, M, cells, cell_stride, C, (const double2*)nullptr, K, ll, ldk, theta, FN, FP);
    return 0;
}

int bnpc_ll_matrix(const uint32_t* x1, const uint32_t* x0, int W, int M, const int32_t* cells,
                   int cell_stride, int C, const double* lp, int K, double* ll, int ldk, void* stream) {
    if (C <= 0 || K <= 0) return 0;
    if (ldk < K) return bad_arg("ldk < K");
    if (W % 4 != 0) return bad_arg("W must be a multiple of 4");
    if (ll_few_fits(K, W)) {
        const size_t smem = sizeof(double2) * (size_t)K * (W * 32 + 1);
        BNPC_LAUNCH(ll_few_kernel, 8 * LLP_CELLS, 0, cdiv(C, LLP_CELLS), 8 * LLP_CELLS, smem, (cudaStream_t)stream,  x1, x0, W, M, cells, cell_stride, C, reinterpret_cast<const double2*>(lp), K, ll, ldk, (const float*)nullptr, 0.0, 0.0);
        return 0;
    }
    dim3 grid(cdiv(C, LL_THREADS), cdiv(K, LL_KT));
    if (grid.y > 65535) return bad_arg("too many clusters for one ll_matrix launch");
    BNPC_LAUNCH(ll_matrix_kernel, LL_THREADS, 0, grid, LL_THREADS, 0, (cudaStream_t)stream,  x1, x0, W, M, cells, cell_stride, C, reinterpret_cast<const double2*>(lp), K, ll, ldk);
    return 0;
}

// perm / u == NULL: the kernel draws the visiting order and the uniforms itself (streams
// stream_id+1, +2 of seed)
static int gibbs_prepare_impl(const int32_t* perm, const double* u, const int32_t* assign, const int32_t* n1,
                              const int32_t* n0, int N, double c1, double c0, double lnew_prior,
                              bnpc_visit_t* visit, uint64_t seed, uint64_t stream_id, void* stream) {
    if (N <= 0) return 0;
    BNPC_LAUNCH(gibbs_prepare_kernel, 0, 0, cdiv(N, 256), 256, 0, (cudaStream_t)stream, perm, u, assign, n1, n0, N, c1, c0, lnew_prior, visit, seed, stream_id, feistel_half_bits(N));
    return 0;
}
int bnpc_gibbs_prepare(const int32_t* perm, const double* u, const int32_t* assign, const int32_t* n1,
                       const int32_t* n0, int N, double c1, double c0, double lnew_prior,
                       bnpc_visit_t* visit, void* stream) {
    if (!perm || !u) return bad_arg("perm/u");
    return gibbs_prepare_impl(perm, u, assign, n1, n0, N, c1, c0, lnew_prior, visit, 0, 0, stream);
}

int bnpc_gibbs_candidates(const double* ll, int ldk, int K, const int32_t* col_of_id,
                          bnpc_visit_t* visit_t0, bnpc_cand_t* cand_t0, int C, double slack,
                          double c_norm, int32_t* blk, void* stream) {
    if (C <= 0) return 0;
    if (K > SW_MAXL) return bad_arg("candidates need K <= 1024");
    if (!blk) return bad_arg("blk");
    BNPC_LAUNCH(gibbs_candidates_kernel, CAND_THREADS, 0, cdiv(C, CAND_THREADS), CAND_THREADS, 0, (cudaStream_t)stream,  ll, ldk, K, col_of_id, visit_t0, cand_t0, C, slack, c_norm, blk);
    return 0;
}

int bnpc_gibbs_compact(const bnpc_visit_t* visit_t0, const bnpc_cand_t* cand_t0, int C, int32_t* blk,
                       bnpc_visit_t* visit_c, bnpc_cand_t* cand_c, int32_t* st, void* stream) {
    if (C <= 0) return 0;
    if (!blk || !visit_c || !cand_c || !st) return bad_arg("compact buffers");
    const int nb = cdiv(C, CAND_THREADS);
    BNPC_LAUNCH(compact_scan_kernel, 1024, 0, 1, 1024, 0, (cudaStream_t)stream, blk, nb, st);
    BNPC_LAUNCH(compact_scatter_kernel, CAND_THREADS, 0, nb, CAND_THREADS, 0, (cudaStream_t)stream, visit_t0, cand_t0, C, blk, visit_c, cand_c);
    return 0;
}

int bnpc_ll_matrix_f32(const uint32_t* x1, const uint32_t* x0, int W, int M, const int32_t* cells,
                       int cell_stride, int C, const double* lp, float* lpf, int K, float* llf, int ldf,
                       void* stream) {
    if (C <= 0 || K <= 0) return 0;
    if (ldf < K) return bad_arg("ldf < K");
    if (W % 4 != 0) return bad_arg("W must be a multiple of 4");
    const long long n = (long long)K * M;
    BNPC_LAUNCH(lp_to_f32_kernel, 0, 0, cdiv(n, 256), 256, 0, (cudaStream_t)stream, reinterpret_cast<const double2*>(lp), n, reinterpret_cast<float2*>(lpf));
    dim3 grid(cdiv(C, LL_THREADS), cdiv(K, LLF_KT));
    BNPC_LAUNCH(ll_matrix_f32_kernel, LL_THREADS, 0, grid, LL_THREADS, 0, (cudaStream_t)stream,  x1, x0, W, M, cells, cell_stride, C, reinterpret_cast<const float2*>(lpf), K, llf, ldf);
    return 0;
}

int bnpc_ll_matrix_tc(const uint32_t* x1, const uint32_t* x0, int W, int M, const int32_t* cells,
                      int cell_stride, int C, const double* lp, uint16_t* bsplit, int K, float* llf, int ldf,
                      void* stream) {
    if (C <= 0 || K <= 0) return 0;
    if (K > BNPC_LEAN_MAXK) return bad_arg("tensor-core rows need K <= BNPC_LEAN_MAXK");
    if (W % 4 != 0) return bad_arg("W must be a multiple of 4");
    const int kpad = (K + 7) & ~7;
    if (ldf < kpad || ldf % 4 != 0) return bad_arg("ldf must be a multiple of 4, >= K rounded up to 8");
    cudaStream_t s = (cudaStream_t)stream;
    const long long total = (long long)W * 2 * kpad * 64;
    BNPC_LAUNCH(lp_split_bf16_kernel, 0, 0, cdiv(total, 256), 256, 0, s, reinterpret_cast<const double2*>(lp), K, M, W, kpad, bsplit);
    switch (kpad) {
        case 8: return launch_ll_tc<8>(x1, x0, W, cells, cell_stride, C, bsplit, llf, ldf, s);
        case 16: return launch_ll_tc<16>(x1, x0, W, cells, cell_stride, C, bsplit, llf, ldf, s);
        case 24: return launch_ll_tc<24>(x1, x0, W, cells, cell_stride, C, bsplit, llf, ldf, s);
        case 32: return launch_ll_tc<32>(x1, x0, W, cells, cell_stride, C, bsplit, llf, ldf, s);
        case 40: return launch_ll_tc<40>(x1, x0, W, cells, cell_stride, C, bsplit, llf, ldf, s);
        case 48: return launch_ll_tc<48>(x1, x0, W, cells, cell_stride, C, bsplit, llf, ldf, s);
        case 56: return launch_ll_tc<56>(x1, x0, W, cells, cell_stride, C, bsplit, llf, ldf, s);
        default: return launch_ll_tc<64>(x1, x0, W, cells, cell_stride, C, bsplit, llf, ldf, s);
    }
}

int bnpc_cocluster_counts(const int32_t* assign, int S, int N, int32_t* counts, void* stream) {
    if (S <= 0 || N < 2) return bad_arg("cocluster_counts needs S > 0 samples of N >= 2 cells");
    const int T = cdiv(N, EST_TILE);
    if (T > 65535) return bad_arg("too many cells for one launch");
    dim3 grid(T, T);
    BNPC_LAUNCH(cocluster_counts_kernel, 256, 0, grid, 256, 0, (cudaStream_t)stream, assign, S, N, counts);
    return 0;
}

int bnpc_mpear_sums(const int32_t* counts, int N, const int32_t* labels, int n_cand, unsigned long long* out,
                    void* stream) {
    return bnpc_mpear_sums_weighted(counts, N, labels, n_cand, nullptr, out, stream);
}

int bnpc_mpear_sums_weighted(const int32_t* counts, int N, const int32_t* labels, int n_cand, const int32_t* weight,
                             unsigned long long* out, void* stream) {
    if (N < 2 || n_cand < 0) return bad_arg("mpear_sums needs N >= 2");
    const int T = cdiv(N, EST_TILE);
    if (T > 65535) return bad_arg("too many cells for one launch");
    cudaError_t ce = cudaMemsetAsync(out, 0, sizeof(unsigned long long) * (1 + 2 * (size_t)n_cand),
                                     (cudaStream_t)stream);
    if (ce != cudaSuccess) return fail("mpear_sums memset", ce);
    dim3 grid(T, T);
    BNPC_LAUNCH(mpear_sums_kernel, 256, 0, grid, 256, 0, (cudaStream_t)stream, counts, N, labels, n_cand, out, weight);
    return 0;
}

int bnpc_debug_set_trace(void* buf) {
    g_t8_trace = reinterpret_cast<long long*>(buf);
    return 0;
}

int bnpc_ll_matrix_i8(const uint32_t* x1, const uint32_t* x0, int W, int M, const int32_t* cells,
                      int cell_stride, int C, const double* lp, uint8_t* bdigits, int K, double vmax,
                      float* llf, int ldf, void* stream) {
    if (C <= 0 || K <= 0) return 0;
    if (K > BNPC_LEAN_MAXK) return bad_arg("tensor-core rows need K <= BNPC_LEAN_MAXK");
    if (W % 4 != 0) return bad_arg("W must be a multiple of 4");
    if (!(vmax > 0.0)) return bad_arg("vmax must be positive");
    if (M >= 32768) return bad_arg("int32 accumulators hold rows of fewer than 32768 mutations");
    const int kpad = (K + 7) & ~7;
    if (ldf < kpad || ldf % 4 != 0) return bad_arg("ldf must be a multiple of 4, >= K rounded up to 8");
    cudaStream_t s = (cudaStream_t)stream;
    const double q = vmax / 65535.0;
    const long long total = (long long)(W / 2) * 2 * kpad * 128;
    BNPC_LAUNCH(lp_split_u8_kernel, 0, 0, cdiv(total, 256), 256, 0, s, reinterpret_cast<const double2*>(lp), K, M, W, kpad, 1.0 / q, bdigits);
    const float nq = -(float)q;
    // rows in cell order (cells = NULL): the kernel that lets chains with equal tiles share the
    // expanded data operand (bnpc_tc_i8s.cuh); BNPC_LL_I8_SINGLE=1 keeps the one-chain kernel of
    // bnpc_tc_i8.cuh, which also serves the gathered rows of later epochs
    static const int single = getenv("BNPC_LL_I8_SINGLE") ? 1 : 0;
    if (!single && cells == nullptr) return launch_ll_i8s(x1, x0, W, C, bdigits, kpad, nq, llf, ldf, s);
    switch (kpad) {
        case 8: return launch_ll_i8<8>(x1, x0, W, cells, cell_stride, C, bdigits, nq, llf, ldf, s);
        case 16: return launch_ll_i8<16>(x1, x0, W, cells, cell_stride, C, bdigits, nq, llf, ldf, s);
        case 24: return launch_ll_i8<24>(x1, x0, W, cells, cell_stride, C, bdigits, nq, llf, ldf, s);
        case 32: return launch_ll_i8<32>(x1, x0, W, cells, cell_stride, C, bdigits, nq, llf, ldf, s);
        case 40: return launch_ll_i8<40>(x1, x0, W, cells, cell_stride, C, bdigits, nq, llf, ldf, s);
        case 48: return launch_ll_i8<48>(x1, x0, W, cells, cell_stride, C, bdigits, nq, llf, ldf, s);
        case 56: return launch_ll_i8<56>(x1, x0, W, cells, cell_stride, C, bdigits, nq, llf, ldf, s);
        default: return launch_ll_i8<64>(x1, x0, W, cells, cell_stride, C, bdigits, nq, llf, ldf, s);
    }
}

int bnpc_ll_shared_plan(int n, const int* kpad, int C, int W, int* group_of, int* column_of, int* tiles_per_supertile,
                        int* ctas_of_group, int* slots_of_group, int* n_groups) {
    if (n <= 0 || n > bnpc::BATCH_MAX) return bad_arg("1..8 chains");
    if (C <= 0 || W <= 0 || W % 4 != 0) return bad_arg("C > 0, W a positive multiple of 4");
    ll_shared_t ones[bnpc::BATCH_MAX];
    memset(ones, 0, sizeof(ones));
    for (int i = 0; i < n; ++i) {
        if (kpad[i] <= 0 || kpad[i] % 8 != 0 || kpad[i] > BNPC_LEAN_MAXK) return bad_arg("kpad: a multiple of 8 in [8, 64]");
        ones[i].W = W; ones[i].C = C; ones[i].nc = 1; ones[i].n_tot = 2 * kpad[i];
        ones[i].ch[0].kpad = kpad[i];
        ones[i].ch[0].ldf = i;                      // tag: which chain this block is
    }
    ll_plan_t plan;
    ll_shared_plan(ones, n, false, plan);
    for (int g = 0; g < plan.ng; ++g) {
        ctas_of_group[g] = plan.groups[g].n_ctas;
        slots_of_group[g] = plan.groups[g].b_slots;
        for (int c = 0; c < plan.groups[g].nc; ++c) {
            group_of[plan.groups[g].ch[c].ldf] = g;
            column_of[plan.groups[g].ch[c].ldf] = plan.groups[g].ch[c].off;
        }
    }
    *tiles_per_supertile = plan.T;
    *n_groups = plan.ng;
    return 0;
}

int bnpc_ll_matrix_i8_shared(const uint32_t* x1, const uint32_t* x0, int W, int M, int C, int n_chains,
                             const double* const* lp, uint8_t* const* bdigits, const int* K, const double* vmax,
                             float* const* llf, const int* ldf, void* stream) {
    if (C <= 0 || n_chains <= 0) return 0;
    if (n_chains > bnpc::BATCH_MAX) return bad_arg("at most 8 chains per call");
    if (bnpc::g_rec.on) return bad_arg("bnpc_ll_matrix_i8_shared launches at once (no recorder)");
    if (W % 4 != 0) return bad_arg("W must be a multiple of 4");
    if (M >= 32768) return bad_arg("int32 accumulators hold rows of fewer than 32768 mutations");
    cudaStream_t s = (cudaStream_t)stream;
    bnpc::Op op[bnpc::BATCH_MAX];
    bnpc::Op* ops[bnpc::BATCH_MAX];
    for (int c = 0; c < n_chains; ++c) {
        if (K[c] <= 0 || K[c] > BNPC_LEAN_MAXK) return bad_arg("tensor-core rows need 0 < K <= BNPC_LEAN_MAXK");
        if (!(vmax[c] > 0.0)) return bad_arg("vmax must be positive");
        const int kpad = (K[c] + 7) & ~7;
        if (ldf[c] < kpad || ldf[c] % 4 != 0) return bad_arg("ldf must be a multiple of 4, >= K rounded up to 8");
        const double q = vmax[c] / 65535.0;
        const long long total = (long long)(W / 2) * 2 * kpad * 128;
        BNPC_LAUNCH(lp_split_u8_kernel, 0, 0, cdiv(total, 256), 256, 0, s, reinterpret_cast<const double2*>(lp[c]), K[c], M, W, kpad, 1.0 / q, bdigits[c]);
        ll_shared_t one;
        memset(&one, 0, sizeof(one));
        one.x1 = x1; one.x0 = x0; one.W = W; one.C = C;
        one.nc = 1; one.n_tot = 2 * kpad;
        one.ch[0].Bg = bdigits[c]; one.ch[0].llf = llf[c]; one.ch[0].neg_q = -(float)q; one.ch[0].ldf = ldf[c];
        one.ch[0].kpad = kpad; one.ch[0].off = 0;
        op[c].kind = 0; op[c].name = "ll_matrix_i8s_kernel"; op[c].block = T8S_THREADS;
        memcpy(op[c].args, &one, sizeof(one));
        ops[c] = &op[c];
    }
    return launch_ll_shared_merged(ops, n_chains, s);
}

// clear = false: the caller has zeroed n_cert already (the composite entry points clear all the
// scratch of an epoch in one launch)
static int gibbs_options_impl(const float* llf, int ldf, int K, const int32_t* col_of_id,
                              const bnpc_visit_t* visit_t0, bnpc_opt_t* opt_t0, int32_t* n_cert, int C,
                              double log_n, double c_norm, int terms, double err_abs, bool clear, void* stream,
                              int by_cell = 0) {
    if (C <= 0) return 0;
    if (K <= 0 || K > BNPC_LEAN_MAXK) return bad_arg("lean epochs need K <= BNPC_LEAN_MAXK");
    if (ldf % 4 != 0 || ldf < ((K + 3) & ~3) || ((uintptr_t)llf & 15)) return bad_arg("llf rows: 16-byte aligned, ldf a multiple of 4 >= K");
    if (clear)
        if (int rc = zero_async(n_cert, sizeof(int32_t) * BNPC_LEAN_MAXK, stream, "gibbs_options memset")) return rc;
    const float err_rel = (float)terms * 2.384185791015625e-07f;      // terms * 2^-22
    BNPC_LAUNCH(gibbs_options_kernel, CAND_THREADS, 0, cdiv(C, CAND_THREADS), CAND_THREADS, 0, (cudaStream_t)stream,  llf, ldf, K, col_of_id, visit_t0, opt_t0, n_cert, C, (float)log_n, c_norm, err_rel, 0.05f + (float)(err_abs > 0.0 ? err_abs : 0.0), by_cell);
    return 0;
}

int bnpc_gibbs_options(const float* llf, int ldf, int K, const int32_t* col_of_id,
                       const bnpc_visit_t* visit_t0, bnpc_opt_t* opt_t0, int32_t* n_cert, int C,
                       double log_n, double c_norm, int terms, double err_abs, void* stream) {
    return gibbs_options_impl(llf, ldf, K, col_of_id, visit_t0, opt_t0, n_cert, C, log_n, c_norm, terms, err_abs, true,
                              stream);
}

static int gibbs_exact_impl(const uint32_t* x1, const uint32_t* x0, int W, int M, const double* lp, int K,
                            const bnpc_visit_t* visit_t0, bnpc_opt_t* opt_t0, const int32_t* n_cert, int C,
                            int32_t* blk, int32_t* idx_c, int32_t* st, bnpc_visit_t* visit_c,
                            bnpc_cand_t* cand_c, double log_n, double c_norm, int32_t* comp, int32_t* order,
                            bool clear, void* stream) {
    if (C <= 0) return 0;
    if (K <= 0 || K > BNPC_LEAN_MAXK) return bad_arg("lean epochs need K <= BNPC_LEAN_MAXK");
    if (!comp || !order) return bad_arg("comp/order");
    const int nb = cdiv(C, CAND_THREADS);
    cudaStream_t s = (cudaStream_t)stream;
    if (clear)
        if (int rc = zero_async(comp, sizeof(int32_t) * 512, stream, "gibbs_exact memset")) return rc;
    BNPC_LAUNCH(gibbs_finalize_kernel, CAND_THREADS, 0, nb, CAND_THREADS, 0, s, opt_t0, n_cert, C, blk);
    BNPC_LAUNCH(compact_scan_kernel, 1024, 0, 1, 1024, 0, s, blk, nb, st);
    BNPC_LAUNCH(compact_index_kernel, CAND_THREADS, 0, nb, CAND_THREADS, 0, s, opt_t0, C, blk, idx_c);
    // default: stage only the columns a CTA's visits use (32 KB of shared memory whatever K is;
    // Gibbs step 2.50 -> 2.09 ms at K = 59 with every visit uncertain); BNPC_EXACT_STAGING=all
    // selects the kernel that stages all K columns per round
    static int all_cols = -1;
    if (all_cols < 0) {
        const char* env = getenv("BNPC_EXACT_STAGING");
        all_cols = (env && env[0] == 'a') ? 1 : 0;
    }
    const size_t smem = all_cols ? sizeof(double2) * 32 * EX_WORDS * (size_t)K : sizeof(double2) * (EX_ROWS + 1) * (EX_COLS + 1);
    // processing order: uncertain visits grouped by their own cluster (see exact_hist_kernel)
    BNPC_LAUNCH(exact_hist_kernel, 256, 0, cdiv(C, 256), 256, 0, s, opt_t0, idx_c, st, comp);
    BNPC_LAUNCH(exact_scan_kernel, BNPC_LEAN_MAXK, 0, 1, BNPC_LEAN_MAXK, 0, s, comp);
    BNPC_LAUNCH(exact_scatter_kernel, 256, 0, cdiv(C, 256), 256, 0, s, opt_t0, idx_c, st, comp, order);
    // the number of uncertain visits lives on the device: blocks beyond it exit at once
    if (all_cols)
        BNPC_LAUNCH(gibbs_exact_allcols_kernel, EX_THREADS, 0, cdiv(C, EX_THREADS), EX_THREADS, smem, s,  x1, x0, W, M, reinterpret_cast<const double2*>(lp), K, visit_t0, opt_t0, idx_c, st, visit_c, cand_c, log_n, c_norm, comp, order);
    else
        BNPC_LAUNCH(gibbs_exact_kernel, EX_THREADS, 0, cdiv(C, EX_THREADS), EX_THREADS, smem, s,  x1, x0, W, M, reinterpret_cast<const double2*>(lp), K, visit_t0, opt_t0, idx_c, st, visit_c, cand_c, log_n, c_norm, comp, order);
    BNPC_LAUNCH(components_kernel, BNPC_LEAN_MAXK, 0, 1, BNPC_LEAN_MAXK, 0, s, comp, K, SW_PAR_WARPS);
    // (the processing order is not needed any more: its buffer receives the owner bytes)
    BNPC_LAUNCH(owner_bytes_kernel, 256, 0, cdiv(C, 256), 256, 0, s, visit_c, st, comp, reinterpret_cast<uint8_t*>(order));
    return 0;
}

int bnpc_gibbs_exact(const uint32_t* x1, const uint32_t* x0, int W, int M, const double* lp, int K,
                     const bnpc_visit_t* visit_t0, bnpc_opt_t* opt_t0, const int32_t* n_cert, int C,
                     int32_t* blk, int32_t* idx_c, int32_t* st, bnpc_visit_t* visit_c,
                     bnpc_cand_t* cand_c, double log_n, double c_norm, int32_t* comp, int32_t* order,
                     void* stream) {
    return gibbs_exact_impl(x1, x0, W, M, lp, K, visit_t0, opt_t0, n_cert, C, blk, idx_c, st, visit_c, cand_c, log_n,
                            c_norm, comp, order, true, stream);
}

int bnpc_gibbs_epoch_begin(const int32_t* live, int K, int32_t* lst, int32_t* cnt, int32_t* col_of_id,
                           int idcap, int32_t* st, int first, void* stream) {
    if (K > idcap) return bad_arg("K > idcap");
    BNPC_LAUNCH(gibbs_epoch_begin_kernel, 0, 0, 1, 1024, 0, (cudaStream_t)stream, live, K, lst, cnt, col_of_id, idcap, st, first);
    return 0;
}

int bnpc_gibbs_sweep(const bnpc_sweep_args_t* a, int block_threads, void* stream) {
    if (!a) return bad_arg("args");
    if (block_threads < 32 || block_threads > 1024 || block_threads % 32) return bad_arg("block_threads");
    if (a->t_begin < a->t_epoch0 || a->t_end - a->t_epoch0 > a->ldx) return bad_arg("sweep range vs ldx");
    // 256 threads leave the full register budget to the sequencer warp; long lists want 1024
    const size_t smem = sizeof(SweepShared);
    if (block_threads <= 256)
        BNPC_LAUNCH(gibbs_sweep_kernel<256>, 256, 1, 1, block_threads, smem, (cudaStream_t)stream, *a);
    else if (block_threads <= 512)
        BNPC_LAUNCH(gibbs_sweep_kernel<512>, 512, 1, 1, block_threads, smem, (cudaStream_t)stream, *a);
    else
        BNPC_LAUNCH(gibbs_sweep_kernel<1024>, 1024, 1, 1, block_threads, smem, (cudaStream_t)stream, *a);
    return 0;
}

int bnpc_set_ranks(const int32_t* ids, int K, int32_t* rank_of_id, void* stream) {
    if (K <= 0) return 0;
    BNPC_LAUNCH(set_ranks_kernel, 0, 0, cdiv(K, 256), 256, 0, (cudaStream_t)stream, ids, K, rank_of_id);
    return 0;
}

static int group_members_impl(const int32_t* assign, int N, const int32_t* rank_of_id, const int32_t* seg_off,
                              int32_t* cursor, int K, int32_t* members, bool clear, void* stream) {
    if (N <= 0) return 0;
    if (clear)
        if (int rc = zero_async(cursor, sizeof(int32_t) * (size_t)K, stream, "group_members memset")) return rc;
    BNPC_LAUNCH(group_members_kernel, 0, 0, cdiv(N, 256), 256, 0, (cudaStream_t)stream, assign, N, rank_of_id, seg_off, cursor, members);
    return 0;
}

int bnpc_group_members(const int32_t* assign, int N, const int32_t* rank_of_id, const int32_t* seg_off,
                       int32_t* cursor, int K, int32_t* members, void* stream) {
    return group_members_impl(assign, N, rank_of_id, seg_off, cursor, K, members, true, stream);
}

static int suffstat_impl(const uint32_t* x1, const uint32_t* x0, int W, int M, const int32_t* members,
                         const int32_t* seg_off, int R, int max_len, int32_t* S1, int32_t* S0, bool clear,
                         void* stream) {
    if (R <= 0) return 0;
    if (clear) {
        if (int rc = zero_async(S1, sizeof(int32_t) * (size_t)R * M, stream, "suffstat memset")) return rc;
        if (int rc = zero_async(S0, sizeof(int32_t) * (size_t)R * M, stream, "suffstat memset")) return rc;
    }
    if (max_len <= 0) return 0;
    int wc_log2 = 2;                       // word columns per CTA row group: 4, 8, 16 or 32
    while ((1 << wc_log2) < W && wc_log2 < 5) ++wc_log2;
    // grid.y = segments x word blocks, limited to 65535: tile the segment axis
    const int n_wblk = cdiv(W, 32);
    const int r_max = 65535 / n_wblk;
    // rows per CTA: SS_CHUNK for the statistics of all clusters; the two sides of a restricted Gibbs
    // scan are a few thousand rows -- a quarter of the chunk gives four times the CTAs to a kernel
    // that is bound by the latency of its row gathers (ncu: 40 CTAs, 12 % of the warp slots)
    const int chunk = ((long long)max_len * R <= 65536) ? SS_CHUNK / 4 : SS_CHUNK;
    for (int r0 = 0; r0 < R; r0 += r_max) {
        const int rr = min(r_max, R - r0);
        dim3 grid(cdiv(max_len, chunk), rr * n_wblk);
        BNPC_LAUNCH(suffstat_kernel, SS_THREADS, 0, grid, SS_THREADS, 0, (cudaStream_t)stream, x1, x0, W, M, members, seg_off + r0, S1 + (size_t)r0 * M, S0 + (size_t)r0 * M, wc_log2, n_wblk, chunk);
    }
    return 0;
}

int bnpc_suffstat(const uint32_t* x1, const uint32_t* x0, int W, int M, const int32_t* members,
                  const int32_t* seg_off, int R, int max_len, int32_t* S1, int32_t* S0, void* stream) {
    return suffstat_impl(x1, x0, W, M, members, seg_off, R, max_len, S1, S0, true, stream);
}

static int beta_rows_impl(const int32_t* S1, const int32_t* S0, int R, int M, double p, double q,
                          const double* tape, uint64_t seed, uint64_t stream_id, float* theta_out,
                          const int32_t* out_ids, double* lp_out, double FN, double FP, void* stream) {
    if (R <= 0) return 0;
    BNPC_LAUNCH(beta_rows_kernel, 0, 0, cdiv((long long)R * M, 128), 128, 0, (cudaStream_t)stream,  S1, S0, R, M, p, q, tape, seed, stream_id, theta_out, out_ids, reinterpret_cast<double2*>(lp_out), FN, FP);
    return 0;
}
int bnpc_beta_rows(const int32_t* S1, const int32_t* S0, int R, int M, double p, double q,
                   const double* tape, uint64_t seed, uint64_t stream_id, float* theta_out,
                   const int32_t* out_ids, void* stream) {
    return beta_rows_impl(S1, S0, R, M, p, q, tape, seed, stream_id, theta_out, out_ids, nullptr, 0.0, 0.0, stream);
}

int bnpc_theta_from_uniform(const double* u, int R, int M, float* theta_out, const int32_t* out_ids,
                            void* stream) {
    if (R <= 0) return 0;
    BNPC_LAUNCH(theta_from_uniform_kernel, 0, 0, cdiv((long long)R * M, 256), 256, 0, (cudaStream_t)stream, u, R, M, theta_out, out_ids);
    return 0;
}

static MhConst make_mh_const(double FN, double FP, double p, double q) {
    MhConst c;
    c.FN = FN; c.FP = FP; c.p = p; c.q = q;
    c.flat = (p == 1.0 && q == 1.0);
    c.betaln = lgamma(p) + lgamma(q) - lgamma(p + q);
    return c;
}

// rnd == NULL: the kernel draws from streams stream_id+1, +2 of seed
static int mh_theta_impl(float* theta, const int32_t* ids, int R, int M, const int32_t* S1, const int32_t* S0,
                         const double* rnd, uint64_t seed, uint64_t stream_id, double FN, double FP, double p,
                         double q, int flags, double* logq, int32_t* declined, void* stream,
                         double* lp_out = nullptr) {
    if (R <= 0) return 0;
    if ((flags & 1) && !logq) return bad_arg("logq required when flags&1");
    BNPC_LAUNCH(mh_theta_kernel, 0, 0, cdiv((long long)R * M, 128), 128, 0, (cudaStream_t)stream,  theta, ids, R, M, S1, S0, rnd, make_mh_const(FN, FP, p, q), flags, logq, declined, seed, stream_id, reinterpret_cast<double2*>(lp_out));
    return 0;
}
int bnpc_mh_theta(float* theta, const int32_t* ids, int R, int M, const int32_t* S1, const int32_t* S0,
                  const double* rnd, double FN, double FP, double p, double q, int flags, double* logq,
                  int32_t* declined, void* stream) {
    if (!rnd) return bad_arg("rnd");
    return mh_theta_impl(theta, ids, R, M, S1, S0, rnd, 0, 0, FN, FP, p, q, flags, logq, declined, stream);
}

// sd_idx == NULL: proposal-sd indices from stream stream_id of seed
static int theta_log_ratio_impl(const float* th_new, const float* th_old, int R, int M, const int32_t* S1,
                                const int32_t* S0, const double* sd_idx, uint64_t seed, uint64_t stream_id,
                                float blo, float bhi, double FN, double FP, double p, double q, double* A,
                                void* stream) {
    if (R <= 0) return 0;
    BNPC_LAUNCH(theta_log_ratio_kernel, 0, 0, cdiv((long long)R * M, 128), 128, 0, (cudaStream_t)stream,  th_new, th_old, R, M, S1, S0, sd_idx, blo, bhi, make_mh_const(FN, FP, p, q), A, seed, stream_id);
    return 0;
}
int bnpc_theta_log_ratio(const float* th_new, const float* th_old, int R, int M, const int32_t* S1,
                         const int32_t* S0, const double* sd_idx, float blo, float bhi, double FN,
                         double FP, double p, double q, double* A, void* stream) {
    if (!sd_idx) return bad_arg("sd_idx");
    return theta_log_ratio_impl(th_new, th_old, R, M, S1, S0, sd_idx, 0, 0, blo, bhi, FN, FP, p, q, A, stream);
}

int bnpc_row_loglik(const float* theta, const int32_t* ids, int R, int M, const int32_t* S1,
                    const int32_t* S0, const double* fn_h, const double* fp_h, int E, double p,
                    double q, double* out, double* prior_out, void* stream) {
    if (R <= 0) return 0;
    if (E < 0 || E > RL_MAXE) return bad_arg("E must be in [0,4]");
    RlArgs g;
    memset(&g, 0, sizeof(g));
    g.E = E;
    for (int e = 0; e < E; ++e) { g.fn[e] = fn_h[e]; g.fp[e] = fp_h[e]; }
    g.p = p; g.q = q;
    g.betaln = lgamma(p) + lgamma(q) - lgamma(p + q);
    BNPC_LAUNCH(row_loglik_kernel, 256, 0, R, 256, 0, (cudaStream_t)stream, theta, ids, R, M, S1, S0, g, out, prior_out);
    return 0;
}

int bnpc_row_sum(const double* v, int R, int M, double* out, void* stream) {
    if (R <= 0) return 0;
    BNPC_LAUNCH(row_sum_kernel, 256, 0, R, 256, 0, (cudaStream_t)stream, v, R, M, out);
    return 0;
}

int bnpc_gather_members(const int32_t* assign, int N, int id_a, int id_b, int32_t* cells_out,
                        int32_t* blk, void* stream) {
    if (N <= 0) return 0;
    const int nb = cdiv(N, 1024);
    BNPC_LAUNCH(gather_count_kernel, 1024, 0, nb, 1024, 0, (cudaStream_t)stream, assign, N, id_a, id_b, blk);
    BNPC_LAUNCH(gather_scan_kernel, 0, 0, 1, 32, 0, (cudaStream_t)stream, blk, nb);
    BNPC_LAUNCH(gather_scatter_kernel, 1024, 0, nb, 1024, 0, (cudaStream_t)stream, assign, N, id_a, id_b, blk, cells_out);
    return 0;
}

int bnpc_anchor_swaps(int32_t* cells, int n, int n_a, int idx_i, int idx_j, int is_merge, void* stream) {
    if (n < 2) return bad_arg("n < 2");
    BNPC_LAUNCH(anchor_swaps_kernel, 0, 0, 1, 32, 0, (cudaStream_t)stream, cells, n, n_a, idx_i, idx_j, is_merge);
    return 0;
}

int bnpc_rg_launch_halves(const uint32_t* x1, const uint32_t* x0, int W, const int32_t* cells, int n,
                          const double* k6_h, int32_t* half, void* stream) {
    if (n <= 2) return 0;
    K6 k;
    for (int i = 0; i < 6; ++i) k.k[i] = k6_h[i];
    BNPC_LAUNCH(rg_launch_halves_kernel, 0, 0, cdiv(n - 2, 128), 128, 0, (cudaStream_t)stream, x1, x0, W, cells, n, k, half);
    return 0;
}

static int rg_sides_impl(const int32_t* cells, int n, const int32_t* half, int32_t* members, int32_t* seg_off,
                         bool clear, void* stream) {
    if (n < 2) return bad_arg("n < 2");
    if (clear)
        if (int rc = zero_async(seg_off, sizeof(int32_t) * 8, stream, "rg_sides memset")) return rc;
    if (n > 2) {
        BNPC_LAUNCH(rg_count_kernel, 0, 0, cdiv(n - 2, 256), 256, 0, (cudaStream_t)stream, half, n - 2, seg_off);
    }
    BNPC_LAUNCH(rg_sides_kernel, 0, 0, cdiv(n, 256), 256, 0, (cudaStream_t)stream, cells, n, half, members, seg_off);
    return 0;
}

// mode 0 with perm / u == NULL: order and uniforms from streams stream_id+1, +2 of seed
static int rg_scan_impl(const double* ll2, int ldk, int n, const int32_t* perm, const double* u, uint64_t seed,
                        uint64_t stream_id, int32_t* half, double alpha, int mode, const int32_t* cells,
                        const int32_t* assign, int id_i, double* lq, int32_t* work, void* stream) {
    if (n <= 2) return 0;
    if (mode == 1 && (!cells || !assign)) return bad_arg("cells/assign required for replay");
    if (!work) return bad_arg("work");
    const int nf = n - 2;
    const int out_off = (nf + 3) & ~3;
    const int hb = feistel_half_bits(nf);
    cudaStream_t st = (cudaStream_t)stream;
    BNPC_LAUNCH(rg_prepare_kernel, 0, 0, cdiv(nf, 128), 128, 0, st, ll2, ldk, n, perm, u, half, alpha, mode, cells, assign, id_i, work, seed, stream_id, hb);
    BNPC_LAUNCH(rg_serial_kernel, 32, 0, 1, 32, 0, st, half, nf, work, out_off);
    BNPC_LAUNCH(rg_finish_kernel, 0, 0, cdiv(nf, 128), 128, 0, st, ll2, ldk, n, perm, half, alpha, mode, work, out_off, lq, seed, stream_id, hb);
    return 0;
}
int bnpc_rg_sides(const int32_t* cells, int n, const int32_t* half, int32_t* members, int32_t* seg_off,
                  void* stream) {
    return rg_sides_impl(cells, n, half, members, seg_off, true, stream);
}

int bnpc_rg_scan(const double* ll2, int ldk, int n, const int32_t* perm, const double* u, int32_t* half,
                 double alpha, int mode, const int32_t* cells, const int32_t* assign, int id_i,
                 double* lq, int32_t* work, void* stream) {
    if (mode == 0 && (!perm || !u)) return bad_arg("perm/u required for a sampled scan");
    return rg_scan_impl(ll2, ldk, n, perm, u, 0, 0, half, alpha, mode, cells, assign, id_i, lq, work, stream);
}

int bnpc_apply_split(const int32_t* cells, int n, const int32_t* half, int new_id, int32_t* assign,
                     void* stream) {
    BNPC_LAUNCH(apply_split_kernel, 0, 0, cdiv(n, 256), 256, 0, (cudaStream_t)stream, cells, n, half, new_id, assign);
    return 0;
}

int bnpc_apply_merge(const int32_t* cells, int n_a, int n, int id, int32_t* assign, void* stream) {
    if (n <= n_a) return 0;
    BNPC_LAUNCH(apply_merge_kernel, 0, 0, cdiv(n - n_a, 256), 256, 0, (cudaStream_t)stream, cells, n_a, n, id, assign);
    return 0;
}

#include "bnpc_chain.cuh"

}  // extern "C"

#include "bnpc_group.cuh"
