// Posterior (MPEAR) estimator kernels -- reference libs/utils.py:90-145.
//
// The reference accumulates, sample by sample, the Hamming distance of every pair of cells
// (scipy pdist over an N(N-1)/2 vector per posterior sample, in a Python loop) and scores every
// candidate cut of the ward dendrogram with three float sums over all pairs.  Here:
//   cocluster_counts_kernel   counts[pair] = number of samples in which the two cells sit in
//                             different clusters, int32, condensed in pdist order (i < j,
//                             row-major).  One CTA per 64 x 64 tile of the upper triangle,
//                             16 pairs per thread, samples staged through shared memory.
//                             Integer compare/add work: S * N^2 / 2 compares, N^2 * 2 bytes out.
//   mpear_sums_kernel         for every candidate labelling c: A_c = #pairs with equal labels and
//                             B_c = sum of counts over those pairs, plus T = sum of all counts --
//                             exact 64-bit integers (order-independent atomics), from which the
//                             host forms I_sum = A, pi_sum = P - T/S, index = A - B/S of
//                             Fritsch & Ickstadt's eq. 13 in float64.
#pragma once

#define EST_TILE 64
#define EST_SCHUNK 32

__device__ __forceinline__ long long condensed_index(long long i, long long j, long long N) {
    // pdist order: (0,1), (0,2), ..., (0,N-1), (1,2), ...
    return i * N - i * (i + 1) / 2 + (j - i - 1);
}

__device__ __forceinline__ void cocluster_counts_kernel(const int32_t* __restrict__ assign, int S, int N, int32_t* __restrict__ counts) {
    const int tj = blockIdx.x, ti = blockIdx.y;
    if (tj < ti) return;
    __shared__ int32_t sa[EST_SCHUNK][EST_TILE], sb[EST_SCHUNK][EST_TILE];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int i0 = ti * EST_TILE, j0 = tj * EST_TILE;
    int cnt[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) cnt[a][b] = 0;
    for (int s0 = 0; s0 < S; s0 += EST_SCHUNK) {
        const int ns = min(EST_SCHUNK, S - s0);
        __syncthreads();
        for (int e = threadIdx.x; e < EST_SCHUNK * EST_TILE; e += 256) {
            const int s = e / EST_TILE, c = e % EST_TILE;            // coalesced along cells
            int va = -1, vb = -2;
            if (s < ns) {
                const long long rowoff = (long long)(s0 + s) * N;
                if (i0 + c < N) va = assign[rowoff + i0 + c];
                if (j0 + c < N) vb = assign[rowoff + j0 + c];
            }
            sa[s][c] = va; sb[s][c] = vb;
        }
        __syncthreads();
        for (int s = 0; s < ns; ++s) {
            int ai[4], bj[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) { ai[a] = sa[s][ty * 4 + a]; bj[a] = sb[s][tx * 4 + a]; }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) cnt[a][b] += (ai[a] != bj[b]) ? 1 : 0;
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const long long i = i0 + ty * 4 + a;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const long long j = j0 + tx * 4 + b;
            if (i < j && j < N) counts[condensed_index(i, j, N)] = cnt[a][b];
        }
    }
}

#define EST_CGROUP 32          /* candidate labellings staged per round */
// weight (or NULL): multiplicity of every point -- the points are the distinct assignment profiles
// of the cells and pair (i, j) stands for weight[i] * weight[j] pairs of cells
__device__ __forceinline__ void mpear_sums_kernel(const int32_t* __restrict__ counts, int N, const int32_t* __restrict__ labels, int n_cand,
                  unsigned long long* __restrict__ out, const int32_t* __restrict__ weight) {
    const int tj = blockIdx.x, ti = blockIdx.y;
    if (tj < ti) return;
    __shared__ int32_t li[EST_CGROUP][EST_TILE], lj[EST_CGROUP][EST_TILE];
    __shared__ unsigned long long accA[EST_CGROUP], accB[EST_CGROUP], accT;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4, lane = threadIdx.x & 31;
    const int i0 = ti * EST_TILE, j0 = tj * EST_TILE;
    int cnt[4][4];
    bool ok[4][4];
    unsigned long long wt[4][4];
    unsigned long long t_loc = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const long long i = i0 + ty * 4 + a;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const long long j = j0 + tx * 4 + b;
            ok[a][b] = i < j && j < N;
            cnt[a][b] = ok[a][b] ? counts[condensed_index(i, j, N)] : 0;
            wt[a][b] = (ok[a][b] && weight) ? (unsigned long long)weight[i] * (unsigned long long)weight[j] : 1ull;
            t_loc += wt[a][b] * (unsigned long long)cnt[a][b];
        }
    }
    if (threadIdx.x == 0) accT = 0ull;
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t_loc += __shfl_xor_sync(0xffffffffu, t_loc, o);
    if (lane == 0 && t_loc) atomicAdd(&accT, t_loc);
    for (int c0 = 0; c0 < n_cand; c0 += EST_CGROUP) {
        const int nc = min(EST_CGROUP, n_cand - c0);
        __syncthreads();
        if (threadIdx.x < EST_CGROUP) { accA[threadIdx.x] = 0ull; accB[threadIdx.x] = 0ull; }
        for (int e = threadIdx.x; e < EST_CGROUP * EST_TILE; e += 256) {
            const int c = e / EST_TILE, x = e % EST_TILE;
            int va = -1, vb = -2;
            if (c < nc) {
                const long long off = (long long)(c0 + c) * N;
                if (i0 + x < N) va = labels[off + i0 + x];
                if (j0 + x < N) vb = labels[off + j0 + x];
            }
            li[c][x] = va; lj[c][x] = vb;
        }
        __syncthreads();
        for (int c = 0; c < nc; ++c) {
            int ai[4], bj[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) { ai[a] = li[c][ty * 4 + a]; bj[a] = lj[c][tx * 4 + a]; }
            unsigned long long A = 0, B = 0;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if (ok[a][b] && ai[a] == bj[b]) { A += wt[a][b]; B += wt[a][b] * (unsigned long long)cnt[a][b]; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                A += __shfl_xor_sync(0xffffffffu, A, o);
                B += __shfl_xor_sync(0xffffffffu, B, o);
            }
            if (lane == 0 && A) { atomicAdd(&accA[c], A); atomicAdd(&accB[c], B); }
        }
        __syncthreads();
        if (threadIdx.x < nc) {
            if (accA[threadIdx.x]) atomicAdd(&out[1 + 2 * (c0 + threadIdx.x)], accA[threadIdx.x]);
            if (accB[threadIdx.x]) atomicAdd(&out[2 + 2 * (c0 + threadIdx.x)], accB[threadIdx.x]);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && accT) atomicAdd(&out[0], accT);
}
