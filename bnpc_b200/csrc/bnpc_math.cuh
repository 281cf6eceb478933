// Device-side scalar math for the BnpC hot path (sm_100a).
//
// The Metropolis-Hastings moves of the reference lean on scipy.stats.truncnorm
// (libs/CRP.py:331,351-357; libs/CRP_learning_errors.py:82-91) and scipy.stats.beta
// (libs/CRP.py:375-376).  These helpers restate the formulas scipy evaluates
// (scipy/stats/_continuous_distns.py: _log_gauss_mass, truncnorm_gen._logpdf/_ppf;
// scipy.special.log_ndtr / ndtr / ndtri_exp) in double precision so that the same
// uniform yields the same proposal and the same acceptance ratio to ~1e-15.
//
// Built with -fmad=false: products and sums round separately, as numpy's do; fused
// multiply-adds appear only where written as fma().
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define BNPC_INF (__longlong_as_double(0x7ff0000000000000LL))

namespace bnpc {

// reference libs/CRP.py:11-14
__device__ __constant__ const double kLogEps = -34.538776394910684;   // np.log(1e-15)
constexpr double kThetaLo = 1e-5;
constexpr double kThetaHi = 1 - 1e-5;
constexpr double kNormLogC = 0.91893853320467274178;                  // log(sqrt(2*pi))

// The special functions below are deliberately NOT inlined: the Metropolis-Hastings kernels call
// them a dozen times per element in straight-line code that each thread runs once; inlined, that
// code is ~290 KB of SASS and the kernel is bound by cold instruction fetches (ncu: 70 % of the
// stall samples "no instructions"), not by the arithmetic.
__device__ __noinline__ double ndtr(double x) {
    // Cephes ndtr as used by scipy.special.ndtr
    const double z = x * 0.70710678118654752440;
    const double az = fabs(z);
    if (az < 0.70710678118654752440) return 0.5 + 0.5 * erf(z);
    double y = 0.5 * erfc(az);
    return (z > 0) ? 1.0 - y : y;
}

__device__ __noinline__ double log_ndtr(double x) {
    // scipy.special.log_ndtr: erfcx form in the left tail, log1p elsewhere
    const double t = x * 0.70710678118654752440;
    if (x < -1.0) return log(erfcx(-t) / 2) - t * t;
    return log1p(-erfc(t) / 2);
}

__device__ __forceinline__ double log_diff_exp(double lp, double lq) {
    // log(exp(lp) - exp(lq)), lp > lq  (scipy _log_diff via complex logsumexp)
    return log(1.0 - exp(lq - lp)) + lp;
}

__device__ __forceinline__ double log_sum_exp2(double lp, double lq) {
    const double m = fmax(lp, lq);
    if (m == -BNPC_INF) return m;
    return log(exp(lp - m) + exp(lq - m)) + m;
}

__device__ __noinline__ double log_gauss_mass(double a, double b) {
    // scipy/stats/_continuous_distns.py:_log_gauss_mass
    if (b <= 0) return log_diff_exp(log_ndtr(b), log_ndtr(a));
    if (a > 0) return log_diff_exp(log_ndtr(-a), log_ndtr(-b));
    return log1p(-ndtr(a) - ndtr(-b));
}

__device__ __forceinline__ double truncnorm_logpdf_mass(double y, double a, double b, double log_mass) {
    // rv_continuous.logpdf support mask (closed interval), then truncnorm_gen._logpdf with
    // log_mass = _log_gauss_mass(a, b) supplied by the caller
    if (!(y >= a && y <= b)) return -BNPC_INF;
    return -(y * y) / 2.0 - kNormLogC - log_mass;
}
__device__ __forceinline__ double truncnorm_logpdf_std(double y, double a, double b) {
    if (!(y >= a && y <= b)) return -BNPC_INF;
    return truncnorm_logpdf_mass(y, a, b, log_gauss_mass(a, b));
}

__device__ __noinline__ double ndtri_exp(double y) {
    // scipy.special.ndtri_exp = ndtri(exp(y)) with the upper branch kept accurate
    if (y > -0.14541345786885906) return -normcdfinv(-expm1(y));
    return normcdfinv(exp(y));
}

__device__ __forceinline__ double truncnorm_ppf_mass(double q, double a, double b, double log_mass) {
    // truncnorm_gen._ppf with log_mass = _log_gauss_mass(a, b) supplied by the caller
    if (a < 0) {
        const double lphi = log_sum_exp2(log_ndtr(a), log(q) + log_mass);
        return ndtri_exp(lphi);
    }
    const double lphi = log_sum_exp2(log_ndtr(-b), log1p(-q) + log_mass);
    return -ndtri_exp(lphi);
}
__device__ __forceinline__ double truncnorm_ppf_std(double q, double a, double b) {
    return truncnorm_ppf_mass(q, a, b, log_gauss_mass(a, b));
}

__device__ __noinline__ double beta_logpdf(double x, double p, double q, double betaln_pq) {
    // scipy beta_gen._logpdf: xlog1py(q-1, -x) + xlogy(p-1, x) - betaln(p, q)
    double t1 = (q - 1.0 == 0.0) ? 0.0 : (q - 1.0) * log1p(-x);
    double t2 = (p - 1.0 == 0.0) ? 0.0 : (p - 1.0) * log(x);
    return (t1 + t2) - betaln_pq;
}

// Bernoulli-with-errors log-probabilities of one (cluster, mutation) entry,
// reference libs/CRP.py:197-212.  theta is float32, (1 - theta) rounds in float32.
__device__ __noinline__ void log_p1_p0(float th, double FN, double FP, double& lp1, double& lp0) {
    const double t = (double)th;
    const double omt = (double)(1.0f - th);
    lp1 = log(t * (1.0 - FN) + omt * FP);
    lp0 = log(t * FN + omt * (1.0 - FP));
}

// ------------------------------------------------------------------ Philox4x32-10
struct Philox {
    uint2 key;
    __device__ __forceinline__ Philox(uint64_t seed) { key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)); }
    __device__ __forceinline__ uint4 operator()(uint64_t c0, uint64_t c1) const {
        uint4 c = make_uint4((uint32_t)c0, (uint32_t)(c0 >> 32), (uint32_t)c1, (uint32_t)(c1 >> 32));
        uint2 k = key;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
            c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
            k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
        }
        return c;
    }
};

__device__ __forceinline__ double u01(uint32_t hi, uint32_t lo) {
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return (double)(v >> 11) * (1.0 / 9007199254740992.0);      // [0,1), 53 bits
}

// Sequential stream of uniforms for rejection samplers: counter (idx, sub++) under
// one (seed, stream) pair; placement-independent.
struct PhiloxStream {
    Philox g; uint64_t c0, c1; uint4 buf; int have;
    __device__ __forceinline__ PhiloxStream(uint64_t seed, uint64_t stream_id, uint64_t idx)
        : g(seed ^ (stream_id * 0x9E3779B97F4A7C15ull)), c0(idx), c1(0), have(0) {}
    __device__ __forceinline__ double next() {
        if (have == 0) { buf = g(c0, c1++); have = 2; return u01(buf.x, buf.y) ; }
        have = 0; return u01(buf.z, buf.w);
    }
    __device__ __forceinline__ double next_open() {          // (0,1)
        double u; do { u = next(); } while (u <= 0.0); return u;
    }
};

__device__ inline double gamma_sample(double shape, PhiloxStream& s) {
    // Marsaglia & Tsang (2000); shape < 1 boosted through shape+1
    double boost = 1.0;
    if (shape < 1.0) { boost = pow(s.next_open(), 1.0 / shape); shape += 1.0; }
    const double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (;;) {
        double x, v;
        do {
            const double u1 = s.next_open(), u2 = s.next();
            x = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
            v = 1.0 + c * x;
        } while (v <= 0.0);
        v = v * v * v;
        const double u = s.next_open();
        if (u < 1.0 - 0.0331 * (x * x) * (x * x)) return boost * d * v;
        if (log(u) < 0.5 * x * x + d * (1.0 - v + log(v))) return boost * d * v;
    }
}

__device__ inline double beta_sample(double a, double b, PhiloxStream& s) {
    const double x = gamma_sample(a, s), y = gamma_sample(b, s);
    const double t = x + y;
    return (t > 0.0) ? x / t : 0.5;
}

__device__ __forceinline__ float clip_theta(double v) {
    // np.clip(v, 1e-5, 1-1e-5).astype(np.float32), libs/CRP.py:180,188
    return (float)fmin(fmax(v, kThetaLo), kThetaHi);
}

}  // namespace bnpc
