// d3d_rng.cuh -- counter-based random stream and truncated-normal sampler.
//
// Replaces the global numpy.random state the reference draws from
// (lib/run.py:313, 435, 578; lib/rtnorm.py:17) by "d3d stream v1":
//   key = (seed_lo, seed_hi), counter = (k>>1, site, sweep, chain), Philox4x32-10,
//   draw k = 53-bit uniform in [0,1) from words (0,1) (k even) or (2,3) (k odd).
// Per (chain, sweep, site) the draws are consumed in the reference's call order:
//   0..2 jump uniforms (lib/run.py:578), 3 acceptance (lib/run.py:435),
//   4..  truncated-normal sub-stream (lib/rtnorm.py:121-218).
// The CPU oracle (oracle/philox.py, oracle/streams.py) implements the same map.
#pragma once
#include <stdint.h>

namespace d3d {

// The double-precision libm routines are 0.5-2 KB of SASS each.  Inlined at every call site
// they blew the sweep kernels up to ~110 KB, far beyond the instruction caches, and the
// scalar warps stalled on instruction fetch.  One out-of-line copy each keeps the per-site
// code footprint small.
__device__ __noinline__ double d_log(double x) { return log(x); }
__device__ __noinline__ double d_exp(double x) { return exp(x); }
__device__ __noinline__ double d_tan(double x) { return tan(x); }
__device__ __noinline__ double d_cos(double x) { return cos(x); }
__device__ __noinline__ double d_div(double a, double b) { return a / b; }
__device__ __noinline__ double d_sqrt(double x) { return sqrt(x); }

// Philox4x32-10 block function (Salmon et al. 2011), out of line for the same reason.
__device__ __noinline__ void philox_block(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1,
                                          uint32_t c2, uint32_t c3, uint32_t* o) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
        uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

__device__ __forceinline__ void philox_block_inl(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1,
                                                 uint32_t c2, uint32_t c3, uint32_t* o) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
        uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

// Rolled form (code footprint of the pipelined sweep kernel, d3d_pipe.cuh): same rounds.
__device__ __forceinline__ void philox_block_rolled(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1,
                                                    uint32_t c2, uint32_t c3, uint32_t* o) {
#pragma unroll 1
    for (int r = 0; r < 10; ++r) {
        const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
        uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

struct Philox {
    uint32_t k0, k1;          // key
    uint32_t site, sweep, chain;
    uint32_t k;               // next draw index
    double   cached;          // odd draw of the current block
    const double* stash;      // optional: draws 4..7 evaluated ahead of time (speculate_draws)

    __device__ __forceinline__ void init(uint64_t seed, uint32_t chain_, uint32_t sweep_,
                                         uint32_t site_) {
        k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32);
        chain = chain_; sweep = sweep_; site = site_; k = 0; cached = 0.0; stash = nullptr;
    }

    __device__ __forceinline__ static double u53(uint32_t hi, uint32_t lo) {
        return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);
    }

    __device__ __forceinline__ double next() {
        if (stash && k >= 4u && k < 8u) { const double u = stash[k - 4u]; ++k; return u; }
        if (k & 1u) { ++k; return cached; }
        uint32_t o[4];
        philox_block(k0, k1, k >> 1, site, sweep, chain, o);
        cached = u53(o[2], o[3]);
        ++k;
        return u53(o[0], o[1]);
    }

    // numpy uniform(low, 1.0): low + (1 - low) * u        (lib/rtnorm.py:17 `rand`)
    __device__ __forceinline__ double rand(double low) { return low + (1.0 - low) * next(); }
    // Box-Muller on two consecutive draws                 (lib/rtnorm.py:17 `randn`)
    __device__ __forceinline__ double randn() {
        double u1 = next(), u2 = next();
        return d_sqrt(-2.0 * d_log(1.0 - u1)) * d_cos(6.283185307179586 * u2);
    }
    // integer in [lo, hi)                                 (lib/rtnorm.py:17 `randi`)
    __device__ __forceinline__ int randi(int lo, int hi) {
        return lo + (int)floor(next() * (double)(hi - lo));
    }
};

// Tables of lib/rtnorm.py:227 (x), :1230 (yu), :2233 (ncell), uploaded at run time.
struct RtTables {
    const double* x;      // [4002]
    const double* yu;     // [4001]
    const int*    ncell;  // [8961]
};

#define D3D_RT_GUARD 100000   // rejection-loop guard (the reference loops forever)

// Transcendentals of the FIRST draws of the truncated-normal sub-stream (draws 4..7 of a
// site), evaluated ahead of time by the proposal warp so that they are off the serial
// accept -> Gibbs -> update chain.  Same values the sampler would compute itself.
enum { SP_U4 = 0, SP_U5, SP_U6, SP_U7, SP_E1, SP_Z1, SP_N1, SP_N2, SP_IRA, SP_LOGU, SP_N };

__device__ __noinline__ void speculate_draws(Philox rng /* by value, at k = 4 */, double* sp) {
    const double u4 = rng.next(), u5 = rng.next(), u6 = rng.next(), u7 = rng.next();
    sp[SP_U4] = u4; sp[SP_U5] = u5; sp[SP_U6] = u6; sp[SP_U7] = u7;
    sp[SP_E1] = -d_log(1e-15 + (1.0 - 1e-15) * u5);            // e of the tail branch, :122
    sp[SP_Z1] = d_log(1.0 + (1e-15 + (1.0 - 1e-15) * u4) * -1.0);  // z when expab == -1, :121
    sp[SP_N1] = d_sqrt(-2.0 * d_log(1.0 - u4)) * d_cos(6.283185307179586 * u5);   // randn #1
    sp[SP_N2] = d_sqrt(-2.0 * d_log(1.0 - u6)) * d_cos(6.283185307179586 * u7);   // randn #2
}

// lib/rtnorm.py:95-223.  `fail` is set when a NaN bound arrives (the reference
// raises at lib/rtnorm.py:144) or a rejection loop exceeds the guard.
__device__ __noinline__ double rtstdnorm(double a, double b, Philox& rng, const RtTables t,
                                         int* fail, const double* spec = nullptr) {
    const double xmin = -2.00443204036, xmax = 3.48672170399;       // :101-102
    if (!(a < b)) { *fail = 1; return a; }                          // :105-106 (and NaN)
    double sign = 1.0;
    if (fabs(a) > fabs(b)) {                                        // :108-109 mirror
        double na = -b, nb = -a; a = na; b = nb; sign = -1.0;
    }
    if (a > xmax) {                                                 // :112-124
        const double twoasq = 2.0 * a * a;
        const double expab = d_exp(-a * (b - a)) - 1.0;
        double z = 0.0;
        if (spec && rng.k == 4) {                                    // pre-evaluated first try
            z = (expab == -1.0) ? spec[SP_Z1]
                                : d_log(1.0 + (1e-15 + (1.0 - 1e-15) * spec[SP_U4]) * expab);
            rng.k = 6;
            if (twoasq * spec[SP_E1] > z * z) return sign * (a - d_div(z, a));
        }
        for (int it = 0; it < D3D_RT_GUARD; ++it) {
            z = d_log(1.0 + rng.rand(1e-15) * expab);
            double e = -d_log(rng.rand(1e-15));
            if (twoasq * e > z * z) return sign * (a - d_div(z, a));
        }
        *fail = 1; return sign * a;
    }
    if (a < xmin) {                                                 // :127-131
        if (spec && rng.k == 4) {                                    // pre-evaluated normals
            double r = spec[SP_N1];
            rng.k = 6;
            if (r >= a && r <= b) return sign * r;
            r = spec[SP_N2];
            rng.k = 8;
            if (r >= a && r <= b) return sign * r;
        }
        for (int it = 0; it < D3D_RT_GUARD; ++it) {
            double r = rng.randn();
            if (r >= a && r <= b) return sign * r;
        }
        *fail = 1; return sign * a;
    }
    // Chopin's table algorithm, :133-222
    const int    kmin = 5, I0 = 3271, N = 4000;
    const double INVH = 1631.73284006, ALPHA = 1.837877066409345;
    const double yl0 = 0.053513975472, ylN = 0.000914116389555;
    int ka = t.ncell[I0 + (int)floor(a * INVH)];
    int kb = (b >= xmax) ? N : t.ncell[I0 + (int)floor(b * INVH)];
    if (abs(kb - ka) < kmin) {                                      // :154-163
        const double twoasq = 2.0 * a * a;
        const double expab = d_exp(-a * (b - a)) - 1.0;
        for (int it = 0; it < D3D_RT_GUARD; ++it) {
            double z = d_log(1.0 + rng.rand(0.0) * expab);
            double e = -d_log(rng.rand(0.0));
            if (twoasq * e > z * z) return sign * (a - d_div(z, a));
        }
        *fail = 1; return sign * a;
    }
    for (int it = 0; it < D3D_RT_GUARD; ++it) {                     // :164-222
        int k = rng.randi(ka, kb + 1);
        if (k == N) {                                               // right tail
            double lbound = t.x[N + 1];
            double z = -d_log(rng.rand(0.0));
            double e = -d_log(rng.rand(0.0));
            z = d_div(z, lbound);
            if (z * z <= 2.0 * e && z < b - lbound) return sign * (lbound + z);
        } else if (k <= ka + 2 || (k >= kb && b < xmax)) {          // edge strips
            double xk = t.x[k], xk1 = t.x[k + 1];
            double sim = xk + (xk1 - xk) * rng.rand(0.0);
            if (sim >= a && sim <= b) {
                double yuk = t.yu[k];
                double simy = yuk * rng.rand(0.0);
                double ylk;
                if (k == 0) ylk = yl0;
                else if (k == N) ylk = ylN;
                else if (k <= 1954) ylk = t.yu[k - 1];
                else ylk = t.yu[k + 1];
                if (simy < ylk || sim * sim + 2.0 * d_log(simy) + ALPHA < 0.0) return sign * sim;
            }
        } else {                                                    // inner strips
            double u = rng.rand(0.0);
            double yuk = t.yu[k];
            double simy = yuk * u;
            double xk = t.x[k];
            double d = t.x[k + 1] - xk;
            double ylk;
            if (k == 1) ylk = yl0;
            else if (k == N) ylk = ylN;
            else if (k <= 1954) ylk = t.yu[k - 1];
            else ylk = t.yu[k + 1];
            if (simy < ylk) return sign * (xk + d_div(u * d * yuk, ylk));
            double sim = xk + d * rng.rand(0.0);
            if (sim * sim + 2.0 * d_log(simy) + ALPHA < 0.0) return sign * sim;
        }
    }
    *fail = 1; return sign * a;
}

// lib/rtnorm.py:21-92 with size=1.
__device__ __forceinline__ double rtnorm(double a, double b, double mu, double sigma,
                                         Philox& rng, const RtTables& t, int* fail) {
    bool scaled = !(mu == 0.0) || !(sigma == 1.0);
    if (scaled) { a = d_div(a - mu, sigma); b = d_div(b - mu, sigma); }     // :74-76
    double r = rtstdnorm(a, b, rng, t, fail);
    if (scaled) r = r * sigma + mu;                                 // :82-83
    return r;
}

// Same sampler for the sweep kernels, with the Gaussian given by (mu, 1/sigma, sigma): the
// bounds are standardised with multiplications (no serial divide) and the first draws come
// from the pre-evaluated `spec` block.
__device__ __forceinline__ double rtnorm_spec(double lo, double hi, double mu, double isg,
                                              double sigma, Philox& rng, const RtTables& t,
                                              int* fail, const double* spec) {
    const double a = (lo - mu) * isg, b = (hi - mu) * isg;
    // draws 4..7 served from the stash when the generic path asks for them
    double r = rtstdnorm(a, b, rng, t, fail, spec);
    return r * sigma + mu;
}

}  // namespace d3d
