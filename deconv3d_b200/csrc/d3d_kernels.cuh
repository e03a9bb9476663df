// d3d_kernels.cuh -- sm_100a kernels of the deconv3d likelihood hot path.
//
// Device data layout (DESIGN.md "Data layout in HBM"): every cube is stored
// z-fastest, [y][x][Dp] with Dp = D rounded up to the 16-byte vector width of
// the storage type T (double2 / float4), so that the FSF window of a proposal
// is wh runs of ww*Dp contiguous elements and a spaxel spectrum is one vector
// run.  Padded channels hold zeros everywhere (data, 1/variance, residual).
//
// Maths of one site update (identical to lib/run.py:367-519, without the
// full-cube temporaries):
//   F = FSF window, Lu(c,w) = lsf (*) exp(-(z-c)^2/(2w^2))   (unit amplitude)
//   h[z]  = sum_s F[s] iv[z,s] e[z,s]       G[z] = sum_s F[s]^2 iv[z,s]
//   dLu   = Lu_old - Lu_new
//   delta = ar_old - ar_new = -( a*sum_z dLu h  +  a^2/2 * sum_z dLu^2 G )
//   Gibbs (L_end = accepted ? Lu_new : Lu_old):
//     S2 = sum_z L_end^2 G ,  S1 = sum_z L_end (h + a Lu_old G)
//     ro = ra/(1+ra S2), mu = ro S1, r ~ TN(mu, sqrt(ro); [a_min, a_max])
//   e <- e + F (a Lu_old - r L_end)
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "d3d_rng.cuh"

namespace d3d {

template <typename T> struct Vec;
template <> struct Vec<double> { typedef double2 V; static const int N = 2; };
template <> struct Vec<float>  { typedef float4  V; static const int N = 4; };

__device__ __forceinline__ void unpack(const double2& v, double* o) { o[0] = v.x; o[1] = v.y; }
__device__ __forceinline__ void unpack(const float4& v, double* o) {
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
__device__ __forceinline__ void pack(double2& v, const double* o) { v.x = o[0]; v.y = o[1]; }
__device__ __forceinline__ void pack(float4& v, const double* o) {
    v.x = (float)o[0]; v.y = (float)o[1]; v.z = (float)o[2]; v.w = (float)o[3];
}

struct Problem {
    int D, Dp, H, W, fh, fw, fhh, fhw, P;
    int n_cubes, chains_per_cube, n_chains;
    int var_is_cube, has_lsf, max_sites;
    const void* data;          // [cube][H][W][Dp] T
    const void* iv;            // [cube][H][W][Dp] T   (var_is_cube)
    const double* iv_scalar;   // [cube]               (!var_is_cube)
    void* err;                 // [chain][H][W][Dp] T
    double* params;            // [chain][H][W][3]
    const uint8_t* mask;       // [cube][H][W]
    const int* sites;          // [cube][max_sites]  linear y*W+x, row-major (lib/run.py:553-566)
    const int* n_sites;        // [cube]
    const double* fsf;         // [fh*fw]
    const double* gtab;        // [cube][H*W][Dp] sum of F^2/sigma^2 over the window of a site (d3d_slide.cuh), or NULL
    const int* run_start;      // [cube][max_sites] index of the first site of the run a list entry belongs to (d3d_pipe.cuh), or NULL
    const int* run_last;       // [cube][max_sites] index of the last site of that run, or NULL
    const double* xtab;        // [cube][xtab_L][max_sites][Dp] cross terms of consecutive sites (d3d_pipe.cuh), or NULL
    int xtab_L;                // look-ahead distances the table holds (>= the L of the kernel reading it)
    double* lucache;           // [chain][H*W][Dp] unit line profile every site ended its last visit with (d3d_pipe.cuh), or NULL
    int lu_valid;              // the cache matches the parameter map at the start of this launch
    const double* kcirc;       // [P] circular LSF kernel (lib/convolution.py:89-160 in direct form)
    const double* ktap_v;      // [ntaps] values and
    const int* ktap_m;         // [ntaps] offsets m of the taps with |K[m]| >= 1e-18 max|K|
    int ntaps;
    // Line model (lib/line_models.py:4-61).  The native family is a TIED MULTIPLET: n_comp Gaussians
    // sharing the centre shift and the width, component k at c + comp_off[k] with relative
    // amplitude comp_ratio[k] (component 0: offset 0, ratio 1).  n_comp = 1 is
    // SingleGaussianLineModel (lib/line_models.py:64-109).  The model stays linear in the
    // amplitude, so the Gibbs step of lib/run.py:456-519 applies unchanged.
    int n_comp;
    double comp_off[4], comp_ratio[4];
    const double* kdense;      // [kd_n] the same kernel as a dense window of signed offsets mhi, mhi-1, ... (d3d_pipe.cuh)
    int kd_n, kd_mhi;
    const double* pmin;        // [cube][3]
    const double* pmax;        // [cube][3]
    const double* prior_var;   // [cube]
    double jump[3];
    RtTables rt;
    unsigned long long seed;
    unsigned int first_chain;
    long long* accepted;       // [chain]  accepted_count           (lib/run.py:341,440)
    long long* iters;          // [chain]  cur_iteration reached
    double* rate;              // [chain]  cur_acceptance_rate      (lib/run.py:356-359)
    int* active;               // [chain]
    int* status;               // [1] sticky numeric-failure flag
    int* dbg;                  // [32] per-warp progress of the pipelined sweep kernel when it gives up
    // Spatial tiling of ONE cube over several contexts/GPUs (coloured mode, d3d_tile.cuh):
    // only sites inside the tile are updated by this context; the residual is kept valid
    // inside the region = tile grown by the FSF half-size.
    int ty0, ty1, tx0, tx1;    // owned sites: ty0 <= y < ty1, tx0 <= x < tx1
    int ry0, ry1, rx0, rx1;    // residual voxels this context keeps up to date
    double* lik_cur;           // [chain][H][W] delta-logL of the latest update (tile mode)
    uint8_t* acc_cur;          // [chain][H][W] its accept flag
};

// One proposal evaluated without touching any state (d3d_delta_logl).
struct EvalReq {
    int enabled;
    double p_new[3];
    double* out;               // [3] delta, ar_old, ar_new
};

// Shared-memory carve-up, all doubles.
struct Smem {
    double* F;      // [fh*fw]
    double* Kv;     // [P]  tap values
    int*    Km;     // [P]  tap offsets
    double* g_o;    // [Dp]  scratch of scalar warp A
    double* g_n;    // [Dp]  scratch of scalar warp B
    double* Lu_o;   // [2][Dp] old / new unit line profiles, double-buffered by site parity
    double* Lu_n;   // [2][Dp]
    double* red;    // [32][8]
    double* bc;     // [8]   decision broadcast: accepted, r, accepted_count, a
    double* prop;   // [2][8] proposal stash
    double* spec;   // [2][16] pre-evaluated truncated-normal draws (d3d_rng.cuh SP_*) of scalar warp B (a, c_old, w_old, a_new, c_new, w_new, log_u, oob)
};

__host__ __device__ inline size_t smem_doubles(int fh, int fw, int P, int Dp) {
    return (size_t)fh * fw + P + (P + 1) / 2 + 2 * (size_t)P + 4 * (size_t)Dp + 32 * 8 + 8 + 16 + 32;
}

__device__ __forceinline__ void carve(Smem& s, double* base, const Problem& pb) {
    s.F = base;
    s.Kv = s.F + pb.fh * pb.fw;
    s.Km = (int*)(s.Kv + pb.P);
    s.g_o = s.Kv + pb.P + (pb.P + 1) / 2;      // [P] each (P >= Dp), zero beyond D
    s.g_n = s.g_o + pb.P;
    s.Lu_o = s.g_n + pb.P;
    s.Lu_n = s.Lu_o + 2 * pb.Dp;
    s.red = s.Lu_n + 2 * pb.Dp;
    s.bc = s.red + 32 * 8;
    s.prop = s.bc + 8;
    s.spec = s.prop + 16;
}

__device__ __forceinline__ void load_constants(Smem& s, const Problem& pb) {
    for (int i = threadIdx.x; i < pb.fh * pb.fw; i += blockDim.x) s.F[i] = pb.fsf[i];
    for (int i = threadIdx.x; i < pb.ntaps; i += blockDim.x) {
        s.Kv[i] = pb.ktap_v[i];
        s.Km[i] = pb.ktap_m[i];
    }
    for (int i = threadIdx.x; i < pb.P; i += blockDim.x) { s.g_o[i] = 0.0; s.g_n[i] = 0.0; }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Components 1 .. n_comp-1 of a tied multiplet at distance d = z - c from the first component:
// sum_k ratio_k exp(-(d - off_k)^2 q), q = 1/(2 w^2).  Out of line and only reached when a
// multiplet is set (d3d_set_line_model): the single-Gaussian paths keep their exact arithmetic.
__device__ __noinline__ double extra_components(const Problem& pb, double d, double q) {
    double s = 0.0;
    for (int k = 1; k < pb.n_comp; ++k) {
        const double dk = d - pb.comp_off[k];
        s = fma(pb.comp_ratio[k], d_exp(-1.0 * (dk * dk) * q), s);
    }
    return s;
}

// lib/line_models.py:98-109 with a = 1
__device__ __forceinline__ double unit_gaussian(const Problem& pb, int z, double c, double w) {
    double d = (double)z - c;
    double g = d_exp(d_div(-1.0 * (d * d), 2.0 * (w * w)));
    if (pb.n_comp > 1) g += extra_components(pb, d, d_div(1.0, 2.0 * (w * w)));
    return g;
}
// Same with 1/(2 w^2) hoisted out of the channel loop (one divide per line instead of one per
// channel; the exponent differs from the reference's by at most one ulp).
__device__ __forceinline__ double unit_gaussian_r(const Problem& pb, int z, double c, double inv2w2) {
    double d = (double)z - c;
    double g = d_exp(-1.0 * (d * d) * inv2w2);
    if (pb.n_comp > 1) g += extra_components(pb, d, inv2w2);
    return g;
}

// lib/convolution.py:89-120 in direct form: out[j] = sum_i g[i] K[(j-i) mod P]
__device__ __forceinline__ double conv_at(const double* g, const double* K, int j, int D, int P) {
    double acc = 0.0;
    for (int i = 0; i < D; ++i) acc = fma(g[i], K[(j - i) & (P - 1)], acc);
    return acc;
}

enum { R_B = 0, R_C, R_PO, R_QOO, R_QON, R_QNN, R_A, R_N };

#ifdef D3D_TRACE
// Debug build only: absolute clock64() stamps of named events for a few consecutive sites.
__device__ unsigned long long g_trace[64 * 32];
#define TR(ev)                                                                              \
    do { if (lane == 0 && blockIdx.x == 0 && it == it0 + 3 && j >= 700 && j < 764)          \
             g_trace[(j - 700) * 32 + (ev)] = clock64(); } while (0)
#else
#define TR(ev)
#endif

#ifdef D3D_PHASE_TIMING
// Debug build only: per-phase clock64() accumulators (window warp 0 / scalar warps, lane 0).
__device__ unsigned long long g_phase[32];
#define PH_T0() unsigned long long ph_t = clock64()
#define PH_ADD(k)                                                                  \
    do { unsigned long long n_ = clock64(); if (lane == 0 && blockIdx.x == 0)      \
             atomicAdd(&g_phase[k], n_ - ph_t); ph_t = n_; } while (0)
#else
#define PH_T0()
#define PH_ADD(k)
#endif

// Spectral convolution on the taps of the circular LSF kernel that matter
// (|K[m]| >= 1e-18 max|K|, listed by the host): out[j] = sum_m K[m] g[(j-m) mod P].
// Four independent partial sums break the DFMA dependency chain.
__device__ __forceinline__ double conv_taps(const double* g, const double* Kv, const int* Km,
                                            int ntaps, int j, int D, int P) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int t = 0;
    for (; t + 3 < ntaps; t += 4) {
        int i0 = (j - Km[t]) & (P - 1), i1 = (j - Km[t + 1]) & (P - 1);
        int i2 = (j - Km[t + 2]) & (P - 1), i3 = (j - Km[t + 3]) & (P - 1);
        a0 = fma(Kv[t], i0 < D ? g[i0] : 0.0, a0);
        a1 = fma(Kv[t + 1], i1 < D ? g[i1] : 0.0, a1);
        a2 = fma(Kv[t + 2], i2 < D ? g[i2] : 0.0, a2);
        a3 = fma(Kv[t + 3], i3 < D ? g[i3] : 0.0, a3);
    }
    for (; t < ntaps; ++t) {
        int i0 = (j - Km[t]) & (P - 1);
        a0 = fma(Kv[t], i0 < D ? g[i0] : 0.0, a0);
    }
    return (a0 + a1) + (a2 + a3);
}

// Two outputs (j0, j1) at once: twice the independent work per tap.
__device__ __forceinline__ void conv_taps2(const double* g, const double* Kv, const int* Km,
                                           int ntaps, int j0, int j1, int D, int P, double& o0,
                                           double& o1) {
    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
    int t = 0;
    for (; t + 1 < ntaps; t += 2) {
        const int m0 = Km[t], m1 = Km[t + 1];
        const double k0 = Kv[t], k1 = Kv[t + 1];
        const int i00 = (j0 - m0) & (P - 1), i01 = (j0 - m1) & (P - 1);
        const int i10 = (j1 - m0) & (P - 1), i11 = (j1 - m1) & (P - 1);
        a0 = fma(k0, i00 < D ? g[i00] : 0.0, a0);
        a1 = fma(k1, i01 < D ? g[i01] : 0.0, a1);
        b0 = fma(k0, i10 < D ? g[i10] : 0.0, b0);
        b1 = fma(k1, i11 < D ? g[i11] : 0.0, b1);
    }
    if (t < ntaps) {
        const int m0 = Km[t];
        const double k0 = Kv[t];
        const int i00 = (j0 - m0) & (P - 1), i10 = (j1 - m0) & (P - 1);
        a0 = fma(k0, i00 < D ? g[i00] : 0.0, a0);
        b0 = fma(k0, i10 < D ? g[i10] : 0.0, b0);
    }
    o0 = a0 + a1;
    o1 = b0 + b1;
}

// Unit-amplitude line profile Lu = lsf (*) gaussian(c, w) computed by ONE warp:
// g -> shared (scratch), then the convolution; out[z] for z in [0, Dp).
__device__ __noinline__ void warp_line_profile(const Problem& pb, const Smem& sm, double c,
                                                  double w, double* g, double* out, int lane,
                                                  int ph_base = -1) {
#ifdef D3D_PHASE_TIMING
    unsigned long long ph_t = clock64();
#define PH_SUB(k) do { unsigned long long n_ = clock64(); if (ph_base >= 0 && lane == 0 && blockIdx.x == 0) atomicAdd(&g_phase[ph_base + k], n_ - ph_t); ph_t = n_; } while (0)
#else
#define PH_SUB(k)
#endif
    const double inv2w2 = d_div(1.0, 2.0 * (w * w));
    PH_SUB(0);
    // two channels per lane and per pass: independent exp() chains overlap
    for (int z = lane; z < pb.Dp; z += 64) {
        const int z1 = z + 32;
        const double g0 = z < pb.D ? unit_gaussian_r(pb, z, c, inv2w2) : 0.0;
        const double g1 = z1 < pb.D ? unit_gaussian_r(pb, z1, c, inv2w2) : 0.0;
        g[z] = g0;
        if (z1 < pb.Dp) g[z1] = g1;
    }
    PH_SUB(1);
    __syncwarp();
    if (pb.has_lsf) {
        for (int z = lane; z < pb.Dp; z += 64) {
            const int z1 = z + 32;
            double o0, o1;
            conv_taps2(g, sm.Kv, sm.Km, pb.ntaps, z, z1, pb.D, pb.P, o0, o1);
            out[z] = z < pb.D ? o0 : 0.0;
            if (z1 < pb.Dp) out[z1] = z1 < pb.D ? o1 : 0.0;
        }
    } else {                                                        // lib/run.py:675-676
        for (int z = lane; z < pb.Dp; z += 32) out[z] = g[z];
    }
    PH_SUB(2);
}

// Residual update coefficient of one channel: e += F * (a * L_old - r * L_end).  Written with
// explicit roundings so that every kernel (and the applier of remote tile updates,
// d3d_tile.cuh) produces the same bits.
__device__ __forceinline__ double upd_coef(double a, double lo, double r, double le) {
    return fma(a, lo, -__dmul_rn(r, le));
}

// Proposal of one site (lib/run.py:370-388, 570-579), evaluated redundantly by
// every lane of the calling warp.
struct Proposal {
    double a, c_old, w_old, a_new, c_new, w_new, log_u;
    int oob;
};

__device__ __noinline__ void make_proposal(const Problem& pb, int chain, int cube, int site,
                                              unsigned sweep, const EvalReq& ev, Philox& rng,
                                              Proposal& p) {
    const double* prm = pb.params + ((size_t)chain * pb.H * pb.W + site) * 3;
    p.a = prm[0]; p.c_old = prm[1]; p.w_old = prm[2];
    p.a_new = p.a; p.log_u = 0.0;
    if (ev.enabled) {
        p.a_new = ev.p_new[0]; p.c_new = ev.p_new[1]; p.w_new = ev.p_new[2];
    } else {
        rng.init(pb.seed, pb.first_chain + (unsigned)chain, sweep, (unsigned)site);
        const double q4 = 1.5707963267948966;                       // CIRCLE_4TH, lib/run.py:35
        const double u0 = rng.next(), u1 = rng.next(), u2 = rng.next();
        // numpy's uniform(-q4, q4) is low + (high - low) * u; amplitude jump is 0 (:262)
        if (pb.jump[0] != 0.0) p.a_new = p.a + pb.jump[0] * d_tan(-q4 + (q4 - (-q4)) * u0);
        p.c_new = p.c_old + pb.jump[1] * d_tan(-q4 + (q4 - (-q4)) * u1);
        p.w_new = p.w_old + pb.jump[2] * d_tan(-q4 + (q4 - (-q4)) * u2);
        p.log_u = d_log(rng.next());                                  // :435
    }
    const double* lo = pb.pmin + cube * 3;
    const double* hi = pb.pmax + cube * 3;
    p.oob = (p.a_new < lo[0]) | (p.c_new < lo[1]) | (p.w_new < lo[2]) |
            (p.a_new > hi[0]) | (p.c_new > hi[1]) | (p.w_new > hi[2]);   // :379-384
}

__device__ __forceinline__ void stash_proposal(const Smem& sm, const Proposal& p) {
    sm.prop[0] = p.a; sm.prop[1] = p.c_old; sm.prop[2] = p.w_old; sm.prop[3] = p.a_new;
    sm.prop[4] = p.c_new; sm.prop[5] = p.w_new; sm.prop[6] = p.log_u; sm.prop[7] = (double)p.oob;
}
__device__ __forceinline__ void fetch_proposal(const Smem& sm, Proposal& p) {
    p.a = sm.prop[0]; p.c_old = sm.prop[1]; p.w_old = sm.prop[2]; p.a_new = sm.prop[3];
    p.c_new = sm.prop[4]; p.w_new = sm.prop[5]; p.log_u = sm.prop[6]; p.oob = sm.prop[7] != 0.0;
}

// Accept test + Gibbs draw (lib/run.py:426-451, 456-499) from the reduced sums.
// Called by all lanes of one warp with identical arguments; lane 0 writes.
// Returns 1 when the proposal is accepted.
__device__ __noinline__ int decide(const Problem& pb, const Smem& sm, int chain, int cube,
                                      int site, const Proposal& p, const double* tot, Philox& rng,
                                      double* chain_row, double* lik_row, const EvalReq& ev,
                                      int lane, const double* spec = nullptr) {
    const double a = p.a, da = p.a_new;
    double delta;
    if (da != a) {
        // amplitude moved too: dL = a Lu_o - a' Lu_n
        const double Pn = tot[R_PO] - tot[R_B];
        const double Bq = a * tot[R_PO] - da * Pn;
        const double Cq = a * a * tot[R_QOO] - 2.0 * a * da * tot[R_QON] + da * da * tot[R_QNN];
        delta = -Bq - 0.5 * Cq;
    } else {
        delta = -(a * tot[R_B]) - 0.5 * (a * a) * tot[R_C];
    }
    if (ev.enabled) {
        if (lane == 0) {
            const double ar_old = 0.5 * tot[R_A];
            ev.out[0] = delta; ev.out[1] = ar_old; ev.out[2] = ar_old - delta;
            sm.bc[0] = 0.0;
        }
        return 0;
    }
    const int accepted = (p.log_u < delta) && !p.oob;               // :438
    const double c_end = accepted ? p.c_new : p.c_old;
    const double w_end = accepted ? p.w_new : p.w_old;
    const double S2 = accepted ? tot[R_QNN] : tot[R_QOO];
    const double S1 = accepted ? (tot[R_PO] - tot[R_B]) + a * tot[R_QON]
                               : tot[R_PO] + a * tot[R_QOO];
    int fail = 0;
    double r;
    if (spec) {
        // ro = ra/(1+ra S2) = 1/q with q = 1/ra + S2; mu = ro S1 = S1/q; sigma = sqrt(ro)
        // (:491-496).  sqrt(q) and S1/q are independent: the serial chain is one divide long
        // instead of divide -> sqrt -> divide.
        const double q = spec[SP_IRA] + S2;
        const double isg = d_sqrt(q);
        const double mu = d_div(S1, q);
        const double sigma = d_div(1.0, isg);
        r = rtnorm_spec(pb.pmin[cube * 3], pb.pmax[cube * 3], mu, isg, sigma, rng, pb.rt, &fail,
                        spec);
    } else {
        const double ra = pb.prior_var[cube];                       // :491
        const double ro = d_div(ra, 1.0 + ra * S2);                 // :492
        const double mu = ro * S1;                                  // :493
        r = rtnorm(pb.pmin[cube * 3], pb.pmax[cube * 3], mu, d_sqrt(ro), rng, pb.rt,
                   &fail);                                          // :495-496
    }
    if (lane == 0) {
        if (fail) atomicExch(pb.status, 1);
        double* prm = pb.params + ((size_t)chain * pb.H * pb.W + site) * 3;
        prm[0] = r; prm[1] = c_end; prm[2] = w_end;                 // :448, :499, :516
        if (chain_row) { chain_row[0] = r; chain_row[1] = c_end; chain_row[2] = w_end; }
        if (lik_row) *lik_row = delta;                              // :430-432
        sm.bc[0] = accepted ? 1.0 : 0.0;
        sm.bc[1] = r;
        sm.bc[3] = a;
    }
    return accepted;
}

__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------
// ROW-MAPPED, WARP-SPECIALISED site update (the fast path, fw*Dp/VEC <= 512).
//
// CTA = NWT "window" threads + 2 "scalar" warps.  Window thread t owns column
// dx = t / ZL and z-vector zp = t % ZL of the FSF window and walks its NE >= fh
// rows: one 16-byte load of the residual (and of 1/variance) per row with a
// constant stride, no index arithmetic in the loop; the residual stays in
// registers until the update.  Scalar warp A computes the old line profile,
// scalar warp B the proposal (Philox, Cauchy jump), the new line profile and,
// after the reduction, the accept test and the truncated-normal Gibbs draw.
// The scalar work of a site overlaps the window loads of the same site and the
// update stores of the previous one.
//   barrier 1 (window warps only): stores of the previous update visible
//   B1: window sums h/g and both profiles ready   B2: warp partials in smem
//   B3: decision broadcast
// ---------------------------------------------------------------------------
template <typename T, bool IVCUBE, int NE, bool WANT_AR>
struct RowSite {
    typedef typename Vec<T>::V V;
    static const int VEC = Vec<T>::N;

    // per-thread constants of the mapping
    int ZL, nwt, dx, zp, lane, warp, nww;    // nww = number of window warps
    bool wt, winwarp, warpA, warpB;
    unsigned parity;                         // site counter: selects the profile buffers

    __device__ __forceinline__ void init(const Problem& pb) {
        const int tid = threadIdx.x;
        ZL = pb.Dp / VEC;
        nwt = pb.fw * ZL;
        nww = (nwt + 31) >> 5;
        lane = tid & 31; warp = tid >> 5;
        dx = tid / ZL; zp = tid - dx * ZL;
        wt = tid < nwt;
        winwarp = warp < nww;             // warp-uniform role test (named barrier is .aligned)
        warpA = warp == nww;
        warpB = warp == nww + 1;
        parity = 0;
    }

    // One site.  Returns (in scalar warp B only) whether the proposal was accepted.
    __device__ __forceinline__ int run(const Problem& pb, const Smem& sm, int chain, int cube,
                                       int site, unsigned sweep, double* chain_row,
                                       double* lik_row, const EvalReq& ev) {
        const int Dp = pb.Dp, W = pb.W, H = pb.H;
        const int y = site / W, x = site - y * W;
        const int y0 = max(y - pb.fhh, 0), y1 = min(y + pb.fhh + 1, H);
        const int x0 = max(x - pb.fhw, 0), x1 = min(x + pb.fhw + 1, W);
        const int wh = y1 - y0, ww = x1 - x0;
        const int oy = y0 - (y - pb.fhh), ox = x0 - (x - pb.fhw);
        const bool active = wt && dx < ww;
        int accepted = 0;
        double* Lu_o = sm.Lu_o + (parity & 1u) * Dp;
        double* Lu_n = sm.Lu_n + (parity & 1u) * Dp;
        ++parity;

        V ecache[NE > 0 ? NE : 1];
        double h[VEC], g[VEC], f2 = 0.0, A = 0.0;
#pragma unroll
        for (int v = 0; v < VEC; ++v) { h[v] = 0.0; g[v] = 0.0; }
        T* erow = nullptr;
        const double* frow = sm.F + oy * pb.fw + ox + dx;
        const size_t rstride = (size_t)W * Dp;

        PH_T0();
        if (winwarp) {
            bar_sync_named(1, nww * 32);          // previous update visible to every window warp
            if (warp == 0) PH_ADD(0);             // [0] named barrier
            if (active) {
                const size_t base = ((size_t)y0 * W + x0 + dx) * Dp + zp * VEC;
                erow = (T*)pb.err + (size_t)chain * H * W * Dp + base;
                const T* ivrow = IVCUBE ? (const T*)pb.iv + (size_t)cube * H * W * Dp + base
                                        : nullptr;
                const double ivs = IVCUBE ? 0.0 : pb.iv_scalar[cube];
                (void)ivs;
                const int CH = 4;
                const T* pe = erow;              // running row pointers: one 64-bit add per row
                const T* pv = ivrow;
                const double* pf = frow;
                if (NE == 0) {
                    // uncached variant: nothing is kept, rows are streamed 4 at a time
                    for (int i0 = 0; i0 < wh; i0 += CH) {
                        V ech[CH];
                        V ivch[IVCUBE ? CH : 1];
                        double fch[CH];
#pragma unroll
                        for (int j = 0; j < CH; ++j) {
                            if (i0 + j < wh) {
                                ech[j] = *(const V*)pe;
                                if (IVCUBE) ivch[j] = *(const V*)pv;
                                fch[j] = *pf;
                                pe += rstride;
                                if (IVCUBE) pv += rstride;
                                pf += pb.fw;
                            } else {
                                ech[j] = V();
                                if (IVCUBE) ivch[j] = V();
                                fch[j] = 0.0;
                            }
                        }
#pragma unroll
                        for (int j = 0; j < CH; ++j) {
                            const double f = fch[j];
                            double e[VEC];
                            unpack(ech[j], e);
                            if (IVCUBE) {
                                double w_[VEC];
                                unpack(ivch[j], w_);
                                const double ff = f * f;
#pragma unroll
                                for (int v = 0; v < VEC; ++v) {
                                    const double t = w_[v] * e[v];
                                    h[v] = fma(f, t, h[v]);
                                    g[v] = fma(ff, w_[v], g[v]);
                                    if (WANT_AR) A = fma(t, e[v], A);
                                }
                            } else {
#pragma unroll
                                for (int v = 0; v < VEC; ++v) {
                                    h[v] = fma(f, e[v], h[v]);
                                    if (WANT_AR) A = fma(e[v], e[v], A);
                                }
                                f2 = fma(f, f, f2);
                            }
                        }
                    }
                }
#pragma unroll
                for (int i0 = 0; i0 < NE; i0 += CH) {
                    V ivch[IVCUBE ? CH : 1];
#pragma unroll
                    for (int j = 0; j < CH; ++j) {
                        const int i = i0 + j;
                        if (i < NE) {
                            if (i < wh) {
                                ecache[i] = *(const V*)pe;
                                if (IVCUBE) ivch[j] = *(const V*)pv;
                                pe += rstride;
                                if (IVCUBE) pv += rstride;
                            } else {
                                ecache[i] = V();
                                if (IVCUBE) ivch[j] = V();
                            }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < CH; ++j) {
                        const int i = i0 + j;
                        if (i < NE) {
                            const double f = i < wh ? *pf : 0.0;
                            pf += pb.fw;
                            double e[VEC];
                            unpack(ecache[i], e);
                            if (IVCUBE) {
                                double w_[VEC];
                                unpack(ivch[j], w_);
                                const double ff = f * f;
#pragma unroll
                                for (int v = 0; v < VEC; ++v) {
                                    const double t = w_[v] * e[v];
                                    h[v] = fma(f, t, h[v]);
                                    g[v] = fma(ff, w_[v], g[v]);
                                    if (WANT_AR) A = fma(t, e[v], A);
                                }
                            } else {
#pragma unroll
                                for (int v = 0; v < VEC; ++v) {
                                    h[v] = fma(f, e[v], h[v]);
                                    if (WANT_AR) A = fma(e[v], e[v], A);
                                }
                                f2 = fma(f, f, f2);
                            }
                        }
                    }
                }
                if (!IVCUBE) {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) { h[v] *= ivs; g[v] = ivs * f2; }
                    A *= ivs;
                }
            }
        } else if (warpA) {
            const double* prm = pb.params + ((size_t)chain * H * W + site) * 3;
            warp_line_profile(pb, sm, prm[1], prm[2], sm.g_o, Lu_o, lane);
            PH_ADD(8);                            // [8] warp A prep
        } else if (warpB) {
            Philox rng;
            Proposal prop;
            make_proposal(pb, chain, cube, site, sweep, ev, rng, prop);
            if (lane == 0) stash_proposal(sm, prop);
            PH_ADD(9);                            // [9] warp B proposal
            warp_line_profile(pb, sm, prop.c_new, prop.w_new, sm.g_n, Lu_n, lane);
            PH_ADD(10);                           // [10] warp B new profile
        }
        if (winwarp && warp == 0) PH_ADD(1);      // [1] window loads + sums
        __syncthreads();                                                        // B1
        if (winwarp && warp == 0) PH_ADD(2);      // [2] wait B1
        if (warpB) PH_ADD(11);                    // [11] warp B wait B1

        if (winwarp) {
            double lo_v[VEC], ln_v[VEC];
            double part[R_N];
#pragma unroll
            for (int j = 0; j < R_N; ++j) part[j] = 0.0;
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                lo_v[v] = Lu_o[zp * VEC + v];
                ln_v[v] = Lu_n[zp * VEC + v];
                if (active) {
                    const double dl = lo_v[v] - ln_v[v];
                    part[R_B] = fma(dl, h[v], part[R_B]);
                    part[R_PO] = fma(lo_v[v], h[v], part[R_PO]);
                    part[R_C] = fma(dl * dl, g[v], part[R_C]);
                    part[R_QOO] = fma(lo_v[v] * lo_v[v], g[v], part[R_QOO]);
                    part[R_QON] = fma(lo_v[v] * ln_v[v], g[v], part[R_QON]);
                    part[R_QNN] = fma(ln_v[v] * ln_v[v], g[v], part[R_QNN]);
                }
            }
            if (WANT_AR) part[R_A] = active ? A : 0.0;
#pragma unroll
            for (int j = 0; j < R_N; ++j) {
                if (j == R_A && !WANT_AR) continue;
                const double s = warp_sum(part[j]);
                if (lane == 0) sm.red[warp * 8 + j] = s;
            }
        }
        if (winwarp && warp == 0) PH_ADD(3);      // [3] partials
        __syncthreads();                                                        // B2
        if (winwarp && warp == 0) PH_ADD(4);      // [4] wait B2
        if (warpB) PH_ADD(12);                    // [12] warp B wait B2

        if (warpB) {
            double tot[R_N];
#pragma unroll
            for (int j = 0; j < R_N; ++j) {
                if (j == R_A && !WANT_AR) { tot[j] = 0.0; continue; }
                tot[j] = warp_sum(lane < nww ? sm.red[lane * 8 + j] : 0.0);
            }
            Proposal prop;
            fetch_proposal(sm, prop);
            Philox rng;                      // draws 0..3 went into the proposal (make_proposal)
            rng.init(pb.seed, pb.first_chain + (unsigned)chain, sweep, (unsigned)site);
            rng.k = 4;
            PH_ADD(13);                           // [13] totals
            accepted = decide(pb, sm, chain, cube, site, prop, tot, rng, chain_row, lik_row, ev,
                              lane);
            PH_ADD(14);                           // [14] decide
        }
        __syncthreads();                                                        // B3
        if (winwarp && warp == 0) PH_ADD(5);      // [5] wait B3
        if (ev.enabled) return 0;

        if (active) {
            const int acc = sm.bc[0] != 0.0;
            const double r = sm.bc[1], a = sm.bc[3];
            double coef[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const double lo = Lu_o[zp * VEC + v];
                coef[v] = upd_coef(a, lo, r, acc ? Lu_n[zp * VEC + v] : lo);
            }
            // laundered base pointers: keeps the compiler from holding the NE row addresses of
            // the load phase in registers across the barriers
            T* pe = erow;
            const double* pf = frow;
            asm volatile("" : "+l"(pe));
            asm volatile("" : "+l"(pf));
            if (NE == 0) {
                for (int i0 = 0; i0 < wh; i0 += 4) {
                    V ech[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (i0 + j < wh) ech[j] = *(const V*)(pe + (size_t)j * rstride);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (i0 + j < wh) {
                            const double f = pf[j * pb.fw];
                            double e[VEC];
                            unpack(ech[j], e);
#pragma unroll
                            for (int v = 0; v < VEC; ++v) e[v] = fma(f, coef[v], e[v]);
                            V o;
                            pack(o, e);
                            *(V*)(pe + (size_t)j * rstride) = o;
                        }
                    }
                    pe += 4 * rstride;
                    pf += 4 * pb.fw;
                }
            }
#pragma unroll
            for (int i = 0; i < NE; ++i) {
                if (i < wh) {
                    const double f = *pf;
                    double e[VEC];
                    unpack(ecache[i], e);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) e[v] = fma(f, coef[v], e[v]);
                    V o;
                    pack(o, e);
                    *(V*)pe = o;
                    pe += rstride;
                    pf += pb.fw;
                }
            }
        }
        if (winwarp && warp == 0) PH_ADD(6);      // [6] update
        return accepted;
    }
};

// ---------------------------------------------------------------------------
// GENERIC site update (any FSF size): flat traversal of the clipped window with a
// block-stride loop; the residual is re-read (L1/L2) for the update.  All warps
// take part in every phase.
// ---------------------------------------------------------------------------
template <typename T, bool IVCUBE, bool WANT_AR>
__device__ __forceinline__ int site_update_generic(const Problem& pb, const Smem& sm, int chain,
                                                   int cube, int site, unsigned int sweep,
                                                   double* chain_row, double* lik_row,
                                                   const EvalReq& ev) {
    typedef typename Vec<T>::V V;
    const int VEC = Vec<T>::N;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = (blockDim.x + 31) >> 5;
    const int Dp = pb.Dp, W = pb.W, H = pb.H;
    const int y = site / W, x = site - y * W;

    Philox rng;
    Proposal prop;
    if (warp == 0) {
        make_proposal(pb, chain, cube, site, sweep, ev, rng, prop);
        warp_line_profile(pb, sm, prop.c_new, prop.w_new, sm.g_n, sm.Lu_n, lane);
    } else if (warp == 1 || nwarps == 1) {
        const double* prm = pb.params + ((size_t)chain * H * W + site) * 3;
        warp_line_profile(pb, sm, prm[1], prm[2], sm.g_o, sm.Lu_o, lane);
    }
    if (warp == 0 && nwarps == 1) {
        const double* prm = pb.params + ((size_t)chain * H * W + site) * 3;
        warp_line_profile(pb, sm, prm[1], prm[2], sm.g_o, sm.Lu_o, lane);
    }

    const int y0 = max(y - pb.fhh, 0), y1 = min(y + pb.fhh + 1, H);
    const int x0 = max(x - pb.fhw, 0), x1 = min(x + pb.fhw + 1, W);
    const int ww = x1 - x0, npos = (y1 - y0) * ww;
    const int oy = y0 - (y - pb.fhh), ox = x0 - (x - pb.fhw);
    const int ZL = Dp / VEC;
    const int NC = blockDim.x / ZL;
    const int col = tid / ZL, zp = tid - col * ZL;
    const bool worker = col < NC;
    const int stepy = NC / ww, stepx = NC - stepy * ww;

    T* err = (T*)pb.err + (size_t)chain * H * W * Dp;
    const T* ivc = IVCUBE ? (const T*)pb.iv + (size_t)cube * H * W * Dp : nullptr;
    const double ivs = IVCUBE ? 0.0 : pb.iv_scalar[cube];

    double h[VEC], g[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) { h[v] = 0.0; g[v] = 0.0; }
    double f2 = 0.0, A = 0.0;
    if (worker) {
        int dy = col / ww, dx = col - dy * ww;
        for (int q = col; q < npos; q += NC) {
            const size_t off = ((size_t)(y0 + dy) * W + (x0 + dx)) * Dp + zp * VEC;
            const double f = sm.F[(oy + dy) * pb.fw + ox + dx];
            double e[VEC];
            unpack(*(const V*)(err + off), e);
            if (IVCUBE) {
                double w_[VEC];
                unpack(*(const V*)(ivc + off), w_);
                const double ff = f * f;
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    const double t = w_[v] * e[v];
                    h[v] = fma(f, t, h[v]);
                    g[v] = fma(ff, w_[v], g[v]);
                    if (WANT_AR) A = fma(t, e[v], A);
                }
            } else {
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    h[v] = fma(f, e[v], h[v]);
                    if (WANT_AR) A = fma(e[v], e[v], A);
                }
                f2 = fma(f, f, f2);
            }
            dx += stepx; dy += stepy;
            if (dx >= ww) { dx -= ww; ++dy; }
        }
        if (!IVCUBE) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) { h[v] *= ivs; g[v] = ivs * f2; }
            A *= ivs;
        }
    }
    __syncthreads();                                   // profiles ready

    double part[R_N];
#pragma unroll
    for (int j = 0; j < R_N; ++j) part[j] = 0.0;
    double lo_v[VEC], ln_v[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) { lo_v[v] = 0.0; ln_v[v] = 0.0; }
    if (worker) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            lo_v[v] = sm.Lu_o[zp * VEC + v];
            ln_v[v] = sm.Lu_n[zp * VEC + v];
            const double dl = lo_v[v] - ln_v[v];
            part[R_B] = fma(dl, h[v], part[R_B]);
            part[R_PO] = fma(lo_v[v], h[v], part[R_PO]);
            part[R_C] = fma(dl * dl, g[v], part[R_C]);
            part[R_QOO] = fma(lo_v[v] * lo_v[v], g[v], part[R_QOO]);
            part[R_QON] = fma(lo_v[v] * ln_v[v], g[v], part[R_QON]);
            part[R_QNN] = fma(ln_v[v] * ln_v[v], g[v], part[R_QNN]);
        }
        if (WANT_AR) part[R_A] = A;
    }
#pragma unroll
    for (int j = 0; j < R_N; ++j) {
        if (j == R_A && !WANT_AR) continue;
        const double s = warp_sum(part[j]);
        if (lane == 0) sm.red[warp * 8 + j] = s;
    }
    __syncthreads();

    int accepted = 0;
    if (warp == 0) {
        double tot[R_N];
#pragma unroll
        for (int j = 0; j < R_N; ++j) {
            if (j == R_A && !WANT_AR) { tot[j] = 0.0; continue; }
            tot[j] = warp_sum(lane < nwarps ? sm.red[lane * 8 + j] : 0.0);
        }
        accepted = decide(pb, sm, chain, cube, site, prop, tot, rng, chain_row, lik_row, ev, lane);
    }
    __syncthreads();
    if (ev.enabled) return 0;
    const int acc = sm.bc[0] != 0.0;
    const double r = sm.bc[1], a = sm.bc[3];

    if (worker) {
        double coef[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) coef[v] = upd_coef(a, lo_v[v], r, acc ? ln_v[v] : lo_v[v]);
        int dy = col / ww, dx = col - dy * ww;
        for (int q = col; q < npos; q += NC) {
            const size_t off = ((size_t)(y0 + dy) * W + (x0 + dx)) * Dp + zp * VEC;
            const double f = sm.F[(oy + dy) * pb.fw + ox + dx];
            double e[VEC];
            unpack(*(const V*)(err + off), e);
#pragma unroll
            for (int v = 0; v < VEC; ++v) e[v] = fma(f, coef[v], e[v]);
            V o;
            pack(o, e);
            *(V*)(err + off) = o;
            dx += stepx; dy += stepy;
            if (dx >= ww) { dx -= ww; ++dy; }
        }
    }
    __syncthreads();                                   // stores visible to the next site
    return accepted;                                   // valid in warp 0
}

// Shared chain-control logic of the SEQ_EXACT kernels (lib/run.py:344-359).
// `decider`: the thread that owns the accepted counter.
#define D3D_SEQ_PROLOGUE()                                                                  \
    extern __shared__ double smem_raw[];                                                    \
    Smem sm;                                                                                \
    carve(sm, smem_raw, pb);                                                                \
    const int chain = blockIdx.x;                                                           \
    if (chain >= pb.n_chains) return;                                                       \
    const int cube = chain / pb.chains_per_cube;                                            \
    if (!pb.active[chain]) return;                                                          \
    load_constants(sm, pb);                                                                 \
    if (threadIdx.x == 0) sm.bc[2] = (double)pb.accepted[chain];                            \
    __syncthreads();                                                                        \
    const int ns = pb.n_sites[cube];                                                        \
    const int* sites = pb.sites + (size_t)cube * pb.max_sites;                              \
    double rate = pb.rate[chain];                                                           \
    const size_t HW = (size_t)pb.H * pb.W;                                                  \
    EvalReq ev; ev.enabled = 0; ev.out = nullptr;                                           \
    long long it = it0;                                                                     \
    int alive = 1;

// ---------------------------------------------------------------------------
// SEQ_EXACT: one CTA per chain walks the masked spaxels in the reference's
// row-major order for iterations [it0, it1) (lib/run.py:344-537).
// ---------------------------------------------------------------------------
template <typename T, bool IVCUBE, int NE>
__global__ void __launch_bounds__(384, 1)
sweep_seq_kernel(const __grid_constant__ Problem pb, long long it0, long long it1, int keep, double min_rate,
                 double* chain_out, double* lik_out, long long row_first, long long rows_local) {
    D3D_SEQ_PROLOGUE()
    RowSite<T, IVCUBE, NE, false> rs;
    rs.init(pb);
    long long accepted = pb.accepted[chain];           // tracked by scalar warp B
    for (; it < it1; ++it) {
        if (!(rate > min_rate || rate == 0.0)) { alive = 0; break; }   // :344-350
        const double max_acc = (double)ns * (double)it;                // :356-359
        if (max_acc > 0.0) rate = sm.bc[2] / max_acc;
        const bool save = (it % keep) == 0;                            // :353
        double* crow = nullptr; double* lrow = nullptr;
        if (save) {
            long long r = it / keep - row_first;
            if (chain_out) crow = chain_out + ((size_t)chain * rows_local + r) * HW * 3;
            if (lik_out) lrow = lik_out + ((size_t)chain * rows_local + r) * HW;
        }
        for (int s = 0; s < ns; ++s) {
            const int site = sites[s];
            accepted += rs.run(pb, sm, chain, cube, site, (unsigned)it,
                               crow ? crow + (size_t)site * 3 : nullptr,
                               lrow ? lrow + site : nullptr, ev);
            if (s == ns - 1 && rs.warpB && rs.lane == 0) sm.bc[2] = (double)accepted;
        }
        __syncthreads();                               // bc[2] of this sweep visible
    }
    if (rs.warpB && rs.lane == 0) {
        pb.accepted[chain] = accepted;
        pb.rate[chain] = rate;
        pb.iters[chain] = it;
        if (!alive) pb.active[chain] = 0;
    }
}

// Uncached row-mapped variant (NE = 0): ~half the registers, two CTAs per SM; used when many
// chains share the GPU so that the serial phases of one chain hide behind another's.
template <typename T, bool IVCUBE>
__global__ void __launch_bounds__(384, 2)
sweep_seq_nc_kernel(const __grid_constant__ Problem pb, long long it0, long long it1, int keep, double min_rate,
                    double* chain_out, double* lik_out, long long row_first, long long rows_local) {
    D3D_SEQ_PROLOGUE()
    RowSite<T, IVCUBE, 0, false> rs;
    rs.init(pb);
    long long accepted = pb.accepted[chain];
    for (; it < it1; ++it) {
        if (!(rate > min_rate || rate == 0.0)) { alive = 0; break; }
        const double max_acc = (double)ns * (double)it;
        if (max_acc > 0.0) rate = sm.bc[2] / max_acc;
        const bool save = (it % keep) == 0;
        double* crow = nullptr; double* lrow = nullptr;
        if (save) {
            long long r = it / keep - row_first;
            if (chain_out) crow = chain_out + ((size_t)chain * rows_local + r) * HW * 3;
            if (lik_out) lrow = lik_out + ((size_t)chain * rows_local + r) * HW;
        }
        for (int s = 0; s < ns; ++s) {
            const int site = sites[s];
            accepted += rs.run(pb, sm, chain, cube, site, (unsigned)it,
                               crow ? crow + (size_t)site * 3 : nullptr,
                               lrow ? lrow + site : nullptr, ev);
            if (s == ns - 1 && rs.warpB && rs.lane == 0) sm.bc[2] = (double)accepted;
        }
        __syncthreads();
    }
    if (rs.warpB && rs.lane == 0) {
        pb.accepted[chain] = accepted;
        pb.rate[chain] = rate;
        pb.iters[chain] = it;
        if (!alive) pb.active[chain] = 0;
    }
}

template <typename T, bool IVCUBE>
__global__ void sweep_seq_generic_kernel(const __grid_constant__ Problem pb, long long it0, long long it1, int keep,
                                         double min_rate, double* chain_out, double* lik_out,
                                         long long row_first, long long rows_local) {
    D3D_SEQ_PROLOGUE()
    long long accepted = pb.accepted[chain];           // tracked by warp 0
    for (; it < it1; ++it) {
        if (!(rate > min_rate || rate == 0.0)) { alive = 0; break; }
        const double max_acc = (double)ns * (double)it;
        if (max_acc > 0.0) rate = sm.bc[2] / max_acc;
        const bool save = (it % keep) == 0;
        double* crow = nullptr; double* lrow = nullptr;
        if (save) {
            long long r = it / keep - row_first;
            if (chain_out) crow = chain_out + ((size_t)chain * rows_local + r) * HW * 3;
            if (lik_out) lrow = lik_out + ((size_t)chain * rows_local + r) * HW;
        }
        for (int s = 0; s < ns; ++s) {
            const int site = sites[s];
            accepted += site_update_generic<T, IVCUBE, false>(
                pb, sm, chain, cube, site, (unsigned)it, crow ? crow + (size_t)site * 3 : nullptr,
                lrow ? lrow + site : nullptr, ev);
            if (s == ns - 1 && threadIdx.x == 0) sm.bc[2] = (double)accepted;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        pb.accepted[chain] = accepted;
        pb.rate[chain] = rate;
        pb.iters[chain] = it;
        if (!alive) pb.active[chain] = 0;
    }
}

// ---------------------------------------------------------------------------
// COLOURED: one launch per colour class (cy, cx) = (y mod fh, x mod fw); the
// windows of the sites of one class are pairwise disjoint, so one CTA per
// (site, chain) updates them concurrently.  sweep_begin_kernel evaluates the
// loop condition of lib/run.py:344-359 once per iteration.
// ---------------------------------------------------------------------------
__global__ void sweep_begin_kernel(const __grid_constant__ Problem pb, long long it, double min_rate) {
    int chain = blockIdx.x * blockDim.x + threadIdx.x;
    if (chain >= pb.n_chains || !pb.active[chain]) return;
    double rate = pb.rate[chain];
    if (!(rate > min_rate || rate == 0.0)) { pb.active[chain] = 0; pb.iters[chain] = it; return; }
    double max_acc = (double)pb.n_sites[chain / pb.chains_per_cube] * (double)it;
    if (max_acc > 0.0) pb.rate[chain] = (double)pb.accepted[chain] / max_acc;
    pb.iters[chain] = it + 1;
}

// (programmatic dependent launch between the colour phases of one sweep, launch_colour_class: the
// next phase may be scheduled while this one drains; it sets up its constants and waits in front of
// the site update, the first thing that reads what a neighbouring phase writes)
#define D3D_COLOUR_PROLOGUE()                                                               \
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");                         \
    extern __shared__ double smem_raw[];                                                    \
    Smem sm;                                                                                \
    carve(sm, smem_raw, pb);                                                                \
    const int chain = blockIdx.y;                                                           \
    const int cube = chain / pb.chains_per_cube;                                            \
    const int iy = blockIdx.x / nlx, ix = blockIdx.x - iy * nlx;                            \
    const int y = cy + iy * pb.fh, x = cx + ix * pb.fw;                                     \
    if (y >= pb.H || x >= pb.W) return;                                                     \
    if (y < pb.ty0 || y >= pb.ty1 || x < pb.tx0 || x >= pb.tx1) return;                     \
    if (!pb.active[chain]) return;                                                          \
    const int site = y * pb.W + x;                                                          \
    if (pb.mask[(size_t)cube * pb.H * pb.W + site] != 1) return;                            \
    load_constants(sm, pb);                                                                 \
    __syncthreads();                                                                        \
    asm volatile("griddepcontrol.wait;" ::: "memory");                                      \
    const size_t HW = (size_t)pb.H * pb.W;                                                  \
    double* crow = crow_base                                                                \
        ? crow_base + (((size_t)chain * rows_local + row_local) * HW + site) * 3 : nullptr; \
    double* lrow = lrow_base                                                                \
        ? lrow_base + ((size_t)chain * rows_local + row_local) * HW + site : nullptr;       \
    EvalReq ev; ev.enabled = 0; ev.out = nullptr;

template <typename T, bool IVCUBE, int NE>
__global__ void __launch_bounds__(384, 1)
sweep_colour_kernel(const __grid_constant__ Problem pb, long long it, int cy, int cx, int nlx, double* crow_base,
                    double* lrow_base, long long rows_local, long long row_local) {
    D3D_COLOUR_PROLOGUE()
    RowSite<T, IVCUBE, NE, false> rs;
    rs.init(pb);
    int acc = rs.run(pb, sm, chain, cube, site, (unsigned)it, crow, lrow, ev);
    if (rs.warpB && rs.lane == 0) {
        if (acc) atomicAdd((unsigned long long*)&pb.accepted[chain], 1ull);
        if (pb.acc_cur) pb.acc_cur[(size_t)chain * HW + site] = (uint8_t)acc;
    }
}

template <typename T, bool IVCUBE>
__global__ void sweep_colour_generic_kernel(const __grid_constant__ Problem pb, long long it, int cy, int cx, int nlx,
                                            double* crow_base, double* lrow_base,
                                            long long rows_local, long long row_local) {
    D3D_COLOUR_PROLOGUE()
    int acc = site_update_generic<T, IVCUBE, false>(pb, sm, chain, cube, site, (unsigned)it, crow,
                                                    lrow, ev);
    if (threadIdx.x == 0) {
        if (acc) atomicAdd((unsigned long long*)&pb.accepted[chain], 1ull);
        if (pb.acc_cur) pb.acc_cur[(size_t)chain * HW + site] = (uint8_t)acc;
    }
}

template <typename T, bool IVCUBE>
__global__ void eval_kernel(const __grid_constant__ Problem pb, int chain, int site, EvalReq ev) {
    extern __shared__ double smem_raw[];
    Smem sm;
    carve(sm, smem_raw, pb);
    load_constants(sm, pb);
    __syncthreads();
    site_update_generic<T, IVCUBE, true>(pb, sm, chain, chain / pb.chains_per_cube, site, 0u,
                                         nullptr, nullptr, ev);
}

// ---------------------------------------------------------------------------
// Forward model (lib/run.py:999-1031 / 623-652)
// ---------------------------------------------------------------------------
// Pass 1 (spectral): lines[chain][y][x][Dp] = mask * a * (lsf (*) gaussian(c,w)); one warp
// per spaxel, the Gaussian is exchanged through shared memory.
__global__ void lines_kernel(const __grid_constant__ Problem pb, const double* params, double* lines, int convolve) {
    // one thread per (spaxel, channel): per-spaxel constants by the channel-0 thread, Gaussian
    // -> shared memory (zero-padded circular buffer of length P per spaxel) -> the taps of the
    // circular LSF kernel -> lines[chain][y][x][z]
    extern __shared__ double smem_raw[];
    double* Kv = smem_raw;                        // [P] tap values
    int* Km = (int*)(Kv + pb.P);                  // [P] tap offsets
    double* g = Kv + pb.P + (pb.P + 1) / 2;       // [spb][P]
    const int D = pb.D, Dp = pb.Dp, P = pb.P, nt = pb.ntaps;
    const int spb = blockDim.x / Dp;              // spaxels per block and pass
    double* par = g + (size_t)spb * P;            // [spb][4]: a, c, 1/(2w^2), on
    const int ls = threadIdx.x / Dp, z = threadIdx.x - ls * Dp;
    const bool mine = ls < spb;
    for (int i = threadIdx.x; i < nt; i += blockDim.x) { Kv[i] = pb.ktap_v[i]; Km[i] = pb.ktap_m[i]; }
    for (int i = threadIdx.x; i < spb * P; i += blockDim.x) g[i] = 0.0;
    const unsigned HW = (unsigned)(pb.H * pb.W);
    const unsigned total = (unsigned)pb.n_chains * HW;      // (host guarantees < 2^31)
    double* gw = g + (size_t)ls * P;
    for (unsigned base = blockIdx.x * spb; base < total; base += gridDim.x * spb) {
        const unsigned sp = base + ls;
        const bool live = mine && sp < total;
        __syncthreads();
        if (live && z == 0) {
            const unsigned chain = sp / HW, site = sp - chain * HW;
            const bool on = pb.mask[(size_t)(chain / pb.chains_per_cube) * HW + site] == 1;
            const double* p = params + (size_t)sp * 3;
            const double w = p[2];
            par[ls * 4 + 0] = p[0]; par[ls * 4 + 1] = p[1];
            par[ls * 4 + 2] = 1.0 / (2.0 * (w * w));
            par[ls * 4 + 3] = on ? 1.0 : 0.0;
        }
        __syncthreads();
        const bool on = live && par[ls * 4 + 3] != 0.0;
        if (on && z < D) {
            const double d = (double)z - par[ls * 4 + 1];
            double gv = exp(-1.0 * (d * d) * par[ls * 4 + 2]);                   // lib/line_models.py:109
            if (pb.n_comp > 1) gv += extra_components(pb, d, par[ls * 4 + 2]);
            gw[z] = par[ls * 4 + 0] * gv;
        }
        __syncthreads();
        if (live) {
            double v = 0.0;
            if (on && z < D) {
                if (convolve && pb.has_lsf) {
                    double a0 = 0.0, a1 = 0.0;
                    int t = 0;
                    for (; t + 1 < nt; t += 2) {
                        a0 = fma(Kv[t], gw[(z - Km[t]) & (P - 1)], a0);
                        a1 = fma(Kv[t + 1], gw[(z - Km[t + 1]) & (P - 1)], a1);
                    }
                    if (t < nt) a0 = fma(Kv[t], gw[(z - Km[t]) & (P - 1)], a0);
                    v = a0 + a1;
                } else {
                    v = gw[z];
                }
            }
            lines[(size_t)sp * Dp + z] = v;
        }
    }
}

// Warp-shuffle form of the spectral pass (same arithmetic and summation order as
// lines_kernel, so the two are bit-identical): one warp per spaxel, lane l keeps channels
// l, l+32, ... of the zero-padded circular buffer (length P <= 32*R) in registers; tap m of
// the circular LSF kernel reads channel (z - m) mod P.  The taps are ascending in m and form
// at most two runs of consecutive offsets (an arc of the ring that wraps at most once), so
// within a run the buffer is simply rotated by one channel per tap: per tap one 8-byte LDS
// (tap value), R 64-bit shuffles with a loop-invariant source lane and R DFMA.  No
// shared-memory buffer, no block barrier and no integer division inside the spaxel loop
// (a CTA works on LINES_SPB consecutive spaxels of ONE chain); a warp stores 32 consecutive
// channels (256 bytes) per instruction.  LSF vectors with gaps (more than two runs) take
// the shared-memory kernel.
#define LINES_SPB 32
template <int R>
__global__ void __launch_bounds__(256) lines_warp_kernel(const __grid_constant__ Problem pb,
                                                         const double* __restrict__ params,
                                                         double* __restrict__ lines, int convolve,
                                                         int run2 /* first tap of the second run, or ntaps */) {
    __shared__ double Kv[32 * R];
    const int D = pb.D, Dp = pb.Dp, P = pb.P, nt = pb.ntaps;
    for (int i = threadIdx.x; i < nt; i += blockDim.x) Kv[i] = pb.ktap_v[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const unsigned HW = (unsigned)(pb.H * pb.W);
    const unsigned nb = (HW + LINES_SPB - 1) / LINES_SPB;   // CTAs per chain
    const unsigned chain = blockIdx.x / nb;
    const unsigned site0 = (blockIdx.x - chain * nb) * LINES_SPB;
    const uint8_t* mask = pb.mask + (size_t)(chain / pb.chains_per_cube) * HW;
    const bool conv = convolve && pb.has_lsf;
    const int src1 = (lane - 1) & (P - 1) & 31;             // rotate by one channel
    const int m0 = nt > 0 ? pb.ktap_m[0] : 0, m1 = run2 < nt ? pb.ktap_m[run2] : 0;
    for (unsigned i = threadIdx.x >> 5; i < LINES_SPB; i += wpb) {
        const unsigned site = site0 + i;
        if (site >= HW) break;
        const size_t sp = (size_t)chain * HW + site;
        const bool on = mask[site] == 1;
        const double* p = params + sp * 3;
        const double a = p[0], c = p[1], w = p[2];
        const double q = 1.0 / (2.0 * (w * w));
        double g[R], v[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int z = 32 * k + lane;
            g[k] = 0.0;
            if (on && z < D) {
                const double d = (double)z - c;
                double gv = exp(-1.0 * (d * d) * q);                          // lib/line_models.py:109
                if (pb.n_comp > 1) gv += extra_components(pb, d, q);
                g[k] = a * gv;
            }
        }
        if (conv) {
            double a0[R], a1[R];                             // taps of even / odd index (lines_kernel)
#pragma unroll
            for (int k = 0; k < R; ++k) a0[k] = a1[k] = 0.0;
            for (int run = 0; run < 2; ++run) {
                const int tb = run ? run2 : 0, te = run ? nt : run2;
                if (tb >= te) continue;
                // r[k][lane] = g[(32k + lane - m) mod P] for the first tap m of the run
                const int j = (lane - (run ? m1 : m0)) & (P - 1);
                double r[R];
                {
                    double s[R];
#pragma unroll
                    for (int k = 0; k < R; ++k) s[k] = __shfl_sync(0xffffffffu, g[k], j & 31);
#pragma unroll
                    for (int k = 0; k < R; ++k) {
                        double x = s[0];
#pragma unroll
                        for (int rr = 1; rr < R; ++rr) x = (((k + (j >> 5)) & (R - 1)) == rr) ? s[rr] : x;
                        r[k] = x;
                    }
                }
                auto tap = [&](int t, double (&acc)[R]) {
                    const double kv = Kv[t];
#pragma unroll
                    for (int k = 0; k < R; ++k) acc[k] = fma(kv, r[k], acc[k]);
                    double s[R];
#pragma unroll
                    for (int k = 0; k < R; ++k) s[k] = __shfl_sync(0xffffffffu, r[k], src1);
#pragma unroll
                    for (int k = 0; k < R; ++k) r[k] = (R > 1 && lane == 0) ? s[(k + R - 1) % R] : s[k];
                };
                int t = tb;
                if (t & 1) { tap(t, a1); ++t; }
                for (; t + 1 < te; t += 2) { tap(t, a0); tap(t + 1, a1); }
                if (t < te) tap(t, a0);
            }
#pragma unroll
            for (int k = 0; k < R; ++k) v[k] = a0[k] + a1[k];
        } else {
#pragma unroll
            for (int k = 0; k < R; ++k) v[k] = g[k];
        }
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int z = 32 * k + lane;
            if (z < Dp) lines[sp * Dp + z] = (on && z < D) ? v[k] : 0.0;
        }
    }
}

// Pass 2 (spatial): true 2-D convolution with the FSF, zero 'same' borders
// (scipy.signal.convolve2d(..., 'same'), lib/run.py:1027-1029, equal to the paste
// at lib/run.py:697-706).  CTA = TY x TX output spaxels x ZC channels; the
// (TY+fh-1) x (TX+fw-1) halo tile of `lines` is staged in shared memory once and
// every output accumulates fh*fw taps from it.  Fused epilogue: residual =
// data - sim, optional sim output in the reference layout, optional chi^2.
template <typename T>
__global__ void stencil_zchunk_kernel(const __grid_constant__ Problem pb, const double* lines, double* sim_out,
                                      int write_err, double* chi2_out, int TY, int TX, int ZC) {
    extern __shared__ double smem_raw[];
    const int fh = pb.fh, fw = pb.fw, Dp = pb.Dp, D = pb.D, H = pb.H, W = pb.W;
    double* F = smem_raw;                                  // [fh*fw]
    double* tile = F + fh * fw;                            // [hy][hx][zc]
    const int ty_n = (H + TY - 1) / TY, tx_n = (W + TX - 1) / TX;
    const int chain = blockIdx.x / (ty_n * tx_n);
    const int trem = blockIdx.x - chain * ty_n * tx_n;
    const int ty0 = (trem / tx_n) * TY, tx0 = (trem % tx_n) * TX;
    const int z0 = blockIdx.y * ZC;
    const int zc = min(ZC, Dp - z0);
    const int cube = chain / pb.chains_per_cube;
    const int hy = TY + fh - 1, hx = TX + fw - 1;
    for (int i = threadIdx.x; i < fh * fw; i += blockDim.x) F[i] = pb.fsf[i];
    const double* lc = lines + (size_t)chain * H * W * Dp;
    for (int i = threadIdx.x; i < hy * hx * zc; i += blockDim.x) {
        int z = i % zc, s = i / zc;
        int sy = s / hx, sx = s - sy * hx;
        int gy = ty0 + sy - pb.fhh, gx = tx0 + sx - pb.fhw;
        double v = 0.0;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = lc[((size_t)gy * W + gx) * Dp + z0 + z];
        tile[i] = v;
    }
    __syncthreads();
    const T* data = (const T*)pb.data + (size_t)cube * H * W * Dp;
    const T* ivc = pb.var_is_cube ? (const T*)pb.iv + (size_t)cube * H * W * Dp : nullptr;
    const double ivs = pb.var_is_cube ? 0.0 : pb.iv_scalar[cube];
    T* err = (T*)pb.err + (size_t)chain * H * W * Dp;
    double chi = 0.0;
    for (int i = threadIdx.x; i < TY * TX * zc; i += blockDim.x) {
        int z = i % zc, s = i / zc;
        int oy = s / TX, ox = s - oy * TX;
        int gy = ty0 + oy, gx = tx0 + ox;
        if (gy >= H || gx >= W) continue;
        // sim[y',x'] = sum_{j,k} F[j][k] lines[y'+fhh-j][x'+fhw-k]   (lib/run.py:697-706)
        double acc = 0.0;
        for (int j = 0; j < fh; ++j) {
            const double* trow = tile + ((size_t)(oy + fh - 1 - j) * hx + ox + fw - 1) * zc + z;
            const double* frow = F + j * fw;
            for (int k = 0; k < fw; ++k) acc = fma(frow[k], trow[-(ptrdiff_t)k * zc], acc);
        }
        const int zz = z0 + z;
        size_t off = ((size_t)gy * W + gx) * Dp + zz;
        if (sim_out && zz < D)
            sim_out[(size_t)chain * D * H * W + ((size_t)zz * H + gy) * W + gx] = acc;
        if (write_err || chi2_out) {
            double e = (double)data[off] - acc;
            if (zz >= D) e = 0.0;
            if (write_err) err[off] = (T)e;
            if (chi2_out) chi += e * e * (ivc ? (double)ivc[off] : ivs);
        }
    }
    if (chi2_out) {
        chi = warp_sum(chi);
        if ((threadIdx.x & 31) == 0) atomicAdd(chi2_out + chain, 0.5 * chi);
    }
}

// un-convolved lines in the reference layout (lib/run.py:597-621)
__global__ void clean_kernel(const __grid_constant__ Problem pb, const double* params, double* out) {
    const size_t HW = (size_t)pb.H * pb.W;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)pb.n_chains * pb.D * HW;
    if (i >= total) return;
    size_t site = i % HW;
    size_t r = i / HW;
    int z = (int)(r % pb.D);
    int chain = (int)(r / pb.D);
    int cube = chain / pb.chains_per_cube;
    double v = 0.0;
    if (pb.mask[(size_t)cube * HW + site] == 1) {
        const double* p = params + ((size_t)chain * HW + site) * 3;
        v = p[0] * unit_gaussian(pb, z, p[1], p[2]);
    }
    out[i] = v;
}

// ---------------------------------------------------------------------------
// Spectral pass, ONE THREAD PER SPAXEL (lib/run.py:1012-1024: line model, then convolve_1d with
// the LSF).  The whole spectrum of a spaxel lives in the registers of its thread: DP_T Gaussian
// values, then out[z] = sum_{|m| <= MH} K[m mod P] * g[(z - m) mod P] -- the circular
// convolution of lib/convolution.py:89-160 in direct form -- fully unrolled with compile-time
// indices (P = smallest power of two >= DP_T, exactly the reference's padded length), so the
// inner work is DFMA only: no shared-memory traffic, no shuffles, no per-tap index arithmetic.
// The warp-per-spaxel kernel above spends ~380 warp instructions per spaxel (taps through
// shuffles, one exp per lane); this one ~70, of which 32 exp() calls.  Output rows are written
// with 16-byte stores (a thread owns DP_T contiguous doubles).
// Taken when every significant tap lies within MH channels and 2 MH + 1 <= P.
// ---------------------------------------------------------------------------
template <int N> struct NextPow2 { static const int value = N <= 1 ? 1 : 2 * NextPow2<(N + 1) / 2>::value; };
template <> struct NextPow2<1> { static const int value = 1; };
template <> struct NextPow2<0> { static const int value = 1; };

// exp(x) for x <= 0, table driven: x = (64 n' + j) ln2/64 + r, |r| <= ln2/128, so that
// exp(x) = 2^n' * 2^(j/64) * exp(r) with a degree-5 polynomial for exp(r) (|r|^6/720 < 4e-17).
// About 14 instructions instead of ~45 for the library routine; relative error <= ~2 ulp.
// Results below 2^-995 are flushed to zero (the tail of a line profile, far below the 1e-12
// relative bar of the convolved cubes).  tab64[j] = 2^(j/64) sits in shared memory.
__device__ __forceinline__ double exp_neg_tab(double x, const double* tab64) {
    const double MAGIC = 6755399441055744.0;                  // 1.5 * 2^52: round to nearest integer
    const double t = fma(x, 92.33248261689366, MAGIC);        // 64 / ln 2
    const int ni = __double2loint(t);
    const double nd = t - MAGIC;
    double r = fma(nd, -1.0830424696223417e-02, x);           // ln2/64, high part
    r = fma(nd, -2.5728046223276688e-14, r);                  // ln2/64, low part
    double p = fma(r, 8.3333333333333332e-03, 4.1666666666666664e-02);
    p = fma(p, r, 1.6666666666666666e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double v = tab64[ni & 63] * p;
    const int e = ni >> 6;                                    // floor(ni / 64) <= 0
    const double scaled = __hiloint2double(__double2hiint(v) + (e << 20), __double2loint(v));
    // x <= -690: below 2^-995, flushed to zero (also keeps ni inside 32 bits for any finite x);
    // NaN (a line of width 0 sitting on a channel: 0 * inf, as in the reference) stays NaN
    return x > -690.0 ? scaled : (x <= -690.0 ? 0.0 : x);
}

template <int DP_T, int MH>
__global__ void __launch_bounds__(128)
lines_lane_kernel(const __grid_constant__ Problem pb, const double* __restrict__ params,
                  double* __restrict__ lines, int convolve) {
    constexpr int P = NextPow2<DP_T>::value;
    constexpr int NT = 2 * MH + 1;
    static_assert(NT <= P, "the tap window must not wrap onto itself");
    __shared__ double sK[NT];
    __shared__ double tab64[64];
    for (int t = threadIdx.x; t < NT; t += blockDim.x) sK[t] = pb.kcirc[((t - MH) % P + P) % P];
    for (int t = threadIdx.x; t < 64; t += blockDim.x) tab64[t] = exp2((double)t * (1.0 / 64.0));
    __syncthreads();
    const size_t HW = (size_t)pb.H * pb.W;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)pb.n_chains * HW) return;
    const size_t chain = i / HW, site = i - chain * HW;
    const bool on = pb.mask[(size_t)(chain / pb.chains_per_cube) * HW + site] == 1;
    const double* p = params + i * 3;
    const double a = p[0], c = p[1], w = p[2];
    const double q = 1.0 / (2.0 * (w * w));
    const int D = pb.D;
    double g[DP_T];
#pragma unroll
    for (int ch = 0; ch < DP_T; ++ch) {
        const double d = (double)ch - c;
        double gv = exp_neg_tab(-1.0 * (d * d) * q, tab64);                 // lib/line_models.py:109
        if (pb.n_comp > 1) gv += extra_components(pb, d, q);
        g[ch] = (on && ch < D) ? a * gv : 0.0;
    }
    double* out = lines + i * DP_T;
    if (!(convolve && pb.has_lsf)) {                                        // lib/run.py:675-676
#pragma unroll
        for (int z = 0; z < DP_T; z += 2) *(double2*)(out + z) = make_double2(g[z], g[z + 1]);
        return;
    }
    double k[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) k[t] = sK[t];
#pragma unroll
    for (int z = 0; z < DP_T; z += 2) {
        double o[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                // tap m = t - MH reads channel (z - m) mod P; channels >= DP_T are zero padding
                const int ch = (((z + u) - (t - MH)) % P + P) % P;
                if (ch < DP_T) {
                    if (t & 1) a1 = fma(k[t], g[ch], a1); else a0 = fma(k[t], g[ch], a0);
                }
            }
            o[u] = (z + u < D) ? a0 + a1 : 0.0;
        }
        *(double2*)(out + z) = make_double2(o[0], o[1]);
    }
}

// ---------------------------------------------------------------------------
// Layout conversion between the reference layout [n][D][H][W] float64 and the
// device layout [n][H][W][Dp] T.
//   mode 0: plain copy (data; NaN -> 0)          mode 1: reciprocal (variance -> 1/var;
//   NaN/zero-weight voxels -> 0, and voxels whose data value is NaN -> 0)
// ---------------------------------------------------------------------------
template <typename T>
__global__ void ingest_kernel(const double* src, const double* data_for_nan, T* dst, int n, int D,
                              int Dp, int H, int W, int mode, int* nan_seen = nullptr) {
    // tile transpose through shared memory: read x-fastest, write z-fastest
    __shared__ double tile[32][33];
    const int nzt = (Dp + 31) / 32, nxt = (W + 31) / 32;
    long long b = blockIdx.x;
    const int xt = (int)(b % nxt); b /= nxt;
    const int zt = (int)(b % nzt); b /= nzt;
    const int y = (int)(b % H); b /= H;
    const int cube = (int)b;
    if (cube >= n) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        int z = zt * 32 + r, x = xt * 32 + tx;
        double v = 0.0;
        if (z < D && x < W) {
            size_t si = (((size_t)cube * D + z) * H + y) * W + x;
            v = src[si];
            if (mode == 1) {
                double dv = data_for_nan ? data_for_nan[si] : 0.0;
                v = (v == v && dv == dv) ? 1.0 / v : 0.0;
                if (!(v == v) || isinf(v)) v = 0.0;
            } else if (!(v == v)) { v = 0.0; if (nan_seen) *nan_seen = 1; }
        }
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        int x = xt * 32 + r, z = zt * 32 + tx;
        if (x < W && z < Dp)
            dst[(((size_t)cube * H + y) * W + x) * Dp + z] = (T)tile[tx][r];
    }
}

// device layout -> reference layout, float64 out
template <typename T>
__global__ void egest_kernel(const T* src, double* dst, int n, int D, int Dp, int H, int W) {
    __shared__ double tile[32][33];
    const int nzt = (Dp + 31) / 32, nxt = (W + 31) / 32;
    long long b = blockIdx.x;
    const int xt = (int)(b % nxt); b /= nxt;
    const int zt = (int)(b % nzt); b /= nzt;
    const int y = (int)(b % H); b /= H;
    const int cube = (int)b;
    if (cube >= n) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        int x = xt * 32 + r, z = zt * 32 + tx;
        double v = 0.0;
        if (x < W && z < Dp) v = (double)src[(((size_t)cube * H + y) * W + x) * Dp + z];
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        int z = zt * 32 + r, x = xt * 32 + tx;
        if (z < D && x < W) dst[(((size_t)cube * D + z) * H + y) * W + x] = tile[tx][r];
    }
}

// Initial parameters uniform in the boundaries (lib/run.py:308-314), sweep 0.
__global__ void init_params_kernel(const __grid_constant__ Problem pb) {
    const size_t HW = (size_t)pb.H * pb.W;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)pb.n_chains * HW) return;
    const int chain = (int)(i / HW);
    const int site = (int)(i - (size_t)chain * HW);
    const int cube = chain / pb.chains_per_cube;
    double* p = pb.params + i * 3;
    if (pb.mask[(size_t)cube * HW + site] != 1) { p[0] = p[1] = p[2] = 0.0; return; }
    Philox rng;
    rng.init(pb.seed, pb.first_chain + (unsigned)chain, 0u, (unsigned)site);
    for (int j = 0; j < 3; ++j) {
        double lo = pb.pmin[cube * 3 + j], hi = pb.pmax[cube * 3 + j];
        p[j] = lo + (hi - lo) * rng.next();
    }
}

// Batched truncated normal (lib/rtnorm.py:21-92), one variate per thread.
__global__ void rtnorm_kernel(RtTables rt, int n, const double* a, const double* b,
                              const double* mu, const double* sigma, unsigned long long seed,
                              unsigned int chain, unsigned int sweep, double* out, int* used,
                              int* status) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Philox rng;
    rng.init(seed, chain, sweep, (unsigned)i);
    int fail = 0;
    out[i] = rtnorm(a[i], b[i], mu[i], sigma[i], rng, rt, &fail);
    if (used) used[i] = (int)rng.k;
    if (fail) atomicExch(status, 1);
}

// Batched spectral convolution (lib/convolution.py:89-120): out[b][j].
__global__ void conv1d_kernel(const double* lines, const double* kcirc, double* out, int n, int P,
                              int batch) {
    extern __shared__ double smem_raw[];
    double* K = smem_raw;
    double* g = K + P;
    for (int i = threadIdx.x; i < P; i += blockDim.x) K[i] = kcirc[i];
    for (int b = blockIdx.x; b < batch; b += gridDim.x) {
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) g[i] = lines[(size_t)b * n + i];
        __syncthreads();
        for (int j = threadIdx.x; j < n; j += blockDim.x)
            out[(size_t)b * n + j] = conv_at(g, K, j, n, P);
    }
}

}  // namespace d3d

#include "d3d_slide.cuh"
#include "d3d_pipe.cuh"
#include "d3d_stencil.cuh"
#include "d3d_tile.cuh"
