// d3d_pipe.cuh -- SEQ_EXACT sweep, decisions pipelined through the linearity of the window sums.
//
// The register-resident sliding window of d3d_slide.cuh removed the memory traffic of the
// reference's row-major sweep (lib/run.py:553-566) but left one serial chain per site:
// window sums -> reduction -> accept test + Gibbs draw -> residual update -> next site's sums.
// This kernel cuts that chain.  The sums of site k are linear in the residual, and the update of
// an earlier site i is rank one, e += F_i (x) coef_i[z] with coef_i = a_i Lu_old,i - r_i L_end,i
// (lib/run.py:402, 441, 508-515).  Hence, with h0_k the sums taken from a residual that does NOT
// yet contain the updates of the last L sites,
//
//     h_k[z] = h0_k[z] + sum_{i = k-L .. k-1} coef_i[z] * X_{i,k}[z],
//     X_{i,k}[z] = sum over the voxels of both windows of F_i F_k / sigma^2     (static: FSF,
//                                                         variance and the pair of sites only)
//
// and the two sums the decision needs (header of d3d_kernels.cuh) become
//
//     sum_z T_k h_k = sum_z T_k h0_k + sum_i ( a_i <T_k X_ik Lu_old,i> - r_i <T_k X_ik L_end,i> ),
//     T_k in {Lu_old,k - Lu_new,k, Lu_old,k}.
//
// The brackets depend on line profiles only, not on any decision: they are formed ahead of time.
// What remains serial is scalar: decision k = f(S0_k, 4L brackets, (a, r, accepted) of the last L
// sites).  The window warps run L sites behind the decisions with their updates and ahead of them
// with their sums; nobody waits for a full sums -> decision -> update round trip any more.
//
// Roles (one CTA per chain, warps communicate through a ring of PIPE_R stages of shared memory
// guarded by mbarriers -- no CTA-wide barrier inside a sweep):
//   W  window warps: (fw + L + 1) column groups of Dp/VEC threads keep the residual of the band
//      [x-fhw-L, x+fhw+1] x fh rows in registers (1/sigma^2 of the same band sits in shared
//      memory, staged with cp.async).  Step k: move the band, sums h0_k -> partial products with
//      the profiles of k -> PART[k]; then the update of site k-L once its decision is there.
//   P  producers -> PROF[k]: draws (Philox), Cauchy jumps, accept uniform, first truncated-normal
//      tries, new line profile (the old one comes from the profile cache, Problem::lucache).
//      Shipped (PMODE 1): two warps, each preparing EIGHT sites per pass lane-parallel over the
//      sites, in lock step (GO barriers) because their code is instruction-cache cold at every
//      pass and the fetches go to the L1.5 all SMs of a GPC share.  PMODE 0 (one warp per site
//      and kind) is compiled with -DD3D_PIPE_PERSITE for A/B runs only.
//   X  quadratic sums over the static G table + the 4L brackets -> SCAL[k]
//   B  the serial scalar chain: totals + corrections, accept test, Gibbs draw -> DEC[k]; also
//      writes the outcome to the parameter map and the saved chain / likelihood rows
// FREE[k] (W after update k, X after its last use of the profiles of k) recycles a stage.
// Measurements behind every one of these choices: profiles/r02_notes.md.
//
// Pipelining runs along "runs" of the site list (consecutive sites of one row, x increasing by
// one); at the end of a run the window warps drain the pending updates, so any mask and any
// order stay exact -- they just do not overlap across the break.  The tables run_start[] and
// X_d[] are built once per problem by the host / xtab_kernel.
#pragma once

namespace d3d {

enum { PB_PROF = 0, PB_SCAL, PB_PART, PB_DEC, PB_FREE, PB_HSUM, PB_GO /* [0]: lock step of the producer warps */, PB_N };
#ifndef PIPE_HS
#define PIPE_HS 4                      // stages of the ring of raw window sums (> L + 1)
#endif
#ifdef D3D_PIPE_RROLE
#define PIPE_NR 1                      // a reducer warp between the window warps and warp B (measured slower: off)
#else
#define PIPE_NR 0
#endif
#ifndef PIPE_R
#define PIPE_R 32                      // stages of the ring: two batches of PIPE_B sites of each of the two producer warps
#endif
#define PIPE_B 8                       // sites a producer warp prepares per pass (lane-parallel over the sites)
#define PIPE_TIMEOUT 6000000000LL      // cycles (~3 s): a stalled wait aborts the launch, never hangs

__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_inval(unsigned long long* b) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_addr(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_addr(b)) : "memory");
}
__device__ __forceinline__ bool mbar_test(unsigned long long* b, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_addr(b)), "r"(parity) : "memory");
    return ok != 0;
}
// The same test with a suspend-time hint: the warp is parked by the hardware until the phase
// completes (or `ns` nanoseconds pass) instead of polling -- waiting warps must not compete with
// the working ones for issue slots.
__device__ __forceinline__ bool mbar_test_sleep(unsigned long long* b, unsigned parity, unsigned ns) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_addr(b)), "r"(parity), "r"(ns) : "memory");
    return ok != 0;
}
// cold path of a wait: bounded; a time-out raises the CTA's abort flag (every later wait then
// falls through) and the launch ends with an error status instead of hanging the GPU
__device__ __noinline__ void mbar_wait_slow(unsigned long long* b, unsigned parity, volatile int* abort_flag, int code) {
    const long long t0 = clock64();
    for (;;) {
#pragma unroll 1
        for (int q = 0; q < 64; ++q)
            if (mbar_test_sleep(b, parity, 100000u)) return;      // up to 100 us asleep per try
        if (*abort_flag) return;
        if (clock64() - t0 > PIPE_TIMEOUT) { *abort_flag = code; return; }
    }
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity, volatile int* abort_flag, int code) {
    if (mbar_test(b, parity)) return;
    mbar_wait_slow(b, parity, abort_flag, code);
}
#ifdef D3D_PIPE_PROF
// Debug build only (profiles/tools/ab_slide.py): per-warp cycle accounting of CTA 0 -- slot 0 the
// cycles spent inside the sweeps, slot c the cycles spent in the waits with code c.
__device__ unsigned long long g_pipe_prof[32 * 16];
__device__ unsigned long long g_pipe_cta[1024 * 2];      // per CTA: cycles inside the kernel, SM id
#define PP_DECL unsigned long long pp_[16]; for (int q_ = 0; q_ < 16; ++q_) pp_[q_] = 0ull
#define MWAIT(b, par, code)                                                     \
    do { const unsigned long long t_ = clock64(); mbar_wait(b, par, abort_flag, (code) | (j << 8)); \
         pp_[code] += clock64() - t_; } while (0)
#define PP_MARK(slot, t_) do { pp_[slot] += clock64() - (t_); } while (0)
#define PP_NOW() clock64()
#define PP_STAMP(slot, t_) do { const unsigned long long n_ = clock64(); pp_[slot] += n_ - (t_); (t_) = n_; } while (0)
#define PP_FLUSH()                                                              \
    do { if (blockIdx.x == 0 && lane == 0)                                      \
             for (int q_ = 0; q_ < 16; ++q_) atomicAdd(&g_pipe_prof[warp * 16 + q_], pp_[q_]); } while (0)
#else
#define PP_DECL
#define MWAIT(b, par, code) mbar_wait(b, par, abort_flag, (code) | (j << 8))   // (j: site index in scope)
#define PP_MARK(slot, t_)
#define PP_NOW() 0ull
#define PP_STAMP(slot, t_)
#define PP_FLUSH()
#endif

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_addr(smem_dst)), "l"(gmem_src) : "memory");
}
// same with `src_bytes` (0 or 16) read from global memory and the rest of the 16 bytes zero-filled
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gmem_src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr(smem_dst)), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Shared-memory layout of the pipelined kernel, in doubles.  Everything whose size is known at
// compile time comes first, so that its address is an immediate (no per-use address arithmetic
// on the serial paths); the arrays sized by the problem follow (pipe_var_layout).
enum { PREC_K0 = 0, PREC_A, PREC_AQOO, PREC_AQON, PREC_SGO, PREC_SGN, PREC_ISGO, PREC_ISGN, PREC_SG2O,
       PREC_SG2N, PREC_X = 10 /* 4 L cross terms */, PREC_RAW = 24 /* C, QOO, QON, QNN */, PREC_N = 28 };
template <int L> struct PipeFix {
    static const int PROP = 0;                         // [R][8]   a, c_old, w_old, a_new, c_new, w_new, -, oob
    static const int SPEC = PROP + PIPE_R * 8;         // [R][16]  d3d_rng.cuh SP_*
    static const int REC = SPEC + PIPE_R * 16;         // [R][PREC_N] what warp B needs of a site, by warp X
    static const int RED = REC + PIPE_R * PREC_N;      // [R][32]  two partial sums per window warp (<= 16 warps; or the totals of warp R)
    static const int DEC = RED + PIPE_R * 32;          // [R][4]   accepted, r, a, delta
    static const int BC = DEC + PIPE_R * 4;            // [8]      [2] accepted_count of the finished sweep
    static const int BAR = BC + 8;                     // [PB_N][R] mbarriers
    static const int ABORT = BAR + PB_N * PIPE_R;      // [2]
    static const int PROG = ABORT + 2;                 // [16]     per-warp progress (list index), reported when a wait times out
    static const int ARGS = PROG + 16;                 // [8]      PipeArgs
    static const int TX = ARGS + 8;                   // [4002]   truncated-normal tables (lib/rtnorm.py:227-2681)
    static const int TYU = TX + 4002;                  // [4002]
    static const int TNC = TYU + 4002;                 // [2242]   ncell as 16-bit
    static const int VAR = TNC + 2242;                 // start of the problem-sized part (even)
    static_assert(PREC_X + 4 * L <= PREC_RAW, "record too small for this look-ahead");
    static_assert(VAR % 2 == 0, "16-byte alignment");
};
struct PipeVar {           // offsets (doubles) from PipeFix::VAR
    int Ft;        // [fw][NE+1]     FSF transposed: a column is contiguous (16-byte pairs)
    int Kd;        // [kd_n padded]  dense window of the circular LSF kernel (host: pb.kdense)
    int G;         // [NA+NP][PIPE_B][GN]  per producer warp and site of the batch: periodically extended Gaussian (old, then new)
    int S;         // [NA+NP][256]   per producer warp: scratch of a batch (uniforms, tan, log, normals, parameters)
    int T64;       // [64]           2^(j/64) for exp_neg_tab
    int Lu_o;      // [R][Dp]
    int Lu_n;      // [R][Dp]
    int iv;        // [NG][NE][ZL]   16-byte vectors of 1/sigma^2 of the resident band
    int hs;        // [PIPE_HS][NG*ZL][VEC] raw window sums h0 of every window thread (read by warp R)
    int GN;        // length of one G buffer (even)
    int total;     // doubles
};
__host__ __device__ inline PipeVar pipe_var_layout(int fw, int NE, int kd_n, int Dp, int nprod, int NG, int ZL,
                                                   bool ivcube) {
    PipeVar v;
    int o = 0;
    const int kdp = (kd_n + 3) & ~1;                  // taps padded to an even count (+1 pair of zeros)
    v.Ft = o; o += fw * (NE + 1); o = (o + 1) & ~1;
    v.Kd = o; o += kdp;
    v.GN = (Dp + kdp + 2 + 1) & ~1;
    v.G = o; o += nprod * PIPE_B * v.GN;     // (per-site producers use the first GN of each warp's block)
    v.S = o; o += nprod * 256;
    v.T64 = o; o += 64;
    v.Lu_o = o; o += PIPE_R * Dp; o = (o + 1) & ~1;
    v.Lu_n = o; o += PIPE_R * Dp; o = (o + 1) & ~1;
    v.iv = o; o += NG * NE * ZL * 2;            // (a scalar variance fills the same ring)
    v.hs = o; o += PIPE_NR * PIPE_HS * NG * Dp;
    v.total = o;
    return v;
}
template <int L>
__host__ __device__ inline size_t pipe_smem_bytes(int fw, int NE, int kd_n, int Dp, int nprod, int NG, int ZL,
                                                  bool ivcube) {
    return ((size_t)PipeFix<L>::VAR + pipe_var_layout(fw, NE, kd_n, Dp, nprod, NG, ZL, ivcube).total) * sizeof(double);
}

// X_d[cube][d-1][j][z] for the pairs (j-d, j) of one run of the site list (zero elsewhere):
// sum over the voxels common to both windows of F_{j-d} F_j / sigma^2.  Both sites lie on one
// row, d columns apart.
template <typename T, bool IVCUBE>
__global__ void xtab_kernel(const __grid_constant__ Problem pb, int L, const int* run_start, double* xtab) {
    const int j = blockIdx.x, d = blockIdx.y + 1, cube = blockIdx.z;
    const int Dp = pb.Dp;
    double* out = xtab + (((size_t)cube * L + (d - 1)) * pb.max_sites + j) * Dp;
    const bool valid = j < pb.n_sites[cube] && j - d >= run_start[(size_t)cube * pb.max_sites + j];
    const int site = valid ? pb.sites[(size_t)cube * pb.max_sites + j] : 0;
    const int y = site / pb.W, x = site - y * pb.W;
    const size_t HW = (size_t)pb.H * pb.W;
    const T* iv = IVCUBE ? (const T*)pb.iv + (size_t)cube * HW * Dp : nullptr;
    const double ivs = IVCUBE ? 0.0 : pb.iv_scalar[cube];
    for (int z = threadIdx.x; z < Dp; z += blockDim.x) {
        double acc = 0.0;
        if (valid && z < pb.D) {
            for (int i = 0; i < pb.fh; ++i) {
                const int Y = y - pb.fhh + i;
                if (Y < 0 || Y >= pb.H) continue;
                // columns of the window of site j that also lie in the window of site j-d
                for (int k = 0; k + d < pb.fw; ++k) {
                    const int Xc = x - pb.fhw + k;
                    if (Xc < 0 || Xc >= pb.W) continue;
                    const double fj = pb.fsf[i * pb.fw + k], fi = pb.fsf[i * pb.fw + k + d];
                    const double w = IVCUBE ? (double)iv[((size_t)Y * pb.W + Xc) * Dp + z] : ivs;
                    acc = fma(fj * fi, w, acc);
                }
            }
        }
        out[z] = acc;
    }
}

// Cold path of warp B's Gibbs draw: every branch of lib/rtnorm.py:95-223 the inline fast paths do
// not cover, and their retries (generic sampler, draws 4..7 served from the pre-evaluated block).
__device__ __noinline__ double pipe_rtnorm_fallback(const Problem& pb, double as, double bs, const double* spec_s,
                                                    unsigned chain, unsigned sweep, unsigned site) {
    int fail = 0;
    Philox rng;
    rng.init(pb.seed, pb.first_chain + chain, sweep, site);
    rng.k = 4;
    rng.stash = spec_s;
    const double r = rtstdnorm(as, bs, rng, pb.rt, &fail, nullptr);
    if (fail) atomicExch(pb.status, 1);
    return r;
}

// One window thread drops its column (written back if it lies in the field) and takes column Xn of
// the rows around y.  Pointer stepping + one range test per row: the block is executed by one
// group per site and its code size matters (see the note on the I-cache).
template <typename T, bool IVCUBE, int NE>
__device__ __forceinline__ void pipe_switch_column(typename Vec<T>::V (&ecache)[NE], int& heldX, int& heldY, bool& colvalid,
                                                   int Xn, int y, bool wanted, int fh, int fhh, int H, int W, int Dp,
                                                   int zoff, size_t rstride, T* errT, const T* ivT,
                                                   typename Vec<T>::V* ivs_mine, int ZL, typename Vec<T>::V ivs_vec) {
    typedef typename Vec<T>::V V;
    if (colvalid) {                                  // the column that left the band is final
        const int ytop = heldY - fhh;
        const int ilo = max(0, -ytop);
        const unsigned span = (unsigned)(min(fh, H - ytop) - 1 - ilo);
        T* p = errT + ((long long)ytop * W + heldX) * Dp + zoff;
#pragma unroll
        for (int i = 0; i < NE; ++i) {
            if ((unsigned)(i - ilo) <= span) *(V*)p = ecache[i];
            p += rstride;
        }
    }
    heldX = Xn; heldY = y;
    colvalid = Xn >= 0 && Xn < W && wanted;
    if (colvalid) {
        const int ytop = y - fhh;
        const int ilo = max(0, -ytop);
        const unsigned span = (unsigned)(min(fh, H - ytop) - 1 - ilo);
        const long long cb = ((long long)ytop * W + Xn) * Dp + zoff;
        const T* p = errT + cb;
        const T* q = IVCUBE ? ivT + cb : nullptr;
        V* sd = ivs_mine;
#pragma unroll
        for (int i = 0; i < NE; ++i) {
            if ((unsigned)(i - ilo) <= span) {
                ecache[i] = *(const V*)p;
                if (IVCUBE) cp_async16(sd, q); else *sd = ivs_vec;
            } else {
                *sd = V();
            }
            p += rstride; sd += ZL;
            if (IVCUBE) q += rstride;
        }
        if (IVCUBE) cp_async_commit();
    }
}

// Kernel arguments every role needs (one copy in shared memory, filled at kernel entry).
struct PipeArgs {
    int keep;
    double min_rate;
    double* chain_out; double* lik_out;
    long long row_first, rows_local;
    long long it_first;            // first sweep of this launch (the profile cache is good from the second on)
};

// Every role is its own __noinline__ function: ptxas then allocates registers role by role (the
// window warps keep 52 registers of residual, the scalar roles need few), instead of one
// allocation for the whole kernel whose long live ranges spilled into every serial path.
// All roles share the skeleton below: same declarations, same sweep loop, same CTA barriers.
#define PIPE_ROLE_PARAMS const Problem& pb, const PipeArgs& ka, int chain, long long it0, long long it1, unsigned gbase
#define PIPE_ROLE_DECLS                                                                          \
    typedef typename Vec<T>::V V;                                                                \
    typedef PipeFix<L> FX;                                                                       \
    const int VEC = Vec<T>::N;                                                                   \
    const int R = PIPE_R;                                                                        \
    extern __shared__ double smem_raw[];                                                         \
    const int fw = SQ ? NE : pb.fw, fh = SQ ? NE : pb.fh;                                        \
    const int fhh = (fh - 1) / 2, fhw = (fw - 1) / 2;                                            \
    const int Dp = pb.Dp, W = pb.W, H = pb.H;                                                    \
    const int ZL = Dp / VEC;                                                                     \
    const int NG = fw + L + 1;                                                                   \
    const int FHP = NE + 1;                                                                      \
    const PipeVar pv = pipe_var_layout(fw, NE, pb.kd_n, Dp, NA + NP, NG, ZL, IVCUBE);            \
    double* const s_var = smem_raw + FX::VAR;                                                    \
    double* const s_prop = smem_raw + FX::PROP;                                                  \
    double* const s_spec = smem_raw + FX::SPEC;                                                  \
    double* const s_rec = smem_raw + FX::REC;                                                    \
    double* const s_red = smem_raw + FX::RED;                                                    \
    double* const s_dec = smem_raw + FX::DEC;                                                    \
    double* const s_bc = smem_raw + FX::BC;                                                      \
    unsigned long long* const s_bar = (unsigned long long*)(smem_raw + FX::BAR);                 \
    volatile int* const abort_flag = (volatile int*)(smem_raw + FX::ABORT);                      \
    volatile int* const s_prog = (volatile int*)(smem_raw + FX::PROG);                           \
    const double* const s_tx = smem_raw + FX::TX;                                                \
    const double* const s_tyu = smem_raw + FX::TYU;                                              \
    const unsigned short* const s_tnc = (const unsigned short*)(smem_raw + FX::TNC);             \
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;                               \
    const int nwt = NG * ZL, nww = (nwt + 31) >> 5;                                              \
    const int wA0 = nww, wP0 = nww + NA, wX0 = wP0 + NP, wR = wX0 + NX, wB = wR + PIPE_NR;       \
    const bool roleA = warp >= wA0 && warp < wP0, roleP = warp >= wP0 && warp < wX0;             \
    const int cube = chain / pb.chains_per_cube;                                                 \
    const int grp = tid / ZL, zp = tid - grp * ZL;                                               \
    const bool wt = tid < nwt;                                                                   \
    const int ns = pb.n_sites[cube];                                                             \
    const unsigned magicW = (unsigned)((0x100000000ull + (unsigned)W - 1u) / (unsigned)W);       \
    const int* sites = pb.sites + (size_t)cube * pb.max_sites;                                   \
    const int* run_start = pb.run_start + (size_t)cube * pb.max_sites;                           \
    const int* run_last = pb.run_last + (size_t)cube * pb.max_sites;                             \
    const size_t HW = (size_t)H * W;                                                             \
    T* const errT = (T*)pb.err + (size_t)chain * HW * Dp;                                        \
    const T* const ivT = IVCUBE ? (const T*)pb.iv + (size_t)cube * HW * Dp : nullptr;            \
    const double ivs = IVCUBE ? 0.0 : pb.iv_scalar[cube];                                        \
    const size_t rstride = (size_t)W * Dp;                                                       \
    const int keep = ka.keep; const double min_rate = ka.min_rate;                               \
    double* const chain_out = ka.chain_out; double* const lik_out = ka.lik_out;                  \
    const long long row_first = ka.row_first, rows_local = ka.rows_local;                        \
    (void)fhh; (void)fhw; (void)FHP; (void)s_prop; (void)s_spec; (void)s_rec; (void)s_red;       \
    (void)s_dec; (void)s_tx; (void)s_tyu; (void)s_tnc; (void)wB; (void)wR; (void)roleA;          \
    (void)roleP; (void)grp; (void)zp; (void)wt; (void)magicW; (void)sites; (void)run_start; (void)run_last;      \
    (void)errT; (void)ivT; (void)ivs; (void)rstride; (void)keep; (void)chain_out; (void)lik_out; \
    (void)row_first; (void)rows_local; (void)s_prog; (void)H; (void)pv; (void)HW; (void)s_var;   \
    PP_DECL;                                                                                     \
    double rate = pb.rate[chain];                                                                \
    long long it = it0;                                                                          \
    int alive = 1
// sweep loop: the acceptance-rate test of lib/run.py:344-359 (every thread forms the same rate
// from the count warp B published), then the role's work, then the CTA barrier of the boundary
#define PIPE_SWEEPS_BEGIN                                                                        \
    for (; it < it1; ++it) {                                                                     \
        if (!(rate > min_rate || rate == 0.0)) { alive = 0; break; }                             \
        const double max_acc = (double)ns * (double)it;                                          \
        if (max_acc > 0.0) rate = s_bc[2] / max_acc;                                             \
        if (*abort_flag) break;                                                                  \
        const unsigned long long pp_sweep0 = PP_NOW(); (void)pp_sweep0;
#define PIPE_SWEEPS_END                                                                          \
        PP_MARK(0, pp_sweep0);                                                                   \
        gbase += (unsigned)ns;                                                                   \
        __syncthreads();                                                                         \
    }                                                                                            \
    (void)alive; PP_FLUSH()
#define PIPE_TMPL template <typename T, bool IVCUBE, int NE, bool SQ, int L, int NA, int NP, int NX, int PMODE>
// Measured (profiles/r02_notes.md): as separate functions the roles reach the kernel parameters
// through a generic pointer and run ~12 % slower; inlined is the default.
#ifdef D3D_PIPE_NOINLINE_ROLES
#define PIPE_ROLE_FN __device__ __noinline__
#else
#define PIPE_ROLE_FN __device__ __forceinline__
#endif

// ---- W: window warps ------------------------------------------------------------------------
PIPE_TMPL PIPE_ROLE_FN unsigned pipe_role_W(PIPE_ROLE_PARAMS) {
    PIPE_ROLE_DECLS;
    // register-resident column of this window thread (rows y-fhh .. y-fhh+fh-1 of column heldX)
    V ecache[NE];
#pragma unroll
    for (int i = 0; i < NE; ++i) ecache[i] = V();  // (rows outside the field are never loaded: keep them finite)
    const int NOCOL = -(1 << 30);
    int heldX = NOCOL, heldY = NOCOL;
    bool colvalid = false;                         // a column inside the field is held
    // 1/sigma^2 of the held column, one 16-byte vector per row (a scalar variance fills the same
    // ring: the sums are one code path, and rows outside the field carry weight zero)
    V* const ivs_mine = (V*)(s_var + pv.iv) + (size_t)grp * NE * ZL + zp;   // + i * ZL per row
    V ivs_vec;
    { double t_[VEC]; for (int v = 0; v < VEC; ++v) t_[v] = ivs; pack(ivs_vec, t_); }
    PIPE_SWEEPS_BEGIN
        // =================================================================================
        int next_u = 0;                              // first site whose update is still pending
        int site = 0, x = 0, y = 0, m = 0;
        int j_last = -1;                             // last list entry of the current run
        for (int j = 0; j < ns; ++j) {
            if (*abort_flag) break;
            if (lane == 0) s_prog[warp] = j;
            unsigned long long pp_t = PP_NOW(); (void)pp_t;
            const bool fresh = j > j_last;           // first site of a run (no global load inside a run)
            const unsigned gj = gbase + (unsigned)j;        // running site count: ring stage + phase
            const int st = (int)(gj & (R - 1));
            const unsigned ph = (gj / R) & 1u;
            // The band [x-fhw-L, x+fhw+1] moves one column per site; group g keeps the column
            // xb + m, m = (g - xb) mod NG: inside a run m just counts down, and the group whose m
            // wraps takes the column that enters the band -- it made that switch at the END of
            // the previous step (below), off the path sums -> decision.
            if (fresh) {
                j_last = run_last[j];
                site = sites[j];
                y = (int)__umulhi((unsigned)site, magicW); x = site - y * W;
                m = (grp - (x - fhw - L)) % NG;
                if (m < 0) m += NG;
                next_u = j;                          // start of a run: nothing is pending
            } else {
                ++site; ++x;
                if (--m < 0) m = NG - 1;
            }
            const bool last = j == j_last;
            // (the group that switched at the end of the previous step holds m == NG-1 now and is idle
            // this step: its staged 1/sigma^2 vectors are first read one step later)
            if (m != NG - 1) cp_async_wait_all();
            if (wt && fresh) {
                // ---- start of a run: every group (re)loads its column; rows outside the field are
                // never tested again below: their 1/sigma^2 slots in shared memory are zero (so they
                // drop out of the sums), their residual registers keep whatever finite value they
                // had and are never written back
                const int Xn = x - fhw - L + m;          // column of this group at this site
                pipe_switch_column<T, IVCUBE, NE>(ecache, heldX, heldY, colvalid, Xn, y,
                                                  Xn >= x - fhw /* no trailing columns yet */, fh, fhh, H, W, Dp,
                                                  zp * VEC, rstride, errT, ivT, ivs_mine, ZL, ivs_vec);
                if (IVCUBE) cp_async_wait_all();         // needed by the sums of this very site
            }
            PP_STAMP(13, pp_t);                      // [13] head
            // ---- window sums h0_j (registers + shared memory only) ------------------------
            const int dxs = m - L;                   // FSF column of this group at site j
            const bool active = colvalid && dxs >= 0 && dxs < fw;
            double h[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) h[v] = 0.0;
            if (active) {
                const double2* fcol = (const double2*)(s_var + pv.Ft + dxs * FHP);   // rows (2k, 2k+1)
#pragma unroll
                for (int i = 0; i < NE; ++i) {
                    const double f = (i & 1) ? fcol[i >> 1].y : fcol[i >> 1].x;   // zero for i >= fh
                    double e[VEC], w_[VEC];
                    unpack(ecache[i], e);
                    unpack(ivs_mine[i * ZL], w_);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) h[v] = fma(f * w_[v], e[v], h[v]);
                }
            }
            PP_STAMP(10, pp_t);                      // [10] window sums
#ifdef D3D_PIPE_RROLE
            // the raw sums go to warp R, which multiplies them with the line profiles and
            // reduces them (no shuffle tree and no wait for the profiles in the window warps)
            if (wt) {
                double* hd = s_var + pv.hs + ((int)(gj & (PIPE_HS - 1)) * nwt + tid) * VEC;
#pragma unroll
                for (int v = 0; v < VEC; v += 2) *(double2*)(hd + v) = make_double2(h[v], h[v + 1]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(s_bar + PB_HSUM * R + st);
#else
            MWAIT(s_bar + PB_PROF * R + st, ph, 1);   // profiles of site j
            {
                double pB = 0.0, pPO = 0.0;
                if (active) {
                    const double* Lo = s_var + pv.Lu_o + st * Dp + zp * VEC;
                    const double* Ln = s_var + pv.Lu_n + st * Dp + zp * VEC;
#pragma unroll
                    for (int v = 0; v < VEC; v += 2) {
                        const double2 lo = *(const double2*)(Lo + v), ln = *(const double2*)(Ln + v);
                        pB = fma(lo.x - ln.x, h[v], pB);
                        pPO = fma(lo.x, h[v], pPO);
                        pB = fma(lo.y - ln.y, h[v + 1], pB);
                        pPO = fma(lo.y, h[v + 1], pPO);
                    }
                }
                const double tot = warp_sum2(pB, pPO, lane);
                if ((lane & 15) == 0) s_red[st * 32 + warp * 2 + (lane >> 4)] = tot;
                __syncwarp();
                if (lane == 0) mbar_arrive(s_bar + PB_PART * R + st);
            }
#endif
            PP_STAMP(11, pp_t);                      // [11] store + arrive
            // ---- pending updates: site j-L in the steady state, everything at a run end ----
            const int u_hi = last ? j : j - L;
            for (int i = next_u; i <= u_hi; ++i) {
                const unsigned gi = gbase + (unsigned)i;
                const int sti = (int)(gi & (R - 1));
                MWAIT(s_bar + PB_DEC * R + sti, (gi / R) & 1u, 2);
                const double* dc = s_dec + sti * 4;
                const bool acc = dc[0] != 0.0;
                const double r = dc[1], a = dc[2];
                const int dxu = m - L + (j - i);     // FSF column of this group at site i
                if (colvalid && dxu >= 0 && dxu < fw) {
                    const double* Lo = s_var + pv.Lu_o + sti * Dp + zp * VEC;
                    const double* Ln = s_var + pv.Lu_n + sti * Dp + zp * VEC;
                    double coef[VEC];
#pragma unroll
                    for (int v = 0; v < VEC; v += 2) {
                        const double2 lo = *(const double2*)(Lo + v), ln = *(const double2*)(Ln + v);
                        coef[v] = a * lo.x - r * (acc ? ln.x : lo.x);
                        coef[v + 1] = a * lo.y - r * (acc ? ln.y : lo.y);
                    }
                    const double2* fcol = (const double2*)(s_var + pv.Ft + dxu * FHP);
#pragma unroll
                    for (int i2 = 0; i2 < NE; ++i2) {
                        const double f = (i2 & 1) ? fcol[i2 >> 1].y : fcol[i2 >> 1].x;
                        double e[VEC];
                        unpack(ecache[i2], e);
#pragma unroll
                        for (int v = 0; v < VEC; ++v) e[v] = fma(f, coef[v], e[v]);
                        pack(ecache[i2], e);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(s_bar + PB_FREE * R + sti);
            }
            if (u_hi >= next_u) next_u = u_hi + 1;
            PP_STAMP(12, pp_t);                      // [12] updates (incl. the wait for the decision)
            // ---- the band moves on: the group whose column leaves it (m == 0; its last update has
            // just been applied) writes it back and takes the column that enters the WINDOW two
            // sites from now -- done here, behind the sums of this site, so that the (rarely
            // executed, instruction-cache-cold) switch never delays a decision
#ifdef D3D_KO_SWITCH      // timing experiment only (wrong results): what the steady-state column switch costs
            if (false)
#else
            if (wt && !last && m == 0)
#endif
                pipe_switch_column<T, IVCUBE, NE>(ecache, heldX, heldY, colvalid, x + fhw + 2, y, true, fh, fhh, H, W,
                                                  Dp, zp * VEC, rstride, errT, ivT, ivs_mine, ZL, ivs_vec);
            PP_STAMP(9, pp_t);                       // [9] column switch
        }
    PIPE_SWEEPS_END;
    // ---- write the resident columns back ----------------------------------------------------
    if (wt && colvalid) {
        const int ytop = heldY - fhh;
        const int ilo = max(0, -ytop);
        const unsigned span = (unsigned)(min(fh, H - ytop) - 1 - ilo);
        T* p = errT + ((long long)ytop * W + heldX) * Dp + zp * VEC;
#pragma unroll
        for (int i = 0; i < NE; ++i) {
            if ((unsigned)(i - ilo) <= span) *(V*)p = ecache[i];
            p += rstride;
        }
    }
    return gbase;
}

// ---- B: the serial scalar chain ---------------------------------------------------------------
PIPE_TMPL PIPE_ROLE_FN unsigned pipe_role_B(PIPE_ROLE_PARAMS) {
    PIPE_ROLE_DECLS;
    long long accepted = pb.accepted[chain];
    PIPE_SWEEPS_BEGIN
        // =================================================================================
        // the serial scalar chain.  Everything that does not depend on the previous decisions
        // arrives precomputed in the record of warp X; the loads below are independent.
        double ha[L], hr[L]; bool hacc[L];           // decisions of the last L sites
#pragma unroll
        for (int d = 0; d < L; ++d) { ha[d] = 0.0; hr[d] = 0.0; hacc[d] = false; }
        const double lo_a = pb.pmin[cube * 3], hi_a = pb.pmax[cube * 3];
        // rows of the chain / likelihood arrays this sweep is saved to (lib/run.py:353); lane 0
        // writes them, and the parameter map, right behind each decision
        double* crow = nullptr; double* lrow = nullptr;
        if ((it % keep) == 0) {
            const long long r = it / keep - row_first;
            if (chain_out) crow = chain_out + ((size_t)chain * rows_local + r) * HW * 3;
            if (lik_out) lrow = lik_out + ((size_t)chain * rows_local + r) * HW;
        }
        for (int j = 0; j < ns; ++j) {
            if (*abort_flag) break;
            if (lane == 0) s_prog[warp] = j;
            const unsigned gj = gbase + (unsigned)j;        // running site count: ring stage + phase
            const int st = (int)(gj & (R - 1));
            const unsigned ph = (gj / R) & 1u;
            MWAIT(s_bar + PB_SCAL * R + st, ph, 3);
            const double* rec = s_rec + st * PREC_N;
            const double* spec_s = s_spec + st * 16;
            const double K0 = rec[PREC_K0], a = rec[PREC_A];
            const double aQOO = rec[PREC_AQOO], aQON = rec[PREC_AQON];
            const double sgO = rec[PREC_SGO], sgN = rec[PREC_SGN];
            const double isgO = rec[PREC_ISGO], isgN = rec[PREC_ISGN];
            const double sg2O = rec[PREC_SG2O], sg2N = rec[PREC_SG2N];
            double xs[4 * L];
#pragma unroll
            for (int q = 0; q < 4 * L; ++q) xs[q] = rec[PREC_X + q];
            const double log_u = spec_s[SP_LOGU];
            const double n1 = spec_s[SP_N1], n2 = spec_s[SP_N2], z1 = spec_s[SP_Z1], e1 = spec_s[SP_E1];
            const double u4 = spec_s[SP_U4], u5 = spec_s[SP_U5];
            const double a_new = s_prop[st * 8 + 3];
            const bool oob = s_prop[st * 8 + 7] != 0.0;
            MWAIT(s_bar + PB_PART * R + st, ph, 4);
#ifdef D3D_PIPE_RROLE
            const double t0 = s_red[st * 32], t1 = s_red[st * 32 + 1];  // the two sums over h (warp R)
#else
            double t0, t1;
            {
                // the two sums over h: one 16-byte load per window warp (unused slots are
                // zero), added in a fixed tree
                const double2* rd = (const double2*)(s_red + st * 32);
                double2 p[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) p[q] = rd[q];
#pragma unroll
                for (int w2 = 8; w2 > 0; w2 >>= 1)
#pragma unroll
                    for (int q = 0; q < w2; ++q) { p[q].x += p[q + w2].x; p[q].y += p[q + w2].y; }
                t0 = p[0].x; t1 = p[0].y;
            }
#endif
            // corrections for the updates the window warps had not applied yet when they
            // formed the sums of this site (the cross terms are zero outside the run)
            double cB = 0.0, cPO = 0.0;
#pragma unroll
            for (int d = L; d >= 1; --d) {
                const double a_i = ha[d - 1], r_i = hr[d - 1];
                const bool ac = hacc[d - 1];
                cB += a_i * xs[4 * (d - 1)] - r_i * (ac ? xs[4 * (d - 1) + 1] : xs[4 * (d - 1)]);
                cPO += a_i * xs[4 * (d - 1) + 2] - r_i * (ac ? xs[4 * (d - 1) + 3] : xs[4 * (d - 1) + 2]);
            }
            const double tB = t0 + cB, tPO = t1 + cPO;
            // accept test (lib/run.py:426-451)
            double delta;
            if (a_new != a) {                                // amplitude jumps too (not the reference's default)
                const double* raw = rec + PREC_RAW;
                const double Pn = tPO - tB;
                const double Bq = a * tPO - a_new * Pn;
                const double Cq = a * a * raw[1] - 2.0 * a * a_new * raw[2] + a_new * a_new * raw[3];
                delta = -Bq - 0.5 * Cq;
            } else {
                delta = fma(-a, tB, K0);                     // -(a B) - a^2 C / 2
            }
            const int acc = (log_u < delta) && !oob;                        // :438
            const double S1 = acc ? (tPO - tB) + aQON : tPO + aQOO;
            // Gibbs draw (lib/run.py:491-496): ro = 1/q, q = 1/ra + S2, sigma = rsqrt(q): formed
            // ahead for both outcomes
            const double sg = acc ? sgN : sgO, isg = acc ? isgN : isgO, sg2 = acc ? sg2N : sg2O;
            const double mu = S1 * sg2;
            const double as = (lo_a - mu) * isg, bs = (hi_a - mu) * isg;   // lib/rtnorm.py:74-76
            double rs_ = 0.0;
            bool done = false;
            const bool plain = (as < bs) && !(fabs(as) > fabs(bs));   // no mirror (:108)
            if (plain && as < -2.00443204036) {                        // :127-131
                if (n1 >= as && n1 <= bs) { rs_ = n1; done = true; }
                else if (n2 >= as && n2 <= bs) { rs_ = n2; done = true; }
            } else if (plain && as > 3.48672170399 && -as * (bs - as) < -40.0) {   // :112-124
                if (2.0 * as * as * e1 > z1 * z1) { rs_ = as - z1 / as; done = true; }
            } else if (plain && as >= -2.00443204036 && as <= 3.48672170399) {     // :133-222
                const int N = 4000;
                const int ka = s_tnc[3271 + (int)floor(as * 1631.73284006)];
                const int kb = bs >= 3.48672170399 ? N : s_tnc[3271 + (int)floor(bs * 1631.73284006)];
                if (kb - ka >= 5) {
                    const int k = ka + (int)floor(u4 * (double)(kb + 1 - ka));
                    if (k != N && !(k <= ka + 2 || (k >= kb && bs < 3.48672170399))) {
                        const double yuk = s_tyu[k], xk = s_tx[k], dk = s_tx[k + 1] - xk;
                        const double ylk = k == 1 ? 0.053513975472 : (k <= 1954 ? s_tyu[k - 1] : s_tyu[k + 1]);
                        if (yuk * u5 < ylk) { rs_ = xk + u5 * dk * yuk / ylk; done = true; }
                    }
                }
            }
            if (!done)                                       // every other branch / retry (cold, out of line)
                rs_ = pipe_rtnorm_fallback(pb, as, bs, spec_s, (unsigned)chain, (unsigned)it, (unsigned)sites[j]);
            const double r = rs_ * sg + mu;                                 // :82-83
            if (lane == 0) {
                double* dc = s_dec + st * 4;
                // (the proposal record is read before DEC is raised: the stage may be recycled
                // as soon as the window warps have applied this decision)
                const double* pr = s_prop + st * 8;
                const double c_end = acc ? pr[4] : pr[1], w_end = acc ? pr[5] : pr[2];
                dc[0] = acc ? 1.0 : 0.0; dc[1] = r; dc[2] = a; dc[3] = delta;
                mbar_arrive(s_bar + PB_DEC * R + st);
                // the outcome goes to the parameter map and the saved rows
                // (lib/run.py:430-432, 448, 499, 516)
                const int site_j = sites[j];
                double* prm = pb.params + ((size_t)chain * HW + site_j) * 3;
                prm[0] = r; prm[1] = c_end; prm[2] = w_end;
                if (crow) { double* cr = crow + (size_t)site_j * 3; cr[0] = r; cr[1] = c_end; cr[2] = w_end; }
                if (lrow) lrow[site_j] = delta;
            }
            accepted += acc;
#pragma unroll
            for (int d = L - 1; d > 0; --d) { ha[d] = ha[d - 1]; hr[d] = hr[d - 1]; hacc[d] = hacc[d - 1]; }
            ha[0] = a; hr[0] = r; hacc[0] = acc != 0;
        }
        if (lane == 0) s_bc[2] = (double)accepted;
    PIPE_SWEEPS_END;
    if (lane == 0) {
        pb.accepted[chain] = accepted;
        pb.rate[chain] = rate;
        pb.iters[chain] = it;
        if (!alive) pb.active[chain] = 0;
        if (*abort_flag) {
            atomicExch(pb.status, 0x40000000 | *abort_flag);
            if (pb.dbg) for (int q = 0; q < 32; ++q) pb.dbg[q] = s_prog[q];
        }
    }
    return gbase;
}

// ---- R: reducer of the raw window sums
PIPE_TMPL PIPE_ROLE_FN unsigned pipe_role_R(PIPE_ROLE_PARAMS) {
    PIPE_ROLE_DECLS;
    PIPE_SWEEPS_BEGIN
        // =================================================================================
        // reducer: raw sums h0[z] of every window thread -> sum over the column groups (lane
        // = z-vector) -> the two sums over z with the profiles of the site -> PART
        const int zl = lane < ZL ? lane : 0;
        for (int j = 0; j < ns; ++j) {
            if (*abort_flag) break;
            if (lane == 0) s_prog[warp] = j;
            const unsigned gj = gbase + (unsigned)j;        // running site count: ring stage + phase
            const int st = (int)(gj & (R - 1));
            const unsigned ph = (gj / R) & 1u;
            MWAIT(s_bar + PB_HSUM * R + st, ph, 1);
            const double* hs = s_var + pv.hs + ((int)(gj & (PIPE_HS - 1)) * nwt + zl) * VEC;
            double acc[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[v] = 0.0;
#pragma unroll 4
            for (int g = 0; g < NG; ++g) {
#pragma unroll
                for (int v = 0; v < VEC; v += 2) {
                    const double2 t = *(const double2*)(hs + g * ZL * VEC + v);
                    acc[v] += t.x; acc[v + 1] += t.y;
                }
            }
            MWAIT(s_bar + PB_PROF * R + st, ph, 8);    // profiles of site j
            double pB = 0.0, pPO = 0.0;
            if (lane < ZL) {
                const double* Lo = s_var + pv.Lu_o + st * Dp + zl * VEC;
                const double* Ln = s_var + pv.Lu_n + st * Dp + zl * VEC;
#pragma unroll
                for (int v = 0; v < VEC; v += 2) {
                    const double2 lo = *(const double2*)(Lo + v), ln = *(const double2*)(Ln + v);
                    pB = fma(lo.x - ln.x, acc[v], pB);
                    pPO = fma(lo.x, acc[v], pPO);
                    pB = fma(lo.y - ln.y, acc[v + 1], pB);
                    pPO = fma(lo.y, acc[v + 1], pPO);
                }
            }
            const double tot = warp_sum2(pB, pPO, lane);
            if ((lane & 15) == 0) s_red[st * 32 + (lane >> 4)] = tot;
            __syncwarp();
            if (lane == 0) mbar_arrive(s_bar + PB_PART * R + st);
        }
    PIPE_SWEEPS_END;
    return gbase;
}

// ---- X: static-table sums and cross terms -> the record warp B reads
PIPE_TMPL PIPE_ROLE_FN unsigned pipe_role_X(PIPE_ROLE_PARAMS) {
    PIPE_ROLE_DECLS;
    PIPE_SWEEPS_BEGIN
        // =================================================================================
        // per site: the quadratic sums over the static G table, the 4 L cross terms with the
        // previous sites of the run, and from them the record warp B reads
        const double ira = 1.0 / pb.prior_var[cube];
        // lane l owns channels l and l + 32 (Dp <= 64); channels beyond Dp read channel 0 and
        // are weighted by zero
        const int zq0 = lane < Dp ? lane : 0, zq1 = lane + 32 < Dp ? lane + 32 : 0;
        const double m0 = lane < Dp ? 1.0 : 0.0, m1 = lane + 32 < Dp ? 1.0 : 0.0;
        const double* const gt0 = pb.gtab + (size_t)cube * HW * Dp;
        const double* const xt0 = pb.xtab + (size_t)cube * pb.xtab_L * pb.max_sites * Dp;
        const int xstride = pb.max_sites * Dp;
        for (int j = warp - wX0; j < ns; j += NX) {
            if (*abort_flag) break;
            if (lane == 0) s_prog[warp] = j;
            const int site = sites[j];
            const int rs = run_start[j];
            const unsigned gj = gbase + (unsigned)j;
            const int st = (int)(gj & (R - 1));
            // static tables of this site (the loads fly while the profiles are produced)
            double gq[2], xq[L][2];
            gq[0] = gt0[site * Dp + zq0] * m0; gq[1] = gt0[site * Dp + zq1] * m1;
#pragma unroll
            for (int d = 1; d <= L; ++d) {
                xq[d - 1][0] = xt0[(d - 1) * xstride + j * Dp + zq0];
                xq[d - 1][1] = xt0[(d - 1) * xstride + j * Dp + zq1];
            }
            // every profile this job may touch must be complete before it arrives on FREE
            // below (keeps the arrivals of several X warps inside the right barrier phase)
#pragma unroll
            for (int d = L; d >= 0; --d)
                if (gj >= (unsigned)d) MWAIT(s_bar + PB_PROF * R + (int)((gj - d) & (R - 1)), ((gj - d) / R) & 1u, 5);
            const double* Lo = s_var + pv.Lu_o + st * Dp;
            const double* Ln = s_var + pv.Lu_n + st * Dp;
            double s4[4], sx[4 * L];
#pragma unroll
            for (int q = 0; q < 4; ++q) s4[q] = 0.0;
#pragma unroll
            for (int q = 0; q < 4 * L; ++q) sx[q] = 0.0;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int z = q ? zq1 : zq0;
                const double lo = Lo[z], ln = Ln[z], dl = lo - ln;
                const double G = gq[q];                          // zero beyond Dp
                s4[0] = fma(dl * dl, G, s4[0]);
                s4[1] = fma(lo * lo, G, s4[1]);
                s4[2] = fma(lo * ln, G, s4[2]);
                s4[3] = fma(ln * ln, G, s4[3]);
#pragma unroll
                for (int d = 1; d <= L; ++d) {
                    if (j - d < rs) continue;                    // (warp-uniform) outside the run: zero
                    const int sti = (int)((gj - d) & (R - 1));
                    const double X = xq[d - 1][q] * (q ? m1 : m0);
                    const double uo = s_var[pv.Lu_o + sti * Dp + z] * X, un = s_var[pv.Lu_n + sti * Dp + z] * X;
                    sx[4 * (d - 1) + 0] = fma(dl, uo, sx[4 * (d - 1) + 0]);
                    sx[4 * (d - 1) + 1] = fma(dl, un, sx[4 * (d - 1) + 1]);
                    sx[4 * (d - 1) + 2] = fma(lo, uo, sx[4 * (d - 1) + 2]);
                    sx[4 * (d - 1) + 3] = fma(lo, un, sx[4 * (d - 1) + 3]);
                }
            }
            double* rec = s_rec + st * PREC_N;
            {
                // C, QOO, QON, QNN end in lanes 0, 16, 8, 24: each of them derives what warp B
                // needs from ITS sum (same code on the four lanes, different slots)
                const double qv = warp_sum4(s4, lane);
                const double a = s_prop[st * 8];
                const double qq = ira + qv;
                const double sg = rsqrt(qq);
                if ((lane & 7) == 0) {
                    const int slot = ((lane >> 3) & 1) * 2 + ((lane >> 4) & 1);   // 0 C, 1 QOO, 2 QON, 3 QNN
                    rec[PREC_RAW + slot] = qv;
                    if (slot == 0) { rec[PREC_K0] = -0.5 * (a * a) * qv; rec[PREC_A] = a; }
                    if (slot == 2) rec[PREC_AQON] = a * qv;
                    if (slot == 1) { rec[PREC_AQOO] = a * qv; rec[PREC_SGO] = sg; rec[PREC_ISGO] = qq * sg; rec[PREC_SG2O] = sg * sg; }
                    if (slot == 3) { rec[PREC_SGN] = sg; rec[PREC_ISGN] = qq * sg; rec[PREC_SG2N] = sg * sg; }
                }
            }
            if (L == 1) {
                const double xv = warp_sum4(sx, lane);
                if ((lane & 7) == 0) rec[PREC_X + ((lane >> 3) & 1) * 2 + ((lane >> 4) & 1)] = xv;
            } else {
                const double xv = warp_sum8(sx, lane);
                if ((lane & 3) == 0) rec[PREC_X + warp_sum8_slot(lane)] = xv;
                if (L == 3) {
                    const double xw = warp_sum4(sx + 8, lane);
                    if ((lane & 7) == 0) rec[PREC_X + 8 + ((lane >> 3) & 1) * 2 + ((lane >> 4) & 1)] = xw;
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(s_bar + PB_SCAL * R + st);
#pragma unroll
                for (int d = 0; d <= L; ++d)
                    if (gj >= (unsigned)d) mbar_arrive(s_bar + PB_FREE * R + (int)((gj - d) & (R - 1)));
            }
        }
    PIPE_SWEEPS_END;
    return gbase;
}

// ---- A / P: per-site producers (PMODE = 0): site k -> warp k mod NA (accept uniform, first
// truncated-normal tries, old profile) and warp k mod NP (proposal, new profile).  Their code runs
// every site and stays warm in the instruction caches: the fastest producers while the chip is
// partly filled (0.62 M evals/s for one chain), four warps' worth of instruction fetch when it is not.
PIPE_TMPL PIPE_ROLE_FN unsigned pipe_role_AP(PIPE_ROLE_PARAMS) {
    PIPE_ROLE_DECLS;
    PIPE_SWEEPS_BEGIN
        // =================================================================================
        const int pw = roleA ? warp - wA0 : NA + (warp - wP0);        // producer index
        double* const G = s_var + pv.G + pw * PIPE_B * pv.GN;
        const int kd_n = pb.kd_n, kd_mhi = pb.kd_mhi, P = pb.P, D = pb.D;
        for (int j = roleA ? warp - wA0 : warp - wP0; j < ns; j += (roleA ? NA : NP)) {
            if (*abort_flag) break;
            if (lane == 0) s_prog[warp] = j;
            const int site = sites[j];
            const unsigned gj = gbase + (unsigned)j;        // running site count: ring stage + phase
            const int st = (int)(gj & (R - 1));
            const unsigned ph = (gj / R) & 1u;
            const double* prm = pb.params + ((size_t)chain * HW + site) * 3;
            const double a = prm[0], c_old = prm[1], w_old = prm[2];
            // Philox blocks 0..3 in lanes 0..3: draws (2b, 2b+1) of this site
            unsigned o[4];
            philox_block_rolled((unsigned)pb.seed, (unsigned)(pb.seed >> 32), (unsigned)(lane & 3),
                                (unsigned)site, (unsigned)it, pb.first_chain + (unsigned)chain, o);
            const double ua = Philox::u53(o[0], o[1]), ub = Philox::u53(o[2], o[3]);
            double c_prof, w_prof;
            double* Lu_out;
            if (gj >= (unsigned)R) MWAIT(s_bar + PB_FREE * R + st, ph ^ 1u, 7);   // stage recycled
            if (roleP) {
                // ---- proposal: Cauchy jump (lib/run.py:570-579), one tan() for all lanes
                const double u0 = __shfl_sync(0xffffffffu, ua, 0);
                const double u1 = __shfl_sync(0xffffffffu, ub, 0);
                const double u2 = __shfl_sync(0xffffffffu, ua, 1);
                const double q4 = 1.5707963267948966;
                const double targ = lane == 0 ? u1 : (lane == 1 ? u2 : u0);
                const double tv = tan(-q4 + (q4 - (-q4)) * targ);
                const double t1 = __shfl_sync(0xffffffffu, tv, 0);
                const double t2 = __shfl_sync(0xffffffffu, tv, 1);
                const double t0 = __shfl_sync(0xffffffffu, tv, 2);
                const double a_new = pb.jump[0] != 0.0 ? a + pb.jump[0] * t0 : a;
                const double c_new = c_old + pb.jump[1] * t1;
                const double w_new = w_old + pb.jump[2] * t2;
                const double* lo = pb.pmin + cube * 3;
                const double* hi = pb.pmax + cube * 3;
                const int oob = (a_new < lo[0]) | (c_new < lo[1]) | (w_new < lo[2]) |
                                (a_new > hi[0]) | (c_new > hi[1]) | (w_new > hi[2]);
                if (lane == 0) {
                    double* prop_s = s_prop + st * 8;
                    prop_s[0] = a; prop_s[1] = c_old; prop_s[2] = w_old;
                    prop_s[3] = a_new; prop_s[4] = c_new; prop_s[5] = w_new;
                    prop_s[7] = (double)oob;
                }
                c_prof = c_new; w_prof = w_new; Lu_out = s_var + pv.Lu_n + st * Dp;
            } else {
                // ---- accept uniform + first truncated-normal draws, lane-parallel:
                // five logs in one log(), two sqrt in one sqrt(), two cos in one cos()
                const double u3 = __shfl_sync(0xffffffffu, ub, 1);
                const double u4 = __shfl_sync(0xffffffffu, ua, 2);
                const double u5 = __shfl_sync(0xffffffffu, ub, 2);
                const double u6 = __shfl_sync(0xffffffffu, ua, 3);
                const double u7 = __shfl_sync(0xffffffffu, ub, 3);
                const double r4 = 1e-15 + (1.0 - 1e-15) * u4, r5 = 1e-15 + (1.0 - 1e-15) * u5;
                const double larg = lane == 0 ? u3 : lane == 1 ? r5 : lane == 2 ? 1.0 + r4 * -1.0
                                  : lane == 3 ? 1.0 - u4 : 1.0 - u6;
                const double lv = log(larg);
                const double l4 = __shfl_sync(0xffffffffu, lv, 3);
                const double l6 = __shfl_sync(0xffffffffu, lv, 4);
                const double sv = sqrt(-2.0 * (lane == 0 ? l4 : l6));
                const double cv = cos(6.283185307179586 * (lane == 0 ? u5 : u7));
                const double nv = sv * cv;
                const double n2 = __shfl_sync(0xffffffffu, nv, 1);
                const double e1 = -__shfl_sync(0xffffffffu, lv, 1);
                const double z1 = __shfl_sync(0xffffffffu, lv, 2);
                if (lane == 0) {
                    double* spec_s = s_spec + st * 16;
                    spec_s[SP_U4] = u4; spec_s[SP_U5] = u5; spec_s[SP_U6] = u6; spec_s[SP_U7] = u7;
                    spec_s[SP_E1] = e1; spec_s[SP_Z1] = z1; spec_s[SP_N1] = nv; spec_s[SP_N2] = n2;
                    spec_s[SP_LOGU] = lv;
                }
                c_prof = c_old; w_prof = w_old; Lu_out = s_var + pv.Lu_o + st * Dp;
            }
            // ---- unit line profile Lu = lsf (*) exp(-(z-c)^2 / (2 w^2)) -------------------
            // (lib/line_models.py:98-109, lib/convolution.py:89-160 in direct form).  Lane l
            // owns channels 2l and 2l+1.  The Gaussian goes to G[] periodically extended
            // (G[i] = g[(i - mhi) mod P]), which turns the circular kernel into a dense FIR
            // window: out[z] = sum_t Kd[t] G[z + t]  (Kd: host, d3d_set_problem).
            {
                const double inv2w2 = 1.0 / (2.0 * (w_prof * w_prof));
                const int z0 = 2 * lane;
                if (z0 < D) {
                    const double d0 = (double)z0 - c_prof, d1 = (double)(z0 + 1) - c_prof;
                    // (one out-of-line copy of exp: code footprint, see the note on the I-cache)
                    double g0 = d_exp(-1.0 * (d0 * d0) * inv2w2);
                    double g1 = z0 + 1 < D ? d_exp(-1.0 * (d1 * d1) * inv2w2) : 0.0;
                    if (pb.n_comp > 1) {                         // tied multiplet (d3d_set_line_model)
                        g0 += extra_components(pb, d0, inv2w2);
                        if (z0 + 1 < D) g1 += extra_components(pb, d1, inv2w2);
                    }
                    if (pb.has_lsf) {
                        const int GN = pv.GN;
#pragma unroll
                        for (int k = -1; k <= 2; ++k) {
                            const int i0 = z0 + kd_mhi + k * P;
                            if (i0 >= 0 && i0 < GN) G[i0] = g0;
                            if (i0 + 1 >= 0 && i0 + 1 < GN) G[i0 + 1] = g1;
                        }
                    } else {                                 // lib/run.py:675-676
                        Lu_out[z0] = g0; Lu_out[z0 + 1] = g1;
                    }
                } else if (z0 < Dp && !pb.has_lsf) {
                    Lu_out[z0] = 0.0; Lu_out[z0 + 1] = 0.0;
                }
                __syncwarp();
                if (pb.has_lsf && z0 < Dp) {
                    const double2* gp = (const double2*)(G + z0);          // 16-byte aligned: z0 even
                    const double2* kp = (const double2*)(s_var + pv.Kd);
                    double o0a = 0.0, o0b = 0.0, o1a = 0.0, o1b = 0.0;
                    double2 gc = gp[0];
                    const int npair = (kd_n + 1) >> 1;                     // Kd is zero-padded
#pragma unroll 2
                    for (int q = 0; q < npair; ++q) {
                        const double2 kk = kp[q];
                        const double2 gn = gp[q + 1];
                        o0a = fma(kk.x, gc.x, o0a);
                        o0b = fma(kk.y, gc.y, o0b);
                        o1a = fma(kk.x, gc.y, o1a);
                        o1b = fma(kk.y, gn.x, o1b);
                        gc = gn;
                    }
                    Lu_out[z0] = z0 < D ? o0a + o0b : 0.0;
                    Lu_out[z0 + 1] = z0 + 1 < D ? o1a + o1b : 0.0;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(s_bar + PB_PROF * R + st);
        }
    PIPE_SWEEPS_END;
    return gbase;
}

// ---- PR: batched producers (PMODE = 1).  A warp prepares PIPE_B = 8 consecutive sites per
// pass, lane-parallel OVER THE SITES: one Philox call gives the 8 x 4 blocks of draws, one tan()
// the 24 Cauchy jumps, two log(), one sqrt(), one cos() the accept uniforms and the first
// truncated-normal tries, and the 16 line profiles (old and new) are evaluated four lanes per
// profile.  ~350 warp instructions per site instead of ~1 100 with one warp per site and role.
// Its code runs once per 8 sites and is instruction-cache cold every time (~39 k cycles per
// batch, profiles/r02_notes.md): the right trade when every SM runs a chain -- the per-site
// producers above, warm but four warps, when the chip is partly filled.  Values are exchanged
// through a 2 KB scratch of the warp.  Same draws and arithmetic per site as d3d_slide.cuh (the
// Gaussian through the table-driven exp of the spectral pass, <= 2 ulp).  A warp's next use of a
// stage lies one lap of the ring later (PIPE_B * warps <= PIPE_R), behind FREE.
PIPE_TMPL PIPE_ROLE_FN unsigned pipe_role_PR(PIPE_ROLE_PARAMS) {
    PIPE_ROLE_DECLS;
    const int NPW = NA + NP;
    static_assert(PIPE_B * (NA + NP) <= PIPE_R && PIPE_R % (PIPE_B * (NA + NP)) == 0,
                  "every producer warp must find its own previous batch in the stages it re-uses");
    const int pw = warp - wA0;
    double* const Gw = s_var + pv.G + pw * PIPE_B * pv.GN;
    double* const sc = s_var + pv.S + pw * 256;
    double* const sU = sc;            // [8][8]  uniforms 0..7 of every site of the batch
    double* const sT = sc + 64;       // [8][4]  tan of the three jump uniforms
    double* const sL = sc + 96;       // [8][8]  the five logs
    double* const sN = sc + 160;      // [8][2]  the two pre-evaluated normals
    double* const sP = sc + 176;      // [8][8]  a, c_old, w_old, a_new, c_new, w_new
    const double* const tab64 = s_var + pv.T64;
    const int kd_n = pb.kd_n, kd_mhi = pb.kd_mhi, P = pb.P, D = pb.D, GN = pv.GN;
    const int QC = 2 * ((Dp + 7) / 8);                  // channels per lane of a profile (even)
    const int nb = (ns + PIPE_B - 1) / PIPE_B;
    // The producer warps run the SAME code on consecutive batches AT THE SAME TIME: this code is
    // instruction-cache cold at every pass (26 KB, once per 8 sites), and the fetches beyond the
    // SM's own cache go to the L1.5 of the GPC, which all its SMs share and which saturates when
    // every SM runs a chain (ncu gcc__*: profiles/r02_notes.md).  In lock step the second warp
    // finds the lines the first one has just fetched.  GO[0] is their two-party barrier; its
    // phase count `go` carries over sweeps and work items in shared memory.
    unsigned go = (unsigned)s_bc[4];
    PIPE_SWEEPS_BEGIN
        for (int b0 = 0; b0 < nb; b0 += NPW) {
            if (*abort_flag) break;
            const int b = b0 + pw;
            const int j0 = b * PIPE_B;
            const int j = j0;                            // (named by the wait diagnostics)
            if (NPW > 1) {
                if (lane == 0) mbar_arrive(s_bar + PB_GO * R);
                MWAIT(s_bar + PB_GO * R, go & 1u, 8);
                ++go;
            }
            // OLD profiles come from the cache unless this is the first sweep after the parameters
            // were set from outside (then they are computed, and every site is written back)
            const bool cached = pb.lucache != nullptr && (pb.lu_valid || it > ka.it_first);
            if (b >= nb) {                               // (odd number of batches: still meet the partner below)
                if (NPW > 1 && lane == 0) mbar_arrive(s_bar + PB_GO * R + 1);
                continue;
            }
            if (lane == 0) s_prog[warp] = j0;
            // ---- lanes 0..7: one site each (stage, parameters) ------------------------------------
            const int js = j0 + (lane & 7);
            const bool mine = lane < PIPE_B && js < ns;
            const unsigned gjs = gbase + (unsigned)js;
            const int st_s = (int)(gjs & (R - 1));
            {
                // Philox: lane = (site of the batch) * 4 + block; draws (2 blk, 2 blk + 1)
                const int s4 = lane >> 2, blk = lane & 3;
                const int jq = j0 + s4;
                const int site_q = jq < ns ? sites[jq] : 0;
                unsigned o[4];
                philox_block_rolled((unsigned)pb.seed, (unsigned)(pb.seed >> 32), (unsigned)blk,
                                    (unsigned)site_q, (unsigned)it, pb.first_chain + (unsigned)chain, o);
                sU[s4 * 8 + 2 * blk] = Philox::u53(o[0], o[1]);
                sU[s4 * 8 + 2 * blk + 1] = Philox::u53(o[2], o[3]);
                if (mine) {
                    const double* prm = pb.params + ((size_t)chain * HW + sites[js]) * 3;
                    sP[lane * 8 + 0] = prm[0]; sP[lane * 8 + 1] = prm[1]; sP[lane * 8 + 2] = prm[2];
                }
            }
            __syncwarp();
            {
                // Cauchy jumps (lib/run.py:570-579): lane = site * 3 + parameter, one tan() for all
                const double q4 = 1.5707963267948966;
                const int s3 = lane / 3, k3 = lane - 3 * s3;
                const double u = lane < 3 * PIPE_B ? sU[s3 * 8 + k3] : 0.5;
                const double tv = tan(-q4 + (q4 - (-q4)) * u);
                if (lane < 3 * PIPE_B) sT[s3 * 4 + k3] = tv;
                // accept uniform + first truncated-normal draws: five logs per site, two passes of
                // four sites (lib/run.py:435; lib/rtnorm.py:121-122, 129)
#pragma unroll 1
                for (int pass = 0; pass < 2; ++pass) {
                    const int s5 = pass * 4 + lane / 5, k5 = lane - 5 * (lane / 5);
                    const bool on5 = lane < 20;
                    const double* uu = sU + (on5 ? s5 : 0) * 8;
                    const double u4 = uu[4];
                    const double r4 = 1e-15 + (1.0 - 1e-15) * u4, r5 = 1e-15 + (1.0 - 1e-15) * uu[5];
                    const double larg = k5 == 0 ? uu[3] : k5 == 1 ? r5 : k5 == 2 ? 1.0 + r4 * -1.0
                                      : k5 == 3 ? 1.0 - u4 : 1.0 - uu[6];
                    const double lv = log(larg);
                    if (on5) sL[s5 * 8 + k5] = lv;
                }
            }
            __syncwarp();
            {
                // the two normals of the Gaussian-proposal branch: lane = site * 2 + k
                const int s2 = (lane >> 1) & 7, k2 = lane & 1;
                const double sv = sqrt(-2.0 * sL[s2 * 8 + 3 + k2]);
                const double cv = cos(6.283185307179586 * sU[s2 * 8 + 5 + 2 * k2]);
                if (lane < 2 * PIPE_B) sN[s2 * 2 + k2] = sv * cv;
            }
            // Everything up to here lives in the warp's own scratch: only now must the stages of the
            // batch be free again (the draws of a pass overlap the consumers' last sites of the
            // half of the ring it refills).
            if (mine && gjs >= (unsigned)R) MWAIT(s_bar + PB_FREE * R + st_s, ((gjs / R) & 1u) ^ 1u, 7);
            __syncwarp();
            if (NPW > 1) {
                // the warps' stages come free at different times: meet again (GO[1]) so that the
                // rest of the pass -- most of its code -- runs in lock step too.  Costs nothing
                // while the producers are ahead: both batches are needed only a half ring later.
                if (lane == 0) mbar_arrive(s_bar + PB_GO * R + 1);
                MWAIT(s_bar + PB_GO * R + 1, (go - 1u) & 1u, 10);
            }
            if (pb.lucache) {
                // Profile cache: the site that held a stage one lap ago is done with it -- the unit
                // profile it ENDED with (new if accepted, old if not) is its "old" profile of the
                // next sweep (lib/run.py:402 recomputes it from the same (c, w): same bits).
                // lane = site of the batch * 4 + quarter of the channels.
                const int s4 = lane >> 2, qd = lane & 3;
                const int jq = j0 + s4, jp = jq - R;
                const bool accp = jq < ns && jp >= 0 && s_dec[(int)((gbase + (unsigned)jq) & (R - 1)) * 4] != 0.0;
                if (jq < ns && jp >= 0 && (accp || !cached)) {   // (a rejected visit leaves the cached profile as it is)
                    const int st_q = (int)((gbase + (unsigned)jq) & (R - 1));
                    const int zlo = qd * QC, zhi = min(zlo + QC, Dp);
                    const double2* src = (const double2*)(s_var + (accp ? pv.Lu_n : pv.Lu_o) + st_q * Dp + zlo);
                    double2* dst = (double2*)(pb.lucache + ((size_t)chain * HW + sites[jp]) * Dp + zlo);
#pragma unroll 1
                    for (int q = 0; 2 * q < zhi - zlo; ++q) dst[q] = src[q];
                }
                __syncwarp();
            }
            if (mine) {
                // proposal, bounds (lib/run.py:374-388) and the records of this site's stage
                const double* pp = sP + lane * 8;
                const double a = pp[0], c_old = pp[1], w_old = pp[2];
                const double t0 = sT[lane * 4 + 0], t1 = sT[lane * 4 + 1], t2 = sT[lane * 4 + 2];
                const double a_new = pb.jump[0] != 0.0 ? a + pb.jump[0] * t0 : a;
                const double c_new = c_old + pb.jump[1] * t1;
                const double w_new = w_old + pb.jump[2] * t2;
                const double* lo = pb.pmin + cube * 3;
                const double* hi = pb.pmax + cube * 3;
                const int oob = (a_new < lo[0]) | (c_new < lo[1]) | (w_new < lo[2]) |
                                (a_new > hi[0]) | (c_new > hi[1]) | (w_new > hi[2]);
                sP[lane * 8 + 3] = a_new; sP[lane * 8 + 4] = c_new; sP[lane * 8 + 5] = w_new;
                double* prop_s = s_prop + st_s * 8;
                prop_s[0] = a; prop_s[1] = c_old; prop_s[2] = w_old;
                prop_s[3] = a_new; prop_s[4] = c_new; prop_s[5] = w_new;
                prop_s[7] = (double)oob;
                const double* uu = sU + lane * 8;
                const double* ll = sL + lane * 8;
                double* spec_s = s_spec + st_s * 16;
                spec_s[SP_U4] = uu[4]; spec_s[SP_U5] = uu[5]; spec_s[SP_U6] = uu[6]; spec_s[SP_U7] = uu[7];
                spec_s[SP_E1] = -ll[1]; spec_s[SP_Z1] = ll[2];
                spec_s[SP_N1] = sN[lane * 2]; spec_s[SP_N2] = sN[lane * 2 + 1];
                spec_s[SP_LOGU] = ll[0];
            }
            __syncwarp();
            // ---- unit line profiles Lu = lsf (*) exp(-(z-c)^2 / (2 w^2)), old and new, of the 8
            // sites: lane = site * 4 + quarter of the channels (lib/line_models.py:98-109,
            // lib/convolution.py:89-160 in direct form).  The Gaussian goes to its G buffer
            // periodically extended (G[i] = g[(i - mhi) mod P]), which turns the circular kernel
            // into a dense FIR window: out[z] = sum_t Kd[t] G[z + t]  (Kd: host, d3d_set_problem).
            {
                const int s4 = lane >> 2, qd = lane & 3;
                const int jq = j0 + s4;
                const bool onq = jq < ns;
                const int st_q = (int)((gbase + (unsigned)jq) & (R - 1));
                const int zlo = qd * QC, zhi = min(zlo + QC, Dp);
                // OLD profile: kept from the site's last visit
                if (cached && onq) {
                    const double2* src = (const double2*)(pb.lucache + ((size_t)chain * HW + sites[jq]) * Dp + zlo);
                    double2* dst = (double2*)(s_var + pv.Lu_o + st_q * Dp + zlo);
#pragma unroll 1
                    for (int q = 0; 2 * q < zhi - zlo; ++q) dst[q] = src[q];
                }
#pragma unroll 1
                for (int which = cached ? 1 : 0; which < 2; ++which) {
                    const double c_prof = sP[s4 * 8 + (which ? 4 : 1)], w_prof = sP[s4 * 8 + (which ? 5 : 2)];
                    const double inv2w2 = 1.0 / (2.0 * (w_prof * w_prof));
                    double* const G = Gw + s4 * GN;
                    double* const Lu_out = s_var + (which ? pv.Lu_n : pv.Lu_o) + st_q * Dp;
                    if (onq) {
#pragma unroll 1
                        for (int z = zlo; z < zhi; ++z) {
                            const double d0 = (double)z - c_prof;
                            double gv = 0.0;
                            if (z < D) {
                                gv = exp_neg_tab(-1.0 * (d0 * d0) * inv2w2, tab64);
                                if (pb.n_comp > 1) gv += extra_components(pb, d0, inv2w2);   // tied multiplet
                            }
                            if (pb.has_lsf) {
                                if (z < D) {
#pragma unroll
                                    for (int k = -1; k <= 2; ++k) {
                                        const int i0 = z + kd_mhi + k * P;
                                        if (i0 >= 0 && i0 < GN) G[i0] = gv;
                                    }
                                }
                            } else {
                                Lu_out[z] = gv;                      // lib/run.py:675-676
                            }
                        }
                    }
                    __syncwarp();
                    if (onq && pb.has_lsf) {
                        const double2* kp = (const double2*)(s_var + pv.Kd);
                        const int npair = (kd_n + 1) >> 1;           // Kd is zero-padded
#pragma unroll 1
                        for (int z = zlo; z < zhi; z += 2) {
                            const double2* gp = (const double2*)(G + z);   // 16-byte aligned: z even
                            double o0a = 0.0, o0b = 0.0, o1a = 0.0, o1b = 0.0;
                            double2 gc = gp[0];
#pragma unroll 2
                            for (int q = 0; q < npair; ++q) {
                                const double2 kk = kp[q];
                                const double2 gn = gp[q + 1];
                                o0a = fma(kk.x, gc.x, o0a);
                                o0b = fma(kk.y, gc.y, o0b);
                                o1a = fma(kk.x, gc.y, o1a);
                                o1b = fma(kk.y, gn.x, o1b);
                                gc = gn;
                            }
                            Lu_out[z] = z < D ? o0a + o0b : 0.0;
                            Lu_out[z + 1] = z + 1 < D ? o1a + o1b : 0.0;
                        }
                    }
                    __syncwarp();                                    // G is re-used by the new profile
                }
            }
            if (mine) mbar_arrive(s_bar + PB_PROF * R + st_s);
        }
        if (pb.lucache && !*abort_flag) {
            // end of the sweep: the sites still in the ring go to the profile cache as they finish
            const bool cached_sweep = pb.lu_valid || it > ka.it_first;
            for (int t = pw; t < R / PIPE_B; t += NPW) {
                const int s4 = lane >> 2, qd = lane & 3;
                const int jf = ns - R + t * PIPE_B + s4;
                const int j = jf;                            // (named by the wait diagnostics)
                if (jf >= 0) {
                    const unsigned gjf = gbase + (unsigned)jf;
                    const int st_q = (int)(gjf & (R - 1));
                    // (decided is enough: only the producers write profiles, and FREE of the last
                    // L sites completes in the next sweep)
                    MWAIT(s_bar + PB_DEC * R + st_q, (gjf / R) & 1u, 9);
                    const int zlo = qd * QC, zhi = min(zlo + QC, Dp);
                    const bool accp = s_dec[st_q * 4] != 0.0;
                    const double2* src = (const double2*)(s_var + (accp ? pv.Lu_n : pv.Lu_o) + st_q * Dp + zlo);
                    double2* dst = (double2*)(pb.lucache + ((size_t)chain * HW + sites[jf]) * Dp + zlo);
                    if (accp || !cached_sweep)
#pragma unroll 1
                        for (int q = 0; 2 * q < zhi - zlo; ++q) dst[q] = src[q];
                }
            }
        }
    PIPE_SWEEPS_END;
    if (pw == 0 && lane == 0) s_bc[4] = (double)go;      // (items are separated by CTA barriers)
    return gbase;
}

template <typename T, bool IVCUBE, int NE, bool SQ, int L, int NA, int NP, int NX, int PMODE, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
sweep_seq_pipe_kernel(const __grid_constant__ Problem pb, long long it0_all, long long it1_all,
                      int keep, double min_rate, double* chain_out, double* lik_out,
                      long long row_first, long long rows_local, const int4* items,
                      const int* item_count, int max_items, volatile long long* progress) {
    typedef PipeFix<L> FX;
    const int VEC = Vec<T>::N;
    const int R = PIPE_R;
    static_assert(L >= 1 && L <= 3 && L + 2 < PIPE_R && L + 1 < PIPE_HS, "look-ahead must fit the rings and the record");
    extern __shared__ double smem_raw[];
    // SQ: square FSF of the template's size -- the window geometry is a compile-time constant
    const int fw = SQ ? NE : pb.fw, fh = SQ ? NE : pb.fh;
    const int Dp = pb.Dp;
    const int ZL = Dp / VEC;
    const int NG = fw + L + 1;                     // resident columns: L trailing + window + next
    const int FHP = NE + 1;
    const PipeVar pv = pipe_var_layout(fw, NE, pb.kd_n, Dp, NA + NP, NG, ZL, IVCUBE);
    double* const s_var = smem_raw + FX::VAR;
    double* const s_bc = smem_raw + FX::BC;
    unsigned long long* const s_bar = (unsigned long long*)(smem_raw + FX::BAR);
    volatile int* const abort_flag = (volatile int*)(smem_raw + FX::ABORT);
    PipeArgs* const s_args = (PipeArgs*)(smem_raw + FX::ARGS);
#ifdef D3D_PIPE_PROF
    const long long t_kernel0 = clock64();
#endif
    const int tid = threadIdx.x, warp = tid >> 5;
    const int nwt = NG * ZL, nww = (nwt + 31) >> 5;
    const int wA0 = nww, wX0 = nww + NA + NP, wR = wX0 + NX, wB = wR + PIPE_NR;

    // ---- constants --------------------------------------------------------------------------
    for (int q = tid; q < fw * FHP; q += blockDim.x) {
        const int dx = q / FHP, i = q - dx * FHP;
        s_var[pv.Ft + q] = i < fh ? pb.fsf[i * fw + dx] : 0.0;
    }
    for (int q = tid; q < pv.G - pv.Kd; q += blockDim.x) s_var[pv.Kd + q] = q < pb.kd_n ? pb.kdense[q] : 0.0;
    for (int q = tid; q < (NA + NP) * PIPE_B * pv.GN; q += blockDim.x) s_var[pv.G + q] = 0.0;
    for (int q = tid; q < (NA + NP) * 256; q += blockDim.x) s_var[pv.S + q] = 0.0;
    for (int q = tid; q < 64; q += blockDim.x) s_var[pv.T64 + q] = exp2((double)q * (1.0 / 64.0));
    for (int q = tid; q < 4002; q += blockDim.x) smem_raw[FX::TX + q] = pb.rt.x[q];
    for (int q = tid; q < 4001; q += blockDim.x) smem_raw[FX::TYU + q] = pb.rt.yu[q];
    for (int q = tid; q < 8961; q += blockDim.x)
        ((unsigned short*)(smem_raw + FX::TNC))[q] = (unsigned short)pb.rt.ncell[q];
    for (int q = tid; q < PIPE_R * 32; q += blockDim.x) smem_raw[FX::RED + q] = 0.0;
    if (tid == 0) {
        *abort_flag = 0;
        s_bc[4] = 0.0;
        s_args->keep = keep; s_args->min_rate = min_rate; s_args->chain_out = chain_out;
        s_args->lik_out = lik_out; s_args->row_first = row_first; s_args->rows_local = rows_local;
        s_args->it_first = it0_all;
        // The barriers are initialised ONCE: stage and phase parity follow a running count of the
        // sites this CTA has worked (gbase + j), which carries on across sweeps and work items, so
        // every phase is completed by the sites that follow (no re-initialisation in flight).
        for (int q = 0; q < PB_N * R; ++q) {
            const int kind = q / R;
#ifdef D3D_PIPE_RROLE
            const int n_part = 1;
#else
            const int n_part = nww;
#endif
            mbar_init(s_bar + q, kind == PB_PROF ? (PMODE ? 1 : 2) : kind == PB_HSUM ? nww : kind == PB_PART ? n_part
                                 : kind == PB_FREE ? nww + (L + 1) : kind == PB_GO ? NA + NP : 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    unsigned gbase = 0;                            // sites worked by this CTA before the current sweep
    __syncthreads();

    const int n_items = items ? item_count[blockIdx.x] : 1;
    for (int item = 0; item < n_items; ++item) {
        int chain; long long it0, it1;
        if (items) {
            const int4 w = items[(size_t)blockIdx.x * max_items + item];
            chain = w.x; it0 = it0_all + w.y; it1 = it0_all + w.z;
            if (tid == 0) while (progress[chain] < it0) __nanosleep(200);
            __syncthreads();
            __threadfence();
        } else {
            chain = blockIdx.x; it0 = it0_all; it1 = it1_all;
            if (chain >= pb.n_chains) return;
        }
        if (!pb.active[chain]) {                       // chain stopped earlier (acceptance rate)
            __syncthreads();
            if (items && tid == 0) progress[chain] = it1;
            continue;
        }
        __syncthreads();                               // previous item fully retired
        if (tid == 0) s_bc[2] = (double)pb.accepted[chain];
        __syncthreads();
        // every role runs all sweeps of the item and returns the advanced site count
        if (warp < nww)       gbase = pipe_role_W<T, IVCUBE, NE, SQ, L, NA, NP, NX, PMODE>(pb, *s_args, chain, it0, it1, gbase);
        else if (warp == wB)  gbase = pipe_role_B<T, IVCUBE, NE, SQ, L, NA, NP, NX, PMODE>(pb, *s_args, chain, it0, it1, gbase);
        else if (PIPE_NR && warp == wR) gbase = pipe_role_R<T, IVCUBE, NE, SQ, L, NA, NP, NX, PMODE>(pb, *s_args, chain, it0, it1, gbase);
        else if (warp >= wX0) gbase = pipe_role_X<T, IVCUBE, NE, SQ, L, NA, NP, NX, PMODE>(pb, *s_args, chain, it0, it1, gbase);
        else if (warp >= wA0) {
            if constexpr (PMODE != 0) gbase = pipe_role_PR<T, IVCUBE, NE, SQ, L, NA, NP, NX, PMODE>(pb, *s_args, chain, it0, it1, gbase);
            else                      gbase = pipe_role_AP<T, IVCUBE, NE, SQ, L, NA, NP, NX, PMODE>(pb, *s_args, chain, it0, it1, gbase);
        }
        if (items) {                                   // hand the chain over to its next owner
            __threadfence();
            __syncthreads();
            if (tid == 0) progress[chain] = it1;
        }
    }
#ifdef D3D_PIPE_PROF
    if (tid == 0 && blockIdx.x < 1024) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_pipe_cta[blockIdx.x * 2] = clock64() - t_kernel0;
        g_pipe_cta[blockIdx.x * 2 + 1] = smid;
    }
#endif
}

}  // namespace d3d
