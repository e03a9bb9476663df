// d3d_slide.cuh -- SEQ_EXACT sweep with a register-resident sliding window.
//
// In the reference's row-major sweep (lib/run.py:553-566) consecutive proposals
// move the FSF window by one spaxel: fw-1 of its fw columns are shared.  This
// kernel keeps the whole window of the chain -- residual AND 1/variance -- in
// registers across sites:
//
//   * (fw+1) column groups of ZL = Dp/VEC threads; group r owns the column X with
//     X = r (mod fw+1) of [x-fhw, x+fhw+1] and keeps its fh rows in registers (one
//     16-byte vector per row);
//   * per site only the group whose column left the window writes it back and
//     loads the column that will ENTER the window at the next site (the window is
//     fw wide, fw+1 columns are resident).  HBM/L2 traffic per site drops from
//     3*fh*fw to ~3*fh vectors (all columns are reloaded when the row changes);
//   * the sums h[z], G[z] (d3d_kernels.cuh header) and the residual update are
//     pure register arithmetic.
//
// Warp roles (named barriers, PTX bar.sync / bar.arrive):
//   W  window warps                       sums, partial reduction, update
//   A  accept uniform + first truncated-normal draws (lane-parallel log/sqrt/cos)
//      + old line profile Lu(c_old,w_old)   } up to two sites ahead of the decisions,
//   P  proposal (Philox, Cauchy jump)       } double-buffered in shared memory
//      + new line profile Lu(c_new,w_new)   }
//   B  accept test + truncated-normal Gibbs draw
// Barrier ids: 2,3 READY[parity] (A,P arrive; W,B wait)  4,5 FREE[parity]
// (W,B arrive two sites later; A,P wait)  6 partials ready (W,B)  7 decision
// broadcast (W,B).  Look-ahead never crosses a sweep boundary, where the
// acceptance-rate test of lib/run.py:344-359 needs the finished sweep.
//
// Code size matters here: the per-site critical path is a latency-bound chain of a few
// hundred instructions, and a kernel image beyond the ~32 KB instruction cache triples the
// cost of every one of them (measured).  Hence rolled loops wherever no register
// array is indexed, one copy of each libm routine, and cold paths out of line.
#pragma once

namespace d3d {

#ifdef D3D_TRACE
// Debug build only (scratch tooling, see profiles/r01_notes.md): per-warp clock64() stamps of
// 16 consecutive sites of chain 0.  TRD stamps behind the arrival of a value (a stamp right
// after a bar.sync only records the issue of the barrier, not its release).
__device__ unsigned long long g_tr2[16 * 16 * 16];
__device__ __forceinline__ int tr_dep(double v) { return __double2hiint(v) == 0x7ff12345 ? 1 : 0; }
#define TRW(ev)                                                                             \
    do { if (lane == 0 && blockIdx.x == 0 && it == it0 + 3 && j >= 700 && j < 716)          \
             g_tr2[((j - 700) * 16 + warp) * 16 + (ev)] = clock64(); } while (0)
#define TRD(ev, v)                                                                          \
    do { int z_ = tr_dep(v);                                                                \
         if (lane == 0 && blockIdx.x == 0 && it == it0 + 3 && j >= 700 && j < 716)          \
             g_tr2[((j - 700) * 16 + warp) * 16 + (ev) + z_] = clock64(); } while (0)
#else
#define TRW(ev)
#define TRD(ev, v)
#endif


__device__ __forceinline__ void bar_arrive_named(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// 8 per-lane partial sums -> 8 warp totals with 7 double shuffles (instead of 8 x 5):
// each round halves the number of slots a lane is responsible for.  On return the
// lanes with (lane & 3) == 0 hold the total of slot warp_sum8_slot(lane).
__device__ __forceinline__ double warp_sum8(const double* v, int lane) {
    double w4[4], w2[2], w1;
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double send = b4 ? v[i] : v[i + 4];
        const double keep = b4 ? v[i + 4] : v[i];
        w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const double send = b3 ? w4[i] : w4[i + 2];
        const double keep = b3 ? w4[i + 2] : w4[i];
        w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    {
        const double send = b2 ? w2[0] : w2[1];
        const double keep = b2 ? w2[1] : w2[0];
        w1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    w1 += __shfl_xor_sync(0xffffffffu, w1, 2);
    w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
    return w1;
}
__device__ __forceinline__ int warp_sum8_slot(int lane) {
    return ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
}

// Two per-lane partial sums -> two warp totals with 5 double shuffles: lane 0 ends with the
// total of p0, lane 16 with the total of p1.
__device__ __forceinline__ double warp_sum2(double p0, double p1, int lane) {
    const bool b4 = lane & 16;
    double w = (b4 ? p1 : p0) + __shfl_xor_sync(0xffffffffu, b4 ? p0 : p1, 16);
    w += __shfl_xor_sync(0xffffffffu, w, 8);
    w += __shfl_xor_sync(0xffffffffu, w, 4);
    w += __shfl_xor_sync(0xffffffffu, w, 2);
    w += __shfl_xor_sync(0xffffffffu, w, 1);
    return w;
}
// Four per-lane partial sums -> four warp totals with 6 double shuffles: every lane ends
// with the total of slot ((lane >> 3) & 1) * 2 + ((lane >> 4) & 1), i.e. lanes 0 / 16 / 8 / 24
// hold q0 / q1 / q2 / q3.
__device__ __forceinline__ double warp_sum4(const double* q, int lane) {
    const bool b4 = lane & 16, b3 = lane & 8;
    const double w01 = (b4 ? q[1] : q[0]) + __shfl_xor_sync(0xffffffffu, b4 ? q[0] : q[1], 16);
    const double w23 = (b4 ? q[3] : q[2]) + __shfl_xor_sync(0xffffffffu, b4 ? q[2] : q[3], 16);
    double w = (b3 ? w23 : w01) + __shfl_xor_sync(0xffffffffu, b3 ? w01 : w23, 8);
    w += __shfl_xor_sync(0xffffffffu, w, 4);
    w += __shfl_xor_sync(0xffffffffu, w, 2);
    w += __shfl_xor_sync(0xffffffffu, w, 1);
    return w;
}

// G[cube][site][z] = sum over the in-cube part of the FSF window of F^2 / sigma^2 (header of
// d3d_kernels.cuh).  It depends on the FSF, the variance and the site only -- not on the residual
// or the parameters -- so the sliding-window kernel reads it from this table (built once per
// problem, H*W*Dp doubles per cube) instead of accumulating it at every site update: the window
// warps are left with h[z] alone, and the four quadratic sums over G are formed by warp B from
// the line profiles while it waits for the window warps.
template <typename T, bool IVCUBE>
__global__ void gtable_kernel(const __grid_constant__ Problem pb, double* gtab) {
    const int HW = pb.H * pb.W;
    const int sc = blockIdx.x;                         // cube * HW + site
    const int cube = sc / HW, site = sc - cube * HW;
    const int y = site / pb.W, x = site - y * pb.W;
    const T* iv = IVCUBE ? (const T*)pb.iv + (size_t)cube * HW * pb.Dp : nullptr;
    const double ivs = IVCUBE ? 0.0 : pb.iv_scalar[cube];
    for (int z = threadIdx.x; z < pb.Dp; z += blockDim.x) {
        double g = 0.0;
        for (int i = 0; i < pb.fh; ++i) {
            const int Y = y - pb.fhh + i;
            if (Y < 0 || Y >= pb.H) continue;
            for (int k = 0; k < pb.fw; ++k) {
                const int X = x - pb.fhw + k;
                if (X < 0 || X >= pb.W) continue;
                const double f = pb.fsf[i * pb.fw + k];
                const double w = IVCUBE ? (double)iv[((size_t)Y * pb.W + X) * pb.Dp + z] : ivs;
                g = fma(f * f, w, g);
            }
        }
        gtab[(size_t)sc * pb.Dp + z] = z < pb.D ? g : 0.0;
    }
}

template <typename T, bool IVCUBE, int NE>
__global__ void __launch_bounds__(384, 1)
sweep_seq_slide_kernel(const __grid_constant__ Problem pb, long long it0_all, long long it1_all,
                       int keep, double min_rate, double* chain_out, double* lik_out,
                       long long row_first, long long rows_local, const int4* items,
                       const int* item_count, int max_items, volatile long long* progress) {
    typedef typename Vec<T>::V V;
    const int VEC = Vec<T>::N;
    extern __shared__ double smem_raw[];
    Smem sm;
    carve(sm, smem_raw, pb);
    load_constants(sm, pb);
    // Work items (chain, first iteration, end iteration) of this CTA.  With more chains than
    // SMs the host lays the chain x sweep rectangle over the CTAs by McNaughton's wrap-around
    // rule, so that every SM gets the same number of sweeps; a chain split over two CTAs is
    // handed over through `progress` (both CTAs are resident: grid <= number of SMs).
    const int n_items = items ? item_count[blockIdx.x] : 1;
    // truncated-normal tables (lib/rtnorm.py:227-2681) in shared memory for the inline table
    // branch of warp B: x[4002], yu[4001] as doubles, ncell[8961] as 16-bit.  Copied once per
    // CTA, BEFORE the item loop: an item whose chain was stopped earlier skips the loop body.
    double* const tx = smem_raw + smem_doubles(pb.fh, pb.fw, pb.P, pb.Dp);
    double* const tyu = tx + 4002;
    unsigned short* const tnc = (unsigned short*)(tyu + 4002);
    for (int q = threadIdx.x; q < 4002; q += blockDim.x) tx[q] = pb.rt.x[q];
    for (int q = threadIdx.x; q < 4001; q += blockDim.x) tyu[q] = pb.rt.yu[q];
    for (int q = threadIdx.x; q < 8961; q += blockDim.x) tnc[q] = (unsigned short)pb.rt.ncell[q];
    __syncthreads();
  for (int item = 0; item < n_items; ++item) {
    int chain; long long it0, it1;
    if (items) {
        const int4 w = items[(size_t)blockIdx.x * max_items + item];
        chain = w.x; it0 = it0_all + w.y; it1 = it0_all + w.z;
        if (threadIdx.x == 0) while (progress[chain] < it0) __nanosleep(200);
        __syncthreads();
        __threadfence();
    } else {
        chain = blockIdx.x; it0 = it0_all; it1 = it1_all;
        if (chain >= pb.n_chains) return;
    }
    const int cube = chain / pb.chains_per_cube;
    if (!pb.active[chain]) {                       // chain stopped earlier (acceptance rate)
        __syncthreads();
        if (items && threadIdx.x == 0) progress[chain] = it1;
        continue;
    }
    __syncthreads();                               // previous item fully retired
    if (threadIdx.x == 0) sm.bc[2] = (double)pb.accepted[chain];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Dp = pb.Dp, W = pb.W, H = pb.H, fw = pb.fw, fh = pb.fh, fhh = pb.fhh;
    const int ZL = Dp / VEC;
    const int ngrp = fw + 1;                       // resident columns: the window + the next one
    const int nwt = ngrp * ZL;
    const int nww = (nwt + 31) >> 5;
    const bool roleW = warp < nww, roleA = warp == nww, roleP = warp == nww + 1,
               roleB = warp == nww + 2;
    const int cntAll = (nww + 3) * 32, cntWB = (nww + 1) * 32;
    const int grp = tid / ZL, zp = tid - grp * ZL;
    const bool wt = tid < nwt;

    const int ns = pb.n_sites[cube];
    // (the host sends one-column fields to the row-mapped kernel: W >= 2 here, the magic fits)
    const unsigned magicW = (unsigned)((0x100000000ull + (unsigned)W - 1u) / (unsigned)W);
    const int* sites = pb.sites + (size_t)cube * pb.max_sites;
    const size_t HW = (size_t)H * W;
    const size_t row_bytes = (size_t)W * Dp * sizeof(T);
    char* const err = (char*)((T*)pb.err + (size_t)chain * HW * Dp);
    const char* const ivc =
        IVCUBE ? (const char*)((const T*)pb.iv + (size_t)cube * HW * Dp) : nullptr;
    const double ivs = IVCUBE ? 0.0 : pb.iv_scalar[cube];
    const double ira = 1.0 / pb.prior_var[cube];

    __syncthreads();

    // register-resident column of this thread: rows y-fhh .. y-fhh+fh-1 of column heldX.
    // fw+1 column groups: group r owns the column X = r (mod fw+1) of [x-fhw, x+fhw+1]; the
    // extra column is the one that enters the window at the next site, so its loads (issued
    // when the group's previous column leaves) have a whole site update to complete.
    V ecache[NE];
    V ivcache[IVCUBE ? NE : 1];
    const int NOCOL = -(1 << 30);
    int heldX = NOCOL, heldY = NOCOL;
    const size_t rstride = (size_t)W * Dp;
    T* const errT = (T*)err;
    const T* const ivT = (const T*)ivc;

    double rate = pb.rate[chain];
    long long accepted = pb.accepted[chain];       // owned by lane 0 of warp B
    long long it = it0;
    int alive = 1;

    for (; it < it1; ++it) {
        if (!(rate > min_rate || rate == 0.0)) { alive = 0; break; }   // lib/run.py:344-350
        const double max_acc = (double)ns * (double)it;                // :356-359
        if (max_acc > 0.0) rate = sm.bc[2] / max_acc;
        const bool save = (it % keep) == 0;                            // :353
        double* crow = nullptr; double* lrow = nullptr;
        if (save) {
            long long r = it / keep - row_first;
            if (chain_out) crow = chain_out + ((size_t)chain * rows_local + r) * HW * 3;
            if (lik_out) lrow = lik_out + ((size_t)chain * rows_local + r) * HW;
        }

        for (int j = 0; j < ns; ++j) {
            const int par = j & 1;
            const int site = sites[j];
            // y = site / W by a multiply-high with ceil(2^32 / W): exact while site * W < 2^32,
            // which the host guarantees (H * W * W < 2^32, d3d_api.cu choose_launch)
            const int y = (int)__umulhi((unsigned)site, magicW), x = site - y * W;
            double* Lu_o = sm.Lu_o + par * Dp;
            double* Lu_n = sm.Lu_n + par * Dp;
            double* prop_s = sm.prop + par * 8;
            double* spec_s = sm.spec + par * 16;

            PH_T0();
            if (roleW) {
                TRW(0);
                // ---- make the resident columns current for (y, x) --------------------
                const int xl = x - pb.fhw;
                // (square FSF of the template's size: modulo by a compile-time constant)
                int m = fw == NE ? (grp - xl) % (NE + 1) : (grp - xl) % ngrp;
                if (m < 0) m += ngrp;
                const int Xn = xl + m;                           // column of this group
                if (wt && Xn != heldX && y == heldY && fh == NE && y >= fhh && y + fhh < H &&
                    heldX >= 0 && heldX < W && Xn >= 0 && Xn < W) {
                    // interior site, same row: all NE rows exist -- straight-line write-back and
                    // reload without per-row tests (this block delays the two warps that hold
                    // the switching group, profiles/r01_notes.md)
                    const size_t rowoff = (size_t)(y - fhh) * W;
                    T* p = errT + (rowoff + heldX) * Dp + zp * VEC;
#pragma unroll
                    for (int i = 0; i < NE; ++i) { *(V*)p = ecache[i]; p += rstride; }
                    const T* pe = errT + (rowoff + Xn) * Dp + zp * VEC;
#pragma unroll
                    for (int i = 0; i < NE; ++i) { ecache[i] = *(const V*)pe; pe += rstride; }
                    if (IVCUBE) {
                        const T* q = ivT + (rowoff + Xn) * Dp + zp * VEC;
#pragma unroll
                        for (int i = 0; i < NE; ++i) { ivcache[i] = *(const V*)q; q += rstride; }
                    }
                    heldX = Xn;
                } else if (wt && (Xn != heldX || y != heldY)) {
                    const int ytop_old = heldY - fhh;
                    if (heldX >= 0 && heldX < W && heldY != NOCOL) {       // write back
                        T* p = errT + ((size_t)max(ytop_old, 0) * W + heldX) * Dp + zp * VEC;
#pragma unroll
                        for (int i = 0; i < NE; ++i) {
                            const int Y = ytop_old + i;
                            if (i < fh && Y >= 0 && Y < H) { *(V*)p = ecache[i]; p += rstride; }
                        }
                    }
                    const int ytop = y - fhh;
                    if (Xn >= 0 && Xn < W) {
                        const size_t off = ((size_t)max(ytop, 0) * W + Xn) * Dp + zp * VEC;
                        const T* p = errT + off;
                        const T* q = IVCUBE ? ivT + off : nullptr;
#pragma unroll
                        for (int i = 0; i < NE; ++i) {
                            const int Y = ytop + i;
                            if (i < fh && Y >= 0 && Y < H) {
                                ecache[i] = *(const V*)p; p += rstride;
                                if (IVCUBE) { ivcache[i] = *(const V*)q; q += rstride; }
                            } else {
                                ecache[i] = V();
                                if (IVCUBE) ivcache[i] = V();
                            }
                        }
                    }
                    heldX = Xn; heldY = y;
                }
                // ---- window sums (registers only) -------------------------------------
                const int dx = heldX - xl;                       // FSF column, valid if active
                const bool active = wt && dx >= 0 && dx < fw && heldX >= 0 && heldX < W;
                double pB = 0.0, pPO = 0.0;
                double h[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) h[v] = 0.0;
                const double* fcol = sm.F + (active ? dx : 0);
                if (active) {
                    // h[z] only: G[z] comes from the table (gtable_kernel), warp B forms its sums
#pragma unroll
                    for (int i = 0; i < NE; ++i) {
                        if (i < fh) {
                            const double f = fcol[i * fw];
                            double e[VEC];
                            unpack(ecache[i], e);
                            if (IVCUBE) {
                                double w_[VEC];
                                unpack(ivcache[i], w_);
#pragma unroll
                                for (int v = 0; v < VEC; ++v) h[v] = fma(f * w_[v], e[v], h[v]);
                            } else {
#pragma unroll
                                for (int v = 0; v < VEC; ++v) h[v] = fma(f, e[v], h[v]);
                            }
                        }
                    }
                    if (!IVCUBE) {
#pragma unroll
                        for (int v = 0; v < VEC; ++v) h[v] *= ivs;
                    }
                }
                if (warp == 0) PH_ADD(0);                        // [0] W switch + sums
                TRD(1, h[0]);
                bar_sync_named(2 + par, cntAll);                 // READY: profiles of site j
                if (warp == 0) PH_ADD(1);                        // [1] W wait READY
                double lo_v[VEC], ln_v[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    lo_v[v] = Lu_o[zp * VEC + v];
                    ln_v[v] = Lu_n[zp * VEC + v];
                    if (active) {
                        pB = fma(lo_v[v] - ln_v[v], h[v], pB);
                        pPO = fma(lo_v[v], h[v], pPO);
                    }
                }
                {
                    const double tot = warp_sum2(pB, pPO, lane);
                    if ((lane & 15) == 0) sm.red[warp * 2 + (lane >> 4)] = tot;
                }
                if (warp == 0) PH_ADD(2);                        // [2] W partials
                TRW(3);
                bar_sync_named(6, cntWB);                        // partials visible to B
                if (warp == 0) PH_ADD(3);                        // [3] W housekeeping
                bar_sync_named(7, cntWB);                        // decision broadcast
                if (warp == 0) PH_ADD(4);                        // [4] W wait decision
                TRD(4, sm.bc[1]);
                if (active) {
                    const int acc = sm.bc[0] != 0.0;
                    const double r = sm.bc[1], a = sm.bc[3];
                    double coef[VEC];
#pragma unroll
                    for (int v = 0; v < VEC; ++v)
                        coef[v] = a * lo_v[v] - r * (acc ? ln_v[v] : lo_v[v]);
#pragma unroll
                    for (int i = 0; i < NE; ++i) {
                        const int Y = y - fhh + i;
                        if (i < fh && Y >= 0 && Y < H) {         // keep the zero padding rows zero
                            const double f = fcol[i * fw];
                            double e[VEC];
                            unpack(ecache[i], e);
#pragma unroll
                            for (int v = 0; v < VEC; ++v) e[v] = fma(f, coef[v], e[v]);
                            pack(ecache[i], e);
                        }
                    }
                }
                if (warp == 0) PH_ADD(5);                        // [5] W update
                TRD(5, reinterpret_cast<const double&>(ecache[0]));
                if (j + 2 < ns) bar_arrive_named(4 + par, cntAll);   // FREE: buffers of site j
            } else if (roleB) {
                // G[z] of this site (static table); the loads fly while B waits for the profiles
                const double* gt = pb.gtab + ((size_t)cube * HW + site) * Dp;
                const double g0 = lane < Dp ? gt[lane] : 0.0;
                const double g1 = lane + 32 < Dp ? gt[lane + 32] : 0.0;
                bar_sync_named(2 + par, cntAll);
                // quadratic sums over G (C, QOO, QON, QNN of d3d_kernels.cuh): they need the two
                // line profiles only, so B forms them while the window warps reduce h
                double qs[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
                for (int z = lane; z < Dp; z += 32) {
                    const double G = z < 32 ? g0 : (z < 64 ? g1 : gt[z]);
                    const double lo = Lu_o[z], ln = Lu_n[z], dl = lo - ln;
                    qs[0] = fma(dl * dl, G, qs[0]);
                    qs[1] = fma(lo * lo, G, qs[1]);
                    qs[2] = fma(lo * ln, G, qs[2]);
                    qs[3] = fma(ln * ln, G, qs[3]);
                }
                const double qr = warp_sum4(qs, lane);
                bar_sync_named(6, cntWB);
                PH_ADD(6);                                       // [6] B waits
                // totals: lanes 0, 1 add the nww warp partials of the two sums over h
                double t = 0.0;
                if (lane < 2)
                    for (int wv = 0; wv < nww; ++wv) t += sm.red[wv * 2 + lane];
                double tot[R_N];
                tot[R_B] = __shfl_sync(0xffffffffu, t, 0);
                tot[R_PO] = __shfl_sync(0xffffffffu, t, 1);
                tot[R_C] = __shfl_sync(0xffffffffu, qr, 0);
                tot[R_QOO] = __shfl_sync(0xffffffffu, qr, 16);
                tot[R_QON] = __shfl_sync(0xffffffffu, qr, 8);
                tot[R_QNN] = __shfl_sync(0xffffffffu, qr, 24);
                const double a = prop_s[0], c_old = prop_s[1], w_old = prop_s[2];
                const double a_new = prop_s[3], c_new = prop_s[4], w_new = prop_s[5];
                const bool oob = prop_s[7] != 0.0;
                const double log_u = spec_s[SP_LOGU];
                PH_ADD(7);                                       // [7] B totals
                TRD(8, tot[0]);
                TRD(10, a);
                // accept test (lib/run.py:426-451)
                double delta;
                if (a_new != a) {
                    const double Pn = tot[R_PO] - tot[R_B];
                    const double Bq = a * tot[R_PO] - a_new * Pn;
                    const double Cq = a * a * tot[R_QOO] - 2.0 * a * a_new * tot[R_QON] +
                                      a_new * a_new * tot[R_QNN];
                    delta = -Bq - 0.5 * Cq;
                } else {
                    delta = -(a * tot[R_B]) - 0.5 * (a * a) * tot[R_C];
                }
                const int acc = (log_u < delta) && !oob;                        // :438
                const double c_end = acc ? c_new : c_old, w_end = acc ? w_new : w_old;
                const double S2 = acc ? tot[R_QNN] : tot[R_QOO];
                const double S1 = acc ? (tot[R_PO] - tot[R_B]) + a * tot[R_QON]
                                      : tot[R_PO] + a * tot[R_QOO];
                // Gibbs draw (lib/run.py:491-496): ro = ra/(1+ra S2) = 1/q, q = 1/ra + S2;
                // sigma = sqrt(ro) = rsqrt(q): ONE serial slow operation instead of
                // divide -> sqrt -> divide; mu = ro S1.
                const double q = ira + S2;
                const double sg = rsqrt(q);
                const double isg = q * sg;
                const double mu = S1 * (sg * sg);
                const double lo = pb.pmin[cube * 3], hi = pb.pmax[cube * 3];
                const double as = (lo - mu) * isg, bs = (hi - mu) * isg;   // lib/rtnorm.py:74-76
                double rs = 0.0;
                int fail = 0;
                bool done = false;
                const bool plain = (as < bs) && !(fabs(as) > fabs(bs));   // no mirror (:108)
                if (plain && as < -2.00443204036) {                        // :127-131
                    // Gaussian proposal; the first two normals were evaluated ahead
                    const double n1 = spec_s[SP_N1], n2 = spec_s[SP_N2];
                    if (n1 >= as && n1 <= bs) { rs = n1; done = true; }
                    else if (n2 >= as && n2 <= bs) { rs = n2; done = true; }
                } else if (plain && as > 3.48672170399 && -as * (bs - as) < -40.0) {   // :112-124
                    // right tail with exp(-a(b-a)) below 2^-54: expab == -1 exactly, so the
                    // first z = log(1 - u) and e = -log(u') were evaluated ahead
                    const double z = spec_s[SP_Z1];
                    if (2.0 * as * as * spec_s[SP_E1] > z * z) { rs = as - z / as; done = true; }
                }
                else if (plain && as >= -2.00443204036 && as <= 3.48672170399) {     // :133-222
                    // Chopin's table method, first try from the pre-drawn uniforms u4 (strip)
                    // and u5 (position); tables in shared memory
                    const int N = 4000;
                    const int ka = tnc[3271 + (int)floor(as * 1631.73284006)];
                    const int kb = bs >= 3.48672170399 ? N : tnc[3271 + (int)floor(bs * 1631.73284006)];
                    if (kb - ka >= 5) {
                        const int k = ka + (int)floor(spec_s[SP_U4] * (double)(kb + 1 - ka));
                        if (k != N && !(k <= ka + 2 || (k >= kb && bs < 3.48672170399))) {
                            const double u = spec_s[SP_U5];
                            const double yuk = tyu[k], xk = tx[k], dk = tx[k + 1] - xk;
                            const double ylk = k == 1 ? 0.053513975472 : (k <= 1954 ? tyu[k - 1] : tyu[k + 1]);
                            if (yuk * u < ylk) { rs = xk + u * dk * yuk / ylk; done = true; }
                        }
                    }
                }
                if (!done) {                                     // every other branch / retry
                    Philox rng;
                    rng.init(pb.seed, pb.first_chain + (unsigned)chain, (unsigned)it,
                             (unsigned)site);
                    rng.k = 4;
                    rng.stash = spec_s;
                    rs = rtstdnorm(as, bs, rng, pb.rt, &fail, nullptr);
                }
                const double r = rs * sg + mu;                                  // :82-83
                if (lane == 0) {
                    if (fail) atomicExch(pb.status, 1);
                    double* prm = pb.params + ((size_t)chain * HW + site) * 3;
                    prm[0] = r; prm[1] = c_end; prm[2] = w_end;                 // :448,:499,:516
                    if (crow) {
                        double* cr = crow + (size_t)site * 3;
                        cr[0] = r; cr[1] = c_end; cr[2] = w_end;
                    }
                    if (lrow) lrow[site] = delta;                               // :430-432
                    sm.bc[0] = acc ? 1.0 : 0.0;
                    sm.bc[1] = r;
                    sm.bc[3] = a;
                    accepted += acc;
                    if (j == ns - 1) sm.bc[2] = (double)accepted;
                }
                PH_ADD(8);                                       // [8] B decide
                TRW(9);
                bar_sync_named(7, cntWB);
                if (j + 2 < ns) bar_arrive_named(4 + par, cntAll);
            } else if (roleA || roleP) {
                if (j >= 2) bar_sync_named(4 + par, cntAll);     // buffers of this parity free
                PH_ADD(9);                                       // [9] A/P wait FREE
                TRW(12);
                const double* prm = pb.params + ((size_t)chain * HW + site) * 3;
                const double a = prm[0], c_old = prm[1], w_old = prm[2];
                // Philox blocks 0..3 in lanes 0..3: draws (2b, 2b+1) of this site
                unsigned o[4];
                philox_block_inl((unsigned)pb.seed, (unsigned)(pb.seed >> 32), (unsigned)(lane & 3),
                                 (unsigned)site, (unsigned)it, pb.first_chain + (unsigned)chain, o);
                const double ua = Philox::u53(o[0], o[1]), ub = Philox::u53(o[2], o[3]);
                double c_prof, w_prof;
                double* g_buf; double* Lu_out;
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
                        prop_s[0] = a; prop_s[1] = c_old; prop_s[2] = w_old;
                        prop_s[3] = a_new; prop_s[4] = c_new; prop_s[5] = w_new;
                        prop_s[7] = (double)oob;
                    }
                    c_prof = c_new; w_prof = w_new; g_buf = sm.g_n; Lu_out = Lu_n;
                    PH_ADD(12);                                  // [12] P proposal
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
                        spec_s[SP_U4] = u4; spec_s[SP_U5] = u5; spec_s[SP_U6] = u6; spec_s[SP_U7] = u7;
                        spec_s[SP_E1] = e1; spec_s[SP_Z1] = z1; spec_s[SP_N1] = nv; spec_s[SP_N2] = n2;
                        spec_s[SP_LOGU] = lv;
                    }
                    c_prof = c_old; w_prof = w_old; g_buf = sm.g_o; Lu_out = Lu_o;
                    PH_ADD(10);                                  // [10] A draws
                }
                // ---- unit line profile Lu = lsf (*) exp(-(z-c)^2 / (2 w^2)) -------------
                {
                    const double inv2w2 = 1.0 / (2.0 * (w_prof * w_prof));
                    const int D = pb.D, P = pb.P;
#pragma unroll 1
                    for (int z = lane; z < D; z += 32) {
                        const double d0 = (double)z - c_prof;
                        double gv = exp(-1.0 * (d0 * d0) * inv2w2);
                        if (pb.n_comp > 1) gv += extra_components(pb, d0, inv2w2);
                        g_buf[z] = gv;
                    }
                    __syncwarp();
                    if (pb.has_lsf) {
                        const int nt = pb.ntaps;
#pragma unroll 1
                        for (int z = lane; z < Dp; z += 32) {
                            double a0 = 0.0, a1 = 0.0;
                            int tq = 0;
#pragma unroll 1
                            for (; tq + 1 < nt; tq += 2) {
                                a0 = fma(sm.Kv[tq], g_buf[(z - sm.Km[tq]) & (P - 1)], a0);
                                a1 = fma(sm.Kv[tq + 1], g_buf[(z - sm.Km[tq + 1]) & (P - 1)], a1);
                            }
                            if (tq < nt) a0 = fma(sm.Kv[tq], g_buf[(z - sm.Km[tq]) & (P - 1)], a0);
                            Lu_out[z] = z < D ? a0 + a1 : 0.0;
                        }
                    } else {                                     // lib/run.py:675-676
                        for (int z = lane; z < Dp; z += 32) Lu_out[z] = z < D ? g_buf[z] : 0.0;
                    }
                }
                PH_ADD(13);                                      // [13] A/P profile
                TRW(13);
                __threadfence_block();
                bar_arrive_named(2 + par, cntAll);
            }
        }
        __syncthreads();                           // sweep boundary: bc[2] visible, pipeline drained
    }

    // ---- write the resident columns back ----------------------------------------------------
    if (roleW && wt && heldX >= 0 && heldX < W && heldY != NOCOL) {
        const int ytop_old = heldY - fhh;
        T* p = errT + ((size_t)max(ytop_old, 0) * W + heldX) * Dp + zp * VEC;
#pragma unroll
        for (int i = 0; i < NE; ++i) {
            const int Y = ytop_old + i;
            if (i < fh && Y >= 0 && Y < H) { *(V*)p = ecache[i]; p += rstride; }
        }
    }
    if (roleB && lane == 0) {
        pb.accepted[chain] = accepted;
        pb.rate[chain] = rate;
        pb.iters[chain] = it;
        if (!alive) pb.active[chain] = 0;
    }
    if (items) {                                   // hand the chain over to its next owner
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) progress[chain] = it1;
    }
  }
}

}  // namespace d3d
