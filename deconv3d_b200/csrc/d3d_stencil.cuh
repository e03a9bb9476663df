// d3d_stencil.cuh -- register-tiled spatial pass of the forward model.
//
// sim[z][y'][x'] = sum_{j,k} F[j][k] * lines[z][y'+fhh-j][x'+fhw-k]   (true 2-D convolution,
// zero 'same' borders: scipy.signal.convolve2d(...,'same') at lib/run.py:1027-1029, equal to
// the paste of lib/run.py:697-706), fused with residual = data - sim and 0.5*sum(e^2/var).
//
// CTA = 256 threads = 8 z-pairs x 4 x-blocks x 8 rows; a thread produces RX = 4 consecutive
// x outputs for one z-pair (8 accumulators).  The (TY+fh-1) x (TX+FW-1) x 16-channel halo tile
// of `lines` sits in shared memory z-fastest: the 8 z-pair lanes of an element read 128
// contiguous bytes (one conflict-free LDS.128 wavefront per quarter warp).  Per FSF row a
// thread loads FW+3 tile vectors once and re-uses each of them for up to 4 outputs, and every
// F[j][k] (one broadcast LDS) for 8 FMAs: ~3.6 DFMA per shared-memory load instead of 0.5
// in the scalar kernel, which is what moves the pass from LDS-bound to the FP64 pipe.
#pragma once

namespace d3d {

template <typename T> struct Pair;
template <> struct Pair<double> { typedef double2 P; };
template <> struct Pair<float>  { typedef float2  P; };

// TMA = true: the halo tile arrives as ONE bulk tensor copy (cp.async.bulk.tensor, a 4-D box of the
// [set][y][x][z] array issued by one thread, completion on an mbarrier): the box may start at
// negative coordinates and end beyond the field -- the TMA unit fills those positions with zeros,
// which IS the 'same' border -- and the ~300 address / predicate / LDGSTS instructions per thread
// of the staging loop disappear.  TMA = false keeps the cp.async loop (no tensor map at hand).
template <typename T, int FW, int RY, bool TMA>
__global__ void __launch_bounds__(256, 2)
stencil_tiled_kernel(const __grid_constant__ Problem pb, const double* __restrict__ lines,
                     double* sim_out, int write_err, double* chi2_out, const __grid_constant__ CUtensorMap tmap) {
    // RY output rows per thread (tile of 8*RY x 16 spaxels): a tile row loaded from shared memory
    // serves RY output rows, and the FSF row is read as 16-byte pairs (row stride FW+1), so that
    // the LDS pipe (4 cycles per LDS.128 of a warp) stays below the FP64 pipe (2 DFMA warps/cycle):
    // RY = 1: 16 tile + 13 FSF loads per 104 DFMA (13 columns); RY = 2: 16 + 14 per 208.
    const int TY = 8 * RY, TX = 16, RX = 4, ZC = 16, FWP = FW + 1;
    extern __shared__ __align__(128) double smem_tile[];
    const int fh = pb.fh, Dp = pb.Dp, D = pb.D, H = pb.H, W = pb.W;
    const int hx = TX + FW - 1, hy = TY + fh - 1;
    double* tile = smem_tile;                               // [hy][hx][ZC]: 128-byte aligned (TMA destination)
    double* F = tile + (size_t)hy * hx * ZC;                // [fh][FW+1], last column 0
    unsigned long long* mbar = (unsigned long long*)(F + ((fh * FWP + 1) & ~1));
    const int ty_n = (H + TY - 1) / TY, tx_n = (W + TX - 1) / TX;
    const int chain = blockIdx.x / (ty_n * tx_n);
    const int trem = blockIdx.x - chain * ty_n * tx_n;
    const int ty0 = (trem / tx_n) * TY, tx0 = (trem % tx_n) * TX;
    const int z0 = blockIdx.y * ZC;
    const int cube = chain / pb.chains_per_cube;
    const int tid = threadIdx.x;

    if (TMA) {
        const unsigned mb = (unsigned)__cvta_generic_to_shared(mbar);
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            const unsigned bytes = (unsigned)(hy * hx * ZC * sizeof(double));
            const unsigned dst = (unsigned)__cvta_generic_to_shared(tile);
            asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(mb), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                         ::"r"(dst), "l"(&tmap), "r"(z0), "r"(tx0 - pb.fhw), "r"(ty0 - pb.fhh), "r"(chain), "r"(mb) : "memory");
        }
    }
    for (int i = tid; i < fh * FWP; i += 256) {
        const int j = i / FWP, k = i - j * FWP;
        F[i] = k < FW ? pb.fsf[j * FW + k] : 0.0;
    }
    const double* lc = lines + (size_t)chain * H * W * Dp;
    if (write_err || chi2_out) {
        // the epilogue's `data` vectors: ask L2 for them now, so that the loads behind the FSF loop
        // pay an L2 hit instead of a DRAM round trip with only 16 warps per SM to hide it
        const T* data = (const T*)pb.data + (size_t)cube * H * W * Dp;
        const int pz = z0 + 2 * (tid & 7), pxb = (tid >> 3) & 3, poy = tid >> 5;
#pragma unroll
        for (int ry = 0; ry < RY; ++ry)
#pragma unroll
            for (int r = 0; r < RX; ++r) {
                const int gy = ty0 + poy * RY + ry, gx = tx0 + pxb * RX + r;
                if (gy < H && gx < W && pz < Dp)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(data + ((size_t)gy * W + gx) * Dp + pz));
            }
    }
    // Halo tile: global -> shared memory with cp.async (LDGSTS, 16 bytes per copy, no register
    // staging): every copy of the thread is in flight at once, so the staging phase costs about one
    // memory round trip instead of one per group of four loads (2 CTAs of 8 warps per SM cannot
    // hide them otherwise).  Positions outside the field are the zero 'same' border.
    if (!TMA)
#pragma unroll 4
    for (int i = tid; i < hy * hx * (ZC / 2); i += 256) {    // double2 granularity
        const int zq = i & 7, s = i >> 3;
        const int sy = s / hx, sx = s - sy * hx;
        const int gy = ty0 + sy - pb.fhh, gx = tx0 + sx - pb.fhw;
        const int z = z0 + 2 * zq;
        double* dst = tile + (size_t)s * ZC + 2 * zq;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W && z < Dp) {
            const unsigned sa = (unsigned)__cvta_generic_to_shared(dst);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(lc + ((size_t)gy * W + gx) * Dp + z) : "memory");
        } else {
            *(double2*)dst = make_double2(0.0, 0.0);
        }
    }
    if (TMA) {
        const unsigned mb = (unsigned)__cvta_generic_to_shared(mbar);
        unsigned ok = 0;
        const long long t0 = clock64();
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(mb) : "memory");
            // (never seen; a tile that does not arrive within ~2 s must end in an error, not in a hung GPU)
            if (!ok && clock64() - t0 > 4000000000LL) { atomicExch(pb.status, 3); break; }
        }
    } else {
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();

    const int zq = tid & 7, xb = (tid >> 3) & 3, oy = tid >> 5;
    double acc[RY][RX][2];
#pragma unroll
    for (int ry = 0; ry < RY; ++ry)
#pragma unroll
        for (int r = 0; r < RX; ++r) { acc[ry][r][0] = 0.0; acc[ry][r][1] = 0.0; }
    // tile row oy*RY + trr feeds FSF row j = ry + fh-1 - trr of output row ry; trr descending keeps
    // the accumulation order of every output at j = 0, 1, ... (k ascending inside)
    for (int trr = fh + RY - 2; trr >= 0; --trr) {
        const double* trow = tile + ((size_t)(oy * RY + trr) * hx + xb * RX) * ZC + 2 * zq;
        double2 v[FW + RX - 1];
#pragma unroll
        for (int u = 0; u < FW + RX - 1; ++u) v[u] = *(const double2*)(trow + (size_t)u * ZC);
#pragma unroll
        for (int ry = 0; ry < RY; ++ry) {
            const int j = ry + fh - 1 - trr;
            if (j < 0 || j >= fh) continue;                  // uniform over the CTA
            const double2* frow = (const double2*)(F + j * FWP);
#pragma unroll
            for (int k2 = 0; k2 < FWP / 2; ++k2) {
                const double2 f2 = frow[k2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int k = 2 * k2 + h;
                    if (k >= FW) continue;
                    const double f = h ? f2.y : f2.x;
#pragma unroll
                    for (int r = 0; r < RX; ++r) {           // output r, tap k -> u = r + FW-1-k
                        acc[ry][r][0] = fma(f, v[r + FW - 1 - k].x, acc[ry][r][0]);
                        acc[ry][r][1] = fma(f, v[r + FW - 1 - k].y, acc[ry][r][1]);
                    }
                }
            }
        }
    }

    const int z = z0 + 2 * zq;
    double chi = 0.0;
#pragma unroll
    for (int ry = 0; ry < RY; ++ry) {
    const int gy = ty0 + oy * RY + ry;
    if (gy < H && z < Dp) {
        const T* data = (const T*)pb.data + (size_t)cube * H * W * Dp;
        const T* ivc = pb.var_is_cube ? (const T*)pb.iv + (size_t)cube * H * W * Dp : nullptr;
        const double ivs = pb.var_is_cube ? 0.0 : pb.iv_scalar[cube];
        T* err = (T*)pb.err + (size_t)chain * H * W * Dp;
        typedef typename Pair<T>::P P2;
#pragma unroll
        for (int r = 0; r < RX; ++r) {
            const int gx = tx0 + xb * RX + r;
            if (gx >= W) continue;
            const size_t off = ((size_t)gy * W + gx) * Dp + z;
            if (sim_out) {
                if (z < D) sim_out[(size_t)chain * D * H * W + ((size_t)z * H + gy) * W + gx] = acc[ry][r][0];
                if (z + 1 < D) sim_out[(size_t)chain * D * H * W + ((size_t)(z + 1) * H + gy) * W + gx] = acc[ry][r][1];
            }
            if (write_err || chi2_out) {
                const P2 d = *(const P2*)(data + off);
                double e0 = z < D ? (double)d.x - acc[ry][r][0] : 0.0;
                double e1 = z + 1 < D ? (double)d.y - acc[ry][r][1] : 0.0;
                if (write_err) { P2 o; o.x = (T)e0; o.y = (T)e1; *(P2*)(err + off) = o; }
                if (chi2_out) {
                    double w0 = ivs, w1 = ivs;
                    if (ivc) { const P2 w = *(const P2*)(ivc + off); w0 = (double)w.x; w1 = (double)w.y; }
                    chi = fma(e0 * e0, w0, chi);
                    chi = fma(e1 * e1, w1, chi);
                }
            }
        }
    }
    }
    if (chi2_out) {
        chi = warp_sum(chi);
        if ((tid & 31) == 0) atomicAdd(chi2_out + chain, 0.5 * chi);
    }
}

// ---------------------------------------------------------------------------
// Any FSF width (e.g. 23, 31, 41 columns): the same register tiling with the FSF row walked in
// chunks of KC = 8 taps, so that the per-thread window stays at KC + RX - 1 vectors whatever
// the width.  CTA = 256 threads = 2 z-pairs x 8 x-blocks x 16 rows (tile 16 x 32 spaxels x 4
// channels: the halo tile of a 41 x 41 FSF is 56 x 72 x 4 doubles = 129 KB).  The tile row
// stride is padded to 1 (mod 4) positions: the 8 lanes of a quarter warp (2 z-pairs x 4 rows)
// then hit 8 different 16-byte slots of a 128-byte bank line -- conflict-free LDS.128.
// ---------------------------------------------------------------------------
__host__ __device__ inline int wide_row_positions(int fw) {
    int hx = 32 + fw - 1;
    while ((hx & 3) != 1) ++hx;
    return hx;
}

template <typename T>
__global__ void __launch_bounds__(256, 1)
stencil_wide_kernel(const __grid_constant__ Problem pb, const double* __restrict__ lines,
                    double* sim_out, int write_err, double* chi2_out) {
    const int TY = 16, TX = 32, RX = 4, ZC = 4, KC = 8;
    extern __shared__ double smem_raw[];
    const int fh = pb.fh, fw = pb.fw, Dp = pb.Dp, D = pb.D, H = pb.H, W = pb.W;
    const int hx = TX + fw - 1, hy = TY + fh - 1, hxp = wide_row_positions(fw);
    double* F = smem_raw;                                   // [fh][fw]
    double* tile = F + ((fh * fw + 1) & ~1);                // [hy][hxp][ZC], 16-byte aligned
    const int ty_n = (H + TY - 1) / TY, tx_n = (W + TX - 1) / TX;
    const int chain = blockIdx.x / (ty_n * tx_n);
    const int trem = blockIdx.x - chain * ty_n * tx_n;
    const int ty0 = (trem / tx_n) * TY, tx0 = (trem % tx_n) * TX;
    const int z0 = blockIdx.y * ZC;
    const int cube = chain / pb.chains_per_cube;
    const int tid = threadIdx.x;

    for (int i = tid; i < fh * fw; i += 256) F[i] = pb.fsf[i];
    const double* lc = lines + (size_t)chain * H * W * Dp;
    for (int i = tid; i < hy * hx * (ZC / 2); i += 256) {    // double2 granularity
        const int zq = i & 1, s = i >> 1;
        const int sy = s / hx, sx = s - sy * hx;
        const int gy = ty0 + sy - pb.fhh, gx = tx0 + sx - pb.fhw;
        const int z = z0 + 2 * zq;
        double* dst = tile + ((size_t)sy * hxp + sx) * ZC + 2 * zq;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W && z < Dp) {      // cp.async staging, as in the tiled kernel
            const unsigned sa = (unsigned)__cvta_generic_to_shared(dst);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(lc + ((size_t)gy * W + gx) * Dp + z) : "memory");
        } else {
            *(double2*)dst = make_double2(0.0, 0.0);
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    const int zq = tid & 1, oy = ((tid >> 6) << 2) | ((tid >> 1) & 3), xb = (tid >> 3) & 7;
    double acc[RX][2];
#pragma unroll
    for (int r = 0; r < RX; ++r) { acc[r][0] = 0.0; acc[r][1] = 0.0; }
    for (int j = 0; j < fh; ++j) {
        const double* trow = tile + ((size_t)(oy + fh - 1 - j) * hxp + xb * RX) * ZC + 2 * zq;
        const double* frow = F + j * fw;
        for (int k0 = 0; k0 < fw; k0 += KC) {
            // taps k0 .. k0+KC-1 (those >= fw count as zero); output r, tap k reads window
            // position r + fw-1-k = base + (r + KC-1 - (k - k0)) with base = fw - k0 - KC
            const int base = fw - k0 - KC;
            double2 v[KC + RX - 1];
#pragma unroll
            for (int u = 0; u < KC + RX - 1; ++u)
                v[u] = base + u >= 0 ? *(const double2*)(trow + (size_t)(base + u) * ZC) : make_double2(0.0, 0.0);
#pragma unroll
            for (int kk = 0; kk < KC; ++kk) {
                const double f = k0 + kk < fw ? frow[k0 + kk] : 0.0;
#pragma unroll
                for (int r = 0; r < RX; ++r) {
                    acc[r][0] = fma(f, v[r + KC - 1 - kk].x, acc[r][0]);
                    acc[r][1] = fma(f, v[r + KC - 1 - kk].y, acc[r][1]);
                }
            }
        }
    }

    const int gy = ty0 + oy;
    const int z = z0 + 2 * zq;
    double chi = 0.0;
    if (gy < H && z < Dp) {
        const T* data = (const T*)pb.data + (size_t)cube * H * W * Dp;
        const T* ivc = pb.var_is_cube ? (const T*)pb.iv + (size_t)cube * H * W * Dp : nullptr;
        const double ivs = pb.var_is_cube ? 0.0 : pb.iv_scalar[cube];
        T* err = (T*)pb.err + (size_t)chain * H * W * Dp;
        typedef typename Pair<T>::P P2;
#pragma unroll
        for (int r = 0; r < RX; ++r) {
            const int gx = tx0 + xb * RX + r;
            if (gx >= W) continue;
            const size_t off = ((size_t)gy * W + gx) * Dp + z;
            if (sim_out) {
                if (z < D) sim_out[(size_t)chain * D * H * W + ((size_t)z * H + gy) * W + gx] = acc[r][0];
                if (z + 1 < D) sim_out[(size_t)chain * D * H * W + ((size_t)(z + 1) * H + gy) * W + gx] = acc[r][1];
            }
            if (write_err || chi2_out) {
                const P2 d = *(const P2*)(data + off);
                double e0 = z < D ? (double)d.x - acc[r][0] : 0.0;
                double e1 = z + 1 < D ? (double)d.y - acc[r][1] : 0.0;
                if (write_err) { P2 o; o.x = (T)e0; o.y = (T)e1; *(P2*)(err + off) = o; }
                if (chi2_out) {
                    double w0 = ivs, w1 = ivs;
                    if (ivc) { const P2 w = *(const P2*)(ivc + off); w0 = (double)w.x; w1 = (double)w.y; }
                    chi = fma(e0 * e0, w0, chi);
                    chi = fma(e1 * e1, w1, chi);
                }
            }
        }
    }
    if (chi2_out) {
        chi = warp_sum(chi);
        if ((tid & 31) == 0) atomicAdd(chi2_out + chain, 0.5 * chi);
    }
}

}  // namespace d3d
