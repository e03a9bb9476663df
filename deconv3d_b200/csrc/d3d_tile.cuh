// d3d_tile.cuh -- ONE oversized cube spread over several contexts / GPUs.
//
// The colour lattice (y mod fh, x mod fw) is global: the sites of one class have pairwise
// disjoint FSF windows wherever they lie, so every context may update the sites of the class
// that fall inside ITS tile concurrently.  What has to travel after a colour phase is only
// the OUTCOME of each site update -- (site, chain, a, c, w, delta-logL, accepted): 64 bytes --
// because the residual change is the rank-1 field F (x) (a_old L(c_old, w_old) - a L(c, w)),
// which the receiver rebuilds from the record and its own copy of the old parameters and adds
// to the part of its residual it will read again (tile grown by the FSF half-size).  The
// exchange itself (NCCL all-gather / peer copies) is the caller's: see deconv3d_b200/dist.py.
#pragma once

namespace d3d {

enum { REC_SITE = 0, REC_CHAIN, REC_A, REC_C, REC_W, REC_LIK, REC_ACC, REC_PAD, REC_N };

// One thread per (chain, lattice slot) of colour class (cy, cx): the record of the site if
// this context owns it (and the chain is running), site = -1 otherwise.
__global__ void pack_records_kernel(const __grid_constant__ Problem pb, int cy, int cx, int nly, int nlx,
                                    double* rec) {
    const int nl = nly * nlx;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pb.n_chains * nl) return;
    const int chain = i / nl, slot = i - chain * nl;
    const int iy = slot / nlx, ix = slot - iy * nlx;
    const int y = cy + iy * pb.fh, x = cx + ix * pb.fw;
    double* r = rec + (size_t)i * REC_N;
    const int cube = chain / pb.chains_per_cube;
    const size_t HW = (size_t)pb.H * pb.W;
    bool live = y < pb.H && x < pb.W && y >= pb.ty0 && y < pb.ty1 && x >= pb.tx0 && x < pb.tx1 &&
                pb.active[chain];
    const int site = y * pb.W + x;
    if (live) live = pb.mask[(size_t)cube * HW + site] == 1;
    if (!live) {
        r[REC_SITE] = -1.0;
#pragma unroll
        for (int k = 1; k < REC_N; ++k) r[k] = 0.0;
        return;
    }
    const double* p = pb.params + ((size_t)chain * HW + site) * 3;
    r[REC_SITE] = (double)site;
    r[REC_CHAIN] = (double)chain;
    r[REC_A] = p[0]; r[REC_C] = p[1]; r[REC_W] = p[2];
    r[REC_LIK] = pb.lik_cur[(size_t)chain * HW + site];
    r[REC_ACC] = (double)pb.acc_cur[(size_t)chain * HW + site];
    r[REC_PAD] = 0.0;
}

// One CTA per record.  Records of sites this context owns (already applied by its own sweep
// kernel) and empty slots are skipped; the others update parameters, likelihood, accept
// counter and the residual inside the region.
template <typename T>
__global__ void apply_records_kernel(const __grid_constant__ Problem pb, const double* rec, int n_rec) {
    typedef typename Vec<T>::V V;
    const int VEC = Vec<T>::N;
    extern __shared__ double smem_raw[];
    const double* r = rec + (size_t)blockIdx.x * REC_N;
    const int site = (int)r[REC_SITE];
    if (site < 0) return;
    const int chain = (int)r[REC_CHAIN];
    const int W = pb.W, H = pb.H, Dp = pb.Dp;
    const int y = site / W, x = site - y * W;
    if (y >= pb.ty0 && y < pb.ty1 && x >= pb.tx0 && x < pb.tx1) return;      // mine

    Smem sm;
    carve(sm, smem_raw, pb);
    load_constants(sm, pb);
    const size_t HW = (size_t)H * W;
    double* prm = pb.params + ((size_t)chain * HW + site) * 3;
    const double a_o = prm[0], c_o = prm[1], w_o = prm[2];
    const double a_n = r[REC_A], c_n = r[REC_C], w_n = r[REC_W];
    __syncthreads();                                       // constants loaded, old parameters read

    // window clipped to the field, then to the region this context keeps valid
    const int y0 = max(max(y - pb.fhh, 0), pb.ry0), y1 = min(min(y + pb.fhh + 1, H), pb.ry1);
    const int x0 = max(max(x - pb.fhw, 0), pb.rx0), x1 = min(min(x + pb.fhw + 1, W), pb.rx1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (y0 < y1 && x0 < x1) {
        if (warp == 0) warp_line_profile(pb, sm, c_o, w_o, sm.g_o, sm.Lu_o, lane);
        else if (warp == 1) warp_line_profile(pb, sm, c_n, w_n, sm.g_n, sm.Lu_n, lane);
        __syncthreads();
        const int ww = x1 - x0, npos = (y1 - y0) * ww;
        const int oy = y0 - (y - pb.fhh), ox = x0 - (x - pb.fhw);
        const int ZL = Dp / VEC, NC = blockDim.x / ZL;
        const int col = tid / ZL, zp = tid - col * ZL;
        if (col < NC) {
            double coef[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v)
                coef[v] = upd_coef(a_o, sm.Lu_o[zp * VEC + v], a_n, sm.Lu_n[zp * VEC + v]);
            T* err = (T*)pb.err + (size_t)chain * HW * Dp;
            for (int q = col; q < npos; q += NC) {
                const int dy = q / ww, dx = q - dy * ww;
                const size_t off = ((size_t)(y0 + dy) * W + (x0 + dx)) * Dp + zp * VEC;
                const double f = sm.F[(oy + dy) * pb.fw + ox + dx];
                double e[VEC];
                unpack(*(const V*)(err + off), e);
#pragma unroll
                for (int v = 0; v < VEC; ++v) e[v] = fma(f, coef[v], e[v]);
                V o;
                pack(o, e);
                *(V*)(err + off) = o;
            }
        }
    }
    if (tid == 0) {
        prm[0] = a_n; prm[1] = c_n; prm[2] = w_n;
        pb.lik_cur[(size_t)chain * HW + site] = r[REC_LIK];
        const int acc = r[REC_ACC] != 0.0;
        pb.acc_cur[(size_t)chain * HW + site] = (uint8_t)acc;
        if (acc) atomicAdd((unsigned long long*)&pb.accepted[chain], 1ull);
    }
}

}  // namespace d3d
