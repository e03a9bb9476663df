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
#include <cooperative_groups.h>

namespace d3d {
namespace cg = cooperative_groups;

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
            const int stepy = NC / ww, stepx = NC - stepy * ww;
            int dy = col / ww, dx = col - dy * ww;          // running position, advanced without divisions
            for (int q = col; q < npos; q += NC, dx += stepx, dy += stepy) {
                if (dx >= ww) { dx -= ww; ++dy; }
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

// ---------------------------------------------------------------------------
// Large windows (generic FSF sizes, e.g. 41x41x64 = 860 KB per array at cfg4): ONE site is
// worked by a thread-block CLUSTER.  Each CTA of the cluster takes a contiguous share of the
// window positions, the six partial sums meet in the leader's shared memory (DSMEM), the leader
// decides, every CTA reads the decision back through DSMEM and updates its share.  All CTAs
// evaluate the proposal and both line profiles redundantly (deterministic, D values).
// grid = (lattice slots * cluster size, chains), cluster = (CS, 1, 1).
// ---------------------------------------------------------------------------
// One site worked by the calling cluster (constants already in shared memory, every CTA of the
// cluster running).  8 window warps + 2 scalar warps per CTA.
template <typename T, bool IVCUBE>
__device__ __forceinline__ void cluster_site_update(const Problem& pb, const Smem& sm, cg::cluster_group& cluster,
                                                    int CS, int cr, int chain, int cube, int site, long long it,
                                                    double* crow, double* lrow) {
    typedef typename Vec<T>::V V;
    const int VEC = Vec<T>::N;
    const int W = pb.W, H = pb.H, Dp = pb.Dp;
    const int y = site / W, x = site - y * W;
    const size_t HW = (size_t)H * W;
    EvalReq ev; ev.enabled = 0; ev.out = nullptr;
    // 8 window warps + 2 scalar warps (proposal / new profile / decision, old profile): the
    // window loads start at once instead of waiting behind the transcendental chains
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NWW = 8, nwarps = NWW;
    Philox rng;
    Proposal prop;
    if (warp == NWW) {
        make_proposal(pb, chain, cube, site, (unsigned)it, ev, rng, prop);
        warp_line_profile(pb, sm, prop.c_new, prop.w_new, sm.g_n, sm.Lu_n, lane);
    } else if (warp == NWW + 1) {
        const double* prm = pb.params + ((size_t)chain * HW + site) * 3;
        warp_line_profile(pb, sm, prm[1], prm[2], sm.g_o, sm.Lu_o, lane);
    }

    const int y0 = max(y - pb.fhh, 0), y1 = min(y + pb.fhh + 1, H);
    const int x0 = max(x - pb.fhw, 0), x1 = min(x + pb.fhw + 1, W);
    const int ww = x1 - x0, npos = (y1 - y0) * ww;
    const int oy = y0 - (y - pb.fhh), ox = x0 - (x - pb.fhw);
    const int ZL = Dp / VEC, NC = (NWW * 32) / ZL;
    const int col = tid / ZL, zp = tid - col * ZL;
    const bool worker = col < NC;
    const int stepy = NC / ww, stepx = NC - stepy * ww;
    // this CTA's share of the positions
    const int share = (npos + CS - 1) / CS;
    const int q0 = min(cr * share, npos), q1 = min(q0 + share, npos);

    T* err = (T*)pb.err + (size_t)chain * HW * Dp;
    const T* ivc = IVCUBE ? (const T*)pb.iv + (size_t)cube * HW * Dp : nullptr;
    const double ivs = IVCUBE ? 0.0 : pb.iv_scalar[cube];

    double h[VEC], g[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) { h[v] = 0.0; g[v] = 0.0; }
    double f2 = 0.0;
    // everything above reads what no colour phase writes (constants, this site's own parameters,
    // the counter-based random stream); from here on the residual of the previous phase is needed
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (worker) {
        const int UN = 4;
        // (dy, dx) of the running position advance by NC without divisions
        int dy = (q0 + col) / ww, dx = (q0 + col) - dy * ww;
        for (int qb = q0 + col; qb < q1; qb += UN * NC) {
            V ev_[UN], wv_[UN];
            double f_[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int q = qb + u * NC;
                if (q < q1) {
                    const size_t off = ((size_t)(y0 + dy) * W + (x0 + dx)) * Dp + zp * VEC;
                    ev_[u] = *(const V*)(err + off);
                    if (IVCUBE) wv_[u] = *(const V*)(ivc + off);
                    f_[u] = sm.F[(oy + dy) * pb.fw + ox + dx];
                }
                dx += stepx; dy += stepy;
                if (dx >= ww) { dx -= ww; ++dy; }
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                if (qb + u * NC < q1) {
                    double e[VEC];
                    unpack(ev_[u], e);
                    const double f = f_[u];
                    if (IVCUBE) {
                        double w_[VEC];
                        unpack(wv_[u], w_);
                        const double ff = f * f;
#pragma unroll
                        for (int v = 0; v < VEC; ++v) {
                            h[v] = fma(f, w_[v] * e[v], h[v]);
                            g[v] = fma(ff, w_[v], g[v]);
                        }
                    } else {
#pragma unroll
                        for (int v = 0; v < VEC; ++v) h[v] = fma(f, e[v], h[v]);
                        f2 = fma(f, f, f2);
                    }
                }
            }
        }
        if (!IVCUBE) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) { h[v] *= ivs; g[v] = ivs * f2; }
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
    }
#pragma unroll
    for (int j = 0; j < R_N; ++j) {
        if (j == R_A) continue;
        const double s = warp_sum(part[j]);
        if (lane == 0 && warp < NWW) sm.red[warp * 8 + j] = s;
    }
    __syncthreads();
    // CTA totals -> slot `cr` of the leader's exchange area (second half of sm.red)
    if (warp == NWW) {
        double* leader_x = cluster.map_shared_rank(sm.red + 128, 0);
#pragma unroll
        for (int j = 0; j < R_N; ++j) {
            if (j == R_A) continue;
            const double s = warp_sum(lane < nwarps ? sm.red[lane * 8 + j] : 0.0);
            if (lane == 0) leader_x[cr * 8 + j] = s;
        }
    }
    cluster.sync();
    int accepted = 0;
    if (cr == 0 && warp == NWW) {
        double tot[R_N];
#pragma unroll
        for (int j = 0; j < R_N; ++j) {
            if (j == R_A) { tot[j] = 0.0; continue; }
            tot[j] = warp_sum(lane < CS ? sm.red[128 + lane * 8 + j] : 0.0);
        }
        accepted = decide(pb, sm, chain, cube, site, prop, tot, rng, crow, lrow, ev, lane);
        if (lane == 0) {
            if (accepted) atomicAdd((unsigned long long*)&pb.accepted[chain], 1ull);
            if (pb.acc_cur) pb.acc_cur[(size_t)chain * HW + site] = (uint8_t)accepted;
        }
    }
    cluster.sync();                                    // decision in the leader's sm.bc
    const double* bc = cluster.map_shared_rank(sm.bc, 0);
    const int acc = bc[0] != 0.0;
    const double r = bc[1], a = bc[3];
    if (worker) {
        double coef[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) coef[v] = upd_coef(a, lo_v[v], r, acc ? ln_v[v] : lo_v[v]);
        const int UN = 4;
        int dy = (q0 + col) / ww, dx = (q0 + col) - dy * ww;
        for (int qb = q0 + col; qb < q1; qb += UN * NC) {
            V ev_[UN];
            double f_[UN];
            size_t off_[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int q = qb + u * NC;
                if (q < q1) {
                    off_[u] = ((size_t)(y0 + dy) * W + (x0 + dx)) * Dp + zp * VEC;
                    ev_[u] = *(const V*)(err + off_[u]);
                    f_[u] = sm.F[(oy + dy) * pb.fw + ox + dx];
                }
                dx += stepx; dy += stepy;
                if (dx >= ww) { dx -= ww; ++dy; }
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                if (qb + u * NC < q1) {
                    double e[VEC];
                    unpack(ev_[u], e);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) e[v] = fma(f_[u], coef[v], e[v]);
                    V o;
                    pack(o, e);
                    *(V*)(err + off_[u]) = o;
                }
            }
        }
    }
    cluster.sync();                                    // nobody leaves while its sm.bc is read
}

template <typename T, bool IVCUBE>
__global__ void __launch_bounds__(320)
sweep_colour_cluster_kernel(const __grid_constant__ Problem pb, long long it, int cy, int cx, int nlx,
                            double* crow_base, double* lrow_base, long long rows_local,
                            long long row_local) {
    // (programmatic dependent launch, see launch_colour_class: the next phase may be scheduled as soon
    // as every CTA of this one runs; it waits in cluster_site_update before it touches the residual)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = (int)cluster.num_blocks(), cr = (int)cluster.block_rank();
    extern __shared__ double smem_raw[];
    Smem sm;
    carve(sm, smem_raw, pb);
    const int chain = blockIdx.y;
    const int cube = chain / pb.chains_per_cube;
    const int slot = blockIdx.x / CS;
    const int iy = slot / nlx, ix = slot - iy * nlx;
    const int y = cy + iy * pb.fh, x = cx + ix * pb.fw;
    // (every exit below is taken by the whole cluster)
    if (y >= pb.H || x >= pb.W) return;
    if (y < pb.ty0 || y >= pb.ty1 || x < pb.tx0 || x >= pb.tx1) return;
    if (!pb.active[chain]) return;
    const int site = y * pb.W + x;
    const size_t HW = (size_t)pb.H * pb.W;
    if (pb.mask[(size_t)cube * HW + site] != 1) return;
    load_constants(sm, pb);
    __syncthreads();
    cluster.sync();                                    // every CTA of the cluster is running (DSMEM)
    double* crow = crow_base ? crow_base + (((size_t)chain * rows_local + row_local) * HW + site) * 3 : nullptr;
    double* lrow = lrow_base ? lrow_base + ((size_t)chain * rows_local + row_local) * HW + site : nullptr;
    cluster_site_update<T, IVCUBE>(pb, sm, cluster, CS, cr, chain, cube, site, it, crow, lrow);
}

// Remote records on large windows: a cluster per record, every CTA rebuilds both profiles and
// adds its share of the rank-1 field.
template <typename T>
__global__ void __launch_bounds__(256)
apply_records_cluster_kernel(const __grid_constant__ Problem pb, const double* rec, int n_rec) {
    typedef typename Vec<T>::V V;
    const int VEC = Vec<T>::N;
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = (int)cluster.num_blocks(), cr = (int)cluster.block_rank();
    extern __shared__ double smem_raw[];
    const double* r = rec + (size_t)(blockIdx.x / CS) * REC_N;
    const int site = (int)r[REC_SITE];
    if (site < 0) return;
    const int chain = (int)r[REC_CHAIN];
    const int W = pb.W, H = pb.H, Dp = pb.Dp;
    const int y = site / W, x = site - y * W;
    if (y >= pb.ty0 && y < pb.ty1 && x >= pb.tx0 && x < pb.tx1) return;      // mine
    const size_t HW = (size_t)H * W;
    double* prm = pb.params + ((size_t)chain * HW + site) * 3;
    const double a_n = r[REC_A], c_n = r[REC_C], w_n = r[REC_W];
    const int y0 = max(max(y - pb.fhh, 0), pb.ry0), y1 = min(min(y + pb.fhh + 1, H), pb.ry1);
    const int x0 = max(max(x - pb.fhw, 0), pb.rx0), x1 = min(min(x + pb.fhw + 1, W), pb.rx1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (!(y0 < y1 && x0 < x1)) {
        // window outside the region this context keeps valid: book-keeping only (whole cluster)
        if (cr == 0 && tid == 0) {
            prm[0] = a_n; prm[1] = c_n; prm[2] = w_n;
            pb.lik_cur[(size_t)chain * HW + site] = r[REC_LIK];
            const int acc = r[REC_ACC] != 0.0;
            pb.acc_cur[(size_t)chain * HW + site] = (uint8_t)acc;
            if (acc) atomicAdd((unsigned long long*)&pb.accepted[chain], 1ull);
        }
        return;
    }
    Smem sm;
    carve(sm, smem_raw, pb);
    load_constants(sm, pb);
    const double a_o = prm[0], c_o = prm[1], w_o = prm[2];
    __syncthreads();
    cluster.sync();                                        // every CTA has read the old parameters
    {
        if (warp == 0) warp_line_profile(pb, sm, c_o, w_o, sm.g_o, sm.Lu_o, lane);
        else if (warp == 1) warp_line_profile(pb, sm, c_n, w_n, sm.g_n, sm.Lu_n, lane);
        __syncthreads();
        const int ww = x1 - x0, npos = (y1 - y0) * ww;
        const int oy = y0 - (y - pb.fhh), ox = x0 - (x - pb.fhw);
        const int ZL = Dp / VEC, NC = blockDim.x / ZL;
        const int col = tid / ZL, zp = tid - col * ZL;
        const int share = (npos + CS - 1) / CS;
        const int q0 = min(cr * share, npos), q1 = min(q0 + share, npos);
        if (col < NC) {
            double coef[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v)
                coef[v] = upd_coef(a_o, sm.Lu_o[zp * VEC + v], a_n, sm.Lu_n[zp * VEC + v]);
            T* err = (T*)pb.err + (size_t)chain * HW * Dp;
            const int stepy = NC / ww, stepx = NC - stepy * ww;
            int dy = (q0 + col) / ww, dx = (q0 + col) - dy * ww;
            for (int q = q0 + col; q < q1; q += NC, dx += stepx, dy += stepy) {
                if (dx >= ww) { dx -= ww; ++dy; }
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
    if (cr == 0 && tid == 0) {
        prm[0] = a_n; prm[1] = c_n; prm[2] = w_n;
        pb.lik_cur[(size_t)chain * HW + site] = r[REC_LIK];
        const int acc = r[REC_ACC] != 0.0;
        pb.acc_cur[(size_t)chain * HW + site] = (uint8_t)acc;
        if (acc) atomicAdd((unsigned long long*)&pb.accepted[chain], 1ull);
    }
}

// ---------------------------------------------------------------------------
// Fused exchange over peer memory (NVLink P2P / same device): no collective library and no host
// round trip inside a phase.  Every context owns a "box": flags[MAXW] (uint64, monotonic phase
// counters, one per source tile) followed by inbox[2][n_tiles][slots][REC_N] doubles (double
// buffered by phase parity).  After its phase kernel, push_records_kernel of tile t stores its
// records into slot t of EVERY box (its own included) and then -- last block, after a
// system-scope fence -- publishes flags[t] = phase + 1 in every box.  The applier of the same
// phase spins (bounded) on the flag of a record's source tile before it reads the record from
// its own box with L2-only loads.  Two phases may overlap across GPUs, never three: a tile can
// only start phase p+2 after its applier of p+1 saw every flag p+1, which the others publish
// after their appliers of phase p.
// ---------------------------------------------------------------------------
enum { TILE_MAXW = 16 };

struct TileBox {
    int n_tiles, my_tile;
    long long slots;                        // records per tile and phase (n_chains * lattice slots)
    unsigned long long* flags[TILE_MAXW];   // flags array of every box (peer pointers)
    double* inbox[TILE_MAXW];               // inbox of every box
    unsigned int* done_counter;             // local: blocks of push_records_kernel that finished
    long long timeout_cycles;               // bound of the flag wait (D3D_TILE_TIMEOUT_S, default 30 s)
};

__device__ __forceinline__ size_t tile_inbox_index(const TileBox& tb, int parity, int src, long long slot) {
    return (((size_t)parity * tb.n_tiles + src) * tb.slots + slot) * REC_N;
}

__global__ void push_records_kernel(const __grid_constant__ Problem pb, const __grid_constant__ TileBox tb,
                                    int cy, int cx, int nly, int nlx, unsigned long long phase) {
    const int nl = nly * nlx;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < pb.n_chains * nl) {
        const int chain = i / nl, slot = i - chain * nl;
        const int iy = slot / nlx, ix = slot - iy * nlx;
        const int y = cy + iy * pb.fh, x = cx + ix * pb.fw;
        const int cube = chain / pb.chains_per_cube;
        const size_t HW = (size_t)pb.H * pb.W;
        bool live = y < pb.H && x < pb.W && y >= pb.ty0 && y < pb.ty1 && x >= pb.tx0 && x < pb.tx1 &&
                    pb.active[chain];
        const int site = y * pb.W + x;
        if (live) live = pb.mask[(size_t)cube * HW + site] == 1;
        double r[REC_N];
#pragma unroll
        for (int k = 0; k < REC_N; ++k) r[k] = 0.0;
        r[REC_SITE] = -1.0;
        if (live) {
            const double* p = pb.params + ((size_t)chain * HW + site) * 3;
            r[REC_SITE] = (double)site; r[REC_CHAIN] = (double)chain;
            r[REC_A] = p[0]; r[REC_C] = p[1]; r[REC_W] = p[2];
            r[REC_LIK] = pb.lik_cur[(size_t)chain * HW + site];
            r[REC_ACC] = (double)pb.acc_cur[(size_t)chain * HW + site];
        }
        const size_t at = tile_inbox_index(tb, (int)(phase & 1ull), tb.my_tile, i);
        for (int t = 0; t < tb.n_tiles; ++t) {
            if (t == tb.my_tile) continue;              // own records are never read back
            double2* dst = (double2*)(tb.inbox[t] + at);
#pragma unroll
            for (int k = 0; k < REC_N / 2; ++k) dst[k] = make_double2(r[2 * k], r[2 * k + 1]);
        }
    }
    // publish: the last block to get here raises this tile's flag in every box
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(tb.done_counter, 1u) + 1u;
        if (done == gridDim.x) {
            *tb.done_counter = 0u;
            __threadfence_system();
            for (int t = 0; t < tb.n_tiles; ++t)
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(tb.flags[t] + tb.my_tile), "l"(phase + 1ull) : "memory");
        }
    }
}

// Waits (bounded: tb.timeout_cycles) until source tile `src` has published `phase`; returns false on timeout.
__device__ __forceinline__ bool tile_wait_flag(const TileBox& tb, int src, unsigned long long phase) {
    const unsigned long long* f = tb.flags[tb.my_tile] + src;
    const long long t0 = clock64();
    for (;;) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
        if (v >= phase + 1ull) return true;
        if (clock64() - t0 > tb.timeout_cycles) return false;
        __nanosleep(200);
    }
}

// Applier of the fused exchange: grid = n_tiles * slots clusters (or CTAs); record (src, slot) is
// read from this context's own box once the source tile has published the phase.
// Triage of the fused exchange: ONE THREAD per (source tile, slot) waits for the source's flag, reads
// the record and does the book-keeping of the remote site (parameters, likelihood, accept flag and
// count); the few records whose window reaches into the region this context keeps valid are appended
// to hits[parity][..] for the cluster applier.  Without it the applier's grid is a cluster per slot
// of every tile (10 000 CTAs at 1024 x 1024 on four GPUs), each of which waits for a flag and passes
// two cluster barriers before all but a few dozen find nothing to do.  Records of one phase have
// disjoint windows: the order of the list does not matter.
__global__ void triage_box_kernel(const __grid_constant__ Problem pb, const __grid_constant__ TileBox tb,
                                  unsigned long long phase, int* hits, unsigned int* hit_count,
                                  long long hits_stride, unsigned int max_hits) {
    const int par = (int)(phase & 1ull);
    if (blockIdx.x == 0 && threadIdx.x == 0) hit_count[par ^ 1] = 0u;   // the next phase's counter
    const long long ridx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ridx >= (long long)tb.n_tiles * tb.slots) return;
    const int src = (int)(ridx / tb.slots);
    if (src == tb.my_tile) return;
    if (*(volatile int*)pb.status == 2 || !tile_wait_flag(tb, src, phase)) { atomicExch(pb.status, 2); return; }
    const double* rp = tb.inbox[tb.my_tile] + tile_inbox_index(tb, par, src, ridx - (long long)src * tb.slots);
    double r[REC_N];
#pragma unroll
    for (int k = 0; k < REC_N; ++k) r[k] = __ldcg(rp + k);
    const int site = (int)r[REC_SITE];
    if (site < 0) return;
    const int chain = (int)r[REC_CHAIN];
    const int W = pb.W, H = pb.H;
    const int y = site / W, x = site - y * W;
    const int y0 = max(max(y - pb.fhh, 0), pb.ry0), y1 = min(min(y + pb.fhh + 1, H), pb.ry1);
    const int x0 = max(max(x - pb.fhw, 0), pb.rx0), x1 = min(min(x + pb.fhw + 1, W), pb.rx1);
    if (y0 < y1 && x0 < x1) {                                // the applier does its book-keeping too
        const unsigned int k = atomicAdd(hit_count + par, 1u);
        if (k < max_hits) hits[(size_t)par * hits_stride + k] = (int)ridx;
        else atomicExch(pb.status, 2);                       // (cannot happen: max_hits bounds the geometry)
        return;
    }
    const size_t HW = (size_t)H * W;
    double* prm = pb.params + ((size_t)chain * HW + site) * 3;
    prm[0] = r[REC_A]; prm[1] = r[REC_C]; prm[2] = r[REC_W];
    pb.lik_cur[(size_t)chain * HW + site] = r[REC_LIK];
    const int acc = r[REC_ACC] != 0.0;
    pb.acc_cur[(size_t)chain * HW + site] = (uint8_t)acc;
    if (acc) atomicAdd((unsigned long long*)&pb.accepted[chain], 1ull);
}

template <typename T, bool CLUSTER>
__global__ void __launch_bounds__(256)
apply_box_kernel(const __grid_constant__ Problem pb, const __grid_constant__ TileBox tb, unsigned long long phase,
                 const int* hits = nullptr, const unsigned int* hit_count = nullptr, long long hits_stride = 0) {
    typedef typename Vec<T>::V V;
    const int VEC = Vec<T>::N;
    int CS = 1, cr = 0;
    if (CLUSTER) {
        cg::cluster_group cluster = cg::this_cluster();
        CS = (int)cluster.num_blocks(); cr = (int)cluster.block_rank();
    }
    extern __shared__ double smem_raw[];
    long long ridx = blockIdx.x / CS;
    if (hits) {
        // behind triage_box_kernel: cluster i takes entry i of this phase's list (flags already seen)
        const int par = (int)(phase & 1ull);
        if (ridx >= (long long)hit_count[par]) return;       // (whole cluster)
        ridx = hits[(size_t)par * hits_stride + ridx];
    }
    const int src = (int)(ridx / tb.slots);
    if (src == tb.my_tile) return;                           // (whole cluster)
    if (!hits) {
    __shared__ int s_ok;
    // (after one time-out every later applier gives up at once: a dead peer costs seconds, not
    // seconds per phase).  In a cluster the LEADER alone decides and shares its verdict through
    // DSMEM, so that either all CTAs of the cluster go on to the cluster barrier below or none.
    if (threadIdx.x == 0 && cr == 0)
        s_ok = (*(volatile int*)pb.status != 2 && tile_wait_flag(tb, src, phase)) ? 1 : 0;
    int ok;
    if (CLUSTER) {
        cg::cluster_group cluster = cg::this_cluster();
        cluster.sync();
        ok = *cluster.map_shared_rank(&s_ok, 0);
        cluster.sync();                                      // the leader's s_ok stays alive until read
    } else {
        __syncthreads();
        ok = s_ok;
    }
    if (!ok) { if (threadIdx.x == 0) atomicExch(pb.status, 2); return; }
    }
    const double* rp = tb.inbox[tb.my_tile] + tile_inbox_index(tb, (int)(phase & 1ull), src, ridx - (long long)src * tb.slots);
    double r[REC_N];
#pragma unroll
    for (int k = 0; k < REC_N; ++k) r[k] = __ldcg(rp + k);
    const int site = (int)r[REC_SITE];
    if (site < 0) return;
    const int chain = (int)r[REC_CHAIN];
    const int W = pb.W, H = pb.H, Dp = pb.Dp;
    const int y = site / W, x = site - y * W;
    const size_t HW = (size_t)H * W;
    double* prm = pb.params + ((size_t)chain * HW + site) * 3;
    const double a_n = r[REC_A], c_n = r[REC_C], w_n = r[REC_W];
    const int y0 = max(max(y - pb.fhh, 0), pb.ry0), y1 = min(min(y + pb.fhh + 1, H), pb.ry1);
    const int x0 = max(max(x - pb.fhw, 0), pb.rx0), x1 = min(min(x + pb.fhw + 1, W), pb.rx1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (!(y0 < y1 && x0 < x1)) {
        // The window of this record misses the region this context keeps valid -- most records of a
        // big field: only the book-keeping of the site, no constants, no profiles, no barrier (the
        // test is the same for every CTA of the cluster).
        if (cr == 0 && tid == 0) {
            prm[0] = a_n; prm[1] = c_n; prm[2] = w_n;
            pb.lik_cur[(size_t)chain * HW + site] = r[REC_LIK];
            const int acc = r[REC_ACC] != 0.0;
            pb.acc_cur[(size_t)chain * HW + site] = (uint8_t)acc;
            if (acc) atomicAdd((unsigned long long*)&pb.accepted[chain], 1ull);
        }
        return;
    }
    Smem sm;
    carve(sm, smem_raw, pb);
    load_constants(sm, pb);
    const double a_o = prm[0], c_o = prm[1], w_o = prm[2];
    __syncthreads();
    if (CLUSTER) cg::this_cluster().sync();                 // every CTA has read the old parameters
    {
        if (warp == 0) warp_line_profile(pb, sm, c_o, w_o, sm.g_o, sm.Lu_o, lane);
        else if (warp == 1) warp_line_profile(pb, sm, c_n, w_n, sm.g_n, sm.Lu_n, lane);
        __syncthreads();
        const int ww = x1 - x0, npos = (y1 - y0) * ww;
        const int oy = y0 - (y - pb.fhh), ox = x0 - (x - pb.fhw);
        const int ZL = Dp / VEC, NC = blockDim.x / ZL;
        const int col = tid / ZL, zp = tid - col * ZL;
        const int share = (npos + CS - 1) / CS;
        const int q0 = min(cr * share, npos), q1 = min(q0 + share, npos);
        if (col < NC) {
            double coef[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v)
                coef[v] = upd_coef(a_o, sm.Lu_o[zp * VEC + v], a_n, sm.Lu_n[zp * VEC + v]);
            T* err = (T*)pb.err + (size_t)chain * HW * Dp;
            const int stepy = NC / ww, stepx = NC - stepy * ww;
            int dy = (q0 + col) / ww, dx = (q0 + col) - dy * ww;
            for (int q = q0 + col; q < q1; q += NC, dx += stepx, dy += stepy) {
                if (dx >= ww) { dx -= ww; ++dy; }
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
    if (cr == 0 && tid == 0) {
        prm[0] = a_n; prm[1] = c_n; prm[2] = w_n;
        pb.lik_cur[(size_t)chain * HW + site] = r[REC_LIK];
        const int acc = r[REC_ACC] != 0.0;
        pb.acc_cur[(size_t)chain * HW + site] = (uint8_t)acc;
        if (acc) atomicAdd((unsigned long long*)&pb.accepted[chain], 1ull);
    }
}

// Posterior summary on the device (lib/run.py:581-593): mean over rows [first_row, n_rows) of a
// chain buffer [n_chains][n_rows][HW*3]; one thread per (chain, element).
__global__ void chain_mean_kernel(const double* chain, int n_chains, long long n_rows, long long first_row,
                                  long long row_elems, double* mean) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n_chains * row_elems) return;
    const long long k = i / row_elems, j = i - k * row_elems;
    const double* p = chain + ((size_t)k * n_rows + first_row) * row_elems + j;
    double s = 0.0;
    for (long long r = first_row; r < n_rows; ++r, p += row_elems) s += *p;
    mean[i] = s / (double)(n_rows - first_row);
}

}  // namespace d3d
