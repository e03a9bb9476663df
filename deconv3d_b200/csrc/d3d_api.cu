// d3d_api.cu -- C ABI of libdeconv3d_b200.so (see include/deconv3d_b200.h).
//
// Host-side runtime of the likelihood hot path: device memory, layout
// conversion, kernel selection and launch, segmentation of the sweep at the
// residual-refresh points of lib/run.py:525-534, CUDA-event timing.
#include <cuda.h>      // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <math.h>
#include <cstring>
#include <vector>
#include <string>
#include <algorithm>
#include <stdlib.h>
#include <chrono>
#include <mutex>
#include <unordered_map>

#include "../../include/deconv3d_b200.h"
#include "d3d_kernels.cuh"

using namespace d3d;

// The pipelined sweep (d3d_pipe.cuh; measured in profiles/r02_notes.md) ships with BATCHED producers
// (D3D_PIPEM_*): two lock-stepped producer warps, each preparing 8 sites per pass, a ring of 32
// stages, look-ahead 2 -- 14 warps at cfg2.  Faster than one warp per site and role at every chain
// count.  The per-site producers (D3D_PIPEF_*) are kept for A/B builds only: -DD3D_PIPE_PERSITE
// compiles them and D3D_PIPE=3 selects them.  Every -D override below is for the A/B builds of
// profiles/tools/build_variant.sh.
#ifndef D3D_PIPE_LMAX
#define D3D_PIPE_LMAX 2         // the cross-term tables are built for this look-ahead
#endif
#ifndef D3D_PIPEF_L
#define D3D_PIPEF_L 2
#endif
#ifndef D3D_PIPEF_NA
#define D3D_PIPEF_NA 2
#endif
#ifndef D3D_PIPEF_NP
#define D3D_PIPEF_NP 2
#endif
#ifndef D3D_PIPEF_MAXT
#define D3D_PIPEF_MAXT 512      // 16 warps at cfg2: 128 registers per thread
#endif
#ifndef D3D_PIPEM_L
#define D3D_PIPEM_L 2
#endif
#ifndef D3D_PIPEM_NPW
#define D3D_PIPEM_NPW 2
#endif
#ifndef D3D_PIPEM_MAXT
#define D3D_PIPEM_MAXT 448      // 14 warps at cfg2: 128 registers per thread
#endif
#ifndef D3D_PIPE_NX
#define D3D_PIPE_NX 1
#endif


static thread_local std::string g_last_error;

static int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CK(call)                                                                        \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess)                                                         \
            return fail(e__ == cudaErrorMemoryAllocation ? D3D_ENOMEM : D3D_ECUDA,      \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),        \
                        __FILE__, __LINE__);                                            \
    } while (0)

// ------------------------------------------------------------------------------
// Device-memory cache.  cudaMalloc / cudaFree of the 10-100 MB state and staging buffers cost
// anything between 1 and 350 ms on the driver side (measured through Run(...), D3D_TIMING=1),
// which is as much as the sweeps of a short run.  Freed blocks are therefore kept (up to
// kPoolCap bytes per process) and handed out again to requests of a similar size; a free keeps
// cudaFree's contract (device idle afterwards).  Everything below allocates through dev_malloc /
// dev_free; the few plain cudaMalloc calls left are deliberate (memory exported by IPC handle).
// ------------------------------------------------------------------------------
namespace {
struct PoolBlock { void* p; size_t bytes; int dev; };
std::mutex g_pool_mu;
std::vector<PoolBlock> g_pool_free;
std::unordered_map<void*, PoolBlock> g_pool_live;
size_t g_pool_cached = 0;
const size_t kPoolCap = (size_t)4 << 30;

cudaError_t pool_release_all_locked() {
    for (auto& b : g_pool_free) { int cur; cudaGetDevice(&cur); if (cur != b.dev) cudaSetDevice(b.dev); cudaFree(b.p); if (cur != b.dev) cudaSetDevice(cur); }
    g_pool_free.clear();
    g_pool_cached = 0;
    return cudaSuccess;
}

cudaError_t pool_malloc(void** out, size_t bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t want = ((bytes ? bytes : 1) + 511) & ~(size_t)511;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    int best = -1;
    for (int i = 0; i < (int)g_pool_free.size(); ++i) {
        const PoolBlock& b = g_pool_free[i];
        if (b.dev != dev || b.bytes < want || b.bytes > want + want / 4 + ((size_t)1 << 20)) continue;
        if (best < 0 || b.bytes < g_pool_free[best].bytes) best = i;
    }
    PoolBlock blk;
    if (best >= 0) {
        blk = g_pool_free[best];
        g_pool_free.erase(g_pool_free.begin() + best);
        g_pool_cached -= blk.bytes;
    } else {
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {                    // give the cache back and try once more
            cudaGetLastError();
            pool_release_all_locked();
            e = cudaMalloc(&p, want);
            if (e != cudaSuccess) return e;
        }
        blk.p = p; blk.bytes = want; blk.dev = dev;
    }
    g_pool_live[blk.p] = blk;
    *out = blk.p;
    return cudaSuccess;
}

cudaError_t pool_free(void* p) {
    if (!p) return cudaSuccess;
    cudaDeviceSynchronize();                       // cudaFree's contract
    std::lock_guard<std::mutex> lk(g_pool_mu);
    auto it = g_pool_live.find(p);
    if (it == g_pool_live.end()) return cudaFree(p);
    PoolBlock blk = it->second;
    g_pool_live.erase(it);
    if (g_pool_cached + blk.bytes <= kPoolCap) { g_pool_free.push_back(blk); g_pool_cached += blk.bytes; return cudaSuccess; }
    return cudaFree(blk.p);
}
}  // namespace
template <typename P>
static cudaError_t dev_malloc(P** p, size_t bytes) { return pool_malloc((void**)p, bytes); }
static cudaError_t dev_free(void* p) { return pool_free(p); }

struct HostTimer {                      // D3D_TIMING=1: host-side stage times on stderr
    bool on; const char* fn; std::chrono::steady_clock::time_point t;
    explicit HostTimer(const char* f) : on(getenv("D3D_TIMING") != nullptr), fn(f), t(std::chrono::steady_clock::now()) {}
    void operator()(const char* what) {
        if (!on) return;
        auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "[%s] %-22s %8.2f ms\n", fn, what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};

struct d3d_ctx {
    int device = 0;
    int dtype = D3D_F64;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    Problem pb;
    bool have_problem = false, have_tables = false, have_params = false;
    std::vector<void*> allocs;          // problem-lifetime allocations
    void* rt_x = nullptr; void* rt_yu = nullptr; void* rt_nc = nullptr;
    double* d_lines = nullptr;          // [n_chains][H][W][Dp] scratch of the forward model
    int tap_runs = 0, tap_run2 = 0;     // runs of consecutive LSF tap offsets; first tap of the second run
    int tap_reach = 0;                  // largest |signed offset| of a significant LSF tap
    int threads = 256, ne = 0;          // sweep launch configuration (row-mapped kernels)
    int generic_threads = 256;
    bool use_slide = false; int slide_threads = 384; size_t slide_smem = 0;
    static size_t sweep_smem_base(const Problem& pb) { return smem_doubles(pb.fh, pb.fw, pb.P, pb.Dp) * sizeof(double); }   // sliding register window (seq mode)
    // pipelined sweep kernel (d3d_pipe.cuh): look-ahead L, producer warps, launch shape
    // pipelined sweep: [0] = per-site producers (A/B builds only), [1] = batched producers (shipped)
    bool use_pipe[2] = {false, false}; int pipe_threads[2] = {0, 0}; size_t pipe_smem[2] = {0, 0};
    int* d_sites_row = nullptr; int* d_run_start = nullptr; double mean_run = 0.0;
    void* d_sched = nullptr; size_t sched_cap = 0;     // work-item lists of the balanced launch
    long long sched_C = -1, sched_S = -1; int sched_G = -1, sched_max_items = 1; size_t sched_flat = 0;
    bool use_nc = false;                // uncached row kernel (2 CTAs/SM) for many chains
    bool colour_attr_set = false, apply_attr_set = false;
    bool cluster_attr_set = false, apply_cluster_attr_set = false;
    int* d_sites_colour = nullptr;      // [cube][max_sites] colour-class order
    // fused tile exchange (d3d_tile.cuh): this context's box and the peers' boxes
    void* box = nullptr; size_t box_bytes = 0; bool box_attr_set = false;
    int* d_hits = nullptr; unsigned int* d_hit_count = nullptr;   // triage of the fused exchange (d3d_tile.cuh):
    long long hits_stride = 0;                                    // records per phase that reach into this tile's region
    TileBox tb;
    std::vector<void*> ipc_opened;
    int cluster = 0;                    // > 1: generic colour kernels work a site with a CTA cluster
    double* d_rec_stage = nullptr; size_t rec_stage_cap = 0;   // staging of host-side record buffers
    size_t sweep_smem = 0;
    int64_t launches = 0, last_bytes = 0, last_updates = 0;
    cudaStream_t copy_stream = nullptr;   // device->host copies of chain rows overlapping the next sweeps (d3d_sweep)
    bool lu_tried = false;               // the profile cache of the pipelined sweep was asked for (it is optional)
    const char* last_kernel = "";        // name of the sweep kernel of the latest d3d_sweep (bench reporting)
    int64_t window_voxels_per_sweep = 0;   // sum over cubes of sum_sites wh*ww*D * chains_per_cube
    std::vector<int> h_nsites;
    size_t elem() const { return dtype == D3D_F64 ? 8 : 4; }
};

template <typename P>
static int dalloc(d3d_ctx* c, P** p, size_t bytes) {
    void* q = nullptr;
    cudaError_t e = dev_malloc(&q, bytes ? bytes : 1);
    if (e != cudaSuccess)
        return fail(D3D_ENOMEM, "dev_malloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    c->allocs.push_back(q);
    *p = (P*)q;
    return 0;
}

static void free_problem(d3d_ctx* c) {
    for (void* p : c->allocs) dev_free(p);
    c->allocs.clear();
    c->d_lines = nullptr;
    c->pb.gtab = nullptr;
    c->pb.xtab = nullptr;
    c->pb.lucache = nullptr; c->pb.lu_valid = 0; c->lu_tried = false;
    c->pb.run_start = nullptr;
    c->pb.run_last = nullptr;
    c->d_sites_row = nullptr; c->d_run_start = nullptr;
    c->have_problem = false;
    c->have_params = false;
}

extern "C" int d3d_abi_version(void) { return D3D_ABI_VERSION; }
extern "C" const char* d3d_last_error(void) { return g_last_error.c_str(); }

extern "C" int d3d_ctx_create(d3d_ctx** out, int device, int dtype) {
    if (!out) return fail(D3D_EINVAL, "d3d_ctx_create: out is NULL");
    if (dtype != D3D_F32 && dtype != D3D_F64)
        return fail(D3D_EINVAL, "d3d_ctx_create: dtype must be D3D_F32 or D3D_F64");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(D3D_ECUDA, "no CUDA device available (%s); deconv3d_b200 has no CPU fallback",
                    cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(D3D_EINVAL, "device %d out of range [0,%d)", device, n);
    CK(cudaSetDevice(device));
    d3d_ctx* c = new d3d_ctx();
    c->device = device;
    c->dtype = dtype;
    memset(&c->pb, 0, sizeof(Problem));
    CK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    CK(cudaEventCreate(&c->ev0));
    CK(cudaEventCreate(&c->ev1));
    *out = c;
    return 0;
}

extern "C" int d3d_ctx_destroy(d3d_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    HostTimer stamp("d3d_ctx_destroy");
    free_problem(c);
    stamp("free problem");
    if (c->d_sched) dev_free(c->d_sched);
    if (c->d_rec_stage) dev_free(c->d_rec_stage);
    for (void* p : c->ipc_opened) cudaIpcCloseMemHandle(p);
    if (c->box) cudaFree(c->box);
    if (c->d_hits) cudaFree(c->d_hits);
    if (c->d_hit_count) cudaFree(c->d_hit_count);
    if (c->rt_x) dev_free(c->rt_x);
    if (c->rt_yu) dev_free(c->rt_yu);
    if (c->rt_nc) dev_free(c->rt_nc);
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    cudaStreamDestroy(c->own_stream);
    delete c;
    return 0;
}

extern "C" int d3d_ctx_set_stream(d3d_ctx* c, void* s) {
    if (!c) return fail(D3D_EINVAL, "ctx is NULL");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return 0;
}

extern "C" int d3d_ctx_synchronize(d3d_ctx* c) {
    if (!c) return fail(D3D_EINVAL, "ctx is NULL");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

// ------------------------------------------------------------------------------
template <typename T>
static int ingest(d3d_ctx* c, const double* src_any, const double* nan_src_dev, void* dst, int n,
                  int mode, double** staged_out, int* nan_seen = nullptr) {
    const Problem& pb = c->pb;
    size_t cnt = (size_t)n * pb.D * pb.H * pb.W;
    double* stage = nullptr;
    CK(dev_malloc(&stage, cnt * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(stage, src_any, cnt * sizeof(double), cudaMemcpyDefault, c->stream);
    if (e != cudaSuccess) { dev_free(stage); return fail(D3D_ECUDA, "copy of cube failed: %s", cudaGetErrorString(e)); }
    long long blocks = (long long)n * pb.H * ((pb.Dp + 31) / 32) * ((pb.W + 31) / 32);
    ingest_kernel<T><<<(unsigned)blocks, 256, 0, c->stream>>>(stage, nan_src_dev, (T*)dst, n, pb.D,
                                                              pb.Dp, pb.H, pb.W, mode, nan_seen);
    c->launches++;
    e = cudaGetLastError();
    if (e != cudaSuccess) { dev_free(stage); return fail(D3D_ECUDA, "ingest launch failed: %s", cudaGetErrorString(e)); }
    if (staged_out) { *staged_out = stage; return 0; }
    CK(cudaStreamSynchronize(c->stream));
    dev_free(stage);
    return 0;
}

static void choose_launch(d3d_ctx* c) {
    // Row-mapped kernels: fw * (Dp/VEC) window threads (<= 512) + 2 scalar warps, the window
    // rows (<= fh) cached in registers: NE in {7, 13, 21}.  Anything bigger: generic kernels.
    const Problem& pb = c->pb;
    const int vec = c->dtype == D3D_F64 ? 2 : 4;
    const int zl = pb.Dp / vec;
    const int nwt = pb.fw * zl;
    c->ne = 0;
    c->threads = 256;
    if (nwt <= 320 && pb.fh <= 21 && !getenv("D3D_FORCE_GENERIC")) {
        c->ne = pb.fh <= 7 ? 7 : pb.fh <= 13 ? 13 : 21;
        c->threads = ((nwt + 31) / 32) * 32 + 64;
    }
    c->generic_threads = 256;
    {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
        c->use_nc = c->ne != 0 && pb.n_chains > sms;
        if (const char* e = getenv("D3D_ROW_VARIANT")) c->use_nc = c->ne != 0 && atoi(e) == 0;
        const int nwt_slide = (pb.fw + 1) * zl;
        c->slide_threads = ((nwt_slide + 31) / 32) * 32 + 96;
        c->slide_smem = c->sweep_smem_base(pb) + 2 * 4002 * sizeof(double) + 8964 * sizeof(unsigned short);
        c->use_slide = c->ne != 0 && c->ne <= 13 && c->slide_threads <= 384 && pb.W >= 2 &&
                       (long long)pb.H * pb.W * pb.W < 0xffffffffLL;   // multiply-high site decode
        if (const char* e = getenv("D3D_SLIDE")) c->use_slide = c->use_slide && atoi(e) != 0;
        // pipelined variant: fw + L + 1 column groups + producer / cross-term / decision warps
        for (int m = 0; m < 2; ++m) {
            const int L = m ? D3D_PIPEM_L : D3D_PIPEF_L;
            const int nprod = m ? D3D_PIPEM_NPW : D3D_PIPEF_NA + D3D_PIPEF_NP;
            const int maxt = m ? D3D_PIPEM_MAXT : D3D_PIPEF_MAXT;
            const int ng = pb.fw + L + 1;
            const int nww = (ng * zl + 31) / 32;
            c->pipe_threads[m] = (nww + nprod + D3D_PIPE_NX + PIPE_NR + 1) * 32;   // (+ reducer) + decision warp
            c->pipe_smem[m] = (m ? pipe_smem_bytes<D3D_PIPEM_L> : pipe_smem_bytes<D3D_PIPEF_L>)(
                pb.fw, c->ne ? c->ne : 7, pb.kd_n, pb.Dp, nprod, ng, zl, pb.var_is_cube != 0);
            // needs runs to pipeline along; Dp <= 64: one producer lane per channel pair, the cross
            // terms of a site in two registers per lane; <= 16 window warps: one record row
            c->use_pipe[m] = c->use_slide && c->pipe_threads[m] <= maxt && c->pipe_smem[m] <= 227 * 1024 &&
                             nww <= 16 && pb.Dp <= 64 && c->mean_run >= 4.0;
            if (const char* e = getenv("D3D_PIPE")) c->use_pipe[m] = c->use_pipe[m] && atoi(e) != 0;
        }
    }
    c->sweep_smem = smem_doubles(pb.fh, pb.fw, pb.P, pb.Dp) * sizeof(double);
    c->colour_attr_set = false;
    c->apply_attr_set = false;
    c->cluster_attr_set = false;
    c->apply_cluster_attr_set = false;
    // big windows: split one site over a thread-block cluster (d3d_tile.cuh)
    c->cluster = (c->ne == 0 && (long long)pb.fh * pb.fw * pb.Dp >= 32768) ? 4 : 0;
    if (c->cluster) {
        // A colour phase is latency-bound (one site's chain of loads, reduction, decision, update):
        // give a site as many CTAs as still leave EVERY site of the largest class resident at once
        // (2 CTAs of 320 threads x 96 registers per SM).  cfg4: 7 x 7 sites -> 6 CTAs per site =
        // 294 of 296 slots, 49.6 ms per sweep instead of 57.7 with 4 (7 and 8 need a second wave:
        // 74 ms); a 128 x 128 field (16 sites): 29.3 / 22.8 / 19.7 us per phase with 4 / 6 / 8
        // (profiles/r02_notes.md).  Taken from the WHOLE field, not from this context's tile:
        // the split of the window sums fixes their rounding, and tiles must reproduce one context.
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
        const long long per_phase = (long long)((pb.H + pb.fh - 1) / pb.fh) * ((pb.W + pb.fw - 1) / pb.fw) * pb.n_chains;
        const long long fit = per_phase > 0 ? (2LL * sms) / per_phase : 0;
        if (fit > 4) c->cluster = (int)std::min<long long>(fit, 8);   // 8: the portable cluster limit
    }
    if (const char* e = getenv("D3D_CLUSTER")) c->cluster = c->ne == 0 ? atoi(e) : 0;
    if (c->cluster < 2) c->cluster = 0;
}

extern "C" int d3d_set_problem(d3d_ctx* c, int n_cubes, int chains_per_cube, int D, int H, int W,
                               const double* data, const double* var, int var_kind,
                               const uint8_t* mask, const double* fsf, int fh, int fw,
                               const double* lsf, const double* pmin, const double* pmax,
                               const double* jump_amp, const double* gibbs_prior_var) {
    if (!c) return fail(D3D_EINVAL, "ctx is NULL");
    if (n_cubes < 1 || chains_per_cube < 1 || D < 1 || H < 1 || W < 1)
        return fail(D3D_EINVAL, "d3d_set_problem: sizes must be positive");
    if (D > 1024) return fail(D3D_EINVAL, "d3d_set_problem: at most 1024 channels are supported");
    if (fh < 1 || fw < 1 || fh % 2 == 0 || fw % 2 == 0)
        return fail(D3D_EINVAL, "FSF *must* be of odd dimensions");              // lib/run.py:210-211
    if (!data || !var || !fsf || !pmin || !pmax || !jump_amp || !gibbs_prior_var)
        return fail(D3D_EINVAL, "d3d_set_problem: NULL argument");
    if (var_kind != D3D_VAR_SCALAR && var_kind != D3D_VAR_CUBE)
        return fail(D3D_EINVAL, "d3d_set_problem: bad var_kind");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    HostTimer stamp("d3d_set_problem");
    free_problem(c);
    stamp("free previous");
    Problem& pb = c->pb;
    RtTables keep_rt = pb.rt;
    unsigned long long keep_seed = pb.seed;
    unsigned int keep_first = pb.first_chain;
    memset(&pb, 0, sizeof pb);
    pb.rt = keep_rt; pb.seed = keep_seed; pb.first_chain = keep_first;
    pb.n_comp = 1; pb.comp_ratio[0] = 1.0;                  // SingleGaussianLineModel until d3d_set_line_model
    const int vec = c->dtype == D3D_F64 ? 2 : 4;
    pb.D = D; pb.H = H; pb.W = W; pb.Dp = ((D + vec - 1) / vec) * vec;
    pb.fh = fh; pb.fw = fw; pb.fhh = (fh - 1) / 2; pb.fhw = (fw - 1) / 2;
    pb.n_cubes = n_cubes; pb.chains_per_cube = chains_per_cube;
    pb.n_chains = n_cubes * chains_per_cube;
    pb.var_is_cube = var_kind == D3D_VAR_CUBE;
    pb.has_lsf = lsf != nullptr;
    // padded spectral length of lib/convolution.py:141-144: 2**len(binary_repr(D-1))
    { int bits = 1; unsigned v = (unsigned)(D - 1); while (v >>= 1) ++bits; pb.P = 1 << bits; }

    const size_t HW = (size_t)H * W, cube_elems = HW * pb.Dp;
    int rc;
    void* d_data; void* d_iv = nullptr; void* d_err;
    if ((rc = dalloc(c, &d_data, (size_t)n_cubes * cube_elems * c->elem()))) return rc;
    if ((rc = dalloc(c, &d_err, (size_t)pb.n_chains * cube_elems * c->elem()))) return rc;
    CK(cudaMemsetAsync(d_err, 0, (size_t)pb.n_chains * cube_elems * c->elem(), c->stream));
    double* staged_data = nullptr;
    int* d_nan = nullptr;
    if ((rc = dalloc(c, &d_nan, sizeof(int)))) return rc;
    CK(cudaMemsetAsync(d_nan, 0, sizeof(int), c->stream));
    if (c->dtype == D3D_F64) rc = ingest<double>(c, data, nullptr, d_data, n_cubes, 0, &staged_data, d_nan);
    else rc = ingest<float>(c, data, nullptr, d_data, n_cubes, 0, &staged_data, d_nan);
    if (rc) return rc;
    if (!pb.var_is_cube) {
        // NaN voxels must drop out of the chi2 sums (nansum, lib/run.py:24-27, 420-425): that needs
        // per-voxel weights, which a scalar variance does not carry -- refuse instead of silently
        // giving such voxels the value 0 with full weight
        int h_nan = 0;
        CK(cudaMemcpyAsync(&h_nan, d_nan, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        if (h_nan) {
            dev_free(staged_data);
            return fail(D3D_EINVAL, "d3d_set_problem: the data holds NaN voxels; pass the variance as a cube "
                                    "(D3D_VAR_CUBE) so that they get zero weight");
        }
    }
    double* d_ivs = nullptr;
    if (pb.var_is_cube) {
        if ((rc = dalloc(c, &d_iv, (size_t)n_cubes * cube_elems * c->elem()))) { dev_free(staged_data); return rc; }
        if (c->dtype == D3D_F64) rc = ingest<double>(c, var, staged_data, d_iv, n_cubes, 1, nullptr);
        else rc = ingest<float>(c, var, staged_data, d_iv, n_cubes, 1, nullptr);
        if (rc) { dev_free(staged_data); return rc; }
    } else {
        std::vector<double> hv(n_cubes);
        CK(cudaMemcpy(hv.data(), var, n_cubes * sizeof(double), cudaMemcpyDefault));
        for (auto& v : hv) v = 1.0 / v;
        if ((rc = dalloc(c, &d_ivs, n_cubes * sizeof(double)))) { dev_free(staged_data); return rc; }
        CK(cudaMemcpy(d_ivs, hv.data(), n_cubes * sizeof(double), cudaMemcpyHostToDevice));
    }
    CK(cudaStreamSynchronize(c->stream));
    dev_free(staged_data);
    pb.data = d_data; pb.iv = d_iv; pb.iv_scalar = d_ivs; pb.err = d_err;
    stamp("cubes alloc + ingest");

    // mask -> row-major site lists (lib/run.py:553-566)
    std::vector<uint8_t> hmask((size_t)n_cubes * HW, 1);
    if (mask) CK(cudaMemcpy(hmask.data(), mask, hmask.size(), cudaMemcpyDefault));
    std::vector<int> nsites(n_cubes, 0);
    int max_sites = 1;
    for (int q = 0; q < n_cubes; ++q) {
        int cnt = 0;
        for (size_t i = 0; i < HW; ++i) cnt += hmask[q * HW + i] == 1;
        nsites[q] = cnt;
        max_sites = std::max(max_sites, cnt);
    }
    std::vector<int> sites((size_t)n_cubes * max_sites, 0);
    int64_t wvox = 0;
    for (int q = 0; q < n_cubes; ++q) {
        int k = 0;
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x)
                if (hmask[q * HW + (size_t)y * W + x] == 1) {
                    sites[(size_t)q * max_sites + k++] = y * W + x;
                    int wh = std::min(y + pb.fhh + 1, H) - std::max(y - pb.fhh, 0);
                    int ww = std::min(x + pb.fhw + 1, W) - std::max(x - pb.fhw, 0);
                    wvox += (int64_t)wh * ww * D * chains_per_cube;
                }
    }
    // the same sites in colour-class order (classes (y mod fh, x mod fw) row-major, sites of a
    // class row-major): coloured mode with many chains walks this list chain by chain
    std::vector<int> csites((size_t)n_cubes * max_sites, 0);
    for (int q = 0; q < n_cubes; ++q) {
        int k = 0;
        for (int cy = 0; cy < std::min(fh, H); ++cy)
            for (int cx = 0; cx < std::min(fw, W); ++cx)
                for (int y = cy; y < H; y += fh)
                    for (int x = cx; x < W; x += fw)
                        if (hmask[q * HW + (size_t)y * W + x] == 1) csites[(size_t)q * max_sites + k++] = y * W + x;
    }
    c->window_voxels_per_sweep = wvox;
    c->h_nsites = nsites;
    pb.max_sites = max_sites;
    // runs of the row-major list: consecutive entries on one row with x increasing by one
    // (d3d_pipe.cuh pipelines the decisions inside a run and drains at its end)
    std::vector<int> runs((size_t)n_cubes * max_sites, 0);
    {
        long long n_runs = 0, n_all = 0;
        for (int q = 0; q < n_cubes; ++q) {
            const int* sl = sites.data() + (size_t)q * max_sites;
            int* rl = runs.data() + (size_t)q * max_sites;
            for (int k = 0; k < nsites[q]; ++k) {
                const bool cont = k > 0 && sl[k] == sl[k - 1] + 1 && sl[k] % W != 0;
                rl[k] = cont ? rl[k - 1] : k;
                n_runs += !cont;
            }
            n_all += nsites[q];
        }
        c->mean_run = n_runs ? (double)n_all / (double)n_runs : 0.0;
    }
    std::vector<int> runl((size_t)n_cubes * max_sites, 0);
    for (int q = 0; q < n_cubes; ++q) {
        const int* rl = runs.data() + (size_t)q * max_sites;
        int* ll = runl.data() + (size_t)q * max_sites;
        for (int k = nsites[q] - 1; k >= 0; --k)
            ll[k] = (k + 1 < nsites[q] && rl[k + 1] == rl[k]) ? ll[k + 1] : k;
    }
    uint8_t* d_mask; int* d_sites; int* d_ns;
    if ((rc = dalloc(c, &d_mask, hmask.size()))) return rc;
    if ((rc = dalloc(c, &d_sites, sites.size() * sizeof(int)))) return rc;
    if ((rc = dalloc(c, &c->d_run_start, runs.size() * sizeof(int)))) return rc;
    CK(cudaMemcpy(c->d_run_start, runs.data(), runs.size() * sizeof(int), cudaMemcpyHostToDevice));
    c->d_sites_row = d_sites;
    pb.run_start = c->d_run_start;
    {
        int* d_rl;
        if ((rc = dalloc(c, &d_rl, runl.size() * sizeof(int)))) return rc;
        CK(cudaMemcpy(d_rl, runl.data(), runl.size() * sizeof(int), cudaMemcpyHostToDevice));
        pb.run_last = d_rl;
    }
    if ((rc = dalloc(c, &c->d_sites_colour, csites.size() * sizeof(int)))) return rc;
    CK(cudaMemcpy(c->d_sites_colour, csites.data(), csites.size() * sizeof(int), cudaMemcpyHostToDevice));
    if ((rc = dalloc(c, &d_ns, n_cubes * sizeof(int)))) return rc;
    CK(cudaMemcpy(d_mask, hmask.data(), hmask.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_sites, sites.data(), sites.size() * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ns, nsites.data(), n_cubes * sizeof(int), cudaMemcpyHostToDevice));
    pb.mask = d_mask; pb.sites = d_sites; pb.n_sites = d_ns;
    stamp("mask + site lists");

    // FSF and the circular LSF kernel.  lib/convolution.py:89-160 in direct form
    // (SURVEY.md 8a-3): out[j] = sum_i line[i]*lsf[t], t = ((j-i+P/2) mod P) - half,
    // kept iff 0 <= t < D  =>  K[m] = lsf[((m+P/2) mod P) - half], m = (j-i) mod P.
    std::vector<double> hfsf((size_t)fh * fw), hk(pb.P, 0.0);
    CK(cudaMemcpy(hfsf.data(), fsf, hfsf.size() * sizeof(double), cudaMemcpyDefault));
    if (lsf) {
        std::vector<double> hl(D);
        CK(cudaMemcpy(hl.data(), lsf, D * sizeof(double), cudaMemcpyDefault));
        int diff = pb.P - D, half = (diff & 1) ? diff / 2 + 1 : diff / 2;
        for (int m = 0; m < pb.P; ++m) {
            int t = ((m + pb.P / 2) % pb.P) - half;
            hk[m] = (t >= 0 && t < D) ? hl[t] : 0.0;
        }
    } else {
        hk[0] = 1.0;
    }
    // taps that matter: |K[m]| >= 1e-18 max|K| (the dropped ones sum to less than the rounding
    // error of the kept ones)
    std::vector<double> tapv; std::vector<int> tapm;
    {
        double kmax = 0.0;
        for (double v : hk) kmax = std::max(kmax, fabs(v));
        for (int m = 0; m < pb.P; ++m)
            if (hk[m] != 0.0 && fabs(hk[m]) >= 1e-18 * kmax) { tapv.push_back(hk[m]); tapm.push_back(m); }
    }
    pb.ntaps = (int)tapv.size();
    // the same taps as ONE dense window of signed offsets m' in [mlo, mhi] (m' = m or m - P):
    // kdense[t] = K[(mhi - t) mod P]; out[z] = sum_t kdense[t] * g[(z - mhi + t) mod P]
    std::vector<double> kdense;
    {
        int mlo = 0, mhi = 0; bool any = false;
        for (int m : tapm) {
            const int ms = m < pb.P / 2 ? m : m - pb.P;
            if (!any) { mlo = mhi = ms; any = true; }
            mlo = std::min(mlo, ms); mhi = std::max(mhi, ms);
        }
        if (!any) { mlo = mhi = 0; }
        for (int t = 0; t <= mhi - mlo; ++t) kdense.push_back(hk[((mhi - t) % pb.P + pb.P) % pb.P]);
        pb.kd_n = (int)kdense.size(); pb.kd_mhi = mhi;
        c->tap_reach = std::max(std::abs(mlo), std::abs(mhi));
    }
    // runs of consecutive tap offsets (lines_warp_kernel rotates its buffer inside a run)
    c->tap_runs = tapm.empty() ? 0 : 1;
    c->tap_run2 = (int)tapm.size();
    for (size_t t = 1; t < tapm.size(); ++t)
        if (tapm[t] != tapm[t - 1] + 1) { if (++c->tap_runs == 2) c->tap_run2 = (int)t; }
    double* d_fsf; double* d_k; double* d_tv; int* d_tm;
    if ((rc = dalloc(c, &d_fsf, hfsf.size() * sizeof(double)))) return rc;
    if ((rc = dalloc(c, &d_k, hk.size() * sizeof(double)))) return rc;
    if ((rc = dalloc(c, &d_tv, (tapv.size() + 1) * sizeof(double)))) return rc;
    if ((rc = dalloc(c, &d_tm, (tapm.size() + 1) * sizeof(int)))) return rc;
    if (!tapv.empty()) {
        CK(cudaMemcpy(d_tv, tapv.data(), tapv.size() * sizeof(double), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_tm, tapm.data(), tapm.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    pb.ktap_v = d_tv; pb.ktap_m = d_tm;
    {
        double* d_kd;
        if ((rc = dalloc(c, &d_kd, kdense.size() * sizeof(double)))) return rc;
        CK(cudaMemcpy(d_kd, kdense.data(), kdense.size() * sizeof(double), cudaMemcpyHostToDevice));
        pb.kdense = d_kd;
    }
    CK(cudaMemcpy(d_fsf, hfsf.data(), hfsf.size() * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_k, hk.data(), hk.size() * sizeof(double), cudaMemcpyHostToDevice));
    pb.fsf = d_fsf; pb.kcirc = d_k;

    double *d_pmin, *d_pmax, *d_prior;
    if ((rc = dalloc(c, &d_pmin, n_cubes * 3 * sizeof(double)))) return rc;
    if ((rc = dalloc(c, &d_pmax, n_cubes * 3 * sizeof(double)))) return rc;
    if ((rc = dalloc(c, &d_prior, n_cubes * sizeof(double)))) return rc;
    CK(cudaMemcpy(d_pmin, pmin, n_cubes * 3 * sizeof(double), cudaMemcpyDefault));
    CK(cudaMemcpy(d_pmax, pmax, n_cubes * 3 * sizeof(double), cudaMemcpyDefault));
    CK(cudaMemcpy(d_prior, gibbs_prior_var, n_cubes * sizeof(double), cudaMemcpyDefault));
    pb.pmin = d_pmin; pb.pmax = d_pmax; pb.prior_var = d_prior;
    CK(cudaMemcpy(pb.jump, jump_amp, 3 * sizeof(double), cudaMemcpyDefault));

    double* d_params; long long* d_acc; long long* d_it; double* d_rate; int* d_active; int* d_status;
    if ((rc = dalloc(c, &d_params, (size_t)pb.n_chains * HW * 3 * sizeof(double)))) return rc;
    if ((rc = dalloc(c, &d_acc, pb.n_chains * sizeof(long long)))) return rc;
    if ((rc = dalloc(c, &d_it, pb.n_chains * sizeof(long long)))) return rc;
    if ((rc = dalloc(c, &d_rate, pb.n_chains * sizeof(double)))) return rc;
    if ((rc = dalloc(c, &d_active, pb.n_chains * sizeof(int)))) return rc;
    if ((rc = dalloc(c, &d_status, 33 * sizeof(int)))) return rc;
    CK(cudaMemset(d_params, 0, (size_t)pb.n_chains * HW * 3 * sizeof(double)));
    CK(cudaMemset(d_status, 0, 33 * sizeof(int)));
    pb.params = d_params; pb.accepted = d_acc; pb.iters = d_it; pb.rate = d_rate;
    pb.active = d_active; pb.status = d_status; pb.dbg = d_status + 1;
    pb.ty0 = 0; pb.ty1 = H; pb.tx0 = 0; pb.tx1 = W;
    pb.ry0 = 0; pb.ry1 = H; pb.rx0 = 0; pb.rx1 = W;
    pb.lik_cur = nullptr; pb.acc_cur = nullptr;              // allocated by d3d_set_tile
    if ((rc = dalloc(c, &c->d_lines, (size_t)pb.n_chains * cube_elems * sizeof(double)))) return rc;

    stamp("small arrays + scratch");
    choose_launch(c);
    c->have_problem = true;
    return 0;
}

static int reset_chain_control(d3d_ctx* c) {
    const Problem& pb = c->pb;
    std::vector<long long> acc(pb.n_chains), it(pb.n_chains, 1);
    std::vector<double> rate(pb.n_chains, 0.0);
    std::vector<int> act(pb.n_chains, 1);
    for (int k = 0; k < pb.n_chains; ++k) acc[k] = c->h_nsites[k / pb.chains_per_cube];  // :341
    CK(cudaMemcpyAsync(pb.accepted, acc.data(), acc.size() * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(pb.iters, it.data(), it.size() * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(pb.rate, rate.data(), rate.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(pb.active, act.data(), act.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int d3d_set_line_model(d3d_ctx* c, int n_components, const double* offsets, const double* ratios) {
    if (c) c->pb.lu_valid = 0;            // (profile cache of the pipelined sweep: parameters move)
    if (!c || !c->have_problem) return fail(D3D_ESTATE, "d3d_set_line_model before d3d_set_problem");
    if (n_components < 1 || n_components > 4)
        return fail(D3D_EINVAL, "d3d_set_line_model: 1 to 4 tied Gaussian components are supported");
    if (n_components > 1 && (!offsets || !ratios)) return fail(D3D_EINVAL, "d3d_set_line_model: NULL argument");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    Problem& pb = c->pb;
    double off[4] = {0, 0, 0, 0}, rat[4] = {1, 0, 0, 0};
    if (n_components > 1) {
        CK(cudaMemcpy(off, offsets, n_components * sizeof(double), cudaMemcpyDefault));
        CK(cudaMemcpy(rat, ratios, n_components * sizeof(double), cudaMemcpyDefault));
    }
    if (!(rat[0] != 0.0)) return fail(D3D_EINVAL, "d3d_set_line_model: the first component carries the amplitude, its ratio must not be 0");
    // normalised to the first component: offset 0, ratio 1 (the amplitude parameter is ITS amplitude)
    pb.n_comp = n_components;
    for (int k = 0; k < 4; ++k) {
        pb.comp_off[k] = k < n_components ? off[k] - off[0] : 0.0;
        pb.comp_ratio[k] = k < n_components ? rat[k] / rat[0] : 0.0;
    }
    return 0;
}

extern "C" int d3d_set_rtnorm_tables(d3d_ctx* c, const double* x, int nx, const double* yu, int nyu,
                                     const int32_t* ncell, int nncell) {
    if (!c) return fail(D3D_EINVAL, "ctx is NULL");
    if (nx != 4002 || nyu != 4001 || nncell != 8961 || !x || !yu || !ncell)
        return fail(D3D_EINVAL, "rtnorm tables must have 4002/4001/8961 entries (lib/rtnorm.py:227-2681)");
    CK(cudaSetDevice(c->device));
    if (!c->rt_x) {
        CK(dev_malloc(&c->rt_x, nx * sizeof(double)));
        CK(dev_malloc(&c->rt_yu, nyu * sizeof(double)));
        CK(dev_malloc(&c->rt_nc, nncell * sizeof(int)));
    }
    CK(cudaMemcpy(c->rt_x, x, nx * sizeof(double), cudaMemcpyDefault));
    CK(cudaMemcpy(c->rt_yu, yu, nyu * sizeof(double), cudaMemcpyDefault));
    CK(cudaMemcpy(c->rt_nc, ncell, nncell * sizeof(int), cudaMemcpyDefault));
    c->pb.rt.x = (const double*)c->rt_x;
    c->pb.rt.yu = (const double*)c->rt_yu;
    c->pb.rt.ncell = (const int*)c->rt_nc;
    c->have_tables = true;
    return 0;
}

extern "C" int d3d_set_rng(d3d_ctx* c, uint64_t seed, uint32_t first_chain_id) {
    if (!c) return fail(D3D_EINVAL, "ctx is NULL");
    c->pb.seed = seed;
    c->pb.first_chain = first_chain_id;
    return 0;
}

extern "C" int d3d_set_params(d3d_ctx* c, const double* params) {
    if (c) c->pb.lu_valid = 0;            // (profile cache of the pipelined sweep: parameters move)
    if (!c || !c->have_problem) return fail(D3D_ESTATE, "d3d_set_params before d3d_set_problem");
    if (!params) return fail(D3D_EINVAL, "params is NULL");
    CK(cudaSetDevice(c->device));
    const Problem& pb = c->pb;
    CK(cudaMemcpyAsync(pb.params, params, (size_t)pb.n_chains * pb.H * pb.W * 3 * sizeof(double),
                       cudaMemcpyDefault, c->stream));
    c->have_params = true;
    return reset_chain_control(c);
}

extern "C" int d3d_get_params(d3d_ctx* c, double* params) {
    if (!c || !c->have_problem) return fail(D3D_ESTATE, "d3d_get_params before d3d_set_problem");
    CK(cudaSetDevice(c->device));
    const Problem& pb = c->pb;
    CK(cudaMemcpyAsync(params, pb.params, (size_t)pb.n_chains * pb.H * pb.W * 3 * sizeof(double),
                       cudaMemcpyDefault, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int d3d_init_params_uniform(d3d_ctx* c) {
    if (c) c->pb.lu_valid = 0;            // (profile cache of the pipelined sweep: parameters move)
    if (!c || !c->have_problem) return fail(D3D_ESTATE, "d3d_init_params_uniform before d3d_set_problem");
    CK(cudaSetDevice(c->device));
    const Problem& pb = c->pb;
    size_t n = (size_t)pb.n_chains * pb.H * pb.W;
    init_params_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(pb);
    c->launches++;
    CK(cudaGetLastError());
    c->have_params = true;
    return reset_chain_control(c);
}

// ------------------------------------------------------------------------------
// forward model
static int stencil_config(const d3d_ctx* c, int* TY, int* TX, int* ZC, size_t* smem) {
    const Problem& pb = c->pb;
    const size_t limit = 200 * 1024;
    int ty = 8, tx = 8;
    while (ty > pb.H && ty > 1) ty /= 2;
    while (tx > pb.W && tx > 1) tx /= 2;
    for (;;) {
        size_t per_ch = (size_t)(ty + pb.fh - 1) * (tx + pb.fw - 1) * sizeof(double);
        size_t fixed = (size_t)pb.fh * pb.fw * sizeof(double);
        long long zc = ((long long)limit - (long long)fixed) / (long long)per_ch;
        if (zc >= 1) {
            int z = (int)std::min<long long>(zc, pb.Dp);
            *TY = ty; *TX = tx; *ZC = z; *smem = fixed + per_ch * z;
            return 0;
        }
        if (ty == 1 && tx == 1) return fail(D3D_EINVAL, "FSF too large for the stencil kernel");
        if (ty >= tx && ty > 1) ty /= 2; else if (tx > 1) tx /= 2; else ty /= 2;
    }
}


// Tensor map (TMA descriptor) of a [n_sets][H][W][Dp] array of doubles for boxes of
// (zc, bx, by, 1) elements; positions outside the array read as zero.  The driver's encoder is
// looked up at run time (no link against libcuda); false when it is not there or refuses.
static bool encode_lines_map(CUtensorMap* map, const double* base, int Dp, int W, int H, int n_sets, int zc, int bx, int by) {
    typedef CUresult (*Encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static Encode enc = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            enc = (Encode)fn;
        else
            cudaGetLastError();
    }
    if (!enc || bx > 256 || by > 256 || zc > 256 || ((size_t)Dp * sizeof(double)) % 16 || ((uintptr_t)base & 15)) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)Dp, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_sets};
    const cuuint64_t strides[3] = {(cuuint64_t)Dp * sizeof(double), (cuuint64_t)W * Dp * sizeof(double),
                                   (cuuint64_t)H * W * Dp * sizeof(double)};
    const cuuint32_t box[4] = {(cuuint32_t)zc, (cuuint32_t)bx, (cuuint32_t)by, 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int run_forward(d3d_ctx* c, const double* d_params, int convolve, double* sim_dev,
                       int write_err, double* chi2_dev, int n_sets = -1) {
    Problem pb = c->pb;
    if (n_sets > 0) pb.n_chains = n_sets;        // explicit parameter sets (first n_sets chain slots)
    const size_t HW = (size_t)pb.H * pb.W;
    {
        const int threads = std::max(256, ((pb.Dp + 31) / 32) * 32);
        const int spb = threads / pb.Dp;
        size_t total = (size_t)pb.n_chains * HW;
        unsigned blocks = (unsigned)std::min<size_t>((total + spb - 1) / spb, 148 * 32);
        size_t smem = ((size_t)pb.P + (pb.P + 1) / 2 + (size_t)spb * pb.P + (size_t)spb * 4) * sizeof(double);
        if (total >= ((size_t)1 << 31)) return fail(D3D_EINVAL, "too many spaxels for one forward call");
        // warp-shuffle spectral pass (one warp per spaxel) whenever the padded depth fits 4 registers
        // per lane; the shared-memory kernel for deeper cubes (and under D3D_LINES_SMEM, for A/B runs)
        const int R = pb.P <= 32 ? 1 : pb.P <= 64 ? 2 : pb.P <= 128 ? 4 : 0;
        const size_t wblocks = (size_t)pb.n_chains * ((HW + LINES_SPB - 1) / LINES_SPB);
        // one thread per spaxel, spectrum in registers (lines_lane_kernel): needs the depth and
        // the reach of the significant LSF taps among the instantiated pairs
        bool lane_done = false;
        if (!getenv("D3D_LINES_WARP") && !getenv("D3D_LINES_SMEM")) {
            const int mh = c->tap_reach;                 // max |signed offset| of a significant tap
            const unsigned lb = (unsigned)((total + 127) / 128);
#define D3D_LANE(DPV, MHV)                                                                          \
            if (!lane_done && pb.Dp == DPV && mh <= MHV) {                                          \
                lines_lane_kernel<DPV, MHV><<<lb, 128, 0, c->stream>>>(pb, d_params, c->d_lines, convolve); \
                lane_done = true;                                                                   \
            }
            D3D_LANE(16, 7) D3D_LANE(30, 8) D3D_LANE(30, 15) D3D_LANE(32, 8) D3D_LANE(32, 15)
            D3D_LANE(40, 8) D3D_LANE(40, 15) D3D_LANE(64, 8) D3D_LANE(64, 15)
#undef D3D_LANE
        }
        if (lane_done) {
        } else if (R && pb.Dp <= 32 * R && pb.ntaps <= 32 * R && c->tap_runs <= 2 && wblocks < ((size_t)1 << 31) &&
            !getenv("D3D_LINES_SMEM")) {
            const unsigned wb = (unsigned)wblocks;
            if (R == 1) lines_warp_kernel<1><<<wb, 256, 0, c->stream>>>(pb, d_params, c->d_lines, convolve, c->tap_run2);
            else if (R == 2) lines_warp_kernel<2><<<wb, 256, 0, c->stream>>>(pb, d_params, c->d_lines, convolve, c->tap_run2);
            else lines_warp_kernel<4><<<wb, 256, 0, c->stream>>>(pb, d_params, c->d_lines, convolve, c->tap_run2);
        } else {
            lines_kernel<<<blocks, threads, smem, c->stream>>>(pb, d_params, c->d_lines, convolve);
        }
        c->launches++;
        CK(cudaGetLastError());
    }
    // register-tiled kernel with compile-time width up to 17 columns (Dp is even: z-pairs), the
    // chunked wide kernel above that (faster from 21 columns on), the scalar one as last resort
    {
        // two output rows per thread (tile 16 x 16) while two CTAs of that tile still fit an SM
        const int TXt = 16, ZCt = 16;
        auto tiled_smem = [&](int ty) {     // tile, FSF (even count), one mbarrier
            return ((size_t)((pb.fh * (pb.fw + 1) + 1) & ~1) + (size_t)(ty + pb.fh - 1) * (TXt + pb.fw - 1) * ZCt + 2) * sizeof(double);
        };
        const int RYt = (pb.H > 8 && tiled_smem(16) <= 110 * 1024 && !getenv("D3D_STENCIL_RY1")) ? 2 : 1;
        const int TYt = 8 * RYt;
        size_t smem_t = tiled_smem(TYt);
        bool ok = smem_t <= 200 * 1024 && !getenv("D3D_STENCIL_SCALAR") && !getenv("D3D_STENCIL_WIDE") && pb.Dp % 2 == 0;
        dim3 grid_t((unsigned)(pb.n_chains * ((pb.H + TYt - 1) / TYt) * ((pb.W + TXt - 1) / TXt)),
                    (unsigned)((pb.Dp + ZCt - 1) / ZCt));
        bool launched = false;
        // tensor map of the [set][y][x][z] array of convolved lines for the halo-tile box of this
        // launch (host-side encoding only; falls back to the cp.async staging loop without it)
        CUtensorMap tmap;
        memset(&tmap, 0, sizeof tmap);
        bool tma = ok && !getenv("D3D_STENCIL_NO_TMA") &&
                   encode_lines_map(&tmap, c->d_lines, pb.Dp, pb.W, pb.H, pb.n_chains, ZCt, TXt + pb.fw - 1, TYt + pb.fh - 1);
#define D3D_TILED_RY(TT, FWV, RYV)                                                                  \
        {                                                                                           \
            if (tma) {                                                                              \
                CK(cudaFuncSetAttribute(stencil_tiled_kernel<TT, FWV, RYV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t)); \
                stencil_tiled_kernel<TT, FWV, RYV, true><<<grid_t, 256, smem_t, c->stream>>>(pb, c->d_lines, sim_dev, write_err, chi2_dev, tmap); \
            } else {                                                                                \
                CK(cudaFuncSetAttribute(stencil_tiled_kernel<TT, FWV, RYV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t)); \
                stencil_tiled_kernel<TT, FWV, RYV, false><<<grid_t, 256, smem_t, c->stream>>>(pb, c->d_lines, sim_dev, write_err, chi2_dev, tmap); \
            }                                                                                       \
        }
#define D3D_TILED(FWV)                                                                              \
        if (ok && !launched && pb.fw == FWV) {                                                      \
            if (c->dtype == D3D_F64) {                                                              \
                if (RYt == 2) D3D_TILED_RY(double, FWV, 2) else D3D_TILED_RY(double, FWV, 1)        \
            } else {                                                                                \
                if (RYt == 2) D3D_TILED_RY(float, FWV, 2) else D3D_TILED_RY(float, FWV, 1)          \
            }                                                                                       \
            launched = true;                                                                        \
        }
        D3D_TILED(3) D3D_TILED(5) D3D_TILED(7) D3D_TILED(9) D3D_TILED(11) D3D_TILED(13)
        D3D_TILED(15) D3D_TILED(17)
#undef D3D_TILED_RY
#undef D3D_TILED
        if (launched) {
            c->launches++;
            CK(cudaGetLastError());
            return 0;
        }
        // any other width: chunked register tiling (tile 16 x 32 spaxels x 4 channels)
        const size_t smem_w = ((size_t)((pb.fh * pb.fw + 1) & ~1) +
                               (size_t)(16 + pb.fh - 1) * wide_row_positions(pb.fw) * 4) * sizeof(double);
        if (smem_w <= 220 * 1024 && !getenv("D3D_STENCIL_SCALAR") && pb.Dp % 2 == 0) {
            dim3 grid_w((unsigned)(pb.n_chains * ((pb.H + 15) / 16) * ((pb.W + 31) / 32)), (unsigned)((pb.Dp + 3) / 4));
            if (c->dtype == D3D_F64) {
                CK(cudaFuncSetAttribute(stencil_wide_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w));
                stencil_wide_kernel<double><<<grid_w, 256, smem_w, c->stream>>>(pb, c->d_lines, sim_dev, write_err, chi2_dev);
            } else {
                CK(cudaFuncSetAttribute(stencil_wide_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w));
                stencil_wide_kernel<float><<<grid_w, 256, smem_w, c->stream>>>(pb, c->d_lines, sim_dev, write_err, chi2_dev);
            }
            c->launches++;
            CK(cudaGetLastError());
            return 0;
        }
    }
    int TY, TX, ZC; size_t smem;
    int rc = stencil_config(c, &TY, &TX, &ZC, &smem);
    if (rc) return rc;
    dim3 grid((unsigned)(pb.n_chains * ((pb.H + TY - 1) / TY) * ((pb.W + TX - 1) / TX)),
              (unsigned)((pb.Dp + ZC - 1) / ZC));
    if (c->dtype == D3D_F64) {
        CK(cudaFuncSetAttribute(stencil_zchunk_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        stencil_zchunk_kernel<double><<<grid, 256, smem, c->stream>>>(pb, c->d_lines, sim_dev, write_err, chi2_dev, TY, TX, ZC);
    } else {
        CK(cudaFuncSetAttribute(stencil_zchunk_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        stencil_zchunk_kernel<float><<<grid, 256, smem, c->stream>>>(pb, c->d_lines, sim_dev, write_err, chi2_dev, TY, TX, ZC);
    }
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

static int forward_common(d3d_ctx* c, const double* params_any, int convolve, double* sim_out,
                          int write_err, double* chi2_out, int n_sets = -1) {
    Problem pb = c->pb;
    if (n_sets > 0) pb.n_chains = n_sets;
    CK(cudaSetDevice(c->device));
    const size_t HW = (size_t)pb.H * pb.W;
    const size_t sim_cnt = (size_t)pb.n_chains * pb.D * HW;
    double* d_params = pb.params;
    double* tmp_params = nullptr;
    if (params_any) {
        CK(dev_malloc(&tmp_params, (size_t)pb.n_chains * HW * 3 * sizeof(double)));
        cudaError_t e = cudaMemcpyAsync(tmp_params, params_any, (size_t)pb.n_chains * HW * 3 * sizeof(double), cudaMemcpyDefault, c->stream);
        if (e != cudaSuccess) { dev_free(tmp_params); return fail(D3D_ECUDA, "params copy failed: %s", cudaGetErrorString(e)); }
        d_params = tmp_params;
    }
    double* d_sim = nullptr; double* d_chi = nullptr;
    int rc = 0;
    if (sim_out) { cudaError_t e = dev_malloc(&d_sim, sim_cnt * sizeof(double)); if (e != cudaSuccess) rc = fail(D3D_ENOMEM, "cudaMalloc sim failed"); }
    if (!rc && chi2_out) {
        cudaError_t e = dev_malloc(&d_chi, pb.n_chains * sizeof(double));
        if (e != cudaSuccess) rc = fail(D3D_ENOMEM, "cudaMalloc chi2 failed");
        else cudaMemsetAsync(d_chi, 0, pb.n_chains * sizeof(double), c->stream);
    }
    if (!rc) rc = run_forward(c, d_params, convolve, d_sim, write_err, d_chi, n_sets);
    if (!rc && sim_out) {
        cudaError_t e = cudaMemcpyAsync(sim_out, d_sim, sim_cnt * sizeof(double), cudaMemcpyDefault, c->stream);
        if (e != cudaSuccess) rc = fail(D3D_ECUDA, "sim copy failed: %s", cudaGetErrorString(e));
    }
    if (!rc && chi2_out) {
        cudaError_t e = cudaMemcpyAsync(chi2_out, d_chi, pb.n_chains * sizeof(double), cudaMemcpyDefault, c->stream);
        if (e != cudaSuccess) rc = fail(D3D_ECUDA, "chi2 copy failed: %s", cudaGetErrorString(e));
    }
    int h_status = 0;
    if (!rc) cudaMemcpyAsync(&h_status, pb.status, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (!rc && e != cudaSuccess) rc = fail(D3D_ECUDA, "forward failed: %s", cudaGetErrorString(e));
    if (!rc && h_status == 3) {
        cudaMemsetAsync(pb.status, 0, sizeof(int), c->stream);
        rc = fail(D3D_ECUDA, "the spatial pass of the forward model never received a halo tile from the TMA unit; "
                             "results of this call are invalid (D3D_STENCIL_NO_TMA=1 selects the cp.async staging)");
    }
    if (tmp_params) dev_free(tmp_params);
    if (d_sim) dev_free(d_sim);
    if (d_chi) dev_free(d_chi);
    return rc;
}

extern "C" int d3d_forward(d3d_ctx* c, double* sim_out, int write_err, double* chi2_out) {
    if (!c || !c->have_problem || !c->have_params)
        return fail(D3D_ESTATE, "d3d_forward needs d3d_set_problem and parameters");
    return forward_common(c, nullptr, 1, sim_out, write_err, chi2_out);
}

extern "C" int d3d_simulate(d3d_ctx* c, const double* params, int n_sets, double* sim_out) {
    if (!c || !c->have_problem) return fail(D3D_ESTATE, "d3d_simulate before d3d_set_problem");
    if (!params || !sim_out) return fail(D3D_EINVAL, "NULL argument");
    if (n_sets < 1 || n_sets > c->pb.n_chains)
        return fail(D3D_EINVAL, "d3d_simulate: n_sets must be in [1, n_chains]");
    return forward_common(c, params, 1, sim_out, 0, nullptr, n_sets);
}


extern "C" int d3d_simulate_clean(d3d_ctx* c, const double* params, int n_sets, double* sim_out) {
    if (!c || !c->have_problem) return fail(D3D_ESTATE, "d3d_simulate_clean before d3d_set_problem");
    if (!params || !sim_out) return fail(D3D_EINVAL, "NULL argument");
    if (n_sets < 1 || n_sets > c->pb.n_chains)
        return fail(D3D_EINVAL, "d3d_simulate_clean: n_sets must be in [1, n_chains]");
    CK(cudaSetDevice(c->device));
    Problem pb = c->pb;
    pb.n_chains = n_sets;
    const size_t HW = (size_t)pb.H * pb.W;
    size_t total = (size_t)pb.n_chains * pb.D * HW;
    double *d_p = nullptr, *d_o = nullptr;
    CK(dev_malloc(&d_p, (size_t)pb.n_chains * HW * 3 * sizeof(double)));
    cudaError_t e = dev_malloc(&d_o, total * sizeof(double));
    if (e != cudaSuccess) { dev_free(d_p); return fail(D3D_ENOMEM, "cudaMalloc failed"); }
    cudaMemcpyAsync(d_p, params, (size_t)pb.n_chains * HW * 3 * sizeof(double), cudaMemcpyDefault, c->stream);
    clean_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(pb, d_p, d_o);
    c->launches++;
    cudaMemcpyAsync(sim_out, d_o, total * sizeof(double), cudaMemcpyDefault, c->stream);
    e = cudaStreamSynchronize(c->stream);
    dev_free(d_p); dev_free(d_o);
    if (e != cudaSuccess) return fail(D3D_ECUDA, "simulate_clean failed: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int d3d_get_residual(d3d_ctx* c, double* err_out) {
    if (!c || !c->have_problem) return fail(D3D_ESTATE, "d3d_get_residual before d3d_set_problem");
    if (!err_out) return fail(D3D_EINVAL, "NULL argument");
    CK(cudaSetDevice(c->device));
    const Problem& pb = c->pb;
    size_t total = (size_t)pb.n_chains * pb.D * pb.H * pb.W;
    double* d_o = nullptr;
    CK(dev_malloc(&d_o, total * sizeof(double)));
    long long blocks = (long long)pb.n_chains * pb.H * ((pb.Dp + 31) / 32) * ((pb.W + 31) / 32);
    if (c->dtype == D3D_F64)
        egest_kernel<double><<<(unsigned)blocks, 256, 0, c->stream>>>((const double*)pb.err, d_o, pb.n_chains, pb.D, pb.Dp, pb.H, pb.W);
    else
        egest_kernel<float><<<(unsigned)blocks, 256, 0, c->stream>>>((const float*)pb.err, d_o, pb.n_chains, pb.D, pb.Dp, pb.H, pb.W);
    c->launches++;
    cudaMemcpyAsync(err_out, d_o, total * sizeof(double), cudaMemcpyDefault, c->stream);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    dev_free(d_o);
    if (e != cudaSuccess) return fail(D3D_ECUDA, "get_residual failed: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int d3d_conv1d(d3d_ctx* c, const double* lines, int n, int batch, const double* lsf,
                          double* out) {
    if (!c) return fail(D3D_EINVAL, "ctx is NULL");
    if (!lines || !lsf || !out || n < 1 || batch < 1) return fail(D3D_EINVAL, "d3d_conv1d: bad argument");
    CK(cudaSetDevice(c->device));
    int bits = 1; { unsigned v = (unsigned)(n - 1); while (v >>= 1) ++bits; }
    const int P = 1 << bits;
    std::vector<double> hl(n), hk(P, 0.0);
    CK(cudaMemcpy(hl.data(), lsf, n * sizeof(double), cudaMemcpyDefault));
    int diff = P - n, half = (diff & 1) ? diff / 2 + 1 : diff / 2;
    for (int m = 0; m < P; ++m) {
        int t = ((m + P / 2) % P) - half;
        hk[m] = (t >= 0 && t < n) ? hl[t] : 0.0;
    }
    double *d_l = nullptr, *d_k = nullptr, *d_o = nullptr;
    size_t cnt = (size_t)n * batch;
    CK(dev_malloc(&d_l, cnt * sizeof(double)));
    CK(dev_malloc(&d_o, cnt * sizeof(double)));
    CK(dev_malloc(&d_k, P * sizeof(double)));
    cudaMemcpyAsync(d_l, lines, cnt * sizeof(double), cudaMemcpyDefault, c->stream);
    cudaMemcpyAsync(d_k, hk.data(), P * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    size_t smem = ((size_t)P + n) * sizeof(double);
    int rc = 0;
    if (smem > 200 * 1024) rc = fail(D3D_EINVAL, "d3d_conv1d: line too long (%d)", n);
    if (!rc) {
        cudaFuncSetAttribute(conv1d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        conv1d_kernel<<<(unsigned)std::min(batch, 148 * 8), 128, smem, c->stream>>>(d_l, d_k, d_o, n, P, batch);
        c->launches++;
        cudaMemcpyAsync(out, d_o, cnt * sizeof(double), cudaMemcpyDefault, c->stream);
        cudaError_t e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail(D3D_ECUDA, "d3d_conv1d failed: %s", cudaGetErrorString(e));
    }
    dev_free(d_l); dev_free(d_o); dev_free(d_k);
    return rc;
}

extern "C" int d3d_rtnorm(d3d_ctx* c, int n, const double* a, const double* b, const double* mu,
                          const double* sigma, uint64_t seed, uint32_t chain, uint32_t sweep,
                          double* out, int32_t* used_out) {
    if (!c) return fail(D3D_EINVAL, "ctx is NULL");
    if (!c->have_tables) return fail(D3D_ESTATE, "d3d_rtnorm needs d3d_set_rtnorm_tables");
    if (n < 1 || !a || !b || !mu || !sigma || !out) return fail(D3D_EINVAL, "d3d_rtnorm: bad argument");
    CK(cudaSetDevice(c->device));
    double* d = nullptr; int* d_used = nullptr; int* d_status = nullptr;
    CK(dev_malloc(&d, (size_t)n * 5 * sizeof(double)));
    CK(dev_malloc(&d_used, (size_t)(n + 1) * sizeof(int)));
    d_status = d_used + n;
    cudaMemsetAsync(d_status, 0, sizeof(int), c->stream);
    const double* src[4] = {a, b, mu, sigma};
    for (int j = 0; j < 4; ++j)
        cudaMemcpyAsync(d + (size_t)j * n, src[j], (size_t)n * sizeof(double), cudaMemcpyDefault, c->stream);
    rtnorm_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(c->pb.rt, n, d, d + n, d + 2 * (size_t)n,
                                                          d + 3 * (size_t)n, seed, chain, sweep,
                                                          d + 4 * (size_t)n, d_used, d_status);
    c->launches++;
    int h_status = 0;
    cudaMemcpyAsync(out, d + 4 * (size_t)n, (size_t)n * sizeof(double), cudaMemcpyDefault, c->stream);
    if (used_out) cudaMemcpyAsync(used_out, d_used, (size_t)n * sizeof(int), cudaMemcpyDefault, c->stream);
    cudaMemcpyAsync(&h_status, d_status, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    dev_free(d); dev_free(d_used);
    if (e != cudaSuccess) return fail(D3D_ECUDA, "d3d_rtnorm failed: %s", cudaGetErrorString(e));
    if (h_status) return fail(D3D_ENUMERIC, "Truncated ndst in [a,b]: b MUST be greater than a (or NaN bounds)");
    return 0;
}

// ------------------------------------------------------------------------------
extern "C" int d3d_delta_logl(d3d_ctx* c, int chain, int y, int x, const double p_new[3],
                              double out[3]) {
    if (!c || !c->have_problem || !c->have_params)
        return fail(D3D_ESTATE, "d3d_delta_logl needs a problem and parameters");
    const Problem& pb = c->pb;
    if (chain < 0 || chain >= pb.n_chains || y < 0 || y >= pb.H || x < 0 || x >= pb.W || !p_new || !out)
        return fail(D3D_EINVAL, "d3d_delta_logl: bad argument");
    CK(cudaSetDevice(c->device));
    double* d_out = nullptr;
    CK(dev_malloc(&d_out, 3 * sizeof(double)));
    EvalReq ev;
    ev.enabled = 1; ev.p_new[0] = p_new[0]; ev.p_new[1] = p_new[1]; ev.p_new[2] = p_new[2];
    ev.out = d_out;
    const int site = y * pb.W + x;
    size_t smem = c->sweep_smem;
#define LAUNCH_EVAL(T, IV)                                                                        \
    do {                                                                                          \
        cudaFuncSetAttribute(eval_kernel<T, IV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        eval_kernel<T, IV><<<1, c->generic_threads, smem, c->stream>>>(pb, chain, site, ev);              \
    } while (0)
    if (c->dtype == D3D_F64) { if (pb.var_is_cube) LAUNCH_EVAL(double, true); else LAUNCH_EVAL(double, false); }
    else { if (pb.var_is_cube) LAUNCH_EVAL(float, true); else LAUNCH_EVAL(float, false); }
#undef LAUNCH_EVAL
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, 3 * sizeof(double), cudaMemcpyDefault, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    dev_free(d_out);
    if (e != cudaSuccess) return fail(D3D_ECUDA, "d3d_delta_logl failed: %s", cudaGetErrorString(e));
    return 0;
}

// ------------------------------------------------------------------------------
__global__ void fill_ll_kernel(long long* p, int n, long long v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

template <typename T, bool IV, int NE>
static cudaError_t launch_seq(d3d_ctx* c, long long it0, long long it1, int keep, double min_rate,
                              double* chain_dev, double* lik_dev, long long row_first,
                              long long rows_local) {
    if (NE != 0 && c->use_slide) {
        const int ne = NE ? NE : 7;
        const int nes = ne > 13 ? 13 : ne;
        if (!c->pb.gtab) {                                      // static G table, once per problem
            double* g = nullptr;
            const size_t n = (size_t)c->pb.n_cubes * c->pb.H * c->pb.W;
            if (dalloc(c, &g, n * c->pb.Dp * sizeof(double))) return cudaErrorMemoryAllocation;
            gtable_kernel<T, IV><<<(unsigned)n, 64, 0, c->stream>>>(c->pb, g);
            c->launches++;
            c->pb.gtab = g;
        }
        // The pipelined kernel walks the ROW-MAJOR list (runs of consecutive sites); the
        // colour-ordered list of the chain-per-CTA coloured mode has no runs: sliding-window kernel.
        // D3D_PIPE=0 forces the sliding-window kernel (A/B runs, tests of that kernel).
        int pm = 1;
#ifdef D3D_PIPE_PERSITE
        if (const char* e = getenv("D3D_PIPE")) { if (atoi(e) == 3) pm = 0; }
#endif
        const bool pipe = c->use_pipe[pm] && c->pb.sites == c->d_sites_row;
        if (pipe && !c->pb.xtab) {                              // static cross-term tables, once per problem
            double* xt = nullptr;
            const size_t n = (size_t)c->pb.n_cubes * D3D_PIPE_LMAX * c->pb.max_sites * c->pb.Dp;
            if (dalloc(c, &xt, n * sizeof(double))) return cudaErrorMemoryAllocation;
            dim3 grid(c->pb.max_sites, D3D_PIPE_LMAX, c->pb.n_cubes);
            xtab_kernel<T, IV><<<grid, 64, 0, c->stream>>>(c->pb, D3D_PIPE_LMAX, c->d_run_start, xt);
            c->launches++;
            c->pb.xtab = xt;
            c->pb.xtab_L = D3D_PIPE_LMAX;
        }
        if (pipe && !c->lu_tried && !getenv("D3D_NO_LUCACHE")) {
            // cache of the unit profile every site ended its last visit with: the next sweep's
            // "old" profile (half of the producers' work).  Optional: without memory for it the
            // kernel recomputes both profiles as before.
            c->lu_tried = true;
            void* q = nullptr;
            const size_t bytes = (size_t)c->pb.n_chains * c->pb.H * c->pb.W * c->pb.Dp * sizeof(double);
            if (dev_malloc(&q, bytes) == cudaSuccess) { c->allocs.push_back(q); c->pb.lucache = (double*)q; }
            else cudaGetLastError();
            c->pb.lu_valid = 0;
        }
        if (!pipe) c->pb.lu_valid = 0;                          // another kernel moves the parameters
        typedef void (*SeqKern)(const Problem, long long, long long, int, double, double*, double*,
                                long long, long long, const int4*, const int*, int, volatile long long*);
        SeqKern kern = sweep_seq_slide_kernel<T, IV, nes>;
        int threads = c->slide_threads;
        size_t smem = c->slide_smem;
        c->last_kernel = !pipe ? "sweep_seq_slide_kernel" : pm ? "sweep_seq_pipe_kernel" : "sweep_seq_pipe_kernel<per-site>";
        if (pipe) {
            // (square FSF of the template's size: the window geometry is a compile-time constant)
            const bool sq = c->pb.fh == nes && c->pb.fw == nes;
#ifdef D3D_PIPE_PERSITE
            if (pm == 0)
                kern = sq ? sweep_seq_pipe_kernel<T, IV, nes, true, D3D_PIPEF_L, D3D_PIPEF_NA, D3D_PIPEF_NP, D3D_PIPE_NX, 0, D3D_PIPEF_MAXT>
                          : sweep_seq_pipe_kernel<T, IV, nes, false, D3D_PIPEF_L, D3D_PIPEF_NA, D3D_PIPEF_NP, D3D_PIPE_NX, 0, D3D_PIPEF_MAXT>;
            else
#endif
                kern = sq ? sweep_seq_pipe_kernel<T, IV, nes, true, D3D_PIPEM_L, D3D_PIPEM_NPW, 0, D3D_PIPE_NX, 1, D3D_PIPEM_MAXT>
                          : sweep_seq_pipe_kernel<T, IV, nes, false, D3D_PIPEM_L, D3D_PIPEM_NPW, 0, D3D_PIPE_NX, 1, D3D_PIPEM_MAXT>;
            threads = c->pipe_threads[pm];
            smem = c->pipe_smem[pm];
        }
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        // balance chains over the SMs (McNaughton wrap-around of the chain x sweep rectangle)
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
        const int C = c->pb.n_chains;
        const long long S = it1 - it0;
        if (C > sms && S > 0 && !getenv("D3D_NO_BALANCE")) {
            const int G = sms;
            // the work-item lists depend on (chains, sweeps, CTAs) only: built and uploaded once,
            // then re-used by every call of the same shape (no host synchronisation per call);
            // the hand-over counters are reset on the stream by a one-block kernel
            if (c->sched_C != C || c->sched_S != S || c->sched_G != G) {
                const long long Tslots = (C * S + G - 1) / G;       // sweeps per CTA
                std::vector<std::vector<int4>> lists(G);
                int g = 0; long long used = 0;
                for (int ch = 0; ch < C; ++ch) {
                    long long room = Tslots - used;
                    if (room >= S) {                                // whole chain on CTA g
                        lists[g].push_back(make_int4(ch, 0, (int)S, 0));
                        used += S;
                        if (used == Tslots && g + 1 < G) { ++g; used = 0; }
                    } else {
                        // split: the LATE slots of CTA g take the chain's last `room` sweeps, the
                        // EARLY slots of CTA g+1 its first S-room sweeps
                        if (room > 0) lists[g].push_back(make_int4(ch, (int)(S - room), (int)S, 0));
                        ++g; used = 0;
                        lists[g].insert(lists[g].begin(), make_int4(ch, 0, (int)(S - room), 0));
                        used = S - room;
                    }
                }
                int max_items = 1;
                for (auto& l : lists) max_items = std::max(max_items, (int)l.size());
                std::vector<int4> flat((size_t)G * max_items, make_int4(0, 0, 0, 0));
                std::vector<int> cnt(G);
                for (int q = 0; q < G; ++q) {
                    cnt[q] = (int)lists[q].size();
                    for (size_t k = 0; k < lists[q].size(); ++k) flat[(size_t)q * max_items + k] = lists[q][k];
                }
                const size_t need = flat.size() * sizeof(int4) + G * sizeof(int) + C * sizeof(long long);
                if (c->sched_cap < need) {
                    if (c->d_sched) { cudaStreamSynchronize(c->stream); dev_free(c->d_sched); }
                    c->sched_cap = 2 * need;
                    if (dev_malloc(&c->d_sched, c->sched_cap) != cudaSuccess) { c->d_sched = nullptr; c->sched_cap = 0; c->sched_C = -1; return cudaErrorMemoryAllocation; }
                }
                char* base = (char*)c->d_sched;
                cudaMemcpyAsync(base, flat.data(), flat.size() * sizeof(int4), cudaMemcpyHostToDevice, c->stream);
                cudaMemcpyAsync(base + flat.size() * sizeof(int4) + C * sizeof(long long), cnt.data(),
                                G * sizeof(int), cudaMemcpyHostToDevice, c->stream);
                cudaStreamSynchronize(c->stream);       // host vectors go out of scope (first call of a shape only)
                c->sched_C = C; c->sched_S = S; c->sched_G = G;
                c->sched_max_items = max_items; c->sched_flat = flat.size();
            }
            const int max_items = c->sched_max_items;
            char* base = (char*)c->d_sched;
            int4* d_items = (int4*)base;
            long long* d_prog = (long long*)(base + c->sched_flat * sizeof(int4));
            int* d_cnt = (int*)(base + c->sched_flat * sizeof(int4) + C * sizeof(long long));
            fill_ll_kernel<<<(C + 255) / 256, 256, 0, c->stream>>>(d_prog, C, it0);
            c->launches++;
            kern<<<G, threads, smem, c->stream>>>(c->pb, it0, it1, keep, min_rate, chain_dev, lik_dev,
                                                  row_first, rows_local, d_items, d_cnt, max_items, d_prog);
        } else {
            kern<<<C, threads, smem, c->stream>>>(c->pb, it0, it1, keep, min_rate, chain_dev, lik_dev,
                                                  row_first, rows_local, nullptr, nullptr, 0, nullptr);
        }
    } else if (NE != 0 && c->use_nc) {
        c->last_kernel = "sweep_seq_nc_kernel";
        cudaFuncSetAttribute(sweep_seq_nc_kernel<T, IV>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->sweep_smem);
        sweep_seq_nc_kernel<T, IV><<<c->pb.n_chains, c->threads, c->sweep_smem, c->stream>>>(
            c->pb, it0, it1, keep, min_rate, chain_dev, lik_dev, row_first, rows_local);
    } else if (NE == 0) {
        c->last_kernel = "sweep_seq_generic_kernel";
        cudaFuncSetAttribute(sweep_seq_generic_kernel<T, IV>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->sweep_smem);
        sweep_seq_generic_kernel<T, IV><<<c->pb.n_chains, c->generic_threads, c->sweep_smem, c->stream>>>(
            c->pb, it0, it1, keep, min_rate, chain_dev, lik_dev, row_first, rows_local);
    } else {
        const int ne = NE ? NE : 7;    // (never instantiates the row kernel with 0 rows)
        c->last_kernel = "sweep_seq_kernel";
        cudaFuncSetAttribute(sweep_seq_kernel<T, IV, ne>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->sweep_smem);
        sweep_seq_kernel<T, IV, ne><<<c->pb.n_chains, c->threads, c->sweep_smem, c->stream>>>(
            c->pb, it0, it1, keep, min_rate, chain_dev, lik_dev, row_first, rows_local);
    }
    c->launches++;
    // the pipelined kernel leaves its profile cache in step with the parameter map; any other
    // kernel moves the parameters without it
    c->pb.lu_valid = (strcmp(c->last_kernel, "sweep_seq_pipe_kernel") == 0 && c->pb.lucache && it1 > it0) ? 1 : 0;
    return cudaGetLastError();
}

template <typename T, bool IV, int NE>
static cudaError_t launch_colour_class(d3d_ctx* c, long long it, int cy, int cx, double* chain_dev,
                                       double* lik_dev, long long rows_local, long long row_local,
                                       bool pdl = false) {
    const Problem& pb = c->pb;
    const int ne = NE ? NE : 7;
    if (!c->colour_attr_set) {
        if (NE == 0)
            cudaFuncSetAttribute(sweep_colour_generic_kernel<T, IV>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->sweep_smem);
        else
            cudaFuncSetAttribute(sweep_colour_kernel<T, IV, ne>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->sweep_smem);
        c->colour_attr_set = true;
    }
    const int nly = (pb.H + pb.fh - 1) / pb.fh, nlx = (pb.W + pb.fw - 1) / pb.fw;
    dim3 grid(nly * nlx, pb.n_chains);
    c->last_kernel = NE == 0 ? (c->cluster ? "sweep_colour_cluster_kernel" : "sweep_colour_generic_kernel")
                             : "sweep_colour_kernel";
    if (NE == 0 && c->cluster) {
        if (!c->cluster_attr_set) {
            cudaFuncSetAttribute(sweep_colour_cluster_kernel<T, IV>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->sweep_smem);
            c->cluster_attr_set = true;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(nly * nlx * c->cluster, pb.n_chains);
        cfg.blockDim = dim3(320);
        cfg.dynamicSmemBytes = c->sweep_smem;
        cfg.stream = c->stream;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = c->cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        // Programmatic dependent launch between the colour phases of one sweep: the CTAs of phase
        // p+1 are scheduled while phase p drains, set up their constants, proposal and line profiles
        // (nothing a neighbouring phase writes) and wait (`griddepcontrol.wait`) right before their
        // first access to the residual.  Only behind another phase kernel (`pdl`).
        at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = pdl ? 2 : 1;
        c->launches++;
        return cudaLaunchKernelEx(&cfg, sweep_colour_cluster_kernel<T, IV>, pb, it, cy, cx, nlx, chain_dev,
                                  lik_dev, rows_local, row_local);
    }
    if (pdl) {                                   // behind another phase of the same sweep (see above)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(NE == 0 ? c->generic_threads : c->threads);
        cfg.dynamicSmemBytes = c->sweep_smem;
        cfg.stream = c->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        c->launches++;
        if (NE == 0)
            return cudaLaunchKernelEx(&cfg, sweep_colour_generic_kernel<T, IV>, pb, it, cy, cx, nlx, chain_dev,
                                      lik_dev, rows_local, row_local);
        return cudaLaunchKernelEx(&cfg, sweep_colour_kernel<T, IV, ne>, pb, it, cy, cx, nlx, chain_dev,
                                  lik_dev, rows_local, row_local);
    }
    if (NE == 0)
        sweep_colour_generic_kernel<T, IV><<<grid, c->generic_threads, c->sweep_smem, c->stream>>>(
            pb, it, cy, cx, nlx, chain_dev, lik_dev, rows_local, row_local);
    else
        sweep_colour_kernel<T, IV, ne><<<grid, c->threads, c->sweep_smem, c->stream>>>(
            pb, it, cy, cx, nlx, chain_dev, lik_dev, rows_local, row_local);
    c->launches++;
    return cudaGetLastError();
}

template <typename T, bool IV, int NE>
static cudaError_t launch_colour(d3d_ctx* c, long long it, double* chain_dev, double* lik_dev,
                                 long long rows_local, long long row_local) {
    const Problem& pb = c->pb;
    cudaError_t e = cudaSuccess;
    const bool pdl_ok = !getenv("D3D_NO_PDL");
    bool behind_phase = false;                       // the previous launch on the stream was a phase of this sweep
    for (int cy = 0; cy < pb.fh && e == cudaSuccess; ++cy)
        for (int cx = 0; cx < pb.fw && e == cudaSuccess; ++cx) {
            if (cy >= pb.H || cx >= pb.W) continue;
            e = launch_colour_class<T, IV, NE>(c, it, cy, cx, chain_dev, lik_dev, rows_local, row_local,
                                               pdl_ok && behind_phase);
            behind_phase = true;
        }
    return e;
}

#define DISPATCH_NE(FN, T, IV, ...)                                         \
    (c->ne == 7    ? FN<T, IV, 7>(__VA_ARGS__)                              \
     : c->ne == 13 ? FN<T, IV, 13>(__VA_ARGS__)                             \
     : c->ne == 21 ? FN<T, IV, 21>(__VA_ARGS__)                             \
                   : FN<T, IV, 0>(__VA_ARGS__))
#define DISPATCH(FN, ...)                                                                   \
    (c->dtype == D3D_F64                                                                    \
         ? (c->pb.var_is_cube ? DISPATCH_NE(FN, double, true, __VA_ARGS__)                  \
                              : DISPATCH_NE(FN, double, false, __VA_ARGS__))                \
         : (c->pb.var_is_cube ? DISPATCH_NE(FN, float, true, __VA_ARGS__)                   \
                              : DISPATCH_NE(FN, float, false, __VA_ARGS__)))

static bool is_device_ptr(const void* p);

extern "C" int d3d_sweep(d3d_ctx* c, int64_t first_iteration, int64_t n_iterations, int mode,
                         int keep_one_in, int refresh_every, double min_acceptance_rate,
                         double* chain_out, double* lik_out, int64_t n_rows,
                         int64_t* accepted_out, int64_t* iterations_out, float* elapsed_ms) {
    if (!c || !c->have_problem || !c->have_params)
        return fail(D3D_ESTATE, "d3d_sweep needs d3d_set_problem and parameters (+ d3d_forward)");
    if (!c->have_tables) return fail(D3D_ESTATE, "d3d_sweep needs d3d_set_rtnorm_tables");
    if (keep_one_in < 1) return fail(D3D_EINVAL, "keep_one_in= MUST be a positive integer");   // :112
    if (first_iteration < 1 || n_iterations < 0) return fail(D3D_EINVAL, "bad iteration range");
    if (mode != D3D_SEQ_EXACT && mode != D3D_COLOURED) return fail(D3D_EINVAL, "bad mode");
    if (first_iteration + n_iterations > 0xffffffffLL) return fail(D3D_EINVAL, "iteration counter exceeds 32 bits");
    CK(cudaSetDevice(c->device));
    const Problem& pb = c->pb;
    const size_t HW = (size_t)pb.H * pb.W;
    const long long it_begin = first_iteration, it_end = first_iteration + n_iterations;

    // rows written by this call: it / keep for it in [it_begin, it_end) with it % keep == 0
    long long row_first = (it_begin + keep_one_in - 1) / keep_one_in;
    long long row_last = (it_end - 1) / keep_one_in;            // inclusive
    long long rows_local = n_iterations > 0 && row_last >= row_first ? row_last - row_first + 1 : 0;
    if ((chain_out || lik_out) && rows_local > 0 && row_last >= n_rows)
        return fail(D3D_EINVAL, "chain_out/lik_out have %lld rows, iteration %lld needs row %lld",
                    (long long)n_rows, (long long)(row_last * keep_one_in), row_last);
    double* chain_dev = nullptr; double* lik_dev = nullptr;
    int rc = 0;
    const bool timing = getenv("D3D_TIMING") != nullptr;       // host-side stage times on stderr
    auto t_host = std::chrono::steady_clock::now();
    auto stamp = [&](const char* what) {
        if (!timing) return;
        auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "[d3d_sweep] %-22s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(n - t_host).count());
        t_host = n;
    };
    if (chain_out && rows_local > 0) {
        if (dev_malloc(&chain_dev, (size_t)pb.n_chains * rows_local * HW * 3 * sizeof(double)) != cudaSuccess)
            return fail(D3D_ENOMEM, "Not enough device memory for that many iterations. Use a higher value in the keep_one_in= parameter.");
        cudaMemsetAsync(chain_dev, 0, (size_t)pb.n_chains * rows_local * HW * 3 * sizeof(double), c->stream);
    }
    if (lik_out && rows_local > 0) {
        if (dev_malloc(&lik_dev, (size_t)pb.n_chains * rows_local * HW * sizeof(double)) != cudaSuccess) {
            if (chain_dev) dev_free(chain_dev);
            return fail(D3D_ENOMEM, "Not enough device memory for the likelihood chain.");
        }
        cudaMemsetAsync(lik_dev, 0, (size_t)pb.n_chains * rows_local * HW * sizeof(double), c->stream);
    }

    bool colour_by_chain = false;
    if (mode == D3D_COLOURED) {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
        const bool whole = pb.ty0 == 0 && pb.tx0 == 0 && pb.ty1 == pb.H && pb.tx1 == pb.W;
        colour_by_chain = whole && pb.n_chains >= (2 * sms) / 3;
        if (const char* ev = getenv("D3D_COLOUR_BY_CHAIN")) colour_by_chain = whole && atoi(ev) != 0;
    }
    stamp("staging alloc");
    // rows of [r0, r1) of every chain from the staging arrays to the caller's arrays: one strided
    // copy per array (row block of chain k -> rows [r0, r1) of chain k)
    auto copy_rows = [&](long long r0, long long r1, cudaStream_t st) {
        if (rc || r1 <= r0) return;
        cudaError_t ce = cudaSuccess;
        if (chain_dev) {
            const size_t rb = HW * 3 * sizeof(double);
            ce = cudaMemcpy2DAsync(chain_out + (size_t)r0 * HW * 3, (size_t)n_rows * rb,
                                   chain_dev + (size_t)(r0 - row_first) * HW * 3, (size_t)rows_local * rb,
                                   (size_t)(r1 - r0) * rb, pb.n_chains, cudaMemcpyDefault, st);
            if (ce != cudaSuccess) rc = fail(D3D_ECUDA, "chain copy failed: %s", cudaGetErrorString(ce));
        }
        if (lik_dev && !rc) {
            const size_t rb = HW * sizeof(double);
            ce = cudaMemcpy2DAsync(lik_out + (size_t)r0 * HW, (size_t)n_rows * rb,
                                   lik_dev + (size_t)(r0 - row_first) * HW, (size_t)rows_local * rb,
                                   (size_t)(r1 - r0) * rb, pb.n_chains, cudaMemcpyDefault, st);
            if (ce != cudaSuccess) rc = fail(D3D_ECUDA, "likelihood copy failed: %s", cudaGetErrorString(ce));
        }
    };
    // Many rows going to HOST memory (the reference's default keep_one_in=1: 272 MB per call of the
    // benched workload): the sweeps are launched in up to four chunks and the rows of a chunk travel
    // on a second stream while the next chunk computes.  A copy to pageable memory blocks the host,
    // so chunk i+1 is always enqueued before the copy of chunk i is started.
    long long chunk = 0;
    {
        const bool to_host = (chain_dev && !is_device_ptr(chain_out)) || (lik_dev && !is_device_ptr(lik_out));
        const double row_bytes = (double)pb.n_chains * rows_local * HW * sizeof(double) *
                                 ((chain_dev ? 3 : 0) + (lik_dev ? 1 : 0));
        if (to_host && mode == D3D_SEQ_EXACT && row_bytes > 64e6 && n_iterations >= 8 && !getenv("D3D_NO_COPY_OVERLAP")) {
            chunk = (n_iterations + 3) / 4;
            if (!c->copy_stream && cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
                cudaGetLastError(); chunk = 0;
            }
        }
    }
    cudaEvent_t seg_done = nullptr;                 // end of the latest chunk whose rows still wait
    long long pend_r0 = 0, pend_r1 = 0;             // its rows
    if (chunk && cudaEventCreateWithFlags(&seg_done, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); chunk = 0; }
    cudaEventRecord(c->ev0, c->stream);
    cudaError_t e = cudaSuccess;
    long long it = it_begin;
    while (it < it_end && e == cudaSuccess && !rc) {
        // segment ends right after an iteration with it % refresh_every == 0 (lib/run.py:525)
        long long seg_end = it_end;
        if (refresh_every > 0) {
            long long next_refresh = ((it + refresh_every - 1) / refresh_every) * refresh_every;
            seg_end = std::min(it_end, next_refresh + 1);
        }
        if (chunk) seg_end = std::min(seg_end, it + chunk);
        if (mode == D3D_SEQ_EXACT) {
            e = DISPATCH(launch_seq, c, it, seg_end, keep_one_in, min_acceptance_rate, chain_dev,
                         lik_dev, row_first, rows_local);
        } else if (colour_by_chain) {
            // many chains: one CTA per chain walks the sites in colour-class order -- the same
            // result as one launch per class (the sites of a class do not interact), without
            // the fh*fw launches and with every SM busy on its own chain.  The sliding-window
            // kernel takes this order too (every site reloads its columns, but its pipelined
            // scalar warps make it 1.7x faster than the row-mapped kernel here).
            const int* row_major = c->pb.sites;
            c->pb.sites = c->d_sites_colour;
            e = DISPATCH(launch_seq, c, it, seg_end, keep_one_in, min_acceptance_rate, chain_dev,
                         lik_dev, row_first, rows_local);
            c->pb.sites = row_major;
        } else {
            c->pb.lu_valid = 0;
            for (long long k = it; k < seg_end && e == cudaSuccess; ++k) {
                sweep_begin_kernel<<<(pb.n_chains + 127) / 128, 128, 0, c->stream>>>(pb, k, min_acceptance_rate);
                c->launches++;
                const bool save = (k % keep_one_in) == 0;
                e = DISPATCH(launch_colour, c, k, save ? chain_dev : nullptr, save ? lik_dev : nullptr,
                             rows_local, save ? k / keep_one_in - row_first : 0);
            }
        }
        if (e == cudaSuccess && refresh_every > 0 && (seg_end - 1) % refresh_every == 0)
            rc = run_forward(c, pb.params, 1, nullptr, 1, nullptr);
        if (chunk && e == cudaSuccess && !rc) {
            // this chunk is enqueued: the rows of the previous one may go (the host blocks in that
            // copy while the GPU works on this chunk), then remember this chunk's rows
            if (pend_r1 > pend_r0) {
                cudaStreamWaitEvent(c->copy_stream, seg_done, 0);
                copy_rows(pend_r0, pend_r1, c->copy_stream);
                cudaStreamSynchronize(c->copy_stream);      // (the event is re-used below)
            }
            cudaEventRecord(seg_done, c->stream);
            pend_r0 = (it + keep_one_in - 1) / keep_one_in;
            pend_r1 = (seg_end - 1) / keep_one_in + 1;
        }
        it = seg_end;
    }
    cudaEventRecord(c->ev1, c->stream);
    if (e != cudaSuccess && !rc) rc = fail(D3D_ECUDA, "sweep launch failed: %s", cudaGetErrorString(e));
    stamp("launches enqueued");
    if (timing) { cudaStreamSynchronize(c->stream); stamp("kernels done"); }

    if (!rc && rows_local > 0) {
        if (chunk) copy_rows(pend_r0, pend_r1, c->stream);               // the last chunk's rows
        else copy_rows(row_first, row_first + rows_local, c->stream);
    }
    if (seg_done) cudaEventDestroy(seg_done);
    std::vector<long long> h_acc(pb.n_chains), h_it(pb.n_chains);
    int h_status = 0;
    if (!rc) {
        cudaMemcpyAsync(h_acc.data(), pb.accepted, pb.n_chains * sizeof(long long), cudaMemcpyDeviceToHost, c->stream);
        cudaMemcpyAsync(h_it.data(), pb.iters, pb.n_chains * sizeof(long long), cudaMemcpyDeviceToHost, c->stream);
        cudaMemcpyAsync(&h_status, pb.status, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
    }
    stamp("row copies enqueued");
    e = cudaStreamSynchronize(c->stream);
    stamp("row copies done");
    if (chain_dev) dev_free(chain_dev);
    if (lik_dev) dev_free(lik_dev);
    stamp("staging free");
    if (rc) return rc;
    if (e != cudaSuccess) return fail(D3D_ECUDA, "sweep failed: %s", cudaGetErrorString(e));
    if (elapsed_ms) cudaEventElapsedTime(elapsed_ms, c->ev0, c->ev1);
    long long updates = 0;
    for (int k = 0; k < pb.n_chains; ++k) {
        if (accepted_out) accepted_out[k] = h_acc[k];
        if (iterations_out) iterations_out[k] = h_it[k];
        // (a chain stopped by the acceptance-rate test in an earlier call has h_it < it_begin)
        updates += std::max(0LL, h_it[k] - it_begin) * (long long)c->h_nsites[k / pb.chains_per_cube];
    }
    c->last_updates = updates;
    c->last_bytes = (int64_t)((pb.var_is_cube ? 3 : 2) * (double)c->elem() *
                              (double)c->window_voxels_per_sweep * (double)n_iterations);
    if (h_status & 0x40000000) {
        int prog[32] = {0};
        cudaMemcpy(prog, pb.dbg, sizeof prog, cudaMemcpyDeviceToHost);
        char buf[400]; int o = 0;
        for (int q = 0; q < 24; ++q) o += snprintf(buf + o, sizeof buf - o, " %d", prog[q]);
        return fail(D3D_ECUDA, "the pipelined sweep kernel stalled on an internal barrier (wait code %d at list "
                               "entry %d; list index reached by warp:%s) and gave up; results of this call are "
                               "invalid (D3D_PIPE=0 selects the sliding-window kernel)", h_status & 0xff,
                    (h_status & 0x3fffffff) >> 8, buf);
    }
    if (h_status)
        return fail(D3D_ENUMERIC, "cannot convert float NaN to integer: a NaN reached the truncated-normal sampler "
                                  "(lib/rtnorm.py:144) or a rejection loop exceeded its guard");
    return 0;
}

#ifdef D3D_TRACE
extern "C" int d3d_debug_trace2(unsigned long long* out) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, d3d::g_tr2, 16 * 16 * 16 * sizeof(unsigned long long));
    return 0;
}
#endif

#ifdef D3D_PIPE_PROF
extern "C" int d3d_debug_pipe_prof(unsigned long long* out512, int reset) {
    cudaDeviceSynchronize();
    if (out512) cudaMemcpyFromSymbol(out512, d3d::g_pipe_prof, 32 * 16 * sizeof(unsigned long long));
    if (reset) { static unsigned long long z[32 * 16]; cudaMemcpyToSymbol(d3d::g_pipe_prof, z, sizeof z); }
    return 0;
}
extern "C" int d3d_debug_pipe_cta(unsigned long long* out2048) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out2048, d3d::g_pipe_cta, 2048 * sizeof(unsigned long long));
    return 0;
}
#endif

#ifdef D3D_PHASE_TIMING
extern "C" int d3d_debug_phases(unsigned long long* out16, int reset) {
    cudaDeviceSynchronize();
    if (out16) cudaMemcpyFromSymbol(out16, d3d::g_phase, 32 * sizeof(unsigned long long));
    if (reset) { unsigned long long z[32] = {0}; cudaMemcpyToSymbol(d3d::g_phase, z, sizeof z); }
    return 0;
}
#endif

// ------------------------------------------------------------------------------
// Spatially tiled coloured sweep of ONE cube over several contexts
// ------------------------------------------------------------------------------
static bool is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

static int rec_stage(d3d_ctx* c, size_t bytes) {
    if (bytes <= c->rec_stage_cap) return 0;
    if (c->d_rec_stage) dev_free(c->d_rec_stage);
    c->d_rec_stage = nullptr; c->rec_stage_cap = 0;
    if (dev_malloc(&c->d_rec_stage, bytes) != cudaSuccess)
        return fail(D3D_ENOMEM, "dev_malloc(%zu bytes) for the record staging buffer failed", bytes);
    c->rec_stage_cap = bytes;
    return 0;
}

extern "C" int d3d_set_tile(d3d_ctx* c, int y0, int y1, int x0, int x1) {
    if (!c || !c->have_problem) return fail(D3D_ESTATE, "d3d_set_tile before d3d_set_problem");
    Problem& pb = c->pb;
    if (y0 < 0 || x0 < 0 || y1 > pb.H || x1 > pb.W || y0 > y1 || x0 > x1)
        return fail(D3D_EINVAL, "d3d_set_tile: tile [%d,%d)x[%d,%d) outside the %dx%d field", y0, y1, x0, x1, pb.H, pb.W);
    CK(cudaSetDevice(c->device));
    if (!pb.lik_cur) {
        const size_t n = (size_t)pb.n_chains * pb.H * pb.W;
        int rc;
        if ((rc = dalloc(c, &pb.lik_cur, n * sizeof(double)))) return rc;
        if ((rc = dalloc(c, &pb.acc_cur, n))) return rc;
        CK(cudaMemsetAsync(pb.lik_cur, 0, n * sizeof(double), c->stream));
        CK(cudaMemsetAsync(pb.acc_cur, 0, n, c->stream));
    }
    pb.ty0 = y0; pb.ty1 = y1; pb.tx0 = x0; pb.tx1 = x1;
    pb.ry0 = std::max(y0 - pb.fhh, 0); pb.ry1 = std::min(y1 + pb.fhh, pb.H);
    pb.rx0 = std::max(x0 - pb.fhw, 0); pb.rx1 = std::min(x1 + pb.fhw, pb.W);
    return 0;
}

extern "C" int d3d_tile_record_slots(d3d_ctx* c, int64_t* n_records) {
    if (!c || !c->have_problem) return fail(D3D_ESTATE, "d3d_tile_record_slots before d3d_set_problem");
    const Problem& pb = c->pb;
    const int nly = (pb.H + pb.fh - 1) / pb.fh, nlx = (pb.W + pb.fw - 1) / pb.fw;
    if (n_records) *n_records = (int64_t)pb.n_chains * nly * nlx;
    return 0;
}

extern "C" int d3d_colour_begin(d3d_ctx* c, int64_t iteration, double min_acceptance_rate) {
    if (c) c->pb.lu_valid = 0;            // (profile cache of the pipelined sweep: parameters move)
    if (!c || !c->have_problem || !c->have_params) return fail(D3D_ESTATE, "d3d_colour_begin needs a problem and parameters");
    CK(cudaSetDevice(c->device));
    const Problem& pb = c->pb;
    sweep_begin_kernel<<<(pb.n_chains + 127) / 128, 128, 0, c->stream>>>(pb, iteration, min_acceptance_rate);
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int d3d_colour_phase(d3d_ctx* c, int64_t iteration, int cy, int cx, double* records_out) {
    if (c) c->pb.lu_valid = 0;            // (profile cache of the pipelined sweep: parameters move)
    if (!c || !c->have_problem || !c->have_params)
        return fail(D3D_ESTATE, "d3d_colour_phase needs d3d_set_problem and parameters (+ d3d_forward)");
    if (!c->have_tables) return fail(D3D_ESTATE, "d3d_colour_phase needs d3d_set_rtnorm_tables");
    const Problem& pb = c->pb;
    if (!pb.lik_cur) return fail(D3D_ESTATE, "d3d_colour_phase needs d3d_set_tile");
    if (cy < 0 || cx < 0 || cy >= pb.fh || cx >= pb.fw) return fail(D3D_EINVAL, "colour class (%d,%d) outside the %dx%d FSF lattice", cy, cx, pb.fh, pb.fw);
    if (iteration < 1 || iteration > 0xffffffffLL) return fail(D3D_EINVAL, "bad iteration");
    CK(cudaSetDevice(c->device));
    const int nly = (pb.H + pb.fh - 1) / pb.fh, nlx = (pb.W + pb.fw - 1) / pb.fw;
    const int n = pb.n_chains * nly * nlx;
    if (cy < pb.H && cx < pb.W) {
        cudaError_t e = DISPATCH(launch_colour_class, c, (long long)iteration, cy, cx, nullptr, pb.lik_cur, 1LL, 0LL);
        if (e != cudaSuccess) return fail(D3D_ECUDA, "colour phase launch failed: %s", cudaGetErrorString(e));
    }
    if (records_out) {
        double* dst = records_out;
        const bool dev = is_device_ptr(records_out);
        if (!dev) { int rc = rec_stage(c, (size_t)n * REC_N * sizeof(double)); if (rc) return rc; dst = c->d_rec_stage; }
        pack_records_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(pb, cy, cx, nly, nlx, dst);
        c->launches++;
        CK(cudaGetLastError());
        if (!dev) {
            CK(cudaMemcpyAsync(records_out, dst, (size_t)n * REC_N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
        }
    }
    return 0;
}

extern "C" int d3d_apply_records(d3d_ctx* c, const double* records, int64_t n_records) {
    if (c) c->pb.lu_valid = 0;            // (profile cache of the pipelined sweep: parameters move)
    if (!c || !c->have_problem || !c->have_params) return fail(D3D_ESTATE, "d3d_apply_records needs a problem and parameters");
    const Problem& pb = c->pb;
    if (!pb.lik_cur) return fail(D3D_ESTATE, "d3d_apply_records needs d3d_set_tile");
    if (n_records < 0 || (n_records && !records)) return fail(D3D_EINVAL, "d3d_apply_records: bad arguments");
    if (n_records == 0) return 0;
    if (n_records > 0x7fffffffLL) return fail(D3D_EINVAL, "too many records");
    CK(cudaSetDevice(c->device));
    const double* src = records;
    if (!is_device_ptr(records)) {
        int rc = rec_stage(c, (size_t)n_records * REC_N * sizeof(double)); if (rc) return rc;
        CK(cudaMemcpyAsync(c->d_rec_stage, records, (size_t)n_records * REC_N * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        src = c->d_rec_stage;
    }
    if (!c->apply_attr_set) {
        cudaFuncSetAttribute(apply_records_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->sweep_smem);
        cudaFuncSetAttribute(apply_records_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->sweep_smem);
        c->apply_attr_set = true;
    }
    if (c->cluster) {
        if (!c->apply_cluster_attr_set) {
            cudaFuncSetAttribute(apply_records_cluster_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->sweep_smem);
            cudaFuncSetAttribute(apply_records_cluster_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->sweep_smem);
            c->apply_cluster_attr_set = true;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)n_records * c->cluster);
        cfg.blockDim = dim3(256);
        cfg.dynamicSmemBytes = c->sweep_smem;
        cfg.stream = c->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = c->cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        c->launches++;
        if (c->dtype == D3D_F64) CK(cudaLaunchKernelEx(&cfg, apply_records_cluster_kernel<double>, pb, src, (int)n_records));
        else CK(cudaLaunchKernelEx(&cfg, apply_records_cluster_kernel<float>, pb, src, (int)n_records));
        return 0;
    }
    if (c->dtype == D3D_F64)
        apply_records_kernel<double><<<(unsigned)n_records, 256, c->sweep_smem, c->stream>>>(pb, src, (int)n_records);
    else
        apply_records_kernel<float><<<(unsigned)n_records, 256, c->sweep_smem, c->stream>>>(pb, src, (int)n_records);
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

// ---- fused exchange over peer memory (no NCCL inside the phase) ---------------------------
static size_t box_flag_bytes() { return 256; }      // TILE_MAXW uint64 flags + the block counter, padded

extern "C" int d3d_tile_fused_init(d3d_ctx* c, int n_tiles, int my_index, void** box_out, int64_t* box_bytes) {
    if (!c || !c->have_problem) return fail(D3D_ESTATE, "d3d_tile_fused_init before d3d_set_problem");
    if (!c->pb.lik_cur) return fail(D3D_ESTATE, "d3d_tile_fused_init needs d3d_set_tile");
    if (n_tiles < 1 || n_tiles > TILE_MAXW || my_index < 0 || my_index >= n_tiles)
        return fail(D3D_EINVAL, "d3d_tile_fused_init: 1 <= n_tiles <= %d and 0 <= my_index < n_tiles", (int)TILE_MAXW);
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    const Problem& pb = c->pb;
    const int nly = (pb.H + pb.fh - 1) / pb.fh, nlx = (pb.W + pb.fw - 1) / pb.fw;
    const long long slots = (long long)pb.n_chains * nly * nlx;
    const size_t bytes = box_flag_bytes() + (size_t)2 * n_tiles * slots * REC_N * sizeof(double);
    for (void* p : c->ipc_opened) cudaIpcCloseMemHandle(p);
    c->ipc_opened.clear();
    if (c->box && c->box_bytes < bytes) { cudaFree(c->box); c->box = nullptr; }
    if (!c->box) {
        // a plain driver allocation (not from the cache): its handle is exported to other processes
        if (cudaMalloc(&c->box, bytes) != cudaSuccess) { c->box = nullptr; return fail(D3D_ENOMEM, "dev_malloc(%zu bytes) for the tile box failed", bytes); }
        c->box_bytes = bytes;
    }
    CK(cudaMemsetAsync(c->box, 0, c->box_bytes, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    memset(&c->tb, 0, sizeof c->tb);
    c->tb.n_tiles = n_tiles; c->tb.my_tile = my_index; c->tb.slots = slots;
    c->tb.flags[my_index] = (unsigned long long*)c->box;
    c->tb.inbox[my_index] = (double*)((char*)c->box + box_flag_bytes());
    c->tb.done_counter = (unsigned int*)((char*)c->box + TILE_MAXW * sizeof(unsigned long long));
    {   // bound of the flag wait: long enough for a healthy but slow peer (Python-driven ranks
        // pause on parameter copies, lazy module loading); D3D_TILE_TIMEOUT_S overrides
        double secs = 30.0;
        if (const char* e = getenv("D3D_TILE_TIMEOUT_S")) secs = atof(e) > 0 ? atof(e) : secs;
        c->tb.timeout_cycles = (long long)(secs * 2.0e9);
    }
    CK(cudaMemset(c->pb.status, 0, sizeof(int)));            // a fresh exchange forgets an earlier time-out
    {   // hit lists of the triage pass, one per phase parity.  Bound: lattice sites of one colour class
        // whose window can reach into the region this context keeps valid.
        const long long ny = (pb.ry1 - pb.ry0 + 2 * pb.fhh) / pb.fh + 2, nx = (pb.rx1 - pb.rx0 + 2 * pb.fhw) / pb.fw + 2;
        c->hits_stride = std::min<long long>((long long)n_tiles * slots, (long long)pb.n_chains * ny * nx);
        if (c->d_hits) { cudaFree(c->d_hits); c->d_hits = nullptr; }
        if (!c->d_hit_count) CK(cudaMalloc(&c->d_hit_count, 2 * sizeof(unsigned int)));
        CK(cudaMalloc(&c->d_hits, (size_t)2 * c->hits_stride * sizeof(int)));
        CK(cudaMemset(c->d_hit_count, 0, 2 * sizeof(unsigned int)));
    }
    if (box_out) *box_out = c->box;
    if (box_bytes) *box_bytes = (int64_t)bytes;
    return 0;
}

extern "C" int d3d_tile_fused_export(d3d_ctx* c, unsigned char* handle64) {
    if (!c || !c->box) return fail(D3D_ESTATE, "d3d_tile_fused_export before d3d_tile_fused_init");
    if (!handle64) return fail(D3D_EINVAL, "handle is NULL");
    CK(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, c->box));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handle64, &h, 64);
    return 0;
}

extern "C" int d3d_tile_fused_connect(d3d_ctx* c, int index, void* peer_box, const unsigned char* handle64) {
    if (!c || !c->box) return fail(D3D_ESTATE, "d3d_tile_fused_connect before d3d_tile_fused_init");
    if (index < 0 || index >= c->tb.n_tiles) return fail(D3D_EINVAL, "tile index %d out of range", index);
    if (index == c->tb.my_tile) return 0;
    CK(cudaSetDevice(c->device));
    void* base = peer_box;
    if (!base) {
        if (!handle64) return fail(D3D_EINVAL, "d3d_tile_fused_connect needs a device pointer or an IPC handle");
        cudaIpcMemHandle_t h;
        memcpy(&h, handle64, 64);
        CK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
        c->ipc_opened.push_back(base);
    } else {
        // same process: make sure this device may store into the peer's memory
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, base) == cudaSuccess && a.type == cudaMemoryTypeDevice && a.device != c->device) {
            cudaError_t e = cudaDeviceEnablePeerAccess(a.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return fail(D3D_ECUDA, "no peer access from device %d to device %d: %s", c->device, a.device, cudaGetErrorString(e));
            cudaGetLastError();
        }
    }
    c->tb.flags[index] = (unsigned long long*)base;
    c->tb.inbox[index] = (double*)((char*)base + box_flag_bytes());
    return 0;
}

extern "C" int d3d_colour_phase_fused(d3d_ctx* c, int64_t iteration, int cy, int cx, int64_t phase_index) {
    if (c) c->pb.lu_valid = 0;            // (profile cache of the pipelined sweep: parameters move)
    if (!c || !c->have_problem || !c->have_params || !c->have_tables)
        return fail(D3D_ESTATE, "d3d_colour_phase_fused needs a problem, parameters and the rtnorm tables");
    if (!c->box) return fail(D3D_ESTATE, "d3d_colour_phase_fused needs d3d_tile_fused_init");
    const Problem& pb = c->pb;
    for (int t = 0; t < c->tb.n_tiles; ++t)
        if (!c->tb.flags[t]) return fail(D3D_ESTATE, "tile %d is not connected (d3d_tile_fused_connect)", t);
    if (cy < 0 || cx < 0 || cy >= pb.fh || cx >= pb.fw) return fail(D3D_EINVAL, "colour class (%d,%d) outside the %dx%d FSF lattice", cy, cx, pb.fh, pb.fw);
    if (iteration < 1 || iteration > 0xffffffffLL || phase_index < 0) return fail(D3D_EINVAL, "bad iteration / phase index");
    CK(cudaSetDevice(c->device));
    const int nly = (pb.H + pb.fh - 1) / pb.fh, nlx = (pb.W + pb.fw - 1) / pb.fw;
    const int n = pb.n_chains * nly * nlx;
    if (cy < pb.H && cx < pb.W) {
        cudaError_t e = DISPATCH(launch_colour_class, c, (long long)iteration, cy, cx, nullptr, pb.lik_cur, 1LL, 0LL);
        if (e != cudaSuccess) return fail(D3D_ECUDA, "colour phase launch failed: %s", cudaGetErrorString(e));
    }
    push_records_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(pb, c->tb, cy, cx, nly, nlx, (unsigned long long)phase_index);
    c->launches++;
    CK(cudaGetLastError());
    if (c->tb.n_tiles == 1) return 0;
    if (!c->box_attr_set) {
        cudaFuncSetAttribute(apply_box_kernel<double, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->sweep_smem);
        cudaFuncSetAttribute(apply_box_kernel<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->sweep_smem);
        cudaFuncSetAttribute(apply_box_kernel<double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->sweep_smem);
        cudaFuncSetAttribute(apply_box_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->sweep_smem);
        c->box_attr_set = true;
    }
    long long n_rec = (long long)c->tb.n_tiles * n;
    const unsigned long long ph = (unsigned long long)phase_index;
    // triage first (one thread per record: flag wait, book-keeping, list of the records that reach
    // into this tile's region), then clusters only for that list; otherwise a cluster per slot of
    // every tile
    const int* hits = nullptr; const unsigned int* hit_count = nullptr;
    const long long hit_bound = std::min<long long>(
        n_rec, (long long)pb.n_chains * ((pb.ry1 - pb.ry0 + 2 * pb.fhh) / pb.fh + 2) * ((pb.rx1 - pb.rx0 + 2 * pb.fhw) / pb.fw + 2));
    // Worth a fourth launch per phase once the cluster-per-slot grid is more than a few waves (the
    // 256 x 256 cube on two GPUs -- 98 slots x 6 CTAs -- is faster without: 93 against 99 ms per sweep).
    // D3D_TILE_TRIAGE=1 / 0 forces / forbids it.
    bool triage = n_rec * std::max(c->cluster, 1) > 1024;
    if (const char* e = getenv("D3D_TILE_TRIAGE")) triage = atoi(e) != 0;
    if (getenv("D3D_TILE_NO_TRIAGE")) triage = false;
    if (triage && c->d_hits && hit_bound <= c->hits_stride) {   // (tile unchanged since the init)
        triage_box_kernel<<<(unsigned)((n_rec + 127) / 128), 128, 0, c->stream>>>(
            pb, c->tb, ph, c->d_hits, c->d_hit_count, c->hits_stride, (unsigned int)c->hits_stride);
        c->launches++;
        CK(cudaGetLastError());
        hits = c->d_hits; hit_count = c->d_hit_count;
        n_rec = c->hits_stride;
    }
    if (c->cluster) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(n_rec * c->cluster));
        cfg.blockDim = dim3(256);
        cfg.dynamicSmemBytes = c->sweep_smem;
        cfg.stream = c->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = c->cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (c->dtype == D3D_F64) CK(cudaLaunchKernelEx(&cfg, apply_box_kernel<double, true>, pb, c->tb, ph, hits, hit_count, c->hits_stride));
        else CK(cudaLaunchKernelEx(&cfg, apply_box_kernel<float, true>, pb, c->tb, ph, hits, hit_count, c->hits_stride));
    } else {
        if (c->dtype == D3D_F64) apply_box_kernel<double, false><<<(unsigned)n_rec, 256, c->sweep_smem, c->stream>>>(pb, c->tb, ph, hits, hit_count, c->hits_stride);
        else apply_box_kernel<float, false><<<(unsigned)n_rec, 256, c->sweep_smem, c->stream>>>(pb, c->tb, ph, hits, hit_count, c->hits_stride);
        CK(cudaGetLastError());
    }
    c->launches++;
    return 0;
}

extern "C" int d3d_sweep_fused(d3d_ctx* c, int64_t first_iteration, int64_t n_iterations, double min_acceptance_rate) {
    if (c) c->pb.lu_valid = 0;            // (profile cache of the pipelined sweep: parameters move)
    if (!c || !c->box) return fail(D3D_ESTATE, "d3d_sweep_fused needs d3d_tile_fused_init");
    if (first_iteration < 1 || n_iterations < 0) return fail(D3D_EINVAL, "bad iteration range");
    const Problem& pb = c->pb;
    const int ncy = std::min(pb.fh, pb.H), ncx = std::min(pb.fw, pb.W);
    for (int64_t it = first_iteration; it < first_iteration + n_iterations; ++it) {
        int rc = d3d_colour_begin(c, it, min_acceptance_rate);
        if (rc) return rc;
        for (int cy = 0; cy < ncy; ++cy)
            for (int cx = 0; cx < ncx; ++cx)
                if ((rc = d3d_colour_phase_fused(c, it, cy, cx, it * ncy * ncx + (int64_t)cy * ncx + cx))) return rc;
    }
    return 0;
}

extern "C" int d3d_get_likelihoods(d3d_ctx* c, double* lik_out) {
    if (!c || !c->have_problem) return fail(D3D_ESTATE, "d3d_get_likelihoods before d3d_set_problem");
    const Problem& pb = c->pb;
    if (!pb.lik_cur) return fail(D3D_ESTATE, "d3d_get_likelihoods needs d3d_set_tile");
    if (!lik_out) return fail(D3D_EINVAL, "lik_out is NULL");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(lik_out, pb.lik_cur, (size_t)pb.n_chains * pb.H * pb.W * sizeof(double), cudaMemcpyDefault, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int d3d_get_chain_control(d3d_ctx* c, int64_t* accepted_out, int64_t* iterations_out, int32_t* active_out) {
    if (!c || !c->have_problem) return fail(D3D_ESTATE, "d3d_get_chain_control before d3d_set_problem");
    const Problem& pb = c->pb;
    CK(cudaSetDevice(c->device));
    if (accepted_out) CK(cudaMemcpyAsync(accepted_out, pb.accepted, pb.n_chains * sizeof(long long), cudaMemcpyDefault, c->stream));
    if (iterations_out) CK(cudaMemcpyAsync(iterations_out, pb.iters, pb.n_chains * sizeof(long long), cudaMemcpyDefault, c->stream));
    if (active_out) CK(cudaMemcpyAsync(active_out, pb.active, pb.n_chains * sizeof(int), cudaMemcpyDefault, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    int h_status = 0;
    CK(cudaMemcpy(&h_status, pb.status, sizeof(int), cudaMemcpyDeviceToHost));
    if (h_status == 2)
        return fail(D3D_ECUDA, "tile exchange timed out: a peer never published its colour phase");
    if (h_status & 0x40000000) {
        int prog[32] = {0};
        cudaMemcpy(prog, pb.dbg, sizeof prog, cudaMemcpyDeviceToHost);
        char buf[400]; int o = 0;
        for (int q = 0; q < 24; ++q) o += snprintf(buf + o, sizeof buf - o, " %d", prog[q]);
        return fail(D3D_ECUDA, "the pipelined sweep kernel stalled on an internal barrier (wait code %d at list "
                               "entry %d; list index reached by warp:%s) and gave up; results of this call are "
                               "invalid (D3D_PIPE=0 selects the sliding-window kernel)", h_status & 0xff,
                    (h_status & 0x3fffffff) >> 8, buf);
    }
    if (h_status)
        return fail(D3D_ENUMERIC, "cannot convert float NaN to integer: a NaN reached the truncated-normal sampler "
                                  "(lib/rtnorm.py:144) or a rejection loop exceeded its guard");
    return 0;
}

extern "C" int d3d_chain_mean(d3d_ctx* c, const double* chain, int64_t n_rows, int64_t first_row,
                              double* mean_out) {
    if (!c || !c->have_problem) return fail(D3D_ESTATE, "d3d_chain_mean before d3d_set_problem");
    if (!chain || !mean_out) return fail(D3D_EINVAL, "d3d_chain_mean: NULL argument");
    if (n_rows < 1 || first_row < 0 || first_row >= n_rows) return fail(D3D_EINVAL, "d3d_chain_mean: bad row range");
    CK(cudaSetDevice(c->device));
    const Problem& pb = c->pb;
    const long long row_elems = (long long)pb.H * pb.W * 3;
    const size_t out_bytes = (size_t)pb.n_chains * row_elems * sizeof(double);
    const size_t in_bytes = (size_t)pb.n_chains * n_rows * row_elems * sizeof(double);
    const bool in_dev = is_device_ptr(chain), out_dev = is_device_ptr(mean_out);
    double* d_in = nullptr; double* d_out = nullptr;
    if (!in_dev) {
        if (dev_malloc(&d_in, in_bytes) != cudaSuccess) return fail(D3D_ENOMEM, "dev_malloc(%zu bytes) failed", in_bytes);
        cudaMemcpyAsync(d_in, chain, in_bytes, cudaMemcpyHostToDevice, c->stream);
    }
    if (!out_dev && dev_malloc(&d_out, out_bytes) != cudaSuccess) {
        if (d_in) dev_free(d_in);
        return fail(D3D_ENOMEM, "dev_malloc(%zu bytes) failed", out_bytes);
    }
    const long long n = (long long)pb.n_chains * row_elems;
    chain_mean_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(in_dev ? chain : d_in, pb.n_chains, n_rows,
                                                                           first_row, row_elems, out_dev ? mean_out : d_out);
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && !out_dev) e = cudaMemcpyAsync(mean_out, d_out, out_bytes, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (d_in) dev_free(d_in);
    if (d_out) dev_free(d_out);
    if (e != cudaSuccess) return fail(D3D_ECUDA, "d3d_chain_mean failed: %s", cudaGetErrorString(e));
    return 0;
}

__global__ void __launch_bounds__(1024) dfma_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

extern "C" int d3d_fp64_peak(d3d_ctx* c, double* tflops_out) {
    if (!c || !tflops_out) return fail(D3D_EINVAL, "d3d_fp64_peak: NULL argument");
    CK(cudaSetDevice(c->device));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    const int blocks = sms * 2, threads = 1024, iters = 2048;
    double* out = nullptr;
    CK(dev_malloc(&out, sizeof(double) * blocks * threads));
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(c->ev0, c->stream);
        dfma_peak_kernel<<<blocks, threads, 0, c->stream>>>(out, iters, 0.999999, 1e-6);
        cudaEventRecord(c->ev1, c->stream);
        cudaEventSynchronize(c->ev1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c->ev0, c->ev1);
        if (rep && ms < best) best = ms;
        c->launches++;
    }
    dev_free(out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(D3D_ECUDA, "d3d_fp64_peak failed: %s", cudaGetErrorString(e));
    *tflops_out = 2.0 * 64.0 * iters * (double)blocks * threads / best / 1e9;
    return 0;
}

extern "C" const char* d3d_last_kernel(d3d_ctx* c) { return c ? c->last_kernel : ""; }

extern "C" int d3d_get_counters(d3d_ctx* c, int64_t* kernel_launches, int64_t* last_sweep_bytes,
                                int64_t* last_sweep_site_updates) {
    if (!c) return fail(D3D_EINVAL, "ctx is NULL");
    if (kernel_launches) *kernel_launches = c->launches;
    if (last_sweep_bytes) *last_sweep_bytes = c->last_bytes;
    if (last_sweep_site_updates) *last_sweep_site_updates = c->last_updates;
    return 0;
}
