/*
 * deconv3d_b200.h -- C ABI of libdeconv3d_b200.so
 *
 * B200 (sm_100a) implementation of the per-iteration likelihood hot path of
 * irap-omp/deconv3d.  The reference is pure Python and has no FFI of its own
 * (SURVEY.md section 8b): its boundary for this path is the Python object API
 * of lib/run.py, lib/convolution.py, lib/line_models.py.  Each entry point
 * below names the reference interface it replaces (paths relative to the
 * reference tree); deconv3d_b200/_native.py is the ctypes binding and
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - Every function returns 0 on success or a negative D3D_E* code; the
 *     message is available from d3d_last_error() (thread-local).
 *   - All array arguments are plain pointers owned by the caller and are only
 *     borrowed for the duration of the call.  They may be host or device
 *     pointers (the library copies with cudaMemcpyDefault).
 *   - Host-visible arrays use the REFERENCE layouts and float64:
 *       cubes        [n][D][H][W]      (z, y, x; x fastest)  lib/run.py:146-149
 *       parameters   [n][H][W][3]      (a, c, w)             lib/run.py:270-272
 *       chain rows   [n][rows][H][W][3], likelihood rows [n][rows][H][W]
 *     Internally the residual lives z-fastest ([y][x][Dp]) in the storage
 *     dtype chosen at d3d_ctx_create.
 *   - A context is bound to one GPU and one stream; calls on one context must
 *     be serialised by the caller.  No global state.
 */
#ifndef DECONV3D_B200_H
#define DECONV3D_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define D3D_ABI_VERSION 2

/* storage dtype of the residual / data / inverse-variance cubes */
#define D3D_F32 0
#define D3D_F64 1

/* sweep modes */
#define D3D_SEQ_EXACT 0   /* row-major masked order of lib/run.py:553-566, one CTA per chain */
#define D3D_COLOURED  1   /* colour classes (y mod fh, x mod fw): disjoint windows per launch */

/* variance kinds for d3d_set_problem */
#define D3D_VAR_SCALAR 0  /* var[n_cubes]            (lib/run.py:186-192 default path) */
#define D3D_VAR_CUBE   1  /* var[n_cubes][D][H][W]   (lib/run.py:171-184)              */

/* error codes */
#define D3D_OK          0
#define D3D_EINVAL     -1  /* bad argument (maps to ValueError / AssertionError)       */
#define D3D_ECUDA      -2  /* CUDA runtime error                                        */
#define D3D_ESTATE     -3  /* call order violated (e.g. sweep before set_problem)       */
#define D3D_ENUMERIC   -4  /* NaN reached the truncated-normal sampler (the reference
                              raises ValueError at lib/rtnorm.py:144) or a rejection
                              loop exceeded its guard                                   */
#define D3D_ENOMEM     -5  /* device allocation failed (MemoryError, lib/run.py:273)    */

typedef struct d3d_ctx d3d_ctx;

/* ---- life cycle ----------------------------------------------------------- */
int d3d_abi_version(void);
const char* d3d_last_error(void);
int d3d_ctx_create(d3d_ctx** out, int device, int dtype);
int d3d_ctx_destroy(d3d_ctx* ctx);
/* Run all work of this context on an existing cudaStream_t (e.g. torch's
 * current stream, so that torch.cuda.Event brackets the kernels). NULL = the
 * context's own stream (default). */
int d3d_ctx_set_stream(d3d_ctx* ctx, void* cuda_stream);
int d3d_ctx_synchronize(d3d_ctx* ctx);

/* ---- problem set-up --------------------------------------------------------
 * Replaces the state built by Run.__init__ at lib/run.py:139-288: data cube(s),
 * variance (-> stored as 1/variance), spatial mask (-> row-major site list,
 * lib/run.py:553-566), FSF image and LSF vector (lib/run.py:208-211; lsf may be
 * NULL = no spectral convolution, lib/run.py:675-676), parameter boundaries
 * (lib/line_models.py:76-90), Cauchy jump amplitudes (lib/run.py:251-262) and
 * the Gibbs a-priori variance (lib/run.py:264-265).
 *   n_cubes            independent cubes ("galaxies"); each has chains_per_cube chains
 *   mask               uint8 [n_cubes][H][W] (1 = iterate), or NULL = all ones
 *   pmin, pmax         [n_cubes][3];   gibbs_prior_var [n_cubes];   jump_amp [3]
 * The reference's H*W full-cube `contributions` array (lib/run.py:285-288) is
 * not materialised: each contribution is rank-1 and recomputed from (a,c,w). */
int d3d_set_problem(d3d_ctx* ctx, int n_cubes, int chains_per_cube,
                    int D, int H, int W,
                    const double* data, const double* var, int var_kind,
                    const uint8_t* mask,
                    const double* fsf, int fh, int fw,
                    const double* lsf,
                    const double* pmin, const double* pmax,
                    const double* jump_amp, const double* gibbs_prior_var);

/* Truncated-normal tables of lib/rtnorm.py:227,1230,2233 (x[4002], yu[4001],
 * ncell[8961]); built on the host by deconv3d_b200/rtnorm_tables.py. */
int d3d_set_rtnorm_tables(d3d_ctx* ctx, const double* x, int nx,
                          const double* yu, int nyu, const int32_t* ncell, int nncell);

/* Counter-based random stream ("d3d stream v1", Philox4x32-10): replaces the
 * global numpy.random state of lib/run.py:313,435,578 and lib/rtnorm.py:17.
 * Chain k of the context draws from (seed, first_chain_id + k). */
int d3d_set_rng(d3d_ctx* ctx, uint64_t seed, uint32_t first_chain_id);

/* ---- chain state -----------------------------------------------------------
 * params: [n_chains][H][W][3] float64, n_chains = n_cubes*chains_per_cube,
 * chain k belongs to cube k / chains_per_cube.  set_params replaces
 * `self.chain[0] = initial_parameters` (lib/run.py:294-307);
 * d3d_init_params_uniform the random initialisation of lib/run.py:308-314
 * (sweep 0 of the stream).  Neither computes the residual: call d3d_forward. */
int d3d_set_params(d3d_ctx* ctx, const double* params);
int d3d_get_params(d3d_ctx* ctx, double* params);
int d3d_init_params_uniform(d3d_ctx* ctx);

/* ---- forward model ---------------------------------------------------------
 * Replaces Run._compute_error_in_one_step (lib/run.py:999-1031) and
 * Run.simulate_convolved (lib/run.py:623-652): per-spaxel Gaussian line
 * (lib/line_models.py:98-109) -> spectral convolution with the LSF with the
 * exact wrap-around of lib/convolution.py:89-160 -> 2-D true convolution with
 * the FSF, zero 'same' borders, sources at masked spaxels only.
 *   sim_out   [n_chains][D][H][W] float64 or NULL
 *   write_err non-zero: residual <- data - sim (lib/run.py:334, 525-534)
 *   chi2_out  [n_chains] float64 or NULL: 0.5*sum(err^2/var) over the cube     */
int d3d_forward(d3d_ctx* ctx, double* sim_out, int write_err, double* chi2_out);
/* Same, for n_sets explicit parameter maps [n_sets][H][W][3] that are not the chain
 * state (Run.simulate_convolved(shape, parameters)); set k uses the cube of chain slot k.
 * Never touches the residual.  sim_out: [n_sets][D][H][W]. */
int d3d_simulate(d3d_ctx* ctx, const double* params, int n_sets, double* sim_out);
/* Un-convolved lines, Run.simulate_clean (lib/run.py:597-621). */
int d3d_simulate_clean(d3d_ctx* ctx, const double* params, int n_sets, double* sim_out);
/* Current residual err_old as [n_chains][D][H][W] float64 (lib/run.py:334). */
int d3d_get_residual(d3d_ctx* ctx, double* err_out);

/* Spectral convolution alone: replaces convolve_1d(line, lsf)
 * (lib/convolution.py:89-120), batch lines of length n, same wrap rule. */
int d3d_conv1d(d3d_ctx* ctx, const double* lines, int n, int batch,
               const double* lsf, double* out);

/* Batched truncated normal: replaces rtnorm(a, b, mu, sigma) (lib/rtnorm.py:21-92).
 * Variate i is drawn from the stream (seed, chain, sweep, site = i), draw index 0
 * onwards; used_out[i] (may be NULL) receives the number of uniform draws consumed. */
int d3d_rtnorm(d3d_ctx* ctx, int n, const double* a, const double* b, const double* mu,
               const double* sigma, uint64_t seed, uint32_t chain, uint32_t sweep,
               double* out, int32_t* used_out);

/* ---- one proposal, no state change ----------------------------------------
 * Replaces lib/run.py:391-426 for a given proposal: out[0] = delta =
 * ar_old - ar_new (the value stored in `likelihoods`, lib/run.py:426-432),
 * out[1] = ar_old, out[2] = ar_new (windowed 0.5*nansum(err^2/var)). */
int d3d_delta_logl(d3d_ctx* ctx, int chain, int y, int x,
                   const double p_new[3], double out[3]);

/* ---- the MH-within-Gibbs sweep --------------------------------------------
 * Replaces the hot loop lib/run.py:344-537 for iterations
 * [first_iteration, first_iteration + n_iterations) (the reference's
 * cur_iteration; the first sweep is iteration 1).  Per site: Cauchy proposal
 * (:570-579), bounds (:379-388), delta-logL on the FSF window (:400-426),
 * accept test (:435-451), Gibbs amplitude draw from the truncated normal
 * (:456-519) and residual update.  The residual refresh of :525-534 happens
 * inside the call whenever cur_iteration % refresh_every == 0 (0 = never).
 * Each chain stops by itself when its acceptance rate falls to
 * min_acceptance_rate (:344-350, 356-359).
 *   chain_out [n_chains][n_rows][H][W][3], lik_out [n_chains][n_rows][H][W]
 *             (float64, host or device, NULL = not recorded): row
 *             it / keep_one_in is written when it % keep_one_in == 0 (:353,:430-451)
 *   accepted_out  [n_chains] int64: running accepted_count (:341,:440)
 *   iterations_out[n_chains] int64: value of cur_iteration when the chain stopped
 *   elapsed_ms    device time of the sweep kernels (CUDA events), or NULL     */
int d3d_sweep(d3d_ctx* ctx, int64_t first_iteration, int64_t n_iterations,
              int mode, int keep_one_in, int refresh_every,
              double min_acceptance_rate,
              double* chain_out, double* lik_out, int64_t n_rows,
              int64_t* accepted_out, int64_t* iterations_out, float* elapsed_ms);

/* ---- one oversized cube tiled over several contexts / GPUs ------------------
 *   The reference has a single process walk every spaxel
 * of the cube (lib/run.py:362-367); here each context owns the sites of one
 * rectangular tile and the coloured sweep is driven phase by phase so that the
 * contexts can exchange the OUTCOMES of their site updates between two phases
 * (the colour lattice (y mod fh, x mod fw) is global, hence the sites of a
 * phase have disjoint windows across tiles too).  Every context holds the
 * whole cube; its residual stays valid inside its tile grown by the FSF
 * half-size, the only part its own updates read.  The exchange (NCCL
 * all-gather, peer copies ...) belongs to the caller: deconv3d_b200/dist.py.
 *
 * A record is D3D_RECORD_DOUBLES float64: site (y*W+x, or -1 = empty slot),
 * chain, a, c, w (the parameters after the update, lib/run.py:448,499,516),
 * delta-logL (:430-432), accepted (:438-440), pad.  One phase fills
 * n_chains * ceil(H/fh) * ceil(W/fw) slots (d3d_tile_record_slots).          */
#define D3D_RECORD_DOUBLES 8
/* Sites owned by this context: y0 <= y < y1, x0 <= x < x1 (default: the whole
 * field).  Also restricts mode D3D_COLOURED of d3d_sweep to those sites. */
int d3d_set_tile(d3d_ctx* ctx, int y0, int y1, int x0, int x1);
int d3d_tile_record_slots(d3d_ctx* ctx, int64_t* n_records);
/* Loop condition and acceptance-rate bookkeeping of lib/run.py:344-359, once
 * per iteration and before its first phase. */
int d3d_colour_begin(d3d_ctx* ctx, int64_t iteration, double min_acceptance_rate);
/* Updates the owned sites of colour class (cy, cx) for `iteration` and writes
 * their records (host or device buffer of d3d_tile_record_slots records, or
 * NULL).  Asynchronous on the context's stream when the buffer is on the device. */
int d3d_colour_phase(d3d_ctx* ctx, int64_t iteration, int cy, int cx, double* records_out);
/* Applies the records of OTHER contexts (own and empty ones are skipped):
 * parameters, delta-logL, accepted_count and the residual inside the region. */
int d3d_apply_records(d3d_ctx* ctx, const double* records, int64_t n_records);
/* Fused exchange: the records go straight into the peers' memory (NVLink P2P
 * stores from the kernel, or plain stores when the tiles share a device) and
 * the applier of a phase waits on per-tile flags in its own memory -- no
 * collective library and no host round trip inside a phase.
 *   d3d_tile_fused_init    allocates this context's box for n_tiles tiles
 *                          (box_out: its device address, for peers in the
 *                          same process)
 *   d3d_tile_fused_export  64-byte CUDA IPC handle of the box (peers in other
 *                          processes)
 *   d3d_tile_fused_connect box of tile `index`: device address or IPC handle
 *   d3d_colour_phase_fused phase kernel + push of the records + applier;
 *                          phase_index must grow by one per call and be the
 *                          same on every tile (iteration * n_classes + class)
 * A peer that never publishes its phase makes the applier give up after ~2 s
 * (reported as D3D_ECUDA by d3d_get_chain_control). */
int d3d_tile_fused_init(d3d_ctx* ctx, int n_tiles, int my_index, void** box_out, int64_t* box_bytes);
int d3d_tile_fused_export(d3d_ctx* ctx, unsigned char* handle64);
int d3d_tile_fused_connect(d3d_ctx* ctx, int index, void* peer_box, const unsigned char* handle64);
int d3d_colour_phase_fused(d3d_ctx* ctx, int64_t iteration, int cy, int cx, int64_t phase_index);
/* Whole iterations of the fused tiled sweep, enqueued without returning to the
 * caller between phases (d3d_colour_begin + every class with
 * phase_index = iteration * n_classes + class). */
int d3d_sweep_fused(d3d_ctx* ctx, int64_t first_iteration, int64_t n_iterations,
                    double min_acceptance_rate);
/* Latest delta-logL of every site [n_chains][H][W] (the row lib/run.py:430-432
 * would store); tile mode only. */
int d3d_get_likelihoods(d3d_ctx* ctx, double* lik_out);
/* accepted_count (:341,:440), cur_iteration and the running flag of every chain
 * (any may be NULL); reports D3D_ENUMERIC like d3d_sweep. */
int d3d_get_chain_control(d3d_ctx* ctx, int64_t* accepted_out, int64_t* iterations_out,
                          int32_t* active_out);

/* Posterior summary without shipping the chain to the host: mean over rows
 * [first_row, n_rows) of chain [n_chains][n_rows][H][W][3] (host or device) ->
 * mean_out [n_chains][H][W][3] (host or device).  extract_parameters,
 * lib/run.py:581-593, with first_row = int((100 - percentage) * n_rows / 100). */
int d3d_chain_mean(d3d_ctx* ctx, const double* chain, int64_t n_rows, int64_t first_row,
                   double* mean_out);

/* Introspection for benches: launches of library kernels so far, algorithmic
 * bytes of the last d3d_sweep (3*s*D*wh*ww per site update with a variance
 * cube, 2*s*D*wh*ww with a scalar variance), site updates done. */
int d3d_get_counters(d3d_ctx* ctx, int64_t* kernel_launches,
                     int64_t* last_sweep_bytes, int64_t* last_sweep_site_updates);

/* Line model (replaces the plug-in hook lib/line_models.py:4-61 -- LineModel.modelize called at
 * lib/run.py:673, 604, 1016 -- for the family the device evaluates): a TIED MULTIPLET of
 * n_components (1..4) Gaussians that share the centre shift and the width,
 *     line(z) = a * sum_k ratios[k]/ratios[0] * exp(-(z - c - (offsets[k]-offsets[0]))^2 / (2 w^2)),
 * parameters (a, c, w) as in SingleGaussianLineModel (lib/line_models.py:64-109; n_components = 1,
 * the state after d3d_set_problem).  a is the amplitude of component 0 and stays the Gibbs
 * parameter (lib/run.py:456-519: the model is linear in it).  Examples: [NII]-Halpha-[NII], the
 * [OII] doublet.  Call after d3d_set_problem, before d3d_forward / d3d_sweep. */
int d3d_set_line_model(d3d_ctx* ctx, int n_components, const double* offsets, const double* ratios);

/* Bench support: measured FP64 FMA peak of the context's device in TFLOP/s (8 independent
 * DFMA chains per thread, every SM full, best of 5) -- the denominator of the FP64-pipe
 * fractions bench.py reports for the forward-model stencil and the sweep kernel.  Nothing in
 * the reference corresponds to it (the reference has no device code). */
int d3d_fp64_peak(d3d_ctx* ctx, double* tflops_out);

/* Bench support: name of the sweep kernel the latest d3d_sweep launched (static string). */
const char* d3d_last_kernel(d3d_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* DECONV3D_B200_H */
