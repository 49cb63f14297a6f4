/*
 * cmf_b200.h - C ABI of the B200-native convolutive-NMF multiplicative-update engine.
 *
 * This is the drop-in boundary for the hot path of degleris1/cmfpy: the solver
 * object that `CMF.fit` builds at reference cmfpy/model.py:146
 * (`ALGORITHMS[alg_name](data, dims, **alg_opts)`) and drives through
 * `.update()` / `.loss` / `.W` / `.H` (model.py:149-176).  The reference has no
 * FFI of its own (it is pure Python + NumPy); each entry point below names the
 * reference interface it replaces.  A reference maintainer binds these with
 * ctypes - see INTEGRATION.md and cmfpy_b200/_lib.py.
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success, non-zero on
 *     failure, and never throws.  cmf_last_error() describes the last failure
 *     on the calling thread.
 *   - host-side array arguments use the REFERENCE layouts (row-major):
 *       X  : N x T        (features x time)        model.py:127-128
 *       W  : L x N x K    (lags x features x comps) model.py:223-229
 *       H  : K x T                                   model.py:239-245
 *   - `dtype` is CMF_F32 or CMF_F64 (the reference works in float64; the
 *     device computes in fp32 / tf32 and converts at the boundary).
 *   - `mem` is CMF_HOST (pageable or pinned host pointer) or CMF_DEVICE
 *     (pointer valid on the solver's device, e.g. a torch tensor's data_ptr()).
 *   - a handle is not thread-safe; one host thread drives one handle.
 *
 * Time sharding: a handle owns the global columns
 * [t_offset, t_offset + t_local) of a problem with t_global columns.  A
 * single-GPU solve is the case t_offset = 0, t_local = t_global.  The phase
 * calls (cmf_mu_w_terms ... cmf_mu_recon) plus the halo import/export and the
 * exposed W-term buffer are what a multi-GPU driver strings together with an
 * all-reduce and a neighbour exchange of its own (NCCL, MPI); cmf_mu_step() is
 * the fused single-GPU iteration and cmf_mu_step_sharded() the multi-GPU one,
 * whose collectives are the library's own kernels over NVLink peer memory.
 */
#ifndef CMF_B200_H
#define CMF_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define CMF_B200_ABI_VERSION 1

enum { CMF_F32 = 0, CMF_F64 = 1 };
enum { CMF_HOST = 0, CMF_DEVICE = 1 };
/* CMF_PREC_FP32: exact fp32 FFMA contractions.
 * CMF_PREC_TF32: tcgen05 tensor-core contractions, operands rounded (RN) to
 *                TF32, fp32 accumulation in tensor memory.
 * CMF_PREC_TF32X3: the same tensor-core kernels with every operand held as a
 *                TF32 pair (hi = RN(x), lo = RN(x - hi)); each product is
 *                a_lo b_hi + a_hi b_lo + a_hi b_hi (error-compensated "3xTF32":
 *                ~22 mantissa bits per operand, fp32-grade results at a third
 *                of the tf32 rate), on either denominator route.             */
enum { CMF_PREC_FP32 = 0, CMF_PREC_TF32 = 1, CMF_PREC_TF32X3 = 2 };
/* How the MU denominators (the est-dependent halves of mult.py:37-38, 46) are formed on the tf32 path.
 * CMF_DEN_DIRECT: contract est, as the reference does.
 * CMF_DEN_GRAM  : exact identity through the small Gram operators
 *                   den_W[l] = sum_l' W[l'] A[l-l'],   A[d] = sum_t H[:,t+d] H[:,t]^T
 *                   den_H[:,t] = sum_d R[d] H[:,t+d],  R[d] = sum_{l-l'=d} W[l]^T W[l']
 *                 (minus the terms of est past the end of the data); K/N of the direct cost, and est is
 *                 then needed once per iteration (for the loss) instead of twice.  Ignored on the fp32 path. */
enum { CMF_DEN_DIRECT = 0, CMF_DEN_GRAM = 1, CMF_DEN_AUTO = 2 /* Gram when the shard is large enough to pay for
                                                              its extra small kernels (2 N K L t_local >= 2e11 and N >= 4 K) */ };

typedef struct cmf_mu_s cmf_mu_t;

typedef struct cmf_mu_params {
  int n_features;        /* N   (ModelDimensions.n_features,   model.py:62) */
  int n_components;      /* K   (ModelDimensions.n_components, model.py:65) */
  int maxlag;            /* L   (ModelDimensions.maxlag,       model.py:64) */
  long long t_local;     /* columns owned by this handle                    */
  long long t_global;    /* T   (ModelDimensions.n_timepoints, model.py:63) */
  long long t_offset;    /* global index of the first owned column          */
  int device;            /* CUDA device ordinal                             */
  int precision;         /* CMF_PREC_*                                      */
  void* stream;          /* cudaStream_t to run on, or NULL for an own one  */
  int denominators;      /* CMF_DEN_*                                       */
} cmf_mu_params;

/* ---- library ---------------------------------------------------------- */
int         cmf_abi_version(void);
const char* cmf_last_error(void);
int         cmf_device_count(int* count);
/* The N x T device buffers of a destroyed solver stay in a small process-wide cache (same device, same size; at
 * most CMF_CACHE_GB GiB, default 16) so that the next solver of the same shape does not pay cudaMalloc again - tens
 * to hundreds of ms for a few GiB.  This returns the cached blocks to the driver.                                    */
int         cmf_release_cached_memory(void);
/* 1 if `precision` has a kernel path for this shape on this build.         */
int         cmf_precision_supported(int precision, int n_features,
                                    int n_components, int maxlag);

/* ---- solver lifecycle: AbstractOptimizer.__init__, algs/base.py:15-43 -- */
int cmf_mu_create(cmf_mu_t** out, const cmf_mu_params* params);
int cmf_mu_destroy(cmf_mu_t* h);

/* self.X = data (base.py:24).  X: N x ncols row-major with leading dimension
 * ld (elements); ncols in [t_local, t_local + L - 1]: the columns past
 * t_local are the static right halo a shard needs for the H step (zero when
 * absent or past t_global).  Also accumulates the local sum of squares used
 * for normX (base.py:25) and flags negative entries (model.py:138).         */
int cmf_mu_set_data(cmf_mu_t* h, const void* X, int dtype, int mem,
                    long long ld, long long ncols);
/* Local sum over owned columns of X^2 and whether any entry was negative.   */
int cmf_mu_data_stats(cmf_mu_t* h, double* sumsq, int* has_negative);
/* Dataset normalisation on the device (the step BEFORE the solver in the
 * reference: datasets/songbird.py:18-19 rows / (1e-6 + L2 norm), datasets/maze.py:
 * 71-72 rows / (1e-8 + L1 norm), datasets/vox_celeb.py:100-102 unit variance per
 * feature).  cmf_mu_row_stats returns the per-feature sums over the OWNED columns
 * (sum x, sum x^2, sum |x|; N doubles each, HOST, any may be NULL - a sharded
 * driver all-reduces them); cmf_mu_scale_rows multiplies feature n by scale[n]
 * and refreshes the local ||X||^2 / negativity flag like cmf_mu_set_data.      */
int cmf_mu_row_stats(cmf_mu_t* h, double* s1, double* s2, double* sabs);
int cmf_mu_scale_rows(cmf_mu_t* h, const double* scale);
/* normX = ||X||_F of the GLOBAL matrix (base.py:25).  Defaults to the local
 * value; a sharded driver sets the all-reduced one.                          */
int cmf_mu_set_norm_x(cmf_mu_t* h, double norm_x);

/* self.W, self.H = initW, initH (base.py:40-41).  W0: L x N x K,
 * H0: K x t_local (owned columns only, leading dimension ldh).  Copies; the
 * caller's arrays are never written (MU rebinds W/H, mult.py:18,22).         */
int cmf_mu_set_factors(cmf_mu_t* h, const void* W0, const void* H0, int dtype,
                       int mem, long long ldh);

/* rand_init (base.py:78-88): with random W, H already set, the two local
 * reductions behind alpha = <X, est> / ||est||^2 over owned columns, and the
 * in-place rescale W *= scale_w, H *= scale_h (base.py:88: both sqrt(alpha)). */
int cmf_mu_init_stats(cmf_mu_t* h, double* x_dot_est, double* est_sumsq);
int cmf_mu_scale_factors(cmf_mu_t* h, double scale_w, double scale_h);

/* ---- halo plumbing for time sharding (no reference counterpart) --------- */
/* Edges to send: left_edge = first L-1 owned columns, right_edge = last L-1
 * owned columns of H, each as (L-1) x halo_ld fp32, time-major, DEVICE.     */
int cmf_mu_halo_width(cmf_mu_t* h, int* n_cols, int* halo_ld);
int cmf_mu_halo_export(cmf_mu_t* h, float* left_edge, float* right_edge);
/* Halos received: left_halo = the L-1 columns before t_offset (from rank-1's
 * right_edge), right_halo = the L-1 columns after the owned range (from
 * rank+1's left_edge).  NULL means zeros (global boundary).                  */
int cmf_mu_halo_import(cmf_mu_t* h, const float* left_halo, const float* right_halo);

/* ---- phases of MultUpdate.update(), algs/mult.py:15-25 ------------------ */
/* est = cmf_predict(W, H) (common.py:50-58) on owned + right-halo columns,
 * and the local sum of squared residuals (cache_resids, base.py:57-62).      */
int cmf_mu_recon(cmf_mu_t* h);
/* Same, but est itself may be left unwritten when no MU step reads it (both
 * denominators from the Gram route): only the residual sum of squares is
 * refreshed.  cmf_mu_get_est and the direct routes recompute est on demand.   */
int cmf_mu_recon_loss(cmf_mu_t* h);
/* num/denom of the W step (_compute_mult_W, mult.py:27-40) from the cached
 * est; local partial sums over owned columns.  The result lives in one
 * DEVICE buffer [num | den], 2 * count fp32, exposed for an in-place
 * all-reduce.                                                                */
int cmf_mu_w_terms(cmf_mu_t* h);
int cmf_mu_w_terms_buffer(cmf_mu_t* h, float** dev_ptr, long long* count);
/* W <- W * num / (den + EPSILON)  (mult.py:18).                             */
int cmf_mu_w_apply(cmf_mu_t* h);
/* num/denom of the H step (_compute_mult_H, mult.py:42-48; tensor_transconv,
 * common.py:61-86) from the cached est, then H <- H * num / (den + EPSILON)
 * (mult.py:22) on owned columns.                                             */
int cmf_mu_h_step(cmf_mu_t* h);
/* Local sum of squared residuals of the last cmf_mu_recon (device -> host). */
int cmf_mu_resid_sumsq(cmf_mu_t* h, double* sumsq);
/* DEVICE address of that double, so a sharded driver can all-reduce it
 * without a host round trip.                                                */
int cmf_mu_resid_sumsq_buffer(cmf_mu_t* h, double** dev_ptr);
/* loss = ||resids||_F / normX (base.py:90-97) from the local residual; only
 * meaningful unsharded or after the driver reduced it itself.               */
int cmf_mu_loss(cmf_mu_t* h, double* loss);

/* 1 if the H step needs a reconstruction with the updated W before it (the
 * direct denominators do; the Gram route does not).  A sharded driver skips
 * the mid-iteration cmf_mu_recon when this returns 0.                         */
int cmf_mu_needs_mid_recon(cmf_mu_t* h, int* needed);

/* ---- the fused single-GPU iteration ------------------------------------ */
/* n_steps x MultUpdate.update() (mult.py:15-25): W terms, W update, recon,
 * H terms, H update, recon + loss.  loss_out[i] = loss after step i
 * (what update() returns, mult.py:25); ms_out[i] = device time of step i
 * (CUDA events; may be NULL).  One host sync at the end.                     */
int cmf_mu_step(cmf_mu_t* h, int n_steps, double* loss_out, float* ms_out);

/* ---- the sharded iteration with its collectives over peer memory --------- */
/* One process per GPU.  Each rank publishes CMF_PEER_BLOB_BYTES (CUDA IPC
 * handles of its W-term buffer, W and a small control block); the host layer
 * all-gathers the blobs (any transport) and hands the world's blobs to
 * cmf_mu_peer_attach, which maps the peers' buffers (NVLink / NVSwitch P2P).
 * cmf_mu_step_sharded then runs n_steps x MultUpdate.update() (mult.py:15-25)
 * with NO host-side collective: the all-reduce of the W terms is fused with the
 * W update in one kernel (reduce-scatter -> update -> all-gather over peer
 * loads/stores), the L-1 halo columns of H are pushed into the neighbours, and
 * the residual sums travel through a ring in every rank's memory.  Every rank
 * must call it with the same n_steps; loss_out[i] is the GLOBAL loss after
 * step i, identical on all ranks.  A host barrier is required between attach
 * and the first step, and before detach / destroy.                            */
#define CMF_PEER_BLOB_BYTES 512
int cmf_mu_peer_export(cmf_mu_t* h, void* blob);
int cmf_mu_peer_attach(cmf_mu_t* h, int rank, int world, const void* blobs);
int cmf_mu_peer_detach(cmf_mu_t* h);
int cmf_mu_step_sharded(cmf_mu_t* h, int n_steps, double* loss_out);
/* The same wiring when all shards live in ONE process (CMF(..., devices=[0, 1, ...]), reference plug-in point
 * cmfpy/model.py:80-82, :146): `handles` are the `world` solvers in rank order, each on its own GPU; peers are
 * reached through cudaDeviceEnablePeerAccess instead of CUDA IPC.  One host thread per device then calls
 * cmf_mu_step_sharded.  cmf_mu_halo_exchange_peer runs one exchange of the L-1 boundary columns of H through
 * peer memory (every rank calls it, concurrently): the halos of the initial factors. */
int cmf_mu_peer_attach_local(cmf_mu_t* h, int rank, int world, cmf_mu_t* const* handles);
int cmf_mu_halo_exchange_peer(cmf_mu_t* h);

/* ---- gradient solvers: GradDescent / BlockDescent, algs/gradient_descent.py -- */
/* Their gradients are the MU terms: gW[l] = s_T_dot(resids, H, l) = den_W - num_W
 * (:41-45) and gH = den_H - num_H (:47-52), so the same contraction kernels serve.
 * Single GPU, direct denominators (CMF_DEN_DIRECT).
 * cmf_gd_cache: cache_resids + cache_gW + cache_gH (:36-38); afterwards
 *   cmf_mu_get_w_terms / cmf_mu_h_terms return num and den of the current factors.
 * cmf_gd_lipschitz_w: lambda_max of the (K L) x (K L) block-Toeplitz matrix of the
 *   lag autocorrelations of H (:54-69), by power iteration on the device.
 * cmf_gd_lipschitz_state: whether the last power iteration (of cmf_gd_lipschitz_w or
 *   cmf_gd_step) settled (two consecutive relative changes <= 1e-7) and how many
 *   iterations it issued (batches of 64, up to 1024); the reference's eigh() is exact.
 * cmf_gd_step: one update(); block_descent = 0: W and H steps from the cached
 *   gradients, then residuals and both gradients (:81-92); 1: W step, residuals,
 *   gH, H step, residuals, gW (:132-147).  x <- max(x - ss g, 0) (:148-159) with
 *   ss = 1 / lipschitz_W for W and step_size_h for H.  Returns the loss (:120-123). */
int cmf_gd_cache(cmf_mu_t* h);
int cmf_gd_lipschitz_w(cmf_mu_t* h, double* lambda_max);
int cmf_gd_lipschitz_state(cmf_mu_t* h, int* settled, int* iterations);
int cmf_gd_step(cmf_mu_t* h, int block_descent, double step_size_h, double* loss_out);

/* ---- HALS: HALSUpdate, algs/hals.py on algs/accelerated.py ----------------- */
/* Hierarchical alternating least squares with the residual est - X kept
 * current on the device after every coordinate block.  Single GPU, T > L.
 *   cmf_hals_begin   : the residual of the current factors (cache_resids).
 *   cmf_hals_sweep_w : update_W (hals.py:47-48, 78-105): all (k, l) columns in
 *                      the reference's order; *diff_norm = ||W_new - W_old||_F,
 *                      what the inner-iteration stop rule compares
 *                      (accelerated.py:50-69); NULL skips it.
 *   cmf_hals_sweep_h : update_H (hals.py:68-70, 113-181): per (k, l) the batch of
 *                      entries t = l (mod L), t < T - L, then the entry T - L + l.
 *   cmf_hals_end     : cache_resids from scratch and the loss (accelerated.py:83-84). */
int cmf_hals_begin(cmf_mu_t* h);
int cmf_hals_sweep_w(cmf_mu_t* h, double* diff_norm);
int cmf_hals_sweep_h(cmf_mu_t* h, double* diff_norm);
int cmf_hals_end(cmf_mu_t* h, double* loss_out);

/* ---- read-back: algorithm.W / algorithm.H (model.py:175-176) ------------ */
int cmf_mu_get_W(cmf_mu_t* h, void* W_out, int dtype, int mem);
int cmf_mu_get_H(cmf_mu_t* h, void* H_out, int dtype, int mem, long long ldh);
/* self.est (base.py:61) / CMF.predict (model.py:191-200): N x t_local.      */
int cmf_mu_get_est(cmf_mu_t* h, void* est_out, int dtype, int mem, long long ld);
/* Debug / kernel-level parity: the H-step terms without applying them,
 * each K x t_local row-major, HOST.                                          */
int cmf_mu_h_terms(cmf_mu_t* h, void* num_out, void* den_out, int dtype);
/* Debug / kernel-level parity: W-step terms, each L x N x K row-major, HOST.*/
int cmf_mu_get_w_terms(cmf_mu_t* h, void* num_out, void* den_out, int dtype);

/* Number of kernels this handle has launched since creation (bench
 * `gpu_launches`) and the name of the contraction path in use.              */
int         cmf_mu_launch_count(cmf_mu_t* h, long long* count);
const char* cmf_mu_path_name(cmf_mu_t* h);
/* Device time (ms) spent in each contraction kernel during the last
 * cmf_mu_step call: [recon, w_terms, h_terms, elementwise], summed over
 * steps (CUDA events on the solver's stream).                               */
int cmf_mu_kernel_ms(cmf_mu_t* h, float out[4]);
/* Toggle per-kernel event timing inside cmf_mu_step (off by default).       */
int cmf_mu_set_profiling(cmf_mu_t* h, int on);
/* How the loss of an iteration (reference base.py:90-97) is formed when nothing else reads the reconstruction (both
 * denominators on the Gram route).
 * 0 (default, "auto"), 3xTF32 mode:
 *   - while loss^2 >= 0.1, from the W terms of the updated factors and without any reconstruction:
 *       ||X - est||^2 = ||X||^2 - 2 <W, num_W> + <W, den_W>
 *     (num_W, den_W of mult.py:35-38 are the derivatives of <X, est> and ||est||^2 / 2 with respect to W, and the next
 *     iteration's W step needs them anyway).  The identity is exact; its cancellation amplifies the ~1e-6 relative
 *     error of the contractions by 1 / (2 loss^2) <= 5, hence the bound; re-decided every 8 iterations;
 *   - otherwise, on large problems (L N K and K T of a million entries or more, K L loss^2 >= 0.2), the
 *     reconstruction runs its hi x hi operand pass alone - the two cross passes it omits, the rounding residuals of
 *     W and H, change the loss by less than 1e-6 relative; else all three passes.
 *   Plain TF32 always reconstructs in auto mode (its contraction errors are too large to amplify).
 * 1 ("full"): the residual is formed explicitly with every operand pass, every iteration.
 * 2 ("wterms"): the identity whenever both denominators are on the Gram route, any precision, any loss.
 * W and H are never affected by this choice. */
int cmf_mu_set_loss_mode(cmf_mu_t* h, int mode);
/* on = 2: additionally one CUDA event after EVERY kernel launch of the following steps; cmf_mu_launch_table then
 * returns one text line "label launches total_ms" per kernel label (a measurement aid of bench.py: the per-kernel
 * roofline table; the reference only has wall-clock time_hist, cmfpy/model.py:160-167). */
int cmf_mu_launch_table(cmf_mu_t* h, char* buf, long long cap);

/* ---- stateless primitives (cmfpy/common.py) ----------------------------- */
/* cmf_predict(W, H) -> est, N x T (common.py:50-58).  Host pointers.        */
int cmf_predict(const void* W, const void* H, void* est_out, int dtype,
                int n_features, long long n_timepoints, int n_components,
                int maxlag, int device, int precision);
/* CMF.score(data) (model.py:202-221): R^2 = 1 - ||cmf_predict(W,H) - X||^2 /
 * ||X||^2, reduced on the device (no N x T read-back).  Host pointers.       */
int cmf_score(const void* W, const void* H, const void* X, int dtype,
              int n_features, long long n_timepoints, int n_components,
              int maxlag, int device, int precision, double* r2_out);
/* tensor_transconv(W, X) -> K x T (common.py:61-86).  Host pointers.        */
int cmf_tensor_transconv(const void* W, const void* X, void* out, int dtype,
                         int n_features, long long n_timepoints,
                         int n_components, int maxlag, int device, int precision);

/* ---- the step before the solver, on the device ---------- */
/* A row-major fp32 matrix in device memory, owned by the library: what the
 * generators below produce and what cmf_mu_set_data accepts with CMF_DEVICE
 * (ptr, ld from cmf_dmat_info), so a data set never visits the host. */
typedef struct cmf_dmat_s cmf_dmat_t;
int cmf_dmat_info(cmf_dmat_t* m, const float** dev_ptr, long long* rows, long long* cols, long long* ld, int* device);
/* copy to a HOST array (rows x cols, leading dimension ld, CMF_F32 or CMF_F64) */
int cmf_dmat_get(cmf_dmat_t* m, void* out, int dtype, long long ld);
int cmf_dmat_destroy(cmf_dmat_t* m);

/* The reference's synthetic data set, Synthetic(...) of cmfpy/datasets/synthetic.py:7-39: H ~ U[0,1) thinned to a
 * fraction 1 - H_sparsity of non-zeros (:22-25), one Gaussian-bump motif per feature on a random component (:27-30,
 * :42-46), noise = noise_scale * U[0,1) (:33), data = cmf_predict(W, H) + noise (:36, the reconstruction on the K1
 * kernel of `precision`).  The reference draws from NumPy's generators (partly the global, unseeded one); here a
 * counter-based generator keyed by (seed, global element index) makes the data set reproducible and identical for any
 * time sharding: a handle holds the columns [t_offset, t_offset + t_local) of the n_timebins-column data set. */
typedef struct cmf_synth_s cmf_synth_t;
typedef struct {
  int n_components, n_features, n_lags;
  long long n_timebins, t_offset, t_local;
  double H_sparsity, noise_scale;
  unsigned long long seed;
  int device, precision;
} cmf_synth_params;
enum { CMF_SYNTH_W = 0, CMF_SYNTH_H = 1, CMF_SYNTH_NOISE = 2, CMF_SYNTH_DATA = 3, CMF_SYNTH_GENERATE = 4 };
int cmf_synth_create(cmf_synth_t** out, const cmf_synth_params* p);
/* `what` to a HOST array in the reference's layout: W (L x N x K), H (K x t_local, ld), NOISE / DATA / GENERATE
 * (N x t_local, ld); GENERATE is the reference's generate() = data + noise (synthetic.py:38-39: the noise a 2nd time). */
int cmf_synth_get(cmf_synth_t* s, int what, void* out, int dtype, long long ld);
/* DATA (shares the handle's buffer) or GENERATE (a new matrix) as a device matrix */
int cmf_synth_matrix(cmf_synth_t* s, int what, cmf_dmat_t** out);
int cmf_synth_destroy(cmf_synth_t* s);

/* The spectrogram step of the audio loader, VoxCeleb.generate of cmfpy/datasets/vox_celeb.py:58-104:
 * scipy.signal.spectrogram(audio, fs, window, nperseg, noverlap) with its defaults (one-sided power spectral density,
 * constant detrend) followed, if `normalize`, by StandardScaler(with_mean=False) over time for every frequency bin
 * (:100-102).  audio: n_samples values (CMF_F32 / CMF_F64, CMF_HOST / CMF_DEVICE); window: nperseg doubles, HOST
 * (the reference's default is scipy's Tukey(0.25) window).  Result: (nperseg/2 + 1) x n_segments, features x time. */
int cmf_spectrogram(const void* audio, int dtype, int mem, long long n_samples, double fs, int nperseg, int noverlap,
                    const double* window, int normalize, int device, cmf_dmat_t** out);

#ifdef __cplusplus
}
#endif
#endif /* CMF_B200_H */
