/*
 * free_hunch_b200 -- C ABI of the B200-native Free Hunch guided-sampling hot path.
 *
 * The reference (AaltoML/free-hunch) is pure Python and has no FFI; this header is the boundary a
 * maintainer would bind (ctypes stub in INTEGRATION.md).  Every entry point names the reference
 * interface it replaces (file:line relative to the reference root).  All pointers named *_dev are
 * caller-owned device buffers of float32 unless stated otherwise; `stream` is a cudaStream_t passed as
 * void*.  Functions return 0 on success and a negative FH_E* code on failure; fh_last_error() gives text.
 * Nothing in this library falls back to the CPU: without a CUDA device every call fails with FH_ECUDA.
 *
 * Threading / streams: one host thread and ONE stream per fh_ctx at a time.  The context owns scratch planes and the
 * buffers of destroyed states (parked for reuse by the next fh_state_create); both are protected by stream order only,
 * so work of one context must not be spread over several streams without the caller ordering them with events.
 *
 * Layout: an "image vector" is (B, C, H, W) contiguous float32, d = C*H*W per image.  In the DCT basis
 * (image_base_covariance = dct_diagonal) H = W = 256 is required.
 */
#ifndef FREE_HUNCH_B200_H
#define FREE_HUNCH_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define FH_OK 0
#define FH_EINVAL (-1)
#define FH_ECUDA (-2)
#define FH_ERANK (-3) /* low-rank capacity exhausted */
#define FH_ESTATE (-4)

#define FH_BASIS_IDENTITY 0 /* CovarianceHessianBFGS     online_update_bfgs.py:7   */
#define FH_BASIS_DCT 1      /* CovarianceHessianBFGSDCT  online_update_bfgs.py:339 */

#define FH_OP_INPAINT 0 /* measurements.py:204 + conditioning_mechanisms.py:384 */
#define FH_OP_BLUR 1    /* measurements.py:126,164 + conditioning_mechanisms.py:489 (gaussian and motion) */
#define FH_OP_SR 2      /* measurements.py:87 + conditioning_mechanisms.py:641 */

typedef struct fh_ctx fh_ctx;     /* per-device workspace: twiddles, scratch planes, reduction buffers */
typedef struct fh_state fh_state; /* covariance state of B images: prior precision diag(pi)+Z^T Gz Z   */
typedef struct fh_op fh_op;       /* measurement operator A / A^T                                       */

const char* fh_last_error(void);
int fh_version(void);

/* ---- context ------------------------------------------------------------------------------- */
int fh_ctx_create(fh_ctx** out, int device, int max_batch, int channels, int height, int width, int max_rank);
int fh_ctx_destroy(fh_ctx* ctx);
/* Which implementation fh_cg_solve uses in the DCT basis (same algorithm, cg.py:230-292, same results to rounding):
 *   FH_CG_STREAMING  five launches per iteration, Z streamed from L2/HBM by TMA, any batch, any rank <= 64;
 *   FH_CG_RESIDENT   one cooperative launch per image and solve, up to 32 rows of Z resident in shared memory (the rest
 *                    re-read from L2), for operators that fold into the DCT passes: inpainting mask, separable blur;
 *   FH_CG_AUTO       resident where it applies and the batch is at most 4 images (the resident kernel takes the images
 *                    one after the other; from 5 images on the streaming iteration is faster per image) (default).
 * Results are bit-identical across batch sizes within one implementation, not between the two.
 * FH_CG_RESIDENT makes fh_cg_solve fail with FH_EINVAL when the resident kernel does not apply. */
#define FH_CG_AUTO 0
#define FH_CG_STREAMING 1
#define FH_CG_RESIDENT 2
#define FH_CG_RESIDENT_F32 3 /* the float32 predecessor of the resident kernel (kept for comparison; not a parity path) */
int fh_ctx_set_cg_mode(fh_ctx* ctx, int mode);
int fh_ctx_cg_last_mode(const fh_ctx* ctx); /* FH_CG_STREAMING or FH_CG_RESIDENT: path of the last DCT-basis solve */
/* 1 when the float64 entry points of the covariance updates (fh_time_update_f64, fh_time_update_from_mean,
 * fh_space_update_f64; online_update_bfgs.py:153-192, :247-312) run as ONE cooperative launch per update and image
 * (DCT basis, at most 4 images per state, a device with one SM per 6-row strip plus two); 0: the multi-launch path. */
int fh_ctx_fused_updates(const fh_ctx* ctx);

/* Orthonormal 2-D DCT-II (inverse=0) / DCT-III (inverse=1) of `nplanes` 256x256 planes.
 * Replaces torch_dct.dct_2d / idct_2d(norm='ortho'), call sites online_update_bfgs.py:352-374. */
int fh_dct2(fh_ctx* ctx, const float* in_dev, float* out_dev, int nplanes, int inverse, void* stream);

/* ---- covariance state ------------------------------------------------------------------------
 * fh_state_create + fh_state_init replace CovarianceHessianBFGS.__init__ (online_update_bfgs.py:12-56)
 * and CovarianceHessianBFGSDCT.__init__ (:341-349).  init_variance_dev holds d values (shared by all B
 * images): the initial diagonal of Sigma (dct_variance.pt, or ones). */
int fh_state_create(fh_ctx* ctx, fh_state** out, int batch, int basis);
int fh_state_destroy(fh_state* st);
int fh_state_init(fh_state* st, const float* init_variance_dev, double init_noise_variance, void* stream);
int fh_state_rank(const fh_state* st); /* m = rows of Z (2 per space update) */
double fh_state_var2(const fh_state* st); /* sigma^2 the Sigma side of the state currently sits at */

/* Sigma . v  -- denoiser_cov_vector_dot, online_update_bfgs.py:194-204 (identity) and :370-374 (DCT:
 * v and the result are in pixel space, DCT -> diag+low-rank -> IDCT happens inside). */
int fh_cov_dot(fh_state* st, const float* v_dev, float* out_dev, void* stream);
/* The same product with float64 in and out and float64 arithmetic throughout (DCT basis; FH_ESTATE otherwise): the
 * mechanism's mu + Sigma mat (conditioning_mechanisms.py:284-285) cancels to ~1 % of its terms at high sigma. */
int fh_cov_dot_f64(fh_state* st, const double* v_dev, double* out_dev, void* stream);

/* update_time_step, online_update_bfgs.py:153-192 (+ DCT wrapper :357-360).  x/score/mean_out/score_out
 * are (B,d).  only_covariance!=0 skips the score/mean update (outputs untouched, may be NULL). */
int fh_time_update(fh_state* st, const float* x_dev, const float* score_dev, double sigma_t, double sigma_tnext,
                   int only_covariance, float* mean_out_dev, float* score_out_dev, void* stream);

/* update_time_step with float64 x / score in and float64 mean / score out: mu' = x + sigma'^2 s' is formed in float64
 * (the reference's x is float64, generate_conditional.py:116, and |x| ~ sigma >> |mu'|). */
int fh_time_update_f64(fh_state* st, const double* x_dev, const double* score_dev, double sigma_t, double sigma_tnext,
                       int only_covariance, double* mean_out_dev, double* score_out_dev, void* stream);

/* update_space_step, online_update_bfgs.py:247-312 (+ DCT wrapper :362-368); project_to_diagonal variant
 * when project_to_diagonal != 0 (:274-277). */
int fh_space_update(fh_state* st, const float* mean_at_x_dev, const float* mean_at_xnext_dev, double sigma_t,
                    const float* x_dev, const float* xnext_dev, int project_to_diagonal, void* stream);

/* update_space_step from float64 inputs: the differences dx = x' - x and de = sigma^2 (mu' - mu_hat) are formed in
 * float64 before they are rounded to the state's float32. */
int fh_space_update_f64(fh_state* st, const double* mean_at_x_dev, const double* mean_at_xnext_dev, double sigma_t,
                        const double* x_dev, const double* xnext_dev, int project_to_diagonal, void* stream);

/* Test hooks of the reference tests (online_update_tests.py:153-155): load Sigma = diag + U U^T - V V^T
 * (U: ku x d rows, V: kv x d rows, identical for all B images) and rebuild everything else
 * (set_others_corresponding_to_current_denoiser_cov, online_update_bfgs.py:327-330). */
int fh_state_set_cov_vectors(fh_state* st, const float* U_dev, int ku, const float* V_dev, int kv, double sigma,
                             void* stream);
/* drop_vectors(0, sigma), online_update_bfgs.py:235-239: forget the low-rank part of Sigma at `sigma`. */
int fh_state_drop_all(fh_state* st, double sigma, void* stream);
/* drop_vectors(0 < n < k, sigma), online_update_bfgs.py:241-245: "keep the last n columns" of the reference's factors
 * U, V of Sigma.  Those columns are combinations of the state's rows -- the host tracks their m x k coefficient
 * matrices (free_hunch_b200/ref_factors.py restates the reference's two Woodbury re-factorisations, :87-119, on them),
 * for which it needs  K = Z diag(delta) Z^T  ((B, m, m) float64, device) at the state's current sigma ... */
int fh_state_gram(fh_state* st, double* K_out_dev, void* stream);
/* ... and when it drops, the state moves to the basis of the kept columns:  Z <- T^T Z  (T: (B, m, m_new) float64,
 * device, row-major), Sigma^-1 = diag(pi + 1/sigma^2) + Z^T Gz_new Z  (Gz_new: (B, m_new, m_new) float64, device),
 * rank <- m_new, Sigma's core rebuilt.  (set_others_corresponding_to_current_denoiser_cov, :327-330.) */
int fh_state_recombine(fh_state* st, int m_new, const double* T_dev, const double* Gz_new_dev, void* stream);
/* Export for get_dense_matrices (online_update_bfgs.py:320-325): pi (B,d) f32, Z (B,m,d) f32,
 * Gz (B,m,m) f64, Gs (B,m,m) f64, delta (B,d) f32; any pointer may be NULL. */
int fh_state_export(fh_state* st, float* pi_dev, float* Z_dev, double* Gz_dev, double* Gs_dev, float* delta_dev,
                    void* stream);

/* update_time_step fed with the previous evaluation: score_prev = (mean_prev - x_prev) / sigma_t^2 is formed inside
 * (conditioning_mechanisms.py:243-244); float64 in and out like fh_time_update_f64. */
int fh_time_update_from_mean(fh_state* st, const double* x_prev_dev, const double* mean_prev_dev, double sigma_t, double sigma_tnext,
                             double* mean_out_dev, double* score_out_dev, void* stream);

/* ---- element-wise glue of a mechanism call / sampler step, fused (float64, same operation order as the reference's
 * PyTorch expressions, hence bit-identical to them) --------------------------------------------------------------
 * fh_mech_combine: conditioning_mechanisms.py:283-288 per image b:
 *     flag_b = std(g_b * sigma2) > threshold ;  out_b = mu_b + (flag_b ? sigv_b * scale / sigma2 : g_b * scale) * sigma2
 *   g = the VJP through the denoiser, sigv = Sigma.mat; flag_dev[b] receives the decision (device memory: no host read).
 * fh_mech_allclose: torch.allclose(a_b, b_b) of :250 per image; notclose_dev[b] = 0 when allclose.
 * fh_sampler_euler / fh_sampler_heun: generate_conditional.py:146-149 and :160-167. */
int fh_mech_combine(fh_ctx* ctx, const double* mu_dev, const double* g_dev, const double* sigv_dev, double sigma2, double threshold,
                    double scale, double* out_dev, int* flag_dev, int batch, void* stream);
/* The same in two steps: fh_mech_flag takes the decision (device memory), fh_cov_dot_f64_if forms Sigma.mat only for the
 * images whose flag is set -- the reference computes it inside the `if` (:283-286) -- and fh_mech_apply combines.  An
 * image whose flag is 0 costs one early-exit launch instead of a Sigma.v product, still without a host read. */
int fh_mech_flag(fh_ctx* ctx, const double* g_dev, double sigma2, double threshold, int* flag_dev, int batch, void* stream);
int fh_cov_dot_f64_if(fh_state* st, const double* v_dev, double* out_dev, const int* flag_dev, void* stream);
int fh_mech_apply(fh_ctx* ctx, const double* mu_dev, const double* g_dev, const double* sigv_dev, double sigma2, double scale,
                  const int* flag_dev, double* out_dev, int batch, void* stream);
int fh_mech_allclose(fh_ctx* ctx, const double* a_dev, const double* b_dev, int* notclose_dev, int batch, void* stream);
int fh_sampler_euler(const double* x_hat_dev, const double* den_dev, double t_cur, double h, double* d_cur_dev, double* x_prime_dev,
                     long n, void* stream);
int fh_sampler_heun(const double* x_hat_dev, const double* d_cur_dev, const double* x_prime_dev, const double* den2_dev, double t_next,
                    double h, double* x_next_dev, long n, void* stream);

/* ---- measurement operators ------------------------------------------------------------------- */
/* Circular convolution with a ks x ks PSF (host pointer, row-major), the FFT operator of
 * measurements.py:140-157,180-198 / utils_sisr.py:22-41 evaluated directly in pixel space:
 * (A x)[i,j] = sum_{a,b} psf[a,b] x[(i-(a-ks/2)) mod H, (j-(b-ks/2)) mod W].  sf>1 adds the x`sf`
 * decimation of utils_sisr.py:44-61 (FH_OP_SR). */
int fh_op_create_conv(fh_ctx* ctx, fh_op** out, const float* psf_host, int ks, int sf);
/* Inpainting mask (measurements.py:204-246): mask_dev is (B,H,W) or (1,H,W) float32 of {0,1},
 * applied to all channels; the pointer is borrowed, not copied. */
int fh_op_create_inpaint(fh_ctx* ctx, fh_op** out, const float* mask_dev, int mask_batch);
int fh_op_destroy(fh_op* op);
/* y = A x (transpose=0; x: (B,C,H,W), y: (B,C,H/sf,W/sf)) or x = A^T y (transpose=1). */
int fh_op_apply(fh_op* op, const float* in_dev, float* out_dev, int batch, int transpose, void* stream);

/* ---- CG guidance solve ------------------------------------------------------------------------
 * Replaces conditioning_utils/cg.py:118-292 specialised to the three `*_mat_generic_customcuda_bfgs_tailored`
 * solvers (conditioning_mechanisms.py:384-419, 489-527, 641-675):
 *   b = A-side residual (inpaint: m*y - m*mu ; else y - A mu),  x0 = b,
 *   solve (sigma_y2 I + A Sigma A^T) u = b until |r| <= rtol*|b|  (or p.Ap <= 1e-16, or maxiter),
 *   mat = u (inpaint) or A^T u.
 * y_dev: (B, y-shape), mu_dev/mat_out_dev: (B,C,H,W).  niter_host/optimal_host/resid_host: B entries each
 * (host memory, may be NULL).  fixed_iters>0 disables the residual test and runs exactly that many
 * iterations (test hook).  Per-image convergence is tracked on the device; the kernels report progress into
 * pinned mapped memory and the host keeps two iterations queued ahead (no stream sync inside the loop); the call
 * returns after one final synchronisation of `stream`. */
int fh_cg_solve(fh_state* st, fh_op* op, const float* y_dev, const float* mu_dev, double sigma_y2, double rtol,
                int maxiter, int fixed_iters, float* mat_out_dev, int* niter_host, int* optimal_host,
                double* resid_host, void* stream);
/* The same solve without any host read-back: returns once the work is queued on `stream`; the outcome of image b is
 * written to info_dev[b] (device memory) when its solve ends.  This is what lets a sampler run the whole guided loop
 * without a host synchronisation (the reference synchronises twice per CG iteration, cg.py:250,259).  Available where
 * the resident kernel applies: DCT basis, all four operators (mask, separable blur, 2-D PSF, x4 super-resolution),
 * batch <= 4 unless FH_CG_RESIDENT is set; otherwise FH_ESTATE -- use fh_cg_solve.  Float32 in and out; the arithmetic
 * in between is that of fh_cg_solve_f64 unless FH_CG_RESIDENT_F32 is set (mask and separable blur only). */
typedef struct { double residual_norm; int niter; int optimal; } fh_cg_info;
int fh_cg_solve_async(fh_state* st, fh_op* op, const float* y_dev, const float* mu_dev, double sigma_y2, double rtol,
                      int maxiter, int fixed_iters, float* mat_out_dev, fh_cg_info* info_dev, void* stream);
/* The guidance solve in the reference's own precision (cg.py runs in float64): y_dev, mu_dev and mat_out_dev
 * are float64; right-hand side, CG vectors, DCTs, operator taps, low-rank products and reductions
 * are all float64 (fh_cg_res64.cuh); only the stored delta and Z are float32.  Same stopping rule and outputs as
 * fh_cg_solve / fh_cg_solve_async.  FH_ESTATE where the resident float64 kernel does not apply. */
int fh_cg_solve_f64(fh_state* st, fh_op* op, const double* y_dev, const double* mu_dev, double sigma_y2, double rtol,
                    int maxiter, int fixed_iters, double* mat_out_dev, int* niter_host, int* optimal_host,
                    double* resid_host, void* stream);
int fh_cg_solve_f64_async(fh_state* st, fh_op* op, const double* y_dev, const double* mu_dev, double sigma_y2, double rtol,
                          int maxiter, int fixed_iters, double* mat_out_dev, fh_cg_info* info_dev, void* stream);
/* bench.py's roofline: average device time (ms) of `reps` back-to-back launches of ONE kernel of the fused CG
 * iteration at the state's current rank, CUDA events on `stream`.  which: 0 fwd_rows, 1 fwd_cols, 2 inv_cols,
 * 3 inv_rows, 4 xr.  Only for operators that fold into the DCT passes (inpainting, separable blur). */
int fh_cg_kernel_time(fh_state* st, fh_op* op, int which, int reps, float* ms_out, void* stream);
/* The same for the resident solve (FH_CG_RESIDENT): average device time (ms) of `reps` launches that each run exactly
 * `iters` CG iterations of image 0 at the state's current rank.  The entry point fills its own right-hand side (a fixed
 * pseudo-random pattern) and uses sigma_y^2 = 0.01 (the reference's noise level 0.1, measurement config); convergence
 * does not end a launch early. */
int fh_cg_resident_time(fh_state* st, fh_op* op, int iters, int reps, float* ms_out, void* stream);
/* Counters for bench.py: kernels launched by this library since the last reset. */
long long fh_launch_count(int reset);
/* Per-kernel device timing for bench.py: while enabled every launch is bracketed by a cudaEvent pair on its
 * stream.  fh_prof_report writes "kernel_name launches total_ms" lines into buf and returns the bytes needed. */
/* Debug (FH_TRACE=1 in the environment at fh_ctx_create): 8 x 16 clock64 checkpoints of CTA 0 of the fused CG kernels. */
int fh_trace_read(fh_ctx* ctx, long long* out_host);
int fh_prof_enable(int on);
int fh_prof_report(char* buf, int buflen);

#ifdef __cplusplus
}
#endif
#endif
