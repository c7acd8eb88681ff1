/* vireo_b200.h -- C ABI of libvireo_b200.so
 *
 * The B200 (sm_100a) implementation of vireoSNP's variational-EM inner loop.
 * The reference has no FFI layer: its "plugin interface" for this path is the
 * Python API (vireoSNP.Vireo / vireo_wrap / BinomMixtureVB).  The package
 * vireo_b200/ mirrors that API and calls the entry points below through
 * ctypes; INTEGRATION.md shows the binding a vireoSNP maintainer would add.
 * Every entry point names the reference code it replaces
 * (paths relative to the reference repository root, vireoSNP v0.5.9).
 *
 * Conventions
 *  - plain C types only; every function returns VB_OK (0) or a negative
 *    VB_E_* code and never throws; vb_last_error() gives the message of the
 *    last failure on the calling thread.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *  - pointers inside vb_vireo_args / vb_bmm_args are DEVICE pointers owned by
 *    the caller (the Python layer uses torch tensors as containers); the
 *    library allocates nothing on the per-iteration path.
 *  - matrices are float64, C order; B = n_batch independent restarts are laid
 *    out with the restart index outermost.
 *  - a vb_counts handle is not thread-safe; use one per (process, GPU).
 */
#ifndef VIREO_B200_H
#define VIREO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VB_OK            0
#define VB_E_CUDA       -1   /* a CUDA runtime call or kernel failed            */
#define VB_E_ARG        -2   /* bad argument (shape, dtype code, NULL pointer)   */
#define VB_E_PATTERN    -3   /* AD has an entry outside DP's sparsity pattern    */
#define VB_E_VALUE      -4   /* negative / non-integer / too large count         */
#define VB_E_UNSUPPORTED -5  /* n_donor or n_GT beyond the compiled kernels      */

/* dtype codes for host arrays handed to vb_counts_create */
#define VB_I32 0
#define VB_I64 1
#define VB_F32 2
#define VB_F64 3

/* phase bits for vb_vireo_step / vb_bmm_step (one EM iteration = all of them, in this order) */
#define VB_PH_SNP        1   /* S1 = AD @ ID_prob, S2 = (DP-AD) @ ID_prob  (SNP-major pass)          */
#define VB_PH_THETA      2   /* theta posterior update from S1,S2 and the OLD GT_prob                */
#define VB_PH_GT         4   /* GT_prob update                                                        */
#define VB_PH_ID         8   /* logLik_ID (cell-major pass) + ID_prob softmax                         */
#define VB_PH_ELBO      16   /* ELBO from the current state and the logLik_ID buffer                  */
#define VB_PH_LOGLIK    32   /* logLik_ID only: fill the buffer, do not touch ID_prob                 */
#define VB_PH_THETA_SUMS 64  /* recompute the theta update's sums over S1*GT, S2*GT from the S1/S2 buffers
                              * (cell-sharded fit: VB_PH_SNP alone, all-reduce S1/S2 across devices, then this
                              * together with VB_PH_THETA / GT / ID / ELBO)                           */

#define VB_MAX_GT        8   /* largest n_GT the theta kernels hold in registers                      */
#define VB_MAX_DONOR   256   /* largest n_donor of the fused cell kernel (32 lanes x 8 registers)     */

typedef struct vb_counts vb_counts;   /* staged AD/DP, both orientations, resident in HBM */

/* Stage AD and DP once.  Replaces the per-call sparse algebra of the reference:
 * `BD = DP - AD` (vireoSNP/utils/vireo_model.py:168,190,228; bmm_model.py:122,136) is folded into a
 * per-nnz (ad, dp) record, and the CSC (cell-major) layout the readers produce
 * (vireoSNP/utils/io_utils.py:57) is transposed on the device to add the SNP-major orientation that
 * `AD @ ID_prob` (vireo_model.py:169-170,207-208) walks.
 *
 * Inputs are HOST arrays of scipy CSC matrices of shape (n_var, n_cell) with sorted indices:
 * indptr has n_cell+1 entries, indices are SNP ids.  pattern(AD) must be a subset of pattern(DP).
 * Counts must be non-negative integers < 2^31; values above 65535 select the wide (12 B/nnz) record. */
int vb_counts_create(int device, int64_t n_cell, int64_t n_var,
                     const void* dp_indptr, int indptr_dtype,
                     const void* dp_indices, int indices_dtype,
                     const void* dp_data, int data_dtype, int64_t dp_nnz,
                     const void* ad_indptr, const void* ad_indices, const void* ad_data, int64_t ad_nnz,
                     void* stream, vb_counts** out);
void vb_counts_destroy(vb_counts* m);
/* shape / layout queries: what = 0 n_cell, 1 n_var, 2 nnz, 3 wide flag, 4 device, 5 bytes resident,
 * 6 grid.x of the cell pass (rows path), 7 grid.x of the SNP pass (rows path), 8 grid.x of the elementwise
 * (V*K) kernels, 9 gather-stream formats built, 10 / 11 stream records of the cell / SNP pass, 12 residual
 * (count >= 32) pairs, 13 bytes of the gather formats, 14 / 15 grid.x of the gather cell / SNP pass,
 * 16 pairs carried by the streams; window-segment formats: 20 + 10 * precision (0 FP64, 1 fixed point) +
 * {0 built, 1 / 2 super-steps of the cell / SNP pass, 3 / 4 largest reads of one row's stream (cell / SNP pass;
 * fixed-point error bound = reads * 2^-33 * table range), 5 / 6 grid.x, 7 bytes, 8 residual pairs, 9 stream pairs} */
int64_t vb_counts_info(const vb_counts* m, int what);

/* sum over nnz(DP>0) of float32(min(log C(dp, ad), 700)), accumulated in float64.
 * Replaces np.sum(get_binom_coeff(AD, DP)) (vireoSNP/utils/vireo_base.py:7-22, vireo_model.py:313,
 * bmm_model.py:239).  `scratch` is a device buffer of >= 1024 doubles. */
int vb_binom_const(const vb_counts* m, double* scratch, double* out_host, void* stream);

/* Sizes of the per-call workspaces, in elements, for a batch of B restarts. */
typedef struct vb_ws_sizes {
    int64_t S;        /* doubles: S1 and S2, each [B, n_var, K]                 */
    int64_t W;        /* doubles: per-allele tables [B, n_var, 2, K] (gather / segment paths: K -> 16; + fixed-point copy) */
    int64_t loglik;   /* doubles: [B, n_cell, K]                               */
    int64_t ab;       /* doubles: [B, T, 2*G] digamma differences              */
    int64_t part;     /* doubles: block partial sums                           */
    int64_t scal;     /* doubles: [B, 8] ELBO terms                            */
    int64_t ctrl;     /* int32:   [B, 4] {done, it, n_decrease, hit_max}       */
    int64_t rpad;     /* doubles: ID_prob in 128-byte rows [B, n_cell, 16] (gather path, else 0) */
    int64_t heavy;    /* doubles: residual sums [B, max(n_cell, 2 n_var), 16] (gather path, else 0) */
} vb_ws_sizes;

typedef struct vb_vireo_args {
    int32_t n_donor, n_gt, n_batch;
    int32_t ase_mode, learn_gt, learn_theta, fix_beta_sum;
    int32_t id_prior_rows;       /* 1 (broadcast over cells) or n_cell                                */
    int32_t theta_prior_rows;    /* 1 or theta rows (n_var in ASE mode)                               */
    int32_t max_iter, min_iter, delay_fit_theta;
    int32_t poll_every;          /* host checks the done flags every this many iterations (0 = 16)    */
    int32_t reserved;
    double  epsilon_conv;
    /* state, in/out */
    double* id_prob;             /* [B, n_cell, K]                                                    */
    double* gt_prob;             /* [B, n_var, K, G]                                                  */
    double* beta_mu;             /* [B, T, G]   T = n_var if ase_mode else 1                          */
    double* beta_sum;            /* [B, T, G]                                                         */
    /* priors, shared by the batch */
    const double* log_id_prior;     /* [id_prior_rows, K] log(ID_prior) as used in the softmax         */
    const double* log_id_prior_kl;  /* same shape, log of the row-normalised prior (scipy.stats.entropy)*/
    const double* log_gt_prior;     /* [n_var, K, G]                                                   */
    const double* log_gt_prior_kl;  /* [n_var, K, G]                                                   */
    const double* s1_prior;         /* [theta_prior_rows, G]                                           */
    const double* s2_prior;
    /* workspace (see vb_vireo_ws_sizes) */
    double *S1, *S2, *W, *loglik, *ab, *part, *scal;
    int32_t* ctrl;
    /* outputs */
    double* elbo;                /* [B, max_iter] every computed ELBO (the reference returns ELBO[:it]) */
    /* gather-path workspace (may be NULL when vb_vireo_ws_sizes reports 0) */
    double *rpad, *heavy;
} vb_vireo_args;

int vb_vireo_ws_sizes(const vb_counts* m, int n_donor, int n_gt, int n_batch, int ase_mode, vb_ws_sizes* out);

/* Priors enter the kernels as logs, in two flavours: log(prior) as the softmax adds it
 * (vireoSNP/utils/vireo_model.py:198,218; bmm_model.py:153) and the log of the row-normalised prior that
 * scipy.stats.entropy compares against (vireo_model.py:237-238; bmm_model.py:166).  prior, log_raw, log_norm:
 * device arrays [n_row, n_col], rows normalised over n_col. */
int vb_log_prior(const double* prior, int64_t n_row, int n_col, double* log_raw, double* log_norm, void* stream);

/* Run the coordinate-ascent loop of Vireo._fit_VB (vireoSNP/utils/vireo_model.py:251-276) for a batch
 * of restarts entirely on the device: per iteration update_theta_size (:165-185), update_GT_prob
 * (:204-219), update_ID_prob (:187-201) and get_ELBO (:222-248), with the reference's convergence
 * rule evaluated on the device.  On return ctrl[b] = {done, it, ...}: `it` is the index of the last
 * executed iteration, so the reference's return value is elbo[b, 0:it].
 * The binomial constant (vireo_model.py:313) is NOT added here; see vb_binom_const. */
int vb_vireo_fit(const vb_counts* m, const vb_vireo_args* a, void* stream);

/* Run selected phases once (teacher-forced single updates: Vireo.update_theta_size / update_GT_prob /
 * update_ID_prob / get_ELBO as separate calls).  ELBO terms land in scal[b, 0:5] =
 * {ELBO, LB_p, KL_ID, KL_GT, KL_theta}. */
int vb_vireo_step(const vb_counts* m, const vb_vireo_args* a, int phases, void* stream);

typedef struct vb_bmm_args {
    int32_t n_donor, n_batch;
    int32_t fix_beta_sum;
    int32_t id_prior_rows;
    int32_t max_iter, min_iter;
    int32_t poll_every, reserved;
    double  epsilon_conv;
    double* id_prob;             /* [B, n_cell, K]                                                    */
    double* beta_mu;             /* [B, n_var, K]                                                     */
    double* beta_sum;            /* [B, n_var, K]                                                     */
    const double* log_id_prior;
    const double* log_id_prior_kl;
    const double* s1_prior;      /* [n_var, K]                                                        */
    const double* s2_prior;
    double *S1, *S2, *W, *loglik, *part, *scal;
    int32_t* ctrl;
    double* elbo;                /* [B, max_iter]                                                     */
    double *rpad, *heavy;        /* gather-path workspace (may be NULL when vb_bmm_ws_sizes reports 0) */
} vb_bmm_args;

int vb_bmm_ws_sizes(const vb_counts* m, int n_donor, int n_batch, vb_ws_sizes* out);

/* BinomMixtureVB._fit_BV (vireoSNP/utils/bmm_model.py:178-201) for a batch of restarts:
 * update_theta_size (:133-144), get_E_logLik (:118-130), update_ID_prob (:147-154), get_ELBO (:157-175). */
int vb_bmm_fit(const vb_counts* m, const vb_bmm_args* a, void* stream);
int vb_bmm_step(const vb_counts* m, const vb_bmm_args* a, int phases, void* stream);

/* Doublet pass of predict_doublet (vireoSNP/utils/vireo_doublet.py:39-68): builds the K + K(K-1)/2
 * column tables from GT_prob [n_var,K,G] and theta (add_doublet_GT :105-136, add_doublet_theta
 * :85-102) on the device, runs the cell-major logLik pass and the softmax with the doublet prior.
 *   loglik_out, prob_out: [n_cell, K2] with K2 = K + K(K-1)/2;  llr_out: [n_cell]
 *   W: workspace [n_var, 2, K2];  log_prior_both: [id_prior_rows, K2]. */
int vb_vireo_doublet(const vb_counts* m, int n_donor, int n_gt, int ase_mode,
                     const double* gt_prob, const double* beta_mu, const double* beta_sum,
                     const double* log_prior_both, int id_prior_rows,
                     double* W, double* loglik_out, double* prob_out, double* llr_out,
                     void* stream);

/* Launch accounting.  Kernel classes: 0 k_snp, 1 k_theta, 2 k_gt, 3 k_cell, 4 k_elbo, 5 k_bmm_theta,
 * 6 k_terms, 7 doublet helpers.  vb_launch_counts: cumulative launches per class since load.
 * vb_profile_enable(1) brackets every launch with CUDA events on its stream; vb_profile_read waits for
 * them, returns summed milliseconds and launch counts per class (arrays of 8) and clears the record. */
void vb_launch_counts(int64_t* n8);
void vb_profile_enable(int on);
int vb_profile_read(double* ms8, int64_t* n8);

/* Kernel family for the two sparse passes: 0 = automatic (window-segment kernels with FP64 tables for large
 * count matrices with n_donor <= 16, row kernels otherwise), 1 = row kernels (one warp per row, table gathered
 * from L2), 2 = gather-stream kernels (lane per row, table streamed through a shared-memory ring),
 * 3 = window-segment kernels, FP64 tables (4 lanes per row, table windows in shared memory),
 * 4 = window-segment kernels with 32-bit fixed-point tables and exact 64-bit integer accumulation (Vireo
 *     without ASE mode; < 1e-7 per update on log-likelihoods; other models use 3).
 * Process-wide; affects workspaces sized afterwards. */
void vb_set_path(int mode);

const char* vb_last_error(void);
/* library build info: "vireo_b200 <version> sm_100a" */
const char* vb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* VIREO_B200_H */
