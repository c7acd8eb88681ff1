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
 *  - every entry point runs on the device of the handle it is given and restores the calling thread's
 *    current CUDA device before it returns.
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
/* A second handle holding the columns (cells) [cell_begin, cell_end) of a staged pair, built on the device from
 * the resident arrays (no host traffic): the cell shard one rank owns in a cell-sharded fit
 * (vb_vireo_fit_sharded).  Equivalent to staging AD[:, cell_begin:cell_end], DP[:, cell_begin:cell_end]. */
int vb_counts_slice(const vb_counts* m, int64_t cell_begin, int64_t cell_end, void* stream, vb_counts** out);
/* shape / layout queries: what = 0 n_cell, 1 n_var, 2 nnz, 3 wide flag, 4 device, 5 bytes resident,
 * 6 grid.x of the cell pass (rows path), 7 grid.x of the SNP pass (rows path), 8 grid.x of the elementwise
 * (V*K) kernels; window-segment formats: 20 + 10 * table kind (0 FP64 rows of 16 columns, 1 fixed point,
 * 2 FP64 rows of 8 columns) + {0 built, 1 / 2 super-steps of the cell / SNP pass, 3 / 4 largest reads of one
 * row's stream (cell / SNP pass; fixed-point error bound = reads * 2^-33 * table range), 5 / 6 grid.x, 7 bytes,
 * 8 residual pairs, 9 stream pairs};
 * 60: why the automatic selector last served this matrix with the row kernels: 0 it did not, 1 small matrix
 *     (by design), 2 building the segment formats FAILED (vb_counts_note has the message; the row kernels are
 *     several times slower on large matrices), 3 residual-dominated counts (by design), 4 n_donor > 16;
 * 61: launches of the row-split cell pass (0: not in use): when a matrix has few cells for the SMs of the device but a
 *     long table (one rank's share of a cell-sharded fit, GT-given fits on mid-sized data) the table rows of the cell
 *     pass are cut into that many ranges that run side by side and a finish kernel adds their partial sums;
 * 62: worst row imbalance of the built window-segment formats, per mille (1000 x pairs of the longest row / mean pairs
 *     per row).  A warp task is as long as its longest row, so a row many times heavier than the mean (coverage of real
 *     data is heavy-tailed) becomes the critical path of its pass; the Python layer warns above 8000.  The builder
 *     cuts rows above twice the mean into parts when the longest exceeds four times the mean (63 counts the parts) */
int64_t vb_counts_info(const vb_counts* m, int what);
/* message of the failed format build behind vb_counts_info(m, 60) == 2 ("" otherwise) */
const char* vb_counts_note(const vb_counts* m);

/* Test hook: builds the window-segment format `table_kind` (0 FP64 rows of 16 columns, 1 fixed point, 2 FP64 rows of
 * 8 columns) if needed and checks it on the host against the staged counts: every (owner, gather row, count) pair of
 * pass `pass` (0 cell pass, 1 SNP pass) must appear exactly once, either as a record that executes while its table
 * row is inside the windows the warp holds, or in the residual list.  out4 = {pairs, errors, super-steps, null
 * slots}.  No reference counterpart (the reference keeps scipy CSC matrices, vireo_model.py:190-196). */
int vb_seg_verify(vb_counts* m, int table_kind, int pass, int64_t* out4);

/* sum over nnz(DP>0) of float32(min(log C(dp, ad), 700)), accumulated in float64.
 * Replaces np.sum(get_binom_coeff(AD, DP)) (vireoSNP/utils/vireo_base.py:7-22, vireo_model.py:313,
 * bmm_model.py:239).  `scratch` is a device buffer of >= 1024 doubles. */
int vb_binom_const(const vb_counts* m, double* scratch, double* out_host, void* stream);

/* Sizes of the per-call workspaces, in elements, for a batch of B restarts. */
typedef struct vb_ws_sizes {
    int64_t S;        /* doubles: S1 and S2, each [B, n_var, K]                 */
    int64_t W;        /* doubles: per-allele tables [B, n_var, 2, K] (segment kernels: rows of 16 or 8 columns; + fixed-point copy) */
    int64_t loglik;   /* doubles: [B, n_cell, K]                               */
    int64_t ab;       /* doubles: [B, T, 2*G] digamma differences              */
    int64_t part;     /* doubles: block partial sums                           */
    int64_t scal;     /* doubles: [B, 8] ELBO terms                            */
    int64_t ctrl;     /* int32:   [B, 8] {done, it_next, last_it, n_decrease, 3 tickets of fused tails, -} */
    int64_t rpad;     /* doubles: ID_prob in 128-byte rows [B, n_cell, 16] (segment kernels, else 0) */
    int64_t heavy;    /* doubles: residual sums [B, max(n_cell, 2 n_var), 16] (segment kernels, else 0) */
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
    /* segment-kernel workspace (may be NULL when vb_vireo_ws_sizes reports 0) */
    double *rpad, *heavy;
    /* element counts of the workspaces as the caller allocated them (copy of what vb_vireo_ws_sizes returned):
     * every entry point checks them against what the kernel family it is about to launch needs and fails with
     * VB_E_ARG instead of writing out of bounds (the family can change between sizing and launching, e.g.
     * through vb_set_path) */
    vb_ws_sizes ws;
} vb_vireo_args;

/* `stream`: formats a kernel family needs are built lazily on first use, on this stream */
int vb_vireo_ws_sizes(const vb_counts* m, int n_donor, int n_gt, int n_batch, int ase_mode, void* stream,
                      vb_ws_sizes* out);

/* Priors enter the kernels as logs, in two flavours: log(prior) as the softmax adds it
 * (vireoSNP/utils/vireo_model.py:198,218; bmm_model.py:153) and the log of the row-normalised prior that
 * scipy.stats.entropy compares against (vireo_model.py:237-238; bmm_model.py:166).  prior, log_raw, log_norm:
 * device arrays [n_row, n_col], rows normalised over n_col. */
int vb_log_prior(const double* prior, int64_t n_row, int n_col, double* log_raw, double* log_norm, void* stream);

/* Run the coordinate-ascent loop of Vireo._fit_VB (vireoSNP/utils/vireo_model.py:251-276) for a batch
 * of restarts entirely on the device: per iteration update_theta_size (:165-185), update_GT_prob
 * (:204-219), update_ID_prob (:187-201) and get_ELBO (:222-248), with the reference's convergence
 * rule evaluated on the device.  On return ctrl[b] = {done, it_next, last_it, n_decrease}: `last_it` is the
 * index of the last executed iteration, so the reference's return value is elbo[b, 0:last_it].
 * Small matrices (launch-bound iterations) replay one captured CUDA graph per group of iterations.
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
    double *rpad, *heavy;        /* segment-kernel workspace (may be NULL when vb_bmm_ws_sizes reports 0) */
    vb_ws_sizes ws;              /* allocated element counts, checked like vb_vireo_args.ws */
} vb_bmm_args;

int vb_bmm_ws_sizes(const vb_counts* m, int n_donor, int n_batch, void* stream, vb_ws_sizes* out);

/* BinomMixtureVB._fit_BV (vireoSNP/utils/bmm_model.py:178-201) for a batch of restarts:
 * update_theta_size (:133-144), get_E_logLik (:118-130), update_ID_prob (:147-154), get_ELBO (:157-175). */
int vb_bmm_fit(const vb_counts* m, const vb_bmm_args* a, void* stream);
int vb_bmm_step(const vb_counts* m, const vb_bmm_args* a, int phases, void* stream);

/* Doublet pass of predict_doublet (vireoSNP/utils/vireo_doublet.py:39-68): builds the K + K(K-1)/2
 * column tables from GT_prob [n_var,K,G] and theta (add_doublet_GT :105-136, add_doublet_theta
 * :85-102) on the device, runs the cell-major logLik pass and the softmax with the doublet prior.
 * Large matrices run the K2 columns as chunks of 16 through the window-segment kernel (one pass of the record
 * stream per chunk), small ones through the row kernels.
 *   loglik_out, prob_out: [n_cell, K2] with K2 = K + K(K-1)/2;  llr_out: [n_cell]
 *   log_prior_both: [id_prior_rows, K2]
 *   W, heavy, ab2: workspaces sized by vb_doublet_ws_sizes (elements, doubles); the call allocates nothing and
 *   does not synchronise the stream. */
typedef struct vb_doublet_ws { int64_t W, heavy, ab2; } vb_doublet_ws;
int vb_doublet_ws_sizes(const vb_counts* m, int n_donor, int n_gt, int ase_mode, void* stream, vb_doublet_ws* out);
int vb_vireo_doublet(const vb_counts* m, int n_donor, int n_gt, int ase_mode,
                     const double* gt_prob, const double* beta_mu, const double* beta_sum,
                     const double* log_prior_both, int id_prior_rows,
                     double* W, double* heavy, double* ab2, const vb_doublet_ws* ws,
                     double* loglik_out, double* prob_out, double* llr_out,
                     void* stream);

/* ---- one fit over several GPUs: cells sharded over ranks ------------------------------------
 * For the fits restart sharding cannot spread: the final fit of vireo_wrap (vireoSNP/utils/vireo_wrap.py:94)
 * and the GT-given mode (one restart, :48-50).  Rank r holds the staged columns of its cells (vb_counts_slice)
 * and their ID_prob rows; GT_prob and theta are replicated.  One EM iteration (vireo_model.py:257-264):
 *     SNP pass over the local cells            S1_r | S2_r                        (:168-170, :207-209)
 *     ONE all-reduce (sum) over the ranks      S1 | S2 | {LB_p, KL_ID} of the previous iteration
 *     ELBO + convergence rule of the previous iteration (:248, :266-274) -- identical on every rank
 *     theta sums, theta, GT update             identical on every rank            (:173-185, :211-219)
 *     cell pass over the local cells           ID_prob_r, local LB_p and KL_ID    (:190-201, :236-237)
 * i.e. the two cell-summed ELBO terms ride on the next iteration's exchange; a fit that stops at iteration t
 * has run one extra SNP pass, which only writes the S workspace.  The whole loop is enqueued on `stream` by
 * this call; the host polls the done flag every poll_every iterations, every rank at the same iterations.
 *
 * vb_comm wraps an NCCL communicator (libnccl.so.2 is bound at run time with dlopen: the library has no
 * link-time dependency on it).  Rank 0 calls vb_comm_unique_id, the 128 bytes travel to the other ranks by any
 * means (the Python layer uses torch.distributed), every rank calls vb_comm_create.  n_ranks == 1 needs no NCCL
 * and no id (the same loop without the exchange: the single-GPU parity test of the deferred ELBO). */
typedef struct vb_comm vb_comm;
#define VB_COMM_ID_BYTES 128
int vb_comm_unique_id(void* id_out);
int vb_comm_create(int device, int n_ranks, int rank, const void* id, vb_comm** out);
void vb_comm_destroy(vb_comm* c);
/* sum-all-reduce / broadcast / all-gather of device doubles on `stream` (the model-selection step and the result
 * gather of the sharded fit use them; NCCL over NVLink) */
int vb_comm_allreduce(vb_comm* c, double* buf, int64_t n, void* stream);
int vb_comm_broadcast(vb_comm* c, double* buf, int64_t n, int root, void* stream);
int vb_comm_allgather(vb_comm* c, const double* send, double* recv, int64_t n_per_rank, void* stream);
/* a: args of a one-restart batch (n_batch == 1, ase_mode == 0) over the LOCAL cells; xbuf: device workspace of
 * 2 * n_var * n_donor + 8 doubles (the exchange buffer; a->S1 / a->S2 are ignored). */
int vb_vireo_fit_sharded(const vb_counts* m_local, const vb_vireo_args* a, vb_comm* c, double* xbuf, void* stream);
/* GT update alone over sharded cells (predict_doublet's final update_GT_prob, vireo_doublet.py:75): local SNP
 * pass, all-reduce of S1 | S2, GT softmax. */
int vb_vireo_gt_sharded(const vb_counts* m_local, const vb_vireo_args* a, vb_comm* c, double* xbuf, void* stream);

/* Launch accounting.  Kernel classes: 0 k_snp, 1 k_theta, 2 k_gt, 3 k_cell, 4 k_elbo, 5 k_bmm_theta,
 * 6 k_terms, 7 helpers (row padding, log priors, doublet tables and softmax, exchange packing).  Kernels replayed
 * from a captured graph are counted per replay.  vb_launch_counts: cumulative launches per class since load.
 * vb_profile_enable(1) brackets every launch with CUDA events on its stream; vb_profile_read waits for
 * them, returns summed milliseconds and launch counts per class (arrays of 8) and clears the record. */
void vb_launch_counts(int64_t* n8);
void vb_profile_enable(int on);
int vb_profile_read(double* ms8, int64_t* n8);

/* Kernel family for the two sparse passes: 0 = automatic (window-segment kernels with FP64 tables for large
 * count matrices with n_donor <= 16, row kernels otherwise), 1 = row kernels (one warp per row, table gathered
 * from L2), 3 = window-segment kernels, FP64 tables (4 lanes per row, table windows in shared memory),
 * 4 = window-segment kernels with 32-bit fixed-point tables and exact 64-bit integer accumulation (Vireo
 *     without ASE mode; < 1e-7 per update on log-likelihoods; other models use 3).
 * Process-wide.  Workspaces sized under one family and used under another are rejected (see vb_vireo_args.ws). */
void vb_set_path(int mode);
/* 1 (default): fits of small matrices replay captured CUDA graphs; 0: plain launches */
void vb_set_graphs(int on);
/* 1 (default): inside the fit loops the theta step runs in the tail of the SNP pass and the ELBO / convergence step in
 * the tail of the cell pass (the CTA that finishes last), three launches per iteration instead of five; 0: stand-alone
 * k_theta / k_elbo launches.  Same arithmetic in the same order: the results are bit-identical. */
void vb_set_fuse(int on);

const char* vb_last_error(void);
/* library build info: "vireo_b200 <version> sm_100a" */
const char* vb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* VIREO_B200_H */
