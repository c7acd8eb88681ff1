// vb_common.cuh -- shared definitions for libvireo_b200 (sm_100a).
//
// Data layout in HBM:
//   counts, cell-major : cell_ptr[C+1] int64, cell_idx[N] int32 (SNP id), cell_cnt[N] uint32 (ad | dp<<16)
//   counts, SNP-major  : snp_ptr[V+1]  int64, snp_idx[N]  int32 (cell id), snp_cnt[N]  uint32
//   wide variant (any count > 65535): *_cnt holds ad, *_dp holds dp (12 B per nnz instead of 8)
//   dense state, float64, restart index outermost: ID_prob [B,C,K], GT_prob [B,V,K,G], tables Wa/Wb [B,V,K].
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vireo_b200.h"

#define VB_THREADS 256
#define VB_WARPS (VB_THREADS / 32)
#define VB_FULL 0xffffffffu
#define VB_CTRL_N 8   // {done, it_next, last_it, n_decrease, ticket of the SNP pass, of the cell pass, of k_theta_sums, -}
#define VB_SCAL_N 8   // {ELBO, LB_p, KL_ID, KL_GT, KL_theta, -, -, -}

// ----------------------------------------------------------------------------------------------
// Gather tables: both sparse passes are "for every owner row: sum over its pairs of count * table[gather row]"
//     cell pass (A): owner = cell j,            gather row = 2*snp + allele  (rows of the W table)
//     SNP  pass (B): owner = 2*snp + allele,    gather row = cell j          (rows of ID_prob)
//   with allele 0 = reference reads (count dp-ad), allele 1 = alternative reads (count ad).
// ----------------------------------------------------------------------------------------------
#define VB_ROW_DOUBLES 16                 // a full gather-table row is 16 doubles = 128 bytes (all 32 banks)
#define VB_SPARSE_LEN 64                  // owners with this few pairs may be routed to the residual kernel

// ----------------------------------------------------------------------------------------------
// Window-segment format (vb_seg.cu), one per pass orientation and table precision.
//   The gather table streams through a ring of `nb` window buffers of `win_rows` rows in shared memory: window w
//   lives in buffer w % nb, so row r always sits at ring row r % (nb * win_rows) -- a record carries that ring row and
//   the kernel needs no per-window base.  Owners are sorted by pair count and packed VB_SEG_OWNERS = 32 to a warp
//   task.  A task's stream is a sequence of "super-steps" of 32 records (one per owner slot, 128 bytes).  While a warp
//   is in *segment* s it holds the windows s .. s + span - 1; a segment carries every pair of window s that is still
//   open and, in the slots lock-step would otherwise pad, pairs of the next span - 1 windows (earliest window first).
//   Record (32 bits):
//       bits 31..28  windows to advance BEFORE this super-step (only in the first slot of every lane group, 0 elsewhere)
//       bits 27..16  ring row
//       bits 15..0   FP64 tables: the upper 16 bits of the double 2 * count (sign, exponent, 4 mantissa bits: every
//                    count with at most 5 significant bits is exact); fixed-point tables: the count (<= 31);  0 = null
//   Pairs whose count has no such code and all pairs of very short owners go to a residual CSR.
//   The kernel binds 4 lanes to one owner slot (4 columns per lane; 2 for the 8-column FP64 tables), so one warp
//   step serves 8 owners and the 32 slots of a super-step take 4 steps.
// ----------------------------------------------------------------------------------------------
#define VB_SEG_OWNERS 32
#define VB_SEG_MAX_WARPS 22               // consumer warps per CTA (+1 producer warp)
#define VB_SEG_DEPTH 4                    // super-steps per record chunk: a task's stream is padded to a multiple of it
#define VB_SEG_ADV_MAX 15                 // window advances one super-step can carry
#define VB_SEG_MAX_RING_ROWS 4096         // 12-bit ring row
#define VB_SEG_MAX_NB 16

struct SegSet;
struct SegSet {
    int built;
    int64_t n_owner, n_gather;
    int64_t n_task;          // ceil(n_owner / 32) tasks in sorted order (longest owners first)
    int64_t n_task_stream;   // the first n_task_stream tasks carry records; the rest only need the epilogue
    int n_win, win_rows, nb; // windows of the table, rows per window, window buffers in shared memory
    int span;                // windows a warp holds at a time (1 + look-ahead windows)
    int fixed;               // count code of the records: 0 upper half of the double 2 * count, 1 the integer
    int64_t n_step;          // super-steps in total
    int64_t n_light;         // pairs carried by the streams
    int64_t n_heavy;         // pairs in the residual
    int64_t max_reads;       // largest sum of counts over one owner's stream pairs (fixed-point error bound)
    int64_t max_len;         // stream pairs of the longest owner (a task is as long as its longest owner)
    int32_t* perm;           // [n_task*32] owner id of each slot, -1 = padding
    int64_t* task_off;       // [n_task_stream + 1] first super-step of each task (multiples of VB_SEG_DEPTH)
    int32_t* tail;           // [n_task_stream] window advances left when a task's stream ends
    // Rows far heavier than the rest (a task is as long as its longest owner): owner o keeps every o_split[o]-th of its
    // stream pairs, the others go to o_split[o] - 1 *virtual owners* of the format `excess`, whose plain sums (vsum) are
    // added to the residual sums H[o] before this format's kernel runs.  nullptr / 0 when no row needed it.
    uint8_t* o_split;        // [n_owner] parts per owner (1: whole)
    int64_t n_virtual;       // owners of `excess`
    int32_t* v_owner;        // [n_virtual] the real owner of a virtual owner (ascending)
    uint8_t* v_part;         // [n_virtual] which part of it (1 .. o_split - 1)
    SegSet* excess;
    double* vsum;            // [B, n_virtual, RW] plain sums of the virtual owners, grown on demand
    int64_t vsum_elems;
    uint32_t* rec;           // [n_step*32 + slack]
    int64_t* hptr;           // [n_owner+1] residual CSR
    int32_t* hrow;
    uint32_t* hcnt;
    int grid, nwarps;
    int64_t bytes;
};

struct SegView {
    int64_t n_owner, n_gather, n_task, n_task_stream;
    int n_win, win_rows, nb, span;
    const int32_t* __restrict__ perm;
    const int64_t* __restrict__ task_off;
    const int32_t* __restrict__ tail;
    const uint32_t* __restrict__ rec;
    const int64_t* __restrict__ hptr;
    const int32_t* __restrict__ hrow;
    const uint32_t* __restrict__ hcnt;
};

// Row-split cell pass (vb_seg.cu): when a matrix has few owner tasks for the SMs of the device but a long table (one
// rank's share of a cell-sharded fit, GT-given fits on mid-sized data), the table rows are cut into R ranges with one
// segment format each; the R launches run side by side on R streams, every CTA streams 1/R of the table, and a small
// kernel adds the partial sums and finishes the cell update.
#define VB_SEG_MAX_SPLIT 8
struct SegSplit {
    int R;                                  // 0: not built / not in use
    int failed;
    SegSet set[VB_SEG_MAX_SPLIT];
    int64_t row_lo[VB_SEG_MAX_SPLIT + 1];   // table rows [row_lo[r], row_lo[r+1]) belong to range r
    cudaStream_t aux[VB_SEG_MAX_SPLIT];
    cudaEvent_t fork, join[VB_SEG_MAX_SPLIT];
};

struct vb_counts {
    int device;
    int sm_count;
    int64_t C, V, N;
    int wide;
    int64_t* cell_ptr;
    int32_t* cell_idx;
    uint32_t* cell_cnt;
    uint32_t* cell_dp;
    int64_t* snp_ptr;
    int32_t* snp_idx;
    uint32_t* snp_cnt;
    uint32_t* snp_dp;
    int grid_cell, grid_snp, grid_elem;
    int64_t bytes;
    // window-segment formats (vb_seg.cu), index = table kind: 0 FP64 rows of 128 B (16 columns), 1 fixed-point rows
    // of 64 B (16 columns), 2 FP64 rows of 64 B (8 columns, n_donor <= 8)
    SegSet sA[3];           // cell pass
    SegSet sB[3];           // SNP pass
    SegSplit rA[3];         // cell pass, table rows split over R launches
    int seg_failed[3];
    // why the automatic selector last chose the row kernels for this matrix: 0 it did not, 1 small matrix (the passes
    // are launch/latency bound either way), 2 building the segment formats failed (message kept in seg_error),
    // 3 most pairs carry counts > 31 (residual-dominated, e.g. mitochondrial clone data), 4 n_donor > 16
    int auto_fallback;
    char seg_error[256];
};

// every extern "C" entry point runs on the handle's device and leaves the caller's current device untouched
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// vb_seg.cu
int vb_seg_build(vb_counts* m, int prec, cudaStream_t st);   // builds sA[prec] and sB[prec] once
void vb_seg_free(vb_counts* m);

// what the kernels see of the staged counts
struct CountsView {
    int64_t C, V, N;
    int64_t g_lo, g_hi;        // segment-format builders: only pairs with a gather row in [g_lo, g_hi), rebased to g_lo
    // segment-format builders, split owners (SegSet::o_split): owner index -> real owner and the part it enumerates
    int fixed;                 // count code of the format (which counts ride in the stream)
    const uint8_t* __restrict__ o_split;
    const int32_t* __restrict__ v_owner;
    const uint8_t* __restrict__ v_part;
    const int64_t* __restrict__ cell_ptr;
    const int32_t* __restrict__ cell_idx;
    const uint32_t* __restrict__ cell_cnt;
    const uint32_t* __restrict__ cell_dp;
    const int64_t* __restrict__ snp_ptr;
    const int32_t* __restrict__ snp_idx;
    const uint32_t* __restrict__ snp_cnt;
    const uint32_t* __restrict__ snp_dp;
};

// EM parameters shared by the Vireo and the binomial-mixture kernels
struct EmP {
    int64_t C, V, T;           // cells, SNPs, theta rows (1 or V)
    int K, G, B;
    int bmm;                   // 1: BinomMixtureVB (theta per (SNP, clone), no genotype layer)
    int ase, learn_gt, learn_theta, fix_beta_sum;
    int id_rows, thp_rows;
    int max_iter, min_iter, delay;
    double eps;
    double *R, *GT, *mu, *sum;
    const double *lidp, *lidp_kl, *lgtp, *lgtp_kl, *s1p, *s2p;
    double *S1, *S2, *Wt, *ll, *ab, *part, *scal, *elbo;
    int* ctrl;
    // segment path (vb_seg.cu): Wt is then [B, 2V, RW] (rows of 128 / 64 bytes, columns replicated RW/KT times),
    // RP the same layout of ID_prob [B, C, 16], H the residual sums [B, max(C, 2V), 16]
    int RW;                    // doubles per row of the padded tables (Wt, RP, H): 16, or 8 for the narrow FP64 segment kernels
    int tiled, KT;             // tiled: 0 row kernels, 2 window-segment kernels (FP64 tables),
                               //        3 window-segment kernels (32-bit fixed-point tables)
    double *RP, *H;
    uint32_t *Wq, *RPq;        // tiled == 3: fixed-point copies of Wt [B, 2V, 16] and RP [B, C, 16]
    double* qscale;            // tiled == 3: [B] power of two with |W| * qscale < 2^32
    // block-partial layout inside part[b * part_stride + ...]
    int64_t part_stride;
    int off_theta, off_klgt, off_cell, off_klth;
    int n_snpblk, n_elemblk, n_cellblk, n_klth;
    // tails of the sparse passes (vb_tail.cuh): 0 none, 1 theta after the SNP pass + ELBO after the cell pass (fit
    // loop), 2 exchange packing after the cell pass (cell-sharded fit; xs = two doubles behind S1 | S2)
    int fuse;
    double* xs;
};

// ----------------------------------------------------------------------------------------------
// scalar math shared by host tests and device code
// ----------------------------------------------------------------------------------------------

// digamma for x > 0: upward recurrence to x >= 10, then the asymptotic series through x^-14
// (next term 3617/8160 x^-16 < 5e-17).  Stands in for scipy.special.digamma
// (vireoSNP/utils/vireo_model.py:152-162, bmm_model.py:126-128, vireo_base.py:102-104).
__host__ __device__ inline double vb_digamma(double x) {
    if (!(x > 0.0)) return NAN;
    double r = 0.0;
    while (x < 10.0) {
        r -= 1.0 / x;
        x += 1.0;
    }
    const double f = 1.0 / (x * x);
    const double t = f * (-1.0 / 12 + f * (1.0 / 120 + f * (-1.0 / 252 + f * (1.0 / 240 + f * (-1.0 / 132 +
                     f * (691.0 / 32760 + f * (-1.0 / 12)))))));
    return r + log(x) - 0.5 / x + t;
}

__host__ __device__ inline double vb_betaln(double a, double b) { return lgamma(a) + lgamma(b) - lgamma(a + b); }

// KL(Beta(p1,p2) || Beta(q1,q2)) written as the reference does: cross(p,q) - cross(p,p)
// (vireoSNP/utils/vireo_base.py:96-127).  psi1/psi2/psis are digamma(p1), digamma(p2), digamma(p1+p2).
__host__ __device__ inline double vb_beta_kl(double p1, double p2, double q1, double q2,
                                             double psi1, double psi2, double psis) {
    const double cq = vb_betaln(q1, q2) - (q1 - 1.0) * psi1 - (q2 - 1.0) * psi2 + ((q1 + q2) - 2.0) * psis;
    const double cp = vb_betaln(p1, p2) - (p1 - 1.0) * psi1 - (p2 - 1.0) * psi2 + ((p1 + p2) - 2.0) * psis;
    return cq - cp;
}

// float32(min(log C(d, a), 700)) as get_binom_coeff does (vireoSNP/utils/vireo_base.py:14-20).
__host__ __device__ inline float vb_binom_term(uint32_t a, uint32_t d) {
    if (a > d) return -INFINITY;      // binom() = 0 -> log = -inf in the reference
    if (a == 0 || a == d) return 0.0f;
    double v = lgamma((double)d + 1.0) - lgamma((double)a + 1.0) - lgamma((double)(d - a) + 1.0);
    if (v > 700.0) v = 700.0;
    return (float)v;
}

#ifdef __CUDACC__
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(VB_FULL, v, off);
    return v;
}

// deterministic block sum; result valid on thread 0.  `sh` must hold VB_WARPS doubles.
__device__ __forceinline__ double block_sum(double v, double* sh) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        for (int i = 0; i < nw; ++i) t += sh[i];
    }
    return t;
}
#endif

#ifdef __CUDACC__
// theta_mode: 0 never, 1 always, 2 when learn_theta and the device iteration counter >= delay
__device__ __forceinline__ bool vb_theta_on(const EmP& p, int b, int theta_mode) {
    if (theta_mode == 1) return true;
    if (theta_mode == 2) return p.learn_theta && p.ctrl[b * VB_CTRL_N + 1] >= p.delay;
    return false;
}
#endif

// launch accounting (vb_em.cu): every kernel launch of the EM path goes through VB_LAUNCH, which counts it
// and, when profiling is enabled, brackets it with CUDA events on the launching stream.
// classes: 0 SNP pass, 1 k_theta, 2 k_gt, 3 cell pass, 4 k_elbo, 5 k_bmm_theta, 6 k_terms, 7 helpers
void vb_launch_begin(int cls, cudaStream_t st, cudaEvent_t* e0, cudaEvent_t* e1);
void vb_launch_end(int cls, cudaStream_t st, cudaEvent_t e0, cudaEvent_t e1);
#define VB_LAUNCH(cls, st, ...)                                         \
    do {                                                                \
        cudaEvent_t e0__ = nullptr, e1__ = nullptr;                     \
        vb_launch_begin(cls, st, &e0__, &e1__);                         \
        __VA_ARGS__;                                                    \
        vb_launch_end(cls, st, e0__, e1__);                             \
    } while (0)

// vb_seg.cu
int vb_pad_rows_launch(const vb_counts* m, const double* src, int64_t n_row, int K, int KT, int RW, int B, double* dst,
                       cudaStream_t st);
enum { GM_CELL = 0, GM_CELL_LL = 1, GM_SNP = 2, GM_PLAIN = 3 };
void vb_seg_geometry(const SegSet& g, int* grid, int* nwarps);
// GM_PLAIN: out[owner * ld + off + column], column < cols; `set` != nullptr: this format instead of sA[prec], its table
// starts `row0` rows into the full table
struct SegPlain { double* out; int64_t ld; int off, cols; const SegSet* set; int64_t row0; };
int vb_seg_launch(const vb_counts* m, const EmP& p, int ori, int mode, int theta_mode, const SegPlain* plain, cudaStream_t st);
int vb_seg_split_rule(const vb_counts* m, int prec);                       // R the row-split cell pass would use (1: none)
int vb_seg_build_split(vb_counts* m, int prec, cudaStream_t st);           // builds rA[prec] once (R from the rule)
int vb_seg_launch_cell_split(const vb_counts* m, const EmP& p, int mode, cudaStream_t st);
int vb_seg_quantise_rows(const vb_counts* m, const EmP& p, cudaStream_t st);   // RP -> RPq before the first SNP pass

// error plumbing -------------------------------------------------------------------------------
void vb_set_error(const char* fmt, ...);
#define VB_CUDA(call)                                                                           \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            vb_set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return VB_E_CUDA;                                                                   \
        }                                                                                       \
    } while (0)
