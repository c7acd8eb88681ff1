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
#define VB_CTRL_N 4   // {done, it_next, last_it, n_decrease}
#define VB_SCAL_N 8   // {ELBO, LB_p, KL_ID, KL_GT, KL_theta, -, -, -}

// ----------------------------------------------------------------------------------------------
// Tiled ("brick") format, one per pass orientation ("tiled kernels"):
//   every nnz with 1 <= dp <= VB_EXPAND_MAX is expanded into dp UNIT records (ad of them carry the
//   alternative allele), so a record is just "add gather row g into owner row o":
//     cell pass: owner = cell j,               gather row = 2*snp + allele   (rows of the Wt table)
//     SNP  pass: owner = 2*snp + allele,       gather row = cell j           (rows of ID_prob)
//   Records are grouped into tiles (owner block of `rpb` rows) x (gather slab of `slab_rows` rows) and
//   stored tile by tile as uint16 slab-local gather indices; each tile starts on a 16-byte boundary.
//   seg[tile][0..rpb] are tile-local record offsets of the owner rows (row stride rpb+4, 16 B aligned).
//   nnz with dp > VB_EXPAND_MAX stay in a small residual in the row (v1) layout ("heavy" stream).
// ----------------------------------------------------------------------------------------------
#define VB_EXPAND_MAX 4
#define VB_TILE_THREADS 512
#define VB_TILE_WARPS (VB_TILE_THREADS / 32)
#define VB_TILE_MAX_DONOR 32

struct TileSet {
    int ok;                 // 0: this orientation cannot use the tiled kernels (reason in vb_last_error)
    int64_t n_owner, n_gather;
    int rpb, cpg, nb, nslab, slab_rows;
    int gl;                 // lanes per group (each lane holds 2 donors)
    int cpg_t;              // compiled accumulator rows per group (4 or 12)
    int recb_max;           // largest tile record payload in bytes (multiple of 16)
    int smem_bytes;         // dynamic shared memory of the kernel
    int64_t n_rec;          // records incl. per-tile padding
    uint32_t* tile_start;   // [nb*nslab + 1] record index where each tile starts
    uint32_t* seg;          // [nb*nslab][rpb+4]
    uint16_t* rec;          // [n_rec]
};

struct TileView {
    int64_t n_owner, n_gather;
    int rpb, cpg, nslab, slab_rows, gstride;   // gstride: doubles per gather row (= n_donor)
    uint32_t slab_bytes, rec_off, ptr_off, buf_stride;
    const uint32_t* __restrict__ tile_start;
    const uint32_t* __restrict__ seg;
    const uint16_t* __restrict__ rec;
};

struct TilePair {
    int K;
    TileSet A;   // cell pass
    TileSet B;   // SNP pass
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
    // residual of nnz with dp > VB_EXPAND_MAX (or ad > dp), same row layouts; built with the first tile set
    int heavy_built;
    int64_t Nh, N_unit, N_unit_rec;
    int64_t* hcell_ptr;
    int32_t* hcell_idx;
    uint32_t* hcell_cnt;
    uint32_t* hcell_dp;
    int64_t* hsnp_ptr;
    int32_t* hsnp_idx;
    uint32_t* hsnp_cnt;
    uint32_t* hsnp_dp;
    // tile sets are specific to n_donor (slab size); a few are cached
    TilePair tiles[4];
    int n_tiles;
    int tile_mode;          // 0 auto, 1 force rows (v1), 2 force tiles
};

// vb_tiles.cu
const TilePair* vb_tiles_get(vb_counts* m, int K, cudaStream_t st);   // nullptr when the tiled path is not usable
void vb_tiles_free(vb_counts* m);

// what the kernels see of the staged counts
struct CountsView {
    int64_t C, V, N;
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
    // block-partial layout inside part[b * part_stride + ...]
    int64_t part_stride;
    int off_theta, off_klgt, off_cell, off_klth;
    int n_snpblk, n_elemblk, n_cellblk, n_klth;
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
