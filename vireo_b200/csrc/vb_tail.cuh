// vb_tail.cuh -- the two tiny single-CTA steps of an EM iteration as device functions, so that the sparse passes
// can run them in the tail of their own launch (the CTA that finishes last) instead of in launches of their own:
//   theta_block  the theta posterior, digamma tables and KL_theta from the block partials of the SNP pass
//                (vireoSNP/utils/vireo_model.py:149-185, vireo_base.py:96-127)            [kernel k_theta]
//   elbo_block   the ELBO sum and the convergence rule (vireo_model.py:248,266-274; bmm_model.py:175,190-199)
//                                                                                         [kernel k_elbo]
// Small matrices are launch-bound (an iteration of the 10k x 5k x 4 configuration is five launches of 5-20 us), and
// so is the cell-sharded fit on many GPUs; the arithmetic and its order are exactly those of the stand-alone kernels,
// so fused and unfused iterations give the same bits.
#pragma once
#include "vb_common.cuh"

// true in exactly one CTA of the launch (per restart): the one whose ticket is drawn last.  The counter re-arms itself.
__device__ __forceinline__ bool vb_last_cta(int* counter) {
    __shared__ int s_last;
    __threadfence();                      // this CTA's partial sums are visible before its ticket is
    __syncthreads();
    if (threadIdx.x == 0) {
        const int t = atomicAdd(counter, 1);
        s_last = t == (int)gridDim.x - 1;
        if (s_last) *counter = 0;
    }
    __syncthreads();
    if (s_last) __threadfence();
    return s_last != 0;
}

struct ThetaOut { double A, B, kl; };

__device__ __forceinline__ ThetaOut theta_finish(const EmP& p, bool do_theta, double sum1, double sum2, double q1,
                                                 double q2, double* mu_io, double* sum_io) {
    double mu = *mu_io, sm = *sum_io;
    if (do_theta) {
        const double s1 = q1 + sum1, s2 = q2 + sum2;          // vireo_model.py:173-181
        mu = s1 / (s1 + s2);                                  // :183
        if (!p.fix_beta_sum) sm = s1 + s2;                    // :184-185
        *mu_io = mu;
        *sum_io = sm;
    }
    const double e1 = mu * sm, e2 = (1.0 - mu) * sm;          // theta_s1 / theta_s2 (:139-147)
    const double es = e1 + e2;
    const double psi1 = vb_digamma(e1), psi2 = vb_digamma(e2), psis = vb_digamma(es);
    ThetaOut o;
    o.A = psi1 - psis;
    o.B = psi2 - psis;
    o.kl = vb_beta_kl(e1, e2, q1, q2, psi1, psi2, psis);
    return o;
}

// shared theta (T = 1).  Any block size that is a multiple of 32; warp w sums slots w, w + nwarps, ... of the
// n_partials block partials, each slot in the same order whatever the block size.
__device__ __forceinline__ void theta_block(const EmP& p, int b, int theta_mode, int n_partials) {
    const bool do_theta = vb_theta_on(p, b, theta_mode);
    const int G = p.G;
    __shared__ double tot[2 * VB_MAX_GT];
    __shared__ double kls[VB_MAX_GT];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (do_theta) {
        for (int slot = w; slot < 2 * VB_MAX_GT; slot += nw) {
            const double* src = p.part + (size_t)b * p.part_stride + p.off_theta + slot;
            double t = 0.0;
            for (int blk = lane; blk < n_partials; blk += 32) t += src[(size_t)blk * 2 * VB_MAX_GT];
            t = warp_sum(t);
            if (lane == 0) tot[slot] = t;
        }
    }
    __syncthreads();
    if (threadIdx.x < G) {
        const int g = threadIdx.x;
        ThetaOut o = theta_finish(p, do_theta, tot[g], tot[VB_MAX_GT + g], p.s1p[g], p.s2p[g],
                                  p.mu + (size_t)b * G + g, p.sum + (size_t)b * G + g);
        double* ab = p.ab + (size_t)b * 2 * G;
        ab[g] = o.A;
        ab[G + g] = o.B;
        kls[g] = o.kl;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int g = 0; g < G; ++g) t += kls[g];
        p.part[(size_t)b * p.part_stride + p.off_klth] = t;
        if (p.tiled == 3) {
            // fixed-point tables: every table entry is a convex combination of the A_g (or of the B_g), so
            // |W| <= max_g max(|A_g|, |B_g|) < 2^e and -W * 2^(32-e) fits an unsigned 32-bit value
            const double* ab = p.ab + (size_t)b * 2 * G;
            double mx = 0.0;
            for (int g = 0; g < 2 * G; ++g) mx = fmax(mx, fabs(ab[g]));
            const int e = (mx > 0.0 && mx < 1e300) ? ilogb(mx) + 1 : 0;
            p.qscale[b] = ldexp(1.0, 32 - e);
        }
    }
}

// final sums + the convergence rule.  advance = 1 inside the fit loop.
// cell_terms != nullptr (cell-sharded fit): {LB_p, KL_ID} already summed over the blocks and over the ranks
__device__ __forceinline__ void elbo_block(const EmP& p, int b, int advance, const double* __restrict__ cell_terms) {
    int* ctrl = p.ctrl + b * VB_CTRL_N;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const double* part = p.part + (size_t)b * p.part_stride;
    __shared__ double etot[4];
    for (int slot = w; slot < 4; slot += nw) {
        double t = 0.0;
        if (slot == 0) { if (cell_terms) { if (lane == 0) t = cell_terms[0]; } else for (int i = lane; i < p.n_cellblk; i += 32) t += part[p.off_cell + 2 * i]; }
        else if (slot == 1) { if (cell_terms) { if (lane == 0) t = cell_terms[1]; } else for (int i = lane; i < p.n_cellblk; i += 32) t += part[p.off_cell + 2 * i + 1]; }
        else if (slot == 2) { if (!p.bmm) for (int i = lane; i < p.n_elemblk; i += 32) t += part[p.off_klgt + i]; }
        else for (int i = lane; i < p.n_klth; i += 32) t += part[p.off_klth + i];
        t = warp_sum(t);
        if (lane == 0) etot[slot] = t;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const double E = etot[0] - etot[1] - etot[2] - etot[3];      // LB_p - KL_ID - KL_GT - KL_theta (:248)
    double* sc = p.scal + (size_t)b * VB_SCAL_N;
    sc[0] = E; sc[1] = etot[0]; sc[2] = etot[1]; sc[3] = etot[2]; sc[4] = etot[3];
    if (!advance) return;
    const int it = ctrl[1];
    double* elbo = p.elbo + (size_t)b * p.max_iter;
    elbo[it] = E;
    bool brk = false;
    if (it > p.min_iter) {                                       // strict, as the reference (:266)
        const double prev = elbo[it - 1];
        const bool dec = p.bmm ? (E - prev < -1e-6)              // bmm_model.py:191
                               : (E < prev - 1e-6);              // vireo_model.py:267
        if (dec) ctrl[3] += 1;                                   // reference only warns
        else if (it == p.max_iter - 1) { /* "did not converge" warning, replayed on the host */ }
        else if (E - prev < p.eps) brk = true;                   // :273
    }
    ctrl[2] = it;                                                // the reference returns ELBO[:it]
    if (brk || it + 1 >= p.max_iter) ctrl[0] = 1;
    else ctrl[1] = it + 1;
}

// cell-sharded fit: the block partials {LB_p, KL_ID} of the local cell pass -> two doubles behind S1 | S2 in the
// exchange buffer (summed over the ranks by the next iteration's all-reduce)
__device__ __forceinline__ void xchg_pack_block(const EmP& p, double* __restrict__ out2) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int slot = w; slot < 2; slot += nw) {
        double t = 0.0;
        for (int i = lane; i < p.n_cellblk; i += 32) t += p.part[p.off_cell + 2 * i + slot];
        t = warp_sum(t);
        if (lane == 0) out2[slot] = t;
    }
}

// what a fused sparse pass does in its tail (EmP.fuse): 0 nothing (stand-alone k_theta / k_elbo follow), 1 theta_block
// after the SNP pass and elbo_block after the cell pass, 2 (cell-sharded fit) xchg_pack_block after the cell pass
__device__ __forceinline__ void snp_pass_tail(const EmP& p, int b, int theta_mode, bool has_partials) {
    if (p.fuse != 1 || p.bmm || p.ase) return;
    if (!has_partials) { if (blockIdx.x == 0) theta_block(p, b, theta_mode, 0); return; }      // theta is not updated: tables only
    if (vb_last_cta(p.ctrl + b * VB_CTRL_N + 4)) theta_block(p, b, theta_mode, (int)gridDim.x);
}
__device__ __forceinline__ void cell_pass_tail(const EmP& p, int b) {
    if (p.fuse == 0) return;
    if (!vb_last_cta(p.ctrl + b * VB_CTRL_N + 5)) return;
    if (p.fuse == 1) elbo_block(p, b, 1, nullptr);
    else xchg_pack_block(p, p.xs);
}
