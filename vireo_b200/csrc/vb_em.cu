// vb_em.cu -- the variational-EM iteration of vireoSNP as sm_100a kernels (FP64, no tensor cores:
// this is a sparse, bandwidth-bound reduction).
//
// One EM iteration of Vireo._fit_VB (vireoSNP/utils/vireo_model.py:257-264) is five launches:
//   k_snp    SNP-major pass over the nnz:  S1 = AD @ R, S2 = (DP-AD) @ R        (:168-170, :207-209)
//            + partial sums of S1*GT_old, S2*GT_old for the theta update          (:175-181)
//   k_theta  reduce those sums, new Beta posterior (beta_mu, beta_sum)            (:183-185),
//            digamma differences A_g = psi(s1)-psi(s1+s2), B_g = psi(s2)-psi(s1+s2) (:149-162), KL_theta
//   k_gt     GT_prob = softmax_g(S1*A_g + S2*B_g + log GT_prior)                  (:211-219), KL_GT,
//            and the per-(SNP, donor) tables  Wa = sum_g GT*A_g,  Wb = sum_g GT*B_g
//   k_cell   cell-major pass over the nnz: logLik_ID = sum_i ad*Wa + (dp-ad)*Wb  (:190-196; the
//            reference's 3 sparse products per genotype collapse into this one pass),
//            ID_prob = softmax_k(logLik_ID + log ID_prior)                        (:198-199),
//            partial sums of logLik_ID*ID_prob and KL(ID_prob || ID_prior)        (:236-237)
//   k_elbo   ELBO = LB_p - KL_ID - KL_GT - KL_theta (:236-248) and the convergence rule (:266-274)
// BinomMixtureVB._fit_BV (vireoSNP/utils/bmm_model.py:183-199) reuses k_snp / k_cell / k_elbo with
// k_bmm_theta in place of k_theta + k_gt.
//
// All reductions are two-stage with a fixed launch geometry, so results are run-to-run identical.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "vb_common.cuh"
#include "vb_tail.cuh"

// ---------------------------------------------------------------------------------------------
// launch accounting: every kernel launch of the EM path goes through VB_LAUNCH, which counts it and,
// when profiling is enabled, brackets it with CUDA events on the launching stream.
// classes: 0 k_snp, 1 k_theta, 2 k_gt, 3 k_cell, 4 k_elbo, 5 k_bmm_theta, 6 k_terms, 7 doublet helpers
// ---------------------------------------------------------------------------------------------
struct ProfEvent { int cls; cudaEvent_t e0, e1; };
static bool g_prof_on = false;
static std::vector<ProfEvent> g_prof_events;
static int64_t g_launches[8] = {0, 0, 0, 0, 0, 0, 0, 0};

void vb_launch_begin(int cls, cudaStream_t st, cudaEvent_t* e0, cudaEvent_t* e1) {
    (void)cls;
    if (g_prof_on) {
        cudaEventCreate(e0);
        cudaEventCreate(e1);
        cudaEventRecord(*e0, st);
    }
}

void vb_launch_end(int cls, cudaStream_t st, cudaEvent_t e0, cudaEvent_t e1) {
    if (g_prof_on && e0 && e1) {
        cudaEventRecord(e1, st);
        g_prof_events.push_back(ProfEvent{cls, e0, e1});
    }
    g_launches[cls] += 1;
}

extern "C" void vb_profile_enable(int on) { g_prof_on = on != 0; }

extern "C" int vb_profile_read(double* ms8, int64_t* n8) {
    for (int i = 0; i < 8; ++i) { if (ms8) ms8[i] = 0.0; if (n8) n8[i] = 0; }
    for (auto& ev : g_prof_events) {
        float ms = 0.f;
        cudaEventSynchronize(ev.e1);
        if (cudaEventElapsedTime(&ms, ev.e0, ev.e1) == cudaSuccess) {
            if (ms8) ms8[ev.cls] += ms;
            if (n8) n8[ev.cls] += 1;
        }
        cudaEventDestroy(ev.e0);
        cudaEventDestroy(ev.e1);
    }
    g_prof_events.clear();
    return VB_OK;
}

extern "C" void vb_launch_counts(int64_t* n8) { for (int i = 0; i < 8; ++i) n8[i] = g_launches[i]; }

// ---------------------------------------------------------------------------------------------
// count decoding
// ---------------------------------------------------------------------------------------------
template <bool WIDE>
__device__ __forceinline__ void decode(uint32_t c, uint32_t d, int& a, int& b) {
    if (WIDE) { a = (int)c; b = (int)d - a; }
    else { a = (int)(c & 0xffffu); b = (int)(c >> 16) - a; }
}

__device__ __forceinline__ double axpy_count(int n, double w, double acc) {
    // unit counts dominate real data (about 90% of DP entries are 1): skip the int->double conversion
    return n == 1 ? acc + w : fma((double)n, w, acc);
}

// ---------------------------------------------------------------------------------------------
// k_cell: one warp per cell.  KT lanes span the donor axis (KR registers each when K > 32),
// so 32/KT nnz are in flight per step; the two tables are gathered by SNP id (L2-resident).
// mode 0: fused softmax, writes ID_prob + logLik_ID + ELBO partials
// mode 1: logLik_ID only, ELBO partials from the existing ID_prob
// FUSED = false: columns [k_off, k_off + KT*KR) of a wider table, logLik only (doublet pass)
// ---------------------------------------------------------------------------------------------
template <int KT, int KR, bool WIDE, bool FUSED>
__global__ void __launch_bounds__(VB_THREADS)
k_cell(const CountsView m, const EmP p, const int mode, const int k_off) {
    const int b = blockIdx.y;
    if (p.ctrl && p.ctrl[b * VB_CTRL_N]) return;
    constexpr int NPW = 32 / KT;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int sub = lane / KT, kl = lane % KT;
    const int K = p.K;
    // Wt[V][2][K]: row 2i = table of a reference-allele read at SNP i (Wb), row 2i+1 = alternative allele (Wa)
    const double* __restrict__ Wt = p.Wt + (size_t)b * p.V * 2 * K;
    double* __restrict__ R = p.R ? p.R + (size_t)b * p.C * K : nullptr;
    double* __restrict__ LL = p.ll + (size_t)b * p.C * K;
    double lbp = 0.0, klid = 0.0;

    const int64_t nw = (int64_t)gridDim.x * VB_WARPS;
    for (int64_t j = (int64_t)blockIdx.x * VB_WARPS + wib; j < p.C; j += nw) {
        double acc[KR];
#pragma unroll
        for (int r = 0; r < KR; ++r) acc[r] = 0.0;
        const int64_t p0 = m.cell_ptr[j], p1 = m.cell_ptr[j + 1];
        // the records of the next 32 nnz are requested before the current 32 are processed (one load latency per
        // batch off the dependent chain index -> table row -> FMA)
        int idx_n = 0;
        uint32_t c_n = 0, d_n = 0;
        if (p0 + lane < p1) {
            idx_n = __ldg(m.cell_idx + p0 + lane);
            c_n = __ldg(m.cell_cnt + p0 + lane);
            if (WIDE) d_n = __ldg(m.cell_dp + p0 + lane);
        }
        for (int64_t base = p0; base < p1; base += 32) {
            const int idx = idx_n;
            const uint32_t c = c_n, d = d_n;
            const int64_t qn = base + 32 + lane;
            idx_n = 0; c_n = 0; d_n = 0;
            if (qn < p1) {
                idx_n = __ldg(m.cell_idx + qn);
                c_n = __ldg(m.cell_cnt + qn);
                if (WIDE) d_n = __ldg(m.cell_dp + qn);
            }
            const int n = (int)((p1 - base) < 32 ? (p1 - base) : 32);
#pragma unroll 4
            for (int s = 0; s < KT; ++s) {
                if (s * NPW >= n) break;                       // warp-uniform
                const int src = s * NPW + sub;
                const int i = __shfl_sync(VB_FULL, idx, src);
                const uint32_t cc = __shfl_sync(VB_FULL, c, src);
                const uint32_t dd = WIDE ? __shfl_sync(VB_FULL, d, src) : 0u;
                if (src < n) {
                    int av, bv;
                    decode<WIDE>(cc, dd, av, bv);
                    const size_t row = (size_t)i * 2 * K + k_off + kl;
#pragma unroll
                    for (int r = 0; r < KR; ++r) {
                        if (k_off + kl + r * KT < K) {
                            if (bv) acc[r] = axpy_count(bv, __ldg(Wt + row + r * KT), acc[r]);
                            if (av) acc[r] = axpy_count(av, __ldg(Wt + row + K + r * KT), acc[r]);
                        }
                    }
                }
            }
        }
        // fold the NPW partial sums of each column (xor butterfly: every lane ends with the same bits)
#pragma unroll
        for (int off = KT; off < 32; off <<= 1)
#pragma unroll
            for (int r = 0; r < KR; ++r) acc[r] += __shfl_xor_sync(VB_FULL, acc[r], off);

        if (!FUSED) {
            if (sub == 0)
#pragma unroll
                for (int r = 0; r < KR; ++r) {
                    const int k = k_off + kl + r * KT;
                    if (k < K) LL[(size_t)j * K + k] = acc[r];
                }
            continue;
        }

        const size_t prow = (size_t)(p.id_rows == 1 ? 0 : j) * K;
        double pr[KR];
        if (mode == 0) {
            double lg[KR], mx = -INFINITY;
#pragma unroll
            for (int r = 0; r < KR; ++r) {
                const int k = kl + r * KT;
                lg[r] = k < K ? acc[r] + p.lidp[prow + k] : -INFINITY;
                mx = fmax(mx, lg[r]);
            }
#pragma unroll
            for (int off = KT / 2; off > 0; off >>= 1) mx = fmax(mx, __shfl_xor_sync(VB_FULL, mx, off));
            double z = 0.0;
#pragma unroll
            for (int r = 0; r < KR; ++r) {
                pr[r] = (kl + r * KT < K) ? exp(lg[r] - mx) : 0.0;
                z += pr[r];
            }
#pragma unroll
            for (int off = KT / 2; off > 0; off >>= 1) z += __shfl_xor_sync(VB_FULL, z, off);
#pragma unroll
            for (int r = 0; r < KR; ++r) pr[r] = pr[r] / z;
        } else {
#pragma unroll
            for (int r = 0; r < KR; ++r) {
                const int k = kl + r * KT;
                pr[r] = k < K ? R[(size_t)j * K + k] : 0.0;
            }
        }
        if (sub == 0) {
#pragma unroll
            for (int r = 0; r < KR; ++r) {
                const int k = kl + r * KT;
                if (k < K) {
                    const size_t e = (size_t)j * K + k;
                    LL[e] = acc[r];
                    if (mode == 0) R[e] = pr[r];
                    lbp += acc[r] * pr[r];
                    if (pr[r] > 0.0) klid += pr[r] * (log(pr[r]) - p.lidp_kl[prow + k]);
                }
            }
        }
    }
    if (!FUSED) return;
    __shared__ double sh[VB_WARPS];
    const double t0 = block_sum(lbp, sh);
    const double t1 = block_sum(klid, sh);
    if (threadIdx.x == 0) {
        double* out = p.part + (size_t)b * p.part_stride + p.off_cell + 2 * blockIdx.x;
        out[0] = t0;
        out[1] = t1;
    }
    cell_pass_tail(p, b);
}

// ---------------------------------------------------------------------------------------------
// k_snp: one warp per SNP.  Gathers ID_prob rows by cell id; two accumulators per column.
// theta_mode: 0 never, 1 always, 2 when learn_theta and the device iteration counter >= delay
// ---------------------------------------------------------------------------------------------
template <int KT, int KR, bool WIDE>
__global__ void __launch_bounds__(VB_THREADS)
k_snp(const CountsView m, const EmP p, const int theta_mode) {
    const int b = blockIdx.y;
    if (p.ctrl && p.ctrl[b * VB_CTRL_N]) return;
    const bool do_theta = !p.bmm && vb_theta_on(p, b, theta_mode);
    if (!p.bmm && !do_theta && !p.learn_gt && theta_mode == 2) {         // nothing consumes S1/S2 this iteration
        snp_pass_tail(p, b, theta_mode, false);
        return;
    }
    constexpr int NPW = 32 / KT;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int sub = lane / KT, kl = lane % KT;
    const int K = p.K, G = p.G;
    const double* __restrict__ R = p.R + (size_t)b * p.C * K;
    const double* __restrict__ GT = p.GT ? p.GT + (size_t)b * p.V * K * G : nullptr;
    double* __restrict__ S1 = p.S1 + (size_t)b * p.V * K;
    double* __restrict__ S2 = p.S2 + (size_t)b * p.V * K;
    double t1[VB_MAX_GT], t2[VB_MAX_GT];
#pragma unroll
    for (int g = 0; g < VB_MAX_GT; ++g) t1[g] = t2[g] = 0.0;

    const int64_t nw = (int64_t)gridDim.x * VB_WARPS;
    for (int64_t i = (int64_t)blockIdx.x * VB_WARPS + wib; i < p.V; i += nw) {
        double a1[KR], a2[KR];
#pragma unroll
        for (int r = 0; r < KR; ++r) a1[r] = a2[r] = 0.0;
        const int64_t p0 = m.snp_ptr[i], p1 = m.snp_ptr[i + 1];
        int idx_n = 0;                 // next batch of records, requested one batch ahead (see k_cell)
        uint32_t c_n = 0, d_n = 0;
        if (p0 + lane < p1) {
            idx_n = __ldg(m.snp_idx + p0 + lane);
            c_n = __ldg(m.snp_cnt + p0 + lane);
            if (WIDE) d_n = __ldg(m.snp_dp + p0 + lane);
        }
        for (int64_t base = p0; base < p1; base += 32) {
            const int idx = idx_n;
            const uint32_t c = c_n, d = d_n;
            const int64_t qn = base + 32 + lane;
            idx_n = 0; c_n = 0; d_n = 0;
            if (qn < p1) {
                idx_n = __ldg(m.snp_idx + qn);
                c_n = __ldg(m.snp_cnt + qn);
                if (WIDE) d_n = __ldg(m.snp_dp + qn);
            }
            const int n = (int)((p1 - base) < 32 ? (p1 - base) : 32);
#pragma unroll 4
            for (int s = 0; s < KT; ++s) {
                if (s * NPW >= n) break;
                const int src = s * NPW + sub;
                const int j = __shfl_sync(VB_FULL, idx, src);
                const uint32_t cc = __shfl_sync(VB_FULL, c, src);
                const uint32_t dd = WIDE ? __shfl_sync(VB_FULL, d, src) : 0u;
                if (src < n) {
                    int av, bv;
                    decode<WIDE>(cc, dd, av, bv);
                    const size_t row = (size_t)j * K + kl;
#pragma unroll
                    for (int r = 0; r < KR; ++r) {
                        if (kl + r * KT < K) {
                            const double w = __ldg(R + row + r * KT);
                            if (av) a1[r] = axpy_count(av, w, a1[r]);
                            if (bv) a2[r] = axpy_count(bv, w, a2[r]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int off = KT; off < 32; off <<= 1)
#pragma unroll
            for (int r = 0; r < KR; ++r) {
                a1[r] += __shfl_xor_sync(VB_FULL, a1[r], off);
                a2[r] += __shfl_xor_sync(VB_FULL, a2[r], off);
            }
        if (sub == 0) {
#pragma unroll
            for (int r = 0; r < KR; ++r) {
                const int k = kl + r * KT;
                if (k < K) {
                    const size_t e = (size_t)i * K + k;
                    S1[e] = a1[r];
                    S2[e] = a2[r];
                    if (do_theta) {
#pragma unroll
                        for (int g = 0; g < VB_MAX_GT; ++g)
                            if (g < G) {
                                const double gt = GT[e * G + g];
                                t1[g] += a1[r] * gt;
                                t2[g] += a2[r] * gt;
                            }
                    }
                }
            }
        }
        if (do_theta && p.ase) {
            // allele-specific mode: theta is per SNP (vireo_model.py:177 `axis=1`): finish this row now.
            // The raw sums are parked in the `ab` rows that k_theta_ase overwrites afterwards.
            double* row = p.ab + ((size_t)b * p.T + i) * 2 * G;
#pragma unroll
            for (int g = 0; g < VB_MAX_GT; ++g)
                if (g < G) {
                    const double u1 = warp_sum(t1[g]), u2 = warp_sum(t2[g]);
                    if (lane == 0) { row[g] = u1; row[G + g] = u2; }
                    t1[g] = t2[g] = 0.0;
                }
        }
    }
    if (!do_theta || p.ase) { snp_pass_tail(p, b, theta_mode, false); return; }
    __shared__ double sh[VB_WARPS][2 * VB_MAX_GT];
#pragma unroll
    for (int g = 0; g < VB_MAX_GT; ++g) {
        if (g < G) {                                                  // G is grid-uniform
            const double u1 = warp_sum(t1[g]), u2 = warp_sum(t2[g]);
            if (lane == 0) { sh[wib][g] = u1; sh[wib][VB_MAX_GT + g] = u2; }
        }
    }
    __syncthreads();
    if (threadIdx.x < 2 * VB_MAX_GT) {
        const int g = threadIdx.x % VB_MAX_GT;
        double t = 0.0;
        if (g < G)
            for (int w = 0; w < VB_WARPS; ++w) t += sh[w][threadIdx.x];
        p.part[(size_t)b * p.part_stride + p.off_theta + (size_t)blockIdx.x * 2 * VB_MAX_GT + threadIdx.x] = t;
    }    snp_pass_tail(p, b, theta_mode, true);
}

// ---------------------------------------------------------------------------------------------
// Row kernels for few donors (K <= 8), one warp per row, one LANE per nnz: every lane gathers the whole K-wide table
// row of its own nnz and keeps K private accumulators.  Small matrices are served by these: their passes are
// instruction- and latency-bound (measured at 10k x 5k x 4 on the KT-lanes-per-nnz kernels: 11.5 warp instructions
// per nnz, issue slots 50% busy, the rest dependent L2 latencies), so
//   * a lane handles U records per round: the U record loads are issued together, then the U table-row gathers
//     (one 16-byte load per two columns), then the FMAs -- two L2 latencies per round instead of two per record;
//   * a record gathers ONE row in the cell pass (most reads are pure reference or pure alternative; the second
//     allele of a heterozygous read takes a rare extra load);
//   * counts become doubles by the 2^52 trick (no conversion pipe), FMAs are unconditional (0 * finite);
//   * the K accumulators are folded across the warp by a reduce-scatter butterfly: every exchange halves the columns
//     a lane carries, K - 1 + log2(32 / K) shuffles instead of 5 K; lanes [c * 32 / KP, (c + 1) * 32 / KP) end up
//     holding the total of column c (all with the same bits), and the softmax runs across those lane groups.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double cnt2dbl(uint32_t c) { return __hiloint2double(0x43300000, (int)c) - 4503599627370496.0; }

template <int KP>
__device__ __forceinline__ double fold_scatter(double (&a)[KP], int lane) {
    // after the step with distance `off` a lane keeps the half of its columns selected by that bit of its id
    int n = KP;
#pragma unroll
    for (int off = 16; off >= 32 / KP; off >>= 1) {
        n >>= 1;
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int k = 0; k < KP / 2; ++k) {
            if (k < n) {
                const double send = up ? a[k] : a[k + n];
                const double keep = up ? a[k + n] : a[k];
                a[k] = keep + __shfl_xor_sync(VB_FULL, send, off);
            }
        }
    }
    double t = a[0];
#pragma unroll
    for (int off = 32 / KP / 2; off > 0; off >>= 1) t += __shfl_xor_sync(VB_FULL, t, off);
    return t;      // total of column lane / (32 / KP)
}


// the K (<= KP) doubles of a table row; 16-byte loads when the row is exactly KP wide (rows are then 16-byte aligned)
template <int KP>
__device__ __forceinline__ void load_row(const double* __restrict__ row, int K, bool on, double (&w)[KP]) {
    if (K == KP) {
        const double2* __restrict__ r2 = reinterpret_cast<const double2*>(row);
#pragma unroll
        for (int k = 0; k < KP / 2; ++k) {
            double2 v = make_double2(0.0, 0.0);
            if (on) v = __ldg(r2 + k);
            w[2 * k] = v.x; w[2 * k + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int k = 0; k < KP; ++k) w[k] = (on && k < K) ? __ldg(row + k) : 0.0;
    }
}

#ifndef VB_LANE_MINB
#define VB_LANE_MINB 4          // resident CTAs per SM the shallow lane kernels are compiled for (64 registers; 5 or 6
                                // -- 48 / 40 registers -- spill and were slower at cfg2 and cfg5)
#endif
// U = records per lane and round (more in flight per lane against fewer warps per SM: more registers)
template <int KP, bool WIDE, int U>
__global__ void __launch_bounds__(VB_THREADS, U * KP <= 8 ? VB_LANE_MINB : 2)
k_cell_lane(const CountsView m, const EmP p, const int mode) {
    constexpr int GW = 32 / KP;                            // lanes per column after the fold
    const int b = blockIdx.y;
    if (p.ctrl && p.ctrl[b * VB_CTRL_N]) return;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int K = p.K;
    const int col = lane / GW;                             // the column this lane owns after the fold
    const double* __restrict__ Wt = p.Wt + (size_t)b * p.V * 2 * K;
    double* __restrict__ R = p.R ? p.R + (size_t)b * p.C * K : nullptr;
    double* __restrict__ LL = p.ll + (size_t)b * p.C * K;
    double lbp = 0.0, klid = 0.0;
    const int64_t nw = (int64_t)gridDim.x * VB_WARPS;
    for (int64_t j = (int64_t)blockIdx.x * VB_WARPS + wib; j < p.C; j += nw) {
        double acc[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) acc[k] = 0.0;
        const int64_t p1 = m.cell_ptr[j + 1];
        for (int64_t q0 = m.cell_ptr[j] + lane; q0 < p1; q0 += 32 * U) {
            int av[U], bv[U], ii[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t q = q0 + 32 * u;
                const bool ok = q < p1;
                const uint32_t c = ok ? __ldg(m.cell_cnt + q) : 0u;
                const uint32_t d = (WIDE && ok) ? __ldg(m.cell_dp + q) : 0u;
                ii[u] = ok ? __ldg(m.cell_idx + q) : 0;
                decode<WIDE>(c, d, av[u], bv[u]);
            }
            double w[U][KP];
#pragma unroll
            for (int u = 0; u < U; ++u)       // alternative-allele row when the read has alternative counts, else reference
                load_row<KP>(Wt + (size_t)ii[u] * 2 * K + (av[u] ? K : 0), K, (av[u] | bv[u]) != 0, w[u]);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double c1 = cnt2dbl((uint32_t)(av[u] ? av[u] : bv[u]));
#pragma unroll
                for (int k = 0; k < KP; ++k) acc[k] = fma(c1, w[u][k], acc[k]);
                if (av[u] && bv[u]) {         // both alleles seen: add the reference row as well
                    double w2[KP];
                    load_row<KP>(Wt + (size_t)ii[u] * 2 * K, K, true, w2);
                    const double c2 = cnt2dbl((uint32_t)bv[u]);
#pragma unroll
                    for (int k = 0; k < KP; ++k) acc[k] = fma(c2, w2[k], acc[k]);
                }
            }
        }
        const double ll = fold_scatter<KP>(acc, lane);     // logLik_ID[j, col]
        const size_t prow = (size_t)(p.id_rows == 1 ? 0 : j) * K;
        const bool live = col < K;
        double pr;
        if (mode == 0) {
            const double lg = live ? ll + p.lidp[prow + col] : -INFINITY;
            double mx = lg;
#pragma unroll
            for (int off = GW; off < 32; off <<= 1) mx = fmax(mx, __shfl_xor_sync(VB_FULL, mx, off));
            const double e = live ? exp(lg - mx) : 0.0;
            double z = e;
#pragma unroll
            for (int off = GW; off < 32; off <<= 1) z += __shfl_xor_sync(VB_FULL, z, off);
            pr = e / z;
        } else {
            pr = live ? R[(size_t)j * K + col] : 0.0;
        }
        if (live && (lane % GW) == 0) {
            const size_t e = (size_t)j * K + col;
            LL[e] = ll;
            if (mode == 0) R[e] = pr;
            lbp += ll * pr;
            if (pr > 0.0) klid += pr * (log(pr) - p.lidp_kl[prow + col]);
        }
    }
    __shared__ double sh[VB_WARPS];
    const double t0 = block_sum(lbp, sh);
    const double t1 = block_sum(klid, sh);
    if (threadIdx.x == 0) {
        double* out = p.part + (size_t)b * p.part_stride + p.off_cell + 2 * blockIdx.x;
        out[0] = t0;
        out[1] = t1;
    }
    cell_pass_tail(p, b);
}

template <int KP, bool WIDE, int U>
__global__ void __launch_bounds__(VB_THREADS, U * KP <= 8 ? VB_LANE_MINB : 2)
k_snp_lane(const CountsView m, const EmP p, const int theta_mode) {
    constexpr int GW = 32 / KP;
    const int b = blockIdx.y;
    if (p.ctrl && p.ctrl[b * VB_CTRL_N]) return;
    const bool do_theta = !p.bmm && vb_theta_on(p, b, theta_mode);
    if (!p.bmm && !do_theta && !p.learn_gt && theta_mode == 2) {         // nothing consumes S1/S2 this iteration
        snp_pass_tail(p, b, theta_mode, false);
        return;
    }
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int K = p.K, G = p.G;
    const int col = lane / GW;
    const double* __restrict__ R = p.R + (size_t)b * p.C * K;
    const double* __restrict__ GT = p.GT ? p.GT + (size_t)b * p.V * K * G : nullptr;
    double* __restrict__ S1 = p.S1 + (size_t)b * p.V * K;
    double* __restrict__ S2 = p.S2 + (size_t)b * p.V * K;
    double t1[VB_MAX_GT], t2[VB_MAX_GT];
#pragma unroll
    for (int g = 0; g < VB_MAX_GT; ++g) t1[g] = t2[g] = 0.0;
    const int64_t nw = (int64_t)gridDim.x * VB_WARPS;
    for (int64_t i = (int64_t)blockIdx.x * VB_WARPS + wib; i < p.V; i += nw) {
        double a1[KP], a2[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) a1[k] = a2[k] = 0.0;
        const int64_t p1 = m.snp_ptr[i + 1];
        for (int64_t q0 = m.snp_ptr[i] + lane; q0 < p1; q0 += 32 * U) {
            int av[U], bv[U], jj[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t q = q0 + 32 * u;
                const bool ok = q < p1;
                const uint32_t c = ok ? __ldg(m.snp_cnt + q) : 0u;
                const uint32_t d = (WIDE && ok) ? __ldg(m.snp_dp + q) : 0u;
                jj[u] = ok ? __ldg(m.snp_idx + q) : 0;
                decode<WIDE>(c, d, av[u], bv[u]);
            }
            double w[U][KP];
#pragma unroll
            for (int u = 0; u < U; ++u) load_row<KP>(R + (size_t)jj[u] * K, K, (av[u] | bv[u]) != 0, w[u]);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double ca = cnt2dbl((uint32_t)av[u]), cb = cnt2dbl((uint32_t)bv[u]);
#pragma unroll
                for (int k = 0; k < KP; ++k) { a1[k] = fma(ca, w[u][k], a1[k]); a2[k] = fma(cb, w[u][k], a2[k]); }
            }
        }
        const double s1 = fold_scatter<KP>(a1, lane), s2 = fold_scatter<KP>(a2, lane);
        if (col < K && (lane % GW) == 0) {
            const size_t e = (size_t)i * K + col;
            S1[e] = s1;
            S2[e] = s2;
            if (do_theta) {
#pragma unroll
                for (int g = 0; g < VB_MAX_GT; ++g)
                    if (g < G) {
                        const double gt = GT[e * G + g];
                        t1[g] += s1 * gt;
                        t2[g] += s2 * gt;
                    }
            }
        }
        if (do_theta && p.ase) {
            // allele-specific mode: theta per SNP (vireo_model.py:177): the raw sums are parked in the ab rows
            double* row = p.ab + ((size_t)b * p.T + i) * 2 * G;
#pragma unroll
            for (int g = 0; g < VB_MAX_GT; ++g)
                if (g < G) {
                    const double u1 = warp_sum(t1[g]), u2 = warp_sum(t2[g]);
                    if (lane == 0) { row[g] = u1; row[G + g] = u2; }
                    t1[g] = t2[g] = 0.0;
                }
        }
    }
    if (!do_theta || p.ase) { snp_pass_tail(p, b, theta_mode, false); return; }
    __shared__ double sh[VB_WARPS][2 * VB_MAX_GT];
#pragma unroll
    for (int g = 0; g < VB_MAX_GT; ++g) {
        if (g < G) {
            const double u1 = warp_sum(t1[g]), u2 = warp_sum(t2[g]);
            if (lane == 0) { sh[wib][g] = u1; sh[wib][VB_MAX_GT + g] = u2; }
        }
    }
    __syncthreads();
    if (threadIdx.x < 2 * VB_MAX_GT) {
        const int g = threadIdx.x % VB_MAX_GT;
        double t = 0.0;
        if (g < G)
            for (int w = 0; w < VB_WARPS; ++w) t += sh[w][threadIdx.x];
        p.part[(size_t)b * p.part_stride + p.off_theta + (size_t)blockIdx.x * 2 * VB_MAX_GT + threadIdx.x] = t;
    }    snp_pass_tail(p, b, theta_mode, true);
}

// ---------------------------------------------------------------------------------------------
// k_theta (shared theta, T = 1): one CTA per restart
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(2 * VB_MAX_GT * 32) k_theta(const EmP p, const int theta_mode) {
    const int b = blockIdx.y;
    if (p.ctrl && p.ctrl[b * VB_CTRL_N]) return;
    theta_block(p, b, theta_mode, p.n_snpblk);
}

// allele-specific mode: theta per SNP; raw sums arrive in the ab rows (see k_snp)
__global__ void __launch_bounds__(VB_THREADS) k_theta_ase(const EmP p, const int theta_mode) {
    const int b = blockIdx.y;
    if (p.ctrl && p.ctrl[b * VB_CTRL_N]) return;
    const bool do_theta = vb_theta_on(p, b, theta_mode);
    const int G = p.G;
    __shared__ double sh[VB_WARPS];
    double kl = 0.0;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < p.T; t += (int64_t)gridDim.x * blockDim.x) {
        double* ab = p.ab + ((size_t)b * p.T + t) * 2 * G;
        double r1[VB_MAX_GT], r2[VB_MAX_GT];
#pragma unroll
        for (int g = 0; g < VB_MAX_GT; ++g)
            if (g < G) { r1[g] = do_theta ? ab[g] : 0.0; r2[g] = do_theta ? ab[G + g] : 0.0; }
#pragma unroll
        for (int g = 0; g < VB_MAX_GT; ++g)
            if (g < G) {
                const size_t pi = (size_t)(p.thp_rows == 1 ? 0 : t) * G + g;
                ThetaOut o = theta_finish(p, do_theta, r1[g], r2[g], p.s1p[pi], p.s2p[pi],
                                          p.mu + ((size_t)b * p.T + t) * G + g, p.sum + ((size_t)b * p.T + t) * G + g);
                ab[g] = o.A;
                ab[G + g] = o.B;
                kl += o.kl;
            }
    }
    const double t = block_sum(kl, sh);
    if (threadIdx.x == 0) p.part[(size_t)b * p.part_stride + p.off_klth + blockIdx.x] = t;
}

// ---------------------------------------------------------------------------------------------
// k_theta_sums: the partial sums sum_{i,k} S1*GT_old, S2*GT_old of the theta update (vireo_model.py:175-181)
// recomputed from S1/S2 -- for callers that reduce S1/S2 across devices after the SNP pass (cell-sharded fit) and
// therefore cannot use the sums the SNP pass folds into its epilogue.  Same output layout as k_snp.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(VB_THREADS) k_theta_sums(const EmP p, const int tail, const int first) {
    const int b = blockIdx.y;
    if (p.ctrl && p.ctrl[b * VB_CTRL_N]) return;
    const int K = p.K, G = p.G;
    const int64_t VK = p.V * K;
    const double* __restrict__ GT = p.GT + (size_t)b * VK * G;
    const double* __restrict__ S1 = p.S1 + (size_t)b * VK;
    const double* __restrict__ S2 = p.S2 + (size_t)b * VK;
    if (p.ase) {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.V; i += (int64_t)gridDim.x * blockDim.x) {
            double* row = p.ab + ((size_t)b * p.T + i) * 2 * G;
            for (int g = 0; g < G; ++g) {
                double u1 = 0.0, u2 = 0.0;
                for (int k = 0; k < K; ++k) {
                    const double gt = GT[((size_t)i * K + k) * G + g];
                    u1 += S1[i * K + k] * gt;
                    u2 += S2[i * K + k] * gt;
                }
                row[g] = u1;
                row[G + g] = u2;
            }
        }
        return;
    }
    double t1[VB_MAX_GT], t2[VB_MAX_GT];
#pragma unroll
    for (int g = 0; g < VB_MAX_GT; ++g) t1[g] = t2[g] = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < VK; e += (int64_t)gridDim.x * blockDim.x) {
        const double s1 = S1[e], s2 = S2[e];
#pragma unroll
        for (int g = 0; g < VB_MAX_GT; ++g)
            if (g < G) {
                const double gt = GT[(size_t)e * G + g];
                t1[g] += s1 * gt;
                t2[g] += s2 * gt;
            }
    }
    __shared__ double sh[VB_WARPS][2 * VB_MAX_GT];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
#pragma unroll
    for (int g = 0; g < VB_MAX_GT; ++g) {
        if (g < G) {
            const double u1 = warp_sum(t1[g]), u2 = warp_sum(t2[g]);
            if (lane == 0) { sh[wib][g] = u1; sh[wib][VB_MAX_GT + g] = u2; }
        }
    }
    __syncthreads();
    if (threadIdx.x < 2 * VB_MAX_GT) {
        const int g = threadIdx.x % VB_MAX_GT;
        double t = 0.0;
        if (g < G)
            for (int w = 0; w < VB_WARPS; ++w) t += sh[w][threadIdx.x];
        p.part[(size_t)b * p.part_stride + p.off_theta + (size_t)blockIdx.x * 2 * VB_MAX_GT + threadIdx.x] = t;
    }
    // cell-sharded fit: the CTA that finishes last closes the previous iteration (ELBO + convergence rule from the
    // exchanged cell terms) and, unless that ended the fit, finishes theta -- instead of two launches of their own
    if (tail && vb_last_cta(p.ctrl + b * VB_CTRL_N + 6)) {
        if (!first) {
            elbo_block(p, b, 1, p.xs);
            __syncthreads();
            if (p.ctrl[b * VB_CTRL_N]) return;
        }
        theta_block(p, b, 2, (int)gridDim.x);
    }
}

// ---------------------------------------------------------------------------------------------
// k_gt: one thread per (SNP, donor)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(VB_THREADS) k_gt(const EmP p, const int do_gt) {
    const int b = blockIdx.y;
    if (p.ctrl && p.ctrl[b * VB_CTRL_N]) return;
    const int K = p.K, G = p.G;
    const int64_t VK = p.V * K;
    double* __restrict__ GT = p.GT + (size_t)b * VK * G;
    const double* __restrict__ S1 = p.S1 + (size_t)b * VK;
    const double* __restrict__ S2 = p.S2 + (size_t)b * VK;
    double* __restrict__ Wt = p.Wt + (size_t)b * VK * 2;
    __shared__ double sh[VB_WARPS];
    double kl = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < VK; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / K;
        const int k = (int)(e - i * K);
        const double* ab = p.ab + ((size_t)b * p.T + (p.ase ? i : 0)) * 2 * G;
        double pr[VB_MAX_GT];
        if (do_gt) {
            const double s1 = S1[e], s2 = S2[e];
            double mx = -INFINITY;
#pragma unroll
            for (int g = 0; g < VB_MAX_GT; ++g)
                if (g < G) {
                    pr[g] = s1 * ab[g] + s2 * ab[G + g] + p.lgtp[(size_t)e * G + g];
                    mx = fmax(mx, pr[g]);
                }
            double z = 0.0;
#pragma unroll
            for (int g = 0; g < VB_MAX_GT; ++g)
                if (g < G) { pr[g] = exp(pr[g] - mx); z += pr[g]; }
#pragma unroll
            for (int g = 0; g < VB_MAX_GT; ++g)
                if (g < G) { pr[g] = pr[g] / z; GT[(size_t)e * G + g] = pr[g]; }
        } else {
#pragma unroll
            for (int g = 0; g < VB_MAX_GT; ++g)
                if (g < G) pr[g] = GT[(size_t)e * G + g];
        }
        double wa = 0.0, wb = 0.0;
#pragma unroll
        for (int g = 0; g < VB_MAX_GT; ++g)
            if (g < G) {
                wa += pr[g] * ab[g];
                wb += pr[g] * ab[G + g];
                if (pr[g] > 0.0) kl += pr[g] * (log(pr[g]) - p.lgtp_kl[(size_t)e * G + g]);
            }
        if (p.tiled) {
            // gather-table rows of RW doubles, columns replicated RW/KT times (vb_seg.cu)
            double* w0 = p.Wt + ((size_t)b * p.V + i) * 2 * p.RW;
            for (int c = k; c < p.RW; c += p.KT) { w0[c] = wb; w0[p.RW + c] = wa; }
            if (p.tiled == 3) {
                const double qs = p.qscale[b];
                const double fb = -wb * qs, fa = -wa * qs;
                const uint32_t qb = fb >= 4294967295.0 ? 0xffffffffu : (fb > 0.0 ? (uint32_t)__double2ull_rn(fb) : 0u);
                const uint32_t qa = fa >= 4294967295.0 ? 0xffffffffu : (fa > 0.0 ? (uint32_t)__double2ull_rn(fa) : 0u);
                uint32_t* q0 = p.Wq + ((size_t)b * p.V + i) * 2 * VB_ROW_DOUBLES;
                for (int c = k; c < VB_ROW_DOUBLES; c += p.KT) { q0[c] = qb; q0[VB_ROW_DOUBLES + c] = qa; }
            }
        } else {
            Wt[(size_t)i * 2 * K + k] = wb;
            Wt[(size_t)i * 2 * K + K + k] = wa;
        }
    }
    const double t = block_sum(kl, sh);
    if (threadIdx.x == 0) p.part[(size_t)b * p.part_stride + p.off_klgt + blockIdx.x] = t;
}

// ---------------------------------------------------------------------------------------------
// k_bmm_theta: BinomMixtureVB.update_theta_size (bmm_model.py:133-144) + the digamma tables of
// get_E_logLik (:125-129) + KL_theta (:166-173); one thread per (variant, clone)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(VB_THREADS) k_bmm_theta(const EmP p, const int do_theta) {
    const int b = blockIdx.y;
    if (p.ctrl && p.ctrl[b * VB_CTRL_N]) return;
    const int64_t VK = p.V * p.K;
    __shared__ double sh[VB_WARPS];
    double kl = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < VK; e += (int64_t)gridDim.x * blockDim.x) {
        const size_t o = (size_t)b * VK + e;
        ThetaOut t = theta_finish(p, do_theta != 0, do_theta ? p.S1[o] : 0.0, do_theta ? p.S2[o] : 0.0, p.s1p[e],
                                  p.s2p[e], p.mu + o, p.sum + o);
        const int64_t i = e / p.K;
        const int k = (int)(e - i * p.K);
        if (p.tiled) {
            double* w0 = p.Wt + ((size_t)b * p.V + i) * 2 * p.RW;
            for (int c = k; c < p.RW; c += p.KT) { w0[c] = t.B; w0[p.RW + c] = t.A; }
        } else {
            double* wt = p.Wt + (size_t)b * VK * 2 + (size_t)i * 2 * p.K + k;
            wt[0] = t.B;
            wt[p.K] = t.A;
        }
        kl += t.kl;
    }
    const double t = block_sum(kl, sh);
    if (threadIdx.x == 0) p.part[(size_t)b * p.part_stride + p.off_klth + blockIdx.x] = t;
}

// ---------------------------------------------------------------------------------------------
// k_terms: LB_p and KL_ID from a caller-supplied logLik_ID buffer (get_ELBO(logLik_ID), :236-237)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(VB_THREADS) k_terms(const EmP p) {
    const int b = blockIdx.y;
    const int K = p.K;
    const int64_t CK = p.C * K;
    const double* __restrict__ R = p.R + (size_t)b * CK;
    const double* __restrict__ LL = p.ll + (size_t)b * CK;
    __shared__ double sh[VB_WARPS];
    double lbp = 0.0, klid = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < CK; e += (int64_t)gridDim.x * blockDim.x) {
        const double r = R[e];
        lbp += LL[e] * r;
        if (r > 0.0) klid += r * (log(r) - p.lidp_kl[p.id_rows == 1 ? (e % K) : e]);
    }
    const double t0 = block_sum(lbp, sh);
    const double t1 = block_sum(klid, sh);
    if (threadIdx.x == 0) {
        double* out = p.part + (size_t)b * p.part_stride + p.off_cell + 2 * blockIdx.x;
        out[0] = t0;
        out[1] = t1;
    }
}

// ---------------------------------------------------------------------------------------------
// k_log_prior: log(prior) as used inside the softmax (vireo_model.py:198,218) and log of the
// row-normalised prior as scipy.stats.entropy sees it (:237-238); one thread per row
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(VB_THREADS) k_log_prior(const double* __restrict__ prior, int64_t n_row, int n_col,
                                                          double* __restrict__ log_raw, double* __restrict__ log_norm) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_row; r += (int64_t)gridDim.x * blockDim.x) {
        const double* p = prior + r * n_col;
        double s = 0.0;
        for (int c = 0; c < n_col; ++c) s += p[c];
        for (int c = 0; c < n_col; ++c) {
            log_raw[r * n_col + c] = log(p[c]);
            log_norm[r * n_col + c] = log(p[c] / s);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// k_elbo: final sums + the convergence rule.  advance = 1 inside the fit loop.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_elbo(const EmP p, const int advance, const double* __restrict__ cell_terms) {
    const int b = blockIdx.y;
    if (advance && p.ctrl[b * VB_CTRL_N]) return;
    elbo_block(p, b, advance, cell_terms);
}

// cell-sharded fit: the block partials {LB_p, KL_ID} of the local cell pass -> two doubles behind S1 | S2 in the
// exchange buffer (summed over the ranks by the next iteration's all-reduce)
__global__ void __launch_bounds__(64) k_xchg_pack(const EmP p, double* __restrict__ out2) {
    if (p.ctrl && p.ctrl[0]) return;
    xchg_pack_block(p, out2);
}

// ---------------------------------------------------------------------------------------------
// doublet tables (vireo_doublet.py:85-136) and the doublet softmax (:64-68)
// ---------------------------------------------------------------------------------------------

// (a, b) of the idx-th pair in itertools.combinations(range(n), 2) order
__device__ __forceinline__ void pair_of(int n, int idx, int& a, int& b) {
    a = 0;
    while (idx >= n - 1 - a) { idx -= n - 1 - a; ++a; }
    b = a + 1 + idx;
}

// columns 0..K-1: singlets, genotype classes G..G2-1 carry zero mass; columns K..K2-1: donor pairs over
// G2 = G + G(G-1)/2 classes.  theta for the mixed classes: mean of mu, geometric mean of sum.
// ab2: [T, 2*G2] digamma differences for the G2 classes, precomputed by k_doublet_theta.
__global__ void __launch_bounds__(64) k_doublet_theta(const double* __restrict__ mu, const double* __restrict__ sum,
                                                      int64_t T, int G, double* __restrict__ ab2) {
    const int G2 = G + G * (G - 1) / 2;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < T * G2; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = e / G2;
        const int c = (int)(e % G2);
        double m_, s_;
        if (c < G) { m_ = mu[t * G + c]; s_ = sum[t * G + c]; }
        else {
            int g1, g2;
            pair_of(G, c - G, g1, g2);
            m_ = (mu[t * G + g1] + mu[t * G + g2]) / 2.0;       // vireo_doublet.py:98
            s_ = sqrt(sum[t * G + g1] * sum[t * G + g2]);       // :99
        }
        const double e1 = s_ * m_, e2 = s_ * (1.0 - m_);        // :49-50
        const double psis = vb_digamma(s_);                     // :51  digamma(beta_sum_both)
        ab2[t * 2 * G2 + c] = vb_digamma(e1) - psis;
        ab2[t * 2 * G2 + G2 + c] = vb_digamma(e2) - psis;
    }
}

// cw == 0: Wt[V][2][K2] (row kernels).  cw = 16 / 8: column chunks of cw as gather tables of the segment kernels,
// Wt[chunk][2V][cw], columns past K2 zero.
__global__ void __launch_bounds__(VB_THREADS) k_doublet_tables(const double* __restrict__ GT, const double* __restrict__ ab2,
                                                               int64_t V, int K, int G, int ase, int cw,
                                                               double* __restrict__ Wt) {
    const int G2 = G + G * (G - 1) / 2;
    const int K2 = K + K * (K - 1) / 2;
    const int KP = cw ? (K2 + cw - 1) / cw * cw : K2;        // columns incl. padding
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < V * KP; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / KP;
        const int c = (int)(e % KP);
        if (c >= K2) {
            const size_t o = ((size_t)(c / cw) * 2 * V + 2 * i) * cw + c % cw;
            Wt[o] = 0.0; Wt[o + cw] = 0.0;
            continue;
        }
        const double* ab = ab2 + (size_t)(ase ? i : 0) * 2 * G2;
        const double* gi = GT + (size_t)i * K * G;
        double wa = 0.0, wb = 0.0;
        if (c < K) {
            for (int g = 0; g < G; ++g) { const double pg = gi[c * G + g]; wa += pg * ab[g]; wb += pg * ab[G2 + g]; }
        } else {
            int k1, k2;
            pair_of(K, c - K, k1, k2);
            const double* A = gi + k1 * G;
            const double* Bq = gi + k2 * G;
            double pr[VB_MAX_GT + VB_MAX_GT * (VB_MAX_GT - 1) / 2];
            double z = 0.0;
            for (int g = 0; g < G; ++g) { pr[g] = A[g] * Bq[g]; z += pr[g]; }                 // :126-127
            int cc = G;
            for (int g1 = 0; g1 < G; ++g1)
                for (int g2 = g1 + 1; g2 < G; ++g2) { pr[cc] = A[g1] * Bq[g2] + A[g2] * Bq[g1]; z += pr[cc]; ++cc; }   // :128-131
            for (int g = 0; g < G2; ++g) { const double pg = pr[g] / z; wa += pg * ab[g]; wb += pg * ab[G2 + g]; }   // :133
        }
        if (cw) {
            const size_t o = ((size_t)(c / cw) * 2 * V + 2 * i) * cw + c % cw;
            Wt[o] = wb; Wt[o + cw] = wa;
        } else {
            Wt[(size_t)i * 2 * K2 + c] = wb;
            Wt[(size_t)i * 2 * K2 + K2 + c] = wa;
        }
    }
}

// one warp per cell: LLR (:64-65) and softmax over all K2 columns with the doublet prior (:67-68)
__global__ void __launch_bounds__(VB_THREADS) k_doublet_softmax(const double* __restrict__ LL, const double* __restrict__ lprior,
                                                                int id_rows, int64_t C, int K, int K2,
                                                                double* __restrict__ prob, double* __restrict__ llr) {
    const int lane = threadIdx.x & 31;
    const int64_t nw = (int64_t)gridDim.x * VB_WARPS;
    for (int64_t j = (int64_t)blockIdx.x * VB_WARPS + (threadIdx.x >> 5); j < C; j += nw) {
        const double* ll = LL + (size_t)j * K2;
        const double* lp = lprior + (size_t)(id_rows == 1 ? 0 : j) * K2;
        double ms = -INFINITY, md = -INFINITY, mx = -INFINITY;
        for (int k = lane; k < K2; k += 32) {
            const double v = ll[k];
            if (k < K) ms = fmax(ms, v); else md = fmax(md, v);
            mx = fmax(mx, v + lp[k]);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            ms = fmax(ms, __shfl_xor_sync(VB_FULL, ms, off));
            md = fmax(md, __shfl_xor_sync(VB_FULL, md, off));
            mx = fmax(mx, __shfl_xor_sync(VB_FULL, mx, off));
        }
        double z = 0.0;
        for (int k = lane; k < K2; k += 32) z += exp(ll[k] + lp[k] - mx);
        z = warp_sum(z);
        for (int k = lane; k < K2; k += 32) prob[(size_t)j * K2 + k] = exp(ll[k] + lp[k] - mx) / z;
        if (lane == 0) llr[j] = md - ms;
    }
}


// ---------------------------------------------------------------------------------------------
// host side: dispatch, iteration driver, C ABI
// ---------------------------------------------------------------------------------------------
static CountsView view_of(const vb_counts* m) {
    CountsView v;
    v.C = m->C; v.V = m->V; v.N = m->N;
    v.g_lo = 0; v.g_hi = INT64_MAX;
    v.fixed = 0; v.o_split = nullptr; v.v_owner = nullptr; v.v_part = nullptr;
    v.cell_ptr = m->cell_ptr; v.cell_idx = m->cell_idx; v.cell_cnt = m->cell_cnt; v.cell_dp = m->cell_dp;
    v.snp_ptr = m->snp_ptr; v.snp_idx = m->snp_idx; v.snp_cnt = m->snp_cnt; v.snp_dp = m->snp_dp;
    return v;
}

// donor-axis tiling: KT lanes x KR registers >= K
static bool tile_for(int K, int& KT, int& KR) {
    if (K < 1 || K > VB_MAX_DONOR) return false;
    KR = 1;
    if (K <= 2) KT = 2; else if (K <= 4) KT = 4; else if (K <= 8) KT = 8; else if (K <= 16) KT = 16;
    else { KT = 32; KR = (K + 31) / 32; }
    return true;
}

#define VB_DISPATCH_TILE(KT_, KR_, WIDE_, ...)                                  \
    do {                                                                          \
        if (KT_ == 2) { constexpr int kt = 2, kr = 1; if (WIDE_) { constexpr bool wd = true; __VA_ARGS__; } else { constexpr bool wd = false; __VA_ARGS__; } } \
        else if (KT_ == 4) { constexpr int kt = 4, kr = 1; if (WIDE_) { constexpr bool wd = true; __VA_ARGS__; } else { constexpr bool wd = false; __VA_ARGS__; } } \
        else if (KT_ == 8) { constexpr int kt = 8, kr = 1; if (WIDE_) { constexpr bool wd = true; __VA_ARGS__; } else { constexpr bool wd = false; __VA_ARGS__; } } \
        else if (KT_ == 16) { constexpr int kt = 16, kr = 1; if (WIDE_) { constexpr bool wd = true; __VA_ARGS__; } else { constexpr bool wd = false; __VA_ARGS__; } } \
        else switch (KR_) {                                                       \
            case 1: { constexpr int kt = 32, kr = 1; if (WIDE_) { constexpr bool wd = true; __VA_ARGS__; } else { constexpr bool wd = false; __VA_ARGS__; } } break; \
            case 2: { constexpr int kt = 32, kr = 2; if (WIDE_) { constexpr bool wd = true; __VA_ARGS__; } else { constexpr bool wd = false; __VA_ARGS__; } } break; \
            case 3: { constexpr int kt = 32, kr = 3; if (WIDE_) { constexpr bool wd = true; __VA_ARGS__; } else { constexpr bool wd = false; __VA_ARGS__; } } break; \
            case 4: { constexpr int kt = 32, kr = 4; if (WIDE_) { constexpr bool wd = true; __VA_ARGS__; } else { constexpr bool wd = false; __VA_ARGS__; } } break; \
            case 5: { constexpr int kt = 32, kr = 5; if (WIDE_) { constexpr bool wd = true; __VA_ARGS__; } else { constexpr bool wd = false; __VA_ARGS__; } } break; \
            case 6: { constexpr int kt = 32, kr = 6; if (WIDE_) { constexpr bool wd = true; __VA_ARGS__; } else { constexpr bool wd = false; __VA_ARGS__; } } break; \
            case 7: { constexpr int kt = 32, kr = 7; if (WIDE_) { constexpr bool wd = true; __VA_ARGS__; } else { constexpr bool wd = false; __VA_ARGS__; } } break; \
            default: { constexpr int kt = 32, kr = 8; if (WIDE_) { constexpr bool wd = true; __VA_ARGS__; } else { constexpr bool wd = false; __VA_ARGS__; } } break; \
        }                                                                         \
    } while (0)

// lane-per-nnz row kernels serve every K <= 8 (VIREO_B200_ROWS_LANE=0 keeps the KT-lanes-per-nnz kernels; =1 the
// round-1 rule: only rows of >= 160 nnz on average)
static bool rows_lane_ok(int K, int64_t nnz, int64_t n_row) {
    static const int mode = getenv("VIREO_B200_ROWS_LANE") ? atoi(getenv("VIREO_B200_ROWS_LANE")) : 2;
    if (mode == 0 || K > 8) return false;
    return mode == 2 || nnz >= 160 * (n_row > 0 ? n_row : 1);
}

// records per lane and round of the lane kernels: 2 (K <= 4) / 1 (K <= 8) by default; VIREO_B200_LANE_DEEP=1 doubles
// them at half the warps per SM (measured slower on the matrices these kernels serve: cfg2 15.7k -> 17.5k it/s,
// cfg5 40 -> 36 ms per BinomMixtureVB.fit call with the shallow variant)
static bool lane_deep() {
    static const bool on = getenv("VIREO_B200_LANE_DEEP") && atoi(getenv("VIREO_B200_LANE_DEEP")) != 0;
    return on;
}

static int launch_cell(const vb_counts* m, const EmP& p, int mode, cudaStream_t st) {
    if (p.tiled == 2 && !p.bmm && m->rA[p.RW == 8 ? 2 : 0].R > 1) return vb_seg_launch_cell_split(m, p, mode, st);
    if (p.tiled >= 2) return vb_seg_launch(m, p, 0, mode == 0 ? GM_CELL : GM_CELL_LL, 0, nullptr, st);
    int KT, KR;
    if (!tile_for(p.K, KT, KR)) { vb_set_error("n_donor=%d outside [1, %d]", p.K, VB_MAX_DONOR); return VB_E_UNSUPPORTED; }
    const CountsView v = view_of(m);
    const dim3 grid(m->grid_cell, p.B);
    if (rows_lane_ok(p.K, m->N, m->C)) {
        const bool deep = lane_deep();
        VB_LAUNCH(3, st, {
            if (p.K <= 4) {
                if (m->wide) { if (deep) k_cell_lane<4, true, 4><<<grid, VB_THREADS, 0, st>>>(v, p, mode); else k_cell_lane<4, true, 2><<<grid, VB_THREADS, 0, st>>>(v, p, mode); }
                else { if (deep) k_cell_lane<4, false, 4><<<grid, VB_THREADS, 0, st>>>(v, p, mode); else k_cell_lane<4, false, 2><<<grid, VB_THREADS, 0, st>>>(v, p, mode); }
            } else {
                if (m->wide) { if (deep) k_cell_lane<8, true, 2><<<grid, VB_THREADS, 0, st>>>(v, p, mode); else k_cell_lane<8, true, 1><<<grid, VB_THREADS, 0, st>>>(v, p, mode); }
                else { if (deep) k_cell_lane<8, false, 2><<<grid, VB_THREADS, 0, st>>>(v, p, mode); else k_cell_lane<8, false, 1><<<grid, VB_THREADS, 0, st>>>(v, p, mode); }
            }
        });
        VB_CUDA(cudaGetLastError());
        return VB_OK;
    }
    VB_LAUNCH(3, st, VB_DISPATCH_TILE(KT, KR, m->wide, k_cell<kt, kr, wd, true><<<grid, VB_THREADS, 0, st>>>(v, p, mode, 0)));
    VB_CUDA(cudaGetLastError());
    return VB_OK;
}

static int launch_snp(const vb_counts* m, const EmP& p, int theta_mode, cudaStream_t st) {
    if (p.tiled >= 2) return vb_seg_launch(m, p, 1, GM_SNP, theta_mode, nullptr, st);
    int KT, KR;
    if (!tile_for(p.K, KT, KR)) { vb_set_error("n_donor=%d outside [1, %d]", p.K, VB_MAX_DONOR); return VB_E_UNSUPPORTED; }
    const CountsView v = view_of(m);
    const dim3 grid(m->grid_snp, p.B);
    if (rows_lane_ok(p.K, m->N, m->V)) {
        const bool deep = lane_deep();
        VB_LAUNCH(0, st, {
            if (p.K <= 4) {
                if (m->wide) { if (deep) k_snp_lane<4, true, 4><<<grid, VB_THREADS, 0, st>>>(v, p, theta_mode); else k_snp_lane<4, true, 2><<<grid, VB_THREADS, 0, st>>>(v, p, theta_mode); }
                else { if (deep) k_snp_lane<4, false, 4><<<grid, VB_THREADS, 0, st>>>(v, p, theta_mode); else k_snp_lane<4, false, 2><<<grid, VB_THREADS, 0, st>>>(v, p, theta_mode); }
            } else {
                if (m->wide) { if (deep) k_snp_lane<8, true, 2><<<grid, VB_THREADS, 0, st>>>(v, p, theta_mode); else k_snp_lane<8, true, 1><<<grid, VB_THREADS, 0, st>>>(v, p, theta_mode); }
                else { if (deep) k_snp_lane<8, false, 2><<<grid, VB_THREADS, 0, st>>>(v, p, theta_mode); else k_snp_lane<8, false, 1><<<grid, VB_THREADS, 0, st>>>(v, p, theta_mode); }
            }
        });
        VB_CUDA(cudaGetLastError());
        return VB_OK;
    }
    VB_LAUNCH(0, st, VB_DISPATCH_TILE(KT, KR, m->wide, k_snp<kt, kr, wd><<<grid, VB_THREADS, 0, st>>>(v, p, theta_mode)));
    VB_CUDA(cudaGetLastError());
    return VB_OK;
}

// ---------------------------------------------------------------------------------------------
// kernel family of the two sparse passes: row kernels (one warp per row, L2 gathers) or the window-segment
// kernels of vb_seg.cu (table windows in shared memory)
// ---------------------------------------------------------------------------------------------
static int g_path = 0;                       // 0 auto, 1 rows, 3 segments (FP64 tables), 4 segments (fixed-point tables)
static int g_graphs = 1;
// theta / ELBO steps run in the tail of the sparse passes inside the fit loops (vb_tail.cuh); VIREO_B200_FUSE=0 keeps
// the stand-alone k_theta / k_elbo launches (same bits either way)
static int g_fuse = -1;
static bool fuse_on() {
    if (g_fuse < 0) g_fuse = (getenv("VIREO_B200_FUSE") && atoi(getenv("VIREO_B200_FUSE")) == 0) ? 0 : 1;
    return g_fuse != 0;
}
extern "C" void vb_set_fuse(int on) { g_fuse = on != 0; }
#define VB_SEG_MIN_NNZ (4ll << 20)           // below this the passes are launch/latency bound either way

extern "C" void vb_set_path(int mode) { g_path = (mode == 1 || mode == 3 || mode == 4) ? mode : 0; }
extern "C" void vb_set_graphs(int on) { g_graphs = on != 0; }

static int kt_for(int K) { return K <= 4 ? 4 : (K <= 8 ? 8 : 16); }

// narrow FP64 rows (8 doubles = 64 bytes) serve n_donor <= 8; VIREO_B200_SEG_NARROW=0 keeps the 16-column rows
static bool seg_narrow_ok(int K) {
    static const bool on = !(getenv("VIREO_B200_SEG_NARROW") && atoi(getenv("VIREO_B200_SEG_NARROW")) == 0);
    return on && kt_for(K) <= 8;
}

// *use = kernel family serving this (counts, K): 0 rows, 2 window segments with FP64 tables, 3 window segments with
// fixed-point tables (`fixed_ok`: the caller can build them).  Formats are built on first use, on `st`.
static int select_family(const vb_counts* mc, int K, int fixed_ok, cudaStream_t st, int* use) {
    vb_counts* m = const_cast<vb_counts*>(mc);
    *use = 0;
    if (g_path == 1) return VB_OK;
    if (K > VB_ROW_DOUBLES) { if (g_path == 0) m->auto_fallback = 4; return VB_OK; }
    const int fp64_fmt = seg_narrow_ok(K) ? 2 : 0;      // FP64 tables: 64-byte rows when 8 columns are enough
    if (g_path == 3 || g_path == 4) {
        int prec = (g_path == 4 && fixed_ok) ? 1 : fp64_fmt;
        int rc = vb_seg_build(m, prec, st);
        if (rc) return rc;
        // the fixed-point kernel keeps odd slots scaled by 2^16: one owner's stream may carry fewer than 2^16 reads
        if (prec == 1 && (m->sA[1].max_reads >= 65536 || m->sB[1].max_reads >= 65536)) {
            prec = fp64_fmt;
            if ((rc = vb_seg_build(m, prec, st))) return rc;
        }
        *use = prec == 1 ? 3 : 2;
        if (*use == 2 && vb_seg_split_rule(m, prec) > 1) vb_seg_build_split(m, prec, st);    // optional: failure keeps the plain cell pass
        return VB_OK;
    }
    // automatic: the window-segment kernels with FP64 tables for large count matrices; row kernels when the passes
    // are launch/latency bound anyway, or when most pairs carry large counts (e.g. mitochondrial clone data) and the
    // residual kernel would do all the work.  A FAILED format build is remembered with its message (vb_counts_info
    // 60 / vb_counts_note): the row kernels then serve the matrix, several times slower, and the caller can tell.
    if (m->N < VB_SEG_MIN_NNZ) { m->auto_fallback = 1; return VB_OK; }
    if (m->seg_failed[fp64_fmt]) { m->auto_fallback = 2; return VB_OK; }
    if (vb_seg_build(m, fp64_fmt, st)) {
        m->auto_fallback = 2;
        snprintf(m->seg_error, sizeof(m->seg_error), "%s", vb_last_error());
        return VB_OK;
    }
    const int64_t pairs = m->sA[fp64_fmt].n_light + m->sA[fp64_fmt].n_heavy;
    if (m->sA[fp64_fmt].n_heavy * 4 > pairs) { m->auto_fallback = 3; return VB_OK; }
    m->auto_fallback = 0;
    *use = 2;
    if (vb_seg_split_rule(m, fp64_fmt) > 1) vb_seg_build_split(m, fp64_fmt, st);             // optional, see above
    return VB_OK;
}

// Block-partial layout.  The capacities cover EVERY kernel family whatever formats happen to be built, so a `part`
// workspace sized once stays valid when another family serves a later call.
static int seg_grid_cap(int64_t n_owner, int sm) {
    const int64_t tasks = (n_owner + VB_SEG_OWNERS - 1) / VB_SEG_OWNERS;
    int64_t cap = (tasks + VB_SEG_MAX_WARPS - 1) / VB_SEG_MAX_WARPS;
    if (cap < sm) cap = sm;
    return (int)cap + 1;
}

static void part_layout(const vb_counts* m, EmP& p) {
    int cap_snp = m->grid_snp > m->grid_elem ? m->grid_snp : m->grid_elem, cap_cell = m->grid_cell, nw;
    const int sc = seg_grid_cap(m->C, m->sm_count), ss = seg_grid_cap(2 * m->V, m->sm_count);
    if (sc > cap_cell) cap_cell = sc;
    if (ss > cap_snp) cap_snp = ss;
    int sa = 0, sb = 0;
    if (p.tiled >= 2) {
        const int fmt = p.tiled == 3 ? 1 : (p.RW == 8 ? 2 : 0);
        vb_seg_geometry(m->sA[fmt], &sa, &nw); vb_seg_geometry(m->sB[fmt], &sb, &nw);
    }
    p.n_snpblk = p.tiled >= 2 ? sb : m->grid_snp;
    p.n_elemblk = m->grid_elem;
    p.n_cellblk = p.tiled >= 2 ? sa : m->grid_cell;
    if (p.tiled == 2 && !p.bmm && m->rA[p.RW == 8 ? 2 : 0].R > 1) p.n_cellblk = m->sm_count;      // row-split cell pass: k_cell_finish
    p.n_klth = (p.bmm || p.ase) ? m->grid_elem : 1;
    p.off_theta = 0;
    p.off_klgt = p.off_theta + cap_snp * 2 * VB_MAX_GT;
    p.off_cell = p.off_klgt + p.n_elemblk;
    p.off_klth = p.off_cell + 2 * cap_cell;
    p.part_stride = p.off_klth + m->grid_elem;
}

// workspace element counts of the family `use`
static void ws_for(const vb_counts* m, int K, int G, int B, int T_is_V, int use, vb_ws_sizes* out) {
    EmP p;
    memset(&p, 0, sizeof(p));
    p.bmm = G == 0; p.ase = T_is_V;
    p.tiled = use;
    p.RW = (use == 2 && seg_narrow_ok(K)) ? 8 : VB_ROW_DOUBLES;
    part_layout(m, p);
    const int64_t T = T_is_V ? m->V : 1;
    const int64_t Kw = use ? p.RW : K;
    out->S = (int64_t)B * m->V * K;
    out->W = (int64_t)B * m->V * Kw * 2;
    out->rpad = use ? (int64_t)B * m->C * p.RW : 0;
    if (use == 3) {   // fixed-point copies (4 bytes per entry) behind the FP64 tables
        out->W += (int64_t)B * m->V * VB_ROW_DOUBLES;
        out->rpad += (int64_t)B * m->C * (VB_ROW_DOUBLES / 2);
    }
    out->heavy = use ? (int64_t)B * (m->C > 2 * m->V ? m->C : 2 * m->V) * p.RW : 0;
    if (use == 2 && !p.bmm) {   // the row-split cell pass keeps its partial sums here: R blocks of [B, C, RW]
        const int64_t split = (int64_t)vb_seg_split_rule(m, p.RW == 8 ? 2 : 0) * B * m->C * p.RW;
        if (split > out->heavy) out->heavy = split;
    }
    out->loglik = (int64_t)B * m->C * K;
    out->ab = (int64_t)B * T * 2 * (G ? G : 1) + (use == 3 ? B : 0);
    out->part = (int64_t)B * p.part_stride;
    out->scal = (int64_t)B * VB_SCAL_N;
    out->ctrl = (int64_t)B * VB_CTRL_N;
}

static int ws_check(const vb_ws_sizes& need, const vb_ws_sizes& have, bool s_external) {
    const bool ok = (s_external || have.S >= need.S) && have.W >= need.W && have.loglik >= need.loglik && have.ab >= need.ab &&
                    have.part >= need.part && have.scal >= need.scal && have.ctrl >= need.ctrl && have.rpad >= need.rpad &&
                    have.heavy >= need.heavy;
    if (!ok) {
        vb_set_error("workspaces do not fit the kernel family in use (need S %lld W %lld loglik %lld ab %lld part %lld rpad %lld "
                     "heavy %lld; have S %lld W %lld loglik %lld ab %lld part %lld rpad %lld heavy %lld): size them with "
                     "vb_*_ws_sizes AFTER the last vb_set_path and pass the sizes in args.ws",
                     (long long)need.S, (long long)need.W, (long long)need.loglik, (long long)need.ab, (long long)need.part,
                     (long long)need.rpad, (long long)need.heavy, (long long)have.S, (long long)have.W, (long long)have.loglik,
                     (long long)have.ab, (long long)have.part, (long long)have.rpad, (long long)have.heavy);
        return VB_E_ARG;
    }
    return VB_OK;
}

static int fill_vireo(const vb_counts* m, const vb_vireo_args* a, EmP& p, cudaStream_t st, bool s_external = false) {
    if (!m || !a) { vb_set_error("NULL argument"); return VB_E_ARG; }
    if (a->n_gt < 1 || a->n_gt > VB_MAX_GT) { vb_set_error("n_GT=%d outside [1, %d]", a->n_gt, VB_MAX_GT); return VB_E_UNSUPPORTED; }
    if (a->n_donor < 1 || a->n_donor > VB_MAX_DONOR) { vb_set_error("n_donor=%d outside [1, %d]", a->n_donor, VB_MAX_DONOR); return VB_E_UNSUPPORTED; }
    if (a->n_batch < 1 || a->n_batch > 65535) { vb_set_error("n_batch=%d outside [1, 65535]", a->n_batch); return VB_E_ARG; }
    memset(&p, 0, sizeof(p));
    p.C = m->C; p.V = m->V; p.T = a->ase_mode ? m->V : 1;
    p.K = a->n_donor; p.G = a->n_gt; p.B = a->n_batch; p.bmm = 0;
    p.ase = a->ase_mode; p.learn_gt = a->learn_gt; p.learn_theta = a->learn_theta; p.fix_beta_sum = a->fix_beta_sum;
    p.id_rows = a->id_prior_rows; p.thp_rows = a->theta_prior_rows;
    p.max_iter = a->max_iter; p.min_iter = a->min_iter; p.delay = a->delay_fit_theta; p.eps = a->epsilon_conv;
    p.R = a->id_prob; p.GT = a->gt_prob; p.mu = a->beta_mu; p.sum = a->beta_sum;
    p.lidp = a->log_id_prior; p.lidp_kl = a->log_id_prior_kl; p.lgtp = a->log_gt_prior; p.lgtp_kl = a->log_gt_prior_kl;
    p.s1p = a->s1_prior; p.s2p = a->s2_prior;
    p.S1 = a->S1; p.S2 = a->S2; p.Wt = a->W; p.ll = a->loglik; p.ab = a->ab; p.part = a->part;
    p.scal = a->scal; p.elbo = a->elbo; p.ctrl = a->ctrl;
    if (!p.R || !p.GT || !p.mu || !p.sum || !p.lidp || !p.lidp_kl || !p.lgtp || !p.lgtp_kl || !p.s1p || !p.s2p ||
        (!s_external && (!p.S1 || !p.S2)) || !p.Wt || !p.ll || !p.ab || !p.part || !p.scal || !p.ctrl) {
        vb_set_error("NULL device pointer in vb_vireo_args");
        return VB_E_ARG;
    }
    int use = 0;
    const int rc = select_family(m, p.K, !p.ase, st, &use);
    if (rc) return rc;
    p.tiled = use; p.KT = kt_for(p.K); p.RP = a->rpad; p.H = a->heavy;
    p.RW = (use == 2 && seg_narrow_ok(p.K)) ? 8 : VB_ROW_DOUBLES;
    if (use && (!p.RP || !p.H)) { vb_set_error("rpad / heavy workspace is NULL (see vb_vireo_ws_sizes)"); return VB_E_ARG; }
    if (use == 3) {   // fixed-point copies live behind the FP64 tables of the same workspaces
        p.Wq = reinterpret_cast<uint32_t*>(p.Wt + (size_t)p.B * p.V * 2 * VB_ROW_DOUBLES);
        p.RPq = reinterpret_cast<uint32_t*>(p.RP + (size_t)p.B * p.C * VB_ROW_DOUBLES);
        p.qscale = p.ab + (size_t)p.B * p.T * 2 * p.G;
    }
    if ((p.id_rows != 1 && p.id_rows != m->C) || (p.thp_rows != 1 && p.thp_rows != p.T)) {
        vb_set_error("prior rows must be 1 or the full extent");
        return VB_E_ARG;
    }
    vb_ws_sizes need;
    ws_for(m, p.K, p.G, p.B, p.ase, use, &need);
    if (ws_check(need, a->ws, s_external)) return VB_E_ARG;
    part_layout(m, p);
    return VB_OK;
}

static int fill_bmm(const vb_counts* m, const vb_bmm_args* a, EmP& p, cudaStream_t st) {
    if (!m || !a) { vb_set_error("NULL argument"); return VB_E_ARG; }
    if (a->n_donor < 1 || a->n_donor > VB_MAX_DONOR) { vb_set_error("n_donor=%d outside [1, %d]", a->n_donor, VB_MAX_DONOR); return VB_E_UNSUPPORTED; }
    if (a->n_batch < 1 || a->n_batch > 65535) { vb_set_error("n_batch=%d outside [1, 65535]", a->n_batch); return VB_E_ARG; }
    memset(&p, 0, sizeof(p));
    p.C = m->C; p.V = m->V; p.T = m->V;
    p.K = a->n_donor; p.G = 1; p.B = a->n_batch; p.bmm = 1;
    p.fix_beta_sum = a->fix_beta_sum; p.learn_theta = 1;
    p.id_rows = a->id_prior_rows; p.thp_rows = (int)m->V;
    p.max_iter = a->max_iter; p.min_iter = a->min_iter; p.eps = a->epsilon_conv;
    p.R = a->id_prob; p.mu = a->beta_mu; p.sum = a->beta_sum;
    p.lidp = a->log_id_prior; p.lidp_kl = a->log_id_prior_kl; p.s1p = a->s1_prior; p.s2p = a->s2_prior;
    p.S1 = a->S1; p.S2 = a->S2; p.Wt = a->W; p.ll = a->loglik; p.part = a->part;
    p.scal = a->scal; p.elbo = a->elbo; p.ctrl = a->ctrl;
    if (!p.R || !p.mu || !p.sum || !p.lidp || !p.lidp_kl || !p.s1p || !p.s2p || !p.S1 || !p.S2 || !p.Wt ||
        !p.ll || !p.part || !p.scal || !p.ctrl) {
        vb_set_error("NULL device pointer in vb_bmm_args");
        return VB_E_ARG;
    }
    int use = 0;
    const int rc = select_family(m, p.K, 0, st, &use);
    if (rc) return rc;
    p.tiled = use; p.KT = kt_for(p.K); p.RP = a->rpad; p.H = a->heavy;
    p.RW = (use == 2 && seg_narrow_ok(p.K)) ? 8 : VB_ROW_DOUBLES;
    if (use && (!p.RP || !p.H)) { vb_set_error("rpad / heavy workspace is NULL (see vb_bmm_ws_sizes)"); return VB_E_ARG; }
    if (p.id_rows != 1 && p.id_rows != m->C) { vb_set_error("id_prior_rows must be 1 or n_cell"); return VB_E_ARG; }
    vb_ws_sizes need;
    ws_for(m, p.K, 0, p.B, 1, use, &need);
    if (ws_check(need, a->ws, false)) return VB_E_ARG;
    part_layout(m, p);
    return VB_OK;
}

static int ws_sizes(const vb_counts* m, int K, int G, int B, int T_is_V, cudaStream_t st, vb_ws_sizes* out) {
    if (!m || !out || K < 1 || B < 1) { vb_set_error("bad argument"); return VB_E_ARG; }
    DeviceGuard dg(m->device);
    int use = 0;
    const int rc = select_family(m, K, G != 0 && !T_is_V, st, &use);
    if (rc) return rc;
    ws_for(m, K, G, B, T_is_V, use, out);
    return VB_OK;
}

extern "C" int vb_vireo_ws_sizes(const vb_counts* m, int n_donor, int n_gt, int n_batch, int ase_mode, void* stream,
                                 vb_ws_sizes* out) {
    if (n_gt < 1) { vb_set_error("n_GT must be >= 1"); return VB_E_ARG; }
    return ws_sizes(m, n_donor, n_gt, n_batch, ase_mode, (cudaStream_t)stream, out);
}
extern "C" int vb_bmm_ws_sizes(const vb_counts* m, int n_donor, int n_batch, void* stream, vb_ws_sizes* out) {
    return ws_sizes(m, n_donor, 0, n_batch, 1, (cudaStream_t)stream, out);
}

// one Vireo iteration; theta_mode / gt flag as documented on the kernels
static int vireo_iteration(const vb_counts* m, const EmP& p, int phases, bool in_loop, cudaStream_t st) {
    int rc;
    const dim3 one(1, p.B), elem(m->grid_elem, p.B);
    const int theta_mode = in_loop ? 2 : ((phases & VB_PH_THETA) ? 1 : 0);
    if (phases & VB_PH_SNP)
        if ((rc = launch_snp(m, p, theta_mode, st))) return rc;
    if (!in_loop && phases == VB_PH_SNP) return VB_OK;      // S1/S2 only (the caller reduces them across devices)
    EmP q = p;
    if (!in_loop && (phases & VB_PH_THETA_SUMS)) {
        VB_LAUNCH(0, st, k_theta_sums<<<elem, VB_THREADS, 0, st>>>(p, 0, 0));
        VB_CUDA(cudaGetLastError());
        q.n_snpblk = m->grid_elem;                           // k_theta sums this kernel's block partials
    }
    // theta always runs: the digamma tables and KL_theta depend on the current beta_mu / beta_sum
    // (p.fuse == 1: the SNP pass did it in its tail)
    if (p.ase) VB_LAUNCH(1, st, k_theta_ase<<<elem, VB_THREADS, 0, st>>>(q, theta_mode));
    else if (p.fuse != 1) VB_LAUNCH(1, st, k_theta<<<one, 2 * VB_MAX_GT * 32, 0, st>>>(q, theta_mode));
    VB_CUDA(cudaGetLastError());
    const int do_gt = in_loop ? p.learn_gt : ((phases & VB_PH_GT) ? 1 : 0);
    VB_LAUNCH(2, st, k_gt<<<elem, VB_THREADS, 0, st>>>(p, do_gt));
    VB_CUDA(cudaGetLastError());
    if (phases & VB_PH_ID) { if ((rc = launch_cell(m, p, 0, st))) return rc; }
    else if (phases & VB_PH_LOGLIK) { if ((rc = launch_cell(m, p, 1, st))) return rc; }
    else if (phases & VB_PH_ELBO) { VB_LAUNCH(6, st, k_terms<<<dim3(p.n_cellblk, p.B), VB_THREADS, 0, st>>>(p)); VB_CUDA(cudaGetLastError()); }
    if ((phases & VB_PH_ELBO) && !(in_loop && p.fuse == 1)) {       // p.fuse == 1: the cell pass closed the iteration in its tail
        VB_LAUNCH(4, st, k_elbo<<<one, 128, 0, st>>>(p, in_loop ? 1 : 0, nullptr));
        VB_CUDA(cudaGetLastError());
    }
    return VB_OK;
}

static int bmm_iteration(const vb_counts* m, const EmP& p, int phases, bool in_loop, cudaStream_t st) {
    int rc;
    const dim3 one(1, p.B), elem(m->grid_elem, p.B);
    if (phases & VB_PH_SNP)
        if ((rc = launch_snp(m, p, 0, st))) return rc;
    VB_LAUNCH(5, st, k_bmm_theta<<<elem, VB_THREADS, 0, st>>>(p, (in_loop || (phases & VB_PH_THETA)) ? 1 : 0));
    VB_CUDA(cudaGetLastError());
    if (phases & VB_PH_ID) { if ((rc = launch_cell(m, p, 0, st))) return rc; }
    else if (phases & VB_PH_LOGLIK) { if ((rc = launch_cell(m, p, 1, st))) return rc; }
    else if (phases & VB_PH_ELBO) { VB_LAUNCH(6, st, k_terms<<<dim3(p.n_cellblk, p.B), VB_THREADS, 0, st>>>(p)); VB_CUDA(cudaGetLastError()); }
    if ((phases & VB_PH_ELBO) && !(in_loop && p.fuse == 1)) {       // p.fuse == 1: the cell pass closed the iteration in its tail
        VB_LAUNCH(4, st, k_elbo<<<one, 128, 0, st>>>(p, in_loop ? 1 : 0, nullptr));
        VB_CUDA(cudaGetLastError());
    }
    return VB_OK;
}

// pinned landing zone for the done flags
static thread_local int32_t* g_pin = nullptr;
static thread_local int g_pin_n = 0;

static int pin_for(int n_ctrl) {
    if (g_pin_n < n_ctrl) {
        if (g_pin) cudaFreeHost(g_pin);
        g_pin = nullptr; g_pin_n = 0;
        VB_CUDA(cudaMallocHost(&g_pin, n_ctrl * sizeof(int32_t)));
        g_pin_n = n_ctrl;
    }
    return VB_OK;
}

static int all_done(const EmP& p, cudaStream_t st, bool* done) {
    const int n_ctrl = p.B * VB_CTRL_N;
    VB_CUDA(cudaMemcpyAsync(g_pin, p.ctrl, n_ctrl * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaStreamSynchronize(st));
    *done = true;
    for (int b = 0; b < p.B; ++b) *done = *done && g_pin[b * VB_CTRL_N];
    return VB_OK;
}

// ---------------------------------------------------------------------------------------------
// Captured iterations.  The kernels of an iteration read the loop state (done flag, iteration counter) from the
// device, so every iteration of a fit is the SAME sequence of launches with the SAME arguments: a group of n
// iterations is captured once into a CUDA graph and replayed.  Worth it where an iteration is launch-bound, i.e.
// on the row-kernel path (small matrices: 5 launches of 5-30 us each); graphs are cached per thread, keyed by every
// launch argument (EmP, the staged arrays, the launch geometry, n).
// ---------------------------------------------------------------------------------------------
struct GraphKey {
    EmP p;
    CountsView v;
    int grid_cell, grid_snp, grid_elem, n_iter, device;
};
struct GraphEntry {
    GraphKey key;
    cudaGraphExec_t exec;
    int64_t launches[8];
    uint64_t stamp;
};
#define VB_GRAPH_SLOTS 8
static thread_local GraphEntry g_graphs_cache[VB_GRAPH_SLOTS];
static thread_local int g_graphs_n = 0;
static thread_local uint64_t g_graph_stamp = 0;
static thread_local cudaStream_t g_cap_stream[64] = {nullptr};

static int graph_for(const vb_counts* m, const EmP& p, int n_iter, GraphEntry** out) {
    GraphKey key;
    memset(&key, 0, sizeof(key));
    key.p = p; key.v = view_of(m);
    key.grid_cell = m->grid_cell; key.grid_snp = m->grid_snp; key.grid_elem = m->grid_elem; key.n_iter = n_iter; key.device = m->device;
    for (int i = 0; i < g_graphs_n; ++i)
        if (!memcmp(&g_graphs_cache[i].key, &key, sizeof(key))) { g_graphs_cache[i].stamp = ++g_graph_stamp; *out = &g_graphs_cache[i]; return VB_OK; }
    const int dslot = m->device >= 0 && m->device < 64 ? m->device : 0;
    if (!g_cap_stream[dslot]) VB_CUDA(cudaStreamCreateWithFlags(&g_cap_stream[dslot], cudaStreamNonBlocking));
    cudaStream_t cs = g_cap_stream[dslot];
    int64_t before[8];
    memcpy(before, g_launches, sizeof(before));
    VB_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    int rc = VB_OK;
    const int all = VB_PH_SNP | VB_PH_THETA | VB_PH_GT | VB_PH_ID | VB_PH_ELBO;
    for (int it = 0; it < n_iter && !rc; ++it) rc = p.bmm ? bmm_iteration(m, p, all, true, cs) : vireo_iteration(m, p, all, true, cs);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(cs, &graph);
    int64_t per[8];
    for (int i = 0; i < 8; ++i) { per[i] = g_launches[i] - before[i]; g_launches[i] = before[i]; }   // counted per replay instead
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) { vb_set_error("stream capture failed: %s", cudaGetErrorString(ce)); cudaGetLastError(); return VB_E_CUDA; }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) { vb_set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(ie)); cudaGetLastError(); return VB_E_CUDA; }
    int slot = g_graphs_n;
    if (g_graphs_n < VB_GRAPH_SLOTS) ++g_graphs_n;
    else {
        slot = 0;
        for (int i = 1; i < VB_GRAPH_SLOTS; ++i) if (g_graphs_cache[i].stamp < g_graphs_cache[slot].stamp) slot = i;
        cudaGraphExecDestroy(g_graphs_cache[slot].exec);
    }
    g_graphs_cache[slot].key = key; g_graphs_cache[slot].exec = exec; g_graphs_cache[slot].stamp = ++g_graph_stamp;
    memcpy(g_graphs_cache[slot].launches, per, sizeof(per));
    *out = &g_graphs_cache[slot];
    return VB_OK;
}

static int run_loop(const vb_counts* m, const EmP& p0, int poll_every, cudaStream_t st) {
    EmP p = p0;
    p.fuse = fuse_on() ? 1 : 0;                 // theta / ELBO in the tails of the sparse passes
    if (p.max_iter < 1) { vb_set_error("max_iter must be >= 1"); return VB_E_ARG; }
    if (!p.elbo) { vb_set_error("elbo output is NULL"); return VB_E_ARG; }
    const int n_ctrl = p.B * VB_CTRL_N;
    int rc = pin_for(n_ctrl);
    if (rc) return rc;
    VB_CUDA(cudaMemsetAsync(p.ctrl, 0, n_ctrl * sizeof(int32_t), st));
    if (p.tiled) {   // the SNP pass gathers ID_prob from its padded-row copy
        rc = vb_pad_rows_launch(m, p.R, p.C, p.K, p.KT, p.RW, p.B, p.RP, st);
        if (!rc && p.tiled == 3) rc = vb_seg_quantise_rows(m, p, st);
        if (rc) return rc;
    }
    if (poll_every <= 0) poll_every = 16;
    const int all = VB_PH_SNP | VB_PH_THETA | VB_PH_GT | VB_PH_ID | VB_PH_ELBO;
    if (g_graphs && !g_prof_on && p.tiled == 0) {
        // launch-bound iterations: replay captured groups of up to 16 iterations; the done flags are polled between groups
        int group = poll_every < 16 ? poll_every : 16;
        if (group > p.max_iter) group = p.max_iter;
        int it = 0;
        bool ok = true;
        while (it < p.max_iter) {
            const int n = p.max_iter - it < group ? p.max_iter - it : group;
            GraphEntry* ge = nullptr;
            if (graph_for(m, p, n, &ge)) { ok = false; break; }       // capture unavailable: plain launches below
            VB_CUDA(cudaGraphLaunch(ge->exec, st));
            for (int i = 0; i < 8; ++i) g_launches[i] += ge->launches[i];
            it += n;
            if (it < p.max_iter && it % poll_every == 0) {
                bool done;
                if ((rc = all_done(p, st, &done))) return rc;
                if (done) break;
            }
        }
        if (ok) return VB_OK;
        if (it > 0) { vb_set_error("graph capture failed mid-fit: %s", vb_last_error()); return VB_E_CUDA; }
    }
    for (int it = 0; it < p.max_iter; ++it) {
        rc = p.bmm ? bmm_iteration(m, p, all, true, st) : vireo_iteration(m, p, all, true, st);
        if (rc) return rc;
        if ((it + 1) % poll_every == 0 && it + 1 < p.max_iter) {
            bool done;
            if ((rc = all_done(p, st, &done))) return rc;
            if (done) break;
        }
    }
    return VB_OK;
}

extern "C" int vb_log_prior(const double* prior, int64_t n_row, int n_col, double* log_raw, double* log_norm, void* stream) {
    if (!prior || !log_raw || !log_norm || n_row < 0 || n_col < 1) { vb_set_error("bad argument"); return VB_E_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    int64_t nb = (n_row + VB_THREADS - 1) / VB_THREADS;
    if (nb > 148 * 8) nb = 148 * 8;
    if (nb < 1) nb = 1;
    VB_LAUNCH(7, st, k_log_prior<<<(unsigned)nb, VB_THREADS, 0, st>>>(prior, n_row, n_col, log_raw, log_norm));
    VB_CUDA(cudaGetLastError());
    return VB_OK;
}

extern "C" int vb_vireo_fit(const vb_counts* m, const vb_vireo_args* a, void* stream) {
    if (!m) { vb_set_error("NULL handle"); return VB_E_ARG; }
    DeviceGuard dg(m->device);
    EmP p;
    int rc = fill_vireo(m, a, p, (cudaStream_t)stream);
    if (rc) return rc;
    return run_loop(m, p, a->poll_every, (cudaStream_t)stream);
}

extern "C" int vb_vireo_step(const vb_counts* m, const vb_vireo_args* a, int phases, void* stream) {
    if (!m) { vb_set_error("NULL handle"); return VB_E_ARG; }
    DeviceGuard dg(m->device);
    EmP p;
    int rc = fill_vireo(m, a, p, (cudaStream_t)stream);
    if (rc) return rc;
    p.ctrl = nullptr;                       // single phases never consult the loop state
    if (p.tiled && (phases & VB_PH_SNP)) {
        rc = vb_pad_rows_launch(m, p.R, p.C, p.K, p.KT, p.RW, p.B, p.RP, (cudaStream_t)stream);
        if (!rc && p.tiled == 3) rc = vb_seg_quantise_rows(m, p, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return vireo_iteration(m, p, phases, false, (cudaStream_t)stream);
}

extern "C" int vb_bmm_fit(const vb_counts* m, const vb_bmm_args* a, void* stream) {
    if (!m) { vb_set_error("NULL handle"); return VB_E_ARG; }
    DeviceGuard dg(m->device);
    EmP p;
    int rc = fill_bmm(m, a, p, (cudaStream_t)stream);
    if (rc) return rc;
    return run_loop(m, p, a->poll_every, (cudaStream_t)stream);
}

extern "C" int vb_bmm_step(const vb_counts* m, const vb_bmm_args* a, int phases, void* stream) {
    if (!m) { vb_set_error("NULL handle"); return VB_E_ARG; }
    DeviceGuard dg(m->device);
    EmP p;
    int rc = fill_bmm(m, a, p, (cudaStream_t)stream);
    if (rc) return rc;
    p.ctrl = nullptr;
    if (p.tiled && (phases & VB_PH_SNP)) {
        rc = vb_pad_rows_launch(m, p.R, p.C, p.K, p.KT, p.RW, p.B, p.RP, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return bmm_iteration(m, p, phases, false, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// doublet pass
// ---------------------------------------------------------------------------------------------
// column-chunk width of the doublet pass: 0 = row kernels, 16 / 8 = window-segment kernels with the FP64 format
// that is (or gets) built for this matrix
static int doublet_chunk(const vb_counts* mc, cudaStream_t st, int* fmt) {
    vb_counts* m = const_cast<vb_counts*>(mc);
    *fmt = 0;
    if (g_path == 1) return 0;
    if (g_path == 0) {
        if (m->N < VB_SEG_MIN_NNZ) return 0;
        for (int f = 0; f <= 2; f += 2) {
            const int64_t pairs = m->sA[f].n_light + m->sA[f].n_heavy;
            if (m->sA[f].built && m->sA[f].n_heavy * 4 > pairs) return 0;        // residual-dominated: rows
        }
    }
    if (m->sA[0].built && m->sB[0].built) { *fmt = 0; return 16; }
    if (m->sA[2].built && m->sB[2].built) { *fmt = 2; return 8; }
    if (m->seg_failed[0] || vb_seg_build(m, 0, st)) return 0;
    if (g_path == 0 && m->sA[0].n_heavy * 4 > m->sA[0].n_light + m->sA[0].n_heavy) return 0;
    *fmt = 0;
    return 16;
}

extern "C" int vb_doublet_ws_sizes(const vb_counts* m, int n_donor, int n_gt, int ase_mode, void* stream, vb_doublet_ws* out) {
    if (!m || !out || n_donor < 1 || n_gt < 1 || n_gt > VB_MAX_GT) { vb_set_error("bad argument"); return VB_E_ARG; }
    DeviceGuard dg(m->device);
    const int64_t K2 = n_donor + (int64_t)n_donor * (n_donor - 1) / 2, G2 = n_gt + n_gt * (n_gt - 1) / 2;
    int fmt;
    const int cw = doublet_chunk(m, (cudaStream_t)stream, &fmt);
    const int64_t KP = cw ? (K2 + cw - 1) / cw * cw : K2;
    out->W = 2 * m->V * KP;
    out->heavy = cw ? m->C * cw : 0;
    out->ab2 = (ase_mode ? m->V : 1) * 2 * G2;
    return VB_OK;
}

extern "C" int vb_vireo_doublet(const vb_counts* m, int n_donor, int n_gt, int ase_mode,
                                const double* gt_prob, const double* beta_mu, const double* beta_sum,
                                const double* log_prior_both, int id_prior_rows,
                                double* W, double* heavy, double* ab2, const vb_doublet_ws* ws,
                                double* loglik_out, double* prob_out, double* llr_out,
                                void* stream) {
    if (!m || !gt_prob || !beta_mu || !beta_sum || !log_prior_both || !W || !ab2 || !ws || !loglik_out || !prob_out || !llr_out) {
        vb_set_error("NULL argument");
        return VB_E_ARG;
    }
    if (n_gt < 1 || n_gt > VB_MAX_GT || n_donor < 1) { vb_set_error("unsupported n_GT=%d / n_donor=%d", n_gt, n_donor); return VB_E_UNSUPPORTED; }
    DeviceGuard dg(m->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int K = n_donor, G = n_gt;
    const int K2 = K + K * (K - 1) / 2, G2 = G + G * (G - 1) / 2;
    const int64_t T = ase_mode ? m->V : 1;
    int fmt;
    const int cw = doublet_chunk(m, st, &fmt);
    const int64_t KP = cw ? (K2 + cw - 1) / cw * cw : K2;
    if (ws->W < 2 * m->V * KP || ws->ab2 < T * 2 * G2 || (cw && (!heavy || ws->heavy < m->C * cw))) {
        vb_set_error("doublet workspaces too small for the kernel family in use (see vb_doublet_ws_sizes)");
        return VB_E_ARG;
    }
    VB_LAUNCH(7, st, k_doublet_theta<<<(int)((T * G2 + 63) / 64 > 4096 ? 4096 : (T * G2 + 63) / 64), 64, 0, st>>>(beta_mu, beta_sum, T, G, ab2));
    VB_CUDA(cudaGetLastError());
    VB_LAUNCH(7, st, k_doublet_tables<<<m->grid_elem, VB_THREADS, 0, st>>>(gt_prob, ab2, m->V, K, G, ase_mode, cw, W));
    VB_CUDA(cudaGetLastError());
    EmP p;
    memset(&p, 0, sizeof(p));
    p.C = m->C; p.V = m->V; p.B = 1; p.ll = loglik_out; p.id_rows = 1;
    if (cw) {
        // one pass of the record stream per chunk of cw columns: chunk c's table is W[c] = [2V, cw]
        p.tiled = 2; p.RW = cw; p.KT = cw; p.K = cw; p.H = heavy;
        for (int k0 = 0, c = 0; k0 < K2; k0 += cw, ++c) {
            p.Wt = W + (size_t)c * 2 * m->V * cw;
            SegPlain pl;
            pl.out = loglik_out; pl.ld = K2; pl.off = k0; pl.cols = K2 - k0 < cw ? K2 - k0 : cw; pl.set = nullptr; pl.row0 = 0;
            const int rc = vb_seg_launch(m, p, 0, GM_PLAIN, 0, &pl, st);
            if (rc) return rc;
        }
    } else {
        // cell-major pass over column chunks of the K2-wide tables, row kernels
        p.K = K2; p.Wt = W;
        const CountsView v = view_of(m);
        const int chunk = K2 <= 16 ? (K2 <= 2 ? 2 : K2 <= 4 ? 4 : K2 <= 8 ? 8 : 16) : 128;
        for (int k0 = 0; k0 < K2; k0 += chunk) {
            const int width = K2 - k0 < chunk ? K2 - k0 : chunk;
            int KT, KR;
            tile_for(width, KT, KR);
            const dim3 grid(m->grid_cell, 1);
            VB_LAUNCH(3, st, VB_DISPATCH_TILE(KT, KR, m->wide, k_cell<kt, kr, wd, false><<<grid, VB_THREADS, 0, st>>>(v, p, 1, k0)));
            VB_CUDA(cudaGetLastError());
        }
    }
    VB_LAUNCH(7, st, k_doublet_softmax<<<m->grid_cell, VB_THREADS, 0, st>>>(loglik_out, log_prior_both, id_prior_rows, m->C, K, K2, prob_out, llr_out));
    VB_CUDA(cudaGetLastError());
    return VB_OK;
}

// ---------------------------------------------------------------------------------------------
// one fit over several GPUs: cells sharded over ranks, NCCL bound at run time
// ---------------------------------------------------------------------------------------------
#include <dlfcn.h>

namespace {
typedef struct ncclComm* nccl_comm_t;
struct nccl_uid { char internal[VB_COMM_ID_BYTES]; };
enum { NCCL_DOUBLE = 8, NCCL_SUM = 0 };            // ncclFloat64, ncclSum (nccl.h, stable since 2.0)
struct NcclApi {
    void* lib;
    int (*GetUniqueId)(nccl_uid*);
    int (*CommInitRank)(nccl_comm_t*, int, nccl_uid, int);
    int (*CommDestroy)(nccl_comm_t);
    int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t);
    int (*Broadcast)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t);
    int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t);
    const char* (*GetErrorString)(int);
};
NcclApi g_nccl = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};

int nccl_load() {
    if (g_nccl.lib) return VB_OK;
    // the copy the process already holds (PyTorch loads its bundled libnccl.so.2) is found by its soname
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { vb_set_error("libnccl.so.2 not found: %s", dlerror()); return VB_E_UNSUPPORTED; }
    NcclApi a;
    a.lib = h;
    *(void**)&a.GetUniqueId = dlsym(h, "ncclGetUniqueId");
    *(void**)&a.CommInitRank = dlsym(h, "ncclCommInitRank");
    *(void**)&a.CommDestroy = dlsym(h, "ncclCommDestroy");
    *(void**)&a.AllReduce = dlsym(h, "ncclAllReduce");
    *(void**)&a.Broadcast = dlsym(h, "ncclBroadcast");
    *(void**)&a.AllGather = dlsym(h, "ncclAllGather");
    *(void**)&a.GetErrorString = dlsym(h, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce || !a.Broadcast || !a.AllGather || !a.GetErrorString) {
        vb_set_error("libnccl.so.2 lacks a required symbol");
        return VB_E_UNSUPPORTED;
    }
    g_nccl = a;
    return VB_OK;
}
}  // namespace

struct vb_comm {
    int device, n_ranks, rank;
    nccl_comm_t comm;
};

#define VB_NCCL(call)                                                                              \
    do {                                                                                           \
        const int r__ = (call);                                                                    \
        if (r__ != 0) { vb_set_error("%s failed: %s", #call, g_nccl.GetErrorString(r__)); return VB_E_CUDA; } \
    } while (0)

extern "C" int vb_comm_unique_id(void* id_out) {
    if (!id_out) { vb_set_error("NULL argument"); return VB_E_ARG; }
    const int rc = nccl_load();
    if (rc) return rc;
    nccl_uid id;
    VB_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return VB_OK;
}

extern "C" int vb_comm_create(int device, int n_ranks, int rank, const void* id, vb_comm** out) {
    if (!out || n_ranks < 1 || rank < 0 || rank >= n_ranks) { vb_set_error("bad argument"); return VB_E_ARG; }
    *out = nullptr;
    vb_comm* c = new vb_comm();
    c->device = device; c->n_ranks = n_ranks; c->rank = rank; c->comm = nullptr;
    if (n_ranks > 1) {
        if (!id) { delete c; vb_set_error("unique id is NULL"); return VB_E_ARG; }
        int rc = nccl_load();
        if (rc) { delete c; return rc; }
        DeviceGuard dg(device);
        nccl_uid uid;
        memcpy(&uid, id, sizeof(uid));
        const int r = g_nccl.CommInitRank(&c->comm, n_ranks, uid, rank);
        if (r != 0) { vb_set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r)); delete c; return VB_E_CUDA; }
    }
    *out = c;
    return VB_OK;
}

extern "C" void vb_comm_destroy(vb_comm* c) {
    if (!c) return;
    if (c->comm) { DeviceGuard dg(c->device); g_nccl.CommDestroy(c->comm); }
    delete c;
}

extern "C" int vb_comm_allreduce(vb_comm* c, double* buf, int64_t n, void* stream) {
    if (!c || !buf || n < 0) { vb_set_error("bad argument"); return VB_E_ARG; }
    if (c->n_ranks == 1 || n == 0) return VB_OK;
    DeviceGuard dg(c->device);
    VB_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, NCCL_DOUBLE, NCCL_SUM, c->comm, (cudaStream_t)stream));
    return VB_OK;
}

extern "C" int vb_comm_broadcast(vb_comm* c, double* buf, int64_t n, int root, void* stream) {
    if (!c || !buf || n < 0 || root < 0 || root >= c->n_ranks) { vb_set_error("bad argument"); return VB_E_ARG; }
    if (c->n_ranks == 1 || n == 0) return VB_OK;
    DeviceGuard dg(c->device);
    VB_NCCL(g_nccl.Broadcast(buf, buf, (size_t)n, NCCL_DOUBLE, root, c->comm, (cudaStream_t)stream));
    return VB_OK;
}

extern "C" int vb_comm_allgather(vb_comm* c, const double* send, double* recv, int64_t n_per_rank, void* stream) {
    if (!c || !send || !recv || n_per_rank < 0) { vb_set_error("bad argument"); return VB_E_ARG; }
    DeviceGuard dg(c->device);
    if (c->n_ranks == 1) {
        if (send != recv && n_per_rank) VB_CUDA(cudaMemcpyAsync(recv, send, n_per_rank * sizeof(double), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        return VB_OK;
    }
    if (n_per_rank == 0) return VB_OK;
    VB_NCCL(g_nccl.AllGather(send, recv, (size_t)n_per_rank, NCCL_DOUBLE, c->comm, (cudaStream_t)stream));
    return VB_OK;
}

static int sharded_args(const vb_counts* m, const vb_vireo_args* a, vb_comm* c, double* xbuf, cudaStream_t st, EmP& p) {
    if (!m || !a || !c || !xbuf) { vb_set_error("NULL argument"); return VB_E_ARG; }
    if (c->device != m->device) { vb_set_error("communicator and counts live on different devices"); return VB_E_ARG; }
    if (a->n_batch != 1 || a->ase_mode) { vb_set_error("cell-sharded fit: one restart, no ASE mode (theta per SNP needs no exchange of S)"); return VB_E_UNSUPPORTED; }
    const int rc = fill_vireo(m, a, p, st, true);
    if (rc) return rc;
    p.S1 = xbuf;
    p.S2 = xbuf + (size_t)p.V * p.K;
    return VB_OK;
}

extern "C" int vb_vireo_fit_sharded(const vb_counts* m, const vb_vireo_args* a, vb_comm* c, double* xbuf, void* stream) {
    if (!m) { vb_set_error("NULL handle"); return VB_E_ARG; }
    DeviceGuard dg(m->device);
    cudaStream_t st = (cudaStream_t)stream;
    EmP p;
    int rc = sharded_args(m, a, c, xbuf, st, p);
    if (rc) return rc;
    if (p.max_iter < 1) { vb_set_error("max_iter must be >= 1"); return VB_E_ARG; }
    if (!p.elbo) { vb_set_error("elbo output is NULL"); return VB_E_ARG; }
    if ((rc = pin_for(VB_CTRL_N))) return rc;
    const int64_t n_x = 2 * p.V * p.K;
    double* xs = xbuf + n_x;                                  // {LB_p, KL_ID} of the previous iteration, summed with S
    VB_CUDA(cudaMemsetAsync(p.ctrl, 0, VB_CTRL_N * sizeof(int32_t), st));
    VB_CUDA(cudaMemsetAsync(xs, 0, 8 * sizeof(double), st));
    if (p.tiled) {
        rc = vb_pad_rows_launch(m, p.R, p.C, p.K, p.KT, p.RW, 1, p.RP, st);
        if (!rc && p.tiled == 3) rc = vb_seg_quantise_rows(m, p, st);
        if (rc) return rc;
    }
    int poll_every = a->poll_every > 0 ? a->poll_every : 16;
    const dim3 one(1, 1), elem(m->grid_elem, 1);
    const bool fused = fuse_on();
    p.xs = xs;
    p.fuse = fused ? 2 : 0;                                   // cell pass packs its two ELBO terms in its tail
    EmP q = p;
    q.n_snpblk = m->grid_elem;                                // k_theta sums the block partials of k_theta_sums
    for (int it = 0; it < p.max_iter; ++it) {
        if ((rc = launch_snp(m, p, 0, st))) return rc;        // local S1 | S2 (theta sums need the reduced S: not here)
        if (c->n_ranks > 1) VB_NCCL(g_nccl.AllReduce(xbuf, xbuf, (size_t)(n_x + 2), NCCL_DOUBLE, NCCL_SUM, c->comm, st));
        if (fused) {
            // theta sums; the last CTA closes iteration it - 1 (ELBO, convergence rule) and finishes theta
            VB_LAUNCH(0, st, k_theta_sums<<<elem, VB_THREADS, 0, st>>>(p, 1, it == 0));
            VB_CUDA(cudaGetLastError());
        } else {
            if (it > 0) { VB_LAUNCH(4, st, k_elbo<<<one, 128, 0, st>>>(p, 1, xs)); VB_CUDA(cudaGetLastError()); }
            VB_LAUNCH(0, st, k_theta_sums<<<elem, VB_THREADS, 0, st>>>(p, 0, 0));
            VB_CUDA(cudaGetLastError());
            VB_LAUNCH(1, st, k_theta<<<one, 2 * VB_MAX_GT * 32, 0, st>>>(q, 2));
            VB_CUDA(cudaGetLastError());
        }
        VB_LAUNCH(2, st, k_gt<<<elem, VB_THREADS, 0, st>>>(p, p.learn_gt));
        VB_CUDA(cudaGetLastError());
        if ((rc = launch_cell(m, p, 0, st))) return rc;
        if (!fused) { VB_LAUNCH(7, st, k_xchg_pack<<<1, 64, 0, st>>>(p, xs)); VB_CUDA(cudaGetLastError()); }
        if ((it + 1) % poll_every == 0 && it + 1 < p.max_iter) {
            // the flag read here is the verdict on iteration it - 1: identical on every rank (same reduced numbers)
            bool done;
            if ((rc = all_done(p, st, &done))) return rc;
            if (done) break;
        }
    }
    // ELBO and convergence rule of the last executed iteration (a no-op if an earlier one already ended the fit)
    if (c->n_ranks > 1) VB_NCCL(g_nccl.AllReduce(xs, xs, 2, NCCL_DOUBLE, NCCL_SUM, c->comm, st));
    VB_LAUNCH(4, st, k_elbo<<<one, 128, 0, st>>>(p, 1, xs));
    VB_CUDA(cudaGetLastError());
    return VB_OK;
}

extern "C" int vb_vireo_gt_sharded(const vb_counts* m, const vb_vireo_args* a, vb_comm* c, double* xbuf, void* stream) {
    if (!m) { vb_set_error("NULL handle"); return VB_E_ARG; }
    DeviceGuard dg(m->device);
    cudaStream_t st = (cudaStream_t)stream;
    EmP p;
    int rc = sharded_args(m, a, c, xbuf, st, p);
    if (rc) return rc;
    p.ctrl = nullptr;
    if (p.tiled) {
        rc = vb_pad_rows_launch(m, p.R, p.C, p.K, p.KT, p.RW, 1, p.RP, st);
        if (!rc && p.tiled == 3) rc = vb_seg_quantise_rows(m, p, st);
        if (rc) return rc;
    }
    if ((rc = launch_snp(m, p, 0, st))) return rc;
    if (c->n_ranks > 1) VB_NCCL(g_nccl.AllReduce(xbuf, xbuf, (size_t)(2 * p.V * p.K), NCCL_DOUBLE, NCCL_SUM, c->comm, st));
    VB_LAUNCH(1, st, k_theta<<<dim3(1, 1), 2 * VB_MAX_GT * 32, 0, st>>>(p, 0));
    VB_CUDA(cudaGetLastError());
    VB_LAUNCH(2, st, k_gt<<<dim3(m->grid_elem, 1), VB_THREADS, 0, st>>>(p, 1));
    VB_CUDA(cudaGetLastError());
    return VB_OK;
}
