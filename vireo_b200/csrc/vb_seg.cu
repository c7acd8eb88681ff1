// vb_seg.cu -- the window-segment kernels: both sparse passes of the EM iteration with the gathered
// table staged window by window through shared memory (cp.async.bulk + mbarrier) and a small lane
// group bound to every owner row.
//
// Both passes are "for every owner: sum over its pairs of count * table[gather row][0:16]":
//   cell pass  logLik_ID[j,:] = sum_i  (dp-ad)_ij * W[2i,:] + ad_ij * W[2i+1,:]
//              (vireoSNP/utils/vireo_model.py:190-196, bmm_model.py:125-129)
//   SNP  pass  S2[i,:] = sum_j (dp-ad)_ij * ID_prob[j,:],  S1[i,:] = sum_j ad_ij * ID_prob[j,:]
//              (vireo_model.py:168-170,207-209, bmm_model.py:136-138)
// What bounds such a pass on B200 is the shared-memory crossbar (128 B/clk/SM): every pair needs one
// table row.  A lane-per-owner layout (the retired gather-stream kernels of round 1) spends 1.5 crossbar
// wavefronts per 128-byte row because a quarter warp is rarely full.  Here
//   * 4 lanes share one owner and read its table row with 16-byte loads (two per lane for a 128-byte
//     FP64 row -- one from each 64-byte half, even lane groups starting in the lower half and odd groups
//     in the upper half so that the two rows of a wavefront never meet in a bank; one per lane for the
//     64-byte rows), so a wavefront always carries whole rows;
//   * a warp task holds 32 owners: 8 of them are served per warp step, the accumulators of all 32
//     stay in registers (static indexing: the slot index is the unrolled loop variable);
//   * the table streams through NB window buffers; a producer warp refills a buffer when every
//     consumer warp has released it (full/empty mbarriers), consumers wait per window, not per record;
//   * the records of a task are stored per window as super-steps of 32 x 16 bits; a segment (the
//     super-steps of one window) may also carry pairs of the NEXT window, which is resident as well,
//     in the slots that lock-step would otherwise pad with null records (look-ahead fill: 59% -> 85%
//     useful slots); segments are padded to groups of DEPTH super-steps so that the register queue
//     that prefetches them DEPTH super-steps ahead needs no rotation;
//   * PREC 0: FP64 tables of 16 columns (rows of 128 bytes);  PREC 2: FP64 tables of 8 columns for
//     n_donor <= 8 (rows of 64 bytes);  PREC 1 (opt-in) keeps 16 columns as unsigned 32-bit fixed point
//     (rows of 64 bytes: half the crossbar traffic) and accumulates count * value exactly in 64-bit
//     integers, so the result does not depend on the summation order; the quantisation error is bounded
//     per owner by sum(count) * 2^-33 * range and reported by vb_counts_info.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <stdlib.h>
#include <string.h>

#include <type_traits>
#include <vector>

#include "vb_common.cuh"
#include "vb_stream.cuh"
#include "vb_tail.cuh"

// ---------------------------------------------------------------------------------------------
// build kernels
// ---------------------------------------------------------------------------------------------

// pairs per owner: stream pairs (count <= 31), residual pairs, reads carried by the stream pairs
template <int ORI, bool WIDE>
__global__ void k_sg_count(const CountsView m, int64_t n_owner, uint32_t* __restrict__ n_light, uint32_t* __restrict__ n_heavy,
                           uint32_t* __restrict__ reads, unsigned int* flags) {
    for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < n_owner; o += (int64_t)gridDim.x * blockDim.x) {
        uint32_t nl = 0, nh = 0, rd = 0;
        const bool ok = for_records<ORI, WIDE>(m, o, [&](int, uint32_t c) {
            if (c > VB_SEG_MAX_COUNT) ++nh;
            else { ++nl; rd += c; }
        });
        if (!ok) atomicOr(&flags[0], 1u);
        n_light[o] = nl;
        n_heavy[o] = nh;
        reads[o] = rd;
    }
}

__global__ void k_sg_iota(int32_t* out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (int32_t)i;
}

// owners at sorted ranks >= first_sparse are served by the residual kernel entirely
__global__ void k_sg_mark_sparse(const int32_t* __restrict__ perm_sorted, int64_t first_sparse, int64_t n_owner,
                                 uint8_t* __restrict__ sparse, uint32_t* __restrict__ n_light, uint32_t* __restrict__ n_heavy,
                                 uint32_t* __restrict__ reads) {
    for (int64_t r = first_sparse + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_owner; r += (int64_t)gridDim.x * blockDim.x) {
        const int32_t o = perm_sorted[r];
        sparse[o] = 1;
        n_heavy[o] += n_light[o];
        n_light[o] = 0;
        reads[o] = 0;
    }
}

__global__ void k_sg_sums(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, const uint32_t* __restrict__ c,
                          int64_t n, unsigned long long* out) {
    unsigned long long sa = 0, sb = 0, mc = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        sa += a[i]; sb += b[i];
        mc = c[i] > mc ? c[i] : mc;
    }
    for (int off = 16; off > 0; off >>= 1) {
        sa += __shfl_xor_sync(VB_FULL, sa, off);
        sb += __shfl_xor_sync(VB_FULL, sb, off);
        const unsigned long long o = __shfl_xor_sync(VB_FULL, mc, off);
        mc = o > mc ? o : mc;
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(out, sa); atomicAdd(out + 1, sb); atomicMax(out + 2, mc); }
}

// pairs of each stream owner per table window; one thread per sorted position
template <int ORI, bool WIDE>
__global__ void k_sg_wincount(const CountsView m, int64_t n_active, const int32_t* __restrict__ perm, int win_rows, int n_win,
                              uint16_t* __restrict__ cnt) {
    for (int64_t pos = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; pos < n_active; pos += (int64_t)gridDim.x * blockDim.x) {
        const int64_t o = perm[pos];
        uint16_t* row = cnt + (size_t)pos * n_win;
        int w = 0;
        uint32_t n = 0;
        for_records<ORI, WIDE>(m, o, [&](int g, uint32_t c) {
            if (c > VB_SEG_MAX_COUNT) return;
            const int gw = g / win_rows;
            if (gw != w) { if (n) row[w] = (uint16_t)n; w = gw; n = 0; }
            ++n;
        });
        if (n) row[w] = (uint16_t)n;
    }
}

// Segment plan, one thread per streaming task.  Segment w holds, for every owner slot, the owner's pairs of
// window w that segment w-1 did not take, followed by up to `slack` pairs of window w+1 (look-ahead fill), where
// slack = segment length - own pairs and the segment length is the largest own-pair count of the 32 owners.
//   nsteps[t][w] super-steps of the segment, take[pos][w] pairs of window w+1 that owner `pos` serves in segment w,
//   wide[t][w]   nsteps rounded up to a multiple of `depth` (stream positions)
__global__ void k_sg_plan(const uint16_t* __restrict__ cnt, int64_t n_active, int64_t n_task_stream, int n_win, int look,
                          int depth, uint16_t* __restrict__ nsteps, uint16_t* __restrict__ take, int64_t* __restrict__ wide) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t <= n_task_stream; t += (int64_t)gridDim.x * blockDim.x) {
        if (t == n_task_stream) { wide[t * n_win] = 0; continue; }
        uint16_t carry[VB_SEG_OWNERS];
        for (int sl = 0; sl < VB_SEG_OWNERS; ++sl) carry[sl] = 0;
        for (int w = 0; w < n_win; ++w) {
            int n = 0;
            for (int sl = 0; sl < VB_SEG_OWNERS; ++sl) {
                const int64_t pos = t * VB_SEG_OWNERS + sl;
                if (pos < n_active) { const int own = (int)cnt[(size_t)pos * n_win + w] - (int)carry[sl]; n = own > n ? own : n; }
            }
            for (int sl = 0; sl < VB_SEG_OWNERS; ++sl) {
                const int64_t pos = t * VB_SEG_OWNERS + sl;
                if (pos >= n_active) continue;
                const int own = (int)cnt[(size_t)pos * n_win + w] - (int)carry[sl];
                int tk = 0;
                if (look && w + 1 < n_win) {
                    const int nxt = cnt[(size_t)pos * n_win + w + 1];
                    tk = n - own < nxt ? n - own : nxt;
                }
                take[(size_t)pos * n_win + w] = (uint16_t)tk;
                carry[sl] = (uint16_t)tk;
            }
            nsteps[t * n_win + w] = (uint16_t)n;
            wide[t * n_win + w] = (int64_t)((n + depth - 1) / depth) * depth;
        }
    }
}

__global__ void k_sg_widen(const uint32_t* __restrict__ in, int64_t n, int64_t* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i <= n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = i < n ? (int64_t)in[i] : 0;
}

// write the records and the residual CSR; one thread per sorted position.  Within window v the owner's first
// take[v-1] pairs belong to segment v-1 (row offsets relative to window v-1, i.e. >= win_rows), the rest to segment v.
// The order of an owner's records inside a segment is free.  With 64-byte table rows (`pair_dist` > 0) two owner
// slots share one shared-memory wavefront (slots s and s + pair_dist of a super-step): rows of equal parity hit the
// same 16 banks.  Slot s therefore lists its even rows from the front and its odd rows from the back of the segment,
// slot s + pair_dist the other way round, so that most steps pair an even with an odd row.
template <int ORI, bool WIDE>
__global__ void k_sg_fill(const CountsView m, int64_t n_owner, int64_t n_active, const int32_t* __restrict__ perm, int win_rows,
                          int n_win, int pair_dist, const int64_t* __restrict__ step_off, const uint16_t* __restrict__ cnt,
                          const uint16_t* __restrict__ take, uint16_t* __restrict__ rec,
                          const int64_t* __restrict__ hptr, int32_t* __restrict__ hrow, uint32_t* __restrict__ hcnt) {
    for (int64_t pos = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; pos < n_owner; pos += (int64_t)gridDim.x * blockDim.x) {
        const int64_t o = perm[pos];
        const bool streams = pos < n_active;
        const int64_t t = pos / VB_SEG_OWNERS;
        const int slot = (int)(pos % VB_SEG_OWNERS);
        const uint32_t front_parity = (pair_dist > 0 && (slot & pair_dist)) ? 1u : 0u;    // row parity listed from the front
        const uint16_t* crow = cnt + (size_t)pos * n_win;
        const uint16_t* trow = take + (size_t)pos * n_win;
        int64_t h = hptr[o];
        int w = -1;
        int k = 0;                 // index of the pair inside its window
        int early = 0;             // pairs of window w served by segment w-1
        // per segment: first super-step, records of this owner, cursors from the front and from the back
        int64_t base_prev = 0, base_cur = 0;
        int tot_prev = 0, tot_cur = 0, f_prev = 0, b_prev = 0, f_cur = 0, b_cur = 0;
        for_records<ORI, WIDE>(m, o, [&](int g, uint32_t c) {
            if (c > VB_SEG_MAX_COUNT || !streams) { hrow[h] = g; hcnt[h] = c; ++h; return; }
            const int gw = g / win_rows;
            if (gw != w) {
                if (gw == w + 1 && w >= 0) { base_prev = base_cur; tot_prev = tot_cur; f_prev = f_cur; b_prev = b_cur; }
                else if (gw > 0) {      // the segment before gw holds only early pairs of this window (none of its own)
                    base_prev = step_off[t * n_win + gw - 1];
                    tot_prev = (int)crow[gw - 1] - (gw > 1 ? (int)trow[gw - 2] : 0) + (int)trow[gw - 1];
                    f_prev = b_prev = 0;
                }
                w = gw; k = 0;
                early = w > 0 ? trow[w - 1] : 0;
                base_cur = step_off[t * n_win + w];
                tot_cur = (int)crow[w] - early + (int)trow[w];
                f_cur = b_cur = 0;
            }
            int64_t s;
            uint32_t rel;
            if (k < early) {
                rel = (uint32_t)(g - (gw - 1) * win_rows);
                s = pair_dist == 0 ? base_prev + f_prev++ : ((rel & 1u) == front_parity ? base_prev + f_prev++ : base_prev + tot_prev - 1 - b_prev++);
            } else {
                rel = (uint32_t)(g - gw * win_rows);
                s = pair_dist == 0 ? base_cur + f_cur++ : ((rel & 1u) == front_parity ? base_cur + f_cur++ : base_cur + tot_cur - 1 - b_cur++);
            }
            rec[(size_t)s * VB_SEG_OWNERS + slot] = (uint16_t)((rel << VB_SEG_CNT_BITS) | c);
            ++k;
        });
    }
}

// ---------------------------------------------------------------------------------------------
// host: build one orientation
// ---------------------------------------------------------------------------------------------
static void seg_set_free(SegSet& g) {
    cudaFree(g.perm); cudaFree(g.nsteps); cudaFree(g.task_off); cudaFree(g.rec);
    cudaFree(g.hptr); cudaFree(g.hrow); cudaFree(g.hcnt);
    memset(&g, 0, sizeof(g));
}

static CountsView sg_view(const vb_counts* m) {
    CountsView v;
    v.C = m->C; v.V = m->V; v.N = m->N;
    v.cell_ptr = m->cell_ptr; v.cell_idx = m->cell_idx; v.cell_cnt = m->cell_cnt; v.cell_dp = m->cell_dp;
    v.snp_ptr = m->snp_ptr; v.snp_idx = m->snp_idx; v.snp_cnt = m->snp_cnt; v.snp_dp = m->snp_dp;
    return v;
}

static int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

// window geometry per table precision: rows per window and resident windows (shared memory: nb * rows * row bytes).
// Look-ahead fill needs the windows w and w+1 resident while w+2 loads: nb >= 3.
static void seg_window(int prec, int* win_rows, int* nb, int* look) {
    if (prec == 0) { *win_rows = env_int("VIREO_B200_SEG_WR64", 512); *nb = env_int("VIREO_B200_SEG_NB64", 3); }
    else { *win_rows = env_int("VIREO_B200_SEG_WR32", 1024); *nb = env_int("VIREO_B200_SEG_NB32", 3); }     // 64-byte rows
    *look = env_int("VIREO_B200_SEG_LOOK", 1) ? 1 : 0;
    if (*win_rows < 32) *win_rows = 32;
    if (*win_rows > VB_SEG_MAX_WIN_ROWS / 2) *win_rows = VB_SEG_MAX_WIN_ROWS / 2;     // 11-bit row offsets span two windows
    *win_rows &= ~7;
    if (*nb < 2) *nb = 2;
    const int row_bytes = prec == 0 ? 128 : 64;
    while ((size_t)*nb * *win_rows * row_bytes > 200 * 1024 && *nb > 3) --*nb;
    while ((size_t)*nb * *win_rows * row_bytes > 200 * 1024) *win_rows -= 8;
    if (*nb < 3) *look = 0;
}

static int seg_depth(int prec) {
    const int d = env_int(prec == 0 ? "VIREO_B200_SEG_DEPTH64" : "VIREO_B200_SEG_DEPTH32", prec == 0 ? VB_SEG_DEPTH64 : VB_SEG_DEPTH32);
    return (d >= 8 && prec != 2) ? 8 : 4;
}

template <int ORI>
static int seg_build_one(vb_counts* m, SegSet& g, int prec, cudaStream_t st) {
    memset(&g, 0, sizeof(g));
    const int sm = m->sm_count;
    const int64_t O = ORI == 0 ? m->C : 2 * m->V;
    const int64_t Gn = ORI == 0 ? 2 * m->V : m->C;
    if (O >= (1ll << 31) - 64 || Gn >= (1ll << 31) - 4096) { vb_set_error("segment format: more than 2^31 rows"); return VB_E_UNSUPPORTED; }
    int win_rows, nb, look;
    seg_window(prec, &win_rows, &nb, &look);
    const int depth = seg_depth(prec);
    int n_win = (int)((Gn + win_rows - 1) / win_rows);
    if (n_win < 1) n_win = 1;
    const int64_t n_task = (O + VB_SEG_OWNERS - 1) / VB_SEG_OWNERS;
    const CountsView v = sg_view(m);
    GsScratch tmp;
    int rc;
    uint32_t *nl, *nh, *rd, *nl_sorted;
    int32_t *ids, *perm_sorted;
    unsigned int* flags;
    unsigned long long* sums;
    uint8_t* sparse;
    if ((rc = tmp.alloc(&nl, O)) || (rc = tmp.alloc(&nh, O + 1)) || (rc = tmp.alloc(&rd, O)) || (rc = tmp.alloc(&nl_sorted, O)) ||
        (rc = tmp.alloc(&ids, O)) || (rc = tmp.alloc(&perm_sorted, O)) || (rc = tmp.alloc(&flags, 4)) ||
        (rc = tmp.alloc(&sums, 4)) || (rc = tmp.alloc(&sparse, O)))
        return rc;
    VB_CUDA(cudaMemsetAsync(flags, 0, 4 * sizeof(unsigned int), st));
    VB_CUDA(cudaMemsetAsync(sums, 0, 4 * sizeof(unsigned long long), st));
    VB_CUDA(cudaMemsetAsync(sparse, 0, O ? O : 1, st));
    unsigned int hflags[4] = {0, 0, 0, 0};
    unsigned long long hsums[4] = {0, 0, 0, 0};
    std::vector<uint32_t> hlen((size_t)O, 0u);
    int64_t n_active = 0;
    if (O) {
        if (m->wide) k_sg_count<ORI, true><<<grid1d(O, sm), 256, 0, st>>>(v, O, nl, nh, rd, flags);
        else k_sg_count<ORI, false><<<grid1d(O, sm), 256, 0, st>>>(v, O, nl, nh, rd, flags);
        VB_CUDA(cudaGetLastError());
        k_sg_iota<<<grid1d(O, sm), 256, 0, st>>>(ids, O);
        VB_CUDA(cudaGetLastError());
        size_t tb = 0;
        VB_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, tb, nl, nl_sorted, ids, perm_sorted, O, 0, 32, st));
        void* cub_tmp;
        if ((rc = tmp.alloc((unsigned char**)&cub_tmp, tb))) return rc;
        VB_CUDA(cub::DeviceRadixSort::SortPairsDescending(cub_tmp, tb, nl, nl_sorted, ids, perm_sorted, O, 0, 32, st));
        VB_CUDA(cudaMemcpyAsync(hflags, flags, sizeof(hflags), cudaMemcpyDeviceToHost, st));
        VB_CUDA(cudaMemcpyAsync(hlen.data(), nl_sorted, O * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        VB_CUDA(cudaStreamSynchronize(st));
        if (hflags[0]) { vb_set_error("segment format: an entry has AD > DP"); return VB_E_VALUE; }
        // Owners with very few pairs would pad a whole task's super-steps for one or two records; when such owners
        // carry a negligible share of the pairs (e.g. alternative-allele rows of homozygous-reference SNPs) they
        // are served by the residual kernel instead.
        double total = 0.0;
        for (int64_t r = 0; r < O; ++r) total += hlen[r];
        int64_t first_sparse = O, moved = 0;
        const int64_t budget = (int64_t)(total / 50.0);
        while (first_sparse > 0 && hlen[first_sparse - 1] <= VB_SPARSE_LEN && moved + hlen[first_sparse - 1] <= budget) {
            moved += hlen[first_sparse - 1];
            --first_sparse;
        }
        if (moved > 0) {
            k_sg_mark_sparse<<<grid1d(O - first_sparse, sm), 256, 0, st>>>(perm_sorted, first_sparse, O, sparse, nl, nh, rd);
            VB_CUDA(cudaGetLastError());
            for (int64_t r = first_sparse; r < O; ++r) hlen[r] = 0;
        }
        n_active = first_sparse;
        while (n_active > 0 && hlen[n_active - 1] == 0) --n_active;
        k_sg_sums<<<grid1d(O, sm), 256, 0, st>>>(nl, nh, rd, O, sums);
        VB_CUDA(cudaGetLastError());
        VB_CUDA(cudaMemcpyAsync(hsums, sums, sizeof(hsums), cudaMemcpyDeviceToHost, st));
        VB_CUDA(cudaStreamSynchronize(st));
    }
    g.n_owner = O; g.n_gather = Gn; g.n_task = n_task; g.n_win = n_win; g.win_rows = win_rows; g.nb = nb; g.look = look; g.depth = depth;
    g.n_light = (int64_t)hsums[0]; g.n_heavy = (int64_t)hsums[1]; g.max_reads = (int64_t)hsums[2];
    g.n_task_stream = (n_active + VB_SEG_OWNERS - 1) / VB_SEG_OWNERS;
    const int64_t nts = g.n_task_stream;

    VB_CUDA(cudaMalloc(&g.perm, (n_task ? n_task : 1) * VB_SEG_OWNERS * sizeof(int32_t)));
    VB_CUDA(cudaMemsetAsync(g.perm, 0xff, (n_task ? n_task : 1) * VB_SEG_OWNERS * sizeof(int32_t), st));
    if (O) VB_CUDA(cudaMemcpyAsync(g.perm, perm_sorted, O * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));

    // super-steps per (task, window) and their offsets
    const int64_t n_tw = nts * n_win;
    uint16_t *cnt_ow, *take;
    int64_t *wide, *step_off;
    const size_t n_cw = (size_t)(n_active ? n_active : 1) * n_win;
    if ((rc = tmp.alloc(&cnt_ow, n_cw)) || (rc = tmp.alloc(&take, n_cw)) || (rc = tmp.alloc(&wide, n_tw + 1)) ||
        (rc = tmp.alloc(&step_off, n_tw + 1)))
        return rc;
    VB_CUDA(cudaMalloc(&g.nsteps, (n_tw ? n_tw : 1) * sizeof(uint16_t)));
    VB_CUDA(cudaMalloc(&g.task_off, (nts + 1) * sizeof(int64_t)));
    VB_CUDA(cudaMemsetAsync(cnt_ow, 0, n_cw * sizeof(uint16_t), st));
    VB_CUDA(cudaMemsetAsync(take, 0, n_cw * sizeof(uint16_t), st));
    if (n_active) {
        if (m->wide) k_sg_wincount<ORI, true><<<grid1d(n_active, sm), 256, 0, st>>>(v, n_active, g.perm, win_rows, n_win, cnt_ow);
        else k_sg_wincount<ORI, false><<<grid1d(n_active, sm), 256, 0, st>>>(v, n_active, g.perm, win_rows, n_win, cnt_ow);
        VB_CUDA(cudaGetLastError());
    }
    k_sg_plan<<<grid1d(nts + 1, sm), 64, 0, st>>>(cnt_ow, n_active, nts, n_win, look, depth, g.nsteps, take, wide);
    VB_CUDA(cudaGetLastError());
    {
        size_t tb = 0;
        VB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, wide, step_off, n_tw + 1, st));
        void* cub_tmp;
        if ((rc = tmp.alloc((unsigned char**)&cub_tmp, tb))) return rc;
        VB_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp, tb, wide, step_off, n_tw + 1, st));
    }
    int64_t total_steps = 0;
    VB_CUDA(cudaMemcpyAsync(&total_steps, step_off + n_tw, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaMemcpy2DAsync(g.task_off, sizeof(int64_t), step_off, (size_t)n_win * sizeof(int64_t), sizeof(int64_t),
                              (size_t)(nts + 1), cudaMemcpyDeviceToDevice, st));
    VB_CUDA(cudaStreamSynchronize(st));
    g.n_step = total_steps;

    // residual CSR
    VB_CUDA(cudaMalloc(&g.hptr, (O + 1) * sizeof(int64_t)));
    VB_CUDA(cudaMalloc(&g.hrow, (g.n_heavy ? g.n_heavy : 1) * sizeof(int32_t)));
    VB_CUDA(cudaMalloc(&g.hcnt, (g.n_heavy ? g.n_heavy : 1) * sizeof(uint32_t)));
    {
        int64_t* hw;
        if ((rc = tmp.alloc(&hw, O + 1))) return rc;
        k_sg_widen<<<grid1d(O + 1, sm), 256, 0, st>>>(nh, O, hw);
        VB_CUDA(cudaGetLastError());
        size_t tb = 0;
        VB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, hw, g.hptr, O + 1, st));
        void* cub_tmp;
        if ((rc = tmp.alloc((unsigned char**)&cub_tmp, tb))) return rc;
        VB_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp, tb, hw, g.hptr, O + 1, st));
    }
    // slack: the register queue runs `depth` super-steps ahead, the L2 prefetch VB_SEG_L2_AHEAD bytes
    const size_t n_rec = ((size_t)total_steps + 2 * depth) * VB_SEG_OWNERS + VB_SEG_L2_AHEAD / 2 + 64;
    VB_CUDA(cudaMalloc(&g.rec, n_rec * sizeof(uint16_t)));
    VB_CUDA(cudaMemsetAsync(g.rec, 0, n_rec * sizeof(uint16_t), st));
    if (O) {
        if (m->wide) k_sg_fill<ORI, true><<<grid1d(O, sm), 256, 0, st>>>(v, O, n_active, g.perm, win_rows, n_win, prec != 0 ? 4 : 0, step_off, cnt_ow, take, g.rec, g.hptr, g.hrow, g.hcnt);
        else k_sg_fill<ORI, false><<<grid1d(O, sm), 256, 0, st>>>(v, O, n_active, g.perm, win_rows, n_win, prec != 0 ? 4 : 0, step_off, cnt_ow, take, g.rec, g.hptr, g.hrow, g.hcnt);
        VB_CUDA(cudaGetLastError());
    }
    VB_CUDA(cudaStreamSynchronize(st));

    // launch geometry: one CTA per SM, as few consumer warps as cover the streaming tasks
    int nw = (int)((nts + sm - 1) / sm);
    if (nw < 1) nw = 1;
    if (nw > VB_SEG_MAX_WARPS) nw = VB_SEG_MAX_WARPS;
    int64_t grid = (nts + nw - 1) / nw;
    if (grid < 1) grid = 1;
    if (grid > 65535 * 16) { vb_set_error("segment format: too many owner rows"); return VB_E_UNSUPPORTED; }
    g.grid = (int)grid; g.nwarps = nw;
    g.bytes = (int64_t)n_rec * 2 + n_task * VB_SEG_OWNERS * 4 + n_tw * 2 + (nts + 1) * 8 + (O + 1) * 8 + g.n_heavy * 8;
    g.built = 1;
    return VB_OK;
}

int vb_seg_build(vb_counts* m, int prec, cudaStream_t st) {
    if (prec < 0 || prec > 2) { vb_set_error("bad table kind"); return VB_E_ARG; }
    if (m->sA[prec].built && m->sB[prec].built) return VB_OK;
    if (m->seg_failed[prec]) return VB_E_UNSUPPORTED;
    DeviceGuard dg(m->device);
    int rc = seg_build_one<0>(m, m->sA[prec], prec, st);
    if (!rc) rc = seg_build_one<1>(m, m->sB[prec], prec, st);
    if (rc) {
        seg_set_free(m->sA[prec]);
        seg_set_free(m->sB[prec]);
        m->seg_failed[prec] = 1;
        cudaGetLastError();
    }
    return rc;
}

void vb_seg_free(vb_counts* m) {
    for (int i = 0; i < 3; ++i) { seg_set_free(m->sA[i]); seg_set_free(m->sB[i]); }
}

void vb_seg_geometry(const SegSet& g, int* grid, int* nwarps) {
    *grid = g.grid > 0 ? g.grid : 1;
    *nwarps = g.nwarps > 0 ? g.nwarps : 1;
}

// ---------------------------------------------------------------------------------------------
// residual pairs: one warp per owner, table rows gathered from L2 (always the FP64 table).
// out[o][0:16] is written for EVERY owner (zeros where there is no residual).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(VB_THREADS)
k_seg_heavy(int64_t n_owner, const int64_t* __restrict__ hptr, const int32_t* __restrict__ hrow, const uint32_t* __restrict__ hcnt,
            const double* __restrict__ table, int64_t table_stride, int RW, double* __restrict__ out, int64_t out_stride,
            const int* __restrict__ ctrl) {
    const int b = blockIdx.y;
    if (ctrl && ctrl[b * VB_CTRL_N]) return;
    const double* __restrict__ T = table + (size_t)b * table_stride;
    double* __restrict__ O = out + (size_t)b * out_stride;
    const int lane = threadIdx.x & 31, kl = lane % RW, sub = lane / RW, nsub = 32 / RW;    // RW = 16 or 8
    const int64_t nw = (int64_t)gridDim.x * VB_WARPS;
    for (int64_t o = (int64_t)blockIdx.x * VB_WARPS + (threadIdx.x >> 5); o < n_owner; o += nw) {
        const int64_t p0 = hptr[o], p1 = hptr[o + 1];
        double acc = 0.0;
        for (int64_t q = p0 + sub; q < p1; q += nsub)
            acc = fma((double)hcnt[q], T[(size_t)hrow[q] * RW + kl], acc);
        for (int off = RW; off < 32; off <<= 1) acc += __shfl_xor_sync(VB_FULL, acc, off);
        if (sub == 0) O[(size_t)o * RW + kl] = acc;
    }
}

// replicate [n_row, K] into the rows (RW doubles each) of a gather table: column c holds source column c % KT
__global__ void __launch_bounds__(VB_THREADS)
k_pad_rows(const double* __restrict__ src, int64_t n_row, int K, int KT, int RW, double* __restrict__ dst) {
    const int b = blockIdx.y;
    const double* __restrict__ S = src + (size_t)b * n_row * K;
    double* __restrict__ D = dst + (size_t)b * n_row * RW;
    const int64_t n = n_row * RW;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / RW;
        const int k = (int)(e % RW) % KT;
        D[e] = k < K ? S[r * K + k] : 0.0;
    }
}

int vb_pad_rows_launch(const vb_counts* m, const double* src, int64_t n_row, int K, int KT, int RW, int B, double* dst,
                       cudaStream_t st) {
    int64_t nb = (n_row * RW + VB_THREADS - 1) / VB_THREADS;
    if (nb > (int64_t)m->sm_count * 8) nb = (int64_t)m->sm_count * 8;
    if (nb < 1) nb = 1;
    VB_LAUNCH(7, st, k_pad_rows<<<dim3((unsigned)nb, B), VB_THREADS, 0, st>>>(src, n_row, K, KT, RW, dst));
    VB_CUDA(cudaGetLastError());
    return VB_OK;
}

// ID_prob rows [B*C, 16] -> unsigned 32-bit fixed point, value * (2^32 - 1): a posterior of exactly 1.0 (most
// cells once the fit has settled) is represented exactly, so the per-SNP sums carry no systematic bias
__device__ __forceinline__ uint32_t quant_unit(double r) {
    const double s = r * 4294967295.0;
    return s >= 4294967295.0 ? 0xffffffffu : (s > 0.0 ? (uint32_t)__double2ull_rn(s) : 0u);
}

__global__ void __launch_bounds__(VB_THREADS) k_seg_quant_rows(const double* __restrict__ src, int64_t n, uint32_t* __restrict__ dst) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
        dst[e] = quant_unit(src[e]);
}

// ---------------------------------------------------------------------------------------------
// k_seg
// ---------------------------------------------------------------------------------------------
#define VB_SG_RED_DOUBLES (2 * VB_MAX_GT)
#define VB_SG_PIECE 16384u          // bytes per bulk copy

struct SegArgs {
    int mode;            // GM_CELL, GM_CELL_LL, GM_SNP
    int theta_mode;      // GM_SNP: 0 never, 1 always, 2 per the device iteration counter
    int nwarps;          // consumer warps per CTA (block size = (nwarps + 1) * 32)
    int nb;              // resident windows
    int has_heavy;       // add p.H[owner] before the epilogue
    int wait_hint_ns;    // > 0: suspend-time hint of the window waits
    int64_t table_stride;    // bytes per restart of the gather table
    const unsigned char* table;
    // GM_PLAIN: the sums of columns [0, plain_cols) go to plain_out[owner * plain_ld + plain_off + column]
    double* plain_out;
    int64_t plain_ld;
    int plain_off, plain_cols;
};

template <int PREC> struct SegCfg;
template <> struct SegCfg<0> { static constexpr int LPO = 4, NC = 4, ROWB = 128; };
template <> struct SegCfg<1> { static constexpr int LPO = 4, NC = 4, ROWB = 64; };
template <> struct SegCfg<2> { static constexpr int LPO = 4, NC = 2, ROWB = 64; };     // FP64, 8 columns

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// One owner slot of one super-step.  `w` holds two 16-bit records; HI selects the upper one.  The row offset
// (bits 15..5 of the record) times the row size gives the byte offset from the start of window w's buffer; a
// record of the look-ahead window lands in the next buffer, which is contiguous unless the ring wraps (WRAP).
// FP64 tables: 4 lanes x 4 columns, two 16-byte loads per lane.
struct SegScratch64 { double v[2][4]; };     // landing registers of the FP64 loads, two sets in flight
struct SegScratch32 {};
struct SegScratchN { double v[2][2]; };         // narrow FP64 rows: one 16-byte load per lane and record

template <bool HI, bool WRAP>
__device__ __forceinline__ void seg_step(uint32_t w, uint32_t base, uint32_t wrap_at, uint32_t ring_bytes, double (&a)[4],
                                         double (&v)[4]) {
    const uint32_t c = HI ? (w & 0x001f0000u) : (w & 0x1fu);
    uint32_t t;
    asm("and.b32 %0, %1, %2;" : "=r"(t) : "r"(w), "r"(HI ? 0xffe00000u : 0xffe0u));   // opaque: keeps the scaling one LEA
    uint32_t off = HI ? __umulhi(t, 1u << 18) : t * 4u;    // row offset * 128 bytes
    if (WRAP && off >= wrap_at) off -= ring_bytes;
    const uint32_t addr = base + off;
    // `base` points at this lane's 16-byte granule in one half of the row; the other half is 64 bytes away (even
    // lane groups start in the lower half, odd groups in the upper half: the two rows that share a wavefront never
    // meet in a bank).
    // count -> double without the conversion pipe: 2^52 + x is exact, subtracting 2^52 leaves x.  The upper record of
    // a word is used in place (count << 16): odd slots carry a factor 2^16 that the epilogue removes.
    const double d = __hiloint2double(0x43300000, (int)c) - 4503599627370496.0;
    // Predicated loads, unconditional FMAs: a null record (count 0) issues no shared-memory wavefront, keeps
    // whatever finite table values its landing registers held, and adds 0 * value.  (An `if` around loads and
    // FMAs is compiled to a branch, which serialises consecutive slots on the load latency.)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.u32 p, %4, 0;\n\t"
        "@p ld.shared.v2.f64 {%0, %1}, [%5];\n\t"
        "@p ld.shared.v2.f64 {%2, %3}, [%6];\n\t}"
        : "+d"(v[0]), "+d"(v[1]), "+d"(v[2]), "+d"(v[3])
        : "r"(c), "r"(addr), "r"(addr ^ 64u));
    a[0] = fma(d, v[0], a[0]);
    a[1] = fma(d, v[1], a[1]);
    a[2] = fma(d, v[2], a[2]);
    a[3] = fma(d, v[3], a[3]);
}
// FP64 tables with 8 columns (n_donor <= 8): rows of 64 bytes, 4 lanes x 2 columns, one 16-byte load per lane.
template <bool HI, bool WRAP>
__device__ __forceinline__ void seg_step(uint32_t w, uint32_t base, uint32_t wrap_at, uint32_t ring_bytes, double (&a)[2],
                                         double (&v)[2]) {
    const uint32_t c = HI ? (w & 0x001f0000u) : (w & 0x1fu);
    uint32_t t;
    asm("and.b32 %0, %1, %2;" : "=r"(t) : "r"(w), "r"(HI ? 0xffe00000u : 0xffe0u));
    uint32_t off = HI ? __umulhi(t, 1u << 17) : t * 2u;    // row offset * 64 bytes
    if (WRAP && off >= wrap_at) off -= ring_bytes;
    const uint32_t addr = base + off;
    const double d = __hiloint2double(0x43300000, (int)c) - 4503599627370496.0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.u32 p, %2, 0;\n\t"
        "@p ld.shared.v2.f64 {%0, %1}, [%3];\n\t}"
        : "+d"(v[0]), "+d"(v[1])
        : "r"(c), "r"(addr));
    a[0] = fma(d, v[0], a[0]);
    a[1] = fma(d, v[1], a[1]);
}
// Fixed point: exact integer accumulation of count * value.  The upper record of a word is used in place
// (count << 16, row offset << 16): the accumulators of odd slots carry a factor 2^16 that the epilogue removes
// (safe while the reads of one owner stay below 2^16, checked when the format is built).
template <bool HI, bool WRAP>
__device__ __forceinline__ void seg_step(uint32_t w, uint32_t base, uint32_t wrap_at, uint32_t ring_bytes, unsigned long long (&a)[4]) {
    const uint32_t c = HI ? (w & 0x001f0000u) : (w & 0x1fu);
    // the address does not depend on the null test: keeping it outside shortens the dependent chain in front of the load
    uint32_t t;
    asm("and.b32 %0, %1, %2;" : "=r"(t) : "r"(w), "r"(HI ? 0xffe00000u : 0xffe0u));
    uint32_t off = HI ? __umulhi(t, 1u << 17) : t * 2u;    // row offset * 64 bytes
    if (WRAP && off >= wrap_at) off -= ring_bytes;
    const uint32_t addr = base + off;
    if (c) {
        uint32_t v0, v1, v2, v3;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(addr));
        a[0] += (unsigned long long)c * v0;
        a[1] += (unsigned long long)c * v1;
        a[2] += (unsigned long long)c * v2;
        a[3] += (unsigned long long)c * v3;
    }
}

template <bool WRAP>
__device__ __forceinline__ void seg_super_step(const uint2& u, uint32_t base, uint32_t wrap_at, uint32_t ring_bytes, double (&acc)[4][4],
                                               SegScratch64& sc) {
    seg_step<false, WRAP>(u.x, base, wrap_at, ring_bytes, acc[0], sc.v[0]); seg_step<true, WRAP>(u.x, base, wrap_at, ring_bytes, acc[1], sc.v[1]);
    seg_step<false, WRAP>(u.y, base, wrap_at, ring_bytes, acc[2], sc.v[0]); seg_step<true, WRAP>(u.y, base, wrap_at, ring_bytes, acc[3], sc.v[1]);
}
template <bool WRAP>
__device__ __forceinline__ void seg_super_step(const uint2& u, uint32_t base, uint32_t wrap_at, uint32_t ring_bytes, double (&acc)[4][2],
                                               SegScratchN& sc) {
    seg_step<false, WRAP>(u.x, base, wrap_at, ring_bytes, acc[0], sc.v[0]); seg_step<true, WRAP>(u.x, base, wrap_at, ring_bytes, acc[1], sc.v[1]);
    seg_step<false, WRAP>(u.y, base, wrap_at, ring_bytes, acc[2], sc.v[0]); seg_step<true, WRAP>(u.y, base, wrap_at, ring_bytes, acc[3], sc.v[1]);
}
template <bool WRAP>
__device__ __forceinline__ void seg_super_step(const uint2& u, uint32_t base, uint32_t wrap_at, uint32_t ring_bytes,
                                               unsigned long long (&acc)[4][4], SegScratch32&) {
    seg_step<false, WRAP>(u.x, base, wrap_at, ring_bytes, acc[0]); seg_step<true, WRAP>(u.x, base, wrap_at, ring_bytes, acc[1]);
    seg_step<false, WRAP>(u.y, base, wrap_at, ring_bytes, acc[2]); seg_step<true, WRAP>(u.y, base, wrap_at, ring_bytes, acc[3]);
}

// the super-steps of one segment; the queue q holds the next DEPTH super-steps of the stream
template <bool WRAP, int DEPTH, typename chunk_t, typename acc_t, typename scratch_t>
__device__ __forceinline__ void seg_segment(uint32_t n, chunk_t (&q)[DEPTH], const unsigned char*& sp, uint32_t base, uint32_t wrap_at,
                                            uint32_t ring_bytes, acc_t& acc, scratch_t& sc) {
    for (uint32_t s = 0; s < n; s += DEPTH) {
        // a group consumes DEPTH * 64 bytes of the warp's stream: one L2 prefetch per 128-byte line of it
#pragma unroll
        for (int l = 0; l < DEPTH * (VB_SEG_OWNERS * 2); l += 128)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(sp + VB_SEG_L2_AHEAD + l));
#pragma unroll
        for (int i = 0; i < DEPTH; ++i) {
            const chunk_t c = q[i];
            q[i] = __ldcs(reinterpret_cast<const chunk_t*>(sp + i * (VB_SEG_OWNERS * 2)));
            if (s + i < n) seg_super_step<WRAP>(c, base, wrap_at, ring_bytes, acc, sc);
        }
        sp += DEPTH * (VB_SEG_OWNERS * 2);
    }
}

// wait for a phase of a window barrier; with a suspend-time hint a waiting warp stays off the issue slots longer
__device__ __forceinline__ void seg_wait(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
    if (hint_ns == 0) { mbar_wait(bar, parity); return; }
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity), "r"(hint_ns) : "memory");
    } while (!ok);
}

template <int PREC, int DEPTH>
__global__ void __launch_bounds__((VB_SEG_MAX_WARPS + 1) * 32, 1)
k_seg(const SegView sv, const EmP p, const SegArgs sa) {
    using Cfg = SegCfg<PREC>;
    constexpr int LPO = Cfg::LPO, M = Cfg::LPO, NC = Cfg::NC, ROWB = Cfg::ROWB;
    typedef typename std::conditional<PREC == 1, unsigned long long, double>::type acc_t;
    typedef uint2 chunk_t;
    extern __shared__ __align__(1024) unsigned char smem[];
    const int b = blockIdx.y;
    if (p.ctrl && p.ctrl[b * VB_CTRL_N]) return;
    const bool do_theta = sa.mode == GM_SNP && !p.bmm && vb_theta_on(p, b, sa.theta_mode);
    if (sa.mode == GM_SNP && !p.bmm && !do_theta && !p.learn_gt && sa.theta_mode == 2) {   // S1/S2 unused this iteration
        snp_pass_tail(p, b, sa.theta_mode, false);
        return;
    }

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int NW = sa.nwarps, NB = sa.nb;
    const int grid = (int)gridDim.x, cta = (int)blockIdx.x;
    const uint32_t win_bytes = (uint32_t)sv.win_rows * ROWB;
    const uint32_t ring = smem_u32(smem);
    const uint32_t bars = ring + (uint32_t)NB * win_bytes;        // full[NB] then empty[NB]
    double* red = reinterpret_cast<double*>(smem + (size_t)NB * win_bytes + 16 * 8);

    // tasks are dealt to CTAs in boustrophedon order of the sorted list: warp ww of CTA c serves task
    // ww * grid + (ww odd ? grid - 1 - c : c), so every CTA gets the same mix of long and short tasks
    int64_t task = -1;
    int nstream = 0;
    for (int ww = 0; ww < NW; ++ww) {
        const int64_t t = (int64_t)ww * grid + ((ww & 1) ? grid - 1 - cta : cta);
        if (t < sv.n_task_stream) { ++nstream; if (ww == w) task = t; }
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < NB; ++i) { mbar_init(bars + 8 * i, 1); mbar_init(bars + 8 * (NB + i), nstream > 0 ? nstream : 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    acc_t acc[M][NC];
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[i][c] = 0;

    const int g = lane / LPO, sub = lane % LPO;
    if (w == NW) {
        // ---------------- producer warp: stream the table through the window buffers
        if (nstream > 0) {
            const unsigned char* __restrict__ T = sa.table + (size_t)b * sa.table_stride;
            int bi = 0;
            uint32_t use = 0;                               // how often buffer bi has been filled before
            for (int wd = 0; wd < sv.n_win; ++wd) {
                if (use > 0) mbar_wait(bars + 8 * (NB + bi), (use - 1) & 1);      // every consumer released the previous fill
#ifdef VB_SEG_CANARY
                // Protocol canary (build.py variants "canary" / "plainfill"): a released buffer is overwritten with NaNs
                // before it is refilled.  A consumer that read a row before its window's fill had completed, or after it
                // had released the window, would pick up a NaN, which no later step can remove from its sums -- so
                // NaN-free, bit-identical results under this build show that no such access happens.
                {
                    unsigned long long* dst = reinterpret_cast<unsigned long long*>(smem + (size_t)bi * win_bytes);
                    for (uint32_t e = lane; e < win_bytes / 8; e += 32) dst[e] = 0x7ff8dead0000beefull;
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // before the bulk copies overwrite it
                    __syncwarp();
                }
#endif
#ifdef VB_SEG_PLAIN_FILL
                // Sanitizer variant (build.py variant "plainfill"): the same full/empty protocol with the window filled
                // by ordinary loads and stores of the producer warp instead of cp.async.bulk, so that racecheck --
                // which follows generic-proxy accesses and mbarrier arrive/wait, but not the completion of
                // async-proxy bulk copies -- can check the protocol itself.  Results are bit-identical.
                {
                    const int64_t r0 = (int64_t)wd * sv.win_rows;
                    const int64_t rows = sv.n_gather - r0 < sv.win_rows ? sv.n_gather - r0 : sv.win_rows;
                    const uint32_t n16 = (uint32_t)(rows * ROWB / 16);
                    const uint4* src = reinterpret_cast<const uint4*>(T + (size_t)r0 * ROWB);
                    uint4* dst = reinterpret_cast<uint4*>(smem + (size_t)bi * win_bytes);
                    for (uint32_t e = lane; e < n16; e += 32) dst[e] = __ldg(src + e);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bars + 8 * bi);
                }
#else
                if (lane == 0) {
                    const int64_t r0 = (int64_t)wd * sv.win_rows;
                    const int64_t rows = sv.n_gather - r0 < sv.win_rows ? sv.n_gather - r0 : sv.win_rows;
                    const uint32_t bytes = (uint32_t)(rows * ROWB);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_expect_tx(bars + 8 * bi, bytes);
                    const unsigned char* src = T + (size_t)r0 * ROWB;
                    const uint32_t dst = ring + (uint32_t)bi * win_bytes;
                    for (uint32_t off = 0; off < bytes; off += VB_SG_PIECE) {
                        const uint32_t sz = bytes - off < VB_SG_PIECE ? bytes - off : VB_SG_PIECE;
                        bulk_g2s(dst + off, src + off, sz, bars + 8 * bi);
                    }
                }
#endif
                __syncwarp();
                if (++bi == NB) { bi = 0; ++use; }
            }
        }
    } else if (task >= 0) {
        // ---------------- consumer warp: 32 owner slots, LPO lanes per slot
        const unsigned char* sp = reinterpret_cast<const unsigned char*>(sv.rec) + (size_t)sv.task_off[task] * (VB_SEG_OWNERS * 2) + g * (M * 2);
        typename std::conditional<PREC == 0, SegScratch64, typename std::conditional<PREC == 1, SegScratch32, SegScratchN>::type>::type sc;
        if constexpr (PREC == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) sc.v[i >> 2][i & 3] = 0.0;
        }
        if constexpr (PREC == 2) {
#pragma unroll
            for (int i = 0; i < 4; ++i) sc.v[i >> 1][i & 1] = 0.0;
        }
        chunk_t q[DEPTH];
#pragma unroll
        for (int i = 0; i < DEPTH; ++i) q[i] = __ldcs(reinterpret_cast<const chunk_t*>(sp + i * (VB_SEG_OWNERS * 2)));
        sp += DEPTH * (VB_SEG_OWNERS * 2);
        const uint16_t* __restrict__ ns = sv.nsteps + (size_t)task * sv.n_win;
        uint32_t n_next = ns[0];
        int bi = 0;                          // buffer of window wd
        uint32_t phase = 0;                  // its fill parity
        // FP64 rows: odd lane groups read the upper half of a row first (see seg_step)
        const uint32_t lane_base = ring + (uint32_t)sub * 16 + (PREC == 0 ? (uint32_t)(g & 1) * 64 : 0u);
        const uint32_t ring_bytes = (uint32_t)NB * win_bytes;
        const int look = sv.look;
        const uint32_t hint = (uint32_t)sa.wait_hint_ns;
        if (look) seg_wait(bars, 0, hint);   // window 0; every segment then waits for the window after its own
        for (int wd = 0; wd < sv.n_win; ++wd) {
            const uint32_t n = n_next;
            if (wd + 1 < sv.n_win) n_next = ns[wd + 1];
            const bool last_buf = bi + 1 == NB;
            if (look) {
                if (wd + 1 < sv.n_win) seg_wait(bars + 8 * (last_buf ? 0 : bi + 1), last_buf ? phase ^ 1 : phase, hint);
            } else {
                seg_wait(bars + 8 * bi, phase, hint);
            }
            const uint32_t base = lane_base + (uint32_t)bi * win_bytes;
            if (look && last_buf) seg_segment<true, DEPTH>(n, q, sp, base, win_bytes, ring_bytes, acc, sc);
            else seg_segment<false, DEPTH>(n, q, sp, base, win_bytes, ring_bytes, acc, sc);
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + 8 * (NB + bi));
            if (last_buf) { bi = 0; phase ^= 1; } else ++bi;
        }
    }

    // ---------------- epilogue: the LPO lanes of a slot hold NC columns each
    const int K = p.K, KT = p.KT;
    double r0 = 0.0, r1 = 0.0;                   // GM_CELL*: LB_p, KL_ID partial sums
    double t1[VB_MAX_GT], t2[VB_MAX_GT];         // GM_SNP: theta partial sums (alternative / reference allele rows)
#pragma unroll
    for (int gq = 0; gq < VB_MAX_GT; ++gq) t1[gq] = t2[gq] = 0.0;
    // columns of this lane: [col0, col0 + 1] and [col0 + col2, col0 + col2 + 1]
    //   fixed point: four consecutive columns;  FP64: one granule in each half of the row, the first from the half
    //   this lane group reads first
    //   narrow FP64: two consecutive columns of 8
    const int col0 = PREC == 0 ? 2 * sub + 8 * (g & 1) : (PREC == 1 ? 4 * sub : 2 * sub);
    const int col2 = PREC == 0 ? ((g & 1) ? -8 : 8) : 2;
    const int RW = PREC == 2 ? 8 : VB_ROW_DOUBLES;
    const double unq = PREC != 1 ? 1.0 : (sa.mode == GM_SNP ? 1.0 / 4294967295.0 : -1.0 / p.qscale[b]);
    const double unq_hi = unq / 65536.0;     // odd slots accumulate count << 16 (exact power of two in either precision)

    // the streaming task first, then a share of the tasks without records (accumulators are zero for those)
    int64_t et = task;
    int64_t extra = sv.n_task_stream + (int64_t)cta * NW + w;
    if (w >= NW) { et = -1; extra = sv.n_task; }
    for (;;) {
        if (et < 0) {
            if (extra >= sv.n_task) break;
            et = extra;
            extra += (int64_t)grid * NW;
        }
#pragma unroll
        for (int mi = 0; mi < M; ++mi) {
            const int owner = sv.perm[et * VB_SEG_OWNERS + g * M + mi];
            double v[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) v[c] = (double)acc[mi][c] * ((mi & 1) ? unq_hi : unq);
            bool valid[NC], primary[NC];
            int kk[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const int col = col0 + (c < 2 ? c : col2 + c - 2);
                kk[c] = col % KT;
                valid[c] = kk[c] < K && owner >= 0;
                primary[c] = col < KT && valid[c];
            }
            if (sa.has_heavy && owner >= 0) {
                const double* __restrict__ H = p.H + ((size_t)b * sv.n_owner + owner) * RW + col0;
#pragma unroll
                for (int c = 0; c < NC; ++c) v[c] += H[c < 2 ? c : col2 + c - 2];
            }
            if (sa.mode == GM_CELL || sa.mode == GM_CELL_LL) {
                const int64_t j = owner >= 0 ? owner : 0;
                const size_t prow = (size_t)(p.id_rows == 1 ? 0 : j) * K;
                double* __restrict__ R = p.R + ((size_t)b * p.C + j) * K;
                double* __restrict__ LL = p.ll + ((size_t)b * p.C + j) * K;
                double pr[NC];
                if (sa.mode == GM_CELL) {
                    double mx = -INFINITY;
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        pr[c] = valid[c] ? v[c] + p.lidp[prow + kk[c]] : -INFINITY;
                        mx = fmax(mx, pr[c]);
                    }
#pragma unroll
                    for (int off = 1; off < LPO; off <<= 1) mx = fmax(mx, __shfl_xor_sync(VB_FULL, mx, off));
                    double z = 0.0;
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        pr[c] = pr[c] == -INFINITY ? 0.0 : exp(pr[c] - mx);
                        if (primary[c]) z += pr[c];
                    }
#pragma unroll
                    for (int off = 1; off < LPO; off <<= 1) z += __shfl_xor_sync(VB_FULL, z, off);
#pragma unroll
                    for (int c = 0; c < NC; ++c) pr[c] = pr[c] / z;
                    if (owner >= 0) {
                        // 128-byte-row copy (columns replicated 16/KT times) for the SNP pass
                        double* __restrict__ RP = p.RP + ((size_t)b * p.C + j) * RW + col0;
                        *reinterpret_cast<double2*>(RP) = make_double2(pr[0], pr[1]);
                        if constexpr (NC == 4) *reinterpret_cast<double2*>(RP + col2) = make_double2(pr[NC - 2], pr[NC - 1]);
                        if constexpr (PREC == 1) {
                            uint32_t* __restrict__ RQ = p.RPq + ((size_t)b * p.C + j) * VB_ROW_DOUBLES + col0;
                            *reinterpret_cast<uint4*>(RQ) = make_uint4(quant_unit(pr[0]), quant_unit(pr[1]), quant_unit(pr[NC - 2]), quant_unit(pr[NC - 1]));
                        }
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < NC; ++c) pr[c] = primary[c] ? R[kk[c]] : 0.0;
                }
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    if (primary[c]) {
                        const double a = v[c], pv = pr[c];
                        LL[kk[c]] = a;
                        if (sa.mode == GM_CELL) R[kk[c]] = pv;
                        r0 += a * pv;
                        if (pv > 0.0) r1 += pv * (log(pv) - p.lidp_kl[prow + kk[c]]);
                    }
                }
            } else if (sa.mode == GM_PLAIN) {
                // column chunk of a wider table (doublet pass): plain sums, no softmax
                if (owner >= 0) {
                    double* __restrict__ O = sa.plain_out + (size_t)owner * sa.plain_ld + sa.plain_off;
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        const int col = col0 + (c < 2 ? c : col2 + c - 2);
                        if (col < sa.plain_cols) O[col] = v[c];
                    }
                }
            } else {   // GM_SNP
                const int64_t i = owner >= 0 ? owner >> 1 : 0;
                const int al = owner & 1;
                const int G = p.G;
                double* __restrict__ S = (al ? p.S1 : p.S2) + ((size_t)b * p.V + i) * K;
                const double* __restrict__ GT = (do_theta && p.GT) ? p.GT + ((size_t)b * p.V + i) * K * G : nullptr;
                double tt[VB_MAX_GT];
#pragma unroll
                for (int gq = 0; gq < VB_MAX_GT; ++gq) tt[gq] = 0.0;
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    if (primary[c]) {
                        S[kk[c]] = v[c];
                        if (GT) {
#pragma unroll
                            for (int gq = 0; gq < VB_MAX_GT; ++gq)
                                if (gq < G) tt[gq] += v[c] * GT[(size_t)kk[c] * G + gq];
                        }
                    }
                }
                if (do_theta && p.ase) {
                    // theta per SNP (vireo_model.py:177 `axis=1`): the raw sums are parked in the ab rows
#pragma unroll
                    for (int gq = 0; gq < VB_MAX_GT; ++gq) {
                        if (gq < G) {
                            double u = tt[gq];
#pragma unroll
                            for (int off = 1; off < LPO; off <<= 1) u += __shfl_xor_sync(VB_FULL, u, off);
                            if (sub == 0 && owner >= 0) p.ab[((size_t)b * p.T + i) * 2 * G + (al ? 0 : G) + gq] = u;
                        }
                    }
                } else if (do_theta && owner >= 0) {
#pragma unroll
                    for (int gq = 0; gq < VB_MAX_GT; ++gq) { if (al) t1[gq] += tt[gq]; else t2[gq] += tt[gq]; }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
            for (int c = 0; c < NC; ++c) acc[i][c] = 0;
        et = -1;
    }

    // ---------------- block-level partial sums (fixed geometry -> run-to-run identical)
    const int nwt = NW + 1;
    if (sa.mode == GM_CELL || sa.mode == GM_CELL_LL) {
        r0 = warp_sum(r0);
        r1 = warp_sum(r1);
        if (lane == 0) { red[w * VB_SG_RED_DOUBLES] = r0; red[w * VB_SG_RED_DOUBLES + 1] = r1; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double a = 0.0, c = 0.0;
            for (int i = 0; i < nwt; ++i) { a += red[i * VB_SG_RED_DOUBLES]; c += red[i * VB_SG_RED_DOUBLES + 1]; }
            double* out = p.part + (size_t)b * p.part_stride + p.off_cell + 2 * blockIdx.x;
            out[0] = a;
            out[1] = c;
        }
        if (sa.mode == GM_CELL) cell_pass_tail(p, b);
    } else if (sa.mode == GM_SNP && do_theta && !p.ase) {
#pragma unroll
        for (int gq = 0; gq < VB_MAX_GT; ++gq) {
            if (gq < p.G) {
                const double u1 = warp_sum(t1[gq]), u2 = warp_sum(t2[gq]);
                if (lane == 0) { red[w * VB_SG_RED_DOUBLES + gq] = u1; red[w * VB_SG_RED_DOUBLES + VB_MAX_GT + gq] = u2; }
            }
        }
        __syncthreads();
        if (threadIdx.x < 2 * VB_MAX_GT) {
            const int gq = threadIdx.x % VB_MAX_GT;
            double t = 0.0;
            if (gq < p.G)
                for (int i = 0; i < nwt; ++i) t += red[i * VB_SG_RED_DOUBLES + threadIdx.x];
            p.part[(size_t)b * p.part_stride + p.off_theta + (size_t)blockIdx.x * 2 * VB_MAX_GT + threadIdx.x] = t;
        }
        snp_pass_tail(p, b, sa.theta_mode, true);
    } else if (sa.mode == GM_SNP) {
        snp_pass_tail(p, b, sa.theta_mode, false);
    }
}

// ---------------------------------------------------------------------------------------------
// host: dispatch
// ---------------------------------------------------------------------------------------------
static SegView view_of_set(const SegSet& g) {
    SegView v;
    v.n_owner = g.n_owner; v.n_gather = g.n_gather; v.n_task = g.n_task; v.n_task_stream = g.n_task_stream;
    v.n_win = g.n_win; v.win_rows = g.win_rows; v.look = g.look;
    v.perm = g.perm; v.nsteps = g.nsteps; v.task_off = g.task_off; v.rec = g.rec;
    v.hptr = g.hptr; v.hrow = g.hrow; v.hcnt = g.hcnt;
    return v;
}

static size_t seg_smem(int prec, int nb, int win_rows) {
    return (size_t)nb * win_rows * (prec == 0 ? 128 : 64) + 16 * 8 + (size_t)(VB_SEG_MAX_WARPS + 1) * VB_SG_RED_DOUBLES * 8;
}

static bool g_seg_attr_set[64] = {false};     // per device: function attributes belong to the context
static size_t g_seg_static[64] = {0};         // largest static shared memory of the k_seg instances
#define VB_SMEM_OPTIN ((size_t)227 * 1024)

// ori 0: cell pass (table = p.Wt / p.Wq), ori 1: SNP pass (table = p.RP / p.RPq)
// mode GM_PLAIN (ori 0, FP64 tables): `plain` says where the sums of this column chunk go
int vb_seg_launch(const vb_counts* m, const EmP& p, int ori, int mode, int theta_mode, const SegPlain* plain, cudaStream_t st) {
    const int prec = p.tiled == 3 ? 1 : (p.RW == 8 ? 2 : 0);
    const SegSet& g = ori ? m->sB[prec] : m->sA[prec];
    if (!g.built) { vb_set_error("segment format was not built"); return VB_E_ARG; }
    const int dev_slot = m->device >= 0 && m->device < 64 ? m->device : 0;
    if (!g_seg_attr_set[dev_slot]) {
        // the opt-in limit covers static + dynamic shared memory: the fused tails (vb_tail.cuh) keep a few hundred
        // bytes of static shared memory in these kernels
        size_t st_max = 0;
        int rc_attr = VB_OK;
        auto opt_in = [&](const void* fn) {
            cudaFuncAttributes fa;
            if (cudaFuncGetAttributes(&fa, fn) != cudaSuccess ||
                cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(VB_SMEM_OPTIN - fa.sharedSizeBytes)) != cudaSuccess) {
                vb_set_error("segment kernel: shared-memory opt-in failed: %s", cudaGetErrorString(cudaGetLastError()));
                rc_attr = VB_E_CUDA;
                return;
            }
            if (fa.sharedSizeBytes > st_max) st_max = fa.sharedSizeBytes;
        };
        opt_in((const void*)k_seg<0, 4>); opt_in((const void*)k_seg<1, 4>); opt_in((const void*)k_seg<0, 8>);
        opt_in((const void*)k_seg<2, 4>); opt_in((const void*)k_seg<1, 8>);
        if (rc_attr) return rc_attr;
        g_seg_static[dev_slot] = st_max;
        g_seg_attr_set[dev_slot] = true;
    }
    const int nb = g.nb;
    const size_t smem = seg_smem(prec, nb, g.win_rows);
    if (smem + g_seg_static[dev_slot] > VB_SMEM_OPTIN) { vb_set_error("segment kernel: window buffers exceed shared memory"); return VB_E_ARG; }
    const SegView sv = view_of_set(g);
    SegArgs sa;
    memset(&sa, 0, sizeof(sa));
    sa.mode = mode; sa.theta_mode = theta_mode; sa.nb = nb;
    const double* tab64 = ori ? p.RP : p.Wt;
    if (prec != 1) { sa.table = reinterpret_cast<const unsigned char*>(tab64); sa.table_stride = g.n_gather * (prec == 0 ? 128 : 64); }
    else { sa.table = reinterpret_cast<const unsigned char*>(ori ? p.RPq : p.Wq); sa.table_stride = g.n_gather * 64; }
    sa.has_heavy = g.n_heavy > 0;
    if (mode == GM_PLAIN) {
        if (!plain || prec == 1 || ori != 0) { vb_set_error("plain segment pass: bad arguments"); return VB_E_ARG; }
        sa.plain_out = plain->out; sa.plain_ld = plain->ld; sa.plain_off = plain->off; sa.plain_cols = plain->cols;
    }
    static const int wait_hint = env_int("VIREO_B200_SEG_WAIT_NS", 0);
    sa.wait_hint_ns = wait_hint;
    int grid_x;
    vb_seg_geometry(g, &grid_x, &sa.nwarps);
    const int cls = ori ? 0 : 3;
    if (sa.has_heavy) {
        int64_t hb = (g.n_owner + VB_WARPS - 1) / VB_WARPS;
        if (hb > (int64_t)m->sm_count * 8) hb = (int64_t)m->sm_count * 8;
        if (hb < 1) hb = 1;
        VB_LAUNCH(cls, st, k_seg_heavy<<<dim3((unsigned)hb, p.B), VB_THREADS, 0, st>>>(
            g.n_owner, g.hptr, g.hrow, g.hcnt, tab64, g.n_gather * p.RW, p.RW, p.H, g.n_owner * p.RW, p.ctrl));
        VB_CUDA(cudaGetLastError());
    }
    const dim3 grid(grid_x, p.B);
    const int threads = (sa.nwarps + 1) * 32;
    VB_LAUNCH(cls, st, {
        if (prec == 2) k_seg<2, 4><<<grid, threads, smem, st>>>(sv, p, sa);
        else if (prec == 0 && g.depth == 4) k_seg<0, 4><<<grid, threads, smem, st>>>(sv, p, sa);
        else if (prec == 0) k_seg<0, 8><<<grid, threads, smem, st>>>(sv, p, sa);
        else if (g.depth == 4) k_seg<1, 4><<<grid, threads, smem, st>>>(sv, p, sa);
        else k_seg<1, 8><<<grid, threads, smem, st>>>(sv, p, sa);
    });
    VB_CUDA(cudaGetLastError());
    return VB_OK;
}

int vb_seg_quantise_rows(const vb_counts* m, const EmP& p, cudaStream_t st) {
    const int64_t n = (int64_t)p.B * p.C * VB_ROW_DOUBLES;
    int64_t nb = (n + VB_THREADS - 1) / VB_THREADS;
    if (nb > (int64_t)m->sm_count * 8) nb = (int64_t)m->sm_count * 8;
    if (nb < 1) nb = 1;
    VB_LAUNCH(7, st, k_seg_quant_rows<<<(unsigned)nb, VB_THREADS, 0, st>>>(p.RP, n, p.RPq));
    VB_CUDA(cudaGetLastError());
    return VB_OK;
}
