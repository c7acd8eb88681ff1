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
// table row.  Here
//   * 4 lanes share one owner and read its table row with 16-byte loads (two per lane for a 128-byte
//     FP64 row -- one from each 64-byte half, even lane groups starting in the lower half and odd groups
//     in the upper half so that the two rows of a wavefront never meet in a bank; one per lane for the
//     64-byte rows), so a wavefront always carries whole rows;
//   * a warp task holds 32 owners: 8 of them are served per warp step, the accumulators of all 32
//     stay in registers (static indexing: the slot index is the unrolled loop variable);
//   * the table streams through a ring of NB window buffers; window w lives in buffer w % NB, so a table row
//     has ONE shared-memory address for the whole pass and the records carry it (no per-window base, no wrap
//     test).  A producer warp refills a buffer when every consumer warp has released it (full/empty mbarriers);
//   * the records of a task are ONE stream of super-steps of 32 x 32 bits.  A warp holds `span` consecutive
//     windows at a time; the planner (k_sg_plan) schedules every owner's pairs earliest-window-first into
//     lock-step super-steps: a super-step is added only while some owner still has a pair in the oldest held
//     window, every other owner uses the slot for its next pair inside the held windows (null record if it has
//     none).  The windows to advance before a super-step are part of the record, so the kernel's loop has no
//     per-window bookkeeping: decode, two loads, four FMAs per slot;
//   * the count is stored as the upper 16 bits of the double 2 * count: one shift makes the FMA operand, and
//     because those bits are the same for every count the address is ONE shift-add of the whole record (the
//     stray bits are a constant folded into the lane base).  The factor 2 leaves in the epilogue (exact);
//   * the record stream of a task reaches its warp through shared memory as well: two 512-byte chunk buffers per
//     consumer warp, refilled by the warp's own lane 0 with cp.async.bulk two chunks ahead -- no register queue and
//     no exposed L2 latency in the loop;
//   * launches with few warps per CTA use an instance with four landing sets; matrices with few cells but a long
//     table split the table rows over several launches side by side and add the partial sums in k_cell_finish;
//   * PREC 0: FP64 tables of 16 columns (rows of 128 bytes);  PREC 2: FP64 tables of 8 columns for
//     n_donor <= 8 (rows of 64 bytes);  PREC 1 (opt-in) keeps 16 columns as unsigned 32-bit fixed point
//     (rows of 64 bytes: half the crossbar traffic) and accumulates count * value exactly in 64-bit
//     integers, so the result does not depend on the summation order; the quantisation error is bounded
//     per owner by sum(count) * 2^-33 * range and reported by vb_counts_info.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <type_traits>
#include <vector>

#include "vb_common.cuh"
#include "vb_stream.cuh"
#include "vb_tail.cuh"

// ---------------------------------------------------------------------------------------------
// count codes
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline uint32_t seg_count_code(uint32_t c, int fixed) {
    if (fixed) return c;
    // 2c = m * 2^e with 16 <= m < 32 (or c < 16: subnormal-free small integers): build the upper half directly
    uint32_t v = 2u * c, e = 0;
    while ((v >> (e + 1)) != 0) ++e;                    // floor(log2(v))
    const uint32_t mant4 = e >= 4 ? (v >> (e - 4)) & 0xfu : (v << (4 - e)) & 0xfu;
    return ((1023u + e) << 4) | mant4;
}
__host__ __device__ inline uint32_t seg_code_count(uint32_t code, int fixed) {
    if (fixed) return code;
    const uint32_t e = (code >> 4) - 1023u, m = 16u | (code & 0xfu);      // value = m * 2^(e-4) = 2 * count
    return (e >= 4 ? m << (e - 4) : m >> (4 - e)) >> 1;
}

// host-callable copy for the CPU test-suite: the code of `count`, or 0xffffffff when the count has no code
extern "C" uint32_t vb_host_seg_count_code(uint32_t count, int fixed) {
    return count != 0 && seg_count_ok(count, fixed) ? seg_count_code(count, fixed) : 0xffffffffu;
}

// ---------------------------------------------------------------------------------------------
// build kernels
// ---------------------------------------------------------------------------------------------

// pairs per owner: stream pairs (count has a code), residual pairs, reads carried by the stream pairs
template <int ORI, bool WIDE>
__global__ void k_sg_count(const CountsView m, int64_t n_owner, int fixed, uint32_t* __restrict__ n_light, uint32_t* __restrict__ n_heavy,
                           uint32_t* __restrict__ reads, unsigned int* flags) {
    for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < n_owner; o += (int64_t)gridDim.x * blockDim.x) {
        uint32_t nl = 0, nh = 0, rd = 0;
        const bool ok = for_records<ORI, WIDE>(m, o, [&](int, uint32_t c) {
            if (!seg_count_ok(c, fixed)) ++nh;
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
                              int fixed, uint16_t* __restrict__ cnt) {
    for (int64_t pos = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; pos < n_active; pos += (int64_t)gridDim.x * blockDim.x) {
        const int64_t o = perm[pos];
        uint16_t* row = cnt + (size_t)pos * n_win;
        int w = 0;
        uint32_t n = 0;
        for_records<ORI, WIDE>(m, o, [&](int g, uint32_t c) {
            if (!seg_count_ok(c, fixed)) return;
            const int gw = g / win_rows;
            if (gw != w) { if (n) row[w] = (uint16_t)n; w = gw; n = 0; }
            ++n;
        });
        if (n) row[w] = (uint16_t)n;
    }
}

// The lock-step schedule of one task, one warp per task, lane = owner slot.  In segment s the warp holds the windows
// s .. s + span - 1.  Every owner runs its pairs earliest window first; the segment lasts as long as some owner still
// has a pair of window s (which leaves the ring next), and in every one of its super-steps each owner with a pair
// inside the held windows executes one.  That is the shortest lock-step schedule for the given span.
//   cnt[pos][s]   in: pairs of the owner in window s;  out: pairs the owner executes during segment s
//   segn[t][s]    super-steps of segment s            adv[t][s]  windows to advance before its first super-step
//   wide[t][s]    = segn, the last one of a task padded so that the task's total is a multiple of `depth`
//   tail[t]       window advances left when the stream ends (the warp releases every window of the table)
// An empty segment passes its advance on to the next super-step; a run of VB_SEG_ADV_MAX empty segments gets a null
// super-step to carry the advances.
__global__ void __launch_bounds__(256)
k_sg_plan(uint16_t* __restrict__ cnt, int64_t n_active, int64_t n_task_stream, int n_win, int span, int depth,
          uint16_t* __restrict__ segn, uint8_t* __restrict__ adv, int64_t* __restrict__ wide, int32_t* __restrict__ tail) {
    const int lane = threadIdx.x & 31;
    const int64_t t = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (t > n_task_stream) return;
    if (t == n_task_stream) { if (lane == 0) wide[t * n_win] = 0; return; }
    const int64_t pos = t * VB_SEG_OWNERS + lane;
    const bool active = pos < n_active;
    uint16_t* row = cnt + (size_t)(active ? pos : 0) * n_win;
    uint32_t consumed = 0, cum_own = 0, cum_held = 0, pending = 0;
    int64_t total = 0;
    for (int w = 0; w < span - 1 && w < n_win; ++w) cum_held += active ? row[w] : 0u;
    for (int s = 0; s < n_win; ++s) {
        cum_own += active ? row[s] : 0u;
        if (s + span - 1 < n_win) cum_held += active ? row[s + span - 1] : 0u;
        uint32_t n = cum_own > consumed ? cum_own - consumed : 0u;      // pairs that must run before window s is released
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) { const uint32_t o = __shfl_xor_sync(VB_FULL, n, off); n = o > n ? o : n; }
        if (s > 0) ++pending;
        if (n == 0 && pending >= VB_SEG_ADV_MAX) n = 1;
        const uint32_t avail = cum_held - consumed;
        const uint32_t e = n < avail ? n : avail;
        consumed += e;
        if (active) row[s] = (uint16_t)e;
        if (lane == 0) {
            segn[t * n_win + s] = (uint16_t)n;
            adv[t * n_win + s] = (uint8_t)(n ? pending : 0u);
            wide[t * n_win + s] = (int64_t)n;
        }
        if (n) pending = 0;
        total += n;
    }
    if (lane == 0) {
        wide[t * n_win + n_win - 1] += (total + depth - 1) / depth * depth - total;
        tail[t] = (int32_t)pending + 1;      // advances the stream does not carry: the warp releases all n_win windows
    }
}

__global__ void k_sg_widen(const uint32_t* __restrict__ in, int64_t n, int64_t* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i <= n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = i < n ? (int64_t)in[i] : 0;
}

// the windows to advance before the first super-step of every non-empty segment, in the first slot of each lane group
__global__ void k_sg_mark_adv(const uint16_t* __restrict__ segn, const uint8_t* __restrict__ adv, const int64_t* __restrict__ step_off,
                              int64_t n_tw, uint32_t* __restrict__ rec) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_tw; i += (int64_t)gridDim.x * blockDim.x) {
        if (segn[i] == 0 || adv[i] == 0) continue;
        uint32_t* ss = rec + (size_t)step_off[i] * VB_SEG_OWNERS;
        for (int j = 0; j < VB_SEG_OWNERS; j += 4) ss[j] = (uint32_t)adv[i] << 28;
    }
}

// write the records and the residual CSR; one thread per sorted position.  The owner's pairs, in gather-row order,
// fill its slot of segment 0's first exec[pos][0] super-steps, then segment 1's, ...  The order of an owner's records
// inside a segment is free.  With 64-byte table rows (`pair_dist` > 0) two owner slots share one shared-memory
// wavefront (slots s and s + pair_dist of a super-step): rows of equal parity hit the same 16 banks.  Slot s
// therefore lists its even rows from the front and its odd rows from the back of the segment, slot s + pair_dist the
// other way round, so that most steps pair an even with an odd row.
template <int ORI, bool WIDE>
__global__ void k_sg_fill(const CountsView m, int64_t n_owner, int64_t n_active, const int32_t* __restrict__ perm, int ring_rows,
                          int n_win, int fixed, int pair_dist, const int64_t* __restrict__ step_off,
                          const uint16_t* __restrict__ exec, uint32_t* __restrict__ rec,
                          const int64_t* __restrict__ hptr, int32_t* __restrict__ hrow, uint32_t* __restrict__ hcnt) {
    for (int64_t pos = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; pos < n_owner; pos += (int64_t)gridDim.x * blockDim.x) {
        const int64_t o = perm[pos];
        const bool streams = pos < n_active;
        const int64_t t = pos / VB_SEG_OWNERS;
        const int slot = (int)(pos % VB_SEG_OWNERS);
        const uint32_t front_parity = (pair_dist > 0 && (slot & pair_dist)) ? 1u : 0u;    // row parity listed from the front
        const uint16_t* erow = exec + (size_t)(streams ? pos : 0) * n_win;
        int64_t h = hptr[o];
        int seg = -1;
        int k = 0, e_cur = 0, f = 0, b = 0;       // pairs placed in the segment, its pairs, cursors from the front / the back
        int64_t base = 0;
        for_records<ORI, WIDE>(m, o, [&](int g, uint32_t c) {
            if (!streams || !seg_count_ok(c, fixed)) { hrow[h] = g; hcnt[h] = c; ++h; return; }
            while (k == e_cur) {
                if (seg + 1 >= n_win) return;       // cannot happen: the plan executes every pair by its own window
                ++seg;
                k = f = b = 0;
                e_cur = erow[seg];
                base = step_off[t * n_win + seg];
            }
            const uint32_t rr = (uint32_t)(g % ring_rows);
            int at;
            if (pair_dist == 0 || (rr & 1u) == front_parity) at = f++;
            else at = e_cur - 1 - b++;
            uint32_t* dst = rec + ((size_t)(base + at) * VB_SEG_OWNERS + slot);
            const uint32_t keep = (slot & 3) == 0 ? (*dst & 0xf0000000u) : 0u;        // k_sg_mark_adv ran before
            *dst = keep | (rr << 16) | seg_count_code(c, fixed);
            ++k;
        });
    }
}

// ---------------------------------------------------------------------------------------------
// host: build one orientation
// ---------------------------------------------------------------------------------------------
static void seg_set_free(SegSet& g) {
    cudaFree(g.perm); cudaFree(g.task_off); cudaFree(g.tail); cudaFree(g.rec);
    cudaFree(g.hptr); cudaFree(g.hrow); cudaFree(g.hcnt);
    cudaFree(g.o_split); cudaFree(g.v_owner); cudaFree(g.v_part); cudaFree(g.vsum);
    if (g.excess) { seg_set_free(*g.excess); delete g.excess; }
    memset(&g, 0, sizeof(g));
}

// what a format covers: the table rows [g_lo, g_hi) (all of them: 0 / -1), on at most max_grid CTAs (0: every SM); split
// owners: o_split alone = part 0 of every real owner, with v_owner / v_part = the n_virtual virtual owners
struct SegBuildOpts {
    int64_t g_lo = 0, g_hi = -1;
    int max_grid = 0;
    const uint8_t* o_split = nullptr;
    const int32_t* v_owner = nullptr;
    const uint8_t* v_part = nullptr;
    int64_t n_virtual = 0;
};

static CountsView sg_view(const vb_counts* m, int64_t g_lo, int64_t g_hi, int fixed, const SegBuildOpts& opt) {
    CountsView v;
    v.C = m->C; v.V = m->V; v.N = m->N;
    v.g_lo = g_lo; v.g_hi = g_hi;
    v.fixed = fixed; v.o_split = opt.o_split; v.v_owner = opt.v_owner; v.v_part = opt.v_part;
    v.cell_ptr = m->cell_ptr; v.cell_idx = m->cell_idx; v.cell_cnt = m->cell_cnt; v.cell_dp = m->cell_dp;
    v.snp_ptr = m->snp_ptr; v.snp_idx = m->snp_idx; v.snp_cnt = m->snp_cnt; v.snp_dp = m->snp_dp;
    return v;
}

static int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

// ring geometry per table precision: rows per window, window buffers (shared memory: nb * rows * row bytes) and the
// windows a warp holds at a time.  nb - span buffers are what the producer can refill while the slowest warp still
// holds its oldest window.
static void seg_window(int prec, int* win_rows, int* nb, int* span) {
    if (prec == 0) {
        *win_rows = env_int("VIREO_B200_SEG_WR64", 256); *nb = env_int("VIREO_B200_SEG_NB64", 6); *span = env_int("VIREO_B200_SEG_SPAN64", 4);
    } else {                                                                                     // 64-byte rows
        *win_rows = env_int("VIREO_B200_SEG_WR32", 512); *nb = env_int("VIREO_B200_SEG_NB32", 6); *span = env_int("VIREO_B200_SEG_SPAN32", 4);
    }
    const int row_bytes = prec == 0 ? 128 : 64;
    if (*nb < 2) *nb = 2;
    if (*nb > VB_SEG_MAX_NB) *nb = VB_SEG_MAX_NB;
    if (*win_rows < 32) *win_rows = 32;
    *win_rows &= ~7;
    while ((size_t)*nb * *win_rows * row_bytes > 192 * 1024 || *nb * *win_rows > VB_SEG_MAX_RING_ROWS) *win_rows -= 8;
    if (*span > *nb - 1) *span = *nb - 1;
    if (*span < 1) *span = 1;
}

template <int ORI>
static int seg_build_one(vb_counts* m, SegSet& g, int prec, cudaStream_t st, const SegBuildOpts& opt = SegBuildOpts()) {
    memset(&g, 0, sizeof(g));
    const int max_grid = opt.max_grid;
    int64_t g_lo = opt.g_lo, g_hi = opt.g_hi;
    const int sm = max_grid > 0 ? max_grid : m->sm_count;
    const int64_t O = opt.v_owner ? opt.n_virtual : (ORI == 0 ? m->C : 2 * m->V);
    const int64_t G_all = ORI == 0 ? 2 * m->V : m->C;
    if (g_hi < 0 || g_hi > G_all) g_hi = G_all;
    const int64_t Gn = g_hi - g_lo;
    if (O >= (1ll << 31) - 64 || Gn >= (1ll << 31) - 4096) { vb_set_error("segment format: more than 2^31 rows"); return VB_E_UNSUPPORTED; }
    int win_rows, nb, span;
    seg_window(prec, &win_rows, &nb, &span);
    const int fixed = prec == 1 ? 1 : 0;
    const int depth = VB_SEG_DEPTH;
    int n_win = (int)((Gn + win_rows - 1) / win_rows);
    if (n_win < 1) n_win = 1;
    const int64_t n_task = (O + VB_SEG_OWNERS - 1) / VB_SEG_OWNERS;
    const CountsView v = sg_view(m, g_lo, g_hi, fixed, opt);
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
        if (m->wide) k_sg_count<ORI, true><<<grid1d(O, sm), 256, 0, st>>>(v, O, fixed, nl, nh, rd, flags);
        else k_sg_count<ORI, false><<<grid1d(O, sm), 256, 0, st>>>(v, O, fixed, nl, nh, rd, flags);
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
        // formats of the row-split pass and of virtual owners keep every owner in the stream
        const int64_t budget = (max_grid > 0 || opt.v_owner) ? 0 : (int64_t)(total / 50.0);
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
    g.n_owner = O; g.n_gather = Gn; g.n_task = n_task; g.n_win = n_win; g.win_rows = win_rows; g.nb = nb; g.span = span; g.fixed = fixed;
    g.n_light = (int64_t)hsums[0]; g.n_heavy = (int64_t)hsums[1]; g.max_reads = (int64_t)hsums[2];
    g.max_len = O ? (int64_t)hlen[0] : 0;
    g.n_task_stream = (n_active + VB_SEG_OWNERS - 1) / VB_SEG_OWNERS;
    const int64_t nts = g.n_task_stream;

    VB_CUDA(cudaMalloc(&g.perm, (n_task ? n_task : 1) * VB_SEG_OWNERS * sizeof(int32_t)));
    VB_CUDA(cudaMemsetAsync(g.perm, 0xff, (n_task ? n_task : 1) * VB_SEG_OWNERS * sizeof(int32_t), st));
    if (O) VB_CUDA(cudaMemcpyAsync(g.perm, perm_sorted, O * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));

    // the schedule: super-steps per (task, segment) and their offsets
    const int64_t n_tw = nts * n_win;
    uint16_t *cnt_ow, *segn;
    uint8_t* adv;
    int64_t *wide, *step_off;
    const size_t n_cw = (size_t)(n_active ? n_active : 1) * n_win;
    if ((rc = tmp.alloc(&cnt_ow, n_cw)) || (rc = tmp.alloc(&segn, n_tw + 1)) || (rc = tmp.alloc(&adv, n_tw + 1)) ||
        (rc = tmp.alloc(&wide, n_tw + 1)) || (rc = tmp.alloc(&step_off, n_tw + 1)))
        return rc;
    VB_CUDA(cudaMalloc(&g.task_off, (nts + 1) * sizeof(int64_t)));
    VB_CUDA(cudaMalloc(&g.tail, (nts + 1) * sizeof(int32_t)));
    VB_CUDA(cudaMemsetAsync(cnt_ow, 0, n_cw * sizeof(uint16_t), st));
    if (n_active) {
        if (m->wide) k_sg_wincount<ORI, true><<<grid1d(n_active, sm), 256, 0, st>>>(v, n_active, g.perm, win_rows, n_win, fixed, cnt_ow);
        else k_sg_wincount<ORI, false><<<grid1d(n_active, sm), 256, 0, st>>>(v, n_active, g.perm, win_rows, n_win, fixed, cnt_ow);
        VB_CUDA(cudaGetLastError());
    }
    k_sg_plan<<<(unsigned)((nts + 1 + 7) / 8), 256, 0, st>>>(cnt_ow, n_active, nts, n_win, span, depth, segn, adv, wide, g.tail);
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
    // the kernel reads whole chunks of `depth` super-steps and nothing past the last one; a little slack all the same
    const size_t n_rec = ((size_t)total_steps + 2 * depth) * VB_SEG_OWNERS + 256;
    VB_CUDA(cudaMalloc(&g.rec, n_rec * sizeof(uint32_t)));
    VB_CUDA(cudaMemsetAsync(g.rec, 0, n_rec * sizeof(uint32_t), st));
    if (n_tw) {
        k_sg_mark_adv<<<grid1d(n_tw, sm), 256, 0, st>>>(segn, adv, step_off, n_tw, g.rec);
        VB_CUDA(cudaGetLastError());
    }
    if (O) {
        const int pair_dist = prec != 0 ? 4 : 0;
        if (m->wide) k_sg_fill<ORI, true><<<grid1d(O, sm), 256, 0, st>>>(v, O, n_active, g.perm, nb * win_rows, n_win, fixed, pair_dist, step_off, cnt_ow, g.rec, g.hptr, g.hrow, g.hcnt);
        else k_sg_fill<ORI, false><<<grid1d(O, sm), 256, 0, st>>>(v, O, n_active, g.perm, nb * win_rows, n_win, fixed, pair_dist, step_off, cnt_ow, g.rec, g.hptr, g.hrow, g.hcnt);
        VB_CUDA(cudaGetLastError());
    }
    VB_CUDA(cudaStreamSynchronize(st));

    // launch geometry: one CTA on every SM when there are enough streaming tasks, as few consumer warps as cover them (the
    // last, partly filled round of the deal holds the lightest tasks of the sorted list)
    int64_t grid = nts < sm ? nts : sm;
    if (grid < 1) grid = 1;
    int nw = (int)((nts + grid - 1) / grid);
    if (nw < 1) nw = 1;
    if (nw > VB_SEG_MAX_WARPS) { nw = VB_SEG_MAX_WARPS; grid = (nts + nw - 1) / nw; }
    if (grid > 65535 * 16) { vb_set_error("segment format: too many owner rows"); return VB_E_UNSUPPORTED; }
    g.grid = (int)grid; g.nwarps = nw;
    g.bytes = (int64_t)n_rec * 4 + n_task * VB_SEG_OWNERS * 4 + (nts + 1) * 8 + (O + 1) * 8 + g.n_heavy * 8;
    g.built = 1;
    return VB_OK;
}

// ---------------------------------------------------------------------------------------------
// rows far heavier than the rest
// ---------------------------------------------------------------------------------------------
// parts per owner: ceil(stream pairs / cap), at most 255; extra[o] = parts - 1 virtual owners
__global__ void k_sg_parts(const uint32_t* __restrict__ n_light, int64_t n_owner, uint32_t cap, uint8_t* __restrict__ parts,
                           uint32_t* __restrict__ extra) {
    for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < n_owner; o += (int64_t)gridDim.x * blockDim.x) {
        uint32_t s = (n_light[o] + cap - 1) / cap;
        s = s < 1u ? 1u : (s > 255u ? 255u : s);
        parts[o] = (uint8_t)s;
        extra[o] = s - 1u;
    }
}
__global__ void k_sg_virtual(const uint8_t* __restrict__ parts, const int64_t* __restrict__ v_off, int64_t n_owner,
                             int32_t* __restrict__ v_owner, uint8_t* __restrict__ v_part) {
    for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < n_owner; o += (int64_t)gridDim.x * blockDim.x)
        for (uint32_t j = 1; j < parts[o]; ++j) { v_owner[v_off[o] + j - 1] = (int32_t)o; v_part[v_off[o] + j - 1] = (uint8_t)j; }
}

// One orientation of one table kind.  When some owner carries more than VB_SEG_SKEW x the mean number of stream pairs
// (a task is as long as its longest owner, so that owner's warp would be the critical path of the pass), the owners
// above twice the mean are cut into parts: part 0 stays in this format, the others become virtual owners of g.excess.
#define VB_SEG_SKEW 4
template <int ORI>
static int seg_build_fmt(vb_counts* m, SegSet& g, int prec, cudaStream_t st) {
    int rc = seg_build_one<ORI>(m, g, prec, st);
    static const int on = env_int("VIREO_B200_SEG_OWNER_SPLIT", 1);
    if (rc || !on || prec == 1 || g.n_owner < 1 || g.n_light < 1) return rc;
    const double mean = (double)g.n_light / (double)g.n_owner;
    static const double skew = getenv("VIREO_B200_SEG_SKEW") ? atof(getenv("VIREO_B200_SEG_SKEW")) : (double)VB_SEG_SKEW;
    static const double capf = getenv("VIREO_B200_SEG_CAP") ? atof(getenv("VIREO_B200_SEG_CAP")) : 2.0;
    if ((double)g.max_len <= skew * mean || g.max_len <= 256) return VB_OK;
    const int64_t O = g.n_owner;
    const int sm = m->sm_count;
    const uint32_t cap = (uint32_t)(capf * mean) > 128u ? (uint32_t)(capf * mean) : 128u;
    GsScratch tmp;
    uint32_t *nl, *nh, *rd, *extra;
    unsigned int* flags;
    int64_t *ew, *v_off;
    uint8_t* parts = nullptr;
    int32_t* v_owner = nullptr;
    uint8_t* v_part = nullptr;
    if ((rc = tmp.alloc(&nl, O)) || (rc = tmp.alloc(&nh, O)) || (rc = tmp.alloc(&rd, O)) || (rc = tmp.alloc(&extra, O + 1)) ||
        (rc = tmp.alloc(&flags, 4)) || (rc = tmp.alloc(&ew, O + 1)) || (rc = tmp.alloc(&v_off, O + 1)))
        return rc;
    const CountsView v = sg_view(m, 0, ORI == 0 ? 2 * m->V : m->C, 0, SegBuildOpts());
    VB_CUDA(cudaMemsetAsync(flags, 0, 4 * sizeof(unsigned int), st));
    if (m->wide) k_sg_count<ORI, true><<<grid1d(O, sm), 256, 0, st>>>(v, O, 0, nl, nh, rd, flags);
    else k_sg_count<ORI, false><<<grid1d(O, sm), 256, 0, st>>>(v, O, 0, nl, nh, rd, flags);
    VB_CUDA(cudaGetLastError());
    VB_CUDA(cudaMalloc(&parts, O));
    k_sg_parts<<<grid1d(O, sm), 256, 0, st>>>(nl, O, cap, parts, extra);
    k_sg_widen<<<grid1d(O + 1, sm), 256, 0, st>>>(extra, O, ew);
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, ew, v_off, O + 1, st);
    void* cub_tmp;
    if ((rc = tmp.alloc((unsigned char**)&cub_tmp, tb))) { cudaFree(parts); return rc; }
    cub::DeviceScan::ExclusiveSum(cub_tmp, tb, ew, v_off, O + 1, st);
    int64_t n_virtual = 0;
    VB_CUDA(cudaMemcpyAsync(&n_virtual, v_off + O, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaStreamSynchronize(st));
    if (n_virtual < 1 || n_virtual >= (1ll << 31) - 64) { cudaFree(parts); return VB_OK; }
    if (cudaMalloc(&v_owner, n_virtual * sizeof(int32_t)) != cudaSuccess || cudaMalloc(&v_part, n_virtual) != cudaSuccess) {
        cudaFree(parts); cudaFree(v_owner); cudaGetLastError();
        return VB_OK;                                   // the unsplit format stays
    }
    k_sg_virtual<<<grid1d(O, sm), 256, 0, st>>>(parts, v_off, O, v_owner, v_part);
    VB_CUDA(cudaGetLastError());
    // part 0 of every owner in this format, the other parts as owners of the excess format
    SegSet main_set, *ex = new SegSet;
    SegBuildOpts om, oe;
    om.o_split = parts;
    oe.o_split = parts; oe.v_owner = v_owner; oe.v_part = v_part; oe.n_virtual = n_virtual;
    rc = seg_build_one<ORI>(m, main_set, prec, st, om);
    if (!rc) rc = seg_build_one<ORI>(m, *ex, prec, st, oe);
    if (!rc && (ex->n_heavy > 0 || ex->n_light < 1)) rc = VB_E_UNSUPPORTED;      // virtual owners stream every pair they have
    if (rc) {                                                                    // keep the unsplit format
        seg_set_free(main_set); seg_set_free(*ex); delete ex;
        cudaFree(parts); cudaFree(v_owner); cudaFree(v_part);
        cudaGetLastError();
        return VB_OK;
    }
    seg_set_free(g);
    g = main_set;
    g.o_split = parts; g.n_virtual = n_virtual; g.v_owner = v_owner; g.v_part = v_part; g.excess = ex;
    g.n_light += ex->n_light;                           // stream pairs of the pass, for the selector's residual rule
    g.bytes += ex->bytes + O + n_virtual * 5;
    return VB_OK;
}

int vb_seg_build(vb_counts* m, int prec, cudaStream_t st) {
    if (prec < 0 || prec > 2) { vb_set_error("bad table kind"); return VB_E_ARG; }
    if (m->sA[prec].built && m->sB[prec].built) return VB_OK;
    if (m->seg_failed[prec]) return VB_E_UNSUPPORTED;
    DeviceGuard dg(m->device);
    int rc = seg_build_fmt<0>(m, m->sA[prec], prec, st);
    if (!rc) rc = seg_build_fmt<1>(m, m->sB[prec], prec, st);
    if (rc) {
        seg_set_free(m->sA[prec]);
        seg_set_free(m->sB[prec]);
        m->seg_failed[prec] = 1;
        cudaGetLastError();
    }
    return rc;
}

static void seg_split_free(SegSplit& sp);

void vb_seg_free(vb_counts* m) {
    for (int i = 0; i < 3; ++i) { seg_set_free(m->sA[i]); seg_set_free(m->sB[i]); seg_split_free(m->rA[i]); }
}

void vb_seg_geometry(const SegSet& g, int* grid, int* nwarps) {
    *grid = g.grid > 0 ? g.grid : 1;
    *nwarps = g.nwarps > 0 ? g.nwarps : 1;
}

// ---------------------------------------------------------------------------------------------
// format check (tests): every pair of the staged counts appears exactly once -- as a record that executes while
// its table row is inside the windows the warp holds, or in the residual CSR.  Host-side, for small matrices.
//   out[0] pairs checked, out[1] errors, out[2] super-steps, out[3] null slots
// ---------------------------------------------------------------------------------------------
struct SegPair { int32_t owner, row; uint32_t count; };
static bool operator<(const SegPair& a, const SegPair& b) {
    if (a.owner != b.owner) return a.owner < b.owner;
    if (a.row != b.row) return a.row < b.row;
    return a.count < b.count;
}
static bool operator==(const SegPair& a, const SegPair& b) { return a.owner == b.owner && a.row == b.row && a.count == b.count; }

template <typename T> static int seg_download(std::vector<T>& h, const T* d, size_t n) {
    h.resize(n);
    if (n) VB_CUDA(cudaMemcpy(h.data(), d, n * sizeof(T), cudaMemcpyDeviceToHost));
    return VB_OK;
}

extern "C" int vb_seg_verify(vb_counts* m, int prec, int ori, int64_t* out) {
    if (!m || !out || prec < 0 || prec > 2) { vb_set_error("vb_seg_verify: bad arguments"); return VB_E_ARG; }
    DeviceGuard dg(m->device);
    int rc = vb_seg_build(m, prec, 0);
    if (rc) return rc;
    VB_CUDA(cudaDeviceSynchronize());
    const SegSet& g = ori ? m->sB[prec] : m->sA[prec];
    std::vector<int64_t> cptr, toff, hptr;
    std::vector<int32_t> cidx, perm, hrow, tail;
    std::vector<uint32_t> ccnt, cdp, rec, hcnt;
    if ((rc = seg_download(cptr, m->cell_ptr, (size_t)m->C + 1)) || (rc = seg_download(cidx, m->cell_idx, (size_t)m->N)) ||
        (rc = seg_download(ccnt, m->cell_cnt, (size_t)m->N)) || (m->wide && (rc = seg_download(cdp, m->cell_dp, (size_t)m->N))) ||
        (rc = seg_download(toff, g.task_off, (size_t)g.n_task_stream + 1)) || (rc = seg_download(tail, g.tail, (size_t)g.n_task_stream)) || (rc = seg_download(perm, g.perm, (size_t)g.n_task * VB_SEG_OWNERS)) ||
        (rc = seg_download(rec, g.rec, (size_t)g.n_step * VB_SEG_OWNERS)) || (rc = seg_download(hptr, g.hptr, (size_t)g.n_owner + 1)) ||
        (rc = seg_download(hrow, g.hrow, (size_t)g.n_heavy)) || (rc = seg_download(hcnt, g.hcnt, (size_t)g.n_heavy)))
        return rc;
    std::vector<SegPair> truth, got;
    for (int64_t j = 0; j < m->C; ++j)
        for (int64_t q = cptr[j]; q < cptr[j + 1]; ++q) {
            const uint32_t a = m->wide ? ccnt[q] : (ccnt[q] & 0xffffu), d = m->wide ? cdp[q] : (ccnt[q] >> 16);
            const int32_t i = cidx[q];
            if (d - a) truth.push_back(ori ? SegPair{2 * i, (int32_t)j, d - a} : SegPair{(int32_t)j, 2 * i, d - a});
            if (a) truth.push_back(ori ? SegPair{2 * i + 1, (int32_t)j, a} : SegPair{(int32_t)j, 2 * i + 1, a});
        }
    int64_t errors = 0, nulls = 0;
    const int64_t ring_rows = (int64_t)g.nb * g.win_rows;
    for (int64_t t = 0; t < g.n_task_stream; ++t) {
        if ((toff[t + 1] - toff[t]) % VB_SEG_DEPTH) ++errors;
        int64_t s = 0;
        for (int64_t u = toff[t]; u < toff[t + 1]; ++u) {
            const uint32_t* ss = &rec[(size_t)u * VB_SEG_OWNERS];
            const uint32_t adv = ss[0] >> 28;
            s += adv;
            if (s >= g.n_win) { ++errors; break; }
            for (int sl = 0; sl < VB_SEG_OWNERS; ++sl) {
                uint32_t w = ss[sl];
                if ((w >> 28) != ((sl & 3) == 0 ? adv : 0u)) ++errors;
                w &= 0x0fffffffu;
                if ((w & 0xffffu) == 0) { ++nulls; if (w) ++errors; continue; }
                const int64_t rr = w >> 16, lo = s * g.win_rows, hi = std::min<int64_t>((s + g.span) * (int64_t)g.win_rows, g.n_gather);
                int64_t r = lo - lo % ring_rows + rr;
                if (r < lo) r += ring_rows;
                const int32_t owner = perm[(size_t)t * VB_SEG_OWNERS + sl];
                if (r >= hi || owner < 0) { ++errors; continue; }
                got.push_back(SegPair{owner, (int32_t)r, seg_code_count(w & 0xffffu, g.fixed)});
            }
        }
        if (s + tail[t] != g.n_win) ++errors;            // stream + tail release every window exactly once
    }
    for (int64_t o = 0; o < g.n_owner; ++o)
        for (int64_t q = hptr[o]; q < hptr[o + 1]; ++q) got.push_back(SegPair{(int32_t)o, hrow[q], hcnt[q]});
    if (g.excess) {      // split owners: the records of the virtual owners belong to their real owners
        const SegSet& e = *g.excess;
        std::vector<int64_t> etoff;
        std::vector<int32_t> eperm, vown;
        std::vector<uint32_t> erec;
        if ((rc = seg_download(etoff, e.task_off, (size_t)e.n_task_stream + 1)) || (rc = seg_download(eperm, e.perm, (size_t)e.n_task * VB_SEG_OWNERS)) ||
            (rc = seg_download(erec, e.rec, (size_t)e.n_step * VB_SEG_OWNERS)) || (rc = seg_download(vown, g.v_owner, (size_t)g.n_virtual)))
            return rc;
        if (e.n_heavy) ++errors;
        for (int64_t t = 0; t < e.n_task_stream; ++t) {
            int64_t s = 0;
            for (int64_t u = etoff[t]; u < etoff[t + 1]; ++u) {
                const uint32_t* ss = &erec[(size_t)u * VB_SEG_OWNERS];
                s += ss[0] >> 28;
                if (s >= e.n_win) { ++errors; break; }
                for (int sl = 0; sl < VB_SEG_OWNERS; ++sl) {
                    const uint32_t w = ss[sl] & 0x0fffffffu;
                    if ((w & 0xffffu) == 0) { ++nulls; continue; }
                    const int64_t rr = w >> 16, lo = s * e.win_rows, hi = std::min<int64_t>((s + e.span) * (int64_t)e.win_rows, e.n_gather);
                    int64_t r = lo - lo % ring_rows + rr;
                    if (r < lo) r += ring_rows;
                    const int32_t vo = eperm[(size_t)t * VB_SEG_OWNERS + sl];
                    if (r >= hi || vo < 0 || vo >= g.n_virtual) { ++errors; continue; }
                    got.push_back(SegPair{vown[vo], (int32_t)r, seg_code_count(w & 0xffffu, e.fixed)});
                }
            }
        }
        out[2] = g.n_step + e.n_step;
    }
    std::sort(truth.begin(), truth.end());
    std::sort(got.begin(), got.end());
    if (truth.size() != got.size()) errors += (int64_t)(truth.size() > got.size() ? truth.size() - got.size() : got.size() - truth.size());
    for (size_t i = 0; i < truth.size() && i < got.size(); ++i)
        if (!(truth[i] == got[i])) ++errors;
    out[0] = (int64_t)truth.size(); out[1] = errors; if (!g.excess) out[2] = g.n_step; out[3] = nulls;
    return VB_OK;
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
#define VB_SEG_FEW_WARPS 11         // consumer warps per CTA up to which the 4-landing-set instance of k_seg is launched
// How the producer / the consumers wait for a window barrier: a warp that spins on try_wait competes with the working
// warps of its scheduler for every issue cycle, so between polls it sleeps (nanoseconds; 0 = spin).  Measured at cfg3:
// 773 -> 782 it/s with 256 / 64; 512 / 128 and 128 / 32 give less, 1024 / 256 loses (build.py variant "spin" = 0 / 0).
#ifndef VB_SEG_SLEEP_P
#define VB_SEG_SLEEP_P 256
#endif
#ifndef VB_SEG_SLEEP_C
#define VB_SEG_SLEEP_C 64
#endif
#if VB_SEG_SLEEP_P > 0
#define VB_SEG_PWAIT(bar, par) mbar_wait_sleep(bar, par, VB_SEG_SLEEP_P)
#else
#define VB_SEG_PWAIT(bar, par) mbar_wait(bar, par)
#endif
#if VB_SEG_SLEEP_C > 0
#define VB_SEG_CWAIT(bar, par) mbar_wait_sleep(bar, par, VB_SEG_SLEEP_C)
#else
#define VB_SEG_CWAIT(bar, par) mbar_wait(bar, par)
#endif
#define VB_SG_RECBAR_OFF 256u        // record-chunk barriers of the consumer warps (two each) behind the window barriers
#define VB_SG_RED_OFF 1024u         // block-reduction scratch
#define VB_SG_PIECE 16384u          // bytes per bulk copy
#define VB_SG_RING_OFF 4096u        // the window ring starts here in dynamic shared memory (barriers and scratch below)

struct SegArgs {
    int mode;            // GM_CELL, GM_CELL_LL, GM_SNP
    int theta_mode;      // GM_SNP: 0 never, 1 always, 2 per the device iteration counter
    int nwarps;          // consumer warps per CTA (block size = (nwarps + 1) * 32)
    int has_heavy;       // add p.H[owner] before the epilogue
    int64_t table_stride;    // bytes per restart of the gather table
    const unsigned char* table;
    // GM_PLAIN: the sums of columns [0, plain_cols) go to plain_out[owner * plain_ld + plain_off + column]
    double* plain_out;
    int64_t plain_ld;
    int plain_off, plain_cols;
    int64_t plain_stride;    // doubles between the outputs of consecutive restarts (0: one restart)
};

template <int PREC> struct SegCfg;
template <> struct SegCfg<0> { static constexpr int LPO = 4, NC = 4, ROWB = 128, SHIFT = 9; };
template <> struct SegCfg<1> { static constexpr int LPO = 4, NC = 4, ROWB = 64, SHIFT = 10; };
template <> struct SegCfg<2> { static constexpr int LPO = 4, NC = 2, ROWB = 64, SHIFT = 10; };     // FP64, 8 columns

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// One owner slot of one super-step.  Record w = ring row << 16 | count code (the advance bits are already cleared).
// The shared-memory address of this lane's granule is ONE shift-add: w >> SHIFT = ring row * row bytes + the upper
// bits of the count code, which are the same for every count (FP64: sign 0 and the exponent bits of 2 <= 2c < 2^26)
// or zero (fixed point: count <= 31) -- `base` has that constant subtracted.  A null record (w = 0) computes an
// address below its row 0 and loads nothing.
// FP64 tables: 4 lanes x 4 columns, two 16-byte loads per lane, one from each 64-byte half of the row (even lane
// groups start in the lower half, odd groups in the upper half: the two rows that share a wavefront never meet in a
// bank).  The count operand is the record shifted into the upper word of a double: 2 * count, the epilogue halves.
#ifdef VB_SEG_DIAG_NOLDS      // timing diagnostic (build.py variant "nolds"): the table loads never execute
#define VB_SEG_SETP4 "setp.eq.u32 p, %4, 0xffffffff;\n\t"
#define VB_SEG_SETP2 "setp.eq.u32 p, %2, 0xffffffff;\n\t"
#else
#define VB_SEG_SETP4 "setp.ne.u32 p, %4, 0;\n\t"
#define VB_SEG_SETP2 "setp.ne.u32 p, %2, 0;\n\t"
#endif
template <int NSET> struct SegScratch64 { double v[NSET][4]; };     // landing registers of the FP64 loads, NSET slots in flight
struct SegScratch32 {};
template <int NSET> struct SegScratchN { double v[NSET][2]; };      // narrow FP64 rows: one 16-byte load per lane and record

__device__ __forceinline__ void seg_step(uint32_t w, uint32_t base, double (&a)[4], double (&v)[4]) {
    const uint32_t xh = w << 16;
    const uint32_t addr = (w >> 9) + base;
    const double d = __hiloint2double((int)xh, 0);
    // Predicated loads, unconditional FMAs: a null record issues no shared-memory wavefront, keeps whatever finite
    // table values its landing registers held, and adds 0 * value.  (An `if` around loads and FMAs is compiled to
    // a branch, which serialises consecutive slots on the load latency.)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        VB_SEG_SETP4
        "@p ld.shared.v2.f64 {%0, %1}, [%5];\n\t"
        "@p ld.shared.v2.f64 {%2, %3}, [%6];\n\t}"
        : "+d"(v[0]), "+d"(v[1]), "+d"(v[2]), "+d"(v[3])
        : "r"(xh), "r"(addr), "r"(addr ^ 64u));
    a[0] = fma(d, v[0], a[0]);
    a[1] = fma(d, v[1], a[1]);
    a[2] = fma(d, v[2], a[2]);
    a[3] = fma(d, v[3], a[3]);
}
// FP64 tables with 8 columns (n_donor <= 8): rows of 64 bytes, 4 lanes x 2 columns, one 16-byte load per lane.
__device__ __forceinline__ void seg_step(uint32_t w, uint32_t base, double (&a)[2], double (&v)[2]) {
    const uint32_t xh = w << 16;
    const uint32_t addr = (w >> 10) + base;
    const double d = __hiloint2double((int)xh, 0);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        VB_SEG_SETP2
        "@p ld.shared.v2.f64 {%0, %1}, [%3];\n\t}"
        : "+d"(v[0]), "+d"(v[1])
        : "r"(xh), "r"(addr));
    a[0] = fma(d, v[0], a[0]);
    a[1] = fma(d, v[1], a[1]);
}
// Fixed point: exact integer accumulation of count * value.
__device__ __forceinline__ void seg_step(uint32_t w, uint32_t base, unsigned long long (&a)[4]) {
    const uint32_t c = w & 0xffffu;
    const uint32_t addr = (w >> 10) + base;
    if (c) {
        uint32_t v0, v1, v2, v3;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(addr));
        a[0] += (unsigned long long)c * v0;
        a[1] += (unsigned long long)c * v1;
        a[2] += (unsigned long long)c * v2;
        a[3] += (unsigned long long)c * v3;
    }
}

// the two halves of a super-step: slots 0, 1 and slots 2, 3 of the lane group; with 4 landing sets every slot of a
// super-step has its own, so the loads of the second half do not wait for the FMAs of the first
template <int H, int NSET>
__device__ __forceinline__ void seg_half_step(uint32_t w0, uint32_t w1, uint32_t base, double (&acc)[4][4], SegScratch64<NSET>& sc) {
    seg_step(w0, base, acc[2 * H], sc.v[(2 * H) % NSET]); seg_step(w1, base, acc[2 * H + 1], sc.v[(2 * H + 1) % NSET]);
}
template <int H, int NSET>
__device__ __forceinline__ void seg_half_step(uint32_t w0, uint32_t w1, uint32_t base, double (&acc)[4][2], SegScratchN<NSET>& sc) {
    seg_step(w0, base, acc[2 * H], sc.v[(2 * H) % NSET]); seg_step(w1, base, acc[2 * H + 1], sc.v[(2 * H + 1) % NSET]);
}
template <int H, int NSET>
__device__ __forceinline__ void seg_half_step(uint32_t w0, uint32_t w1, uint32_t base, unsigned long long (&acc)[4][4], SegScratch32&) {
    seg_step(w0, base, acc[2 * H]); seg_step(w1, base, acc[2 * H + 1]);
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void lds64(uint32_t addr, uint32_t& a, uint32_t& b) {
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr));
}

// 23 warps per SM leave 80 registers per lane (registers are handed out in units of 512 per warp: 88 would count as 96).
// NSET = 2: that budget, two landing sets.  NSET = 4: launches with at most VB_SEG_FEW_WARPS consumer warps per CTA
// (matrices with few owner rows per SM: GT-given fits on mid-sized data, one rank's share of a cell-sharded fit) are
// bound by the dependent chain of ONE warp, not by registers -- every slot of a super-step gets its own landing set.
template <int PREC, int NSET>
__global__ void __launch_bounds__((NSET == 2 ? VB_SEG_MAX_WARPS + 1 : VB_SEG_FEW_WARPS + 2) * 32, 1)     // NSET 4: one warp of margin
k_seg(const SegView sv, const EmP p, const SegArgs sa) {
    using Cfg = SegCfg<PREC>;
    constexpr int LPO = Cfg::LPO, M = Cfg::LPO, NC = Cfg::NC, ROWB = Cfg::ROWB, CH = VB_SEG_DEPTH;
    constexpr uint32_t CHB = CH * VB_SEG_OWNERS * 4;             // bytes of a record chunk
    typedef typename std::conditional<PREC == 1, unsigned long long, double>::type acc_t;
    extern __shared__ __align__(1024) unsigned char smem[];
    const int b = blockIdx.y;
    if (p.ctrl && p.ctrl[b * VB_CTRL_N]) return;
    const bool do_theta = sa.mode == GM_SNP && !p.bmm && vb_theta_on(p, b, sa.theta_mode);
    if (sa.mode == GM_SNP && !p.bmm && !do_theta && !p.learn_gt && sa.theta_mode == 2) {   // S1/S2 unused this iteration
        snp_pass_tail(p, b, sa.theta_mode, false);
        return;
    }

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int NW = sa.nwarps, NB = sv.nb;
    const int grid = (int)gridDim.x, cta = (int)blockIdx.x;
    const uint32_t win_bytes = (uint32_t)sv.win_rows * ROWB;
    // dynamic shared memory: full[NB] and empty[NB] window barriers, two record-chunk barriers per consumer warp, the
    // block-reduction scratch, (4 KB in) the window ring, and behind it two record chunks per consumer warp
    const uint32_t bars = smem_u32(smem);                         // warp-uniform: lives in the uniform register file
    const uint32_t ring = bars + VB_SG_RING_OFF;
    double* red = reinterpret_cast<double*>(smem + VB_SG_RED_OFF);

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
        for (int i = 0; i < 2 * NW; ++i) mbar_init(bars + VB_SG_RECBAR_OFF + 8 * i, 1);
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
                if (use > 0) VB_SEG_PWAIT(bars + 8 * (NB + bi), (use - 1) & 1);   // every consumer released the previous fill
#ifdef VB_SEG_CANARY
                // Protocol canary (build.py variants "canary" / "plainfill"): a released buffer is overwritten with NaNs
                // before it is refilled.  A consumer that read a row before its window's fill had completed, or after it
                // had released the window, would pick up a NaN, which no later step can remove from its sums -- so
                // NaN-free, bit-identical results under this build show that no such access happens.
                {
                    unsigned long long* dst = reinterpret_cast<unsigned long long*>(smem + VB_SG_RING_OFF + (size_t)bi * win_bytes);
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
                    uint4* dst = reinterpret_cast<uint4*>(smem + VB_SG_RING_OFF + (size_t)bi * win_bytes);
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
            // a warp that advances near the end of the table waits for windows past it: complete their barriers empty
            for (int wd = 0; wd < sv.span; ++wd) {
                if (use > 0) VB_SEG_PWAIT(bars + 8 * (NB + bi), (use - 1) & 1);
                if (lane == 0) mbar_arrive(bars + 8 * bi);
                __syncwarp();
                if (++bi == NB) { bi = 0; ++use; }
            }
        }
    } else if (task >= 0) {
        // ---------------- consumer warp: 32 owner slots, LPO lanes per slot
        const int64_t s0 = sv.task_off[task], s1 = sv.task_off[task + 1];
        typename std::conditional<PREC == 0, SegScratch64<NSET>, typename std::conditional<PREC == 1, SegScratch32, SegScratchN<NSET>>::type>::type sc;
        if constexpr (PREC == 0) {
#pragma unroll
            for (int i = 0; i < 4 * NSET; ++i) sc.v[i >> 2][i & 3] = 0.0;
        }
        if constexpr (PREC == 2) {
#pragma unroll
            for (int i = 0; i < 2 * NSET; ++i) sc.v[i >> 1][i & 1] = 0.0;
        }
        // this lane's granule of ring row 0, minus the constant upper bits of the count code (see seg_step); FP64
        // rows: odd lane groups read the upper half of a row first
        uint32_t base = ring + (uint32_t)sub * 16 + (PREC == 0 ? (uint32_t)(g & 1) * 64 : 0u) - (PREC == 0 ? 32u : (PREC == 2 ? 16u : 0u));
        asm volatile("" : "+r"(base));
        const int span = sv.span;
        // The warp holds `span` consecutive windows.  rs: buffer of the oldest one, with the fill parity of that buffer's
        // current window in bit 31.  An advance releases the oldest window and waits for the one behind the newest.
        for (int wd = 0; wd < span; ++wd) mbar_wait(bars + 8 * wd, 0);
        uint32_t rs = 0;
        auto advance = [&]() {
            __syncwarp();                                 // every lane's reads of the oldest window are done
            const uint32_t ri = rs & 0xffffu;
            uint32_t ni = ri + (uint32_t)span, np = rs >> 31;
            if (ni >= (uint32_t)NB) { ni -= (uint32_t)NB; np ^= 1u; }
            if (lane == 0) mbar_arrive(bars + 8 * ((uint32_t)NB + ri));
            VB_SEG_CWAIT(bars + 8 * ni, np);
            if ((++rs & 0xffffu) == (uint32_t)NB) rs = (rs & 0x80000000u) ^ 0x80000000u;
        };
        // The record stream of the task arrives in chunks of CH super-steps (512 bytes), bulk-copied into this warp's
        // two chunk buffers two chunks ahead of their use: a record costs the warp one shared-memory load and no
        // register queue.  Chunk gi sits in buffer gi & 1; its barrier completes with parity (gi >> 1) & 1.
        const int ngroups = (int)((s1 - s0) / CH);
        const uint32_t chunk0 = (uint32_t)(s0 / CH);             // chunk index in the whole stream (512-byte units)
        const unsigned char* grec = reinterpret_cast<const unsigned char*>(sv.rec);
        const uint32_t rbuf = ring + (uint32_t)NB * win_bytes + (uint32_t)w * (2 * CHB);
        const uint32_t rbar = bars + VB_SG_RECBAR_OFF + (uint32_t)w * 16;
        if (lane == 0) {
            for (int j = 0; j < 2 && j < ngroups; ++j) {
                mbar_expect_tx(rbar + 8 * j, CHB);
                bulk_g2s(rbuf + j * CHB, grec + (size_t)(chunk0 + j) * CHB, CHB, rbar + 8 * j);
            }
        }
        for (int gi = 0; gi < ngroups; ++gi) {
            const uint32_t buf = (uint32_t)gi & 1u;
            mbar_wait(rbar + 8 * buf, ((uint32_t)gi >> 1) & 1u);
            uint32_t cb = rbuf + buf * CHB + (uint32_t)g * 16;
            asm volatile("" : "+r"(cb));                         // one register per chunk, not recomputed per load
            uint4 c = lds128(cb);
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                const uint32_t adv = c.x >> 28;
                if (adv) {
                    uint32_t a = adv;
                    do advance(); while (--a);
                    c.x &= 0x0fffffffu;
                }
                // the records of the next super-step are loaded half by half as soon as this one's are decoded
                seg_half_step<0, NSET>(c.x, c.y, base, acc, sc);
                if (i + 1 < CH) lds64(cb + (i + 1) * (VB_SEG_OWNERS * 4), c.x, c.y);
                seg_half_step<1, NSET>(c.z, c.w, base, acc, sc);
                if (i + 1 < CH) lds64(cb + (i + 1) * (VB_SEG_OWNERS * 4) + 8, c.z, c.w);
            }
            __syncwarp();                                 // every lane has read the chunk: refill its buffer
#ifdef VB_SEG_CANARY
            // protocol canary: the released chunk is wiped (null records) before its refill; a lane that read it too
            // late, or the next chunk too early, would lose pairs and change the results
            for (uint32_t e = lane; e < CHB / 4; e += 32)
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(rbuf + buf * CHB + 4 * e), "r"(0u) : "memory");
            __syncwarp();
#endif
            if (lane == 0 && gi + 2 < ngroups) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(rbar + 8 * buf, CHB);
                bulk_g2s(rbuf + buf * CHB, grec + (size_t)(chunk0 + (uint32_t)gi + 2u) * CHB, CHB, rbar + 8 * buf);
            }
        }
        for (int a = sv.tail[task]; a > 0; --a) advance();       // release the windows still held
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
    // FP64 tables: the records carry 2 * count;  fixed point: undo the table scale
    const double unq = PREC != 1 ? 0.5 : (sa.mode == GM_SNP ? 1.0 / 4294967295.0 : -1.0 / p.qscale[b]);

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
            for (int c = 0; c < NC; ++c) v[c] = (double)acc[mi][c] * unq;
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
                    double* __restrict__ O = sa.plain_out + (size_t)b * sa.plain_stride + (size_t)owner * sa.plain_ld + sa.plain_off;
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
// row-split cell pass: finish kernel.  partial[r][b][cell][RW] holds the sums of table-row range r (written by k_seg
// in GM_PLAIN mode); this kernel adds them and does what the cell-mode epilogue of k_seg does: softmax over the donors
// with the ID prior, ID_prob, logLik_ID, the padded-row copy for the SNP pass, the LB_p / KL_ID block partials
// (vireoSNP/utils/vireo_model.py:198-201,236-237), then the fused tail.  mode 1: logLik + partials from the present ID_prob.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(VB_THREADS)
k_cell_finish(const EmP p, const double* __restrict__ partial, int R, int mode) {
    const int b = blockIdx.y;
    if (p.ctrl && p.ctrl[b * VB_CTRL_N]) return;
    __shared__ double sh[2 * VB_WARPS];
    const int K = p.K, KT = p.KT, RW = p.RW;
    const size_t range_stride = (size_t)p.B * p.C * RW;
    double r0 = 0.0, r1 = 0.0;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < p.C; j += (int64_t)gridDim.x * blockDim.x) {
        const double* __restrict__ src = partial + ((size_t)b * p.C + j) * RW;
        double v[VB_ROW_DOUBLES];
#pragma unroll
        for (int k = 0; k < VB_ROW_DOUBLES; ++k) v[k] = 0.0;
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int k = 0; k < VB_ROW_DOUBLES; ++k)
                if (k < K) v[k] += src[(size_t)r * range_stride + k];
        const size_t prow = (size_t)(p.id_rows == 1 ? 0 : j) * K;
        double* __restrict__ Rj = p.R + ((size_t)b * p.C + j) * K;
        double* __restrict__ LL = p.ll + ((size_t)b * p.C + j) * K;
        double pr[VB_ROW_DOUBLES];
        if (mode == 0) {
            double mx = -INFINITY;
#pragma unroll
            for (int k = 0; k < VB_ROW_DOUBLES; ++k) { pr[k] = k < K ? v[k] + p.lidp[prow + k] : -INFINITY; mx = fmax(mx, pr[k]); }
            double z = 0.0;
#pragma unroll
            for (int k = 0; k < VB_ROW_DOUBLES; ++k) { pr[k] = pr[k] == -INFINITY ? 0.0 : exp(pr[k] - mx); z += pr[k]; }
#pragma unroll
            for (int k = 0; k < VB_ROW_DOUBLES; ++k) pr[k] = pr[k] / z;
            double* __restrict__ RP = p.RP + ((size_t)b * p.C + j) * RW;
#pragma unroll
            for (int c = 0; c < VB_ROW_DOUBLES; ++c)
                if (c < RW) RP[c] = (c % KT) < K ? pr[c % KT] : 0.0;
        } else {
#pragma unroll
            for (int k = 0; k < VB_ROW_DOUBLES; ++k) pr[k] = k < K ? Rj[k] : 0.0;
        }
#pragma unroll
        for (int k = 0; k < VB_ROW_DOUBLES; ++k) {
            if (k < K) {
                LL[k] = v[k];
                if (mode == 0) Rj[k] = pr[k];
                r0 += v[k] * pr[k];
                if (pr[k] > 0.0) r1 += pr[k] * (log(pr[k]) - p.lidp_kl[prow + k]);
            }
        }
    }
    r0 = warp_sum(r0);
    r1 = warp_sum(r1);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { sh[2 * w] = r0; sh[2 * w + 1] = r1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
        for (int i = 0; i < VB_WARPS; ++i) { a += sh[2 * i]; c += sh[2 * i + 1]; }
        double* out = p.part + (size_t)b * p.part_stride + p.off_cell + 2 * blockIdx.x;
        out[0] = a;
        out[1] = c;
    }
    if (mode == 0) cell_pass_tail(p, b);
}

// ---------------------------------------------------------------------------------------------
// host: row-split cell pass
// ---------------------------------------------------------------------------------------------
// R = how many launches side by side fill the SMs with 22-warp CTAs, provided every range keeps enough windows (the
// partial sums live in the caller's `heavy` workspace, which ws_for sizes for them).  VIREO_B200_SEG_SPLIT=0 turns it off.
// Measured on one rank's share of cfg3 (scripts/time_shard.py): 4 ranks 0.546 -> 0.468 ms per iteration, 8 ranks 0.471 -> 0.363.
int vb_seg_split_rule(const vb_counts* m, int prec) {
    static const int on = env_int("VIREO_B200_SEG_SPLIT", 1);
    if (!on || prec == 1 || m->C < 1) return 1;
    int win_rows, nb, span;
    seg_window(prec, &win_rows, &nb, &span);
    const int64_t tasks = (m->C + VB_SEG_OWNERS - 1) / VB_SEG_OWNERS;
    const int64_t n_win = (2 * m->V + win_rows - 1) / win_rows;
    int64_t R = (int64_t)m->sm_count * VB_SEG_MAX_WARPS / (tasks > 0 ? tasks : 1);
    if (R > VB_SEG_MAX_SPLIT) R = VB_SEG_MAX_SPLIT;
    while (R > 1 && n_win / R < 4 * span) --R;
    return R < 3 ? 1 : (int)R;      // two launches do not pay for the fork / join and the finish kernel (measured at cfg4 and on half of cfg3)
}

static void seg_split_free(SegSplit& sp) {
    for (int r = 0; r < VB_SEG_MAX_SPLIT; ++r) {
        seg_set_free(sp.set[r]);
        if (sp.aux[r]) cudaStreamDestroy(sp.aux[r]);
        if (sp.join[r]) cudaEventDestroy(sp.join[r]);
    }
    if (sp.fork) cudaEventDestroy(sp.fork);
    memset(&sp, 0, sizeof(sp));
}

int vb_seg_build_split(vb_counts* m, int prec, cudaStream_t st) {
    SegSplit& sp = m->rA[prec];
    if (sp.R > 0) return VB_OK;
    if (sp.failed) return VB_E_UNSUPPORTED;
    const int R = vb_seg_split_rule(m, prec);
    if (R < 2) { sp.failed = 1; return VB_E_UNSUPPORTED; }
    DeviceGuard dg(m->device);
    int win_rows, nb, span;
    seg_window(prec, &win_rows, &nb, &span);
    const int64_t n_win = (2 * m->V + win_rows - 1) / win_rows;
    const int64_t wpr = (n_win + R - 1) / R;
    int rc = VB_OK;
    for (int r = 0; r <= R; ++r) sp.row_lo[r] = std::min<int64_t>((int64_t)r * wpr * win_rows, 2 * m->V);
    for (int r = 0; r < R && !rc; ++r) {
        SegBuildOpts ro;
        ro.g_lo = sp.row_lo[r]; ro.g_hi = sp.row_lo[r + 1]; ro.max_grid = m->sm_count / R;
        rc = seg_build_one<0>(m, sp.set[r], prec, st, ro);
        if (!rc && sp.set[r].n_heavy > 0) { vb_set_error("row-split cell pass: residual pairs"); rc = VB_E_UNSUPPORTED; }   // the partial sums use the residual workspace
    }
    for (int r = 0; r < R && !rc; ++r) {
        if (cudaStreamCreateWithFlags(&sp.aux[r], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&sp.join[r], cudaEventDisableTiming) != cudaSuccess) rc = VB_E_CUDA;
    }
    if (!rc && cudaEventCreateWithFlags(&sp.fork, cudaEventDisableTiming) != cudaSuccess) rc = VB_E_CUDA;
    if (rc) {
        seg_split_free(sp);
        sp.failed = 1;
        cudaGetLastError();
        return rc;
    }
    sp.R = R;
    return VB_OK;
}

int vb_seg_launch_cell_split(const vb_counts* m, const EmP& p, int mode, cudaStream_t st) {
    const int prec = p.RW == 8 ? 2 : 0;
    const SegSplit& sp = m->rA[prec];
    if (sp.R < 2 || p.tiled != 2) { vb_set_error("row-split cell pass was not built"); return VB_E_ARG; }
    const size_t range_stride = (size_t)p.B * p.C * p.RW;
    VB_CUDA(cudaEventRecord(sp.fork, st));
    for (int r = 0; r < sp.R; ++r) {
        VB_CUDA(cudaStreamWaitEvent(sp.aux[r], sp.fork, 0));
        SegPlain pl;
        pl.out = p.H + (size_t)r * range_stride; pl.ld = p.RW; pl.off = 0; pl.cols = p.RW;
        pl.set = &sp.set[r]; pl.row0 = sp.row_lo[r];
        const int rc = vb_seg_launch(m, p, 0, GM_PLAIN, 0, &pl, sp.aux[r]);
        if (rc) return rc;
        VB_CUDA(cudaEventRecord(sp.join[r], sp.aux[r]));
    }
    for (int r = 0; r < sp.R; ++r) VB_CUDA(cudaStreamWaitEvent(st, sp.join[r], 0));
    VB_LAUNCH(3, st, k_cell_finish<<<dim3((unsigned)p.n_cellblk, p.B), VB_THREADS, 0, st>>>(p, p.H, sp.R, mode));
    VB_CUDA(cudaGetLastError());
    return VB_OK;
}

// split owners: the plain sums of an owner's virtual parts (format `excess`) are added to its residual sums H[o], which
// the epilogue of the owner's own format adds to its accumulators.  One thread per (first virtual part, column); the
// parts of an owner are consecutive, so the order of the additions is fixed.
__global__ void __launch_bounds__(VB_THREADS)
k_seg_fold(int64_t n_virtual, const int32_t* __restrict__ v_owner, const uint8_t* __restrict__ v_part, const uint8_t* __restrict__ o_split,
           const double* __restrict__ vsum, double* __restrict__ H, int64_t n_owner, int RW, const int* __restrict__ ctrl) {
    const int b = blockIdx.y;
    if (ctrl && ctrl[b * VB_CTRL_N]) return;
    const int64_t n = n_virtual * RW;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = e / RW;
        const int c = (int)(e % RW);
        if (v_part[v] != 1) continue;
        const int64_t o = v_owner[v];
        const int extra = (int)o_split[o] - 1;
        double* h = H + ((size_t)b * n_owner + o) * RW + c;
        double t = *h;
        for (int j = 0; j < extra; ++j) t += vsum[((size_t)b * n_virtual + v + j) * RW + c];
        *h = t;
    }
}

// ---------------------------------------------------------------------------------------------
// host: dispatch
// ---------------------------------------------------------------------------------------------
static SegView view_of_set(const SegSet& g) {
    SegView v;
    v.n_owner = g.n_owner; v.n_gather = g.n_gather; v.n_task = g.n_task; v.n_task_stream = g.n_task_stream;
    v.n_win = g.n_win; v.win_rows = g.win_rows; v.nb = g.nb; v.span = g.span;
    v.perm = g.perm; v.task_off = g.task_off; v.tail = g.tail; v.rec = g.rec;
    v.hptr = g.hptr; v.hrow = g.hrow; v.hcnt = g.hcnt;
    return v;
}

static size_t seg_smem(int prec, int nb, int win_rows) {
    static_assert(2 * VB_SEG_MAX_NB * 8 <= VB_SG_RECBAR_OFF && VB_SG_RECBAR_OFF + VB_SEG_MAX_WARPS * 16 <= VB_SG_RED_OFF &&
                  VB_SG_RED_OFF + (VB_SEG_MAX_WARPS + 1) * VB_SG_RED_DOUBLES * 8 <= VB_SG_RING_OFF, "shared-memory layout of k_seg");
    return VB_SG_RING_OFF + (size_t)nb * win_rows * (prec == 0 ? 128 : 64) + (size_t)VB_SEG_MAX_WARPS * 2 * VB_SEG_DEPTH * VB_SEG_OWNERS * 4;
}

// VIREO_B200_SEG_FEW=0 keeps the two-landing-set instance for every launch
static bool seg_few_on() {
    static const bool on = !(getenv("VIREO_B200_SEG_FEW") && atoi(getenv("VIREO_B200_SEG_FEW")) == 0);
    return on;
}

static bool g_seg_attr_set[64] = {false};     // per device: function attributes belong to the context
static size_t g_seg_static[64] = {0};         // largest static shared memory of the k_seg instances
#define VB_SMEM_OPTIN ((size_t)227 * 1024)

// ori 0: cell pass (table = p.Wt / p.Wq), ori 1: SNP pass (table = p.RP / p.RPq)
// mode GM_PLAIN (FP64 tables): `plain` says where the plain sums go (a column chunk of the doublet pass, a row range
// of the row-split cell pass, the virtual owners of a format with split owners)
int vb_seg_launch(const vb_counts* m, const EmP& p, int ori, int mode, int theta_mode, const SegPlain* plain, cudaStream_t st) {
    const int prec = p.tiled == 3 ? 1 : (p.RW == 8 ? 2 : 0);
    const SegSet& g = (plain && plain->set) ? *plain->set : (ori ? m->sB[prec] : m->sA[prec]);
    if (!g.built) { vb_set_error("segment format was not built"); return VB_E_ARG; }
    // split owners: the virtual parts first (plain sums of the excess format), folded into H below
    const bool has_excess = g.excess != nullptr && prec != 1;
    if (has_excess) {
        SegSet& gm = const_cast<SegSet&>(g);
        const int64_t need = (int64_t)p.B * g.n_virtual * p.RW;
        if (gm.vsum_elems < need) {
            cudaFree(gm.vsum); gm.vsum = nullptr; gm.vsum_elems = 0;
            VB_CUDA(cudaMalloc(&gm.vsum, (size_t)need * sizeof(double)));
            gm.vsum_elems = need;
        }
        SegPlain pe;
        pe.out = g.vsum; pe.ld = p.RW; pe.off = 0; pe.cols = p.RW; pe.set = g.excess; pe.row0 = 0;
        const int rc_e = vb_seg_launch(m, p, ori, GM_PLAIN, 0, &pe, st);
        if (rc_e) return rc_e;
    }
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
        opt_in((const void*)k_seg<0, 2>); opt_in((const void*)k_seg<1, 2>); opt_in((const void*)k_seg<2, 2>);
        opt_in((const void*)k_seg<0, 4>); opt_in((const void*)k_seg<2, 4>);
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
    sa.mode = mode; sa.theta_mode = theta_mode;
    const double* tab64 = ori ? p.RP : p.Wt;
    const int64_t rows_all = ori ? p.C : 2 * p.V;                  // the table of one restart, whatever part of it `g` covers
    if (prec != 1) {
        sa.table = reinterpret_cast<const unsigned char*>(tab64) + (size_t)((plain && plain->set) ? plain->row0 : 0) * (prec == 0 ? 128 : 64);
        sa.table_stride = rows_all * (prec == 0 ? 128 : 64);
    }
    else { sa.table = reinterpret_cast<const unsigned char*>(ori ? p.RPq : p.Wq); sa.table_stride = rows_all * 64; }
    sa.has_heavy = g.n_heavy > 0 || has_excess;
    if (mode == GM_PLAIN) {
        if (!plain || prec == 1) { vb_set_error("plain segment pass: bad arguments"); return VB_E_ARG; }
        sa.plain_out = plain->out; sa.plain_ld = plain->ld; sa.plain_off = plain->off; sa.plain_cols = plain->cols;
        sa.plain_stride = plain->set ? g.n_owner * plain->ld : 0;         // a format of its own: one block of sums per restart
    }
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
        if (has_excess) {
            int64_t fb = (g.n_virtual * p.RW + VB_THREADS - 1) / VB_THREADS;
            if (fb > (int64_t)m->sm_count * 8) fb = (int64_t)m->sm_count * 8;
            VB_LAUNCH(cls, st, k_seg_fold<<<dim3((unsigned)fb, p.B), VB_THREADS, 0, st>>>(
                g.n_virtual, g.v_owner, g.v_part, g.o_split, g.vsum, p.H, g.n_owner, p.RW, p.ctrl));
            VB_CUDA(cudaGetLastError());
        }
    }
    const dim3 grid(grid_x, p.B);
    const int threads = (sa.nwarps + 1) * 32;
    VB_LAUNCH(cls, st, {
        const bool few = sa.nwarps <= VB_SEG_FEW_WARPS && seg_few_on();
        if (prec == 2) { if (few) k_seg<2, 4><<<grid, threads, smem, st>>>(sv, p, sa); else k_seg<2, 2><<<grid, threads, smem, st>>>(sv, p, sa); }
        else if (prec == 0) { if (few) k_seg<0, 4><<<grid, threads, smem, st>>>(sv, p, sa); else k_seg<0, 2><<<grid, threads, smem, st>>>(sv, p, sa); }
        else k_seg<1, 2><<<grid, threads, smem, st>>>(sv, p, sa);
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
