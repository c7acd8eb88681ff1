// vb_gather.cu -- the ring-slab gather kernels: both sparse passes of the EM iteration with the
// gathered table staged through shared memory by bulk async copies (TMA, cp.async.bulk + mbarrier).
//
// Both passes are "for every owner: sum over its records of count * table[gather row][0:K]":
//   cell pass  logLik_ID[j,:] = sum_i  (dp-ad)_ij * W[2i,:] + ad_ij * W[2i+1,:]
//              (vireoSNP/utils/vireo_model.py:190-196, bmm_model.py:125-129)
//   SNP  pass  S2[i,:] = sum_j (dp-ad)_ij * ID_prob[j,:],  S1[i,:] = sum_j ad_ij * ID_prob[j,:]
//              (vireo_model.py:168-170,207-209, bmm_model.py:136-138)
// In FP64 with K = 16 donors a table row is 128 bytes; gathering it from L2 per nnz (the v1 row
// kernels in vb_em.cu) is latency bound at ~3% of the HBM roofline.  Here
//   * one LANE owns one owner row and keeps its K accumulators in registers,
//   * the gather table streams through a ring of VB_RING_SLABS x VB_SLAB_ROWS rows in shared memory,
//     filled slab by slab by a producer warp with cp.async.bulk (completion on an mbarrier),
//   * every lane walks its own delta-coded record stream (16 bits per record, vb_common.cuh) and
//     consumes a record as soon as its gather row is resident, so lanes drift inside the ring window
//     instead of synchronising per tile,
//   * a row is read with 16-byte loads whose granule is XOR-rotated by the lane id, so the 8 lanes of
//     a quarter warp always hit 8 different bank groups whatever rows they address (conflict free).
// The bound of this design is the shared-memory crossbar (128 B/clk/SM = one K=16 record per clock
// per SM), not HBM: the record streams are 2 B per nnz-allele pair.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "vb_common.cuh"
#include "vb_stream.cuh"

__device__ __forceinline__ uint32_t n_skips(int delta) {
    return delta > VB_REC_MAX_DELTA ? (uint32_t)((delta - 1) / VB_REC_MAX_DELTA) : 0u;
}

// pass 1: stream length (records incl. skips), light pairs and heavy pairs per owner
template <int ORI, bool WIDE>
__global__ void k_gs_count(const CountsView m, int64_t n_owner, uint32_t* __restrict__ len, uint32_t* __restrict__ n_light,
                           uint32_t* __restrict__ n_heavy, unsigned int* flags) {
    for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < n_owner; o += (int64_t)gridDim.x * blockDim.x) {
        uint32_t n = 0, nl = 0, nh = 0;
        int prev = 0;
        const bool ok = for_records<ORI, WIDE>(m, o, [&](int g, uint32_t c) {
            if (c > VB_REC_MAX_COUNT) { ++nh; return; }
            n += n_skips(g - prev) + 1;
            ++nl;
            prev = g;
        });
        if (!ok) atomicOr(&flags[0], 1u);
        len[o] = n;
        n_light[o] = nl;
        n_heavy[o] = nh;
    }
}

// owners routed to the residual kernel entirely (sorted ranks >= first_sparse): their stream is empty
__global__ void k_gs_mark_sparse(const int32_t* __restrict__ perm_sorted, int64_t first_sparse, int64_t n_owner,
                                 uint8_t* __restrict__ sparse, uint32_t* __restrict__ len_sorted, uint32_t* __restrict__ len,
                                 uint32_t* __restrict__ n_light, uint32_t* __restrict__ n_heavy) {
    for (int64_t r = first_sparse + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_owner; r += (int64_t)gridDim.x * blockDim.x) {
        const int32_t o = perm_sorted[r];
        sparse[o] = 1;
        len_sorted[r] = 0;
        len[o] = 0;
        n_heavy[o] += n_light[o];
        n_light[o] = 0;
    }
}

__global__ void k_gs_iota(int32_t* out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (int32_t)i;
}

// blocks of each warp slot: owners are sorted by stream length, descending, so lane 0 holds the maximum
__global__ void k_gs_slot_blocks(const uint32_t* __restrict__ len_sorted, int64_t n_slot, uint32_t* __restrict__ nblk) {
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s <= n_slot; s += (int64_t)gridDim.x * blockDim.x)
        nblk[s] = s < n_slot ? (len_sorted[s * 32] + 3u) / 4u : 0u;
}

// pass 2: write the streams and the residual CSR; one thread per sorted position
template <int ORI, bool WIDE>
__global__ void k_gs_fill(const CountsView m, int64_t n_pos, const int32_t* __restrict__ perm,
                          const uint32_t* __restrict__ slot_blk, uint16_t* __restrict__ rec,
                          const int64_t* __restrict__ hptr, int32_t* __restrict__ hrow, uint32_t* __restrict__ hcnt,
                          const uint8_t* __restrict__ sparse) {
    for (int64_t pos = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; pos < n_pos; pos += (int64_t)gridDim.x * blockDim.x) {
        const int64_t o = perm[pos];
        if (o < 0) continue;
        const bool all_residual = sparse[o] != 0;
        const int lane = (int)(pos & 31);
        uint16_t* base = rec + (size_t)slot_blk[pos >> 5] * 128 + lane * 4;
        uint32_t t = 0;
        int prev = 0;
        int64_t h = hptr[o];
        for_records<ORI, WIDE>(m, o, [&](int g, uint32_t c) {
            if (c > VB_REC_MAX_COUNT || all_residual) { hrow[h] = g; hcnt[h] = c; ++h; return; }
            int d = g - prev;
            while (d > VB_REC_MAX_DELTA) {
                base[(size_t)(t >> 2) * 128 + (t & 3)] = (uint16_t)(VB_REC_MAX_DELTA << VB_REC_COUNT_BITS);
                ++t;
                d -= VB_REC_MAX_DELTA;
            }
            base[(size_t)(t >> 2) * 128 + (t & 3)] = (uint16_t)((d << VB_REC_COUNT_BITS) | c);
            ++t;
            prev = g;
        });
    }
}

__global__ void k_gs_sum3(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, const uint32_t* __restrict__ c,
                          int64_t n, unsigned long long* out) {
    unsigned long long sa = 0, sb = 0, sc = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        sa += a[i]; sb += b[i]; sc += c[i];
    }
    for (int off = 16; off > 0; off >>= 1) {
        sa += __shfl_xor_sync(VB_FULL, sa, off);
        sb += __shfl_xor_sync(VB_FULL, sb, off);
        sc += __shfl_xor_sync(VB_FULL, sc, off);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(out, sa); atomicAdd(out + 1, sb); atomicAdd(out + 2, sc); }
}

__global__ void k_gs_widen(const uint32_t* __restrict__ in, int64_t n, int64_t* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i <= n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = i < n ? (int64_t)in[i] : 0;
}

static void gather_set_free(GatherSet& g) {
    cudaFree(g.perm); cudaFree(g.len); cudaFree(g.slot_blk); cudaFree(g.rec);
    cudaFree(g.hptr); cudaFree(g.hrow); cudaFree(g.hcnt); cudaFree(g.cta_start);
    memset(&g, 0, sizeof(g));
}

static CountsView gs_view(const vb_counts* m) {
    CountsView v;
    v.C = m->C; v.V = m->V; v.N = m->N;
    v.cell_ptr = m->cell_ptr; v.cell_idx = m->cell_idx; v.cell_cnt = m->cell_cnt; v.cell_dp = m->cell_dp;
    v.snp_ptr = m->snp_ptr; v.snp_idx = m->snp_idx; v.snp_cnt = m->snp_cnt; v.snp_dp = m->snp_dp;
    return v;
}

template <int ORI>
static int build_one(vb_counts* m, GatherSet& g, cudaStream_t st) {
    memset(&g, 0, sizeof(g));
    const int sm = m->sm_count;
    const int64_t O = ORI == 0 ? m->C : 2 * m->V;
    const int64_t Gn = ORI == 0 ? 2 * m->V : m->C;
    if (O >= (1ll << 31) - 64 || Gn >= (1ll << 23)) { vb_set_error("gather format: more than 2^31 owner rows or 2^23 gather rows"); return VB_E_UNSUPPORTED; }
    const int64_t n_slot = (O + 31) / 32;
    const int64_t n_pos = n_slot * 32;
    const CountsView v = gs_view(m);
    GsScratch tmp;
    int rc;
    uint32_t *len, *nl, *nh, *len_sorted, *nblk;
    int32_t *ids, *perm_sorted;
    unsigned int* flags;
    unsigned long long* sums;
    if ((rc = tmp.alloc(&len, O)) || (rc = tmp.alloc(&nl, O)) || (rc = tmp.alloc(&nh, O + 1)) ||
        (rc = tmp.alloc(&len_sorted, n_pos)) || (rc = tmp.alloc(&nblk, n_slot + 1)) || (rc = tmp.alloc(&ids, O)) ||
        (rc = tmp.alloc(&perm_sorted, O)) || (rc = tmp.alloc(&flags, 4)) || (rc = tmp.alloc(&sums, 4)))
        return rc;
    VB_CUDA(cudaMemsetAsync(flags, 0, 4 * sizeof(unsigned int), st));
    VB_CUDA(cudaMemsetAsync(sums, 0, 4 * sizeof(unsigned long long), st));
    VB_CUDA(cudaMemsetAsync(len_sorted, 0, n_pos * sizeof(uint32_t), st));
    if (O) {
        if (m->wide) k_gs_count<ORI, true><<<grid1d(O, sm), 256, 0, st>>>(v, O, len, nl, nh, flags);
        else k_gs_count<ORI, false><<<grid1d(O, sm), 256, 0, st>>>(v, O, len, nl, nh, flags);
        VB_CUDA(cudaGetLastError());
        k_gs_sum3<<<grid1d(O, sm), 256, 0, st>>>(len, nl, nh, O, sums);
        VB_CUDA(cudaGetLastError());
        k_gs_iota<<<grid1d(O, sm), 256, 0, st>>>(ids, O);
        VB_CUDA(cudaGetLastError());
    }
    unsigned int hflags[4];
    unsigned long long hsums[4];
    VB_CUDA(cudaMemcpyAsync(hflags, flags, sizeof(hflags), cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaMemcpyAsync(hsums, sums, sizeof(hsums), cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaStreamSynchronize(st));
    if (hflags[0]) { vb_set_error("gather format: an entry has AD > DP"); return VB_E_VALUE; }
    g.n_owner = O; g.n_gather = Gn; g.n_slot = n_slot;
    g.n_rec = (int64_t)hsums[0]; g.n_light = (int64_t)hsums[1]; g.n_heavy = (int64_t)hsums[2];

    // owners sorted by stream length, descending -> lanes of a warp slot do equal work
    VB_CUDA(cudaMalloc(&g.perm, (n_pos ? n_pos : 1) * sizeof(int32_t)));
    VB_CUDA(cudaMalloc(&g.len, (n_pos ? n_pos : 1) * sizeof(uint32_t)));
    VB_CUDA(cudaMalloc(&g.slot_blk, (n_slot + 1) * sizeof(uint32_t)));
    VB_CUDA(cudaMemsetAsync(g.perm, 0xff, (n_pos ? n_pos : 1) * sizeof(int32_t), st));
    uint8_t* sparse;
    if ((rc = tmp.alloc(&sparse, O))) return rc;
    VB_CUDA(cudaMemsetAsync(sparse, 0, O ? O : 1, st));
    std::vector<uint32_t> hlen((size_t)n_pos, 0u);
    if (O) {
        size_t tb = 0;
        VB_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, tb, len, len_sorted, ids, perm_sorted, O, 0, 32, st));
        void* cub_tmp;
        if ((rc = tmp.alloc((unsigned char**)&cub_tmp, tb))) return rc;
        VB_CUDA(cub::DeviceRadixSort::SortPairsDescending(cub_tmp, tb, len, len_sorted, ids, perm_sorted, O, 0, 32, st));
        VB_CUDA(cudaMemcpyAsync(g.perm, perm_sorted, O * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
        VB_CUDA(cudaMemcpyAsync(hlen.data(), len_sorted, O * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        VB_CUDA(cudaStreamSynchronize(st));
        // Owners with a short stream ride the front edge of the ring and spend a whole warp step on one or two
        // records; when such owners carry a negligible share of the records (e.g. the alternative-allele rows of
        // homozygous-reference SNPs) their pairs go to the residual kernel instead.
        int64_t first_sparse = O, moved = 0;
        const int64_t budget = g.n_rec / 50;
        while (first_sparse > 0 && hlen[first_sparse - 1] <= VB_SPARSE_LEN && moved + hlen[first_sparse - 1] <= budget) {
            moved += hlen[first_sparse - 1];
            --first_sparse;
        }
        while (first_sparse < O && hlen[first_sparse] == 0) ++first_sparse;   // nothing to move for empty streams
        if (moved > 0) {
            k_gs_mark_sparse<<<grid1d(O - first_sparse, sm), 256, 0, st>>>(perm_sorted, first_sparse, O, sparse, len_sorted, len, nl, nh);
            VB_CUDA(cudaGetLastError());
            for (int64_t r = first_sparse; r < O; ++r) hlen[r] = 0;
            VB_CUDA(cudaMemsetAsync(sums, 0, 4 * sizeof(unsigned long long), st));
            k_gs_sum3<<<grid1d(O, sm), 256, 0, st>>>(len, nl, nh, O, sums);
            VB_CUDA(cudaGetLastError());
            VB_CUDA(cudaMemcpyAsync(hsums, sums, sizeof(hsums), cudaMemcpyDeviceToHost, st));
            VB_CUDA(cudaStreamSynchronize(st));
            g.n_rec = (int64_t)hsums[0]; g.n_light = (int64_t)hsums[1]; g.n_heavy = (int64_t)hsums[2];
        }
    }
    VB_CUDA(cudaMemcpyAsync(g.len, len_sorted, (n_pos ? n_pos : 1) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    {
        // Warp slots are dealt alternately from the long and the short end of the sorted list (position p ->
        // slot p/2 or n_slot-1-p/2), and a CTA is a contiguous range of positions sized so that every CTA
        // carries about the same number of records with at most VB_GATHER_MAX_WARPS slots.  Mixing both ends
        // keeps bimodal length distributions (reference- vs alternative-allele rows) within the slot limit.
        std::vector<int32_t> start;
        start.push_back(0);
        const double target = (double)g.n_rec / (double)sm;
        double acc_rec = 0.0, done_rec = 0.0;
        int in_cta = 0;
        for (int64_t s = 0; s < n_slot; ++s) {
            double r = 0.0;
            const int64_t sl = (s & 1) ? n_slot - 1 - (s >> 1) : (s >> 1);
            for (int l = 0; l < 32; ++l) r += hlen[(size_t)sl * 32 + l];
            const int64_t ctas = (int64_t)start.size();                 // CTAs opened so far (incl. the current one)
            const bool full = in_cta >= VB_GATHER_MAX_WARPS;
            // close the current CTA when adding this slot would overshoot its share of what is left
            const bool share = in_cta > 0 && target > 0.0 && ctas < sm && acc_rec + 0.5 * r > (double)ctas * target - done_rec;
            if (full || share) {
                start.push_back((int32_t)s);
                done_rec += acc_rec;
                acc_rec = 0.0;
                in_cta = 0;
            }
            acc_rec += r;
            ++in_cta;
        }
        start.push_back((int32_t)n_slot);
        if (n_slot == 0) { start.clear(); start.push_back(0); start.push_back(0); }
        g.grid = (int)start.size() - 1;
        g.max_warps = 1;
        for (size_t i = 0; i + 1 < start.size(); ++i)
            if (start[i + 1] - start[i] > g.max_warps) g.max_warps = start[i + 1] - start[i];
        VB_CUDA(cudaMalloc(&g.cta_start, start.size() * sizeof(int32_t)));
        VB_CUDA(cudaMemcpyAsync(g.cta_start, start.data(), start.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        VB_CUDA(cudaStreamSynchronize(st));
    }
    k_gs_slot_blocks<<<grid1d(n_slot + 1, sm), 256, 0, st>>>(len_sorted, n_slot, nblk);
    VB_CUDA(cudaGetLastError());
    {
        size_t tb = 0;
        VB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, nblk, g.slot_blk, n_slot + 1, st));
        void* cub_tmp;
        if ((rc = tmp.alloc((unsigned char**)&cub_tmp, tb))) return rc;
        VB_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp, tb, nblk, g.slot_blk, n_slot + 1, st));
    }
    uint32_t total_blk = 0;
    VB_CUDA(cudaMemcpyAsync(&total_blk, g.slot_blk + n_slot, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaStreamSynchronize(st));
    g.n_block = total_blk;
    if ((int64_t)g.n_rec > (int64_t)total_blk * 128) { vb_set_error("gather format: inconsistent block count"); return VB_E_CUDA; }
    if (g.n_rec >= (1ll << 37)) { vb_set_error("gather format: too many records"); return VB_E_UNSUPPORTED; }

    // residual CSR
    VB_CUDA(cudaMalloc(&g.hptr, (O + 1) * sizeof(int64_t)));
    VB_CUDA(cudaMalloc(&g.hrow, (g.n_heavy ? g.n_heavy : 1) * sizeof(int32_t)));
    VB_CUDA(cudaMalloc(&g.hcnt, (g.n_heavy ? g.n_heavy : 1) * sizeof(uint32_t)));
    {
        int64_t* wide;
        if ((rc = tmp.alloc(&wide, O + 1))) return rc;
        k_gs_widen<<<grid1d(O + 1, sm), 256, 0, st>>>(nh, O, wide);
        VB_CUDA(cudaGetLastError());
        size_t tb = 0;
        VB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, wide, g.hptr, O + 1, st));
        void* cub_tmp;
        if ((rc = tmp.alloc((unsigned char**)&cub_tmp, tb))) return rc;
        VB_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp, tb, wide, g.hptr, O + 1, st));
    }
    VB_CUDA(cudaMalloc(&g.rec, ((size_t)total_blk * 128 + 64) * sizeof(uint16_t)));
    VB_CUDA(cudaMemsetAsync(g.rec, 0, ((size_t)total_blk * 128 + 64) * sizeof(uint16_t), st));
    if (O) {
        if (m->wide) k_gs_fill<ORI, true><<<grid1d(n_pos, sm), 256, 0, st>>>(v, n_pos, g.perm, g.slot_blk, g.rec, g.hptr, g.hrow, g.hcnt, sparse);
        else k_gs_fill<ORI, false><<<grid1d(n_pos, sm), 256, 0, st>>>(v, n_pos, g.perm, g.slot_blk, g.rec, g.hptr, g.hrow, g.hcnt, sparse);
        VB_CUDA(cudaGetLastError());
    }
    VB_CUDA(cudaStreamSynchronize(st));
    g.bytes = (int64_t)total_blk * 256 + n_pos * 8 + (n_slot + 1) * 4 + (O + 1) * 8 + g.n_heavy * 8;
    g.built = 1;
    return VB_OK;
}

int vb_gather_build(vb_counts* m, cudaStream_t st) {
    if (m->gA.built && m->gB.built) return VB_OK;
    if (m->gather_failed) return VB_E_UNSUPPORTED;
    VB_CUDA(cudaSetDevice(m->device));
    int rc = build_one<0>(m, m->gA, st);
    if (!rc) rc = build_one<1>(m, m->gB, st);
    if (rc) {
        gather_set_free(m->gA);
        gather_set_free(m->gB);
        m->gather_failed = 1;
        cudaGetLastError();
    }
    return rc;
}

void vb_gather_free(vb_counts* m) {
    gather_set_free(m->gA);
    gather_set_free(m->gB);
}

// ---------------------------------------------------------------------------------------------
// k_heavy: residual pairs with count >= 32, one warp per owner, table rows gathered from L2.
// out[o][0:16] is written for EVERY owner (zeros where there is no residual).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(VB_THREADS)
k_heavy(const GatherView gv, const double* __restrict__ table, int64_t table_stride, double* __restrict__ out,
        int64_t out_stride, const int* __restrict__ ctrl) {
    const int b = blockIdx.y;
    if (ctrl && ctrl[b * VB_CTRL_N]) return;
    const double* __restrict__ T = table + (size_t)b * table_stride;
    double* __restrict__ O = out + (size_t)b * out_stride;
    const int lane = threadIdx.x & 31, kl = lane & 15, sub = lane >> 4;
    const int64_t nw = (int64_t)gridDim.x * VB_WARPS;
    for (int64_t o = (int64_t)blockIdx.x * VB_WARPS + (threadIdx.x >> 5); o < gv.n_owner; o += nw) {
        const int64_t p0 = gv.hptr[o], p1 = gv.hptr[o + 1];
        double acc = 0.0;
        for (int64_t q = p0 + sub; q < p1; q += 2)
            acc = fma((double)gv.hcnt[q], T[(size_t)gv.hrow[q] * VB_ROW_DOUBLES + kl], acc);
        acc += __shfl_xor_sync(VB_FULL, acc, 16);
        if (sub == 0) O[(size_t)o * VB_ROW_DOUBLES + kl] = acc;
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

// ---------------------------------------------------------------------------------------------
// k_gather
// ---------------------------------------------------------------------------------------------
#define VB_G_EPOCH 4
#define VB_G_RING_MAGIC ((uint32_t)((0x100000000ull + VB_RING_ROWS - 1) / VB_RING_ROWS))
#define VB_G_RING_BYTES (VB_RING_ROWS * VB_ROW_DOUBLES * 8)
#define VB_G_SLAB_BYTES (VB_SLAB_ROWS * VB_ROW_DOUBLES * 8)
#define VB_G_RED_DOUBLES (2 * VB_MAX_GT)
// dynamic shared memory: ring | full barriers | progress | reduction scratch
#define VB_G_OFF_BAR VB_G_RING_BYTES
#define VB_G_OFF_LANDED (VB_G_OFF_BAR + VB_RING_SLABS * 8)
#define VB_G_OFF_PROG (VB_G_OFF_LANDED + 16)
#define VB_G_OFF_RED (VB_G_OFF_PROG + 32 * 4)
#define VB_G_SMEM (VB_G_OFF_RED + (VB_GATHER_MAX_WARPS + 1) * VB_G_RED_DOUBLES * 8)

struct GatherArgs {
    int mode;            // GM_*
    int theta_mode;      // GM_SNP: 0 never, 1 always, 2 per the device iteration counter
    int nwarps;          // largest number of consumer warps of a CTA (block size = (nwarps + 1) * 32)
    int has_heavy;       // add p.H[owner] before the epilogue
    int KT;              // accumulators per lane = 2 * NL
    int64_t table_stride;   // doubles per restart of the gather table
    const double* table;
    double* plain_out;   // GM_PLAIN: [n_owner, 16]
};

template <int NL>
__global__ void __launch_bounds__((VB_GATHER_MAX_WARPS + 1) * 32, 1)
k_gather(const GatherView gv, const EmP p, const GatherArgs ga) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int b = blockIdx.y;
    if (p.ctrl && p.ctrl[b * VB_CTRL_N]) return;
    const bool do_theta = ga.mode == GM_SNP && !p.bmm && vb_theta_on(p, b, ga.theta_mode);
    if (ga.mode == GM_SNP && !p.bmm && !do_theta && !p.learn_gt && ga.theta_mode == 2) return;   // S1/S2 unused this iteration

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int slot0 = gv.cta_start[blockIdx.x];
    const int nwarps = gv.cta_start[blockIdx.x + 1] - slot0;      // consumer warps of this CTA; warp `nwarps` produces
    const uint32_t ring = smem_u32(smem);
    const uint32_t bars = ring + VB_G_OFF_BAR;
    volatile int* progress = reinterpret_cast<volatile int*>(smem + VB_G_OFF_PROG);
    double* red = reinterpret_cast<double*>(smem + VB_G_OFF_RED);
    const int nslab = (int)((gv.n_gather + VB_SLAB_ROWS - 1) / VB_SLAB_ROWS);

    if (threadIdx.x < 32) progress[threadIdx.x] = threadIdx.x < nwarps ? 0 : 0x7fffffff;
    if (threadIdx.x == 0) {
        *reinterpret_cast<volatile int*>(smem + VB_G_OFF_LANDED) = 0;
        for (int i = 0; i < VB_RING_SLABS; ++i) mbar_init(bars + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    constexpr int NA = 2 * NL;
    double acc[NA];
#pragma unroll
    for (int i = 0; i < NA; ++i) acc[i] = 0.0;
    int owner = -1;

    if (w == nwarps) {
        // ---------------- producer warp: stream the table through the ring.  It is also the only waiter on
        // the mbarriers (in slab order, so the parity test is exact) and republishes completions as the
        // monotonic counter `landed` that the consumer warps poll.
        const double* __restrict__ T = ga.table + (size_t)b * ga.table_stride;
        int issued = 0, landed = 0;
        bool stop = nslab == 0;
        while (!(stop && landed == issued)) {
            bool did = false;
            if (!stop) {
                const int mn = __reduce_min_sync(VB_FULL, progress[lane]);
                if (mn == 0x7fffffff) stop = true;                      // nobody reads the rest of the table
                else if (mn >= issued - VB_RING_SLABS + 1) {            // slab issued-RING released by every warp
                    if (lane == 0) {
                        const int64_t r0 = (int64_t)issued * VB_SLAB_ROWS;
                        const int64_t rows = gv.n_gather - r0 < VB_SLAB_ROWS ? gv.n_gather - r0 : VB_SLAB_ROWS;
                        const uint32_t bytes = (uint32_t)(rows * VB_ROW_DOUBLES * 8);
                        const int buf = issued % VB_RING_SLABS;
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        mbar_expect_tx(bars + 8 * buf, bytes);
                        bulk_g2s(ring + buf * VB_G_SLAB_BYTES, T + (size_t)r0 * VB_ROW_DOUBLES, bytes, bars + 8 * buf);
                    }
                    ++issued;
                    did = true;
                    if (issued == nslab) stop = true;
                }
            }
            if (landed < issued) {
                if (mbar_test(bars + 8 * (landed % VB_RING_SLABS), (uint32_t)((landed / VB_RING_SLABS) & 1))) {
                    ++landed;
                    did = true;
                    __syncwarp();
                    if (lane == 0) asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(ring + VB_G_OFF_LANDED), "r"(landed) : "memory");
                }
            }
            if (!did) __nanosleep(32);
        }
    } else if (w < nwarps) {
        // ---------------- consumer warp: one owner per lane
        const int64_t pos = (int64_t)slot0 + w;       // dealing position -> warp slot (long and short ends alternate)
        const int64_t slot = (pos & 1) ? gv.n_slot - 1 - (pos >> 1) : (pos >> 1);
        uint32_t rem = 0;
        const uint2* sp = reinterpret_cast<const uint2*>(gv.rec);
        {
            owner = gv.perm[slot * 32 + lane];
            rem = gv.len[slot * 32 + lane];
            sp += (size_t)gv.slot_blk[slot] * 32 + lane;
        }
        uint32_t blk_left = (rem + 3u) >> 2;       // 4 records per 8-byte block; one block is prefetched ahead
        uint2 cur = make_uint2(0, 0), nxt = cur;
        if (blk_left) { cur = __ldg(sp); sp += 32; --blk_left; }
        if (blk_left) { nxt = __ldg(sp); sp += 32; --blk_left; }
        uint32_t phase = 4;                        // records left in `cur`
        int g = 0;                                 // gather row of the last consumed record
        int released = 0, landed_seen = -1;
        const uint32_t lane_base = ring + ((uint32_t)(lane & 7) << 4);   // ring is 1024-byte aligned
        const uint32_t prog_addr = smem_u32((const void*)(progress + w));
        for (;;) {
            // ---- bookkeeping, once per VB_G_EPOCH steps: landed edge, slabs every lane has passed
            int ld;
            asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(ld) : "r"(ring + VB_G_OFF_LANDED) : "memory");
            const int hi = ld * VB_SLAB_ROWS;
            const int lo = __reduce_min_sync(VB_FULL, rem ? g + (int)((cur.x & 0xffffu) >> VB_REC_COUNT_BITS) : 0x7fffffff);
            if (lo == 0x7fffffff) break;
            const int rel = lo / VB_SLAB_ROWS;
            if (rel > released) {
                released = rel;
                if (lane == 0) asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(prog_addr), "r"(rel) : "memory");
            }
            if (lo >= hi) {                        // every lane waits for a slab that has not landed
                if (ld == landed_seen) __nanosleep(32);
                landed_seen = ld;
                continue;
            }
            landed_seen = ld;
#pragma unroll
            for (int e = 0; e < VB_G_EPOCH; ++e) {
                const uint32_t r = cur.x & 0xffffu;
                const int gn = g + (int)(r >> VB_REC_COUNT_BITS);
                if (rem != 0 && gn < hi) {
                    g = gn;
                    const uint32_t c = r & VB_REC_MAX_COUNT;
                    cur.x = __funnelshift_r(cur.x, cur.y, 16);
                    cur.y >>= 16;
                    --rem;
                    if (--phase == 0) {
                        cur = nxt;
                        phase = 4;
                        if (blk_left) { nxt = __ldg(sp); sp += 32; --blk_left; }
                    }
                    if (c) {
                        const double dm = (double)c;
                        // ring position of row gn: gn mod VB_RING_ROWS by multiply-shift (exact for gn < 2^23)
                        const uint32_t q1536 = __umulhi((uint32_t)gn, VB_G_RING_MAGIC);
                        const uint32_t row = lane_base + (((uint32_t)gn - q1536 * VB_RING_ROWS) << 7);
#pragma unroll
                        for (int t = 0; t < NL; ++t) {
                            double vx, vy;
                            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(vx), "=d"(vy) : "r"(row ^ (uint32_t)(t << 4)));
                            acc[2 * t] = fma(dm, vx, acc[2 * t]);
                            acc[2 * t + 1] = fma(dm, vy, acc[2 * t + 1]);
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(prog_addr), "r"(0x7fffffff) : "memory");
    }

    // ---------------- epilogue: lane-local, K values per owner
    const int K = p.K;
    const int q = lane & 7;
    double r0 = 0.0, r1 = 0.0;                   // GM_CELL*: LB_p, KL_ID partial sums
    double t1[VB_MAX_GT], t2[VB_MAX_GT];         // GM_SNP: theta partial sums
#pragma unroll
    for (int gq = 0; gq < VB_MAX_GT; ++gq) t1[gq] = t2[gq] = 0.0;

    if (owner >= 0) {
        if (ga.has_heavy) {
            const double* __restrict__ H = p.H + ((size_t)b * gv.n_owner + owner) * VB_ROW_DOUBLES;
#pragma unroll
            for (int t = 0; t < NL; ++t) {
                const int k = 2 * ((q ^ t) & (NL - 1));
                acc[2 * t] += H[k];
                acc[2 * t + 1] += H[k + 1];
            }
        }
        if (ga.mode == GM_PLAIN) {
            double* __restrict__ out = ga.plain_out + ((size_t)b * gv.n_owner + owner) * VB_ROW_DOUBLES;
#pragma unroll
            for (int t = 0; t < NL; ++t) {
                const int k = 2 * ((q ^ t) & (NL - 1));
                *reinterpret_cast<double2*>(out + k) = make_double2(acc[2 * t], acc[2 * t + 1]);
            }
        } else if (ga.mode == GM_CELL || ga.mode == GM_CELL_LL) {
            const int64_t j = owner;
            const size_t prow = (size_t)(p.id_rows == 1 ? 0 : j) * K;
            double* __restrict__ R = p.R + ((size_t)b * p.C + j) * K;
            double* __restrict__ LL = p.ll + ((size_t)b * p.C + j) * K;
            double pr[NA];
            if (ga.mode == GM_CELL) {
                double mx = -INFINITY;
#pragma unroll
                for (int t = 0; t < NL; ++t) {
                    const int k = 2 * ((q ^ t) & (NL - 1));
                    pr[2 * t] = k < K ? acc[2 * t] + p.lidp[prow + k] : -INFINITY;
                    pr[2 * t + 1] = k + 1 < K ? acc[2 * t + 1] + p.lidp[prow + k + 1] : -INFINITY;
                    mx = fmax(mx, fmax(pr[2 * t], pr[2 * t + 1]));
                }
                double z = 0.0;
#pragma unroll
                for (int i = 0; i < NA; ++i) {
                    pr[i] = pr[i] == -INFINITY ? 0.0 : exp(pr[i] - mx);
                    z += pr[i];
                }
#pragma unroll
                for (int i = 0; i < NA; ++i) pr[i] = pr[i] / z;
                // padded, replicated copy for the SNP pass
                double* __restrict__ RP = p.RP + ((size_t)b * p.C + j) * VB_ROW_DOUBLES;
#pragma unroll
                for (int t = 0; t < NL; ++t) {
                    const int c2 = 2 * ((q ^ t) & (NL - 1));
#pragma unroll
                    for (int rep = 0; rep < VB_ROW_DOUBLES / NA; ++rep)
                        *reinterpret_cast<double2*>(RP + rep * NA + c2) = make_double2(pr[2 * t], pr[2 * t + 1]);
                }
            } else {
#pragma unroll
                for (int t = 0; t < NL; ++t) {
                    const int k = 2 * ((q ^ t) & (NL - 1));
                    pr[2 * t] = k < K ? R[k] : 0.0;
                    pr[2 * t + 1] = k + 1 < K ? R[k + 1] : 0.0;
                }
            }
#pragma unroll
            for (int t = 0; t < NL; ++t) {
                const int k0 = 2 * ((q ^ t) & (NL - 1));
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int k = k0 + h;
                    if (k < K) {
                        const double a = acc[2 * t + h], pv = pr[2 * t + h];
                        LL[k] = a;
                        if (ga.mode == GM_CELL) R[k] = pv;
                        r0 += a * pv;
                        if (pv > 0.0) r1 += pv * (log(pv) - p.lidp_kl[prow + k]);
                    }
                }
            }
        } else {   // GM_SNP
            const int64_t i = owner >> 1;
            const int al = owner & 1;
            const int G = p.G;
            double* __restrict__ S = (al ? p.S1 : p.S2) + ((size_t)b * p.V + i) * K;
            const double* __restrict__ GT = (do_theta && p.GT) ? p.GT + ((size_t)b * p.V + i) * K * G : nullptr;
#pragma unroll
            for (int t = 0; t < NL; ++t) {
                const int k0 = 2 * ((q ^ t) & (NL - 1));
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int k = k0 + h;
                    if (k < K) {
                        const double s = acc[2 * t + h];
                        S[k] = s;
                        if (GT) {
#pragma unroll
                            for (int gq = 0; gq < VB_MAX_GT; ++gq)
                                if (gq < G) t1[gq] += s * GT[(size_t)k * G + gq];
                        }
                    }
                }
            }
            if (do_theta && p.ase) {
                // theta per SNP (vireo_model.py:177 `axis=1`): the raw sums are parked in the ab rows
                double* row = p.ab + ((size_t)b * p.T + i) * 2 * G + (al ? 0 : G);
#pragma unroll
                for (int gq = 0; gq < VB_MAX_GT; ++gq)
                    if (gq < G) row[gq] = t1[gq];
            } else if (do_theta && !al) {
#pragma unroll
                for (int gq = 0; gq < VB_MAX_GT; ++gq) { t2[gq] = t1[gq]; t1[gq] = 0.0; }
            }
        }
    }

    // ---------------- block-level partial sums (fixed geometry -> run-to-run identical)
    if (ga.mode == GM_CELL || ga.mode == GM_CELL_LL) {
        r0 = warp_sum(r0);
        r1 = warp_sum(r1);
        if (lane == 0) { red[w * VB_G_RED_DOUBLES] = r0; red[w * VB_G_RED_DOUBLES + 1] = r1; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double a = 0.0, c = 0.0;
            for (int i = 0; i < nwarps; ++i) { a += red[i * VB_G_RED_DOUBLES]; c += red[i * VB_G_RED_DOUBLES + 1]; }
            double* out = p.part + (size_t)b * p.part_stride + p.off_cell + 2 * blockIdx.x;
            out[0] = a;
            out[1] = c;
        }
    } else if (ga.mode == GM_SNP && do_theta && !p.ase) {
#pragma unroll
        for (int gq = 0; gq < VB_MAX_GT; ++gq) {
            if (gq < p.G) {
                const double u1 = warp_sum(t1[gq]), u2 = warp_sum(t2[gq]);
                if (lane == 0) { red[w * VB_G_RED_DOUBLES + gq] = u1; red[w * VB_G_RED_DOUBLES + VB_MAX_GT + gq] = u2; }
            }
        }
        __syncthreads();
        if (threadIdx.x < 2 * VB_MAX_GT) {
            const int gq = threadIdx.x % VB_MAX_GT;
            double t = 0.0;
            if (gq < p.G)
                for (int i = 0; i < nwarps; ++i) t += red[i * VB_G_RED_DOUBLES + threadIdx.x];
            p.part[(size_t)b * p.part_stride + p.off_theta + (size_t)blockIdx.x * 2 * VB_MAX_GT + threadIdx.x] = t;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host: launch geometry and dispatch
// ---------------------------------------------------------------------------------------------
void vb_gather_geometry(const vb_counts* m, const GatherSet& g, int* grid, int* nwarps) {
    (void)m;
    *grid = g.grid > 0 ? g.grid : 1;
    *nwarps = g.max_warps > 0 ? g.max_warps : 1;
}

static bool g_attr_set[64] = {false};         // per device: function attributes belong to the context

template <int NL>
static int launch_nl(const GatherView& gv, const EmP& p, const GatherArgs& ga, dim3 grid, cudaStream_t st) {
    k_gather<NL><<<grid, (ga.nwarps + 1) * 32, VB_G_SMEM, st>>>(gv, p, ga);
    return VB_OK;
}

static GatherView view_of_set(const GatherSet& g) {
    GatherView v;
    v.n_owner = g.n_owner; v.n_gather = g.n_gather; v.n_slot = g.n_slot;
    v.perm = g.perm; v.len = g.len; v.slot_blk = g.slot_blk; v.rec = g.rec;
    v.hptr = g.hptr; v.hrow = g.hrow; v.hcnt = g.hcnt; v.cta_start = g.cta_start;
    return v;
}

// ori 0: cell pass (table = p.Wt padded), ori 1: SNP pass (table = p.RP)
int vb_gather_launch(const vb_counts* m, const EmP& p, int ori, int mode, int theta_mode, double* plain_out,
                     cudaStream_t st) {
    const GatherSet& g = ori ? m->gB : m->gA;
    if (!g.built) { vb_set_error("gather format was not built"); return VB_E_ARG; }
    const int dev_slot = m->device >= 0 && m->device < 64 ? m->device : 0;
    if (!g_attr_set[dev_slot]) {
        VB_CUDA(cudaFuncSetAttribute(k_gather<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, VB_G_SMEM));
        VB_CUDA(cudaFuncSetAttribute(k_gather<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, VB_G_SMEM));
        VB_CUDA(cudaFuncSetAttribute(k_gather<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, VB_G_SMEM));
        g_attr_set[dev_slot] = true;
    }
    const GatherView gv = view_of_set(g);
    GatherArgs ga;
    memset(&ga, 0, sizeof(ga));
    ga.mode = mode; ga.theta_mode = theta_mode;
    ga.KT = p.KT;
    ga.table = ori ? p.RP : p.Wt;
    ga.table_stride = g.n_gather * VB_ROW_DOUBLES;
    ga.plain_out = plain_out;
    ga.has_heavy = g.n_heavy > 0;
    int grid_x;
    vb_gather_geometry(m, g, &grid_x, &ga.nwarps);
    const int cls = ori ? 0 : 3;
    if (ga.has_heavy) {
        int64_t hb = (g.n_owner + VB_WARPS - 1) / VB_WARPS;
        if (hb > (int64_t)m->sm_count * 8) hb = (int64_t)m->sm_count * 8;
        if (hb < 1) hb = 1;
        VB_LAUNCH(cls, st, k_heavy<<<dim3((unsigned)hb, p.B), VB_THREADS, 0, st>>>(gv, ga.table, ga.table_stride, p.H,
                                                                                 g.n_owner * VB_ROW_DOUBLES, p.ctrl));
        VB_CUDA(cudaGetLastError());
    }
    const dim3 grid(grid_x, p.B);
    VB_LAUNCH(cls, st, {
        if (p.KT == 4) launch_nl<2>(gv, p, ga, grid, st);
        else if (p.KT == 8) launch_nl<4>(gv, p, ga, grid, st);
        else launch_nl<8>(gv, p, ga, grid, st);
    });
    VB_CUDA(cudaGetLastError());
    return VB_OK;
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
