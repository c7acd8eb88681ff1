// vb_stream.cuh -- pieces shared by the two shared-memory gather families (vb_gather.cu, vb_seg.cu):
// enumeration of the (owner, gather row, count) pairs of a pass, scratch bookkeeping, and the
// mbarrier / bulk-copy PTX wrappers (SASS: SYNCS / UBLKCP).
#pragma once
#include "vb_common.cuh"

// ---------------------------------------------------------------------------------------------
// record enumeration in stream order
// ---------------------------------------------------------------------------------------------
template <bool WIDE>
__device__ __forceinline__ void counts_at(const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ dp, int64_t q,
                                          uint32_t& a, uint32_t& d) {
    if (WIDE) { a = cnt[q]; d = dp[q]; }
    else { const uint32_t c = cnt[q]; a = c & 0xffffu; d = c >> 16; }
}

// FP64 tables: a count rides in a record when 2 * count is exact in a double cut to 4 mantissa bits, i.e. when it has
// at most 5 significant bits (1..31, 32, 34, ..., 62, 64, 68, ...); fixed-point tables: counts up to 31.
__host__ __device__ inline bool seg_count_ok(uint32_t c, int fixed) {
    if (fixed) return c <= 31u;
    if (c >= (1u << 24)) return false;
    uint32_t t = c;
    while (t >= 32u) { if (t & 1u) return false; t >>= 1; }
    return true;
}

// ORI 0: owner = cell j, gather row = 2*snp + allele;  ORI 1: owner = 2*snp + allele, gather row = cell.
// Calls f(gather_row - m.g_lo, count) for every pair with count > 0 and a gather row in [m.g_lo, m.g_hi), in ascending
// gather-row order; returns false if a count pair with ad > dp was met (the formats do not represent it).
// Split owners (m.o_split): of the pairs whose count rides in the stream, the k-th goes to part k % parts; `o` names
// part 0 of a real owner, or -- with m.v_owner -- a virtual owner (real owner m.v_owner[o], part m.v_part[o]).  Pairs
// for the residual list stay with part 0.
template <int ORI, bool WIDE, typename F>
__device__ __forceinline__ bool for_records(const CountsView& m, int64_t o, F&& f) {
    bool ok = true;
    const int64_t lo = m.g_lo, hi = m.g_hi;
    uint32_t parts = 1, part = 0, k = 0;
    if (m.v_owner) { part = m.v_part[o]; o = m.v_owner[o]; parts = m.o_split[o]; }
    else if (m.o_split) parts = m.o_split[o];
    auto emit = [&](int64_t g, uint32_t c) {
        if (g < lo || g >= hi) return;
        if (parts > 1) {
            if (seg_count_ok(c, m.fixed)) { const bool mine = (k % parts) == part; ++k; if (!mine) return; }
            else if (part != 0) return;
        }
        f((int)(g - lo), c);
    };
    if (ORI == 0) {
        const int64_t p0 = m.cell_ptr[o], p1 = m.cell_ptr[o + 1];
        for (int64_t q = p0; q < p1; ++q) {
            uint32_t a, d;
            counts_at<WIDE>(m.cell_cnt, m.cell_dp, q, a, d);
            if (a > d) { ok = false; continue; }
            const int64_t g = 2 * (int64_t)m.cell_idx[q];
            if (d - a) emit(g, d - a);
            if (a) emit(g + 1, a);
        }
    } else {
        const int64_t i = o >> 1;
        const int al = (int)(o & 1);
        const int64_t p0 = m.snp_ptr[i], p1 = m.snp_ptr[i + 1];
        for (int64_t q = p0; q < p1; ++q) {
            uint32_t a, d;
            counts_at<WIDE>(m.snp_cnt, m.snp_dp, q, a, d);
            if (a > d) { ok = false; continue; }
            const uint32_t c = al ? a : d - a;
            if (c) emit(m.snp_idx[q], c);
        }
    }
    return ok;
}


// ---------------------------------------------------------------------------------------------
// host: build helpers
// ---------------------------------------------------------------------------------------------
static int grid1d(int64_t n, int sm) {
    int64_t b = (n + 255) / 256;
    if (b < 1) b = 1;
    const int64_t cap = (int64_t)sm * 16;
    return (int)(b > cap ? cap : b);
}

struct GsScratch {
    void* p[24];
    int n = 0;
    template <typename T> int alloc(T** out, size_t count) {
        cudaError_t e = cudaMalloc((void**)out, (count ? count : 1) * sizeof(T));
        if (e != cudaSuccess) { vb_set_error("cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e)); return VB_E_CUDA; }
        p[n++] = *out;
        return VB_OK;
    }
    ~GsScratch() { for (int i = 0; i < n; ++i) cudaFree(p[i]); }
};


// ---------------------------------------------------------------------------------------------
// device helpers: mbarrier + bulk copy (PTX; SASS shows UBLKCP / SYNCS)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
// the same wait with the warp off the issue slots between polls (a spinning warp competes with the working warps of
// its scheduler for every issue cycle)
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity, uint32_t ns) {
    while (!mbar_test(bar, parity)) __nanosleep(ns);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

