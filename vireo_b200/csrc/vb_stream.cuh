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

// ORI 0: owner = cell j, gather row = 2*snp + allele;  ORI 1: owner = 2*snp + allele, gather row = cell.
// Calls f(gather_row - m.g_lo, count) for every pair with count > 0 and a gather row in [m.g_lo, m.g_hi), in ascending
// gather-row order; returns false if a count pair with ad > dp was met (the formats do not represent it).
template <int ORI, bool WIDE, typename F>
__device__ __forceinline__ bool for_records(const CountsView& m, int64_t o, F&& f) {
    bool ok = true;
    const int64_t lo = m.g_lo, hi = m.g_hi;
    if (ORI == 0) {
        const int64_t p0 = m.cell_ptr[o], p1 = m.cell_ptr[o + 1];
        for (int64_t q = p0; q < p1; ++q) {
            uint32_t a, d;
            counts_at<WIDE>(m.cell_cnt, m.cell_dp, q, a, d);
            if (a > d) { ok = false; continue; }
            const int64_t g = 2 * (int64_t)m.cell_idx[q];
            if (d - a && g >= lo && g < hi) f((int)(g - lo), d - a);
            if (a && g + 1 >= lo && g + 1 < hi) f((int)(g + 1 - lo), a);
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
            const int64_t g = m.snp_idx[q];
            if (c && g >= lo && g < hi) f((int)(g - lo), c);
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

