// vb_stage.cu -- one-off staging of the AD/DP count matrices into HBM, both orientations.
//
// Replaces, for the hot path, what the reference re-derives on every call from two scipy CSC
// matrices: `BD = DP - AD` (vireoSNP/utils/vireo_model.py:168,190,228; bmm_model.py:122,136) and the
// implicit CSC->CSR walk inside `AD @ ID_prob` (vireo_model.py:169-170).  Also hosts the binomial
// constant (vireoSNP/utils/vireo_base.py:7-22).
#include <cub/device/device_radix_sort.cuh>
#include <stdarg.h>
#include <string.h>

#include "vb_common.cuh"

static thread_local char g_err[512] = "";

void vb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* vb_last_error(void) { return g_err; }
extern "C" const char* vb_version(void) { return "vireo_b200 0.1.0 sm_100a"; }

// ---------------------------------------------------------------------------------------------
// conversion kernels
// ---------------------------------------------------------------------------------------------

// flags[0] = bad value seen, flags[1] = max value, flags[2] = AD entry outside DP's pattern
template <typename T>
__global__ void k_to_count(const T* __restrict__ in, uint32_t* __restrict__ out, int64_t n, unsigned int* flags) {
    unsigned int mx = 0, bad = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const T v = in[i];
        const double dv = (double)v;
        uint32_t u = 0;
        if (!(dv >= 0.0) || dv >= 2147483648.0 || dv != floor(dv)) bad = 1; else u = (uint32_t)dv;
        out[i] = u;
        mx = max(mx, u);
    }
    if (bad) atomicOr(&flags[0], 1u);
    atomicMax(&flags[1], mx);
}

template <typename TI, typename TO>
__global__ void k_cast(const TI* __restrict__ in, TO* __restrict__ out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (TO)in[i];
}

// index of the column (cell) holding nnz position q: largest j with ptr[j] <= q
__device__ __forceinline__ int64_t owner_of(const int64_t* __restrict__ ptr, int64_t n_col, int64_t q) {
    int64_t lo = 0, hi = n_col;   // invariant: ptr[lo] <= q < ptr[hi]
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (ptr[mid] <= q) lo = mid; else hi = mid;
    }
    return lo;
}

// scatter AD's values into DP's pattern: one thread per AD nnz, binary search inside the cell's DP segment
__global__ void k_merge_ad(const int64_t* __restrict__ ad_ptr, const int32_t* __restrict__ ad_idx,
                           const uint32_t* __restrict__ ad_val, int64_t ad_nnz, int64_t n_cell,
                           const int64_t* __restrict__ dp_ptr, const int32_t* __restrict__ dp_idx,
                           uint32_t* __restrict__ ad_at_dp, unsigned int* flags) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < ad_nnz; t += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t v = ad_val[t];
        if (v == 0) continue;                       // explicit zero: nothing to place
        const int64_t j = owner_of(ad_ptr, n_cell, t);
        const int32_t snp = ad_idx[t];
        int64_t lo = dp_ptr[j], hi = dp_ptr[j + 1];
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (dp_idx[mid] < snp) lo = mid + 1; else hi = mid;
        }
        if (lo < dp_ptr[j + 1] && dp_idx[lo] == snp) ad_at_dp[lo] = v;
        else atomicOr(&flags[2], 1u);
    }
}

__global__ void k_pack(const uint32_t* __restrict__ ad, const uint32_t* __restrict__ dp, uint32_t* __restrict__ out,
                       int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = ad[i] | (dp[i] << 16);
}

__global__ void k_iota(uint32_t* out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (uint32_t)i;
}

// after the stable sort by SNP id: build the SNP-major arrays
__global__ void k_gather_snp_major(const uint32_t* __restrict__ perm, int64_t n, const int64_t* __restrict__ cell_ptr,
                                   int64_t n_cell, const uint32_t* __restrict__ cell_cnt,
                                   const uint32_t* __restrict__ cell_dp, int32_t* __restrict__ snp_idx,
                                   uint32_t* __restrict__ snp_cnt, uint32_t* __restrict__ snp_dp) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = perm[t];
        snp_idx[t] = (int32_t)owner_of(cell_ptr, n_cell, q);
        snp_cnt[t] = cell_cnt[q];
        if (cell_dp) snp_dp[t] = cell_dp[q];
    }
}

// snp_ptr[i] = first position in the sorted key array with key >= i
__global__ void k_row_starts(const uint32_t* __restrict__ keys, int64_t n, int64_t n_row, int64_t* __restrict__ ptr) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i <= n_row; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = n;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if ((int64_t)keys[mid] < i) lo = mid + 1; else hi = mid;
        }
        ptr[i] = lo;
    }
}

// ---------------------------------------------------------------------------------------------
// host helpers
// ---------------------------------------------------------------------------------------------

static size_t dtype_size(int code) { return (code == VB_I32 || code == VB_F32) ? 4 : 8; }

static int grid_for(int64_t n, int sm) {
    int64_t b = (n + 255) / 256;
    if (b < 1) b = 1;
    const int64_t cap = (int64_t)sm * 16;
    return (int)(b > cap ? cap : b);
}

struct Scratch {   // frees whatever is still registered when it goes out of scope
    void* p[32];
    int n = 0;
    void* keep(void* q) { p[n++] = q; return q; }
    ~Scratch() { for (int i = 0; i < n; ++i) if (p[i]) cudaFree(p[i]); }
};

static int upload(const void* host, size_t bytes, cudaStream_t st, void** dev) {
    VB_CUDA(cudaMalloc(dev, bytes ? bytes : 8));
    if (bytes) VB_CUDA(cudaMemcpyAsync(*dev, host, bytes, cudaMemcpyHostToDevice, st));
    return VB_OK;
}

static int to_counts(const void* raw, int dtype, int64_t n, uint32_t* out, unsigned int* flags, int sm, cudaStream_t st) {
    const int g = grid_for(n, sm);
    switch (dtype) {
        case VB_I32: k_to_count<int32_t><<<g, 256, 0, st>>>((const int32_t*)raw, out, n, flags); break;
        case VB_I64: k_to_count<int64_t><<<g, 256, 0, st>>>((const int64_t*)raw, out, n, flags); break;
        case VB_F32: k_to_count<float><<<g, 256, 0, st>>>((const float*)raw, out, n, flags); break;
        case VB_F64: k_to_count<double><<<g, 256, 0, st>>>((const double*)raw, out, n, flags); break;
        default: vb_set_error("bad data dtype code %d", dtype); return VB_E_ARG;
    }
    VB_CUDA(cudaGetLastError());
    return VB_OK;
}

template <typename TO>
static int to_index(const void* raw, int dtype, int64_t n, TO* out, int sm, cudaStream_t st) {
    const int g = grid_for(n, sm);
    if (dtype == VB_I32) k_cast<int32_t, TO><<<g, 256, 0, st>>>((const int32_t*)raw, out, n);
    else if (dtype == VB_I64) k_cast<int64_t, TO><<<g, 256, 0, st>>>((const int64_t*)raw, out, n);
    else { vb_set_error("bad index dtype code %d", dtype); return VB_E_ARG; }
    VB_CUDA(cudaGetLastError());
    return VB_OK;
}

extern "C" void vb_counts_destroy(vb_counts* m) {
    if (!m) return;
    DeviceGuard dg(m->device);
    cudaFree(m->cell_ptr); cudaFree(m->cell_idx); cudaFree(m->cell_cnt); cudaFree(m->cell_dp);
    cudaFree(m->snp_ptr); cudaFree(m->snp_idx); cudaFree(m->snp_cnt); cudaFree(m->snp_dp);
    vb_seg_free(m);
    delete m;
}

extern "C" int64_t vb_counts_info(const vb_counts* m, int what) {
    if (!m) return -1;
    switch (what) {
        case 0: return m->C;
        case 1: return m->V;
        case 2: return m->N;
        case 3: return m->wide;
        case 4: return m->device;
        case 5: return m->bytes;
        case 6: return m->grid_cell;
        case 7: return m->grid_snp;
        case 8: return m->grid_elem;
        case 60: return m->auto_fallback;
        case 61: return m->rA[0].R > m->rA[2].R ? m->rA[0].R : m->rA[2].R;      // launches of the row-split cell pass (0: not in use)
        case 63: {   // virtual owners of the built segment formats (rows cut into parts because they were far heavier than the rest)
            int64_t n = 0;
            for (int q = 0; q < 3; ++q) n += m->sA[q].n_virtual + m->sB[q].n_virtual;
            return n;
        }
        case 62: {   // worst row imbalance of the built segment formats, per mille: 1000 * longest owner / mean owner
            int64_t worst = 0;
            for (int q = 0; q < 3; ++q)
                for (const SegSet* g : {&m->sA[q], &m->sB[q]})
                    if (g->built && g->n_light > 0 && g->n_owner > 0) {
                        const int64_t r = (int64_t)(1000.0 * (double)g->max_len * (double)g->n_owner / (double)g->n_light);
                        if (r > worst) worst = r;
                    }
            return worst;
        }
        // window-segment formats: 20 + 10 * precision + {0 built, 1 / 2 super-steps of the cell / SNP pass,
        // 3 / 4 largest reads of one owner's stream (cell / SNP pass), 5 / 6 grid.x, 7 bytes, 8 residual pairs, 9 stream pairs}
        default: break;
    }
    if (what >= 20 && what < 50) {
        const int q = (what - 20) / 10;          // 0 FP64 16 columns, 1 fixed point, 2 FP64 8 columns
        switch ((what - 20) % 10) {
            case 0: return m->sA[q].built && m->sB[q].built;
            case 1: return m->sA[q].n_step;
            case 2: return m->sB[q].n_step;
            case 3: return m->sA[q].max_reads;
            case 4: return m->sB[q].max_reads;
            case 5: return m->sA[q].grid;
            case 6: return m->sB[q].grid;
            case 7: return m->sA[q].bytes + m->sB[q].bytes;
            case 8: return m->sA[q].n_heavy;
            case 9: return m->sA[q].n_light;
        }
    }
    return -1;
}

extern "C" const char* vb_counts_note(const vb_counts* m) { return m ? m->seg_error : ""; }

// SNP-major orientation from the finished cell-major arrays (stable radix sort of the nnz positions by SNP id) and the
// launch geometry of the row kernels
static int finish_counts(vb_counts* m, cudaStream_t st) {
    const int sm = m->sm_count;
    const int64_t C = m->C, V = m->V, N = m->N;
    const int64_t Nz = N ? N : 1;
    Scratch tmp;
    VB_CUDA(cudaMalloc(&m->snp_ptr, (V + 1) * sizeof(int64_t)));
    VB_CUDA(cudaMalloc(&m->snp_idx, Nz * sizeof(int32_t)));
    VB_CUDA(cudaMalloc(&m->snp_cnt, Nz * sizeof(uint32_t)));
    if (m->wide) VB_CUDA(cudaMalloc(&m->snp_dp, Nz * sizeof(uint32_t)));
    if (N) {
        uint32_t *keys_out, *vals_in, *vals_out;
        VB_CUDA(cudaMalloc(&keys_out, N * sizeof(uint32_t))); tmp.keep(keys_out);
        VB_CUDA(cudaMalloc(&vals_in, N * sizeof(uint32_t))); tmp.keep(vals_in);
        VB_CUDA(cudaMalloc(&vals_out, N * sizeof(uint32_t))); tmp.keep(vals_out);
        k_iota<<<grid_for(N, sm), 256, 0, st>>>(vals_in, N);
        VB_CUDA(cudaGetLastError());
        int bits = 1;
        while ((1ll << bits) < V) ++bits;
        size_t tb = 0;
        const uint32_t* keys_in = (const uint32_t*)m->cell_idx;
        VB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys_in, keys_out, vals_in, vals_out, (int64_t)N, 0, bits, st));
        void* cub_tmp;
        VB_CUDA(cudaMalloc(&cub_tmp, tb ? tb : 8)); tmp.keep(cub_tmp);
        VB_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, tb, keys_in, keys_out, vals_in, vals_out, (int64_t)N, 0, bits, st));
        k_gather_snp_major<<<grid_for(N, sm), 256, 0, st>>>(vals_out, N, m->cell_ptr, C, m->cell_cnt, m->cell_dp,
                                                          m->snp_idx, m->snp_cnt, m->snp_dp);
        VB_CUDA(cudaGetLastError());
        k_row_starts<<<grid_for(V + 1, sm), 256, 0, st>>>(keys_out, N, V, m->snp_ptr);
        VB_CUDA(cudaGetLastError());
    } else {
        VB_CUDA(cudaMemsetAsync(m->snp_ptr, 0, (V + 1) * sizeof(int64_t), st));
    }
    VB_CUDA(cudaStreamSynchronize(st));

    // launch geometry: one warp per row, persistent grid capped at 8 CTAs per SM
    auto rows_grid = [&](int64_t rows) {
        int64_t b = (rows + VB_WARPS - 1) / VB_WARPS;
        if (b < 1) b = 1;
        const int64_t cap = (int64_t)sm * 8;
        return (int)(b > cap ? cap : b);
    };
    m->grid_cell = rows_grid(C);
    m->grid_snp = rows_grid(V);
    m->grid_elem = sm * 8;
    m->bytes = (C + 1 + V + 1) * 8 + N * (m->wide ? 24 : 16);
    return VB_OK;
}

__global__ void k_rebase_ptr(const int64_t* __restrict__ in, int64_t n, int64_t base, int64_t* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = in[i] - base;
}

extern "C" int vb_counts_slice(const vb_counts* src, int64_t c0, int64_t c1, void* stream, vb_counts** out) {
    if (!out) { vb_set_error("out is NULL"); return VB_E_ARG; }
    *out = nullptr;
    if (!src || c0 < 0 || c1 < c0 || c1 > src->C) { vb_set_error("bad cell range [%lld, %lld)", (long long)c0, (long long)c1); return VB_E_ARG; }
    DeviceGuard dg(src->device);
    cudaStream_t st = (cudaStream_t)stream;
    int64_t ends[2] = {0, 0};
    VB_CUDA(cudaMemcpyAsync(&ends[0], src->cell_ptr + c0, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaMemcpyAsync(&ends[1], src->cell_ptr + c1, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaStreamSynchronize(st));
    vb_counts* m = new vb_counts();
    memset(m, 0, sizeof(*m));
    m->device = src->device; m->sm_count = src->sm_count; m->C = c1 - c0; m->V = src->V; m->N = ends[1] - ends[0];
    m->wide = src->wide;       // keep the record layout of the parent (a shard without wide counts still reads them correctly)
    struct Guard { vb_counts* m; ~Guard() { if (m) vb_counts_destroy(m); } } guard{m};
    const int64_t Nz = m->N ? m->N : 1;
    VB_CUDA(cudaMalloc(&m->cell_ptr, (m->C + 1) * sizeof(int64_t)));
    VB_CUDA(cudaMalloc(&m->cell_idx, Nz * sizeof(int32_t)));
    VB_CUDA(cudaMalloc(&m->cell_cnt, Nz * sizeof(uint32_t)));
    if (m->wide) VB_CUDA(cudaMalloc(&m->cell_dp, Nz * sizeof(uint32_t)));
    k_rebase_ptr<<<grid_for(m->C + 1, m->sm_count), 256, 0, st>>>(src->cell_ptr + c0, m->C + 1, ends[0], m->cell_ptr);
    VB_CUDA(cudaGetLastError());
    if (m->N) {
        VB_CUDA(cudaMemcpyAsync(m->cell_idx, src->cell_idx + ends[0], m->N * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
        VB_CUDA(cudaMemcpyAsync(m->cell_cnt, src->cell_cnt + ends[0], m->N * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
        if (m->wide) VB_CUDA(cudaMemcpyAsync(m->cell_dp, src->cell_dp + ends[0], m->N * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    }
    const int rc = finish_counts(m, st);
    if (rc) return rc;
    guard.m = nullptr;
    *out = m;
    return VB_OK;
}

extern "C" int vb_counts_create(int device, int64_t n_cell, int64_t n_var,
                                const void* dp_indptr, int indptr_dtype,
                                const void* dp_indices, int indices_dtype,
                                const void* dp_data, int data_dtype, int64_t dp_nnz,
                                const void* ad_indptr, const void* ad_indices, const void* ad_data, int64_t ad_nnz,
                                void* stream, vb_counts** out) {
    if (!out) { vb_set_error("out is NULL"); return VB_E_ARG; }
    *out = nullptr;
    if (n_cell < 0 || n_var < 0 || dp_nnz < 0 || ad_nnz < 0 || n_cell >= (1ll << 31) || n_var >= (1ll << 31) ||
        dp_nnz >= (1ll << 31) || ad_nnz > dp_nnz + (1ll << 31)) {
        vb_set_error("unsupported shape: n_cell=%lld n_var=%lld nnz(DP)=%lld nnz(AD)=%lld (each must be < 2^31)",
                     (long long)n_cell, (long long)n_var, (long long)dp_nnz, (long long)ad_nnz);
        return VB_E_ARG;
    }
    if (!dp_indptr || !ad_indptr || (dp_nnz && (!dp_indices || !dp_data)) || (ad_nnz && (!ad_indices || !ad_data))) {
        vb_set_error("NULL input array");
        return VB_E_ARG;
    }
    DeviceGuard dg(device);
    cudaStream_t st = (cudaStream_t)stream;
    cudaDeviceProp prop;
    VB_CUDA(cudaGetDeviceProperties(&prop, device));
    const int sm = prop.multiProcessorCount;
    const int64_t C = n_cell, V = n_var, N = dp_nnz;
    const int64_t Nz = N ? N : 1;

    Scratch tmp;
    vb_counts* m = new vb_counts();
    memset(m, 0, sizeof(*m));
    m->device = device; m->sm_count = sm; m->C = C; m->V = V; m->N = N;
    struct Guard { vb_counts* m; ~Guard() { if (m) vb_counts_destroy(m); } } guard{m};

    // ---- raw uploads
    void *r_dp_ptr, *r_dp_idx, *r_dp_val, *r_ad_ptr, *r_ad_idx, *r_ad_val;
    int rc;
    if ((rc = upload(dp_indptr, (C + 1) * dtype_size(indptr_dtype), st, &r_dp_ptr))) return rc; tmp.keep(r_dp_ptr);
    if ((rc = upload(dp_indices, N * dtype_size(indices_dtype), st, &r_dp_idx))) return rc; tmp.keep(r_dp_idx);
    if ((rc = upload(dp_data, N * dtype_size(data_dtype), st, &r_dp_val))) return rc; tmp.keep(r_dp_val);
    if ((rc = upload(ad_indptr, (C + 1) * dtype_size(indptr_dtype), st, &r_ad_ptr))) return rc; tmp.keep(r_ad_ptr);
    if ((rc = upload(ad_indices, ad_nnz * dtype_size(indices_dtype), st, &r_ad_idx))) return rc; tmp.keep(r_ad_idx);
    if ((rc = upload(ad_data, ad_nnz * dtype_size(data_dtype), st, &r_ad_val))) return rc; tmp.keep(r_ad_val);

    // ---- typed device copies
    unsigned int* flags;
    VB_CUDA(cudaMalloc(&flags, 8 * sizeof(unsigned int))); tmp.keep(flags);
    VB_CUDA(cudaMemsetAsync(flags, 0, 8 * sizeof(unsigned int), st));
    VB_CUDA(cudaMalloc(&m->cell_ptr, (C + 1) * sizeof(int64_t)));
    VB_CUDA(cudaMalloc(&m->cell_idx, Nz * sizeof(int32_t)));
    int64_t* ad_ptr; int32_t* ad_idx; uint32_t *ad_val, *dp_val, *ad_at_dp;
    VB_CUDA(cudaMalloc(&ad_ptr, (C + 1) * sizeof(int64_t))); tmp.keep(ad_ptr);
    VB_CUDA(cudaMalloc(&ad_idx, (ad_nnz ? ad_nnz : 1) * sizeof(int32_t))); tmp.keep(ad_idx);
    VB_CUDA(cudaMalloc(&ad_val, (ad_nnz ? ad_nnz : 1) * sizeof(uint32_t))); tmp.keep(ad_val);
    VB_CUDA(cudaMalloc(&dp_val, Nz * sizeof(uint32_t)));
    VB_CUDA(cudaMalloc(&ad_at_dp, Nz * sizeof(uint32_t)));
    // dp_val / ad_at_dp may become members (wide) -- track them manually
    struct Pair { uint32_t *a, *b; ~Pair() { cudaFree(a); cudaFree(b); } } pair{dp_val, ad_at_dp};

    if ((rc = to_index<int64_t>(r_dp_ptr, indptr_dtype, C + 1, m->cell_ptr, sm, st))) return rc;
    if ((rc = to_index<int64_t>(r_ad_ptr, indptr_dtype, C + 1, ad_ptr, sm, st))) return rc;
    if (N && (rc = to_index<int32_t>(r_dp_idx, indices_dtype, N, m->cell_idx, sm, st))) return rc;
    if (ad_nnz && (rc = to_index<int32_t>(r_ad_idx, indices_dtype, ad_nnz, ad_idx, sm, st))) return rc;
    if (N && (rc = to_counts(r_dp_val, data_dtype, N, dp_val, flags, sm, st))) return rc;
    if (ad_nnz && (rc = to_counts(r_ad_val, data_dtype, ad_nnz, ad_val, flags, sm, st))) return rc;

    // ---- AD into DP's pattern
    VB_CUDA(cudaMemsetAsync(ad_at_dp, 0, Nz * sizeof(uint32_t), st));
    if (ad_nnz) {
        k_merge_ad<<<grid_for(ad_nnz, sm), 256, 0, st>>>(ad_ptr, ad_idx, ad_val, ad_nnz, C, m->cell_ptr, m->cell_idx,
                                                       ad_at_dp, flags);
        VB_CUDA(cudaGetLastError());
    }
    unsigned int hflags[8];
    VB_CUDA(cudaMemcpyAsync(hflags, flags, sizeof(hflags), cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaStreamSynchronize(st));
    if (hflags[0]) { vb_set_error("AD/DP hold a negative, non-integer or >= 2^31 count"); return VB_E_VALUE; }
    if (hflags[2]) { vb_set_error("AD has a non-zero entry where DP stores nothing (pattern(AD) must be within pattern(DP))"); return VB_E_PATTERN; }
    m->wide = hflags[1] > 65535u ? 1 : 0;

    if (m->wide) {
        m->cell_cnt = ad_at_dp; m->cell_dp = dp_val;
        pair.a = pair.b = nullptr;
    } else {
        VB_CUDA(cudaMalloc(&m->cell_cnt, Nz * sizeof(uint32_t)));
        if (N) { k_pack<<<grid_for(N, sm), 256, 0, st>>>(ad_at_dp, dp_val, m->cell_cnt, N); VB_CUDA(cudaGetLastError()); }
    }

    if ((rc = finish_counts(m, st))) return rc;
    guard.m = nullptr;
    *out = m;
    return VB_OK;
}

// ---------------------------------------------------------------------------------------------
// binomial constant
// ---------------------------------------------------------------------------------------------

template <bool WIDE>
__global__ void __launch_bounds__(VB_THREADS) k_binom(const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ dp,
                                                      int64_t n, double* __restrict__ part) {
    __shared__ double sh[VB_WARPS];
    double acc = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t a, d;
        if (WIDE) { a = cnt[i]; d = dp[i]; } else { const uint32_t c = cnt[i]; a = c & 0xffffu; d = c >> 16; }
        if (d > 0) acc += (double)vb_binom_term(a, d);
    }
    const double t = block_sum(acc, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = t;
}

__global__ void k_sum_partials(const double* __restrict__ part, int n, double* __restrict__ out) {
    __shared__ double sh[VB_WARPS];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += part[i];
    const double t = block_sum(acc, sh);
    if (threadIdx.x == 0) out[0] = t;
}

extern "C" int vb_binom_const(const vb_counts* m, double* scratch, double* out_host, void* stream) {
    if (!m || !scratch || !out_host) { vb_set_error("NULL argument"); return VB_E_ARG; }
    DeviceGuard dg(m->device);
    cudaStream_t st = (cudaStream_t)stream;
    int g = grid_for(m->N, m->sm_count);
    if (g > 1023) g = 1023;
    if (m->wide) k_binom<true><<<g, VB_THREADS, 0, st>>>(m->cell_cnt, m->cell_dp, m->N, scratch + 1);
    else k_binom<false><<<g, VB_THREADS, 0, st>>>(m->cell_cnt, nullptr, m->N, scratch + 1);
    VB_CUDA(cudaGetLastError());
    k_sum_partials<<<1, VB_THREADS, 0, st>>>(scratch + 1, g, scratch);
    VB_CUDA(cudaGetLastError());
    VB_CUDA(cudaMemcpyAsync(out_host, scratch, sizeof(double), cudaMemcpyDeviceToHost, st));
    VB_CUDA(cudaStreamSynchronize(st));
    return VB_OK;
}

// host-callable scalar math, so the CPU test-suite can check the device formulas without a GPU
extern "C" double vb_host_digamma(double x) { return vb_digamma(x); }
extern "C" double vb_host_beta_kl(double p1, double p2, double q1, double q2) {
    return vb_beta_kl(p1, p2, q1, q2, vb_digamma(p1), vb_digamma(p2), vb_digamma(p1 + p2));
}
extern "C" float vb_host_binom_term(uint32_t a, uint32_t d) { return vb_binom_term(a, d); }
