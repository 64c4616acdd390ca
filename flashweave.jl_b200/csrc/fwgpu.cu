// fwgpu.cu — libfwgpu.so: context, C ABI (include/fwgpu.h) and kernel launches.
// Built for sm_100a only (see build.py).  No torch types cross this boundary.
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/fwgpu.h"
#include "common.cuh"
#include "fz.cuh"
#include "subsets.cuh"
#include "hiton.cuh"
#include "pairwise.cuh"
#include "cor_tc.cuh"
#include "cor_tc3.cuh"
#include "comm.cuh"
#include "mi.cuh"
#include "hiton_mi.cuh"
#include "prep.cuh"

static_assert(sizeof(fw_test_result) == 32, "TestResult layout (src/types.jl:140-145)");
static_assert(sizeof(DevResult) == 32, "DevResult layout");

static thread_local std::string g_create_error;

template <class T>
struct DevBuf {
    T* ptr = nullptr;
    size_t cap = 0;     // elements
    bool owned = true;
    cudaError_t reserve(size_t n) {
        if (n <= cap && ptr && owned) return cudaSuccess;
        release();
        if (n == 0) n = 1;
        cudaError_t e = cudaMalloc((void**)&ptr, n * sizeof(T));
        if (e == cudaSuccess) { cap = n; owned = true; } else { ptr = nullptr; cap = 0; }
        return e;
    }
    void adopt(T* p, size_t n) { release(); ptr = p; cap = n; owned = false; }
    void release() { if (ptr && owned) cudaFree(ptr); ptr = nullptr; cap = 0; owned = true; }
    ~DevBuf() { release(); }
};

struct fw_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;      // H2D chunks of fw_upload_cor_f32
    cudaEvent_t chunk_ev[12] = {nullptr};
    std::string err;
    int index_base = 0;
    i64 launches = 0;

    // data
    i64 n = 0, p = 0, ld = 0;
    int data_kind = -1;                  // 0: f32 continuous, 1: i32 discrete, -1 none
    DevBuf<float> d_data_f32;
    DevBuf<int> d_data_i32;
    i64 n_obs = -1;                      // rows used by Fisher-z tests
    // discrete table: bit planes + per-variable level statistics (mi.cuh)
    DevBuf<unsigned int> d_planes; DevBuf<int> d_levels, d_maxvals, d_nnz, d_bad; DevBuf<double> d_logtab;
    std::vector<int> h_levels, h_maxvals;
    int disc_L = 0, disc_W = 0;
    int sparse_sem = 0;                  // fw_set_semantics
    // fz_nz: non-zero planes of the continuous table (fznz.cuh), built on first use
    DevBuf<unsigned int> d_nzmask; DevBuf<int> d_nnz_f; bool nz_ready = false;

    std::vector<uint8_t> meta_mask;     // meta_variable_mask of the resident table (carried to the host side's edgelist; no device use)
    // cor_mat: the full symmetric matrix (d_cor), or - in a multi-GPU group - this rank's row shard (grp, CorView)
    DevBuf<float> d_cor; i64 cor_p = 0;
    bool cor_sharded = false;
    // raw candidates of the univariate Fisher-z stage (PwRec, common.cuh): collected by the cor_mat GEMM epilogue when
    // fw_pairwise_prefetch announced the parameters, else by one pass over the resident matrix
    struct Collect { bool armed = false, valid = false; double alpha = 0.0; i64 n_obs_min = 0; float r_lo = 2.0f; i64 cap = 0; } col;
    DevBuf<PwRec> d_list; DevBuf<u64> d_listcnt; DevBuf<PwRec> d_gathered;
    // multi-GPU group (comm.cuh)
    fwcomm::Group grp;
    DevBuf<float> g_table; DevBuf<float> g_cor; DevBuf<PwRec> g_list; DevBuf<u64> g_flags; DevBuf<int> g_err;

    // univariate neighbour lists (device-resident CSR + host copy of the offsets)
    DevBuf<i64> d_uni_off, d_uni_nbr; DevBuf<double> d_uni_stat, d_uni_p;
    std::vector<i64> h_uni_off; i64 uni_entries = -1;
    i64 pw_tests = 0, pw_reliable = 0, pw_raw_sig = 0;
    PwCollected part; int part_kind = -1;   // fw_pairwise_partial: this rank's raw-significant records (device, scratch slots 7-10)
    i64 exec_by_k[3] = {0, 0, 0};     // tests executed by the last fw_hiton_pc with |Zs| = 1, 2, 3

    // per-phase device timing (CUDA events on `stream`), see fw_last_timing
    cudaEvent_t ev[8] = {nullptr};
    bool ev_valid[4] = {false, false, false, false};
    cudaEvent_t evx[3] = {nullptr, nullptr, nullptr};   // fw_multi_cor: after the standardising kernel, before / after the closing group barrier
    bool evx_valid = false;

    // scratch
    DevBuf<int> d_counter; DevBuf<u64> d_exec;
    PairwiseScratch pw;
    cortc::Scratch tc;
    fznztc::Planes nzplanes;             // bf16 operand planes of the fz_nz pairwise pre-filter
    cortc::Prepared tcp;                 // standardised table + TMA descriptors of the row-sharded cor_mat path
    // fw_hiton_pc work buffers (grow-only, reused across calls)
    struct HitonBufs {
        DevBuf<i64> dt, doff, dpcn, dtpcn, dpcc, dtpcc, dnt; DevBuf<double> dpcs, dpcp, dtpcs, dtpcp; DevBuf<int> dsel, dorder, dstatus; DevBuf<float> gs;
    } hb;
};

static int fail(fw_ctx* c, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ctx, FW_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define NEED(cond, code, ...) do { if (!(cond)) return fail(ctx, code, __VA_ARGS__); } while (0)

// a new resident table invalidates everything derived from the previous one: cor_mat, the standardised split table, the
// non-zero planes and the univariate neighbour lists (fw_hiton_pc / fw_pairwise_copy then fail with FW_ERR_STATE instead of
// silently running against the lists of another table)
static void table_changed(fw_ctx* c) {
    c->meta_mask.clear();
    c->nz_ready = false; c->tcp.valid = false; c->cor_p = 0; c->cor_sharded = false; c->col.valid = false;
    c->uni_entries = -1; c->h_uni_off.clear();
}
static bool has_cor(const fw_ctx* c) { return c->cor_p > 0 && (c->cor_sharded ? c->grp.attached : c->d_cor.ptr != nullptr); }
static CorView make_cor_view(const fw_ctx* c) {
    CorView v;
    for (int q = 0; q < FW_MAX_RANKS; ++q) v.shard[q] = nullptr;
    v.p = c->cor_p;
    if (c->cor_sharded) { v.world = c->grp.world; v.h = c->grp.h; for (int q = 0; q < c->grp.world; ++q) v.shard[q] = c->grp.cor_of(q); }
    else { v.world = 1; v.h = 1; v.shard[0] = c->d_cor.ptr; }
    return v;
}


// ---- multi-GPU group: device-side barrier (comm.cuh) -------------------------------------------------------------------------
static const long long kCommTimeoutClk = 20000000000LL;          // ~10 s of SM clocks: a dead peer becomes an error, never a hang
static int comm_barrier(fw_ctx* ctx) {
    fwcomm::Group& G = ctx->grp;
    G.seq++;
    fwcomm::signal_kernel<<<1, 1, 0, ctx->stream>>>(G.flags_of(G.rank) + fwcomm::F_SEQ, G.seq);
    fwcomm::PeerFlags pf;
    for (int q = 0; q < FW_MAX_RANKS; ++q) pf.f[q] = q < G.world ? G.flags_of(q) + fwcomm::F_SEQ : nullptr;
    fwcomm::wait_kernel<<<1, 32, 0, ctx->stream>>>(pf, G.world, G.seq, kCommTimeoutClk, G.d_err);
    ctx->launches += 2;
    cudaError_t e_ = cudaGetLastError();
    if (e_ != cudaSuccess) { ctx->err = std::string("group barrier launch failed: ") + cudaGetErrorString(e_); return FW_ERR_CUDA; }
    return FW_OK;
}
// after a stream synchronisation: did a barrier time out?
static int comm_check(fw_ctx* ctx) {
    if (!ctx->grp.attached) return FW_OK;
    int h = 0;
    if (cudaMemcpy(&h, ctx->grp.d_err, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) { ctx->err = "group error flag unreadable"; return FW_ERR_CUDA; }
    if (h) { ctx->err = "group barrier timed out waiting for rank " + std::to_string(h - 1) + " (every rank must issue the same sequence of collective calls)"; return FW_ERR_CUDA; }
    return FW_OK;
}

static i64 list_capacity(i64 n_pairs) { return std::max<i64>(1, std::min<i64>(n_pairs, std::max<i64>((i64)1 << 16, n_pairs / 16))); }

static FzConsts make_fz_consts(i64 n_rows, i64 n_obs_min) {
    FzConsts fc;
    i64 sf = n_rows - 0 - 3;                 // len_z is hard-coded 0 (src/tests.jl:156,256)
    fc.sf_pos = sf > 0 ? 1 : 0;
    fc.half_sqrt_sf = sf > 0 ? std::sqrt((double)sf) / 2.0 : 0.0;
    fc.rows_ok = n_rows >= n_obs_min ? 1 : 0;
    return fc;
}

// Decision band of the p-value-free scan (fz.cuh, FzConsts): the |stat| at which fz_pval crosses alpha (bisection on the same
// formula, statfuns.jl:3-17), widened by 1e-9 relative on both sides - far more than the ulp-level differences between this
// host evaluation and the device's log/erfc - and the |stat| beyond which the p-value drops below 1e-290.
static double host_fz_pval(double s, const FzConsts& fc) {
    double fz = fc.sf_pos ? fc.half_sqrt_sf * std::log((1.0 + s) / (1.0 - s)) : 0.0;
    return std::erfc(std::fabs(fz) * 0.70710678118654752440) / 2.0 * 2.0;
}
static double host_fz_cross(const FzConsts& fc, double level) {       // smallest s in [0, 1] with pval(s) < level (1.0 if none)
    if (!(host_fz_pval(1.0, fc) < level)) return 2.0;
    double lo = 0.0, hi = 1.0;                                         // pval(lo) >= level > pval(hi)
    for (int it = 0; it < 200; ++it) { double mid = 0.5 * (lo + hi); if (host_fz_pval(mid, fc) < level) hi = mid; else lo = mid; }
    return hi;
}
static void fill_fz_band(FzConsts& fc, double alpha) {
    fc.s_lo = -1.0; fc.s_hi = 1e300; fc.s_under = 0.0;                // exact p-value everywhere
    if (!(alpha > 1e-280 && alpha < 1.0)) return;
    if (!(host_fz_pval(0.0, fc) >= alpha)) return;
    const double sa = host_fz_cross(fc, alpha);                        // 2.0: never significant
    fc.s_lo = sa >= 2.0 ? 1e300 : sa * (1.0 - 1e-9);
    fc.s_hi = sa >= 2.0 ? 1e300 : sa * (1.0 + 1e-9);
    const double su = host_fz_cross(fc, 1e-290);
    fc.s_under = su >= 2.0 ? 1e300 : su * (1.0 - 1e-9);
    if (fc.s_under < fc.s_hi) fc.s_under = fc.s_hi;
}

static MiTable make_mi_table(const fw_ctx* c, int kind) {
    MiTable t;
    t.planes = c->d_planes.ptr; t.levels = c->d_levels.ptr; t.max_vals = c->d_maxvals.ptr; t.nnz = c->d_nnz.ptr;
    t.p = c->p; t.n = (int)c->n; t.W = c->disc_W; t.L = c->disc_L; t.nz = (kind == FW_MI_NZ) ? 1 : 0;
    const int rem = (int)(c->n & 31);
    t.tail_mask = rem ? ((1u << rem) - 1u) : 0xffffffffu;
    t.lgt = c->d_logtab.ptr;
    t.sparse_sem = (c->sparse_sem && kind == FW_MI_NZ) ? 1 : 0;
    return t;
}

static int ensure_nz_table(fw_ctx* ctx, NzTable* t) {
    if (ctx->data_kind != 0) return fail(ctx, FW_ERR_STATE, "fz_nz: no continuous table resident (call fw_set_data_f32 first)");
    const int W = (int)((ctx->n + 31) / 32);
    if (!ctx->nz_ready) {
        cudaError_t e = ctx->d_nzmask.reserve((size_t)ctx->p * W);
        if (e == cudaSuccess) e = ctx->d_nnz_f.reserve(ctx->p);
        if (e == cudaSuccess) e = cudaMemsetAsync(ctx->d_nnz_f.ptr, 0, sizeof(int) * ctx->p, ctx->stream);
        if (e != cudaSuccess) return fail(ctx, FW_ERR_CUDA, "fz_nz: mask allocation failed: %s", cudaGetErrorString(e));
        const int T = 256; const i64 warps = ctx->p * W; const i64 blocks = (warps * 32 + T - 1) / T;
        nz_mask_kernel<<<(unsigned)blocks, T, 0, ctx->stream>>>(ctx->d_data_f32.ptr, ctx->n, ctx->ld, ctx->p, W, ctx->d_nzmask.ptr, ctx->d_nnz_f.ptr);
        ctx->launches++;
        e = cudaGetLastError();
        if (e != cudaSuccess) return fail(ctx, FW_ERR_CUDA, "nz_mask_kernel: %s", cudaGetErrorString(e));
        ctx->nz_ready = true;
    }
    t->data = ctx->d_data_f32.ptr; t->nzmask = ctx->d_nzmask.ptr; t->nnz = ctx->d_nnz_f.ptr; t->p = ctx->p; t->ld = ctx->ld; t->n = (int)ctx->n; t->W = W;
    return FW_OK;
}

// longest jobs first: stable order of `sel` by descending candidate count (hoff[t + 1] - hoff[t]); a counting sort, because
// this runs on the host in front of every HITON launch (50 000 targets at C4) while the GPU waits
static void sort_by_candidates_desc(std::vector<int>& sel, const std::vector<i64>& hoff) {
    i64 mx = 0;
    for (int t : sel) mx = std::max<i64>(mx, hoff[t + 1] - hoff[t]);
    if (mx > (i64)4 * (i64)sel.size() + 1024) {                                  // sparse key range: comparison sort
        std::stable_sort(sel.begin(), sel.end(), [&](int x, int y) { return (hoff[x + 1] - hoff[x]) > (hoff[y + 1] - hoff[y]); });
        return;
    }
    std::vector<i64> start((size_t)mx + 2, 0);
    for (int t : sel) start[(size_t)(mx - (hoff[t + 1] - hoff[t])) + 1]++;       // bucket b = mx - count: descending count
    for (size_t b = 1; b < start.size(); ++b) start[b] += start[b - 1];
    std::vector<int> out(sel.size());
    for (int t : sel) out[(size_t)start[(size_t)(mx - (hoff[t + 1] - hoff[t]))]++] = t;
    sel.swap(out);
}

// ---- capacity classes shared by the subset-search and HITON launches -------------------------
static const int kCaps[4] = {32, 64, 128, 224};
static size_t hiton_smem_bytes(int cap, bool r_in_smem, int nz_words = -1, bool cache = false) {
    size_t o = r_in_smem ? sizeof(float) * (size_t)cap * cap : 0;
    o = (o + 15) & ~(size_t)15;
    o += sizeof(i64) * (cap + 1) + 4 * sizeof(double) * cap + sizeof(i64) * cap + 3 * sizeof(int) * cap + (size_t)cap;
    o = (o + 15) & ~(size_t)15;
    if (nz_words >= 0) o += sizeof(i64) * cap + 2 * sizeof(double) * cap + sizeof(unsigned int) * nz_words + 16 + FZNZ_GRAM_BYTES;
    o = (o + 15) & ~(size_t)15;
    if (cache) o = (size_t)fz_tab_layout((int)o).end;
    return o + 16;
}
static size_t subsets_smem_bytes(int cap, bool r_in_smem, int nz_words = -1) {
    size_t o = r_in_smem ? sizeof(float) * (size_t)cap * cap : 0;
    o = (o + 15) & ~(size_t)15;
    o += sizeof(i64) * (cap + 1) + sizeof(int) * cap;
    o = (o + 15) & ~(size_t)15;
    if (nz_words >= 0) o += sizeof(i64) * cap + 2 * sizeof(double) * cap + sizeof(unsigned int) * nz_words + 16 + FZNZ_GRAM_BYTES;
    return o + 16;
}

template <class K>
static cudaError_t grid_for(K kernel, int threads, size_t smem, int sm_count, i64 n_items, int* grid) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
    i64 g = (i64)sm_count * occ;      // persistent CTAs: a whole number of CTAs per SM
    if (g > n_items) g = n_items;
    if (g < 1) g = 1;
    *grid = (int)g;
    return cudaSuccess;
}


// SparseMatrixCSC{Float32,Int64} / SparseMatrixCSC{Int32,Int64} input (the reference's default make_sparse = true tables):
// uploaded as the CSC triple (nnz-proportional PCIe traffic) and densified on the device into the resident table
template <class T>
static int set_data_csc(fw_ctx* ctx, const int64_t* colptr, const int64_t* rowval, const T* nzval, int64_t n, int64_t p, DevBuf<T>& dst, const char* who) {
    NEED(colptr && n > 0 && p > 0, FW_ERR_INVALID, "%s: bad arguments", who);
    const i64 base = ctx->index_base;
    const i64 nnz = colptr[p] - colptr[0];
    NEED(colptr[0] == base && nnz >= 0 && (nnz == 0 || (rowval && nzval)), FW_ERR_INVALID, "%s: bad CSC structure (colptr[0] must equal the index base %lld)", who, (long long)base);
    for (i64 v = 0; v < p; ++v) NEED(colptr[v + 1] >= colptr[v], FW_ERR_INVALID, "%s: colptr not monotone", who);
    CK(cudaSetDevice(ctx->device));
    DevBuf<i64> dcp, drv; DevBuf<T> dnz;
    CK(dcp.reserve(p + 1)); CK(drv.reserve(nnz)); CK(dnz.reserve(nnz)); CK(dst.reserve((size_t)n * p)); CK(ctx->d_bad.reserve(4));
    CK(cudaMemcpyAsync(dcp.ptr, colptr, sizeof(i64) * (p + 1), cudaMemcpyHostToDevice, ctx->stream));
    if (nnz) {
        CK(cudaMemcpyAsync(drv.ptr, rowval, sizeof(i64) * nnz, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(dnz.ptr, nzval, sizeof(T) * nnz, cudaMemcpyHostToDevice, ctx->stream));
    }
    CK(cudaMemsetAsync(dst.ptr, 0, sizeof(T) * (size_t)n * p, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_bad.ptr, 0, sizeof(int), ctx->stream));
    csc_scatter_kernel<T><<<(unsigned)p, 256, 0, ctx->stream>>>(dcp.ptr, drv.ptr, dnz.ptr, n, p, base, dst.ptr, ctx->d_bad.ptr);
    ctx->launches++;
    CK(cudaGetLastError());
    int bad = 0;
    CK(cudaMemcpyAsync(&bad, ctx->d_bad.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    NEED(!bad, FW_ERR_INVALID, "%s: row index out of range", who);
    return FW_OK;
}

// ---- raw candidates of the univariate Fisher-z stage ---------------------------------------------------------------------------
// fw_pairwise_prefetch announced (alpha, n_obs_min): the cor_mat GEMM about to be launched appends every |r| >= r_lo of its
// tiles to the candidate list in its epilogue (the list of the group buffer when the matrix is row-sharded)
static int arm_collect(fw_ctx* ctx, bool sharded, PwEmit* em) {
    em->list = nullptr; em->counters = nullptr; em->cap = 0; em->r_lo = 2.0f; em->on = 0;
    ctx->col.valid = false;
    if (!ctx->col.armed) return FW_OK;
    const i64 p = ctx->p, n_pairs = p * (p - 1) / 2;
    const FzConsts fc = make_fz_consts(ctx->n_obs, ctx->col.n_obs_min);
    ctx->col.r_lo = pairwise_fz_r_lo(fc, ctx->n_obs, ctx->col.n_obs_min, ctx->col.alpha);
    CK(ctx->d_listcnt.reserve(4));
    if (sharded) { em->list = ctx->grp.list_of(ctx->grp.rank); em->cap = ctx->grp.list_cap; }
    else { const i64 cap = std::max<i64>(list_capacity(n_pairs), (i64)ctx->d_list.cap); CK(ctx->d_list.reserve((size_t)cap)); em->list = ctx->d_list.ptr; em->cap = (i64)ctx->d_list.cap; }
    CK(cudaMemsetAsync(ctx->d_listcnt.ptr, 0, 4 * sizeof(u64), ctx->stream));
    em->counters = ctx->d_listcnt.ptr; em->r_lo = ctx->col.r_lo; em->on = 1;
    ctx->col.cap = em->cap;
    return FW_OK;
}

// pw_univar_neighbors for test_name "fz" on the resident cor_mat (tests.jl:470-478 + :372-407 + statfuns.jl:326-350)
static int run_pairwise_fz(fw_ctx* ctx, double alpha, i64 n_obs_min, bool fdr, bool rel_only, PairwiseOut* po, int* nl) {
    const i64 p = ctx->cor_p, n_pairs = p * (p - 1) / 2;
    const FzConsts fc = make_fz_consts(ctx->n_obs, n_obs_min);
    const float r_lo = pairwise_fz_r_lo(fc, ctx->n_obs, n_obs_min, alpha);
    const bool multi = ctx->cor_sharded;
    fwcomm::Group& G = ctx->grp;
    std::string msg;
    cudaStream_t st = ctx->stream;
    const bool reuse = ctx->col.valid && ctx->col.alpha == alpha && ctx->col.n_obs_min == n_obs_min && ctx->col.r_lo == r_lo;
    const PwRec* recs = nullptr; i64 nrec = 0, n_nan = 0;
    u64 hc[2] = {0, 0};
    CK(ctx->d_listcnt.reserve(4));
    if (!multi) {
        bool have = reuse;
        i64 cap = (i64)ctx->d_list.cap;
        if (have) {
            CK(cudaMemcpyAsync(hc, ctx->d_listcnt.ptr, 2 * sizeof(u64), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if ((i64)hc[0] > ctx->col.cap) { have = false; cap = (i64)hc[0]; }          // the GEMM epilogue ran out of list space: re-scan with the exact size
        } else cap = std::max<i64>(cap, list_capacity(n_pairs));
        for (int attempt = 0; !have && attempt < 2; ++attempt) {
            CK(ctx->d_list.reserve((size_t)cap));
            CK(cudaMemsetAsync(ctx->d_listcnt.ptr, 0, 4 * sizeof(u64), st));
            PwEmit em; em.list = ctx->d_list.ptr; em.counters = ctx->d_listcnt.ptr; em.cap = (i64)ctx->d_list.cap; em.r_lo = r_lo; em.on = 1;
            cudaError_t e = pairwise_fz_emit(ctx->d_cor.ptr, p, 0, 0, p - 1, em, st, nl, &msg);
            if (e != cudaSuccess) return fail(ctx, FW_ERR_CUDA, "fw_pairwise: %s: %s", msg.c_str(), cudaGetErrorString(e));
            CK(cudaMemcpyAsync(hc, ctx->d_listcnt.ptr, 2 * sizeof(u64), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if ((i64)hc[0] <= em.cap) have = true; else cap = (i64)hc[0];
        }
        NEED(have, FW_ERR_CUDA, "fw_pairwise: candidate list overflow");
        ctx->col.valid = false;                                                         // (a re-scan replaced the collected list)
        if (reuse && (i64)hc[0] <= ctx->col.cap) ctx->col.valid = true;
        recs = ctx->d_list.ptr; nrec = (i64)hc[0]; n_nan = (i64)hc[1];
    } else {
        NEED(G.attached, FW_ERR_STATE, "fw_pairwise: the row-sharded cor_mat needs an attached group");
        if (!reuse) {
            // collective: every rank scans its own tile rows (peers may still be pulling the previous list: barrier first)
            int st_ = comm_barrier(ctx); if (st_ != FW_OK) return st_;
            CK(cudaMemsetAsync(ctx->d_listcnt.ptr, 0, 4 * sizeof(u64), st));
            PwEmit em; em.list = G.list_of(G.rank); em.counters = ctx->d_listcnt.ptr; em.cap = G.list_cap; em.r_lo = r_lo; em.on = 1;
            const int grp_id[2] = {G.rank, 2 * G.world - 1 - G.rank};
            for (int k = 0; k < 2; ++k) {
                const i64 g0 = (i64)grp_id[k] * G.h * 128, g1 = std::min<i64>(g0 + (i64)G.h * 128, p);
                cudaError_t e = pairwise_fz_emit(G.cor_of(G.rank), p, g0, (i64)k * G.h * 128, g1 - g0, em, st, nl, &msg);
                if (e != cudaSuccess) return fail(ctx, FW_ERR_CUDA, "fw_pairwise: %s: %s", msg.c_str(), cudaGetErrorString(e));
            }
            CK(cudaMemcpyAsync(G.flags_of(G.rank) + fwcomm::F_LIST_N, ctx->d_listcnt.ptr, 2 * sizeof(u64), cudaMemcpyDeviceToDevice, st));
            st_ = comm_barrier(ctx); if (st_ != FW_OK) return st_;
            ctx->col.valid = false;
        }
        // pull the peers' compact lists (global m and rank order of Benjamini-Hochberg need all of them, statfuns.jl:326-350)
        u64 cnt[FW_MAX_RANKS][2];
        for (int q = 0; q < G.world; ++q) CK(cudaMemcpyAsync(cnt[q], G.flags_of(q) + fwcomm::F_LIST_N, 2 * sizeof(u64), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        { int st_ = comm_check(ctx); if (st_ != FW_OK) return st_; }
        for (int q = 0; q < G.world; ++q) {
            NEED((i64)cnt[q][0] <= G.list_cap, FW_ERR_UNSUPPORTED, "fw_pairwise: rank %d collected %llu raw candidates, more than the group's list capacity %lld (univariate network too dense)",
                 q, (unsigned long long)cnt[q][0], (long long)G.list_cap);
            nrec += (i64)cnt[q][0]; n_nan += (i64)cnt[q][1];
        }
        CK(ctx->d_gathered.reserve((size_t)std::max<i64>(nrec, 1)));
        i64 o = 0;
        for (int q = 0; q < G.world; ++q) {
            if (cnt[q][0]) CK(cudaMemcpyAsync(ctx->d_gathered.ptr + o, G.list_of(q), sizeof(PwRec) * cnt[q][0], cudaMemcpyDeviceToDevice, st));
            o += (i64)cnt[q][0];
        }
        recs = ctx->d_gathered.ptr;
    }
    cudaError_t e = pairwise_fz_tail(ctx->pw, recs, nrec, n_nan, p, fc, ctx->n_obs, n_obs_min, alpha, fdr, rel_only, st, po, nl, &msg);
    if (e != cudaSuccess && msg.find("unsupported size") != std::string::npos) return fail(ctx, FW_ERR_UNSUPPORTED, "fw_pairwise: %s", msg.c_str());
    if (e != cudaSuccess) return fail(ctx, FW_ERR_CUDA, "fw_pairwise: %s: %s", msg.c_str(), cudaGetErrorString(e));
    return FW_OK;
}

extern "C" {

const char* fw_build_info(void) {
#define FW_STR2(x) #x
#define FW_STR(x) FW_STR2(x)
    return "libfwgpu sm_100a, nvcc " FW_STR(__CUDACC_VER_MAJOR__) "." FW_STR(__CUDACC_VER_MINOR__) "." FW_STR(__CUDACC_VER_BUILD__) ", host gcc " __VERSION__;
}

int32_t fw_create(int32_t device, fw_ctx** out) {
    fw_ctx* ctx = nullptr;
    if (!out) return fail(nullptr, FW_ERR_INVALID, "fw_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, FW_ERR_CUDA, "fw_create: no CUDA device available (%s); libfwgpu has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, FW_ERR_INVALID, "fw_create: device %d out of range [0,%d)", device, ndev);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, FW_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(nullptr, FW_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10) return fail(nullptr, FW_ERR_UNSUPPORTED, "fw_create: device is sm_%d%d; libfwgpu is built for sm_100a only", prop.major, prop.minor);
    ctx = new fw_ctx();
    ctx->device = device; ctx->sm_count = prop.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete ctx; return fail(nullptr, FW_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    if (ctx->d_counter.reserve(16) != cudaSuccess || ctx->d_exec.reserve(16) != cudaSuccess) { delete ctx; return fail(nullptr, FW_ERR_NOMEM, "scratch allocation failed"); }
    for (int i = 0; i < 8; ++i) if (cudaEventCreate(&ctx->ev[i]) != cudaSuccess) { delete ctx; return fail(nullptr, FW_ERR_CUDA, "cudaEventCreate failed"); }
    for (int i = 0; i < 3; ++i) if (cudaEventCreate(&ctx->evx[i]) != cudaSuccess) { delete ctx; return fail(nullptr, FW_ERR_CUDA, "cudaEventCreate failed"); }
    if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return fail(nullptr, FW_ERR_CUDA, "cudaStreamCreate failed"); }
    for (int i = 0; i < 12; ++i) if (cudaEventCreateWithFlags(&ctx->chunk_ev[i], cudaEventDisableTiming) != cudaSuccess) { delete ctx; return fail(nullptr, FW_ERR_CUDA, "cudaEventCreate failed"); }
    *out = ctx;
    return FW_OK;
}

int32_t fw_destroy(fw_ctx* ctx) {
    if (!ctx) return FW_OK;
    cudaSetDevice(ctx->device);
    fw_comm_detach(ctx);
    if (ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
    for (int i = 0; i < 8; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i < 3; ++i) if (ctx->evx[i]) cudaEventDestroy(ctx->evx[i]);
    for (int i = 0; i < 12; ++i) if (ctx->chunk_ev[i]) cudaEventDestroy(ctx->chunk_ev[i]);
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    delete ctx;
    return FW_OK;
}

// page-lock / unlock a caller-owned host buffer (result arrays that are reused across calls then move at full PCIe speed)
int32_t fw_host_register(fw_ctx* ctx, void* ptr, int64_t bytes) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(ptr && bytes > 0, FW_ERR_INVALID, "fw_host_register: bad arguments");
    CK(cudaSetDevice(ctx->device));
    CK(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault));
    return FW_OK;
}
int32_t fw_host_unregister(fw_ctx* ctx, void* ptr) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(ptr, FW_ERR_INVALID, "fw_host_unregister: NULL");
    CK(cudaSetDevice(ctx->device));
    CK(cudaHostUnregister(ptr));
    return FW_OK;
}

const char* fw_last_error(fw_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int32_t fw_set_index_base(fw_ctx* ctx, int32_t base) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(base == 0 || base == 1, FW_ERR_INVALID, "index base must be 0 or 1");
    ctx->index_base = base; return FW_OK;
}
void* fw_stream(fw_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int32_t fw_synchronize(fw_ctx* ctx) { if (!ctx) return FW_ERR_INVALID; CK(cudaSetDevice(ctx->device)); CK(cudaStreamSynchronize(ctx->stream)); return comm_check(ctx); }
int64_t fw_launch_count(fw_ctx* ctx) { return ctx ? ctx->launches : 0; }

// device time (ms, CUDA events on the context's stream) of the last run of each phase:
// out[0] = cor_mat kernels, out[1] = pairwise stage kernels, out[2] = HITON-PC kernel(s), out[3] = subset-search kernel(s); -1 = not run
int32_t fw_last_timing(fw_ctx* ctx, double* out_ms, int32_t n) {
    if (!ctx || !out_ms) return FW_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    for (int i = 0; i < n && i < 4; ++i) {
        out_ms[i] = -1.0;
        if (!ctx->ev_valid[i]) continue;
        float ms = 0.f;
        CK(cudaEventSynchronize(ctx->ev[2 * i + 1]));
        CK(cudaEventElapsedTime(&ms, ctx->ev[2 * i], ctx->ev[2 * i + 1]));
        out_ms[i] = (double)ms;
    }
    // [4] = standardise + split part of out[0] in fw_multi_cor (reads the peers' table slices), [5] = wait in its closing group barrier
    for (int i = 4; i < n && i < 6; ++i) {
        out_ms[i] = -1.0;
        if (!ctx->evx_valid) continue;
        float ms = 0.f;
        CK(cudaEventSynchronize(ctx->evx[2]));
        if (i == 4) CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->evx[0])); else CK(cudaEventElapsedTime(&ms, ctx->evx[1], ctx->evx[2]));
        out_ms[i] = (double)ms;
    }
    return FW_OK;
}

// ---- data -----------------------------------------------------------------------------
int32_t fw_set_data_f32(fw_ctx* ctx, const float* host, int64_t n, int64_t p, int64_t ld) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(host && n > 0 && p > 0 && ld >= n, FW_ERR_INVALID, "fw_set_data_f32: bad arguments (n=%lld p=%lld ld=%lld)", (long long)n, (long long)p, (long long)ld);
    CK(cudaSetDevice(ctx->device));
    CK(ctx->d_data_f32.reserve((size_t)n * p));
    CK(cudaMemcpy2DAsync(ctx->d_data_f32.ptr, n * sizeof(float), host, ld * sizeof(float), n * sizeof(float), p, cudaMemcpyHostToDevice, ctx->stream));
    ctx->n = n; ctx->p = p; ctx->ld = n; ctx->data_kind = 0; ctx->n_obs = n;
    table_changed(ctx);
    CK(cudaStreamSynchronize(ctx->stream));          // host pointers are borrowed for the duration of the call only (page-locked sources copy asynchronously)
    return FW_OK;
}
int32_t fw_adopt_data_f32_device(fw_ctx* ctx, const float* dev, int64_t n, int64_t p, int64_t ld) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(dev && n > 0 && p > 0 && ld >= n, FW_ERR_INVALID, "fw_adopt_data_f32_device: bad arguments");
    ctx->d_data_f32.adopt(const_cast<float*>(dev), (size_t)ld * p);
    ctx->n = n; ctx->p = p; ctx->ld = ld; ctx->data_kind = 0; ctx->n_obs = n;
    table_changed(ctx);
    return FW_OK;
}
// get_levels / get_max_vals (src/misc.jl:64-97) and the bit-plane table of the level codes resident in ctx->d_data_i32 ([p][n])
static int install_discrete_table(fw_ctx* ctx, int64_t n, int64_t p, const char* who) {
    NEED(n < ((i64)1 << 31) - 64, FW_ERR_UNSUPPORTED, "%s: more than 2^31 rows", who);
    CK(ctx->d_levels.reserve(p)); CK(ctx->d_maxvals.reserve(p)); CK(ctx->d_nnz.reserve(p)); CK(ctx->d_bad.reserve(4));
    CK(cudaMemsetAsync(ctx->d_bad.ptr, 0, sizeof(int), ctx->stream));
    {
        const int T = 256; const i64 blocks = (p * 32 + T - 1) / T;
        mi_levels_kernel<<<(unsigned)blocks, T, 0, ctx->stream>>>(ctx->d_data_i32.ptr, n, n, p, ctx->d_levels.ptr, ctx->d_maxvals.ptr, ctx->d_nnz.ptr, ctx->d_bad.ptr);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    ctx->h_levels.resize(p); ctx->h_maxvals.resize(p);
    int bad = 0;
    CK(cudaMemcpyAsync(ctx->h_levels.data(), ctx->d_levels.ptr, sizeof(int) * p, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_maxvals.data(), ctx->d_maxvals.ptr, sizeof(int) * p, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&bad, ctx->d_bad.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->data_kind = -1;
    NEED(!(bad & 1), FW_ERR_INVALID, "%s: negative level codes", who);
    int mx = 0; for (i64 v = 0; v < p; ++v) mx = std::max(mx, ctx->h_maxvals[v]);
    NEED(mx + 1 <= FW_MAX_L, FW_ERR_UNSUPPORTED, "%s: %d levels; the bit-plane engine supports at most %d (maximum(max_vals)+1)", who, mx + 1, FW_MAX_L);
    ctx->disc_L = std::max(mx + 1, 2); ctx->disc_W = (int)((n + 31) / 32);
    CK(ctx->d_planes.reserve((size_t)p * (ctx->disc_L - 1) * ctx->disc_W));
    {
        const int T = 256; const i64 warps = p * ctx->disc_W; const i64 blocks = (warps * 32 + T - 1) / T;
        mi_pack_planes_kernel<<<(unsigned)blocks, T, 0, ctx->stream>>>(ctx->d_data_i32.ptr, n, n, p, ctx->disc_L, ctx->disc_W, ctx->d_planes.ptr);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    CK(ctx->d_logtab.reserve((size_t)n + 1));
    mi_logtab_kernel<<<(unsigned)((n + 256) / 256), 256, 0, ctx->stream>>>((int)n, ctx->d_logtab.ptr);
    ctx->launches++;
    CK(cudaGetLastError());
    ctx->n = n; ctx->p = p; ctx->ld = n; ctx->data_kind = 1; ctx->n_obs = n;
    table_changed(ctx);
    return FW_OK;
}
int32_t fw_set_data_i32(fw_ctx* ctx, const int32_t* host, int64_t n, int64_t p, int64_t ld) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(host && n > 0 && p > 0 && ld >= n, FW_ERR_INVALID, "fw_set_data_i32: bad arguments");
    CK(cudaSetDevice(ctx->device));
    CK(ctx->d_data_i32.reserve((size_t)n * p));
    CK(cudaMemcpy2DAsync(ctx->d_data_i32.ptr, n * sizeof(int), host, ld * sizeof(int), n * sizeof(int), p, cudaMemcpyHostToDevice, ctx->stream));
    return install_discrete_table(ctx, n, p, "fw_set_data_i32");
}
int32_t fw_set_data_csc_f32(fw_ctx* ctx, const int64_t* colptr, const int64_t* rowval, const float* nzval, int64_t n, int64_t p) {
    if (!ctx) return FW_ERR_INVALID;
    int st_ = set_data_csc<float>(ctx, colptr, rowval, nzval, n, p, ctx->d_data_f32, "fw_set_data_csc_f32");
    if (st_ != FW_OK) return st_;
    ctx->n = n; ctx->p = p; ctx->ld = n; ctx->data_kind = 0; ctx->n_obs = n;
    table_changed(ctx);
    return FW_OK;
}
int32_t fw_set_data_csc_i32(fw_ctx* ctx, const int64_t* colptr, const int64_t* rowval, const int32_t* nzval, int64_t n, int64_t p) {
    if (!ctx) return FW_ERR_INVALID;
    int st_ = set_data_csc<int>(ctx, colptr, rowval, nzval, n, p, ctx->d_data_i32, "fw_set_data_csc_i32");
    if (st_ != FW_OK) return st_;
    return install_discrete_table(ctx, n, p, "fw_set_data_csc_i32");
}
// meta_variable_mask (src/preprocessing.jl:418-446, written by src/io.jl:338-346): which variables are meta variables.  The tests treat
// meta variables like any other variable (they differ in preprocessing only), so the mask is carried for the writers, not used on the device.
int32_t fw_set_meta_mask(fw_ctx* ctx, const uint8_t* mask, int64_t p) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(p >= 0 && (p == 0 || mask) && (ctx->p == 0 || p == ctx->p), FW_ERR_INVALID, "fw_set_meta_mask: %lld entries for %lld variables", (long long)p, (long long)ctx->p);
    ctx->meta_mask.assign(mask, mask + p);
    return FW_OK;
}
int32_t fw_get_meta_mask(fw_ctx* ctx, uint8_t* mask_out, int64_t p) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(mask_out && p == ctx->p, FW_ERR_INVALID, "fw_get_meta_mask: %lld entries for %lld variables", (long long)p, (long long)ctx->p);
    for (i64 i = 0; i < p; ++i) mask_out[i] = i < (i64)ctx->meta_mask.size() ? ctx->meta_mask[i] : 0;
    return FW_OK;
}
int32_t fw_set_semantics(fw_ctx* ctx, int32_t semantics) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(semantics == FW_SEMANTICS_DENSE || semantics == FW_SEMANTICS_SPARSE, FW_ERR_INVALID, "fw_set_semantics: unknown value %d", semantics);
    ctx->sparse_sem = semantics;
    return FW_OK;
}
int32_t fw_set_n_obs(fw_ctx* ctx, int64_t n) { if (!ctx) return FW_ERR_INVALID; NEED(n >= 0, FW_ERR_INVALID, "n_obs < 0"); ctx->n_obs = n; return FW_OK; }

int32_t fw_levels(fw_ctx* ctx, int32_t* levels, int32_t* max_vals) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(ctx->data_kind == 1, FW_ERR_STATE, "fw_levels: no discrete table resident (call fw_set_data_i32 first)");
    if (levels) memcpy(levels, ctx->h_levels.data(), sizeof(int) * ctx->p);
    if (max_vals) memcpy(max_vals, ctx->h_maxvals.data(), sizeof(int) * ctx->p);
    return FW_OK;
}

// ---- normalisation (the step in front of the hot path; prep.cuh) -----------------------------------------------------
int32_t fw_normalize_f32(fw_ctx* ctx, const float* host, int64_t n, int64_t p, int64_t ld, int32_t norm, int32_t n_bins,
                         int64_t* n_out, int64_t* p_out, uint8_t* row_mask, uint8_t* col_mask) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(host && n > 0 && p > 0 && ld >= n, FW_ERR_INVALID, "fw_normalize_f32: bad arguments (n=%lld p=%lld ld=%lld)", (long long)n, (long long)p, (long long)ld);
    NEED(norm >= PREP_ROWS && norm <= PREP_BINNED_NZ_ROWS, FW_ERR_INVALID, "fw_normalize_f32: unknown normalisation mode %d", norm);
    const bool binned = norm == PREP_BINNED_NZ_CLR || norm == PREP_BINNED_NZ_ROWS;
    NEED(!binned || (n_bins >= 2 && n_bins <= FW_MAX_L), FW_ERR_UNSUPPORTED, "fw_normalize_f32: n_bins = %d not in 2..%d", n_bins, FW_MAX_L);
    NEED(n < ((i64)1 << 31) - 64 && p < ((i64)1 << 31) - 64 && n <= (i64)65535 * 256, FW_ERR_UNSUPPORTED, "fw_normalize_f32: table too large");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int T = 256;
    DevBuf<float> raw; DevBuf<unsigned char> dflag;
    CK(raw.reserve((size_t)n * p)); CK(dflag.reserve(p));
    CK(cudaMemcpy2DAsync(raw.ptr, n * sizeof(float), host, ld * sizeof(float), n * sizeof(float), p, cudaMemcpyHostToDevice, st));
    // filter_by_variance (preprocessing.jl:367-409): variables first, then samples on the kept variables
    prep_col_distinct_kernel<<<(unsigned)p, T, 0, st>>>(raw.ptr, n, n, dflag.ptr); ctx->launches++;
    CK(cudaGetLastError());
    std::vector<unsigned char> cflag(p);
    CK(cudaMemcpyAsync(cflag.data(), dflag.ptr, p, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    std::vector<int> cols; for (i64 v = 0; v < p; ++v) if (cflag[v] && n > 1) cols.push_back((int)v);
    i64 p1 = (i64)cols.size();
    std::vector<unsigned char> rflag(n, 0);
    std::vector<int> rows;
    DevBuf<int> dcols, drows, dnz0; DevBuf<float> dsum32; DevBuf<double> dsum64, dslog, dg, dpc;
    std::vector<float> hsum32(n); std::vector<double> hsum64(n), hslog(n), hg(n, 1.0), hpc(n, 0.0); std::vector<int> hnz0(n);
    if (p1 > 0) {
        CK(dcols.reserve(p1)); CK(dnz0.reserve(n)); CK(dsum32.reserve(n)); CK(dsum64.reserve(n)); CK(dslog.reserve(n));
        CK(cudaMemcpyAsync(dcols.ptr, cols.data(), sizeof(int) * p1, cudaMemcpyHostToDevice, st));
        prep_row_stats_kernel<<<(unsigned)((n + T - 1) / T), T, 0, st>>>(raw.ptr, n, n, dcols.ptr, p1, dsum32.ptr, dsum64.ptr, dnz0.ptr, dslog.ptr); ctx->launches++;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(hsum32.data(), dsum32.ptr, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(hsum64.data(), dsum64.ptr, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(hslog.data(), dslog.ptr, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(hnz0.data(), dnz0.ptr, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (i64 r = 0; r < n; ++r) if (hsum32[r] > 0.0f) { rflag[r] = 1; rows.push_back((int)r); }
    }
    // per-sample parameters on the host (O(n); fp64 exp/log as the reference evaluates them)
    if (!rows.empty() && (norm == PREP_CLR_NZ || norm == PREP_BINNED_NZ_CLR)) {
        for (int r : rows) hg[r] = std::exp(hslog[r] / (double)(p1 - hnz0[r]));                       // geomean of the non-zero entries
    } else if (!rows.empty() && norm == PREP_CLR_ADAPT) {
        // adaptive_pseudocount! (preprocessing.jl:157-175): deepest sample, smallest abundance, per-sample pseudo-counts
        int md = rows[0]; for (int r : rows) if (hsum64[r] > hsum64[md]) md = r;                     // findmax: first maximum
        CK(drows.reserve(rows.size()));
        CK(cudaMemcpyAsync(drows.ptr, rows.data(), sizeof(int) * rows.size(), cudaMemcpyHostToDevice, st));
        unsigned int hbits = 0x7f800000u;
        CK(cudaMemcpyAsync(ctx->d_counter.ptr, &hbits, sizeof(unsigned int), cudaMemcpyHostToDevice, st));
        prep_min_nonzero_kernel<<<(unsigned)p1, T, 0, st>>>(raw.ptr, n, dcols.ptr, p1, drows.ptr, (i64)rows.size(), reinterpret_cast<unsigned int*>(ctx->d_counter.ptr)); ctx->launches++;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&hbits, ctx->d_counter.ptr, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        float min_ab; memcpy(&min_ab, &hbits, sizeof(float));
        const double base = min_ab >= 1.0f ? 1.0 : (double)min_ab / 10.0;
        const double k = (double)hnz0[md], nprod1 = hslog[md], pv = (double)p1;
        std::vector<int> kept;
        for (int r : rows) {
            const double nz0 = (double)hnz0[r];
            const double pcv = std::exp((1.0 / (nz0 - pv)) * ((k - pv) * std::log(base) + nprod1 - hslog[r]));
            if (pcv != 0.0) { hpc[r] = pcv; hg[r] = std::exp((hslog[r] + nz0 * std::log(pcv)) / pv); kept.push_back(r); }
            else rflag[r] = 0;                                                                       // pseudo-count below machine precision: sample removed
        }
        rows.swap(kept);
    }
    const i64 n1 = (i64)rows.size();
    std::vector<unsigned char> cmask(p, 0);
    for (int v : cols) cmask[v] = 1;
    i64 p2 = p1;
    if (n1 == 0 || p1 == 0) {
        ctx->data_kind = -1; ctx->n = 0; ctx->p = 0; table_changed(ctx);
        if (n_out) *n_out = 0; if (p_out) *p_out = 0;
        if (row_mask) memcpy(row_mask, rflag.data(), n);
        if (col_mask) memset(col_mask, 0, p);
        return FW_OK;
    }
    CK(drows.reserve(n1));
    CK(cudaMemcpyAsync(drows.ptr, rows.data(), sizeof(int) * n1, cudaMemcpyHostToDevice, st));
    CK(dg.reserve(n)); CK(dpc.reserve(n));
    CK(cudaMemcpyAsync(dg.ptr, hg.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dpc.ptr, hpc.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st));
    const dim3 grid((unsigned)p1, (unsigned)((n1 + T - 1) / T));
    if (norm == PREP_ROWS || norm == PREP_CLR_NZ || norm == PREP_CLR_ADAPT) {
        CK(ctx->d_data_f32.reserve((size_t)n1 * p1));
        prep_transform_kernel<<<grid, T, 0, st>>>(raw.ptr, n, dcols.ptr, drows.ptr, n1, norm, dsum32.ptr, dg.ptr, dpc.ptr, ctx->d_data_f32.ptr); ctx->launches++;
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(st));
        ctx->n = n1; ctx->p = p1; ctx->ld = n1; ctx->data_kind = 0; ctx->n_obs = n1;
        table_changed(ctx);
    } else {
        DevBuf<int> tmp; DevBuf<unsigned int> dseen;
        CK(tmp.reserve((size_t)n1 * p1)); CK(dseen.reserve(p1));
        if (norm == PREP_BINARY) {
            prep_binary_kernel<<<grid, T, 0, st>>>(raw.ptr, n, dcols.ptr, drows.ptr, n1, tmp.ptr); ctx->launches++;
        } else {
            NEED((i64)n1 * p1 < ((i64)1 << 31), FW_ERR_UNSUPPORTED, "fw_normalize_f32: more than 2^31 entries in a binned table");
            DevBuf<double> vals, sorted; DevBuf<int> dnnz, doffs;
            CK(vals.reserve((size_t)n1 * p1)); CK(sorted.reserve((size_t)n1 * p1)); CK(dnnz.reserve(p1)); CK(doffs.reserve(p1 + 1));
            CK(cudaMemsetAsync(dnnz.ptr, 0, sizeof(int) * p1, st));
            prep_rank_values_kernel<<<grid, T, 0, st>>>(raw.ptr, n, dcols.ptr, drows.ptr, n1, norm, dsum32.ptr, dg.ptr, vals.ptr, dnnz.ptr); ctx->launches++;
            CK(cudaGetLastError());
            std::vector<int> offs(p1 + 1); for (i64 j = 0; j <= p1; ++j) offs[j] = (int)(j * n1);
            CK(cudaMemcpyAsync(doffs.ptr, offs.data(), sizeof(int) * (p1 + 1), cudaMemcpyHostToDevice, st));
            size_t need = 0;
            CK(cub::DeviceSegmentedSort::SortKeys(nullptr, need, vals.ptr, sorted.ptr, (int)(n1 * p1), (int)p1, doffs.ptr, doffs.ptr + 1, st));
            DevBuf<unsigned char> tmpsort; CK(tmpsort.reserve(need));
            CK(cub::DeviceSegmentedSort::SortKeys(tmpsort.ptr, need, vals.ptr, sorted.ptr, (int)(n1 * p1), (int)p1, doffs.ptr, doffs.ptr + 1, st)); ctx->launches += 3;
            prep_bin_kernel<<<grid, T, 0, st>>>(vals.ptr, sorted.ptr, dnnz.ptr, n1, n_bins, tmp.ptr); ctx->launches++;
            CK(cudaGetLastError());
            CK(cudaStreamSynchronize(st));                       // the sort scratch is released at scope exit
        }
        CK(cudaGetLastError());
        // level filter: exactly 2 levels (binary, preprocessing.jl:479-481) / exactly n_bins - 1 distinct non-zero levels (:511-513)
        CK(cudaMemsetAsync(dseen.ptr, 0, sizeof(unsigned int) * p1, st));
        prep_col_levels_kernel<<<(unsigned)p1, T, 0, st>>>(tmp.ptr, n1, dseen.ptr); ctx->launches++;
        CK(cudaGetLastError());
        std::vector<unsigned int> seen(p1);
        CK(cudaMemcpyAsync(seen.data(), dseen.ptr, sizeof(unsigned int) * p1, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        std::vector<int> colmap;
        for (i64 j = 0; j < p1; ++j) {
            const int lv_all = __builtin_popcount(seen[j]), lv_nz = __builtin_popcount(seen[j] & ~1u);
            const bool keep = norm == PREP_BINARY ? lv_all == 2 : lv_nz == n_bins - 1;
            if (keep) colmap.push_back((int)j); else cmask[cols[j]] = 0;
        }
        p2 = (i64)colmap.size();
        if (p2 == 0) {
            ctx->data_kind = -1; ctx->n = 0; ctx->p = 0; table_changed(ctx);
            if (n_out) *n_out = n1; if (p_out) *p_out = 0;
            if (row_mask) memcpy(row_mask, rflag.data(), n);
            if (col_mask) memset(col_mask, 0, p);
            return FW_OK;
        }
        DevBuf<int> dmap; CK(dmap.reserve(p2));
        CK(cudaMemcpyAsync(dmap.ptr, colmap.data(), sizeof(int) * p2, cudaMemcpyHostToDevice, st));
        CK(ctx->d_data_i32.reserve((size_t)n1 * p2));
        prep_gather_cols_kernel<<<dim3((unsigned)p2, (unsigned)((n1 + T - 1) / T)), T, 0, st>>>(tmp.ptr, dmap.ptr, n1, ctx->d_data_i32.ptr); ctx->launches++;
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(st));
        int st_ = install_discrete_table(ctx, n1, p2, "fw_normalize_f32");
        if (st_ != FW_OK) return st_;
    }
    if (n_out) *n_out = n1; if (p_out) *p_out = p2;
    if (row_mask) memcpy(row_mask, rflag.data(), n);
    if (col_mask) memcpy(col_mask, cmask.data(), p);
    return FW_OK;
}

int32_t fw_get_data_f32(fw_ctx* ctx, float* host_out, int64_t ld) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(ctx->data_kind == 0 && ctx->d_data_f32.ptr, FW_ERR_STATE, "fw_get_data_f32: no continuous table resident");
    NEED(host_out && ld >= ctx->n, FW_ERR_INVALID, "fw_get_data_f32: bad arguments");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy2DAsync(host_out, ld * sizeof(float), ctx->d_data_f32.ptr, ctx->ld * sizeof(float), ctx->n * sizeof(float), ctx->p, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return FW_OK;
}
int32_t fw_get_data_i32(fw_ctx* ctx, int32_t* host_out, int64_t ld) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(ctx->data_kind == 1 && ctx->d_data_i32.ptr, FW_ERR_STATE, "fw_get_data_i32: no discrete table resident");
    NEED(host_out && ld >= ctx->n, FW_ERR_INVALID, "fw_get_data_i32: bad arguments");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy2DAsync(host_out, ld * sizeof(int), ctx->d_data_i32.ptr, ctx->n * sizeof(int), ctx->n * sizeof(int), ctx->p, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return FW_OK;
}

// ---- cor_mat ----------------------------------------------------------------------------
int32_t fw_set_cor_f32(fw_ctx* ctx, const float* host_cor, int64_t p) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(host_cor && p > 0, FW_ERR_INVALID, "fw_set_cor_f32: bad arguments");
    CK(cudaSetDevice(ctx->device));
    CK(ctx->d_cor.reserve((size_t)p * p));
    CK(cudaMemcpyAsync(ctx->d_cor.ptr, host_cor, sizeof(float) * (size_t)p * p, cudaMemcpyHostToDevice, ctx->stream));
    ctx->cor_p = p; if (ctx->p == 0) ctx->p = p;
    ctx->cor_sharded = false; ctx->col.valid = false;
    CK(cudaStreamSynchronize(ctx->stream));          // the host buffer is borrowed for the duration of the call only
    return FW_OK;
}
int32_t fw_adopt_cor_device(fw_ctx* ctx, const float* dev_cor, int64_t p) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(dev_cor && p > 0, FW_ERR_INVALID, "fw_adopt_cor_device: bad arguments");
    ctx->d_cor.adopt(const_cast<float*>(dev_cor), (size_t)p * p);
    ctx->cor_p = p; if (ctx->p == 0) ctx->p = p;
    ctx->cor_sharded = false; ctx->col.valid = false;
    return FW_OK;
}
int32_t fw_adopt_cor_device_rows(fw_ctx* ctx, const float* dev_cor, int64_t p, int64_t rows_allocated) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(dev_cor && p > 0 && rows_allocated >= p, FW_ERR_INVALID, "fw_adopt_cor_device_rows: bad arguments");
    ctx->d_cor.adopt(const_cast<float*>(dev_cor), (size_t)rows_allocated * p);
    ctx->cor_p = p; if (ctx->p == 0) ctx->p = p;
    ctx->cor_sharded = false; ctx->col.valid = false;
    return FW_OK;
}
void* fw_cor_device_ptr(fw_ctx* ctx) { return ctx ? (void*)ctx->d_cor.ptr : nullptr; }

// fw_set_data_f32 + fw_cor_matrix in one call, with the PCIe upload hidden behind the GEMM (column chunks on a copy stream)
int32_t fw_upload_cor_f32(fw_ctx* ctx, const float* host, int64_t n, int64_t p, int64_t ld, float* host_out) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(host && n > 0 && p > 0 && ld >= n, FW_ERR_INVALID, "fw_upload_cor_f32: bad arguments (n=%lld p=%lld ld=%lld)", (long long)n, (long long)p, (long long)ld);
    CK(cudaSetDevice(ctx->device));
    CK(ctx->d_data_f32.reserve((size_t)n * p));
    if (!(ctx->d_cor.ptr && ctx->d_cor.owned && ctx->d_cor.cap >= (size_t)p * p)) CK(ctx->d_cor.reserve((size_t)p * p));
    std::string msg; int nl = 0;
    ctx->n = n; ctx->p = p; ctx->ld = n; ctx->data_kind = 0; ctx->n_obs = n;
    table_changed(ctx);
    PwEmit em; { int st_ = arm_collect(ctx, false, &em); if (st_ != FW_OK) return st_; }
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    cudaError_t e = cortc::run_overlapped(ctx->tc, host, ctx->d_data_f32.ptr, n, p, ld, ctx->d_cor.ptr, ctx->stream, ctx->copy_stream, ctx->chunk_ev, 12, &nl, &msg, em);
    ctx->launches += nl;
    if (e != cudaSuccess) return fail(ctx, FW_ERR_CUDA, "fw_upload_cor_f32: %s: %s", msg.c_str(), cudaGetErrorString(e));
    CK(cudaEventRecord(ctx->ev[1], ctx->stream)); ctx->ev_valid[0] = true;
    ctx->cor_p = p; ctx->cor_sharded = false; ctx->col.valid = ctx->col.armed;
    CK(cudaStreamSynchronize(ctx->copy_stream));     // every chunk has left the (borrowed) host buffer; the GEMM tail may still be running
    if (host_out) {
        CK(cudaMemcpyAsync(host_out, ctx->d_cor.ptr, sizeof(float) * (size_t)p * p, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return FW_OK;
}

// ---- row-sharded cor_mat (multi-GPU) -------------------------------------------------------------
// fw_cor_prepare: standardise + split the resident table once.  fw_cor_rows: the upper-triangular 128x128 tiles of tile rows
// [tile_row_begin, tile_row_end) into the resident / adopted cor_mat buffer (which must hold at least tile_row_end*128 rows of
// p floats).  After the ranks have exchanged their row blocks (NCCL all-gather on the adopted buffer), fw_cor_symmetrize copies
// the upper triangle to the lower one; the result is bit-identical to fw_cor_matrix on one GPU.
int32_t fw_cor_prepare(fw_ctx* ctx, int32_t* n_tile_rows) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(ctx->data_kind == 0, FW_ERR_STATE, "fw_cor_prepare: no continuous table resident (call fw_set_data_f32 first)");
    CK(cudaSetDevice(ctx->device));
    std::string msg; int nl = 0;
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    cudaError_t e = cortc::prepare(ctx->tc, ctx->tcp, ctx->d_data_f32.ptr, ctx->n, ctx->p, ctx->ld, ctx->stream, &nl, &msg);
    ctx->launches += nl;
    if (e != cudaSuccess) return fail(ctx, FW_ERR_CUDA, "fw_cor_prepare: %s: %s", msg.c_str(), cudaGetErrorString(e));
    if (n_tile_rows) *n_tile_rows = ctx->tcp.nb;
    return FW_OK;
}
int32_t fw_cor_rows(fw_ctx* ctx, int32_t tile_row_begin, int32_t tile_row_end) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(ctx->tcp.valid, FW_ERR_STATE, "fw_cor_rows: call fw_cor_prepare first");
    NEED(ctx->d_cor.ptr && ctx->cor_p == ctx->p, FW_ERR_STATE, "fw_cor_rows: no cor_mat buffer (fw_adopt_cor_device)");
    NEED(tile_row_begin >= 0 && tile_row_end >= tile_row_begin, FW_ERR_INVALID, "fw_cor_rows: bad tile-row range");
    NEED(ctx->d_cor.cap >= (size_t)std::min<i64>((i64)tile_row_end * 128, ctx->p) * ctx->p, FW_ERR_INVALID, "fw_cor_rows: cor_mat buffer too small");
    CK(cudaSetDevice(ctx->device));
    std::string msg; int nl = 0;
    cudaError_t e = cortc::run_rows(ctx->tcp, ctx->d_cor.ptr, ctx->p, tile_row_begin, tile_row_end, false, ctx->stream, &nl, &msg);
    ctx->launches += nl;
    if (e != cudaSuccess) return fail(ctx, FW_ERR_CUDA, "fw_cor_rows: %s: %s", msg.c_str(), cudaGetErrorString(e));
    CK(cudaEventRecord(ctx->ev[1], ctx->stream)); ctx->ev_valid[0] = true;
    return FW_OK;
}
int32_t fw_cor_symmetrize(fw_ctx* ctx) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(has_cor(ctx), FW_ERR_STATE, "fw_cor_symmetrize: no cor_mat resident");
    CK(cudaSetDevice(ctx->device));
    const unsigned nb = (unsigned)((ctx->cor_p + 31) / 32);
    cortc::cor_symmetrize_kernel<<<dim3(nb, nb), 256, 0, ctx->stream>>>(ctx->d_cor.ptr, ctx->cor_p);
    ctx->launches++;
    CK(cudaGetLastError());
    return FW_OK;
}

int32_t fw_cor_matrix(fw_ctx* ctx, float* host_out) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(ctx->data_kind == 0, FW_ERR_STATE, "fw_cor_matrix: no continuous table resident (call fw_set_data_f32 first)");
    CK(cudaSetDevice(ctx->device));
    i64 p = ctx->p;
    if (!(ctx->d_cor.ptr && ctx->d_cor.owned && ctx->d_cor.cap >= (size_t)p * p)) CK(ctx->d_cor.reserve((size_t)p * p));
    std::string msg;
    int nl = 0;
    PwEmit em; { int st_ = arm_collect(ctx, false, &em); if (st_ != FW_OK) return st_; }
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    cudaError_t e = cortc::run(ctx->tc, ctx->d_data_f32.ptr, ctx->n, p, ctx->ld, ctx->d_cor.ptr, ctx->stream, &nl, &msg, em);
    ctx->launches += nl;
    if (e != cudaSuccess) return fail(ctx, FW_ERR_CUDA, "fw_cor_matrix: %s: %s", msg.c_str(), cudaGetErrorString(e));
    CK(cudaEventRecord(ctx->ev[1], ctx->stream)); ctx->ev_valid[0] = true;
    ctx->cor_p = p; ctx->cor_sharded = false; ctx->col.valid = ctx->col.armed;
    if (host_out) {
        CK(cudaMemcpyAsync(host_out, ctx->d_cor.ptr, sizeof(float) * (size_t)p * p, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return FW_OK;
}

int32_t fw_pairwise_prefetch(fw_ctx* ctx, double alpha, int64_t n_obs_min) {
    if (!ctx) return FW_ERR_INVALID;
    ctx->col.armed = alpha > 0.0;                    // alpha <= 0 disarms
    ctx->col.alpha = alpha; ctx->col.n_obs_min = n_obs_min; ctx->col.valid = false;
    return FW_OK;
}

int32_t fw_cor_gather(fw_ctx* ctx, const int64_t* idx, int64_t m, float* host_out) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(has_cor(ctx), FW_ERR_STATE, "fw_cor_gather: no cor_mat resident");
    NEED(m >= 0 && (m == 0 || (idx && host_out)), FW_ERR_INVALID, "fw_cor_gather: NULL argument");
    if (m == 0) return FW_OK;
    CK(cudaSetDevice(ctx->device));
    std::vector<i64> h(m);
    for (i64 i = 0; i < m; ++i) { h[i] = idx[i] - ctx->index_base; NEED(h[i] >= 0 && h[i] < ctx->cor_p, FW_ERR_INVALID, "fw_cor_gather: variable index out of range"); }
    DevBuf<i64> di; DevBuf<float> dout;
    CK(di.reserve(m)); CK(dout.reserve((size_t)m * m));
    CK(cudaMemcpyAsync(di.ptr, h.data(), sizeof(i64) * m, cudaMemcpyHostToDevice, ctx->stream));
    fwcomm::cor_gather_kernel<<<(unsigned)((m * m + 255) / 256), 256, 0, ctx->stream>>>(make_cor_view(ctx), di.ptr, m, dout.ptr);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(host_out, dout.ptr, sizeof(float) * (size_t)m * m, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return FW_OK;
}

// ---- multi-GPU group (comm.cuh) ---------------------------------------------------------------------------------------------
int32_t fw_comm_handle_bytes(void) { return (int32_t)sizeof(fwcomm::Handle); }

int32_t fw_comm_export(fw_ctx* ctx, int32_t rank, int32_t world, int64_t n, int64_t p, void* handle_out) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(handle_out && world >= 1 && world <= FW_MAX_RANKS && rank >= 0 && rank < world && n > 0 && p >= world, FW_ERR_INVALID,
         "fw_comm_export: bad arguments (rank %d of %d, n = %lld, p = %lld; at most %d ranks)", rank, world, (long long)n, (long long)p, FW_MAX_RANKS);
    NEED(!ctx->grp.attached, FW_ERR_STATE, "fw_comm_export: context is attached to a group (fw_comm_detach first)");
    CK(cudaSetDevice(ctx->device));
    fwcomm::Group& G = ctx->grp;
    G.rank = rank; G.world = world; G.n = n; G.p = p;
    G.nb = (int)((p + 127) / 128); G.h = (G.nb + 2 * world - 1) / (2 * world);
    const i64 n_pairs = p * (p - 1) / 2;
    G.list_cap = std::max<i64>((i64)1 << 16, std::min<i64>(n_pairs, n_pairs / (16 * (i64)world) + ((i64)1 << 16)));
    const i64 cols = G.col0(rank + 1) - G.col0(rank);
    CK(ctx->g_table.reserve((size_t)std::max<i64>(cols, 1) * n));
    CK(ctx->g_cor.reserve((size_t)G.shard_rows() * p));
    CK(ctx->g_list.reserve((size_t)G.list_cap));
    CK(ctx->g_flags.reserve(fwcomm::N_FLAGS));
    CK(ctx->g_err.reserve(4));
    CK(cudaMemset(ctx->g_flags.ptr, 0, sizeof(u64) * fwcomm::N_FLAGS));
    CK(cudaMemset(ctx->g_err.ptr, 0, sizeof(int) * 4));
    G.own[fwcomm::B_TABLE] = ctx->g_table.ptr; G.own[fwcomm::B_COR] = ctx->g_cor.ptr; G.own[fwcomm::B_LIST] = ctx->g_list.ptr; G.own[fwcomm::B_FLAGS] = ctx->g_flags.ptr;
    G.d_err = ctx->g_err.ptr; G.seq = 0;
    fwcomm::Handle H; memset(&H, 0, sizeof(H));
    H.pid = (int64_t)getpid(); H.device = ctx->device; H.rank = rank; H.world = world; H.n = n; H.p = p; H.list_cap = G.list_cap;
    for (int b = 0; b < 4; ++b) { CK(cudaIpcGetMemHandle(&H.ipc[b], G.own[b])); H.raw[b] = G.own[b]; }
    memcpy(handle_out, &H, sizeof(H));
    G.exported = true;
    return FW_OK;
}

int32_t fw_comm_detach(fw_ctx* ctx) {
    if (!ctx) return FW_ERR_INVALID;
    fwcomm::Group& G = ctx->grp;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (int q = 0; q < FW_MAX_RANKS; ++q) for (int b = 0; b < 4; ++b) {
        if (G.opened[q][b] && G.peer[q][b]) cudaIpcCloseMemHandle(G.peer[q][b]);
        G.opened[q][b] = false; G.peer[q][b] = nullptr;
    }
    G.attached = false;
    if (ctx->cor_sharded) { ctx->cor_sharded = false; ctx->cor_p = 0; }
    ctx->col.valid = false;
    return FW_OK;
}

int32_t fw_comm_attach(fw_ctx* ctx, const void* handles) {
    if (!ctx) return FW_ERR_INVALID;
    fwcomm::Group& G = ctx->grp;
    NEED(handles, FW_ERR_INVALID, "fw_comm_attach: handles is NULL");
    NEED(G.exported && !G.attached, FW_ERR_STATE, "fw_comm_attach: call fw_comm_export first (and fw_comm_detach before re-attaching)");
    CK(cudaSetDevice(ctx->device));
    const int64_t me = (int64_t)getpid();
    for (int q = 0; q < G.world; ++q) {
        fwcomm::Handle H; memcpy(&H, (const char*)handles + (size_t)q * sizeof(H), sizeof(H));
        NEED(H.rank == q && H.world == G.world && H.n == G.n && H.p == G.p && H.list_cap == G.list_cap, FW_ERR_INVALID,
             "fw_comm_attach: handle %d does not describe rank %d of this group (rank %d, world %d, n %lld, p %lld)", q, q, H.rank, H.world, (long long)H.n, (long long)H.p);
        if (q == G.rank) { for (int b = 0; b < 4; ++b) G.peer[q][b] = G.own[b]; continue; }
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, ctx->device, H.device));
        NEED(can, FW_ERR_UNSUPPORTED, "fw_comm_attach: device %d cannot access device %d (no NVLink / P2P path)", ctx->device, H.device);
        if (H.pid == me) {
            cudaError_t e_ = cudaDeviceEnablePeerAccess(H.device, 0);
            if (e_ == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e_ = cudaSuccess; }
            if (e_ != cudaSuccess) return fail(ctx, FW_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", H.device, cudaGetErrorString(e_));
            for (int b = 0; b < 4; ++b) G.peer[q][b] = H.raw[b];
        } else {
            for (int b = 0; b < 4; ++b) {
                cudaError_t e_ = cudaIpcOpenMemHandle(&G.peer[q][b], H.ipc[b], cudaIpcMemLazyEnablePeerAccess);
                if (e_ != cudaSuccess) { fw_comm_detach(ctx); return fail(ctx, FW_ERR_CUDA, "cudaIpcOpenMemHandle (rank %d, buffer %d): %s", q, b, cudaGetErrorString(e_)); }
                G.opened[q][b] = true;
            }
        }
    }
    G.attached = true; G.seq = 0;
    return FW_OK;
}

// upload this rank's columns [p*rank/world, p*(rank+1)/world) of the table; after the call every rank's slice is in place
int32_t fw_multi_set_data_f32(fw_ctx* ctx, const float* host_slice, int64_t ld) {
    if (!ctx) return FW_ERR_INVALID;
    fwcomm::Group& G = ctx->grp;
    NEED(G.attached, FW_ERR_STATE, "fw_multi_set_data_f32: no group attached (fw_comm_export / fw_comm_attach)");
    NEED(host_slice && ld >= G.n, FW_ERR_INVALID, "fw_multi_set_data_f32: bad arguments");
    CK(cudaSetDevice(ctx->device));
    const i64 cols = G.col0(G.rank + 1) - G.col0(G.rank);
    { int st_ = comm_barrier(ctx); if (st_ != FW_OK) return st_; }       // every rank has finished reading the previous table / cor_mat / lists
    if (cols > 0) CK(cudaMemcpy2DAsync(ctx->g_table.ptr, G.n * sizeof(float), host_slice, ld * sizeof(float), G.n * sizeof(float), (size_t)cols, cudaMemcpyHostToDevice, ctx->stream));
    { int st_ = comm_barrier(ctx); if (st_ != FW_OK) return st_; }       // every slice is in place
    ctx->n = G.n; ctx->p = G.p; ctx->ld = G.n; ctx->data_kind = 2; ctx->n_obs = G.n;
    table_changed(ctx);
    CK(cudaStreamSynchronize(ctx->stream));                               // the host buffer is borrowed for the duration of the call only
    return comm_check(ctx);
}

// cor_mat = Float32.(cor(data)) (learning.jl:42-44) computed by the group: this rank's tile rows into its own shard
int32_t fw_multi_cor(fw_ctx* ctx) {
    if (!ctx) return FW_ERR_INVALID;
    fwcomm::Group& G = ctx->grp;
    NEED(G.attached && ctx->data_kind == 2, FW_ERR_STATE, "fw_multi_cor: no group table resident (fw_multi_set_data_f32)");
    CK(cudaSetDevice(ctx->device));
    const i64 n = G.n, p = G.p;
    const i64 kp = (n + cortc::BK - 1) / cortc::BK * cortc::BK, p_pad = (p + cortc::BM - 1) / cortc::BM * cortc::BM;
    CK(ctx->tc.reserve((size_t)2 * p_pad * kp));
    __nv_bfloat16* zhi = ctx->tc.z; __nv_bfloat16* zlo = ctx->tc.z + (size_t)p_pad * kp;
    std::string msg; int nl = 0;
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    {   // fused all-gather + standardise: every column is read from its owner's slice over NVLink
        fwcomm::PeerSlices ps;
        for (int q = 0; q <= FW_MAX_RANKS; ++q) ps.col0[q] = G.col0(q < G.world ? q : G.world);
        for (int q = 0; q < FW_MAX_RANKS; ++q) ps.s[q] = q < G.world ? (const float*)G.peer[q][fwcomm::B_TABLE] : nullptr;
        const int staged = n * (i64)sizeof(float) <= 200 * 1024 ? 1 : 0;
        const size_t smem = staged ? (size_t)n * sizeof(float) : 0;
        CK(cudaFuncSetAttribute(fwcomm::standardize_split_peer_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
        fwcomm::standardize_split_peer_kernel<256><<<(unsigned)p_pad, 256, smem, ctx->stream>>>(ps, G.world, n, p, kp, zhi, zlo, staged, G.col0((G.rank + 1) % G.world));
        ctx->launches++;
        CK(cudaGetLastError());
        CK(cudaEventRecord(ctx->evx[0], ctx->stream));
    }
    cudaError_t e = cortc::encode_map(&ctx->tcp.tm_hi, zhi, kp, p_pad, &msg);
    if (e == cudaSuccess) e = cortc::encode_map(&ctx->tcp.tm_lo, zlo, kp, p_pad, &msg);
    if (e != cudaSuccess) return fail(ctx, FW_ERR_CUDA, "fw_multi_cor: %s", msg.c_str());
    ctx->tcp.kp = kp; ctx->tcp.p_pad = p_pad; ctx->tcp.nb = (int)(p_pad / cortc::BM); ctx->tcp.valid = true;
    PwEmit em; { int st_ = arm_collect(ctx, true, &em); if (st_ != FW_OK) return st_; }
    const int grp_id[2] = {G.rank, 2 * G.world - 1 - G.rank};
    for (int k = 0; k < 2; ++k) {
        e = cortc::run_rows(ctx->tcp, ctx->g_cor.ptr, p, grp_id[k] * G.h, (grp_id[k] + 1) * G.h, false, ctx->stream, &nl, &msg, em, G.world, G.h);
        if (e != cudaSuccess) return fail(ctx, FW_ERR_CUDA, "fw_multi_cor: %s: %s", msg.c_str(), cudaGetErrorString(e));
    }
    ctx->launches += nl;
    CK(cudaEventRecord(ctx->ev[1], ctx->stream)); ctx->ev_valid[0] = true;
    if (em.on) CK(cudaMemcpyAsync(G.flags_of(G.rank) + fwcomm::F_LIST_N, ctx->d_listcnt.ptr, 2 * sizeof(u64), cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaEventRecord(ctx->evx[1], ctx->stream));
    { int st_ = comm_barrier(ctx); if (st_ != FW_OK) return st_; }       // the whole distributed cor_mat (and every rank's candidate list) is complete
    CK(cudaEventRecord(ctx->evx[2], ctx->stream)); ctx->evx_valid = true;
    ctx->cor_p = p; ctx->cor_sharded = true; ctx->col.valid = ctx->col.armed;
    return FW_OK;
}

// ---- single tests -------------------------------------------------------------------------
int32_t fw_test_batch(fw_ctx* ctx, int32_t kind, int64_t n_tests, const int64_t* X, const int64_t* Y,
                      const int32_t* k, const int64_t* Zs, int64_t hps, int64_t n_obs_min, fw_test_result* out) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(kind >= FW_MI && kind <= FW_FZ_NZ, FW_ERR_INVALID, "fw_test_batch: unknown test kind %d", kind);
    NEED(n_tests >= 0 && (n_tests == 0 || (X && Y && k && Zs && out)), FW_ERR_INVALID, "fw_test_batch: NULL argument");
    const bool disc = kind == FW_MI || kind == FW_MI_NZ;
    const bool nzk = kind == FW_FZ_NZ;
    NzTable nzt;
    if (nzk) { int st_ = ensure_nz_table(ctx, &nzt); if (st_ != FW_OK) return st_; }
    if (disc) NEED(ctx->data_kind == 1, FW_ERR_STATE, "fw_test_batch: no discrete table resident (fw_set_data_i32)");
    else if (!nzk) {
        NEED(has_cor(ctx), FW_ERR_STATE, "fw_test_batch: no cor_mat resident (fw_cor_matrix / fw_set_cor_f32)");
        NEED(ctx->n_obs >= 0, FW_ERR_STATE, "fw_test_batch: number of observations unknown (fw_set_data_f32 / fw_set_n_obs)");
    }
    if (n_tests == 0) return FW_OK;
    CK(cudaSetDevice(ctx->device));
    const i64 p = (disc || nzk) ? ctx->p : ctx->cor_p, base = ctx->index_base;
    std::vector<i64> hx(n_tests), hy(n_tests), hz((size_t)n_tests * 3);
    for (i64 t = 0; t < n_tests; ++t) {
        NEED(k[t] >= 0 && k[t] <= 3, FW_ERR_UNSUPPORTED, "fw_test_batch: |Zs| = %d not in 0..3", k[t]);
        hx[t] = X[t] - base; hy[t] = Y[t] - base;
        NEED(hx[t] >= 0 && hx[t] < p && hy[t] >= 0 && hy[t] < p, FW_ERR_INVALID, "fw_test_batch: variable index out of range at test %lld", (long long)t);
        for (int j = 0; j < 3; ++j) {
            i64 z = j < k[t] ? Zs[t * 3 + j] - base : 0;
            NEED(z >= 0 && z < p, FW_ERR_INVALID, "fw_test_batch: conditioning index out of range at test %lld", (long long)t);
            hz[(size_t)t * 3 + j] = z;
        }
    }
    DevBuf<i64> dx, dy, dz; DevBuf<int> dk; DevBuf<DevResult> dout;
    CK(dx.reserve(n_tests)); CK(dy.reserve(n_tests)); CK(dz.reserve((size_t)n_tests * 3)); CK(dk.reserve(n_tests)); CK(dout.reserve(n_tests));
    CK(cudaMemcpyAsync(dx.ptr, hx.data(), sizeof(i64) * n_tests, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dy.ptr, hy.data(), sizeof(i64) * n_tests, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dz.ptr, hz.data(), sizeof(i64) * n_tests * 3, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dk.ptr, k, sizeof(int) * n_tests, cudaMemcpyHostToDevice, ctx->stream));
    if (disc) {
        MiTable t = make_mi_table(ctx, kind);
        const int WARPS = 8;
        size_t smem = (size_t)WARPS * t.L * t.L * t.L * t.L * t.L * sizeof(int);
        i64 blocks = std::min<i64>((n_tests + WARPS - 1) / WARPS, (i64)ctx->sm_count * 8);
        mi_test_batch_kernel<WARPS><<<(unsigned)blocks, WARPS * 32, smem, ctx->stream>>>(t, n_tests, dx.ptr, dy.ptr, dk.ptr, dz.ptr, hps, n_obs_min, dout.ptr);
    } else if (nzk) {
        i64 blocks = std::min<i64>(n_tests, (i64)ctx->sm_count * 8);
        const size_t nz_smem = sizeof(unsigned int) * nzt.W + 16 + FZNZ_GRAM_BYTES;
        NEED(nz_smem <= 200 * 1024, FW_ERR_UNSUPPORTED, "fw_test_batch: %lld rows do not fit the fz_nz kernel's shared memory", (long long)ctx->n);
        CK(cudaFuncSetAttribute(fznz_test_batch_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nz_smem));
        fznz_test_batch_kernel<256><<<(unsigned)blocks, 256, nz_smem, ctx->stream>>>(nzt, n_tests, dx.ptr, dy.ptr, dk.ptr, dz.ptr, n_obs_min, dout.ptr);
    } else {
        FzConsts fc = make_fz_consts(ctx->n_obs, n_obs_min);
        int threads = 128; i64 blocks = (n_tests + threads - 1) / threads;
        fz_test_batch_kernel<<<(unsigned)blocks, threads, 0, ctx->stream>>>(make_cor_view(ctx), p, n_tests, dx.ptr, dy.ptr, dk.ptr, dz.ptr, fc, ctx->n_obs, n_obs_min, dout.ptr);
    }
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, dout.ptr, sizeof(DevResult) * n_tests, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return FW_OK;
}

int32_t fw_test_subsets_batch(fw_ctx* ctx, int32_t kind, int64_t n_jobs, const int64_t* X, const int64_t* Y,
                              const int64_t* z_off, const int64_t* z_idx,
                              int32_t max_k, double alpha, int64_t hps, int64_t n_obs_min, int64_t max_tests,
                              fw_test_result* out_result, int64_t* out_Zs, int32_t* out_k, int64_t* num_tests, double* frac) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(kind >= FW_MI && kind <= FW_FZ_NZ, FW_ERR_INVALID, "fw_test_subsets: unknown test kind %d", kind);
    NEED(max_k >= 1 && max_k <= 3, FW_ERR_UNSUPPORTED, "fw_test_subsets: max_k = %d not in 1..3", max_k);
    NEED(n_jobs >= 0 && (n_jobs == 0 || (X && Y && z_off && out_result && out_Zs && out_k && num_tests && frac)), FW_ERR_INVALID, "fw_test_subsets: NULL argument");
    const bool disc = kind == FW_MI || kind == FW_MI_NZ;
    const bool nzk = kind == FW_FZ_NZ;
    NzTable nzt;
    if (nzk) { int st_ = ensure_nz_table(ctx, &nzt); if (st_ != FW_OK) return st_; }
    if (disc) NEED(ctx->data_kind == 1, FW_ERR_STATE, "fw_test_subsets: no discrete table resident (fw_set_data_i32)");
    else if (!nzk) {
        NEED(has_cor(ctx), FW_ERR_STATE, "fw_test_subsets: no cor_mat resident");
        NEED(ctx->n_obs >= 0, FW_ERR_STATE, "fw_test_subsets: number of observations unknown");
    }
    if (n_jobs == 0) return FW_OK;
    CK(cudaSetDevice(ctx->device));
    const i64 p = (disc || nzk) ? ctx->p : ctx->cor_p, base = ctx->index_base;
    const i64 nz = z_off[n_jobs];
    NEED(nz == 0 || z_idx, FW_ERR_INVALID, "fw_test_subsets: z_idx is NULL");
    std::vector<i64> hx(n_jobs), hy(n_jobs), hz((size_t)std::max<i64>(nz, 1));
    for (i64 j = 0; j < n_jobs; ++j) {
        hx[j] = X[j] - base; hy[j] = Y[j] - base;
        NEED(hx[j] >= 0 && hx[j] < p && hy[j] >= 0 && hy[j] < p, FW_ERR_INVALID, "fw_test_subsets: variable index out of range in job %lld", (long long)j);
        NEED(z_off[j + 1] >= z_off[j], FW_ERR_INVALID, "fw_test_subsets: z_off not monotone");
    }
    for (i64 i = 0; i < nz; ++i) { hz[i] = z_idx[i] - base; NEED(hz[i] >= 0 && hz[i] < p, FW_ERR_INVALID, "fw_test_subsets: conditioning index out of range"); }

    // jobs with empty Z_total: the reference's sentinel (src/tests.jl:285); others by capacity class
    std::vector<int> cls[5];
    int max_need = 0;
    for (i64 j = 0; j < n_jobs; ++j) {
        i64 m = z_off[j + 1] - z_off[j];
        if (m == 0) {
            fw_test_result r; memset(&r, 0, sizeof(r)); r.stat = NAN; r.pval = NAN; r.df = -1; r.suff_power = 1;
            out_result[j] = r; out_Zs[j * 3] = -1; out_Zs[j * 3 + 1] = -1; out_Zs[j * 3 + 2] = -1; out_k[j] = 1; num_tests[j] = -1; frac[j] = NAN;
            continue;
        }
        NEED(m + 2 < (i64)1 << 20, FW_ERR_UNSUPPORTED, "fw_test_subsets: |Z_total| too large");
        int need = (int)m + 2, c = 4;
        for (int q = 0; q < 4; ++q) if (need <= kCaps[q]) { c = q; break; }
        cls[c].push_back((int)j);
        if (c == 4) max_need = std::max(max_need, need);
    }
    DevBuf<i64> dx, dy, dzo, dzi, dZs, dnt; DevBuf<int> dsel, dk; DevBuf<DevResult> dres; DevBuf<double> dfr; DevBuf<float> gs;
    CK(dx.reserve(n_jobs)); CK(dy.reserve(n_jobs)); CK(dzo.reserve(n_jobs + 1)); CK(dzi.reserve(nz)); CK(dZs.reserve((size_t)n_jobs * 3));
    CK(dnt.reserve(n_jobs)); CK(dsel.reserve(n_jobs)); CK(dk.reserve(n_jobs)); CK(dres.reserve(n_jobs)); CK(dfr.reserve(n_jobs));
    CK(cudaMemcpyAsync(dx.ptr, hx.data(), sizeof(i64) * n_jobs, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dy.ptr, hy.data(), sizeof(i64) * n_jobs, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dzo.ptr, z_off, sizeof(i64) * (n_jobs + 1), cudaMemcpyHostToDevice, ctx->stream));
    if (nz) CK(cudaMemcpyAsync(dzi.ptr, hz.data(), sizeof(i64) * nz, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_exec.ptr, 0, 4 * sizeof(u64), ctx->stream));
    if (disc) {
        SubsetsMiArgs ma;
        ma.t = make_mi_table(ctx, kind); ma.hps = hps;
        ma.X = dx.ptr; ma.Y = dy.ptr; ma.z_off = dzo.ptr; ma.z_idx = dzi.ptr; ma.n_jobs = (int)n_jobs; ma.counter = ctx->d_counter.ptr;
        ma.max_k = max_k; ma.alpha = alpha; ma.max_tests = max_tests;
        int need = 2; for (int c = 0; c < 5; ++c) for (int j : cls[c]) need = std::max<int>(need, (int)(z_off[j + 1] - z_off[j]) + 2);
        ma.cap = need;
        ma.out = dres.ptr; ma.out_Zs = dZs.ptr; ma.out_k = dk.ptr; ma.num_tests = dnt.ptr; ma.frac = dfr.ptr; ma.executed_total = ctx->d_exec.ptr;
        const int TH = 256; const int L = ma.t.L;
        size_t smem = ((sizeof(i64) * (need + 1) + sizeof(i64) * need + sizeof(int) * need + 15) & ~(size_t)15) + (size_t)(TH / 32) * L * L * L * L * L * sizeof(int)
                      + (size_t)(TH / 32) * MI_BIN_WARP_BYTES + 16;      // + the count buffers of the batched binary scan
        NEED(smem <= 200 * 1024, FW_ERR_UNSUPPORTED, "fw_test_subsets: |Z_total| = %d exceeds the supported maximum", need - 2);
        CK(cudaMemsetAsync(ctx->d_counter.ptr, 0, sizeof(int), ctx->stream));
        int grid = 1;
        CK(grid_for(subsets_mi_kernel<256, 1>, TH, smem, ctx->sm_count, n_jobs, &grid));
        subsets_mi_kernel<256, 1><<<grid, TH, smem, ctx->stream>>>(ma);
        ctx->launches++;
        CK(cudaGetLastError());
        for (int c = 0; c < 5; ++c) cls[c].clear();
    }
    SubsetsArgs a;
    a.cv = make_cor_view(ctx); a.p = p; a.X = dx.ptr; a.Y = dy.ptr; a.z_off = dzo.ptr; a.z_idx = dzi.ptr;
    a.max_k = max_k; a.alpha = alpha; a.max_tests = max_tests; a.fc = make_fz_consts(ctx->n_obs, n_obs_min);
    a.out = dres.ptr; a.out_Zs = dZs.ptr; a.out_k = dk.ptr; a.num_tests = dnt.ptr; a.frac = dfr.ptr; a.executed_total = ctx->d_exec.ptr;
    a.counter = ctx->d_counter.ptr;
    if (nzk) { a.nzt = nzt; a.n_obs_min = n_obs_min; }
    const int nzw = nzk ? nzt.W : -1;
    size_t sel_off = 0;
    for (int c = 0; c < 5; ++c) {
        if (cls[c].empty()) continue;
        int n_sel = (int)cls[c].size();
        CK(cudaMemcpyAsync(dsel.ptr + sel_off, cls[c].data(), sizeof(int) * n_sel, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemsetAsync(ctx->d_counter.ptr, 0, sizeof(int), ctx->stream));
        a.sel = dsel.ptr + sel_off; a.n_sel = n_sel; sel_off += n_sel;
        int grid = 1;
        if (c < 4) {
            a.cap = kCaps[c]; a.gscratch = nullptr;
            size_t smem = subsets_smem_bytes(a.cap, true, nzw);
            NEED(smem <= 220 * 1024, FW_ERR_UNSUPPORTED, "fw_test_subsets: job does not fit shared memory (n = %lld rows, %d slots)", (long long)ctx->n, a.cap);
            if (nzk) { CK(grid_for(subsets_fz_kernel<256, 2, true, false>, 256, smem, ctx->sm_count, n_sel, &grid)); subsets_fz_kernel<256, 2, true, false><<<grid, 256, smem, ctx->stream>>>(a); }
            else if (c < 2) { CK(grid_for(subsets_fz_kernel<128, 2, false, false>, 128, smem, ctx->sm_count, n_sel, &grid)); subsets_fz_kernel<128, 2, false, false><<<grid, 128, smem, ctx->stream>>>(a); }
            else { CK(grid_for(subsets_fz_kernel<256, 2, false, false>, 256, smem, ctx->sm_count, n_sel, &grid)); subsets_fz_kernel<256, 2, false, false><<<grid, 256, smem, ctx->stream>>>(a); }
        } else {
            a.cap = max_need;
            size_t smem = subsets_smem_bytes(a.cap, false, nzw);
            NEED(smem <= 200 * 1024, FW_ERR_UNSUPPORTED, "fw_test_subsets: |Z_total| = %d exceeds the supported maximum", max_need - 2);
            if (nzk) CK(grid_for(subsets_fz_kernel<256, 2, true, true>, 256, smem, ctx->sm_count, n_sel, &grid));
            else CK(grid_for(subsets_fz_kernel<256, 2, false, true>, 256, smem, ctx->sm_count, n_sel, &grid));
            size_t per = (size_t)a.cap * a.cap;
            while (grid > 1 && per * grid * sizeof(float) > ((size_t)4 << 30)) grid = (grid + 1) / 2;
            CK(gs.reserve(per * grid));
            a.gscratch = gs.ptr;
            if (nzk) subsets_fz_kernel<256, 2, true, true><<<grid, 256, smem, ctx->stream>>>(a);
            else subsets_fz_kernel<256, 2, false, true><<<grid, 256, smem, ctx->stream>>>(a);
        }
        ctx->launches++;
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<DevResult> hres(n_jobs); std::vector<i64> hZs((size_t)n_jobs * 3), hnt(n_jobs); std::vector<int> hk(n_jobs); std::vector<double> hfr(n_jobs);
    CK(cudaMemcpy(hres.data(), dres.ptr, sizeof(DevResult) * n_jobs, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hZs.data(), dZs.ptr, sizeof(i64) * n_jobs * 3, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hnt.data(), dnt.ptr, sizeof(i64) * n_jobs, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hk.data(), dk.ptr, sizeof(int) * n_jobs, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hfr.data(), dfr.ptr, sizeof(double) * n_jobs, cudaMemcpyDeviceToHost));
    for (i64 j = 0; j < n_jobs; ++j) {
        if (z_off[j + 1] == z_off[j]) continue;
        memcpy(&out_result[j], &hres[j], sizeof(fw_test_result));
        for (int i = 0; i < 3; ++i) out_Zs[j * 3 + i] = hZs[(size_t)j * 3 + i] >= 0 ? hZs[(size_t)j * 3 + i] + base : -1;
        out_k[j] = hk[j]; num_tests[j] = hnt[j]; frac[j] = hfr[j];
    }
    return FW_OK;
}

int32_t fw_test_subsets(fw_ctx* ctx, int32_t kind, int64_t X, int64_t Y, const int64_t* Z_total, int64_t m,
                        int32_t max_k, double alpha, int64_t hps, int64_t n_obs_min, int64_t max_tests,
                        fw_test_result* out_result, int64_t* out_Zs, int32_t* out_k, int64_t* num_tests, double* frac) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(m >= 0, FW_ERR_INVALID, "fw_test_subsets: m < 0");
    int64_t z_off[2] = {0, m};
    return fw_test_subsets_batch(ctx, kind, 1, &X, &Y, z_off, Z_total, max_k, alpha, hps, n_obs_min, max_tests, out_result, out_Zs, out_k, num_tests, frac);
}

// ---- pairwise stage -------------------------------------------------------------------------
// the resident neighbour lists <- the CSR a pairwise stage produced (device copies + host offsets)
static int adopt_pairwise(fw_ctx* ctx, const PairwiseOut& po, i64 p) {
    CK(ctx->d_uni_off.reserve(p + 1)); CK(ctx->d_uni_nbr.reserve(po.n_entries)); CK(ctx->d_uni_stat.reserve(po.n_entries)); CK(ctx->d_uni_p.reserve(po.n_entries));
    CK(cudaMemcpyAsync(ctx->d_uni_off.ptr, po.d_off, sizeof(i64) * (p + 1), cudaMemcpyDeviceToDevice, ctx->stream));
    if (po.n_entries) {
        CK(cudaMemcpyAsync(ctx->d_uni_nbr.ptr, po.d_nbr, sizeof(i64) * po.n_entries, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_uni_stat.ptr, po.d_stat, sizeof(double) * po.n_entries, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_uni_p.ptr, po.d_adjp, sizeof(double) * po.n_entries, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    ctx->h_uni_off.resize(p + 1);
    CK(cudaMemcpyAsync(ctx->h_uni_off.data(), po.d_off, sizeof(i64) * (p + 1), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->uni_entries = po.n_entries; ctx->pw_tests = po.n_tests; ctx->pw_reliable = po.n_reliable; ctx->pw_raw_sig = po.n_raw_sig;
    return FW_OK;
}

int32_t fw_pairwise(fw_ctx* ctx, int32_t kind, double alpha, int64_t hps, int64_t n_obs_min,
                    int32_t fdr, int32_t correct_reliable_only, int64_t* n_entries) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(kind >= FW_MI && kind <= FW_FZ_NZ, FW_ERR_INVALID, "fw_pairwise: unknown test kind %d", kind);
    const bool disc = kind == FW_MI || kind == FW_MI_NZ;
    const bool nzk = kind == FW_FZ_NZ;
    NzTable nzt;
    if (nzk) { int st_ = ensure_nz_table(ctx, &nzt); if (st_ != FW_OK) return st_; }
    if (disc) NEED(ctx->data_kind == 1, FW_ERR_STATE, "fw_pairwise: no discrete table resident (fw_set_data_i32)");
    else if (!nzk) {
        NEED(has_cor(ctx), FW_ERR_STATE, "fw_pairwise: no cor_mat resident (fw_cor_matrix / fw_set_cor_f32)");
        NEED(ctx->n_obs >= 0, FW_ERR_STATE, "fw_pairwise: number of observations unknown");
    }
    CK(cudaSetDevice(ctx->device));
    const i64 p = (disc || nzk) ? ctx->p : ctx->cor_p;
    PairwiseOut po;
    std::string msg; int nl = 0;
    CK(cudaEventRecord(ctx->ev[2], ctx->stream));
    cudaError_t e;
    if (disc) {
        MiTable t = make_mi_table(ctx, kind);
        e = pairwise_mi_run(ctx->pw, t, hps, n_obs_min, alpha, fdr != 0, correct_reliable_only != 0, ctx->stream, &po, &nl, &msg);
    } else if (nzk) {
        e = pairwise_fznz_run(ctx->pw, ctx->nzplanes, nzt, n_obs_min, alpha, fdr != 0, correct_reliable_only != 0, ctx->stream, &po, &nl, &msg);
    } else {
        int st_ = run_pairwise_fz(ctx, alpha, n_obs_min, fdr != 0, correct_reliable_only != 0, &po, &nl);
        ctx->launches += nl; nl = 0;
        if (st_ != FW_OK) return st_;
        e = cudaSuccess;
    }
    ctx->launches += nl;
    if (e != cudaSuccess && msg.find("unsupported size") != std::string::npos) return fail(ctx, FW_ERR_UNSUPPORTED, "fw_pairwise: %s", msg.c_str());
    if (e != cudaSuccess) return fail(ctx, FW_ERR_CUDA, "fw_pairwise: %s: %s", msg.c_str(), cudaGetErrorString(e));
    CK(cudaEventRecord(ctx->ev[3], ctx->stream)); ctx->ev_valid[1] = true;
    { int st_ = adopt_pairwise(ctx, po, p); if (st_ != FW_OK) return st_; }
    if (n_entries) *n_entries = po.n_entries;
    return FW_OK;
}

// ---- pairwise stage of the table-based kinds split over the ranks of a job (include/fwgpu.h "multi-GPU, table-based kinds") -------
int32_t fw_pairwise_partial(fw_ctx* ctx, int32_t kind, double alpha, int64_t hps, int64_t n_obs_min, int32_t correct_reliable_only,
                            int32_t rank, int32_t world, int64_t* n_raw, int64_t* n_reliable) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(kind == FW_MI || kind == FW_MI_NZ || kind == FW_FZ_NZ, FW_ERR_UNSUPPORTED, "fw_pairwise_partial: kind %d (FW_FZ shares the work through the group path, fw_multi_cor + fw_pairwise)", kind);
    NEED(world >= 1 && world <= FW_MAX_RANKS && rank >= 0 && rank < world, FW_ERR_INVALID, "fw_pairwise_partial: rank %d of %d", rank, world);
    const bool disc = kind != FW_FZ_NZ;
    NzTable nzt;
    if (!disc) { int st_ = ensure_nz_table(ctx, &nzt); if (st_ != FW_OK) return st_; }
    else NEED(ctx->data_kind == 1, FW_ERR_STATE, "fw_pairwise_partial: no discrete table resident (fw_set_data_i32)");
    CK(cudaSetDevice(ctx->device));
    std::string msg; int nl = 0;
    CK(cudaEventRecord(ctx->ev[2], ctx->stream));
    cudaError_t e;
    ctx->part = PwCollected(); ctx->part_kind = -1;
    if (disc) { MiTable t = make_mi_table(ctx, kind); e = pairwise_mi_collect(ctx->pw, t, hps, n_obs_min, alpha, correct_reliable_only != 0, rank, world, ctx->stream, &ctx->part, &nl, &msg); }
    else e = pairwise_fznz_collect(ctx->pw, ctx->nzplanes, nzt, n_obs_min, alpha, correct_reliable_only != 0, rank, world, ctx->stream, &ctx->part, &nl, &msg);
    ctx->launches += nl;
    if (e != cudaSuccess) return fail(ctx, FW_ERR_CUDA, "fw_pairwise_partial: %s: %s", msg.c_str(), cudaGetErrorString(e));
    CK(cudaEventRecord(ctx->ev[3], ctx->stream)); ctx->ev_valid[1] = true;
    ctx->part_kind = kind;
    if (n_raw) *n_raw = ctx->part.nf;
    if (n_reliable) *n_reliable = ctx->part.n_rel;
    return FW_OK;
}

int32_t fw_pairwise_partial_copy(fw_ctx* ctx, int32_t* x, int32_t* y, double* pval, double* stat) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(ctx->part_kind >= 0, FW_ERR_STATE, "fw_pairwise_partial_copy: no partial pairwise result (fw_pairwise_partial)");
    NEED(ctx->part.nf == 0 || (x && y && pval && stat), FW_ERR_INVALID, "fw_pairwise_partial_copy: NULL output");
    CK(cudaSetDevice(ctx->device));
    const i64 nf = ctx->part.nf;
    if (nf) {
        CK(cudaMemcpyAsync(x, ctx->part.c_x, sizeof(int) * nf, cudaMemcpyDefault, ctx->stream));
        CK(cudaMemcpyAsync(y, ctx->part.c_y, sizeof(int) * nf, cudaMemcpyDefault, ctx->stream));
        CK(cudaMemcpyAsync(pval, ctx->part.c_p, sizeof(double) * nf, cudaMemcpyDefault, ctx->stream));
        CK(cudaMemcpyAsync(stat, ctx->part.c_stat, sizeof(double) * nf, cudaMemcpyDefault, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return FW_OK;
}

int32_t fw_pairwise_merge(fw_ctx* ctx, int32_t kind, double alpha, int32_t fdr, int64_t n_raw_total, const int32_t* x, const int32_t* y,
                          const double* pval, const double* stat, int64_t m_tests, int64_t* n_entries) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(kind == FW_MI || kind == FW_MI_NZ || kind == FW_FZ_NZ, FW_ERR_UNSUPPORTED, "fw_pairwise_merge: kind %d", kind);
    NEED(ctx->data_kind == (kind == FW_FZ_NZ ? 0 : 1) && ctx->p > 0, FW_ERR_STATE, "fw_pairwise_merge: no table of this kind resident");
    NEED(n_raw_total >= 0 && (n_raw_total == 0 || (x && y && pval && stat)), FW_ERR_INVALID, "fw_pairwise_merge: NULL records");
    NEED(n_raw_total < ((i64)1 << 31) - 1, FW_ERR_UNSUPPORTED, "fw_pairwise_merge: more than 2^31 raw-significant pairs");
    CK(cudaSetDevice(ctx->device));
    const i64 p = ctx->p, nf = n_raw_total;
    bool on_device = false;                                           // records may live in host or device memory (unified addressing)
    if (nf) { cudaPointerAttributes at; if (cudaPointerGetAttributes(&at, x) == cudaSuccess) on_device = at.type == cudaMemoryTypeDevice; else cudaGetLastError(); }
    if (!on_device) for (i64 i = 0; i < nf; ++i) NEED(x[i] >= 0 && x[i] < y[i] && y[i] < p, FW_ERR_INVALID, "fw_pairwise_merge: record %lld is not a pair x < y < p", (long long)i);
    int *d_x, *d_y; double *d_p, *d_s;
    const i64 cap = std::max<i64>(nf, 16);
    CK(ctx->pw.get(7, sizeof(int) * cap, (void**)&d_x)); CK(ctx->pw.get(8, sizeof(int) * cap, (void**)&d_y));
    CK(ctx->pw.get(9, sizeof(double) * cap, (void**)&d_p)); CK(ctx->pw.get(10, sizeof(double) * cap, (void**)&d_s));
    ctx->part = PwCollected(); ctx->part_kind = -1;                   // the slots of the partial result are overwritten
    if (nf) {
        CK(cudaMemcpyAsync(d_x, x, sizeof(int) * nf, cudaMemcpyDefault, ctx->stream));
        CK(cudaMemcpyAsync(d_y, y, sizeof(int) * nf, cudaMemcpyDefault, ctx->stream));
        CK(cudaMemcpyAsync(d_p, pval, sizeof(double) * nf, cudaMemcpyDefault, ctx->stream));
        CK(cudaMemcpyAsync(d_s, stat, sizeof(double) * nf, cudaMemcpyDefault, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));                       // borrowed pointers: valid for the duration of the call only
    }
    PairwiseOut po; std::string msg; int nl = 0;
    po.n_tests = p * (p - 1) / 2; po.n_raw_sig = nf; po.n_reliable = m_tests;
    cudaError_t e = pairwise_order_finish(ctx->pw, d_x, d_y, d_p, d_s, nf, m_tests, p, alpha, fdr != 0, ctx->stream, &po, &nl, &msg);
    ctx->launches += nl;
    if (e != cudaSuccess && msg.find("unsupported size") != std::string::npos) return fail(ctx, FW_ERR_UNSUPPORTED, "fw_pairwise_merge: %s", msg.c_str());
    if (e != cudaSuccess) return fail(ctx, FW_ERR_CUDA, "fw_pairwise_merge: %s: %s", msg.c_str(), cudaGetErrorString(e));
    { int st_ = adopt_pairwise(ctx, po, p); if (st_ != FW_OK) return st_; }
    if (n_entries) *n_entries = po.n_entries;
    return FW_OK;
}

int32_t fw_pairwise_copy(fw_ctx* ctx, int64_t* offsets, int64_t* nbr, double* stat, double* adjp) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(ctx->uni_entries >= 0, FW_ERR_STATE, "fw_pairwise_copy: no neighbour lists resident");
    CK(cudaSetDevice(ctx->device));
    const i64 p = (i64)ctx->h_uni_off.size() - 1, ne = ctx->uni_entries, base = ctx->index_base;
    if (offsets) memcpy(offsets, ctx->h_uni_off.data(), sizeof(i64) * (p + 1));
    if (ne) {
        if (nbr) { CK(cudaMemcpyAsync(nbr, ctx->d_uni_nbr.ptr, sizeof(i64) * ne, cudaMemcpyDeviceToHost, ctx->stream)); }
        if (stat) CK(cudaMemcpyAsync(stat, ctx->d_uni_stat.ptr, sizeof(double) * ne, cudaMemcpyDeviceToHost, ctx->stream));
        if (adjp) CK(cudaMemcpyAsync(adjp, ctx->d_uni_p.ptr, sizeof(double) * ne, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    if (nbr && base) for (i64 i = 0; i < ne; ++i) nbr[i] += base;
    return FW_OK;
}

int32_t fw_set_univar_nbrs(fw_ctx* ctx, const int64_t* offsets, const int64_t* nbr, const double* stat, const double* adjp) {
    if (!ctx) return FW_ERR_INVALID;
    i64 p = ctx->p > 0 ? ctx->p : ctx->cor_p;
    NEED(p > 0, FW_ERR_STATE, "fw_set_univar_nbrs: number of variables unknown (install data or cor_mat first)");
    NEED(offsets, FW_ERR_INVALID, "fw_set_univar_nbrs: offsets is NULL");
    CK(cudaSetDevice(ctx->device));
    i64 ne = offsets[p];
    NEED(offsets[0] == 0 && ne >= 0 && (ne == 0 || (nbr && stat && adjp)), FW_ERR_INVALID, "fw_set_univar_nbrs: bad CSR");
    std::vector<i64> hn((size_t)std::max<i64>(ne, 1));
    for (i64 i = 0; i < ne; ++i) { hn[i] = nbr[i] - ctx->index_base; NEED(hn[i] >= 0 && hn[i] < p, FW_ERR_INVALID, "fw_set_univar_nbrs: neighbour index out of range"); }
    for (i64 v = 0; v < p; ++v) NEED(offsets[v + 1] >= offsets[v], FW_ERR_INVALID, "fw_set_univar_nbrs: offsets not monotone");
    CK(ctx->d_uni_off.reserve(p + 1)); CK(ctx->d_uni_nbr.reserve(ne)); CK(ctx->d_uni_stat.reserve(ne)); CK(ctx->d_uni_p.reserve(ne));
    CK(cudaMemcpyAsync(ctx->d_uni_off.ptr, offsets, sizeof(i64) * (p + 1), cudaMemcpyHostToDevice, ctx->stream));
    if (ne) {
        CK(cudaMemcpyAsync(ctx->d_uni_nbr.ptr, hn.data(), sizeof(i64) * ne, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_uni_stat.ptr, stat, sizeof(double) * ne, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_uni_p.ptr, adjp, sizeof(double) * ne, cudaMemcpyHostToDevice, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->h_uni_off.assign(offsets, offsets + p + 1);
    ctx->uni_entries = ne;
    return FW_OK;
}

int32_t fw_pairwise_stats(fw_ctx* ctx, int64_t* n_tests, int64_t* n_reliable, int64_t* n_raw_sig) {
    if (!ctx) return FW_ERR_INVALID;
    if (n_tests) *n_tests = ctx->pw_tests;
    if (n_reliable) *n_reliable = ctx->pw_reliable;
    if (n_raw_sig) *n_raw_sig = ctx->pw_raw_sig;
    return FW_OK;
}

// ---- HITON-PC -------------------------------------------------------------------------------
int32_t fw_hiton_exec_by_k(fw_ctx* ctx, int64_t* out3) {
    if (!ctx || !out3) return FW_ERR_INVALID;
    for (int i = 0; i < 3; ++i) out3[i] = ctx->exec_by_k[i];
    return FW_OK;
}

int32_t fw_hiton_pc_capacity(fw_ctx* ctx, int64_t n_targets, const int64_t* targets, int64_t* capacity) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(ctx->uni_entries >= 0, FW_ERR_STATE, "fw_hiton_pc: no neighbour lists resident (fw_pairwise / fw_set_univar_nbrs)");
    NEED(capacity && (n_targets == 0 || targets), FW_ERR_INVALID, "fw_hiton_pc_capacity: NULL argument");
    const i64 p = (i64)ctx->h_uni_off.size() - 1;
    i64 tot = 0;
    for (i64 t = 0; t < n_targets; ++t) {
        i64 T = targets[t] - ctx->index_base;
        NEED(T >= 0 && T < p, FW_ERR_INVALID, "fw_hiton_pc: target index out of range");
        tot += ctx->h_uni_off[T + 1] - ctx->h_uni_off[T];
    }
    *capacity = tot;
    return FW_OK;
}

int32_t fw_hiton_pc(fw_ctx* ctx, int32_t kind, int64_t n_targets, const int64_t* targets,
                    int32_t max_k, double alpha, int64_t hps, int64_t n_obs_min, int64_t max_tests,
                    int64_t* pc_off, int64_t* pc_count, int64_t* pc_nbr, double* pc_stat, double* pc_p,
                    int64_t* tpc_count, int64_t* tpc_nbr, double* tpc_stat, double* tpc_p,
                    int64_t* num_tests, int64_t* tests_executed_total) {
    return fw_hiton_pc_ex(ctx, kind, n_targets, targets, max_k, alpha, hps, n_obs_min, max_tests, nullptr, nullptr, nullptr, nullptr,
                          pc_off, pc_count, pc_nbr, pc_stat, pc_p, tpc_count, tpc_nbr, tpc_stat, tpc_p, num_tests, tests_executed_total,
                          nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
}

int32_t fw_hiton_pc_ex(fw_ctx* ctx, int32_t kind, int64_t n_targets, const int64_t* targets,
                       int32_t max_k, double alpha, int64_t hps, int64_t n_obs_min, int64_t max_tests,
                       const int64_t* wl_off, const int64_t* wl_idx, const int64_t* bl_off, const int64_t* bl_idx,
                       int64_t* pc_off, int64_t* pc_count, int64_t* pc_nbr, double* pc_stat, double* pc_p,
                       int64_t* tpc_count, int64_t* tpc_nbr, double* tpc_stat, double* tpc_p,
                       int64_t* num_tests, int64_t* tests_executed_total,
                       int64_t* rej_count, int64_t* rej_nbr, int64_t* rej_Zs, int32_t* rej_k, fw_test_result* rej_result,
                       int64_t* rej_num_tests, double* rej_frac) {
    if (!ctx) return FW_ERR_INVALID;
    NEED(kind >= FW_MI && kind <= FW_FZ_NZ, FW_ERR_INVALID, "fw_hiton_pc: unknown test kind %d", kind);
    NEED(max_k >= 0 && max_k <= 3, FW_ERR_UNSUPPORTED, "fw_hiton_pc: max_k = %d not in 0..3", max_k);
    const bool disc = kind == FW_MI || kind == FW_MI_NZ;
    const bool nzk = kind == FW_FZ_NZ;
    NzTable nzt;
    if (nzk) { int st_ = ensure_nz_table(ctx, &nzt); if (st_ != FW_OK) return st_; }
    if (disc) NEED(ctx->data_kind == 1, FW_ERR_STATE, "fw_hiton_pc: no discrete table resident (fw_set_data_i32)");
    else if (!nzk) {
        NEED(has_cor(ctx), FW_ERR_STATE, "fw_hiton_pc: no cor_mat resident");
        NEED(ctx->n_obs >= 0, FW_ERR_STATE, "fw_hiton_pc: number of observations unknown");
    }
    NEED(ctx->uni_entries >= 0, FW_ERR_STATE, "fw_hiton_pc: no neighbour lists resident (fw_pairwise / fw_set_univar_nbrs)");
    NEED(n_targets >= 0 && (n_targets == 0 || targets), FW_ERR_INVALID, "fw_hiton_pc: NULL targets");
    if (n_targets == 0) { if (pc_off) pc_off[0] = 0; if (tests_executed_total) *tests_executed_total = 0; return FW_OK; }
    CK(cudaSetDevice(ctx->device));
    const i64 p = (disc || nzk) ? ctx->p : ctx->cor_p, base = ctx->index_base;
    NEED((i64)ctx->h_uni_off.size() == p + 1, FW_ERR_STATE, "fw_hiton_pc: neighbour lists and cor_mat disagree on the number of variables");

    std::vector<i64> ht(n_targets), hoff(n_targets + 1);
    hoff[0] = 0;
    for (i64 t = 0; t < n_targets; ++t) {
        ht[t] = targets[t] - base;
        NEED(ht[t] >= 0 && ht[t] < p, FW_ERR_INVALID, "fw_hiton_pc: target index out of range");
        hoff[t + 1] = hoff[t] + (ctx->h_uni_off[ht[t] + 1] - ctx->h_uni_off[ht[t]]);
    }
    const i64 cap_total = std::max<i64>(hoff[n_targets], 1);

    fw_ctx::HitonBufs& B = ctx->hb;
    DevBuf<i64>&dt = B.dt, &doff = B.doff, &dpcn = B.dpcn, &dtpcn = B.dtpcn, &dpcc = B.dpcc, &dtpcc = B.dtpcc, &dnt = B.dnt;
    DevBuf<double>&dpcs = B.dpcs, &dpcp = B.dpcp, &dtpcs = B.dtpcs, &dtpcp = B.dtpcp;
    DevBuf<int>&dsel = B.dsel, &dorder = B.dorder, &dstatus = B.dstatus; DevBuf<float>& gs = B.gs;
    CK(dt.reserve(n_targets)); CK(doff.reserve(n_targets + 1)); CK(dpcn.reserve(cap_total)); CK(dtpcn.reserve(cap_total));
    CK(dpcc.reserve(n_targets)); CK(dtpcc.reserve(n_targets)); CK(dnt.reserve(n_targets));
    CK(dpcs.reserve(cap_total)); CK(dpcp.reserve(cap_total)); CK(dtpcs.reserve(cap_total)); CK(dtpcp.reserve(cap_total));
    CK(dsel.reserve(n_targets)); CK(dorder.reserve(cap_total)); CK(dstatus.reserve(n_targets));
    CK(cudaMemcpyAsync(dt.ptr, ht.data(), sizeof(i64) * n_targets, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(doff.ptr, hoff.data(), sizeof(i64) * (n_targets + 1), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_exec.ptr, 0, 4 * sizeof(u64), ctx->stream));
    CK(cudaMemsetAsync(dpcc.ptr, 0, sizeof(i64) * n_targets, ctx->stream));
    CK(cudaMemsetAsync(dtpcc.ptr, 0, sizeof(i64) * n_targets, ctx->stream));
    CK(cudaMemsetAsync(dnt.ptr, 0, sizeof(i64) * n_targets, ctx->stream));

    std::vector<i64> h_pcc(n_targets, 0), h_tpcc(n_targets, 0), h_nt(n_targets, 0);
    if (max_k == 0) {
        // hiton.jl:394-397: PC = univariate neighbours, no conditioning (handled on the host side of the ABI:
        // it is a copy of the resident lists, there is nothing to compute)
        std::vector<i64> un((size_t)std::max<i64>(ctx->uni_entries, 1)); std::vector<double> us(un.size()), up(un.size());
        if (ctx->uni_entries) {
            CK(cudaMemcpy(un.data(), ctx->d_uni_nbr.ptr, sizeof(i64) * ctx->uni_entries, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(us.data(), ctx->d_uni_stat.ptr, sizeof(double) * ctx->uni_entries, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(up.data(), ctx->d_uni_p.ptr, sizeof(double) * ctx->uni_entries, cudaMemcpyDeviceToHost));
        }
        for (i64 t = 0; t < n_targets; ++t) {
            i64 e0 = ctx->h_uni_off[ht[t]], cnt = ctx->h_uni_off[ht[t] + 1] - e0;
            if (pc_count) pc_count[t] = cnt;
            if (tpc_count) tpc_count[t] = 0;
            if (num_tests) num_tests[t] = 0;
            if (rej_count) rej_count[t] = 0;
            for (i64 i = 0; i < cnt; ++i) {
                if (pc_nbr) pc_nbr[hoff[t] + i] = un[e0 + i] + base;
                if (pc_stat) pc_stat[hoff[t] + i] = us[e0 + i];
                if (pc_p) pc_p[hoff[t] + i] = up[e0 + i];
            }
        }
        if (pc_off) memcpy(pc_off, hoff.data(), sizeof(i64) * (n_targets + 1));
        if (tests_executed_total) *tests_executed_total = 0;
        return FW_OK;
    }

    // whitelists / blacklists (CSR over the target list, src/hiton.jl:20-38) and rejection records (src/hiton.jl:72-74)
    const bool track = rej_count != nullptr;
    NEED(!track || (rej_nbr && rej_Zs && rej_k && rej_result && rej_num_tests && rej_frac), FW_ERR_INVALID, "fw_hiton_pc_ex: rejection tracking needs all rej_* arrays");
    HitonLists lists; memset(&lists, 0, sizeof(lists));
    DevBuf<i64> dwlo, dwli, dblo, dbli, drejc, drejn, drejz, drejt; DevBuf<int> drejk; DevBuf<DevResult> drejr; DevBuf<double> drejf;
    {
        const int64_t* offs[2] = {wl_off, bl_off}; const int64_t* idxs[2] = {wl_idx, bl_idx};
        DevBuf<i64>* doff[2] = {&dwlo, &dblo}; DevBuf<i64>* didx[2] = {&dwli, &dbli};
        for (int w = 0; w < 2; ++w) {
            if (!offs[w]) continue;
            NEED(offs[w][0] == 0, FW_ERR_INVALID, "fw_hiton_pc_ex: list offsets must start at 0");
            for (i64 t = 0; t < n_targets; ++t) NEED(offs[w][t + 1] >= offs[w][t], FW_ERR_INVALID, "fw_hiton_pc_ex: list offsets not monotone");
            const i64 nl = offs[w][n_targets];
            if (nl == 0) continue;
            NEED(idxs[w], FW_ERR_INVALID, "fw_hiton_pc_ex: list indices are NULL");
            std::vector<i64> h(nl);
            for (i64 i = 0; i < nl; ++i) { h[i] = idxs[w][i] - base; NEED(h[i] >= 0 && h[i] < p, FW_ERR_INVALID, "fw_hiton_pc_ex: list entry out of range"); }
            CK(doff[w]->reserve(n_targets + 1)); CK(didx[w]->reserve(nl));
            CK(cudaMemcpyAsync(doff[w]->ptr, offs[w], sizeof(i64) * (n_targets + 1), cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemcpyAsync(didx[w]->ptr, h.data(), sizeof(i64) * nl, cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));                       // h goes out of scope
            if (w == 0) { lists.wl_off = dwlo.ptr; lists.wl_idx = dwli.ptr; } else { lists.bl_off = dblo.ptr; lists.bl_idx = dbli.ptr; }
        }
    }
    if (track) {
        CK(drejc.reserve(n_targets)); CK(drejn.reserve(cap_total)); CK(drejz.reserve((size_t)cap_total * 3)); CK(drejk.reserve(cap_total));
        CK(drejr.reserve(cap_total)); CK(drejt.reserve(cap_total)); CK(drejf.reserve(cap_total));
        CK(cudaMemsetAsync(drejc.ptr, 0, sizeof(i64) * n_targets, ctx->stream));
        lists.rej_count = drejc.ptr; lists.rej_nbr = drejn.ptr; lists.rej_Zs = drejz.ptr; lists.rej_k = drejk.ptr; lists.rej_res = drejr.ptr;
        lists.rej_ntests = drejt.ptr; lists.rej_frac = drejf.ptr;
    }

    if (disc) {
        HitonMiArgs ma;
        ma.lists = lists;
        ma.t = make_mi_table(ctx, kind); ma.hps = hps;
        ma.uni_off = ctx->d_uni_off.ptr; ma.uni_nbr = ctx->d_uni_nbr.ptr; ma.uni_stat = ctx->d_uni_stat.ptr; ma.uni_p = ctx->d_uni_p.ptr;
        ma.targets = dt.ptr; ma.out_off = doff.ptr; ma.counter = ctx->d_counter.ptr;
        ma.max_k = max_k; ma.alpha = alpha; ma.max_tests = max_tests;
        ma.cand_order = dorder.ptr;
        ma.pc_nbr = dpcn.ptr; ma.pc_stat = dpcs.ptr; ma.pc_p = dpcp.ptr; ma.pc_count = dpcc.ptr;
        ma.tpc_nbr = dtpcn.ptr; ma.tpc_stat = dtpcs.ptr; ma.tpc_p = dtpcp.ptr; ma.tpc_count = dtpcc.ptr;
        ma.num_tests = dnt.ptr; ma.executed_total = ctx->d_exec.ptr; ma.status = dstatus.ptr;
        std::vector<int> sel(n_targets);
        std::iota(sel.begin(), sel.end(), 0);
        sort_by_candidates_desc(sel, hoff);
        i64 need = 2; for (i64 t = 0; t < n_targets; ++t) need = std::max<i64>(need, hoff[t + 1] - hoff[t] + 2);
        const int TH = 256; const int L = ma.t.L;
        const size_t tabs = (size_t)(TH / 32) * L * L * L * L * L * sizeof(int) + (size_t)(TH / 32) * MI_BIN_WARP_BYTES;   // + count buffers of the batched binary scan
        // optimistic capacity (accepted sets are far smaller than candidate lists); re-run overflowing targets in the next class
        int caps[3] = {(int)std::min<i64>(need, 32), (int)std::min<i64>(need, 64), (int)need};
        const int n_rounds = need <= 32 ? 1 : (need <= 64 ? 2 : 3);
        CK(cudaEventRecord(ctx->ev[4], ctx->stream));
        std::vector<int> hstatus(n_targets);
        const int Wp = ma.t.W | 1;                                        // odd row stride: conflict-free per-lane plane reads
        for (int round = 0; round < n_rounds && !sel.empty(); ++round) {
            ma.cap = caps[round];
            const size_t fixed = sizeof(i64) * (ma.cap + 1) + 4 * sizeof(double) * ma.cap + sizeof(i64) * ma.cap + 2 * sizeof(int) * ma.cap + (((size_t)ma.cap + 15) & ~(size_t)15);
            // binary tables: the planes of every slot staged in shared memory + the positive-AND count tables (mi_lane.cuh); the
            // region is shared with the buffers of the warp-cooperative scan, which remains the path for everything else
            const size_t lane_bytes = sizeof(unsigned int) * (size_t)ma.cap * Wp + sizeof(int) * MI_LANE_TAB_INTS;
            ma.Wp = Wp;
            ma.lane_ok = (L == 2 && !ma.t.nz && ma.cap <= 64 && fixed + lane_bytes + 16 <= 100 * 1024) ? 1 : 0;
            const size_t gen_tabs = (size_t)(TH / 32) * L * L * L * L * L * sizeof(int);
            size_t smem = fixed + gen_tabs + (ma.lane_ok ? lane_bytes : tabs - gen_tabs) + 16;      // (the batched scan's buffers are unused on the lane path)
            NEED(smem <= 220 * 1024, FW_ERR_UNSUPPORTED, "fw_hiton_pc: a target has %lld candidates; too many for the discrete kernel", (long long)need - 2);
            int n_sel = (int)sel.size();
            CK(cudaMemcpyAsync(dsel.ptr, sel.data(), sizeof(int) * n_sel, cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemsetAsync(ctx->d_counter.ptr, 0, sizeof(int), ctx->stream));
            ma.sel = dsel.ptr; ma.n_sel = n_sel;
            int grid = 1;
            CK(grid_for(hiton_mi_kernel<256, 1>, TH, smem, ctx->sm_count, n_sel, &grid));
            hiton_mi_kernel<256, 1><<<grid, TH, smem, ctx->stream>>>(ma);
            ctx->launches++;
            CK(cudaGetLastError());
            CK(cudaEventRecord(ctx->ev[5], ctx->stream)); ctx->ev_valid[2] = true;
            CK(cudaMemcpyAsync(hstatus.data(), dstatus.ptr, sizeof(int) * n_targets, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            std::vector<int> again;
            for (int t : sel) if (hstatus[t] == 1) again.push_back(t);
            NEED(again.empty() || round + 1 < n_rounds, FW_ERR_UNSUPPORTED, "fw_hiton_pc: capacity overflow in the full-bound class");
            sel.swap(again);
        }
    }
    HitonArgs a;
    a.cv = make_cor_view(ctx); a.p = p;
    a.uni_off = ctx->d_uni_off.ptr; a.uni_nbr = ctx->d_uni_nbr.ptr; a.uni_stat = ctx->d_uni_stat.ptr; a.uni_p = ctx->d_uni_p.ptr;
    a.targets = dt.ptr; a.out_off = doff.ptr; a.counter = ctx->d_counter.ptr;
    a.max_k = max_k; a.alpha = alpha; a.max_tests = max_tests; a.fc = make_fz_consts(ctx->n_obs < 0 ? 0 : ctx->n_obs, n_obs_min);
    fill_fz_band(a.fc, alpha);
    a.cand_order = dorder.ptr;
    a.pc_nbr = dpcn.ptr; a.pc_stat = dpcs.ptr; a.pc_p = dpcp.ptr; a.pc_count = dpcc.ptr;
    a.tpc_nbr = dtpcn.ptr; a.tpc_stat = dtpcs.ptr; a.tpc_p = dtpcp.ptr; a.tpc_count = dtpcc.ptr;
    a.num_tests = dnt.ptr; a.executed_total = ctx->d_exec.ptr; a.status = dstatus.ptr;
    if (nzk) { a.nzt = nzt; a.n_obs_min = n_obs_min; }
    a.lists = lists;
    const int nzw = nzk ? nzt.W : -1;

    // capacity classes: a target needs at most (#candidates + 2) slots; start optimistic (<= 64) and
    // escalate the few targets whose accepted set outgrows the class
    std::vector<int> pending[5];
    for (i64 t = 0; t < n_targets && !disc; ++t) {
        // the accepted set is usually far smaller than the candidate list: up to 94 candidates start in the 32-slot class
        // (tables + p-value-free scan) and are re-run in a larger class only if more than 30 get accepted
        i64 need = (hoff[t + 1] - hoff[t]) + 2;
        pending[need <= 96 ? 0 : 1].push_back((int)t);
    }
    std::vector<int> hstatus(n_targets);
    if (!disc) CK(cudaEventRecord(ctx->ev[4], ctx->stream));
    for (int c = 0; c < 5; ++c) {
        if (pending[c].empty()) continue;
        std::vector<int>& sel = pending[c];
        // longest jobs first (candidate count is the work proxy)
        sort_by_candidates_desc(sel, hoff);
        int n_sel = (int)sel.size();
        CK(cudaMemcpyAsync(dsel.ptr, sel.data(), sizeof(int) * n_sel, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemsetAsync(ctx->d_counter.ptr, 0, sizeof(int), ctx->stream));
        a.sel = dsel.ptr; a.n_sel = n_sel;
        int grid = 1;
        if (c < 4) {
            a.cap = kCaps[c]; a.gscratch = nullptr;
            // per-candidate level-1/level-2 tables fit next to R in the 32-slot class; the table scan does not recover the positions of
            // the returned subset, which rejection records need: tracking runs the generic scan
            const bool cache = !nzk && c == 0 && !track;
            size_t smem = hiton_smem_bytes(a.cap, true, nzw, cache);
            NEED(smem <= 220 * 1024, FW_ERR_UNSUPPORTED, "fw_hiton_pc: target does not fit shared memory (n = %lld rows, %d slots)", (long long)ctx->n, a.cap);
            if (nzk) { CK(grid_for(hiton_fz_kernel<256, 2, true, false, false>, 256, smem, ctx->sm_count, n_sel, &grid)); hiton_fz_kernel<256, 2, true, false, false><<<grid, 256, smem, ctx->stream>>>(a); }
            else if (cache && !lists.wl_off && !lists.bl_off) { CK(grid_for(hiton_fz_kernel<128, 4, false, false, true, false>, 128, smem, ctx->sm_count, n_sel, &grid)); hiton_fz_kernel<128, 4, false, false, true, false><<<grid, 128, smem, ctx->stream>>>(a); }
            else if (cache) { CK(grid_for(hiton_fz_kernel<128, 4, false, false, true>, 128, smem, ctx->sm_count, n_sel, &grid)); hiton_fz_kernel<128, 4, false, false, true><<<grid, 128, smem, ctx->stream>>>(a); }
            else if (c <= 1) { CK(grid_for(hiton_fz_kernel<128, 2, false, false, false>, 128, smem, ctx->sm_count, n_sel, &grid)); hiton_fz_kernel<128, 2, false, false, false><<<grid, 128, smem, ctx->stream>>>(a); }
            else { CK(grid_for(hiton_fz_kernel<256, 2, false, false, false>, 256, smem, ctx->sm_count, n_sel, &grid)); hiton_fz_kernel<256, 2, false, false, false><<<grid, 256, smem, ctx->stream>>>(a); }
        } else {
            i64 need = 0; for (int t : sel) need = std::max<i64>(need, hoff[t + 1] - hoff[t] + 2);
            NEED(need <= 3000, FW_ERR_UNSUPPORTED, "fw_hiton_pc: a target has %lld candidates; more than 2998 accepted neighbours are not supported", (long long)need - 2);
            a.cap = (int)need;
            size_t smem = hiton_smem_bytes(a.cap, false, nzw);
            if (nzk) CK(grid_for(hiton_fz_kernel<256, 2, true, true, false>, 256, smem, ctx->sm_count, n_sel, &grid));
            else CK(grid_for(hiton_fz_kernel<256, 2, false, true, false>, 256, smem, ctx->sm_count, n_sel, &grid));
            size_t per = (size_t)a.cap * a.cap;
            while (grid > 1 && per * grid * sizeof(float) > ((size_t)8 << 30)) grid = (grid + 1) / 2;
            CK(gs.reserve(per * grid));
            a.gscratch = gs.ptr;
            if (nzk) hiton_fz_kernel<256, 2, true, true, false><<<grid, 256, smem, ctx->stream>>>(a);
            else hiton_fz_kernel<256, 2, false, true, false><<<grid, 256, smem, ctx->stream>>>(a);
        }
        ctx->launches++;
        CK(cudaGetLastError());
        CK(cudaEventRecord(ctx->ev[5], ctx->stream)); ctx->ev_valid[2] = true;
        CK(cudaMemcpyAsync(hstatus.data(), dstatus.ptr, sizeof(int) * n_targets, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (int t : sel) if (hstatus[t] == 1) {
            NEED(c < 4, FW_ERR_UNSUPPORTED, "fw_hiton_pc: capacity overflow in the unbounded class");
            pending[c + 1].push_back(t);
        }
    }
    CK(cudaStreamSynchronize(ctx->stream));
    // copy out
    if (pc_off) memcpy(pc_off, hoff.data(), sizeof(i64) * (n_targets + 1));
    if (pc_count) CK(cudaMemcpyAsync(pc_count, dpcc.ptr, sizeof(i64) * n_targets, cudaMemcpyDeviceToHost, ctx->stream));
    if (tpc_count) CK(cudaMemcpyAsync(tpc_count, dtpcc.ptr, sizeof(i64) * n_targets, cudaMemcpyDeviceToHost, ctx->stream));
    if (num_tests) CK(cudaMemcpyAsync(num_tests, dnt.ptr, sizeof(i64) * n_targets, cudaMemcpyDeviceToHost, ctx->stream));
    if (hoff[n_targets] > 0) {
        i64 ne = hoff[n_targets];
        if (pc_nbr) CK(cudaMemcpyAsync(pc_nbr, dpcn.ptr, sizeof(i64) * ne, cudaMemcpyDeviceToHost, ctx->stream));
        if (pc_stat) CK(cudaMemcpyAsync(pc_stat, dpcs.ptr, sizeof(double) * ne, cudaMemcpyDeviceToHost, ctx->stream));
        if (pc_p) CK(cudaMemcpyAsync(pc_p, dpcp.ptr, sizeof(double) * ne, cudaMemcpyDeviceToHost, ctx->stream));
        if (tpc_nbr) CK(cudaMemcpyAsync(tpc_nbr, dtpcn.ptr, sizeof(i64) * ne, cudaMemcpyDeviceToHost, ctx->stream));
        if (tpc_stat) CK(cudaMemcpyAsync(tpc_stat, dtpcs.ptr, sizeof(double) * ne, cudaMemcpyDeviceToHost, ctx->stream));
        if (tpc_p) CK(cudaMemcpyAsync(tpc_p, dtpcp.ptr, sizeof(double) * ne, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (track) {
        CK(cudaMemcpyAsync(rej_count, drejc.ptr, sizeof(i64) * n_targets, cudaMemcpyDeviceToHost, ctx->stream));
        if (hoff[n_targets] > 0) {
            const i64 ne = hoff[n_targets];
            CK(cudaMemcpyAsync(rej_nbr, drejn.ptr, sizeof(i64) * ne, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaMemcpyAsync(rej_Zs, drejz.ptr, sizeof(i64) * ne * 3, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaMemcpyAsync(rej_k, drejk.ptr, sizeof(int) * ne, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaMemcpyAsync(rej_result, drejr.ptr, sizeof(DevResult) * ne, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaMemcpyAsync(rej_num_tests, drejt.ptr, sizeof(i64) * ne, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaMemcpyAsync(rej_frac, drejf.ptr, sizeof(double) * ne, cudaMemcpyDeviceToHost, ctx->stream));
        }
    }
    u64 hexec[4] = {0, 0, 0, 0};
    CK(cudaMemcpyAsync(hexec, ctx->d_exec.ptr, 4 * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    { int st_ = comm_check(ctx); if (st_ != FW_OK) return st_; }
    if (tests_executed_total) *tests_executed_total = (i64)hexec[0];
    for (int i = 0; i < 3; ++i) ctx->exec_by_k[i] = (i64)hexec[i + 1];
    if (base) {
        // only the valid prefix of each target's slot range holds variable ids
        std::vector<i64> pcc(n_targets), tpcc(n_targets);
        CK(cudaMemcpy(pcc.data(), dpcc.ptr, sizeof(i64) * n_targets, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(tpcc.data(), dtpcc.ptr, sizeof(i64) * n_targets, cudaMemcpyDeviceToHost));
        for (i64 t = 0; t < n_targets; ++t) {
            if (pc_nbr) for (i64 i = 0; i < pcc[t]; ++i) pc_nbr[hoff[t] + i] += base;
            if (tpc_nbr) for (i64 i = 0; i < tpcc[t]; ++i) tpc_nbr[hoff[t] + i] += base;
            if (track) for (i64 i = 0; i < rej_count[t]; ++i) {
                rej_nbr[hoff[t] + i] += base;
                for (int j = 0; j < 3; ++j) if (rej_Zs[(hoff[t] + i) * 3 + j] >= 0) rej_Zs[(hoff[t] + i) * 3 + j] += base;
            }
        }
    }
    return FW_OK;
}

}  // extern "C"
