// pairwise.cuh — the univariate stage: pw_univar_neighbors (src/tests.jl:436-532) for the
// Fisher-z test on a resident cor_mat (the `test_name == "fz"` lookup branch, tests.jl:149-156,
// 470-478), Benjamini-Hochberg (src/statfuns.jl:326-350) and the neighbour lists of
// condensed_stats_to_dict (tests.jl:372-388), all on the device.
//
// HBM-bound: the algorithmic traffic is one 4-byte correlation per pair (2*p*(p-1) bytes
// for the upper triangle).  Raw candidates (|r| within reach of the alpha threshold: a
// conservative float compare) are collected either for free in the epilogue of the cor_mat
// GEMM (cor_tc.cuh, when fw_pairwise_prefetch announced the parameters) or by ONE coalesced
// pass over the resident matrix (pw_fz_emit_kernel); the exact p-value (fp64 log + erfc) is
// evaluated for the candidates only (~1 % of the pairs), densely, by pw_fz_select_kernel.
// The candidates are unordered: Benjamini-Hochberg's result does not depend on the order of
// tied p-values (every member of a tie receives the same adjusted value: the running minimum
// of p*m/i reaches all of them from the last rank of the tie), and the neighbour lists are
// sorted by (row, column) key afterwards.
#pragma once
#include <cub/cub.cuh>
#include <string>
#include "common.cuh"
#include "fz.cuh"

struct PairwiseOut {
    i64* d_off = nullptr; i64* d_nbr = nullptr; double* d_stat = nullptr; double* d_adjp = nullptr;
    i64 n_entries = 0;      // directed entries (2 per significant pair)
    i64 n_tests = 0, n_reliable = 0, n_raw_sig = 0;
};

struct PairwiseScratch {
    void* bufs[20] = {nullptr}; size_t sizes[20] = {0};
    cudaError_t get(int i, size_t bytes, void** out) {
        if (bytes == 0) bytes = 16;
        if (sizes[i] < bytes) {
            if (bufs[i]) cudaFree(bufs[i]);
            bufs[i] = nullptr; sizes[i] = 0;
            cudaError_t e = cudaMalloc(&bufs[i], bytes);
            if (e != cudaSuccess) return e;
            sizes[i] = bytes;
        }
        *out = bufs[i];
        return cudaSuccess;
    }
    ~PairwiseScratch() { for (int i = 0; i < 20; ++i) if (bufs[i]) cudaFree(bufs[i]); }
};

// One pass over the upper triangle of the locally held cor_mat rows: one CTA per row X, coalesced 128-byte row segments, every
// |r| >= r_lo (or NaN: counted) becomes a raw candidate.  Hits are compacted into a per-warp shared-memory queue and flushed
// with ONE global atomic per >= 64 records, so the append costs ~2e5 atomics at C4 instead of one per warp iteration.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) pw_fz_emit_kernel(const float* __restrict__ rows, i64 p, i64 row_global0, i64 row_local0, PwEmit e) {
    constexpr int WARPS = THREADS / 32, QCAP = 96;     // a flush happens at >= 64 queued records; one ballot adds at most 32
    __shared__ PwRec q[WARPS][QCAP];
    const i64 X = row_global0 + blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* row = rows + (row_local0 + blockIdx.x) * p;
    int qn = 0;                                         // warp-uniform
    unsigned int n_nan = 0;
    auto flush = [&]() {
        u64 pos0 = 0;
        if (lane == 0) pos0 = atomicAdd(&e.counters[0], (u64)qn);
        pos0 = __shfl_sync(0xffffffffu, pos0, 0);
        for (int i = lane; i < qn; i += 32) if ((i64)(pos0 + i) < e.cap) e.list[pos0 + i] = q[warp][i];
        __syncwarp();
        qn = 0;
    };
    auto consider = [&](i64 Y, float r, bool in_range) {
        bool hit = false;
        if (in_range) { if (r != r) ++n_nan; else hit = fabsf(r) >= e.r_lo; }
        const unsigned int bal = __ballot_sync(0xffffffffu, hit);
        if (bal) {
            if (hit) { PwRec rec; rec.x = (int)X; rec.y = (int)Y; rec.r = r; q[warp][qn + __popc(bal & ((1u << lane) - 1u))] = rec; }
            qn += __popc(bal);
            __syncwarp();
            if (qn >= 64) flush();
        }
    };
    if ((p & 3) == 0) {
        // 16-byte loads: the row base is 16-byte aligned when p % 4 == 0; the first vector is masked below Y = X + 1
        const float4* row4 = reinterpret_cast<const float4*>(row);
        for (i64 v0 = (X + 1) >> 2; v0 < (p >> 2); v0 += THREADS) {
            const i64 v = v0 + tid;
            const bool in = v < (p >> 2);
            float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
            if (in) f = __ldg(row4 + v);
            const i64 Y = v << 2;
            consider(Y, f.x, in && Y > X); consider(Y + 1, f.y, in && Y + 1 > X); consider(Y + 2, f.z, in && Y + 2 > X); consider(Y + 3, f.w, in && Y + 3 > X);
        }
    } else {
        for (i64 y0 = X + 1; y0 < p; y0 += THREADS) {
            const i64 Y = y0 + tid;
            consider(Y, Y < p ? row[Y] : 0.0f, Y < p);
        }
    }
    if (qn) flush();
    n_nan = __reduce_add_sync(0xffffffffu, n_nan);
    if (lane == 0 && n_nan) atomicAdd(&e.counters[1], (u64)n_nan);
}

// exact p-value of every raw candidate; the raw-significant ones (p < alpha) are appended (unordered, one atomic per warp) to the
// compacted arrays the Benjamini-Hochberg stage works on
__global__ void pw_fz_select_kernel(const PwRec* __restrict__ recs, i64 n, FzConsts fc, double alpha, u64* counter,
                                    int* c_x, int* c_y, double* c_p, double* c_stat) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool sig = false; PwRec rec; double pv = 0.0;
    if (i < n) { rec = recs[i]; pv = fz_pval_dev((double)rec.r, fc); sig = pv < alpha; }
    const unsigned int bal = __ballot_sync(0xffffffffu, sig);
    if (!bal) return;
    u64 pos0 = 0;
    if (lane == 0) pos0 = atomicAdd(counter, (u64)__popc(bal));
    pos0 = __shfl_sync(0xffffffffu, pos0, 0);
    if (sig) {
        const u64 pos = pos0 + (u64)__popc(bal & ((1u << lane) - 1u));
        c_x[pos] = rec.x; c_y[pos] = rec.y; c_p[pos] = pv; c_stat[pos] = (double)rec.r;
    }
}

__global__ void pw_u32_to_i64(const unsigned int* in, i64* out, i64 n) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (i64)in[i];
}
__global__ void pw_make_keys(const double* c_p, u64* keys, unsigned int* vals, i64 n) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { keys[i] = (u64)__double_as_longlong(c_p[i]); vals[i] = (unsigned int)i; }
}
// statfuns.jl:335-344: v[i] = p_(i) * m / i (1-based rank), last one additionally capped at 1.0;
// written reversed so that a forward inclusive min-scan yields the step-up adjustment.
__global__ void pw_bh_terms(const u64* sorted_keys, double* rev, i64 nf, double m) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf) return;
    double pv = __longlong_as_double((long long)sorted_keys[i]);
    double v = __ddiv_rn(__dmul_rn(pv, m), (double)(i + 1));
    if (i == nf - 1) v = fmin(v, 1.0);
    rev[nf - 1 - i] = v;
}
struct MinOp { __device__ __forceinline__ double operator()(double a, double b) const { return fmin(a, b); } };
// scatter the adjusted p back to compaction (condensed-index) order and flag survivors (tests.jl:381)
__global__ void pw_bh_scatter(const double* rev_scanned, const unsigned int* sorted_vals, i64 nf, double alpha, double* adj, unsigned char* keep) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf) return;
    double a = rev_scanned[nf - 1 - i];
    unsigned int pos = sorted_vals[i];
    adj[pos] = a; keep[pos] = (!isnan(a) && a < alpha) ? 1 : 0;
}
__global__ void pw_nofdr_keep(const double* c_p, i64 nf, double alpha, double* adj, unsigned char* keep) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf) return;
    adj[i] = c_p[i]; keep[i] = (c_p[i] < alpha) ? 1 : 0;
}
// directed entries (X->Y) and (Y->X) keyed by row*p + col for the survivors
__global__ void pw_emit_directed(const int* c_x, const int* c_y, const unsigned char* keep, const i64* keep_pos, i64 nf, i64 p,
                                 u64* keys, unsigned int* src) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf || !keep[i]) return;
    i64 o = keep_pos[i] * 2;
    keys[o] = (u64)((i64)c_x[i] * p + c_y[i]); src[o] = (unsigned int)i;
    keys[o + 1] = (u64)((i64)c_y[i] * p + c_x[i]); src[o + 1] = (unsigned int)i;
}
__global__ void pw_fill_csr(const u64* sorted_keys, const unsigned int* sorted_src, i64 ne, i64 p, const double* c_stat, const double* adj,
                            i64* nbr, double* stat, double* adjp) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ne) return;
    u64 k = sorted_keys[i];
    nbr[i] = (i64)(k % (u64)p);
    unsigned int s = sorted_src[i];
    stat[i] = c_stat[s]; adjp[i] = adj[s];
}
// offsets[v] = first entry whose row >= v
__global__ void pw_row_offsets(const u64* sorted_keys, i64 ne, i64 p, i64* off) {
    i64 v = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (v > p) return;
    u64 target = (u64)v * (u64)p;
    i64 lo = 0, hi = ne;
    while (lo < hi) { i64 mid = (lo + hi) >> 1; if (sorted_keys[mid] < target) lo = mid + 1; else hi = mid; }
    off[v] = lo;
}
__global__ void pw_flags_to_i64(const unsigned char* keep, i64* out, i64 n) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = keep[i];
}

#define PWCK(call, what) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { if (msg) *msg = what; return e_; } } while (0)

static inline unsigned pw_blocks(i64 n, int t) { return (unsigned)((n + t - 1) / t); }

// Common second half of the pairwise stage: Benjamini-Hochberg on the compacted raw-significant pairs (which must be in
// condensed-index order: x ascending, then y ascending), then the symmetric neighbour CSR.
static cudaError_t pairwise_finish(PairwiseScratch& S, int* c_x, int* c_y, double* c_p, double* c_stat, i64 nf, i64 m, i64 p, double alpha, bool fdr,
                                   cudaStream_t st, PairwiseOut* out, int* n_launch, std::string* msg) {
    const int T = 256;
    // cub's sorts / scans take 32-bit item counts here: a denser univariate network than 2^31 - 1 raw-significant pairs is refused, not truncated
    if (nf >= ((i64)1 << 31) - 1) { if (msg) *msg = "more than 2^31 raw-significant pairs are not supported (unsupported size)"; return cudaErrorInvalidValue; }
    void* tmp = nullptr; size_t tmp_bytes = 0, need = 0;
    double *adj, *rev; unsigned char* keep; u64 *keys, *keys2; unsigned int *vals, *vals2; i64 *keep64, *keep_pos;
    i64* d_off = out->d_off;
    // one arena for the BH / CSR temporaries
    size_t a_adj = sizeof(double) * nf, a_rev = sizeof(double) * nf, a_keep = (size_t)nf, a_keys = sizeof(u64) * 2 * nf, a_vals = sizeof(unsigned int) * 2 * nf, a_k64 = sizeof(i64) * nf;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t arena_bytes = al(a_adj) + al(a_rev) + al(a_keep) + 2 * al(a_keys) + 2 * al(a_vals) + 2 * al(a_k64);
    unsigned char* arena; PWCK(S.get(11, arena_bytes, (void**)&arena), "alloc");
    size_t o = 0;
    adj = (double*)(arena + o); o += al(a_adj);
    rev = (double*)(arena + o); o += al(a_rev);
    keep = (unsigned char*)(arena + o); o += al(a_keep);
    keys = (u64*)(arena + o); o += al(a_keys);
    keys2 = (u64*)(arena + o); o += al(a_keys);
    vals = (unsigned int*)(arena + o); o += al(a_vals);
    vals2 = (unsigned int*)(arena + o); o += al(a_vals);
    keep64 = (i64*)(arena + o); o += al(a_k64);
    keep_pos = (i64*)(arena + o); o += al(a_k64);

    if (fdr) {
        pw_make_keys<<<pw_blocks(nf, T), T, 0, st>>>(c_p, keys, vals, nf); (*n_launch)++;
        cub::DeviceRadixSort::SortPairs(nullptr, need, keys, keys2, vals, vals2, (int)nf, 0, 64, st);
        if (need > tmp_bytes) { tmp_bytes = need; PWCK(S.get(5, tmp_bytes, &tmp), "alloc"); }
        PWCK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, vals, vals2, (int)nf, 0, 64, st), "sort"); (*n_launch) += 8;
        pw_bh_terms<<<pw_blocks(nf, T), T, 0, st>>>(keys2, rev, nf, (double)m); (*n_launch)++;
        cub::DeviceScan::InclusiveScan(nullptr, need, rev, rev, MinOp(), (int)nf, st);
        if (need > tmp_bytes) { tmp_bytes = need; PWCK(S.get(5, tmp_bytes, &tmp), "alloc"); }
        PWCK(cub::DeviceScan::InclusiveScan(tmp, tmp_bytes, rev, rev, MinOp(), (int)nf, st), "minscan"); (*n_launch)++;
        pw_bh_scatter<<<pw_blocks(nf, T), T, 0, st>>>(rev, vals2, nf, alpha, adj, keep); (*n_launch)++;
    } else {
        pw_nofdr_keep<<<pw_blocks(nf, T), T, 0, st>>>(c_p, nf, alpha, adj, keep); (*n_launch)++;
    }
    pw_flags_to_i64<<<pw_blocks(nf, T), T, 0, st>>>(keep, keep64, nf); (*n_launch)++;
    cub::DeviceScan::ExclusiveSum(nullptr, need, keep64, keep_pos, (int)nf, st);
    if (need > tmp_bytes) { tmp_bytes = need; PWCK(S.get(5, tmp_bytes, &tmp), "alloc"); }
    PWCK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, keep64, keep_pos, (int)nf, st), "scan"); (*n_launch)++;
    i64 last_pos = 0; unsigned char last_keep = 0;
    PWCK(cudaMemcpyAsync(&last_pos, keep_pos + nf - 1, sizeof(i64), cudaMemcpyDeviceToHost, st), "d2h");
    PWCK(cudaMemcpyAsync(&last_keep, keep + nf - 1, 1, cudaMemcpyDeviceToHost, st), "d2h");
    PWCK(cudaStreamSynchronize(st), "sync");
    const i64 n_keep = last_pos + last_keep, ne = 2 * n_keep;
    if (ne >= ((i64)1 << 31) - 1) { if (msg) *msg = "more than 2^31 neighbour-list entries are not supported (unsupported size)"; return cudaErrorInvalidValue; }
    out->n_entries = ne;
    if (ne == 0) { PWCK(cudaMemsetAsync(d_off, 0, sizeof(i64) * (p + 1), st), "memset"); return cudaSuccess; }
    pw_emit_directed<<<pw_blocks(nf, T), T, 0, st>>>(c_x, c_y, keep, keep_pos, nf, p, keys, vals); (*n_launch)++;
    int end_bit = 1; while (((u64)1 << end_bit) < (u64)p * (u64)p && end_bit < 64) ++end_bit;
    cub::DeviceRadixSort::SortPairs(nullptr, need, keys, keys2, vals, vals2, (int)ne, 0, end_bit, st);
    if (need > tmp_bytes) { tmp_bytes = need; PWCK(S.get(5, tmp_bytes, &tmp), "alloc"); }
    PWCK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, vals, vals2, (int)ne, 0, end_bit, st), "sort"); (*n_launch) += 8;
    i64* d_nbr; double *d_stat, *d_adjp;
    PWCK(S.get(12, sizeof(i64) * ne, (void**)&d_nbr), "alloc");
    PWCK(S.get(13, sizeof(double) * ne, (void**)&d_stat), "alloc");
    PWCK(S.get(14, sizeof(double) * ne, (void**)&d_adjp), "alloc");
    pw_fill_csr<<<pw_blocks(ne, T), T, 0, st>>>(keys2, vals2, ne, p, c_stat, adj, d_nbr, d_stat, d_adjp); (*n_launch)++;
    pw_row_offsets<<<pw_blocks(p + 1, T), T, 0, st>>>(keys2, ne, p, d_off); (*n_launch)++;
    PWCK(cudaGetLastError(), "csr kernels");
    out->d_nbr = d_nbr; out->d_stat = d_stat; out->d_adjp = d_adjp;
    return cudaSuccess;
}

// gather the unordered emission of the discrete pairwise kernel into condensed-index order
__global__ void pw_pair_keys(const int* c_x, const int* c_y, i64 p, u64* keys, unsigned int* vals, i64 n) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { keys[i] = (u64)((i64)c_x[i] * p + c_y[i]); vals[i] = (unsigned int)i; }
}
__global__ void pw_gather_pairs(const unsigned int* perm, const int* sx, const int* sy, const double* sp, const double* ss,
                                int* dx, int* dy, double* dp, double* ds, i64 n) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { unsigned int j = perm[i]; dx[i] = sx[j]; dy[i] = sy[j]; dp[i] = sp[j]; ds[i] = ss[j]; }
}

// conservative |r| pre-filter of the univariate Fisher-z test: p < alpha  <=>  |r| > tanh(z_alpha / sqrt(n-3)); 1e-4 relative safety band.
// 2.0f: nothing can be significant (n - 3 <= 0: z = 0, p = 1; rows < n_obs_min: p = 1)
static float pairwise_fz_r_lo(const FzConsts& fc, i64 n_rows, i64 n_obs_min, double alpha) {
    if (!(fc.sf_pos && n_rows >= n_obs_min) || !(alpha > 0.0)) return 2.0f;
    if (alpha >= 1.0) return 0.0f;
    double lo = 0.0, hi = 40.0;                    // z_alpha from erfc(z/sqrt2) = alpha by bisection (host, once)
    for (int it = 0; it < 200; ++it) { double mid = 0.5 * (lo + hi); if (std::erfc(mid * 0.70710678118654752440) > alpha) lo = mid; else hi = mid; }
    return (float)(std::tanh(lo / (2.0 * fc.half_sqrt_sf)) * (1.0 - 1e-4));
}

// one pass over `n_rows` locally held rows (global row ids row_global0.., stored from local row row_local0 of `rows`)
static cudaError_t pairwise_fz_emit(const float* rows, i64 p, i64 row_global0, i64 row_local0, i64 n_rows, const PwEmit& e, cudaStream_t st,
                                    int* n_launch, std::string* msg) {
    if (n_rows <= 0) return cudaSuccess;
    pw_fz_emit_kernel<256><<<(unsigned)n_rows, 256, 0, st>>>(rows, p, row_global0, row_local0, e);
    (*n_launch)++;
    PWCK(cudaGetLastError(), "pw_fz_emit_kernel");
    return cudaSuccess;
}

// second half: exact p-values of the nrec raw candidates (of all ranks), Benjamini-Hochberg, neighbour CSR
static cudaError_t pairwise_fz_tail(PairwiseScratch& S, const PwRec* recs, i64 nrec, i64 n_nan, i64 p, FzConsts fc, i64 n_rows, i64 n_obs_min, double alpha,
                                    bool fdr, bool reliable_only, cudaStream_t st, PairwiseOut* out, int* n_launch, std::string* msg) {
    const int T = 256;
    const i64 n_pairs = p * (p - 1) / 2;
    out->n_tests = n_pairs;
    // suff_power of every univariate fz test is (n >= n_obs_min) (tests.jl:159); tests.jl:397-402: unreliable tests are stored as NaN,
    // a NaN correlation gives a NaN p-value; tests.jl:521-526: m = number of tests, minus the NaN ones when correct_reliable_only
    const bool suff_all = n_rows >= n_obs_min;
    const i64 n_rel = (suff_all || !reliable_only) ? n_pairs - n_nan : 0;
    const i64 m = reliable_only ? n_rel : n_pairs;
    out->n_reliable = n_rel;
    i64* d_off; PWCK(S.get(6, sizeof(i64) * (p + 1), (void**)&d_off), "alloc");
    out->d_off = d_off;
    i64 nf = 0;
    int *c_x = nullptr, *c_y = nullptr; double *c_p = nullptr, *c_stat = nullptr;
    if (nrec > 0 && n_rel > 0) {
        u64* cnt; PWCK(S.get(4, sizeof(u64) * 4, (void**)&cnt), "alloc");
        PWCK(S.get(7, sizeof(int) * nrec, (void**)&c_x), "alloc");
        PWCK(S.get(8, sizeof(int) * nrec, (void**)&c_y), "alloc");
        PWCK(S.get(9, sizeof(double) * nrec, (void**)&c_p), "alloc");
        PWCK(S.get(10, sizeof(double) * nrec, (void**)&c_stat), "alloc");
        PWCK(cudaMemsetAsync(cnt, 0, sizeof(u64), st), "memset");
        pw_fz_select_kernel<<<pw_blocks(nrec, T), T, 0, st>>>(recs, nrec, fc, alpha, cnt, c_x, c_y, c_p, c_stat); (*n_launch)++;
        PWCK(cudaGetLastError(), "pw_fz_select_kernel");
        u64 h = 0;
        PWCK(cudaMemcpyAsync(&h, cnt, sizeof(u64), cudaMemcpyDeviceToHost, st), "d2h");
        PWCK(cudaStreamSynchronize(st), "sync");
        nf = (i64)h;
    }
    out->n_raw_sig = nf;
    if (nf == 0) {
        PWCK(cudaMemsetAsync(d_off, 0, sizeof(i64) * (p + 1), st), "memset");
        out->n_entries = 0; out->d_nbr = nullptr; out->d_stat = nullptr; out->d_adjp = nullptr;
        return cudaSuccess;
    }
    return pairwise_finish(S, c_x, c_y, c_p, c_stat, nf, m, p, alpha, fdr, st, out, n_launch, msg);
}
