// pairwise.cuh — the univariate stage: pw_univar_neighbors (src/tests.jl:436-532) for the
// Fisher-z test on a resident cor_mat (the `test_name == "fz"` lookup branch, tests.jl:149-156,
// 470-478), Benjamini-Hochberg (src/statfuns.jl:326-350) and the neighbour lists of
// condensed_stats_to_dict (tests.jl:372-388), all on the device.
//
// HBM-bound: the algorithmic traffic is one 4-byte correlation per pair (2*p*(p-1) bytes
// for the upper triangle), read twice (count pass + write pass) with fully coalesced
// 128-byte row segments.  The exact p-value (fp64 log + erfc) is only evaluated for pairs
// whose |r| is within reach of the alpha threshold; everything else is rejected by a
// conservative float compare, so the fp64 pipe stays off the critical path.
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

// One CTA per row X: counts (pass 0) or writes in ascending-Y order (pass 1) the pairs with raw p < alpha.
// row_cnt[X] = {#raw-significant, #reliable (non-NaN p)}.
template <int THREADS, int PASS>
__global__ void __launch_bounds__(THREADS) pw_fz_rows_kernel(const float* __restrict__ cor, i64 p, FzConsts fc, double alpha, float r_lo,
                                                             int reliable_only, int suff_all,
                                                             unsigned int* row_sig, unsigned int* row_rel, const i64* row_base,
                                                             int* c_x, int* c_y, double* c_p, double* c_stat) {
    const i64 X = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* row = cor + X * p;
    __shared__ unsigned int s_w[THREADS / 32];
    __shared__ unsigned int s_base;
    unsigned int n_sig = 0, n_rel = 0;
    if (PASS == 1 && tid == 0) s_base = 0;
    if (PASS == 1) __syncthreads();
    const i64 base_out = PASS == 1 ? row_base[X] : 0;
    for (i64 y0 = X + 1; y0 < p; y0 += THREADS) {
        const i64 Y = y0 + tid;
        bool sig = false, rel = false;
        double pv = 0.0; float r = 0.0f;
        if (Y < p) {
            r = __ldg(row + Y);
            // tests.jl:397-402: unreliable tests are stored as NaN; a NaN correlation gives a NaN p-value
            rel = !isnan(r) && (suff_all || !reliable_only);
            if (rel && fabsf(r) >= r_lo) {
                pv = fz_pval_dev((double)r, fc);
                sig = pv < alpha;
            }
        }
        if (PASS == 0) {
            n_sig += sig; n_rel += rel;
        } else {
            unsigned int bal = __ballot_sync(0xffffffffu, sig);
            if (lane == 0) s_w[warp] = __popc(bal);
            __syncthreads();
            unsigned int woff = 0, tot = 0;
            for (int w = 0; w < THREADS / 32; ++w) { unsigned int c = s_w[w]; if (w < warp) woff += c; tot += c; }
            unsigned int b = s_base;
            if (sig) {
                i64 pos = base_out + b + woff + __popc(bal & ((1u << lane) - 1u));
                c_x[pos] = (int)X; c_y[pos] = (int)Y; c_p[pos] = pv; c_stat[pos] = (double)r;
            }
            __syncthreads();
            if (tid == 0) s_base = b + tot;
        }
    }
    if (PASS == 0) {
        typedef cub::BlockReduce<unsigned int, THREADS> BR;
        __shared__ typename BR::TempStorage tmp;
        unsigned int ts = BR(tmp).Sum(n_sig);
        __syncthreads();
        unsigned int tr = BR(tmp).Sum(n_rel);
        if (tid == 0) { row_sig[X] = ts; row_rel[X] = tr; }
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

static cudaError_t pairwise_fz_run(PairwiseScratch& S, const float* d_cor, i64 p, FzConsts fc, i64 n_rows, i64 n_obs_min, double alpha,
                                   bool fdr, bool reliable_only, int sm_count, cudaStream_t st, PairwiseOut* out, int* n_launch, std::string* msg) {
    (void)sm_count;
    const int T = 256;
    const i64 n_pairs = p * (p - 1) / 2;
    out->n_tests = n_pairs;
    // suff_power of every univariate fz test is (n >= n_obs_min) (tests.jl:159); below that the stat is 0 and p is 1
    const int suff_all = n_rows >= n_obs_min ? 1 : 0;
    // conservative pre-filter: p < alpha  <=>  |r| > tanh(z_alpha / sqrt(n-3)); keep a 1e-4 relative safety band
    float r_lo = 2.0f;   // nothing can be significant when n - 3 <= 0 (z = 0, p = 1) or rows < n_obs_min (p = 1)
    if (fc.sf_pos && suff_all) {
        // z_alpha from erfc(z/sqrt2) = alpha by bisection (host, once)
        double lo = 0.0, hi = 40.0;
        for (int it = 0; it < 200; ++it) { double mid = 0.5 * (lo + hi); if (std::erfc(mid * 0.70710678118654752440) > alpha) lo = mid; else hi = mid; }
        double rc = std::tanh(lo / (2.0 * fc.half_sqrt_sf));
        r_lo = (float)(rc * (1.0 - 1e-4));
        if (!(alpha > 0.0)) r_lo = 2.0f;
        if (alpha >= 1.0) r_lo = 0.0f;
    }
    unsigned int *row_sig, *row_rel; i64 *row_sig64, *row_base, *tot2;
    PWCK(S.get(0, sizeof(unsigned int) * (p + 1), (void**)&row_sig), "alloc");
    PWCK(S.get(1, sizeof(unsigned int) * (p + 1), (void**)&row_rel), "alloc");
    PWCK(S.get(2, sizeof(i64) * (p + 1), (void**)&row_sig64), "alloc");
    PWCK(S.get(3, sizeof(i64) * (p + 1), (void**)&row_base), "alloc");
    PWCK(S.get(4, sizeof(i64) * 4, (void**)&tot2), "alloc");
    PWCK(cudaMemsetAsync(row_sig, 0, sizeof(unsigned int) * (p + 1), st), "memset");
    PWCK(cudaMemsetAsync(row_rel, 0, sizeof(unsigned int) * (p + 1), st), "memset");
    pw_fz_rows_kernel<T, 0><<<(unsigned)p, T, 0, st>>>(d_cor, p, fc, alpha, r_lo, reliable_only ? 1 : 0, suff_all, row_sig, row_rel, nullptr, nullptr, nullptr, nullptr, nullptr);
    (*n_launch)++;
    PWCK(cudaGetLastError(), "pw_fz_rows_kernel<0>");
    // exclusive scan of the per-row counts (+ totals)
    void* tmp = nullptr; size_t tmp_bytes = 0, need = 0;
    pw_u32_to_i64<<<pw_blocks(p + 1, T), T, 0, st>>>(row_sig, row_sig64, p + 1); (*n_launch)++;
    cub::DeviceScan::ExclusiveSum(nullptr, need, row_sig64, row_base, (int)(p + 1), st);
    tmp_bytes = need;
    PWCK(S.get(5, tmp_bytes, &tmp), "alloc");
    PWCK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, row_sig64, row_base, (int)(p + 1), st), "scan"); (*n_launch)++;
    i64* rel64 = row_sig64;   // reuse after the scan
    pw_u32_to_i64<<<pw_blocks(p + 1, T), T, 0, st>>>(row_rel, rel64, p + 1); (*n_launch)++;
    cub::DeviceReduce::Sum(nullptr, need, rel64, tot2, (int)(p + 1), st);
    if (need > tmp_bytes) { tmp_bytes = need; PWCK(S.get(5, tmp_bytes, &tmp), "alloc"); }
    PWCK(cub::DeviceReduce::Sum(tmp, tmp_bytes, rel64, tot2, (int)(p + 1), st), "reduce"); (*n_launch)++;
    i64 h_tot[2] = {0, 0};
    PWCK(cudaMemcpyAsync(&h_tot[0], row_base + p, sizeof(i64), cudaMemcpyDeviceToHost, st), "d2h");
    PWCK(cudaMemcpyAsync(&h_tot[1], tot2, sizeof(i64), cudaMemcpyDeviceToHost, st), "d2h");
    PWCK(cudaStreamSynchronize(st), "sync");
    const i64 nf = h_tot[0];
    // tests.jl:521-526: m = number of tests, minus the NaN ones when correct_reliable_only
    i64 n_nan = n_pairs - h_tot[1];
    const i64 m = reliable_only ? n_pairs - n_nan : n_pairs;
    // (without correct_reliable_only the NaN count only contains NaN correlations, which the reference keeps in m)
    out->n_reliable = h_tot[1]; out->n_raw_sig = nf;

    i64* d_off; PWCK(S.get(6, sizeof(i64) * (p + 1), (void**)&d_off), "alloc");
    out->d_off = d_off;
    if (nf == 0) {
        PWCK(cudaMemsetAsync(d_off, 0, sizeof(i64) * (p + 1), st), "memset");
        out->n_entries = 0; out->d_nbr = nullptr; out->d_stat = nullptr; out->d_adjp = nullptr;
        return cudaSuccess;
    }
    int *c_x, *c_y; double *c_p, *c_stat;
    PWCK(S.get(7, sizeof(int) * nf, (void**)&c_x), "alloc");
    PWCK(S.get(8, sizeof(int) * nf, (void**)&c_y), "alloc");
    PWCK(S.get(9, sizeof(double) * nf, (void**)&c_p), "alloc");
    PWCK(S.get(10, sizeof(double) * nf, (void**)&c_stat), "alloc");
    pw_fz_rows_kernel<T, 1><<<(unsigned)p, T, 0, st>>>(d_cor, p, fc, alpha, r_lo, reliable_only ? 1 : 0, suff_all, nullptr, nullptr, row_base, c_x, c_y, c_p, c_stat);
    (*n_launch)++;
    PWCK(cudaGetLastError(), "pw_fz_rows_kernel<1>");
    return pairwise_finish(S, c_x, c_y, c_p, c_stat, nf, m, p, alpha, fdr, st, out, n_launch, msg);
}
