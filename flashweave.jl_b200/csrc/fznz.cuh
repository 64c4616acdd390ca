// fznz.cuh — zero-ignoring Fisher-z tests (test_name "fz_nz", FlashWeaveHE-S).
//
// Replaces (reference paths relative to the FlashWeave.jl checkout):
//   cor_subset!                        src/statfuns.jl:138-155   (called from test_subsets, src/tests.jl:293-308)
//   prepare_nzdata / needs_nz_view     src/hiton.jl:41-50,85, src/misc.jl:103-107
//   univariate fz_nz test              src/tests.jl:108-160 (dense branch :127-147) + pw_univar_kernel :410-423
//
// For fz_nz the correlations are not global: every (X, Y) job uses the Pearson sub-matrix of [X, Y, Z_total...]
// on the rows where X != 0 and Y != 0, and n for fz_pval is the number of those rows.  Here the row view is a
// bit mask (AND of two precomputed non-zero planes), the sub-matrix is recomputed per job into the same
// shared-memory block R that the plain Fisher-z kernels gather from cor_mat, and the tests themselves are the
// unchanged pcor_rec code of fz.cuh.  Moments are accumulated in fp64 in two passes (mean, then centred
// products) like Statistics.cor, and the result is rounded to Float32 (cor_mat's eltype, learning.jl:127-129).
#pragma once
#include "common.cuh"
#include "fz.cuh"

struct NzTable {
    const float* data;            // column-major n x p (one variable per contiguous row of length ld)
    const unsigned int* nzmask;   // [p][W] bit r of word w: row 32w+r is non-zero
    const int* nnz;               // per variable
    i64 p; i64 ld; int n; int W;
};

__global__ void nz_mask_kernel(const float* __restrict__ data, i64 n, i64 ld, i64 p, int W, unsigned int* __restrict__ mask, int* __restrict__ nnz) {
    const i64 gw = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= p * W) return;
    const i64 v = gw / W; const int w = (int)(gw % W);
    const i64 row = (i64)w * 32 + lane;
    const bool nzb = row < n && data[v * ld + row] != 0.0f;
    unsigned int b = __ballot_sync(0xffffffffu, nzb);
    if (lane == 0) { mask[v * W + w] = b; if (b) atomicAdd(&nnz[v], __popc(b)); }
}

__device__ __forceinline__ FzConsts nz_consts(i64 rows, i64 n_obs_min) {
    FzConsts fc;
    const i64 sf = rows - 3;
    fc.sf_pos = sf > 0 ? 1 : 0;
    fc.half_sqrt_sf = sf > 0 ? __ddiv_rn(__dsqrt_rn((double)sf), 2.0) : 0.0;
    fc.rows_ok = rows >= n_obs_min ? 1 : 0;
    return fc;
}

// ---- univariate test, one warp (tests.jl:108-160 on the X-trimmed view of tests.jl:412-416) ----------------
struct NzUni { double stat; double pval; bool suff; };
__device__ NzUni fznz_uni_warp(const NzTable& t, i64 X, i64 Y, i64 n_obs_min) {
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    NzUni r;
    const i64 rows_x = t.nnz[X];
    if (rows_x < n_obs_min) { r.stat = 0.0; r.pval = 1.0; r.suff = (0 >= n_obs_min); return r; }     // tests.jl:111-115,159
    const float* x = t.data + X * t.ld; const float* y = t.data + Y * t.ld;
    const unsigned int* mx = t.nzmask + X * t.W; const unsigned int* my = t.nzmask + Y * t.W;
    int cnt = 0;
    for (int w = lane; w < t.W; w += 32) cnt += __popc(mx[w] & my[w]);
    const i64 n_obs = __reduce_add_sync(full, cnt);
    double p_stat = 0.0;
    if (n_obs > 0 && n_obs >= n_obs_min) {
        double sx = 0.0, sy = 0.0;
        for (int i = lane; i < t.n; i += 32) { float a = x[i], b = y[i]; if (a != 0.0f && b != 0.0f) { sx += (double)a; sy += (double)b; } }
        for (int o = 16; o > 0; o >>= 1) { sx += __shfl_xor_sync(full, sx, o); sy += __shfl_xor_sync(full, sy, o); }
        const double mxv = sx / (double)n_obs, myv = sy / (double)n_obs;
        double sxx = 0.0, syy = 0.0, sxy = 0.0;
        for (int i = lane; i < t.n; i += 32) {
            float a = x[i], b = y[i];
            if (a != 0.0f && b != 0.0f) { double da = (double)a - mxv, db = (double)b - myv; sxx += da * da; syy += db * db; sxy += da * db; }
        }
        for (int o = 16; o > 0; o >>= 1) { sxx += __shfl_xor_sync(full, sxx, o); syy += __shfl_xor_sync(full, syy, o); sxy += __shfl_xor_sync(full, sxy, o); }
        double rr = sxy / (sqrt(sxx) * sqrt(syy));
        if (rr > 1.0) rr = 1.0; else if (rr < -1.0) rr = -1.0;         // clampcor; NaN passes through (tests.jl:143, :381)
        p_stat = (double)(float)rr;                                      // eltype of the data (Float32)
    }
    r.stat = p_stat;
    r.pval = fz_pval_dev(p_stat, nz_consts(n_obs, n_obs_min));
    r.suff = n_obs >= n_obs_min;
    return r;
}

// ---- block-cooperative cor_subset! (statfuns.jl:138-155) on the rows where var[0] != 0 and var[1] != 0 ------
// Fills R[i*ld + j] for all slot pairs i != j < nv (slot -> variable id in var[]), NaN -> 0.  mask: W words of
// shared memory, mom: 2*nv doubles of shared memory.  Returns the number of rows of the view (all threads).
template <int THREADS>
__device__ int fznz_subcor_block(const NzTable& t, const i64* var, int nv, int xs, int ys, float* R, int ld, unsigned int* mask, double* mom, int* s_cnt) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned full = 0xffffffffu;
    const unsigned int* mx = t.nzmask + var[xs] * t.W; const unsigned int* my = t.nzmask + var[ys] * t.W;
    if (tid == 0) *s_cnt = 0;
    __syncthreads();
    int c = 0;
    for (int w = tid; w < t.W; w += THREADS) { unsigned int m = mx[w] & my[w]; mask[w] = m; c += __popc(m); }
    c = __reduce_add_sync(full, c);
    if (lane == 0 && c) atomicAdd(s_cnt, c);
    __syncthreads();
    const int rows = *s_cnt;
    if (rows == 0) {                                    // empty view: every correlation is NaN -> 0 (statfuns.jl:150)
        for (int e = tid; e < nv * nv; e += THREADS) R[(e / nv) * ld + (e % nv)] = 0.0f;
        __syncthreads();
        return 0;
    }
    // means and centred norms: one warp per variable, two passes (Statistics.cor: corm -> covzm)
    for (int a = warp; a < nv; a += THREADS / 32) {
        const float* x = t.data + var[a] * t.ld;
        double s = 0.0;
        for (int i = lane; i < t.n; i += 32) if ((mask[i >> 5] >> (i & 31)) & 1u) s += (double)x[i];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(full, s, o);
        const double mu = s / (double)rows;
        double ss = 0.0;
        for (int i = lane; i < t.n; i += 32) if ((mask[i >> 5] >> (i & 31)) & 1u) { double d = (double)x[i] - mu; ss += d * d; }
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(full, ss, o);
        if (lane == 0) { mom[2 * a] = mu; mom[2 * a + 1] = sqrt(ss); }
    }
    __syncthreads();
    // centred cross products: one warp per slot pair (a < b)
    const int n_pairs = nv * (nv - 1) / 2;
    for (int e = warp; e < n_pairs; e += THREADS / 32) {
        int a, b; unrank2(e, nv, a, b);
        const float* xa = t.data + var[a] * t.ld; const float* xb = t.data + var[b] * t.ld;
        const double ma = mom[2 * a], mb = mom[2 * b];
        double s = 0.0;
        for (int i = lane; i < t.n; i += 32) if ((mask[i >> 5] >> (i & 31)) & 1u) s += ((double)xa[i] - ma) * ((double)xb[i] - mb);
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(full, s, o);
        if (lane == 0) {
            double rr = s / (mom[2 * a + 1] * mom[2 * b + 1]);
            if (rr > 1.0) rr = 1.0; else if (rr < -1.0) rr = -1.0;
            float rf = isnan(rr) ? 0.0f : (float)rr;                    // statfuns.jl:150: NaN -> 0; cor_mat eltype Float32
            R[a * ld + b] = rf; R[b * ld + a] = rf;
        }
    }
    __syncthreads();
    return rows;
}

// ---- pairwise stage: one warp per pair, unordered emission (tests.jl:410-433, :391-407) -----------------------
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) pw_fznz_rows_kernel(NzTable t, i64 n_obs_min, double alpha, int reliable_only,
                                                                  u64* counters, i64 cap, int* c_x, int* c_y, double* c_p, double* c_stat) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const i64 X = blockIdx.x;
    i64 n_rel = 0;
    for (i64 Y = X + 1 + warp; Y < t.p; Y += WARPS) {
        NzUni r = fznz_uni_warp(t, X, Y, n_obs_min);
        const bool rel = (r.suff || !reliable_only) && !isnan(r.pval);
        n_rel += rel;
        if (rel && r.pval < alpha && lane == 0) {
            u64 pos = atomicAdd(&counters[0], 1ull);
            if ((i64)pos < cap) { c_x[pos] = (int)X; c_y[pos] = (int)Y; c_p[pos] = r.pval; c_stat[pos] = r.stat; }
        }
        __syncwarp();
    }
    if (lane == 0 && n_rel) atomicAdd(&counters[1], (u64)n_rel);
}

// ---- independent tests: one CTA per test (fw_test_batch, kind fz_nz) --------------------------------------------
template <int THREADS>
__global__ void __launch_bounds__(THREADS) fznz_test_batch_kernel(NzTable t, i64 n_tests, const i64* X, const i64* Y, const int* k, const i64* Zs,
                                                                  i64 n_obs_min, DevResult* out) {
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned int* mask = reinterpret_cast<unsigned int*>(smem);
    __shared__ float R[25];
    __shared__ double mom[10];
    __shared__ i64 var[5];
    __shared__ int s_cnt;
    for (i64 tix = blockIdx.x; tix < n_tests; tix += gridDim.x) {
        __syncthreads();
        const int kk = k[tix];
        if (kk == 0) {
            if (threadIdx.x < 32) {
                NzUni r = fznz_uni_warp(t, X[tix], Y[tix], n_obs_min);
                if (threadIdx.x == 0) out[tix] = make_result(r.stat, r.pval, 0, r.suff);
            }
            continue;
        }
        if (threadIdx.x == 0) { var[0] = X[tix]; var[1] = Y[tix]; for (int j = 0; j < 3; ++j) var[2 + j] = j < kk ? Zs[tix * 3 + j] : X[tix]; }
        __syncthreads();
        const int rows = fznz_subcor_block<THREADS>(t, var, kk + 2, 0, 1, R, 5, mask, mom, &s_cnt);
        if (threadIdx.x == 0) {
            // tests.jl:250-265 on the (X, Y)-trimmed view: n = rows of the view
            FzConsts fc = nz_consts(rows, n_obs_min);
            CorSlots cs; cs.R = R; cs.ld = 5;
            FzTest ft = fz_cond_test(cs, 0, 1, 2, 3, 4, kk, fc);
            out[tix] = make_result(ft.stat, ft.pval, 0, ft.suff);
        }
    }
}
