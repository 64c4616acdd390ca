// mi_lane.cuh — the conditioning-subset scan of one candidate for binary tables (kind "mi", 2 levels), one LANE per test.
//
// Replaces the inner loops of test_subsets (src/tests.jl:281-346) over test(X, Y, Zs) (src/tests.jl:184-229) with its contingency
// table (src/contingency.jl:7-56) and mutual_information / adjust_df / mi_pval (src/statfuns.jl:157-305) for that case.
//
// Why it is cheap.  A 2 x 2 x 2^k table is the Moebius transform of the POSITIVE-AND counts N(S & T) = |rows where all of S and
// all of T are 1| with S a subset of {X, Y}, T a subset of the conditioning set.  Of the 4 * 2^k positive counts of a test only
// those with T = Zs itself are new: the 4 * (2^k - 1) others belong to smaller conditioning sets, which all subsets of one
// candidate share.  So, per candidate, the counts for |T| <= 2 are tabulated once (P0, P1[member], P2[pair of members]) and a
// k = 3 test costs one pass over the rows with 2 AND + 4 POPC + 3 AND per 32 rows (the tables of the old scan cost 32 POPC and
// 32 warp reductions per 32 rows and test); k = 1 and k = 2 tests are pure table look-ups.  The bit planes of the target and its
// accepted members are staged in shared memory once (row stride odd => the per-lane plane reads are bank-conflict free), every
// lane walks its own subset, and the fp64 statistics (32 logs for k = 3) run per lane on registers - no warp reductions at all.
// Counts are exact integers; the statistics accumulate the same terms in the same order as mi_stats_bin_thread (hiton_mi.cuh), with
// each logarithm taken from a table of log(i) (see mi_lane_stats).
#pragma once
#include "common.cuh"
#include "mi.cuh"
#include "subsets.cuh"

constexpr int MI_LANE_MAX_M = 30;                                 // accepted members the tables are sized for
constexpr int MI_LANE_P2 = MI_LANE_MAX_M * (MI_LANE_MAX_M - 1) / 2;
// shared-memory ints behind the staged planes: P0[4], P1[30][4], P2[435][4]
constexpr int MI_LANE_TAB_INTS = 4 + 4 * MI_LANE_MAX_M + 4 * MI_LANE_P2;

__device__ __forceinline__ int mi_lane_pidx(int a, int b, int m) { return a * m - ((a * (a + 1)) >> 1) + (b - a - 1); }   // a < b, lexicographic

// positive-AND counts of (X = slot xs, Y = slot ys) against the accepted members acc[0..m): all threads of the CTA
template <int THREADS>
__device__ void mi_lane_build(const unsigned int* sp, int Wp, int W, int n_rows, int xs, int ys, const int* acc, int m, int* tabs) {
    const int tid = threadIdx.x;
    const unsigned int* px = sp + xs * Wp;
    const unsigned int* py = sp + ys * Wp;
    int* P0 = tabs; int* P1 = tabs + 4; int* P2 = tabs + 4 + 4 * MI_LANE_MAX_M;
    const int n_items = 1 + m + m * (m - 1) / 2;
    for (int it = tid; it < n_items; it += THREADS) {
        const unsigned int* pa = nullptr; const unsigned int* pb = nullptr; int* dst = P0;
        if (it >= 1 && it <= m) { pa = sp + acc[it - 1] * Wp; dst = P1 + 4 * (it - 1); }
        else if (it > m) { int a, b; unrank2_small(it - 1 - m, m, a, b); pa = sp + acc[a] * Wp; pb = sp + acc[b] * Wp; dst = P2 + 4 * (it - 1 - m); }
        int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
        for (int w = 0; w < W; ++w) {
            unsigned int t = pa ? pa[w] : 0xffffffffu;
            if (pb) t &= pb[w];
            const unsigned int x = px[w], tx = t & x, ty = t & py[w];
            c0 += __popc(t); c1 += __popc(tx); c2 += __popc(ty); c3 += __popc(tx & ty);
        }
        dst[0] = it == 0 ? n_rows : c0; dst[1] = c1; dst[2] = c2; dst[3] = c3;
    }
}

// statistics of one 2 x 2 x S table in registers; cell[4 * s + 2 * b + a] = N(X = a, Y = b, stratum s).  Terms are accumulated in
// the order of mi_stats_bin_thread: the diagonal sum takes the (0,0) cells of all strata, then the (1,1) cells; the off-diagonal sum
// the (0,1) cells, then the (1,0) cells (statfuns.jl:181 loops x level, y level, strata innermost).  Every logarithm of
// statfuns.jl:186-188, log(mk * c / (ma * mb)), has integer arguments <= n: it is evaluated as (lg[mk] - lg[ma]) + (lg[c] - lg[mb])
// from the table lg[i] = log(i) (32 fp64 log calls per k = 3 test were half of this kernel's instructions).  The rearrangement
// changes the statistic by <= ~1e-13 relative (4 table entries of magnitude <= log n, each correctly rounded to 1 ulp).
template <int S>
__device__ __forceinline__ MiResult mi_lane_stats(const int (&cell)[4 * S], i64 hps, const double* __restrict__ lg) {
    MiResult r;
    i64 n_obs = 0; int levels_z = 0;
#pragma unroll
    for (int s = 0; s < S; ++s) { const int tot = cell[4 * s] + cell[4 * s + 1] + cell[4 * s + 2] + cell[4 * s + 3]; n_obs += tot; levels_z += tot > 0; }
    if (!(((double)n_obs / (double)(4 * levels_z)) > (double)hps)) { r.stat = 0.0; r.pval = 1.0; r.df = 0; r.suff = false; return r; }
    double pos = 0.0, neg = 0.0; i64 n_pos = 0, n_neg = 0; int df = 0;
    double lk[S], la1[S], lb0[S], lb1[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const int c00 = cell[4 * s], c10 = cell[4 * s + 1], c01 = cell[4 * s + 2], c11 = cell[4 * s + 3];
        const int ma0 = c00 + c01, ma1 = c10 + c11, mb0 = c00 + c10, mb1 = c01 + c11, mk = ma0 + ma1;
        lk[s] = __ldg(lg + mk); la1[s] = __ldg(lg + ma1); lb0[s] = __ldg(lg + mb0); lb1[s] = __ldg(lg + mb1);
        const double la0 = __ldg(lg + ma0);
        if (c00 != 0) { pos += ((lk[s] - la0) + (__ldg(lg + c00) - lb0[s])) * (double)c00; n_pos += c00; }
        if (c01 != 0) { neg += ((lk[s] - la0) + (__ldg(lg + c01) - lb1[s])) * (double)c01; n_neg += c01; }
        n_pos += c11; n_neg += c10;
        df += (ma0 > 0 && ma1 > 0 && mb0 > 0 && mb1 > 0) ? 1 : 0;
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const int c10 = cell[4 * s + 1], c11 = cell[4 * s + 3];
        if (c11 != 0) pos += ((lk[s] - la1[s]) + (__ldg(lg + c11) - lb1[s])) * (double)c11;
        if (c10 != 0) neg += ((lk[s] - la1[s]) + (__ldg(lg + c10) - lb0[s])) * (double)c10;
    }
    const i64 n_mi = n_pos + n_neg;
    double mi = (pos + neg) / (double)n_mi;
    if (neg * ((double)n_neg / (double)n_mi) > pos * ((double)n_pos / (double)n_mi)) mi *= -1.0;
    r.stat = mi; r.df = df; r.pval = mi_pval_dev(fabs(mi), df, n_obs); r.suff = true;
    return r;
}

// f[S][T] positive-AND counts (T = bit mask over the K conditioning variables) -> cells, in place: superset Moebius over T, then
// over S = {x, y}
template <int K>
__device__ __forceinline__ void mi_lane_cells(int (&f)[4][1 << K], int (&cell)[4 << K]) {
    constexpr int S = 1 << K;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int bit = 0; bit < K; ++bit) {
#pragma unroll
            for (int t = 0; t < S; ++t) if (!(t & (1 << bit))) f[q][t] -= f[q][t | (1 << bit)];
        }
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const int g0 = f[0][s], gx = f[1][s], gy = f[2][s], gxy = f[3][s];
        cell[4 * s + 3] = gxy; cell[4 * s + 1] = gx - gxy; cell[4 * s + 2] = gy - gxy; cell[4 * s] = g0 - gx - gy + gxy;
    }
}

// one test of the scan: positions (a < b < c) in the accepted list, k of them used
__device__ MiResult mi_lane_test(const unsigned int* sp, int Wp, int W, int xs, int ys, const int* acc, int m, const int* tabs,
                                 int k, int a, int b, int c, i64 hps, const double* __restrict__ lg) {
    const int* P0 = tabs; const int* P1 = tabs + 4; const int* P2 = tabs + 4 + 4 * MI_LANE_MAX_M;
    if (k == 3) {
        const unsigned int* pa = sp + acc[a] * Wp; const unsigned int* pb = sp + acc[b] * Wp; const unsigned int* pc = sp + acc[c] * Wp;
        const unsigned int* px = sp + xs * Wp; const unsigned int* py = sp + ys * Wp;
        int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma unroll 4
        for (int w = 0; w < W; ++w) {
            const unsigned int t = pa[w] & pb[w] & pc[w];
            const unsigned int tx = t & px[w], ty = t & py[w];
            c0 += __popc(t); c1 += __popc(tx); c2 += __popc(ty); c3 += __popc(tx & ty);
        }
        const int ab = mi_lane_pidx(a, b, m), ac = mi_lane_pidx(a, c, m), bc = mi_lane_pidx(b, c, m);
        int f[4][8], cell[32];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            f[q][0] = P0[q]; f[q][1] = P1[4 * a + q]; f[q][2] = P1[4 * b + q]; f[q][4] = P1[4 * c + q];
            f[q][3] = P2[4 * ab + q]; f[q][5] = P2[4 * ac + q]; f[q][6] = P2[4 * bc + q];
        }
        f[0][7] = c0; f[1][7] = c1; f[2][7] = c2; f[3][7] = c3;
        mi_lane_cells<3>(f, cell);
        return mi_lane_stats<8>(cell, hps, lg);
    }
    if (k == 2) {
        const int ab = mi_lane_pidx(a, b, m);
        int f[4][4], cell[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) { f[q][0] = P0[q]; f[q][1] = P1[4 * a + q]; f[q][2] = P1[4 * b + q]; f[q][3] = P2[4 * ab + q]; }
        mi_lane_cells<2>(f, cell);
        return mi_lane_stats<4>(cell, hps, lg);
    }
    int f[4][2], cell[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) { f[q][0] = P0[q]; f[q][1] = P1[4 * a + q]; }
    mi_lane_cells<1>(f, cell);
    return mi_lane_stats<2>(cell, hps, lg);
}

// Same contract as eval_subsets (subsets.cuh) / eval_subsets_mi_bin: chunks of THREADS consecutive subsets of the reference's
// enumeration order, one per thread; the first failing index and the arg-max are recovered with block reductions.  The planes of
// slots xs, ys and acc[] must be staged in sp; tabs: MI_LANE_TAB_INTS ints of shared memory.  m <= MI_LANE_MAX_M.
template <int THREADS>
__device__ void eval_subsets_mi_lane(const unsigned int* sp, int Wp, int W, int n_rows, int xs, int ys, const int* acc, int m, int max_k, double alpha,
                                     i64 max_tests, i64 hps, i64* tri_off, int* tabs, EvalShared* sh, EvalOut* out, const double* __restrict__ lg) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = THREADS / 32;
    const unsigned full = 0xffffffffu;
    const SubsetCounts sc = subset_counts(m, max_k);
    if (sc.c3 > 0) for (int i = tid; i <= m; i += THREADS) tri_off[i] = sc.c3 - choose3(m - i);
    if (tid == 0) sh->fail_idx = (u64)FW_INF_IDX;
    mi_lane_build<THREADS>(sp, Wp, W, n_rows, xs, ys, acc, m, tabs);
    __syncthreads();
    const i64 limit = (max_tests > 0 && max_tests < sc.total) ? max_tests : sc.total;
    i64 my_fail = FW_INF_IDX; double f_stat = 0.0, f_p = 0.0; i64 f_df = 0; int f_suff = 0;
    i64 best_idx = -1; double b_stat = 0.0, b_p = -1.0; i64 b_df = 0;
    i64 executed = 0;
    bool any_fail = false;
    for (i64 base = 0; base < limit; base += THREADS) {
        const i64 idx = base + tid;
        if (idx < limit) {
            int k, a, b, c;
            unrank_subset(idx, m, sc, tri_off, k, a, b, c);
            const MiResult r = mi_lane_test(sp, Wp, W, xs, ys, acc, m, tabs, k, a, b, c, hps, lg);
            const bool sig = (r.pval < alpha) && r.suff;
            const bool stop = !sig || (max_tests > 0 && idx + 1 >= max_tests);
            if (stop) { if (idx < my_fail) { my_fail = idx; f_stat = r.stat; f_p = r.pval; f_df = r.df; f_suff = r.suff ? 1 : 0; } }
            else if (r.pval >= b_p) { best_idx = idx; b_stat = r.stat; b_p = r.pval; b_df = r.df; }
        }
        const i64 end = base + THREADS;
        executed = end < limit ? end : limit;
        any_fail = __syncthreads_or(my_fail != FW_INF_IDX);
        if (any_fail) break;
    }
    if (any_fail) {
        if (my_fail != FW_INF_IDX) atomicMin(&sh->fail_idx, (u64)my_fail);
        __syncthreads();
        if ((u64)my_fail == sh->fail_idx) {
            int k, a, b, c;
            unrank_subset(my_fail, m, sc, tri_off, k, a, b, c);
            out->stat = f_stat; out->pval = f_p; out->df = f_df; out->suff = f_suff;
            out->sig = ((f_p < alpha) && f_suff) ? 1 : 0;
            out->k = k; out->pos[0] = a; out->pos[1] = b; out->pos[2] = c;
            out->num_tests = my_fail + 1; out->total = sc.total; out->executed = executed; executed_by_k(sc, executed, out->ex_k);
        }
        __syncthreads();
        return;
    }
    // all significant: arg-max p-value, ties -> larger index (tests.jl:338-341)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double op = __shfl_down_sync(full, b_p, off);
        double os = __shfl_down_sync(full, b_stat, off);
        i64 oi = __shfl_down_sync(full, best_idx, off);
        i64 od = __shfl_down_sync(full, b_df, off);
        if (op > b_p || (op == b_p && oi > best_idx)) { b_p = op; b_stat = os; best_idx = oi; b_df = od; }
    }
    if (lane == 0) { sh->w_p[warp] = b_p; sh->w_stat[warp] = b_stat; sh->w_idx[warp] = best_idx; sh->w_df[warp] = b_df; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < NW; ++w) {
            double op = sh->w_p[w]; i64 oi = sh->w_idx[w];
            if (op > b_p || (op == b_p && oi > best_idx)) { b_p = op; b_stat = sh->w_stat[w]; best_idx = oi; b_df = sh->w_df[w]; }
        }
        int k = 0, a = 0, b = 0, c = 0;
        if (best_idx >= 0) unrank_subset(best_idx, m, sc, tri_off, k, a, b, c);
        out->stat = b_stat; out->pval = b_p; out->df = b_df; out->suff = 1;
        out->sig = (b_p < alpha) ? 1 : 0;
        out->k = k; out->pos[0] = a; out->pos[1] = b; out->pos[2] = c;
        out->num_tests = limit; out->total = sc.total; out->executed = executed; executed_by_k(sc, executed, out->ex_k);
    }
    __syncthreads();
}
