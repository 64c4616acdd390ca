// subsets.cuh — block-cooperative restatement of test_subsets (src/tests.jl:281-346).
//
// The reference walks the conditioning subsets of Z_total sequentially: sizes max_k..1,
// lexicographic inside a size, stops at the first non-significant test (or at max_tests),
// otherwise keeps the maximum-p-value result (ties -> later subset).  Here a CTA
// evaluates the same index space in ascending chunks of THREADS*TPT tests; the first
// failing index and the arg-max are recovered with block reductions, so the returned
// (result, Zs, num_tests) is exactly the reference's.  Tests past the first failure inside
// a chunk are speculative work (counted in `executed`, never in `num_tests`).
#pragma once
#include "common.cuh"

struct EvalOut {
    double stat, pval;
    i64 df;
    int suff;
    int sig;            // issig(result, alpha)  (tests.jl:1-3)
    int k;              // size of the returned subset (0: none)
    int pos[3];         // positions of the returned subset inside Z_total
    i64 num_tests;      // tests.jl:322 counter at return
    i64 total;          // num_tests_total (tests.jl:310-333)
    i64 executed;       // tests actually evaluated on the device
    i64 ex_k[3];        // ... of which with |Zs| = 1, 2, 3
};

// Per-target whitelists / blacklists of si_HITON_PC (src/hiton.jl:20-38: a whitelisted candidate is accepted untested with
// (NaN, NaN) - and, in the elimination phase, pushed a SECOND time onto `accepted`; a blacklisted one is skipped) and the optional
// rejection records of track_rejections (src/hiton.jl:72-74: candidate -> (Zs, TestResult, (num_tests, frac))).  Lists are CSR
// over the launch's target list; rejection slots share the targets' output ranges (a target rejects at most its candidate count).
struct HitonLists {
    const i64* wl_off; const i64* wl_idx;
    const i64* bl_off; const i64* bl_idx;
    i64* rej_count; i64* rej_nbr; i64* rej_Zs; int* rej_k; DevResult* rej_res; i64* rej_ntests; double* rej_frac;
};
// bit 0: candidate is whitelisted, bit 1: blacklisted
__device__ __forceinline__ int hiton_list_flags(const HitonLists& L, int tsel, i64 cand) {
    int f = 0;
    if (L.wl_off) for (i64 i = L.wl_off[tsel]; i < L.wl_off[tsel + 1]; ++i) if (L.wl_idx[i] == cand) { f |= 1; break; }
    if (L.bl_off) for (i64 i = L.bl_off[tsel]; i < L.bl_off[tsel + 1]; ++i) if (L.bl_idx[i] == cand) { f |= 2; break; }
    return f;
}
constexpr int HITON_ORDER_MASK = 0x0fffffff;            // cand_order[rank] = position in the univariate list | list flags << 28

struct EvalShared {
    u64 fail_idx;
    double w_p[32];
    double w_stat[32];
    i64 w_idx[32];
    i64 w_df[32];
};

struct SubsetCounts { i64 c3, c2, c1, total; };
__device__ __forceinline__ void executed_by_k(const SubsetCounts& sc, i64 executed, i64* ex_k) {
    i64 e3 = executed < sc.c3 ? executed : sc.c3;
    i64 r = executed - e3;
    i64 e2 = r < sc.c2 ? r : sc.c2;
    ex_k[2] = e3; ex_k[1] = e2; ex_k[0] = r - e2;
}
__device__ __forceinline__ SubsetCounts subset_counts(int m, int max_k) {
    SubsetCounts s;
    s.c3 = (max_k >= 3) ? choose3(m) : 0;
    s.c2 = (max_k >= 2) ? choose2(m) : 0;
    s.c1 = (max_k >= 1) ? (i64)m : 0;
    s.total = s.c3 + s.c2 + s.c1;
    return s;
}

// index -> (k, positions) in the reference's enumeration order.  tri_off[i] = number of
// triples whose first position is < i (shared-memory table, m+1 entries).
// 32-bit variant, valid when sc.total < 2^31 and m <= 4096
__device__ __forceinline__ void unrank_subset32(int idx, int m, int c3, int c2, const i64* tri_off, int& k, int& a, int& b, int& c) {
    if (idx < c3) {
        int lo = 0, hi = m - 3;
        while (lo < hi) { int mid = (lo + hi + 1) >> 1; if ((int)tri_off[mid] <= idx) lo = mid; else hi = mid - 1; }
        int pa, pb;
        unrank2_small(idx - (int)tri_off[lo], m - lo - 1, pa, pb);
        k = 3; a = lo; b = lo + 1 + pa; c = lo + 1 + pb;
    } else if (idx < c3 + c2) {
        unrank2_small(idx - c3, m, a, b); k = 2; c = 0;
    } else {
        k = 1; a = idx - c3 - c2; b = 0; c = 0;
    }
}
__device__ __forceinline__ void unrank_subset(i64 idx, int m, const SubsetCounts& sc, const i64* tri_off, int& k, int& a, int& b, int& c) {
    if (idx < sc.c3) {
        int lo = 0, hi = m - 3;              // largest i with tri_off[i] <= idx
        while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (tri_off[mid] <= idx) lo = mid; else hi = mid - 1; }
        int i = lo, pa, pb;
        unrank2(idx - tri_off[i], m - i - 1, pa, pb);
        k = 3; a = i; b = i + 1 + pa; c = i + 1 + pb;
    } else if (idx < sc.c3 + sc.c2) {
        unrank2(idx - sc.c3, m, a, b); k = 2; c = 0;
    } else {
        k = 1; a = (int)(idx - sc.c3 - sc.c2); b = 0; c = 0;
    }
}

// TestFn: {stat,pval,suff,df} operator()(int k, int za, int zb, int zc) with z* = slots.
// GROUP = threads that evaluate one test together (1: thread-per-test, Fisher-z; 32: warp-per-test, discrete);
// the functor is called by all GROUP lanes and must return the same result in each.
// All threads of the CTA must call this with identical arguments.  `out` lives in shared memory.
template <int THREADS, int TPT, int GROUP, class TestFn>
__device__ void eval_subsets(const TestFn& test, const int* acc, int m, int max_k, double alpha, i64 max_tests,
                             i64* tri_off, EvalShared* sh, EvalOut* out) {
    const int tid = threadIdx.x;
    constexpr int NG = THREADS / GROUP;              // tests in flight per pass
    const int grp = tid / GROUP;
    const bool leader = (tid % GROUP) == 0;
    const SubsetCounts sc = subset_counts(m, max_k);
    if (sc.c3 > 0) for (int i = tid; i <= m; i += THREADS) tri_off[i] = sc.c3 - choose3(m - i);
    if (tid == 0) sh->fail_idx = (u64)FW_INF_IDX;
    __syncthreads();
    const i64 limit = (max_tests > 0 && max_tests < sc.total) ? max_tests : sc.total;

    i64 my_fail = FW_INF_IDX; double f_stat = 0.0, f_p = 0.0; i64 f_df = 0; int f_suff = 0;
    i64 best_idx = -1; double b_stat = 0.0, b_p = -1.0; i64 b_df = 0;
    i64 executed = 0;
    bool any_fail = false;
    const bool small = sc.total < ((i64)1 << 30) && m <= 4096;     // 32-bit index arithmetic in the hot loop
    const int c3s = (int)(small ? sc.c3 : 0), c2s = (int)(small ? sc.c2 : 0);
    // chunk schedule: a first small chunk (NG tests) catches the early exits of the reference cheaply; later chunks are
    // NG*TPT tests, so an all-significant scan pays few barriers
    int tpt = 1;
    for (i64 base = 0; base < limit;) {
        for (int u = 0; u < tpt; ++u) {
            i64 idx = base + (i64)u * NG + grp;
            if (idx < limit && my_fail == FW_INF_IDX) {
                int k, a, b, c;
                if (small) unrank_subset32((int)idx, m, c3s, c2s, tri_off, k, a, b, c);
                else unrank_subset(idx, m, sc, tri_off, k, a, b, c);
                auto r = test(k, acc[a], acc[b], acc[c]);
                bool sig = (r.pval < alpha) && r.suff;
                bool stop = !sig || (max_tests > 0 && idx + 1 >= max_tests);
                if (stop) { my_fail = idx; f_stat = r.stat; f_p = r.pval; f_df = r.df; f_suff = r.suff ? 1 : 0; }
                else if (r.pval >= b_p) { best_idx = idx; b_stat = r.stat; b_p = r.pval; b_df = r.df; }
            }
        }
        i64 end = base + (i64)NG * tpt;
        executed = end < limit ? end : limit;
        any_fail = __syncthreads_or(my_fail != FW_INF_IDX);
        if (any_fail) break;
        base = end; tpt = TPT;
    }
    if (any_fail) {
        if (my_fail != FW_INF_IDX && leader) atomicMin(&sh->fail_idx, (u64)my_fail);
        __syncthreads();
        if ((u64)my_fail == sh->fail_idx && leader) {
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
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double op = __shfl_down_sync(full, b_p, off);
        double os = __shfl_down_sync(full, b_stat, off);
        i64 oi = __shfl_down_sync(full, best_idx, off);
        i64 od = __shfl_down_sync(full, b_df, off);
        if (op > b_p || (op == b_p && oi > best_idx)) { b_p = op; b_stat = os; best_idx = oi; b_df = od; }
    }
    const int warp = tid >> 5, lane = tid & 31;
    if (lane == 0) { sh->w_p[warp] = b_p; sh->w_stat[warp] = b_stat; sh->w_idx[warp] = best_idx; sh->w_df[warp] = b_df; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < THREADS / 32; ++w) {
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
