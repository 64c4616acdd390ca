// hiton.cuh — device-resident si_HITON_PC (src/hiton.jl:283-400) for Fisher-z tests, and the
// batched test_subsets job kernel (src/tests.jl:281-346).
//
// One CTA owns one target variable T and runs the whole interleaving + elimination loop
// (hiton.jl:109-149) without leaving the SM: the correlations among {T, accepted members,
// current candidate} live in a shared-memory sub-block R that grows by one gathered row per
// candidate (the reference re-reads cor_mat through pcor_rec for every test instead), the
// conditioning subsets of a candidate are evaluated by all threads (subsets.cuh), and the
// accept/reject bookkeeping of update_sig_result! (hiton.jl:53-78) is done by thread 0
// between barriers.  Targets are pulled from a global atomic queue, so the grid is a
// multiple of the SM count regardless of how many targets there are.
#pragma once
#include "common.cuh"
#include "fz.cuh"
#include "subsets.cuh"
#include "fznz.cuh"

struct HitonArgs {
    // resident inputs
    CorView cv; i64 p;                       // cor_mat (p x p, symmetric; row-sharded over the GPUs of the group)
    const i64* uni_off; const i64* uni_nbr;  // univariate neighbour CSR over all p variables
    const double* uni_stat; const double* uni_p;
    // work list
    const i64* targets;                      // 0-based variable ids, indexed by tsel
    const int* sel; int n_sel;               // which entries of targets[] this launch handles
    const i64* out_off;                      // per target: offset of its output/scratch slots
    int* counter;                            // atomic work queue
    // parameters
    int max_k; double alpha; i64 max_tests; FzConsts fc;
    int cap;                                 // slot capacity of R (cap x cap floats)
    float* gscratch;                         // when non-null: R lives here (cap*cap floats per CTA)
    // scratch / outputs (indexed by out_off[tsel] + i)
    int* cand_order;
    i64* pc_nbr; double* pc_stat; double* pc_p; i64* pc_count;
    i64* tpc_nbr; double* tpc_stat; double* tpc_p; i64* tpc_count;
    i64* num_tests; u64* executed_total;      // executed_total[0] = all, [1..3] = by |Zs|
    int* status;                             // per target: 0 ok, 1 capacity overflow (re-run with larger cap)
    // fz_nz only: the table and its non-zero planes; correlations are recomputed per (T, candidate) view
    NzTable nzt; i64 n_obs_min;
    HitonLists lists;                        // whitelists / blacklists / rejection records (all optional)
};

struct FzSlotTest {
    CorSlots r; int x, y; FzConsts fc;
    __device__ __forceinline__ FzTest operator()(int k, int za, int zb, int zc) const {
        return fz_cond_test(r, x, y, za, zb, zc, k, fc);
    }
};

// ---- per-candidate tables (capacity class 32) ------------------------------------------------------------------------------
// Every conditioning subset (Z1, Z2, Z3) of one candidate shares its level-1 terms with all subsets that start with the same
// Z1, and its (X,Y|Z1,Z2) level-2 term with all subsets that start with the same (Z1, Z2).  They are built once per candidate
// in shared memory, which leaves 1 level-1 + 2 level-2 + 1 level-3 step per k = 3 test (the recursion has 6 + 3 + 1) with
// bit-identical arithmetic (same operations in the same order, just not repeated).  A Float64 literal (special) anywhere in
// the tables disables them for that candidate, so the generic typed path keeps handling the degenerate cases.
//
// The tables are indexed by POSITION in the accepted list (Z1 < Z2 < Z3 by position, the reference's enumeration), only the
// upper triangles are needed, and two float triangles share one 32 x 33 square (row stride 33: a warp whose lanes differ in
// the first position reads stride-33 words, lanes that differ in the second read consecutive words - both conflict-free):
//   S1: upper RP[a][b] = r(Za, Zb)                  lower SQ[a][b] = sqrt(1f0 - r(Za,Zb)^2)       (stored at [b][a])
//   S2: upper BX[a][b] = pcor(X, Zb | Za) (Float32) lower BY[a][b] = pcor(Y, Zb | Za)
//   S3: upper SBX[a][b] = sqrt(1f0 - BX^2)          lower SBY[a][b]
//   A2[a][b] = pcor(X, Y | Za, Zb) (Float64), A1[a] = pcor(X, Y | Za) (Float32)
// CLX: idx -> (a, b, c) of the idx-th triple in COLEXICOGRAPHIC order (idx = C(c,3) + C(b,2) + a), 5 bits each.  That order
// does not depend on m, so one table serves every candidate of every target; consecutive lanes get consecutive a and the
// same (b, c), which is what makes the reads above broadcast or conflict-free.
constexpr int FZ_TLD = 33;
constexpr int FZ_CACHE_CAP = 32;                       // slots; at most 30 accepted members
constexpr int FZ_CLX_N = 30 * 29 * 28 / 6;             // 4060 triples
constexpr int FZ_PLX_N = 30 * 29 / 2;                  // 435 pairs, colex as well: idx = C(b,2) + a
struct FzTab { int S1, S2, S3, A2, A1, CLX, PLX, BND, end; };    // byte offsets into dynamic shared memory
__host__ __device__ inline FzTab fz_tab_layout(int o) {
    FzTab t;
    o = (o + 15) & ~15;
    t.A2 = o; o += 8 * FZ_CACHE_CAP * FZ_TLD;
    t.S1 = o; o += 4 * FZ_CACHE_CAP * FZ_TLD;
    t.S2 = o; o += 4 * FZ_CACHE_CAP * FZ_TLD;
    t.S3 = o; o += 4 * FZ_CACHE_CAP * FZ_TLD;
    t.A1 = o; o += 4 * FZ_CACHE_CAP;
    t.BND = o; o += 16;                                 // u64: bits of the smallest |stat| seen so far in the current scan
    t.CLX = o; o += 2 * ((FZ_CLX_N + 7) & ~7);
    t.PLX = o; o += 2 * ((FZ_PLX_N + 7) & ~7);
    t.end = o;
    return t;
}

// CTA-shared state of one scan: first failing index, and the queue of exact maximum-p candidates (see eval_subsets_fz_cached)
struct FzScanShared { u64 fail_idx; int n_cont; int c_idx[128]; double c_stat[128]; };    // THREADS <= 128
// The accepted list of a scan without materialising it: a run of consecutive slots followed by a stored tail.  Interleaving
// phase: slots 1..M (no tail).  Elimination phase of candidate slot c: the untested slots c+1..M, then the candidates accepted
// so far in acceptance order (hiton.jl:134-147: the candidate is deleted from `accepted` and pushed back when it survives).
// With whitelists the processed whitelisted slots are never deleted from `accepted` (hiton.jl:124-131 skips check_candidate!), so
// they precede the untested run (prefix list `pre`) and appear once more among the pushed entries of the tail.
struct AccView {
    int n_head, head_base; const int* tail;
    int n_pre; const int* pre;
    __device__ __forceinline__ int operator[](int j) const {
        if (j < n_pre) return pre[j];
        j -= n_pre;
        return j < n_head ? head_base + j : tail[j - n_head];
    }
};

template <int THREADS>
__device__ void fz_build_colex(const FzTab tab) {
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned short* clx = reinterpret_cast<unsigned short*>(smem + tab.CLX);
    for (int idx = threadIdx.x; idx < FZ_CLX_N; idx += THREADS) {
        int c = 2; while ((c + 1) * c * (c - 1) / 6 <= idx) ++c;          // largest c with C(c,3) <= idx
        int rem = idx - c * (c - 1) * (c - 2) / 6;
        int b = 1; while ((b + 1) * b / 2 <= rem) ++b;                    // largest b with C(b,2) <= rem
        int a = rem - b * (b - 1) / 2;
        clx[idx] = (unsigned short)(a | (b << 5) | (c << 10));
    }
    unsigned short* plx = reinterpret_cast<unsigned short*>(smem + tab.PLX);
    for (int idx = threadIdx.x; idx < FZ_PLX_N; idx += THREADS) {
        int b = 1; while ((b + 1) * b / 2 <= idx) ++b;
        plx[idx] = (unsigned short)((idx - b * (b - 1) / 2) | (b << 5));
    }
}

// acc[0..m) = slots of the accepted members in list order; xs / ys = slots of X (the target) and Y (the candidate)
template <int THREADS>
__device__ bool fz_build_tables(const float* R, int ld, int xs, int ys, const AccView acc, int m, const FzTab tab, i64* tri_off, FzScanShared* sh) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* S1 = reinterpret_cast<float*>(smem + tab.S1);
    float* S2 = reinterpret_cast<float*>(smem + tab.S2);
    float* S3 = reinterpret_cast<float*>(smem + tab.S3);
    double* A2 = reinterpret_cast<double*>(smem + tab.A2);
    float* A1 = reinterpret_cast<float*>(smem + tab.A1);
    const unsigned short* plx = reinterpret_cast<const unsigned short*>(smem + tab.PLX);
    const int tid = threadIdx.x;
    const int npairs = m * (m - 1) / 2;
    if (tid == 0) { *reinterpret_cast<u64*>(smem + tab.BND) = (u64)__double_as_longlong(1e300); sh->fail_idx = (u64)FW_INF_IDX; sh->n_cont = 0; }
    {   // scan state of eval_subsets_fz_cached, set up here so that it costs no barrier of its own
        const int c3 = m * (m - 1) * (m - 2) / 6;
        for (int i = tid; i <= m; i += THREADS) tri_off[i] = c3 - choose3(m - i);
    }
    bool ok = true;
    const float rxy = R[xs * ld + ys];
    for (int e = tid; e < npairs; e += THREADS) {
        const int ia = plx[e] & 31, ib = plx[e] >> 5;
        const int z = acc[ia], sl = acc[ib];
        const float rsz = R[z * ld + sl], ssz = sq1mf(rsz);
        S1[ia * FZ_TLD + ib] = rsz; S1[ib * FZ_TLD + ia] = ssz;
        const float rxz = R[xs * ld + z], ryz = R[ys * ld + z];
        const float sxz = sq1mf(rxz), syz = sq1mf(ryz);
        float a1, bx, by;
        ok &= p1f(rxy, rxz, ryz, sxz, syz, a1);                              // pcor(X, Y | z)   (every thread of row ia computes the same value)
        ok &= p1f(R[xs * ld + sl], rxz, rsz, sxz, ssz, bx);                  // pcor(X, s | z)
        ok &= p1f(R[ys * ld + sl], ryz, rsz, syz, ssz, by);                  // pcor(Y, s | z)
        const float sbx = sq1mf(bx), sby = sq1mf(by);
        S2[ia * FZ_TLD + ib] = bx; S2[ib * FZ_TLD + ia] = by;
        S3[ia * FZ_TLD + ib] = sbx; S3[ib * FZ_TLD + ia] = sby;
        if (ib == ia + 1) A1[ia] = a1;
        // pcor(X, Y | z, s) = p2f(A1[z], bx, by)
        A2[ia * FZ_TLD + ib] = p2f_pre(a1, bx, by, sbx, __dsqrt_rn(__dsub_rn(1.0, __dmul_rn((double)by, (double)by))));
    }
    if (tid == 0) {                                                          // the last position starts no pair
        const int z = acc[m - 1];
        const float rxz = R[xs * ld + z], ryz = R[ys * ld + z];
        float a1;
        ok &= p1f(rxy, rxz, ryz, sq1mf(rxz), sq1mf(ryz), a1);
        A1[m - 1] = a1;
    }
    return __syncthreads_or(!ok) == 0;
}

// ---- p-value-free scan of one candidate's conditioning subsets (capacity class 32, tables built) --------------------------
// Same contract as eval_subsets (subsets.cuh) for max_k = 3, m >= 3 and a non-binding max_tests; the returned
// (result, Zs, num_tests) is exactly the reference's:
//  * schedule: the first THREADS subsets of the reference order (size 3 first, lexicographic) are evaluated as one chunk -
//    this is where the reference's early exit almost always happens (the chunk is skipped when the previous candidate of the
//    same target had no early exit: speculation only changes `executed`, never the result); if all of them are significant, every remaining subset is
//    evaluated in one pass (triples in colex order through CLX, then pairs and singles from the tables), each carrying its
//    reference index, and the first failing index / the arg-max are recovered by block reductions as in eval_subsets;
//  * significance is decided on |stat| against the band [s_lo, s_hi] around the alpha threshold (the p-value is a decreasing
//    function of |stat|; inside the band, and for NaN, the exact p-value decides);
//  * the running maximum p-value (ties -> later subset, tests.jl:338-341) is tracked as the minimum |stat|; whenever two values
//    are closer than 1e-8 relative, or both p-values are below 1e-290, the exact p-values are compared instead.  Exact
//    p-values (fp64 log + erfc) are therefore evaluated a handful of times per candidate instead of once per test.
struct FzScanState {
    int my_fail; double f_stat;
    int best_idx; double best_abs, b_stat, b_p; bool bp_valid;
};
// slow side of the scan (a handful of calls per candidate).  *bound (shared memory) holds the smallest |stat| any thread of the
// CTA has accepted so far in this scan: a later |stat| more than TOL above it (and >= s_hi) is significant and cannot be the
// maximum p-value, so the caller skips it without any bookkeeping.
__device__ __noinline__ void fz_scan_consider(FzScanState& st, int idx, double s, double alpha, const FzConsts& fc, u64* bound) {
    constexpr double TOL = 1e-8;
    const double t = fabs(s);
    bool sig;
    double p_new = -1.0;
    if (t >= fc.s_hi) sig = true;
    else if (t <= fc.s_lo) sig = false;
    else { p_new = fz_pval_dev(s, fc); sig = p_new < alpha; }
    if (!sig) { if (idx < st.my_fail) { st.my_fail = idx; st.f_stat = s; } return; }
    bool take;
    if (st.best_idx < 0) take = true;
    else if (t < st.best_abs * (1.0 - TOL) && t < fc.s_under) take = true;                 // strictly larger p-value
    else if (t > st.best_abs * (1.0 + TOL) && st.best_abs < fc.s_under) take = false;      // strictly smaller p-value
    else {
        // near tie (or both p-values in the underflow range): exact comparison, ties -> later subset
        if (p_new < 0.0) p_new = fz_pval_dev(s, fc);
        if (!st.bp_valid) { st.b_p = fz_pval_dev(st.b_stat, fc); st.bp_valid = true; }
        take = p_new > st.b_p || (p_new == st.b_p && idx > st.best_idx);
    }
    if (!take) return;
    st.best_idx = idx; st.best_abs = t; st.b_stat = s; st.b_p = p_new; st.bp_valid = p_new >= 0.0;
    if (t < fc.s_under) atomicMin(bound, (u64)__double_as_longlong(t));       // non-negative doubles order like their bit patterns
}

// out->pval = -1.0 means "deferred": the maximum was unique, so no p-value had to be evaluated to find it; the caller evaluates
// fz_pval_dev(out->stat, fc) when it needs the number.  NEED_POS: also report the positions of the returned subset.
template <int THREADS, bool NEED_POS>
__device__ void eval_subsets_fz_cached(const CorSlots r, const FzTab tab, int xs, int ys, const AccView acc, int m,
                                       double alpha, const FzConsts fc, i64* tri_off, FzScanShared* sh, EvalOut* out, bool skip0) {
    extern __shared__ __align__(16) unsigned char smem[];
    const float* S1 = reinterpret_cast<const float*>(smem + tab.S1);
    const float* S2 = reinterpret_cast<const float*>(smem + tab.S2);
    const float* S3 = reinterpret_cast<const float*>(smem + tab.S3);
    const double* A2 = reinterpret_cast<const double*>(smem + tab.A2);
    const float* A1 = reinterpret_cast<const float*>(smem + tab.A1);
    const unsigned short* clx = reinterpret_cast<const unsigned short*>(smem + tab.CLX);
    const unsigned short* plx = reinterpret_cast<const unsigned short*>(smem + tab.PLX);
    const int tid = threadIdx.x;
    const int c3 = m * (m - 1) * (m - 2) / 6, c2 = m * (m - 1) / 2, total = c3 + c2 + m;
    constexpr int NOFAIL = 0x7fffffff;
    FzScanState st;
    st.my_fail = NOFAIL; st.f_stat = 0.0; st.best_idx = -1; st.best_abs = 0.0; st.b_stat = 0.0; st.b_p = -1.0; st.bp_valid = false;
    u64* bound = reinterpret_cast<u64*>(smem + tab.BND);       // reset by fz_build_tables

    // one k = 3 test from the tables; (a < b < c) are positions in the accepted list
    auto triple = [&](int a, int b, int c) -> double {
        const int ab = a * FZ_TLD + b, ac = a * FZ_TLD + c, bc = b * FZ_TLD + c;
        const int ba = b * FZ_TLD + a, ca = c * FZ_TLD + a;
        float z3z2;
        if (p1f(S1[bc], S1[ac], S1[ab], S1[ca], S1[ba], z3z2)) {                           // pcor(Z3, Z2 | Z1)
            const double sc = __dsqrt_rn(__dsub_rn(1.0, __dmul_rn((double)z3z2, (double)z3z2)));
            const double B = p2f_pre(S2[ac], S2[ab], z3z2, S3[ab], sc);                    // pcor(X, Z3 | Z1, Z2)
            const double C = p2f_pre(S2[ca], S2[ba], z3z2, S3[ba], sc);                    // pcor(Y, Z3 | Z1, Z2)
            return p3d(A2[ab], B, C);
        }
        return pcor_generic(r, xs, ys, acc[a], acc[b], acc[c], 3);
    };
    auto consider = [&](int idx, double s) {
        const double t = fabs(s);
        const double bnd = __longlong_as_double((i64)*reinterpret_cast<volatile u64*>(bound));
        if (t >= fc.s_hi && t > bnd * (1.0 + 1e-8)) return;                                // significant, cannot be the maximum p
        fz_scan_consider(st, idx, s, alpha, fc, bound);
    };

    // ---- chunk 0: reference indices [0, THREADS); skipped (skip0, CTA-uniform) when the caller expects no early exit ----
    const int n0 = skip0 ? 0 : (total < THREADS ? total : THREADS);
    int executed = n0;
    bool any_fail = false;
    if (!skip0) {
        if (tid < n0) {
            int k, a, b, c;
            unrank_subset32(tid, m, c3, c2, tri_off, k, a, b, c);
            const double s = k == 3 ? triple(a, b, c) : (k == 2 ? A2[a * FZ_TLD + b] : (double)A1[a]);
            consider(tid, s);
        }
        any_fail = __syncthreads_or(st.my_fail != NOFAIL);
    }
    if (!any_fail && total > n0) {
        // ---- everything else in one pass ----
        // two independent tests per thread and iteration, evaluated by the straight-line fz_k3_straight (fz.cuh): without the
        // slow-path branches of the division / square-root intrinsics the compiler overlaps the two dependent chains
        auto decode = [&](int q, int& a, int& b, int& c) {
            const unsigned int e = clx[q];
            a = e & 31; b = (e >> 5) & 31; c = e >> 10;
        };
        // reference (lexicographic) index: triples that start before a, pairs of the suffix that start before b, then c.  Needed only
        // for a test that reaches the bookkeeping (or may belong to chunk 0): computed lazily
        auto ref_index = [&](int a, int b, int c) {
            const int na = m - a - 1, pa = b - a - 1;
            return c3 - ((na + 1) * na * (na - 1)) / 6 + pa * na - ((pa * (pa + 1)) >> 1) + (c - b - 1);
        };
        auto k3 = [&](int a, int b, int c, bool& special) {
            const int ab = a * FZ_TLD + b, ac = a * FZ_TLD + c, bc = b * FZ_TLD + c, ba = b * FZ_TLD + a, ca = c * FZ_TLD + a;
            return fz_k3_straight(S1[bc], S1[ac], S1[ab], S1[ca], S1[ba], S2[ac], S2[ab], S2[ca], S2[ba], S3[ab], S3[ba], A2[ab], special);
        };
        // chunk 0 holds the reference indices [0, n0): a triple whose first position is >= a_lim0 starts at tri_off[a] >= n0
        int a_lim0 = 0;
        while (a_lim0 < m - 2 && (int)tri_off[a_lim0] < n0) ++a_lim0;
        // tests in flight per thread and iteration.  Round 1 measured 2 best (the dependent fp64 chains of two tests overlap); since the
        // hot loop lost its selects / eager index arithmetic, 1 is faster (C4: 45.7 -> 39.1 ms): half the loop body, fewer instruction-cache
        // misses (stall_no_inst was 18 %) at the same 5 CTAs per SM
#ifndef FW_HITON_INFLIGHT
#define FW_HITON_INFLIGHT 1
#endif
        constexpr int NF = FW_HITON_INFLIGHT;
        for (int q = tid; q < c3; q += NF * THREADS) {
            int ia[NF], ib[NF], ic[NF]; bool has[NF], sp[NF]; double sv[NF];
            // the smallest |stat| accepted so far in this scan (any thread): a larger significant |stat| needs no bookkeeping at all
            const double skip_above = __longlong_as_double((i64)*reinterpret_cast<volatile u64*>(bound)) * (1.0 + 1e-8);
#pragma unroll
            for (int u = 0; u < NF; ++u) {
                has[u] = q + u * THREADS < c3;
                decode(has[u] ? q + u * THREADS : q, ia[u], ib[u], ic[u]);
                sp[u] = false;
            }
#pragma unroll
            for (int u = 0; u < NF; ++u) sv[u] = k3(ia[u], ib[u], ic[u], sp[u]);
            bool slow = false;
#pragma unroll
            for (int u = 0; u < NF; ++u) {
                const double t = fabs(sv[u]);
                slow |= has[u] && (sp[u] || ia[u] < a_lim0 || !(t >= fc.s_hi && t > skip_above));
            }
            if (slow) {
#pragma unroll
                for (int u = 0; u < NF; ++u) {
                    if (!has[u]) continue;
                    const int idx = ref_index(ia[u], ib[u], ic[u]);
                    if (idx < n0) continue;                                     // already evaluated in chunk 0
                    if (sp[u]) sv[u] = pcor_generic(r, xs, ys, acc[ia[u]], acc[ib[u]], acc[ic[u]], 3);
                    consider(idx, sv[u]);
                }
            }
        }
        for (int q = tid; q < c2; q += THREADS) {
            const int a = plx[q] & 31, b = plx[q] >> 5;
            const int idx = c3 + a * m - ((a * (a + 1)) >> 1) + (b - a - 1);
            if (idx < n0) continue;
            consider(idx, A2[a * FZ_TLD + b]);
        }
        for (int q = tid; q < m; q += THREADS) {
            if (c3 + c2 + q < n0) continue;
            consider(c3 + c2 + q, (double)A1[q]);
        }
        executed = total;
        any_fail = __syncthreads_or(st.my_fail != NOFAIL);
    }
    SubsetCounts sc; sc.c3 = c3; sc.c2 = c2; sc.c1 = m; sc.total = total;
    if (any_fail) {
        if (st.my_fail != NOFAIL) atomicMin(&sh->fail_idx, (u64)st.my_fail);
        __syncthreads();
        if (st.my_fail != NOFAIL && (u64)st.my_fail == sh->fail_idx) {
            int k = 0, a = 0, b = 0, c = 0;
            if constexpr (NEED_POS) unrank_subset((i64)st.my_fail, m, sc, tri_off, k, a, b, c);
            const double f_p = fz_pval_dev(st.f_stat, fc);
            out->stat = st.f_stat; out->pval = f_p; out->df = 0; out->suff = 1;
            out->sig = (f_p < alpha) ? 1 : 0;
            out->k = k; out->pos[0] = a; out->pos[1] = b; out->pos[2] = c;
            out->num_tests = (i64)st.my_fail + 1; out->total = total; out->executed = executed; executed_by_k(sc, executed, out->ex_k);
        }
        __syncthreads();
        return;
    }
    // all significant: only threads whose minimum |stat| is within the tie tolerance of the CTA's minimum (or everything is in
    // the underflow range, where the bound is never lowered) can hold the maximum p-value; they evaluate it exactly and queue
    // it for thread 0, which picks the maximum (ties -> later subset) - usually out of a single entry
    constexpr double TOL = 1e-8;
    const double wmin = __longlong_as_double((i64)*reinterpret_cast<volatile u64*>(bound));
    if (st.best_idx >= 0 && (st.best_abs <= wmin * (1.0 + TOL) || st.best_abs >= fc.s_under)) {
        const int slot = atomicAdd(&sh->n_cont, 1);
        sh->c_stat[slot] = st.b_stat; sh->c_idx[slot] = st.best_idx;
    }
    __syncthreads();
    if (tid == 0) {
        const int nc = sh->n_cont;
        double b_p = -1.0, b_stat = sh->c_stat[0]; i64 bidx = sh->c_idx[0];
        if (nc > 1) {
            bidx = -1;
            for (int w = 0; w < nc; ++w) {
                const double op = fz_pval_dev(sh->c_stat[w], fc); const i64 oi = sh->c_idx[w];
                if (op > b_p || (op == b_p && oi > bidx)) { b_p = op; b_stat = sh->c_stat[w]; bidx = oi; }
            }
        }
        int k = 0, a = 0, b = 0, c = 0;
        if constexpr (NEED_POS) unrank_subset(bidx, m, sc, tri_off, k, a, b, c);
        out->stat = b_stat; out->pval = b_p; out->df = 0; out->suff = 1;
        out->sig = 1;                                   // every subset was significant
        out->k = k; out->pos[0] = a; out->pos[1] = b; out->pos[2] = c;
        out->num_tests = total; out->total = total; out->executed = executed; executed_by_k(sc, executed, out->ex_k);
    }
    // no trailing barrier: `out` is complete for thread 0 only; the caller's next barrier publishes it
}

// GS: R lives in global scratch (the unbounded capacity class); otherwise it is a plain shared-memory array, which lets
// the compiler emit LDS instead of generic loads in the test arithmetic.
// LISTS = false compiles the whitelist / blacklist / rejection-record handling out (no lists passed: the default call).
template <int THREADS, int TPT, bool NZ, bool GS, bool CACHE, bool LISTS = true>
// resident CTAs per SM of the 32-slot class: 4 (124 registers, no spills) measured best with one test in flight (C4: 39.1 -> 37.6 ms against 5 CTAs / 96 registers)
#ifndef FW_HITON_MINB
#define FW_HITON_MINB 4
#endif
__global__ void __launch_bounds__(THREADS, (THREADS == 128 && !NZ) ? (CACHE ? FW_HITON_MINB : 8) : ((NZ && !GS) ? 2 : 1)) hiton_fz_kernel(HitonArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    const int cap = a.cap;
    // carve shared memory
    size_t o = 0;
    float* Rs = reinterpret_cast<float*>(smem + o);
    if (!GS) o += sizeof(float) * (size_t)cap * cap;
    o = (o + 15) & ~(size_t)15;
    i64* tri_off = reinterpret_cast<i64*>(smem + o); o += sizeof(i64) * (cap + 1);
    double* tpc_stat = reinterpret_cast<double*>(smem + o); o += sizeof(double) * cap;
    double* tpc_p = reinterpret_cast<double*>(smem + o); o += sizeof(double) * cap;
    double* pcs_stat = reinterpret_cast<double*>(smem + o); o += sizeof(double) * cap;
    double* pcs_p = reinterpret_cast<double*>(smem + o); o += sizeof(double) * cap;
    i64* member = reinterpret_cast<i64*>(smem + o); o += sizeof(i64) * cap;
    int* acc = reinterpret_cast<int*>(smem + o); o += sizeof(int) * cap;
    int* pc_slot = reinterpret_cast<int*>(smem + o); o += sizeof(int) * cap;
    int* wl_slot = reinterpret_cast<int*>(smem + o); o += sizeof(int) * cap;           // processed whitelisted slots (elimination phase)
    unsigned char* sflag = smem + o; o += (size_t)cap;                                  // per slot: list flags of the member
    o = (o + 15) & ~(size_t)15;
    // fz_nz scratch: slot -> variable, per-variable (mean, norm), row mask of the current view
    i64* slotvar = reinterpret_cast<i64*>(smem + o); if (NZ) o += sizeof(i64) * cap;
    double* mom = reinterpret_cast<double*>(smem + o); if (NZ) o += sizeof(double) * 2 * cap;
    unsigned int* vmask = reinterpret_cast<unsigned int*>(smem + o);
    if (NZ) o += sizeof(unsigned int) * a.nzt.W;
    o = (o + 15) & ~(size_t)15;
    const FzTab tb = fz_tab_layout((int)o);      // only carved (and only valid) when CACHE
    __shared__ EvalShared sh;
    __shared__ EvalOut ev;
    __shared__ int s_ti, s_nc, s_M, s_macc, s_npc, s_cnt[2], s_spec, s_npre, s_nrej;
    __shared__ FzScanShared fsh;
    __shared__ i64 s_ntests;
    __shared__ u64 s_exec, s_exk[3];

    float* R;
    if constexpr (GS) R = a.gscratch + (size_t)blockIdx.x * cap * cap; else R = Rs;
    const int ld = cap;
    // the p-value-free scan assumes max_tests cannot bind: C(30,3) + C(30,2) + 30 = 4525 subsets at most in this class
    const bool max_tests_free = a.max_tests <= 0 || a.max_tests > 4525;
    if constexpr (CACHE) fz_build_colex<THREADS>(tb);         // first use is behind the __syncthreads() of the target loop

    for (;;) {
        __syncthreads();
        if (tid == 0) s_ti = atomicAdd(a.counter, 1);
        __syncthreads();
        const int ti = s_ti;
        if (ti >= a.n_sel) break;
        const int tsel = a.sel[ti];
        const i64 T = a.targets[tsel];
        const i64 e0 = a.uni_off[T];
        const int n_uni = (int)(a.uni_off[T + 1] - e0);
        const i64 o0 = a.out_off[tsel];
        int* order = a.cand_order + o0;

        // ---- prepare_interleaving_phase (hiton.jl:199-220): p < alpha, stable sort by p ----
        if (tid == 0) { s_nc = 0; s_M = 0; s_ntests = 0; s_exec = 0; s_exk[0] = s_exk[1] = s_exk[2] = 0; s_spec = 0; s_npre = 0; s_nrej = 0; }
        __syncthreads();
        for (int i = tid; i < n_uni; i += THREADS) {
            double pi = a.uni_p[e0 + i];
            if (pi < a.alpha) {
                int rank = 0;
                for (int j = 0; j < n_uni; ++j) {
                    double pj = a.uni_p[e0 + j];
                    if (pj < a.alpha && (pj < pi || (pj == pi && j < i))) ++rank;
                }
                order[rank] = LISTS ? (i | (hiton_list_flags(a.lists, tsel, a.uni_nbr[e0 + i]) << 28)) : i;
                atomicAdd(&s_nc, 1);
            }
        }
        __syncthreads();
        const int n_c = s_nc;
        bool overflow = false;
        const bool track = LISTS && a.lists.rej_count != nullptr;
        // rejection record of the candidate just scanned (thread 0; hiton.jl:72-74).  pos -> variable through the accepted list
        auto reject = [&](i64 cand, const AccView& av) {
            const i64 r = o0 + s_nrej;
            a.lists.rej_nbr[r] = cand; a.lists.rej_k[r] = ev.k;
            for (int i = 0; i < 3; ++i) a.lists.rej_Zs[r * 3 + i] = i < ev.k ? member[av[ev.pos[i]] - 1] : -1;
            a.lists.rej_res[r] = make_result(ev.stat, ev.pval, ev.df, ev.suff != 0);
            a.lists.rej_ntests[r] = ev.num_tests; a.lists.rej_frac[r] = ev.total > 0 ? (double)ev.num_tests / (double)ev.total : 0.0;
            s_nrej = s_nrej + 1;
        };

        // One barrier-separated serial section per candidate: thread 0 consumes the scan result (`ev`), does the accept /
        // reject bookkeeping of update_sig_result! (hiton.jl:53-78) and prepares the accepted list of the next candidate.
        // ---- interleaving phase (hiton.jl:109-149, phase 'I') ------------------------------
        for (int ci = 0; ci < n_c; ++ci) {
            const int M = s_M;                         // accepted so far; acc[0..M) = slots 1..M (kept by thread 0)
            if (M + 2 > cap) { overflow = true; break; }
            const int ui = LISTS ? (order[ci] & HITON_ORDER_MASK) : order[ci], lf = LISTS ? (order[ci] >> 28) : 0;
            if (lf == 2) continue;                     // blacklisted (and not whitelisted): skipped untested (hiton.jl:31-34)
            const i64 cand = a.uni_nbr[e0 + ui];
            const int ys = M + 1;
            if constexpr (!NZ) {
                // gather the candidate's correlations with T and the members into slot ys
                // (incl. the diagonal entry cor_mat[cand, cand]: a whitelisted member can occur twice in a conditioning set, hiton.jl:20-29)
                for (int s = tid; s <= M + 1; s += THREADS) {
                    i64 other = (s == 0) ? T : (s == ys ? cand : member[s - 1]);
                    float v = a.cv.at(cand, other);
                    R[ys * ld + s] = v; R[s * ld + ys] = v;
                }
            } else {
                for (int s = tid; s <= M + 1; s += THREADS) slotvar[s] = (s == 0) ? T : (s == ys ? cand : member[s - 1]);
            }
            __syncthreads();
            bool accept = false;                       // meaningful in thread 0 only
            if (lf & 1) {
                // whitelisted: accepted untested with (NaN, NaN) (hiton.jl:20-29)
                if (tid == 0) { tpc_stat[M] = __longlong_as_double(0x7ff8000000000000LL); tpc_p[M] = tpc_stat[M]; accept = true; }
            } else if (M == 0) {
                // accepted empty: accept with the univariate result (hiton.jl:57-59)
                if (tid == 0) { tpc_stat[0] = a.uni_stat[e0 + ui]; tpc_p[0] = a.uni_p[e0 + ui]; accept = true; }
            } else {
                FzSlotTest tf; tf.r.R = R; tf.r.ld = ld; tf.x = 0; tf.y = ys; tf.fc = a.fc;
                bool run = true;
                AccView av; av.n_head = M; av.head_base = 1; av.tail = pc_slot; av.n_pre = 0; av.pre = wl_slot;
                if constexpr (NZ) {
                    // cor_subset! on the rows where T != 0 and candidate != 0 (tests.jl:293-308; hiton.jl:41-50,85)
                    const int rows = fznz_subcor_block<THREADS>(a.nzt, slotvar, M + 2, 0, ys, R, ld, vmask, mom, s_cnt);
                    run = !(a.n_obs_min > (i64)rows);                     // else (0, 1, 0, false), zero tests: rejected
                    tf.fc = nz_consts(rows, a.n_obs_min);
                    if (!run && track && tid == 0) { ev.stat = 0.0; ev.pval = 1.0; ev.df = 0; ev.suff = 0; ev.k = 0; ev.num_tests = 0; ev.total = 0; reject(cand, av); }
                }
                if (run) {
                    bool cached = false;
                    if constexpr (CACHE) {
                        cached = (M >= 3 && a.max_k >= 3 && max_tests_free) && fz_build_tables<THREADS>(R, ld, 0, ys, av, M, tb, tri_off, &fsh);
                        if (cached) eval_subsets_fz_cached<THREADS, false>(tf.r, tb, 0, ys, av, M, a.alpha, tf.fc, tri_off, &fsh, &ev, s_spec != 0);
                    }
                    if (!cached) eval_subsets<THREADS, TPT, 1>(tf, acc, M, a.max_k, a.alpha, a.max_tests, tri_off, &sh, &ev);
                    if (tid == 0) {
                        s_ntests += ev.num_tests; s_exec += (u64)ev.executed; s_exk[0] += (u64)ev.ex_k[0]; s_exk[1] += (u64)ev.ex_k[1]; s_exk[2] += (u64)ev.ex_k[2];
                        if (ev.sig) { tpc_stat[M] = ev.stat; tpc_p[M] = ev.pval; accept = true; }
                        else if (track) reject(cand, av);
                        s_spec = (cached && ev.sig && ev.num_tests == ev.total) ? 1 : 0;
                    }
                }
            }
            if (tid == 0 && accept) { member[M] = cand; acc[M] = M + 1; if (LISTS) sflag[M + 1] = (unsigned char)lf; s_M = M + 1; }
            __syncthreads();
        }
        if (overflow) {
            if (tid == 0) a.status[tsel] = 1;
            continue;
        }

        // ---- elimination phase (phase 'E', fast_elim = true) ---------------------------------
        // accepted list of candidate slot c = AccView{untested slots c+1..M, then pc_slot[0..npc)}: deleteat!(accepted, candidate)
        // and push!(accepted, candidate) of hiton.jl:134-147 without moving anything
        const int M = s_M;
        if (tid == 0) s_npc = 0;
        __syncthreads();
        for (int c = 1; c <= M; ++c) {
            const int npc0 = s_npc, npre = LISTS ? s_npre : 0;
            const int macc = npre + (M - c) + npc0;
            AccView av; av.n_head = M - c; av.head_base = c + 1; av.tail = pc_slot; av.n_pre = npre; av.pre = wl_slot;
            bool accept = false;                       // thread 0 only
            if (LISTS && (sflag[c] & 1)) {
                // whitelisted: (NaN, NaN), pushed onto `accepted` again while its original entry stays (hiton.jl:20-29, 124-131)
                __syncthreads();
                if (tid == 0) { pcs_stat[npc0] = __longlong_as_double(0x7ff8000000000000LL); pcs_p[npc0] = pcs_stat[npc0]; wl_slot[npre] = c; s_npre = npre + 1; accept = true; }
            } else if (macc == 0) {
                __syncthreads();                       // every thread has read s_npc before thread 0 may advance it (no scan, hence no barrier, on this path)
                if (tid == 0) { pcs_stat[npc0] = tpc_stat[c - 1]; pcs_p[npc0] = tpc_p[c - 1]; accept = true; }   // support_dict = TPC_dict
            } else {
                if (macc + 2 > cap) { overflow = true; break; }          // duplicates of whitelisted members outgrew the class: re-run in the next one
                FzSlotTest tf; tf.r.R = R; tf.r.ld = ld; tf.x = 0; tf.y = c; tf.fc = a.fc;
                bool run = true;
                if constexpr (NZ) {
                    for (int s = tid; s <= M; s += THREADS) slotvar[s] = (s == 0) ? T : member[s - 1];
                    __syncthreads();
                    const int rows = fznz_subcor_block<THREADS>(a.nzt, slotvar, M + 1, 0, c, R, ld, vmask, mom, s_cnt);
                    run = !(a.n_obs_min > (i64)rows);
                    tf.fc = nz_consts(rows, a.n_obs_min);
                    if (!run && track && tid == 0) { ev.stat = 0.0; ev.pval = 1.0; ev.df = 0; ev.suff = 0; ev.k = 0; ev.num_tests = 0; ev.total = 0; reject(member[c - 1], av); }
                }
                if (run) {
                    bool cached = false;
                    if constexpr (CACHE) {
                        cached = (macc >= 3 && macc <= FZ_CACHE_CAP - 2 && a.max_k >= 3 && max_tests_free) && fz_build_tables<THREADS>(R, ld, 0, c, av, macc, tb, tri_off, &fsh);
                        if (cached) eval_subsets_fz_cached<THREADS, false>(tf.r, tb, 0, c, av, macc, a.alpha, tf.fc, tri_off, &fsh, &ev, s_spec != 0);
                    }
                    if (!cached) {
                        for (int j = tid; j < macc; j += THREADS) acc[j] = av[j];
                        __syncthreads();
                        eval_subsets<THREADS, TPT, 1>(tf, acc, macc, a.max_k, a.alpha, a.max_tests, tri_off, &sh, &ev);
                    }
                    if (tid == 0) {
                        s_ntests += ev.num_tests; s_exec += (u64)ev.executed; s_exk[0] += (u64)ev.ex_k[0]; s_exk[1] += (u64)ev.ex_k[1]; s_exk[2] += (u64)ev.ex_k[2];
                        if (ev.sig) { pcs_stat[npc0] = ev.stat; pcs_p[npc0] = ev.pval; accept = true; }
                        else if (track) reject(member[c - 1], av);
                        s_spec = (cached && ev.sig && ev.num_tests == ev.total) ? 1 : 0;
                    }
                }
            }
            if (tid == 0 && accept) { pc_slot[npc0] = c; s_npc = npc0 + 1; }
            __syncthreads();
        }
        if (overflow) {
            if (tid == 0) a.status[tsel] = 1;
            continue;
        }

        // ---- update_PC_dict! (hiton.jl:249-256) and write-out ---------------------------------
        const int npc = s_npc;
        if constexpr (CACHE) {
            // p-values the scans did not need (unique maximum |stat| ordering) are evaluated here, one thread each
            for (int i = tid; i < M + npc; i += THREADS) {
                double* pp = i < M ? &tpc_p[i] : &pcs_p[i - M];
                const double st_ = i < M ? tpc_stat[i] : pcs_stat[i - M];
                if (*pp < 0.0) *pp = fz_pval_dev(st_, a.fc);
            }
            __syncthreads();
        }
        for (int i = tid; i < npc; i += THREADS) {
            int c = pc_slot[i];
            double s = pcs_stat[i], pp = pcs_p[i];
            double ts = tpc_stat[c - 1], tp = tpc_p[c - 1];
            if (tp > pp || isnan(pp)) { s = ts; pp = tp; }
            if (a.pc_nbr) { a.pc_nbr[o0 + i] = member[c - 1]; a.pc_stat[o0 + i] = s; a.pc_p[o0 + i] = pp; }
        }
        for (int i = tid; i < M; i += THREADS) {
            if (a.tpc_nbr) { a.tpc_nbr[o0 + i] = member[i]; a.tpc_stat[o0 + i] = tpc_stat[i]; a.tpc_p[o0 + i] = tpc_p[i]; }
        }
        if (tid == 0) {
            a.pc_count[tsel] = npc; a.tpc_count[tsel] = M; a.num_tests[tsel] = s_ntests; a.status[tsel] = 0;
            if (track) a.lists.rej_count[tsel] = s_nrej;
            atomicAdd(a.executed_total, s_exec);
            atomicAdd(a.executed_total + 1, s_exk[0]); atomicAdd(a.executed_total + 2, s_exk[1]); atomicAdd(a.executed_total + 3, s_exk[2]);
        }
    }
}

// -------------------------------------------------------------------------------------------
// Batched test_subsets: one CTA per (X, Y, Z_total) job.
// -------------------------------------------------------------------------------------------
struct SubsetsArgs {
    CorView cv; i64 p;
    const i64* X; const i64* Y; const i64* z_off; const i64* z_idx;
    const int* sel; int n_sel; int* counter;
    int max_k; double alpha; i64 max_tests; FzConsts fc;
    int cap; float* gscratch;
    DevResult* out; i64* out_Zs; int* out_k; i64* num_tests; double* frac; u64* executed_total;
    NzTable nzt; i64 n_obs_min;             // fz_nz only
};

template <int THREADS, int TPT, bool NZ, bool GS>
__global__ void __launch_bounds__(THREADS) subsets_fz_kernel(SubsetsArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    const int cap = a.cap;
    size_t o = 0;
    float* Rs = reinterpret_cast<float*>(smem + o);
    if (!GS) o += sizeof(float) * (size_t)cap * cap;
    o = (o + 15) & ~(size_t)15;
    i64* tri_off = reinterpret_cast<i64*>(smem + o); o += sizeof(i64) * (cap + 1);
    int* acc = reinterpret_cast<int*>(smem + o); o += sizeof(int) * cap;
    o = (o + 15) & ~(size_t)15;
    i64* slotvar = reinterpret_cast<i64*>(smem + o); if (NZ) o += sizeof(i64) * cap;
    double* mom = reinterpret_cast<double*>(smem + o); if (NZ) o += sizeof(double) * 2 * cap;
    unsigned int* vmask = reinterpret_cast<unsigned int*>(smem + o);
    __shared__ EvalShared sh;
    __shared__ EvalOut ev;
    __shared__ int s_ji, s_cnt[2];
    float* R;
    if constexpr (GS) R = a.gscratch + (size_t)blockIdx.x * cap * cap; else R = Rs;
    const int ld = cap;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_ji = atomicAdd(a.counter, 1);
        __syncthreads();
        if (s_ji >= a.n_sel) break;
        const int job = a.sel[s_ji];
        const i64 z0 = a.z_off[job];
        const int m = (int)(a.z_off[job + 1] - z0);
        const int nv = m + 2;
        FzSlotTest tf; tf.r.R = R; tf.r.ld = ld; tf.x = 0; tf.y = 1; tf.fc = a.fc;
        for (int s = tid; s < m; s += THREADS) acc[s] = s + 2;
        if constexpr (!NZ) {
            // gather the (m+2)^2 sub-block: slot 0 = X, 1 = Y, 2.. = Z_total
            for (int e = tid; e < nv * nv; e += THREADS) {
                int i = e / nv, j = e % nv;
                i64 vi = i == 0 ? a.X[job] : (i == 1 ? a.Y[job] : a.z_idx[z0 + i - 2]);
                i64 vj = j == 0 ? a.X[job] : (j == 1 ? a.Y[job] : a.z_idx[z0 + j - 2]);
                R[i * ld + j] = a.cv.at(vi, vj);
            }
            __syncthreads();
        } else {
            for (int s = tid; s < nv; s += THREADS) slotvar[s] = s == 0 ? a.X[job] : (s == 1 ? a.Y[job] : a.z_idx[z0 + s - 2]);
            __syncthreads();
            const int rows = fznz_subcor_block<THREADS>(a.nzt, slotvar, nv, 0, 1, R, ld, vmask, mom, s_cnt);
            if (a.n_obs_min > (i64)rows) {                                  // tests.jl:293-296
                if (tid == 0) {
                    a.out[job] = make_result(0.0, 1.0, 0, false);
                    for (int i = 0; i < 3; ++i) a.out_Zs[job * 3 + i] = -1;
                    a.out_k[job] = 0; a.num_tests[job] = 0; a.frac[job] = 0.0;
                }
                continue;
            }
            tf.fc = nz_consts(rows, a.n_obs_min);
        }
        eval_subsets<THREADS, TPT, 1>(tf, acc, m, a.max_k, a.alpha, a.max_tests, tri_off, &sh, &ev);
        if (tid == 0) {
            a.out[job] = make_result(ev.stat, ev.pval, ev.df, ev.suff != 0);
            for (int i = 0; i < 3; ++i) a.out_Zs[job * 3 + i] = (i < ev.k) ? a.z_idx[z0 + ev.pos[i]] : -1;
            a.out_k[job] = ev.k;
            a.num_tests[job] = ev.num_tests;
            a.frac[job] = (double)ev.num_tests / (double)ev.total;
            atomicAdd(a.executed_total, (u64)ev.executed);
        }
    }
}

// -------------------------------------------------------------------------------------------
// Independent conditional tests, one thread each (fw_test_batch, kind fz): test(X,Y,Zs,...)
// -------------------------------------------------------------------------------------------
__global__ void fz_test_batch_kernel(const CorView cv, i64 p, i64 n_tests, const i64* X, const i64* Y, const int* k,
                                     const i64* Zs, FzConsts fc, i64 n_rows, i64 n_obs_min, DevResult* out) {
    i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tests) return;
    i64 var[5] = {X[t], Y[t], Zs[t * 3], Zs[t * 3 + 1], Zs[t * 3 + 2]};
    int kk = k[t];
    if (kk == 0) {
        // tests.jl:108-160 with a precomputed cor_mat (:149-156)
        double stat = (n_rows >= n_obs_min) ? (double)cv.at(var[0], var[1]) : 0.0;
        if (n_rows < n_obs_min) { out[t] = make_result(0.0, 1.0, 0, 0 >= n_obs_min); return; }
        out[t] = make_result(stat, fz_pval_dev(stat, fc), 0, true);
        return;
    }
    for (int i = kk + 2; i < 5; ++i) var[i] = var[0];
    CorGlobal r; r.cv = cv; r.var = var;
    FzTest ft = fz_cond_test(r, 0, 1, 2, 3, 4, kk, fc);
    out[t] = make_result(ft.stat, ft.pval, 0, ft.suff);
}
