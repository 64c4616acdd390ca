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
    const float* cor; i64 p;                 // cor_mat (p x p, symmetric)
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
};

struct FzSlotTest {
    CorSlots r; int x, y; FzConsts fc;
    __device__ __forceinline__ FzTest operator()(int k, int za, int zb, int zc) const {
        return fz_cond_test(r, x, y, za, zb, zc, k, fc);
    }
};

// ---- per-candidate tables (capacity class 32): every conditioning subset of one candidate shares its level-1 terms with all
// subsets that start with the same Z1, and its (X,Y|Z1,Z2) level-2 term with all subsets that start with the same (Z1,Z2).
// Caching them in shared memory leaves 1 level-1 + 2 level-2 + 1 level-3 step per k = 3 test (was 6 + 3 + 1) with bit-identical
// arithmetic (same operations, same order, just not repeated).  A Float64 literal (special) anywhere in the tables disables the
// cache for that candidate, so the generic typed path keeps handling the degenerate cases.
struct FzTables {
    float* SQ;      // SQ[z*ld + s] = sqrt(1f0 - r(s,z)^2)
    float* BX;      // BX[z*ld + s] = pcor(X, s | z)   (level 1, Float32)
    float* BY;      // BY[z*ld + s] = pcor(Y, s | z)
    float* A1;      // A1[z]        = pcor(X, Y | z)
    double* A2;     // A2[z1*ld + z2] = pcor(X, Y | z1, z2) for z1 before z2 in the accepted list (level 2, Float64)
    int ld;
};

template <int THREADS>
__device__ bool fz_build_tables(const float* R, int ld, int xs, int ys, const int* acc, int m, const FzTables& T, int* s_special) {
    const int tid = threadIdx.x;
    if (tid == 0) *s_special = 0;
    // SQ for z in acc, s in acc + {x, y}
    for (int e = tid; e < m * (m + 2); e += THREADS) {
        const int ia = e / (m + 2), ib = e % (m + 2);
        const int z = acc[ia], sl = ib < m ? acc[ib] : (ib == m ? xs : ys);
        if (sl != z) T.SQ[z * ld + sl] = sq1mf(R[sl * ld + z]);
    }
    __syncthreads();
    bool ok = true;
    for (int e = tid; e < m * m; e += THREADS) {
        const int ia = e / m, ib = e % m;
        const int z = acc[ia];
        if (ia == ib) {
            float a1;
            ok &= p1f(R[xs * ld + ys], R[xs * ld + z], R[ys * ld + z], T.SQ[z * ld + xs], T.SQ[z * ld + ys], a1);
            T.A1[z] = a1;
        } else {
            const int sl = acc[ib];
            float bx, by;
            ok &= p1f(R[xs * ld + sl], R[xs * ld + z], R[sl * ld + z], T.SQ[z * ld + xs], T.SQ[z * ld + sl], bx);
            ok &= p1f(R[ys * ld + sl], R[ys * ld + z], R[sl * ld + z], T.SQ[z * ld + ys], T.SQ[z * ld + sl], by);
            T.BX[z * ld + sl] = bx; T.BY[z * ld + sl] = by;
        }
    }
    if (!ok) *s_special = 1;
    __syncthreads();
    for (int e = tid; e < m * m; e += THREADS) {
        const int ia = e / m, ib = e % m;
        if (ia < ib) { const int z1 = acc[ia], z2 = acc[ib]; T.A2[z1 * ld + z2] = p2f(T.A1[z1], T.BX[z1 * ld + z2], T.BY[z1 * ld + z2]); }
    }
    __syncthreads();
    return *s_special == 0;
}

struct FzCachedTest {
    CorSlots r; FzTables T; int x, y; FzConsts fc;
    __device__ __forceinline__ FzTest operator()(int k, int za, int zb, int zc) const {
        FzTest t; t.df = 0;
        if (!fc.rows_ok) { t.stat = 0.0; t.pval = 1.0; t.suff = false; return t; }
        if (k == 1) t.stat = (double)T.A1[za];
        else if (k == 2) t.stat = T.A2[za * T.ld + zb];
        else {
            float z3z2;
            const bool ok = p1f(r(zc, zb), r(zc, za), r(zb, za), T.SQ[za * T.ld + zc], T.SQ[za * T.ld + zb], z3z2);
            if (ok) {
                const double B = p2f(T.BX[za * T.ld + zc], T.BX[za * T.ld + zb], z3z2);     // pcor(X, Z3 | Z1, Z2)
                const double C = p2f(T.BY[za * T.ld + zc], T.BY[za * T.ld + zb], z3z2);     // pcor(Y, Z3 | Z1, Z2)
                t.stat = p3d(T.A2[za * T.ld + zb], B, C);
            } else t.stat = pcor_generic(r, x, y, za, zb, zc, 3);
        }
        t.pval = fz_pval_dev(t.stat, fc);
        t.suff = true;
        return t;
    }
};

// GS: R lives in global scratch (the unbounded capacity class); otherwise it is a plain shared-memory array, which lets
// the compiler emit LDS instead of generic loads in the test arithmetic.
template <int THREADS, int TPT, bool NZ, bool GS, bool CACHE>
__global__ void __launch_bounds__(THREADS, (THREADS == 128 && !NZ) ? 8 : 1) hiton_fz_kernel(HitonArgs a) {
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
    o = (o + 15) & ~(size_t)15;
    // fz_nz scratch: slot -> variable, per-variable (mean, norm), row mask of the current view
    i64* slotvar = reinterpret_cast<i64*>(smem + o); if (NZ) o += sizeof(i64) * cap;
    double* mom = reinterpret_cast<double*>(smem + o); if (NZ) o += sizeof(double) * 2 * cap;
    unsigned int* vmask = reinterpret_cast<unsigned int*>(smem + o);
    if (NZ) o += sizeof(unsigned int) * a.nzt.W;
    o = (o + 15) & ~(size_t)15;
    FzTables tb;
    tb.ld = cap;
    tb.A2 = reinterpret_cast<double*>(smem + o); if (CACHE) o += sizeof(double) * (size_t)cap * cap;
    tb.SQ = reinterpret_cast<float*>(smem + o); if (CACHE) o += sizeof(float) * (size_t)cap * cap;
    tb.BX = reinterpret_cast<float*>(smem + o); if (CACHE) o += sizeof(float) * (size_t)cap * cap;
    tb.BY = reinterpret_cast<float*>(smem + o); if (CACHE) o += sizeof(float) * (size_t)cap * cap;
    tb.A1 = reinterpret_cast<float*>(smem + o);
    __shared__ EvalShared sh;
    __shared__ EvalOut ev;
    __shared__ int s_ti, s_nc, s_M, s_macc, s_npc, s_accept, s_cnt, s_special;
    __shared__ i64 s_ntests;
    __shared__ u64 s_exec, s_exk[3];

    float* R;
    if constexpr (GS) R = a.gscratch + (size_t)blockIdx.x * cap * cap; else R = Rs;
    const int ld = cap;

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
        if (tid == 0) { s_nc = 0; s_M = 0; s_ntests = 0; s_exec = 0; s_exk[0] = s_exk[1] = s_exk[2] = 0; }
        __syncthreads();
        for (int i = tid; i < n_uni; i += THREADS) {
            double pi = a.uni_p[e0 + i];
            if (pi < a.alpha) {
                int rank = 0;
                for (int j = 0; j < n_uni; ++j) {
                    double pj = a.uni_p[e0 + j];
                    if (pj < a.alpha && (pj < pi || (pj == pi && j < i))) ++rank;
                }
                order[rank] = i;
                atomicAdd(&s_nc, 1);
            }
        }
        __syncthreads();
        const int n_c = s_nc;
        bool overflow = false;

        // ---- interleaving phase (hiton.jl:109-149, phase 'I') ------------------------------
        for (int ci = 0; ci < n_c; ++ci) {
            const int M = s_M;
            if (M + 2 > cap) { overflow = true; break; }
            const int ui = order[ci];
            const i64 cand = a.uni_nbr[e0 + ui];
            const int ys = M + 1;
            if constexpr (!NZ) {
                // gather the candidate's correlations with T and the members into slot ys
                for (int s = tid; s <= M; s += THREADS) {
                    i64 other = (s == 0) ? T : member[s - 1];
                    float v = __ldg(a.cor + cand * a.p + other);
                    R[ys * ld + s] = v; R[s * ld + ys] = v;
                }
            } else {
                for (int s = tid; s <= M + 1; s += THREADS) slotvar[s] = (s == 0) ? T : (s == ys ? cand : member[s - 1]);
            }
            if (tid == 0) s_accept = 0;
            __syncthreads();
            if (M == 0) {
                // accepted empty: accept with the univariate result (hiton.jl:57-59)
                if (tid == 0) { tpc_stat[0] = a.uni_stat[e0 + ui]; tpc_p[0] = a.uni_p[e0 + ui]; s_accept = 1; }
            } else {
                if (tid < M) acc[tid] = tid + 1;
                for (int s = tid + THREADS; s < M; s += THREADS) acc[s] = s + 1;
                __syncthreads();
                FzSlotTest tf; tf.r.R = R; tf.r.ld = ld; tf.x = 0; tf.y = ys; tf.fc = a.fc;
                bool run = true;
                if constexpr (NZ) {
                    // cor_subset! on the rows where T != 0 and candidate != 0 (tests.jl:293-308; hiton.jl:41-50,85)
                    const int rows = fznz_subcor_block<THREADS>(a.nzt, slotvar, M + 2, 0, ys, R, ld, vmask, mom, &s_cnt);
                    run = !(a.n_obs_min > (i64)rows);                     // else (0, 1, 0, false), zero tests: rejected
                    tf.fc = nz_consts(rows, a.n_obs_min);
                }
                if (run) {
                    bool cached = false;
                    if constexpr (CACHE) cached = (M >= 3 && a.max_k >= 3) && fz_build_tables<THREADS>(R, ld, 0, ys, acc, M, tb, &s_special);
                    if (cached) {
                        FzCachedTest tc; tc.r = tf.r; tc.T = tb; tc.x = 0; tc.y = ys; tc.fc = tf.fc;
                        eval_subsets<THREADS, TPT, 1>(tc, acc, M, a.max_k, a.alpha, a.max_tests, tri_off, &sh, &ev);
                    } else {
                        eval_subsets<THREADS, TPT, 1>(tf, acc, M, a.max_k, a.alpha, a.max_tests, tri_off, &sh, &ev);
                    }
                    if (tid == 0) {
                        s_ntests += ev.num_tests; s_exec += (u64)ev.executed; s_exk[0] += (u64)ev.ex_k[0]; s_exk[1] += (u64)ev.ex_k[1]; s_exk[2] += (u64)ev.ex_k[2];
                        if (ev.sig) { tpc_stat[M] = ev.stat; tpc_p[M] = ev.pval; s_accept = 1; }
                    }
                }
            }
            __syncthreads();
            if (s_accept) { if (tid == 0) { member[M] = cand; s_M = M + 1; } }
            __syncthreads();
        }
        if (overflow) {
            if (tid == 0) a.status[tsel] = 1;
            continue;
        }

        // ---- elimination phase (phase 'E', fast_elim = true) ---------------------------------
        const int M = s_M;
        for (int s = tid; s < M; s += THREADS) acc[s] = s + 1;
        if (tid == 0) { s_macc = M; s_npc = 0; }
        __syncthreads();
        for (int c = 1; c <= M; ++c) {
            if (tid == 0) {
                // deleteat!(accepted, findall(in(candidate), accepted))  (hiton.jl:134-136)
                int w = 0, macc = s_macc;
                for (int j = 0; j < macc; ++j) { int v = acc[j]; if (v != c) acc[w++] = v; }
                s_macc = w; s_accept = 0;
            }
            __syncthreads();
            const int macc = s_macc;
            if (macc == 0) {
                if (tid == 0) { pcs_stat[s_npc] = tpc_stat[c - 1]; pcs_p[s_npc] = tpc_p[c - 1]; s_accept = 1; }   // support_dict = TPC_dict
            } else {
                FzSlotTest tf; tf.r.R = R; tf.r.ld = ld; tf.x = 0; tf.y = c; tf.fc = a.fc;
                bool run = true;
                if constexpr (NZ) {
                    for (int s = tid; s <= M; s += THREADS) slotvar[s] = (s == 0) ? T : member[s - 1];
                    __syncthreads();
                    const int rows = fznz_subcor_block<THREADS>(a.nzt, slotvar, M + 1, 0, c, R, ld, vmask, mom, &s_cnt);
                    run = !(a.n_obs_min > (i64)rows);
                    tf.fc = nz_consts(rows, a.n_obs_min);
                }
                if (run) {
                    bool cached = false;
                    if constexpr (CACHE) cached = (macc >= 3 && a.max_k >= 3) && fz_build_tables<THREADS>(R, ld, 0, c, acc, macc, tb, &s_special);
                    if (cached) {
                        FzCachedTest tc; tc.r = tf.r; tc.T = tb; tc.x = 0; tc.y = c; tc.fc = tf.fc;
                        eval_subsets<THREADS, TPT, 1>(tc, acc, macc, a.max_k, a.alpha, a.max_tests, tri_off, &sh, &ev);
                    } else {
                        eval_subsets<THREADS, TPT, 1>(tf, acc, macc, a.max_k, a.alpha, a.max_tests, tri_off, &sh, &ev);
                    }
                    if (tid == 0) {
                        s_ntests += ev.num_tests; s_exec += (u64)ev.executed; s_exk[0] += (u64)ev.ex_k[0]; s_exk[1] += (u64)ev.ex_k[1]; s_exk[2] += (u64)ev.ex_k[2];
                        if (ev.sig) { pcs_stat[s_npc] = ev.stat; pcs_p[s_npc] = ev.pval; s_accept = 1; }
                    }
                }
            }
            __syncthreads();
            if (tid == 0 && s_accept) { acc[s_macc] = c; s_macc = s_macc + 1; pc_slot[s_npc] = c; s_npc = s_npc + 1; }
            __syncthreads();
        }

        // ---- update_PC_dict! (hiton.jl:249-256) and write-out ---------------------------------
        const int npc = s_npc;
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
            atomicAdd(a.executed_total, s_exec);
            atomicAdd(a.executed_total + 1, s_exk[0]); atomicAdd(a.executed_total + 2, s_exk[1]); atomicAdd(a.executed_total + 3, s_exk[2]);
        }
    }
}

// -------------------------------------------------------------------------------------------
// Batched test_subsets: one CTA per (X, Y, Z_total) job.
// -------------------------------------------------------------------------------------------
struct SubsetsArgs {
    const float* cor; i64 p;
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
    __shared__ int s_ji, s_cnt;
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
                R[i * ld + j] = __ldg(a.cor + vi * a.p + vj);
            }
            __syncthreads();
        } else {
            for (int s = tid; s < nv; s += THREADS) slotvar[s] = s == 0 ? a.X[job] : (s == 1 ? a.Y[job] : a.z_idx[z0 + s - 2]);
            __syncthreads();
            const int rows = fznz_subcor_block<THREADS>(a.nzt, slotvar, nv, 0, 1, R, ld, vmask, mom, &s_cnt);
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
__global__ void fz_test_batch_kernel(const float* cor, i64 p, i64 n_tests, const i64* X, const i64* Y, const int* k,
                                     const i64* Zs, FzConsts fc, i64 n_rows, i64 n_obs_min, DevResult* out) {
    i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tests) return;
    i64 var[5] = {X[t], Y[t], Zs[t * 3], Zs[t * 3 + 1], Zs[t * 3 + 2]};
    int kk = k[t];
    if (kk == 0) {
        // tests.jl:108-160 with a precomputed cor_mat (:149-156)
        double stat = (n_rows >= n_obs_min) ? (double)__ldg(cor + var[0] * p + var[1]) : 0.0;
        if (n_rows < n_obs_min) { out[t] = make_result(0.0, 1.0, 0, 0 >= n_obs_min); return; }
        out[t] = make_result(stat, fz_pval_dev(stat, fc), 0, true);
        return;
    }
    for (int i = kk + 2; i < 5; ++i) var[i] = var[0];
    CorGlobal r; r.cor = cor; r.p = p; r.var = var;
    FzTest ft = fz_cond_test(r, 0, 1, 2, 3, 4, kk, fc);
    out[t] = make_result(ft.stat, ft.pval, 0, ft.suff);
}
