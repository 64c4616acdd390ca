// hiton_mi.cuh — device-resident si_HITON_PC (src/hiton.jl:283-400) and batched test_subsets
// (src/tests.jl:281-346) for the discrete kinds (mi / mi_nz).  Same control skeleton as hiton.cuh; the
// conditioning subsets of a candidate are evaluated one test per warp on the bit-plane table (mi.cuh).
#pragma once
#include "common.cuh"
#include "mi.cuh"
#include "fznz_tc.cuh"
#include "subsets.cuh"
#include "hiton.cuh"
#include "mi_lane.cuh"

struct MiSlotTest {
    MiTable t; const i64* var; int x, y; i64 hps; int* tab;
    __device__ __forceinline__ MiResult operator()(int k, int za, int zb, int zc) const {
        i64 Z[3] = {var[za], var[zb], var[zc]};
        return mi_test_warp(t, var[x], var[y], Z, k, hps, 0, tab);
    }
};

// ---- binary fast path of the subset scan: counting per warp, statistics per lane -------------------------------------------------
// mi_test_warp spends most of its instructions after the counting: the MI sum (fp64 log), the degrees of freedom and the
// chi^2 tail are scalar work that a warp executes for ONE test.  Here a warp takes a batch of B consecutive subsets: it counts the
// 2 x 2 x 2^k table of each of them cooperatively (lanes stride over the words; mask tree + inclusion-exclusion popcounts as in
// mi_test_warp_bin) into shared memory, then lane j evaluates the statistics of subset j - the scalar part is amortised over the
// batch.  Batches grow 1 -> 8 -> 32 subsets per warp, so the reference's early exit (which almost always happens in the first
// few subsets) still costs one small chunk.  Per-lane results feed the same first-failure / arg-max reductions as eval_subsets
// with one thread per test.  Valid for kind "mi", 2-level tables (L = 2) and 2-level X and Y; same integers and formulas as
// mi_test_warp_bin (statfuns.jl:163-305, tests.jl:184-229), summed in stratum order.
template <int K>
__device__ __forceinline__ void mi_count_bin_to_smem(const unsigned int* __restrict__ planes, int W, unsigned int tail_mask, i64 X, i64 Y, i64 z0, i64 z1, i64 z2, int* out /* 4 * 2^K cells */) {
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    constexpr int S = 1 << K;
    int cn[S], cx[S], cy[S], cxy[S];
#pragma unroll
    for (int s = 0; s < S; ++s) { cn[s] = 0; cx[s] = 0; cy[s] = 0; cxy[s] = 0; }
    const unsigned int* px = planes + (size_t)X * W;
    const unsigned int* py = planes + (size_t)Y * W;
    const unsigned int* pz0 = planes + (size_t)z0 * W;
    const unsigned int* pz1 = planes + (size_t)z1 * W;
    const unsigned int* pz2 = planes + (size_t)z2 * W;
    for (int w = lane; w < W; w += 32) {
        const unsigned int valid = (w == W - 1) ? tail_mask : 0xffffffffu;
        const unsigned int x = __ldg(px + w), y = __ldg(py + w);
        unsigned int m[S];
        m[0] = valid;
        if (K > 0) { const unsigned int z = __ldg(pz0 + w); m[1] = m[0] & z; m[0] &= ~z; }
        if (K > 1) { const unsigned int z = __ldg(pz1 + w);
#pragma unroll
            for (int s = 0; s < 2; ++s) { m[s + 2] = m[s] & z; m[s] &= ~z; } }
        if (K > 2) { const unsigned int z = __ldg(pz2 + w);
#pragma unroll
            for (int s = 0; s < 4; ++s) { m[s + 4] = m[s] & z; m[s] &= ~z; } }
        const unsigned int xy = x & y;
#pragma unroll
        for (int s = 0; s < S; ++s) { cn[s] += __popc(m[s]); cx[s] += __popc(m[s] & x); cy[s] += __popc(m[s] & y); cxy[s] += __popc(m[s] & xy); }
    }
    int mine = 0;                                   // lane c owns cell (a = c & 1, b = (c >> 1) & 1, stratum = c >> 2)
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const int n = __reduce_add_sync(full, cn[s]), nx = __reduce_add_sync(full, cx[s]);
        const int ny = __reduce_add_sync(full, cy[s]), nxy = __reduce_add_sync(full, cxy[s]);
        if ((lane >> 2) == s) {
            const int a = lane & 1, b = (lane >> 1) & 1;
            mine = a ? (b ? nxy : nx - nxy) : (b ? ny - nxy : n - nx - ny + nxy);
        }
    }
    if (lane < 4 * S) out[lane] = mine;
}

// statistics of one 2 x 2 x S table (S = 2^K strata, K >= 1), one thread: tests.jl:200-221, statfuns.jl:163-254, 281-305.
// The fp64 terms are accumulated in the order of the reference's loops `for i, j, k` (statfuns.jl:181: x level, y level, strata
// innermost; strata in raw-key order here - the reference walks them in first-seen order, which can differ in the last ulp, see
// DESIGN.md 4.5): the diagonal sum takes the (0,0) cells of all strata, then the (1,1) cells; the off-diagonal sum (0,1), then (1,0).
// cs: this lane's 4 S counts, uv: 2 S doubles of scratch for the deferred terms (both in shared memory, bank-conflict-free strides)
__device__ MiResult mi_stats_bin_thread(const int* cs /* [S][b][a] */, double* uv, int S, i64 hps) {
    MiResult r;
    i64 n_obs = 0; int levels_z = 0;
    for (int s = 0; s < S; ++s) { const int tot = cs[4 * s] + cs[4 * s + 1] + cs[4 * s + 2] + cs[4 * s + 3]; n_obs += tot; levels_z += tot > 0; }
    if (!(((double)n_obs / (double)(4 * levels_z)) > (double)hps)) { r.stat = 0.0; r.pval = 1.0; r.df = 0; r.suff = false; return r; }
    double pos = 0.0, neg = 0.0; i64 n_pos = 0, n_neg = 0; int df = 0;
    for (int s = 0; s < S; ++s) {
        const int c00 = cs[4 * s], c10 = cs[4 * s + 1], c01 = cs[4 * s + 2], c11 = cs[4 * s + 3];   // index = 2 b + a (a: X level, b: Y level)
        const int ma0 = c00 + c01, ma1 = c10 + c11;          // margins over Y for X = 0, 1 (marg_i)
        const int mb0 = c00 + c10, mb1 = c01 + c11;          // margins over X for Y = 0, 1 (marg_j)
        const int mk = ma0 + ma1;
        if (c00 != 0) { pos += log((double)((i64)mk * c00) / (double)((i64)ma0 * mb0)) * (double)c00; n_pos += c00; }
        if (c01 != 0) { neg += log((double)((i64)mk * c01) / (double)((i64)ma0 * mb1)) * (double)c01; n_neg += c01; }
        const double u = c11 != 0 ? log((double)((i64)mk * c11) / (double)((i64)ma1 * mb1)) * (double)c11 : 0.0;
        const double v = c10 != 0 ? log((double)((i64)mk * c10) / (double)((i64)ma1 * mb0)) * (double)c10 : 0.0;
        n_pos += c11; n_neg += c10;
        df += (ma0 > 0 && ma1 > 0 && mb0 > 0 && mb1 > 0) ? 1 : 0;
        uv[2 * s] = u; uv[2 * s + 1] = v;                      // (1,1) and (1,0) terms of stratum s
    }
    for (int s = 0; s < S; ++s) { pos += uv[2 * s]; neg += uv[2 * s + 1]; }       // adding an exact 0.0 term changes nothing
    const i64 n_mi = n_pos + n_neg;
    double mi = (pos + neg) / (double)n_mi;
    if (neg * ((double)n_neg / (double)n_mi) > pos * ((double)n_pos / (double)n_mi)) mi *= -1.0;
    r.stat = mi; r.df = df; r.pval = mi_pval_dev(fabs(mi), df, n_obs); r.suff = true;
    return r;
}

// cnt: MI_BIN_WARP_BYTES of shared memory per warp: 32 subsets x 33 ints of counts (row stride 33: the cell-per-lane stores and
// the subset-per-lane loads are both conflict-free) + 32 x 17 doubles of scratch.  All threads of the CTA call this with identical arguments.
constexpr int MI_BIN_CNT_LD = 33, MI_BIN_UV_LD = 17;
constexpr int MI_BIN_WARP_BYTES = 32 * MI_BIN_CNT_LD * 4 + 32 * MI_BIN_UV_LD * 8;
template <int THREADS>
__device__ void eval_subsets_mi_bin(const unsigned int* __restrict__ planes, int W, unsigned int tail_mask, const i64* var, int xs, int ys, const int* acc, int m, int max_k, double alpha, i64 max_tests,
                                    i64 hps, i64* tri_off, int* cnt, EvalShared* sh, EvalOut* out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = THREADS / 32;
    const unsigned full = 0xffffffffu;
    const SubsetCounts sc = subset_counts(m, max_k);
    if (sc.c3 > 0) for (int i = tid; i <= m; i += THREADS) tri_off[i] = sc.c3 - choose3(m - i);
    if (tid == 0) sh->fail_idx = (u64)FW_INF_IDX;
    __syncthreads();
    const i64 limit = (max_tests > 0 && max_tests < sc.total) ? max_tests : sc.total;
    const i64 X = var[xs], Y = var[ys];
    int* wcnt = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(cnt) + (size_t)warp * MI_BIN_WARP_BYTES);
    double* wuv = reinterpret_cast<double*>(wcnt + 32 * MI_BIN_CNT_LD);
    i64 my_fail = FW_INF_IDX; double f_stat = 0.0, f_p = 0.0; i64 f_df = 0; int f_suff = 0;
    i64 best_idx = -1; double b_stat = 0.0, b_p = -1.0; i64 b_df = 0;
    i64 executed = 0;
    bool any_fail = false;
    int B = 1;                                                   // subsets per warp in this chunk
    for (i64 base = 0; base < limit;) {
        const i64 my_idx = base + (i64)warp * B + lane;           // the subset this lane will evaluate
        const bool mine = lane < B && my_idx < limit;
        int k = 0, pa = 0, pb = 0, pc = 0;
        if (mine) unrank_subset(my_idx, m, sc, tri_off, k, pa, pb, pc);
        // counting: the warp walks through its batch
        for (int j = 0; j < B; ++j) {
            const i64 idx = base + (i64)warp * B + j;
            if (idx >= limit) break;                             // warp-uniform
            const int kj = __shfl_sync(full, k, j);
            const i64 z0 = var[acc[__shfl_sync(full, pa, j)]];
            const i64 z1 = var[acc[__shfl_sync(full, pb, j)]];   // positions beyond k are 0: harmless valid members
            const i64 z2 = var[acc[__shfl_sync(full, pc, j)]];
            if (kj == 3) mi_count_bin_to_smem<3>(planes, W, tail_mask, X, Y, z0, z1, z2, wcnt + j * MI_BIN_CNT_LD);
            else if (kj == 2) mi_count_bin_to_smem<2>(planes, W, tail_mask, X, Y, z0, z1, z2, wcnt + j * MI_BIN_CNT_LD);
            else mi_count_bin_to_smem<1>(planes, W, tail_mask, X, Y, z0, z1, z2, wcnt + j * MI_BIN_CNT_LD);
        }
        __syncwarp();
        if (mine) {
            const MiResult r = mi_stats_bin_thread(wcnt + lane * MI_BIN_CNT_LD, wuv + lane * MI_BIN_UV_LD, 1 << k, hps);
            const bool sig = (r.pval < alpha) && r.suff;
            const bool stop = !sig || (max_tests > 0 && my_idx + 1 >= max_tests);
            if (stop) { my_fail = my_idx; f_stat = r.stat; f_p = r.pval; f_df = r.df; f_suff = r.suff ? 1 : 0; }
            else if (r.pval >= b_p) { best_idx = my_idx; b_stat = r.stat; b_p = r.pval; b_df = r.df; }
        }
        __syncwarp();
        const i64 end = base + (i64)NW * B;
        executed = end < limit ? end : limit;
        any_fail = __syncthreads_or(my_fail != FW_INF_IDX);
        if (any_fail) break;
        base = end; B = B == 1 ? 8 : 32;
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

struct HitonMiArgs {
    MiTable t; i64 hps;
    const i64* uni_off; const i64* uni_nbr; const double* uni_stat; const double* uni_p;
    const i64* targets; const int* sel; int n_sel; const i64* out_off; int* counter;
    int max_k; double alpha; i64 max_tests; int cap;
    int* cand_order;
    i64* pc_nbr; double* pc_stat; double* pc_p; i64* pc_count;
    i64* tpc_nbr; double* tpc_stat; double* tpc_p; i64* tpc_count;
    i64* num_tests; u64* executed_total; int* status;
    HitonLists lists;                        // whitelists / blacklists / rejection records (all optional; subsets.cuh)
    int lane_ok, Wp;                         // binary tables: the planes of `cap` slots fit shared memory (row stride Wp, odd): one lane per test (mi_lane.cuh)
};

#ifndef FW_MI_MINB
#define FW_MI_MINB 2
#endif
template <int THREADS, int TPT>
__global__ void __launch_bounds__(THREADS, FW_MI_MINB) hiton_mi_kernel(HitonMiArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int cap = a.cap, L = a.t.L;
    const int tab_ints = L * L * L * L * L;
    size_t o = 0;
    i64* tri_off = reinterpret_cast<i64*>(smem + o); o += sizeof(i64) * (cap + 1);
    double* tpc_stat = reinterpret_cast<double*>(smem + o); o += sizeof(double) * cap;
    double* tpc_p = reinterpret_cast<double*>(smem + o); o += sizeof(double) * cap;
    double* pcs_stat = reinterpret_cast<double*>(smem + o); o += sizeof(double) * cap;
    double* pcs_p = reinterpret_cast<double*>(smem + o); o += sizeof(double) * cap;
    i64* var = reinterpret_cast<i64*>(smem + o); o += sizeof(i64) * cap;           // slot -> variable: 0 = T, 1..M members, M+1 candidate
    int* acc = reinterpret_cast<int*>(smem + o); o += sizeof(int) * cap;
    int* pc_slot = reinterpret_cast<int*>(smem + o); o += sizeof(int) * cap;
    unsigned char* sflag = smem + o; o += ((size_t)cap + 15) & ~(size_t)15;          // per slot: list flags of the member
    int* tabs = reinterpret_cast<int*>(smem + o);
    int* tab = tabs + warp * tab_ints;
    int* cnt = tabs + (THREADS / 32) * tab_ints;                                     // batched binary scan: [warps][32][32] ints
    const bool bin_table = (L == 2 && !a.t.nz);
    // one-lane-per-test scan (mi_lane.cuh): staged planes [cap][Wp] + count tables; shares the region of the batched scan's count buffers
    const bool use_lane = bin_table && a.lane_ok;
    const int Wp = a.Wp, W = a.t.W;
    unsigned int* sp = reinterpret_cast<unsigned int*>(cnt);                          // (the per-warp tables `tab` of the generic scan stay usable)
    int* ltabs = reinterpret_cast<int*>(sp + (size_t)cap * Wp);
    __shared__ EvalShared sh;
    __shared__ EvalOut ev;
    __shared__ int s_ti, s_nc, s_M, s_macc, s_npc, s_accept, s_nrej;
    __shared__ i64 s_ntests;
    __shared__ u64 s_exec, s_exk[3];

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
        if (tid == 0) { s_nc = 0; s_M = 0; s_ntests = 0; s_exec = 0; s_exk[0] = s_exk[1] = s_exk[2] = 0; var[0] = T; s_nrej = 0; }
        __syncthreads();
        const bool track = a.lists.rej_count != nullptr;
        if (use_lane) for (int w = tid; w < W; w += THREADS) sp[w] = __ldg(a.t.planes + (size_t)T * W + w);      // slot 0 = the target
        // hiton.jl:182-183,300-302: a discrete target with fewer than 2 levels has no neighbours
        if (a.t.levels[T] < 2) {
            if (tid == 0) { a.pc_count[tsel] = 0; a.tpc_count[tsel] = 0; a.num_tests[tsel] = 0; a.status[tsel] = 0; if (track) a.lists.rej_count[tsel] = 0; }
            continue;
        }
        // rejection record of the candidate just scanned (thread 0; hiton.jl:72-74): positions of `ev` refer to acc[0..)
        auto reject = [&](i64 cand) {
            const i64 r = o0 + s_nrej;
            a.lists.rej_nbr[r] = cand; a.lists.rej_k[r] = ev.k;
            for (int i = 0; i < 3; ++i) a.lists.rej_Zs[r * 3 + i] = i < ev.k ? var[acc[ev.pos[i]]] : -1;
            a.lists.rej_res[r] = make_result(ev.stat, ev.pval, ev.df, ev.suff != 0);
            a.lists.rej_ntests[r] = ev.num_tests; a.lists.rej_frac[r] = ev.total > 0 ? (double)ev.num_tests / (double)ev.total : 0.0;
            s_nrej = s_nrej + 1;
        };
        for (int i = tid; i < n_uni; i += THREADS) {
            double pi = a.uni_p[e0 + i];
            if (pi < a.alpha) {
                int rank = 0;
                for (int j = 0; j < n_uni; ++j) {
                    double pj = a.uni_p[e0 + j];
                    if (pj < a.alpha && (pj < pi || (pj == pi && j < i))) ++rank;
                }
                order[rank] = i | (hiton_list_flags(a.lists, tsel, a.uni_nbr[e0 + i]) << 28);
                atomicAdd(&s_nc, 1);
            }
        }
        __syncthreads();
        const int n_c = s_nc;
        bool overflow = false;
        // ---- interleaving phase ----
        for (int ci = 0; ci < n_c; ++ci) {
            const int M = s_M;
            if (M + 2 > cap) { overflow = true; break; }
            const int ui = order[ci] & HITON_ORDER_MASK, lf = order[ci] >> 28;
            if (lf == 2) continue;                     // blacklisted (and not whitelisted): skipped untested (hiton.jl:31-34)
            const i64 cand = a.uni_nbr[e0 + ui];
            const int ys = M + 1;
            if (tid == 0) { var[ys] = cand; s_accept = 0; }
            if (use_lane) for (int w = tid; w < W; w += THREADS) sp[(size_t)ys * Wp + w] = __ldg(a.t.planes + (size_t)cand * W + w);
            __syncthreads();
            if (lf & 1) {
                // whitelisted: accepted untested with (NaN, NaN) (hiton.jl:20-29)
                if (tid == 0) { tpc_stat[M] = __longlong_as_double(0x7ff8000000000000LL); tpc_p[M] = tpc_stat[M]; s_accept = 1; }
            } else if (M == 0) {
                if (tid == 0) { tpc_stat[0] = a.uni_stat[e0 + ui]; tpc_p[0] = a.uni_p[e0 + ui]; s_accept = 1; }
            } else {
                for (int s = tid; s < M; s += THREADS) acc[s] = s + 1;
                __syncthreads();
                MiSlotTest tf; tf.t = a.t; tf.var = var; tf.x = 0; tf.y = ys; tf.hps = a.hps; tf.tab = tab;
                if (use_lane && M <= MI_LANE_MAX_M && a.t.levels[T] == 2 && a.t.levels[cand] == 2)
                    eval_subsets_mi_lane<THREADS>(sp, Wp, W, a.t.n, 0, ys, acc, M, a.max_k, a.alpha, a.max_tests, a.hps, tri_off, ltabs, &sh, &ev, a.t.lgt);
                else if (bin_table && !use_lane && a.t.levels[T] == 2 && a.t.levels[cand] == 2)
                    eval_subsets_mi_bin<THREADS>(a.t.planes, a.t.W, a.t.tail_mask, var, 0, ys, acc, M, a.max_k, a.alpha, a.max_tests, a.hps, tri_off, cnt, &sh, &ev);
                else
                    eval_subsets<THREADS, TPT, 32>(tf, acc, M, a.max_k, a.alpha, a.max_tests, tri_off, &sh, &ev);
                if (tid == 0) {
                    s_ntests += ev.num_tests; s_exec += (u64)ev.executed; s_exk[0] += (u64)ev.ex_k[0]; s_exk[1] += (u64)ev.ex_k[1]; s_exk[2] += (u64)ev.ex_k[2];
                    if (ev.sig) { tpc_stat[M] = ev.stat; tpc_p[M] = ev.pval; s_accept = 1; }
                    else if (track) reject(cand);
                }
            }
            __syncthreads();
            if (s_accept) { if (tid == 0) { s_M = M + 1; sflag[M + 1] = (unsigned char)lf; } }      // var[M+1] already holds the candidate = member M
            __syncthreads();
        }
        if (overflow) { if (tid == 0) a.status[tsel] = 1; continue; }
        // ---- elimination phase ----
        const int M = s_M;
        for (int s = tid; s < M; s += THREADS) acc[s] = s + 1;
        if (tid == 0) { s_macc = M; s_npc = 0; }
        __syncthreads();
        for (int c = 1; c <= M; ++c) {
            const bool wl = (sflag[c] & 1) != 0;
            if (tid == 0) {
                // deleteat!(accepted, findall(in(candidate), accepted)) (hiton.jl:134-136); a whitelisted member stays and is pushed again
                if (!wl) { int w = 0, macc = s_macc; for (int j = 0; j < macc; ++j) { int v = acc[j]; if (v != c) acc[w++] = v; } s_macc = w; }
                s_accept = 0;
            }
            __syncthreads();
            const int macc = s_macc;
            if (macc + 2 > cap) { overflow = true; break; }              // duplicates of whitelisted members outgrew the class
            if (wl) {
                if (tid == 0) { pcs_stat[s_npc] = __longlong_as_double(0x7ff8000000000000LL); pcs_p[s_npc] = pcs_stat[s_npc]; s_accept = 1; }
            } else if (macc == 0) {
                if (tid == 0) { pcs_stat[s_npc] = tpc_stat[c - 1]; pcs_p[s_npc] = tpc_p[c - 1]; s_accept = 1; }
            } else {
                MiSlotTest tf; tf.t = a.t; tf.var = var; tf.x = 0; tf.y = c; tf.hps = a.hps; tf.tab = tab;
                if (use_lane && macc <= MI_LANE_MAX_M && a.t.levels[T] == 2 && a.t.levels[var[c]] == 2)
                    eval_subsets_mi_lane<THREADS>(sp, Wp, W, a.t.n, 0, c, acc, macc, a.max_k, a.alpha, a.max_tests, a.hps, tri_off, ltabs, &sh, &ev, a.t.lgt);
                else if (bin_table && !use_lane && a.t.levels[T] == 2 && a.t.levels[var[c]] == 2)
                    eval_subsets_mi_bin<THREADS>(a.t.planes, a.t.W, a.t.tail_mask, var, 0, c, acc, macc, a.max_k, a.alpha, a.max_tests, a.hps, tri_off, cnt, &sh, &ev);
                else
                    eval_subsets<THREADS, TPT, 32>(tf, acc, macc, a.max_k, a.alpha, a.max_tests, tri_off, &sh, &ev);
                if (tid == 0) {
                    s_ntests += ev.num_tests; s_exec += (u64)ev.executed; s_exk[0] += (u64)ev.ex_k[0]; s_exk[1] += (u64)ev.ex_k[1]; s_exk[2] += (u64)ev.ex_k[2];
                    if (ev.sig) { pcs_stat[s_npc] = ev.stat; pcs_p[s_npc] = ev.pval; s_accept = 1; }
                    else if (track) reject(var[c]);
                }
            }
            __syncthreads();
            if (tid == 0 && s_accept) { acc[s_macc] = c; s_macc = s_macc + 1; pc_slot[s_npc] = c; s_npc = s_npc + 1; }
            __syncthreads();
        }
        if (overflow) { if (tid == 0) a.status[tsel] = 1; continue; }
        const int npc = s_npc;
        for (int i = tid; i < npc; i += THREADS) {
            int c = pc_slot[i];
            double s = pcs_stat[i], pp = pcs_p[i];
            double ts = tpc_stat[c - 1], tp = tpc_p[c - 1];
            if (tp > pp || isnan(pp)) { s = ts; pp = tp; }
            if (a.pc_nbr) { a.pc_nbr[o0 + i] = var[c]; a.pc_stat[o0 + i] = s; a.pc_p[o0 + i] = pp; }
        }
        for (int i = tid; i < M; i += THREADS) {
            if (a.tpc_nbr) { a.tpc_nbr[o0 + i] = var[i + 1]; a.tpc_stat[o0 + i] = tpc_stat[i]; a.tpc_p[o0 + i] = tpc_p[i]; }
        }
        if (tid == 0) {
            a.pc_count[tsel] = npc; a.tpc_count[tsel] = M; a.num_tests[tsel] = s_ntests; a.status[tsel] = 0;
            if (track) a.lists.rej_count[tsel] = s_nrej;
            atomicAdd(a.executed_total, s_exec);
            atomicAdd(a.executed_total + 1, s_exk[0]); atomicAdd(a.executed_total + 2, s_exk[1]); atomicAdd(a.executed_total + 3, s_exk[2]);
        }
    }
}

struct SubsetsMiArgs {
    MiTable t; i64 hps;
    const i64* X; const i64* Y; const i64* z_off; const i64* z_idx;
    int n_jobs; int* counter;
    int max_k; double alpha; i64 max_tests; int cap;
    DevResult* out; i64* out_Zs; int* out_k; i64* num_tests; double* frac; u64* executed_total;
};

template <int THREADS, int TPT>
__global__ void __launch_bounds__(THREADS) subsets_mi_kernel(SubsetsMiArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int cap = a.cap, L = a.t.L;
    const int tab_ints = L * L * L * L * L;
    size_t o = 0;
    i64* tri_off = reinterpret_cast<i64*>(smem + o); o += sizeof(i64) * (cap + 1);
    i64* var = reinterpret_cast<i64*>(smem + o); o += sizeof(i64) * cap;
    int* acc = reinterpret_cast<int*>(smem + o); o += sizeof(int) * cap;
    o = (o + 15) & ~(size_t)15;
    int* tab = reinterpret_cast<int*>(smem + o) + warp * tab_ints;
    int* cnt = reinterpret_cast<int*>(smem + o) + (THREADS / 32) * tab_ints;
    const bool bin_table = (L == 2 && !a.t.nz);
    __shared__ EvalShared sh;
    __shared__ EvalOut ev;
    __shared__ int s_ji;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_ji = atomicAdd(a.counter, 1);
        __syncthreads();
        const int job = s_ji;
        if (job >= a.n_jobs) break;
        const i64 z0 = a.z_off[job];
        const int m = (int)(a.z_off[job + 1] - z0);
        if (m == 0) continue;                                   // sentinel written by the host (tests.jl:285)
        for (int s = tid; s < m + 2; s += THREADS) var[s] = s == 0 ? a.X[job] : (s == 1 ? a.Y[job] : a.z_idx[z0 + s - 2]);
        for (int s = tid; s < m; s += THREADS) acc[s] = s + 2;
        __syncthreads();
        MiSlotTest tf; tf.t = a.t; tf.var = var; tf.x = 0; tf.y = 1; tf.hps = a.hps; tf.tab = tab;
        if (bin_table && a.t.levels[a.X[job]] == 2 && a.t.levels[a.Y[job]] == 2)
            eval_subsets_mi_bin<THREADS>(a.t.planes, a.t.W, a.t.tail_mask, var, 0, 1, acc, m, a.max_k, a.alpha, a.max_tests, a.hps, tri_off, cnt, &sh, &ev);
        else
            eval_subsets<THREADS, TPT, 32>(tf, acc, m, a.max_k, a.alpha, a.max_tests, tri_off, &sh, &ev);
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

// univariate test of two 2-level variables from the four counts, one thread (tests.jl:28-77, statfuns.jl:207-254; same
// formulas as mi_test_warp_bin<0>, terms accumulated in the order of the reference's loops `for i, for j`)
__device__ MiResult mi_pair_bin_thread(int n, int nx, int ny, int nxy, i64 hps, i64 n_obs_min) {
    MiResult r;
    if ((i64)n < n_obs_min || n == 0) { r.stat = 0.0; r.pval = 1.0; r.df = 0; r.suff = false; return r; }     // weak pre-check, tests.jl:9-20
    const i64 n_obs = n;
    if (!(!(n_obs < n_obs_min) && (((double)n_obs / 4.0) > (double)hps))) { r.stat = 0.0; r.pval = 1.0; r.df = 0; r.suff = false; return r; }   // tests.jl:58
    const int c11 = nxy, c10 = nx - nxy, c01 = ny - nxy, c00 = n - nx - ny + nxy;       // c[a = X level][b = Y level]
    const int ma0 = n - nx, ma1 = nx, mb0 = n - ny, mb1 = ny;
    double pos = 0.0, neg = 0.0; i64 n_pos = 0, n_neg = 0;
    if (c00 != 0) { pos += log((double)(n_obs * c00) / (double)((i64)ma0 * mb0)) * (double)c00; n_pos += c00; }
    if (c01 != 0) { neg += log((double)(n_obs * c01) / (double)((i64)ma0 * mb1)) * (double)c01; n_neg += c01; }
    if (c10 != 0) { neg += log((double)(n_obs * c10) / (double)((i64)ma1 * mb0)) * (double)c10; n_neg += c10; }
    if (c11 != 0) { pos += log((double)(n_obs * c11) / (double)((i64)ma1 * mb1)) * (double)c11; n_pos += c11; }
    const int df = (ma0 > 0 && ma1 > 0 && mb0 > 0 && mb1 > 0) ? 1 : 0;
    double mi = (pos + neg) / (double)n_obs;
    if (neg * ((double)n_neg / (double)n_obs) > pos * ((double)n_pos / (double)n_obs)) mi *= -1.0;
    r.stat = mi; r.df = df; r.pval = mi_pval_dev(fabs(mi), df, n_obs); r.suff = true;
    return r;
}

// ---- pairwise stage for the discrete kinds: one warp per pair, unordered emission of raw-significant pairs ----
// (pw_univar_kernel!, tests.jl:410-433: X-trimmed view when needs_nz_view(X); add_pwresults_to_matrix!, :391-407)
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) pw_mi_rows_kernel(MiTable t, i64 hps, i64 n_obs_min, double alpha, int reliable_only,
                                                                u64* counters /* [0] emitted, [1] reliable */, i64 cap,
                                                                int* c_x, int* c_y, double* c_p, double* c_stat, int sh_rank, int sh_world) {
    extern __shared__ int smem_tab[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int* tab = smem_tab + warp * (t.L * t.L);
    const i64 X = blockIdx.x;
    if (!pw_owns_group(X / PW_X_GROUP, sh_rank, sh_world)) return;
    i64 n_rel = 0;
    if (t.L == 2 && !t.nz && t.levels[X] == 2) {
        // binary table: one LANE per pair.  N(X=1,Y=1) = popc(X & Y) over the words is the only count that is not a per-variable
        // constant (N(X=1) = nnz[X], N(Y=1) = nnz[Y], N = n), so a lane walks its own Y plane (its cache line serves the next 31
        // words) against the broadcast X plane and then evaluates MI / df / p by itself - no reductions, no idle lanes.
        // Pairs whose Y is not a 2-level variable take the warp-cooperative generic test below.
        const unsigned int* px = t.planes + (size_t)X * t.W;
        const int nx = t.nnz[X];
        u64 n_rel_lane = 0;
        for (i64 y0 = X + 1 + (i64)warp * 32; y0 < t.p; y0 += (i64)WARPS * 32) {
            const i64 Y = y0 + lane;
            const bool in_range = Y < t.p;
            const bool bin_y = in_range && t.levels[Y] == 2;
            if (bin_y) {
                const unsigned int* py = t.planes + (size_t)Y * t.W;
                int nxy = 0;
                for (int w = 0; w < t.W; ++w) nxy += __popc(__ldg(px + w) & __ldg(py + w));
                const MiResult r = mi_pair_bin_thread(t.n, nx, t.nnz[Y], nxy, hps, n_obs_min);
                const bool rel = r.suff || !reliable_only;
                n_rel_lane += rel;
                if (rel && r.pval < alpha) {
                    u64 pos = atomicAdd(&counters[0], 1ull);
                    if ((i64)pos < cap) { c_x[pos] = (int)X; c_y[pos] = (int)Y; c_p[pos] = r.pval; c_stat[pos] = r.stat; }
                }
            }
            unsigned int rest = __ballot_sync(0xffffffffu, in_range && !bin_y);
            while (rest) {
                const int j = __ffs(rest) - 1; rest &= rest - 1;
                MiResult r = mi_test_warp(t, X, y0 + j, nullptr, 0, hps, n_obs_min, tab);
                const bool rel = r.suff || !reliable_only;
                if (lane == 0) {
                    n_rel_lane += rel;
                    if (rel && r.pval < alpha) {
                        u64 pos = atomicAdd(&counters[0], 1ull);
                        if ((i64)pos < cap) { c_x[pos] = (int)X; c_y[pos] = (int)(y0 + j); c_p[pos] = r.pval; c_stat[pos] = r.stat; }
                    }
                }
                __syncwarp();
            }
        }
        for (int o = 16; o > 0; o >>= 1) n_rel_lane += __shfl_xor_sync(0xffffffffu, n_rel_lane, o);
        if (lane == 0 && n_rel_lane) atomicAdd(&counters[1], n_rel_lane);
        return;
    }
    for (i64 Y = X + 1 + warp; Y < t.p; Y += WARPS) {
        MiResult r = mi_test_warp(t, X, Y, nullptr, 0, hps, n_obs_min, tab);
        // unreliable tests become NaN and are excluded from BH's m (tests.jl:397-402, 521-526)
        const bool rel = r.suff || !reliable_only;
        n_rel += rel;
        if (rel && r.pval < alpha && lane == 0) {
            u64 pos = atomicAdd(&counters[0], 1ull);
            if ((i64)pos < cap) { c_x[pos] = (int)X; c_y[pos] = (int)Y; c_p[pos] = r.pval; c_stat[pos] = r.stat; }
        }
        __syncwarp();
    }
    if (lane == 0 && n_rel) atomicAdd(&counters[1], (u64)n_rel);
}

// ---- pw_univar_neighbors for the table-based kinds (tests.jl:436-532) in two halves ------------------------------------------------
//   collect:      all pairs (X, Y > X) with X in this rank's X groups -> unordered raw-significant records + the number of reliable tests
//   order_finish: records (of ALL ranks) -> condensed-index order (x, then y ascending) -> BH with the global m -> neighbour CSR
// One GPU: both halves back to back (pairwise_*_run).  Several GPUs: fw_pairwise_partial / host all-gather / fw_pairwise_merge.
struct PwCollected { int* c_x = nullptr; int* c_y = nullptr; double* c_p = nullptr; double* c_stat = nullptr; i64 nf = 0, n_rel = 0; };

static cudaError_t pairwise_mi_collect(PairwiseScratch& S, const MiTable& t, i64 hps, i64 n_obs_min, double alpha, bool reliable_only, int sh_rank, int sh_world,
                                       cudaStream_t st, PwCollected* pc, int* n_launch, std::string* msg) {
    const int WARPS = 8;
    const i64 p = t.p, n_pairs = p * (p - 1) / 2;
    u64* counters; PWCK(S.get(0, sizeof(u64) * 4, (void**)&counters), "alloc");
    i64 cap = std::max<i64>((i64)1 << 16, std::min<i64>(n_pairs, n_pairs / 8 + 1024));
    u64 h_cnt[2] = {0, 0};
    for (int attempt = 0; attempt < 2; ++attempt) {
        PWCK(S.get(7, sizeof(int) * cap, (void**)&pc->c_x), "alloc");
        PWCK(S.get(8, sizeof(int) * cap, (void**)&pc->c_y), "alloc");
        PWCK(S.get(9, sizeof(double) * cap, (void**)&pc->c_p), "alloc");
        PWCK(S.get(10, sizeof(double) * cap, (void**)&pc->c_stat), "alloc");
        PWCK(cudaMemsetAsync(counters, 0, sizeof(u64) * 4, st), "memset");
        pw_mi_rows_kernel<WARPS><<<(unsigned)p, WARPS * 32, WARPS * t.L * t.L * sizeof(int), st>>>(t, hps, n_obs_min, alpha, reliable_only ? 1 : 0, counters, cap,
                                                                                             pc->c_x, pc->c_y, pc->c_p, pc->c_stat, sh_rank, sh_world);
        (*n_launch)++;
        PWCK(cudaGetLastError(), "pw_mi_rows_kernel");
        PWCK(cudaMemcpyAsync(h_cnt, counters, sizeof(u64) * 2, cudaMemcpyDeviceToHost, st), "d2h");
        PWCK(cudaStreamSynchronize(st), "sync");
        if ((i64)h_cnt[0] <= cap) break;
        cap = (i64)h_cnt[0];
    }
    pc->nf = (i64)h_cnt[0]; pc->n_rel = (i64)h_cnt[1];
    return cudaSuccess;
}

// fz_nz: per-pair correlations on the co-non-zero rows.  Default: tensor-core pre-filter (fznz_tc.cuh) + exact fp64 test of the
// candidate pairs; FWGPU_FZNZ_TC=0 runs the exact test on every pair instead - same neighbour lists, used by the tests to check the pre-filter.
static bool fznz_use_tc() {
    const char* e = getenv("FWGPU_FZNZ_TC");       // read on every call: the tests switch between the two paths
    return e ? atoi(e) != 0 : true;
}
static cudaError_t pairwise_fznz_collect(PairwiseScratch& S, fznztc::Planes& planes, const NzTable& t, i64 n_obs_min, double alpha, bool reliable_only,
                                         int sh_rank, int sh_world, cudaStream_t st, PwCollected* pc, int* n_launch, std::string* msg) {
    const int WARPS = 8;
    const i64 p = t.p, n_pairs = p * (p - 1) / 2;
    u64* counters; PWCK(S.get(0, sizeof(u64) * 4, (void**)&counters), "alloc");
    u64 h_cnt[3] = {0, 0, 0};
    if (fznz_use_tc() && t.n >= 64 && p >= 2) {
        i64 ccap = std::max<i64>((i64)1 << 16, std::min<i64>(n_pairs, n_pairs / 16 + 1024));
        int *k_x, *k_y;
        for (int attempt = 0; attempt < 2; ++attempt) {
            PWCK(S.get(16, sizeof(int) * ccap, (void**)&k_x), "alloc");
            PWCK(S.get(17, sizeof(int) * ccap, (void**)&k_y), "alloc");
            PWCK(cudaMemsetAsync(counters, 0, sizeof(u64) * 4, st), "memset");
            { cudaError_t e_ = fznztc::run_prefilter(planes, t, n_obs_min, alpha, reliable_only, counters, ccap, k_x, k_y, attempt > 0, st, n_launch, msg, sh_rank, sh_world); if (e_ != cudaSuccess) return e_; }
            PWCK(cudaMemcpyAsync(h_cnt, counters, sizeof(u64) * 3, cudaMemcpyDeviceToHost, st), "d2h");
            PWCK(cudaStreamSynchronize(st), "sync (fznz_prefilter_kernel)");
            if ((i64)h_cnt[0] <= ccap) break;
            ccap = (i64)h_cnt[0];
        }
        const i64 n_cand = (i64)h_cnt[0];
        if (getenv("FWGPU_VERBOSE")) fprintf(stderr, "[fwgpu] fz_nz pairwise: %lld candidates of %lld pairs after the tensor-core pre-filter (rank %d of %d)\n", (long long)n_cand, (long long)n_pairs, sh_rank, sh_world);
        const i64 cap = std::max<i64>(n_cand, 16);
        PWCK(S.get(7, sizeof(int) * cap, (void**)&pc->c_x), "alloc");
        PWCK(S.get(8, sizeof(int) * cap, (void**)&pc->c_y), "alloc");
        PWCK(S.get(9, sizeof(double) * cap, (void**)&pc->c_p), "alloc");
        PWCK(S.get(10, sizeof(double) * cap, (void**)&pc->c_stat), "alloc");
        if (n_cand > 0) {
            if (n_cand >= ((i64)1 << 31) - 1) { if (msg) *msg = "more than 2^31 candidate pairs are not supported (unsupported size)"; return cudaErrorInvalidValue; }
            // sort the candidates by (X tile, Y tile, X, Y): see fznz_cand_keys_kernel
            auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
            unsigned char* karena; PWCK(S.get(18, 2 * al(sizeof(u64) * n_cand), (void**)&karena), "alloc");
            u64* keys = (u64*)karena; u64* keys2 = (u64*)(karena + al(sizeof(u64) * n_cand));
            fznztc::fznz_cand_keys_kernel<<<pw_blocks(n_cand, 256), 256, 0, st>>>(k_x, k_y, n_cand, keys); (*n_launch)++;
            int tb = 1; while (((u64)1 << tb) <= (u64)((p + 127) >> 7) && tb < 20) ++tb;
            void* tmp = nullptr; size_t need = 0;
            cub::DeviceRadixSort::SortKeys(nullptr, need, keys, keys2, (int)n_cand, 0, 33 + tb, st);
            PWCK(S.get(5, need, &tmp), "alloc");
            PWCK(cub::DeviceRadixSort::SortKeys(tmp, need, keys, keys2, (int)n_cand, 0, 33 + tb, st), "sort (candidates)"); (*n_launch) += 7;
            const i64 blocks = std::min<i64>((n_cand + WARPS * 4 - 1) / (WARPS * 4), (i64)148 * 32);
            const size_t wmb = fznz_warp_scratch_bytes(WARPS, t.W);
            if (wmb > 48 * 1024) PWCK(cudaFuncSetAttribute(fznztc::fznz_candidates_kernel<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wmb), "cudaFuncSetAttribute(fznz_candidates_kernel)");
            fznztc::fznz_candidates_kernel<WARPS><<<(unsigned)blocks, WARPS * 32, wmb, st>>>(t, n_cand, keys2, n_obs_min, alpha, reliable_only ? 1 : 0,
                                                                                         counters, cap, pc->c_x, pc->c_y, pc->c_p, pc->c_stat, wmb ? 1 : 0);
            (*n_launch)++;
            PWCK(cudaGetLastError(), "fznz_candidates_kernel");
        }
        PWCK(cudaMemcpyAsync(h_cnt, counters, sizeof(u64) * 3, cudaMemcpyDeviceToHost, st), "d2h");
        PWCK(cudaStreamSynchronize(st), "sync (fznz_candidates_kernel)");
        h_cnt[0] = h_cnt[2];                                             // raw-significant pairs
    } else {
        i64 cap = std::max<i64>((i64)1 << 16, std::min<i64>(n_pairs, n_pairs / 8 + 1024));
        for (int attempt = 0; attempt < 2; ++attempt) {
            PWCK(S.get(7, sizeof(int) * cap, (void**)&pc->c_x), "alloc");
            PWCK(S.get(8, sizeof(int) * cap, (void**)&pc->c_y), "alloc");
            PWCK(S.get(9, sizeof(double) * cap, (void**)&pc->c_p), "alloc");
            PWCK(S.get(10, sizeof(double) * cap, (void**)&pc->c_stat), "alloc");
            PWCK(cudaMemsetAsync(counters, 0, sizeof(u64) * 4, st), "memset");
            const size_t wmb = fznz_warp_scratch_bytes(WARPS, t.W);
            if (wmb > 48 * 1024) PWCK(cudaFuncSetAttribute(pw_fznz_rows_kernel<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wmb), "cudaFuncSetAttribute(pw_fznz_rows_kernel)");
            pw_fznz_rows_kernel<WARPS><<<(unsigned)p, WARPS * 32, wmb, st>>>(t, n_obs_min, alpha, reliable_only ? 1 : 0, counters, cap, pc->c_x, pc->c_y, pc->c_p, pc->c_stat, sh_rank, sh_world, wmb ? 1 : 0);
            (*n_launch)++;
            PWCK(cudaGetLastError(), "pw_fznz_rows_kernel");
            PWCK(cudaMemcpyAsync(h_cnt, counters, sizeof(u64) * 2, cudaMemcpyDeviceToHost, st), "d2h");
            PWCK(cudaStreamSynchronize(st), "sync");
            if ((i64)h_cnt[0] <= cap) break;
            cap = (i64)h_cnt[0];
        }
    }
    pc->nf = (i64)h_cnt[0]; pc->n_rel = (i64)h_cnt[1];
    return cudaSuccess;
}

// records in any order -> condensed-index order (sort by x*p + y, gather) -> BH (m tests) + neighbour CSR
static cudaError_t pairwise_order_finish(PairwiseScratch& S, const int* c_x, const int* c_y, const double* c_p, const double* c_stat, i64 nf, i64 m, i64 p,
                                         double alpha, bool fdr, cudaStream_t st, PairwiseOut* out, int* n_launch, std::string* msg) {
    const int T = 256;
    i64* d_off; PWCK(S.get(6, sizeof(i64) * (p + 1), (void**)&d_off), "alloc");
    out->d_off = d_off;
    if (nf == 0) {
        PWCK(cudaMemsetAsync(d_off, 0, sizeof(i64) * (p + 1), st), "memset");
        out->n_entries = 0; out->d_nbr = nullptr; out->d_stat = nullptr; out->d_adjp = nullptr;
        return cudaSuccess;
    }
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    unsigned char* arena; PWCK(S.get(11, 2 * al(sizeof(u64) * nf) + 2 * al(sizeof(unsigned int) * nf), (void**)&arena), "alloc");
    u64* keys = (u64*)arena; u64* keys2 = (u64*)(arena + al(sizeof(u64) * nf));
    unsigned int* vals = (unsigned int*)(arena + 2 * al(sizeof(u64) * nf)); unsigned int* vals2 = (unsigned int*)(arena + 2 * al(sizeof(u64) * nf) + al(sizeof(unsigned int) * nf));
    pw_pair_keys<<<pw_blocks(nf, T), T, 0, st>>>(c_x, c_y, p, keys, vals, nf); (*n_launch)++;
    int end_bit = 1; while (((u64)1 << end_bit) < (u64)p * (u64)p && end_bit < 64) ++end_bit;
    void* tmp = nullptr; size_t need = 0;
    if (nf >= ((i64)1 << 31) - 1) { if (msg) *msg = "more than 2^31 raw-significant pairs are not supported (unsupported size)"; return cudaErrorInvalidValue; }
    cub::DeviceRadixSort::SortPairs(nullptr, need, keys, keys2, vals, vals2, (int)nf, 0, end_bit, st);
    PWCK(S.get(5, need, &tmp), "alloc");
    PWCK(cub::DeviceRadixSort::SortPairs(tmp, need, keys, keys2, vals, vals2, (int)nf, 0, end_bit, st), "sort"); (*n_launch) += 8;
    unsigned char* ord; PWCK(S.get(15, 2 * al(sizeof(int) * nf) + 2 * al(sizeof(double) * nf), (void**)&ord), "alloc");
    int* o_x = (int*)ord; int* o_y = (int*)(ord + al(sizeof(int) * nf));
    double* o_p = (double*)(ord + 2 * al(sizeof(int) * nf)); double* o_s = (double*)(ord + 2 * al(sizeof(int) * nf) + al(sizeof(double) * nf));
    pw_gather_pairs<<<pw_blocks(nf, T), T, 0, st>>>(vals2, c_x, c_y, c_p, c_stat, o_x, o_y, o_p, o_s, nf); (*n_launch)++;
    PWCK(cudaGetLastError(), "pw_gather_pairs");
    PWCK(cudaStreamSynchronize(st), "sync");      // the arena in slot 11 is re-used by pairwise_finish
    return pairwise_finish(S, o_x, o_y, o_p, o_s, nf, m, p, alpha, fdr, st, out, n_launch, msg);
}

static cudaError_t pairwise_mi_run(PairwiseScratch& S, const MiTable& t, i64 hps, i64 n_obs_min, double alpha, bool fdr, bool reliable_only,
                                   cudaStream_t st, PairwiseOut* out, int* n_launch, std::string* msg) {
    const i64 p = t.p, n_pairs = p * (p - 1) / 2;
    PwCollected pc;
    cudaError_t e = pairwise_mi_collect(S, t, hps, n_obs_min, alpha, reliable_only, 0, 1, st, &pc, n_launch, msg);
    if (e != cudaSuccess) return e;
    out->n_tests = n_pairs; out->n_raw_sig = pc.nf;
    out->n_reliable = reliable_only ? pc.n_rel : n_pairs;
    const i64 m = reliable_only ? pc.n_rel : n_pairs;             // tests.jl:521-526
    return pairwise_order_finish(S, pc.c_x, pc.c_y, pc.c_p, pc.c_stat, pc.nf, m, p, alpha, fdr, st, out, n_launch, msg);
}

static cudaError_t pairwise_fznz_run(PairwiseScratch& S, fznztc::Planes& planes, const NzTable& t, i64 n_obs_min, double alpha, bool fdr, bool reliable_only,
                                     cudaStream_t st, PairwiseOut* out, int* n_launch, std::string* msg) {
    const i64 p = t.p, n_pairs = p * (p - 1) / 2;
    PwCollected pc;
    cudaError_t e = pairwise_fznz_collect(S, planes, t, n_obs_min, alpha, reliable_only, 0, 1, st, &pc, n_launch, msg);
    if (e != cudaSuccess) return e;
    out->n_tests = n_pairs; out->n_raw_sig = pc.nf;
    out->n_reliable = reliable_only ? pc.n_rel : n_pairs;
    const i64 m = reliable_only ? pc.n_rel : n_pairs;
    return pairwise_order_finish(S, pc.c_x, pc.c_y, pc.c_p, pc.c_stat, pc.nf, m, p, alpha, fdr, st, out, n_launch, msg);
}
