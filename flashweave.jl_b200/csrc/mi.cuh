// mi.cuh — discrete (mutual-information / G^2) conditional-independence tests on bit planes.
//
// Replaces (reference paths relative to the FlashWeave.jl checkout):
//   contingency_table! dense 2-way / 3-way + level_map!   src/contingency.jl:7-17,42-56, src/misc.jl:162-184
//   nz_adjust_cont_tab / offset_levels                     src/statfuns.jl:307-323
//   mutual_information 2-D / 3-D, adjust_df, mi_pval        src/statfuns.jl:157-305
//   test(X,Y,data,::MiTest..) / test(X,Y,Zs,data,::MiTestCond,..)   src/tests.jl:28-92,184-229
//   get_levels / get_max_vals / needs_nz_view               src/misc.jl:64-107
//
// Layout: the level codes (0..L-1, L = maximum(max_vals)+1 <= 4) are stored as L-1 one-bit planes per
// variable, planes[v][lvl-1][W] with W = ceil(n/32) words (bit r of word w = row 32w+r has level lvl;
// rows >= n are 0 in every plane).  A contingency cell is popc(X_a & Y_b & Z-stratum mask): 32 rows per
// AND+POPC instead of one scattered increment per row, counts are exact integers, and the row trimming
// of the _nz kinds (hiton.jl:41-50,85; tests.jl:412-416) is one more AND with a "valid rows" mask instead
// of a materialised view.  One warp evaluates one test: lanes stride over the words, per-stratum cell
// counters live in registers, totals are formed with __reduce_add_sync, and the MI / df / chi^2 epilogue is
// spread over the lanes (one cell per lane-iteration, fp64 logs) and reduced with shuffles.
//
// The z-strata are indexed by the raw mixed-radix key; the reference maps keys to first-seen dense indices
// (level_map!), which only permutes the slices: MI, df and levels_z are invariant.
#pragma once
#include "common.cuh"

#define FW_MAX_L 4

struct MiTable {
    const unsigned int* planes;   // [p][L-1][W]
    const int* levels;            // per variable: number of distinct values (misc.jl:64-82)
    const int* max_vals;          // per variable: maximum value (misc.jl:84-97)
    const int* nnz;               // per variable: rows with a non-zero code
    i64 p; int n; int W; int L; int nz;   // nz: zero-adjusted kind (mi_nz)
    unsigned int tail_mask;       // valid bits of the last word
    const double* lgt;            // lgt[i] = log(i), i = 0..n (lgt[0] = 0): the cells and margins of a table are integers <= n (mi_lane.cuh)
    int sparse_sem;               // mi_nz: semantics of the reference's SPARSE-input code path (contingency.jl:182-258, 300-480), see mi_epilogue_warp
};

__global__ void mi_logtab_kernel(int n, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) out[i] = i > 0 ? log((double)i) : 0.0;
}

__device__ __forceinline__ const unsigned int* mi_plane(const MiTable& t, i64 v, int lvl /*1..L-1*/) {
    return t.planes + ((size_t)v * (t.L - 1) + (lvl - 1)) * (size_t)t.W;
}
// bits of word w where variable v has level `lvl` (0..L-1); `valid` masks the rows >= n
__device__ __forceinline__ unsigned int mi_level_word(const MiTable& t, i64 v, int lvl, int w, unsigned int valid) {
    if (lvl > 0) return __ldg(mi_plane(t, v, lvl) + w);
    unsigned int any = 0;
    for (int l = 1; l < t.L; ++l) any |= __ldg(mi_plane(t, v, l) + w);
    return ~any & valid;
}
__device__ __forceinline__ unsigned int mi_nonzero_word(const MiTable& t, i64 v, int w) {
    unsigned int any = 0;
    for (int l = 1; l < t.L; ++l) any |= __ldg(mi_plane(t, v, l) + w);
    return any;
}
// misc.jl:103-107 needs_nz_view for dense discrete data
__device__ __forceinline__ bool mi_needs_nz_view(const MiTable& t, i64 v) { return t.nz && t.levels[v] > 2; }

// ---- chi^2 upper tail for integer df: ccdf(Chisq(df), x) = Q(df/2, x/2) (statfuns.jl:157-161) ----------
// closed forms: even df: e^{-h} sum_{j<df/2} h^j/j!;  odd df: erfc(sqrt(h)) + e^{-h} sqrt(h) * sum_{j=0}^{(df-3)/2} h^j / Gamma(j+3/2)
__device__ double chisq_sf_dev(i64 df, double x) {
    if (!(x > 0.0)) return (x <= 0.0) ? 1.0 : x;       // NaN propagates
    const double h = 0.5 * x;
    if (isinf(x)) return 0.0;
    if ((df & 1) == 0) {
        double term = 1.0, sum = 1.0;
        for (i64 j = 1; j < df / 2; ++j) { term *= h / (double)j; sum += term; }
        return exp(-h) * sum;
    }
    double q = erfc(sqrt(h));
    if (df >= 3) {
        // term_0 = h^{1/2} / Gamma(3/2) = 2 sqrt(h/pi); term_{j} = term_{j-1} * h / (j + 1/2)
        double term = 2.0 * sqrt(h / 3.14159265358979323846), sum = term;
        for (i64 j = 1; j <= (df - 3) / 2; ++j) { term *= h / ((double)j + 0.5); sum += term; }
        q += exp(-h) * sum;
    }
    return q;
}
__device__ __forceinline__ double mi_pval_dev(double mi_abs, i64 df, i64 n_obs) {
    double g = 2.0 * mi_abs * (double)n_obs;
    return df > 0 ? chisq_sf_dev(df, g) : 1.0;
}

struct MiResult { double stat; double pval; i64 df; bool suff; };

// ---- warp-cooperative epilogue: MI, df, p from a dense count table in (per-warp) shared memory ---------
// tab[s*L*L + b*L + a] = N(X=a, Y=b, stratum s), S strata (S = 1: univariate).  All 32 lanes call this.
// Follows tests.jl:48-68 (univariate) / :200-221 (conditional) after the table has been built.
// sparse_k > 0: a conditional test with |Zs| = sparse_k under the reference's sparse-input semantics (its default tables for
// sensitive=false, learning.jl:470).  The table itself equals the dense one on the rows the test uses; what differs is levels_z in
// the power rule (tests.jl:210): the k = 1 specialisation indexes slices by the raw z value and reports max(z) + 1
// (contingency.jl:171-173, 229); the generic merge back-fills the rows it never visited (all-zero rows and, under Nz, the rows
// where X or Y is zero) into the stratum of the all-zero key, creating that stratum if no visited row had it (contingency.jl:461-477).
__device__ MiResult mi_epilogue_warp(const int* tab, int L, int S, int lvx, int lvy, int mvx, int mvy, int nz, bool conditional,
                                     i64 hps, i64 n_obs_min, int sparse_k = 0, i64 n_total = 0) {
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    int ox = 0, oy = 0, lx = lvx, ly = lvy, sx = L, sy = L;
    if (nz) { ox = mvx > 1 ? 1 : 0; oy = mvy > 1 ? 1 : 0; sx = L - ox; sy = L - oy; lx = sx; ly = sy; }   // statfuns.jl:307-323
    // levels_z (strata seen in the data view, all X/Y levels) and n_obs (sum of the sub-table over all slices)
    int lz_loc = 0; i64 nobs_loc = 0;
    for (int s = lane; s < S; s += 32) {
        const int* t = tab + s * L * L;
        int tot = 0, sub = 0;
        for (int b = 0; b < L; ++b) for (int a = 0; a < L; ++a) { int c = t[b * L + a]; tot += c; if (a >= ox && b >= oy) sub += c; }
        lz_loc += tot > 0; nobs_loc += sub;
    }
    int levels_z = conditional ? __reduce_add_sync(full, lz_loc) : 1;
    if (conditional && sparse_k > 0 && nz && (ox || oy)) {
        int hi = -1; i64 cnt = 0;
        for (int s = lane; s < S; s += 32) {
            const int* t = tab + s * L * L;
            int tot = 0;
            for (int e = 0; e < L * L; ++e) tot += t[e];
            if (tot > 0) hi = s;
            cnt += tot;
        }
        hi = __reduce_max_sync(full, hi);
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(full, cnt, o);
        if (sparse_k == 1) levels_z = hi + 1 > 1 ? hi + 1 : 1;
        else {
            int tot0 = 0;
            for (int e = 0; e < L * L; ++e) tot0 += tab[e];
            if (tot0 == 0 && n_total - cnt > 0) levels_z += 1;
        }
    }
    i64 n_obs = nobs_loc;
    for (int o = 16; o > 0; o >>= 1) n_obs += __shfl_xor_sync(full, n_obs, o);
    MiResult r;
    bool ok;
    if (conditional) ok = ((double)n_obs / (double)((i64)lx * ly * levels_z)) > (double)hps;                 // tests.jl:210 (n_obs_min not consulted)
    else ok = !(n_obs < n_obs_min) && (((double)n_obs / (double)((i64)lx * ly)) > (double)hps);             // tests.jl:58
    if (!ok) { r.stat = 0.0; r.pval = 1.0; r.df = 0; r.suff = false; return r; }
    // marginals are recomputed from the table by the lane that needs them (L <= 4: at most 4 adds each)
    double pos = 0.0, neg = 0.0; i64 n_pos = 0, n_neg = 0, df_loc = 0;
    const int cells = S * sx * sy;
    for (int e = lane; e < cells; e += 32) {
        const int s = e / (sx * sy), rem = e % (sx * sy), j = rem / sx, i = rem % sx;   // i, j: indices inside the (offset) view
        const int* t = tab + s * L * L;
        const int c = t[(j + oy) * L + (i + ox)];
        int mik = 0, mjk = 0, mk = 0;
        if (i < lx) for (int jj = 0; jj < ly; ++jj) mik += t[(jj + oy) * L + (i + ox)];
        if (j < ly) for (int ii = 0; ii < lx; ++ii) mjk += t[(j + oy) * L + (ii + ox)];
        for (int jj = 0; jj < ly; ++jj) for (int ii = 0; ii < lx; ++ii) mk += t[(jj + oy) * L + (ii + ox)];
        if (c != 0 && mik != 0 && mjk != 0) {
            // statfuns.jl:187 (3-D): log((marg_k*c)/(marg_ik*marg_jk))*c ; :232 (2-D): c*log((n_obs*c)/(marg_i*marg_j))
            const double num = conditional ? (double)((i64)mk * c) : (double)(n_obs * c);
            const double tt = log(num / (double)((i64)mik * mjk)) * (double)c;
            if (i == j) { pos += tt; n_pos += c; } else { neg += tt; n_neg += c; }
        }
    }
    // df: per stratum (#non-empty row margins - 1)(#non-empty col margins - 1), statfuns.jl:281-305
    for (int s = lane; s < S; s += 32) {
        const int* t = tab + s * L * L;
        int alx = 0, aly = 0;
        for (int i = 0; i < lx; ++i) { int m = 0; for (int jj = 0; jj < ly; ++jj) m += t[(jj + oy) * L + (i + ox)]; alx += m > 0; }
        for (int j = 0; j < ly; ++j) { int m = 0; for (int ii = 0; ii < lx; ++ii) m += t[(j + oy) * L + (ii + ox)]; aly += m > 0; }
        int tot = 0;
        for (int b = 0; b < L; ++b) for (int a = 0; a < L; ++a) tot += t[b * L + a];
        // the reference only visits k in 1:levels_z = the strata present in the data view
        if (!conditional || tot > 0) df_loc += (i64)(max(1, alx) - 1) * (max(1, aly) - 1);
    }
    for (int o = 16; o > 0; o >>= 1) {
        pos += __shfl_xor_sync(full, pos, o); neg += __shfl_xor_sync(full, neg, o);
        n_pos += __shfl_xor_sync(full, n_pos, o); n_neg += __shfl_xor_sync(full, n_neg, o);
        df_loc += __shfl_xor_sync(full, df_loc, o);
    }
    const i64 n_mi = conditional ? (n_pos + n_neg) : n_obs;            // statfuns.jl:197 vs :223
    double mi = (pos + neg) / (double)n_mi;
    if (neg * ((double)n_neg / (double)n_mi) > pos * ((double)n_pos / (double)n_mi)) mi *= -1.0;   // sign heuristic, statfuns.jl:202,249
    r.stat = mi; r.df = df_loc; r.pval = mi_pval_dev(fabs(mi), df_loc, n_obs); r.suff = true;
    return r;
}

// ---- warp-cooperative table build: N(X=a, Y=b, Z-stratum s) on the rows of the data view ---------------
// view = rows where X != 0 (if needs_nz_view(X)) and Y != 0 (if needs_nz_view(Y)); tab must hold L*L*L^k ints.
// extra_valid (may be null): an additional row mask of W words (the X-trimmed view of the pairwise stage is
// implied by X itself, so it is only used by callers that pass explicit row views).
__device__ void mi_count_warp(const MiTable& t, i64 X, i64 Y, const i64* Z, int k, bool trim_x, bool trim_y, int* tab) {
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    const int L = t.L;
    int S = 1; for (int j = 0; j < k; ++j) S *= L;
    // outer loop over strata; inner over this lane's words.  Counters for one stratum: L*L <= 16 registers.
    for (int s = 0; s < S; ++s) {
        int zl[3] = {0, 0, 0};
        { int q = s; for (int j = 0; j < k; ++j) { zl[j] = q % L; q /= L; } }       // key = sum z_j * L^j  (misc.jl:168-171)
        int cnt[FW_MAX_L * FW_MAX_L];
#pragma unroll
        for (int c = 0; c < FW_MAX_L * FW_MAX_L; ++c) cnt[c] = 0;
        for (int w = lane; w < t.W; w += 32) {
            const unsigned int valid = (w == t.W - 1) ? t.tail_mask : 0xffffffffu;
            unsigned int m = valid;
            if (trim_x) m &= mi_nonzero_word(t, X, w);
            if (trim_y) m &= mi_nonzero_word(t, Y, w);
            for (int j = 0; j < k; ++j) m &= mi_level_word(t, Z[j], zl[j], w, valid);
            if (m == 0) continue;
            unsigned int xw[FW_MAX_L], yw[FW_MAX_L];
            unsigned int anyx = 0, anyy = 0;
#pragma unroll
            for (int l = 1; l < FW_MAX_L; ++l) {
                xw[l] = (l < L) ? __ldg(mi_plane(t, X, l) + w) : 0u; anyx |= xw[l];
                yw[l] = (l < L) ? __ldg(mi_plane(t, Y, l) + w) : 0u; anyy |= yw[l];
            }
            xw[0] = ~anyx; yw[0] = ~anyy;
#pragma unroll
            for (int b = 0; b < FW_MAX_L; ++b)
#pragma unroll
                for (int a = 0; a < FW_MAX_L; ++a)
                    if (a < L && b < L) cnt[b * FW_MAX_L + a] += __popc(m & xw[a] & yw[b]);
        }
#pragma unroll
        for (int b = 0; b < FW_MAX_L; ++b)
#pragma unroll
            for (int a = 0; a < FW_MAX_L; ++a)
                if (a < L && b < L) {
                    int tot = __reduce_add_sync(full, cnt[b * FW_MAX_L + a]);
                    if (lane == 0) tab[s * L * L + b * L + a] = tot;
                }
    }
    __syncwarp();
}

// ---- binary fast path (L = 2, kind "mi", both variables with 2 levels) ------------------------------------------------------
// Words outer / strata inner: per word the 2^K stratum masks come from a mask tree over the Z planes, and each stratum needs
// only popc(m), popc(m&x), popc(m&y), popc(m&x&y) (the other cells follow by inclusion-exclusion).  After the warp reduction
// lane c owns cell (a = c&1, b = (c>>1)&1, stratum = c>>2); margins are 2 shuffles away, so the MI / df epilogue needs no
// shared-memory table.  Same integers, same formulas as the generic path (statfuns.jl:163-305).
template <int K>
__device__ MiResult mi_test_warp_bin(const MiTable& t, i64 X, i64 Y, const i64* Z, i64 hps, i64 n_obs_min) {
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    constexpr int S = 1 << K;
    int cn[S], cx[S], cy[S], cxy[S];
#pragma unroll
    for (int s = 0; s < S; ++s) { cn[s] = 0; cx[s] = 0; cy[s] = 0; cxy[s] = 0; }
    const unsigned int* px = t.planes + (size_t)X * t.W;
    const unsigned int* py = t.planes + (size_t)Y * t.W;
    const unsigned int* pz0 = K > 0 ? t.planes + (size_t)Z[0] * t.W : nullptr;
    const unsigned int* pz1 = K > 1 ? t.planes + (size_t)Z[1] * t.W : nullptr;
    const unsigned int* pz2 = K > 2 ? t.planes + (size_t)Z[2] * t.W : nullptr;
    for (int w = lane; w < t.W; w += 32) {
        const unsigned int valid = (w == t.W - 1) ? t.tail_mask : 0xffffffffu;
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
    int mine = 0;                                   // N(X = a, Y = b | stratum) for this lane's cell
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const int n = __reduce_add_sync(full, cn[s]), nx = __reduce_add_sync(full, cx[s]);
        const int ny = __reduce_add_sync(full, cy[s]), nxy = __reduce_add_sync(full, cxy[s]);
        if ((lane >> 2) == s) {
            const int a = lane & 1, b = (lane >> 1) & 1;
            mine = a ? (b ? nxy : nx - nxy) : (b ? ny - nxy : n - nx - ny + nxy);
        }
    }
    const int a = lane & 1, b = (lane >> 1) & 1;
    const int mik = mine + __shfl_xor_sync(full, mine, 2);       // sum over b
    const int mjk = mine + __shfl_xor_sync(full, mine, 1);       // sum over a
    const int mk = mik + __shfl_xor_sync(full, mik, 1);
    const i64 n_obs = __reduce_add_sync(full, mine);
    const unsigned int present = __ballot_sync(full, mk > 0 && (lane & 3) == 0);
    const int levels_z = K > 0 ? __popc(present) : 1;
    MiResult r;
    bool ok;
    if (K > 0) ok = ((double)n_obs / (double)(4 * levels_z)) > (double)hps;                       // tests.jl:210
    else ok = !(n_obs < n_obs_min) && (((double)n_obs / 4.0) > (double)hps);                      // tests.jl:58
    if (!ok) { r.stat = 0.0; r.pval = 1.0; r.df = 0; r.suff = false; return r; }
    double pos = 0.0, neg = 0.0; int n_pos = 0, n_neg = 0;
    if (mine != 0 && mik != 0 && mjk != 0) {
        const double num = K > 0 ? (double)((i64)mk * mine) : (double)(n_obs * mine);
        const double tt = log(num / (double)((i64)mik * mjk)) * (double)mine;
        if (a == b) { pos = tt; n_pos = mine; } else { neg = tt; n_neg = mine; }
    }
    for (int o = 16; o > 0; o >>= 1) { pos += __shfl_xor_sync(full, pos, o); neg += __shfl_xor_sync(full, neg, o); }
    n_pos = __reduce_add_sync(full, n_pos); n_neg = __reduce_add_sync(full, n_neg);
    // df: strata whose 2x2 slice has two non-empty rows and two non-empty columns (statfuns.jl:281-305)
    const unsigned int rows_nz = __ballot_sync(full, mik > 0 && b == 0);      // bits 4s + a
    const unsigned int cols_nz = __ballot_sync(full, mjk > 0 && a == 0);      // bits 4s + 2b
    int df = 0;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const unsigned int rb = (rows_nz >> (4 * s)) & 3u, cb = (cols_nz >> (4 * s)) & 5u;
        df += (rb == 3u && cb == 5u) ? 1 : 0;
    }
    const i64 n_mi = K > 0 ? (i64)(n_pos + n_neg) : n_obs;
    double mi = (pos + neg) / (double)n_mi;
    if (neg * ((double)n_neg / (double)n_mi) > pos * ((double)n_pos / (double)n_mi)) mi *= -1.0;
    r.stat = mi; r.df = df; r.pval = mi_pval_dev(fabs(mi), df, n_obs); r.suff = true;
    return r;
}

// full single test, one warp: tests.jl:28-77 (k = 0; the caller's X-trimmed view, tests.jl:412-416) and :184-229 (k >= 1;
// the view trimmed for X and Y as hiton.jl:41-50,85 does)
__device__ MiResult mi_test_warp(const MiTable& t, i64 X, i64 Y, const i64* Z, int k, i64 hps, i64 n_obs_min, int* tab) {
    MiResult r;
    const bool trim_x = mi_needs_nz_view(t, X);
    const int lvx = t.levels[X], lvy = t.levels[Y];
    if (t.L == 2 && !t.nz && lvx == 2 && lvy == 2) {
        if (k == 0) {
            // weak pre-check of tests.jl:9-20: rows / ((2-2)(2-2)) = Inf > hps unless there are no rows
            if ((i64)t.n < n_obs_min || t.n == 0) { r.stat = 0.0; r.pval = 1.0; r.df = 0; r.suff = false; return r; }
            return mi_test_warp_bin<0>(t, X, Y, Z, hps, n_obs_min);
        }
        if (k == 1) return mi_test_warp_bin<1>(t, X, Y, Z, hps, n_obs_min);
        if (k == 2) return mi_test_warp_bin<2>(t, X, Y, Z, hps, n_obs_min);
        return mi_test_warp_bin<3>(t, X, Y, Z, hps, n_obs_min);
    }
    if (k == 0) {
        // tests.jl:86-92 and the weak pre-check sufficient_power(X, Y, data, ...) of tests.jl:9-20 on the X-trimmed view
        if (lvx < 2) { r.stat = 0.0; r.pval = 1.0; r.df = 0; r.suff = false; return r; }
        const i64 rows = trim_x ? (i64)t.nnz[X] : (i64)t.n;
        bool pre = !(rows < n_obs_min);
        if (pre) {
            const int ox = lvx > 1 ? 2 : 1, oy = lvy > 1 ? 2 : 1;                  // offset_levels applied to *levels* (tests.jl:16)
            const double q = (double)rows / (double)((i64)(lvx - ox) * (lvy - oy));  // Inf / NaN semantics as in Julia
            pre = q > (double)hps;
        }
        if (!pre) { r.stat = 0.0; r.pval = 1.0; r.df = 0; r.suff = false; return r; }
        mi_count_warp(t, X, Y, Z, 0, trim_x, false, tab);
        return mi_epilogue_warp(tab, t.L, 1, lvx, lvy, t.max_vals[X], t.max_vals[Y], t.nz, false, hps, n_obs_min);
    }
    // rows of the test: the dense path trims the view for a variable with more than 2 levels (misc.jl:103-107, hiton.jl:41-50,85),
    // the sparse path skips the zero rows of a variable with max_val > 1 (contingency.jl:243-248)
    const bool sp = t.sparse_sem && t.nz;
    const bool tx = sp ? t.max_vals[X] > 1 : trim_x;
    const bool trim_y = sp ? t.max_vals[Y] > 1 : mi_needs_nz_view(t, Y);
    mi_count_warp(t, X, Y, Z, k, tx, trim_y, tab);
    int S = 1; for (int j = 0; j < k; ++j) S *= t.L;
    return mi_epilogue_warp(tab, t.L, S, lvx, lvy, t.max_vals[X], t.max_vals[Y], t.nz, true, hps, n_obs_min, sp ? k : 0, (i64)t.n);
}

// ---- table preparation ---------------------------------------------------------------------------------
// one warp per (variable, word): ballot of (code == lvl) over 32 rows
__global__ void mi_pack_planes_kernel(const int* __restrict__ data, i64 n, i64 ld, i64 p, int L, int W, unsigned int* __restrict__ planes) {
    const i64 gw = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= p * W) return;
    const i64 v = gw / W; const int w = (int)(gw % W);
    const i64 row = (i64)w * 32 + lane;
    const int code = row < n ? data[v * ld + row] : 0;
    for (int l = 1; l < L; ++l) {
        unsigned int b = __ballot_sync(0xffffffffu, code == l);
        if (lane == 0) planes[((size_t)v * (L - 1) + (l - 1)) * (size_t)W + w] = b;
    }
}
// one warp per variable: max value, set of values seen (codes 0..31), validity (0 <= code), non-zero rows
__global__ void mi_levels_kernel(const int* __restrict__ data, i64 n, i64 ld, i64 p, int* levels, int* max_vals, int* nnz, int* bad) {
    const i64 v = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (v >= p) return;
    unsigned int seen = 0; int mx = 0, nz = 0, neg = 0, big = 0;
    for (i64 i = lane; i < n; i += 32) {
        int c = data[v * ld + i];
        if (c < 0) neg = 1; else if (c > 31) big = 1; else seen |= 1u << c;
        mx = max(mx, c); nz += c != 0;
    }
    seen = __reduce_or_sync(0xffffffffu, seen); mx = __reduce_max_sync(0xffffffffu, mx); nz = __reduce_add_sync(0xffffffffu, nz);
    neg = __reduce_or_sync(0xffffffffu, (unsigned)neg); big = __reduce_or_sync(0xffffffffu, (unsigned)big);
    if (lane == 0) {
        levels[v] = __popc(seen); max_vals[v] = mx; nnz[v] = nz;
        if (neg) atomicOr(bad, 1);
        if (big) atomicOr(bad, 2);
    }
}

// independent tests, one warp each (fw_test_batch, kinds mi / mi_nz)
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) mi_test_batch_kernel(MiTable t, i64 n_tests, const i64* X, const i64* Y, const int* k, const i64* Zs,
                                                                   i64 hps, i64 n_obs_min, DevResult* out) {
    extern __shared__ int smem_tab[];
    int S = 1; for (int j = 0; j < 3; ++j) S *= t.L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int* tab = smem_tab + warp * (t.L * t.L * S);
    for (i64 tix = (i64)blockIdx.x * WARPS + warp; tix < n_tests; tix += (i64)gridDim.x * WARPS) {
        i64 Z[3] = {Zs[tix * 3], Zs[tix * 3 + 1], Zs[tix * 3 + 2]};
        MiResult r = mi_test_warp(t, X[tix], Y[tix], Z, k[tix], hps, n_obs_min, tab);
        if (lane == 0) out[tix] = make_result(r.stat, r.pval, r.df, r.suff);
        __syncwarp();
    }
}
