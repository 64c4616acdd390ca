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
// unchanged pcor_rec code of fz.cuh.
//
// Canonical summation (shared with oracle/fw_oracle.cpp::cor_view, operation for operation, so the Float32 correlations are
// bit-equal on both sides): every variable is shifted by its value in the FIRST row of the view; the raw moments
// G_ab = sum x'_a x'_b and S_a = sum x'_a are accumulated with fma in increasing row order in 8 interleaved partial sums
// (class = row index mod 8), the partials are added in order 0..7 starting from 0.0; c_ab = G_ab - (S_a*S_b)/n and
// r_ab = c_ab / (sqrt(c_aa)*sqrt(c_bb)) with individually rounded operations, clamped to [-1, 1] and rounded to Float32
// (cor_mat's eltype, learning.jl:127-129).  All three code paths below (register-blocked Gram, pair items of the large
// capacity classes, the univariate warp) produce exactly these partial sums.
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
// canonical correlation from the combined moments (see the header): NaN passes through
__device__ __forceinline__ double fznz_r_from_moments(double gaa, double gbb, double gab, double sa, double sb, double n) {
    const double caa = __dsub_rn(gaa, __ddiv_rn(__dmul_rn(sa, sa), n));
    const double cbb = __dsub_rn(gbb, __ddiv_rn(__dmul_rn(sb, sb), n));
    const double cab = __dsub_rn(gab, __ddiv_rn(__dmul_rn(sa, sb), n));
    double rr = __ddiv_rn(cab, __dmul_rn(__dsqrt_rn(caa), __dsqrt_rn(cbb)));
    if (rr > 1.0) rr = 1.0; else if (rr < -1.0) rr = -1.0;             // clampcor; NaN passes through
    return rr;
}
// One pass over the rows where X != 0 and Y != 0.  Lane (c = lane & 7, j = lane >> 3) accumulates class c (row mod 8) of one
// moment: j = 0: Sx (and Sy in a second register), 1: Sxx, 2: Syy, 3: Sxy, over the view rows of its class in increasing row
// order (the canonical summation order shared with the oracle).  Every lane walks the SET bits of its class only
// (the four moment lanes of a class move together): the view holds 10-30 % of the rows in FlashWeaveHE tables, and a loop over all
// 32 rows of every mask word spent most of its fp64 issue slots on rows outside the view.  wm: W words of per-warp scratch (shared
// memory) for the combined mask, or nullptr (the two masks are then re-read from global memory).
__device__ NzUni fznz_uni_warp(const NzTable& t, i64 X, i64 Y, i64 n_obs_min, unsigned int* wm) {
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    NzUni r;
    const i64 rows_x = t.nnz[X];
    if (rows_x < n_obs_min) { r.stat = 0.0; r.pval = 1.0; r.suff = (0 >= n_obs_min); return r; }     // tests.jl:111-115,159
    const float* x = t.data + X * t.ld; const float* y = t.data + Y * t.ld;
    const unsigned int* mx = t.nzmask + X * t.W; const unsigned int* my = t.nzmask + Y * t.W;
    int cnt = 0, first = 0x7fffffff;
    if (wm) __syncwarp();                                                // the previous pair's cursors are done with wm
    for (int w = lane; w < t.W; w += 32) {
        const unsigned int m = mx[w] & my[w];
        if (wm) wm[w] = m;
        cnt += __popc(m);
        if (m && first == 0x7fffffff) first = w * 32 + __ffs(m) - 1;
    }
    if (wm) __syncwarp();                                                // wm is complete and visible to every lane
    const i64 n_obs = __reduce_add_sync(full, cnt);
    first = __reduce_min_sync(full, first);
    double p_stat = 0.0;
    if (n_obs > 0 && n_obs >= n_obs_min) {
        const double cx = (double)x[first], cy = (double)y[first];
        const int c = lane & 7, j = lane >> 3;
        double acc = 0.0, acc2 = 0.0;
        // Class words: the membership bits of class c (rows c, c + 8, c + 16, ...) of 8 consecutive mask words gathered into one
        // 32-bit word, bit b <-> row 256 k + 8 b + c (increasing b = increasing row).  Built with uniform control flow (4 spaced bits
        // of a word -> one nibble by a multiply); the only divergent loop left is the walk over the set bits of a class word, whose
        // trip count differs between the 8 classes by their popcounts.  U rows per trip: the row indices come from the mask alone,
        // so the 2U loads of a trip are independent and in flight together; the moments are still accumulated row by row.
        constexpr int U = 4;
        for (int k = 0; k * 8 < t.W; ++k) {
            unsigned int cw = 0u;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int w = 8 * k + i;
                const unsigned int m = w < t.W ? (wm ? wm[w] : (mx[w] & my[w])) : 0u;
                cw |= ((((m >> c) & 0x01010101u) * 0x10204080u) >> 28) << (4 * i);
            }
            while (cw) {
                int rows[U]; int nr = 0;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (cw) { rows[u] = k * 256 + 8 * (__ffs(cw) - 1) + c; cw &= cw - 1u; nr = u + 1; } else rows[u] = first;
                }
                float xa[U], ya[U];
#pragma unroll
                for (int u = 0; u < U; ++u) { xa[u] = x[rows[u]]; ya[u] = y[rows[u]]; }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (u < nr) {
                        const double da = __dsub_rn((double)xa[u], cx), db = __dsub_rn((double)ya[u], cy);
                        const double uu = (j == 2) ? db : da;
                        const double vv = (j == 0) ? 1.0 : ((j == 1) ? da : db);
                        acc = fma(uu, vv, acc);
                        acc2 = fma(db, 1.0, acc2);
                    }
                }
            }
        }
        __syncwarp();
        double tot = 0.0, tot2 = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { tot += __shfl_sync(full, acc, (lane & 24) + k); tot2 += __shfl_sync(full, acc2, (lane & 24) + k); }
        const double sx = __shfl_sync(full, tot, 0), sy = __shfl_sync(full, tot2, 0);
        const double sxx = __shfl_sync(full, tot, 8), syy = __shfl_sync(full, tot, 16), sxy = __shfl_sync(full, tot, 24);
        const double rr = fznz_r_from_moments(sxx, syy, sxy, sx, sy, (double)n_obs);      // NaN passes through (tests.jl:143, :381)
        p_stat = (double)(float)rr;                                      // eltype of the data (Float32)
    }
    r.stat = p_stat;
    r.pval = fz_pval_dev(p_stat, nz_consts(n_obs, n_obs_min));
    r.suff = n_obs >= n_obs_min;
    return r;
}

// Four pairs per warp: each group of 8 lanes (g = lane >> 3) tests its own pair, lane c = lane & 7 accumulates ALL five moments of
// class c (one load / conversion / subtraction per row instead of one per moment lane).  The per-(class, moment) sums and their
// combination are the ones of fznz_uni_warp, bit for bit.  Control flow outside the set-bit walk is warp-uniform (an invalid or
// skipped group walks an empty mask).  wm: this GROUP's W words of scratch, or nullptr.
__device__ NzUni fznz_uni_g8(const NzTable& t, i64 X, i64 Y, i64 n_obs_min, unsigned int* wm, bool valid) {
    const int lane = threadIdx.x & 31, c = lane & 7, g0 = lane & 24;
    const unsigned full = 0xffffffffu;
    NzUni r; r.stat = 0.0; r.pval = 1.0; r.suff = false;
    if (!valid) { X = 0; Y = 0; }
    const bool short_x = valid && t.nnz[X] < n_obs_min;                                  // tests.jl:111-115,159
    const float* x = t.data + X * t.ld; const float* y = t.data + Y * t.ld;
    const unsigned int* mx = t.nzmask + X * t.W; const unsigned int* my = t.nzmask + Y * t.W;
    int cnt = 0, first = 0x7fffffff;
    __syncwarp();                                                        // the previous pairs are done with wm
    for (int w = c; w < t.W; w += 8) {
        const unsigned int m = (valid && !short_x) ? (mx[w] & my[w]) : 0u;
        if (wm) wm[w] = m;
        cnt += __popc(m);
        if (m && first == 0x7fffffff) first = w * 32 + __ffs(m) - 1;
    }
    __syncwarp();
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) { cnt += __shfl_xor_sync(full, cnt, o); first = min(first, __shfl_xor_sync(full, first, o)); }
    const i64 n_obs = cnt;
    const bool act = valid && !short_x && n_obs > 0 && n_obs >= n_obs_min;
    double sx = 0.0, sy = 0.0, sxx = 0.0, syy = 0.0, sxy = 0.0;
    const int frow = act ? first : 0;
    const double cx = (double)x[frow], cy = (double)y[frow];
    constexpr int U = 4;
    for (int k = 0; k * 8 < t.W; ++k) {
        unsigned int cw = 0u;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int w = 8 * k + i;
            const unsigned int m = (act && w < t.W) ? (wm ? wm[w] : (mx[w] & my[w])) : 0u;
            cw |= ((((m >> c) & 0x01010101u) * 0x10204080u) >> 28) << (4 * i);           // class word: see fznz_uni_warp
        }
        while (cw) {
            int rows[U]; int nr = 0;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (cw) { rows[u] = k * 256 + 8 * (__ffs(cw) - 1) + c; cw &= cw - 1u; nr = u + 1; } else rows[u] = frow;
            }
            float xa[U], ya[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { xa[u] = x[rows[u]]; ya[u] = y[rows[u]]; }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (u < nr) {
                    const double da = __dsub_rn((double)xa[u], cx), db = __dsub_rn((double)ya[u], cy);
                    sx = fma(da, 1.0, sx); sy = fma(db, 1.0, sy);
                    sxx = fma(da, da, sxx); syy = fma(db, db, syy); sxy = fma(da, db, sxy);
                }
            }
        }
    }
    __syncwarp();
    double tsx = 0.0, tsy = 0.0, tsxx = 0.0, tsyy = 0.0, tsxy = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        tsx += __shfl_sync(full, sx, g0 + k); tsy += __shfl_sync(full, sy, g0 + k);
        tsxx += __shfl_sync(full, sxx, g0 + k); tsyy += __shfl_sync(full, syy, g0 + k); tsxy += __shfl_sync(full, sxy, g0 + k);
    }
    if (!valid) return r;
    if (short_x) { r.stat = 0.0; r.pval = 1.0; r.suff = (0 >= n_obs_min); return r; }
    double p_stat = 0.0;
    if (act) {
        const double rr = fznz_r_from_moments(tsxx, tsyy, tsxy, tsx, tsy, (double)n_obs);  // NaN passes through (tests.jl:143, :381)
        p_stat = (double)(float)rr;                                      // eltype of the data (Float32)
    }
    r.stat = p_stat;
    r.pval = fz_pval_dev(p_stat, nz_consts(n_obs, n_obs_min));
    r.suff = n_obs >= n_obs_min;
    return r;
}

// ---- register-blocked Gram matrix of one job (nv <= 32 variables) ------------------------------------------------------------
// All moments of cor_subset! in ONE pass over the table: with v = (x_0 .. x_{nv-1}, 1) the upper triangle of G = sum over the
// view rows of v v^T holds every cross product, every sum of squares and every sum.  Only the VIEW rows are staged (a FlashWeaveHE
// view holds 20-45 % of the rows): a tile is the next FZNZ_CROWS view rows of each of the 8 row classes (class = row mod 8), found
// by warp c walking the class words of the mask (fznz_uni_warp); the values are gathered, converted to fp64 once and stored as
// [slot][33] in shared memory, the next tile's global reads are in flight while the current one is consumed.  Warp c consumes the
// slots of class c in increasing row order and each lane owns one B x B block of the upper triangle of G in fp64 registers
// (B = 4; more than 32 blocks - 29 to 33 entries of v - take a second pass), so a row costs 2B shared-memory reads and B*B DFMA
// per lane.  Partial sums of the warps (= classes) are added in warp order: the canonical summation order of the header.  Then
// r_ab = (G_ab - S_a S_b / n) / sqrt((G_aa - S_a^2/n)(G_bb - S_b^2/n)) in fp64, rounded to Float32 (cor_mat's eltype), NaN -> 0.
#ifndef FW_FZNZ_CROWS
#define FW_FZNZ_CROWS 16
#endif
constexpr int FZNZ_CROWS = FW_FZNZ_CROWS;                // view rows per class and tile
constexpr int FZNZ_TROWS = 8 * FZNZ_CROWS;              // slots of a tile
constexpr int FZNZ_TLD = 33;
constexpr int FZNZ_GMAX = 36;                           // row stride of G (9 blocks of 4)
constexpr int FZNZ_NVMAX = 32;                          // variables of a job on this path (+ the ones column = 33 = FZNZ_TLD)
constexpr int FZNZ_GRAM_BYTES = (FZNZ_TROWS * FZNZ_TLD + 8) * 8 + FZNZ_GMAX * FZNZ_GMAX * 8 + 32 + 2 * (FZNZ_TROWS + 8) * 4;

template <int THREADS>
__device__ void fznz_gram_block(const float* __restrict__ data, i64 ldv, int n, int W, const i64* var, int nv, int rows,
                                float* R, int ld, const unsigned int* mask, unsigned int buf_off, const double* piv) {
    static_assert(THREADS == 256, "canonical summation: 8 warps = 8 row classes (row index mod 8), see the header");
    extern __shared__ __align__(16) unsigned char smem[];      // buf_off: offset of the scratch from the dynamic shared-memory base (keeps LDS/STS)
    unsigned char* buf = smem + buf_off;
    constexpr int B = 4;
    constexpr int NW = THREADS / 32;
    constexpr int PF = (FZNZ_NVMAX * FZNZ_TROWS + THREADS - 1) / THREADS;                // staged values per thread and tile
    static_assert(NW * 32 * B * B <= FZNZ_TROWS * FZNZ_TLD, "warp partials reuse the tile buffer");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* tile = reinterpret_cast<double*>(buf);                                       // [FZNZ_TROWS][FZNZ_TLD] (+ 8 of slack: blocks may over-read)
    double* G = tile + FZNZ_TROWS * FZNZ_TLD + 8;                                        // [ne][FZNZ_GMAX], upper triangle
    int* trow = reinterpret_cast<int*>(G + FZNZ_GMAX * FZNZ_GMAX + 4);                   // [2][FZNZ_TROWS]: table row of each slot (double-buffered)
    int* tcnt = trow + 2 * FZNZ_TROWS;                                                   // [2][8]: slots filled per class
    const int ne = nv + 1;                                                               // entries of v; v[nv] = 1
    const int g = (ne + B - 1) / B;                                                      // block grid of the upper triangle
    const int nblk = g * (g + 1) / 2;                                                    // <= 45: one or two passes of 32 blocks
    const int n_stage = nv * FZNZ_TROWS;
    const int KW = (W + 7) >> 3;                                                         // class words
    auto block_of = [&](int q, int& bi, int& bj) { for (bi = 0; bi < g; ++bi) { const int len = g - bi; if (q < len) { bj = bi + q; return; } q -= len; } bi = bj = 0; };
    // class word kw of class `warp`: bit b <-> table row 256 kw + 8 b + warp (see fznz_uni_warp); warp-uniform
    auto class_word = [&](int kw) {
        unsigned int cw = 0u;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int w = 8 * kw + i;
            const unsigned int m = w < W ? mask[w] : 0u;
            cw |= ((((m >> warp) & 0x01010101u) * 0x10204080u) >> 28) << (4 * i);
        }
        return cw;
    };
    for (int pass = 0; pass * 32 < nblk; ++pass) {
        const int q = lane + 32 * pass;
        const bool live = q < nblk;
        int bi, bj; block_of(live ? q : 0, bi, bj);
        const int i0 = bi * B, j0 = bj * B;
        double acc[B][B];
#pragma unroll
        for (int i = 0; i < B; ++i) {
#pragma unroll
            for (int j = 0; j < B; ++j) acc[i][j] = 0.0;
        }
        __syncthreads();                                                                 // the tile buffer is free (previous pass / caller)
        for (int e = tid; e < FZNZ_TROWS; e += THREADS) {                                // the ones column and the padding never change
#pragma unroll 1
            for (int a = nv; a < FZNZ_TLD; ++a) tile[e * FZNZ_TLD + a] = (a == nv) ? 1.0 : 0.0;
        }
        // cursor of this warp over the view rows of its class (warp-uniform registers)
        int kw = 0; unsigned int cw = class_word(0);
        auto fill = [&](int bsel) {                                                      // the next FZNZ_CROWS view rows of class `warp` -> trow[bsel]
            int filled = 0;
            while (filled < FZNZ_CROWS) {
                if (!cw) { if (++kw >= KW) break; cw = class_word(kw); continue; }
                const int cnt = __popc(cw);
                const int take = cnt < FZNZ_CROWS - filled ? cnt : FZNZ_CROWS - filled;
                if (lane < take) trow[bsel * FZNZ_TROWS + warp * FZNZ_CROWS + filled + lane] = kw * 256 + 8 * (int)__fns(cw, 0, lane + 1) + warp;
                cw = take == cnt ? 0u : (cw & ~((1u << __fns(cw, 0, take + 1)) - 1u));   // drop the `take` lowest set bits
                filled += take;
            }
            if (lane == 0) tcnt[bsel * 8 + warp] = filled;
        };
        float pf[PF];
        auto fetch = [&](int bsel) {
#pragma unroll
            for (int k = 0; k < PF; ++k) {
                const int e = tid + k * THREADS;
                const int a = e / FZNZ_TROWS, slot = e % FZNZ_TROWS;
                const bool ok = e < n_stage && (slot % FZNZ_CROWS) < tcnt[bsel * 8 + slot / FZNZ_CROWS];
                pf[k] = ok ? __ldg(data + var[a] * ldv + trow[bsel * FZNZ_TROWS + slot]) : 0.0f;
            }
        };
        auto any_rows = [&](int bsel) { int t_ = 0;
#pragma unroll
            for (int c = 0; c < 8; ++c) t_ |= tcnt[bsel * 8 + c];
            return t_ != 0; };
        int cur = 0;
        fill(0);
        __syncthreads();                                                                 // trow[0] / tcnt[0] visible
        bool have = any_rows(0);
        if (have) fetch(0);
        while (have) {
            __syncthreads();                                                             // previous tile fully consumed
#pragma unroll
            for (int k = 0; k < PF; ++k) {
                const int e = tid + k * THREADS;
                if (e < n_stage) tile[(e % FZNZ_TROWS) * FZNZ_TLD + e / FZNZ_TROWS] = __dsub_rn((double)pf[k], piv[e / FZNZ_TROWS]);
            }
            fill(cur ^ 1);                                                               // rows of the next tile
            __syncthreads();
            const bool have_next = any_rows(cur ^ 1);
            if (have_next) fetch(cur ^ 1);                                               // in flight while this tile is consumed
            const int mine = tcnt[cur * 8 + warp];
            if (live) {
                for (int r = 0; r < mine; ++r) {
                    const double* v = tile + (warp * FZNZ_CROWS + r) * FZNZ_TLD;
                    double vi[B], vj[B];
#pragma unroll
                    for (int i = 0; i < B; ++i) { vi[i] = v[i0 + i]; vj[i] = v[j0 + i]; }
#pragma unroll
                    for (int i = 0; i < B; ++i) {
#pragma unroll
                        for (int j = 0; j < B; ++j) acc[i][j] = fma(vi[i], vj[j], acc[i][j]);
                    }
                }
            }
            cur ^= 1; have = have_next;
        }
    // cross-warp reduction in a fixed order (deterministic): partials through the tile buffer
        __syncthreads();
#pragma unroll
        for (int i = 0; i < B; ++i) {
#pragma unroll
            for (int j = 0; j < B; ++j) tile[(warp * 32 + lane) * (B * B) + i * B + j] = acc[i][j];
        }
        __syncthreads();
        for (int e = tid; e < 32 * B * B; e += THREADS) {
            const int blk = e / (B * B), ij = e % (B * B);
            const int qq = blk + 32 * pass;
            if (qq >= nblk) continue;
            int ci, cj; block_of(qq, ci, cj);
            const int gi = ci * B + ij / B, gj = cj * B + ij % B;
            if (gi < ne && gj < ne && gi <= gj) {
                double sum = 0.0;
#pragma unroll
                for (int w = 0; w < NW; ++w) sum += tile[(w * 32 + blk) * (B * B) + ij];
                G[gi * FZNZ_GMAX + gj] = sum;
            }
        }
    }
    __syncthreads();
    const int n_pairs = nv * (nv - 1) / 2;
    // Diagonal: only read when a whitelisted member occurs twice in a conditioning set (hiton.jl:20-29).  cor_subset! then sees the
    // variable in two columns of the view and stores their computed correlation at cor_mat[Z, Z] (statfuns.jl:146-152 with X == Y):
    // the same moments with G_ab = G_aa, i.e. 1 up to rounding, or 0 for a variable that is constant on the view (NaN -> 0)
    for (int a = tid; a < nv; a += THREADS) {
        const double g = G[a * FZNZ_GMAX + a], sa = G[a * FZNZ_GMAX + nv];
        const double rr = fznz_r_from_moments(g, g, g, sa, sa, (double)rows);
        R[a * ld + a] = isnan(rr) ? 0.0f : (float)rr;
    }
    for (int e = tid; e < n_pairs; e += THREADS) {
        int a, b; unrank2_small(e, nv, a, b);
        const double rr = fznz_r_from_moments(G[a * FZNZ_GMAX + a], G[b * FZNZ_GMAX + b], G[a * FZNZ_GMAX + b],
                                              G[a * FZNZ_GMAX + nv], G[b * FZNZ_GMAX + nv], (double)rows);
        const float rf = isnan(rr) ? 0.0f : (float)rr;                                   // statfuns.jl:150: NaN -> 0; cor_mat eltype Float32
        R[a * ld + b] = rf; R[b * ld + a] = rf;
    }
    __syncthreads();
}

// ---- block-cooperative cor_subset! (statfuns.jl:138-155) on the rows where var[xs] != 0 and var[ys] != 0 ------
// Fills R[i*ld + j] for all slot pairs i != j < nv (slot -> variable id in var[]), NaN -> 0.  mask: W words of shared memory
// followed (16-byte aligned) by FZNZ_GRAM_BYTES of scratch, mom: 2*nv doubles of shared memory.  Returns the number of rows of
// the view (all threads).
template <int THREADS>
__device__ int fznz_subcor_block(const NzTable& t, const i64* var, int nv, int xs, int ys, float* R, int ld, unsigned int* mask, double* mom, int* s_cnt) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned full = 0xffffffffu;
    const unsigned int* mx = t.nzmask + var[xs] * t.W; const unsigned int* my = t.nzmask + var[ys] * t.W;
    if (tid == 0) { s_cnt[0] = 0; s_cnt[1] = 0x7fffffff; }              // rows of the view, first row of the view
    __syncthreads();
    int c = 0, first = 0x7fffffff;
    for (int w = tid; w < t.W; w += THREADS) {
        unsigned int m = mx[w] & my[w]; mask[w] = m; c += __popc(m);
        if (m && first == 0x7fffffff) first = w * 32 + __ffs(m) - 1;
    }
    c = __reduce_add_sync(full, c);
    first = __reduce_min_sync(full, first);
    if (lane == 0 && c) { atomicAdd(&s_cnt[0], c); atomicMin(&s_cnt[1], first); }
    __syncthreads();
    const int rows = s_cnt[0];
    if (rows == 0) {                                    // empty view: every correlation is NaN -> 0 (statfuns.jl:150)
        for (int e = tid; e < nv * nv; e += THREADS) R[(e / nv) * ld + (e % nv)] = 0.0f;
        __syncthreads();
        return 0;
    }
    const int row0 = s_cnt[1];
    if (nv <= FZNZ_NVMAX) {
        // pivots (the variables' values in the first row of the view) in mom[0..nv)
        for (int a = tid; a < nv; a += THREADS) mom[a] = (double)__ldg(t.data + var[a] * t.ld + row0);
        extern __shared__ __align__(16) unsigned char smem[];
        const unsigned int off = (((unsigned int)__cvta_generic_to_shared(mask + t.W) + 15u) & ~15u) - (unsigned int)__cvta_generic_to_shared(smem);
        fznz_gram_block<THREADS>(t.data, t.ld, t.n, t.W, var, nv, rows, R, ld, mask, off, mom);   // (its first barrier publishes the pivots)
        return rows;
    }
    // larger capacity classes: the same canonical partial sums, one (item, class) per lane - 4 items per warp, lane (j = lane >> 3,
    // c = lane & 7) walks the rows of class c (row mod 8) of item j in increasing order.  First the per-variable moments
    // (S_a, G_aa) into mom[2a], mom[2a+1], then one item per slot pair (a < b).
    const int cls = lane & 7, sub = lane >> 3;
    auto item_sum = [&](const float* xa, double pa, const float* xb, double pb) {        // xb == nullptr: the ones column
        double acc = 0.0;
        for (int row = cls; row < t.n; row += 8) {
            if (!((mask[row >> 5] >> (row & 31)) & 1u)) continue;
            const double da = __dsub_rn((double)xa[row], pa);
            const double db = xb ? __dsub_rn((double)xb[row], pb) : 1.0;
            acc = fma(da, db, acc);
        }
        double tot = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) tot += __shfl_sync(full, acc, (lane & 24) + k);
        return tot;
    };
    for (int it = warp * 4 + sub; it < ((2 * nv + 3) & ~3); it += (THREADS / 32) * 4) {   // item 2a: S_a, item 2a+1: G_aa (padded to whole warps)
        const bool live = it < 2 * nv;
        const int a = live ? it >> 1 : 0;
        const float* xa = t.data + var[a] * t.ld;
        const double pa = (double)xa[row0];
        const double v = item_sum(xa, pa, (it & 1) ? xa : nullptr, pa);
        if (live && cls == 0) mom[it] = v;
    }
    __syncthreads();
    for (int a = tid; a < nv; a += THREADS) {                                 // the diagonal: see fznz_gram_block
        const double rr = fznz_r_from_moments(mom[2 * a + 1], mom[2 * a + 1], mom[2 * a + 1], mom[2 * a], mom[2 * a], (double)rows);
        R[a * ld + a] = isnan(rr) ? 0.0f : (float)rr;
    }
    const int n_pairs = nv * (nv - 1) / 2;
    for (int e = warp * 4 + sub; e < ((n_pairs + 3) & ~3); e += (THREADS / 32) * 4) {
        const bool live = e < n_pairs;
        int a = 0, b = 1; if (live) unrank2(e, nv, a, b);
        const float* xa = t.data + var[a] * t.ld; const float* xb = t.data + var[b] * t.ld;
        const double gab = item_sum(xa, (double)xa[row0], xb, (double)xb[row0]);
        if (live && cls == 0) {
            const double rr = fznz_r_from_moments(mom[2 * a + 1], mom[2 * b + 1], gab, mom[2 * a], mom[2 * b], (double)rows);
            const float rf = isnan(rr) ? 0.0f : (float)rr;                    // statfuns.jl:150: NaN -> 0; cor_mat eltype Float32
            R[a * ld + b] = rf; R[b * ld + a] = rf;
        }
    }
    __syncthreads();
    return rows;
}

// per-warp mask scratch of the pairwise kernels: WARPS x W words of dynamic shared memory when that fits (0 = re-read the masks)
static inline size_t fznz_warp_scratch_bytes(int warps, int W) { const size_t b = (size_t)warps * 4 * W * sizeof(unsigned int); return b <= 96 * 1024 ? b : 0; }   // four pairs per warp

// ---- pairwise stage: one warp per pair, unordered emission (tests.jl:410-433, :391-407) -----------------------
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) pw_fznz_rows_kernel(NzTable t, i64 n_obs_min, double alpha, int reliable_only,
                                                                  u64* counters, i64 cap, int* c_x, int* c_y, double* c_p, double* c_stat, int sh_rank, int sh_world, int use_wm) {
    extern __shared__ unsigned int pw_wm[];                             // WARPS x 4 x W words (or nothing: see fznz_warp_scratch_bytes)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 3;
    const i64 X = blockIdx.x;
    if (!pw_owns_group(X / PW_X_GROUP, sh_rank, sh_world)) return;
    unsigned int* wm = use_wm ? pw_wm + (size_t)(warp * 4 + g) * t.W : nullptr;
    i64 n_rel = 0;
    for (i64 Y0 = X + 1 + warp * 4; Y0 < t.p; Y0 += WARPS * 4) {        // four pairs per warp: one per group of 8 lanes
        const i64 Y = Y0 + g;
        const bool valid = Y < t.p;
        NzUni r = fznz_uni_g8(t, X, Y, n_obs_min, wm, valid);
        const bool rel = valid && (r.suff || !reliable_only) && !isnan(r.pval);
        if ((lane & 7) == 0) {
            n_rel += rel;
            if (rel && r.pval < alpha) {
                u64 pos = atomicAdd(&counters[0], 1ull);
                if ((i64)pos < cap) { c_x[pos] = (int)X; c_y[pos] = (int)Y; c_p[pos] = r.pval; c_stat[pos] = r.stat; }
            }
        }
        __syncwarp();
    }
    n_rel = __reduce_add_sync(0xffffffffu, (unsigned int)n_rel);
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
    __shared__ int s_cnt[2];
    for (i64 tix = blockIdx.x; tix < n_tests; tix += gridDim.x) {
        __syncthreads();
        const int kk = k[tix];
        if (kk == 0) {
            if (threadIdx.x < 32) {
                NzUni r = fznz_uni_warp(t, X[tix], Y[tix], n_obs_min, mask);
                if (threadIdx.x == 0) out[tix] = make_result(r.stat, r.pval, 0, r.suff);
            }
            continue;
        }
        if (threadIdx.x == 0) { var[0] = X[tix]; var[1] = Y[tix]; for (int j = 0; j < 3; ++j) var[2 + j] = j < kk ? Zs[tix * 3 + j] : X[tix]; }
        __syncthreads();
        const int rows = fznz_subcor_block<THREADS>(t, var, kk + 2, 0, 1, R, 5, mask, mom, s_cnt);
        if (threadIdx.x == 0) {
            // tests.jl:250-265 on the (X, Y)-trimmed view: n = rows of the view
            FzConsts fc = nz_consts(rows, n_obs_min);
            CorSlots cs; cs.R = R; cs.ld = 5;
            FzTest ft = fz_cond_test(cs, 0, 1, 2, 3, 4, kk, fc);
            out[tix] = make_result(ft.stat, ft.pval, 0, ft.suff);
        }
    }
}
