// fw_oracle.cpp — CPU restatement of FlashWeave.jl's conditional-independence hot path.
//
// *** TEST INFRASTRUCTURE ONLY. ***  This file is the ORACLE: the checker the CUDA
// path is compared against, and the "port" CPU baseline bench.py times.  Nothing in
// the product path (flashweave.jl_b200/, libfwgpu.so) may include, link or call it.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs load liboracle.so.
//
// Parity status: PINNED.  tests/test_oracle_golden.py checks this file against the
// reference's own fixtures (test/data/tests_expected.tsv: 204 TestResults;
// test/statfuns.jl and test/contingency.jl known answers; the 8 expected graphs in
// test/data/learning_expected/), converted to tests/golden/*.npz|json by
// tests/golden/make_golden.py.
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference).  Indices are 0-based here; the reference is 1-based.
// Build: g++ -O2 -fopenmp -ffp-contract=off (no FMA contraction: Julia does not fuse).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <numeric>
#include <set>
#include <vector>
#include <map>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef int64_t i64;
typedef int32_t i32;

extern "C" {
// src/types.jl:140-145 (TestResult: stat Float64, pval Float64, df Int, suff_power Bool)
typedef struct {
    double stat;
    double pval;
    i64 df;
    uint8_t suff_power;
    uint8_t pad[7];
} fwo_result;
}

enum Kind { MI = 0, MI_NZ = 1, FZ = 2, FZ_NZ = 3 };
static inline bool is_discrete(int kind) { return kind == MI || kind == MI_NZ; }   // types.jl:66-72
static inline bool is_nz(int kind) { return kind == MI_NZ || kind == FZ_NZ; }      // types.jl:64

static const double NaN = std::numeric_limits<double>::quiet_NaN();

static inline fwo_result mk(double s, double p, i64 df, bool sp) {
    fwo_result r;
    memset(&r, 0, sizeof(r));
    r.stat = s; r.pval = p; r.df = df; r.suff_power = sp ? 1 : 0;
    return r;
}

// ---------------------------------------------------------------------------------
// statfuns.jl:3-17  fisher_z_transform / fz_pval.  ccdf(Normal(), x) is StatsFuns'
// normccdf(x) = erfc(x/sqrt2)/2 (Distributions.jl is not vendored; published formula).
// ---------------------------------------------------------------------------------
static double fz_pval(double stat, i64 n, i64 len_z) {
    i64 sample_factor = n - len_z - 3;
    double fz = 0.0;
    if (sample_factor > 0) fz = (std::sqrt((double)sample_factor) / 2.0) * std::log((1.0 + stat) / (1.0 - stat));
    return (std::erfc(std::fabs(fz) * 0.70710678118654752440) / 2.0) * 2.0;
}

// ---------------------------------------------------------------------------------
// statfuns.jl:157-161 mi_pval: ccdf(Chisq(df), g) = Q(df/2, g/2), the regularised upper
// incomplete gamma function (Distributions.jl -> StatsFuns.chisqccdf; published
// definition).  Evaluated by series / continued fraction (Lentz), each to ~1e-16.
// ---------------------------------------------------------------------------------
static double gamma_q(double a, double x) {
    if (!(x > 0.0)) return (x == 0.0 || x < 0.0) ? 1.0 : NaN;
    if (std::isinf(x)) return 0.0;
    double lg = std::lgamma(a);
    if (x < a + 1.0) {
        // P by series, Q = 1 - P
        double ap = a, sum = 1.0 / a, del = sum;
        for (int it = 0; it < 100000; ++it) {
            ap += 1.0; del *= x / ap; sum += del;
            if (std::fabs(del) < std::fabs(sum) * 1e-17) break;
        }
        double P = sum * std::exp(-x + a * std::log(x) - lg);
        return 1.0 - P;
    }
    // continued fraction for Q (modified Lentz)
    const double tiny = 1e-300;
    double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
    for (int i = 1; i < 100000; ++i) {
        double an = -(double)i * ((double)i - a);
        b += 2.0;
        d = an * d + b; if (std::fabs(d) < tiny) d = tiny;
        c = b + an / c; if (std::fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        double del = d * c; h *= del;
        if (std::fabs(del - 1.0) < 1e-16) break;
    }
    return std::exp(-x + a * std::log(x) - lg) * h;
}
static double chisq_sf(i64 df, double x) { return gamma_q(0.5 * (double)df, 0.5 * x); }

static double mi_pval(double mi, i64 df, i64 n_obs) {
    double g = 2.0 * mi * (double)n_obs;
    return df > 0 ? chisq_sf(df, g) : 1.0;
}

// ---------------------------------------------------------------------------------
// statfuns.jl:23-75 pcor_rec.  Julia's arithmetic is type-driven: cor_mat eltype
// ContType is Float32 in LGL (learning.jl:30-31,44) and Float64 in the test-suite's
// convenience call (tests.jl:272).  A value here carries its Julia type (f64 flag) so
// that promotions happen exactly where Julia's would:
//   * `denom == 0.0 ? 0.0 : e/d`, and the clamps to -1.0 / 1.0, yield Float64 literals;
//   * `pYZ0_nZ0^2.0` (statfuns.jl:52) promotes to Float64, `pXZ0_nZ0^2` does not;
//   * round(x, digits=5) is round(x*10^5)/10^5 in x's own type, ties-to-even.
// ---------------------------------------------------------------------------------
struct TV { double v; bool f64; };
static inline TV tv32(float x) { TV t; t.v = (double)x; t.f64 = false; return t; }
static inline TV tv64(double x) { TV t; t.v = x; t.f64 = true; return t; }
static inline TV tmul(TV a, TV b) {
    if (!a.f64 && !b.f64) return tv32((float)a.v * (float)b.v);
    return tv64(a.v * b.v);
}
static inline TV tsub(TV a, TV b) {
    if (!a.f64 && !b.f64) return tv32((float)a.v - (float)b.v);
    return tv64(a.v - b.v);
}
static inline TV tdiv(TV a, TV b) {
    if (!a.f64 && !b.f64) return tv32((float)a.v / (float)b.v);
    return tv64(a.v / b.v);
}
static inline TV tsqrt(TV a) {
    if (!a.f64) return tv32(std::sqrt((float)a.v));
    return tv64(std::sqrt(a.v));
}
static inline TV tround5(TV a) {
    if (!a.f64) {
        float y = std::nearbyintf((float)a.v * 100000.0f) / 100000.0f;
        if (!std::isfinite(y)) return a;
        return tv32(y);
    }
    double y = std::nearbyint(a.v * 100000.0) / 100000.0;
    if (!std::isfinite(y)) return a;
    return tv64(y);
}

struct CorMat {
    const double* m;   // p x p, column-major (symmetric), values are Float32-representable if cont32
    i64 p;
    bool cont32;
    inline TV at(i64 i, i64 j) const {
        double v = m[i + j * p];
        return cont32 ? tv32((float)v) : tv64(v);
    }
};

static TV pcor_rec(i64 X, i64 Y, const i64* Zs, int k, const CorMat& C, i64* n_steps) {
    TV one = C.cont32 ? tv32(1.0f) : tv64(1.0);
    TV p;
    if (n_steps) (*n_steps)++;
    if (k == 1) {
        i64 Z = Zs[0];
        TV pXY = C.at(X, Y), pXZ = C.at(X, Z), pYZ = C.at(Y, Z);
        TV e = tsub(pXY, tmul(pXZ, pYZ));
        e = tround5(e);
        TV d = tmul(tsqrt(tsub(one, tmul(pXZ, pXZ))), tsqrt(tsub(one, tmul(pYZ, pYZ))));
        p = (d.v == 0.0) ? tv64(0.0) : tdiv(e, d);
    } else {
        i64 Z0 = Zs[k - 1];
        TV a = pcor_rec(X, Y, Zs, k - 1, C, n_steps);
        TV b = pcor_rec(X, Z0, Zs, k - 1, C, n_steps);
        TV c = pcor_rec(Y, Z0, Zs, k - 1, C, n_steps);
        TV e = tsub(a, tmul(b, c));
        e = tround5(e);
        TV c2 = tv64(c.v * c.v);   // c^2.0: always Float64 (exactly rounded square)
        TV d = tmul(tsqrt(tsub(one, tmul(b, b))), tsqrt(tsub(one, c2)));
        p = (d.v == 0.0) ? tv64(0.0) : tdiv(e, d);
    }
    if (p.v < -1.0) p = tv64(-1.0);
    else if (p.v >= 1.0) p = tv64(1.0);
    return p;
}

// ---------------------------------------------------------------------------------
// Data views.  The reference trims rows with `@view data[data[:, V] .!= 0, :]`
// (hiton.jl:41-50, tests.jl:412-416, tests.jl:127-131); a view here is a row-index list.
// ---------------------------------------------------------------------------------
struct Data {
    i64 n, p;
    const double* cont;   // column-major n x p (continuous kinds) or null
    const i32* disc;      // column-major n x p (discrete kinds) or null
    inline bool nonzero(i64 row, i64 col) const {
        return cont ? (cont[row + col * n] != 0.0) : (disc[row + col * n] != 0);
    }
};
struct Rows {
    bool all;
    i64 n_all;
    std::vector<i32> idx;
    i64 size() const { return all ? n_all : (i64)idx.size(); }
    inline i64 operator[](i64 i) const { return all ? i : (i64)idx[i]; }
};
static Rows all_rows(const Data& D) { Rows r; r.all = true; r.n_all = D.n; return r; }
static Rows trim_rows(const Data& D, const Rows& in, i64 V) {
    Rows r; r.all = false; r.n_all = D.n;
    i64 m = in.size();
    r.idx.reserve(m);
    for (i64 i = 0; i < m; ++i) { i64 row = in[i]; if (D.nonzero(row, V)) r.idx.push_back((i32)row); }
    return r;
}

// SparseMatrixCSC image of the discrete table (0-based), built when the sparse semantics are switched on
struct Csc { std::vector<i64> colptr; std::vector<i32> rowval, nzval; };

struct Ctx {
    int kind;
    Data D;
    bool sparse_sem = false;             // discrete kinds: follow the reference's sparse-input code path (contingency.jl:80-480)
    Csc csc;
    std::vector<i32> levels, max_vals;   // misc.jl:64-97 (discrete only)
    i64 max_level;                       // types.jl:88-91,110: maximum(max_vals)+1
    CorMat C;                            // precomputed cor_mat (fz) or scratch (fz_nz)
    double* cor_mut;                     // writable p x p scratch for fz_nz (learning.jl:127-129), else null
    bool cont32;
};

// misc.jl:103-107 needs_nz_view (dense data only: the canonical semantics)
static bool needs_nz_view(const Ctx& c, i64 X) {
    bool nz = is_nz(c.kind);
    bool is_nz_var = !is_discrete(c.kind) || c.levels[X] > 2;
    // misc.jl:103-107: `!issparse(data) || isa(test_obj, FzTestCond)` - a sparse discrete table is never row-trimmed, the sparse
    // contingency functions skip the zero rows of X / Y themselves
    if (is_discrete(c.kind) && c.sparse_sem) return false;
    return nz && is_nz_var;
}

// ---------------------------------------------------------------------------------
// Statistics.cor (stdlib; corm -> covzm -> cov2cor!): centred cross products, then
// C_ij / (sd_i * sd_j) clamped to [-1,1], unit diagonal.  Accumulated in Float64; the
// result is rounded to Float32 when cont32 (learning.jl:44 `convert(Matrix{Float32}, ..)`).
// ---------------------------------------------------------------------------------
static void cor_columns(const Data& D, const Rows& R, const i64* vars, i64 nv, double* out /* nv x nv col-major */, bool cont32) {
    i64 m = R.size();
    std::vector<double> xc((size_t)(m * nv));
    std::vector<double> sd(nv);
    for (i64 a = 0; a < nv; ++a) {
        const double* col = D.cont + vars[a] * D.n;
        double s = 0.0;
        for (i64 i = 0; i < m; ++i) s += col[R[i]];
        double mu = s / (double)m;
        double ss = 0.0;
        double* x = &xc[(size_t)(a * m)];
        for (i64 i = 0; i < m; ++i) { x[i] = col[R[i]] - mu; ss += x[i] * x[i]; }
        sd[a] = std::sqrt(ss);
    }
#pragma omp parallel for schedule(dynamic, 4) if (nv > 64)
    for (i64 a = 0; a < nv; ++a) {
        out[a + a * nv] = 1.0;
        const double* xa = &xc[(size_t)(a * m)];
        for (i64 b = a + 1; b < nv; ++b) {
            const double* xb = &xc[(size_t)(b * m)];
            double s = 0.0;
            for (i64 i = 0; i < m; ++i) s += xa[i] * xb[i];
            double r = s / (sd[a] * sd[b]);
            if (r > 1.0) r = 1.0; else if (r < -1.0) r = -1.0;   // clampcor (NaN passes through)
            if (cont32) r = (double)(float)r;
            out[a + b * nv] = r; out[b + a * nv] = r;
        }
    }
}

// ---------------------------------------------------------------------------------
// View correlations of the zero-ignoring Fisher-z kind: cor_subset! (statfuns.jl:138-155, called from
// tests.jl:293-308 on the hiton.jl:41-50,85 views) and the univariate `cor(sub_x, sub_y)` of tests.jl:127-147.
// The reference evaluates Statistics.cor on a Float32 view (Float32 arithmetic, BLAS/pairwise summation order:
// not reproducible outside Julia); here the same Pearson correlation is evaluated in Float64 and rounded to
// Float32, in ONE summation order that the engine (csrc/fznz.cuh) follows operation for operation, so that the
// Float32 correlations - and with them pcor_rec's 5-digit roundings - are bit-equal on both sides:
//   * every variable is shifted by its value in the first row of the view (exact zero variance for a variable
//     that is constant on the view; no cancellation when |mean| >> sd);
//   * raw moments G_ab = sum x'_a x'_b and S_a = sum x'_a are accumulated with fma in increasing row order in
//     8 interleaved partial sums (class = row index in the full table mod 8), the partials added in order 0..7;
//   * c_ab = G_ab - (S_a*S_b)/n, r_ab = c_ab / (sqrt(c_aa)*sqrt(c_bb)), clamped to [-1, 1], NaN kept.
// ---------------------------------------------------------------------------------
static void cor_view(const Data& D, const Rows& R, const i64* vars, i64 nv, double* out /* nv x nv col-major */, bool cont32) {
    const i64 m = R.size();
    const i64 ne = nv + 1;                                   // entry nv is the constant 1 (sums)
    std::vector<double> P((size_t)(8 * ne * ne), 0.0), piv((size_t)nv, 0.0), v((size_t)ne, 1.0);
    if (m > 0) for (i64 a = 0; a < nv; ++a) piv[a] = D.cont[R[0] + vars[a] * D.n];
    for (i64 i = 0; i < m; ++i) {
        const i64 row = R[i];
        double* Pw = &P[(size_t)((row & 7) * ne * ne)];
        for (i64 a = 0; a < nv; ++a) v[a] = D.cont[row + vars[a] * D.n] - piv[a];
        for (i64 a = 0; a < nv; ++a) for (i64 b = a; b < ne; ++b) Pw[a * ne + b] = std::fma(v[a], v[b], Pw[a * ne + b]);
    }
    std::vector<double> G((size_t)(ne * ne), 0.0);
    for (i64 e = 0; e < ne * ne; ++e) { double s = 0.0; for (int w = 0; w < 8; ++w) s += P[(size_t)(w * ne * ne + e)]; G[e] = s; }
    const double n = (double)m;
    for (i64 a = 0; a < nv; ++a) {
        out[a + a * nv] = 1.0;
        const double sa = G[a * ne + nv];
        const double caa = G[a * ne + a] - (sa * sa) / n;
        for (i64 b = a + 1; b < nv; ++b) {
            const double sb = G[b * ne + nv];
            const double cbb = G[b * ne + b] - (sb * sb) / n;
            const double cab = G[a * ne + b] - (sa * sb) / n;
            double r = cab / (std::sqrt(caa) * std::sqrt(cbb));
            if (r > 1.0) r = 1.0; else if (r < -1.0) r = -1.0;   // clampcor (NaN passes through)
            if (cont32) r = (double)(float)r;
            out[a + b * nv] = r; out[b + a * nv] = r;
        }
    }
}

// ---------------------------------------------------------------------------------
// Discrete machinery.
// ---------------------------------------------------------------------------------
struct DiscScratch {
    i64 L, K, nz_slices;                 // L = max_level, K = max_k
    std::vector<i64> ctab;               // L x L x L^K, column-major (types.jl:112)
    std::vector<i64> marg_i, marg_j, marg_k;
    std::vector<i32> z_map;              // types.jl:32-46 ZMapper
    std::vector<i32> z;
    std::vector<i64> cum_levels;
    void init(i64 L_, i64 K_, i64 n) {
        L = L_; K = K_ < 1 ? 1 : K_;
        nz_slices = 1; for (i64 j = 0; j < K; ++j) nz_slices *= L;
        ctab.assign((size_t)(L * L * nz_slices), 0);
        marg_i.assign((size_t)(L * nz_slices), 0);
        marg_j.assign((size_t)(L * nz_slices), 0);
        marg_k.assign((size_t)nz_slices, 0);
        cum_levels.assign((size_t)K, 0);
        cum_levels[0] = 1;
        for (i64 j = 1; j < K; ++j) cum_levels[j] = cum_levels[j - 1] * L;
        i64 mm = L; for (i64 j = 0; j < K; ++j) mm += L * cum_levels[j];
        z_map.assign((size_t)mm, -1);
        z.assign((size_t)n, 0);
    }
};

// statfuns.jl:281-297 adjust_df (one slice)
static i64 adjust_df_slice(const i64* mi, const i64* mj, i64 lx, i64 ly) {
    i64 alx = 0, aly = 0;
    for (i64 i = 0; i < lx; ++i) alx += (mi[i] > 0) - (mi[i] < 0);
    for (i64 j = 0; j < ly; ++j) aly += (mj[j] > 0) - (mj[j] < 0);
    alx = std::max<i64>(1, alx); aly = std::max<i64>(1, aly);
    return (alx - 1) * (aly - 1);
}

// statfuns.jl:209-254 mutual_information (2-D).  `sub` is a (ox,oy)-offset view of an
// L x L table: cell(i,j) = tab[(i+ox) + (j+oy)*L]; view extent (sx, sy).
static double mutual_information_2d(const i64* tab, i64 L, i64 ox, i64 oy, i64 sx, i64 sy, i64 lx, i64 ly,
                                    i64* marg_i, i64* marg_j) {
    for (i64 i = 0; i < L; ++i) { marg_i[i] = 0; marg_j[i] = 0; }
    for (i64 i = 0; i < lx; ++i)
        for (i64 j = 0; j < ly; ++j) {
            i64 c = tab[(i + ox) + (j + oy) * L];
            marg_i[i] += c; marg_j[j] += c;
        }
    double pos = 0.0, neg = 0.0;
    i64 n_pos = 0, n_neg = 0, n_obs = 0;
    for (i64 j = 0; j < sy; ++j) for (i64 i = 0; i < sx; ++i) n_obs += tab[(i + ox) + (j + oy) * L];
    for (i64 i = 0; i < lx; ++i) {
        i64 mii = marg_i[i];
        for (i64 j = 0; j < ly; ++j) {
            i64 c = tab[(i + ox) + (j + oy) * L];
            i64 mjj = marg_j[j];
            if (c != 0 && mii != 0 && mjj != 0) {
                double cell_mi = (double)c * std::log((double)(n_obs * c) / (double)(mii * mjj));
                if (i == j) { pos += cell_mi; n_pos += c; } else { neg += cell_mi; n_neg += c; }
            }
        }
    }
    double mi = (pos + neg) / (double)n_obs;
    if (neg * ((double)n_neg / (double)n_obs) > pos * ((double)n_pos / (double)n_obs)) mi *= -1.0;
    return mi;
}

// statfuns.jl:163-207 mutual_information (3-D) on an offset view of an L x L x S table.
// Marginals over (lx, ly, lz); the MI loop runs over the whole view extent (sx, sy, S).
static double mutual_information_3d(const i64* tab, i64 L, i64 S, i64 ox, i64 oy, i64 sx, i64 sy,
                                    i64 lx, i64 ly, i64 lz, i64* marg_i, i64* marg_j, i64* marg_k) {
    for (i64 t = 0; t < L * S; ++t) { marg_i[t] = 0; marg_j[t] = 0; }
    for (i64 t = 0; t < S; ++t) marg_k[t] = 0;
    for (i64 i = 0; i < lx; ++i)
        for (i64 j = 0; j < ly; ++j)
            for (i64 k = 0; k < lz; ++k) {
                i64 c = tab[(i + ox) + (j + oy) * L + k * L * L];
                marg_i[i + k * L] += c; marg_j[j + k * L] += c; marg_k[k] += c;
            }
    double pos = 0.0, neg = 0.0;
    i64 n_pos = 0, n_neg = 0;
    for (i64 i = 0; i < sx; ++i)
        for (i64 j = 0; j < sy; ++j)
            for (i64 k = 0; k < S; ++k) {
                i64 c = tab[(i + ox) + (j + oy) * L + k * L * L];
                i64 mik = marg_i[i + k * L], mjk = marg_j[j + k * L];
                if (c != 0 && mik != 0 && mjk != 0) {
                    double t = std::log((double)(marg_k[k] * c) / (double)(mik * mjk)) * (double)c;
                    if (i == j) { pos += t; n_pos += c; } else { neg += t; n_neg += c; }
                }
            }
    i64 n_obs = n_pos + n_neg;
    double mi = (pos + neg) / (double)n_obs;
    if (neg * ((double)n_neg / (double)n_obs) > pos * ((double)n_pos / (double)n_obs)) mi *= -1.0;
    return mi;
}

// tests.jl:5-20 sufficient_power (weak pre-check on the data view)
static bool suff_power_pre(const Ctx& c, i64 X, i64 Y, i64 rows, i64 n_obs_min, i64 hps) {
    if (rows < n_obs_min) return false;
    if (is_discrete(c.kind)) {
        i64 lx = c.levels[X], ly = c.levels[Y];
        i64 ox = lx > 1 ? 2 : 1, oy = ly > 1 ? 2 : 1;   // statfuns.jl:307-311 applied to *levels*
        double q = (double)rows / (double)((lx - ox) * (ly - oy));   // Inf/NaN semantics as in Julia
        if (!(q > (double)hps)) return false;
    }
    return true;
}

// tests.jl:28-77 univariate discrete test on a row view
static fwo_result test_mi_uni(const Ctx& c, DiscScratch& s, i64 X, i64 Y, const Rows& R, i64 hps, i64 n_obs_min) {
    i64 rows = R.size();
    if (!suff_power_pre(c, X, Y, rows, n_obs_min, hps)) return mk(0.0, 1.0, 0, false);
    i64 L = c.max_level;
    i64* tab = s.ctab.data();
    for (i64 t = 0; t < L * L; ++t) tab[t] = 0;                      // contingency.jl:7-17
    const i32* cx = c.D.disc + X * c.D.n; const i32* cy = c.D.disc + Y * c.D.n;
    for (i64 i = 0; i < rows; ++i) { i64 r = R[i]; tab[cx[r] + cy[r] * L] += 1; }
    i64 lx = c.levels[X], ly = c.levels[Y], ox = 0, oy = 0, sx = L, sy = L;
    if (is_nz(c.kind)) {                                             // statfuns.jl:307-323
        ox = c.max_vals[X] > 1 ? 1 : 0; oy = c.max_vals[Y] > 1 ? 1 : 0;
        sx = L - ox; sy = L - oy; lx = sx; ly = sy;
    }
    i64 n_obs = 0;
    for (i64 j = 0; j < sy; ++j) for (i64 i = 0; i < sx; ++i) n_obs += tab[(i + ox) + (j + oy) * L];
    if (n_obs < n_obs_min || !((double)n_obs / (double)(lx * ly) > (double)hps)) return mk(0.0, 1.0, 0, false);
    double mi = mutual_information_2d(tab, L, ox, oy, sx, sy, lx, ly, s.marg_i.data(), s.marg_j.data());
    i64 df = adjust_df_slice(s.marg_i.data(), s.marg_j.data(), lx, ly);
    double pval = mi_pval(std::fabs(mi), df, n_obs);
    return mk(mi, pval, df, true);
}

// tests.jl:184-229 conditional discrete test; contingency.jl:42-56 + misc.jl:162-184
// ---- sparse 3-way tables (the reference's default input for sensitive=false, learning.jl:470) -----------------------------
// contingency.jl:182-237: the max_k = 1 / heterogeneous specialisation.  Slices are indexed by the RAW z value and
// levels_z = largest z value seen + 1 (:171-173, :229) - not the number of distinct values the dense path reports.
static i64 sparse_ctab_k1(const Ctx& c, i64 X, i64 Y, i64 Z, bool X_nz, bool Y_nz, i64* tab) {
    const Csc& A = c.csc; const i64 L = c.max_level, n = c.D.n;
    i64 levels_z = 1;
    i64 pX = A.colptr[X], pY = A.colptr[Y], pZ = A.colptr[Z];
    const i64 eX = A.colptr[X + 1], eY = A.colptr[Y + 1], eZ = A.colptr[Z + 1];
    i64 vX = 0, vY = 0, vZ = 0;                                       // 0-based values (the reference's val - 1)
    i64 rZ = pZ < eZ ? A.rowval[pZ] : n;                              // n = "beyond the last row"
    auto zupd = [&](i64 row) {                                        // make_Zupd_expression, :144-161
        while (pZ < eZ - 1 && rZ < row) { ++pZ; rZ = A.rowval[pZ]; }
        if (rZ == row) { vZ = A.nzval[pZ]; if (vZ + 1 > levels_z) levels_z = vZ + 1; } else vZ = 0;
    };
    while (pX < eX && pY < eY) {                                      // :203-218
        const i64 rX = A.rowval[pX], rY = A.rowval[pY];
        if (rX == rY) { vX = A.nzval[pX]; vY = A.nzval[pY]; zupd(rX); tab[vX + vY * L + vZ * L * L] += 1; ++pX; ++pY; }
        else if (rX < rY) { if (!Y_nz) { vX = A.nzval[pX]; vY = 0; zupd(rX); tab[vX + vY * L + vZ * L * L] += 1; } ++pX; }
        else { if (!X_nz) { vY = A.nzval[pY]; vX = 0; zupd(rY); tab[vX + vY * L + vZ * L * L] += 1; } ++pY; }
    }
    if (!Y_nz) { vY = 0; while (pX < eX) { vX = A.nzval[pX]; zupd(A.rowval[pX]); tab[vX + vY * L + vZ * L * L] += 1; ++pX; } }   // :125-141
    if (!X_nz) { vX = 0; while (pY < eY) { vY = A.nzval[pY]; zupd(A.rowval[pY]); tab[vX + vY * L + vZ * L * L] += 1; ++pY; } }
    return levels_z;
}

// contingency.jl:300-480 sparse_ctab_backend! for cols = (X, Y, Zs...): merge over the rows with at least one non-zero among the
// columns; under Nz a row where an adjusted X / Y is zero is skipped; z keys are mapped first-seen (:262-284); the rows never
// visited (all-zero or skipped) are back-filled into ctab[1, 1, slice of the all-zero key], which may ADD a stratum (:461-477).
static i64 sparse_ctab_backend(const Ctx& c, DiscScratch& s, const i64* cols, int N, bool nz_type, bool X_nz, bool Y_nz, i64* tab) {
    const Csc& A = c.csc; const i64 L = c.max_level, n = c.D.n;
    std::fill(s.z_map.begin(), s.z_map.end(), -1);
    i64 levels_total = 0, n_oob = 0, min_ind = n - 1, counted = 0;   // rows 0-based: the reference's n_rows <-> n - 1, n_rows + 1 <-> n
    bool break_loop = false;
    i64 ptr[5], bound[5], rowind[5], val[5];
    for (int i = 0; i < N; ++i) {
        ptr[i] = A.colptr[cols[i]]; bound[i] = A.colptr[cols[i] + 1];
        if (ptr[i] < bound[i]) { rowind[i] = A.rowval[ptr[i]]; if (rowind[i] < min_ind) min_ind = rowind[i]; }
        else { if (nz_type && i < 2 && (i == 0 ? X_nz : Y_nz)) break_loop = true; rowind[i] = n; ++n_oob; }
    }
    while (true) {
        bool skip_row = false;
        i64 next_min = n - 1;
        for (int i = 0; i < N; ++i) {
            if (nz_type && i >= 2 && skip_row) {                      // :395-410: catch up without reading a value
                while (rowind[i] < next_min) { ++ptr[i]; if (ptr[i] >= bound[i]) { ++n_oob; rowind[i] = n; } else rowind[i] = A.rowval[ptr[i]]; }
            } else {
                if (rowind[i] == min_ind) {
                    val[i] = A.nzval[ptr[i]]; ++ptr[i];
                    if (ptr[i] >= bound[i]) { if (nz_type && i < 2 && (i == 0 ? X_nz : Y_nz)) break_loop = true; ++n_oob; rowind[i] = n; }
                    else rowind[i] = A.rowval[ptr[i]];
                } else {
                    val[i] = 0;
                    if (nz_type && i < 2 && (i == 0 ? X_nz : Y_nz)) skip_row = true;
                }
            }
            if (rowind[i] < next_min) next_min = rowind[i];
        }
        if (!skip_row) {
            i64 key = 0; for (int i = 2; i < N; ++i) key += val[i] * s.cum_levels[i - 2];
            i32 zv = s.z_map[(size_t)key];
            if (zv == -1) { zv = (i32)levels_total; s.z_map[(size_t)key] = zv; ++levels_total; }
            tab[val[0] + val[1] * L + (i64)zv * L * L] += 1; ++counted;
        }
        if (nz_type && break_loop) break;
        if (n_oob >= N) break;
        min_ind = next_min;
    }
    const i64 all_zero_obs = n - counted;                             // :461-477
    if (all_zero_obs > 0) {
        i64 idx;
        if (s.z_map[0] != -1) idx = s.z_map[0]; else { idx = levels_total; ++levels_total; }
        tab[idx * L * L] += all_zero_obs;
    }
    return levels_total;
}

static fwo_result mi_cond_from_table(const Ctx& c, DiscScratch& s, i64 X, i64 Y, i64 levels_z, i64 hps);

// tests.jl:184-229 on a sparse table: contingency.jl:240-258 picks the specialisation
static fwo_result test_mi_cond_sparse(const Ctx& c, DiscScratch& s, i64 X, i64 Y, const i64* Zs, int k, i64 hps,
                                      i64* out_levels_z = nullptr, i64* out_ctab = nullptr) {
    const i64 L = c.max_level, S = s.nz_slices;
    i64* tab = s.ctab.data();
    for (i64 t = 0; t < L * L * S; ++t) tab[t] = 0;
    const bool nzk = is_nz(c.kind);
    const bool X_nz = nzk && c.max_vals[X] > 1, Y_nz = nzk && c.max_vals[Y] > 1;
    i64 levels_z;
    if (k == 1 && (X_nz || Y_nz)) levels_z = sparse_ctab_k1(c, X, Y, Zs[0], X_nz, Y_nz, tab);
    else { i64 cols[5] = {X, Y, 0, 0, 0}; for (int j = 0; j < k; ++j) cols[2 + j] = Zs[j]; levels_z = sparse_ctab_backend(c, s, cols, 2 + k, nzk, X_nz, Y_nz, tab); }
    if (out_levels_z) *out_levels_z = levels_z;
    if (out_ctab) memcpy(out_ctab, tab, sizeof(i64) * (size_t)(L * L * S));
    return mi_cond_from_table(c, s, X, Y, levels_z, hps);
}

static fwo_result test_mi_cond(const Ctx& c, DiscScratch& s, i64 X, i64 Y, const i64* Zs, int k, const Rows& R, i64 hps,
                               i64* out_levels_z = nullptr, i64* out_ctab = nullptr) {
    if (c.sparse_sem) return test_mi_cond_sparse(c, s, X, Y, Zs, k, hps, out_levels_z, out_ctab);   // (R is the untrimmed table: needs_nz_view)
    i64 L = c.max_level, S = s.nz_slices, rows = R.size();
    i64* tab = s.ctab.data();
    for (i64 t = 0; t < L * L * S; ++t) tab[t] = 0;
    std::fill(s.z_map.begin(), s.z_map.end(), -1);
    i64 levels_z = 0;
    for (i64 i = 0; i < rows; ++i) {                                  // level_map!
        i64 r = R[i];
        i64 key = 0;   // reference key is 1-based: 1 + sum
        for (int j = 0; j < k; ++j) key += (i64)c.D.disc[r + Zs[j] * c.D.n] * s.cum_levels[j];
        i32 lv = s.z_map[(size_t)key];
        if (lv == -1) { lv = (i32)levels_z; s.z_map[(size_t)key] = lv; levels_z++; }
        s.z[(size_t)i] = lv;
    }
    const i32* cx = c.D.disc + X * c.D.n; const i32* cy = c.D.disc + Y * c.D.n;
    for (i64 i = 0; i < rows; ++i) { i64 r = R[i]; tab[cx[r] + cy[r] * L + (i64)s.z[(size_t)i] * L * L] += 1; }
    if (out_levels_z) *out_levels_z = levels_z;
    if (out_ctab) memcpy(out_ctab, tab, sizeof(i64) * (size_t)(L * L * S));
    return mi_cond_from_table(c, s, X, Y, levels_z, hps);
}

// second half of tests.jl:184-229: nz adjustment, power rule, MI, df, p from the table in s.ctab
static fwo_result mi_cond_from_table(const Ctx& c, DiscScratch& s, i64 X, i64 Y, i64 levels_z, i64 hps) {
    const i64 L = c.max_level, S = s.nz_slices;
    i64* tab = s.ctab.data();
    i64 lx = c.levels[X], ly = c.levels[Y], ox = 0, oy = 0, sx = L, sy = L;
    if (is_nz(c.kind)) {
        ox = c.max_vals[X] > 1 ? 1 : 0; oy = c.max_vals[Y] > 1 ? 1 : 0;
        sx = L - ox; sy = L - oy; lx = sx; ly = sy;
    }
    i64 n_obs = 0;
    for (i64 kk = 0; kk < S; ++kk) for (i64 j = 0; j < sy; ++j) for (i64 i = 0; i < sx; ++i)
        n_obs += tab[(i + ox) + (j + oy) * L + kk * L * L];
    if (!((double)n_obs / (double)(lx * ly * levels_z) > (double)hps)) return mk(0.0, 1.0, 0, false);  // tests.jl:210
    double mi = mutual_information_3d(tab, L, S, ox, oy, sx, sy, lx, ly, levels_z,
                                      s.marg_i.data(), s.marg_j.data(), s.marg_k.data());
    i64 df = 0;                                                      // statfuns.jl:299-305
    for (i64 kk = 0; kk < levels_z; ++kk) df += adjust_df_slice(&s.marg_i[(size_t)(kk * L)], &s.marg_j[(size_t)(kk * L)], lx, ly);
    double pval = mi_pval(std::fabs(mi), df, n_obs);
    return mk(mi, pval, df, true);
}

// tests.jl:108-160 univariate Fisher-z test.  R is the caller's (X-trimmed for _nz) view.
static fwo_result test_fz_uni(const Ctx& c, i64 X, i64 Y, const Rows& R, i64 n_obs_min) {
    i64 rows = R.size();
    if (rows < n_obs_min) return mk(0.0, 1.0, 0, 0 >= n_obs_min);     // tests.jl:111-115,159
    double p_stat; i64 n_obs;
    if (c.kind == FZ && c.C.m) {                                      // tests.jl:149-152
        n_obs = rows;
        p_stat = n_obs >= n_obs_min ? c.C.m[X + Y * c.C.p] : 0.0;
    } else {
        Rows sub = is_nz(c.kind) ? trim_rows(c.D, R, Y) : R;          // tests.jl:127-131
        if (sub.size() == 0) { p_stat = 0.0; n_obs = 0; }
        else {
            n_obs = sub.size();
            if (n_obs >= n_obs_min) {
                i64 vars[2] = {X, Y}; double out[4];
                if (is_nz(c.kind)) cor_view(c.D, sub, vars, 2, out, c.cont32);
                else cor_columns(c.D, sub, vars, 2, out, c.cont32);
                p_stat = out[2];
            } else p_stat = 0.0;
        }
    }
    double pval = fz_pval(p_stat, n_obs, 0);
    return mk(p_stat, pval, 0, n_obs >= n_obs_min);
}

// tests.jl:250-265 conditional Fisher-z test (len_z hard-coded 0 at :256)
static fwo_result test_fz_cond(const CorMat& C, i64 X, i64 Y, const i64* Zs, int k, i64 rows, i64 n_obs_min, i64* n_steps = nullptr) {
    if (rows < n_obs_min) return mk(0.0, 1.0, 0, false);
    TV p = pcor_rec(X, Y, Zs, k, C, n_steps);
    double pval = fz_pval(p.v, rows, 0);
    return mk(p.v, pval, 0, true);
}

static inline bool issig(const fwo_result& r, double alpha) { return r.pval < alpha && r.suff_power; }   // tests.jl:1-3

static double binom(i64 n, i64 k) {
    if (k < 0 || k > n) return 0.0;
    double r = 1.0;
    for (i64 i = 1; i <= k; ++i) r = r * (double)(n - k + i) / (double)i;
    return std::floor(r + 0.5);
}

struct SubsetsOut { fwo_result res; i64 Zs[3]; int k; i64 num_tests; double frac; };

struct Params {
    int kind; int max_k; double alpha; i64 hps; i64 n_obs_min; i64 max_tests;
    bool fdr; bool correct_reliable_only; bool fast_elim;
};

// tests.jl:281-346 test_subsets (exhaustive).  R = view trimmed for X and Y as hiton.jl:85 does.
static SubsetsOut test_subsets(Ctx& c, DiscScratch* s, i64 X, i64 Y, const std::vector<i64>& Z_total, const Rows& R, const Params& P) {
    SubsetsOut o; o.k = 0; o.Zs[0] = o.Zs[1] = o.Zs[2] = -1;
    i64 m = (i64)Z_total.size();
    if (m == 0) { o.res = mk(NaN, NaN, -1, true); o.k = 1; o.Zs[0] = -1; o.num_tests = -1; o.frac = NaN; return o; }  // :285
    fwo_result lowest = mk(0.0, 0.0, 0, true); i64 lowZ[3] = {-1, -1, -1}; int lowk = 0;
    bool disc = is_discrete(c.kind), nz = is_nz(c.kind);
    i64 rows = R.size();
    i64 num_tests = 0;
    if (!disc && nz) {                                               // :293-308
        if (P.n_obs_min > rows) { o.res = mk(0.0, 1.0, 0, false); o.num_tests = 0; o.frac = 0.0; return o; }
        std::vector<i64> vars; vars.push_back(X); vars.push_back(Y);
        for (i64 z : Z_total) vars.push_back(z);
        i64 nv = (i64)vars.size();
        std::vector<double> sub((size_t)(nv * nv));
        cor_view(c.D, R, vars.data(), nv, sub.data(), c.cont32);      // statfuns.jl:138-155 cor_subset!
        for (i64 a = 0; a < nv - 1; ++a) for (i64 b = a + 1; b < nv; ++b) {
            double v = sub[(size_t)(a + b * nv)]; if (std::isnan(v)) v = 0.0;
            c.cor_mut[(size_t)(vars[a] + vars[b] * c.D.p)] = v; c.cor_mut[(size_t)(vars[b] + vars[a] * c.D.p)] = v;
        }
    }
    double total = 0.0;
    for (int ss = P.max_k; ss >= 1; --ss) {
        total += binom(m, ss);
        if (ss > m) continue;
        std::vector<i64> comb(ss); for (int i = 0; i < ss; ++i) comb[i] = i;   // lexicographic, Combinatorics.combinations
        while (true) {
            i64 Zs[3]; for (int i = 0; i < ss; ++i) Zs[i] = Z_total[(size_t)comb[i]];
            fwo_result r = disc ? test_mi_cond(c, *s, X, Y, Zs, ss, R, P.hps) : test_fz_cond(c.C, X, Y, Zs, ss, rows, P.n_obs_min);
            num_tests++;
            if (!issig(r, P.alpha) || (P.max_tests > 0 && num_tests >= P.max_tests)) {   // :326-336
                for (int rs = ss - 1; rs >= 1; --rs) total += binom(m, rs);
                o.res = r; o.k = ss; for (int i = 0; i < ss; ++i) o.Zs[i] = Zs[i];
                o.num_tests = num_tests; o.frac = (double)num_tests / total; return o;
            } else if (r.pval >= lowest.pval) { lowest = r; lowk = ss; for (int i = 0; i < ss; ++i) lowZ[i] = Zs[i]; }
            int i = ss - 1;
            while (i >= 0 && comb[i] == m - ss + i) --i;
            if (i < 0) break;
            comb[i]++; for (int j = i + 1; j < ss; ++j) comb[j] = comb[j - 1] + 1;
        }
    }
    o.res = lowest; o.k = lowk; for (int i = 0; i < 3; ++i) o.Zs[i] = lowZ[i];
    o.num_tests = num_tests; o.frac = (double)num_tests / total;
    return o;
}

// statfuns.jl:326-350 benjamini_hochberg!
static void benjamini_hochberg(double* pvals, i64 n, double alpha, i64 m) {
    if (n == 0) return;
    std::vector<std::pair<i64, double>> sp;
    for (i64 i = 0; i < n; ++i) if (pvals[i] < alpha) sp.push_back(std::make_pair(i, pvals[i]));
    if (sp.empty()) return;
    std::stable_sort(sp.begin(), sp.end(), [](const std::pair<i64, double>& a, const std::pair<i64, double>& b) { return a.second < b.second; });
    i64 nf = (i64)sp.size();
    sp[(size_t)(nf - 1)].second = std::min(sp[(size_t)(nf - 1)].second * (double)m / (double)nf, 1.0);
    for (i64 i = nf - 2; i >= 0; --i) {
        double next_adj = sp[(size_t)(i + 1)].second;
        double new_adj = sp[(size_t)i].second * (double)m / (double)(i + 1);
        sp[(size_t)i].second = std::min(next_adj, new_adj);
    }
    for (i64 i = 0; i < n; ++i) pvals[i] = NaN;
    for (auto& e : sp) pvals[(size_t)e.first] = e.second;
}

// tests.jl:395 condensed index, 0-based: pairs (X<Y) row-major upper triangle
static inline i64 pair_index(i64 X, i64 Y, i64 p) { return X * p - X * (X + 1) / 2 + (Y - X - 1); }

struct Nbr { i64 v; double stat; double pval; };
typedef std::vector<std::vector<Nbr>> NbrLists;   // per variable, insertion-ordered (OrderedDict)

// tests.jl:410-532 pw_univar_neighbors (+ :372-407)
static void pairwise(Ctx& c, const Params& P, NbrLists& out, std::vector<double>* raw_stats = nullptr, std::vector<double>* raw_pvals = nullptr,
                     i64* n_tests_out = nullptr) {
    i64 p = c.D.p, n_pairs = p * (p - 1) / 2;
    std::vector<double> stats((size_t)n_pairs, NaN), pvals((size_t)n_pairs, NaN);
    bool disc = is_discrete(c.kind);
#pragma omp parallel
    {
        DiscScratch s; if (disc) s.init(c.max_level, 1, c.D.n);
#pragma omp for schedule(dynamic, 8)
        for (i64 X = 0; X < p - 1; ++X) {
            Rows R = needs_nz_view(c, X) ? trim_rows(c.D, all_rows(c.D), X) : all_rows(c.D);   // tests.jl:412-416
            for (i64 Y = X + 1; Y < p; ++Y) {
                fwo_result r;
                if (disc) r = (c.levels[X] < 2) ? mk(0.0, 1.0, 0, false) : test_mi_uni(c, s, X, Y, R, P.hps, P.n_obs_min);   // tests.jl:86-92
                else r = test_fz_uni(c, X, Y, R, P.n_obs_min);
                i64 pi = pair_index(X, Y, p);
                if (P.correct_reliable_only && !r.suff_power) { stats[(size_t)pi] = NaN; pvals[(size_t)pi] = NaN; }  // :397-402
                else { stats[(size_t)pi] = r.stat; pvals[(size_t)pi] = r.pval; }
            }
        }
    }
    if (n_tests_out) *n_tests_out = n_pairs;
    if (raw_stats) *raw_stats = stats;
    if (raw_pvals) *raw_pvals = pvals;
    if (P.fdr) {                                                     // :521-529
        i64 m = n_pairs;
        if (P.correct_reliable_only) { i64 nn = 0; for (double v : pvals) nn += std::isnan(v) ? 1 : 0; m -= nn; }
        benjamini_hochberg(pvals.data(), n_pairs, P.alpha, m);
    }
    out.assign((size_t)p, std::vector<Nbr>());
    for (i64 X = 0; X < p - 1; ++X) for (i64 Y = X + 1; Y < p; ++Y) {   // :372-388
        i64 pi = pair_index(X, Y, p);
        double pv = pvals[(size_t)pi];
        if (!std::isnan(pv) && pv < P.alpha) {
            Nbr a = {Y, stats[(size_t)pi], pv}; Nbr b = {X, stats[(size_t)pi], pv};
            out[(size_t)X].push_back(a); out[(size_t)Y].push_back(b);
        }
    }
}

// ---------------------------------------------------------------------------------
// hiton.jl:283-400 si_HITON_PC with time_limit = 0 (no preemption), bnb = false.
// ---------------------------------------------------------------------------------
struct HitonOut {
    std::vector<Nbr> PC;      // state_results, insertion order
    std::vector<Nbr> TPC;     // inter_results
    i64 num_tests;            // Σ num_tests over check_candidate! calls (tests.jl:322)
};

struct OrderedNbrDict {
    std::vector<Nbr> items;
    int find(i64 v) const { for (size_t i = 0; i < items.size(); ++i) if (items[i].v == v) return (int)i; return -1; }
    void set(i64 v, double s, double p) { int i = find(v); if (i >= 0) { items[(size_t)i].stat = s; items[(size_t)i].pval = p; } else { Nbr n = {v, s, p}; items.push_back(n); } }
};

// hiton.jl:109-149 hiton_backend
static void hiton_backend(Ctx& c, DiscScratch* s, i64 T, const std::vector<i64>& candidates, const Rows& RT, const Params& P,
                          const std::set<i64>& whitelist, char phase, const OrderedNbrDict& support, OrderedNbrDict& accepted_dict,
                          i64& num_tests) {
    std::vector<i64> accepted;
    if (phase == 'E') accepted = candidates;
    for (size_t ci = 0; ci < candidates.size(); ++ci) {
        i64 cand = candidates[ci];
        if (!whitelist.empty() && whitelist.count(cand)) {            // hiton.jl:20-38
            accepted.push_back(cand);
            accepted_dict.set(cand, NaN, NaN);
            continue;
        }
        if (phase == 'E') accepted.erase(std::remove(accepted.begin(), accepted.end(), cand), accepted.end());   // :134-136
        Rows R2 = needs_nz_view(c, cand) ? trim_rows(c.D, RT, cand) : RT;                                        // :85
        SubsetsOut so = test_subsets(c, s, T, cand, accepted, R2, P);
        if (so.num_tests > 0) num_tests += so.num_tests;
        if (accepted.empty()) {                                        // hiton.jl:57-59
            accepted.push_back(cand);
            int si = support.find(cand);
            accepted_dict.set(cand, support.items[(size_t)si].stat, support.items[(size_t)si].pval);
        } else if (issig(so.res, P.alpha)) {
            accepted.push_back(cand);
            accepted_dict.set(cand, so.res.stat, so.res.pval);
        } else {
            if (phase == 'E' && !P.fast_elim) accepted.push_back(cand);
        }
    }
}

static HitonOut si_hiton_pc(Ctx& c, DiscScratch* s, i64 T, const std::vector<Nbr>& univar_nbrs, const Params& P, const std::set<i64>& whitelist) {
    HitonOut out; out.num_tests = 0;
    if (is_discrete(c.kind) && c.levels[T] < 2) return out;           // hiton.jl:182-183, 300-302
    Rows RT = needs_nz_view(c, T) ? trim_rows(c.D, all_rows(c.D), T) : all_rows(c.D);   // :193
    if (P.max_k == 0) { out.PC = univar_nbrs; return out; }           // :394-397
    // prepare_interleaving_phase (hiton.jl:199-220): candidates with p < alpha, stable sort by p
    std::vector<std::pair<i64, double>> cp;
    for (const Nbr& nb : univar_nbrs) if (nb.pval < P.alpha) cp.push_back(std::make_pair(nb.v, nb.pval));
    std::stable_sort(cp.begin(), cp.end(), [](const std::pair<i64, double>& a, const std::pair<i64, double>& b) { return a.second < b.second; });
    std::vector<i64> candidates; for (auto& e : cp) candidates.push_back(e.first);
    if (candidates.empty()) return out;                               // :336-338
    OrderedNbrDict uni; uni.items = univar_nbrs;
    OrderedNbrDict TPC;
    hiton_backend(c, s, T, candidates, RT, P, whitelist, 'I', uni, TPC, out.num_tests);
    // prepare_elimination_phase (:223-246): candidates = keys(TPC) in insertion order
    std::vector<i64> pc_cands; for (const Nbr& nb : TPC.items) pc_cands.push_back(nb.v);
    OrderedNbrDict PC;
    hiton_backend(c, s, T, pc_cands, RT, P, whitelist, 'E', TPC, PC, out.num_tests);
    // update_PC_dict! (:249-256)  (no_red_tests=true / fast_elim=true defaults, learning.jl:207)
    for (Nbr& nb : PC.items) {
        int ti = TPC.find(nb.v);
        if (ti >= 0 && (TPC.items[(size_t)ti].pval > nb.pval || std::isnan(nb.pval))) { nb.stat = TPC.items[(size_t)ti].stat; nb.pval = TPC.items[(size_t)ti].pval; }
    }
    out.PC = PC.items; out.TPC = TPC.items;
    return out;
}

// misc.jl:137-159 make_weights (weight_type "cond_stat")
static double make_weight(int kind, const Nbr& pc, const std::vector<Nbr>& univar) {
    if (is_discrete(kind)) {
        double us = 0.0; for (const Nbr& u : univar) if (u.v == pc.v) { us = u.stat; break; }
        double sg = (us > 0.0) - (us < 0.0);
        return sg * std::fabs(pc.stat);
    }
    return pc.stat;
}

// misc.jl:201-218 maxweight
static double maxweight(double w1, double w2) {
    if (std::isnan(w1)) return w2;
    if (std::isnan(w2)) return w1;
    double s1 = (w1 > 0) - (w1 < 0), s2 = (w2 > 0) - (w2 < 0);
    if (s1 * s2 < 0) return w1;
    return std::max(std::fabs(w1), std::fabs(w2)) * s1;
}

struct Edge { i64 a, b; double w; };

// ---------------------------------------------------------------------------------
// learning.jl:203-279 LGL.  mode 0 = parallel="single" (learning.jl:137-138; the GPU
// parity target); mode 1 = one-worker "single_il" emulation with the feed-forward
// whitelist (interleaved.jl:60-86,113-179; used only to pin the oracle to the
// committed edgelists).
// ---------------------------------------------------------------------------------
static double now_s() {
#ifdef _OPENMP
    return omp_get_wtime();
#else
    return 0.0;
#endif
}

static void lgl(Ctx& c, Params P, int mode, std::vector<Edge>& edges, i64* cond_tests, i64* pair_tests, int n_threads,
                const i64* target_subset, i64 n_target_subset, std::vector<HitonOut>* per_target_out, double* phase_secs) {
    i64 p = c.D.p;
    NbrLists uni;
    double t0 = now_s();
    pairwise(c, P, uni, nullptr, nullptr, pair_tests);
    double t1 = now_s();
    if (phase_secs) phase_secs[1] = t1 - t0;
    // learning.jl:97-98 target order: ascending univariate degree, stable
    std::vector<i64> order((size_t)p); std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](i64 a, i64 b) { return uni[(size_t)a].size() < uni[(size_t)b].size(); });
    std::vector<HitonOut> res((size_t)p);
    i64 total_tests = 0;
    if (P.max_k == 0) {
        for (i64 t = 0; t < p; ++t) res[(size_t)t].PC = uni[(size_t)t];
    } else if (mode == 0) {
        std::vector<i64> targets;
        if (target_subset) targets.assign(target_subset, target_subset + n_target_subset); else targets = order;
        i64 nt = (i64)targets.size();
#pragma omp parallel num_threads(n_threads > 0 ? n_threads : 1) reduction(+ : total_tests)
        {
            DiscScratch s; if (is_discrete(c.kind)) s.init(c.max_level, P.max_k, c.D.n);
            // fz_nz mutates the scratch cor_mat (learning.jl:127-129): one per worker
            Ctx local = c;
            std::vector<double> scratch;
            if (c.kind == FZ_NZ) { scratch.assign((size_t)(p * p), 0.0); local.cor_mut = scratch.data(); local.C.m = scratch.data(); }
            std::set<i64> empty;
#pragma omp for schedule(dynamic, 1)
            for (i64 ti = 0; ti < nt; ++ti) {
                i64 T = targets[(size_t)ti];
                res[(size_t)T] = si_hiton_pc(local, &s, T, uni[(size_t)T], P, empty);
                total_tests += res[(size_t)T].num_tests;
            }
        }
    } else {
        DiscScratch s; if (is_discrete(c.kind)) s.init(c.max_level, P.max_k, c.D.n);
        std::vector<double> scratch;
        if (c.kind == FZ_NZ) { scratch.assign((size_t)(p * p), 0.0); c.cor_mut = scratch.data(); c.C.m = scratch.data(); }
        std::vector<std::set<i64>> graph((size_t)p);
        std::vector<i64> sched;
        if (p >= 2) { sched.push_back(order[1]); sched.push_back(order[0]); for (i64 i = 2; i < p; ++i) sched.push_back(order[(size_t)i]); }
        else sched = order;
        // FIFO of two initial jobs with empty whitelists, then one-at-a-time with whitelist =
        // neighbours in the graph of all finished targets (interleaved.jl:136-141,166-179)
        for (size_t i = 0; i < sched.size(); ++i) {
            i64 T = sched[i];
            std::set<i64> wl; if (i >= 2) wl = graph[(size_t)T];
            res[(size_t)T] = si_hiton_pc(c, &s, T, uni[(size_t)T], P, wl);
            total_tests += res[(size_t)T].num_tests;
            for (const Nbr& nb : res[(size_t)T].PC) { graph[(size_t)T].insert(nb.v); graph[(size_t)nb.v].insert(T); }
        }
    }
    if (cond_tests) *cond_tests = total_tests;
    if (phase_secs) phase_secs[2] = now_s() - t1;
    // weights + make_symmetric_graph (misc.jl:230-272), OR rule
    std::vector<std::map<i64, double>> W((size_t)p);
    for (i64 t = 0; t < p; ++t) for (const Nbr& nb : res[(size_t)t].PC) W[(size_t)t][nb.v] = make_weight(c.kind, nb, uni[(size_t)t]);
    std::set<std::pair<i64, i64>> seen;
    for (i64 a = 0; a < p; ++a) for (auto& kv : W[(size_t)a]) {
        i64 b = kv.first;
        std::pair<i64, i64> e = a >= b ? std::make_pair(a, b) : std::make_pair(b, a);
        if (seen.count(e)) continue;
        double w = kv.second, rw = NaN;
        auto it = W[(size_t)b].find(a); if (it != W[(size_t)b].end()) rw = it->second;
        double sw = maxweight(w, rw);
        if (std::isnan(sw)) continue;
        Edge ed = {std::min(a, b), std::max(a, b), sw}; edges.push_back(ed); seen.insert(e);
    }
    if (per_target_out) *per_target_out = res;
}

// misc.jl:64-97 get_levels / get_max_vals (dense)
static void compute_levels(const Data& D, std::vector<i32>& levels, std::vector<i32>& max_vals) {
    levels.assign((size_t)D.p, 0); max_vals.assign((size_t)D.p, 0);
    for (i64 v = 0; v < D.p; ++v) {
        std::set<i32> u; i32 mx = std::numeric_limits<i32>::min();
        for (i64 i = 0; i < D.n; ++i) { i32 x = D.disc[i + v * D.n]; u.insert(x); mx = std::max(mx, x); }
        levels[(size_t)v] = (i32)u.size(); max_vals[(size_t)v] = D.n ? mx : 0;
    }
}

// =================================================================================
// C API (ctypes)
// =================================================================================
extern "C" {

struct fwo_ctx {
    Ctx c;
    std::vector<double> cor_store;
    void ensure_scratch() {
        if (cor_store.empty()) cor_store.assign((size_t)(c.D.p * c.D.p), 0.0);
        c.cor_mut = cor_store.data(); c.C.m = cor_store.data();
    }
};

// kind: 0 mi, 1 mi_nz, 2 fz, 3 fz_nz.  data_f64 (continuous) or data_i32 (discrete),
// column-major n x p, borrowed.  cont32 != 0: ContType = Float32 (LGL, learning.jl:30-31);
// cont32 == 0: Float64 (the prec=64 convenience calls in test/tests.jl).
fwo_ctx* fwo_create(int kind, i64 n, i64 p, const double* data_f64, const i32* data_i32, int cont32) {
    fwo_ctx* h = new fwo_ctx();
    Ctx& c = h->c;
    c.kind = kind; c.D.n = n; c.D.p = p; c.D.cont = data_f64; c.D.disc = data_i32; c.cont32 = cont32 != 0;
    c.max_level = 0; c.C.m = nullptr; c.cor_mut = nullptr; c.C.p = p; c.C.cont32 = c.cont32;
    if (is_discrete(kind)) {
        compute_levels(c.D, c.levels, c.max_vals);
        i32 mx = 0; for (i32 v : c.max_vals) mx = std::max(mx, v);
        c.max_level = (i64)mx + 1;
    }
    return h;
}
void fwo_destroy(fwo_ctx* h) { delete h; }
// discrete kinds: follow the reference's sparse-input code path (SparseMatrixCSC tables, the default for sensitive=false)
void fwo_set_sparse_semantics(fwo_ctx* h, int on) {
    Ctx& c = h->c;
    c.sparse_sem = on != 0 && is_discrete(c.kind);
    if (c.sparse_sem && c.csc.colptr.empty()) {
        c.csc.colptr.assign((size_t)c.D.p + 1, 0);
        for (i64 v = 0; v < c.D.p; ++v) {
            for (i64 i = 0; i < c.D.n; ++i) { i32 x = c.D.disc[i + v * c.D.n]; if (x != 0) { c.csc.rowval.push_back((i32)i); c.csc.nzval.push_back(x); } }
            c.csc.colptr[(size_t)v + 1] = (i64)c.csc.rowval.size();
        }
    }
}
void fwo_get_levels(fwo_ctx* h, i32* levels, i32* max_vals) {
    for (i64 i = 0; i < h->c.D.p; ++i) { levels[i] = h->c.levels[(size_t)i]; max_vals[i] = h->c.max_vals[(size_t)i]; }
}

// learning.jl:42-44: cor_mat = Float32.(cor(data)) (or Float64 when cont32 == 0)
void fwo_compute_cor(fwo_ctx* h) {
    Ctx& c = h->c; i64 p = c.D.p;
    h->cor_store.assign((size_t)(p * p), 0.0);
    std::vector<i64> vars((size_t)p); std::iota(vars.begin(), vars.end(), 0);
    cor_columns(c.D, all_rows(c.D), vars.data(), p, h->cor_store.data(), c.cont32);
    c.C.m = h->cor_store.data();
}
// install a caller-provided cor_mat (p x p doubles holding ContType values)
void fwo_set_cor(fwo_ctx* h, const double* cor) {
    Ctx& c = h->c; i64 p = c.D.p;
    h->cor_store.assign(cor, cor + p * p); c.C.m = h->cor_store.data();
}
void fwo_get_cor(fwo_ctx* h, double* out) { memcpy(out, h->cor_store.data(), sizeof(double) * h->cor_store.size()); }
void fwo_alloc_scratch_cor(fwo_ctx* h) { h->ensure_scratch(); }

double fwo_fz_pval(double stat, i64 n, i64 len_z) { return fz_pval(stat, n, len_z); }
double fwo_chisq_sf(i64 df, double x) { return chisq_sf(df, x); }
double fwo_mi_pval(double mi, i64 df, i64 n_obs) { return mi_pval(mi, df, n_obs); }
void fwo_benjamini_hochberg(double* pvals, i64 n, double alpha, i64 m) { benjamini_hochberg(pvals, n, alpha, m); }

double fwo_pcor_rec(const double* cor, i64 p, int cont32, i64 X, i64 Y, const i64* Zs, int k, i64* n_steps) {
    CorMat C; C.m = cor; C.p = p; C.cont32 = cont32 != 0;
    if (n_steps) *n_steps = 0;
    return pcor_rec(X, Y, Zs, k, C, n_steps).v;
}

// convenience MI of a dense table (statfuns.jl:258-279): dims (lx, ly[, lz]) column-major
double fwo_mutual_information(const i64* ctab, i64 lx, i64 ly, i64 lz) {
    i64 L = std::max(lx, ly);
    if (lz <= 0) {
        std::vector<i64> tab((size_t)(L * L), 0), mi((size_t)L), mj((size_t)L);
        for (i64 i = 0; i < lx; ++i) for (i64 j = 0; j < ly; ++j) tab[(size_t)(i + j * L)] = ctab[i + j * lx];
        return mutual_information_2d(tab.data(), L, 0, 0, lx, ly, lx, ly, mi.data(), mj.data());
    }
    std::vector<i64> tab((size_t)(L * L * lz), 0), mi((size_t)(L * lz)), mj((size_t)(L * lz)), mk_((size_t)lz);
    for (i64 i = 0; i < lx; ++i) for (i64 j = 0; j < ly; ++j) for (i64 k = 0; k < lz; ++k)
        tab[(size_t)(i + j * L + k * L * L)] = ctab[i + j * lx + k * lx * ly];
    return mutual_information_3d(tab.data(), L, lz, 0, 0, lx, ly, lx, ly, lz, mi.data(), mj.data(), mk_.data());
}

// Single tests.  rows: optional row-index view (n_rows < 0: all rows).  For the
// univariate forms the view is the caller's X-trimmed one (tests.jl:412-416); pass
// trim_x != 0 to let the oracle apply needs_nz_view(X) itself.
static Rows make_rows(const Ctx& c, const i32* rows, i64 n_rows) {
    if (n_rows < 0) return all_rows(c.D);
    Rows r; r.all = false; r.n_all = c.D.n; r.idx.assign(rows, rows + n_rows); return r;
}

void fwo_test_uni(fwo_ctx* h, i64 X, const i64* Ys, i64 nY, i64 hps, i64 n_obs_min, int trim_x, const i32* rows, i64 n_rows, fwo_result* out) {
    Ctx& c = h->c;
    Rows R = make_rows(c, rows, n_rows);
    if (trim_x && needs_nz_view(c, X)) R = trim_rows(c.D, R, X);
    DiscScratch s; if (is_discrete(c.kind)) s.init(c.max_level, 1, c.D.n);
    for (i64 i = 0; i < nY; ++i) {
        if (is_discrete(c.kind)) out[i] = (c.levels[(size_t)X] < 2) ? mk(0.0, 1.0, 0, false) : test_mi_uni(c, s, X, Ys[i], R, hps, n_obs_min);
        else out[i] = test_fz_uni(c, X, Ys[i], R, n_obs_min);
    }
}

// conditional single test on a row view (hiton.jl:85: the caller trims for T and candidate;
// trim_xy != 0 applies needs_nz_view to X then Y here).
void fwo_test_cond(fwo_ctx* h, i64 X, i64 Y, const i64* Zs, int k, i64 hps, i64 n_obs_min, int max_k, int trim_xy,
                   const i32* rows, i64 n_rows, fwo_result* out, i64* levels_z, i64* ctab_out) {
    Ctx& c = h->c;
    Rows R = make_rows(c, rows, n_rows);
    if (trim_xy) { if (needs_nz_view(c, X)) R = trim_rows(c.D, R, X); if (needs_nz_view(c, Y)) R = trim_rows(c.D, R, Y); }
    if (is_discrete(c.kind)) {
        DiscScratch s; s.init(c.max_level, std::max(max_k, k), c.D.n);
        *out = test_mi_cond(c, s, X, Y, Zs, k, R, hps, levels_z, ctab_out);
    } else {
        if (c.kind == FZ_NZ) {
            // tests.jl:293-308: cor_subset! on [X, Y, Zs...] (here Z_total = Zs)
            std::vector<i64> vars; vars.push_back(X); vars.push_back(Y); for (int i = 0; i < k; ++i) vars.push_back(Zs[i]);
            i64 nv = (i64)vars.size(); std::vector<double> sub((size_t)(nv * nv));
            h->ensure_scratch();
            cor_view(c.D, R, vars.data(), nv, sub.data(), c.cont32);
            for (i64 a = 0; a < nv - 1; ++a) for (i64 b = a + 1; b < nv; ++b) {
                double v = sub[(size_t)(a + b * nv)]; if (std::isnan(v)) v = 0.0;
                c.cor_mut[(size_t)(vars[a] + vars[b] * c.D.p)] = v; c.cor_mut[(size_t)(vars[b] + vars[a] * c.D.p)] = v;
            }
        }
        *out = test_fz_cond(c.C, X, Y, Zs, k, R.size(), n_obs_min);
    }
}

// test_subsets (tests.jl:281-346); rows as in fwo_test_cond.
void fwo_test_subsets(fwo_ctx* h, i64 X, i64 Y, const i64* Z_total, i64 m, int max_k, double alpha, i64 hps, i64 n_obs_min, i64 max_tests,
                      int trim_xy, const i32* rows, i64 n_rows, fwo_result* out, i64* out_Zs, int* out_k, i64* num_tests, double* frac) {
    Ctx& c = h->c;
    Rows R = make_rows(c, rows, n_rows);
    if (trim_xy) { if (needs_nz_view(c, X)) R = trim_rows(c.D, R, X); if (needs_nz_view(c, Y)) R = trim_rows(c.D, R, Y); }
    Params P; P.kind = c.kind; P.max_k = max_k; P.alpha = alpha; P.hps = hps; P.n_obs_min = n_obs_min; P.max_tests = max_tests;
    P.fdr = true; P.correct_reliable_only = true; P.fast_elim = true;
    DiscScratch s; if (is_discrete(c.kind)) s.init(c.max_level, max_k, c.D.n);
    if (c.kind == FZ_NZ) h->ensure_scratch();
    std::vector<i64> Z(Z_total, Z_total + m);
    SubsetsOut so = test_subsets(c, &s, X, Y, Z, R, P);
    *out = so.res; *out_k = so.k; for (int i = 0; i < 3; ++i) out_Zs[i] = so.Zs[i];
    *num_tests = so.num_tests; *frac = so.frac;
}

// pairwise stage -> CSR (offsets[p+1]; nbr/stat/adjp arrays sized by a first call with
// nbr == NULL, which returns the total count).  raw_* optional (n_pairs each).
i64 fwo_pairwise(fwo_ctx* h, double alpha, i64 hps, i64 n_obs_min, int fdr, int correct_reliable_only,
                 i64* offsets, i64* nbr, double* stat, double* adjp, double* raw_stats, double* raw_pvals) {
    Ctx& c = h->c;
    Params P; P.kind = c.kind; P.max_k = 0; P.alpha = alpha; P.hps = hps; P.n_obs_min = n_obs_min; P.max_tests = 0;
    P.fdr = fdr != 0; P.correct_reliable_only = correct_reliable_only != 0; P.fast_elim = true;
    NbrLists L; std::vector<double> rs, rp;
    pairwise(c, P, L, &rs, &rp);
    i64 tot = 0; for (auto& l : L) tot += (i64)l.size();
    if (offsets) { i64 o = 0; for (i64 v = 0; v < c.D.p; ++v) { offsets[v] = o; o += (i64)L[(size_t)v].size(); } offsets[c.D.p] = o; }
    if (nbr) { i64 o = 0; for (auto& l : L) for (auto& e : l) { nbr[o] = e.v; stat[o] = e.stat; adjp[o] = e.pval; ++o; } }
    if (raw_stats) memcpy(raw_stats, rs.data(), sizeof(double) * rs.size());
    if (raw_pvals) memcpy(raw_pvals, rp.data(), sizeof(double) * rp.size());
    return tot;
}

// learning.jl:51-61 automatic n_obs_min (applies to every kind because of the precedence quirk)
i64 fwo_auto_n_obs_min(fwo_ctx* h, int max_k, i64 hps) {
    Ctx& c = h->c;
    if (is_discrete(c.kind)) {
        i64 ml = 0; for (i32 v : c.levels) ml = std::max<i64>(ml, v);
        double pw = std::pow((double)ml, (double)max_k);
        i64 n_strata = (i64)std::min(pw, 8.0);
        return hps * 2 * 2 * n_strata;
    }
    return 20;
}

// HITON-PC for one target given its univariate neighbour list (hiton.jl:283-400), mode "single".
// Returns |PC|; pc_* sized >= n_uni.  whitelist optional.
i64 fwo_hiton_pc(fwo_ctx* h, i64 T, const i64* uni_nbr, const double* uni_stat, const double* uni_p, i64 n_uni,
                 int max_k, double alpha, i64 hps, i64 n_obs_min, i64 max_tests, const i64* whitelist, i64 n_wl,
                 i64* pc_nbr, double* pc_stat, double* pc_p, i64* num_tests) {
    Ctx& c = h->c;
    Params P; P.kind = c.kind; P.max_k = max_k; P.alpha = alpha; P.hps = hps; P.n_obs_min = n_obs_min; P.max_tests = max_tests;
    P.fdr = true; P.correct_reliable_only = true; P.fast_elim = true;
    DiscScratch s; if (is_discrete(c.kind)) s.init(c.max_level, max_k, c.D.n);
    if (c.kind == FZ_NZ) h->ensure_scratch();
    std::vector<Nbr> uni; for (i64 i = 0; i < n_uni; ++i) { Nbr nb = {uni_nbr[i], uni_stat[i], uni_p[i]}; uni.push_back(nb); }
    std::set<i64> wl; for (i64 i = 0; i < n_wl; ++i) wl.insert(whitelist[i]);
    HitonOut o = si_hiton_pc(c, &s, T, uni, P, wl);
    for (size_t i = 0; i < o.PC.size(); ++i) { pc_nbr[i] = o.PC[i].v; pc_stat[i] = o.PC[i].stat; pc_p[i] = o.PC[i].pval; }
    *num_tests = o.num_tests;
    return (i64)o.PC.size();
}

// The same for many targets at once (one OpenMP thread per target = the reference's one worker per target job,
// interleaved.jl:90): univariate lists as CSR over the listed targets (uni_off[n_targets + 1]); outputs use the same offsets
// (a target's PC is never longer than its candidate list).  Used by the sampled GPU-vs-oracle parity checks.
void fwo_hiton_pc_batch(fwo_ctx* h, i64 n_targets, const i64* targets, const i64* uni_off, const i64* uni_nbr, const double* uni_stat,
                        const double* uni_p, int max_k, double alpha, i64 hps, i64 n_obs_min, i64 max_tests, int n_threads,
                        i64* pc_count, i64* pc_nbr, double* pc_stat, double* pc_p, i64* num_tests) {
    Ctx& c = h->c;
    Params P; P.kind = c.kind; P.max_k = max_k; P.alpha = alpha; P.hps = hps; P.n_obs_min = n_obs_min; P.max_tests = max_tests;
    P.fdr = true; P.correct_reliable_only = true; P.fast_elim = true;
    const i64 p = c.D.p;
#pragma omp parallel num_threads(n_threads > 0 ? n_threads : 1)
    {
        DiscScratch s; if (is_discrete(c.kind)) s.init(c.max_level, max_k, c.D.n);
        Ctx local = c;
        std::vector<double> scratch;
        if (c.kind == FZ_NZ) { scratch.assign((size_t)(p * p), 0.0); local.cor_mut = scratch.data(); local.C.m = scratch.data(); }
        std::set<i64> empty;
#pragma omp for schedule(dynamic, 1)
        for (i64 t = 0; t < n_targets; ++t) {
            std::vector<Nbr> uni;
            for (i64 i = uni_off[t]; i < uni_off[t + 1]; ++i) { Nbr nb = {uni_nbr[i], uni_stat[i], uni_p[i]}; uni.push_back(nb); }
            HitonOut o = si_hiton_pc(local, &s, targets[t], uni, P, empty);
            pc_count[t] = (i64)o.PC.size(); num_tests[t] = o.num_tests;
            for (size_t i = 0; i < o.PC.size(); ++i) { pc_nbr[uni_off[t] + (i64)i] = o.PC[i].v; pc_stat[uni_off[t] + (i64)i] = o.PC[i].stat; pc_p[uni_off[t] + (i64)i] = o.PC[i].pval; }
        }
    }
}

// Full LGL (learning.jl:203-279).  n_obs_min < 0: automatic.  mode 0 "single", 1 "single_il" emulation.
// Edges returned as (a < b, weight); call with edge_a == NULL to get the count only is NOT supported:
// pass capacity >= p*(p-1)/2 or a known bound via max_edges (returns -1 if exceeded).
// target_subset (mode 0 only): run HITON only for these targets (bounded CPU-baseline sample).
// pc_offsets/pc_nbr/pc_stat/pc_p optional per-target PC lists (pc capacity = max_pc entries).
i64 fwo_lgl(fwo_ctx* h, int max_k, double alpha, i64 hps, i64 n_obs_min, i64 max_tests, int fdr, int mode, int n_threads,
            const i64* target_subset, i64 n_target_subset,
            i64* edge_a, i64* edge_b, double* edge_w, i64 max_edges, i64* cond_tests, i64* pair_tests,
            i64* pc_offsets, i64* pc_nbr, double* pc_stat, double* pc_p, i64 max_pc, double* phase_secs /* cor, pairwise, hiton */) {
    Ctx& c = h->c;
    Params P; P.kind = c.kind; P.max_k = max_k; P.alpha = alpha; P.hps = hps; P.max_tests = max_tests;
    P.fdr = fdr != 0; P.correct_reliable_only = true; P.fast_elim = true;
    P.n_obs_min = n_obs_min < 0 ? fwo_auto_n_obs_min(h, max_k, hps) : n_obs_min;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    if (phase_secs) phase_secs[0] = phase_secs[1] = phase_secs[2] = 0.0;
    if (c.kind == FZ && !c.C.m) { double t0 = now_s(); fwo_compute_cor(h); if (phase_secs) phase_secs[0] = now_s() - t0; }
    std::vector<Edge> edges; std::vector<HitonOut> per;
    lgl(c, P, mode, edges, cond_tests, pair_tests, n_threads, target_subset, n_target_subset, &per, phase_secs);
    if ((i64)edges.size() > max_edges) return -1;
    std::sort(edges.begin(), edges.end(), [](const Edge& x, const Edge& y) { return x.a != y.a ? x.a < y.a : x.b < y.b; });
    for (size_t i = 0; i < edges.size(); ++i) { edge_a[i] = edges[i].a; edge_b[i] = edges[i].b; edge_w[i] = edges[i].w; }
    if (pc_offsets) {
        i64 o = 0;
        for (i64 t = 0; t < c.D.p; ++t) {
            pc_offsets[t] = o;
            for (const Nbr& nb : per[(size_t)t].PC) { if (o < max_pc) { pc_nbr[o] = nb.v; pc_stat[o] = nb.stat; pc_p[o] = nb.pval; } ++o; }
        }
        pc_offsets[c.D.p] = o;
    }
    return (i64)edges.size();
}

// Exhaustive check of the divider-free round(x, digits=5) used by the CUDA path (csrc/fz.cuh round5f / round5d): for every
// integer |k| <= limit, q0 = k * RN(1e-5), r = fma(-q0, 1e5, k), q = fma(r, RN(1e-5), q0) must equal the correctly rounded k / 1e5
// (what Julia's round(x * 1e5) / 1e5 computes) in Float32 and in Float64.  Returns the number of mismatches.
long long fwo_round5_shortcut_mismatches(long long limit) {
    long long bad = 0;
    const double invd = 1.0 / 100000.0; const float invf = 1.0f / 100000.0f;
    for (long long k = -limit; k <= limit; ++k) {
        const double y = (double)k, q0 = y * invd, r = std::fma(-q0, 100000.0, y), q = std::fma(r, invd, q0);
        if ((k != 0 && q != y / 100000.0) || (k == 0 && q != 0.0)) ++bad;
        const float yf = (float)k, q0f = yf * invf, rf = std::fmaf(-q0f, 100000.0f, yf), qf = std::fmaf(rf, invf, q0f);
        if ((k != 0 && qf != yf / 100000.0f) || (k == 0 && qf != 0.0f)) ++bad;
    }
    return bad;
}

int fwo_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
