// common.cuh — shared types for the fwgpu kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef long long i64;
typedef unsigned long long u64;

#define FW_INF_IDX 0x7fffffffffffffffLL

// device-side image of src/types.jl:140-145 TestResult (same 32-byte layout as fw_test_result)
struct DevResult {
    double stat;
    double pval;
    i64 df;
    unsigned char suff_power;
    unsigned char pad_[7];
};

__device__ __forceinline__ DevResult make_result(double s, double p, i64 df, bool sp) {
    DevResult r;
    r.stat = s; r.pval = p; r.df = df; r.suff_power = sp ? 1 : 0;
#pragma unroll
    for (int i = 0; i < 7; ++i) r.pad_[i] = 0;
    return r;
}

// ---- the resident correlation matrix as the test kernels see it ------------------------------------------------------------
// One GPU (world == 1): the full symmetric p x p Float32 matrix at shard[0].
// Several GPUs: cor_mat stays ROW-SHARDED where the GEMM produced it and is never exchanged.  Only the upper triangle exists:
// the 128-row tile rows are split into 2*world contiguous groups of h tile rows, rank r owns groups r and 2*world-1-r (equal
// tile counts), stored back to back in its shard; shard[q] is rank q's buffer mapped into this process (CUDA IPC / peer access),
// so a lookup is one NVLink read of the owner's HBM.  r(a, b) lives in row min(a, b), column max(a, b).
constexpr int FW_MAX_RANKS = 8;
struct CorView {
    const float* shard[FW_MAX_RANKS];
    i64 p;
    int world, h;
    __device__ __forceinline__ float at(i64 a, i64 b) const {
        if (world == 1) return __ldg(shard[0] + a * p + b);
        const i64 lo = a < b ? a : b, hi = a < b ? b : a;
        const int tr = (int)(lo >> 7), g = tr / h;
        const int owner = g < world ? g : 2 * world - 1 - g;
        const i64 lrow = (i64)((g < world ? 0 : h) + (tr - g * h)) * 128 + (lo & 127);
        return shard[owner][lrow * p + hi];
    }
    // first local row of tile row `tr` in its owner's shard, and the owner (host + device)
    __host__ __device__ __forceinline__ int owner_of_tile_row(int tr) const { const int g = tr / h; return g < world ? g : 2 * world - 1 - g; }
    __host__ __device__ __forceinline__ i64 local_row_of_tile_row(int tr) const { const int g = tr / h; return (i64)((g < world ? 0 : h) + (tr - g * h)) * 128; }
};

// ---- pairwise stage of the table-based kinds (mi, mi_nz, fz_nz) split over the GPUs of a group ------------------------------------
// The X variables (first member of a pair X < Y) are dealt in groups of PW_X_GROUP = 1024 (8 tile rows of the tensor-core pre-filter)
// in a snake 0..N-1, N-1..0: row X has p - X - 1 partners, so pairing an early with a late group balances the ranks.
constexpr int PW_X_GROUP = 1024;
__host__ __device__ __forceinline__ bool pw_owns_group(i64 g, int rank, int world) {
    if (world <= 1) return true;
    const int c = (int)(g % (2 * world));
    return (c < world ? c : 2 * world - 1 - c) == rank;
}

// ---- raw candidates of the univariate Fisher-z stage ---------------------------------------------------------------------------
// A pair whose |r| reaches the (conservatively lowered) significance threshold, as the cor_mat GEMM epilogue (cor_tc.cuh) or the
// one-pass scan of the resident matrix (pairwise.cuh) appends it: unordered, 12 bytes.  counters[0] = records appended (it keeps
// counting past `cap`: the host then knows the size to retry with), counters[1] = NaN correlations seen in the upper triangle.
struct PwRec { int x, y; float r; };
struct PwEmit { PwRec* list; u64* counters; i64 cap; float r_lo; int on; };

__device__ __forceinline__ i64 choose2(i64 n) { return n < 2 ? 0 : n * (n - 1) / 2; }
__device__ __forceinline__ i64 choose3(i64 n) { return n < 3 ? 0 : n * (n - 1) * (n - 2) / 6; }

// Lexicographic unranking of pairs (a < b) out of n elements: rank q in [0, C(n,2)).
// Number of pairs whose first element is < a:  off(a) = a*n - a*(a+1)/2.
__device__ __forceinline__ void unrank2(i64 q, int n, int& a, int& b) {
    double fn = 2.0 * (double)n - 1.0;
    double disc = fn * fn - 8.0 * (double)q;
    int aa = (int)((fn - sqrt(disc > 0.0 ? disc : 0.0)) * 0.5);
    if (aa < 0) aa = 0;
    if (aa > n - 2) aa = n - 2;
    while (aa > 0 && (i64)aa * n - (i64)aa * (aa + 1) / 2 > q) --aa;
    while (aa < n - 2 && (i64)(aa + 1) * n - (i64)(aa + 1) * (aa + 2) / 2 <= q) ++aa;
    a = aa;
    b = (int)(q - ((i64)aa * n - (i64)aa * (aa + 1) / 2)) + aa + 1;
}
// 32-bit variant for n <= 4096 (q < 2^23: exactly representable in fp32, the estimate is off by at most 1)
__device__ __forceinline__ void unrank2_small(int q, int n, int& a, int& b) {
    const float fn = 2.0f * (float)n - 1.0f;
    const float disc = fmaf(fn, fn, -8.0f * (float)q);
    int aa = (int)((fn - sqrtf(fmaxf(disc, 0.0f))) * 0.5f);
    aa = max(0, min(aa, n - 2));
    int off = aa * n - ((aa * (aa + 1)) >> 1);
    if (off > q) { --aa; off = aa * n - ((aa * (aa + 1)) >> 1); }
    if (off > q) { --aa; off = aa * n - ((aa * (aa + 1)) >> 1); }
    int off1 = (aa + 1) * n - (((aa + 1) * (aa + 2)) >> 1);
    if (aa < n - 2 && off1 <= q) { ++aa; off = off1; off1 = (aa + 1) * n - (((aa + 1) * (aa + 2)) >> 1); }
    if (aa < n - 2 && off1 <= q) { ++aa; off = off1; }
    a = aa;
    b = q - off + aa + 1;
}
