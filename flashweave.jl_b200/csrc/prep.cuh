// prep.cuh — the normalisation step in front of the CI-test hot path, on the device.
//
// Replaces, for dense tables without meta variables (reference paths relative to the FlashWeave.jl checkout):
//   preprocess_data / normalize_data           src/preprocessing.jl:412-563, 660-684
//   filter_by_variance                         src/preprocessing.jl:367-409
//   rownorm! ("tss")                           src/preprocessing.jl:348
//   clr! (ignore_zeros) / adaptive_clr!        src/preprocessing.jl:133-215
//   discretize / discretize_nz (tied ranks)    src/preprocessing.jl:217-292
//   presabs_norm! + level filters              src/preprocessing.jl:364-365, 475-520
//
// The table is the column-major n x p Matrix{Float32} the reference converts its input to (check_convert_sparse), one
// variable per contiguous row here.  O(n p) work runs on the device — distinct-value / row-sum filters, per-sample log sums,
// the element-wise transforms, the per-variable tied ranks (one segmented radix sort of the non-zero values, then two binary
// searches per element) — and the O(n) per-sample parameters (geometric means, adaptive pseudo-counts; fp64 exp/log exactly as
// the reference evaluates them) on the host between two launches.  The normalised table is written straight into the
// context's resident table (fw_set_data_f32 / fw_set_data_i32 semantics), so it never travels back to the host unless asked for.
#pragma once
#include "common.cuh"

enum { PREP_ROWS = 0, PREP_CLR_ADAPT = 1, PREP_CLR_NZ = 2, PREP_BINARY = 3, PREP_BINNED_NZ_CLR = 4, PREP_BINNED_NZ_ROWS = 5 };

// var(column) > 0  <=>  the column holds at least two distinct values (exact for count data)
__global__ void __launch_bounds__(256) prep_col_distinct_kernel(const float* __restrict__ x, i64 n, i64 ld, unsigned char* __restrict__ colflag) {
    const i64 v = blockIdx.x;
    const float* c = x + v * ld;
    const float first = c[0];
    int any = 0;
    for (i64 i = threadIdx.x; i < n; i += blockDim.x) any |= (c[i] != first);
    any = __syncthreads_or(any);
    if (threadIdx.x == 0) colflag[v] = any ? 1 : 0;
}

// one thread per sample over the kept variables (ascending): Float32 row sum (rownorm!'s divisor), Float64 row sum (depth),
// number of zeros and sum of log over the non-zero entries (pseudocount_vars_from_sample, preprocessing.jl:133-139)
__global__ void __launch_bounds__(256) prep_row_stats_kernel(const float* __restrict__ x, i64 n, i64 ld, const int* __restrict__ cols, i64 p1,
                                                             float* __restrict__ sum32, double* __restrict__ sum64, int* __restrict__ nzero, double* __restrict__ slog) {
    const i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    float s32 = 0.0f; double s64 = 0.0, sl = 0.0; int nz0 = 0;
    for (i64 j = 0; j < p1; ++j) {
        const float v = x[(i64)cols[j] * ld + r];
        s32 += v; s64 += (double)v;
        if (v == 0.0f) ++nz0; else sl += log((double)v);
    }
    sum32[r] = s32; sum64[r] = s64; nzero[r] = nz0; slog[r] = sl;
}

// global minimum of the non-zero entries of the kept block (adaptive_pseudocount!, preprocessing.jl:160): float bits of a
// positive float order like unsigned integers
__global__ void __launch_bounds__(256) prep_min_nonzero_kernel(const float* __restrict__ x, i64 ld, const int* __restrict__ cols, i64 p1,
                                                               const int* __restrict__ rows, i64 n1, unsigned int* __restrict__ out_bits) {
    const i64 j = blockIdx.x;
    const float* c = x + (i64)cols[j] * ld;
    unsigned int best = 0x7f800000u;
    for (i64 i = threadIdx.x; i < n1; i += blockDim.x) { const float v = c[rows[i]]; if (v != 0.0f) best = min(best, __float_as_uint(fabsf(v))); }
    best = __reduce_min_sync(0xffffffffu, best);
    if ((threadIdx.x & 31) == 0 && best != 0x7f800000u) atomicMin(out_bits, best);
}

// continuous transforms into the compacted table out[c][r] (n1 rows per variable)
//   PREP_ROWS:      x / sum32[row]                       (Float32 division, rownorm!)
//   PREP_CLR_NZ:    x != 0 ? log(x / g[row]) : 0         (Float64, then Float32)
//   PREP_CLR_ADAPT: log((x != 0 ? x : pc[row]) / g[row])
__global__ void __launch_bounds__(256) prep_transform_kernel(const float* __restrict__ x, i64 ld, const int* __restrict__ cols, const int* __restrict__ rows, i64 n1,
                                                             int mode, const float* __restrict__ sum32, const double* __restrict__ g, const double* __restrict__ pc,
                                                             float* __restrict__ out) {
    const i64 j = blockIdx.x;                                 // grid: (variables, row blocks)
    const i64 i = (i64)blockIdx.y * blockDim.x + threadIdx.x;
    if (i >= n1) return;
    const int r = rows[i];
    const float v = x[(i64)cols[j] * ld + r];
    float o;
    if (mode == PREP_ROWS) o = __fdiv_rn(v, sum32[r]);
    else if (mode == PREP_CLR_NZ) o = v != 0.0f ? (float)log((double)v / g[r]) : 0.0f;
    else o = (float)log((v != 0.0f ? (double)v : pc[r]) / g[r]);
    out[j * n1 + i] = o;
}

// values to be ranked (Float64): the normalised value where the RAW entry is non-zero (nz_mask of preprocessing.jl:495),
// +inf elsewhere so that a sort puts them behind the segment's real values; counts the non-zero entries per variable
__global__ void __launch_bounds__(256) prep_rank_values_kernel(const float* __restrict__ x, i64 ld, const int* __restrict__ cols, const int* __restrict__ rows, i64 n1,
                                                               int mode, const float* __restrict__ sum32, const double* __restrict__ g,
                                                               double* __restrict__ vals, int* __restrict__ nnz) {
    const i64 j = blockIdx.x;                                 // grid: (variables, row blocks)
    const i64 i = (i64)blockIdx.y * blockDim.x + threadIdx.x;
    int c = 0;
    if (i < n1) {
        const int r = rows[i];
        const float v = x[(i64)cols[j] * ld + r];
        double o = __longlong_as_double(0x7ff0000000000000LL);
        if (v != 0.0f) { o = (mode == PREP_BINNED_NZ_CLR) ? log((double)v / g[r]) : (double)__fdiv_rn(v, sum32[r]); c = 1; }
        vals[j * n1 + i] = o;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&nnz[j], c);
}

// discretize_nz (preprocessing.jl:280-292, :238-253): tied rank of the value among the variable's non-zero entries
// (average of the 1-based positions of equal values), divided by the largest rank, floor(. / step) + 1; 0 where the raw entry is 0
__global__ void __launch_bounds__(256) prep_bin_kernel(const double* __restrict__ vals, const double* __restrict__ sorted, const int* __restrict__ nnz, i64 n1,
                                                       int n_bins, int* __restrict__ out) {
    const i64 j = blockIdx.x;                                 // grid: (variables, row blocks)
    const i64 i = (i64)blockIdx.y * blockDim.x + threadIdx.x;
    if (i >= n1) return;
    const double v = vals[j * n1 + i];
    const int m = nnz[j];
    int o = 0;
    if (!isinf(v) && m > 0) {
        const double* s = sorted + j * n1;
        int lo = 0, hi = m;                                   // first index with s[idx] >= v
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (s[mid] < v) lo = mid + 1; else hi = mid; }
        const int first = lo;
        hi = m;                                               // first index with s[idx] > v
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (s[mid] <= v) lo = mid + 1; else hi = mid; }
        const int last = lo;                                  // equal values occupy [first, last)
        const double rank = 0.5 * (double)((first + 1) + last);
        // largest rank = tied rank of the maximum
        const double vmax = s[m - 1];
        int l2 = 0, h2 = m;
        while (l2 < h2) { const int mid = (l2 + h2) >> 1; if (s[mid] < vmax) l2 = mid + 1; else h2 = mid; }
        const double rmax = 0.5 * (double)((l2 + 1) + m);
        const double step = (1.0 / (double)(n_bins - 1)) + 1e-5;
        o = (int)floor((rank / rmax) / step) + 1;
    }
    out[j * n1 + i] = o;
}

__global__ void __launch_bounds__(256) prep_binary_kernel(const float* __restrict__ x, i64 ld, const int* __restrict__ cols, const int* __restrict__ rows, i64 n1,
                                                          int* __restrict__ out) {
    const i64 j = blockIdx.x;                                 // grid: (variables, row blocks)
    const i64 i = (i64)blockIdx.y * blockDim.x + threadIdx.x;
    if (i >= n1) return;
    const float v = x[(i64)cols[j] * ld + rows[i]];
    out[j * n1 + i] = v > 0.0f ? 1 : (v < 0.0f ? -1 : 0);    // sign
}

// set of level codes (0 .. 31) seen per variable, as a bit mask
__global__ void __launch_bounds__(256) prep_col_levels_kernel(const int* __restrict__ t, i64 n1, unsigned int* __restrict__ seen) {
    const i64 j = blockIdx.x;
    unsigned int m = 0;
    for (i64 i = threadIdx.x; i < n1; i += blockDim.x) { const int c = t[j * n1 + i]; m |= (c >= 0 && c < 31) ? (1u << c) : 0x80000000u; }
    m = __reduce_or_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m) atomicOr(&seen[j], m);
}

__global__ void __launch_bounds__(256) prep_gather_cols_kernel(const int* __restrict__ src, const int* __restrict__ colmap, i64 n1, int* __restrict__ dst) {
    const i64 j = blockIdx.x;                                 // grid: (variables, row blocks)
    const i64 i = (i64)blockIdx.y * blockDim.x + threadIdx.x;
    if (i < n1) dst[j * n1 + i] = src[(i64)colmap[j] * n1 + i];
}

// ---- sparse input: SparseMatrixCSC{T,Int64} (colptr, rowval, nzval; 0- or 1-based) -> the dense column-major table ------------
// (the reference densifies column views the same way for the dense test kernels; stored zeros stay zeros)
template <class T>
__global__ void __launch_bounds__(256) csc_scatter_kernel(const i64* __restrict__ colptr, const i64* __restrict__ rowval, const T* __restrict__ nzval,
                                                          i64 n, i64 p, i64 base, T* __restrict__ dense, int* __restrict__ bad) {
    const i64 v = blockIdx.x;
    const i64 b = colptr[v] - base, e = colptr[v + 1] - base;
    for (i64 k = b + threadIdx.x; k < e; k += blockDim.x) {
        const i64 r = rowval[k] - base;
        if (r < 0 || r >= n) { atomicOr(bad, 1); continue; }
        dense[v * n + r] = nzval[k];
    }
}
