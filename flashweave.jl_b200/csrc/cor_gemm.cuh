// cor_gemm.cuh — cor_mat = Float32.(cor(data))  (src/learning.jl:42-44; Statistics.cor).
//
// Step 1 (HBM-bound, one pass over the table + one write): per-column mean and centred norm in
// fp64, standardised columns z = (x - mean) / ||x - mean|| written K-major ([p][n_pad]).
// Step 2: C = Z Z^T (upper-triangular tiles, mirrored), clamped to [-1, 1], unit diagonal
// (cov2cor!).  A constant column has ||.|| = 0 -> NaN row/column, as in the reference.
#pragma once
#include <string>
#include "common.cuh"

struct CorGemmScratch {
    float* z = nullptr; size_t z_elems = 0;
    cudaError_t reserve(size_t n) {
        if (n <= z_elems && z) return cudaSuccess;
        if (z) cudaFree(z);
        z = nullptr; z_elems = 0;
        cudaError_t e = cudaMalloc((void**)&z, n * sizeof(float));
        if (e == cudaSuccess) z_elems = n;
        return e;
    }
    ~CorGemmScratch() { if (z) cudaFree(z); }
};

// one CTA per column
template <int THREADS>
__global__ void __launch_bounds__(THREADS) cor_standardize_kernel(const float* __restrict__ data, i64 n, i64 ld, i64 kp, float* __restrict__ z) {
    const i64 col = blockIdx.x;
    const float* x = data + col * ld;
    float* zc = z + col * kp;
    __shared__ double red[THREADS / 32];
    __shared__ double s_mean, s_inv;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double s = 0.0;
    for (i64 i = tid; i < n; i += THREADS) s += (double)x[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (tid == 0) { double t = 0.0; for (int w = 0; w < THREADS / 32; ++w) t += red[w]; s_mean = t / (double)n; }
    __syncthreads();
    const double mean = s_mean;
    double ss = 0.0;
    for (i64 i = tid; i < n; i += THREADS) { double d = (double)x[i] - mean; ss += d * d; }
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_down_sync(0xffffffffu, ss, o);
    __syncthreads();
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    if (tid == 0) { double t = 0.0; for (int w = 0; w < THREADS / 32; ++w) t += red[w]; s_inv = 1.0 / sqrt(t); }
    __syncthreads();
    const double inv = s_inv;
    for (i64 i = tid; i < kp; i += THREADS) zc[i] = (i < n) ? (float)(((double)x[i] - mean) * inv) : 0.0f;
}

// fp32 SIMT tile GEMM (first correct version; the tensor-core kernel replaces it)
__global__ void __launch_bounds__(256) cor_gemm_simt_kernel(const float* __restrict__ Z, i64 p, i64 kp, float* __restrict__ C) {
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj < bi) return;
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int lrow = tid >> 2, lk = (tid & 3) * 4;
    const i64 arow = (i64)bi * 64 + lrow, brow = (i64)bj * 64 + lrow;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    for (i64 k0 = 0; k0 < kp; k0 += 16) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (arow < p) a = *reinterpret_cast<const float4*>(Z + arow * kp + k0 + lk);
        if (brow < p) b = *reinterpret_cast<const float4*>(Z + brow * kp + k0 + lk);
        As[lk + 0][lrow] = a.x; As[lk + 1][lrow] = a.y; As[lk + 2][lrow] = a.z; As[lk + 3][lrow] = a.w;
        Bs[lk + 0][lrow] = b.x; Bs[lk + 1][lrow] = b.y; Bs[lk + 2][lrow] = b.z; Bs[lk + 3][lrow] = b.w;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float ar[4], br[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { ar[i] = As[kk][ty * 4 + i]; br[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            i64 r = (i64)bi * 64 + ty * 4 + i, c = (i64)bj * 64 + tx * 4 + j;
            if (r < p && c < p) {
                float v = acc[i][j];
                v = v > 1.0f ? 1.0f : (v < -1.0f ? -1.0f : v);   // clampcor; NaN passes through
                if (r == c) v = 1.0f;                             // cov2cor!: C[j,j] = 1
                C[r * p + c] = v;
                C[c * p + r] = v;
            }
        }
}

static cudaError_t cor_gemm_run(CorGemmScratch& S, const float* d_data, i64 n, i64 p, i64 ld, float* d_cor, int sm_count, cudaStream_t st,
                                int* n_launch, std::string* msg) {
    (void)sm_count;
    const i64 kp = (n + 15) / 16 * 16;
    cudaError_t e = S.reserve((size_t)p * kp);
    if (e != cudaSuccess) { *msg = "scratch allocation"; return e; }
    cor_standardize_kernel<256><<<(unsigned)p, 256, 0, st>>>(d_data, n, ld, kp, S.z);
    (*n_launch)++;
    e = cudaGetLastError(); if (e != cudaSuccess) { *msg = "cor_standardize_kernel"; return e; }
    const unsigned nb = (unsigned)((p + 63) / 64);
    cor_gemm_simt_kernel<<<dim3(nb, nb), 256, 0, st>>>(S.z, p, kp, d_cor);
    (*n_launch)++;
    e = cudaGetLastError(); if (e != cudaSuccess) { *msg = "cor_gemm_simt_kernel"; return e; }
    return cudaSuccess;
}
