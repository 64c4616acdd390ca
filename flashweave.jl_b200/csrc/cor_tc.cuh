// cor_tc.cuh — correlation matrix on the 5th-gen tensor cores (tcgen05 / TMEM / TMA), sm_100a.
//
// cor_mat = Float32.(cor(data))  (src/learning.jl:42-44).  With standardised columns
// z_i = (x_i - mean_i)/||x_i - mean_i|| the matrix is the Gram matrix C = Z Z^T (SYRK-shaped).
// bf16 inputs alone are too coarse for the 1e-5 target, so each z is split
// z = hi + lo (two bf16) and three MMAs accumulate hi*hi + hi*lo + lo*hi into one fp32 TMEM
// accumulator (the lo*lo term is below 2^-16 relative).
//
// Layout: Zhi, Zlo are K-major [p_pad][kp] bf16 (one variable per row; kp = n rounded up to 64,
// p_pad = p rounded up to 128, zero padded), so A and B tiles of C = A B^T are both K-major TMA
// boxes of 128 rows x 64 elements (128 bytes, SWIZZLE_128B).  One CTA computes one 128x128 tile of
// the upper triangle and mirrors it; warp 0 = TMA producer, warp 1 = MMA issuer (one elected
// thread, tcgen05.mma.cta_group::1.kind::f16, M=128 N=128 K=16), warps 2-5 = epilogue
// (tcgen05.ld 32x32b drains into fp32 registers, clamp to [-1,1], unit diagonal, stores of tile and mirror).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <algorithm>
#include <cstdlib>
#include <string>
#include "common.cuh"

namespace cortc {

constexpr int BM = 128, BN = 128, BK = 64, STAGES = 3;
constexpr int TILE_BYTES = BM * BK * 2;                // 16 KB per operand tile
constexpr int STAGE_BYTES = 4 * TILE_BYTES;            // A_hi, A_lo, B_hi, B_lo
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start >> 4 | LBO (1 = 16 B, ignored for swizzled K-major) << 16 | SBO (1024 B = 8 rows of 128 B) << 32 | version 1 << 46 | SWIZZLE_128B (2) << 61
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// one CTA per column: standardise and split into bf16 hi/lo, K-major rows of length kp (zero padded)
template <int THREADS>
__global__ void __launch_bounds__(THREADS) standardize_split_kernel(const float* __restrict__ data, i64 n, i64 ld, i64 p, i64 kp,
                                                                    __nv_bfloat16* __restrict__ zhi, __nv_bfloat16* __restrict__ zlo, i64 col0) {
    const i64 col = col0 + blockIdx.x;
    __nv_bfloat16* hi = zhi + col * kp;
    __nv_bfloat16* lo = zlo + col * kp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (col >= p) {                                            // padding rows of the last tile
        for (i64 i = tid; i < kp; i += THREADS) { hi[i] = __float2bfloat16_rn(0.f); lo[i] = __float2bfloat16_rn(0.f); }
        return;
    }
    const float* x = data + col * ld;
    __shared__ double red[THREADS / 32];
    __shared__ double s_mean, s_inv;
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
    for (i64 i = tid; i < kp; i += THREADS) {
        float z = (i < n) ? (float)(((double)x[i] - mean) * inv) : 0.0f;
        __nv_bfloat16 h = __float2bfloat16_rn(z);
        hi[i] = h;
        lo[i] = __float2bfloat16_rn(z - __bfloat162float(h));
    }
}

// Accumulation precision.  tcgen05.mma adds each K=16 partial product into the fp32 TMEM accumulator with
// truncation (measured on B200: bias = -|r| * (#accumulations) * ulp/2, i.e. -5.6e-5 at n = 10^4 when all three
// split terms share one accumulator).  Therefore:
//   * the large hi*hi sum goes to a double-buffered "main" accumulator that the epilogue warps drain into fp32
//     registers (round-to-nearest adds) every CHUNK k-blocks, so a TMEM accumulator only ever holds a small
//     partial sum (64 accumulations of magnitude <= |r| * CHUNK*64/n);
//   * the cross terms hi*lo + lo*hi (2^-8 smaller) go to their own accumulator, read once at the end.
// Residual bias at n = 10^4, |r| -> 1: ~2e-6 (was 5.6e-5).
constexpr int CHUNK = 16;                                  // k-blocks per drained partial sum (1024 samples)
constexpr int COR_GROUP = 8;                               // tile rows per rasterisation group of the cluster kernel
constexpr int NTHREADS = 192;                              // warp 0: TMA, warp 1: MMA, warps 2-5: epilogue
constexpr int TMEM_ALLOC = 512;                            // main[0] @0, main[1] @128, cross @256

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Univariate Fisher-z stage fused into the epilogue (tests.jl:470-478 looks the same correlations up again): the 32 finished
// correlations of this thread's row are still in registers (v[j] = clamped r(row, col0 + j), hitmask bit j = "|r| >= r_lo"), so
// the raw candidates are appended to the list here instead of re-reading the 10 GB matrix: per warp and 32-column chunk one
// exclusive scan of the hit counts and ONE global atomic, then every lane writes its own records.
__device__ __forceinline__ void emit_chunk(const PwEmit& em, const uint32_t* v, unsigned int hitmask, i64 row, i64 col0, int lane) {
    const unsigned int n_hit = __popc(hitmask);
    unsigned int incl = n_hit;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    const unsigned int total = __shfl_sync(0xffffffffu, incl, 31);
    if (!total) return;                                  // warp-uniform
    u64 base = 0;
    if (lane == 31) base = atomicAdd(&em.counters[0], (u64)total);
    base = __shfl_sync(0xffffffffu, base, 31);
    u64 pos = base + incl - n_hit;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        if ((hitmask >> j) & 1u) {
            if ((i64)pos < em.cap) { PwRec rec; rec.x = (int)row; rec.y = (int)(col0 + j); rec.r = __uint_as_float(v[j]); em.list[pos] = rec; }
            ++pos;
        }
    }
}

// Tiles: upper-triangular tile rows [bi0, bi0 + nbi) of an nb x nb tile grid (blockIdx.x enumerates them row-major).  mirror != 0
// also writes the transposed tile (single-GPU mode); mirror == 0 (row-sharded mode) leaves the lower triangle to
// cor_symmetrize_kernel after the ranks have exchanged their row blocks, and only mirrors inside diagonal tiles.
__global__ void __launch_bounds__(NTHREADS, 1) cor_tc_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
                                                             float* __restrict__ C, i64 p, int num_kb, int nb, int bi0, int mirror, const PwEmit em, const int sh_world, const int sh_h) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;                       // SWIZZLE_128B tiles need 1024-byte alignment
    uint8_t* gen = smem_raw + (base - raw);
    // barriers: full[STAGES], empty[STAGES], cfull[2], cempty[2]; then the TMEM base-address slot
    const uint32_t bar0 = base + STAGES * STAGE_BYTES;
    const uint32_t bar_full = bar0, bar_empty = bar0 + 8 * STAGES, bar_cfull = bar0 + 16 * STAGES, bar_cempty = bar0 + 16 * STAGES + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + STAGES * STAGE_BYTES + 16 * STAGES + 32);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // tile of the upper triangle: row bi has nb - bi tiles
    int bi, bj;
    {
        // global index of this tile in the row-major enumeration of the whole upper triangle
        const long long t = (long long)blockIdx.x + ((long long)bi0 * nb - (long long)bi0 * (bi0 - 1) / 2);
        const double f = 2.0 * nb + 1.0;
        int r = (int)((f - sqrt(f * f - 8.0 * (double)t)) * 0.5);
        if (r < 0) r = 0;
        if (r > nb - 1) r = nb - 1;
        while (r > 0 && (long long)r * nb - (long long)r * (r - 1) / 2 > t) --r;
        while (r < nb - 1 && (long long)(r + 1) * nb - (long long)(r + 1) * r / 2 <= t) ++r;
        bi = r; bj = r + (int)(t - ((long long)r * nb - (long long)r * (r - 1) / 2));
    }
    const int n_chunks = (num_kb + CHUNK - 1) / CHUNK;
    // row-sharded mode (several GPUs): tile row bi is stored at its local position in this rank's shard (common.cuh, CorView)
    float* Cw = C;
    if (sh_world > 1) { const int g = bi / sh_h; Cw = C + ((i64)((g < sh_world ? 0 : sh_h) + (bi - g * sh_h)) * 128 - (i64)bi * 128) * p; }

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_cfull + 8 * b, 1); mbar_init(bar_cempty + 8 * b, 4); }   // 4 epilogue warps arrive
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_lo) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_ALLOC) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (uint32_t)((kb / STAGES) & 1);
                mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                const uint32_t full = bar_full + 8 * s;
                mbar_expect_tx(full, STAGE_BYTES);
                const uint32_t st = base + s * STAGE_BYTES;
                tma_load_2d(st, &tm_hi, full, kb * BK, bi * BM);
                tma_load_2d(st + TILE_BYTES, &tm_lo, full, kb * BK, bi * BM);
                tma_load_2d(st + 2 * TILE_BYTES, &tm_hi, full, kb * BK, bj * BN);
                tma_load_2d(st + 3 * TILE_BYTES, &tm_lo, full, kb * BK, bj * BN);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            // cute::UMMA::InstrDescriptor: c_format F32 (1) << 4 | a_format BF16 (1) << 7 | b_format BF16 (1) << 10 | K-major A, B | N >> 3 << 17 | M >> 4 << 24
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            const uint32_t t_cross = tmem_base + 256;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (uint32_t)((kb / STAGES) & 1);
                const int c = kb / CHUNK, b = c & 1, u = c >> 1;
                const bool first = (kb % CHUNK) == 0;
                if (first && c >= 2) {                                            // main[b] must have been drained (drain u-1)
                    mbar_wait(bar_cempty + 8 * b, (uint32_t)((u - 1) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                mbar_wait(bar_full + 8 * s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t t_main = tmem_base + 128u * (uint32_t)b;
                const uint32_t st = base + s * STAGE_BYTES;
                const uint64_t a_hi = make_desc(st), a_lo = make_desc(st + TILE_BYTES);
                const uint64_t b_hi = make_desc(st + 2 * TILE_BYTES), b_lo = make_desc(st + 3 * TILE_BYTES);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);          // 32 bytes per UMMA_K step inside the 128-byte swizzle row
                    umma_bf16(t_main, a_hi + adv, b_hi + adv, idesc, (first && k == 0) ? 0u : 1u);
                    umma_bf16(t_cross, a_hi + adv, b_lo + adv, idesc, (kb == 0 && k == 0) ? 0u : 1u);
                    umma_bf16(t_cross, a_lo + adv, b_hi + adv, idesc, 1u);
                }
                umma_commit(bar_empty + 8 * s);                                   // frees the smem stage when these MMAs retire
                if ((kb % CHUNK) == CHUNK - 1 || kb == num_kb - 1) umma_commit(bar_cfull + 8 * b);   // partial sum of chunk c complete
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue warps: drain main[b] per chunk into registers, add the cross accumulator, write tile + mirror =====
        const int q = warp & 3;                                                   // TMEM lane quarter this warp may access
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        float acc[BN];
#pragma unroll
        for (int j = 0; j < BN; ++j) acc[j] = 0.0f;
        for (int c = 0; c < n_chunks; ++c) {
            const int b = c & 1, u = c >> 1;
            mbar_wait(bar_cfull + 8 * b, (uint32_t)(u & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + lane_base + 128u * (uint32_t)b + (uint32_t)c0, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(v[j]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_cempty + 8 * b);
        }
        // the last chunk's commit also covers every earlier MMA, including the cross accumulator
        const i64 row = (i64)bi * BM + q * 32 + lane;
        const bool diag = (bi == bj);
        unsigned int n_nan = 0;
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + lane_base + 256u + (uint32_t)c0, v);
            unsigned int hitmask = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const i64 col = (i64)bj * BN + c0 + j;
                float x = acc[c0 + j] + __uint_as_float(v[j]);
                x = x > 1.0f ? 1.0f : (x < -1.0f ? -1.0f : x);        // clampcor (NaN passes through)
                if (row == col) x = 1.0f;                              // cov2cor!: unit diagonal
                const bool ok = row < p && col < p && (!diag || col >= row);
                if (ok) {
                    Cw[row * p + col] = x;
                    if (mirror || diag) Cw[col * p + row] = x;         // mirror (coalesced across the warp: consecutive rows)
                }
                v[j] = __float_as_uint(x);
                if (em.on && ok && col > row) { if (x != x) ++n_nan; else if (fabsf(x) >= em.r_lo) hitmask |= 1u << j; }
            }
            if (em.on) emit_chunk(em, v, hitmask, row, (i64)bj * BN + c0, lane);
        }
        if (em.on) { n_nan = __reduce_add_sync(0xffffffffu, n_nan); if (lane == 0 && n_nan) atomicAdd(&em.counters[1], (u64)n_nan); }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_ALLOC) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static cudaError_t encode_map(CUtensorMap* tm, void* ptr, i64 kp, i64 p_pad, std::string* msg) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
        if (e != cudaSuccess || !f) { *msg = "cuTensorMapEncodeTiled entry point not found"; return e != cudaSuccess ? e : cudaErrorUnknown; }
        fn = (EncodeTiledFn)f;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)kp, (cuuint64_t)p_pad};
    cuuint64_t gstr[1] = {(cuuint64_t)kp * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { *msg = "cuTensorMapEncodeTiled failed (CUresult " + std::to_string((int)r) + ")"; return cudaErrorInvalidValue; }
    return cudaSuccess;
}

struct Scratch {
    __nv_bfloat16* z = nullptr; size_t elems = 0;     // [2][p_pad][kp]
    cudaError_t reserve(size_t n) {
        if (n <= elems && z) return cudaSuccess;
        if (z) cudaFree(z);
        z = nullptr; elems = 0;
        cudaError_t e = cudaMalloc((void**)&z, n * sizeof(__nv_bfloat16));
        if (e == cudaSuccess) elems = n;
        return e;
    }
    ~Scratch() { if (z) cudaFree(z); }
};

// ---- cluster variant: two CTAs (same tile row bi, adjacent tile columns bj, bj+1) share the A operand -------------------------
// CTA c of the pair loads one half of A (c = 0: A_hi, c = 1: A_lo) and TMA-multicasts it into both CTAs' shared memory, so each
// SM pulls 48 KB instead of 64 KB per k-block through L2 (the kernel is L2-feed bound: profiles/).  Stage release is
// cluster-wide: every MMA thread's tcgen05.commit multicast-arrives on the `empty` barrier of both CTAs (count 2), because a
// producer's multicast writes into its peer's stage buffer too.  Everything else is cor_tc_kernel.
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                 ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
cor_tc2_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
               float* __restrict__ C, i64 p, int num_kb, int nb, int bi0, int bi1, int mirror, int bjlo, int bjhi, const PwEmit em, const int sh_world, const int sh_h) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - raw);
    const uint32_t bar0 = base + STAGES * STAGE_BYTES;
    const uint32_t bar_full = bar0, bar_empty = bar0 + 8 * STAGES, bar_cfull = bar0 + 16 * STAGES, bar_cempty = bar0 + 16 * STAGES + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + STAGES * STAGE_BYTES + 16 * STAGES + 32);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t crank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));

    // Tile order (L2 rasterisation).  The standardised table (2 GB at C4) is far larger than L2, and every tile streams its 128
    // A rows and 128 B rows over the whole K range; CTAs that run at the same time advance through K roughly in lock step, so an
    // operand row block is fetched from DRAM once per GROUP of co-resident tiles that share it.  Tile rows are therefore taken in
    // groups of COR_GROUP; inside a group the column pairs are the outer loop and the rows the inner one: the ~74 clusters of a
    // wave cover COR_GROUP rows x ~9 column pairs (COR_GROUP + 18 row blocks of operands instead of 1 + 148 in row-major order).
    // cluster q -> (bi, column pair); a pair covers columns (bj, bj + 1) of the rectangle rows [r0, r0 + rows) x columns
    // [max(r0, bjlo), bjhi); CTAs whose tile lies below the diagonal or beyond bjhi still feed their half of A to the peer,
    // compute a redundant tile and write nothing (row-range mode: bjlo = 0, bjhi = nb; column-band mode of the
    // upload-overlapped path: bi0 = 0, bi1 = bjhi).
    int bi = bi0, bj;
    bool live = true;
    {
        long long q = blockIdx.x >> 1;
        int r0 = bi0, rows = 1, lo = 0;
        for (;; r0 += COR_GROUP) {
            rows = bi1 - r0 < COR_GROUP ? bi1 - r0 : COR_GROUP;
            lo = r0 > bjlo ? r0 : bjlo;
            const long long cnt = (long long)((bjhi - lo + 1) >> 1) * rows;
            if (q < cnt || r0 + COR_GROUP >= bi1) break;
            q -= cnt;
        }
        bi = r0 + (int)(q % rows);
        bj = lo + 2 * (int)(q / rows) + (int)crank;
        if (bj >= bjhi) { bj = bjhi - 1; live = false; }
        if (bj < bi) live = false;
    }
    const int n_chunks = (num_kb + CHUNK - 1) / CHUNK;
    // row-sharded mode (several GPUs): tile row bi is stored at its local position in this rank's shard (common.cuh, CorView)
    float* Cw = C;
    if (sh_world > 1) { const int g = bi / sh_h; Cw = C + ((i64)((g < sh_world ? 0 : sh_h) + (bi - g * sh_h)) * 128 - (i64)bi * 128) * p; }

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 2); }   // both CTAs release a stage
        for (int b = 0; b < 2; ++b) { mbar_init(bar_cfull + 8 * b, 1); mbar_init(bar_cempty + 8 * b, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_lo) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_ALLOC) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                                 // the peer's barriers exist before any multicast can arrive on them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (uint32_t)((kb / STAGES) & 1);
                mbar_wait(bar_empty + 8 * s, ph ^ 1u);                      // both CTAs have drained this stage
                const uint32_t full = bar_full + 8 * s;
                mbar_expect_tx(full, STAGE_BYTES);                            // A_hi + A_lo (one of them from the peer) + B_hi + B_lo
                const uint32_t st = base + s * STAGE_BYTES;
                if (crank == 0) tma_load_2d_mc(st, &tm_hi, full, kb * BK, bi * BM, (uint16_t)3);
                else tma_load_2d_mc(st + TILE_BYTES, &tm_lo, full, kb * BK, bi * BM, (uint16_t)3);
                tma_load_2d(st + 2 * TILE_BYTES, &tm_hi, full, kb * BK, bj * BN);
                tma_load_2d(st + 3 * TILE_BYTES, &tm_lo, full, kb * BK, bj * BN);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            const uint32_t t_cross = tmem_base + 256;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (uint32_t)((kb / STAGES) & 1);
                const int c = kb / CHUNK, b = c & 1, u = c >> 1;
                const bool first = (kb % CHUNK) == 0;
                if (first && c >= 2) {
                    mbar_wait(bar_cempty + 8 * b, (uint32_t)((u - 1) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                mbar_wait(bar_full + 8 * s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t t_main = tmem_base + 128u * (uint32_t)b;
                const uint32_t st = base + s * STAGE_BYTES;
                const uint64_t a_hi = make_desc(st), a_lo = make_desc(st + TILE_BYTES);
                const uint64_t b_hi = make_desc(st + 2 * TILE_BYTES), b_lo = make_desc(st + 3 * TILE_BYTES);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);
                    umma_bf16(t_main, a_hi + adv, b_hi + adv, idesc, (first && k == 0) ? 0u : 1u);
                    umma_bf16(t_cross, a_hi + adv, b_lo + adv, idesc, (kb == 0 && k == 0) ? 0u : 1u);
                    umma_bf16(t_cross, a_lo + adv, b_hi + adv, idesc, 1u);
                }
                umma_commit_mc(bar_empty + 8 * s, (uint16_t)3);               // release the stage in both CTAs
                if ((kb % CHUNK) == CHUNK - 1 || kb == num_kb - 1) umma_commit(bar_cfull + 8 * b);
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        float acc[BN];
#pragma unroll
        for (int j = 0; j < BN; ++j) acc[j] = 0.0f;
        for (int c = 0; c < n_chunks; ++c) {
            const int b = c & 1, u = c >> 1;
            mbar_wait(bar_cfull + 8 * b, (uint32_t)(u & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + lane_base + 128u * (uint32_t)b + (uint32_t)c0, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(v[j]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_cempty + 8 * b);
        }
        const i64 row = (i64)bi * BM + q * 32 + lane;
        const bool diag = (bi == bj);
        unsigned int n_nan = 0;
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + lane_base + 256u + (uint32_t)c0, v);
            unsigned int hitmask = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const i64 col = (i64)bj * BN + c0 + j;
                float x = acc[c0 + j] + __uint_as_float(v[j]);
                x = x > 1.0f ? 1.0f : (x < -1.0f ? -1.0f : x);
                if (row == col) x = 1.0f;
                const bool ok = live && row < p && col < p && (!diag || col >= row);
                if (ok) {
                    Cw[row * p + col] = x;
                    if (mirror || diag) Cw[col * p + row] = x;
                }
                v[j] = __float_as_uint(x);
                if (em.on && ok && col > row) { if (x != x) ++n_nan; else if (fabsf(x) >= em.r_lo) hitmask |= 1u << j; }
            }
            if (em.on) emit_chunk(em, v, hitmask, row, (i64)bj * BN + c0, lane);
        }
        if (em.on) { n_nan = __reduce_add_sync(0xffffffffu, n_nan); if (lane == 0 && n_nan) atomicAdd(&em.counters[1], (u64)n_nan); }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                                 // no CTA leaves while its peer can still multicast into it / arrive on its barriers
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_ALLOC) : "memory");
    }
}

struct Prepared { CUtensorMap tm_hi, tm_lo; i64 kp = 0, p_pad = 0; int nb = 0; bool valid = false; };

// standardise + split the resident table (once per table) and build the TMA descriptors
static cudaError_t prepare(Scratch& S, Prepared& P, const float* d_data, i64 n, i64 p, i64 ld, cudaStream_t st, int* n_launch, std::string* msg) {
    const i64 kp = (n + BK - 1) / BK * BK;
    const i64 p_pad = (p + BM - 1) / BM * BM;
    cudaError_t e = S.reserve((size_t)2 * p_pad * kp);
    if (e != cudaSuccess) { *msg = "scratch allocation"; return e; }
    __nv_bfloat16* zhi = S.z;
    __nv_bfloat16* zlo = S.z + (size_t)p_pad * kp;
    standardize_split_kernel<256><<<(unsigned)p_pad, 256, 0, st>>>(d_data, n, ld, p, kp, zhi, zlo, 0);
    (*n_launch)++;
    e = cudaGetLastError(); if (e != cudaSuccess) { *msg = "standardize_split_kernel"; return e; }
    e = encode_map(&P.tm_hi, zhi, kp, p_pad, msg); if (e != cudaSuccess) return e;
    e = encode_map(&P.tm_lo, zlo, kp, p_pad, msg); if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(cor_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) { *msg = "cudaFuncSetAttribute(cor_tc_kernel)"; return e; }
    P.kp = kp; P.p_pad = p_pad; P.nb = (int)(p_pad / BM); P.valid = true;
    return cudaSuccess;
}

// upper-triangular tiles of tile rows [bi0, bi1).  FWGPU_COR_CLUSTER: 0 = cor_tc_kernel (one CTA per 128 x 128 tile), 1 = cor_tc2_kernel
// (CTA pair sharing A by multicast), 2 = cor_tc3_kernel (cta_group::2, 256 x 256 super-tile per CTA pair; cor_tc3.cuh, the default)
static int cluster_kernel_mode() {
    static const int v = [] { const char* e = getenv("FWGPU_COR_CLUSTER"); return e ? atoi(e) : 2; }();
    return v;
}
static bool use_cluster_kernel() { return cluster_kernel_mode() != 0; }
static cudaError_t launch_tc3(const Prepared& P, float* d_cor, i64 p, int bi0, int bi1, bool mirror, int bjlo, int bjhi, cudaStream_t st, int* n_launch,
                              std::string* msg, const PwEmit& em, int sh_world, int sh_h);
// clusters of cor_tc2_kernel for tile rows [bi0, bi1) x columns [bjlo, bjhi): see the tile-order comment in the kernel
static long long grouped_clusters(int bi0, int bi1, int bjlo, int bjhi) {
    long long c = 0;
    for (int r0 = bi0; r0 < bi1; r0 += COR_GROUP) {
        const int rows = bi1 - r0 < COR_GROUP ? bi1 - r0 : COR_GROUP;
        const int lo = r0 > bjlo ? r0 : bjlo;
        c += (long long)((bjhi - lo + 1) >> 1) * rows;
    }
    return c;
}
static cudaError_t run_rows(const Prepared& P, float* d_cor, i64 p, int bi0, int bi1, bool mirror, cudaStream_t st, int* n_launch, std::string* msg,
                            const PwEmit& em = PwEmit{nullptr, nullptr, 0, 2.0f, 0}, int sh_world = 1, int sh_h = 1) {
    if (bi1 > P.nb) bi1 = P.nb;
    if (bi0 >= bi1) return cudaSuccess;
    if (cluster_kernel_mode() == 2) return launch_tc3(P, d_cor, p, bi0, bi1, mirror, 0, P.nb, st, n_launch, msg, em, sh_world, sh_h);
    if (use_cluster_kernel()) {
        const long long clusters = grouped_clusters(bi0, bi1, 0, P.nb);
        cudaError_t e = cudaFuncSetAttribute(cor_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) { *msg = "cudaFuncSetAttribute(cor_tc2_kernel)"; return e; }
        cor_tc2_kernel<<<(unsigned)(2 * clusters), NTHREADS, SMEM_BYTES, st>>>(P.tm_hi, P.tm_lo, d_cor, p, (int)(P.kp / BK), P.nb, bi0, bi1, mirror ? 1 : 0, 0, P.nb, em, sh_world, sh_h);
        (*n_launch)++;
        e = cudaGetLastError(); if (e != cudaSuccess) { *msg = "cor_tc2_kernel"; return e; }
        return cudaSuccess;
    }
    const long long first = (long long)bi0 * P.nb - (long long)bi0 * (bi0 - 1) / 2;
    const long long last = (long long)bi1 * P.nb - (long long)bi1 * (bi1 - 1) / 2;
    cor_tc_kernel<<<(unsigned)(last - first), NTHREADS, SMEM_BYTES, st>>>(P.tm_hi, P.tm_lo, d_cor, p, (int)(P.kp / BK), P.nb, bi0, mirror ? 1 : 0, em, sh_world, sh_h);
    (*n_launch)++;
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) { *msg = "cor_tc_kernel"; return e; }
    return cudaSuccess;
}

static cudaError_t run(Scratch& S, const float* d_data, i64 n, i64 p, i64 ld, float* d_cor, cudaStream_t st, int* n_launch, std::string* msg,
                       const PwEmit& em = PwEmit{nullptr, nullptr, 0, 2.0f, 0}) {
    Prepared P;
    cudaError_t e = prepare(S, P, d_data, n, p, ld, st, n_launch, msg);
    if (e != cudaSuccess) return e;
    return run_rows(P, d_cor, p, 0, P.nb, true, st, n_launch, msg, em);
}

// Upload-overlapped cor_mat (one GPU): the host table is copied in column chunks on `copy_st`; as soon as chunk c has landed the
// compute stream standardises its columns and runs the tiles (bi <= bj, bj in the chunk's tile columns) — a growing column band
// of the upper triangle — so the PCIe transfer hides behind the GEMM.  Same tiles, same arithmetic as run().
static cudaError_t run_overlapped(Scratch& S, const float* host, float* d_data, i64 n, i64 p, i64 host_ld, float* d_cor,
                                  cudaStream_t st, cudaStream_t copy_st, cudaEvent_t* evs, int n_evs, int* n_launch, std::string* msg,
                                  const PwEmit& em = PwEmit{nullptr, nullptr, 0, 2.0f, 0}) {
    const i64 kp = (n + BK - 1) / BK * BK;
    const i64 p_pad = (p + BM - 1) / BM * BM;
    const int nb = (int)(p_pad / BM);
    cudaError_t e = S.reserve((size_t)2 * p_pad * kp);
    if (e != cudaSuccess) { *msg = "scratch allocation"; return e; }
    __nv_bfloat16* zhi = S.z;
    __nv_bfloat16* zlo = S.z + (size_t)p_pad * kp;
    Prepared P;
    e = encode_map(&P.tm_hi, zhi, kp, p_pad, msg); if (e != cudaSuccess) return e;
    e = encode_map(&P.tm_lo, zlo, kp, p_pad, msg); if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(cor_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) { *msg = "cudaFuncSetAttribute(cor_tc2_kernel)"; return e; }
    int chunks = n_evs < 12 ? n_evs : 12;
    if (chunks > nb) chunks = nb;
    if (chunks < 1) chunks = 1;
    // the copy stream must not overtake work that still reads d_data / Z from a previous call
    e = cudaEventRecord(evs[0], st); if (e != cudaSuccess) { *msg = "event"; return e; }
    e = cudaStreamWaitEvent(copy_st, evs[0], 0); if (e != cudaSuccess) { *msg = "wait"; return e; }
    int t0 = 0;
    for (int c = 0; c < chunks; ++c) {
        const int t1 = (int)((long long)nb * (c + 1) / chunks);          // tile columns [t0, t1)
        const i64 c0 = (i64)t0 * BM, c1 = std::min<i64>((i64)t1 * BM, p);
        if (c1 > c0) {
            e = cudaMemcpy2DAsync(d_data + c0 * n, n * sizeof(float), host + c0 * host_ld, host_ld * sizeof(float), n * sizeof(float), (size_t)(c1 - c0),
                                  cudaMemcpyHostToDevice, copy_st);
            if (e != cudaSuccess) { *msg = "cudaMemcpy2DAsync"; return e; }
        }
        e = cudaEventRecord(evs[c], copy_st); if (e != cudaSuccess) { *msg = "event"; return e; }
        e = cudaStreamWaitEvent(st, evs[c], 0); if (e != cudaSuccess) { *msg = "wait"; return e; }
        const i64 s1 = (c == chunks - 1) ? p_pad : (i64)t1 * BM;          // the last chunk also zero-fills the padding rows
        standardize_split_kernel<256><<<(unsigned)(s1 - c0), 256, 0, st>>>(d_data, n, n, p, kp, zhi, zlo, c0);
        (*n_launch)++;
        if (cluster_kernel_mode() == 2) {
            P.kp = kp; P.p_pad = p_pad; P.nb = nb;
            e = launch_tc3(P, d_cor, p, 0, t1, true, t0, t1, st, n_launch, msg, em, 1, 1); if (e != cudaSuccess) return e;
        } else {
            const long long clusters = grouped_clusters(0, t1, t0, t1);
            cor_tc2_kernel<<<(unsigned)(2 * clusters), NTHREADS, SMEM_BYTES, st>>>(P.tm_hi, P.tm_lo, d_cor, p, (int)(kp / BK), nb, 0, t1, 1, t0, t1, em, 1, 1);
            (*n_launch)++;
            e = cudaGetLastError(); if (e != cudaSuccess) { *msg = "cor_tc2_kernel (band)"; return e; }
        }
        t0 = t1;
    }
    return cudaSuccess;
}

// lower triangle <- upper triangle (after the row blocks of all ranks are in place): 32x32 tiles through shared memory
__global__ void __launch_bounds__(256) cor_symmetrize_kernel(float* __restrict__ C, i64 p) {
    __shared__ float tile[32][33];
    const int bx = blockIdx.x, by = blockIdx.y;         // source tile (rows by, cols bx) with bx >= by
    if (bx < by) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const i64 row = (i64)by * 32 + r, col = (i64)bx * 32 + tx;
        tile[r][tx] = (row < p && col < p) ? C[row * p + col] : 0.0f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const i64 row = (i64)bx * 32 + r, col = (i64)by * 32 + tx;      // destination (transposed position)
        if (row < p && col < p && row > col) C[row * p + col] = tile[tx][r];
    }
}

}  // namespace cortc
