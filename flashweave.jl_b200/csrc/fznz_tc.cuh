// fznz_tc.cuh — tensor-core pre-filter of the pairwise fz_nz stage (tcgen05 / TMEM / TMA, sm_100a).
//
// Replaces the all-pairs loop of pw_univar_kernel for test_name "fz_nz" (src/tests.jl:410-433 calling the dense univariate
// test :108-160): for every pair (X, Y) the correlation is taken on the rows where both are non-zero.  With the 0/1 non-zero
// indicator m and the table x (exactly 0 where m = 0) every moment of that view is a dot product over ALL rows:
//     n   = sum m_x m_y      Sx  = sum x m_y      Sxx = sum x^2 m_y      Sxy = sum x y
//                            Sy  = sum m_x y      Syy = sum m_x y^2
// i.e. six p x p x n contractions of the three bf16 operand planes {M, X, X2} - tensor-core work.  bf16 operands give r to
// ~1e-4, which is not the parity target, so this kernel only CLASSIFIES.  From the worst-case rounding errors of the operands
// (bf16: |dx'| <= 2^-9 |x'|, the x'^2 plane 3 * 2^-9, a product x'y' 2 * 2^-9; Cauchy-Schwarz turns sums of |x'| into
// sqrt(n Sxx)) and of the truncating fp32 accumulation it derives an interval for each view variance and for the covariance,
//     vx in [vx~ - ex, .],  ex = (E2 + T) qx + 2 |mx| (E1 + T) sqrt(qx),   qx = Sxx / n
//     |cov| <= |cov~| + ec, ec = (E3 + T) sqrt(qx qy) + (E1 + T) (|mx| sqrt(qy) + |my| sqrt(qx)),
// and a pair is dismissed only if the resulting UPPER bound of |r| is below the significance threshold; a pair whose variance
// interval reaches zero (r may be NaN or arbitrarily ill-conditioned: few shared rows far from the column mean, constant
// columns) always goes to the exact fp64 warp-per-pair test (fznz_uni_warp), as does every pair above the threshold.  Counts (n, exact:
// sums of 0/1 products in fp32) decide the "too few rows" cases exactly as the reference does.  The resulting neighbour lists are
// identical to the exhaustive exact kernel (tests/test_gpu_fznz.py compares both), at ~3 % of its fp64 work.
//
// One CTA = one 128 (X block) x 64 (Y block) tile of the upper triangle, six fp32 accumulators of 64 TMEM columns, a 3-stage
// TMA -> mbarrier -> tcgen05.mma pipeline (72 KB / stage: 3 A planes of 128 rows, 3 B planes of 64 rows, K-major, SWIZZLE_128B);
// warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue (tcgen05.ld, classification, warp-aggregated appends).
// The planes hold x' = (x - mean_nz) / sd_nz on the non-zero rows (correlations are shift/scale invariant; centring keeps
// the raw-moment cancellation benign) and x'^2, zero elsewhere.
#pragma once
#include "cor_tc.cuh"
#include "fznz.cuh"

namespace fznztc {

using cortc::smem_u32; using cortc::mbar_init; using cortc::mbar_expect_tx; using cortc::mbar_wait; using cortc::tma_load_2d;
using cortc::make_desc; using cortc::umma_bf16; using cortc::umma_commit;

constexpr int BM = 128, BN = 64, BK = 64, STAGES = 3;
constexpr int A_TILE = BM * BK * 2;                    // 16 KB
constexpr int B_TILE = BN * BK * 2;                    // 8 KB
constexpr int STAGE_BYTES = 3 * A_TILE + 3 * B_TILE;   // 72 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
constexpr int NTHREADS = 192;
constexpr int TC_GROUP = 8;                            // X blocks per rasterisation group

// one CTA per variable: mean / sd over the non-zero rows (fixed-order reductions: deterministic), then the three bf16 planes
template <int THREADS>
__global__ void __launch_bounds__(THREADS) nz_planes_kernel(const float* __restrict__ data, i64 n, i64 ld, i64 p, i64 kp,
                                                            __nv_bfloat16* __restrict__ pm, __nv_bfloat16* __restrict__ px, __nv_bfloat16* __restrict__ px2) {
    const i64 col = blockIdx.x;
    __nv_bfloat16* m = pm + col * kp; __nv_bfloat16* x1 = px + col * kp; __nv_bfloat16* x2 = px2 + col * kp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
    if (col >= p) {
        for (i64 i = tid; i < kp; i += THREADS) { m[i] = zero; x1[i] = zero; x2[i] = zero; }
        return;
    }
    const float* x = data + col * ld;
    __shared__ double red[THREADS / 32]; __shared__ double red2[THREADS / 32];
    __shared__ double s_mean, s_inv;
    double s = 0.0, c = 0.0;
    for (i64 i = tid; i < n; i += THREADS) { const float v = x[i]; if (v != 0.0f) { s += (double)v; c += 1.0; } }
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_down_sync(0xffffffffu, s, o); c += __shfl_down_sync(0xffffffffu, c, o); }
    if (lane == 0) { red[warp] = s; red2[warp] = c; }
    __syncthreads();
    if (tid == 0) { double t = 0.0, k = 0.0; for (int w = 0; w < THREADS / 32; ++w) { t += red[w]; k += red2[w]; } s_mean = k > 0.0 ? t / k : 0.0; red2[0] = k; }
    __syncthreads();
    const double mean = s_mean, cnt = red2[0];
    __syncthreads();
    double ss = 0.0;
    for (i64 i = tid; i < n; i += THREADS) { const float v = x[i]; if (v != 0.0f) { const double d = (double)v - mean; ss += d * d; } }
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_down_sync(0xffffffffu, ss, o);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    if (tid == 0) { double t = 0.0; for (int w = 0; w < THREADS / 32; ++w) t += red[w]; s_inv = (t > 0.0 && cnt > 1.0) ? 1.0 / sqrt(t / cnt) : 1.0; }
    __syncthreads();
    const double inv = s_inv;
    const __nv_bfloat16 one = __float2bfloat16_rn(1.f);
    for (i64 i = tid; i < kp; i += THREADS) {
        const float v = i < n ? x[i] : 0.0f;
        if (v != 0.0f) {
            const __nv_bfloat16 h = __float2bfloat16_rn((float)(((double)v - mean) * inv));
            const float hf = __bfloat162float(h);
            m[i] = one; x1[i] = h; x2[i] = __float2bfloat16_rn(hf * hf);
        } else { m[i] = zero; x1[i] = zero; x2[i] = zero; }
    }
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}

struct PrefilterArgs {
    i64 p; int num_kb; int nb_a, nb_b;          // tile grid: nb_a blocks of 128 X, nb_b blocks of 64 Y
    const int* nnz;                             // per-variable non-zero count (rows of the X-trimmed view, tests.jl:412-416)
    i64 n_obs_min; int reliable_only;
    float z_alpha;                              // two-sided normal quantile of alpha: p < alpha  <=>  sqrt(n-3) * atanh|r| > z_alpha
    float trunc_rel;                            // T: bound of the accumulated truncation of the fp32 TMEM accumulators, relative to sum |terms|
    u64* counters;                              // [0] candidates appended, [1] pairs counted reliable here (non-candidates)
    i64 cand_cap; int* cand_x; int* cand_y;
    int sh_rank, sh_world;                      // X groups of this rank (pw_owns_group, common.cuh)
};

__global__ void __launch_bounds__(NTHREADS, 1) fznz_prefilter_kernel(const __grid_constant__ CUtensorMap tm_m, const __grid_constant__ CUtensorMap tm_x,
                                                                     const __grid_constant__ CUtensorMap tm_x2, PrefilterArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - raw);
    const uint32_t bar0 = base + STAGES * STAGE_BYTES;
    const uint32_t bar_full = bar0, bar_empty = bar0 + 8 * STAGES, bar_done = bar0 + 16 * STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + STAGES * STAGE_BYTES + 16 * STAGES + 16);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // tile (I, J): X block I (128 variables) against Y block J (64 variables), J >= 2 I (some Y above some X).  Tile order as in
    // cor_tc2_kernel: X blocks in groups of TC_GROUP, inside a group the Y blocks are the outer loop, so the tiles that run at
    // the same time share TC_GROUP X row blocks and ~18 Y row blocks of the operand planes (3 GB at C5, far larger than L2)
    // instead of streaming a different Y block each; tiles of the group's rectangle below the diagonal exit at once.
    int I = 0, J = 0;
    {
        long long q = blockIdx.x;
        int r0 = 0, rows = 1;
        for (;; r0 += TC_GROUP) {
            rows = a.nb_a - r0 < TC_GROUP ? a.nb_a - r0 : TC_GROUP;
            const long long cnt = pw_owns_group(r0 / TC_GROUP, a.sh_rank, a.sh_world) ? (long long)(a.nb_b - 2 * r0) * rows : 0;
            if (q < cnt) break;
            if (r0 + TC_GROUP >= a.nb_a) return;
            q -= cnt;
        }
        I = r0 + (int)(q % rows);
        J = 2 * r0 + (int)(q / rows);
        if (J < 2 * I || J >= a.nb_b) return;
    }

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_m) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_x2) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer: boxes of 64 rows x 64 elements; an A plane is two boxes stacked (same canonical layout) =====
            for (int kb = 0; kb < a.num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (uint32_t)((kb / STAGES) & 1);
                mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                const uint32_t full = bar_full + 8 * s;
                mbar_expect_tx(full, STAGE_BYTES);
                const uint32_t st = base + s * STAGE_BYTES;
                const CUtensorMap* maps[3] = {&tm_m, &tm_x, &tm_x2};
#pragma unroll
                for (int pl = 0; pl < 3; ++pl) {
                    tma_load_2d(st + pl * A_TILE, maps[pl], full, kb * BK, I * BM);
                    tma_load_2d(st + pl * A_TILE + B_TILE, maps[pl], full, kb * BK, I * BM + 64);
                    tma_load_2d(st + 3 * A_TILE + pl * B_TILE, maps[pl], full, kb * BK, J * BN);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer: M = 128, N = 64, K = 16, bf16 x bf16 -> fp32 =====
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            for (int kb = 0; kb < a.num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (uint32_t)((kb / STAGES) & 1);
                mbar_wait(bar_full + 8 * s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = base + s * STAGE_BYTES;
                const uint64_t a_m = make_desc(st), a_x = make_desc(st + A_TILE), a_x2 = make_desc(st + 2 * A_TILE);
                const uint64_t b_m = make_desc(st + 3 * A_TILE), b_x = make_desc(st + 3 * A_TILE + B_TILE), b_x2 = make_desc(st + 3 * A_TILE + 2 * B_TILE);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);
                    const uint32_t acc = (kb == 0 && k == 0) ? 0u : 1u;
                    umma_bf16(tmem_base + 0, a_m + adv, b_m + adv, idesc, acc);          // n
                    umma_bf16(tmem_base + 64, a_x + adv, b_m + adv, idesc, acc);         // Sx  = sum x m_y
                    umma_bf16(tmem_base + 128, a_m + adv, b_x + adv, idesc, acc);        // Sy  = sum m_x y
                    umma_bf16(tmem_base + 192, a_x2 + adv, b_m + adv, idesc, acc);       // Sxx = sum x^2 m_y
                    umma_bf16(tmem_base + 256, a_m + adv, b_x2 + adv, idesc, acc);       // Syy = sum m_x y^2
                    umma_bf16(tmem_base + 320, a_x + adv, b_x + adv, idesc, acc);        // Sxy = sum x y
                }
                umma_commit(bar_empty + 8 * s);
            }
            umma_commit(bar_done);
        }
        __syncwarp();
    } else {
        // ===== epilogue: classify the 128 x 64 pairs of the tile =====
        const int q = warp & 3;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        mbar_wait(bar_done, 0u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const i64 X = (i64)I * BM + q * 32 + lane;
        const bool x_ok = X < a.p;
        const bool x_short = x_ok && (i64)a.nnz[x_ok ? X : 0] < a.n_obs_min;              // tests.jl:111-115: too few rows in the X view
        unsigned int n_rel = 0;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
            uint32_t v[6][16];
#pragma unroll
            for (int m = 0; m < 6; ++m) tmem_ld16(tmem_base + lane_base + (uint32_t)(64 * m + c0), v[m]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const i64 Y = (i64)J * BN + c0 + j;
                bool cand = false;
                if (x_ok && Y < a.p && Y > X) {
                    const float n = __uint_as_float(v[0][j]);
                    if (x_short) n_rel += (0 >= a.n_obs_min || !a.reliable_only) ? 1u : 0u;
                    else if (!(n > 0.0f && (i64)n >= a.n_obs_min)) n_rel += ((i64)n >= a.n_obs_min || !a.reliable_only) ? 1u : 0u;
                    else {
                        const float inv = 1.0f / n;
                        const float mx = __uint_as_float(v[1][j]) * inv, my = __uint_as_float(v[2][j]) * inv;
                        const float vx = __uint_as_float(v[3][j]) * inv - mx * mx, vy = __uint_as_float(v[4][j]) * inv - my * my;
                        const float cov = __uint_as_float(v[5][j]) * inv - mx * my;
                        // error intervals (see the header); 1.02: the bounds use the computed moments in place of the true ones
                        const float E1 = 1.02f / 512.0f + a.trunc_rel, E2 = 1.02f / 128.0f + a.trunc_rel, E3 = 1.02f / 256.0f + a.trunc_rel;
                        const float qx = __uint_as_float(v[3][j]) * inv, qy = __uint_as_float(v[4][j]) * inv;
                        const float sqx = sqrtf(qx), sqy = sqrtf(qy);
                        const float lo_vx = vx - (E2 * qx + 2.0f * fabsf(mx) * E1 * sqx);
                        const float lo_vy = vy - (E2 * qy + 2.0f * fabsf(my) * E1 * sqy);
                        const float hi_cov = fabsf(cov) + E3 * sqx * sqy + E1 * (fabsf(mx) * sqy + fabsf(my) * sqx);
                        const bool var_ok = lo_vx > 0.0f && lo_vy > 0.0f;
                        const float r_hi = hi_cov * rsqrtf(fmaxf(lo_vx * lo_vy, 1e-30f)) * 1.0001f;   // fp32 evaluation of these formulas
                        const float thr = n > 3.5f ? tanhf(a.z_alpha * rsqrtf(n - 3.0f)) : 2.0f;   // n - 3 <= 0: the p-value is 1
                        cand = !var_ok || !(r_hi < thr * 0.999f);                                  // NaN-safe: anything odd is a candidate
                        if (!cand) n_rel += 1u;
                    }
                }
                const unsigned int bal = __ballot_sync(0xffffffffu, cand);
                if (bal) {
                    u64 pos0 = 0;
                    if (lane == 0) pos0 = atomicAdd(&a.counters[0], (u64)__popc(bal));
                    pos0 = __shfl_sync(0xffffffffu, pos0, 0);
                    if (cand) {
                        const u64 pos = pos0 + (u64)__popc(bal & ((1u << lane) - 1u));
                        if ((i64)pos < a.cand_cap) { a.cand_x[pos] = (int)X; a.cand_y[pos] = (int)Y; }
                    }
                }
            }
        }
        n_rel = __reduce_add_sync(0xffffffffu, n_rel);
        if (lane == 0 && n_rel) atomicAdd(&a.counters[1], (u64)n_rel);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// Candidate order.  The exact test streams both columns of a pair through L1 (2 x 4n bytes, an L2 round trip per sector); the 8 warps of
// a CTA take 8 consecutive candidates, so the list is sorted by (X tile, Y tile, X, Y): the warps of a CTA then share X (its sectors hit
// L1 after the first warp) and the CTAs that run together stay inside a few pre-filter tiles (L2).  key = tileX:20 | tileY:20 | x&127:7 | y&63:6.
__global__ void fznz_cand_keys_kernel(const int* __restrict__ cand_x, const int* __restrict__ cand_y, i64 n, u64* __restrict__ keys) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 x = (u64)cand_x[i], y = (u64)cand_y[i];
    keys[i] = ((x >> 7) << 33) | ((y >> 6) << 13) | ((x & 127u) << 6) | (y & 63u);
}

// exact fp64 test of the candidate pairs, one warp each; same outputs as pw_fznz_rows_kernel
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) fznz_candidates_kernel(NzTable t, i64 n_cand, const u64* __restrict__ keys,
                                                                     i64 n_obs_min, double alpha, int reliable_only,
                                                                     u64* counters, i64 cap, int* c_x, int* c_y, double* c_p, double* c_stat, int use_wm) {
    extern __shared__ unsigned int cand_wm[];                           // WARPS x 4 x W words (fznz_warp_scratch_bytes)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 3;
    unsigned int* wm = use_wm ? cand_wm + (size_t)(warp * 4 + g) * t.W : nullptr;
    i64 n_rel = 0;
    for (i64 i0 = ((i64)blockIdx.x * WARPS + warp) * 4; i0 < n_cand; i0 += (i64)gridDim.x * WARPS * 4) {   // four candidates per warp
        const i64 i = i0 + g;
        const bool valid = i < n_cand;
        const u64 k = keys[valid ? i : i0];
        const i64 X = (i64)(((k >> 33) << 7) | ((k >> 6) & 127u)), Y = (i64)((((k >> 13) & 0xFFFFFu) << 6) | (k & 63u));
        NzUni r = fznz_uni_g8(t, X, Y, n_obs_min, wm, valid);
        const bool rel = valid && (r.suff || !reliable_only) && !isnan(r.pval);
        if ((lane & 7) == 0) {
            n_rel += rel;
            if (rel && r.pval < alpha) {
                u64 pos = atomicAdd(&counters[2], 1ull);
                if ((i64)pos < cap) { c_x[pos] = (int)X; c_y[pos] = (int)Y; c_p[pos] = r.pval; c_stat[pos] = r.stat; }
            }
        }
        __syncwarp();
    }
    n_rel = __reduce_add_sync(0xffffffffu, (unsigned int)n_rel);
    if (lane == 0 && n_rel) atomicAdd(&counters[1], (u64)n_rel);
}

struct Planes {
    __nv_bfloat16* z = nullptr; size_t elems = 0;
    cudaError_t reserve(size_t n) {
        if (n <= elems && z) return cudaSuccess;
        if (z) cudaFree(z);
        z = nullptr; elems = 0;
        cudaError_t e = cudaMalloc((void**)&z, n * sizeof(__nv_bfloat16));
        if (e == cudaSuccess) elems = n;
        return e;
    }
    ~Planes() { if (z) cudaFree(z); }
};

static cudaError_t encode_map64(CUtensorMap* tm, void* ptr, i64 kp, i64 p_pad, std::string* msg) {
    static cortc::EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
        if (e != cudaSuccess || !f) { *msg = "cuTensorMapEncodeTiled entry point not found"; return e != cudaSuccess ? e : cudaErrorUnknown; }
        fn = (cortc::EncodeTiledFn)f;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)kp, (cuuint64_t)p_pad};
    cuuint64_t gstr[1] = {(cuuint64_t)kp * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, 64u};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { *msg = "cuTensorMapEncodeTiled failed (CUresult " + std::to_string((int)r) + ")"; return cudaErrorInvalidValue; }
    return cudaSuccess;
}

// two-sided standard-normal quantile: erfc(z / sqrt 2) = alpha (bisection, host)
static double z_of_alpha(double alpha) {
    double lo = 0.0, hi = 40.0;
    for (int it = 0; it < 200; ++it) { const double mid = 0.5 * (lo + hi); if (std::erfc(mid * 0.70710678118654752440) < alpha) hi = mid; else lo = mid; }
    return lo;                                                 // slightly below the exact quantile: the conservative side
}

// planes + pre-filter launch; on return counters[0] = #candidates (possibly > cand_cap: caller re-runs), counters[1] = #reliable counted
static cudaError_t run_prefilter(Planes& P, const NzTable& t, i64 n_obs_min, double alpha, bool reliable_only, u64* counters,
                                 i64 cand_cap, int* cand_x, int* cand_y, bool planes_ready, cudaStream_t st, int* n_launch, std::string* msg,
                                 int sh_rank = 0, int sh_world = 1) {
    const i64 p = t.p, n = t.n;
    const i64 kp = (n + BK - 1) / BK * BK;
    const i64 p_pad = (p + BM - 1) / BM * BM;
    cudaError_t e = P.reserve((size_t)3 * p_pad * kp);
    if (e != cudaSuccess) { *msg = "plane allocation"; return e; }
    __nv_bfloat16* pm = P.z; __nv_bfloat16* px = P.z + (size_t)p_pad * kp; __nv_bfloat16* px2 = P.z + (size_t)2 * p_pad * kp;
    if (!planes_ready) {
        nz_planes_kernel<256><<<(unsigned)p_pad, 256, 0, st>>>(t.data, n, t.ld, p, kp, pm, px, px2);
        (*n_launch)++;
        e = cudaGetLastError(); if (e != cudaSuccess) { *msg = "nz_planes_kernel"; return e; }
    }
    CUtensorMap tm_m, tm_x, tm_x2;
    e = encode_map64(&tm_m, pm, kp, p_pad, msg); if (e != cudaSuccess) return e;
    e = encode_map64(&tm_x, px, kp, p_pad, msg); if (e != cudaSuccess) return e;
    e = encode_map64(&tm_x2, px2, kp, p_pad, msg); if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(fznz_prefilter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) { *msg = "cudaFuncSetAttribute(fznz_prefilter_kernel)"; return e; }
    PrefilterArgs a;
    a.p = p; a.num_kb = (int)(kp / BK); a.nb_a = (int)(p_pad / BM); a.nb_b = (int)(p_pad / BN);
    a.nnz = t.nnz; a.n_obs_min = n_obs_min; a.reliable_only = reliable_only ? 1 : 0;
    a.z_alpha = (float)(z_of_alpha(alpha) * (1.0 - 1e-6));
    a.trunc_rel = (float)(2.0 * (double)(kp / 16) / 8388608.0);       // one truncation (< 2^-23 relative) per K = 16 accumulation, x2 margin
    a.counters = counters; a.cand_cap = cand_cap; a.cand_x = cand_x; a.cand_y = cand_y;
    a.sh_rank = sh_rank; a.sh_world = sh_world;
    static_assert(TC_GROUP * BM == PW_X_GROUP, "the X groups of the sharded pairwise stage are the rasterisation groups of the pre-filter");
    long long tiles = 0;
    for (int r0 = 0; r0 < a.nb_a; r0 += TC_GROUP)
        if (pw_owns_group(r0 / TC_GROUP, sh_rank, sh_world)) tiles += (long long)(a.nb_b - 2 * r0) * std::min(TC_GROUP, a.nb_a - r0);
    if (tiles == 0) return cudaSuccess;
    fznz_prefilter_kernel<<<(unsigned)tiles, NTHREADS, SMEM_BYTES, st>>>(tm_m, tm_x, tm_x2, a);
    (*n_launch)++;
    e = cudaGetLastError(); if (e != cudaSuccess) { *msg = "fznz_prefilter_kernel"; return e; }
    return cudaSuccess;
}

}  // namespace fznztc
