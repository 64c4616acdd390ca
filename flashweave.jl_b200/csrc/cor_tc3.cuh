// cor_tc3.cuh — cor_mat on a CTA PAIR: tcgen05.mma.cta_group::2, one 256 x 256 super-tile per 2-CTA cluster.
//
// Why: an SS-mode 128x128x16 MMA reads 8 KB of shared memory per 64 tensor-pipe cycles (the whole 128 B/clk port) while TMA fills
// the next stage through the same port; cor_tc2_kernel (cor_tc.cuh) therefore tops out at ~65 % tensor-pipe activity.  With
// cta_group::2 the pair computes D[256 x 256] = A[256 x K] B[256 x K]^T: each CTA holds its own 128 rows of A and HALF of B
// (128 of the 256 columns), the hardware feeds both SMs from the two halves, so the shared-memory reads per flop halve (A 4 KB +
// B-half 4 KB per 128 cycles) and the L2 -> SM fill drops from 48 KB to 32 KB per 128 x 128 x 64 block of MMAs.
//
// Accuracy scheme: z = hi + lo in bf16 and hi*hi + hi*lo + lo*hi as in cor_tc.cuh; because the tensor core truncates when it adds a
// K = 16 partial product into the fp32 accumulator, a TMEM accumulator only ever holds the partial sum of CHUNK3 k-blocks, which the
// epilogue warps drain into fp32 registers.  512 TMEM columns = two 256-column accumulators used alternately (see CHUNK3 below).
//
// Roles per CTA (320 threads): warp 0 = TMA producer (own A rows, own half of B; the transaction bytes of BOTH CTAs are counted on the
// leader's `full` barrier, cp.async.bulk.tensor ... .cta_group::2), warp 1 = TMEM allocation + (leader CTA only) the MMA-issuing
// thread, warps 2-9 = epilogue (lane quarter = warp % 4, column half = (warp - 2) / 4): chunk drains, clamp, unit diagonal, then the
// tile is staged in the idle pipeline buffers and both the direct and the mirrored tile leave in 512-byte runs from compact loops
// (mirror_block; the raw candidates of the univariate Fisher-z stage are collected there when armed).
#pragma once
#include "cor_tc.cuh"

namespace cortc {

#ifndef FW_COR3_SLEEP_NS
#define FW_COR3_SLEEP_NS 0
#endif
constexpr int NTHREADS3 = 320;
constexpr int COR_GROUP3 = 8;                              // super-rows (of 256 matrix rows) per rasterisation group
constexpr int STG_LD = 129;                                // staging row stride in floats (conflict-free both ways)

#ifdef FW_COR3_DEBUG
// experiment builds only (scripts/cor3_trace.py): per-cluster cycle stamps of the leader CTA, 8 x i64 per cluster
__device__ long long* g_cor3_dbg = nullptr;
#define COR3_DBG(slot, val) do { if (g_cor3_dbg) g_cor3_dbg[(size_t)(blockIdx.x >> 1) * 8 + (slot)] = (long long)(val); } while (0)
#else
#define COR3_DBG(slot, val) do { } while (0)
#endif

__device__ __forceinline__ uint32_t mapa_cta(uint32_t addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
    return r;
}
// TMA load of this CTA's operand tile; the bytes are counted on `bar_cluster` (a shared::cluster address: the leader CTA's barrier)
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const CUtensorMap* tm, uint32_t bar_cluster, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(tm), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_bf16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_cg2(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
// bounded spin (a protocol bug must trap, not hang the GPU): ~2^27 polls is seconds, far beyond any legitimate wait
// SLEEP_NS > 0: back off between polls (the epilogue warps and the producer wait for microseconds; the kernel is power-limited and
// eight warps polling at full issue rate are not free)
template <int SLEEP_NS = 0>
__device__ __forceinline__ void mbar_wait_b(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    unsigned int spins = 0;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok) { if (++spins > (1u << 27)) __trap(); if (SLEEP_NS > 0) __nanosleep(SLEEP_NS); }
    } while (!ok);
}
__device__ __forceinline__ void mbar_wait_cluster_b(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    unsigned int spins = 0;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 27)) __trap();
    } while (!ok);
}

// raw candidates of one 32-column chunk of the staged tile (emit_chunk of cor_tc.cuh with the values read back from shared memory)
__device__ __forceinline__ void emit_chunk_stg(const PwEmit& em, const float* stg_row, int c0, unsigned int hitmask, i64 row, i64 col0, int lane) {
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
    while (hitmask) {
        const int j = __ffs(hitmask) - 1; hitmask &= hitmask - 1;
        if ((i64)pos < em.cap) { PwRec rec; rec.x = (int)row; rec.y = (int)(col0 + c0 + j); rec.r = stg_row[c0 + j]; em.list[pos] = rec; }
        ++pos;
    }
}

// Mirrored tile + raw candidates of one 128 x 128 tile staged in S (row stride STG_LD): warp q writes the matrix rows col0 + q*32 .. +31
// (= tile columns), 128 consecutive entries each (= the tile rows, read down a column of S: conflict-free with the odd stride).
// DIAG: diagonal tile, only col >= row is written / tested.  EMIT: collect the raw candidates of the univariate stage for the warp's
// 128 x 32 block (lane = row within a 32-row group k, bit = column).  Everything row- or column-uniform is hoisted: the loop body is
// one shared-memory load and one predicated store per element.
template <bool DIAG, bool EMIT>
__device__ __forceinline__ void mirror_block(const float* S, float* Cw, i64 p, i64 trow0, i64 col0, int q, int lane, bool do_mirror, const PwEmit& em) {
    const int nrow = (int)((p - trow0) < 128 ? (p - trow0) : 128), ncol = (int)((p - col0) < 128 ? (p - col0) : 128);
    unsigned int hm[4] = {0u, 0u, 0u, 0u}, n_nan = 0;
    float* mp = Cw + (col0 + q * 32) * p + trow0 + lane;
    const float* sp = S + lane * STG_LD + q * 32;
    const int cend = ncol - q * 32 < 32 ? ncol - q * 32 : 32;
    for (int cc = 0; cc < cend; ++cc, mp += p, ++sp) {
        const int c = q * 32 + cc;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int r = k * 32 + lane;
            const float x = sp[k * 32 * STG_LD];
            const bool ok = r < nrow && (!DIAG || c >= r);
            if (ok && do_mirror) mp[k * 32] = x;
            if (EMIT) { if (ok && (!DIAG || c > r)) { if (x != x) ++n_nan; else if (fabsf(x) >= em.r_lo) hm[k] |= 1u << cc; } }
        }
    }
    if (EMIT) {
#pragma unroll
        for (int k = 0; k < 4; ++k) emit_chunk_stg(em, S + (k * 32 + lane) * STG_LD, q * 32, hm[k], trow0 + k * 32 + lane, col0, lane);
        n_nan = __reduce_add_sync(0xffffffffu, n_nan);
        if (lane == 0 && n_nan) atomicAdd(&em.counters[1], (u64)n_nan);
    }
}

// Accumulation: all three split terms of a chunk of CHUNK3 k-blocks (256 samples) go to ONE TMEM accumulator, the two accumulators
// (256 columns each = all 512) alternate by chunk, and the epilogue warps drain the finished one into fp32 registers (round-to-nearest
// adds) while the MMAs of the next chunk run.  The truncation bias of the tensor core's accumulate is proportional to
// (#accumulations per chunk) x (magnitude of the partial sum): 48 x |r| 256/n here against 64 x |r| 1024/n of cor_tc2_kernel's
// main accumulator, i.e. ~5x smaller, which is why the cross terms no longer need an accumulator of their own.
constexpr int CHUNK3 = 4;

// Tile order.  Tile rows [bi0, bi1) (units of 128 matrix rows) are taken two at a time = one super-row per cluster (CTA r of the pair
// owns tile row R0 + r); a cluster covers the tile columns (lo + 2j, lo + 2j + 1).  Super-rows are grouped COR_GROUP3 at a time, the
// column pairs are the outer loop inside a group and the super-rows the inner one (L2 rasterisation, see cor_tc2_kernel): a wave of
// 74 clusters then streams 8 + ~9 operand row blocks of 256 rows.  lo = max(first tile row of the group, bjlo); 128-tiles below the
// diagonal, beyond bjhi or beyond bi1 are dead (computed, not written); clusters that are dead as a whole exit at once.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS3, 1)
cor_tc3_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
               float* __restrict__ C, i64 p, int num_kb, int nb, int bi0, int bi1, int mirror, int bjlo, int bjhi, const PwEmit em, const int sh_world, const int sh_h) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - raw);
    // barriers: full[STAGES], empty[STAGES], cfull[2], cempty[2]; then the TMEM base-address slot
    const uint32_t bar0 = base + STAGES * STAGE_BYTES;
    const uint32_t bar_full = bar0, bar_empty = bar0 + 8 * STAGES, bar_cfull = bar0 + 16 * STAGES, bar_cempty = bar0 + 16 * STAGES + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + STAGES * STAGE_BYTES + 16 * STAGES + 32);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t crank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
#ifdef FW_COR3_DEBUG
    const long long t_entry = clock64();
#endif

    int R0, cp0;                                           // first tile row of the super-row, first tile column of the pair
    {
        long long q = blockIdx.x >> 1;
        int r0 = bi0, rows = 1, lo = 0;
        for (;; r0 += 2 * COR_GROUP3) {
            const int left = (bi1 - r0 + 1) >> 1;          // super-rows left from r0
            rows = left < COR_GROUP3 ? left : COR_GROUP3;
            lo = r0 > bjlo ? r0 : bjlo;
            const long long cnt = bjhi > lo ? (long long)((bjhi - lo + 1) >> 1) * rows : 0;
            if (q < cnt || r0 + 2 * COR_GROUP3 >= bi1) break;
            q -= cnt;
        }
        R0 = r0 + 2 * (int)(q % rows);
        cp0 = lo + 2 * (int)(q / rows);
    }
    if (cp0 + 1 < R0 || cp0 >= bjhi || R0 >= bi1) return;  // nothing of this cluster is on or above the diagonal (both CTAs agree)
    const int bi = R0 + (int)crank;                        // this CTA's tile row (A operand, TMEM lanes)
    const int n_chunks = (num_kb + CHUNK3 - 1) / CHUNK3;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_cfull + 8 * b, 1); mbar_init(bar_cempty + 8 * b, 16); }   // 8 epilogue warps of each CTA arrive on the LEADER's cempty
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_lo) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                                    // the peer's barriers exist before anything can arrive on them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer: A rows of tile row bi, B rows (= matrix columns) of tile column cp0 + crank =====
            const int bcol = cp0 + (int)crank;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (uint32_t)((kb / STAGES) & 1);
                mbar_wait_b<FW_COR3_SLEEP_NS / 4>(bar_empty + 8 * s, ph ^ 1u);
                const uint32_t full_leader = mapa_cta(bar_full + 8 * s, 0);
                if (crank == 0) mbar_expect_tx(bar_full + 8 * s, 2 * STAGE_BYTES);       // both CTAs' four tiles
                const uint32_t st = base + s * STAGE_BYTES;
                tma_load_2d_cg2(st, &tm_hi, full_leader, kb * BK, bi * BM);
                tma_load_2d_cg2(st + TILE_BYTES, &tm_lo, full_leader, kb * BK, bi * BM);
                tma_load_2d_cg2(st + 2 * TILE_BYTES, &tm_hi, full_leader, kb * BK, bcol * BN);
                tma_load_2d_cg2(st + 3 * TILE_BYTES, &tm_lo, full_leader, kb * BK, bcol * BN);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0 && crank == 0) {
            // ===== MMA issuer (leader CTA): M = 256 (128 per CTA), N = 256 (128 B rows from each CTA), K = 16 =====
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
#ifdef FW_COR3_DEBUG
            long long w_full = 0, w_drain = 0, t_first = 0;
            COR3_DBG(0, t_entry);
#endif
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (uint32_t)((kb / STAGES) & 1);
                const int c = kb / CHUNK3, b = c & 1, u = c >> 1;
                const bool first = (kb % CHUNK3) == 0;
                if (first && c >= 2) {                                            // the epilogue warps of both CTAs have drained chunk c - 2
#ifdef FW_COR3_DEBUG
                    const long long td0 = clock64();
#endif
                    mbar_wait_cluster_b(bar_cempty + 8 * b, (uint32_t)((u - 1) & 1));
#ifdef FW_COR3_DEBUG
                    w_drain += clock64() - td0;
#endif
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
#ifdef FW_COR3_DEBUG
                const long long tw0 = clock64();
#endif
                mbar_wait_b(bar_full + 8 * s, ph);
#ifdef FW_COR3_DEBUG
                if (kb == 0) t_first = clock64(); else w_full += clock64() - tw0;
#endif
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t t_acc = tmem_base + 256u * (uint32_t)b;
                const uint32_t st = base + s * STAGE_BYTES;
                const uint64_t a_hi = make_desc(st), a_lo = make_desc(st + TILE_BYTES);
                const uint64_t b_hi = make_desc(st + 2 * TILE_BYTES), b_lo = make_desc(st + 3 * TILE_BYTES);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);
                    umma_bf16_cg2(t_acc, a_hi + adv, b_hi + adv, idesc, (first && k == 0) ? 0u : 1u);
                    umma_bf16_cg2(t_acc, a_hi + adv, b_lo + adv, idesc, 1u);
                    umma_bf16_cg2(t_acc, a_lo + adv, b_hi + adv, idesc, 1u);
                }
                umma_commit_cg2(bar_empty + 8 * s, (uint16_t)3);                  // frees the stage in both CTAs
                if ((kb % CHUNK3) == CHUNK3 - 1 || kb == num_kb - 1) umma_commit_cg2(bar_cfull + 8 * b, (uint16_t)3);
            }
#ifdef FW_COR3_DEBUG
            COR3_DBG(1, t_first); COR3_DBG(2, clock64()); (void)w_full; (void)w_drain;
#endif
        }
        __syncwarp();
    } else {
        // ===== epilogue warps =====
        const int q = warp & 3, hh = (warp - 2) >> 2;                             // TMEM lane quarter, column half
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const uint32_t col_base = (uint32_t)(hh * 128);
        float acc[128];
#pragma unroll
        for (int j = 0; j < 128; ++j) acc[j] = 0.0f;
        for (int c = 0; c < n_chunks; ++c) {
            const int b = c & 1, u = c >> 1;
            mbar_wait_b<FW_COR3_SLEEP_NS>(bar_cfull + 8 * b, (uint32_t)(u & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int c0 = 0; c0 < 128; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + lane_base + 256u * (uint32_t)b + col_base + (uint32_t)c0, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(v[j]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0 && c + 2 < n_chunks) mbar_arrive_cluster(mapa_cta(bar_cempty + 8 * b, 0));
        }
        // every MMA has retired: the pipeline buffers are idle from here on and serve as the staging area of the tile
#ifdef FW_COR3_DEBUG
        if (warp == 2 && lane == 0 && crank == 0) COR3_DBG(5, clock64());
#endif
        const int bj = cp0 + hh;                                                  // tile column of this warp's 128 columns
        const bool live = bi < bi1 && bj >= bi && bj < bjhi;
        const bool diag = (bi == bj);
        const i64 trow0 = (i64)bi * BM, col0 = (i64)bj * BN;                      // first matrix row / column of the 128 x 128 tile
        // staged tile of this column half: S[r][c], r = row in the tile (all four lane quarters), c = column in the tile
        float* S = reinterpret_cast<float*>(gen) + (size_t)hh * 128 * STG_LD;
        float* stg_row = S + (q * 32 + lane) * STG_LD;
#pragma unroll
        for (int j = 0; j < 128; ++j) {
            float x = acc[j];
            x = x > 1.0f ? 1.0f : (x < -1.0f ? -1.0f : x);            // clampcor (NaN passes through)
            stg_row[j] = x;
        }
        if (diag) stg_row[q * 32 + lane] = 1.0f;                       // cov2cor!: unit diagonal (row == col)
        asm volatile("bar.sync %0, 128;" ::"r"(1 + hh) : "memory");    // the four warps of this column half
#ifdef FW_COR3_DEBUG
        long long t_s1 = clock64(), t_s2 = 0;
#endif
        if (live) {
            // row-sharded mode (several GPUs): tile row t is stored at its local position in this rank's shard (common.cuh, CorView)
            float* Cw = C;
            if (sh_world > 1) { const int g = bi / sh_h; Cw = C + ((i64)((g < sh_world ? 0 : sh_h) + (bi - g * sh_h)) * 128 - (i64)bi * 128) * p; }
            // mirrored tile (+ raw candidates of the univariate stage), then the direct tile: 32 matrix rows of 512 bytes per warp each
            const bool do_mirror = (mirror != 0) || diag;
            if (do_mirror || em.on) {
                if (diag) { if (em.on) mirror_block<true, true>(S, Cw, p, trow0, col0, q, lane, do_mirror, em); else mirror_block<true, false>(S, Cw, p, trow0, col0, q, lane, do_mirror, em); }
                else      { if (em.on) mirror_block<false, true>(S, Cw, p, trow0, col0, q, lane, do_mirror, em); else mirror_block<false, false>(S, Cw, p, trow0, col0, q, lane, do_mirror, em); }
            }
#ifdef FW_COR3_DEBUG
            t_s2 = clock64();
#endif
            // direct tile: warp q writes the tile rows q*32 .. +31, one 128-byte row segment per store instruction
            const int ncol = (int)((p - col0) < 128 ? (p - col0) : 128);              // valid columns of this tile
            for (int rr = 0; rr < 32; ++rr) {
                const i64 r_ = trow0 + q * 32 + rr;
                if (r_ >= p) break;
                float* dst = Cw + r_ * p + col0;
                const float* src = S + (q * 32 + rr) * STG_LD;
                const int cmin = diag ? q * 32 + rr : 0;                              // diagonal tile: upper triangle only
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int c = k * 32 + lane;
                    if (c < ncol && c >= cmin) dst[c] = src[c];
                }
            }
        }
#ifdef FW_COR3_DEBUG
        if (warp == 2 && lane == 0 && crank == 0) { COR3_DBG(6, clock64()); COR3_DBG(3, t_s1); COR3_DBG(4, t_s2); }
#endif
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                                    // no CTA leaves while its peer can still arrive on its barriers / read its smem
#ifdef FW_COR3_DEBUG
    if (warp == 2 && lane == 0 && crank == 0) COR3_DBG(7, clock64());
#endif
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// clusters of cor_tc3_kernel for tile rows [bi0, bi1) x tile columns [bjlo, bjhi)
static long long grouped_clusters3(int bi0, int bi1, int bjlo, int bjhi) {
    long long c = 0;
    for (int r0 = bi0; r0 < bi1; r0 += 2 * COR_GROUP3) {
        const int left = (bi1 - r0 + 1) >> 1;
        const int rows = left < COR_GROUP3 ? left : COR_GROUP3;
        const int lo = r0 > bjlo ? r0 : bjlo;
        if (bjhi > lo) c += (long long)((bjhi - lo + 1) >> 1) * rows;
    }
    return c;
}

static cudaError_t launch_tc3(const Prepared& P, float* d_cor, i64 p, int bi0, int bi1, bool mirror, int bjlo, int bjhi, cudaStream_t st, int* n_launch,
                              std::string* msg, const PwEmit& em, int sh_world, int sh_h) {
    cudaError_t e0 = cudaFuncSetAttribute(cor_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);    // per device
    if (e0 != cudaSuccess) { *msg = "cudaFuncSetAttribute(cor_tc3_kernel)"; return e0; }
    const long long clusters = grouped_clusters3(bi0, bi1, bjlo, bjhi);
    if (clusters <= 0) return cudaSuccess;
    cor_tc3_kernel<<<(unsigned)(2 * clusters), NTHREADS3, SMEM_BYTES, st>>>(P.tm_hi, P.tm_lo, d_cor, p, (int)(P.kp / BK), P.nb, bi0, bi1, mirror ? 1 : 0, bjlo, bjhi, em, sh_world, sh_h);
    (*n_launch)++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { *msg = "cor_tc3_kernel"; return e; }
    return cudaSuccess;
}

}  // namespace cortc

#ifdef FW_COR3_DEBUG
extern "C" int fw_debug_cor3_trace(long long* dev_buf) {
    return (int)cudaMemcpyToSymbol(cortc::g_cor3_dbg, &dev_buf, sizeof(dev_buf));
}
#endif
