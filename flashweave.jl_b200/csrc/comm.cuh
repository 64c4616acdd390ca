// comm.cuh — the GPUs of one node as a group: replaces the worker pool / RemoteChannel plumbing of src/interleaved.jl:76-93
// and the SharedArray table of src/learning.jl:553-560 for parallel="single" semantics (targets are independent given the
// pairwise stage, src/learning.jl:137-138).
//
// One fw_ctx per GPU (normally one process per GPU).  Every rank exports four buffers - its slice of the table, its row shard of
// cor_mat, its list of raw pairwise candidates, a few flag words - as CUDA IPC handles (plain pointers + peer access when the
// contexts live in one process); the host language exchanges the fixed-size handle blobs once (fw_comm_export / fw_comm_attach).
// From then on the data path needs no collective library and no host round trip:
//   * table: every rank uploads 1/world of the columns over its own PCIe link; the standardise+split kernel of the cor_mat GEMM
//     reads each column straight from its owner's HBM over NVLink (staged once in shared memory), i.e. the "all-gather" is fused
//     into the kernel that consumes it;
//   * cor_mat: stays row-sharded where the GEMM produced it (CorView, common.cuh); HITON-PC's gathers dereference the owner's
//     shard through the peer mapping - ~72 MB of reads per pass over all targets at C4 instead of a 10 GB exchange;
//   * pairwise stage: every rank collects the raw candidates of its own tile rows (GEMM epilogue), the compact lists (12 B per
//     candidate) are pulled from the peers and Benjamini-Hochberg (global m and rank order, src/statfuns.jl:326-350) is
//     evaluated redundantly on every rank;
//   * ordering between ranks: a device-side barrier (one flag word per rank, signalled and awaited by tiny kernels on the
//     ranks' own streams, with a time-out so a dead peer can never hang the GPU).
#pragma once
#include <unistd.h>
#include <cuda_runtime.h>
#include "common.cuh"

namespace fwcomm {

struct Handle {                       // what one rank publishes; plain bytes, fixed size (fw_comm_handle_bytes)
    int64_t pid; int32_t device, rank, world, pad_; int64_t n, p, list_cap;
    cudaIpcMemHandle_t ipc[4];        // table slice, cor shard, candidate list, flags
    void* raw[4];                     // the same buffers as device pointers (contexts of one process)
};

enum { B_TABLE = 0, B_COR = 1, B_LIST = 2, B_FLAGS = 3 };
enum { F_SEQ = 0, F_LIST_N = 1, F_NAN_N = 2, N_FLAGS = 16 };

struct PeerFlags { const u64* f[FW_MAX_RANKS]; };
struct PeerSlices { const float* s[FW_MAX_RANKS]; i64 col0[FW_MAX_RANKS + 1]; };

__global__ void signal_kernel(u64* flag, u64 v) {
    __threadfence_system();
    *reinterpret_cast<volatile u64*>(flag) = v;
    __threadfence_system();
}
// thread t waits until rank t has reached barrier `v`; gives up after timeout_clk cycles (err = 1 + rank of the missing peer)
__global__ void wait_kernel(PeerFlags pf, int world, u64 v, long long timeout_clk, int* err) {
    const int t = threadIdx.x;
    if (t < world) {
        const volatile u64* f = reinterpret_cast<const volatile u64*>(pf.f[t]);
        const long long t0 = clock64();
        while (*f < v) {
            if (clock64() - t0 > timeout_clk) { atomicExch(err, 1 + t); break; }
            __nanosleep(256);
        }
    }
    __threadfence_system();
}

// standardise + split (cor_tc.cuh, standardize_split_kernel) with the column read from its OWNER's table slice: one NVLink read
// per element, staged in shared memory (n floats) for the three passes (mean, centred norm, write)
template <int THREADS>
__global__ void __launch_bounds__(THREADS) standardize_split_peer_kernel(PeerSlices ps, int world, i64 n, i64 p, i64 kp,
                                                                         __nv_bfloat16* __restrict__ zhi, __nv_bfloat16* __restrict__ zlo, int staged, i64 rot) {
    extern __shared__ float s_col[];
    // every rank walks the columns starting at a different owner (rot = first column of the next rank's slice): at any moment the
    // 8 GPUs read from 8 different peers instead of all draining one owner's NVLink egress (measured: 11 ms -> 3 ms on the last rank)
    const i64 col = ((i64)blockIdx.x + rot) % (i64)gridDim.x;
    __nv_bfloat16* hi = zhi + col * kp;
    __nv_bfloat16* lo = zlo + col * kp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (col >= p) {
        for (i64 i = tid; i < kp; i += THREADS) { hi[i] = __float2bfloat16_rn(0.f); lo[i] = __float2bfloat16_rn(0.f); }
        return;
    }
    int owner = 0;
    while (owner + 1 < world && col >= ps.col0[owner + 1]) ++owner;
    const float* x = ps.s[owner] + (col - ps.col0[owner]) * n;
    __shared__ double red[THREADS / 32];
    __shared__ double s_mean, s_inv;
    double s = 0.0;
    for (i64 i = tid; i < n; i += THREADS) { const float v = x[i]; if (staged) s_col[i] = v; s += (double)v; }
    const float* xs = staged ? s_col : x;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (tid == 0) { double t = 0.0; for (int w = 0; w < THREADS / 32; ++w) t += red[w]; s_mean = t / (double)n; }
    __syncthreads();
    const double mean = s_mean;
    double ss = 0.0;
    for (i64 i = tid; i < n; i += THREADS) { double d = (double)xs[i] - mean; ss += d * d; }
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_down_sync(0xffffffffu, ss, o);
    __syncthreads();
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    if (tid == 0) { double t = 0.0; for (int w = 0; w < THREADS / 32; ++w) t += red[w]; s_inv = 1.0 / sqrt(t); }
    __syncthreads();
    const double inv = s_inv;
    for (i64 i = tid; i < kp; i += THREADS) {
        float z = (i < n) ? (float)(((double)xs[i] - mean) * inv) : 0.0f;
        __nv_bfloat16 h = __float2bfloat16_rn(z);
        hi[i] = h;
        lo[i] = __float2bfloat16_rn(z - __bfloat162float(h));
    }
}

// cor(idx[i], idx[j]) for a list of variables (fw_cor_gather): works on the full and on the row-sharded matrix
__global__ void cor_gather_kernel(CorView cv, const i64* __restrict__ idx, i64 m, float* __restrict__ out) {
    const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m * m) return;
    out[e] = cv.at(idx[e / m], idx[e % m]);
}

struct Group {
    bool exported = false, attached = false;
    int rank = 0, world = 1, nb = 0, h = 1;
    i64 n = 0, p = 0, list_cap = 0;
    void* own[4] = {nullptr, nullptr, nullptr, nullptr};       // device buffers this rank exports (owned by the context)
    void* peer[FW_MAX_RANKS][4];                               // all ranks' buffers as this device sees them (own included)
    bool opened[FW_MAX_RANKS][4];
    u64 seq = 0;
    int* d_err = nullptr;
    Group() { for (int q = 0; q < FW_MAX_RANKS; ++q) for (int b = 0; b < 4; ++b) { peer[q][b] = nullptr; opened[q][b] = false; } }
    i64 col0(int q) const { return p * q / world; }
    i64 shard_rows() const { return (i64)2 * h * 128; }
    float* cor_of(int q) const { return (float*)peer[q][B_COR]; }
    u64* flags_of(int q) const { return (u64*)peer[q][B_FLAGS]; }
    PwRec* list_of(int q) const { return (PwRec*)peer[q][B_LIST]; }
};

}  // namespace fwcomm
