/* fwgpu.h — C ABI of libfwgpu.so: the B200 (sm_100a) conditional-independence test engine
 * that replaces the inner test batches of FlashWeave.jl's HITON-PC skeleton search.
 *
 * The reference has no FFI of its own; its seam is a set of Julia generic functions
 * (SURVEY.md §8b).  Each entry point below names the reference function it replaces
 * (paths relative to the reference checkout).  INTEGRATION.md shows the Julia `ccall`
 * glue a maintainer would add.
 *
 * Conventions
 *  - every function returns an int32 status (FW_OK = 0); nothing throws; the message of
 *    the last failure on a context is fw_last_error(ctx) (fw_last_error(NULL) for
 *    fw_create failures);
 *  - one context per host worker, calls on a context are serialised by the caller
 *    (mirrors one single-threaded Julia process per worker, src/interleaved.jl:90);
 *  - matrices are column-major, n samples x p variables (Julia Matrix layout);
 *  - variable indices are 0-based unless fw_set_index_base(ctx, 1) was called (the Julia
 *    glue does that once);
 *  - host pointers are borrowed for the duration of the call only;
 *  - "insufficient power" is not an error: it is TestResult(0,1,0,false) as in the
 *    reference (src/tests.jl:36-40, 58-62, 210-214, 258-262).
 */
#ifndef FWGPU_H
#define FWGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fw_ctx fw_ctx;

/* src/types.jl:140-145  struct TestResult {stat::Float64; pval::Float64; df::Int; suff_power::Bool}
 * (isbits, 32 bytes with padding: identical layout, so a Julia Vector{TestResult} can be
 * passed directly as the output buffer). */
typedef struct fw_test_result {
    double stat;
    double pval;
    int64_t df;
    uint8_t suff_power;
    uint8_t pad_[7];
} fw_test_result;

/* test_name strings of the reference (src/types.jl:64-72) */
enum fw_test_kind { FW_MI = 0, FW_MI_NZ = 1, FW_FZ = 2, FW_FZ_NZ = 3 };

enum fw_status {
    FW_OK = 0,
    FW_ERR_INVALID = 1,      /* bad argument */
    FW_ERR_CUDA = 2,         /* CUDA runtime / driver failure (message has the CUDA error string) */
    FW_ERR_STATE = 3,        /* call order: e.g. tests before data / cor_mat were provided */
    FW_ERR_UNSUPPORTED = 4,  /* e.g. max_k > 3 */
    FW_ERR_NOMEM = 5
};

/* ---- lifecycle ------------------------------------------------------------------------ */
int32_t fw_create(int32_t device, fw_ctx** out);
int32_t fw_destroy(fw_ctx* ctx);
const char* fw_last_error(fw_ctx* ctx);
int32_t fw_set_index_base(fw_ctx* ctx, int32_t base);          /* 0 (default) or 1 */
/* cudaStream_t all kernels of this context are launched on (for CUDA-event timing). */
void* fw_stream(fw_ctx* ctx);
int32_t fw_synchronize(fw_ctx* ctx);
/* cudaHostRegister / cudaHostUnregister of a caller-owned buffer (e.g. the result arrays of fw_hiton_pc when they are reused
 * across calls, or a SharedArray table): page-locked buffers move at full PCIe speed and asynchronously */
int32_t fw_host_register(fw_ctx* ctx, void* ptr, int64_t bytes);
int32_t fw_host_unregister(fw_ctx* ctx, void* ptr);
/* number of kernel launches issued by this context since creation (bench.py's gpu_launches) */
int64_t fw_launch_count(fw_ctx* ctx);
/* device time in ms (CUDA events on the context's stream) of the last run of each phase:
 * out[0] cor_mat kernels, out[1] pairwise-stage kernels, out[2] HITON-PC kernel(s), out[3] reserved, out[4] the standardising
 * kernel inside out[0] of fw_multi_cor, out[5] the wait in fw_multi_cor's closing group barrier; -1 = not run */
int32_t fw_last_timing(fw_ctx* ctx, double* out_ms, int32_t n);

/* ---- data (the `data` argument of every reference test function) ----------------------- */
/* continuous table (fz / fz_nz): Matrix{Float32}, prec=32 (src/learning.jl:470, misc.jl:54-62) */
int32_t fw_set_data_f32(fw_ctx* ctx, const float* host, int64_t n, int64_t p, int64_t ld);
/* discrete table (mi / mi_nz): Matrix{Int32} level codes >= 0 */
int32_t fw_set_data_i32(fw_ctx* ctx, const int32_t* host, int64_t n, int64_t p, int64_t ld);
/* sparse tables: SparseMatrixCSC{Float32,Int64} / SparseMatrixCSC{Int32,Int64} as (colptr[p+1], rowval[nnz], nzval[nnz]) with the
 * context's index base (fw_set_index_base(ctx, 1) for Julia's arrays).  Only the triple crosses PCIe; the table is densified
 * on the device and then behaves exactly like the dense one (the dense test semantics are the canonical ones). */
int32_t fw_set_data_csc_f32(fw_ctx* ctx, const int64_t* colptr, const int64_t* rowval, const float* nzval, int64_t n, int64_t p);
int32_t fw_set_data_csc_i32(fw_ctx* ctx, const int64_t* colptr, const int64_t* rowval, const int32_t* nzval, int64_t n, int64_t p);
/* Which of the reference's two code paths the discrete zero-ignoring kind (mi_nz) follows.  FW_SEMANTICS_DENSE (default): the
 * Matrix{Int32} path (src/contingency.jl:7-56 + level_map!), the canonical semantics.  FW_SEMANTICS_SPARSE: the SparseMatrixCSC path
 * the reference takes by default for sensitive=false (make_sparse, src/learning.jl:470): identical tables, but levels_z of the power
 * rule (src/tests.jl:210) is max(z)+1 in the max_k = 1 specialisation (src/contingency.jl:171-173, 229) and may count the back-fill
 * stratum of the generic merge (src/contingency.jl:461-477); rows are skipped per max_val > 1 instead of trimmed per levels > 2.
 * Applies to the table however it was uploaded (dense or fw_set_data_csc_i32); mi / fz / fz_nz are unaffected. */
enum fw_semantics { FW_SEMANTICS_DENSE = 0, FW_SEMANTICS_SPARSE = 1 };
int32_t fw_set_semantics(fw_ctx* ctx, int32_t semantics);
/* same, device-resident inputs owned by the caller (e.g. a tensor that was NCCL-broadcast);
 * the pointer must stay valid until the next fw_set_data / fw_adopt_data / fw_destroy */
int32_t fw_adopt_data_f32_device(fw_ctx* ctx, const float* dev, int64_t n, int64_t p, int64_t ld);
/* meta_variable_mask of the resident table (src/preprocessing.jl:418-446; `# meta mask` line of src/io.jl:338-346).  Meta variables
 * are ordinary variables for every test; the mask is stored with the table (cleared when a new table is installed) and handed back to
 * the host side that writes the network. */
int32_t fw_set_meta_mask(fw_ctx* ctx, const uint8_t* mask, int64_t p);
int32_t fw_get_meta_mask(fw_ctx* ctx, uint8_t* mask_out, int64_t p);
/* number of observations used by Fisher-z tests when only a cor_mat is installed
 * (size(data,1) in src/tests.jl:150,256) */
int32_t fw_set_n_obs(fw_ctx* ctx, int64_t n);

/* ---- normalisation: the step in front of the hot path --------------------------------------
 * normalize_data / preprocess_data for dense tables without meta variables (src/preprocessing.jl:412-563, 660-684):
 * filter_by_variance (:367-409; variables with a single value, then samples without reads), the normalisation proper, the
 * level filters of the discrete modes, conversion to the target precision (Float32 / Int32).  `host` is the column-major n x p
 * table as Float32 (check_convert_sparse, :579-594).  The result becomes the context's resident table exactly as if it had been
 * passed to fw_set_data_f32 (FW_NORM_ROWS, _CLR_ADAPT, _CLR_NZ) or fw_set_data_i32 (the others); fw_get_data_* copies it out.
 * row_mask[n] / col_mask[p] (may be NULL) receive the obs_filter_mask and the kept-variable mask; *n_out x *p_out is the new shape
 * (0 variables or samples left: FW_OK, no table resident). */
enum fw_norm_mode {
    FW_NORM_ROWS = 0,             /* "tss"                 rownorm!                                   :348        */
    FW_NORM_CLR_ADAPT = 1,        /* "clr-adapt"   (fz)    adaptive_clr!                              :133-215    */
    FW_NORM_CLR_NZ = 2,           /* "clr-nonzero" (fz_nz) clr!(ignore_zeros = true)                  :192-207    */
    FW_NORM_BINARY = 3,           /* "pres-abs"    (mi)    presabs_norm! + exactly-2-levels filter    :364, :475  */
    FW_NORM_BINNED_NZ_CLR = 4,    /* "clr-nonzero-binned" (mi_nz)  clr_nz + discretize_nz (tied ranks) :217-292   */
    FW_NORM_BINNED_NZ_ROWS = 5    /* "tss-nonzero-binned"          rownorm! + discretize_nz                       */
};
int32_t fw_normalize_f32(fw_ctx* ctx, const float* host, int64_t n, int64_t p, int64_t ld, int32_t norm_mode, int32_t n_bins,
                         int64_t* n_out, int64_t* p_out, uint8_t* row_mask, uint8_t* col_mask);
/* copy the resident table to the host (column-major, leading dimension ld >= n) */
int32_t fw_get_data_f32(fw_ctx* ctx, float* host_out, int64_t ld);
int32_t fw_get_data_i32(fw_ctx* ctx, int32_t* host_out, int64_t ld);

/* ---- precompute (src/learning.jl:33-47 prepare_lgl) ------------------------------------ */
/* get_levels / get_max_vals, src/misc.jl:64-97 */
int32_t fw_levels(fw_ctx* ctx, int32_t* levels, int32_t* max_vals);
/* cor_mat = convert(Matrix{Float32}, cor(data)), src/learning.jl:42-44.  Computed on the
 * tensor cores from the resident table and kept resident; host_out may be NULL. */
int32_t fw_cor_matrix(fw_ctx* ctx, float* host_out);
/* fw_set_data_f32 followed by fw_cor_matrix as one call: the table is uploaded in column chunks on a copy stream while the
 * tiles whose operands have arrived are already being computed (same result, the PCIe transfer hides behind the GEMM). */
int32_t fw_upload_cor_f32(fw_ctx* ctx, const float* host, int64_t n, int64_t p, int64_t ld, float* host_out);
/* install a caller-provided cor_mat (the `cor_mat` field of FzTest/FzTestCond, src/types.jl:126-136) */
int32_t fw_set_cor_f32(fw_ctx* ctx, const float* host_cor, int64_t p);
int32_t fw_adopt_cor_device(fw_ctx* ctx, const float* dev_cor, int64_t p);
/* device pointer of the resident cor_mat (p*p floats), e.g. to broadcast it with NCCL */
void* fw_cor_device_ptr(fw_ctx* ctx);
/* Row-sharded cor_mat for several GPUs (same src/learning.jl:42-44 result, split over ranks): adopt a buffer of
 * rows_allocated >= p rows, standardise the table once (fw_cor_prepare returns the number of 128-row tile rows), compute the
 * upper-triangular tiles of this rank's tile rows (fw_cor_rows), exchange the row blocks with NCCL (all-gather on the adopted
 * buffer, done by the host language), then copy the upper triangle to the lower one (fw_cor_symmetrize).  The result is
 * bit-identical to fw_cor_matrix on one GPU. */
int32_t fw_adopt_cor_device_rows(fw_ctx* ctx, const float* dev_cor, int64_t p, int64_t rows_allocated);
int32_t fw_cor_prepare(fw_ctx* ctx, int32_t* n_tile_rows);
int32_t fw_cor_rows(fw_ctx* ctx, int32_t tile_row_begin, int32_t tile_row_end);
int32_t fw_cor_symmetrize(fw_ctx* ctx);

/* cor(idx[i], idx[j]) for a list of m variables (row-major m x m): a sub-matrix of the resident cor_mat without moving all of
 * it - the only way to look at a row-sharded matrix from the host */
int32_t fw_cor_gather(fw_ctx* ctx, const int64_t* idx, int64_t m, float* host_out);

/* ---- multi-GPU: the GPUs of one node as a group -------------------------------------------------------------------------
 * Replaces the worker pool / RemoteChannels of src/interleaved.jl:76-93 and the SharedArray table of src/learning.jl:553-560
 * for parallel="single" semantics (targets are independent given the pairwise stage, src/learning.jl:137-138).  One context per
 * GPU, normally one process per GPU.  Setup: every rank calls fw_comm_export (allocates its share of an n x p job and writes a
 * fixed-size opaque handle of fw_comm_handle_bytes() bytes), the host language all-gathers the handles (Distributed /
 * torch.distributed / a file), every rank calls fw_comm_attach with the world x handle_bytes blob in rank order (CUDA IPC
 * mappings; plain peer access between contexts of one process).  From then on no collective library and no host round trip is
 * involved in the data path (csrc/comm.cuh): each rank uploads 1/world of the table columns over its own PCIe link
 * (fw_multi_set_data_f32: host_slice points at column p*rank/world of the column-major table), fw_multi_cor computes the rank's
 * tile rows of cor_mat - reading every column from its owner's HBM over NVLink inside the standardising kernel - and leaves the
 * matrix ROW-SHARDED; fw_pairwise (FW_FZ) pulls the peers' compact candidate lists for the global Benjamini-Hochberg step;
 * fw_hiton_pc, fw_test_batch, fw_test_subsets and fw_cor_gather dereference the owners' shards through the peer mappings.
 * The fw_multi_* calls and fw_pairwise are COLLECTIVE in a group: every rank must issue the same sequence (ordering between
 * ranks is a device-side barrier on the ranks' own streams, with a time-out: a missing peer is an error, never a hang).
 * Results are identical to one GPU.  Contexts of one process need one host thread per context. */
int32_t fw_comm_handle_bytes(void);
int32_t fw_comm_export(fw_ctx* ctx, int32_t rank, int32_t world, int64_t n, int64_t p, void* handle_out);
int32_t fw_comm_attach(fw_ctx* ctx, const void* handles /* world x fw_comm_handle_bytes(), rank order */);
int32_t fw_comm_detach(fw_ctx* ctx);
int32_t fw_multi_set_data_f32(fw_ctx* ctx, const float* host_slice, int64_t ld);
int32_t fw_multi_cor(fw_ctx* ctx);

/* ---- single tests ---------------------------------------------------------------------- */
/* test(X, Y, Zs, data, test_obj, ...) for a batch of independent tests
 * (src/tests.jl:28 / :108 with k = 0, :184 / :250 with k in 1..3).  Zs is n_tests x 3,
 * row-major, entries beyond k[i] ignored.  Row trimming for the _nz kinds is applied by
 * the engine exactly as the callers in src/tests.jl:412-416 and src/hiton.jl:41-50,85 do. */
int32_t fw_test_batch(fw_ctx* ctx, int32_t kind, int64_t n_tests, const int64_t* X, const int64_t* Y,
                      const int32_t* k, const int64_t* Zs, int64_t hps, int64_t n_obs_min,
                      fw_test_result* out);

/* ---- subset search: drop-in for test_subsets, src/tests.jl:281-346 ---------------------- */
/* Returns what the reference returns: the first non-significant result in the reference's
 * enumeration order (subset size max_k..1, lexicographic in Z_total's order), or the
 * maximum-p-value result (ties: later subset) when all are significant; out_Zs[0..out_k)
 * the subset; num_tests as counted at src/tests.jl:322; frac = num_tests / total.
 * Empty Z_total gives the sentinel (NaN, NaN, -1, true), Zs = (-1,), num_tests = -1, frac = NaN. */
int32_t fw_test_subsets(fw_ctx* ctx, int32_t kind, int64_t X, int64_t Y, const int64_t* Z_total, int64_t m,
                        int32_t max_k, double alpha, int64_t hps, int64_t n_obs_min, int64_t max_tests,
                        fw_test_result* out_result, int64_t* out_Zs, int32_t* out_k,
                        int64_t* num_tests, double* frac);
/* many (X, Y, Z_total) jobs in one launch; Z lists as CSR (z_off[n_jobs+1], z_idx[]) */
int32_t fw_test_subsets_batch(fw_ctx* ctx, int32_t kind, int64_t n_jobs, const int64_t* X, const int64_t* Y,
                              const int64_t* z_off, const int64_t* z_idx,
                              int32_t max_k, double alpha, int64_t hps, int64_t n_obs_min, int64_t max_tests,
                              fw_test_result* out_result, int64_t* out_Zs, int32_t* out_k,
                              int64_t* num_tests, double* frac);

/* ---- pairwise stage: pw_univar_neighbors, src/tests.jl:436-532 -------------------------- */
/* All p(p-1)/2 univariate tests, Benjamini-Hochberg (src/statfuns.jl:326-350) when fdr != 0,
 * neighbour lists (var -> nbr -> (stat, adjusted p)), neighbours in ascending index
 * (src/tests.jl:372-388).  The result stays resident on the device (it feeds fw_hiton_pc)
 * and can be copied out as CSR. */
int32_t fw_pairwise(fw_ctx* ctx, int32_t kind, double alpha, int64_t hps, int64_t n_obs_min,
                    int32_t fdr, int32_t correct_reliable_only, int64_t* n_entries);
/* Announce the (alpha, n_obs_min) of the next fw_pairwise(FW_FZ): fw_cor_matrix / fw_upload_cor_f32 / fw_multi_cor then collect
 * the raw candidates (|r| within reach of the threshold) in the GEMM epilogue, where every tile is still in registers, and
 * fw_pairwise does not read the p x p matrix again (src/tests.jl:470-478 looks each correlation up a second time).  Purely an
 * optimisation: with other parameters, or without the announcement, fw_pairwise scans the resident matrix once.  alpha <= 0 disarms. */
int32_t fw_pairwise_prefetch(fw_ctx* ctx, double alpha, int64_t n_obs_min);
int32_t fw_pairwise_copy(fw_ctx* ctx, int64_t* offsets /* p+1 */, int64_t* nbr, double* stat, double* adjp);
/* Multi-GPU, table-based kinds (FW_MI, FW_MI_NZ, FW_FZ_NZ): pw_univar_neighbors split over `world` ranks that each hold the whole
 * table (src/tests.jl:464, 494 hands the X rows of the pairwise loop to the workers with @distributed; here the X variables are dealt
 * to the ranks in balanced groups).  fw_pairwise_partial evaluates every pair (X, Y > X) whose X belongs to `rank` and keeps the
 * raw-significant records (p < alpha) on the device: n_raw of them, and n_reliable = the rank's tests that enter the FDR correction
 * (meaningful with correct_reliable_only; otherwise m = p (p - 1) / 2).  fw_pairwise_partial_copy returns the records (opaque to
 * the host, 0-based).  The host language concatenates the records of all ranks in any order (Distributed / torch.distributed
 * all-gather: this is the one exchange step of the pairwise stage, statfuns.jl:326-350 needs every p-value) and every rank calls
 * fw_pairwise_merge with the concatenation and m_tests = sum of n_reliable (or p (p - 1) / 2): condensed-index order,
 * Benjamini-Hochberg and the neighbour lists exactly as fw_pairwise would produce them on one GPU.  The record arrays of
 * fw_pairwise_partial_copy / fw_pairwise_merge may be host or device memory (unified addressing): with NCCL the records never
 * leave the GPUs.  Host records are validated (0 <= x < y < p, FW_ERR_INVALID otherwise); device records are taken as the ranks'
 * fw_pairwise_partial produced them. */
int32_t fw_pairwise_partial(fw_ctx* ctx, int32_t kind, double alpha, int64_t hps, int64_t n_obs_min, int32_t correct_reliable_only,
                            int32_t rank, int32_t world, int64_t* n_raw, int64_t* n_reliable);
int32_t fw_pairwise_partial_copy(fw_ctx* ctx, int32_t* x, int32_t* y, double* pval, double* stat);
int32_t fw_pairwise_merge(fw_ctx* ctx, int32_t kind, double alpha, int32_t fdr, int64_t n_raw_total, const int32_t* x, const int32_t* y,
                          const double* pval, const double* stat, int64_t m_tests, int64_t* n_entries);
/* install caller-provided neighbour lists (the `all_univar_nbrs` argument of LGL, src/learning.jl:213,235-242) */
int32_t fw_set_univar_nbrs(fw_ctx* ctx, const int64_t* offsets /* p+1 */, const int64_t* nbr,
                           const double* stat, const double* adjp);
/* counters of the last fw_pairwise: tests evaluated, reliable (non-NaN) tests = BH's m, raw p < alpha */
int32_t fw_pairwise_stats(fw_ctx* ctx, int64_t* n_tests, int64_t* n_reliable, int64_t* n_raw_sig);

/* ---- per-target loop: si_HITON_PC, src/hiton.jl:283-400 (time_limit = 0) ---- */
/* Runs interleaving + elimination for every target in targets[] on the device, one CTA per
 * target, against the resident neighbour lists (fw_pairwise / fw_set_univar_nbrs).
 * Outputs per target t (slot range pc_off[t]..pc_off[t+1], capacity = its candidate count):
 *   PC  = HitonState.state_results (after update_PC_dict!, src/hiton.jl:249-256), insertion order
 *   TPC = HitonState.inter_results
 *   num_tests[t] = sum of test_subsets' num_tests (src/tests.jl:322) over the target's candidates.
 * Any output pointer may be NULL. */
int32_t fw_hiton_pc(fw_ctx* ctx, int32_t kind, int64_t n_targets, const int64_t* targets,
                    int32_t max_k, double alpha, int64_t hps, int64_t n_obs_min, int64_t max_tests,
                    int64_t* pc_off /* n_targets+1 */, int64_t* pc_count, int64_t* pc_nbr, double* pc_stat, double* pc_p,
                    int64_t* tpc_count, int64_t* tpc_nbr, double* tpc_stat, double* tpc_p,
                    int64_t* num_tests, int64_t* tests_executed_total);
/* The same with the remaining keyword arguments of si_HITON_PC (src/hiton.jl:283-292):
 *  - whitelist / blacklist per target as CSR over targets[] (wl_off[n_targets+1], wl_idx[]; NULL = none), src/hiton.jl:20-38:
 *    a whitelisted candidate is accepted untested with (NaN, NaN) in both phases - in the elimination phase it is pushed onto
 *    `accepted` a second time, exactly as the reference does - a blacklisted one is skipped.  This is what the feed-forward
 *    schedule of src/interleaved.jl:124-128 (the reference's default single_il / multi_il modes) passes per target job;
 *  - track_rejections (src/hiton.jl:72-74): rej_count != NULL requests HitonState.state_rejections, per target in its slot range
 *    pc_off[t]..: rejected candidate, the subset Zs (rej_k entries of rej_Zs[3], -1 padded) and TestResult that rejected it,
 *    (num_tests, frac) of that test_subsets call. */
int32_t fw_hiton_pc_ex(fw_ctx* ctx, int32_t kind, int64_t n_targets, const int64_t* targets,
                       int32_t max_k, double alpha, int64_t hps, int64_t n_obs_min, int64_t max_tests,
                       const int64_t* wl_off, const int64_t* wl_idx, const int64_t* bl_off, const int64_t* bl_idx,
                       int64_t* pc_off, int64_t* pc_count, int64_t* pc_nbr, double* pc_stat, double* pc_p,
                       int64_t* tpc_count, int64_t* tpc_nbr, double* tpc_stat, double* tpc_p,
                       int64_t* num_tests, int64_t* tests_executed_total,
                       int64_t* rej_count, int64_t* rej_nbr, int64_t* rej_Zs, int32_t* rej_k, fw_test_result* rej_result,
                       int64_t* rej_num_tests, double* rej_frac);
/* tests executed by the last fw_hiton_pc with |Zs| = 1, 2, 3 (for the algorithmic-bytes figure of the roofline) */
int32_t fw_hiton_exec_by_k(fw_ctx* ctx, int64_t* out3);
/* capacity query for the arrays above: sum over targets of their candidate counts */
int32_t fw_hiton_pc_capacity(fw_ctx* ctx, int64_t n_targets, const int64_t* targets, int64_t* capacity);

/* build information: "sm_100a", compiler version */
const char* fw_build_info(void);

#ifdef __cplusplus
}
#endif
#endif /* FWGPU_H */
