/* Minimal C host for libfwgpu.so: the same sequence the Julia glue of INTEGRATION.md issues through ccall.
 *   gcc -std=c99 -Iinclude examples/fw_demo.c -Lflashweave.jl_b200 -lfwgpu -Wl,-rpath,$PWD/flashweave.jl_b200 -lm -o fw_demo
 * Column-major n x p Float32 table in, neighbour lists of every variable out (fz, max_k = 3). */
#include <stdio.h>
#include <stdlib.h>
#include "fwgpu.h"

#define CK(call) do { int32_t st_ = (call); if (st_ != FW_OK) { fprintf(stderr, "%s failed [%d]: %s\n", #call, st_, fw_last_error(ctx)); return 1; } } while (0)

int main(void) {
    const int64_t n = 400, p = 64;
    fw_ctx* ctx = NULL;
    if (fw_create(0, &ctx) != FW_OK) { fprintf(stderr, "fw_create: %s\n", fw_last_error(NULL)); return 1; }
    float* x = (float*)malloc(sizeof(float) * n * p);
    unsigned s = 12345u;
    for (int64_t v = 0; v < p; ++v)                                  /* blocks of 8 variables sharing a factor */
        for (int64_t i = 0; i < n; ++i) {
            unsigned f = (unsigned)(i * 2654435761u) ^ (unsigned)((v / 8) * 40503u);
            s = s * 1664525u + 1013904223u;
            x[v * n + i] = 0.8f * ((float)(f % 1000) / 1000.0f - 0.5f) + 0.6f * ((float)(s >> 8) / 16777216.0f - 0.5f);
        }
    CK(fw_set_data_f32(ctx, x, n, p, n));
    CK(fw_cor_matrix(ctx, NULL));                                     /* cor_mat stays on the device */
    int64_t n_entries = 0;
    CK(fw_pairwise(ctx, FW_FZ, 0.01, 5, 20, 1, 1, &n_entries));
    int64_t* targets = (int64_t*)malloc(sizeof(int64_t) * p);
    for (int64_t v = 0; v < p; ++v) targets[v] = v;
    int64_t cap = 0;
    CK(fw_hiton_pc_capacity(ctx, p, targets, &cap));
    int64_t* off = (int64_t*)malloc(sizeof(int64_t) * (p + 1));
    int64_t* cnt = (int64_t*)malloc(sizeof(int64_t) * p);
    int64_t* nbr = (int64_t*)malloc(sizeof(int64_t) * (cap > 0 ? cap : 1));
    double* st = (double*)malloc(sizeof(double) * (cap > 0 ? cap : 1));
    double* pv = (double*)malloc(sizeof(double) * (cap > 0 ? cap : 1));
    int64_t* ntests = (int64_t*)malloc(sizeof(int64_t) * p);
    int64_t executed = 0;
    CK(fw_hiton_pc(ctx, FW_FZ, p, targets, 3, 0.01, 5, 20, 10000000, off, cnt, nbr, st, pv, NULL, NULL, NULL, NULL, ntests, &executed));
    int64_t edges = 0, tests = 0;
    for (int64_t v = 0; v < p; ++v) { edges += cnt[v]; tests += ntests[v]; }
    printf("%lld univariate entries, %lld directed PC entries, %lld conditional tests (%lld executed)\n",
           (long long)n_entries, (long long)edges, (long long)tests, (long long)executed);
    fw_destroy(ctx);
    free(x); free(targets); free(off); free(cnt); free(nbr); free(st); free(pv); free(ntests);
    return 0;
}
