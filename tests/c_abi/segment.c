/* Plain C client of the C ABI (include/fds_b200.h): no Python, no torch.
 *   gcc -std=c99 -I include tests/c_abi/segment.c -L fds_b200 -lfds_b200 -Wl,-rpath,$PWD/fds_b200 -lm
 * Reads u[N], V[m][N], v[N] (binary doubles) from argv[1] and does what one iteration of the reference's loop
 * does (fds/fds.py:128-170) through the library:
 *   A  run_segment (fds/segment.py:14-82) on the host buffers          -> u', V', v', Jsum[steps][m+2][2]
 *   B  the boundary that ends it (time dilation, projection, QR)        -> R, b, G_dil, g_dil, Q, vq
 *      through the enqueue-only entry points (page-locked result block, one sync)
 * and writes  u'[N] V'[m][N] v'[N] Jsum R[m][m] b[m] G_dil[m] g_dil Q[m][N] vq[N]  to argv[2]. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "fds_b200.h"

#define CK(call) do { if ((call) != 0) { fprintf(stderr, "%s: %s\n", #call, fds_last_error(ctx)); return 1; } } while (0)

int main(int argc, char** argv) {
    if (argc < 8) { fprintf(stderr, "usage: in out N m steps eps s\n"); return 2; }
    const long N = atol(argv[3]);
    const int m = atoi(argv[4]);
    const long steps = atol(argv[5]);
    const double eps = atof(argv[6]), s = atof(argv[7]);
    double* u = malloc(sizeof(double) * N);
    double* V = malloc(sizeof(double) * m * N);
    double* v = malloc(sizeof(double) * N);
    double* J = malloc(sizeof(double) * steps * (m + 2) * 2);
    double* Q = malloc(sizeof(double) * m * N);
    double* vq = malloc(sizeof(double) * N);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(u, 8, N, f) != (size_t)N || fread(V, 8, (size_t)m * N, f) != (size_t)m * N ||
        fread(v, 8, N, f) != (size_t)N) { fprintf(stderr, "bad input\n"); return 2; }
    fclose(f);

    fds_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.system = FDS_SYS_L96_RK4; cfg.math = FDS_MATH_EXACT; cfg.m = m;
    cfg.n_local = N; cfg.n_global = N; cfg.world = 1;
    fds_ctx* ctx = NULL;
    if (fds_ctx_create(&cfg, &ctx) != 0) { fprintf(stderr, "create: %s\n", fds_last_error(NULL)); return 1; }
    /* A: one run_segment call on host buffers, in place */
    CK(fds_run_segment_host(ctx, u, V, v, s, eps, steps, J));
    /* B: the boundary, enqueued; results land in page-locked memory, checked after ONE sync */
    const long nout = (long)fds_boundary_out_doubles(ctx);
    double* out = fds_host_alloc(sizeof(double) * nout);
    if (!out) { fprintf(stderr, "fds_host_alloc failed\n"); return 1; }
    CK(fds_boundary_enqueue(ctx, s, eps, 1, 1, 0, 0, out));
    CK(fds_sync(ctx));
    CK(fds_boundary_check(ctx, out, 1));
    CK(fds_get_tangents(ctx, Q, vq));
    printf("launches %lld version %s qr_passes %d\n", (long long)fds_kernel_launches(ctx), fds_version(),
           fds_last_qr_passes(ctx));

    f = fopen(argv[2], "wb");
    fwrite(u, 8, N, f); fwrite(V, 8, (size_t)m * N, f); fwrite(v, 8, N, f);
    fwrite(J, 8, steps * (m + 2) * 2, f);
    fwrite(out, 8, (size_t)m * m + 2 * m + 1, f);          /* R, b, G_dil, g_dil */
    fwrite(Q, 8, (size_t)m * N, f); fwrite(vq, 8, N, f);
    fclose(f);
    fds_host_free(out);
    fds_ctx_destroy(ctx);
    return 0;
}
