/* Plain C client of the C ABI (include/fds_b200.h): no Python, no torch.
 *   gcc -std=c99 -I include tests/c_abi/segment.c -L fds_b200 -lfds_b200 -Wl,-rpath,$PWD/fds_b200 -lm
 * Reads u[N], V[m][N], v[N] (binary doubles) from argv[1], runs one boundary (time dilation + projection)
 * + one segment + the closing boundary with QR through the library, writes
 * u'[N], V'[m][N], v'[N], Jsum[steps][m+2][2], R[m][m], b[m], G_dil[m], g_dil to argv[2]. */
#include <stdio.h>
#include <stdlib.h>
#include "fds_b200.h"

#define CK(call) do { if ((call) != 0) { fprintf(stderr, "%s: %s\n", #call, fds_last_error(ctx)); return 1; } } while (0)

int main(int argc, char** argv) {
    if (argc < 7) { fprintf(stderr, "usage: in out N m steps eps\n"); return 2; }
    const long N = atol(argv[3]);
    const int m = atoi(argv[4]);
    const long steps = atol(argv[5]);
    const double eps = atof(argv[6]), s = 0.1;
    double* u = malloc(sizeof(double) * N);
    double* V = malloc(sizeof(double) * m * N);
    double* v = malloc(sizeof(double) * N);
    double* J = malloc(sizeof(double) * steps * (m + 2) * 2);
    double* R = malloc(sizeof(double) * m * m);
    double* b = malloc(sizeof(double) * m);
    double* Gd = malloc(sizeof(double) * m);
    double gd = 0.0;
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(u, 8, N, f) != (size_t)N || fread(V, 8, (size_t)m * N, f) != (size_t)m * N ||
        fread(v, 8, N, f) != (size_t)N) { fprintf(stderr, "bad input\n"); return 2; }
    fclose(f);

    fds_config cfg = {0};
    cfg.system = FDS_SYS_L96_RK4; cfg.math = FDS_MATH_EXACT; cfg.m = m;
    cfg.n_local = N; cfg.n_global = N; cfg.world = 1;
    fds_ctx* ctx = NULL;
    if (fds_ctx_create(&cfg, &ctx) != 0) { fprintf(stderr, "create: %s\n", fds_last_error(NULL)); return 1; }
    CK(fds_set_state(ctx, u));
    CK(fds_set_tangents(ctx, V, v));
    /* first boundary of a run: time dilation + projection only, builds u + eps [V; v]        */
    CK(fds_boundary(ctx, s, eps, 0, 0, FDS_BUILD_ENSEMBLE, 0, NULL, NULL, Gd, &gd));
    CK(fds_segment(ctx, s, eps, steps, J));
    /* closing boundary: dilation, projection, QR; keep the tangents for the download           */
    CK(fds_boundary(ctx, s, eps, 1, 1, 0, 0, R, b, Gd, &gd));
    CK(fds_get_state(ctx, u));
    CK(fds_get_tangents(ctx, V, v));
    printf("launches %lld version %s\n", (long long)fds_kernel_launches(ctx), fds_version());
    fds_ctx_destroy(ctx);

    f = fopen(argv[2], "wb");
    fwrite(u, 8, N, f); fwrite(V, 8, (size_t)m * N, f); fwrite(v, 8, N, f);
    fwrite(J, 8, steps * (m + 2) * 2, f); fwrite(R, 8, m * m, f); fwrite(b, 8, m, f);
    fwrite(Gd, 8, m, f); fwrite(&gd, 8, 1, f);
    fclose(f);
    return 0;
}
