/* tests/capi/c_abi_batch.c — a plain-C program (no Python, no C++ runtime of its own) that binds the drop-in boundary the way
 * a C host would: dlopen(libtinympc_b200.so), look the entry points of include/tinympc_b200.h up by name, set a problem
 * up (the reference's cartpole example, examples/cartpole_example.cpp:32-70, fp64), solve a batch of B instances from HOST
 * buffers and dump x0 | sol_x | sol_u | iter | solved | residuals to a binary file that tests/test_gpu_capi.py compares
 * with the oracle.   usage: c_abi_batch <path/to/libtinympc_b200.so> <B> <out.bin>                                       */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tinympc_b200.h"

#define NX 4
#define NU 1
#define NH 10

typedef int (*fn_default_settings)(tinympc_settings_t *);
typedef int (*fn_precompute)(int32_t, int32_t, int32_t, double, const void *, const void *, const void *, const void *, const void *, void *,
                             void *, void *, void *, void *, void *);
typedef int (*fn_create)(const tinympc_problem_t *, int32_t, tinympc_b200_solver_t **);
typedef int (*fn_destroy)(tinympc_b200_solver_t *);
typedef int (*fn_update_settings)(tinympc_b200_solver_t *, const tinympc_settings_t *);
typedef int (*fn_solve_host)(tinympc_b200_solver_t *, const tinympc_batch_t *);
typedef int (*fn_get_stats)(const tinympc_b200_solver_t *, tinympc_b200_stats_t *);
typedef const char *(*fn_str)(void);

static void *must(void *lib, const char *name) {
    void *p = dlsym(lib, name);
    if (!p) {
        fprintf(stderr, "missing symbol %s\n", name);
        exit(2);
    }
    return p;
}

int main(int argc, char **argv) {
    if (argc < 4) return 64;
    void *lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
    if (!lib) {
        fprintf(stderr, "dlopen: %s\n", dlerror());
        return 2;
    }
    const long B = atol(argv[2]);
    fn_default_settings default_settings = (fn_default_settings)must(lib, "tinympc_b200_default_settings");
    fn_precompute precompute = (fn_precompute)must(lib, "tinympc_b200_precompute_cache");
    fn_create create = (fn_create)must(lib, "tinympc_b200_create");
    fn_destroy destroy = (fn_destroy)must(lib, "tinympc_b200_destroy");
    fn_update_settings update_settings = (fn_update_settings)must(lib, "tinympc_b200_update_settings");
    fn_solve_host solve_host = (fn_solve_host)must(lib, "tinympc_b200_solve_host");
    fn_get_stats get_stats = (fn_get_stats)must(lib, "tinympc_b200_get_stats");
    fn_str last_error = (fn_str)must(lib, "tinympc_b200_last_error");

    /* column-major A (4x4), B (4x1); Q + rho, R + rho as tiny_setup stores them (tiny_api.cpp:117-118) */
    const double rho = 1.0;
    const double A[NX * NX] = {1.0, 0.0, 0.0, 0.0, 0.01, 1.0, 0.0, 0.0, 0.0, 0.039, 1.002, 0.458, 0.0, 0.0, 0.01, 1.002};
    const double Bm[NX * NU] = {0.0, 0.02, 0.0, 0.067};
    const double f[NX] = {0, 0, 0, 0};
    const double Q[NX] = {10.0 + 1.0, 1.0 + 1.0, 10.0 + 1.0, 1.0 + 1.0};
    const double R[NU] = {1.0 + 1.0};
    double Kinf[NU * NX], Pinf[NX * NX], Quu_inv[NU * NU], AmBKt[NX * NX], APf[NX], BPf[NU];
    int sweeps = precompute(TINYMPC_F64, NX, NU, rho, A, Bm, f, Q, R, Kinf, Pinf, Quu_inv, AmBKt, APf, BPf);
    if (sweeps <= 0) {
        fprintf(stderr, "precompute: %s\n", last_error());
        return 3;
    }
    double x_min[NX * NH], x_max[NX * NH], u_min[NU * (NH - 1)], u_max[NU * (NH - 1)];
    for (int i = 0; i < NX * NH; ++i) { x_min[i] = -5.0; x_max[i] = 5.0; }
    for (int i = 0; i < NU * (NH - 1); ++i) { u_min[i] = -3.0; u_max[i] = 3.0; }

    tinympc_problem_t p;
    memset(&p, 0, sizeof(p));
    p.nx = NX; p.nu = NU; p.N = NH; p.dtype = TINYMPC_F64; p.rho = rho;
    p.Adyn = A; p.Bdyn = Bm; p.fdyn = f; p.Q = Q; p.R = R;
    p.Kinf = Kinf; p.Pinf = Pinf; p.Quu_inv = Quu_inv; p.AmBKt = AmBKt; p.APf = APf; p.BPf = BPf;
    p.x_min = x_min; p.x_max = x_max; p.u_min = u_min; p.u_max = u_max;
    tinympc_b200_solver_t *h = NULL;
    if (create(&p, 0, &h)) {
        fprintf(stderr, "create: %s\n", last_error());
        return 4;
    }
    tinympc_settings_t st;
    default_settings(&st);
    st.max_iter = 100; /* cartpole_example.cpp:58 */
    update_settings(h, &st);

    double *x0 = malloc(sizeof(double) * B * NX), *Xref = malloc(sizeof(double) * NH * NX);
    double *sol_x = malloc(sizeof(double) * B * NH * NX), *sol_u = malloc(sizeof(double) * B * (NH - 1) * NU);
    double *res = malloc(sizeof(double) * B * 4);
    int32_t *iter = malloc(sizeof(int32_t) * B), *solved = malloc(sizeof(int32_t) * B);
    for (long b = 0; b < B; ++b) { /* a fan of initial cart positions / pole angles */
        x0[b * NX + 0] = 0.5 - 0.01 * (double)(b % 97);
        x0[b * NX + 1] = 0.0;
        x0[b * NX + 2] = 0.002 * (double)(b % 31);
        x0[b * NX + 3] = 0.0;
    }
    for (int k = 0; k < NH; ++k) { Xref[k * NX + 0] = 1.0; Xref[k * NX + 1] = Xref[k * NX + 2] = Xref[k * NX + 3] = 0.0; }
    tinympc_batch_t io;
    memset(&io, 0, sizeof(io));
    io.B = B; io.x0 = x0; io.Xref = Xref; io.xref_per_instance = 0; io.Uref = NULL; io.cold_start = 1;
    io.sol_x = sol_x; io.sol_u = sol_u; io.iter = iter; io.solved = solved; io.residuals = res;
    if (solve_host(h, &io)) {
        fprintf(stderr, "solve_host: %s\n", last_error());
        return 5;
    }
    tinympc_b200_stats_t s;
    get_stats(h, &s);
    FILE *fo = fopen(argv[3], "wb");
    if (!fo) return 6;
    fwrite(x0, sizeof(double), B * NX, fo);
    fwrite(sol_x, sizeof(double), B * NH * NX, fo);
    fwrite(sol_u, sizeof(double), B * (NH - 1) * NU, fo);
    fwrite(iter, sizeof(int32_t), B, fo);
    fwrite(solved, sizeof(int32_t), B, fo);
    fwrite(res, sizeof(double), B * 4, fo);
    fclose(fo);
    long nsolved = 0, iters = 0;
    for (long b = 0; b < B; ++b) { nsolved += solved[b]; iters += iter[b]; }
    printf("c_abi_batch: B=%ld solved=%ld iters=%ld riccati_sweeps=%d kernel_family=%d launches=%lld\n", B, nsolved, iters, sweeps,
           s.kernel_family, (long long)s.kernel_launches);
    destroy(h);
    dlclose(lib);
    return 0;
}
