/* casadi_external_test.c -- drives libkite_casadi.so the way casadi::external("NAME", lib) does: dlopen, resolve NAME,
 * NAME_n_in, NAME_n_out, NAME_sparsity_in / _out, NAME_work by dlsym, size the arg / res arrays from NAME_work, call
 * NAME(arg, res, iw, w, 0).  Plain C (no engine headers): only what a CasADi host sees.
 *   casadi_external_test <lib> --patterns            no GPU: symbol table, arities, sparsity patterns
 *   casadi_external_test <lib> --golden <file>       GPU: values against golden vectors (tests/test_casadi_external.py) */
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef int (*eval_t)(const double**, double**, int*, double*, int);
typedef int (*getint_t)(void);
typedef const int* (*sparsity_t)(int);
typedef int (*work_t)(int*, int*, int*, int*);

static int failures = 0;
#define CHECK(c) do { if (!(c)) { printf("CHECK FAILED line %d: %s\n", __LINE__, #c); ++failures; } } while (0)

typedef struct { eval_t f; getint_t n_in, n_out; sparsity_t sp_in, sp_out; work_t work; } ext_fn;

static ext_fn resolve(void* lib, const char* name) {
    ext_fn e; char buf[128];
    memset(&e, 0, sizeof e);
    e.f = (eval_t)dlsym(lib, name);
    snprintf(buf, sizeof buf, "%s_n_in", name); e.n_in = (getint_t)dlsym(lib, buf);
    snprintf(buf, sizeof buf, "%s_n_out", name); e.n_out = (getint_t)dlsym(lib, buf);
    snprintf(buf, sizeof buf, "%s_sparsity_in", name); e.sp_in = (sparsity_t)dlsym(lib, buf);
    snprintf(buf, sizeof buf, "%s_sparsity_out", name); e.sp_out = (sparsity_t)dlsym(lib, buf);
    snprintf(buf, sizeof buf, "%s_work", name); e.work = (work_t)dlsym(lib, buf);
    CHECK(e.f && e.n_in && e.n_out && e.sp_in && e.sp_out && e.work);
    return e;
}
static int nnz_of(const int* sp) { return sp[2 + sp[1]]; }

static void check_dense(const int* sp, int n) {
    int i;
    CHECK(sp && sp[0] == n && sp[1] == 1 && sp[2] == 0 && sp[3] == n);
    for (i = 0; sp && i < n; ++i) CHECK(sp[4 + i] == i);
}

static double* read_vec(FILE* f, const char* tag, int* n_out) {
    char t[64]; int n, i; double* v;
    if (fscanf(f, "%63s %d", t, &n) != 2 || strcmp(t, tag) != 0) { printf("golden file: expected %s\n", tag); exit(2); }
    v = (double*)malloc(sizeof(double) * (size_t)n);
    for (i = 0; i < n; ++i) if (fscanf(f, "%lf", &v[i]) != 1) exit(2);
    *n_out = n;
    return v;
}
static int close_to(const double* a, const double* b, int n, double rtol) {
    int i;
    for (i = 0; i < n; ++i) if (fabs(a[i] - b[i]) > rtol * fmax(fabs(b[i]), 1.0)) { printf("  mismatch at %d: %.17g vs %.17g\n", i, a[i], b[i]); return 0; }
    return 1;
}

int main(int argc, char** argv) {
    void* lib; ext_fn dyn, jac, aero, rk4, dyn_id, jac_id;
    if (argc < 3) { printf("usage: %s <libkite_casadi.so> --patterns | --golden <file>\n", argv[0]); return 2; }
    lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
    if (!lib) { printf("dlopen: %s\n", dlerror()); return 2; }
    dyn = resolve(lib, "dynamics"); jac = resolve(lib, "dyn_jacobian"); aero = resolve(lib, "Aero"); rk4 = resolve(lib, "RK4");
    dyn_id = resolve(lib, "dynamics_id"); jac_id = resolve(lib, "dyn_jacobian_id");

    /* arities and work sizes (casadi::External::init) */
    { int a, r, iw, w;
      CHECK(dyn.n_in() == 2 && dyn.n_out() == 1 && jac.n_in() == 2 && jac.n_out() == 1 && aero.n_in() == 2 && rk4.n_in() == 3 && rk4.n_out() == 1);
      CHECK(dyn_id.n_in() == 3 && jac_id.n_in() == 3);
      CHECK(rk4.work(&a, &r, &iw, &w) == 0 && a == 3 && r == 1 && iw == 0 && w == 0);
      CHECK(jac.work(&a, &r, &iw, &w) == 0 && a == 2 && r == 1); }
    /* patterns */
    check_dense(dyn.sp_in(0), 13); check_dense(dyn.sp_in(1), 3); check_dense(dyn.sp_out(0), 13);
    check_dense(rk4.sp_in(2), 1); check_dense(aero.sp_out(0), 3); check_dense(dyn_id.sp_in(2), 21);
    CHECK(dyn.sp_in(2) == NULL && dyn.sp_out(1) == NULL);
    { const int* sp = jac.sp_out(0); int j, t, nnz;
      CHECK(sp && sp[0] == 13 && sp[1] == 13);
      nnz = nnz_of(sp);
      CHECK(nnz == 104);                                           /* zero tether arm */
      for (j = 0; j < 13; ++j) for (t = sp[2 + j] + 1; t < sp[2 + j + 1]; ++t) CHECK(sp[2 + 14 + t] > sp[2 + 14 + t - 1]);   /* rows ascending */
      /* spot checks of the structure: d v_dot0 / d w0 and d v_dot2 / d w2 are structural zeros; d r_dot / d r = 0 */
      { int has_0_3 = 0, has_6_6 = 0, has_9_9 = 0;
        for (j = 0; j < 13; ++j) for (t = sp[2 + j]; t < sp[2 + j + 1]; ++t) { int i = sp[2 + 14 + t]; if (i == 0 && j == 3) has_0_3 = 1; if (i == 6 && j == 6) has_6_6 = 1; if (i == 9 && j == 9) has_9_9 = 1; }
        CHECK(!has_0_3 && !has_6_6 && has_9_9); } }

    if (strcmp(argv[2], "--golden") == 0 && argc >= 4) {
        FILE* gf = fopen(argv[3], "r"); int n, k;
        double *x, *u, *f_ref, *jx_ref, *aero_ref, *xn_ref, *h, *p, *fid_ref;
        double out[169]; const double* arg[3]; double* res[1];
        if (!gf) { printf("cannot open %s\n", argv[3]); return 2; }
        x = read_vec(gf, "x", &n); u = read_vec(gf, "u", &n); f_ref = read_vec(gf, "f", &n); jx_ref = read_vec(gf, "Jx", &n);   /* Jx row-major 13x13 */
        aero_ref = read_vec(gf, "aero", &n); h = read_vec(gf, "h", &n); xn_ref = read_vec(gf, "xn", &n);
        p = read_vec(gf, "p", &n); fid_ref = read_vec(gf, "f_id", &n);
        arg[0] = x; arg[1] = u; arg[2] = h; res[0] = out;
        CHECK(dyn.f(arg, res, NULL, NULL, 0) == 0); CHECK(close_to(out, f_ref, 13, 1e-9));
        CHECK(aero.f(arg, res, NULL, NULL, 0) == 0); CHECK(close_to(out, aero_ref, 3, 1e-9));
        CHECK(rk4.f(arg, res, NULL, NULL, 0) == 0); CHECK(close_to(out, xn_ref, 13, 1e-9));
        CHECK(jac.f(arg, res, NULL, NULL, 0) == 0);
        { const int* sp = jac.sp_out(0); int j, t; double dense[169];
          memset(dense, 0, sizeof dense);
          for (j = 0; j < 13; ++j) for (t = sp[2 + j]; t < sp[2 + j + 1]; ++t) dense[sp[2 + 14 + t] * 13 + j] = out[t];
          CHECK(close_to(dense, jx_ref, 169, 1e-9)); }
        /* a null input reads as zeros (CasADi convention): dynamics(x, 0) == dynamics(x, [0,0,0]) */
        { double zero[3] = {0, 0, 0}, o2[13];
          arg[1] = NULL; CHECK(dyn.f(arg, res, NULL, NULL, 0) == 0); memcpy(o2, out, sizeof o2);
          arg[1] = zero; CHECK(dyn.f(arg, res, NULL, NULL, 0) == 0);
          for (k = 0; k < 13; ++k) CHECK(o2[k] == out[k]);
          arg[1] = u; }
        /* a null output is skipped */
        res[0] = NULL; CHECK(dyn.f(arg, res, NULL, NULL, 0) == 0); res[0] = out;
        /* identification variant */
        arg[2] = p; CHECK(dyn_id.f(arg, res, NULL, NULL, 0) == 0); CHECK(close_to(out, fid_ref, 13, 1e-9));
        CHECK(jac_id.f(arg, res, NULL, NULL, 0) == 0);
        fclose(gf);
    }
    printf("casadi_external_test: %d failures\n", failures);
    return failures ? 1 : 0;
}
