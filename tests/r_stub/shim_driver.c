/* TEST INFRASTRUCTURE: drives the .Call entry points of r/cardelino_b200_shim.c the way R would (lookup in the
 * registration table, argument lists, an error context that Rf_error longjmps to), on top of fake_r.c.
 *
 *   shim_driver errors                 CPU-safe: registration, argument-count check, released handle, and -- on a box
 *                                      without a GPU -- cdl_create's error arriving as an R error
 *   shim_driver run <in.bin> <out.txt> on a B200: load (dense and dgCMatrix), Gibbs (full list, light list, relabel,
 *                                      two chains, verbose output), error paths, a user interrupt in mid-run, EM,
 *                                      get_logLik, finalizer.  Results go to <out.txt> as "name count values..." lines
 *                                      that tests/test_r_shim.py compares with the ctypes path.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <setjmp.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

extern jmp_buf fake_r_top;
extern int fake_r_top_set, fake_r_protect_depth, fake_r_protect_max, fake_r_interrupt_at, fake_r_interrupt_polls;
extern char fake_r_last_error[1024], fake_r_stdout[1 << 16];
DL_FUNC fake_r_lookup(const char* name, int nargs);
void fake_r_run_finalizer(SEXP s);
void R_init_cardelino(DllInfo* dll);

typedef SEXP (*F1)(SEXP);
typedef SEXP (*F3)(SEXP, SEXP, SEXP);
typedef SEXP (*F4)(SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F5)(SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F7)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F8)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);

/* .Call under an error context: returns NULL and leaves the message in fake_r_last_error if the callee raised */
#define TRY(expr) (fake_r_top_set = 1, setjmp(fake_r_top) == 0 ? (expr) : (SEXP)NULL)

static FILE* out;
static void put(const char* name, const double* v, long n) {
    fprintf(out, "%s %ld", name, n);
    for (long q = 0; q < n; ++q) fprintf(out, " %.17g", v[q]);
    fprintf(out, "\n");
}
static void put_text(const char* name, const char* s) {
    fprintf(out, "%s -1 ", name);
    for (; *s; ++s) fputc(*s == '\n' ? '|' : *s, out);
    fprintf(out, "\n");
}
static SEXP get(SEXP list, const char* name) {
    SEXP names = Rf_getAttrib(list, R_NamesSymbol);
    for (int q = 0; q < Rf_length(list); ++q)
        if (!strcmp(CHAR(STRING_ELT(names, q)), name)) return VECTOR_ELT(list, q);
    fprintf(stderr, "driver: no field %s\n", name);
    exit(3);
}
static void put_field(const char* prefix, SEXP list, const char* name) {
    char key[128];
    snprintf(key, sizeof(key), "%s.%s", prefix, name);
    SEXP v = get(list, name);
    if (Rf_isNull(v)) { fprintf(out, "%s 0\n", key); return; }
    if (Rf_length(v) && ((int*)0 == 0) && 1) {
        /* INTSXP scalars (n_iter, stop_rule) are written as doubles */
        double tmp;
        if (Rf_length(v) == 1 && (strcmp(name, "n_iter") == 0 || strcmp(name, "stop_rule") == 0)) {
            tmp = (double)Rf_asInteger(v);
            put(key, &tmp, 1);
            return;
        }
    }
    put(key, REAL(v), Rf_length(v));
}

static SEXP mk_params(int n, const char** names, SEXP* vals) {
    const char* nm[32];
    for (int q = 0; q < n; ++q) nm[q] = names[q];
    nm[n] = "";
    SEXP p = Rf_mkNamed(VECSXP, nm);
    for (int q = 0; q < n; ++q) SET_VECTOR_ELT(p, q, vals[q]);
    return p;
}
static SEXP str1(const char* s) {
    SEXP v = Rf_allocVector(STRSXP, 1);
    SET_STRING_ELT(v, 0, Rf_mkChar(s));
    return v;
}
static SEXP real2(double a, double b) { SEXP v = Rf_allocVector(REALSXP, 2); REAL(v)[0] = a; REAL(v)[1] = b; return v; }

static int mode_errors(void) {
    int fails = 0;
    R_init_cardelino(NULL);
    /* argument-count check of the registration table */
    if (TRY((SEXP)fake_r_lookup("b200_gibbs", 3)) != NULL || !strstr(fake_r_last_error, "Incorrect number of arguments")) ++fails;
    if (TRY((SEXP)fake_r_lookup("b200_nope", 1)) != NULL || !strstr(fake_r_last_error, "not in load table")) ++fails;
    printf("lookup checks: %s\n", fails ? "FAIL" : "ok");
    /* a released handle */
    SEXP dead = R_MakeExternalPtr(NULL, R_NilValue, R_NilValue);
    SEXP m = Rf_allocMatrix(REALSXP, 2, 2);
    F3 load = (F3)fake_r_lookup("b200_load_dense", 3);
    if (TRY(load(dead, m, m)) != NULL || !strstr(fake_r_last_error, "handle already released")) { ++fails; printf("released-handle check: FAIL\n"); }
    else printf("released-handle check: ok (%s)\n", fake_r_last_error);
    /* cdl_create: an external pointer on a GPU box, an R error (not a crash, not a CPU fallback) elsewhere */
    F1 create = (F1)fake_r_lookup("b200_create", 1);
    SEXP h = TRY(create(Rf_ScalarInteger(0)));
    if (h) { printf("b200_create: handle created (GPU present)\n"); fake_r_run_finalizer(h); }
    else if (strstr(fake_r_last_error, "CUDA") || strstr(fake_r_last_error, "device")) printf("b200_create: R error as expected without a GPU: %s\n", fake_r_last_error);
    else { ++fails; printf("b200_create: unexpected failure: %s\n", fake_r_last_error); }
    if (fake_r_protect_depth != 0) { ++fails; printf("protect stack imbalance: %d\n", fake_r_protect_depth); }
    printf(fails ? "ERRORS MODE FAILED\n" : "ERRORS MODE OK\n");
    return fails ? 1 : 0;
}

int main(int argc, char** argv) {
    if (argc >= 2 && !strcmp(argv[1], "errors")) return mode_errors();
    if (argc < 4 || strcmp(argv[1], "run")) { fprintf(stderr, "usage: shim_driver errors | run <in.bin> <out.txt>\n"); return 2; }
    FILE* in = fopen(argv[2], "rb");
    out = fopen(argv[3], "w");
    if (!in || !out) { perror("open"); return 2; }
    int64_t hdr[6];                          /* N, M, K, min_iter, max_iter, seed */
    if (fread(hdr, sizeof(int64_t), 6, in) != 6) return 2;
    const int N = (int)hdr[0], M = (int)hdr[1], K = (int)hdr[2];
    SEXP A = Rf_allocMatrix(REALSXP, N, M), D = Rf_allocMatrix(REALSXP, N, M), Config = Rf_allocMatrix(REALSXP, N, K);
    if (fread(REAL(A), 8, (size_t)N * M, in) != (size_t)N * M || fread(REAL(D), 8, (size_t)N * M, in) != (size_t)N * M ||
        fread(REAL(Config), 8, (size_t)N * K, in) != (size_t)N * K) return 2;
    fclose(in);
    R_init_cardelino(NULL);
    F1 create = (F1)fake_r_lookup("b200_create", 1);
    F3 load = (F3)fake_r_lookup("b200_load_dense", 3);
    F8 load_dgc = (F8)fake_r_lookup("b200_load_dgc", 8);
    F4 gibbs = (F4)fake_r_lookup("b200_gibbs", 4);
    F7 em = (F7)fake_r_lookup("b200_em", 7);
    F5 loglik = (F5)fake_r_lookup("b200_get_loglik", 5);

    SEXP h = TRY(create(Rf_ScalarInteger(0)));
    if (!h) { fprintf(stderr, "create failed: %s\n", fake_r_last_error); return 1; }
    if (!TRY(load(h, A, D))) { fprintf(stderr, "load failed: %s\n", fake_r_last_error); return 1; }

    const char* names[] = {"relax_Config", "relax_rate_fixed", "relax_rate_prior", "keep_base_clone", "prior0", "prior1",
                           "min_iter", "max_iter", "buin_frac", "wise", "n_chain", "seed", "relabel", "verbose", "light"};
    SEXP vals[15] = {Rf_ScalarLogical(1), R_NilValue, real2(1, 9), Rf_ScalarLogical(1), real2(0.2, 99.8), real2(0.45, 0.55),
                     Rf_ScalarInteger((int)hdr[3]), Rf_ScalarInteger((int)hdr[4]), Rf_ScalarReal(0.5), str1("variant"),
                     Rf_ScalarInteger(2), Rf_ScalarReal((double)hdr[5]), Rf_ScalarLogical(0), Rf_ScalarLogical(1),
                     Rf_ScalarLogical(0)};
    const char* fields[] = {"theta0", "theta1_elements", "theta0_all", "theta1_all", "logLik", "prob_all_t", "prob", "prob_variant",
                            "relax_rate", "Config_prob", "Config_all_t", "relax_rate_all", "DIC", "n_iter", "percent_converged",
                            "stop_rule", "logLik_post"};
    /* 1. the full list, two chains, verbose */
    fake_r_stdout[0] = 0;
    SEXP res = TRY(gibbs(h, Config, R_NilValue, mk_params(15, names, vals)));
    if (!res) { fprintf(stderr, "gibbs failed: %s\n", fake_r_last_error); return 1; }
    for (int c = 0; c < 2; ++c) {
        char pre[16];
        snprintf(pre, sizeof(pre), "full%d", c);
        for (unsigned q = 0; q < sizeof(fields) / sizeof(fields[0]); ++q) put_field(pre, VECTOR_ELT(res, c), fields[q]);
    }
    put_text("full.stdout", fake_r_stdout);
    /* 2. light list */
    vals[10] = Rf_ScalarInteger(1); vals[13] = Rf_ScalarLogical(0); vals[14] = Rf_ScalarLogical(1);
    fake_r_stdout[0] = 0;
    res = TRY(gibbs(h, Config, R_NilValue, mk_params(15, names, vals)));
    if (!res) { fprintf(stderr, "light gibbs failed: %s\n", fake_r_last_error); return 1; }
    for (unsigned q = 0; q < sizeof(fields) / sizeof(fields[0]); ++q) put_field("light", VECTOR_ELT(res, 0), fields[q]);
    put_text("light.stdout", fake_r_stdout);
    /* 3. relabel */
    vals[12] = Rf_ScalarLogical(1); vals[14] = Rf_ScalarLogical(0);
    res = TRY(gibbs(h, Config, R_NilValue, mk_params(15, names, vals)));
    if (!res) { fprintf(stderr, "relabel gibbs failed: %s\n", fake_r_last_error); return 1; }
    for (unsigned q = 0; q < sizeof(fields) / sizeof(fields[0]); ++q) put_field("relabel", VECTOR_ELT(res, 0), fields[q]);
    /* 4. error paths: each must come back as an R error with the reference's wording and leave the handle usable */
    vals[14] = Rf_ScalarLogical(1);                       /* relabel + light */
    put_text("err.relabel_light", TRY(gibbs(h, Config, R_NilValue, mk_params(15, names, vals))) ? "NO ERROR" : fake_r_last_error);
    vals[12] = Rf_ScalarLogical(0); vals[14] = Rf_ScalarLogical(0);
    vals[9] = str1("foo");
    put_text("err.wise", TRY(gibbs(h, Config, R_NilValue, mk_params(15, names, vals))) ? "NO ERROR" : fake_r_last_error);
    vals[9] = str1("variant");
    vals[1] = Rf_ScalarReal(1.5);
    put_text("err.relax_rate_fixed", TRY(gibbs(h, Config, R_NilValue, mk_params(15, names, vals))) ? "NO ERROR" : fake_r_last_error);
    vals[1] = R_NilValue;
    { SEXP bad = Rf_allocMatrix(REALSXP, N + 3, 2);       /* prior1 with the wrong number of rows */
      for (int q = 0; q < 2 * (N + 3); ++q) REAL(bad)[q] = 0.5;
      SEXP keep = vals[5]; vals[5] = bad;
      put_text("err.prior1_rows", TRY(gibbs(h, Config, R_NilValue, mk_params(15, names, vals))) ? "NO ERROR" : fake_r_last_error);
      vals[5] = keep; }
    { const char* n2[] = {"min_iter", "bogus"}; SEXP v2[] = {Rf_ScalarInteger(10), Rf_ScalarInteger(1)};
      put_text("err.unknown_param", TRY(gibbs(h, Config, R_NilValue, mk_params(2, n2, v2))) ? "NO ERROR" : fake_r_last_error); }
    /* 5. a user interrupt in mid-run: the third poll fires */
    vals[6] = Rf_ScalarInteger(4000); vals[7] = Rf_ScalarInteger(4000);
    fake_r_interrupt_polls = 0; fake_r_interrupt_at = 3;
    put_text("err.interrupt", TRY(gibbs(h, Config, R_NilValue, mk_params(15, names, vals))) ? "NO ERROR" : fake_r_last_error);
    { double polls = fake_r_interrupt_polls; put("interrupt.polls", &polls, 1); }
    vals[6] = Rf_ScalarInteger((int)hdr[3]); vals[7] = Rf_ScalarInteger((int)hdr[4]);
    /* 6. EM and get_logLik on the same handle (still usable after the errors and the interrupt) */
    res = TRY(em(h, Config, R_NilValue, Rf_ScalarInteger(10), Rf_ScalarInteger(1000), Rf_ScalarReal(1e-5), real2(0.1, 0.5)));
    if (!res) { fprintf(stderr, "em failed: %s\n", fake_r_last_error); return 1; }
    put("em.theta", REAL(get(res, "theta")), 2);
    put("em.prob", REAL(get(res, "prob")), (long)M * K);
    put("em.logLik", REAL(get(res, "logLik")), 1);
    { double t = Rf_asInteger(get(res, "n_iter")); put("em.n_iter", &t, 1); }
    { SEXP lab = Rf_allocVector(INTSXP, M);
      for (int j = 0; j < M; ++j) INTEGER(lab)[j] = 1 + j % K;
      SEXP th1 = Rf_allocVector(REALSXP, N);
      for (int i = 0; i < N; ++i) REAL(th1)[i] = 0.3 + 0.4 * i / N;
      SEXP v = TRY(loglik(h, Config, lab, Rf_ScalarReal(0.02), th1));
      if (!v) { fprintf(stderr, "get_loglik failed: %s\n", fake_r_last_error); return 1; }
      put("loglik.labels", REAL(v), 1);
      v = TRY(loglik(h, Config, get(res, "prob"), Rf_ScalarReal(0.02), th1));
      if (!v) { fprintf(stderr, "get_loglik(prob) failed: %s\n", fake_r_last_error); return 1; }
      put("loglik.prob", REAL(v), 1); }
    /* 7. the dgCMatrix loader: D with explicit zeros dropped, A with its own pattern */
    { int nd = 0, na = 0;
      SEXP Dp = Rf_allocVector(INTSXP, M + 1), Ap = Rf_allocVector(INTSXP, M + 1);
      for (long q = 0; q < (long)N * M; ++q) { const double d = REAL(D)[q], a = REAL(A)[q]; nd += (d == d && d > 0); na += (a == a && a > 0); }
      SEXP Di = Rf_allocVector(INTSXP, nd), Dx = Rf_allocVector(REALSXP, nd), Ai = Rf_allocVector(INTSXP, na), Ax = Rf_allocVector(REALSXP, na);
      nd = na = 0;
      for (int j = 0; j < M; ++j) {
          INTEGER(Dp)[j] = nd; INTEGER(Ap)[j] = na;
          for (int i = 0; i < N; ++i) {
              const double d = REAL(D)[i + (long)j * N], a = REAL(A)[i + (long)j * N];
              if (d == d && d > 0) { INTEGER(Di)[nd] = i; REAL(Dx)[nd++] = d; }
              if (a == a && a > 0) { INTEGER(Ai)[na] = i; REAL(Ax)[na++] = a; }
          }
      }
      INTEGER(Dp)[M] = nd; INTEGER(Ap)[M] = na;
      SEXP dims = Rf_allocVector(INTSXP, 2); INTEGER(dims)[0] = N; INTEGER(dims)[1] = M;
      if (!TRY(load_dgc(h, dims, Dp, Di, Dx, Ap, Ai, Ax))) { fprintf(stderr, "load_dgc failed: %s\n", fake_r_last_error); return 1; }
      vals[10] = Rf_ScalarInteger(1); vals[14] = Rf_ScalarLogical(1);
      res = TRY(gibbs(h, Config, R_NilValue, mk_params(15, names, vals)));
      if (!res) { fprintf(stderr, "gibbs after load_dgc failed: %s\n", fake_r_last_error); return 1; }
      put_field("dgc", VECTOR_ELT(res, 0), "prob"); put_field("dgc", VECTOR_ELT(res, 0), "logLik"); }
    /* 8. finalizer, then a call on the released handle */
    fake_r_run_finalizer(h);
    put_text("err.released", TRY(load(h, A, D)) ? "NO ERROR" : fake_r_last_error);
    { double d[2] = {(double)fake_r_protect_depth, (double)fake_r_protect_max}; put("protect", d, 2); }
    fclose(out);
    printf("RUN MODE OK\n");
    return 0;
}
