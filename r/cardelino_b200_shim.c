/*
 * .Call shim between R and libcardelino_b200.so (include/cardelino_b200.h).
 *
 * NOT built against R in the build image (no R there); it is compiled and EXECUTED in the test-suite against a
 * small stand-in for R's C API (tests/r_stub/fake_r.c, tests/test_r_shim.py), and shipped as source for a maintainer
 * who has R.  Build inside the cardelino package tree:
 *     src/cardelino_b200_shim.c   (this file)
 *     src/Makevars:  PKG_CPPFLAGS = -I<repo>/include
 *                    PKG_LIBS     = -L<repo>/cardelino_b200 -lcardelino_b200 -Wl,-rpath,<repo>/cardelino_b200
 *     NAMESPACE:     useDynLib(cardelino, .registration = TRUE)
 * and replace the bodies of clone_id_Gibbs / clone_id_EM (R/clone_id.R:351-675, :240-326) with
 * r/clone_id_b200.R.
 *
 * Rules followed here: every SEXP is PROTECTed while allocating; the handle is an
 * external pointer with a finalizer; errors are raised with Rf_error only AFTER the failing call has returned
 * (the library never longjmps and never calls R from another thread).  The library calls back on R's main
 * thread every check_every sweeps (cdl_progress_fn): the shim prints the reference's "x% converged." line with
 * Rprintf when verbose, and polls for a user interrupt WITHOUT a longjmp (R_ToplevelExec around
 * R_CheckUserInterrupt); an interrupt makes the library stop cleanly (CDL_ERR_INTERRUPTED) and is then re-raised.
 * doMC/fork must not be used once a handle exists (CUDA contexts do not survive fork): n_chain is passed down
 * and the chains run inside the library.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <math.h>
#include <string.h>
#include "cardelino_b200.h"

static void handle_finalizer(SEXP ptr) {
    cdl_handle h = (cdl_handle)R_ExternalPtrAddr(ptr);
    if (h) { cdl_destroy(h); R_ClearExternalPtr(ptr); }
}

static cdl_handle get_handle(SEXP ptr) {
    cdl_handle h = (cdl_handle)R_ExternalPtrAddr(ptr);
    if (!h) Rf_error("cardelino_b200: handle already released");
    return h;
}

static void check(cdl_handle h, int rc) {
    if (rc != CDL_OK) {
        char msg[1024];
        strncpy(msg, cdl_last_error(h), sizeof(msg) - 1);
        msg[sizeof(msg) - 1] = 0;
        Rf_error("%s", msg);          /* same wording as the reference's stop() where one exists */
    }
}

SEXP b200_create(SEXP device) {
    cdl_handle h = NULL;
    int rc = cdl_create(Rf_asInteger(device), &h);
    if (rc != CDL_OK) Rf_error("%s", cdl_last_error(NULL));
    SEXP ptr = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ptr, handle_finalizer, TRUE);
    UNPROTECT(1);
    return ptr;
}

/* A, D: dense double matrices (variants x cells), NA_real_ allowed -- replaces as.matrix() + the NA
 * normalisation of R/clone_id.R:121-122,380-391 */
SEXP b200_load_dense(SEXP ptr, SEXP A, SEXP D) {
    cdl_handle h = get_handle(ptr);
    SEXP dim = Rf_getAttrib(D, R_DimSymbol);
    check(h, cdl_load_dense(h, INTEGER(dim)[0], INTEGER(dim)[1], REAL(A), REAL(D)));
    return R_NilValue;
}

/* dgCMatrix slots (load_cellSNP_vcf output, R/read_data.R:293-294): p is integer in R -> widen */
SEXP b200_load_dgc(SEXP ptr, SEXP dims, SEXP Dp, SEXP Di, SEXP Dx, SEXP Ap, SEXP Ai, SEXP Ax) {
    cdl_handle h = get_handle(ptr);
    int64_t N = INTEGER(dims)[0], M = INTEGER(dims)[1];
    int64_t* dp = (int64_t*)R_alloc(M + 1, sizeof(int64_t));
    int64_t* ap = (int64_t*)R_alloc(M + 1, sizeof(int64_t));
    for (int64_t j = 0; j <= M; ++j) { dp[j] = INTEGER(Dp)[j]; ap[j] = INTEGER(Ap)[j]; }
    check(h, cdl_load_csc(h, N, M, dp, INTEGER(Di), REAL(Dx), ap, INTEGER(Ai), REAL(Ax)));
    return R_NilValue;
}

/* ---- progress callback: printing and interrupts, both on R's main thread ---- */
typedef struct { int verbose; int interrupted; } progress_state;

static void poll_interrupt(void* unused) { (void)unused; R_CheckUserInterrupt(); }

static int on_progress(void* user, int32_t chain, int32_t iteration, double fraction) {
    progress_state* s = (progress_state*)user;
    (void)chain; (void)iteration;
    if (s->verbose && fraction == fraction)                     /* cat(round(mean(Converged_all), 3) * 100, "% converged.\n") */
        Rprintf("%g%% converged.\n", floor(fraction * 1000.0 + 0.5) / 1000.0 * 100.0);
    if (!R_ToplevelExec(poll_interrupt, NULL)) {                /* FALSE: the callee was interrupted (no longjmp escapes) */
        s->interrupted = 1;
        return 1;
    }
    return 0;
}

/* params: named list mirroring clone_id_Gibbs's arguments (+ n_chain, seed, light); returns one list per chain
 * with the fields of the reference's return list (R/clone_id.R:662-673).
 *   light = FALSE  the full list: prob_all (kept on the device, so the reference's stop rule and relabel are exact),
 *                  Config_all, dense prob_variant
 *   light = TRUE   large donors: no prob_all / Config_all / prob_variant (an N x M or max_iter x M*K double array
 *                  would not fit R's memory: 80 GB and 2.5 PB at the BASELINE C4 shape); the stop rule then runs from
 *                  running sums (cdl_gibbs_info.stop_rule says how) -- a documented deviation of the return list */
SEXP b200_gibbs(SEXP ptr, SEXP Config, SEXP Psi, SEXP params) {
    cdl_handle h = get_handle(ptr);
    cdl_gibbs_params p;
    cdl_gibbs_defaults(&p);
    int light = 0;
    progress_state ps = {0, 0};
    SEXP names = Rf_getAttrib(params, R_NamesSymbol);
    for (int q = 0; q < Rf_length(params); ++q) {
        const char* nm = CHAR(STRING_ELT(names, q));
        SEXP v = VECTOR_ELT(params, q);
        if (!strcmp(nm, "relax_Config")) p.relax_Config = Rf_asLogical(v);
        else if (!strcmp(nm, "relax_rate_fixed")) {
            p.relax_rate_fixed_set = !Rf_isNull(v);
            if (p.relax_rate_fixed_set) p.relax_rate_fixed = Rf_asReal(v);
        } else if (!strcmp(nm, "relax_rate_prior")) { p.relax_rate_prior[0] = REAL(v)[0]; p.relax_rate_prior[1] = REAL(v)[1]; }
        else if (!strcmp(nm, "keep_base_clone")) p.keep_base_clone = Rf_asLogical(v);
        else if (!strcmp(nm, "prior0")) { p.prior0[0] = REAL(v)[0]; p.prior0[1] = REAL(v)[1]; }
        else if (!strcmp(nm, "prior1")) {
            if (Rf_isMatrix(v)) {                /* n_element x 2 (R/clone_id.R:418-424); the library checks the row count */
                SEXP d = Rf_getAttrib(v, R_DimSymbol);
                if (INTEGER(d)[1] != 2) Rf_error("prior1 need to be a matrix of n_element x 2");
                p.prior1_matrix = REAL(v);
                p.prior1_matrix_rows = INTEGER(d)[0];
            } else {
                if (Rf_length(v) != 2) Rf_error("prior1 need to be a matrix of n_element x 2");
                p.prior1[0] = REAL(v)[0]; p.prior1[1] = REAL(v)[1];
            }
        } else if (!strcmp(nm, "min_iter")) p.min_iter = Rf_asInteger(v);
        else if (!strcmp(nm, "max_iter")) p.max_iter = Rf_asInteger(v);
        else if (!strcmp(nm, "buin_frac")) p.buin_frac = Rf_asReal(v);
        else if (!strcmp(nm, "wise")) {
            const char* w = CHAR(STRING_ELT(v, 0));
            if (!strcmp(w, "global")) p.wise = CDL_WISE_GLOBAL;
            else if (!strcmp(w, "variant")) p.wise = CDL_WISE_VARIANT;
            else if (!strcmp(w, "element")) p.wise = CDL_WISE_ELEMENT;
            else Rf_error("Input wise mode: %s, while only supporting: element, variant, global", w);
        } else if (!strcmp(nm, "n_chain")) p.n_chain = Rf_asInteger(v);
        else if (!strcmp(nm, "seed")) p.seed = (uint64_t)Rf_asReal(v);     /* wrapper: sample.int(.Machine$integer.max, 1) */
        else if (!strcmp(nm, "relabel")) p.relabel = Rf_asLogical(v) ? 1 : 0;
        else if (!strcmp(nm, "verbose")) ps.verbose = Rf_asLogical(v) ? 1 : 0;
        else if (!strcmp(nm, "light")) light = Rf_asLogical(v) ? 1 : 0;
        else Rf_error("cardelino_b200: unknown parameter '%s'", nm);
    }
    if (light && p.relabel) Rf_error("relabel needs the prob_all trace: use light = FALSE");
    p.keep_prob_all = light ? 0 : 1;      /* the reference returns prob_all; large donors cannot */
    p.progress = on_progress;
    p.progress_user = &ps;
    SEXP cdim = Rf_getAttrib(Config, R_DimSymbol);
    const int N = INTEGER(cdim)[0], K = INTEGER(cdim)[1];
    const int rc = cdl_gibbs_run(h, REAL(Config), K, Rf_isNull(Psi) ? NULL : REAL(Psi), &p, NULL);
    if (rc == CDL_ERR_INTERRUPTED && ps.interrupted)      /* raised only now that the library call has returned; */
        Rf_error("interrupted by the user");              /* the handle stays valid for the next call             */
    check(h, rc);

    int64_t Nn, M, nnz; double W;
    check(h, cdl_data_info(h, &Nn, &M, &nnz, &W));
    SEXP chains = PROTECT(Rf_allocVector(VECSXP, p.n_chain));
    for (int c = 0; c < p.n_chain; ++c) {
        cdl_gibbs_info info;
        check(h, cdl_gibbs_info_get(h, c, &info));
        const int T = info.n_iter, ne = info.n_element;
        const char* fields[] = {"theta0", "theta1_elements", "theta0_all", "theta1_all", "logLik", "prob_all_t", "prob",
                                "prob_variant", "relax_rate", "Config_prob", "Config_all_t", "relax_rate_all", "DIC",
                                "n_iter", "percent_converged", "element", "stop_rule", "logLik_post", ""};
        SEXP out = PROTECT(Rf_mkNamed(VECSXP, fields));
        int np = 1;
        SEXP theta0 = PROTECT(Rf_allocVector(REALSXP, 1)); ++np;
        SEXP th1 = PROTECT(Rf_allocVector(REALSXP, ne)); ++np;
        check(h, cdl_fetch_theta(h, c, REAL(theta0), REAL(th1)));
        SEXP t0all = PROTECT(Rf_allocMatrix(REALSXP, T, 1)); ++np;
        SEXP t1all = PROTECT(Rf_allocMatrix(REALSXP, T, ne)); ++np;
        check(h, cdl_fetch_theta_traces(h, c, REAL(t0all), REAL(t1all)));
        SEXP ll = PROTECT(Rf_allocVector(REALSXP, T)); ++np;
        check(h, cdl_fetch_loglik_trace(h, c, REAL(ll)));
        SEXP prob = PROTECT(Rf_allocMatrix(REALSXP, (int)M, K)); ++np;
        check(h, cdl_fetch_prob(h, c, REAL(prob)));
        SEXP rr = PROTECT(Rf_allocVector(REALSXP, 1)); ++np;
        SEXP rrall = PROTECT(Rf_allocVector(REALSXP, T)); ++np;
        check(h, cdl_fetch_relax_rate(h, c, REAL(rr), REAL(rrall)));
        SEXP cprob = PROTECT(Rf_allocMatrix(REALSXP, N, K)); ++np;
        check(h, cdl_fetch_config_prob(h, c, REAL(cprob)));
        SEXP dic = PROTECT(Rf_allocVector(REALSXP, 6)); ++np;
        memcpy(REAL(dic), info.dic, sizeof(double) * 6);
        SEXP pall = R_NilValue, pv = R_NilValue, call = R_NilValue;
        if (!light) {
            /* big traces come back row-major (one iteration contiguous): allocate transposed, t() in R */
            pall = PROTECT(Rf_allocMatrix(REALSXP, (int)(M * K), T)); ++np;
            check(h, cdl_fetch_prob_all(h, c, REAL(pall)));
            pv = PROTECT(Rf_allocMatrix(REALSXP, N, (int)M)); ++np;
            check(h, cdl_fetch_prob_variant(h, c, REAL(pv)));
            call = PROTECT(Rf_allocMatrix(REALSXP, N * K, T)); ++np;
            check(h, cdl_fetch_config_all(h, c, REAL(call)));
        }
        /* wise = "element": 1-based positions of the covered entries, which(!is.na(D)) (R/clone_id.R:413-415) */
        SEXP element = PROTECT(Rf_allocVector(REALSXP, p.wise == CDL_WISE_ELEMENT ? (R_xlen_t)nnz : 0)); ++np;
        if (p.wise == CDL_WISE_ELEMENT) {
            int64_t* el = (int64_t*)R_alloc((size_t)nnz, sizeof(int64_t));
            check(h, cdl_fetch_element_index(h, el));
            for (int64_t q = 0; q < nnz; ++q) REAL(element)[q] = (double)(el[q] + 1);
        }
        SEXP n_iter = PROTECT(Rf_ScalarInteger(T)); ++np;
        SEXP pc = PROTECT(Rf_ScalarReal(info.percent_converged)); ++np;
        SEXP rule = PROTECT(Rf_ScalarInteger(info.stop_rule)); ++np;
        SEXP llp = PROTECT(Rf_ScalarReal(info.logLik_post)); ++np;
        SEXP objs[] = {theta0, th1, t0all, t1all, ll, pall, prob, pv, rr, cprob, call, rrall, dic, n_iter, pc, element, rule, llp};
        for (int q = 0; q < 18; ++q) SET_VECTOR_ELT(out, q, objs[q]);
        SET_VECTOR_ELT(chains, c, out);
        UNPROTECT(np);
    }
    UNPROTECT(1);
    return chains;
}

SEXP b200_em(SEXP ptr, SEXP Config, SEXP Psi, SEXP min_iter, SEXP max_iter, SEXP thr, SEXP theta_init) {
    cdl_handle h = get_handle(ptr);
    SEXP cdim = Rf_getAttrib(Config, R_DimSymbol);
    const int K = INTEGER(cdim)[1];
    int64_t N, M, nnz; double W;
    check(h, cdl_data_info(h, &N, &M, &nnz, &W));
    const char* fields[] = {"theta", "prob", "logLik", "n_iter", ""};
    SEXP out = PROTECT(Rf_mkNamed(VECSXP, fields));
    SEXP theta = PROTECT(Rf_allocVector(REALSXP, 2));
    SEXP prob = PROTECT(Rf_allocMatrix(REALSXP, (int)M, K));
    double ll = 0; int nit = 0;
    check(h, cdl_em_run(h, REAL(Config), K, Rf_isNull(Psi) ? NULL : REAL(Psi), Rf_asInteger(min_iter),
                        Rf_asInteger(max_iter), Rf_asReal(thr), REAL(theta_init) /* runif() drawn in R, :278 */, 0,
                        REAL(theta), REAL(prob), &ll, &nit));
    SET_VECTOR_ELT(out, 0, theta);
    SET_VECTOR_ELT(out, 1, prob);
    SET_VECTOR_ELT(out, 2, Rf_ScalarReal(ll));
    SET_VECTOR_ELT(out, 3, Rf_ScalarInteger(nit));
    UNPROTECT(3);
    return out;
}

SEXP b200_get_loglik(SEXP ptr, SEXP Config, SEXP Assign, SEXP theta0, SEXP theta1) {
    cdl_handle h = get_handle(ptr);
    SEXP cdim = Rf_getAttrib(Config, R_DimSymbol);
    double out = 0;
    if (Rf_isMatrix(Assign))
        check(h, cdl_get_loglik(h, REAL(Config), INTEGER(cdim)[1], NULL, REAL(Assign), Rf_asReal(theta0), REAL(theta1),
                                Rf_length(theta1), &out));
    else
        check(h, cdl_get_loglik(h, REAL(Config), INTEGER(cdim)[1], INTEGER(Assign), NULL, Rf_asReal(theta0), REAL(theta1),
                                Rf_length(theta1), &out));
    return Rf_ScalarReal(out);
}

static const R_CallMethodDef call_methods[] = {
    {"b200_create", (DL_FUNC)&b200_create, 1},       {"b200_load_dense", (DL_FUNC)&b200_load_dense, 3},
    {"b200_load_dgc", (DL_FUNC)&b200_load_dgc, 8},   {"b200_gibbs", (DL_FUNC)&b200_gibbs, 4},
    {"b200_em", (DL_FUNC)&b200_em, 7},               {"b200_get_loglik", (DL_FUNC)&b200_get_loglik, 5},
    {NULL, NULL, 0}};

void R_init_cardelino(DllInfo* dll) {
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
