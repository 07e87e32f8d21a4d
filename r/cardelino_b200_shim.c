/*
 * .Call shim between R and libcardelino_b200.so (include/cardelino_b200.h).
 *
 * NOT compiled in the build image (no R headers there); shipped as source for a maintainer who has R.
 * Build inside the cardelino package tree:
 *     src/cardelino_b200_shim.c   (this file)
 *     src/Makevars:  PKG_CPPFLAGS = -I<repo>/include
 *                    PKG_LIBS     = -L<repo>/cardelino_b200 -lcardelino_b200 -Wl,-rpath,<repo>/cardelino_b200
 *     NAMESPACE:     useDynLib(cardelino, .registration = TRUE)
 * and replace the bodies of clone_id_Gibbs / clone_id_EM (R/clone_id.R:351-675, :240-326) with
 * r/clone_id_b200.R.
 *
 * Rules followed here: every SEXP is PROTECTed while allocating; the handle is an
 * external pointer with a finalizer; errors are raised with Rf_error only AFTER library resources of the
 * failing call are released (the library never longjmps); all calls arrive on R's main thread and the
 * library never calls back into R.  doMC/fork must not be used once a handle exists (CUDA contexts do not
 * survive fork): n_chain is passed down and the chains run inside the library.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
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

/* params: named list mirroring clone_id_Gibbs's arguments; returns the reference's return list */
SEXP b200_gibbs(SEXP ptr, SEXP Config, SEXP Psi, SEXP params) {
    cdl_handle h = get_handle(ptr);
    cdl_gibbs_params p;
    cdl_gibbs_defaults(&p);
#define GETI(name) do { SEXP v = Rf_getAttrib(params, Rf_install(#name)); (void)v; } while (0)
    /* list access helper */
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
            if (Rf_isMatrix(v)) p.prior1_matrix = REAL(v);
            else { p.prior1[0] = REAL(v)[0]; p.prior1[1] = REAL(v)[1]; }
        } else if (!strcmp(nm, "min_iter")) p.min_iter = Rf_asInteger(v);
        else if (!strcmp(nm, "max_iter")) p.max_iter = Rf_asInteger(v);
        else if (!strcmp(nm, "buin_frac")) p.buin_frac = Rf_asReal(v);
        else if (!strcmp(nm, "wise")) {
            const char* w = CHAR(STRING_ELT(v, 0));
            p.wise = !strcmp(w, "global") ? CDL_WISE_GLOBAL : !strcmp(w, "variant") ? CDL_WISE_VARIANT : CDL_WISE_ELEMENT;
        } else if (!strcmp(nm, "n_chain")) p.n_chain = Rf_asInteger(v);
        else if (!strcmp(nm, "seed")) p.seed = (uint64_t)Rf_asReal(v);     /* wrapper: sample.int(.Machine$integer.max, 1) */
    }
    p.keep_prob_all = 1;      /* the reference returns prob_all */
    SEXP cdim = Rf_getAttrib(Config, R_DimSymbol);
    const int N = INTEGER(cdim)[0], K = INTEGER(cdim)[1];
    check(h, cdl_gibbs_run(h, REAL(Config), K, Rf_isNull(Psi) ? NULL : REAL(Psi), &p, NULL));

    int64_t Nn, M, nnz; double W;
    check(h, cdl_data_info(h, &Nn, &M, &nnz, &W));
    SEXP chains = PROTECT(Rf_allocVector(VECSXP, p.n_chain));
    for (int c = 0; c < p.n_chain; ++c) {
        cdl_gibbs_info info;
        check(h, cdl_gibbs_info_get(h, c, &info));
        const int T = info.n_iter, ne = info.n_element;
        const char* fields[] = {"theta0", "theta1_elements", "theta0_all", "theta1_all", "logLik", "prob_all_t", "prob",
                                "prob_variant", "relax_rate", "Config_prob", "Config_all_t", "relax_rate_all", "DIC",
                                "n_iter", "percent_converged", "element", ""};
        SEXP out = PROTECT(Rf_mkNamed(VECSXP, fields));
        SEXP theta0 = PROTECT(Rf_allocVector(REALSXP, 1));
        SEXP th1 = PROTECT(Rf_allocVector(REALSXP, ne));
        check(h, cdl_fetch_theta(h, c, REAL(theta0), REAL(th1)));
        SEXP t0all = PROTECT(Rf_allocMatrix(REALSXP, T, 1));
        SEXP t1all = PROTECT(Rf_allocMatrix(REALSXP, T, ne));
        check(h, cdl_fetch_theta_traces(h, c, REAL(t0all), REAL(t1all)));
        SEXP ll = PROTECT(Rf_allocVector(REALSXP, T));
        check(h, cdl_fetch_loglik_trace(h, c, REAL(ll)));
        /* big traces come back row-major (one iteration contiguous): allocate transposed, t() in R */
        SEXP pall = PROTECT(Rf_allocMatrix(REALSXP, (int)(M * K), T));
        check(h, cdl_fetch_prob_all(h, c, REAL(pall)));
        SEXP prob = PROTECT(Rf_allocMatrix(REALSXP, (int)M, K));
        check(h, cdl_fetch_prob(h, c, REAL(prob)));
        SEXP pv = PROTECT(Rf_allocMatrix(REALSXP, N, (int)M));
        check(h, cdl_fetch_prob_variant(h, c, REAL(pv)));
        SEXP rr = PROTECT(Rf_allocVector(REALSXP, 1));
        SEXP rrall = PROTECT(Rf_allocVector(REALSXP, T));
        check(h, cdl_fetch_relax_rate(h, c, REAL(rr), REAL(rrall)));
        SEXP cprob = PROTECT(Rf_allocMatrix(REALSXP, N, K));
        check(h, cdl_fetch_config_prob(h, c, REAL(cprob)));
        SEXP call = PROTECT(Rf_allocMatrix(REALSXP, N * K, T));
        check(h, cdl_fetch_config_all(h, c, REAL(call)));
        SEXP dic = PROTECT(Rf_allocVector(REALSXP, 6));
        memcpy(REAL(dic), info.dic, sizeof(double) * 6);
        /* wise = "element": 1-based positions of the covered entries, which(!is.na(D)) (R/clone_id.R:413-415) */
        SEXP element = PROTECT(Rf_allocVector(REALSXP, p.wise == CDL_WISE_ELEMENT ? (R_xlen_t)nnz : 0));
        if (p.wise == CDL_WISE_ELEMENT) {
            int64_t* el = (int64_t*)R_alloc((size_t)nnz, sizeof(int64_t));
            check(h, cdl_fetch_element_index(h, el));
            for (int64_t q = 0; q < nnz; ++q) REAL(element)[q] = (double)(el[q] + 1);
        }
        SEXP objs[] = {theta0, th1, t0all, t1all, ll, pall, prob, pv, rr, cprob, call, rrall, dic,
                       PROTECT(Rf_ScalarInteger(T)), PROTECT(Rf_ScalarReal(info.percent_converged)), element};
        for (int q = 0; q < 16; ++q) SET_VECTOR_ELT(out, q, objs[q]);
        SET_VECTOR_ELT(chains, c, out);
        UNPROTECT(17);
        R_CheckUserInterrupt();
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
