/* TEST INFRASTRUCTURE: a tiny implementation of the slice of R's C API declared in Rinternals.h / R.h of this
 * directory, so that r/cardelino_b200_shim.c runs without R.  Objects are tagged vectors with `names` and `dim`
 * attributes; nothing is garbage collected (PROTECT only counts); Rf_error longjmps to the driver's handler like
 * R's does to the top level; a "user interrupt" can be armed to fire at the n-th R_CheckUserInterrupt. */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct fake_sexp {
    unsigned int type;
    R_xlen_t n;
    void* data;            /* double / int / SEXP / char payload */
    SEXP names, dim;
    void* ext;             /* external pointer address */
    void (*fin)(SEXP);
};

static struct fake_sexp nil_obj = {NILSXP, 0, NULL, NULL, NULL, NULL, NULL};
static struct fake_sexp dim_sym = {SYMSXP, 0, (void*)"dim", NULL, NULL, NULL, NULL};
static struct fake_sexp names_sym = {SYMSXP, 0, (void*)"names", NULL, NULL, NULL, NULL};
SEXP R_NilValue = &nil_obj, R_DimSymbol = &dim_sym, R_NamesSymbol = &names_sym;

/* ---- driver-visible state ---- */
jmp_buf fake_r_top;                 /* where Rf_error lands */
int fake_r_top_set = 0;
char fake_r_last_error[1024];
char fake_r_stdout[1 << 16];        /* what Rprintf wrote */
int fake_r_protect_depth = 0, fake_r_protect_max = 0;
int fake_r_interrupt_at = 0;        /* fire a user interrupt at the n-th R_CheckUserInterrupt (0 = never) */
int fake_r_interrupt_polls = 0;
static jmp_buf* toplevel_exec_ctx = NULL;

static SEXP new_obj(unsigned int type, R_xlen_t n, size_t elt) {
    SEXP s = (SEXP)calloc(1, sizeof(struct fake_sexp));
    s->type = type; s->n = n;
    s->data = calloc((size_t)(n > 0 ? n : 1), elt);
    s->names = R_NilValue; s->dim = R_NilValue;
    return s;
}

SEXP Rf_allocVector(unsigned int type, R_xlen_t n) {
    switch (type) {
        case REALSXP: return new_obj(type, n, sizeof(double));
        case INTSXP: case LGLSXP: return new_obj(type, n, sizeof(int));
        case VECSXP: case STRSXP: {
            SEXP s = new_obj(type, n, sizeof(SEXP));
            for (R_xlen_t q = 0; q < n; ++q) ((SEXP*)s->data)[q] = R_NilValue;
            return s;
        }
        default: Rf_error("fake_r: allocVector of type %u", type);
    }
}
SEXP Rf_allocMatrix(unsigned int type, int nr, int nc) {
    if (nr < 0 || nc < 0) Rf_error("negative extents to matrix");
    SEXP s = Rf_allocVector(type, (R_xlen_t)nr * nc);
    s->dim = Rf_allocVector(INTSXP, 2);
    INTEGER(s->dim)[0] = nr; INTEGER(s->dim)[1] = nc;
    return s;
}
SEXP Rf_mkChar(const char* c) {
    SEXP s = new_obj(CHARSXP, (R_xlen_t)strlen(c), 1);
    free(s->data);
    s->data = strdup(c);
    return s;
}
SEXP Rf_mkNamed(unsigned int type, const char** names) {
    R_xlen_t n = 0;
    while (names[n][0]) ++n;
    SEXP s = Rf_allocVector(type, n);
    s->names = Rf_allocVector(STRSXP, n);
    for (R_xlen_t q = 0; q < n; ++q) SET_STRING_ELT(s->names, q, Rf_mkChar(names[q]));
    return s;
}
SEXP Rf_ScalarInteger(int v) { SEXP s = Rf_allocVector(INTSXP, 1); INTEGER(s)[0] = v; return s; }
SEXP Rf_ScalarLogical(int v) { SEXP s = Rf_allocVector(LGLSXP, 1); LOGICAL(s)[0] = v; return s; }
SEXP Rf_ScalarReal(double v) { SEXP s = Rf_allocVector(REALSXP, 1); REAL(s)[0] = v; return s; }

static void need(SEXP s, unsigned int t1, unsigned int t2, const char* what) {
    if (!s || (s->type != t1 && s->type != t2)) Rf_error("fake_r: %s() on an object of type %u", what, s ? s->type : 999u);
}
double* REAL(SEXP s) { need(s, REALSXP, REALSXP, "REAL"); return (double*)s->data; }
int* INTEGER(SEXP s) { need(s, INTSXP, LGLSXP, "INTEGER"); return (int*)s->data; }
int* LOGICAL(SEXP s) { need(s, LGLSXP, LGLSXP, "LOGICAL"); return (int*)s->data; }
SEXP VECTOR_ELT(SEXP s, R_xlen_t q) {
    need(s, VECSXP, VECSXP, "VECTOR_ELT");
    if (q < 0 || q >= s->n) Rf_error("fake_r: VECTOR_ELT out of bounds");
    return ((SEXP*)s->data)[q];
}
SEXP SET_VECTOR_ELT(SEXP s, R_xlen_t q, SEXP v) {
    need(s, VECSXP, VECSXP, "SET_VECTOR_ELT");
    if (q < 0 || q >= s->n) Rf_error("fake_r: SET_VECTOR_ELT out of bounds");
    ((SEXP*)s->data)[q] = v;
    return v;
}
SEXP STRING_ELT(SEXP s, R_xlen_t q) {
    need(s, STRSXP, STRSXP, "STRING_ELT");
    if (q < 0 || q >= s->n) Rf_error("fake_r: STRING_ELT out of bounds");
    return ((SEXP*)s->data)[q];
}
void SET_STRING_ELT(SEXP s, R_xlen_t q, SEXP v) { need(s, STRSXP, STRSXP, "SET_STRING_ELT"); ((SEXP*)s->data)[q] = v; }
const char* CHAR(SEXP s) { need(s, CHARSXP, CHARSXP, "CHAR"); return (const char*)s->data; }
int Rf_length(SEXP s) { return s ? (int)s->n : 0; }
Rboolean Rf_isNull(SEXP s) { return s == R_NilValue || s->type == NILSXP; }
Rboolean Rf_isMatrix(SEXP s) { return s && s->dim != R_NilValue && s->dim->n == 2; }
SEXP Rf_install(const char* name) {
    if (!strcmp(name, "dim")) return R_DimSymbol;
    if (!strcmp(name, "names")) return R_NamesSymbol;
    Rf_error("fake_r: unknown symbol %s", name);
}
SEXP Rf_getAttrib(SEXP s, SEXP sym) {
    if (sym == R_DimSymbol) return s->dim;
    if (sym == R_NamesSymbol) return s->names;
    return R_NilValue;
}
SEXP Rf_setAttrib(SEXP s, SEXP sym, SEXP v) {
    if (sym == R_DimSymbol) s->dim = v;
    else if (sym == R_NamesSymbol) s->names = v;
    return v;
}
int Rf_asInteger(SEXP s) {
    if (s->n < 1) return (int)0x80000000;
    return s->type == REALSXP ? (int)REAL(s)[0] : INTEGER(s)[0];
}
int Rf_asLogical(SEXP s) {
    if (s->n < 1) return (int)0x80000000;
    return s->type == REALSXP ? REAL(s)[0] != 0.0 : INTEGER(s)[0] != 0;
}
double Rf_asReal(SEXP s) {
    if (s->n < 1) return 0.0 / 0.0;
    return s->type == REALSXP ? REAL(s)[0] : (double)INTEGER(s)[0];
}

SEXP Rf_protect(SEXP s) {
    if (++fake_r_protect_depth > fake_r_protect_max) fake_r_protect_max = fake_r_protect_depth;
    return s;
}
void Rf_unprotect(int n) {
    fake_r_protect_depth -= n;
    if (fake_r_protect_depth < 0) { fprintf(stderr, "fake_r: protect stack imbalance\n"); abort(); }
}
char* R_alloc(size_t n, int size) { return (char*)calloc(n ? n : 1, (size_t)size); }

SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot) {
    (void)tag; (void)prot;
    SEXP s = new_obj(EXTPTRSXP, 0, 1);
    s->ext = p;
    return s;
}
void* R_ExternalPtrAddr(SEXP s) { return s->type == EXTPTRSXP ? s->ext : NULL; }
void R_ClearExternalPtr(SEXP s) { s->ext = NULL; }
void R_RegisterCFinalizerEx(SEXP s, void (*fin)(SEXP), Rboolean onexit) { (void)onexit; s->fin = fin; }
void fake_r_run_finalizer(SEXP s) { if (s->fin) s->fin(s); }

void Rf_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(fake_r_last_error, sizeof(fake_r_last_error), fmt, ap);
    va_end(ap);
    if (!fake_r_top_set) { fprintf(stderr, "fake_r: Rf_error outside a top-level context: %s\n", fake_r_last_error); abort(); }
    fake_r_protect_depth = 0;           /* R unwinds the protect stack to the context's level */
    longjmp(fake_r_top, 1);
}
void Rprintf(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    const size_t used = strlen(fake_r_stdout);
    vsnprintf(fake_r_stdout + used, sizeof(fake_r_stdout) - used, fmt, ap);
    va_end(ap);
}

/* R_CheckUserInterrupt longjmps to the innermost top-level context when an interrupt is pending */
void R_CheckUserInterrupt(void) {
    ++fake_r_interrupt_polls;
    if (fake_r_interrupt_at && fake_r_interrupt_polls >= fake_r_interrupt_at) {
        fake_r_interrupt_at = 0;                 /* delivered once */
        if (toplevel_exec_ctx) longjmp(*toplevel_exec_ctx, 1);
        snprintf(fake_r_last_error, sizeof(fake_r_last_error), "<user interrupt>");
        if (!fake_r_top_set) abort();
        longjmp(fake_r_top, 2);
    }
}
Rboolean R_ToplevelExec(void (*fun)(void*), void* data) {
    jmp_buf ctx;
    jmp_buf* saved = toplevel_exec_ctx;
    Rboolean ok = TRUE;
    toplevel_exec_ctx = &ctx;
    if (setjmp(ctx) == 0) fun(data);
    else ok = FALSE;
    toplevel_exec_ctx = saved;
    return ok;
}

static const R_CallMethodDef* registered = NULL;
int R_registerRoutines(DllInfo* dll, const void* c, const R_CallMethodDef* call, const void* f, const void* e) {
    (void)dll; (void)c; (void)f; (void)e;
    registered = call;
    return 1;
}
Rboolean R_useDynamicSymbols(DllInfo* dll, Rboolean v) { (void)dll; (void)v; return TRUE; }
/* .Call("name", ...) lookup in the registration table, with R's argument-count check */
DL_FUNC fake_r_lookup(const char* name, int nargs) {
    for (const R_CallMethodDef* m = registered; m && m->name; ++m)
        if (!strcmp(m->name, name)) {
            if (m->numArgs != nargs) Rf_error("Incorrect number of arguments (%d), expecting %d for '%s'", nargs, m->numArgs, name);
            return m->fun;
        }
    Rf_error("C symbol name \"%s\" not in load table", name);
}
