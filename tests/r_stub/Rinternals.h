/* Minimal stand-in for R's C API (NOT R's headers): the entry points r/cardelino_b200_shim.c uses, with R's
 * signatures, so that the shim can be compiled AND executed in an image without R.  fake_r.c implements them over a
 * tiny tagged-vector object; tests/test_r_shim.py drives the .Call entry points through shim_driver.c. */
#ifndef FAKE_RINTERNALS_H
#define FAKE_RINTERNALS_H
#include <stddef.h>
typedef struct fake_sexp* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef enum { FALSE = 0, TRUE } Rboolean;
#define NILSXP 0
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
#define CHARSXP 9
#define EXTPTRSXP 22
#define SYMSXP 1
extern SEXP R_NilValue, R_DimSymbol, R_NamesSymbol;
void* R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
SEXP R_MakeExternalPtr(void*, SEXP, SEXP);
void R_RegisterCFinalizerEx(SEXP, void (*)(SEXP), Rboolean);
void Rf_error(const char*, ...) __attribute__((noreturn));
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
int Rf_asInteger(SEXP);
double Rf_asReal(SEXP);
int Rf_asLogical(SEXP);
SEXP Rf_getAttrib(SEXP, SEXP);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
int* INTEGER(SEXP);
int* LOGICAL(SEXP);
double* REAL(SEXP);
char* R_alloc(size_t, int);
SEXP Rf_install(const char*);
int Rf_length(SEXP);
const char* CHAR(SEXP);
SEXP STRING_ELT(SEXP, R_xlen_t);
void SET_STRING_ELT(SEXP, R_xlen_t, SEXP);
SEXP Rf_mkChar(const char*);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
Rboolean Rf_isNull(SEXP);
Rboolean Rf_isMatrix(SEXP);
SEXP Rf_allocVector(unsigned int, R_xlen_t);
SEXP Rf_mkNamed(unsigned int, const char**);
SEXP Rf_allocMatrix(unsigned int, int, int);
SEXP Rf_ScalarInteger(int);
SEXP Rf_ScalarReal(double);
SEXP Rf_ScalarLogical(int);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
void R_CheckUserInterrupt(void);
Rboolean R_ToplevelExec(void (*fun)(void*), void* data);
#endif
