/* Minimal stand-in for R's C API: just enough declarations to syntax-check r/cardelino_b200_shim.c in an image
 * without R (tests/test_host.py::test_r_shim_compiles_against_stub_headers).  Not R's headers. */
typedef void* SEXP; typedef ptrdiff_t R_xlen_t;
#define VECSXP 19
#define REALSXP 14
extern SEXP R_NilValue, R_DimSymbol, R_NamesSymbol;
void* R_ExternalPtrAddr(SEXP); void R_ClearExternalPtr(SEXP); SEXP R_MakeExternalPtr(void*, SEXP, SEXP);
void R_RegisterCFinalizerEx(SEXP, void (*)(SEXP), int); void Rf_error(const char*, ...);
SEXP PROTECT(SEXP); void UNPROTECT(int); int Rf_asInteger(SEXP); double Rf_asReal(SEXP); int Rf_asLogical(SEXP);
SEXP Rf_getAttrib(SEXP, SEXP); int* INTEGER(SEXP); double* REAL(SEXP); char* R_alloc(size_t, int);
SEXP Rf_install(const char*); int Rf_length(SEXP); const char* CHAR(SEXP); SEXP STRING_ELT(SEXP, R_xlen_t);
SEXP VECTOR_ELT(SEXP, R_xlen_t); int Rf_isNull(SEXP); int Rf_isMatrix(SEXP); SEXP Rf_allocVector(int, R_xlen_t);
SEXP Rf_mkNamed(int, const char**); SEXP Rf_allocMatrix(int, int, int); SEXP Rf_ScalarInteger(int); SEXP Rf_ScalarReal(double);
void SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP); void R_CheckUserInterrupt(void);
#define TRUE 1
#define FALSE 0
