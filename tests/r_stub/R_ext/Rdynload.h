/* stand-in for <R_ext/Rdynload.h>: see ../Rinternals.h */
#ifndef FAKE_RDYNLOAD_H
#define FAKE_RDYNLOAD_H
typedef void* (*DL_FUNC)();
typedef struct { const char* name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct _DllInfo DllInfo;
int R_registerRoutines(DllInfo*, const void*, const R_CallMethodDef*, const void*, const void*);
Rboolean R_useDynamicSymbols(DllInfo*, Rboolean);
#endif
