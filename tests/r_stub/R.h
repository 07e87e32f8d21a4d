/* stand-in for <R.h>: see Rinternals.h in this directory */
#ifndef FAKE_R_H
#define FAKE_R_H
#include <stddef.h>
void Rprintf(const char*, ...);
#endif
