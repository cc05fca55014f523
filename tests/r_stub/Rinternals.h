#ifndef SHARP_RINTERNALS_STUB_H
#define SHARP_RINTERNALS_STUB_H
#include "R.h"
SEXP Rf_allocVector(unsigned int, R_xlen_t);
SEXP Rf_allocMatrix(unsigned int, int, int);
SEXP Rf_alloc3DArray(unsigned int, int, int, int);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
int *INTEGER(SEXP);
double *REAL(SEXP);
typedef unsigned char Rbyte;
Rbyte *RAW(SEXP);
R_xlen_t XLENGTH(SEXP);
int TYPEOF(SEXP);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP STRING_ELT(SEXP, R_xlen_t);
const char *CHAR(SEXP);
int Rf_length(SEXP);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
int Rf_asInteger(SEXP);
int Rf_asLogical(SEXP);
double Rf_asReal(SEXP);
Rboolean Rf_isNull(SEXP);
Rboolean Rf_isNumeric(SEXP);
Rboolean Rf_isMatrix(SEXP);
Rboolean Rf_isReal(SEXP);
Rboolean Rf_inherits(SEXP, const char *);
SEXP Rf_install(const char *);
SEXP R_do_slot(SEXP, SEXP);
SEXP Rf_mkString(const char *);
SEXP Rf_ScalarInteger(int);
SEXP Rf_ScalarReal(double);
SEXP R_MakeExternalPtr(void *, SEXP, SEXP);
void *R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
#endif
