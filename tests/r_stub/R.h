/* Minimal forward declarations of the stable R C API used by r_glue/sharp_r_glue.c -- ONLY so that the glue can be
 * syntax-checked in an image without R (tests/test_abi.py).  Not R's headers; never linked. */
#ifndef SHARP_R_STUB_H
#define SHARP_R_STUB_H
#include <stddef.h>
typedef struct SEXPREC *SEXP;
typedef ptrdiff_t R_xlen_t;
typedef int Rboolean;
#define TRUE 1
#define FALSE 0
extern SEXP R_NilValue;
enum { INTSXP = 13, REALSXP = 14, VECSXP = 19, RAWSXP = 24 };
void Rf_error(const char *, ...);
char *R_alloc(size_t, int);
#endif
