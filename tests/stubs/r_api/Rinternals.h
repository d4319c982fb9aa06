/* Rinternals.h -- DECLARATIONS ONLY of the part of R's C API that shim/b200nn_r.c uses, written from the
 * documented signatures ("Writing R Extensions", section 5 / 6).  TEST INFRASTRUCTURE: R is not in the build
 * image, so tests/test_host_logic.py compiles the shim against these prototypes (gcc -fsyntax-only -Wall
 * -Werror) to check argument counts and types of every call; nothing here is ever linked or shipped. */
#ifndef B200NN_STUB_RINTERNALS_H
#define B200NN_STUB_RINTERNALS_H
#include <stddef.h>
#include <stdlib.h>
typedef struct SEXPREC *SEXP;
typedef ptrdiff_t R_xlen_t;
typedef unsigned int SEXPTYPE;
typedef enum { FALSE = 0, TRUE } Rboolean;
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
extern SEXP R_NilValue, R_NamesSymbol, R_DimSymbol;
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
SEXP Rf_allocVector(SEXPTYPE, R_xlen_t);
SEXP Rf_allocMatrix(SEXPTYPE, int, int);
SEXP Rf_coerceVector(SEXP, SEXPTYPE);
SEXP Rf_duplicate(SEXP);
SEXP Rf_install(const char *);
SEXP Rf_mkChar(const char *);
SEXP Rf_getAttrib(SEXP, SEXP);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
SEXP Rf_ScalarLogical(int);
int Rf_asInteger(SEXP);
int Rf_asLogical(SEXP);
double Rf_asReal(SEXP);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
R_xlen_t Rf_xlength(SEXP);
Rboolean Rf_isNull(SEXP);
Rboolean Rf_isReal(SEXP);
Rboolean Rf_isInteger(SEXP);
Rboolean Rf_isLogical(SEXP);
Rboolean Rf_isMatrix(SEXP);
int *INTEGER(SEXP);
double *REAL(SEXP);
const char *CHAR(SEXP);
SEXP STRING_ELT(SEXP, R_xlen_t);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
void SET_STRING_ELT(SEXP, R_xlen_t, SEXP);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP R_do_MAKE_CLASS(const char *);
SEXP R_do_new_object(SEXP);
SEXP R_do_slot(SEXP, SEXP);
SEXP R_do_slot_assign(SEXP, SEXP, SEXP);
SEXP R_MakeExternalPtr(void *, SEXP, SEXP);
void *R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
void R_CheckUserInterrupt(void);
char *R_alloc(size_t, int);
#if defined(__GNUC__)
void Rf_error(const char *, ...) __attribute__((noreturn, format(printf, 1, 2)));
void Rprintf(const char *, ...) __attribute__((format(printf, 1, 2)));
void REprintf(const char *, ...) __attribute__((format(printf, 1, 2)));
#else
void Rf_error(const char *, ...);
void Rprintf(const char *, ...);
void REprintf(const char *, ...);
#endif
#endif
