/* Minimal stand-in for <Rinternals.h> (see R.h in this directory). */
#ifndef R_STUB_RINTERNALS_H
#define R_STUB_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC *SEXP;
typedef ptrdiff_t R_xlen_t;
typedef int Rboolean;
#define TRUE 1
#define FALSE 0
#define REALSXP 14
#define VECSXP 19
extern SEXP R_NilValue, R_NamesSymbol;
extern double R_NaReal;
#define NA_REAL R_NaReal
SEXP Rf_protect(SEXP); void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
SEXP Rf_allocVector(unsigned int type, R_xlen_t n);
SEXP Rf_getAttrib(SEXP, SEXP);
SEXP Rf_mkNamed(unsigned int type, const char **names);
SEXP Rf_ScalarReal(double);
double Rf_asReal(SEXP); int Rf_asInteger(SEXP); int Rf_isNull(SEXP); int Rf_nrows(SEXP);
double *REAL(SEXP); R_xlen_t XLENGTH(SEXP);
SEXP STRING_ELT(SEXP, R_xlen_t); SEXP VECTOR_ELT(SEXP, R_xlen_t); SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
const char *CHAR(SEXP);
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot);
void *R_ExternalPtrAddr(SEXP); void R_ClearExternalPtr(SEXP);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit);
#endif
