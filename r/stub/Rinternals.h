#ifndef MFB_STUB_RINTERNALS_H
#define MFB_STUB_RINTERNALS_H
#include "R.h"
#define REALSXP 14
#define INTSXP 13
#define LGLSXP 10
#define VECSXP 19
extern SEXP R_NilValue, R_DimSymbol;
extern double R_NaReal;
#define NA_REAL R_NaReal
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
double *REAL(SEXP);
int *INTEGER(SEXP);
int *LOGICAL(SEXP);
int LENGTH(SEXP);
int Rf_asInteger(SEXP);
int Rf_asLogical(SEXP);
double Rf_asReal(SEXP);
SEXP Rf_getAttrib(SEXP, SEXP);
SEXP Rf_allocVector(unsigned int, ptrdiff_t);
SEXP Rf_allocMatrix(unsigned int, int, int);
SEXP Rf_ScalarInteger(int);
SEXP SET_VECTOR_ELT(SEXP, ptrdiff_t, SEXP);
SEXP R_MakeExternalPtr(void *, SEXP, SEXP);
void *R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
#endif
