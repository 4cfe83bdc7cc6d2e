#ifndef R_STUB_INTERNALS_H
#define R_STUB_INTERNALS_H
typedef struct SEXPREC* SEXP;
typedef int Rboolean;
typedef long R_xlen_t;
#define TRUE 1
#define FALSE 0
#define REALSXP 14
#define INTSXP 13
#define VECSXP 19
extern SEXP R_NilValue;
double* REAL(SEXP);
int* INTEGER(SEXP);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
int Rf_length(SEXP);
int Rf_asInteger(SEXP);
int Rf_asLogical(SEXP);
double Rf_asReal(SEXP);
int Rf_isNull(SEXP);
SEXP Rf_allocVector(unsigned int, long);
SEXP Rf_allocMatrix(unsigned int, int, int);
SEXP Rf_mkNamed(unsigned int, const char**);
SEXP Rf_ScalarReal(double);
SEXP Rf_ScalarInteger(int);
SEXP Rf_getAttrib(SEXP, SEXP);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
SEXP Rf_install(const char*);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
SEXP SET_VECTOR_ELT(SEXP, long, SEXP);
SEXP R_MakeExternalPtr(void*, SEXP, SEXP);
void* R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
#endif
