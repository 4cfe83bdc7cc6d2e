/* Minimal stand-in for R's C API: exactly the subset r/gpedm_b200_glue.c uses.  R is not installed in the
 * build image; tests/r_stub/r_stub_runtime.c implements these functions so that the glue can be compiled,
 * loaded and EXECUTED (tests/test_r_glue.py) without R.  Semantics follow "Writing R Extensions" for the
 * calls below; garbage collection is replaced by an arena that the test harness frees. */
#ifndef R_STUB_INTERNALS_H
#define R_STUB_INTERNALS_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct SEXPREC* SEXP;
typedef int Rboolean;
typedef ptrdiff_t R_xlen_t;
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
#define NILSXP 0
#define SYMSXP 1
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
#define EXTPTRSXP 22
extern SEXP R_NilValue;
double* REAL(SEXP);
int* INTEGER(SEXP);
int* LOGICAL(SEXP);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
int Rf_length(SEXP);
int Rf_asInteger(SEXP);
int Rf_asLogical(SEXP);
double Rf_asReal(SEXP);
int Rf_isNull(SEXP);
SEXP Rf_allocVector(unsigned int, R_xlen_t);
SEXP Rf_allocMatrix(unsigned int, int, int);
SEXP Rf_mkNamed(unsigned int, const char**);
SEXP Rf_ScalarReal(double);
SEXP Rf_ScalarInteger(int);
SEXP Rf_ScalarLogical(int);
SEXP Rf_getAttrib(SEXP, SEXP);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
SEXP Rf_install(const char*);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
SEXP R_MakeExternalPtr(void*, SEXP, SEXP);
void* R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
char* R_alloc(size_t, int);
#ifdef __GNUC__
void Rf_error(const char*, ...) __attribute__((noreturn, format(printf, 1, 2)));
#else
void Rf_error(const char*, ...);
#endif
#ifdef __cplusplus
}
#endif
#endif
