/* A tiny fake R runtime: the ~30 R C-API functions r/gpedm_b200_glue.c uses, plus a harness
 * (rstub_*) through which tests build SEXP arguments, invoke a registered .Call entry by name the way
 * R's .Call does (R_registerRoutines table lookup, argument-count check, Rf_error -> condition), read
 * results back, and run the external-pointer finalizers R's garbage collector would run.
 * Test infrastructure only; it makes the reference-side binding EXECUTABLE in an image without R. */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "Rinternals.h"
#include "R_ext/Rdynload.h"

struct SEXPREC {
  int type;
  R_xlen_t length;
  int nrow, ncol;      /* ncol < 0: no dim attribute */
  void* data;          /* REAL / INTEGER / LOGICAL payload, SEXP[] for VECSXP, char* for SYMSXP */
  char** names;        /* VECSXP names */
  void* extptr;
  SEXP attr_name[4], attr_val[4];
  int nattr;
  struct SEXPREC* next; /* arena list */
};

static struct SEXPREC nil_rec = {NILSXP, 0, 0, -1, NULL, NULL, NULL, {0}, {0}, 0, NULL};
SEXP R_NilValue = &nil_rec;
static SEXP arena = NULL;
static int protect_depth = 0;
static char err_msg[1024];
static jmp_buf err_jmp;
static int err_armed = 0;
static struct { SEXP p; R_CFinalizer_t fn; } fins[256];
static int nfins = 0;
static const R_CallMethodDef* call_table = NULL;
static int dynamic_symbols = -1;
static void* ralloc_list[256];
static int nralloc = 0;

static SEXP new_sexp(int type, R_xlen_t n) {
  SEXP s = (SEXP)calloc(1, sizeof(struct SEXPREC));
  s->type = type;
  s->length = n;
  s->ncol = -1;
  s->next = arena;
  arena = s;
  return s;
}

void Rf_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_msg, sizeof(err_msg), fmt, ap);
  va_end(ap);
  if (!err_armed) {
    fprintf(stderr, "r_stub: Rf_error outside .Call: %s\n", err_msg);
    abort();
  }
  longjmp(err_jmp, 1);
}

char* R_alloc(size_t n, int size) {
  void* p = calloc(n ? n : 1, (size_t)size);
  if (nralloc < 256) ralloc_list[nralloc++] = p;
  return (char*)p;
}

static void need(SEXP s, int type, const char* what) {
  if (s->type != type) Rf_error("r_stub: %s applied to a SEXP of type %d", what, s->type);
}
double* REAL(SEXP s) { need(s, REALSXP, "REAL()"); return (double*)s->data; }
int* INTEGER(SEXP s) {
  if (s->type != INTSXP && s->type != LGLSXP) Rf_error("r_stub: INTEGER() applied to a SEXP of type %d", s->type);
  return (int*)s->data;
}
int* LOGICAL(SEXP s) { need(s, LGLSXP, "LOGICAL()"); return (int*)s->data; }
int Rf_nrows(SEXP s) { return s->ncol >= 0 ? s->nrow : (int)s->length; }
int Rf_ncols(SEXP s) { return s->ncol >= 0 ? s->ncol : 1; }
int Rf_length(SEXP s) { return (int)s->length; }
int Rf_isNull(SEXP s) { return s->type == NILSXP; }
int Rf_asInteger(SEXP s) {
  if (s->length < 1) return -2147483647 - 1; /* NA_INTEGER */
  if (s->type == INTSXP || s->type == LGLSXP) return ((int*)s->data)[0];
  if (s->type == REALSXP) return (int)((double*)s->data)[0];
  Rf_error("r_stub: asInteger on type %d", s->type);
}
int Rf_asLogical(SEXP s) { return Rf_asInteger(s) != 0; }
double Rf_asReal(SEXP s) {
  if (s->length < 1) Rf_error("r_stub: asReal on a zero-length vector");
  if (s->type == REALSXP) return ((double*)s->data)[0];
  if (s->type == INTSXP || s->type == LGLSXP) return (double)((int*)s->data)[0];
  Rf_error("r_stub: asReal on type %d", s->type);
}
SEXP Rf_allocVector(unsigned int type, R_xlen_t n) {
  SEXP s = new_sexp((int)type, n);
  size_t el = type == REALSXP ? sizeof(double) : (type == VECSXP ? sizeof(SEXP) : sizeof(int));
  s->data = calloc(n > 0 ? (size_t)n : 1, el);
  if (type == VECSXP)
    for (R_xlen_t i = 0; i < n; ++i) ((SEXP*)s->data)[i] = R_NilValue;
  return s;
}
SEXP Rf_allocMatrix(unsigned int type, int nr, int nc) {
  SEXP s = Rf_allocVector(type, (R_xlen_t)nr * nc);
  s->nrow = nr;
  s->ncol = nc;
  return s;
}
SEXP Rf_mkNamed(unsigned int type, const char** names) {
  int n = 0;
  while (names[n][0] != '\0') ++n;
  SEXP s = Rf_allocVector(type, n);
  s->names = (char**)calloc((size_t)n + 1, sizeof(char*));
  for (int i = 0; i < n; ++i) s->names[i] = strdup(names[i]);
  return s;
}
SEXP Rf_ScalarReal(double v) { SEXP s = Rf_allocVector(REALSXP, 1); ((double*)s->data)[0] = v; return s; }
SEXP Rf_ScalarInteger(int v) { SEXP s = Rf_allocVector(INTSXP, 1); ((int*)s->data)[0] = v; return s; }
SEXP Rf_ScalarLogical(int v) { SEXP s = Rf_allocVector(LGLSXP, 1); ((int*)s->data)[0] = v != 0; return s; }
SEXP Rf_install(const char* name) {
  SEXP s = new_sexp(SYMSXP, 1);
  s->data = strdup(name);
  return s;
}
SEXP Rf_getAttrib(SEXP s, SEXP name) {
  for (int i = 0; i < s->nattr; ++i)
    if (strcmp((char*)s->attr_name[i]->data, (char*)name->data) == 0) return s->attr_val[i];
  return R_NilValue;
}
SEXP Rf_setAttrib(SEXP s, SEXP name, SEXP val) {
  if (s->nattr < 4) {
    s->attr_name[s->nattr] = name;
    s->attr_val[s->nattr++] = val;
  }
  return val;
}
SEXP Rf_protect(SEXP s) { ++protect_depth; return s; }
void Rf_unprotect(int n) { protect_depth -= n; }
SEXP SET_VECTOR_ELT(SEXP s, R_xlen_t i, SEXP v) {
  need(s, VECSXP, "SET_VECTOR_ELT");
  if (i < 0 || i >= s->length) Rf_error("r_stub: SET_VECTOR_ELT index %ld out of range", (long)i);
  ((SEXP*)s->data)[i] = v;
  return v;
}
SEXP VECTOR_ELT(SEXP s, R_xlen_t i) {
  need(s, VECSXP, "VECTOR_ELT");
  if (i < 0 || i >= s->length) Rf_error("r_stub: VECTOR_ELT index %ld out of range", (long)i);
  return ((SEXP*)s->data)[i];
}
SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot) {
  (void)tag; (void)prot;
  SEXP s = new_sexp(EXTPTRSXP, 1);
  s->extptr = p;
  return s;
}
void* R_ExternalPtrAddr(SEXP s) { need(s, EXTPTRSXP, "R_ExternalPtrAddr"); return s->extptr; }
void R_ClearExternalPtr(SEXP s) { need(s, EXTPTRSXP, "R_ClearExternalPtr"); s->extptr = NULL; }
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fn, Rboolean onexit) {
  (void)onexit;
  if (nfins < 256) { fins[nfins].p = s; fins[nfins].fn = fn; ++nfins; }
}
int R_registerRoutines(DllInfo* dll, const void* c, const R_CallMethodDef* call, const void* f, const void* e) {
  (void)dll; (void)c; (void)f; (void)e;
  call_table = call;
  return 1;
}
int R_useDynamicSymbols(DllInfo* dll, int v) { (void)dll; dynamic_symbols = v; return 1; }

/* ------------------------------------------------------------------ harness (called from the tests) */
void R_init_gpedmb200(DllInfo*);
int rstub_load(void) {  /* what dyn.load + library.dynam do: run the init routine, return #registered entries */
  R_init_gpedmb200(NULL);
  int n = 0;
  while (call_table && call_table[n].name) ++n;
  return n;
}
int rstub_dynamic_symbols(void) { return dynamic_symbols; }
const char* rstub_entry_name(int i) { return call_table[i].name; }
int rstub_entry_nargs(int i) { return call_table[i].numArgs; }
SEXP rstub_nil(void) { return R_NilValue; }
SEXP rstub_real(const double* v, R_xlen_t n, int nrow, int ncol) {
  SEXP s = Rf_allocVector(REALSXP, n);
  if (n) memcpy(s->data, v, (size_t)n * sizeof(double));
  if (ncol >= 0) { s->nrow = nrow; s->ncol = ncol; }
  return s;
}
SEXP rstub_int(const int* v, R_xlen_t n, int logical) {
  SEXP s = Rf_allocVector(logical ? LGLSXP : INTSXP, n);
  if (n) memcpy(s->data, v, (size_t)n * sizeof(int));
  return s;
}
int rstub_type(SEXP s) { return s->type; }
long rstub_length(SEXP s) { return (long)s->length; }
int rstub_nrow(SEXP s) { return Rf_nrows(s); }
int rstub_ncol(SEXP s) { return Rf_ncols(s); }
void* rstub_data(SEXP s) { return s->data; }
SEXP rstub_list_get(SEXP s, const char* name) {
  if (s->type != VECSXP || !s->names) return NULL;
  for (R_xlen_t i = 0; i < s->length; ++i)
    if (strcmp(s->names[i], name) == 0) return ((SEXP*)s->data)[i];
  return NULL;
}
SEXP rstub_list_elt(SEXP s, long i) { return (s->type == VECSXP && i >= 0 && i < s->length) ? ((SEXP*)s->data)[i] : NULL; }
const char* rstub_last_error(void) { return err_msg; }
int rstub_protect_depth(void) { return protect_depth; }

/* .Call(name, ...): look the entry up in the registered table, check the argument count like R does,
 * convert Rf_error into a return code (1) with the message in rstub_last_error(). */
int rstub_dot_call(const char* name, int nargs, SEXP* args, SEXP* out) {
  typedef SEXP (*F1)(SEXP);
  typedef SEXP (*F2)(SEXP, SEXP);
  typedef SEXP (*F4)(SEXP, SEXP, SEXP, SEXP);
  typedef SEXP (*F5)(SEXP, SEXP, SEXP, SEXP, SEXP);
  typedef SEXP (*F6)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
  typedef SEXP (*F8)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
  typedef SEXP (*F10)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
  err_msg[0] = '\0';
  const R_CallMethodDef* e = NULL;
  for (int i = 0; call_table && call_table[i].name; ++i)
    if (strcmp(call_table[i].name, name) == 0) e = &call_table[i];
  if (!e) { snprintf(err_msg, sizeof(err_msg), "C symbol name \"%s\" not in load table", name); return 2; }
  if (e->numArgs != nargs) {
    snprintf(err_msg, sizeof(err_msg), "Incorrect number of arguments (%d), expecting %d for '%s'", nargs, e->numArgs, name);
    return 3;
  }
  const int depth0 = protect_depth;
  err_armed = 1;
  if (setjmp(err_jmp)) {
    err_armed = 0;
    protect_depth = depth0; /* R unwinds the protect stack on error */
    return 1;
  }
  SEXP r = NULL;
  SEXP* a = args;
  switch (nargs) {
    case 1: r = ((F1)e->fun)(a[0]); break;
    case 2: r = ((F2)e->fun)(a[0], a[1]); break;
    case 4: r = ((F4)e->fun)(a[0], a[1], a[2], a[3]); break;
    case 5: r = ((F5)e->fun)(a[0], a[1], a[2], a[3], a[4]); break;
    case 6: r = ((F6)e->fun)(a[0], a[1], a[2], a[3], a[4], a[5]); break;
    case 8: r = ((F8)e->fun)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]); break;
    case 10: r = ((F10)e->fun)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9]); break;
    default: err_armed = 0; snprintf(err_msg, sizeof(err_msg), "r_stub: %d-argument entries are not wired", nargs); return 4;
  }
  err_armed = 0;
  if (protect_depth != depth0) {
    snprintf(err_msg, sizeof(err_msg), "stack imbalance in '%s': %d then %d", name, depth0, protect_depth);
    protect_depth = depth0;
    return 5;
  }
  *out = r;
  return 0;
}

/* gc(): run the finalizers of every external pointer (the tests drop all references first), free the arena */
int rstub_gc(void) {
  int ran = 0;
  for (int i = 0; i < nfins; ++i)
    if (fins[i].p->extptr) { fins[i].fn(fins[i].p); ++ran; }
  nfins = 0;
  while (arena) {
    SEXP n = arena->next;
    if (arena->type == VECSXP && arena->names) {
      for (R_xlen_t i = 0; i < arena->length; ++i) free(arena->names[i]);
      free(arena->names);
    }
    free(arena->data);
    free(arena);
    arena = n;
  }
  for (int i = 0; i < nralloc; ++i) free(ralloc_list[i]);
  nralloc = 0;
  return ran;
}
