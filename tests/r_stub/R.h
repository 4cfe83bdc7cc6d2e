/* Minimal declarations of the R C API used by r/gpedm_b200_glue.c, ONLY so that the glue can be
 * type-checked against include/gpedm.h in an image without R (tests/test_host_logic.py). */
#ifndef R_STUB_H
#define R_STUB_H
#include <stddef.h>
void Rf_error(const char*, ...);
char* R_alloc(size_t, int);
#endif
