/* Stand-in for <R.h> (see Rinternals.h in this directory). */
#ifndef R_STUB_H
#define R_STUB_H
#include <stddef.h>
#include "Rinternals.h"
#endif
