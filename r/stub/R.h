/* Minimal stand-in for R's C API: just the declarations r/src/r_shim.c uses, so that the shim can be type-checked
 * (gcc -fsyntax-only) on a machine without R.  NOT a replacement for R's headers. */
#ifndef MFB_STUB_R_H
#define MFB_STUB_R_H
#include <stddef.h>
#include <stdint.h>
typedef struct SEXPREC *SEXP;
typedef int Rboolean;
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
void Rf_error(const char *fmt, ...) __attribute__((noreturn, format(printf, 1, 2)));
#endif
