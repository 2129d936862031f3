/* Minimal stand-in for R's <R.h>: just enough declarations for `gcc -fsyntax-only r/src/rshim.c` in an image without R
 * (tests/test_abi_cpu.py::test_r_shim_compiles_against_stub_headers).  NOT R: prototypes only, written from R's documented
 * C API ("Writing R Extensions"), no code from R. */
#ifndef R_STUB_R_H
#define R_STUB_R_H
#include <stddef.h>
void Rf_error(const char *fmt, ...);
char *R_alloc(size_t n, int size);
void R_CheckUserInterrupt(void);
#endif
