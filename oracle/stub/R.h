/* Minimal stand-in for R's <R.h>, used ONLY to compile the reference's own search core
 * (under /root/reference/src, where it lies) into oracle/_ref/libcatnet_ref.so.
 * Test infrastructure: never linked into the product library.
 * The reference needs: error, warning, Rprintf, GetRNGstate, PutRNGstate, unif_rand, norm_rand, SEXP. */
#ifndef CNB_ORACLE_STUB_R_H
#define CNB_ORACLE_STUB_R_H
#include <stdio.h>
#include <stdlib.h>
#include <stdarg.h>
#include <string.h>
#include <limits.h>
#include <math.h>
#include <errno.h>
#ifdef __cplusplus
extern "C" {
#endif
void cnb_stub_error(const char *fmt, ...);
void cnb_stub_warning(const char *fmt, ...);
void Rprintf(const char *fmt, ...);
void GetRNGstate(void);
void PutRNGstate(void);
double unif_rand(void);
double norm_rand(void);
#ifdef __cplusplus
}
#endif
#define error cnb_stub_error
#define warning cnb_stub_warning
typedef void *SEXP;
#endif
