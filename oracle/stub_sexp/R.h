/* Functional miniature of R's C API, used ONLY to compile the reference's own catnet_entropy.cpp (under
 * /root/reference/src, where it lies, unmodified) into oracle/_ref/libcatnet_ref_pairwise.so.
 * Test infrastructure: never linked into the product library.
 *
 * An SEXP here is a typed vector with an optional dim attribute; the reference's pairwise functions use nothing else
 * (isMatrix, isNull, AS_INTEGER, GET_DIM, INTEGER, INTEGER_POINTER, NEW_NUMERIC, NEW_INTEGER, NUMERIC_POINTER, PROTECT).
 * One deliberate difference from R: AS_INTEGER(R_NilValue) stays R_NilValue (R's coerceVector would return integer(0),
 * which the reference then takes for a perturbation matrix, catnet_entropy.cpp:139-143); the reference's evident
 * intention -- NULL means "no perturbations" -- is what the oracle pins. */
#ifndef CNB_ORACLE_STUB_SEXP_R_H
#define CNB_ORACLE_STUB_SEXP_R_H
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

typedef struct cnb_sexp {
	int type;        /* 13 = INTSXP, 14 = REALSXP */
	int n;
	int ndim;        /* 0 or 2 */
	int dim[2];
	void *data;
	struct cnb_sexp *dim_sexp;
} *SEXP;
SEXP cnb_stub_alloc(int type, int n);
SEXP cnb_stub_as_integer(SEXP x);
SEXP cnb_stub_get_dim(SEXP x);
void cnb_stub_free_all(void);
#ifdef __cplusplus
}
#endif
#define error cnb_stub_error
#define warning cnb_stub_warning
#define R_NilValue ((SEXP)0)
#define isNull(x) ((x) == R_NilValue)
#define isMatrix(x) ((x) != R_NilValue && (x)->ndim == 2)
#define PROTECT(x) (x)
#define UNPROTECT(n) ((void)(n))
#define AS_INTEGER(x) cnb_stub_as_integer(x)
#define GET_DIM(x) cnb_stub_get_dim(x)
#define INTEGER(x) ((int *)(x)->data)
#define INTEGER_POINTER(x) ((int *)(x)->data)
#define NUMERIC_POINTER(x) ((double *)(x)->data)
#define NEW_NUMERIC(n) cnb_stub_alloc(14, (n))
#define NEW_INTEGER(n) cnb_stub_alloc(13, (n))
#define AS_NUMERIC(x) (x)                     /* the drivers hand over REALSXP already */
#define R_NaInt INT_MIN
#define R_IsNA(x) cnb_stub_is_na((double)(x))
#define R_NegInf (-HUGE_VAL)
/* the rest of the API rcatnet.cpp touches (S4 slots, lists, strings: RCatnet(SEXP), genRcatnet, prob_string ...) is only
 * DECLARED: the drivers never reach those functions, and the definitions in ref_samples.cpp abort if they ever do */
#define VECSXP 19
#define STRSXP 16
#ifdef __cplusplus
extern "C" {
#endif
int cnb_stub_is_na(double x);
int length(SEXP x);
SEXP VECTOR_ELT(SEXP x, long i);
SEXP SET_VECTOR_ELT(SEXP x, long i, SEXP v);
SEXP allocVector(unsigned int type, long n);
const char *CHAR(SEXP x);
SEXP AS_LIST(SEXP x);
SEXP AS_CHARACTER(SEXP x);
SEXP mkChar(const char *s);
SEXP mkString(const char *s);
SEXP install(const char *s);
SEXP asChar(SEXP x);
SEXP STRING_ELT(SEXP x, long i);
void SET_STRING_ELT(SEXP x, long i, SEXP v);
int IS_VECTOR(SEXP x);
int isS4(SEXP x);
SEXP SET_SLOT(SEXP x, SEXP what, SEXP v);
SEXP GET_SLOT(SEXP x, SEXP what);
SEXP NEW_OBJECT(SEXP cls);
SEXP MAKE_CLASS(const char *name);
#ifdef __cplusplus
}
#endif
#endif
