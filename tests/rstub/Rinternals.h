/* tests/rstub -- stand-in for R's C API, just enough for catnet_b200/csrc/rshim.cpp in an image without R:
 * tests/test_rshim.py syntax-checks the shim against these declarations and RUNS it against their miniature
 * implementation in minir.cpp (values, coercion, attributes, S4 slots, Rf_error as a C++ exception).  Test
 * infrastructure only. */
#ifndef CNB_RSTUB_RINTERNALS_H
#define CNB_RSTUB_RINTERNALS_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct SEXPREC *SEXP;
typedef int Rboolean;
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
extern SEXP R_NilValue, R_DimSymbol;
extern double R_NegInf;
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
SEXP Rf_coerceVector(SEXP, unsigned int);
typedef long R_xlen_t;                       /* R: ptrdiff_t */
SEXP Rf_allocVector(unsigned int, long);
SEXP Rf_getAttrib(SEXP, SEXP);
SEXP Rf_install(const char *);
SEXP Rf_mkChar(const char *);
SEXP Rf_mkString(const char *);
SEXP Rf_ScalarInteger(int);
SEXP Rf_ScalarReal(double);
int Rf_isNull(SEXP);
int Rf_isMatrix(SEXP);
int Rf_length(SEXP);
int Rf_asInteger(SEXP);
int Rf_asLogical(SEXP);
double Rf_asReal(SEXP);
int *INTEGER(SEXP);
double *REAL(SEXP);
const char *CHAR(SEXP);
SEXP VECTOR_ELT(SEXP, long);
SEXP SET_VECTOR_ELT(SEXP, long, SEXP);
SEXP STRING_ELT(SEXP, long);
void SET_STRING_ELT(SEXP, long, SEXP);
SEXP R_do_slot_assign(SEXP, SEXP, SEXP);
SEXP R_do_slot(SEXP, SEXP);
SEXP Rf_duplicate(SEXP);
int Rf_isNewList(SEXP);
int IS_S4_OBJECT(SEXP);
SEXP R_do_MAKE_CLASS(const char *);
SEXP R_do_new_object(SEXP);
typedef void (*R_CFinalizer_t)(SEXP);
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot);
void *R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean onexit);
void Rf_error(const char *, ...);
void Rf_warning(const char *, ...);
void GetRNGstate(void);
void PutRNGstate(void);
double unif_rand(void);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
#define coerceVector Rf_coerceVector
#define allocVector Rf_allocVector
#define getAttrib Rf_getAttrib
#define install Rf_install
#define mkChar Rf_mkChar
#define mkString Rf_mkString
#define ScalarInteger Rf_ScalarInteger
#define ScalarReal Rf_ScalarReal
#define isNull Rf_isNull
#define isMatrix Rf_isMatrix
#define length Rf_length
#define asInteger Rf_asInteger
#define asLogical Rf_asLogical
#define asReal Rf_asReal
#define duplicate Rf_duplicate
#define isNewList Rf_isNewList
#define isS4(x) IS_S4_OBJECT(x)
#define error Rf_error
#define warning Rf_warning
#ifdef __cplusplus
}
#endif
#endif
