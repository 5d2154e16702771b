/* tests/rstub/minir.cpp -- a miniature, FUNCTIONAL implementation of the part of R's C API that
 * catnet_b200/csrc/rshim.cpp uses (declared in tests/rstub/Rinternals.h), so that the shim can be compiled, linked and
 * executed in an image without R (tests/test_rshim.py).  Test infrastructure only.
 *
 * Values are reference-counted nowhere and freed nowhere: a test process builds a few small objects and exits.
 * Rf_error throws a C++ exception that mr_call catches (R would longjmp to its top level); Rf_warning records the
 * message.  `mr_*` functions are the handles the Python test uses to build arguments and to read results. */
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>
#include "Rinternals.h"
#include "R_ext/Rdynload.h"

#define MR_NILSXP 0
#define MR_SYMSXP 1
#define MR_CHARSXP 9
#define MR_LGLSXP 10
#define MR_S4SXP 25

struct SEXPREC {
	int type = MR_NILSXP;
	std::vector<int> ints;                 /* LGLSXP, INTSXP */
	std::vector<double> reals;             /* REALSXP */
	std::vector<SEXP> elts;                /* VECSXP, STRSXP (of CHARSXP) */
	std::string str;                       /* CHARSXP, SYMSXP, class name of an S4 object */
	std::map<std::string, SEXP> attrs;     /* "dim", and the slots of an S4 object */
	void *ext = nullptr;                   /* EXTPTRSXP: the address, and its C finalizer */
	void (*fin)(SEXP) = nullptr;
};

static SEXPREC g_nil;
static std::string g_error, g_warnings;
static std::vector<SEXP> g_extptrs;         /* every external pointer made: mr_run_finalizers plays the garbage collector */
static long g_alloc_countdown = -1;        /* mr_fail_alloc_after: the n-th allocVector from now raises an R error */
static unsigned long long g_rng = 0x9e3779b97f4a7c15ull;

extern "C" {

SEXP R_NilValue = &g_nil;
SEXP R_DimSymbol = NULL;
double R_NegInf = -HUGE_VAL;

static SEXP mk(int type) {
	SEXP s = new SEXPREC;
	s->type = type;
	return s;
}
static SEXP dim_symbol(void) {
	if (!R_DimSymbol) { R_DimSymbol = mk(MR_SYMSXP); R_DimSymbol->str = "dim"; }
	return R_DimSymbol;
}

SEXP Rf_protect(SEXP s) { return s; }
void Rf_unprotect(int) {}
SEXP Rf_install(const char *name) {
	if (!strcmp(name, "dim")) return dim_symbol();
	SEXP s = mk(MR_SYMSXP);
	s->str = name;
	return s;
}
SEXP Rf_mkChar(const char *s) { SEXP c = mk(MR_CHARSXP); c->str = s; return c; }
SEXP Rf_allocVector(unsigned int type, long n) {
	if (g_alloc_countdown >= 0 && g_alloc_countdown-- == 0) Rf_error("cannot allocate vector of size %ld", n);
	SEXP s = mk((int)type);
	if (type == INTSXP || type == MR_LGLSXP) s->ints.assign(n, 0);
	else if (type == REALSXP) s->reals.assign(n, 0.0);
	else if (type == VECSXP) s->elts.assign(n, R_NilValue);
	else if (type == STRSXP) s->elts.assign(n, Rf_mkChar(""));
	return s;
}
SEXP Rf_mkString(const char *s) { SEXP v = Rf_allocVector(STRSXP, 1); v->elts[0] = Rf_mkChar(s); return v; }
SEXP Rf_ScalarInteger(int x) { SEXP v = Rf_allocVector(INTSXP, 1); v->ints[0] = x; return v; }
SEXP Rf_ScalarReal(double x) { SEXP v = Rf_allocVector(REALSXP, 1); v->reals[0] = x; return v; }
int Rf_isNull(SEXP s) { return s == R_NilValue || s->type == MR_NILSXP; }
int Rf_length(SEXP s) {
	switch (s->type) {
	case INTSXP: case MR_LGLSXP: return (int)s->ints.size();
	case REALSXP: return (int)s->reals.size();
	case VECSXP: case STRSXP: return (int)s->elts.size();
	case MR_NILSXP: return 0;
	default: return 1;
	}
}
SEXP Rf_getAttrib(SEXP s, SEXP what) {
	auto it = s->attrs.find(what->str);
	return it == s->attrs.end() ? R_NilValue : it->second;
}
int Rf_isMatrix(SEXP s) {
	if (Rf_isNull(s)) return 0;
	SEXP d = Rf_getAttrib(s, dim_symbol());
	return !Rf_isNull(d) && Rf_length(d) == 2;
}
/* R's NA_INTEGER is INT_MIN, NA_REAL a NaN: as.integer(NA_real_) is NA_integer_ */
SEXP Rf_coerceVector(SEXP s, unsigned int type) {
	if ((unsigned)s->type == type) return s;
	SEXP out = mk((int)type);
	out->attrs = s->attrs;
	if (type == INTSXP && s->type == REALSXP) {
		for (double v : s->reals) out->ints.push_back(v != v ? (int)0x80000000 : (int)v);
	} else if (type == INTSXP && s->type == MR_LGLSXP) {
		out->ints = s->ints;
	} else if (type == REALSXP && (s->type == INTSXP || s->type == MR_LGLSXP)) {
		for (int v : s->ints) out->reals.push_back(v == (int)0x80000000 ? NAN : (double)v);
	} else if (type == INTSXP && s->type == MR_NILSXP) {
		/* integer(0) */
	} else if (type == STRSXP && s->type == STRSXP) {
		out->elts = s->elts;
	} else {
		throw std::runtime_error("minir: unsupported coercion");
	}
	return out;
}
int Rf_asInteger(SEXP s) {
	if (s->type == INTSXP || s->type == MR_LGLSXP) return s->ints.empty() ? (int)0x80000000 : s->ints[0];
	if (s->type == REALSXP) return s->reals.empty() || s->reals[0] != s->reals[0] ? (int)0x80000000 : (int)s->reals[0];
	return (int)0x80000000;
}
int Rf_asLogical(SEXP s) {
	int v = Rf_asInteger(s);
	return v == (int)0x80000000 ? v : (v != 0);
}
double Rf_asReal(SEXP s) {
	if (s->type == REALSXP) return s->reals.empty() ? NAN : s->reals[0];
	if (s->type == INTSXP || s->type == MR_LGLSXP) return s->ints.empty() || s->ints[0] == (int)0x80000000 ? NAN : (double)s->ints[0];
	return NAN;
}
int *INTEGER(SEXP s) {
	if (s->type != INTSXP && s->type != MR_LGLSXP) throw std::runtime_error("minir: INTEGER() of a non-integer");
	return s->ints.data();
}
double *REAL(SEXP s) {
	if (s->type != REALSXP) throw std::runtime_error("minir: REAL() of a non-numeric");
	return s->reals.data();
}
const char *CHAR(SEXP s) { return s->str.c_str(); }
SEXP VECTOR_ELT(SEXP s, long i) {
	if (s->type != VECSXP || i < 0 || i >= (long)s->elts.size()) throw std::runtime_error("minir: VECTOR_ELT out of range");
	return s->elts[i];
}
SEXP SET_VECTOR_ELT(SEXP s, long i, SEXP v) {
	if (s->type != VECSXP || i < 0 || i >= (long)s->elts.size()) throw std::runtime_error("minir: SET_VECTOR_ELT out of range");
	return s->elts[i] = v;
}
SEXP STRING_ELT(SEXP s, long i) {
	if (s->type != STRSXP || i < 0 || i >= (long)s->elts.size()) throw std::runtime_error("minir: STRING_ELT out of range");
	return s->elts[i];
}
void SET_STRING_ELT(SEXP s, long i, SEXP v) {
	if (s->type != STRSXP || i < 0 || i >= (long)s->elts.size()) throw std::runtime_error("minir: SET_STRING_ELT out of range");
	s->elts[i] = v;
}
SEXP R_do_MAKE_CLASS(const char *name) { SEXP c = mk(MR_SYMSXP); c->str = name; return c; }
SEXP R_do_new_object(SEXP cls) { SEXP o = mk(MR_S4SXP); o->str = cls->str; return o; }
SEXP R_do_slot_assign(SEXP x, SEXP what, SEXP v) { x->attrs[what->str] = v; return x; }
SEXP R_do_slot(SEXP x, SEXP what) {
	auto it = x->attrs.find(what->str);
	if (it == x->attrs.end()) throw std::runtime_error("minir: no slot " + what->str);
	return it->second;
}
SEXP Rf_duplicate(SEXP s) {
	if (Rf_isNull(s)) return s;
	SEXP c = new SEXPREC(*s);
	for (auto &e : c->elts) e = Rf_duplicate(e);
	for (auto &a : c->attrs) a.second = Rf_duplicate(a.second);
	return c;
}
int Rf_isNewList(SEXP s) { return s->type == VECSXP || s->type == MR_NILSXP; }
int IS_S4_OBJECT(SEXP s) { return s->type == MR_S4SXP; }

void Rf_error(const char *fmt, ...) {
	char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	throw std::runtime_error(std::string("Rf_error: ") + buf);
}
void Rf_warning(const char *fmt, ...) {
	char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	g_warnings += buf;
	g_warnings += "\n";
}
#define MR_EXTPTRSXP 22
SEXP R_MakeExternalPtr(void *p, SEXP, SEXP) { SEXP s = mk(MR_EXTPTRSXP); s->ext = p; g_extptrs.push_back(s); return s; }
void *R_ExternalPtrAddr(SEXP s) { return s->ext; }
void R_ClearExternalPtr(SEXP s) { s->ext = nullptr; }
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t f, Rboolean) { s->fin = f; }
void GetRNGstate(void) {}
void PutRNGstate(void) {}
double unif_rand(void) {                                 /* splitmix64, 53 bits */
	unsigned long long z = (g_rng += 0x9e3779b97f4a7c15ull);
	z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
	z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
	z ^= z >> 31;
	return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}
int R_registerRoutines(DllInfo *, const void *, const R_CallMethodDef *, const void *, const void *) { return 1; }
int R_useDynamicSymbols(DllInfo *, int) { return 1; }

/* ---- handles for the Python test ---- */
SEXP mr_nil(void) { return R_NilValue; }
SEXP mr_int_vector(const int *v, int n) { SEXP s = Rf_allocVector(INTSXP, n); if (n) memcpy(s->ints.data(), v, n * sizeof(int)); return s; }
SEXP mr_real_vector(const double *v, int n) { SEXP s = Rf_allocVector(REALSXP, n); if (n) memcpy(s->reals.data(), v, n * sizeof(double)); return s; }
SEXP mr_logical(int v) { SEXP s = Rf_allocVector(MR_LGLSXP, 1); s->ints[0] = v; return s; }
SEXP mr_string(const char *v) { return Rf_mkString(v); }
SEXP mr_list(int n) { return Rf_allocVector(VECSXP, n); }
void mr_list_set(SEXP l, int i, SEXP v) { SET_VECTOR_ELT(l, i, v); }
/* matrix(v, nrow, ncol): column-major */
SEXP mr_set_dim(SEXP s, int nrow, int ncol) {
	SEXP d = Rf_allocVector(INTSXP, 2);
	d->ints[0] = nrow;
	d->ints[1] = ncol;
	s->attrs["dim"] = d;
	return s;
}
SEXP mr_new_s4(const char *cls) { return R_do_new_object(R_do_MAKE_CLASS(cls)); }
void mr_set_slot(SEXP obj, const char *name, SEXP v) { obj->attrs[name] = v; }
/* c("1", ..., "n") */
SEXP mr_category_names(int n) {
	SEXP v = Rf_allocVector(STRSXP, n);
	for (int k = 0; k < n; k++) v->elts[k] = Rf_mkChar(std::to_string(k + 1).c_str());
	return v;
}
void mr_set_string(SEXP v, int i, const char *s) { SET_STRING_ELT(v, i, Rf_mkChar(s)); }
int mr_type(SEXP s) { return s->type; }
int mr_length(SEXP s) { return Rf_length(s); }
SEXP mr_elt(SEXP s, int i) { return s->elts.at(i); }
const char *mr_chars(SEXP s) { return s->str.c_str(); }
const int *mr_ints(SEXP s) { return s->ints.data(); }
const double *mr_reals(SEXP s) { return s->reals.data(); }
SEXP mr_slot(SEXP s, const char *name) {
	auto it = s->attrs.find(name);
	return it == s->attrs.end() ? NULL : it->second;
}
const char *mr_last_error(void) { return g_error.c_str(); }
const char *mr_warnings(void) { return g_warnings.c_str(); }
void mr_clear(void) { g_error.clear(); g_warnings.clear(); }
void mr_set_rng(unsigned long long s) { g_rng = s; }
/* the n-th allocVector from now fails like R's allocator does: with an R error, i.e. a non-local exit out of the shim */
void mr_fail_alloc_after(long n) { g_alloc_countdown = n; }
/* what R's garbage collector does with unreachable external pointers: run their finalizers.  Returns how many of them still
 * held an address (= resources the shim had not released itself when it was left) */
int mr_run_finalizers(void) {
	int held = 0;
	for (SEXP s : g_extptrs) {
		if (s->ext) held++;
		if (s->fin) s->fin(s);
	}
	g_extptrs.clear();
	return held;
}

/* .Call(name, args...) through the table the shim registers; NULL + mr_last_error() if it raised an R error */
const R_CallMethodDef *cnb_r_call_methods(void);
SEXP mr_call(const char *name, SEXP *a, int nargs) {
	g_error.clear();
	const R_CallMethodDef *m = cnb_r_call_methods();
	for (; m->name; m++)
		if (!strcmp(m->name, name)) break;
	if (!m->name) { g_error = std::string("no such routine: ") + name; return NULL; }
	if (m->numArgs != nargs) { g_error = "wrong number of arguments"; return NULL; }
	try {
		typedef SEXP (*F0)(void);
		typedef SEXP (*F1)(SEXP);
		typedef SEXP (*F2)(SEXP, SEXP);
		typedef SEXP (*F3)(SEXP, SEXP, SEXP);
		typedef SEXP (*F4)(SEXP, SEXP, SEXP, SEXP);
		typedef SEXP (*F12)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
		typedef SEXP (*F14)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
		typedef SEXP (*F23)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP,
				SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
		switch (nargs) {
		case 0: return ((F0)m->fun)();
		case 1: return ((F1)m->fun)(a[0]);
		case 2: return ((F2)m->fun)(a[0], a[1]);
		case 3: return ((F3)m->fun)(a[0], a[1], a[2]);
		case 4: return ((F4)m->fun)(a[0], a[1], a[2], a[3]);
		case 12: return ((F12)m->fun)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11]);
		case 14: return ((F14)m->fun)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], a[13]);
		case 23: return ((F23)m->fun)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], a[13],
				a[14], a[15], a[16], a[17], a[18], a[19], a[20], a[21], a[22]);
		}
		g_error = "unsupported arity";
		return NULL;
	} catch (const std::exception &e) {
		g_error = e.what();
		return NULL;
	}
}

} /* extern "C" */
