/* ref_samples.cpp -- drives the reference's OWN sampler, RCatnet::genSamples (/root/reference/src/rcatnet.cpp:524-650,
 * compiled where it lies, unmodified), from plain C arrays through the miniature R API of oracle/stub_sexp/R.h.
 * R's unif_rand is replaced by the injected splitmix64 stream the product uses (catnet_b200/csrc/cnb_rng.h), so that
 * samples can be compared draw for draw.  TEST INFRASTRUCTURE (oracle/_ref/libcatnet_ref_samples.so). */
#include <R.h>
#include <stdint.h>
#include <vector>
#include "rcatnet.h"

size_t g_memcounter = 0; /* declared extern in reference utils.cpp:35 */

static std::vector<SEXP> g_all;
static uint64_t g_rng_state = 0;
static long g_rng_calls = 0;

extern "C" {
void cnb_stub_error(const char *fmt, ...) { (void)fmt; }
void cnb_stub_warning(const char *fmt, ...) { (void)fmt; }
void Rprintf(const char *fmt, ...) { (void)fmt; }
void GetRNGstate(void) {}
void PutRNGstate(void) {}
double norm_rand(void) { return 0.0; }
/* splitmix64 -> uniform in (0,1): ((x>>11)+0.5)*2^-53, the stream of ref_driver.cpp and cnb_rng.h */
double unif_rand(void) {
	uint64_t z = (g_rng_state += 0x9E3779B97F4A7C15ull);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	z = z ^ (z >> 31);
	g_rng_calls++;
	return ((double)(z >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

SEXP cnb_stub_alloc(int type, int n) {
	SEXP s = (SEXP)calloc(1, sizeof(*s));
	s->type = type;
	s->n = n;
	s->data = calloc(n > 0 ? n : 1, type == 14 ? sizeof(double) : sizeof(int));
	g_all.push_back(s);
	return s;
}
SEXP cnb_stub_as_integer(SEXP x) { return x; }
SEXP cnb_stub_get_dim(SEXP x) { (void)x; return R_NilValue; }
int cnb_stub_is_na(double x) { return x != x; }   /* an int is never NA_real_; NA_INTEGER fails the range test that follows */

/* S4 / list / string API: reached only by RCatnet(SEXP), genRcatnet, prob_string, which these drivers never call */
#define UNREACHED(name) do { fprintf(stderr, "oracle/ref_samples.cpp: " name " is not implemented in the R API miniature\n"); abort(); } while (0)
int length(SEXP x) { return x ? x->n : 0; }
SEXP VECTOR_ELT(SEXP, long) { UNREACHED("VECTOR_ELT"); }
SEXP SET_VECTOR_ELT(SEXP, long, SEXP) { UNREACHED("SET_VECTOR_ELT"); }
SEXP allocVector(unsigned int, long) { UNREACHED("allocVector"); }
const char *CHAR(SEXP) { UNREACHED("CHAR"); }
SEXP AS_LIST(SEXP) { UNREACHED("AS_LIST"); }
SEXP AS_CHARACTER(SEXP) { UNREACHED("AS_CHARACTER"); }
SEXP mkChar(const char *) { UNREACHED("mkChar"); }
SEXP mkString(const char *) { UNREACHED("mkString"); }
SEXP install(const char *) { UNREACHED("install"); }
SEXP asChar(SEXP) { UNREACHED("asChar"); }
SEXP STRING_ELT(SEXP, long) { UNREACHED("STRING_ELT"); }
void SET_STRING_ELT(SEXP, long, SEXP) { UNREACHED("SET_STRING_ELT"); }
int IS_VECTOR(SEXP) { UNREACHED("IS_VECTOR"); }
int isS4(SEXP) { UNREACHED("isS4"); }
SEXP SET_SLOT(SEXP, SEXP, SEXP) { UNREACHED("SET_SLOT"); }
SEXP GET_SLOT(SEXP, SEXP) { UNREACHED("GET_SLOT"); }
SEXP NEW_OBJECT(SEXP) { UNREACHED("NEW_OBJECT"); }
SEXP MAKE_CLASS(const char *) { UNREACHED("MAKE_CLASS"); }

/* cnSamples: out[M][N] 1-based categories (INT_MIN = NA) of the network (numCats, parents0 [N][maxParents] 0-based in
 * listed order, tables by offsets) drawn by RCatnet::genSamples with the uniform stream seeded by `seed`; pert [M][N]
 * (values in 1..C fix the node in that sample) or NULL; returns the number of uniforms consumed, < 0 on failure */
long ref_samples(int N, const int *numCats, int maxParents, const int *numParents, const int *parents0,
		const double *probs, const long long *offsets, int M, const int *pert, double naRate, unsigned long long seed,
		int *out) {
	int maxcats = 0, i;
	for (i = 0; i < N; i++)
		if (numCats[i] > maxcats) maxcats = numCats[i];
	RCatnet net;
	net.init(N, maxParents, maxcats, 0, 0, 0, numCats);
	for (i = 0; i < N; i++) {
		if (numParents[i] > 0) net.setParents(i, const_cast<int *>(parents0 + (size_t)i * maxParents), numParents[i]);
		net.setCondProb(i, const_cast<double *>(probs + offsets[i]), (int)(offsets[i + 1] - offsets[i]));
	}
	SEXP rnum = cnb_stub_alloc(13, 1), rna = cnb_stub_alloc(14, 1), rpert = R_NilValue;
	((int *)rnum->data)[0] = M;
	((double *)rna->data)[0] = naRate;
	if (pert) {
		rpert = cnb_stub_alloc(13, N * M);
		memcpy(rpert->data, pert, (size_t)N * M * sizeof(int));
	}
	g_rng_state = seed;
	g_rng_calls = 0;
	SEXP r = net.genSamples(rnum, rpert, rna);
	long calls = -1;
	if (r != R_NilValue && r->n == N * M) {
		memcpy(out, r->data, (size_t)N * M * sizeof(int));
		calls = g_rng_calls;
	}
	for (SEXP s : g_all) { free(s->data); free(s); }
	g_all.clear();
	return calls;
}
}
