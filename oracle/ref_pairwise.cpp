/* ref_pairwise.cpp -- drives the reference's OWN pairwise functions (catnetEntropyPairwise, catnetEntropyOrder,
 * catnetKLpairwise, catnetPearsonPairwise; /root/reference/src/catnet_entropy.cpp:38-911, compiled where it lies,
 * unmodified) from plain C arrays, through the miniature R API of oracle/stub_sexp/R.h.
 * TEST INFRASTRUCTURE (oracle/_ref/libcatnet_ref_pairwise.so): never linked into the product. */
#include <R.h>
#include <vector>

extern "C" {
SEXP catnetEntropyPairwise(SEXP rSamples, SEXP rPerturbations);
SEXP catnetEntropyOrder(SEXP rSamples, SEXP rPerturbations);
SEXP catnetKLpairwise(SEXP rSamples, SEXP rPerturbations);
SEXP catnetPearsonPairwise(SEXP rSamples, SEXP rPerturbations);
}

size_t g_memcounter = 0; /* declared extern in reference utils.cpp:35 */

static std::vector<SEXP> g_all;
static char g_last_error[512];
static int g_error_count = 0;

extern "C" {
void cnb_stub_error(const char *fmt, ...) {
	va_list ap; va_start(ap, fmt);
	vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
	va_end(ap);
	g_error_count++;
}
void cnb_stub_warning(const char *fmt, ...) { (void)fmt; }
void Rprintf(const char *fmt, ...) { (void)fmt; }
void GetRNGstate(void) {}
void PutRNGstate(void) {}
double unif_rand(void) { return 0.5; }
double norm_rand(void) { return 0.0; }

SEXP cnb_stub_alloc(int type, int n) {
	SEXP s = (SEXP)calloc(1, sizeof(*s));
	s->type = type;
	s->n = n;
	s->data = calloc(n > 0 ? n : 1, type == 14 ? sizeof(double) : sizeof(int));
	g_all.push_back(s);
	return s;
}
SEXP cnb_stub_as_integer(SEXP x) { return x; }   /* inputs are integer already; NULL stays NULL (see stub_sexp/R.h) */
SEXP cnb_stub_get_dim(SEXP x) {
	if (!x || x->ndim != 2) return R_NilValue;
	if (!x->dim_sexp) {
		x->dim_sexp = cnb_stub_alloc(13, 2);
		((int *)x->dim_sexp->data)[0] = x->dim[0];
		((int *)x->dim_sexp->data)[1] = x->dim[1];
	}
	return x->dim_sexp;
}
void cnb_stub_free_all(void) {
	for (SEXP s : g_all) { free(s->data); free(s); }
	g_all.clear();
}

static SEXP int_matrix(const int *v, int N, int M) {
	if (!v) return R_NilValue;
	SEXP s = cnb_stub_alloc(13, N * M);
	s->ndim = 2;
	s->dim[0] = N;
	s->dim[1] = M;
	memcpy(s->data, v, (size_t)N * M * sizeof(int));   /* the reference recodes its input in place */
	return s;
}

/* kind 0 entropy, 1 KL, 2 Pearson: out[N*N] as the reference returns it (mat[n2*N + n1], i.e. R's matrix(mat, N, N)
 * has [n1, n2]); kind 3 entropy order: order_out[N], 1-based.  samples/pert are R's N x M column-major matrices,
 * i.e. [M][N] in C; samples 1-based categories, no NA.  Returns 0, or -1 if the reference raised an R error. */
int ref_pairwise(int kind, int N, int M, const int *samples, const int *pert, double *out, int *order_out) {
	g_error_count = 0;
	g_last_error[0] = 0;
	SEXP s = int_matrix(samples, N, M), p = int_matrix(pert, N, M), r = R_NilValue;
	switch (kind) {
	case 0: r = catnetEntropyPairwise(s, p); break;
	case 1: r = catnetKLpairwise(s, p); break;
	case 2: r = catnetPearsonPairwise(s, p); break;
	case 3: r = catnetEntropyOrder(s, p); break;
	default: break;
	}
	int rc = (g_error_count || r == R_NilValue) ? -1 : 0;
	if (rc == 0) {
		if (kind == 3) memcpy(order_out, r->data, (size_t)N * sizeof(int));
		else memcpy(out, r->data, (size_t)N * N * sizeof(double));
	}
	cnb_stub_free_all();
	return rc;
}
const char *ref_pairwise_last_error(void) { return g_last_error; }
}
