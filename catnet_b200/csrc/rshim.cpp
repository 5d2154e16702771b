/* rshim.cpp -- the R-facing side of the drop-in: a catnet.so replacement's .Call entry points for the search path.
 *
 * Exports what R's useDynLib(catnet) resolves for this path (reference src/ccnInit.c:29-52, prototypes
 * src/catnet_rexport.h:31-64):
 *     ccnOptimalNetsForOrder/12   ccnOptimalNetsSA/23   ccnParHistogram/14   ccnReleaseCache/0   ccnSetSeed/1
 *     ccnEntropyPairwise/2   ccnEntropyOrder/2   ccnKLPairwise/2   ccnPearsonPairwise/2
 *     ccnLoglik/4   ccnNodeLoglik/4   ccnSetProb/3   ccnSamples/4
 * and registers them in R_init_catnet.  Every other routine of the reference's table keeps its reference
 * implementation (this shim is linked INTO the package next to the reference sources).
 *
 * SEXP <-> C ABI only: all computation happens behind include/catnet_b200.h.  R API calls (coercion, allocation,
 * Rf_error) stay on the calling thread.  Rf_error and (under options(warn = 2)) Rf_warning leave by longjmp, which skips
 * C++ destructors: so no routine raises either while its std::vectors are alive -- problems found on the way are thrown
 * as a C++ exception or noted, the routine's body unwinds normally, and `guarded` raises the R error / warnings
 * afterwards with nothing of the call left on the C++ side.  What R itself may still do in the middle (an allocation that
 * fails) is covered for the library's result by an external pointer with a finalizer (export_result).
 *
 * This file needs R's headers.  It is NOT part of libcatnet_b200.so; the image used to develop it has no R, so there
 * it is compiled against tests/rstub and run against the miniature R API of tests/rstub/minir.cpp
 * (tests/test_rshim.py on the CPU with the C ABI answered by the oracle, tests/test_zz_rshim_gpu.py over the library).
 */
#include <R.h>
#include <Rinternals.h>
#include <Rdefines.h>
#include <R_ext/Rdynload.h>
#include <float.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <vector>
#include "../../include/catnet_b200.h"

namespace {

/* ---- errors and warnings without longjmp over live C++ objects (see the header) ---- */
struct ShimError {};
thread_local char g_fail[640];
thread_local char g_warn[4][160];
thread_local int g_nwarn;

[[noreturn]] void shim_fail(const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_fail, sizeof(g_fail), fmt, ap);
	va_end(ap);
	throw ShimError();
}
void shim_warn(const char *msg) {
	if (g_nwarn < 4) {
		strncpy(g_warn[g_nwarn], msg, sizeof(g_warn[0]) - 1);
		g_warn[g_nwarn][sizeof(g_warn[0]) - 1] = 0;
		g_nwarn++;
	}
}
/* run the body of a .Call routine; raise what it noted once its frame is gone */
template <class Body>
SEXP guarded(Body body) {
	g_fail[0] = 0;
	g_nwarn = 0;
	SEXP out = R_NilValue;
	try {
		out = body();
	} catch (const ShimError &) {
		out = R_NilValue;
	}
	if (g_fail[0]) error("%s", g_fail);            /* R restores its protection stack when it unwinds to the caller */
	if (g_nwarn) {
		PROTECT(out);
		for (int i = 0; i < g_nwarn; i++) warning("%s", g_warn[i]);
		UNPROTECT(1);
	}
	return out;
}
/* elements of an N x M matrix as R counts them; refuses what an int-indexed buffer cannot address */
R_xlen_t cells(int rows, int cols) {
	const R_xlen_t n = (R_xlen_t)rows * (R_xlen_t)cols;
	if (rows < 0 || cols < 0 || n > (R_xlen_t)2147483647) shim_fail("matrix of %d x %d is too large", rows, cols);
	return n;
}
/* a fresh value into a slot: protected across install(), which may allocate */
void set_slot(SEXP obj, const char *name, SEXP value) {
	PROTECT(value);
	SET_SLOT(obj, install(name), value);
	UNPROTECT(1);
}

/* list (length N) of integer vectors, or R_NilValue -> [N][N] rows of 1-based labels, 0 padded */
bool pool_matrix(SEXP rPool, int N, std::vector<int32_t> &out) {
	if (isNull(rPool) || length(rPool) != N) return false;
	out.assign((size_t)N * N, 0);
	for (int i = 0; i < N; i++) {
		SEXP el = PROTECT(coerceVector(VECTOR_ELT(rPool, i), INTSXP));
		int len = length(el);
		if (len > 0 && len <= N)
			for (int j = 0; j < len; j++) out[(size_t)i * N + j] = INTEGER(el)[j];
		UNPROTECT(1);
	}
	return true;
}

/* R's uniform generator behind the C ABI's function-pointer form (cnb_sa_params.unif / cnb_hist_params.unif) */
double r_unif(void *) { return unif_rand(); }

struct Marshalled {
	cnb_params p;
	std::vector<int32_t> parent_sizes, order, num_cats, pool, fixed;
	int nprotect;
};

/* the 12 arguments of ccnOptimalNetsForOrder (catnet_rexport.h:36-40) -> cnb_params */
void marshal(Marshalled &m, SEXP rSamples, SEXP rPerturbations, SEXP rMaxParents, SEXP rParentSizes,
		SEXP rMaxComplexity, SEXP rOrder, SEXP rNodeCats, SEXP rParentsPool, SEXP rFixedParentsPool,
		SEXP rMatEdgeLiks, SEXP rEcho) {
	memset(&m.p, 0, sizeof(m.p));
	m.nprotect = 0;
	if (!isMatrix(rSamples)) shim_fail("Data is not a matrix");                 /* rcatnet_search.cpp:69-70 */
	rSamples = PROTECT(coerceVector(rSamples, INTSXP));
	m.nprotect++;
	SEXP dim = getAttrib(rSamples, R_DimSymbol);
	const int N = INTEGER(dim)[0], M = INTEGER(dim)[1];
	m.p.num_nodes = N;
	m.p.num_samples = M;
	m.p.samples = INTEGER(rSamples);                /* N x M column-major == [sample][node], NA_INTEGER == INT_MIN */
	if (!isNull(rPerturbations)) {
		rPerturbations = PROTECT(coerceVector(rPerturbations, INTSXP));
		m.nprotect++;
		m.p.perturbations = INTEGER(rPerturbations);
	}
	m.p.max_parent_set = asInteger(rMaxParents);
	m.p.max_complexity = asInteger(rMaxComplexity);
	m.p.echo = asLogical(rEcho);
	if (!isNull(rParentSizes) && length(rParentSizes) == N) {
		SEXP v = PROTECT(coerceVector(rParentSizes, INTSXP));
		m.parent_sizes.assign(INTEGER(v), INTEGER(v) + N);
		UNPROTECT(1);
		m.p.parent_sizes = m.parent_sizes.data();
	}
	if (!isNull(rOrder) && length(rOrder) >= N) {
		SEXP v = PROTECT(coerceVector(rOrder, INTSXP));
		m.order.assign(INTEGER(v), INTEGER(v) + N);
		UNPROTECT(1);
		m.p.order = m.order.data();
	} else if (!isNull(rOrder)) {
		shim_warn("Invalid nodeOrder parameter - reset to default node order.");   /* rcatnet_search.cpp:112-116 */
	}
	if (!isNull(rNodeCats) && length(rNodeCats) == N) {                     /* catIndices = 1..C per node */
		m.num_cats.resize(N);
		for (int i = 0; i < N; i++) m.num_cats[i] = length(VECTOR_ELT(rNodeCats, i));
		m.p.num_cats = m.num_cats.data();
	}
	if (pool_matrix(rParentsPool, N, m.pool)) m.p.parents_pool = m.pool.data();
	if (pool_matrix(rFixedParentsPool, N, m.fixed)) m.p.fixed_parents = m.fixed.data();
	if (!isNull(rMatEdgeLiks) && (R_xlen_t)length(rMatEdgeLiks) == cells(N, N)) {
		rMatEdgeLiks = PROTECT(coerceVector(rMatEdgeLiks, REALSXP));
		m.nprotect++;
		m.p.edge_prob = REAL(rMatEdgeLiks);           /* R's own column-major buffer, exactly what the ABI expects */
	}
}

/* nested probability lists, first listed parent outermost (rcatnet.cpp:318-345) */
SEXP prob_list(const double *probs, const int *pc, int d, int cc, int level, int offset) {
	if (level >= d) {
		SEXP v = PROTECT(allocVector(REALSXP, cc));
		memcpy(REAL(v), probs + (size_t)offset * cc, cc * sizeof(double));
		UNPROTECT(1);
		return v;
	}
	SEXP l = PROTECT(allocVector(VECSXP, pc[level]));
	for (int j = 0; j < pc[level]; j++)
		SET_VECTOR_ELT(l, j, prob_list(probs, pc, d, cc, level + 1, offset * pc[level] + j));
	UNPROTECT(1);
	return l;
}

/* one S4 catNetwork with the slots RCatnet::genRcatnet fills (rcatnet.cpp:174-316) */
SEXP make_catnetwork(const cnb_result *r, int i, const int32_t *cats_pos, SEXP rNodeNames) {
	const int N = cnb_result_num_nodes(r);
	std::vector<int32_t> order(N), npar(N), pars((size_t)N * 8);
	cnb_result_order(r, i, order.data());
	cnb_result_parents(r, i, npar.data(), pars.data(), 8);
	int maxpar = 0, maxcat = 0, complexity = 0;
	double loglik = 0;
	cnb_result_network(r, i, &complexity, &loglik);
	for (int n = 0; n < N; n++) {
		if (npar[n] > maxpar) maxpar = npar[n];
		if (cats_pos[n] > maxcat) maxcat = cats_pos[n];
	}
	SEXP cnet = PROTECT(NEW_OBJECT(MAKE_CLASS("catNetwork")));
	set_slot(cnet, "objectName", mkString("catNetwork"));
	set_slot(cnet, "numnodes", ScalarInteger(N));
	SEXP names = PROTECT(allocVector(STRSXP, N));
	char buf[64];
	for (int n = 0; n < N; n++) {
		if (!isNull(rNodeNames) && length(rNodeNames) == N) {
			SET_STRING_ELT(names, n, STRING_ELT(rNodeNames, order[n] - 1));        /* rcatnet_sa.cpp:902 */
		} else {
			snprintf(buf, sizeof(buf), "N%d", (int)order[n]);                      /* setNodesOrder, catnet_class.h:290 */
			SET_STRING_ELT(names, n, mkChar(buf));
		}
	}
	SET_SLOT(cnet, install("nodes"), names);
	set_slot(cnet, "maxParents", ScalarInteger(maxpar));
	SEXP plist = PROTECT(allocVector(VECSXP, N));
	for (int n = 0; n < N; n++) {
		if (npar[n] <= 0) continue;                                                /* NULL entry */
		SEXP pv = PROTECT(allocVector(INTSXP, npar[n]));
		for (int k = 0; k < npar[n]; k++) INTEGER(pv)[k] = pars[(size_t)n * 8 + k] + 1;   /* 1-based positions */
		SET_VECTOR_ELT(plist, n, pv);
		UNPROTECT(1);
	}
	SET_SLOT(cnet, install("parents"), plist);
	set_slot(cnet, "maxCategories", ScalarInteger(maxcat));
	SEXP clist = PROTECT(allocVector(VECSXP, N));
	for (int n = 0; n < N; n++) {
		SEXP cv = PROTECT(allocVector(STRSXP, cats_pos[n]));
		for (int k = 0; k < cats_pos[n]; k++) {
			snprintf(buf, sizeof(buf), "%d", k + 1);                               /* catIndices 1..C */
			SET_STRING_ELT(cv, k, mkChar(buf));
		}
		SET_VECTOR_ELT(clist, n, cv);
		UNPROTECT(1);
	}
	SET_SLOT(cnet, install("categories"), clist);
	SEXP problists = PROTECT(allocVector(VECSXP, N));
	SEXP sizes = PROTECT(allocVector(INTSXP, N));
	std::vector<double> probs;
	for (int n = 0; n < N; n++) {
		double ll, pr;
		int32_t ss = 0;
		int ts = cnb_result_node(r, i, n, &ll, &pr, &ss, NULL, 0);
		probs.assign(ts > 0 ? ts : 1, 0.0);
		cnb_result_node(r, i, n, &ll, &pr, &ss, probs.data(), ts);
		int pc[8];
		for (int k = 0; k < npar[n]; k++) pc[k] = cats_pos[pars[(size_t)n * 8 + k]];
		SET_VECTOR_ELT(problists, n, prob_list(probs.data(), pc, npar[n], cats_pos[n], 0, 0));
		INTEGER(sizes)[n] = ss;
	}
	SET_SLOT(cnet, install("probabilities"), problists);
	set_slot(cnet, "meta", mkString("catNetwork object"));
	set_slot(cnet, "complexity", ScalarInteger(complexity));
	set_slot(cnet, "likelihood", ScalarReal(loglik > -(double)FLT_MAX ? loglik : R_NegInf));   /* :295-300 */
	SET_SLOT(cnet, install("nodeSampleSizes"), sizes);
	UNPROTECT(6);
	return cnet;
}

/* The library's result outlives several R allocations (every network is an S4 object with nested lists), and any of them
 * may leave the shim by a non-local exit (allocation failure, a warning turned into an error by options(warn = 2)): the
 * result therefore hangs on an external pointer whose finalizer frees it -- R's collector releases what the shim could not. */
void result_finalizer(SEXP ptr) {
	cnb_result *r = (cnb_result *)R_ExternalPtrAddr(ptr);
	if (r) {
		cnb_result_free(r);
		R_ClearExternalPtr(ptr);
	}
}

SEXP export_result(cnb_result *r, const cnb_params &p, SEXP rNodeNames) {
	SEXP guard = PROTECT(R_MakeExternalPtr(r, R_NilValue, R_NilValue));
	R_RegisterCFinalizerEx(guard, result_finalizer, TRUE);
	const int N = p.num_nodes, K = cnb_result_num_networks(r);
	if (K < 1) {
		result_finalizer(guard);
		UNPROTECT(1);
		shim_warn("No networks are found");                                          /* rcatnet_search.cpp:279-282 */
		return R_NilValue;
	}
	std::vector<int32_t> cats_pos(N);
	SEXP out = PROTECT(allocVector(VECSXP, K));
	for (int i = 0; i < K; i++) {
		cnb_result_cats(r, i, cats_pos.data());
		SET_VECTOR_ELT(out, i, make_catnetwork(r, i, cats_pos.data(), rNodeNames));
	}
	result_finalizer(guard);                                                       /* released here in the normal course */
	UNPROTECT(2);
	return out;
}

[[noreturn]] void fail(int rc) {
	shim_fail("catnet (B200): %s [code %d]", cnb_last_error(), rc);
}

} /* namespace */

extern "C" {

/* catnetOptimalNetsForOrder (catnet_rexport.h:36-40); rUseCache changes nothing for ONE order: R releases the cache before
 * the call (R/catnet.search.R:204) and no key comes up twice inside an order (it also passes FALSE, :214) */
SEXP ccnOptimalNetsForOrder(SEXP rSamples, SEXP rPerturbations, SEXP rMaxParents, SEXP rParentSizes,
		SEXP rMaxComplexity, SEXP rOrder, SEXP rNodeCats, SEXP rParentsPool, SEXP rFixedParentsPool,
		SEXP rMatEdgeLiks, SEXP rUseCache, SEXP rEcho) {
	return guarded([&]() -> SEXP {
		(void)rUseCache;
		Marshalled m;
		marshal(m, rSamples, rPerturbations, rMaxParents, rParentSizes, rMaxComplexity, rOrder, rNodeCats, rParentsPool,
				rFixedParentsPool, rMatEdgeLiks, rEcho);
		cnb_result *res = NULL;
		int rc = cnb_search_order(&m.p, &res);
		UNPROTECT(m.nprotect);
		if (rc != CNB_OK) fail(rc);
		return export_result(res, m.p, R_NilValue);
	});
}

/* catnetOptimalNetsSA (catnet_rexport.h:42-48) */
SEXP ccnOptimalNetsSA(SEXP rNodeNames, SEXP rSamples, SEXP rPerturbations, SEXP rMaxParents, SEXP rParentSizes,
		SEXP rMaxComplexity, SEXP rNodeCats, SEXP rParentsPool, SEXP rFixedParentsPool, SEXP rMaxParentsPool,
		SEXP rMatEdgeLiks, SEXP rDirProbs, SEXP rModel, SEXP rStartOrder, SEXP rTempStart, SEXP rTempCoolFact,
		SEXP rTempCheckOrders, SEXP rMaxIter, SEXP rOrderShuffles, SEXP rStopDiff, SEXP rThreads, SEXP rUseCache,
		SEXP rEcho) {
	return guarded([&]() -> SEXP {
		(void)rMaxParentsPool;            /* R hard-codes maxParentsPool <- 0 (R/catnet.search.R:523) */
		Marshalled m;
		marshal(m, rSamples, rPerturbations, rMaxParents, rParentSizes, rMaxComplexity, R_NilValue, rNodeCats, rParentsPool,
				rFixedParentsPool, rMatEdgeLiks, rEcho);
		const int N = m.p.num_nodes;
		cnb_sa_params sa;
		memset(&sa, 0, sizeof(sa));
		const char *model = CHAR(STRING_ELT(PROTECT(coerceVector(rModel, STRSXP)), 0));
		sa.select_mode = !strncmp(model, "AIC", 3) ? 1 : (!strncmp(model, "BIC", 3) ? 2 : 0);   /* rcatnet_sa.cpp:323-327 */
		UNPROTECT(1);
		std::vector<int32_t> start(N);
		if (length(rStartOrder) < N) {
			shim_warn("Invalid nodeStartOrder parameter - reset to default node order.");
			for (int i = 0; i < N; i++) start[i] = i + 1;
		} else {
			SEXP v = PROTECT(coerceVector(rStartOrder, INTSXP));
			start.assign(INTEGER(v), INTEGER(v) + N);
			UNPROTECT(1);
		}
		sa.start_order = start.data();
		sa.temp_start = asReal(rTempStart);
		sa.temp_cool_fact = asReal(rTempCoolFact);
		sa.temp_check_orders = asInteger(rTempCheckOrders);
		sa.max_iter = asInteger(rMaxIter);
		sa.order_shuffles = asReal(rOrderShuffles);
		sa.stop_diff = asReal(rStopDiff);
		sa.num_threads = asInteger(rThreads);
		sa.use_cache = asLogical(rUseCache);
		int nprot = 0;
		if (!isNull(rDirProbs) && (R_xlen_t)length(rDirProbs) == cells(N, N)) {
			rDirProbs = PROTECT(coerceVector(rDirProbs, REALSXP));
			nprot++;
			sa.dir_prob = REAL(rDirProbs);
		}
		/* the chain draws from R's own generator on this thread, exactly where stock catnet does (rcatnet_sa.cpp:146,151,225,
		 * 745; utils.h:156,190): set.seed(s); cnSearchSA(...) proposes the same orders as the reference */
		sa.unif = r_unif;
		GetRNGstate();
		cnb_result *res = NULL;
		int rc = cnb_search_sa(&m.p, &sa, &res);
		PutRNGstate();
		UNPROTECT(m.nprotect + nprot);
		if (rc != CNB_OK) fail(rc);
		return export_result(res, m.p, rNodeNames);
	});
}

/* catnetParHistogram (catnet_rexport.h:49-54): cnSearchHist, R/catnet.histo.R:116-129 */
SEXP ccnParHistogram(SEXP rSamples, SEXP rPerturbations, SEXP rMaxParents, SEXP rParentSizes, SEXP rMaxComplexity,
		SEXP rNodeCats, SEXP rParentsPool, SEXP rFixedParentsPool, SEXP rScore, SEXP rWeight, SEXP rMaxIter,
		SEXP rThreads, SEXP rUseCache, SEXP rEcho) {
	return guarded([&]() -> SEXP {
		Marshalled m;
		marshal(m, rSamples, rPerturbations, rMaxParents, rParentSizes, rMaxComplexity, R_NilValue, rNodeCats, rParentsPool,
				rFixedParentsPool, R_NilValue, rEcho);
		const int N = m.p.num_nodes;
		cnb_hist_params hp;
		memset(&hp, 0, sizeof(hp));
		const char *score = CHAR(STRING_ELT(PROTECT(coerceVector(rScore, STRSXP)), 0));
		hp.score = !strncmp(score, "AIC", 3) ? 1 : (!strncmp(score, "BIC", 3) ? 2 : 0);        /* rcatnet_hist.cpp:212-216 */
		UNPROTECT(1);
		hp.weight = asInteger(rWeight);
		hp.max_iter = asInteger(rMaxIter);
		hp.num_threads = asInteger(rThreads);
		hp.use_cache = asLogical(rUseCache);
		hp.unif = r_unif;                                                              /* rcatnet_hist.cpp:121-159 */
		SEXP rmat = PROTECT(allocVector(REALSXP, cells(N, N)));                              /* :229-233 */
		GetRNGstate();
		int rc = cnb_search_hist(&m.p, &hp, REAL(rmat), NULL);
		PutRNGstate();
		UNPROTECT(m.nprotect + 1);
		if (rc != CNB_OK) fail(rc);
		return rmat;
	});
}

/* catnetEntropyPairwise / catnetKLpairwise / catnetPearsonPairwise / catnetEntropyOrder (catnet_rexport.h;
 * catnet_entropy.cpp:38-911), called from R/catnet.entropy.R:60-62, :94-96, :125-127 and R/catnet.chisq.R:27-29 */
static void pairwise_args(SEXP &rSamples, SEXP &rPerturbations, int &N, int &M, int &nprot) {
	if (!isMatrix(rSamples)) shim_fail("Data should be a matrix");                          /* catnet_entropy.cpp:47-48 */
	if (!isNull(rPerturbations) && !isMatrix(rPerturbations)) shim_fail("Perturbations should be a matrix");
	rSamples = PROTECT(coerceVector(rSamples, INTSXP));
	nprot = 1;
	SEXP dim = getAttrib(rSamples, R_DimSymbol);
	N = INTEGER(dim)[0];
	M = INTEGER(dim)[1];
	if (!isNull(rPerturbations)) {
		rPerturbations = PROTECT(coerceVector(rPerturbations, INTSXP));
		nprot++;
	}
}

static SEXP pairwise_call(int kind, SEXP rSamples, SEXP rPerturbations) {
	int N, M, nprot;
	pairwise_args(rSamples, rPerturbations, N, M, nprot);
	SEXP rvec = PROTECT(allocVector(REALSXP, cells(N, N)));
	int rc = cnb_pairwise(kind, N, M, INTEGER(rSamples), isNull(rPerturbations) ? NULL : INTEGER(rPerturbations), 0,
			REAL(rvec));
	UNPROTECT(nprot + 1);
	if (rc != CNB_OK) fail(rc);
	return rvec;
}

SEXP ccnEntropyPairwise(SEXP rSamples, SEXP rPerturbations) {
	return guarded([&]() -> SEXP {
		return pairwise_call(CNB_PAIRWISE_ENTROPY, rSamples, rPerturbations);
	});
}

SEXP ccnKLPairwise(SEXP rSamples, SEXP rPerturbations) {
	return guarded([&]() -> SEXP {
		return pairwise_call(CNB_PAIRWISE_KL, rSamples, rPerturbations);
	});
}

SEXP ccnPearsonPairwise(SEXP rSamples, SEXP rPerturbations) {
	return guarded([&]() -> SEXP {
		return pairwise_call(CNB_PAIRWISE_PEARSON, rSamples, rPerturbations);
	});
}

SEXP ccnEntropyOrder(SEXP rSamples, SEXP rPerturbations) {
	return guarded([&]() -> SEXP {
		int N, M, nprot;
		pairwise_args(rSamples, rPerturbations, N, M, nprot);
		SEXP rvec = PROTECT(allocVector(INTSXP, N));
		int rc = cnb_entropy_order(N, M, INTEGER(rSamples), isNull(rPerturbations) ? NULL : INTEGER(rPerturbations), 0,
				INTEGER(rvec));
		UNPROTECT(nprot + 1);
		if (rc != CNB_OK) fail(rc);
		return rvec;
	});
}

/* ---- a given catNetwork against data: ccnLoglik, ccnNodeLoglik, ccnSetProb (src/ccnInit.c:32-34) ----
 * The reference's glue for these is absent from the snapshot; the object is read the way RCatnet::RCatnet(SEXP) reads
 * it (rcatnet.cpp:38-172): parents 1-based in listed order, categories by length, nested probability lists flattened
 * first-parent-outermost (gen_prob_vector). */
struct RNetwork {
	cnb_network net;
	std::vector<int32_t> cats, npar, pars;
	std::vector<double> probs;
	std::vector<int64_t> off;
};

static void flatten_probs(SEXP node, const int *pc, int d, int level, int cc, std::vector<double> &out) {
	if (level >= d) {
		SEXP v = PROTECT(coerceVector(node, REALSXP));
		for (int k = 0; k < cc; k++) out.push_back(k < length(v) ? REAL(v)[k] : 0.0);
		UNPROTECT(1);
		return;
	}
	for (int j = 0; j < pc[level]; j++) {
		if (!isNewList(node) || j >= length(node)) shim_fail("probabilities do not match the parents' categories");
		flatten_probs(VECTOR_ELT(node, j), pc, d, level + 1, cc, out);
	}
}

static void read_network(SEXP cnet, RNetwork &r, bool want_probs) {
	if (!isS4(cnet)) shim_fail("cnet should be a catNetwork");
	const int N = INTEGER(GET_SLOT(cnet, install("numnodes")))[0];
	SEXP rparents = GET_SLOT(cnet, install("parents")), rcats = GET_SLOT(cnet, install("categories"));
	SEXP rprobs = GET_SLOT(cnet, install("probabilities"));
	if (length(rparents) != N || length(rcats) != N || (want_probs && length(rprobs) != N)) shim_fail("invalid catNetwork");
	int maxp = 1;
	for (int i = 0; i < N; i++) maxp = std::max(maxp, (int)length(VECTOR_ELT(rparents, i)));
	r.cats.resize(N);
	r.npar.resize(N);
	r.pars.assign((size_t)N * maxp, 0);
	r.off.assign(N + 1, 0);
	r.probs.clear();
	for (int i = 0; i < N; i++) r.cats[i] = length(VECTOR_ELT(rcats, i));
	for (int i = 0; i < N; i++) {
		SEXP pv = PROTECT(coerceVector(VECTOR_ELT(rparents, i), INTSXP));
		const int d = length(pv);
		r.npar[i] = d;
		int pc[64];
		int64_t size = r.cats[i];
		if (d > 8) shim_fail("more than 8 parents");
		for (int k = 0; k < d; k++) {
			const int q = INTEGER(pv)[k];
			if (q < 1 || q > N) shim_fail("invalid parent");
			r.pars[(size_t)i * maxp + k] = q;
			pc[k] = r.cats[q - 1];
			size *= pc[k];
		}
		UNPROTECT(1);
		r.off[i + 1] = r.off[i] + size;
		if (want_probs) flatten_probs(VECTOR_ELT(rprobs, i), pc, d, 0, r.cats[i], r.probs);
	}
	memset(&r.net, 0, sizeof(r.net));
	r.net.num_nodes = N;
	r.net.max_parents = maxp;
	r.net.num_cats = r.cats.data();
	r.net.num_parents = r.npar.data();
	r.net.parents = r.pars.data();
	r.net.probs = want_probs ? r.probs.data() : NULL;
	r.net.prob_offsets = r.off.data();
}

static void network_data(SEXP &rSamples, SEXP &rPerturbations, int N, int &M, int &nprot) {
	if (!isMatrix(rSamples)) shim_fail("'data' should be a matrix or data frame of categories");   /* R/catnet.loglik.R:57-58 */
	rSamples = PROTECT(coerceVector(rSamples, INTSXP));
	nprot = 1;
	SEXP dim = getAttrib(rSamples, R_DimSymbol);
	if (INTEGER(dim)[0] != N) shim_fail("The number of nodes in  the object and data should be equal");   /* :69-70 */
	M = INTEGER(dim)[1];
	if (!isNull(rPerturbations)) {
		rPerturbations = PROTECT(coerceVector(rPerturbations, INTSXP));
		nprot++;
	}
}

/* -FLT_MAX is the library's "minus infinity" (rcatnet.cpp:295-300 exports it as -Inf) */
static void export_neg_inf(double *v, int n) {
	for (int k = 0; k < n; k++)
		if (!(v[k] > -(double)FLT_MAX)) v[k] = R_NegInf;
}

SEXP ccnLoglik(SEXP cnet, SEXP rSamples, SEXP rPerturbations, SEXP rBySample) {
	return guarded([&]() -> SEXP {
		RNetwork r;
		read_network(cnet, r, true);
		int M, nprot;
		network_data(rSamples, rPerturbations, r.net.num_nodes, M, nprot);
		const int bysample = asLogical(rBySample) == 1;
		const int nout = bysample ? M : r.net.num_nodes;
		SEXP rvec = PROTECT(allocVector(REALSXP, nout));
		int rc = cnb_loglik(&r.net, M, INTEGER(rSamples), isNull(rPerturbations) ? NULL : INTEGER(rPerturbations),
				bysample ? CNB_LOGLIK_BYSAMPLE : CNB_LOGLIK_NODES, 0, REAL(rvec));
		if (rc == CNB_OK) export_neg_inf(REAL(rvec), nout);
		UNPROTECT(nprot + 1);
		if (rc != CNB_OK) fail(rc);
		return rvec;                                   /* R sums the per-node values (R/catnet.loglik.R:103-105) */
	});
}

SEXP ccnNodeLoglik(SEXP cnet, SEXP rNodes, SEXP rSamples, SEXP rPerturbations) {
	return guarded([&]() -> SEXP {
		RNetwork r;
		read_network(cnet, r, true);
		int M, nprot;
		network_data(rSamples, rPerturbations, r.net.num_nodes, M, nprot);
		SEXP nodes = PROTECT(coerceVector(rNodes, INTSXP));
		const int nsel = length(nodes);
		SEXP rvec = PROTECT(allocVector(REALSXP, nsel));
		int rc = cnb_node_loglik(&r.net, nsel, INTEGER(nodes), M, INTEGER(rSamples),
				isNull(rPerturbations) ? NULL : INTEGER(rPerturbations), 0, REAL(rvec));
		if (rc == CNB_OK) export_neg_inf(REAL(rvec), nsel);
		UNPROTECT(nprot + 2);
		if (rc != CNB_OK) fail(rc);
		return rvec;
	});
}

SEXP ccnSetProb(SEXP cnet, SEXP rSamples, SEXP rPerturbations) {
	return guarded([&]() -> SEXP {
		RNetwork r;
		read_network(cnet, r, false);
		const int N = r.net.num_nodes;
		int M, nprot;
		network_data(rSamples, rPerturbations, N, M, nprot);
		std::vector<double> probs((size_t)r.off[N]), ll(N);
		std::vector<int32_t> sizes(N);
		int rc = cnb_set_prob(&r.net, M, INTEGER(rSamples), isNull(rPerturbations) ? NULL : INTEGER(rPerturbations), 0,
				probs.data(), ll.data(), sizes.data());
		UNPROTECT(nprot);
		if (rc != CNB_OK) fail(rc);
		/* the object with its tables, likelihood and sample sizes replaced (R keeps nodes / categories, catnet.samples.R:427-436) */
		SEXP out = PROTECT(duplicate(cnet));
		SEXP problists = PROTECT(allocVector(VECSXP, N));
		SEXP rsizes = PROTECT(allocVector(INTSXP, N));
		double total = 0;
		bool neg_inf = false;
		for (int i = 0; i < N; i++) {
			int pc[8];
			for (int k = 0; k < r.npar[i]; k++) pc[k] = r.cats[r.pars[(size_t)i * r.net.max_parents + k] - 1];
			SET_VECTOR_ELT(problists, i, prob_list(probs.data() + r.off[i], pc, r.npar[i], r.cats[i], 0, 0));
			INTEGER(rsizes)[i] = sizes[i];
			if (!(ll[i] > -(double)FLT_MAX)) neg_inf = true;
			total += ll[i];                                /* CATNET::loglik: nodes in index order (catnet_class.h:469-485) */
		}
		SET_SLOT(out, install("probabilities"), problists);
		set_slot(out, "likelihood", ScalarReal(neg_inf ? R_NegInf : total));
		SET_SLOT(out, install("nodeSampleSizes"), rsizes);
		UNPROTECT(3);
		return out;
	});
}

/* cnSamples: .Call("ccnSamples", object, numsamples, perturbations, naRate) (R/catnet.samples.R:60-62) ->
 * RCatnet::genSamples (rcatnet.cpp:524-650).  R's RNG cannot cross the C ABI: one uniform from it seeds the library's
 * stream (set.seed() still makes runs reproducible; the stream differs from stock catnet). */
SEXP ccnSamples(SEXP cnet, SEXP rNumSamples, SEXP rPerturbations, SEXP rNaRate) {
	return guarded([&]() -> SEXP {
		RNetwork r;
		read_network(cnet, r, true);
		const int N = r.net.num_nodes, M = asInteger(rNumSamples);
		if (M < 1) shim_fail("The number of samples should be integer");
		int nprot = 0;
		const int32_t *pert = NULL;
		if (!isNull(rPerturbations)) {
			rPerturbations = PROTECT(coerceVector(rPerturbations, INTSXP));
			nprot++;
			if ((R_xlen_t)length(rPerturbations) == cells(N, M)) pert = INTEGER(rPerturbations);
		}
		GetRNGstate();
		const uint64_t seed = (uint64_t)(unif_rand() * 9007199254740992.0);
		PutRNGstate();
		SEXP rsamples = PROTECT(allocVector(INTSXP, cells(N, M)));           /* R reshapes it with matrix(data, ncol = numsamples) */
		int rc = cnb_samples(&r.net, M, pert, asReal(rNaRate), seed, 0, INTEGER(rsamples));
		UNPROTECT(nprot + 1);
		if (rc != CNB_OK) fail(rc);
		return rsamples;
	});
}

/* R calls this right BEFORE every search (R/catnet.search.R:204, 279) to drop the reference's memoised scores.  The
 * library has none; its idle device buffers are worth keeping across searches and go at unload (R_unload_catnet). */
SEXP ccnReleaseCache(void) {
	return R_NilValue;
}

SEXP ccnSetSeed(SEXP rSeed) {
	cnb_set_seed(asInteger(rSeed));
	return R_NilValue;
}

/* registration of the routines this shim owns; the package's ccnInit.c keeps registering the rest and calls this */
static const R_CallMethodDef cnb_call_methods[] = {
	{"ccnOptimalNetsForOrder", (DL_FUNC)&ccnOptimalNetsForOrder, 12},
	{"ccnOptimalNetsSA", (DL_FUNC)&ccnOptimalNetsSA, 23},
	{"ccnParHistogram", (DL_FUNC)&ccnParHistogram, 14},
	{"ccnReleaseCache", (DL_FUNC)&ccnReleaseCache, 0},
	{"ccnSetSeed", (DL_FUNC)&ccnSetSeed, 1},
	{"ccnEntropyPairwise", (DL_FUNC)&ccnEntropyPairwise, 2},
	{"ccnEntropyOrder", (DL_FUNC)&ccnEntropyOrder, 2},
	{"ccnKLPairwise", (DL_FUNC)&ccnKLPairwise, 2},
	{"ccnPearsonPairwise", (DL_FUNC)&ccnPearsonPairwise, 2},
	{"ccnLoglik", (DL_FUNC)&ccnLoglik, 4},
	{"ccnNodeLoglik", (DL_FUNC)&ccnNodeLoglik, 4},
	{"ccnSetProb", (DL_FUNC)&ccnSetProb, 3},
	{"ccnSamples", (DL_FUNC)&ccnSamples, 4},
	{NULL, NULL, 0}};

const R_CallMethodDef *cnb_r_call_methods(void) { return cnb_call_methods; }

#ifdef CNB_STANDALONE_RSHIM
/* build as a complete catnet.so for the search path only (ccnMarginalProb is then unavailable) */
void R_init_catnet(DllInfo *info) {
	R_registerRoutines(info, NULL, cnb_call_methods, NULL, NULL);
	R_useDynamicSymbols(info, (Rboolean)1);
}
void R_unload_catnet(DllInfo *info) {
	(void)info;
	cnb_release_cache();
}
#endif

} /* extern "C" */
