/*
 * oracle/ref_driver.cpp  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * C-callable driver around the UNMODIFIED reference search core, compiled from the sources
 * where they lie under /root/reference/src (catnet_search2.h, catnet_class.h, problist.h,
 * search_params.h, cache.{h,cpp}, thread.h/pthread.cpp, utils.{h,cpp}) against the stub R
 * headers in oracle/stub/.  Output goes to oracle/_ref/libcatnet_ref.so (git-ignored).
 *
 * Three entry groups:
 *   ref_family_score   -> CATNET::setParents + CATNET::setNodeSampleProb  (catnet_class.h:393-433, 818-879)
 *   ref_search_order   -> CATNET_SEARCH2::estimate                         (catnet_search2.h:152-897)
 *                         fed the way RCatnetSearch::estimateCatnets feeds it (rcatnet_search.cpp:132-273)
 *   ref_search_sa      -> the SA loop of RCatnetSearchSA::search (rcatnet_sa.cpp:471-820) re-stated around
 *                         the reference's own estimate()/c_thread/c_cache, because the original is welded to
 *                         SEXP.  R's RNG is replaced by an injected, documented uniform stream (splitmix64).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 */
#include <stdint.h>
#include <vector>
#include "catnet_search2.h"

#ifndef MAX_NODE_NAME
#define MAX_NODE_NAME 16
#endif

typedef CATNET<char, MAX_NODE_NAME, double> RefNet;
typedef CATNET_SEARCH2<char, MAX_NODE_NAME, double> RefSearch;

size_t g_memcounter = 0; /* declared extern in reference utils.cpp:35 */
void ReleaseCache();       /* reference cache.cpp:62 */

/* ---------------- injected RNG + R API stand-ins ---------------- */
static uint64_t g_rng_state = 0x9E3779B97F4A7C15ull;
static long g_rng_calls = 0;
static int g_error_count = 0;
static char g_last_error[512];

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

/* splitmix64 -> uniform in (0,1): ((x>>11)+0.5)*2^-53.  The product's SA driver uses the same stream
 * (catnet_b200/csrc/cnb_rng.h) so that trajectories can be compared draw for draw. */
double unif_rand(void) {
	uint64_t z = (g_rng_state += 0x9E3779B97F4A7C15ull);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	z = z ^ (z >> 31);
	g_rng_calls++;
	return ((double)(z >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
double norm_rand(void) { return 0.0; }

void ref_set_seed(uint64_t seed) { g_rng_state = seed; g_rng_calls = 0; }
long ref_rng_calls(void) { return g_rng_calls; }
int ref_error_count(void) { return g_error_count; }
const char *ref_last_error(void) { return g_last_error; }
void ref_release_cache(void) { ReleaseCache(); }

/* ---------------- family-level hook ---------------- */
/* samples0: [M][N] zero-based category codes (anything outside [0,numCats[node]) is skipped, as NA is).
 * counts_out: table of size numCats[child]*prod(numCats[parents]) (reference layout, problist.h:143-155).
 * returns the table size, or <0 on error. */
int ref_family_score(const int *samples0, int N, int M, const int *numCats, int child,
		const int *parents, int d, double *counts_out, double *loglik_out, int *ncount_out) {
	int maxcats = 0, i;
	for (i = 0; i < N; i++)
		if (numCats[i] > maxcats) maxcats = numCats[i];
	RefNet net(N, d, maxcats, 0, 0, 0, numCats);
	std::vector<int> pars(parents, parents + d);
	/* d == 0: the constructor already made the parent-less table; setParents(.., 0) would free it and bail out
	 * on CATNET_MALLOC(0) (catnet_class.h:414-419), and estimate() never calls it that way (:387-389) */
	if (d > 0) net.setParents(child, &pars[0], d);
	double ll = net.setNodeSampleProb(child, const_cast<int*>(samples0), M);
	PROB_LIST<double> *pl = net.getNodeProb(child);
	if (!pl) return -1;
	if (counts_out)
		for (i = 0; i < pl->nProbSize; i++) counts_out[i] = pl->pProbs[i];
	if (loglik_out) *loglik_out = ll;
	if (ncount_out) *ncount_out = pl->sampleSize;
	return pl->nProbSize;
}

/* ---------------- result handle ---------------- */
struct RefResult {
	int nslots;
	RefNet **nets;  /* owned */
	int N;
	std::vector<int> order; /* SA: accepted order (1-based labels) */
	int niter, naccept;
	std::vector<double> trace; /* SA: fLogLik of every evaluated drive, in evaluation order */
	std::vector<int> trace_accept;
};

static void fill_params(SEARCH_PARAMETERS *sp, int N, int M,
		const int *samples_lbl, const int *pert_lbl, const int *order /*1-based labels per position*/,
		const int *parentSizes_lbl, const int *numCats_lbl,
		const int *parentsPool_lbl /*[N][N] 1-based labels, 0 padded; row all-zero = no parents allowed */,
		const int *fixedPool_lbl, const double *edgeLiks_lbl /* R layout: [col*N+row] */, int *pMaxParentSet) {
	int i, j, k;
	/* rcatnet_search.cpp:141-149 */
	if (parentSizes_lbl && sp->m_pParentSizes)
		for (i = 0; i < N; i++) sp->m_pParentSizes[i] = parentSizes_lbl[order[i] - 1];
	/* rcatnet_search.cpp:151-160 */
	for (j = 0; j < M; j++)
		for (i = 0; i < N; i++) {
			int v = samples_lbl[j * N + order[i] - 1];
			if (v == INT_MIN /* R NA_INTEGER */ || v < 1) v = CATNET_NAN;
			sp->m_pSamples[j * N + i] = v;
		}
	/* :163-174 */
	if (pert_lbl && sp->m_pPerturbations)
		for (j = 0; j < M; j++)
			for (i = 0; i < N; i++)
				sp->m_pPerturbations[j * N + i] = pert_lbl[j * N + order[i] - 1];
	/* :176-189 ; R passes catIndices = 1..C (catnet.search.R:199-202) */
	if (numCats_lbl && sp->m_pNodeNumCats && sp->m_pNodeCats)
		for (i = 0; i < N; i++) {
			int len = numCats_lbl[order[i] - 1];
			sp->m_pNodeNumCats[i] = len;
			if (sp->m_pNodeCats[i]) CATNET_FREE(sp->m_pNodeCats[i]);
			sp->m_pNodeCats[i] = (int*)CATNET_MALLOC(len * sizeof(int));
			for (j = 0; j < len; j++) sp->m_pNodeCats[i][j] = j + 1;
		}
	/* :191-224 */
	if (parentsPool_lbl && sp->m_parentsPool)
		for (i = 0; i < N; i++) {
			const int *row = parentsPool_lbl + (size_t)(order[i] - 1) * N;
			int len = 0;
			while (len < N && row[len] > 0) len++;
			if (len > 0) {
				for (j = 0; j < len; j++) {
					for (k = 0; k < N; k++)
						if (row[j] == order[k]) break;
					sp->m_parentsPool[i][j] = (k < N) ? k : -1;
				}
				for (; j < N; j++) sp->m_parentsPool[i][j] = -1;
			} else {
				for (j = 0; j < N; j++) sp->m_parentsPool[i][j] = -1;
				sp->m_parentsPool[i][0] = i;
			}
		}
	/* :226-256 */
	if (fixedPool_lbl && sp->m_fixedParentsPool)
		for (i = 0; i < N; i++) {
			const int *row = fixedPool_lbl + (size_t)(order[i] - 1) * N;
			int len = 0;
			while (len < N && row[len] > 0) len++;
			if (len > 0) {
				if (pMaxParentSet && *pMaxParentSet < len) *pMaxParentSet = len;
				for (j = 0; j < len; j++) {
					for (k = 0; k < N; k++)
						if (row[j] == order[k]) break;
					sp->m_fixedParentsPool[i][j] = (k < N) ? k : -1;
				}
				for (; j < N; j++) sp->m_fixedParentsPool[i][j] = -1;
			} else {
				for (j = 0; j < N; j++) sp->m_fixedParentsPool[i][j] = -1;
			}
		}
	/* :258-268 */
	if (edgeLiks_lbl && sp->m_matEdgeLiks)
		for (j = 0; j < N; j++)
			for (i = 0; i < N; i++)
				sp->m_matEdgeLiks[j * N + i] = edgeLiks_lbl[(size_t)(order[j] - 1) * N + order[i] - 1];
}

/* All "label space" inputs are indexed by original node label (as R hands them over);
 * `order` maps position -> 1-based label.  samples_lbl: [M][N] 1-based categories, NA = INT_MIN or <1. */
void *ref_search_order(int N, int M, const int *samples_lbl, const int *pert_lbl,
		int maxParentSet, const int *parentSizes_lbl, int maxComplexity, const int *order,
		const int *numCats_lbl, const int *parentsPool_lbl, const int *fixedPool_lbl,
		const double *edgeLiks_lbl, int useCache) {
	std::vector<int> inv(N);
	for (int i = 0; i < N; i++) inv[order[i] - 1] = i + 1;
	SEARCH_PARAMETERS *sp = new SEARCH_PARAMETERS(N, M, maxParentSet, maxComplexity, 0,
			numCats_lbl != 0, parentSizes_lbl != 0, pert_lbl != 0, parentsPool_lbl != 0,
			fixedPool_lbl != 0, edgeLiks_lbl != 0, 0, NULL, NULL);
	int mps = maxParentSet;
	fill_params(sp, N, M, samples_lbl, pert_lbl, order, parentSizes_lbl, numCats_lbl,
			parentsPool_lbl, fixedPool_lbl, edgeLiks_lbl, &mps);
	RefSearch *s = new RefSearch();
	if (useCache)
		s->setCacheParams(N, mps, const_cast<int*>(order), &inv[0]);
	int rc = s->estimate(sp);
	delete sp;
	RefResult *r = new RefResult();
	r->N = N;
	r->niter = r->naccept = 0;
	r->nslots = 0;
	r->nets = 0;
	if (rc == CATNET_ERR_OK && s->numCatnets() > 0 && s->catnets()) {
		r->nslots = s->numCatnets();
		r->nets = (RefNet**)malloc(r->nslots * sizeof(RefNet*));
		memcpy(r->nets, s->catnets(), r->nslots * sizeof(RefNet*));
		memset(s->catnets(), 0, r->nslots * sizeof(RefNet*)); /* steal, as rcatnet_sa.cpp:729-731 does */
	}
	delete s;
	return r;
}

int ref_result_nslots(void *h) { return ((RefResult*)h)->nslots; }
int ref_result_exists(void *h, int k) {
	RefResult *r = (RefResult*)h;
	return k >= 0 && k < r->nslots && r->nets[k] != 0;
}
int ref_result_net(void *h, int k, int *complexity, double *loglik) {
	RefResult *r = (RefResult*)h;
	if (!ref_result_exists(h, k)) return -1;
	*complexity = r->nets[k]->complexity();
	*loglik = r->nets[k]->loglik();
	return 0;
}
/* per-node view: parents (order positions), node log-lik, prior, sample size, CPT (normalised, problist.h:193-222) */
int ref_result_node(void *h, int k, int node, int *npars, int *pars /*cap N*/, double *loglik,
		double *priorlik, int *sampleSize, int *probsize, double *probs /*cap*/, int probs_cap) {
	RefResult *r = (RefResult*)h;
	if (!ref_result_exists(h, k)) return -1;
	RefNet *net = r->nets[k];
	int np = net->numParents(node);
	*npars = np;
	const int *pp = net->getParents(node);
	for (int i = 0; i < np; i++) pars[i] = pp ? pp[i] : -1;
	PROB_LIST<double> *pl = net->getNodeProb(node);
	if (!pl) return -2;
	*loglik = pl->loglik;
	*priorlik = pl->priorlik;
	*sampleSize = pl->sampleSize;
	*probsize = pl->nProbSize;
	for (int i = 0; i < pl->nProbSize && i < probs_cap; i++) probs[i] = pl->pProbs[i];
	return 0;
}
void ref_result_free(void *h) {
	RefResult *r = (RefResult*)h;
	if (!r) return;
	for (int i = 0; i < r->nslots; i++)
		if (r->nets[i]) delete r->nets[i];
	free(r->nets);
	delete r;
}
int ref_result_sa_info(void *h, int *order_out, int *niter, int *naccept) {
	RefResult *r = (RefResult*)h;
	for (size_t i = 0; i < r->order.size(); i++) order_out[i] = r->order[i];
	*niter = r->niter;
	*naccept = r->naccept;
	return (int)r->trace.size();
}
int ref_result_sa_trace(void *h, double *trace, int *accept, int cap) {
	RefResult *r = (RefResult*)h;
	int n = (int)r->trace.size();
	for (int i = 0; i < n && i < cap; i++) { trace[i] = r->trace[i]; accept[i] = r->trace_accept[i]; }
	return n;
}

/* ---------------- SA (rcatnet_sa.cpp) ---------------- */
THREAD_PROC_DEFINE(RefSaThreadProc, pParam) {          /* rcatnet_sa.cpp:37-46 */
	SEARCH_PARAMETERS *p = (SEARCH_PARAMETERS*)pParam;
	if (p && p->m_pCaller) {
		RefSearch *c = (RefSearch*)p->m_pCaller;
		c->estimate(p);
		c->_exit_thread((void*)0);
	}
	return 0;
}

/* rcatnet_sa.cpp:123-184, verbatim semantics (insertion move n1 -> n2) */
static int *sa_gen_order(const int *porder, int norder, int shuffles, int bjump) {
	int i, sh, n1, n2;
	double u;
	if (norder < 1) return 0;
	int *neworder = (int*)malloc(norder * sizeof(int));
	if (shuffles <= 0) {
		_gen_permutation<int>(neworder, norder);     /* utils.h:144-180 */
		return neworder;
	}
	int *torder = (int*)malloc(norder * sizeof(int));
	memcpy(neworder, porder, norder * sizeof(int));
	for (sh = 0; sh < shuffles; sh++) {
		memcpy(torder, neworder, norder * sizeof(int));
		u = unif_rand();
		n1 = (int)(u * norder);
		if (bjump) {
			n2 = n1;
			while (n2 == n1) { u = unif_rand(); n2 = (int)(u * norder); }
		} else {
			n2 = n1 + 1;
			if (n1 >= norder - 1) n2 = 0;
		}
		if (n1 < n2) {
			for (i = 0; i < n1; i++) neworder[i] = torder[i];
			for (i = n1; i < n2; i++) neworder[i] = torder[i + 1];
			neworder[n2] = torder[n1];
			for (i = n2 + 1; i < norder; i++) neworder[i] = torder[i];
		} else {
			for (i = 0; i < n2; i++) neworder[i] = torder[i];
			for (i = n2; i < n1; i++) neworder[i + 1] = torder[i];
			neworder[n2] = torder[n1];
			for (i = n1 + 1; i < norder; i++) neworder[i] = torder[i];
		}
	}
	free(torder);
	return neworder;
}

/* rcatnet_sa.cpp:186-251 */
static int *sa_gen_order_dirprobs(int numnodes, const double *mat, double *pOrderProb) {
	int i, j, k;
	int *neworder = (int*)calloc(numnodes, sizeof(int));
	int *flags = (int*)malloc(numnodes * sizeof(int));
	double *probs = (double*)malloc(numnodes * sizeof(double)), faux, fsum;
	*pOrderProb = 1;
	neworder[0] = 0;
	for (k = 1; k < numnodes; k++) {
		fsum = 0;
		for (i = 0; i <= k; i++) {
			faux = 1;
			for (j = 0; j < i; j++) faux *= mat[neworder[j] * numnodes + k];
			for (j = i; j < k; j++) faux *= mat[k * numnodes + neworder[j]];
			probs[i] = faux;
			fsum += faux;
		}
		faux = fsum * unif_rand();
		fsum = 0;
		for (i = 0; i < k; i++) {
			fsum += probs[i];
			if (fsum >= faux) break;
		}
		*pOrderProb *= probs[i];
		if (i > 0) memcpy(flags, neworder, i * sizeof(int));
		flags[i] = k;
		if (i < k) memcpy(flags + i + 1, neworder + i, (k - i) * sizeof(int));
		memcpy(neworder, flags, (k + 1) * sizeof(int));
	}
	for (i = 0; i < numnodes; i++) neworder[i]++;
	if (*pOrderProb > 0) *pOrderProb = log(*pOrderProb);
	else *pOrderProb = -FLT_MAX;
	free(flags);
	free(probs);
	return neworder;
}

void *ref_search_sa(int N, int M, const int *samples_lbl, const int *pert_lbl,
		int maxParentSet, const int *parentSizes_lbl, int maxComplexity,
		const int *numCats_lbl, const int *parentsPool_lbl, const int *fixedPool_lbl,
		const double *edgeLiks_lbl, const double *dirProbs,
		int selectMode /*0 max-complexity,1 AIC,2 BIC*/, const int *startOrder,
		double tempStart, double tempCoolFact, int tempCheckOrders, int maxIter,
		double orderShuffles, double stopDiff, int numThreads, int useCache) {
	int n, nn, i, nDrives = numThreads < 1 ? 1 : numThreads;
	int bUseCache = useCache, m_maxParentSet = maxParentSet;
	if (bUseCache && nDrives > 8) bUseCache = 0;                  /* :315-319 */
	MUTEX cache_mutex;
	std::vector<int> optOrder(startOrder, startOrder + N);
	int bjump = 0;
	if (orderShuffles < 0) { bjump = 1; orderShuffles = -orderShuffles; }     /* :371-377 */
	int minShuffles = (int)orderShuffles;
	orderShuffles = orderShuffles - minShuffles;

	std::vector<int*> testOrder(nDrives, (int*)0);
	std::vector<std::vector<int> > testInv(nDrives, std::vector<int>(N));
	std::vector<double> newOrdProb(nDrives, 0.0);
	std::vector<RefSearch*> drives(nDrives);
	std::vector<SEARCH_PARAMETERS*> sps(nDrives);
	for (n = 0; n < nDrives; n++) drives[n] = new RefSearch();
	if (bUseCache) MUTEX_INIT(cache_mutex);
	for (n = 0; n < nDrives; n++)                                  /* :441-450 */
		sps[n] = new SEARCH_PARAMETERS(N, M, m_maxParentSet, maxComplexity, 0,
				numCats_lbl != 0, parentSizes_lbl != 0, pert_lbl != 0, parentsPool_lbl != 0,
				fixedPool_lbl != 0, edgeLiks_lbl != 0, 0,
				bUseCache ? &cache_mutex : NULL, drives[n]);

	RefResult *res = new RefResult();
	res->N = N; res->nslots = 0; res->nets = 0;
	double fOptLogLik = -FLT_MAX, fLogLik, complx, ftemp, deltaLogLik, acceptprob, ordProb = 0;
	RefNet *pOptNet = 0, *pCurNet, **pCatnets;
	int nstop = 0, nstopCounter = 0, naccept = 0, nLastChangedTemp = 0, niter = 0;
	int stepiters, stepaccept, nCatnets, nCurNet;
	double tempCur = tempStart;

	while (niter < maxIter) {                                      /* :472 */
		for (n = 0; n < nDrives; n++) {
			if (testOrder[n]) free(testOrder[n]);
			if (dirProbs)
				testOrder[n] = sa_gen_order_dirprobs(N, dirProbs, &newOrdProb[n]);
			else
				testOrder[n] = sa_gen_order(&optOrder[0], N,
						minShuffles + _gen_binomial<double>(1, orderShuffles), bjump);
			nstop = 0;
			for (nn = 0; nn < n; nn++) {                               /* :491-506 */
				nstop = 1;
				for (i = 0; i < N; i++)
					if (testOrder[nn][i] != testOrder[n][i]) { nstop = 0; break; }
				if (nstop) break;
			}
			if (nstop) continue;
			for (i = 0; i < N; i++) testInv[n][testOrder[n][i] - 1] = i + 1;
			if (bUseCache)
				drives[n]->setCacheParams(N, m_maxParentSet, testOrder[n], &testInv[n][0]);
			fill_params(sps[n], N, M, samples_lbl, pert_lbl, testOrder[n], parentSizes_lbl, numCats_lbl,
					parentsPool_lbl, fixedPool_lbl, edgeLiks_lbl, &m_maxParentSet);
			drives[n]->_start_thread(RefSaThreadProc, sps[n]);        /* :639 */
		}
		nstop = 1;
		stepiters = nDrives;
		stepaccept = 0;
		for (n = 0; n < nDrives; n++) {                              /* :646 */
			if (!drives[n]->_is_running()) continue;
			int threadres = drives[n]->_join_thread();
			threadres = drives[n]->_stop_thread();
			if (stepaccept > 0) continue;
			nstop = 0;
			nCatnets = 0;
			pCatnets = 0;
			if (threadres == 0) { nCatnets = drives[n]->numCatnets(); pCatnets = drives[n]->catnets(); }
			if (nCatnets > 0 && pCatnets) {
				pCurNet = NULL;
				switch (selectMode) {
				case 1:
					fLogLik = -FLT_MAX;
					for (nCurNet = 0; nCurNet < nCatnets; nCurNet++) {
						if (!pCatnets[nCurNet]) continue;
						complx = M * pCatnets[nCurNet]->loglik() - pCatnets[nCurNet]->complexity();
						if (complx > fLogLik) { fLogLik = complx; pCurNet = pCatnets[nCurNet]; }
					}
					break;
				case 2:
					fLogLik = -FLT_MAX;
					ftemp = 0.5 * log((double)M);
					for (nCurNet = 0; nCurNet < nCatnets; nCurNet++) {
						if (!pCatnets[nCurNet]) continue;
						complx = M * pCatnets[nCurNet]->loglik() - ftemp * pCatnets[nCurNet]->complexity();
						if (complx > fLogLik) { fLogLik = complx; pCurNet = pCatnets[nCurNet]; }
					}
					break;
				}
				if (!pCurNet)
					for (nCurNet = nCatnets - 1; nCurNet >= 0; nCurNet--)
						if (pCatnets[nCurNet]) { pCurNet = pCatnets[nCurNet]; break; }
				if (!pCurNet) continue;
				fLogLik = pCurNet->loglik();
				int accepted = 0;
				if (!pOptNet) {                                          /* :714-737 */
					accepted = 1;
					stepaccept++;
					ordProb = newOrdProb[n];
				} else {                                                 /* :738-798 */
					deltaLogLik = fLogLik - fOptLogLik;
					deltaLogLik += (newOrdProb[n] - ordProb);
					acceptprob = 0;
					if (tempCur > 0 && deltaLogLik <= 0)
						acceptprob = exp(deltaLogLik / tempCur);
					ftemp = unif_rand();
					if (deltaLogLik > 0 || ftemp < acceptprob) {
						accepted = 1;
						ordProb = newOrdProb[n];
						stepaccept++;
						stepiters = n + 1;
						nstopCounter = 0;
					} else {
						deltaLogLik = 0;
					}
					if (deltaLogLik < 0) deltaLogLik = -deltaLogLik;
					if (deltaLogLik < stopDiff) nstopCounter++;
					if (nstopCounter >= tempCheckOrders) nstop = 1;
				}
				res->trace.push_back(fLogLik);
				res->trace_accept.push_back(accepted);
				if (accepted) {
					pOptNet = pCurNet;
					fOptLogLik = fLogLik;
					for (i = 0; i < res->nslots; i++)
						if (res->nets[i]) delete res->nets[i];
					free(res->nets);
					res->nslots = nCatnets;
					res->nets = (RefNet**)malloc(nCatnets * sizeof(RefNet*));
					memcpy(res->nets, pCatnets, nCatnets * sizeof(RefNet*));
					memset(pCatnets, 0, nCatnets * sizeof(RefNet*));
					memcpy(&optOrder[0], testOrder[n], N * sizeof(int));
				}
			}
		}
		if (stepaccept > 0) naccept++;
		niter += stepiters;
		if (niter - nLastChangedTemp >= tempCheckOrders) {            /* :807-814 */
			tempCur *= tempCoolFact;
			nLastChangedTemp = niter;
		}
		if (nstop) break;
	}
	for (n = 0; n < nDrives; n++) {
		if (testOrder[n]) free(testOrder[n]);
		delete sps[n];
		delete drives[n];
	}
	if (bUseCache) MUTEX_DESTROY(cache_mutex);
	res->order = optOrder;
	res->niter = niter;
	res->naccept = naccept;
	return res;
}

/* ---------------- cnSearchHist (rcatnet_hist.cpp:164-573) ----------------
 * The reference's loop is welded to SEXP; restated here around its own estimate(), threads and cache: batches of
 * numThreads fresh random orders (_gen_permutation, duplicates inside a batch skipped :300-316), one estimate() per
 * order, network picked by AIC / BIC / highest complexity (:463-503), weight 1 / exp(loglik/(M N)) / exp(score/(M N))
 * (:510-515) added to hist[parent][child] in label space (:517-528).  Returns the iteration count. */
int ref_search_hist(int N, int M, const int *samples_lbl, const int *pert_lbl, int maxParentSet,
		const int *parentSizes_lbl, int maxComplexity, const int *numCats_lbl, const int *parentsPool_lbl,
		const int *fixedPool_lbl, int scoreSel, int weight, int maxIter, int numThreads, int useCache,
		double *hist /* [N*N], [parent*N + child] */) {
	int n, nn, i, j, nDrives = numThreads < 1 ? 1 : numThreads;
	int bUseCache = useCache, m_maxParentSet = maxParentSet;
	if (bUseCache && nDrives > 8) bUseCache = 0;                  /* :205-209 */
	MUTEX cache_mutex;
	memset(hist, 0, sizeof(double) * N * N);
	std::vector<int*> testOrder(nDrives, (int*)0);
	std::vector<std::vector<int> > testInv(nDrives, std::vector<int>(N));
	std::vector<RefSearch*> drives(nDrives);
	std::vector<SEARCH_PARAMETERS*> sps(nDrives);
	for (n = 0; n < nDrives; n++) drives[n] = new RefSearch();
	if (bUseCache) MUTEX_INIT(cache_mutex);
	for (n = 0; n < nDrives; n++)                                  /* :268-277 */
		sps[n] = new SEARCH_PARAMETERS(N, M, m_maxParentSet, maxComplexity, 0,
				numCats_lbl != 0, parentSizes_lbl != 0, pert_lbl != 0, parentsPool_lbl != 0,
				fixedPool_lbl != 0, 0, 0, bUseCache ? &cache_mutex : NULL, drives[n]);
	int nstop = 0, niter = 0, nCatnets, nCurNet;
	double fLogLik, complx, ftemp;
	RefNet *pCurNet, **pCatnets;
	while (niter < maxIter) {                                      /* :282 */
		for (n = 0; n < nDrives; n++) {
			if (testOrder[n]) free(testOrder[n]);
			testOrder[n] = sa_gen_order(0, N, 0, 0);                   /* _gen_permutation */
			nstop = 0;
			for (nn = 0; nn < n; nn++) {
				nstop = 1;
				for (i = 0; i < N; i++)
					if (testOrder[nn][i] != testOrder[n][i]) { nstop = 0; break; }
				if (nstop) break;
			}
			if (nstop) continue;
			for (i = 0; i < N; i++) testInv[n][testOrder[n][i] - 1] = i + 1;
			if (bUseCache)
				drives[n]->setCacheParams(N, m_maxParentSet, testOrder[n], &testInv[n][0]);
			fill_params(sps[n], N, M, samples_lbl, pert_lbl, testOrder[n], parentSizes_lbl, numCats_lbl,
					parentsPool_lbl, fixedPool_lbl, 0, &m_maxParentSet);
			drives[n]->_start_thread(RefSaThreadProc, sps[n]);        /* :443 */
		}
		nstop = 1;
		for (n = 0; n < nDrives; n++) {                              /* :449 */
			if (!drives[n]->_is_running()) continue;
			nstop = 0;
			int threadres = drives[n]->_join_thread();
			threadres = drives[n]->_stop_thread();
			nCatnets = 0;
			pCatnets = 0;
			if (threadres == 0) { nCatnets = drives[n]->numCatnets(); pCatnets = drives[n]->catnets(); }
			if (nCatnets > 0 && pCatnets) {
				fLogLik = 0;
				pCurNet = NULL;
				switch (scoreSel) {
				case 1:
					fLogLik = -FLT_MAX;
					for (nCurNet = 0; nCurNet < nCatnets; nCurNet++) {
						if (!pCatnets[nCurNet]) continue;
						complx = M * pCatnets[nCurNet]->loglik() - pCatnets[nCurNet]->complexity();
						if (complx > fLogLik) { fLogLik = complx; pCurNet = pCatnets[nCurNet]; }
					}
					break;
				case 2:
					fLogLik = -FLT_MAX;
					ftemp = 0.5 * log((double)M);
					for (nCurNet = 0; nCurNet < nCatnets; nCurNet++) {
						if (!pCatnets[nCurNet]) continue;
						complx = M * pCatnets[nCurNet]->loglik() - ftemp * pCatnets[nCurNet]->complexity();
						if (complx > fLogLik) { fLogLik = complx; pCurNet = pCatnets[nCurNet]; }
					}
					break;
				}
				if (!pCurNet)
					for (nCurNet = nCatnets - 1; nCurNet >= 0; nCurNet--)
						if (pCatnets[nCurNet]) { pCurNet = pCatnets[nCurNet]; break; }
				if (!pCurNet) continue;
				if (weight <= 0) ftemp = 1;
				else if (weight == 1) ftemp = (double)exp(pCurNet->loglik() / (double)(M * N));
				else ftemp = (double)exp((double)fLogLik / (double)(M * N));
				const int *pNumParents = pCurNet->numParents();
				const int **pParents = pCurNet->parents();
				for (j = 0; j < N; j++) {
					if (pNumParents[j] <= 0) continue;
					for (i = 0; i < pNumParents[j]; i++)
						hist[(testOrder[n][pParents[j][i]] - 1) * N + testOrder[n][j] - 1] += ftemp;
				}
			}
			niter++;
		}
		if (nstop) break;
	}
	for (n = 0; n < nDrives; n++) {
		if (testOrder[n]) free(testOrder[n]);
		delete sps[n];
		delete drives[n];
	}
	if (bUseCache) MUTEX_DESTROY(cache_mutex);
	return niter;
}

/* CPU baseline helper: score `nfam` explicit families back to back with the reference primitive
 * (setParents + setNodeSampleProb), returns the sum of log-liks (to defeat dead-code elimination). */
double ref_score_family_list(const int *samples0, int N, int M, const int *numCats,
		const int *children, const int *parents /*[nfam][d]*/, int d, int nfam, double *scores_out) {
	int maxcats = 0, i;
	for (i = 0; i < N; i++)
		if (numCats[i] > maxcats) maxcats = numCats[i];
	RefNet net(N, d, maxcats, 0, 0, 0, numCats);
	double acc = 0;
	std::vector<int> pars(d > 0 ? d : 1);
	for (int f = 0; f < nfam; f++) {
		for (i = 0; i < d; i++) pars[i] = parents[(size_t)f * d + i];
		if (d > 0) net.setParents(children[f], &pars[0], d);
		double ll = net.setNodeSampleProb(children[f], const_cast<int*>(samples0), M);
		if (scores_out) scores_out[f] = ll;
		acc += ll;
	}
	return acc;
}

/* ---------------- fixed network: cnLoglik / cnNodeLoglik / cnSetProb ----------------
 * The .Call glue of these (catnetLoglik, catnetNodeLoglik, catnetSetProb; catnet_rexport.h) is absent from the
 * snapshot, the class methods they forward to are not: CATNET::sampleLoglikVector / bySampleLoglikVector /
 * sampleNodeLoglik (catnet_class.h:616-815) and setNodeSampleProb (:818-879).  These entry points build the CATNET the
 * glue would build from the S4 object (setParents + setCondProb, as RCatnet::RCatnet(SEXP) does, rcatnet.cpp:38-172)
 * and call those methods.  samples0: [M][N] zero-based codes, anything outside [0, numCats) is missing. */
static RefNet *make_net(int N, const int *numCats, int maxParents, const int *numParents, const int *parents0,
		const double *probs, const long long *offsets) {
	int maxcats = 0, i;
	for (i = 0; i < N; i++)
		if (numCats[i] > maxcats) maxcats = numCats[i];
	RefNet *net = new RefNet(N, maxParents, maxcats, 0, 0, 0, numCats);
	for (i = 0; i < N; i++) {
		if (numParents[i] > 0) net->setParents(i, const_cast<int *>(parents0 + (size_t)i * maxParents), numParents[i]);
		if (probs) net->setCondProb(i, const_cast<double *>(probs + offsets[i]), (int)(offsets[i + 1] - offsets[i]));
	}
	return net;
}

/* mode 0: sampleLoglikVector -> out[N]; 1: bySampleLoglikVector -> out[M]; 2: sampleNodeLoglik of every node -> out[N]
 * (no perturbations there); 3: sampleLoglik -> out[0] */
int ref_net_loglik(int N, int M, const int *samples0, const int *pert, const int *numCats, int maxParents,
		const int *numParents, const int *parents0, const double *probs, const long long *offsets, int mode,
		double *out) {
	RefNet *net = make_net(N, numCats, maxParents, numParents, parents0, probs, offsets);
	int *s = const_cast<int *>(samples0), *p = const_cast<int *>(pert);
	double *v = 0;
	int i;
	switch (mode) {
	case 0:
		v = net->sampleLoglikVector(s, M, p);
		if (v) memcpy(out, v, (size_t)N * sizeof(double));
		break;
	case 1:
		v = net->bySampleLoglikVector(s, M, p);
		if (v) memcpy(out, v, (size_t)M * sizeof(double));
		break;
	case 2:
		for (i = 0; i < N; i++) out[i] = net->sampleNodeLoglik(i, s, M);
		break;
	default:
		out[0] = net->sampleLoglik(s, M);
		break;
	}
	if (v) CATNET_FREE(v);
	delete net;
	return (mode < 2 && !v) ? -1 : 0;
}

/* cnSetProb: setNodeSampleProb(node, samples, M, bNormalize = 1) for every node; with perturbations the node sees the
 * sub-sample where it is not perturbed (the way estimate() feeds it, catnet_search2.h:458-472).  probs_out is laid out
 * by `offsets`; loglik_out[N] the returned family log-liks, sizes_out[N] the sample sizes. */
int ref_net_set_prob(int N, int M, const int *samples0, const int *pert, const int *numCats, int maxParents,
		const int *numParents, const int *parents0, const long long *offsets, double *probs_out, double *loglik_out,
		int *sizes_out) {
	RefNet *net = make_net(N, numCats, maxParents, numParents, parents0, 0, offsets);
	std::vector<int> sub;
	for (int i = 0; i < N; i++) {
		const int *s = samples0;
		int m = M;
		if (pert) {
			sub.clear();
			for (int j = 0; j < M; j++)
				if (!pert[(size_t)j * N + i]) sub.insert(sub.end(), samples0 + (size_t)j * N, samples0 + (size_t)(j + 1) * N);
			s = sub.empty() ? 0 : &sub[0];
			m = (int)(sub.size() / N);
		}
		double ll = net->setNodeSampleProb(i, const_cast<int *>(s), m, 1);
		PROB_LIST<double> *pl = net->getNodeProb(i);
		if (!pl || pl->nProbSize != offsets[i + 1] - offsets[i]) { delete net; return -1; }
		if (m < 1) pl->normalize();      /* setNodeSampleProb bails out before normalising an empty sample */
		memcpy(probs_out + offsets[i], pl->pProbs, (size_t)pl->nProbSize * sizeof(double));
		loglik_out[i] = ll;
		sizes_out[i] = m < 1 ? 0 : pl->sampleSize;
	}
	delete net;
	return 0;
}

} /* extern "C" */
