/*
 * oracle/catnet_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's structure-search scoring path (catnet 1.16.0, /root/reference/src).
 * It exists to check the CUDA path; nothing under catnet_b200/ links, imports or calls it.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * Parity status: PINNED against the reference itself -- tests/test_oracle.py compares every function here with
 * oracle/_ref/libcatnet_ref.so (the reference's own C++ compiled unmodified) and with the committed fixtures in
 * tests/golden/ that were generated from it (tools/make_golden.py).  The reference ships no golden vectors of its
 * own.  The reference's score cache (cache.cpp) is restated too (cache_get / cache_set below):
 * pinned on 9,300 fuzzed cnSearchSA / cnSearchHist problems against the compiled reference with its cache ON
 * (tools/fuzz_oracle.py sacacheport) and by tests/golden/ref_search_sa_cache.json.gz.
 *
 * Each function cites the reference lines it follows.  The arithmetic that decides ties is kept operation for
 * operation: counts held exactly, pp = 1/N_k, loglik += n*log(n*pp) in ascending table order, division by the
 * sample count last, network log-lik re-summed over nodes in index order, strict '<' comparisons.
 */
#include "catnet_oracle.h"
#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NEG_INF (-(double)FLT_MAX)

/* ------------------------------------------------------------------ family score ---- */
/* CATNET::setNodeSampleProb (catnet_class.h:818-879) + PROB_LIST::find_slot (problist.h:252-263) +
 * PROB_LIST::loglikelihood (problist.h:224-250).  samples0: [M][N] zero-based codes; a sample is skipped when the
 * child or any parent code is outside [0, cats).  keep (optional, [M]): 0 = sample excluded (perturbation). */
int orc_family_score(const int *samples0, int N, int M, const int *cats, int child, const int *pars, int d,
		const unsigned char *keep, double *counts, double *loglik_out, int *ncount_out) {
	int tsize = cats[child], j, i, k;
	for (i = 0; i < d; i++) tsize *= cats[pars[i]];
	memset(counts, 0, (size_t)tsize * sizeof(double));
	int ncount = 0;
	for (j = 0; j < M; j++) {
		if (keep && !keep[j]) continue;
		const int *row = samples0 + (size_t)j * N;
		int idx = 0, ok = 1;
		for (i = 0; i < d; i++) {                       /* first listed parent most significant */
			int v = row[pars[i]];
			if (v < 0 || v >= cats[pars[i]]) { ok = 0; break; }
			idx = idx * cats[pars[i]] + v;
		}
		int x = row[child];
		if (ok && x >= 0 && x < cats[child]) {
			counts[idx * cats[child] + x] += 1;
			ncount++;
		}
	}
	double ll = 0;
	int cc = cats[child];
	for (k = 0; k < tsize; k += cc) {
		double pp = 0;
		for (i = 0; i < cc; i++) pp += counts[k + i];
		if (pp <= 0) continue;
		pp = 1 / pp;
		for (i = 0; i < cc; i++)
			if (counts[k + i] > 0) ll += counts[k + i] * log(counts[k + i] * pp);
	}
	if (ncount > 1) ll /= (double)ncount;              /* catnet_class.h:867-870 */
	*loglik_out = ll;
	*ncount_out = ncount;
	return tsize;
}

/* ------------------------------------------------------------------ result containers ---- */
typedef struct {
	int child, d, pars[ORC_MAXP], tsize, ncount, cc;
	double *counts;
	double loglik, prior;
} fam_t;

typedef struct {
	int exists;
	int *fam;          /* [N] index into the family table */
	double m_loglik;   /* cached; 0 means "recompute" exactly like CATNET::loglik (catnet_class.h:469-473) */
} net_t;

struct orc_result {
	int N, nslots;
	net_t *nets;
	fam_t *fams;
	int nfams, cap;
	int *order;        /* SA: accepted order */
	int niter, naccept, ntrace, trace_cap;
	double *trace;
	int *trace_accept;
};

static int add_family(orc_result *r, const fam_t *f) {
	if (r->nfams == r->cap) {
		r->cap = r->cap ? 2 * r->cap : 256;
		r->fams = (fam_t *)realloc(r->fams, (size_t)r->cap * sizeof(fam_t));
	}
	fam_t *g = &r->fams[r->nfams];
	*g = *f;
	g->counts = (double *)malloc((size_t)f->tsize * sizeof(double));
	memcpy(g->counts, f->counts, (size_t)f->tsize * sizeof(double));
	return r->nfams++;
}

/* CATNET::findLoglik (catnet_class.h:475-485): sum over nodes in index order of (loglik + priorlik) */
static double net_loglik(const orc_result *r, net_t *n) {
	if (n->m_loglik == 0) {
		double s = 0;
		for (int i = 0; i < r->N; i++) s += (r->fams[n->fam[i]].loglik + r->fams[n->fam[i]].prior);
		n->m_loglik = s;
	}
	return n->m_loglik;
}

static int fam_complexity(const fam_t *f, const int *cats) {          /* catnet_class.h:457-467 */
	int c = cats[f->child] - 1;
	for (int j = 0; j < f->d; j++) c *= cats[f->pars[j]];
	return c;
}

/* Bernoulli edge prior of a candidate (catnet_search2.h:584-611; base network :408-435) */
static double edge_prior(const double *edge, int N, int node, const int *pars, int d) {
	double t = 1, t2 = 1, prior;
	int i, j;
	for (i = 0; i < d; i++) t *= edge[(size_t)pars[i] * N + node];
	prior = t > 0 ? log(t) : NEG_INF;
	for (i = 0; i < N; i++) {
		int isp = 0;
		if (i == node) continue;
		for (j = 0; j < d; j++) isp |= (pars[j] == i);
		if (!isp) t2 *= (1 - edge[(size_t)i * N + node]);
	}
	t2 = t2 > 0 ? log(t2) : NEG_INF;
	if (prior == NEG_INF || t2 == NEG_INF) return NEG_INF;
	return prior + t2;
}

static void score_family(fam_t *f, const int *S, int N, int M, const int *cats, const unsigned char *keep,
		const double *edge, double *buf) {
	f->counts = buf;
	f->cc = cats[f->child];
	f->tsize = orc_family_score(S, N, M, cats, f->child, f->pars, f->d, keep, buf, &f->loglik, &f->ncount);
	f->prior = edge ? edge_prior(edge, N, f->child, f->pars, f->d) : 0;
}

/* next combination in lexicographic order (what combinationSets enumerates, catnet_search2.h:102-147) */
static int next_comb(int *idx, int k, int n) {
	int i = k - 1;
	while (i >= 0 && idx[i] == n - k + i) i--;
	if (i < 0) return 0;
	idx[i]++;
	for (int j = i + 1; j < k; j++) idx[j] = idx[j - 1] + 1;
	return 1;
}

/* ------------------------------------------------------------------ score cache ---- */
/* c_cache (cache.cpp:142-317): one process-global table of CATNET_CACHE_EL pointers, no chaining -- a store replaces
 * whatever sits in its slot.  R switches it on for cnSearchSA and cnSearchHist (R/catnet.search.R:293-294) after
 * releasing it (:279), so every call starts empty.  An entry is keyed by (node, sorted pool, parent count) in LABEL
 * space and keeps the parent set in the sequence it was LISTED in when it was scored, its table (laid out in that
 * sequence), its log-likelihood (summed in that table order) and its prior (multiplied up in the node order of the
 * order it was scored under): a hit under a later order returns all of that unchanged, so cache-ON results differ from
 * cache-OFF ones in the listing order of parents and in the last bits of log-likelihoods. */
typedef struct {
	int node, npars, npool, *pool;     /* labels 1-based; pool sorted ascending */
	int pars[ORC_MAXP];                /* labels 1-based, in listing order */
	fam_t fam;                         /* counts owned (NULL for an entry that was never scored) */
} centry_t;
static centry_t **g_cache = NULL;
static unsigned g_ncache = 0, g_cbits = 0;
static int g_cache_on = 0;
static unsigned g_primes[180];             /* PRIMES_1000 (cache.cpp:35-54): the first 180 primes */

/* test hook: what the cache did, for checking the library's host-side model of it (catnet_b200/csrc/cnb_cache.cpp) on the
 * CPU.  rows of ORC_TRACE_ROW ints: kind (1 hit on the equal-categories path, 2 hit otherwise, 3 store of an
 * equal-categories winner), serial number of the search, child label, parent count, listed parent labels; the prior of a
 * hit beside it; the node order of every search. */
#define ORC_TRACE_ROW (4 + ORC_MAXP)
static int *g_tr_rows = NULL, *g_tr_orders = NULL, g_tr_cap = 0, g_tr_n = 0, g_tr_ocap = 0, g_tr_serial = 0;
static double *g_tr_prior = NULL;
void orc_cache_trace(int *rows, double *priors, int cap_rows, int *orders, int cap_orders) {
	g_tr_rows = rows; g_tr_prior = priors; g_tr_cap = cap_rows; g_tr_orders = orders; g_tr_ocap = cap_orders;
	g_tr_n = 0; g_tr_serial = 0;
}
void orc_cache_trace_counts(int *nrows, int *nsearches) { *nrows = g_tr_n; *nsearches = g_tr_serial; }
static void trace_row(int kind, int child, int d, const int *pars_lbl, double prior) {
	if (!g_tr_rows) return;
	if (g_tr_n < g_tr_cap) {
		int *r = g_tr_rows + (size_t)g_tr_n * ORC_TRACE_ROW;
		memset(r, 0, ORC_TRACE_ROW * sizeof(int));
		r[0] = kind; r[1] = g_tr_serial; r[2] = child; r[3] = d;
		for (int i = 0; i < d && i < ORC_MAXP; i++) r[4 + i] = pars_lbl[i];
		if (g_tr_prior) g_tr_prior[g_tr_n] = prior;
	}
	g_tr_n++;
}

static void centry_free(centry_t *e) { if (e) { free(e->pool); free(e->fam.counts); free(e); } }
void orc_cache_release(void) {               /* ReleaseCache (cache.cpp:62-76) */
	if (g_cache) for (unsigned i = 0; i < g_ncache; i++) centry_free(g_cache[i]);
	free(g_cache);
	g_cache = NULL;
	g_ncache = g_cbits = 0;
}
void orc_set_cache(int on) { g_cache_on = on; }

static void cache_primes(void) {
	if (g_primes[0]) return;
	int n = 0;
	for (unsigned v = 2; n < 180; v++) {
		int isp = 1;
		for (unsigned q = 2; q * q <= v; q++) if (v % q == 0) { isp = 0; break; }
		if (isp) g_primes[n++] = v;
	}
}
/* table size chosen at the first store (cache.cpp:225-264), unsigned 32-bit arithmetic as there */
static void cache_init(int N, int maxParentSet) {
	cache_primes();
	unsigned nl = (unsigned)N;
	int i;
	for (i = 0; i < maxParentSet; i++) nl *= (unsigned)N;
	if (nl < g_primes[179]) {
		for (i = 0; i < 180; i++) if (nl >= g_primes[i]) { nl = g_primes[i]; break; }
	} else {
		unsigned prime = g_primes[179];
		for (i = 0; i < 180; i++) if (g_primes[i] >= (unsigned)N) { prime = g_primes[i]; break; }
		nl = prime;
		for (i = 0; i < maxParentSet; i++) nl *= prime;
	}
	if (nl < g_primes[10]) nl = g_primes[10];
	g_cbits = 1;
	i = 1;
	while (i < N * maxParentSet) { g_cbits++; i <<= 1; }
	if (nl < 1) nl = 1;                       /* InitializeCache (cache.cpp:78-88) */
	g_ncache = nl;
	g_cache = (centry_t **)calloc(g_ncache, sizeof(centry_t *));
}
static int cmp_int(const void *a, const void *b) { return *(const int *)a - *(const int *)b; }
/* slot of (sorted pool, node, parent count): cache.cpp:174-184 = :296-306 */
static unsigned cache_slot(const int *pool, int npool, int node, int npars, int N) {
	unsigned nl = 1;
	for (int i = 0; i < npool; i++) {
		int j = (pool[i] - 1) % 180;
		nl *= g_primes[180 - j - 1];
		nl %= g_ncache;
	}
	nl = (nl << g_cbits) + node + N * npars;
	return nl % g_ncache;
}
/* getCachedProb (cache.cpp:142-216).  pool: labels, sorted here */
static centry_t *cache_get(int *pool, int npool, int node, int npars, int N) {
	if (!g_cache) return NULL;
	qsort(pool, npool, sizeof(int), cmp_int);
	centry_t *e = g_cache[cache_slot(pool, npool, node, npars, N)];
	if (!e || e->node != node || e->npars != npars || e->npool != npool) return NULL;
	for (int i = 0; i < npool; i++) if (e->pool[i] != pool[i]) return NULL;
	return e;
}
/* setCachedProb (cache.cpp:218-317): f = the scored family with parents as order positions, ord = position -> label0 */
static void cache_set(int *pool, int npool, int node, const fam_t *f, const int *ord, int N, int maxParentSet) {
	if (!g_cache) cache_init(N, maxParentSet);
	qsort(pool, npool, sizeof(int), cmp_int);
	centry_t *e = (centry_t *)calloc(1, sizeof(centry_t));
	e->node = node; e->npars = f->d; e->npool = npool;
	e->pool = (int *)malloc((npool ? npool : 1) * sizeof(int));
	memcpy(e->pool, pool, npool * sizeof(int));
	for (int i = 0; i < f->d; i++) e->pars[i] = ord[f->pars[i]] + 1;
	e->fam = *f;
	e->fam.counts = NULL;
	if (f->counts && f->tsize > 0) {
		e->fam.counts = (double *)malloc((size_t)f->tsize * sizeof(double));
		memcpy(e->fam.counts, f->counts, (size_t)f->tsize * sizeof(double));
	}
	unsigned s = cache_slot(pool, npool, node, f->d, N);
	centry_free(g_cache[s]);
	g_cache[s] = e;
}
/* a hit as a family of the current order: parents back to positions, in the entry's listing sequence (cache.cpp:203-205) */
static void cache_fam(const centry_t *e, int child, const int *pos_of, fam_t *out) {
	*out = e->fam;
	out->child = child;
	for (int i = 0; i < e->npars; i++) out->pars[i] = pos_of[e->pars[i] - 1];
}

/* ------------------------------------------------------------------ search for one order ---- */
/* RCatnetSearch::estimateCatnets marshalling (rcatnet_search.cpp:111-268) + CATNET_SEARCH2::estimate
 * (catnet_search2.h:152-897).  All label-space inputs as R passes them; see catnet_oracle.h. */
orc_result *orc_search_order(int N, int M, const int *samples_lbl, const int *pert_lbl, int maxParentSet,
		const int *parentSizes_lbl, int maxComplexity, const int *order, const int *numCats_lbl,
		const int *parentsPool_lbl, const int *fixedPool_lbl, const double *edge_lbl) {
	int i, j, k, d, n;
	if (N < 1 || M < 1) return NULL;
	if (maxComplexity < N) maxComplexity = N;                           /* :182-183 */
	int *ord = (int *)malloc(N * sizeof(int)), *pos_of = (int *)malloc(N * sizeof(int));
	for (i = 0; i < N; i++) { ord[i] = order ? order[i] - 1 : i; pos_of[ord[i]] = i; }
	if (g_tr_orders && g_cache_on && g_tr_serial < g_tr_ocap)
		for (i = 0; i < N; i++) g_tr_orders[(size_t)g_tr_serial * N + i] = ord[i] + 1;

	/* samples into order space, NA -> INT_MAX (rcatnet_search.cpp:151-160); categories given as 1..C or inferred
	 * as min..max (catnet_search2.h:201-235); recode to 0-based, NA -> C (:237-257) */
	int *S = (int *)malloc((size_t)N * M * sizeof(int)), *cats = (int *)malloc(N * sizeof(int));
	for (i = 0; i < N; i++) {
		int lo = INT_MAX, hi = -INT_MAX;
		for (j = 0; j < M; j++) {
			int v = samples_lbl[(size_t)j * N + ord[i]];
			if (v == INT_MIN || v < 1) v = INT_MAX;
			S[(size_t)j * N + i] = v;
			if (v != INT_MAX) { if (v < lo) lo = v; if (v > hi) hi = v; }
		}
		if (numCats_lbl) { cats[i] = numCats_lbl[ord[i]]; lo = 1; }
		else cats[i] = hi - lo + 1;
		for (j = 0; j < M; j++) {
			int v = S[(size_t)j * N + i];
			S[(size_t)j * N + i] = (v == INT_MAX) ? cats[i] : v - lo;
		}
	}
	double *edge = NULL;
	if (edge_lbl) {                                                       /* rcatnet_search.cpp:258-268 */
		edge = (double *)malloc((size_t)N * N * sizeof(double));
		for (j = 0; j < N; j++)
			for (i = 0; i < N; i++) edge[(size_t)j * N + i] = edge_lbl[(size_t)ord[j] * N + ord[i]];
	}
	unsigned char *allow = NULL, *fixd = NULL;                             /* [child][candidate] in order space */
	if (parentsPool_lbl) {
		allow = (unsigned char *)calloc((size_t)N * N, 1);
		for (i = 0; i < N; i++)
			for (j = 0; j < N && parentsPool_lbl[(size_t)ord[i] * N + j] > 0; j++) {
				int l = parentsPool_lbl[(size_t)ord[i] * N + j];
				if (l <= N) allow[(size_t)i * N + pos_of[l - 1]] = 1;
			}
	}
	if (fixedPool_lbl) {
		fixd = (unsigned char *)calloc((size_t)N * N, 1);
		for (i = 0; i < N; i++)
			for (j = 0; j < N && fixedPool_lbl[(size_t)ord[i] * N + j] > 0; j++) {
				int l = fixedPool_lbl[(size_t)ord[i] * N + j];
				if (l <= N) fixd[(size_t)i * N + pos_of[l - 1]] = 1;
			}
	}

	orc_result *R = (orc_result *)calloc(1, sizeof(orc_result));
	R->N = N;
	R->nslots = maxComplexity + 1;
	R->nets = (net_t *)calloc(R->nslots, sizeof(net_t));
	net_t *cur = (net_t *)calloc(R->nslots, sizeof(net_t));
	int maxcat = 0;
	for (i = 0; i < N; i++) if (cats[i] > maxcat) maxcat = cats[i];
	size_t maxtab = maxcat;                      /* largest table: child x up to max(maxParentSet, N-1) parents */
	for (i = 0; i < N - 1 && maxtab < ((size_t)1 << 26); i++) maxtab *= maxcat;
	double *cbuf = (double *)malloc(maxtab * sizeof(double)), *bestc = (double *)malloc(maxtab * sizeof(double));
	unsigned char *keep = pert_lbl ? (unsigned char *)malloc(M) : NULL;
	int *fix = (int *)malloc(N * sizeof(int)), *fre = (int *)malloc(N * sizeof(int)), *idx = (int *)malloc(N * sizeof(int));
	int *basefam = (int *)malloc(N * sizeof(int)), *plbl = (int *)malloc((N + ORC_MAXP) * sizeof(int));

#define SET_KEEP(node) do { if (keep) for (int q_ = 0; q_ < M; q_++) keep[q_] = pert_lbl[(size_t)q_ * N + ord[node]] == 0; } while (0)
	/* base network: every node with its fixed parents (catnet_search2.h:359-441) */
	int basecx = 0;
	for (n = 0; n < N; n++) {
		int nf = 0;
		for (j = 0; j < n; j++) {
			int a = !allow || allow[(size_t)n * N + j];
			if (a && fixd && fixd[(size_t)n * N + j]) fix[nf++] = j;
		}
		fam_t f;
		memset(&f, 0, sizeof(f));
		f.child = n; f.d = nf;
		for (j = 0; j < nf; j++) f.pars[j] = fix[j];
		SET_KEEP(n);
		score_family(&f, S, N, M, cats, keep, edge, cbuf);
		basefam[n] = add_family(R, &f);
		basecx += fam_complexity(&f, cats);
	}
	if (basecx > maxComplexity) { /* the reference would write out of bounds here; refuse */
		free(cur); orc_result_free(R); R = NULL; goto done;
	}
	R->nets[basecx].exists = 1;
	R->nets[basecx].fam = (int *)malloc(N * sizeof(int));
	memcpy(R->nets[basecx].fam, basefam, N * sizeof(int));

	for (n = 1; n < N; n++) {                                           /* :444 */
		int nf = 0, nr = 0;
		for (j = 0; j < n; j++) {                                       /* :456-486 */
			int a = !allow || allow[(size_t)n * N + j];
			if (!a) continue;
			if (fixd && fixd[(size_t)n * N + j]) fix[nf++] = j; else fre[nr++] = j;
		}
		int equal = 1, c0 = nr ? cats[fre[0]] : (nf ? cats[fix[0]] : 0);   /* :491-496 over free ++ fixed */
		for (j = 0; j < nr; j++) if (cats[fre[j]] != c0) equal = 0;
		for (j = 0; j < nf; j++) if (cats[fix[j]] != c0) equal = 0;
		int maxpars = maxParentSet;                                    /* :498-503 */
		if (parentSizes_lbl && parentSizes_lbl[ord[n]] < maxpars) maxpars = parentSizes_lbl[ord[n]];
		if (maxpars > nr + nf) maxpars = nr + nf;
		for (k = 0; k < R->nslots; k++) { cur[k].exists = 0; }
		SET_KEEP(n);
		/* a node that is perturbed in EVERY sample has no sample to score a candidate with: setNodeSampleProb returns
		 * -FLT_MAX for nsamples < 1 (catnet_class.h:824-826), so no candidate beats the initial fMaxLogLik = -FLT_MAX and
		 * the node keeps its base family (equal categories: ncombMaxLogLik < 0 -> continue, catnet_search2.h:637-640;
		 * mixed: tempLogLik == -FLT_MAX never creates or replaces a slot, :813-826) */
		int dead = 0;
		if (keep) {
			int nkept = 0;
			for (j = 0; j < M; j++) nkept += keep[j];
			if (nkept < 1) dead = 1;
		}
		/* ... but with the cache on, the unequal-categories path still stores an entry per candidate (:789-797), and a store
		 * can evict what a later order would have hit: those stores are kept, the rest of the node's work is dropped */
		if (dead && !(g_cache_on && !equal)) maxpars = 0;
		const fam_t *bf = &R->fams[basefam[n]];
		int base_node_cx = fam_complexity(bf, cats);
		double base_term = bf->loglik + bf->prior;

		for (d = nf + 1; d <= maxpars; d++) {
			int kk = d - nf;
			double fMax = NEG_INF;
			fam_t best, cand;
			int have_best = 0;
			memset(&best, 0, sizeof(best));
			int hit = 0;
			if (g_cache_on && equal) {
				/* equal categories: ONE entry per (node, pool, d) holding the winner (:514-524, :642-650) */
				for (j = 0; j < nr; j++) plbl[j] = ord[fre[j]] + 1;
				for (j = 0; j < nf; j++) plbl[nr + j] = ord[fix[j]] + 1;
				const centry_t *e = cache_get(plbl, nr + nf, ord[n] + 1, d, N);
				if (e) {
					cache_fam(e, n, pos_of, &best);
					fMax = best.loglik;
					if (edge) fMax += best.prior;
					have_best = hit = 1;
					trace_row(1, ord[n] + 1, d, e->pars, best.prior);
				}
			}
			for (j = 0; j < kk; j++) idx[j] = j;
			if (!hit) do {
				memset(&cand, 0, sizeof(cand));
				cand.child = n; cand.d = d;
				for (j = 0; j < nf; j++) cand.pars[j] = fix[j];       /* fixed first (:532-553) */
				for (j = 0; j < kk; j++) cand.pars[nf + j] = fre[idx[j]];
				const centry_t *ce = NULL;
				if (g_cache_on && !equal) {
					/* unequal categories: one entry per candidate, its own parents as the pool (:724-733, :789-797) */
					for (j = 0; j < d; j++) plbl[j] = ord[cand.pars[j]] + 1;
					ce = cache_get(plbl, d, ord[n] + 1, d, N);
				}
				if (ce) { cache_fam(ce, n, pos_of, &cand); if (!dead) trace_row(2, ord[n] + 1, d, ce->pars, cand.prior); }
				else if (dead) { cand.loglik = NEG_INF; cand.cc = cats[n]; }
				else score_family(&cand, S, N, M, cats, keep, edge, cbuf);
				if (g_cache_on && !equal && !ce) {
					for (j = 0; j < d; j++) plbl[j] = ord[cand.pars[j]] + 1;
					cache_set(plbl, d, ord[n] + 1, &cand, ord, N, maxParentSet);
				}
				if (dead) continue;
				double f = cand.loglik;
				if (edge) f += cand.prior;                              /* :610 */
				if (equal) {
					if (fMax < f) {                                     /* :613 strict, first maximum wins */
						fMax = f;
						best = cand;
						best.counts = bestc;
						memcpy(bestc, cbuf, (size_t)cand.tsize * sizeof(double));
						have_best = 1;
					}
					continue;
				}
				/* unequal categories: every candidate runs the DP update (:800-827) */
				int ncx = fam_complexity(&cand, cats), famid = -1;
				for (k = 0; k < R->nslots; k++) {
					if (!R->nets[k].exists) continue;
					int cx = k - base_node_cx + ncx;
					if (cx > maxComplexity) continue;
					double t = net_loglik(R, &R->nets[k]) - base_term + f;
					if (!cur[cx].exists && t > NEG_INF) {
						cur[cx].exists = 1;
						if (!cur[cx].fam) cur[cx].fam = (int *)malloc(N * sizeof(int));
						cur[cx].fam[0] = -1;                            /* empty CATNET: loglik() = -FLT_MAX */
					}
					if (cur[cx].exists && (cur[cx].fam[0] < 0 ? NEG_INF : net_loglik(R, &cur[cx])) < t) {
						if (famid < 0) famid = add_family(R, &cand);
						memcpy(cur[cx].fam, R->nets[k].fam, N * sizeof(int));
						cur[cx].fam[n] = famid;
						cur[cx].m_loglik = 0;                            /* setParents invalidates (:429-430) */
					}
				}
			} while (next_comb(idx, kk, nr));
			if (!equal || !have_best) continue;                         /* :637-640 */
			if (g_cache_on && !hit) {
				for (j = 0; j < nr; j++) plbl[j] = ord[fre[j]] + 1;
				for (j = 0; j < nf; j++) plbl[nr + j] = ord[fix[j]] + 1;
				cache_set(plbl, nr + nf, ord[n] + 1, &best, ord, N, maxParentSet);
				for (j = 0; j < d; j++) plbl[j] = ord[best.pars[j]] + 1;
				trace_row(3, ord[n] + 1, d, plbl, best.prior);
			}
			int ncx = fam_complexity(&best, cats), famid = -1;
			for (k = 0; k < R->nslots; k++) {                           /* :657-676 */
				if (!R->nets[k].exists) continue;
				int cx = k - base_node_cx + ncx;
				if (cx > maxComplexity) continue;
				double t = net_loglik(R, &R->nets[k]) - base_term + fMax;
				if (!cur[cx].exists && t > NEG_INF) {
					cur[cx].exists = 1;
					if (!cur[cx].fam) cur[cx].fam = (int *)malloc(N * sizeof(int));
					cur[cx].fam[0] = -1;
				}
				if (cur[cx].exists && (cur[cx].fam[0] < 0 ? NEG_INF : net_loglik(R, &cur[cx])) < t) {
					if (famid < 0) famid = add_family(R, &best);
					memcpy(cur[cx].fam, R->nets[k].fam, N * sizeof(int));
					cur[cx].fam[n] = famid;
					cur[cx].m_loglik = 0;
				}
			}
		}
		for (j = 0; j < R->nslots; j++) {                              /* :847-865 */
			if (!cur[j].exists || cur[j].fam[0] < 0) { cur[j].exists = 0; continue; }
			int take = !R->nets[j].exists || net_loglik(R, &R->nets[j]) < net_loglik(R, &cur[j]);
			if (take) {
				int *t = R->nets[j].fam;
				R->nets[j].fam = cur[j].fam;
				R->nets[j].m_loglik = cur[j].m_loglik;
				R->nets[j].exists = 1;
				cur[j].fam = t;
			}
			cur[j].exists = 0;
		}
	}
	for (k = 0; k < R->nslots; k++) free(cur[k].fam);
	free(cur);
done:
	if (g_cache_on) g_tr_serial++;
	free(ord); free(pos_of); free(S); free(cats); free(edge); free(allow); free(fixd); free(cbuf); free(bestc);
	free(keep); free(fix); free(fre); free(idx); free(basefam); free(plbl);
	return R;
}

/* ------------------------------------------------------------------ accessors ---- */
int orc_result_nslots(const orc_result *r) { return r ? r->nslots : 0; }
int orc_result_exists(const orc_result *r, int k) { return r && k >= 0 && k < r->nslots && r->nets[k].exists; }
int orc_result_net(orc_result *r, int k, int *complexity, double *loglik) {
	if (!orc_result_exists(r, k)) return -1;
	*complexity = k;
	*loglik = net_loglik(r, &r->nets[k]);
	return 0;
}
/* node view; probs = counts normalised as PROB_LIST::normalize does (problist.h:193-222) */
int orc_result_node(const orc_result *r, int k, int node, int *npars, int *pars, double *loglik, double *prior,
		int *sampleSize, int *probsize, double *probs, int cap) {
	if (!orc_result_exists(r, k)) return -1;
	const fam_t *f = &r->fams[r->nets[k].fam[node]];
	*npars = f->d;
	for (int i = 0; i < f->d; i++) pars[i] = f->pars[i];
	*loglik = f->loglik;
	*prior = f->prior;
	*sampleSize = f->ncount;
	*probsize = f->tsize;
	for (int b = 0; probs && b < f->tsize && b + f->cc <= cap; b += f->cc) {
		double pp = 0;
		for (int i = 0; i < f->cc; i++) pp += f->counts[b + i];
		if (pp > 0) {
			pp = 1 / pp;
			for (int i = 0; i < f->cc; i++) probs[b + i] = f->counts[b + i] * pp;
		} else {
			double favg = (double)1 / (double)f->cc, acc = 0;
			for (int i = 0; i < f->cc - 1; i++) { probs[b + i] = favg; acc += favg; }
			probs[b + f->cc - 1] = 1 - acc;
		}
	}
	return 0;
}
int orc_result_counts(const orc_result *r, int k, int node, double *counts, int cap) {
	if (!orc_result_exists(r, k)) return -1;
	const fam_t *f = &r->fams[r->nets[k].fam[node]];
	for (int i = 0; i < f->tsize && i < cap; i++) counts[i] = f->counts[i];
	return f->tsize;
}
void orc_result_free(orc_result *r) {
	if (!r) return;
	for (int k = 0; k < r->nslots; k++) free(r->nets[k].fam);
	for (int i = 0; i < r->nfams; i++) free(r->fams[i].counts);
	free(r->nets); free(r->fams); free(r->order); free(r->trace); free(r->trace_accept);
	free(r);
}

/* ------------------------------------------------------------------ SA ---- */
static unsigned long long g_state = 0x9E3779B97F4A7C15ull;
static double (*g_unif_fn)(void *) = 0;
static void *g_unif_ctx = 0;
void orc_set_seed(unsigned long long seed) { g_state = seed; g_unif_fn = 0; }
/* a caller-owned generator instead (what R's unif_rand is to the reference); cleared by orc_set_seed */
void orc_set_unif(double (*fn)(void *), void *ctx) { g_unif_fn = fn; g_unif_ctx = ctx; }
/* splitmix64 -> (0,1): the injected stand-in for R's unif_rand, identical to oracle/ref_driver.cpp */
static double unif(void) {
	if (g_unif_fn) return g_unif_fn(g_unif_ctx);
	unsigned long long z = (g_state += 0x9E3779B97F4A7C15ull);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	z = z ^ (z >> 31);
	return ((double)(z >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

static void gen_permutation(int *out, int n) {                          /* utils.h:144-180 */
	int *aux = (int *)malloc(n * sizeof(int)), cc = 0, brep = 0, i, j;
	while (++cc < 1e5) {
		brep = 0;
		for (i = 0; i < n; i++) aux[i] = (int)(n * n * unif());
		for (j = 0; j < n; j++) {
			out[j] = 0;
			for (i = 0; i < n; i++) {
				if (i != j && aux[j] == aux[i]) { brep = 1; break; }
				if (aux[j] >= aux[i]) out[j]++;
			}
			if (brep) break;
		}
		if (!brep) break;
	}
	if (cc >= 1e5 - 1) for (j = 0; j < n; j++) out[j] = j + 1;
	free(aux);
}

static void gen_order(const int *curo, int *out, int n, int shuffles, int bjump) {   /* rcatnet_sa.cpp:123-184 */
	int sh, i, n1, n2;
	if (shuffles <= 0) { gen_permutation(out, n); return; }
	int *t = (int *)malloc(n * sizeof(int));
	memcpy(out, curo, n * sizeof(int));
	for (sh = 0; sh < shuffles; sh++) {
		memcpy(t, out, n * sizeof(int));
		n1 = (int)(unif() * n);
		if (bjump) { n2 = n1; while (n2 == n1) n2 = (int)(unif() * n); }
		else { n2 = n1 + 1; if (n1 >= n - 1) n2 = 0; }
		if (n1 < n2) for (i = n1; i < n2; i++) out[i] = t[i + 1];
		else for (i = n2; i < n1; i++) out[i + 1] = t[i];
		out[n2] = t[n1];
	}
	free(t);
}

/* the SA loop of RCatnetSearchSA::search (rcatnet_sa.cpp:471-820), one evaluation after the other, with the score
 * cache when orc_set_cache(1) (R's setting, R/catnet.search.R:293-294); dirProbs proposals are not restated here */
orc_result *orc_search_sa(int N, int M, const int *samples_lbl, const int *pert_lbl, int maxParentSet,
		const int *parentSizes_lbl, int maxComplexity, const int *numCats_lbl, const int *parentsPool_lbl,
		const int *fixedPool_lbl, const double *edge_lbl, int selectMode, const int *startOrder, double tempStart,
		double tempCoolFact, int tempCheckOrders, int maxIter, double orderShuffles, double stopDiff, int numThreads) {
	int nDrives = numThreads < 1 ? 1 : numThreads, n, nn, i, bjump = 0;
	if (orderShuffles < 0) { bjump = 1; orderShuffles = -orderShuffles; }
	int minShuffles = (int)orderShuffles;
	orderShuffles -= minShuffles;
	int *opt = (int *)malloc(N * sizeof(int));
	memcpy(opt, startOrder, N * sizeof(int));
	int **test = (int **)malloc(nDrives * sizeof(int *)), *skip = (int *)malloc(nDrives * sizeof(int));
	for (n = 0; n < nDrives; n++) test[n] = (int *)malloc(N * sizeof(int));
	orc_result *best = NULL;
	double fOpt = NEG_INF, tempCur = tempStart, *trace = NULL;
	int *tacc = NULL, ntrace = 0, tcap = 0;
	int nstop = 0, nstopCounter = 0, naccept = 0, nLast = 0, niter = 0;
	/* score cache (orc_set_cache): off beyond 8 threads (rcatnet_sa.cpp:315-319).  The reference starts all the step's
	 * searches on threads before it looks at any of them, so every proposal leaves its entries in the cache whether or not
	 * an earlier one is accepted; with one thread that is this loop, with more the threads' stores interleave as they
	 * come (the reference's own result then depends on timing) and the restatement takes them proposal after proposal. */
	const int cache_was = g_cache_on;
	if (nDrives > 8) g_cache_on = 0;
	orc_result **pre = (orc_result **)calloc(nDrives, sizeof(orc_result *));
	while (niter < maxIter) {
		for (n = 0; n < nDrives; n++) {
			int extra = 0;
			if (orderShuffles > 0 && unif() <= orderShuffles) extra = 1;   /* _gen_binomial(1, frac) */
			gen_order(opt, test[n], N, minShuffles + extra, bjump);
			skip[n] = 0;
			for (nn = 0; nn < n && !skip[n]; nn++) skip[n] = !memcmp(test[nn], test[n], N * sizeof(int));
		}
		if (g_cache_on)
			for (n = 0; n < nDrives; n++)
				if (!skip[n]) pre[n] = orc_search_order(N, M, samples_lbl, pert_lbl, maxParentSet, parentSizes_lbl,
						maxComplexity, test[n], numCats_lbl, parentsPool_lbl, fixedPool_lbl, edge_lbl);
		nstop = 1;
		int stepiters = nDrives, stepaccept = 0;
		for (n = 0; n < nDrives; n++) {
			if (skip[n]) continue;
			if (stepaccept > 0) { orc_result_free(pre[n]); pre[n] = NULL; continue; }
			nstop = 0;
			orc_result *r = g_cache_on ? pre[n] : orc_search_order(N, M, samples_lbl, pert_lbl, maxParentSet, parentSizes_lbl,
					maxComplexity, test[n], numCats_lbl, parentsPool_lbl, fixedPool_lbl, edge_lbl);
			pre[n] = NULL;
			if (!r) continue;
			int pick = -1, k;
			double fl = NEG_INF;
			if (selectMode == 1 || selectMode == 2) {
				double ft = selectMode == 1 ? 1.0 : 0.5 * log((double)M);
				for (k = 0; k < r->nslots; k++) {
					if (!r->nets[k].exists) continue;
					double v = selectMode == 1 ? M * net_loglik(r, &r->nets[k]) - k : M * net_loglik(r, &r->nets[k]) - ft * k;
					if (v > fl) { fl = v; pick = k; }
				}
			}
			if (pick < 0) for (k = r->nslots - 1; k >= 0; k--) if (r->nets[k].exists) { pick = k; break; }
			if (pick < 0) { orc_result_free(r); continue; }
			fl = net_loglik(r, &r->nets[pick]);
			int accepted = 0;
			if (!best) { accepted = 1; stepaccept++; }
			else {
				double delta = fl - fOpt, ap = 0;
				if (tempCur > 0 && delta <= 0) ap = exp(delta / tempCur);
				double u = unif();
				if (delta > 0 || u < ap) { accepted = 1; stepaccept++; stepiters = n + 1; nstopCounter = 0; }
				else delta = 0;
				if (delta < 0) delta = -delta;
				if (delta < stopDiff) nstopCounter++;
				if (nstopCounter >= tempCheckOrders) nstop = 1;
			}
			if (ntrace == tcap) {
				tcap = tcap ? 2 * tcap : 256;
				trace = (double *)realloc(trace, tcap * sizeof(double));
				tacc = (int *)realloc(tacc, tcap * sizeof(int));
			}
			trace[ntrace] = fl;
			tacc[ntrace++] = accepted;
			if (accepted) {
				fOpt = fl;
				orc_result_free(best);
				best = r;
				memcpy(opt, test[n], N * sizeof(int));
			} else orc_result_free(r);
		}
		if (stepaccept > 0) naccept++;
		niter += stepiters;
		if (niter - nLast >= tempCheckOrders) { tempCur *= tempCoolFact; nLast = niter; }
		if (nstop) break;
	}
	if (best) {
		best->order = opt;
		best->niter = niter;
		best->naccept = naccept;
		best->trace = trace;
		best->trace_accept = tacc;
		best->ntrace = ntrace;
	} else { free(opt); free(trace); free(tacc); }
	for (n = 0; n < nDrives; n++) free(test[n]);
	free(test); free(skip); free(pre);
	g_cache_on = cache_was;
	return best;
}

/* cnSearchHist: the loop of RCatnetSearchHist::search (rcatnet_hist.cpp:282-548).  Batches of numThreads fresh random
 * orders (duplicates inside a batch skipped, :300-316), one search per order, network picked by AIC / BIC / highest
 * complexity (:463-503), weight 1 / exp(loglik/(M N)) / exp(score/(M N)) (:510-515) added to
 * hist[parent * N + child] in label space (:517-528).  Returns the iteration count (:530). */
int orc_search_hist(int N, int M, const int *samples_lbl, const int *pert_lbl, int maxParentSet,
		const int *parentSizes_lbl, int maxComplexity, const int *numCats_lbl, const int *parentsPool_lbl,
		const int *fixedPool_lbl, int scoreSel, int weight, int maxIter, int numThreads, double *hist) {
	int nDrives = numThreads < 1 ? 1 : numThreads, n, nn, k, i, j;
	int **test = (int **)malloc(nDrives * sizeof(int *)), *skip = (int *)malloc(nDrives * sizeof(int));
	for (n = 0; n < nDrives; n++) test[n] = (int *)malloc(N * sizeof(int));
	memset(hist, 0, sizeof(double) * N * N);
	int niter = 0, nstop = 0;
	const int cache_was = g_cache_on;
	if (nDrives > 8) g_cache_on = 0;                                 /* rcatnet_hist.cpp:205-209 */
	while (niter < maxIter) {
		for (n = 0; n < nDrives; n++) {
			gen_permutation(test[n], N);
			skip[n] = 0;
			for (nn = 0; nn < n && !skip[n]; nn++) skip[n] = !memcmp(test[nn], test[n], N * sizeof(int));
		}
		nstop = 1;
		for (n = 0; n < nDrives; n++) {
			if (skip[n]) continue;
			nstop = 0;
			orc_result *r = orc_search_order(N, M, samples_lbl, pert_lbl, maxParentSet, parentSizes_lbl, maxComplexity,
					test[n], numCats_lbl, parentsPool_lbl, fixedPool_lbl, NULL);
			int pick = -1;
			double fl = 0;
			if (r && (scoreSel == 1 || scoreSel == 2)) {
				double ft = 0.5 * log((double)M);
				fl = NEG_INF;
				for (k = 0; k < r->nslots; k++) {
					if (!r->nets[k].exists) continue;
					double v = scoreSel == 1 ? M * net_loglik(r, &r->nets[k]) - k : M * net_loglik(r, &r->nets[k]) - ft * k;
					if (v > fl) { fl = v; pick = k; }
				}
			}
			if (r && pick < 0) for (k = r->nslots - 1; k >= 0; k--) if (r->nets[k].exists) { pick = k; break; }
			if (!r) { niter++; continue; }                       /* no slots at all: the iteration still counts (:530) */
			if (pick < 0) { orc_result_free(r); continue; }      /* slots but no network: `continue` before niter++ (:506) */
			double w;
			if (weight <= 0) w = 1;
			else if (weight == 1) w = exp(net_loglik(r, &r->nets[pick]) / (double)(M * N));
			else w = exp(fl / (double)(M * N));
			for (j = 0; j < N; j++) {
				const fam_t *f = &r->fams[r->nets[pick].fam[j]];
				for (i = 0; i < f->d; i++) hist[(test[n][f->pars[i]] - 1) * N + test[n][j] - 1] += w;
			}
			orc_result_free(r);
			niter++;
		}
		if (nstop) break;
	}
	for (n = 0; n < nDrives; n++) free(test[n]);
	free(test); free(skip);
	g_cache_on = cache_was;
	return niter;
}

int orc_result_sa_info(const orc_result *r, int *order_out, int *niter, int *naccept) {
	if (!r || !r->order) return -1;
	memcpy(order_out, r->order, r->N * sizeof(int));
	*niter = r->niter;
	*naccept = r->naccept;
	return r->ntrace;
}
int orc_result_sa_trace(const orc_result *r, double *trace, int *accept, int cap) {
	for (int i = 0; i < r->ntrace && i < cap; i++) { trace[i] = r->trace[i]; accept[i] = r->trace_accept[i]; }
	return r->ntrace;
}

/* ------------------------------------------------------------------ pairwise metrics ---- */
/* catnet_entropy.cpp:38-911 (catnetEntropyPairwise, catnetEntropyOrder, catnetKLpairwise, catnetPearsonPairwise).
 * All four are functions of two kinds of two-way tables per ordered node pair (n1, n2):
 *   kept[i2][i1] = samples with x_n2 = i2, x_n1 = i1 among those where n1 is NOT perturbed (every sample without
 *                  perturbations; :144-157), and
 *   full[i2][i1] = the same over all samples (:609-611, :832-834).
 * Categories are the value range of the node, code = value - min (:63-107).  No missing values (the reference
 * decrements NA_INTEGER, :61-63, which is undefined); this restatement returns -1 for any value < 1.
 * out[n2*N + n1] as the reference fills klmat.  kind 0 entropy, 1 KL, 2 Pearson, 3 entropy order (order_out, 1-based). */
static void orc_pair_table(const int *code, const int *pert, int N, int M, int n1, int n2, int C, int masked, int *tab) {
	int j;
	memset(tab, 0, (size_t)C * C * sizeof(int));
	for (j = 0; j < M; j++) {
		if (masked && pert && pert[(size_t)j * N + n1]) continue;
		if (n1 == n2) tab[code[(size_t)j * N + n1]]++;                                  /* :165-167 marginal */
		else tab[C * code[(size_t)j * N + n2] + code[(size_t)j * N + n1]]++;             /* :184-185 */
	}
}

/* :163-203 -- conditional entropy H(n1 | n2) (n1 == n2: marginal entropy) from integer counts */
static double orc_entropy_cell(const int *tab, int C, int c1, int c2, int same) {
	double floglik = 0, fsum = 0, faux;
	int i, j;
	if (same) {
		for (j = 0; j < c1; j++) {
			fsum += tab[j];
			if (tab[j] > 0) floglik += tab[j] * (double)log((double)tab[j]);
		}
		if (fsum > 0) {
			floglik -= fsum * (double)log((double)fsum);
			floglik /= fsum;
		}
		return -floglik;
	}
	for (i = 0; i < c2; i++) {
		faux = 0;
		for (j = 0; j < c1; j++) {
			faux += tab[C * i + j];
			if (tab[C * i + j] > 0) floglik += tab[C * i + j] * (double)log((double)tab[C * i + j]);
		}
		fsum += faux;
		if (faux > 0) floglik -= faux * (double)log((double)faux);
	}
	if (fsum > 0) floglik /= fsum;
	return -floglik;
}

/* row-normalise a table of doubles (:613-634, :836-846) */
static void orc_row_normalise(double *p, int C, int c1, int c2) {
	int i, j;
	for (i = 0; i < c2; i++) {
		double fsum = 0, faux;
		for (j = 0; j < c1; j++) fsum += p[C * i + j];
		if (fsum <= 0) continue;
		faux = 1 / fsum;
		for (j = 0; j < c1; j++) p[C * i + j] *= faux;
	}
}

/* the reference's _order (utils.h:97-133) for decreasing = 1, holes and all: equal values mark each other used, so
 * a run of t equal ranks fills one position (with the LAST of the tied indices) and leaves t-1 positions at -1 */
static int orc_cmp_double(const void *a, const void *b) {
	double x = *(const double *)a, y = *(const double *)b;
	return x < y ? -1 : x > y;
}
static void orc_order_decreasing(const double *v, int n, int *porder) {
	double *sorted = (double *)malloc((size_t)n * sizeof(double)), *aux = (double *)malloc((size_t)n * sizeof(double));
	int i, j;
	memcpy(sorted, v, (size_t)n * sizeof(double));
	memcpy(aux, v, (size_t)n * sizeof(double));
	qsort(sorted, (size_t)n, sizeof(double), orc_cmp_double);      /* _quick_sort: ascending */
	if (n < 2) { free(sorted); free(aux); return; }                /* utils.h:98-99: porder untouched */
	for (j = 1; j < n; j++) porder[j] = -1;
	for (j = n - 1; j >= 0; j--)
		for (i = 0; i < n; i++)
			if (aux[i] == sorted[j]) {
				porder[n - 1 - j] = i;
				aux[i] = (double)FLT_MAX;
			}
	free(sorted);
	free(aux);
}

int orc_pairwise(int kind, int N, int M, const int *samples, const int *pert, double *out, int *order_out) {
	int *code = (int *)malloc((size_t)N * M * sizeof(int)), *cats = (int *)malloc((size_t)N * sizeof(int));
	int i, j, n1, n2, C = 1;
	for (i = 0; i < N; i++) {                                      /* :81-107 */
		int lo = INT_MAX, hi = -INT_MAX;
		for (j = 0; j < M; j++) {
			int v = samples[(size_t)j * N + i];
			if (v < 1) { free(code); free(cats); return -1; }
			if (v < lo) lo = v;
			if (v > hi) hi = v;
		}
		cats[i] = hi - lo + 1;
		for (j = 0; j < M; j++) code[(size_t)j * N + i] = samples[(size_t)j * N + i] - lo;
		if (cats[i] > C) C = cats[i];
	}
	int *tab = (int *)malloc((size_t)C * C * sizeof(int));
	double *p1 = (double *)malloc((size_t)C * C * sizeof(double)), *p2 = (double *)malloc((size_t)C * C * sizeof(double));
	double *mat = (double *)calloc((size_t)N * N, sizeof(double));
	for (n1 = 0; n1 < N; n1++)
		for (n2 = 0; n2 < N; n2++) {
			if (kind == 0 || kind == 3) {
				orc_pair_table(code, pert, N, M, n1, n2, C, 1, tab);
				mat[(size_t)n2 * N + n1] = orc_entropy_cell(tab, C, cats[n1], cats[n2], n1 == n2);
				continue;
			}
			if (n1 == n2) continue;                                /* :604-605, :827-828 */
			/* without perturbations the "perturbed sub-sample" is EMPTY here (numsamplesPert = 0, :594-602), not
			 * the whole sample as in the entropy functions */
			if (pert) orc_pair_table(code, pert, N, M, n1, n2, C, 1, tab);
			else memset(tab, 0, (size_t)C * C * sizeof(int));
			for (i = 0; i < C * C; i++) p1[i] = tab[i];
			orc_pair_table(code, pert, N, M, n1, n2, C, 0, tab);
			for (i = 0; i < C * C; i++) p2[i] = tab[i];
			double floglik = 0;
			if (kind == 1) {                                       /* :613-647 */
				orc_row_normalise(p1, C, cats[n1], cats[n2]);
				orc_row_normalise(p2, C, cats[n1], cats[n2]);
				for (i = 0; i < cats[n2]; i++)
					for (j = 0; j < cats[n1]; j++) {
						double a = p1[C * i + j], b = p2[C * i + j];
						if (a > 0 && b > 0) floglik += a * (double)log((double)a / (double)b);
						else if (a != 0 && b == 0) floglik = FLT_MAX;
					}
			} else {                                               /* :836-864 */
				orc_row_normalise(p2, C, cats[n1], cats[n2]);
				for (i = 0; i < cats[n2]; i++) {
					double fsum = 0, faux;
					for (j = 0; j < cats[n1]; j++) fsum += p1[C * i + j];
					if (fsum <= 0) continue;
					for (j = 0; j < cats[n1]; j++) {
						double b = p2[C * i + j];
						faux = p1[C * i + j] - fsum * b;
						if (b > 0) floglik += (double)(faux * faux) / (double)(fsum * b);
						else if (faux != 0 && b == 0) floglik = FLT_MAX;
					}
				}
			}
			mat[(size_t)n2 * N + n1] += floglik;
		}
	if (kind == 3) {                                               /* :428-440 */
		double *ranks = (double *)malloc((size_t)N * sizeof(double));
		for (n1 = 0; n1 < N; n1++) {
			double faux = mat[(size_t)n1 * N + n1], s = 0;
			for (n2 = 0; n2 < N; n2++) {
				mat[(size_t)n2 * N + n1] = faux - mat[(size_t)n2 * N + n1];
				s += mat[(size_t)n2 * N + n1];
			}
			ranks[n1] = s;
		}
		/* porder comes from CATNET_MALLOC without initialisation; position 0 is always written for N >= 2 */
		for (i = 0; i < N; i++) order_out[i] = 0;
		orc_order_decreasing(ranks, N, order_out);
		for (i = 0; i < N; i++) order_out[i]++;
		free(ranks);
	} else
		memcpy(out, mat, (size_t)N * N * sizeof(double));
	free(code); free(cats); free(tab); free(p1); free(p2); free(mat);
	return 0;
}

/* ------------------------------------------------------------------ fixed network ---- */
/* cnLoglik / cnNodeLoglik / cnSetProb on a given network.  The reference's .Call glue for these is
 * missing from the snapshot; what is restated here are the CATNET methods it forwards to.
 * Network: cats[N], numParents[N], parents0[N][maxParents] (0-based, listed order: first parent = most significant
 * table index, problist.h:252-263), tables concatenated by offsets[N+1] (child category fastest).
 * samples0 [M][N] zero-based codes, outside [0, cats) = missing. */
static int orc_net_cell(const int *row, const int *cats, int node, const int *pars, int d) {
	int idx = 0, i;
	for (i = 0; i < d; i++) {
		int v = row[pars[i]];
		if (v < 0 || v >= cats[pars[i]]) return -1;              /* find_slot returns 0 -> sample skipped */
		idx = idx * cats[pars[i]] + v;
	}
	int x = row[node];
	if (x < 0 || x >= cats[node]) return -1;
	return idx * cats[node] + x;
}

/* mode 0: CATNET::sampleLoglikVector (catnet_class.h:616-687) -> out[N]: per node, the log-probabilities of its
 *         unperturbed, complete samples summed IN SAMPLE ORDER, -FLT_MAX at the first zero probability, divided by
 *         the count if > 1.
 * mode 1: CATNET::bySampleLoglikVector (:689-765) -> out[M]: nodes outer, samples inner; a zero probability sets
 *         that sample's value to -FLT_MAX and ENDS the node's pass over the samples (the `break` leaves the sample
 *         loop), later nodes keep adding to every sample.
 * mode 2: CATNET::sampleNodeLoglik (:767-815) for every node -> out[N] (= mode 0 without perturbations).
 * mode 3: CATNET::sampleLoglik (:551-614) -> out[0]: sum over nodes in index order, -FLT_MAX as soon as one node is. */
int orc_net_loglik(int N, int M, const int *samples0, const int *pert, const int *cats, int maxParents,
		const int *numParents, const int *parents0, const double *probs, const long long *offsets, int mode,
		double *out) {
	int i, j;
	if (mode == 1) memset(out, 0, (size_t)M * sizeof(double));
	double total = 0;
	for (i = 0; i < N; i++) {
		const int *pars = parents0 + (size_t)i * maxParents;
		const double *tab = probs + offsets[i];
		double acc = 0;
		int ncount = 0;
		for (j = 0; j < M; j++) {
			if (pert && mode < 2 && pert[(size_t)j * N + i]) continue;
			int cell = orc_net_cell(samples0 + (size_t)j * N, cats, i, pars, numParents[i]);
			if (cell < 0) continue;
			if (mode == 1) {
				if (tab[cell] > 0) out[j] += log(tab[cell]);
				else { out[j] = NEG_INF; break; }
				continue;
			}
			if (tab[cell] > 0) acc += log(tab[cell]);
			else { acc = NEG_INF; break; }
			ncount++;
		}
		if (mode == 1) continue;
		if (ncount > 1 && acc > NEG_INF) acc /= (double)ncount;
		if (mode == 3) {
			if (acc == NEG_INF) { total = NEG_INF; break; }
			total += acc;
		} else
			out[i] = acc;
	}
	if (mode == 3) out[0] = total;
	return 0;
}

/* cnSetProb: CATNET::setNodeSampleProb(node, samples, M, bNormalize = 1) (catnet_class.h:818-879) for every node, on
 * the sub-sample where the node is not perturbed; PROB_LIST::normalize (problist.h:193-222): rows without a sample
 * become uniform with the last cell 1 - sum.  An empty sub-sample gives -FLT_MAX, size 0 and uniform rows. */
int orc_net_set_prob(int N, int M, const int *samples0, const int *pert, const int *cats, int maxParents,
		const int *numParents, const int *parents0, const long long *offsets, double *probs_out, double *loglik_out,
		int *sizes_out) {
	int i, j, k;
	for (i = 0; i < N; i++) {
		const int *pars = parents0 + (size_t)i * maxParents;
		double *tab = probs_out + offsets[i];
		int size = (int)(offsets[i + 1] - offsets[i]), cc = cats[i], ncount = 0, m = 0;
		memset(tab, 0, (size_t)size * sizeof(double));
		for (j = 0; j < M; j++) {
			if (pert && pert[(size_t)j * N + i]) continue;
			m++;
			int cell = orc_net_cell(samples0 + (size_t)j * N, cats, i, pars, numParents[i]);
			if (cell < 0) continue;
			tab[cell] += 1;
			ncount++;
		}
		double ll = 0;
		for (k = 0; k < size; k += cc) {                          /* problist.h:224-250 */
			double pp = 0;
			for (j = 0; j < cc; j++) pp += tab[k + j];
			if (pp <= 0) continue;
			pp = 1 / pp;
			for (j = 0; j < cc; j++)
				if (tab[k + j] > 0) ll += tab[k + j] * log(tab[k + j] * pp);
		}
		if (ncount > 1) ll /= (double)ncount;
		loglik_out[i] = m < 1 ? NEG_INF : ll;
		sizes_out[i] = m < 1 ? 0 : ncount;
		for (k = 0; k < size; k += cc) {                          /* problist.h:193-222 */
			double pp = 0;
			for (j = 0; j < cc; j++) pp += tab[k + j];
			if (pp > 0) {
				pp = 1 / pp;
				for (j = 0; j < cc; j++) tab[k + j] *= pp;
			} else {
				double favg = (double)1 / (double)cc;
				pp = 0;
				for (j = 0; j < cc - 1; j++) { tab[k + j] = favg; pp += favg; }
				tab[k + cc - 1] = 1 - pp;
			}
		}
	}
	return 0;
}

/* ------------------------------------------------------------------ cnSamples ---- */
/* RCatnet::genSamples (rcatnet.cpp:524-650) under the injected uniform stream (orc_unif = splitmix64, as the SA
 * restatement above).  Sampling order = CATNET::getOrder (catnet_class.h:1119-1157).  out [M][N] 1-based, INT_MIN = NA.
 * Returns the number of uniforms consumed, -1 for a cyclic network. */
static double orc_stream(unsigned long long *state) {
	unsigned long long z = (*state += 0x9E3779B97F4A7C15ull);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	z = z ^ (z >> 31);
	return ((double)(z >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

long orc_net_samples(int N, const int *cats, int maxParents, const int *numParents, const int *parents0,
		const double *probs, const long long *offsets, int M, const int *pert, double naRate, unsigned long long seed,
		int *out) {
	int *order = (int *)malloc((size_t)N * sizeof(int)), *found = (int *)calloc((size_t)N, sizeof(int));
	int i, j, k, p;
	long calls = 0;
	unsigned long long state = seed;
	for (k = 0; k < N; k++) {                                      /* getOrder */
		for (j = 0; j < N; j++) {
			if (found[j]) continue;
			int ready = 1;
			for (p = 0; p < numParents[j]; p++)
				if (!found[parents0[(size_t)j * maxParents + p]]) { ready = 0; break; }
			if (ready) break;
		}
		if (j >= N) { free(order); free(found); return -1; }
		order[k] = j;
		found[j] = 1;
	}
	for (k = 0; k < N; k++) {                                      /* :569-599 */
		const int node = order[k];
		const int *pars = parents0 + (size_t)node * maxParents;
		for (j = 0; j < M; j++) {
			int *row = out + (size_t)j * N;
			if (pert) {
				int v = pert[(size_t)j * N + node];
				if (v >= 1 && v <= cats[node]) { row[node] = v; continue; }
			}
			long long idx = 0;
			for (p = 0; p < numParents[node]; p++) idx = idx * cats[pars[p]] + (row[pars[p]] - 1);
			const double *tab = probs + offsets[node] + idx * cats[node];
			double u = orc_stream(&state), v = 0;
			calls++;
			for (i = 0; i < cats[node]; i++) {
				v += tab[i];
				if (u <= v) break;
			}
			row[node] = i + 1;
		}
	}
	k = (int)(naRate * N);                                         /* :601-626 */
	if (k > 0 && k < N) {
		int *aux = (int *)malloc((size_t)N * sizeof(int));
		for (j = 0; j < M; j++) {
			for (i = 0; i < N; i++) { aux[i] = (int)(N * orc_stream(&state)); calls++; }
			for (p = 0; p < k; p++) {
				int node = 0, fmax = -(int)RAND_MAX;
				for (i = 0; i < N; i++)
					if (fmax < aux[i]) { fmax = aux[i]; node = i; }
				aux[node] = -(int)RAND_MAX;
				out[(size_t)j * N + node] = INT_MIN;
			}
		}
		free(aux);
	}
	free(order);
	free(found);
	return calls;
}
