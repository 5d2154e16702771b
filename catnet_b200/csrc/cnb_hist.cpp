/* cnb_hist.cpp -- cnSearchHist: parent histogram over random node orders, host driver around the device per-order
 * evaluation (the caller right next to cnSearchSA on the same path).
 *
 * Restates the loop of RCatnetSearchHist::search (/root/reference/src/rcatnet_hist.cpp:282-548):
 *   orders      batches of numThreads fresh random permutations (_gen_permutation, utils.h:144-180), duplicates
 *               inside a batch skipped (:300-316)
 *   evaluation  one CATNET_SEARCH2::estimate per order -- here plan + K2..K5 on the device (cnb_ctx_search_light)
 *   selection   AIC / BIC / highest complexity (:463-503)
 *   weight      1, exp(loglik / (M N)) or exp(score / (M N)) (:510-515)
 *   histogram   hist[parent label][child label] += weight for every edge of the selected network (:517-528)
 * The reference evaluates the orders of a batch on pthreads and memoises family scores in a process-wide cache;
 * the device evaluates them side by side and re-scores.  With use_cache (what R passes) a host model of the reference's
 * cache walks the orders one after the other, so that the selected networks and their log-likelihoods -- the histogram's
 * weights -- are those of the reference with its cache on (cnb_cache.h, cnb_sa.cpp).  The uniform stream is drawn in the
 * reference's order.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <memory>
#include <vector>
#include "cnb_ctx.h"
#include "cnb_rng.h"

#define RC(call) do { int rc_ = (call); if (rc_ != CNB_OK) return rc_; } while (0)

void cnb_gen_permutation(std::vector<int> &out, int n, CnbRng &rng);      /* cnb_sa.cpp */

/* one order on one device: evaluate, pick the network (AIC / BIC / highest complexity), its weight and parents.
 * *found = false when the search left no network (`continue` before niter++, rcatnet_hist.cpp:506). */
int cnb_hist_eval_order(cnb_ctx *c, const cnb_params *q, const cnb_hist_params *hp, CnbHistEval *ev) {
	RC(cnb_ctx_search_light(c, q));
	return cnb_hist_pick(c, hp, ev);
}

/* the second half: the context holds an evaluated order */
int cnb_hist_pick(cnb_ctx *c, const cnb_hist_params *hp, CnbHistEval *ev) {
	const int N = c->N, M = c->M;
	std::vector<int32_t> cx;
	std::vector<double> ll;
	ev->found = false;
	RC(cnb_ctx_dp_summary(c, &cx, &ll));
	if (cx.empty()) return CNB_OK;
	int pick = -1;
	double fLogLik = 0;
	if (hp->score == 1) {                    /* AIC :464-476 */
		fLogLik = -(double)FLT_MAX;
		for (size_t i = 0; i < cx.size(); i++) {
			double v = M * ll[i] - cx[i];
			if (v > fLogLik) { fLogLik = v; pick = (int)i; }
		}
	} else if (hp->score == 2) {             /* BIC :477-489 */
		fLogLik = -(double)FLT_MAX;
		double ft = 0.5 * log((double)M);
		for (size_t i = 0; i < cx.size(); i++) {
			double v = M * ll[i] - ft * cx[i];
			if (v > fLogLik) { fLogLik = v; pick = (int)i; }
		}
	}
	if (pick < 0) pick = (int)cx.size() - 1; /* highest complexity :492-500 */
	if (hp->weight <= 0) ev->w = 1;
	else if (hp->weight == 1) ev->w = exp(ll[pick] / (double)(M * N));
	else ev->w = exp(fLogLik / (double)(M * N));
	RC(cnb_ctx_network_parents(c, cx[pick], &ev->npar, &ev->pars));
	ev->found = true;
	return CNB_OK;
}

/* hist[parent label][child label] += weight for every edge of the selected network (:517-528) */
void cnb_hist_add(double *hist, int N, const std::vector<int> &order, const CnbHistEval &ev) {
	for (int j = 0; j < N; j++)
		for (int i = 0; i < ev.npar[j]; i++)
			hist[(size_t)(order[ev.pars[(size_t)j * CNB_MAXP + i]] - 1) * N + order[j] - 1] += ev.w;
}

extern "C" int cnb_ctx_search_hist(cnb_ctx *c, const cnb_params *p, const cnb_hist_params *hp, double *hist,
		int32_t *iterations) {
	if (!c || !p || !hp || !hist) return CNB_ERR_PARAM;
	if (!c->have_samples) { cnb_set_error("context has no samples"); return CNB_ERR_PARAM; }
	const int N = c->N;
	const int nDrives = hp->num_threads < 1 ? 1 : hp->num_threads;
	memset(hist, 0, sizeof(double) * (size_t)N * N);
	CnbRng rng(hp->seed ? hp->seed : cnb_default_seed(), hp->unif, hp->unif_ctx);
	std::vector<std::vector<int>> test(nDrives, std::vector<int>(N));
	std::vector<char> skip(nDrives);
	CnbHistEval ev;
	cnb_params q = *p;
	q.edge_prob = nullptr;                       /* rcatnet_hist.cpp:268-277: no edge priors on this entry point */
	int niter = 0, nstop = 0;
	/* the orders of a batch are independent: evaluated concurrently on sibling contexts when the data set is small */
	std::vector<cnb_ctx *> lanes, where(nDrives, c);
	RC(cnb_ctx_batch(c, nDrives, &lanes));
	std::unique_ptr<CnbScoreCache> cache;
	if (cnb_cache_wanted(hp->use_cache, nDrives)) cache.reset(new CnbScoreCache(N, p->max_parent_set));
	const bool batched = (int)lanes.size() >= nDrives && nDrives > 1 && !cache;
	struct CacheGuard {
		cnb_ctx *c; std::vector<cnb_ctx *> &l;
		~CacheGuard() { c->cache = nullptr; c->quiet = false; for (cnb_ctx *s : l) s->quiet = false; }
	} guard{c, lanes};
	c->cache = cache.get();                      /* every order passes through the model, one after the other */
	const bool quiet = !getenv("CNB_SA_TIMING"); /* no per-phase timing events: fewer driver calls per evaluated order */
	c->quiet = quiet;
	for (cnb_ctx *s : lanes) s->quiet = quiet;
	std::vector<const int32_t *> todo_orders;
	while (niter < hp->max_iter) {
		for (int n = 0; n < nDrives; n++) {
			cnb_gen_permutation(test[n], N, rng);
			skip[n] = 0;
			for (int nn = 0; nn < n && !skip[n]; nn++) skip[n] = (test[nn] == test[n]);
		}
		nstop = 1;
		if (batched) {
			todo_orders.clear();
			int t = 0;
			for (int n = 0; n < nDrives; n++)
				if (!skip[n]) { where[n] = lanes[t++]; todo_orders.push_back(test[n].data()); }
			RC(cnb_ctx_batch_run(c, lanes, &q, t, todo_orders.data()));
		}
		for (int n = 0; n < nDrives; n++) {
			if (skip[n]) continue;
			nstop = 0;
			if (batched) RC(cnb_hist_pick(where[n], hp, &ev));
			else {
				q.order = test[n].data();
				RC(cnb_hist_eval_order(c, &q, hp, &ev));
			}
			if (!ev.found) continue;                 /* slots but no network: `continue` before niter++ (:506) */
			cnb_hist_add(hist, N, test[n], ev);
			niter++;
			if (p->echo) fprintf(stderr, "Iteration %d\\%d\n", niter, hp->max_iter);
		}
		if (nstop) break;
	}
	if (iterations) *iterations = niter;
	return CNB_OK;
}

extern "C" int cnb_search_hist(const cnb_params *p, const cnb_hist_params *hp, double *hist, int32_t *iterations) {
	if (!p || !hp || !hist) return CNB_ERR_PARAM;
	if (cnb_multi_devices(p) > 1 || cnb_multi_devices(p) < -1) return cnb_multi_search_hist(p, hp, hist, iterations);
	cnb_ctx *c = nullptr;
	RC(cnb_ctx_acquire(p->device, &c));
	const char *fb = getenv("CNB_FORCE_BITS");                /* kernel-selection switches for fuzzing (cnb_search_order, cnb_api.cu) */
	cnb_ctx_force_generic(c, fb ? atoi(fb) : 0);
	int rc = cnb_ctx_set_samples(c, p->num_nodes, p->num_samples, p->samples, p->perturbations, p->num_cats);
	if (rc == CNB_OK) rc = cnb_ctx_search_hist(c, p, hp, hist, iterations);
	cnb_ctx_park(c);
	return rc;
}
