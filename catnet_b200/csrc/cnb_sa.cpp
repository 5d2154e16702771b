/* cnb_sa.cpp -- simulated annealing over node orders, host driver around the device per-order evaluation.
 *
 * Restates the loop of RCatnetSearchSA::search (/root/reference/src/rcatnet_sa.cpp:471-820):
 *   proposals   _genOrder (:123-184) / _genOrderFormDirProbs (:186-251), duplicates inside a batch skipped (:491-506)
 *   evaluation  one CATNET_SEARCH2::estimate per proposal -- here plan + K2..K5 on the device (cnb_ctx_search_light)
 *   selection   AIC / BIC / highest complexity (:671-713)
 *   acceptance  first evaluation unconditionally (:714-737), then Metropolis (:738-798); the first accepted
 *               proposal of a batch wins and the later ones are discarded (:655-659, :784-786)
 *   cooling     every tempCheckOrders iterations (:807-814)
 * The reference evaluates the numThreads proposals of a batch on pthreads; the device evaluates them side by side on
 * sibling contexts.  The uniform stream is drawn in the reference's order.
 *   score cache with use_cache (what R passes): the reference memoises family scores in a process-wide table and a hit
 *               returns the family as it was listed and summed under the order that stored it (cache.cpp:142-216).  A host
 *               model of that table (cnb_cache.h) walks every evaluated order; the device re-scores the candidates the cache
 *               answers differently.  Every proposal of a step passes through the cache, accepted or not, one after the
 *               other (the reference's threads interleave their stores as they come: with more than one thread its own
 *               result depends on timing, and this is the interleaving "proposal after proposal").
 */
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <string.h>
#include <memory>
#include <chrono>
#include <vector>
#include "cnb_ctx.h"
#include "cnb_rng.h"

#define RC(call) do { int rc_ = (call); if (rc_ != CNB_OK) return rc_; } while (0)

/* utils.h:144-180 */
void cnb_gen_permutation(std::vector<int> &out, int n, CnbRng &rng);
static void gen_permutation(std::vector<int> &out, int n, CnbRng &rng) { cnb_gen_permutation(out, n, rng); }
void cnb_gen_permutation(std::vector<int> &out, int n, CnbRng &rng) {
	out.resize(n);
	std::vector<int> aux(n);
	int cc = 0, brep = 0;
	while (++cc < 1e5) {
		brep = 0;
		for (int i = 0; i < n; i++) aux[i] = (int)(n * n * rng.unif());
		for (int j = 0; j < n; j++) {
			out[j] = 0;
			for (int i = 0; i < n; i++) {
				if (i != j && aux[j] == aux[i]) { brep = 1; break; }
				if (aux[j] >= aux[i]) out[j]++;
			}
			if (brep) break;
		}
		if (!brep) break;
	}
	if (cc >= 1e5 - 1)
		for (int j = 0; j < n; j++) out[j] = j + 1;
}

/* rcatnet_sa.cpp:123-184: `shuffles` moves of one node from position n1 to n2 (adjacent, or anywhere if bjump) */
static void gen_order(const std::vector<int> &cur, std::vector<int> &out, int shuffles, int bjump, CnbRng &rng) {
	const int n = (int)cur.size();
	out.resize(n);
	if (shuffles <= 0) {
		gen_permutation(out, n, rng);
		return;
	}
	out = cur;
	std::vector<int> t(n);
	for (int sh = 0; sh < shuffles; sh++) {
		t = out;
		int n1 = (int)(rng.unif() * n), n2;
		if (bjump) {
			n2 = n1;
			while (n2 == n1) n2 = (int)(rng.unif() * n);
		} else {
			n2 = n1 + 1;
			if (n1 >= n - 1) n2 = 0;
		}
		if (n1 < n2) {
			for (int i = n1; i < n2; i++) out[i] = t[i + 1];
		} else {
			for (int i = n2; i < n1; i++) out[i + 1] = t[i];
		}
		out[n2] = t[n1];
	}
}

/* rcatnet_sa.cpp:186-251: build an order node by node from pairwise direction probabilities */
static void gen_order_dirprobs(int n, const double *mat, std::vector<int> &out, double *order_prob, CnbRng &rng) {
	std::vector<int> ord(n, 0), tmp(n);
	std::vector<double> probs(n);
	double op = 1;
	for (int k = 1; k < n; k++) {
		double fsum = 0;
		for (int i = 0; i <= k; i++) {
			double f = 1;
			for (int j = 0; j < i; j++) f *= mat[ord[j] * n + k];
			for (int j = i; j < k; j++) f *= mat[k * n + ord[j]];
			probs[i] = f;
			fsum += f;
		}
		double u = fsum * rng.unif();
		fsum = 0;
		int i;
		for (i = 0; i < k; i++) {
			fsum += probs[i];
			if (fsum >= u) break;
		}
		op *= probs[i];
		for (int j = 0; j < i; j++) tmp[j] = ord[j];
		tmp[i] = k;
		for (int j = i; j < k; j++) tmp[j + 1] = ord[j];
		for (int j = 0; j <= k; j++) ord[j] = tmp[j];
	}
	out.resize(n);
	for (int i = 0; i < n; i++) out[i] = ord[i] + 1;
	*order_prob = op > 0 ? log(op) : -(double)FLT_MAX;
}

extern "C" int cnb_ctx_search_sa(cnb_ctx *c, const cnb_params *p, const cnb_sa_params *sa, cnb_result **out) {
	if (!c || !p || !sa || !out || !sa->start_order) return CNB_ERR_PARAM;
	*out = nullptr;
	const int N = c->N, M = c->M;
	int nDrives = sa->num_threads < 1 ? 1 : sa->num_threads;
	double orderShuffles = sa->order_shuffles;
	int bjump = 0;
	if (orderShuffles < 0) { bjump = 1; orderShuffles = -orderShuffles; }        /* :371-377 */
	int minShuffles = (int)orderShuffles;
	orderShuffles -= minShuffles;
	std::vector<int> optOrder(sa->start_order, sa->start_order + N);
	for (int i = 0; i < N; i++)
		if (optOrder[i] <= 0 || optOrder[i] > N) { cnb_set_error("Invalid startOrder parameter"); return CNB_ERR_PARAM; }
	CnbRng rng(sa->seed ? sa->seed : cnb_default_seed(), sa->unif, sa->unif_ctx);
	std::vector<std::vector<int>> test(nDrives);
	std::vector<double> newOrdProb(nDrives, 0.0);
	std::vector<char> skip(nDrives);
	std::unique_ptr<cnb_result> best;
	std::vector<double> trace;
	std::vector<int32_t> trace_acc;
	bool haveOpt = false;
	double fOptLogLik = -(double)FLT_MAX, ordProb = 0, tempCur = sa->temp_start;
	int nstop = 0, nstopCounter = 0, naccept = 0, nLastChangedTemp = 0, niter = 0;
	std::vector<int32_t> cx;
	std::vector<double> ll;
	cnb_params q = *p;
	/* the numThreads proposals of a step are independent evaluations (the reference hands one to each worker thread,
	 * rcatnet_sa.cpp:475-652): here they run concurrently on this device, one sibling context and host thread each, when the
	 * data set is small enough to copy; the decisions below then look at them in proposal order, exactly as before */
	std::vector<cnb_ctx *> lanes;
	RC(cnb_ctx_batch(c, nDrives, &lanes));
	std::unique_ptr<CnbScoreCache> cache;
	if (cnb_cache_wanted(sa->use_cache, nDrives)) cache.reset(new CnbScoreCache(N, p->max_parent_set));
	const bool have_lanes = (int)lanes.size() >= nDrives && nDrives > 1;
	const bool batched = have_lanes && !cache;
	/* cache mode without sibling contexts and several proposals per step: each evaluation is exported before the next one
	 * overwrites the context */
	const bool eager = cache && !have_lanes && nDrives > 1;
	const char *no_pipe = getenv("CNB_CACHE_NO_PIPELINE");            /* A/B switch: the proposals of a step strictly one after the other */
	const bool pipelined = !no_pipe;
	std::vector<cnb_ctx *> where(nDrives, c);
	std::vector<const int32_t *> todo_orders;
	std::vector<std::unique_ptr<cnb_result>> pre(nDrives);
	std::vector<std::vector<int32_t>> pre_cx(nDrives);
	std::vector<std::vector<double>> pre_ll(nDrives);
	struct CacheGuard {                              /* no context keeps a pointer to the model, or stays quiet, past this call */
		cnb_ctx *c; std::vector<cnb_ctx *> &l;
		~CacheGuard() { c->cache = nullptr; c->quiet = false; for (cnb_ctx *s : l) { s->cache = nullptr; s->quiet = false; } }
	} guard{c, lanes};
	/* an evaluated order costs about as much as the driver calls it makes: no per-phase timing events (CNB_SA_TIMING=1 keeps them) */
	const bool quiet = !getenv("CNB_SA_TIMING");
	c->quiet = quiet;
	for (cnb_ctx *s : lanes) s->quiet = quiet;

	while (niter < sa->max_iter) {
		for (int n = 0; n < nDrives; n++) {
			if (sa->dir_prob) {
				gen_order_dirprobs(N, sa->dir_prob, test[n], &newOrdProb[n], rng);
			} else {
				int extra = 0;                               /* _gen_binomial(1, frac), utils.h:182-195 */
				if (orderShuffles > 0 && rng.unif() <= orderShuffles) extra = 1;
				gen_order(optOrder, test[n], minShuffles + extra, bjump, rng);
			}
			skip[n] = 0;
			for (int nn = 0; nn < n && !skip[n]; nn++) skip[n] = (test[nn] == test[n]);
		}
		nstop = 1;
		int stepiters = nDrives, stepaccept = 0;
		if (batched) {
			todo_orders.clear();
			int t = 0;
			for (int n = 0; n < nDrives; n++)
				if (!skip[n]) { where[n] = lanes[t++]; todo_orders.push_back(test[n].data()); }
			RC(cnb_ctx_batch_run(c, lanes, &q, t, todo_orders.data()));
		}
		if (cache && have_lanes && pipelined) {
			/* every proposal of the step goes through the cache, in proposal order, whether or not an earlier one is accepted.
			 * What the cache does not touch -- plan, scores, option table -- runs for all of them side by side; then each in
			 * turn is walked through the model on this thread, its overrides and its DP are queued on its own stream, and the
			 * next one's walk runs while that DP does */
			todo_orders.clear();
			int t = 0;
			for (int n = 0; n < nDrives; n++)
				if (!skip[n]) { where[n] = lanes[t++]; todo_orders.push_back(test[n].data()); }
			RC(cnb_ctx_batch_run(c, lanes, &q, t, todo_orders.data(), true));
			for (int n = 0; n < nDrives; n++) {
				if (skip[n]) continue;
				where[n]->cache = cache.get();
				int rc = cnb_ctx_dp_launch(where[n]);
				where[n]->cache = nullptr;
				if (rc != CNB_OK) {                               /* nothing may stay in flight on the sibling contexts */
					for (int k = 0; k < n; k++) if (!skip[k]) cnb_ctx_dp_wait(where[k]);
					return rc;
				}
			}
			int rc = CNB_OK;
			for (int n = 0; n < nDrives; n++)
				if (!skip[n]) { int r1 = cnb_ctx_dp_wait(where[n]); if (rc == CNB_OK) rc = r1; }
			RC(rc);
		} else if (cache) {
			int t = 0;
			for (int n = 0; n < nDrives; n++) {
				if (skip[n]) continue;
				cnb_ctx *ce = have_lanes ? lanes[t++] : c;
				where[n] = ce;
				ce->cache = cache.get();
				q.order = test[n].data();
				RC(cnb_ctx_search_light(ce, &q));
				ce->cache = nullptr;
				if (eager) {
					RC(cnb_ctx_dp_summary(ce, &pre_cx[n], &pre_ll[n]));
					cnb_result *r = nullptr;
					if (!pre_cx[n].empty()) RC(cnb_ctx_collect(ce, &r));
					pre[n].reset(r);
				}
			}
		}
		for (int n = 0; n < nDrives; n++) {
			if (skip[n]) continue;
			if (stepaccept > 0) continue;                /* a previous drive was accepted: discard the rest */
			nstop = 0;
			cnb_ctx *ce = (batched || cache) ? where[n] : c;   /* the context that holds this proposal's evaluation */
			if (!batched && !cache) {
				q.order = test[n].data();
				RC(cnb_ctx_search_light(c, &q));
			}
			if (eager) { cx = pre_cx[n]; ll = pre_ll[n]; }
			else RC(cnb_ctx_dp_summary(ce, &cx, &ll));
			if (cx.empty()) continue;
			int pick = -1;
			double fLogLik = -(double)FLT_MAX;
			if (sa->select_mode == 1) {                  /* AIC :672-684 */
				for (size_t i = 0; i < cx.size(); i++) {
					double v = M * ll[i] - cx[i];
					if (v > fLogLik) { fLogLik = v; pick = (int)i; }
				}
			} else if (sa->select_mode == 2) {           /* BIC :685-698 */
				double ft = 0.5 * log((double)M);
				for (size_t i = 0; i < cx.size(); i++) {
					double v = M * ll[i] - ft * cx[i];
					if (v > fLogLik) { fLogLik = v; pick = (int)i; }
				}
			}
			if (pick < 0) pick = (int)cx.size() - 1;     /* highest complexity :700-708 */
			fLogLik = ll[pick];
			int accepted = 0;
			if (!haveOpt) {
				accepted = 1;
				stepaccept++;
				ordProb = newOrdProb[n];
			} else {
				double delta = fLogLik - fOptLogLik;
				delta += newOrdProb[n] - ordProb;
				double acceptprob = 0;
				if (tempCur > 0 && delta <= 0) acceptprob = exp(delta / tempCur);
				double u = rng.unif();
				if (delta > 0 || u < acceptprob) {
					accepted = 1;
					ordProb = newOrdProb[n];
					stepaccept++;
					stepiters = n + 1;
					nstopCounter = 0;
				} else {
					delta = 0;
				}
				if (delta < 0) delta = -delta;
				if (delta < sa->stop_diff) nstopCounter++;
				if (nstopCounter >= sa->temp_check_orders) nstop = 1;
			}
			trace.push_back(fLogLik);
			trace_acc.push_back(accepted);
			if (accepted) {
				haveOpt = true;
				fOptLogLik = fLogLik;
				cnb_result *r = nullptr;
				if (eager) r = pre[n].release();
				else RC(cnb_ctx_collect(ce, &r));
				best.reset(r);
				optOrder = test[n];
			}
		}
		if (stepaccept > 0) naccept++;
		niter += stepiters;
		if (niter - nLastChangedTemp >= sa->temp_check_orders) {      /* :807-814 */
			tempCur *= sa->temp_cool_fact;
			nLastChangedTemp = niter;
		}
		if (p->echo) fprintf(stderr, "Iteration %d\\%d\n", niter, sa->max_iter);
		if (nstop) break;
	}
	if (!best) { cnb_set_error("No networks are found"); return CNB_ERR_PROC; }
	best->sa_order.assign(optOrder.begin(), optOrder.end());
	best->niter = niter;
	best->naccept = naccept;
	best->trace = trace;
	best->trace_accept = trace_acc;
	*out = best.release();
	return CNB_OK;
}

extern "C" int cnb_search_sa(const cnb_params *p, const cnb_sa_params *sa, cnb_result **out) {
	if (!p || !sa || !out) return CNB_ERR_PARAM;
	*out = nullptr;
	if (cnb_multi_devices(p) > 1 || cnb_multi_devices(p) < -1) return cnb_multi_search_sa(p, sa, out);      /* one chain per device: cnb_multi.cu */
	cnb_ctx *c = nullptr;
	static const bool trace = getenv("CNB_TRACE") != nullptr;
	auto now = [] { return std::chrono::steady_clock::now(); };
	auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
		return std::chrono::duration<double, std::milli>(b - a).count(); };
	auto t0 = now();
	RC(cnb_ctx_acquire(p->device, &c));
	const char *fb = getenv("CNB_FORCE_BITS");                /* kernel-selection switches for fuzzing (cnb_search_order, cnb_api.cu) */
	cnb_ctx_force_generic(c, fb ? atoi(fb) : 0);
	auto t1 = now();
	int rc = cnb_ctx_set_samples(c, p->num_nodes, p->num_samples, p->samples, p->perturbations, p->num_cats);
	auto t2 = now();
	if (rc == CNB_OK) rc = cnb_ctx_search_sa(c, p, sa, out);
	auto t3 = now();
	cnb_ctx_park(c);
	if (trace)
		fprintf(stderr, "[cnb trace] cnb_search_sa: create %.2f set_samples %.2f search %.2f destroy %.2f ms\n", ms(t0, t1), ms(t1, t2),
				ms(t2, t3), ms(t3, now()));
	return rc;
}

/* The reference keeps a process-global score cache between .Call's (cache.cpp:58-76) and R releases it before every
 * search (R/catnet.search.R:204,279, R/catnet.histo.R:119): the model of it (cnb_cache.h) lives for one cnb_search_sa /
 * cnb_search_hist call only, which is the same thing.  What the library does keep between one-shot calls is an idle
 * context per device with its device buffers (cnb_api.cu: cnb_ctx_park); this frees them. */
extern "C" void cnb_release_cache(void) { cnb_multi_release(); cnb_pool_release(); }
/* ccnSetSeed (R/catnet.def.R:743-747): the seed of searches that pass none (cnb_sa_params.seed / cnb_hist_params.seed = 0) */
static int32_t g_seed = 0;
extern "C" void cnb_set_seed(int32_t seed) { g_seed = seed; }
uint64_t cnb_default_seed(void) { return (uint64_t)(uint32_t)g_seed; }
