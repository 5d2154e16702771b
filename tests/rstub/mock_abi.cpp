/* tests/rstub/mock_abi.cpp -- the C ABI of include/catnet_b200.h answered by the ORACLE (oracle/catnet_oracle.c)
 * instead of the GPU library, for the CPU test of the R shim (tests/test_rshim.py): with it, rshim.cpp + minir.cpp run
 * end to end in an image with neither R nor a GPU, and the S4 objects the shim builds can be compared with the
 * oracle's own answer.  Test infrastructure only: nothing in the product links this file.  On a GPU box the same test
 * links the shim against the real libcatnet_b200.so instead.
 *
 * Implemented: cnb_search_order + the result accessors the shim reads, cnb_pairwise, cnb_entropy_order, cnb_loglik,
 * cnb_node_loglik, cnb_set_prob, cnb_samples, cnb_search_sa (without dirProb), cnb_search_hist, cnb_set_seed,
 * cnb_release_cache, cnb_last_error. */
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/catnet_b200.h"
extern "C" {
#include "../../oracle/catnet_oracle.h"
}

struct cnb_result {
	orc_result *h = NULL;
	int N = 0;
	std::vector<int> slots;                /* oracle slots that hold a network, ascending complexity */
	std::vector<int32_t> order, cats_pos;  /* position -> 1-based label; categories per position */
};

static int g_live_results = 0;
static std::string g_err;
static int g_released = 0, g_seed = 0;

extern "C" {

const char *cnb_last_error(void) { return g_err.c_str(); }
const char *cnb_version(void) { return "mock (oracle-backed)"; }
void cnb_release_cache(void) { g_released++; }
void cnb_set_seed(int32_t seed) { g_seed = seed; }
int mock_released(void) { return g_released; }
int mock_seed(void) { return g_seed; }

static int not_in_mock(const char *what) {
	g_err = std::string(what) + ": not in the mock";
	return CNB_ERR_PARAM;
}

/* categories per node: given, or inferred like catnet_search2.h:208-235 (max - min + 1 over the non-missing values) */
static std::vector<int32_t> mock_cats(const cnb_params *p) {
	const int N = p->num_nodes, M = p->num_samples;
	std::vector<int32_t> cats(N);
	if (p->num_cats) { cats.assign(p->num_cats, p->num_cats + N); return cats; }
	for (int i = 0; i < N; i++) {
		int lo = 0x7fffffff, hi = -0x7fffffff;
		for (int j = 0; j < M; j++) {
			int v = p->samples[(size_t)j * N + i];
			if (v < 1) continue;
			if (v < lo) lo = v;
			if (v > hi) hi = v;
		}
		cats[i] = hi >= lo ? hi - lo + 1 : 1;
	}
	return cats;
}

int cnb_search_order(const cnb_params *p, cnb_result **out) {
	const int N = p->num_nodes, M = p->num_samples;
	std::vector<int32_t> order(N);
	for (int i = 0; i < N; i++) order[i] = p->order ? p->order[i] : i + 1;
	std::vector<int32_t> cats = mock_cats(p);
	orc_result *h = orc_search_order(N, M, p->samples, p->perturbations, p->max_parent_set, p->parent_sizes,
			p->max_complexity, order.data(), p->num_cats, p->parents_pool, p->fixed_parents, p->edge_prob);
	if (!h) { g_err = "oracle: search failed"; return CNB_ERR_PARAM; }
	cnb_result *r = new cnb_result;
	g_live_results++;
	r->h = h;
	r->N = N;
	r->order = order;
	r->cats_pos.resize(N);
	for (int i = 0; i < N; i++) r->cats_pos[i] = cats[order[i] - 1];
	for (int k = 0; k < orc_result_nslots(h); k++)
		if (orc_result_exists(h, k)) r->slots.push_back(k);
	*out = r;
	return CNB_OK;
}

int32_t cnb_result_num_networks(const cnb_result *r) { return (int32_t)r->slots.size(); }
int32_t cnb_result_num_nodes(const cnb_result *r) { return r->N; }
int cnb_result_network(const cnb_result *r, int32_t i, int32_t *complexity, double *loglik) {
	int cx;
	double ll;
	orc_result_net(r->h, r->slots[i], &cx, &ll);
	*complexity = cx;
	*loglik = ll;
	return CNB_OK;
}
int cnb_result_parents(const cnb_result *r, int32_t i, int32_t *num_parents, int32_t *parents, int32_t cap) {
	for (int n = 0; n < r->N; n++) {
		int np = 0, pars[64], ss, ps;
		double ll, pr;
		orc_result_node(r->h, r->slots[i], n, &np, pars, &ll, &pr, &ss, &ps, NULL, 0);
		num_parents[n] = np;
		for (int k = 0; k < np && k < cap; k++) parents[(size_t)n * cap + k] = pars[k];
	}
	return CNB_OK;
}
int cnb_result_order(const cnb_result *r, int32_t, int32_t *order_out) {
	memcpy(order_out, r->order.data(), r->N * sizeof(int32_t));
	return CNB_OK;
}
int cnb_result_cats(const cnb_result *r, int32_t, int32_t *cats_out) {
	memcpy(cats_out, r->cats_pos.data(), r->N * sizeof(int32_t));
	return CNB_OK;
}
int cnb_result_node(const cnb_result *r, int32_t i, int32_t node, double *loglik, double *prior, int32_t *sample_size,
		double *probs, int32_t cap) {
	int np = 0, pars[64], ss = 0, ps = 0;
	std::vector<double> buf(1 << 16);
	orc_result_node(r->h, r->slots[i], node, &np, pars, loglik, prior, &ss, &ps, buf.data(), (int)buf.size());
	*sample_size = ss;
	if (probs)
		for (int k = 0; k < ps && k < cap; k++) probs[k] = buf[k];
	return ps;
}
/* test hook: results handed out and not yet freed */
int mock_live_results(void) { return g_live_results; }
void cnb_result_free(cnb_result *r) {
	if (!r) return;
	orc_result_free(r->h);
	delete r;
	g_live_results--;
}

int cnb_pairwise(int32_t kind, int32_t N, int32_t M, const int32_t *samples, const int32_t *pert, int32_t, double *out) {
	std::vector<int> order(N);
	if (orc_pairwise(kind, N, M, samples, pert, out, order.data())) { g_err = "oracle: missing or non-positive category"; return CNB_ERR_PARAM; }
	return CNB_OK;
}
int cnb_entropy_order(int32_t N, int32_t M, const int32_t *samples, const int32_t *pert, int32_t, int32_t *order_out) {
	std::vector<double> mat((size_t)N * N);
	if (orc_pairwise(3, N, M, samples, pert, mat.data(), order_out)) { g_err = "oracle: missing or non-positive category"; return CNB_ERR_PARAM; }
	return CNB_OK;
}

/* one chain; the returned networks are those of the accepted order (rcatnet_sa.cpp:873-905) */
int cnb_search_sa(const cnb_params *p, const cnb_sa_params *sa, cnb_result **out) {
	const int N = p->num_nodes;
	if (sa->dir_prob) return not_in_mock("cnb_search_sa with dirProb");
	orc_set_seed(sa->seed);
	if (sa->unif) orc_set_unif(sa->unif, sa->unif_ctx);      /* the shim's R generator, draw for draw */
	orc_cache_release();                                     /* R releases the cache before the call; rUseCache switches it on */
	orc_set_cache(sa->use_cache);
	orc_result *h = orc_search_sa(N, p->num_samples, p->samples, p->perturbations, p->max_parent_set, p->parent_sizes,
			p->max_complexity, p->num_cats, p->parents_pool, p->fixed_parents, p->edge_prob, sa->select_mode, sa->start_order,
			sa->temp_start, sa->temp_cool_fact, sa->temp_check_orders, sa->max_iter, sa->order_shuffles, sa->stop_diff,
			sa->num_threads);
	orc_set_cache(0);
	orc_cache_release();
	if (!h) { g_err = "oracle: search failed"; return CNB_ERR_PARAM; }
	cnb_result *r = new cnb_result;
	g_live_results++;
	r->h = h;
	r->N = N;
	r->order.resize(N);
	int niter, naccept;
	std::vector<int> ord(N);
	orc_result_sa_info(h, ord.data(), &niter, &naccept);
	std::vector<int32_t> cats = mock_cats(p);
	r->cats_pos.resize(N);
	for (int i = 0; i < N; i++) { r->order[i] = ord[i]; r->cats_pos[i] = cats[ord[i] - 1]; }
	for (int k = 0; k < orc_result_nslots(h); k++)
		if (orc_result_exists(h, k)) r->slots.push_back(k);
	*out = r;
	return CNB_OK;
}
int cnb_search_hist(const cnb_params *p, const cnb_hist_params *hp, double *hist, int32_t *iterations) {
	orc_set_seed(hp->seed);
	if (hp->unif) orc_set_unif(hp->unif, hp->unif_ctx);
	orc_cache_release();
	orc_set_cache(hp->use_cache);
	int it = orc_search_hist(p->num_nodes, p->num_samples, p->samples, p->perturbations, p->max_parent_set, p->parent_sizes,
			p->max_complexity, p->num_cats, p->parents_pool, p->fixed_parents, hp->score, hp->weight, hp->max_iter,
			hp->num_threads, hist);
	orc_set_cache(0);
	orc_cache_release();
	if (it < 0) { g_err = "oracle: histogram failed"; return CNB_ERR_PARAM; }
	if (iterations) *iterations = it;
	return CNB_OK;
}
/* cnb_network + 1-based samples (values < 1 = missing) -> the oracle's zero-based forms */
struct MockNet {
	std::vector<int> s0, pars0, cats, npar;
	std::vector<long long> off;
	int N, maxp;
	MockNet(const cnb_network *net, int M, const int32_t *samples) : N(net->num_nodes), maxp(net->max_parents) {
		cats.assign(net->num_cats, net->num_cats + N);
		npar.assign(net->num_parents, net->num_parents + N);
		pars0.assign((size_t)N * maxp, 0);
		for (int i = 0; i < N; i++)
			for (int k = 0; k < npar[i]; k++) pars0[(size_t)i * maxp + k] = net->parents[(size_t)i * maxp + k] - 1;
		off.assign(net->prob_offsets, net->prob_offsets + N + 1);
		if (samples) {
			s0.resize((size_t)M * N);
			for (size_t j = 0; j < s0.size(); j++) s0[j] = samples[j] < 1 ? 99 : samples[j] - 1;
		}
	}
};
int cnb_loglik(const cnb_network *net, int32_t M, const int32_t *samples, const int32_t *pert, int32_t mode, int32_t, double *out) {
	MockNet m(net, M, samples);
	return orc_net_loglik(m.N, M, m.s0.data(), pert, m.cats.data(), m.maxp, m.npar.data(), m.pars0.data(), net->probs,
			m.off.data(), mode == CNB_LOGLIK_BYSAMPLE ? 1 : 0, out) ? not_in_mock("orc_net_loglik failed") : CNB_OK;
}
int cnb_node_loglik(const cnb_network *net, int32_t nsel, const int32_t *nodes, int32_t M, const int32_t *samples,
		const int32_t *pert, int32_t, double *out) {
	MockNet m(net, M, samples);
	std::vector<double> all(m.N);
	if (orc_net_loglik(m.N, M, m.s0.data(), pert, m.cats.data(), m.maxp, m.npar.data(), m.pars0.data(), net->probs,
			m.off.data(), 0, all.data())) return not_in_mock("orc_net_loglik failed");
	for (int k = 0; k < nsel; k++) {
		if (nodes[k] < 1 || nodes[k] > m.N) { g_err = "invalid node"; return CNB_ERR_PARAM; }
		out[k] = all[nodes[k] - 1];
	}
	return CNB_OK;
}
int cnb_set_prob(const cnb_network *net, int32_t M, const int32_t *samples, const int32_t *pert, int32_t, double *probs_out,
		double *loglik_out, int32_t *sizes_out) {
	MockNet m(net, M, samples);
	return orc_net_set_prob(m.N, M, m.s0.data(), pert, m.cats.data(), m.maxp, m.npar.data(), m.pars0.data(), m.off.data(),
			probs_out, loglik_out, sizes_out) ? not_in_mock("orc_net_set_prob failed") : CNB_OK;
}
int cnb_samples(const cnb_network *net, int32_t M, const int32_t *pert, double na_rate, uint64_t seed, int32_t, int32_t *out) {
	MockNet m(net, 0, NULL);
	long calls = orc_net_samples(m.N, m.cats.data(), m.maxp, m.npar.data(), m.pars0.data(), net->probs, m.off.data(), M, pert,
			na_rate, seed, out);
	if (calls < 0) { g_err = "cyclic network"; return CNB_ERR_PARAM; }
	return CNB_OK;
}

} /* extern "C" */
