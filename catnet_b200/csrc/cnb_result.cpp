/* cnb_result.cpp -- accessors over the returned networks: the catNetwork slots RCatnet::genRcatnet exports
 * (/root/reference/src/rcatnet.cpp:174-316: parents :214-229, probabilities :269-280 + genProbList :318-345,
 * complexity/likelihood :290-303, nodeSampleSizes :305-312). */
#include <string.h>
#include "cnb_result.h"

extern "C" {

int32_t cnb_result_num_networks(const cnb_result *r) { return r ? (int32_t)r->nets.size() : 0; }
int32_t cnb_result_num_nodes(const cnb_result *r) { return r ? r->N : 0; }

int cnb_result_network(const cnb_result *r, int32_t i, int32_t *complexity, double *loglik) {
	if (!r || i < 0 || i >= (int32_t)r->nets.size()) return CNB_ERR_PARAM;
	if (complexity) *complexity = r->nets[i].complexity;
	if (loglik) *loglik = r->nets[i].loglik;
	return CNB_OK;
}

int cnb_result_order(const cnb_result *r, int32_t i, int32_t *order_out) {
	if (!r || i < 0 || i >= (int32_t)r->nets.size() || !order_out) return CNB_ERR_PARAM;
	const std::vector<int32_t> &o = r->orders[r->nets[i].order_id];
	memcpy(order_out, o.data(), o.size() * sizeof(int32_t));
	return CNB_OK;
}

int cnb_result_cats(const cnb_result *r, int32_t i, int32_t *cats_out) {
	if (!r || i < 0 || i >= (int32_t)r->nets.size() || !cats_out) return CNB_ERR_PARAM;
	const std::vector<int32_t> &c = r->order_cats[r->nets[i].order_id];
	memcpy(cats_out, c.data(), c.size() * sizeof(int32_t));
	return CNB_OK;
}

int cnb_result_parents(const cnb_result *r, int32_t i, int32_t *num_parents, int32_t *parents, int32_t cap) {
	if (!r || i < 0 || i >= (int32_t)r->nets.size() || !num_parents) return CNB_ERR_PARAM;
	const CnbNet &net = r->nets[i];
	for (int n = 0; n < r->N; n++) {
		const CnbFamily &f = r->fams[net.fam[n]];
		num_parents[n] = f.d;
		if (parents)
			for (int k = 0; k < f.d && k < cap; k++) parents[(size_t)n * cap + k] = f.pars[k];
	}
	return CNB_OK;
}

/* CPT = counts normalised per parent configuration; an unobserved configuration becomes uniform with the last
 * cell 1 - sum (problist.h:193-222) */
int cnb_result_node(const cnb_result *r, int32_t i, int32_t node, double *loglik, double *prior, int32_t *sample_size,
		double *probs, int32_t probs_cap) {
	if (!r || i < 0 || i >= (int32_t)r->nets.size() || node < 0 || node >= r->N) return CNB_ERR_PARAM;
	const CnbFamily &f = r->fams[r->nets[i].fam[node]];
	if (loglik) *loglik = f.loglik;
	if (prior) *prior = f.prior;
	if (sample_size) *sample_size = f.ncount;
	int ts = (int)f.counts.size();
	if (probs) {
		for (int k = 0; k < ts; k += f.cc) {
			double pp = 0;
			for (int c = 0; c < f.cc; c++) pp += (double)f.counts[k + c];
			if (pp > 0) {
				pp = 1 / pp;
				for (int c = 0; c < f.cc && k + c < probs_cap; c++) probs[k + c] = (double)f.counts[k + c] * pp;
			} else {
				double favg = (double)1 / (double)f.cc, acc = 0;
				for (int c = 0; c < f.cc - 1; c++) {
					if (k + c < probs_cap) probs[k + c] = favg;
					acc += favg;
				}
				if (k + f.cc - 1 < probs_cap) probs[k + f.cc - 1] = 1 - acc;
			}
		}
	}
	return ts;
}

int cnb_result_counts(const cnb_result *r, int32_t i, int32_t node, int32_t *counts, int32_t cap) {
	if (!r || i < 0 || i >= (int32_t)r->nets.size() || node < 0 || node >= r->N) return CNB_ERR_PARAM;
	const CnbFamily &f = r->fams[r->nets[i].fam[node]];
	for (int k = 0; k < (int)f.counts.size() && k < cap; k++) counts[k] = f.counts[k];
	return (int)f.counts.size();
}

int cnb_result_sa_info(const cnb_result *r, int32_t *order_out, int32_t *niter, int32_t *naccept) {
	if (!r) return CNB_ERR_PARAM;
	if (order_out) for (size_t i = 0; i < r->sa_order.size(); i++) order_out[i] = r->sa_order[i];
	if (niter) *niter = r->niter;
	if (naccept) *naccept = r->naccept;
	return CNB_OK;
}

int cnb_result_sa_trace(const cnb_result *r, double *loglik, int32_t *accepted, int32_t cap) {
	if (!r) return CNB_ERR_PARAM;
	int n = (int)r->trace.size();
	for (int i = 0; i < n && i < cap; i++) {
		if (loglik) loglik[i] = r->trace[i];
		if (accepted) accepted[i] = r->trace_accept[i];
	}
	return n;
}

/* per complexity keep the network with the higher likelihood (R/catnet.search.R:317-332): a's list is the
 * "optnets" of the newer search, b the earlier evaluation; a slot of a is replaced when b's is strictly better. */
int cnb_result_merge(cnb_result *a, const cnb_result *b) {
	if (!a || !b || a->N != b->N) return CNB_ERR_PARAM;
	int order_base = (int)a->orders.size();
	int fam_base = (int)a->fams.size();
	bool copied = false;
	for (const CnbNet &nb : b->nets) {
		for (CnbNet &na : a->nets) {
			if (na.complexity != nb.complexity || !(nb.loglik > na.loglik)) continue;
			if (!copied) {
				a->orders.insert(a->orders.end(), b->orders.begin(), b->orders.end());
				a->order_cats.insert(a->order_cats.end(), b->order_cats.begin(), b->order_cats.end());
				a->fams.insert(a->fams.end(), b->fams.begin(), b->fams.end());
				copied = true;
			}
			na = nb;
			na.order_id += order_base;
			for (int32_t &f : na.fam) f += fam_base;
		}
	}
	return CNB_OK;
}

void cnb_result_free(cnb_result *r) { delete r; }

} /* extern "C" */

/* chains of one multi-device cnSearchSA: per complexity the higher likelihood (a's on ties), and complexities only b
 * reached are added -- each chain's list is a full evaluation of its own accepted order */
int cnb_result_union(cnb_result *a, const cnb_result *b) {
	if (!a || !b || a->N != b->N) return CNB_ERR_PARAM;
	const int order_base = (int)a->orders.size(), fam_base = (int)a->fams.size();
	bool copied = false;
	std::vector<CnbNet> out;
	size_t i = 0, j = 0;
	auto take_b = [&](const CnbNet &nb) {
		if (!copied) {
			a->orders.insert(a->orders.end(), b->orders.begin(), b->orders.end());
			a->order_cats.insert(a->order_cats.end(), b->order_cats.begin(), b->order_cats.end());
			a->fams.insert(a->fams.end(), b->fams.begin(), b->fams.end());
			copied = true;
		}
		CnbNet n = nb;
		n.order_id += order_base;
		for (int32_t &f : n.fam) f += fam_base;
		out.push_back(std::move(n));
	};
	while (i < a->nets.size() || j < b->nets.size()) {
		if (j >= b->nets.size() || (i < a->nets.size() && a->nets[i].complexity < b->nets[j].complexity)) out.push_back(a->nets[i++]);
		else if (i >= a->nets.size() || b->nets[j].complexity < a->nets[i].complexity) take_b(b->nets[j++]);
		else {
			if (b->nets[j].loglik > a->nets[i].loglik) take_b(b->nets[j]);
			else out.push_back(a->nets[i]);
			i++;
			j++;
		}
	}
	a->nets.swap(out);
	return CNB_OK;
}
