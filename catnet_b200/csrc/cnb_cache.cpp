/* cnb_cache.cpp -- the reference's score cache as a host-side model: see cnb_cache.h.  No CUDA calls. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include "cnb_cache.h"
#include "cnb_log.h"

/* PRIMES_1000 (cache.cpp:35-54) is the list of the first 180 primes */
static const uint32_t *first_primes() {
	struct Table {
		uint32_t v[180];
		Table() {
			int n = 0;
			for (uint32_t x = 2; n < 180; x++) {
				bool isp = true;
				for (uint32_t q = 2; q * q <= x && isp; q++) isp = x % q != 0;
				if (isp) v[n++] = x;
			}
		}
	};
	static const Table t;                                  /* initialised once, also when several chains start at the same time */
	return t.v;
}

/* table size and shift, chosen at the first store from the node count and maxParentSet in unsigned 32-bit arithmetic
 * (cache.cpp:225-264) */
void CnbScoreCache::init() {
	const uint32_t *pr = first_primes();
	uint32_t nl = (uint32_t)N;
	for (int i = 0; i < P; i++) nl *= (uint32_t)N;
	if (nl < pr[179]) {
		for (int i = 0; i < 180; i++)
			if (nl >= pr[i]) { nl = pr[i]; break; }
	} else {
		uint32_t prime = pr[179];
		for (int i = 0; i < 180; i++)
			if (pr[i] >= (uint32_t)N) { prime = pr[i]; break; }
		nl = prime;
		for (int i = 0; i < P; i++) nl *= prime;
	}
	if (nl < pr[10]) nl = pr[10];
	nbits = 1;
	int i = 1;
	while (i < N * P) { nbits++; i <<= 1; }
	if (nl < 1) nl = 1;                                   /* InitializeCache (cache.cpp:78-88) */
	ncache = nl;
	if (ncache <= (1u << 24)) flat.assign(ncache, 0);
	static const char *nd = getenv("CNB_CACHE_NO_DENSE");
	dense_on = N <= 160 && !nd;
	if (dense_on) {
		dense1.assign((size_t)N * N, 0);
		dense2.assign((size_t)N * N * N, 0);
	}
}

int32_t *CnbScoreCache::dense_cell(const int32_t *pool, int npool, int node, int npars) {
	if (!dense_on || npool != npars) return nullptr;
	if (npars == 1) return &dense1[(size_t)(node - 1) * N + pool[0] - 1];
	if (npars == 2) return &dense2[((size_t)(node - 1) * N + pool[0] - 1) * N + pool[1] - 1];
	return nullptr;
}

/* cache.cpp:174-184 (= :296-306) */
uint32_t CnbScoreCache::slot_of(const int32_t *pool, int npool, int node, int npars) const {
	const uint32_t *pr = first_primes();
	uint32_t nl = 1;
	for (int i = 0; i < npool; i++) {
		const int j = (pool[i] - 1) % 180;
		nl *= pr[180 - j - 1];
		nl %= ncache;
	}
	nl = (nl << nbits) + (uint32_t)node + (uint32_t)(N * npars);
	return nl % ncache;
}

CnbScoreCache::Entry *CnbScoreCache::lookup(const int32_t *pool, int npool, int node, int npars) {
	st.lookups++;
	if (!ncache) return nullptr;                          /* nothing stored yet (cache.cpp:152-153) */
	if (const int32_t *cell = dense_cell(pool, npool, node, npars)) {
		if (!*cell) return nullptr;
		st.hits++;
		return &entries[*cell - 1];
	}
	const uint32_t s = slot_of(pool, npool, node, npars);
	int32_t id = 0;
	if (!flat.empty()) id = flat[s];
	else {
		auto it = sparse.find(s);
		if (it != sparse.end()) id = it->second;
	}
	if (!id) return nullptr;
	Entry &e = entries[id - 1];
	if (e.node != node || e.npars != npars || e.npool != npool) return nullptr;
	if (memcmp(pool_of(e), pool, (size_t)npool * sizeof(int32_t))) return nullptr;
	st.hits++;
	return &e;
}

/* setCachedProb (cache.cpp:218-317): the new element replaces whatever the slot held */
CnbScoreCache::Entry *CnbScoreCache::store(const int32_t *pool, int npool, int node, int npars) {
	if (!ncache) init();
	st.stores++;
	const uint32_t s = slot_of(pool, npool, node, npars);
	int32_t *ref = nullptr;
	if (!flat.empty()) ref = &flat[s];
	else ref = &sparse[s];
	Entry *e;
	if (*ref) {
		e = &entries[*ref - 1];
		st.evictions++;
		if (int32_t *gone = dense_cell(pool_of(*e), e->npool, e->node, e->npars)) *gone = 0;
		if (e->pool_off >= 0 && npool <= CNB_MAXP) { free_blocks.push_back(e->pool_off); e->pool_off = -1; }
	} else {
		entries.emplace_back();
		*ref = (int32_t)entries.size();
		e = &entries.back();
		e->pool_off = -1;
	}
	if (npool > CNB_MAXP && e->pool_off < 0) {
		if (!free_blocks.empty()) { e->pool_off = free_blocks.back(); free_blocks.pop_back(); }
		else { e->pool_off = (int32_t)arena.size(); arena.resize(arena.size() + N); }
	}
	e->node = node;
	e->npars = npars;
	e->npool = npool;
	memcpy(e->pool_off < 0 ? e->pool_small : &arena[e->pool_off], pool, (size_t)npool * sizeof(int32_t));
	e->prior = 0;
	if (int32_t *cell = dense_cell(pool, npool, node, npars)) *cell = *ref;
	return e;
}

double cnb_host_prior(const CnbPlan &pl, int child, const int32_t *pars, int d) {
	if (!pl.has_prior) return 0.0;
	const int N = pl.N;
	const double *edge = pl.edge.data();
	double t = 1.0;
	for (int i = 0; i < d; i++) t = t * edge[(size_t)pars[i] * N + child];
	double prior = t > 0 ? cnb_log(t) : CNB_NEG_INF;
	double t2 = 1.0;
	for (int i = 0; i < N; i++) {
		if (i == child) continue;
		bool isp = false;
		for (int j = 0; j < d; j++) isp = isp || pars[j] == i;
		if (!isp) t2 = t2 * (1.0 - edge[(size_t)i * N + child]);
	}
	t2 = t2 > 0 ? cnb_log(t2) : CNB_NEG_INF;
	if (prior == CNB_NEG_INF || t2 == CNB_NEG_INF) return CNB_NEG_INF;
	return prior + t2;
}

/* rUseCache as the reference reads it: off beyond 8 threads (rcatnet_sa.cpp:315-319, rcatnet_hist.cpp:205-209).
 * CNB_SCORE_CACHE=off ignores the flag (the cache-off result, proposals of a step evaluated side by side). */
bool cnb_cache_wanted(int use_cache, int num_threads) {
	static const char *env = getenv("CNB_SCORE_CACHE");
	if (env && (!strcmp(env, "off") || !strcmp(env, "0"))) return false;
	return use_cache && num_threads <= 8;
}

static bool next_comb(int *idx, int k, int n) {
	int i = k - 1;
	while (i >= 0 && idx[i] == n - k + i) i--;
	if (i < 0) return false;
	idx[i]++;
	for (int j = i + 1; j < k; j++) idx[j] = idx[j - 1] + 1;
	return true;
}

int CnbScoreCache::walk(const CnbPlan &pl, const cnb_params *params, const uint8_t *never_kept, const int64_t *winners,
		std::vector<CnbOverride> *out) {
	if (pl.N != N) { cnb_set_error("score cache: %d nodes, plan has %d", N, pl.N); return CNB_ERR_PARAM; }
	std::vector<int32_t> pos_of(N), pool(N + CNB_MAXP), listing(CNB_MAXP);
	for (int i = 0; i < N; i++) pos_of[pl.order[i]] = i;
	const size_t first = out->size();
	int idx[CNB_MAXP];
	auto complexity = [&](int child, const int32_t *pars, int d) {
		int64_t cx = pl.cats[child] - 1;
		for (int k = 0; k < d; k++) cx *= pl.cats[pars[k]];
		return (int32_t)cx;
	};
	for (int c = 1; c < N; c++) {                          /* catnet_search2.h:444 */
		const CnbNodePlan &nd = pl.nodes[c];
		const int node = pl.order[c] + 1;
		const int32_t *fix = pl.lists.data() + nd.fix_off, *fre = pl.lists.data() + nd.free_off;
		const int nf = nd.nfix, nr = nd.nfree;
		const bool dead = never_kept && never_kept[pl.order[c]];
		if (nd.equal_branch) {
			/* ONE entry per (node, pool, parent count): the winner (:509-524, :642-650).  A node without a sample to score
			 * with never stores (ncombMaxLogLik < 0, :637-640), and its lookups change nothing */
			if (dead) continue;
			for (int bi = nd.first_blk; bi < nd.first_blk + nd.nblk; bi++) {
				const CnbBlock &b = pl.blocks[bi];
				for (int j = 0; j < nr; j++) pool[j] = pl.order[fre[j]] + 1;
				for (int j = 0; j < nf; j++) pool[nr + j] = pl.order[fix[j]] + 1;
				std::sort(pool.begin(), pool.begin() + nr + nf);
				Entry *e = lookup(pool.data(), nr + nf, node, b.d);
				if (e) {
					CnbOverride ov;
					memset(&ov, 0, sizeof(ov));
					ov.opt = b.opt_off;
					ov.child = c;
					ov.d = b.d;
					ov.prior = e->prior;
					/* the cached set as a candidate of this block: its free parents, ascending, give the rank */
					int k = 0;
					bool ok = true;
					for (int q = 0; q < b.d; q++) {
						const int par = pos_of[e->pars[q] - 1];
						ov.pars[q] = par;
						bool fixed = false;
						for (int t = 0; t < nf; t++) fixed = fixed || fix[t] == par;
						if (fixed) continue;
						const int32_t *at = std::lower_bound(fre, fre + nr, par);
						if (at == fre + nr || *at != par || k >= b.d - nf) { ok = false; break; }
						idx[k++] = (int)(at - fre);
					}
					if (!ok || k != b.d - nf) { cnb_set_error("score cache: a cached parent set is not a candidate of its pool"); return CNB_ERR_PROC; }
					std::sort(idx, idx + k);
					int64_t rank = 0;
					int prev = -1;
					for (int i = 0; i < k; i++) {
						for (int j = prev + 1; j < idx[i]; j++) rank += cnb_binom(nr - 1 - j, k - 1 - i);
						prev = idx[i];
					}
					ov.fid = b.start + rank;
					ov.cx = complexity(c, ov.pars, b.d);
					out->push_back(ov);
					continue;
				}
				const int64_t w = winners ? winners[b.opt_off] : -1;
				if (w < 0) continue;                          /* no candidate with a finite score: nothing is stored (:637-640) */
				cnb_unrank_lex(nr, b.d - nf, w - b.start, idx);
				for (int j = 0; j < nf; j++) listing[j] = fix[j];
				for (int j = 0; j < b.d - nf; j++) listing[nf + j] = fre[idx[j]];
				e = store(pool.data(), nr + nf, node, b.d);
				for (int j = 0; j < b.d; j++) e->pars[j] = pl.order[listing[j]] + 1;
				e->prior = cnb_host_prior(pl, c, listing.data(), b.d);
			}
			continue;
		}
		/* unequal categories: one entry per candidate, its own parents as the pool (:721-733, :789-797), every candidate
		 * in enumeration order.  A node without samples has no blocks in the plan, but the reference still stores its
		 * candidates (with -FLT_MAX), and a store can evict what a later order would have hit */
		int maxpars = params->max_parent_set;
		if (params->parent_sizes && params->parent_sizes[pl.order[c]] < maxpars) maxpars = params->parent_sizes[pl.order[c]];
		if (maxpars > nr + nf) maxpars = nr + nf;
		if (maxpars > CNB_MAXP) maxpars = CNB_MAXP;
		int bi = nd.first_blk;
		for (int d = nf + 1; d <= maxpars; d++) {
			const int kk = d - nf;
			const CnbBlock *b = nullptr;
			if (!dead) {
				if (bi >= nd.first_blk + nd.nblk || pl.blocks[bi].d != d) { cnb_set_error("score cache: plan and walk disagree on node %d", c + 1); return CNB_ERR_PROC; }
				b = &pl.blocks[bi++];
			}
			if (nf == 0 && d <= 2 && dense_on && !emit_all_hits && !dead) {
				/* the bulk of an SA chain's lookups: one or two free parents, keys that the dense index answers -- plain
				 * loops in the enumeration order (i, then j > i = lexicographic), no hash unless the key has to be stored */
				const size_t nbase = (size_t)(node - 1) * N;
				const int ccx = pl.cats[c] - 1;
				const int64_t opt0 = b->opt_off;
				int64_t r = 0;
				for (int i = 0; i < (d == 1 ? nr : nr - 1); i++) {
					const int a = fre[i], la = pl.order[a] + 1;
					for (int j = (d == 1 ? i : i + 1); j < (d == 1 ? i + 1 : nr); j++, r++) {
						const int b = fre[j], lb = pl.order[b] + 1;
						const int32_t cell = d == 1 ? dense1[nbase + la - 1]
								: (la < lb ? dense2[(nbase + la - 1) * N + lb - 1] : dense2[(nbase + lb - 1) * N + la - 1]);
						st.lookups++;
						if (cell) {
							st.hits++;
							const Entry &h = entries[cell - 1];
							const bool fwd = h.pars[0] == la;               /* d == 2: listed (a, b) as this order would, or (b, a) */
							if (!pl.has_prior && fwd) continue;
							out->emplace_back();
							CnbOverride &ov = out->back();
							ov.opt = opt0 + r;
							ov.fid = -1;
							ov.child = c;
							ov.d = d;
							ov.prior = h.prior;
							ov.pars[0] = fwd ? a : b;
							ov.pars[1] = d == 2 ? (fwd ? b : a) : 0;
							for (int q = 2; q < CNB_MAXP; q++) ov.pars[q] = 0;
							ov.cx = d == 1 ? ccx * pl.cats[a] : ccx * pl.cats[a] * pl.cats[b];
							continue;
						}
						st.lookups--;                                      /* lookup() below counts it */
						pool[0] = d == 1 ? la : std::min(la, lb);
						pool[1] = std::max(la, lb);
						Entry *e = lookup(pool.data(), d, node, d);        /* a miss (or nothing stored at all yet) */
						if (e) { cnb_set_error("score cache: dense index out of step"); return CNB_ERR_PROC; }
						e = store(pool.data(), d, node, d);
						e->pars[0] = la;
						e->pars[1] = d == 2 ? lb : 0;
						listing[0] = a;
						listing[1] = b;
						e->prior = cnb_host_prior(pl, c, listing.data(), d);
					}
				}
				continue;
			}
			for (int j = 0; j < kk; j++) idx[j] = j;
			int64_t r = 0;
			do {
				for (int j = 0; j < nf; j++) listing[j] = fix[j];
				for (int j = 0; j < kk; j++) listing[nf + j] = fre[idx[j]];
				for (int j = 0; j < d; j++) pool[j] = pl.order[listing[j]] + 1;
				if (d == 2) { if (pool[0] > pool[1]) std::swap(pool[0], pool[1]); }
				else if (d > 2) std::sort(pool.begin(), pool.begin() + d);
				if (dense_on && d <= 2 && !dead && !emit_all_hits && !pl.has_prior) {
					/* the common case once the chain has seen its families: the key is in the table with the listing this order
					 * would give it -- nothing to do, no hash */
					const int32_t cell = d == 1 ? dense1[(size_t)(node - 1) * N + pool[0] - 1]
							: dense2[((size_t)(node - 1) * N + pool[0] - 1) * N + pool[1] - 1];
					if (cell) {
						const Entry &h = entries[cell - 1];
						if (h.pars[0] == pl.order[listing[0]] + 1 && (d == 1 || h.pars[1] == pl.order[listing[1]] + 1)) {
							st.lookups++;
							st.hits++;
							r++;
							continue;
						}
					}
				}
				Entry *e = lookup(pool.data(), d, node, d);
				if (!e) {
					e = store(pool.data(), d, node, d);
					for (int j = 0; j < d; j++) e->pars[j] = pl.order[listing[j]] + 1;
					if (!dead) e->prior = cnb_host_prior(pl, c, listing.data(), d);
				} else if (!dead) {
					bool same = !pl.has_prior && !emit_all_hits;   /* with priors the stored prior belongs to the storing order: always an override */
					for (int j = 0; j < d && same; j++) same = e->pars[j] == pl.order[listing[j]] + 1;
					if (!same) {
						CnbOverride ov;
						memset(&ov, 0, sizeof(ov));
						ov.opt = b->opt_off + r;
						ov.fid = -1;
						ov.child = c;
						ov.d = d;
						ov.prior = e->prior;
						for (int j = 0; j < d; j++) ov.pars[j] = pos_of[e->pars[j] - 1];
						ov.cx = complexity(c, ov.pars, d);
						out->push_back(ov);
					}
				}
				r++;
			} while (next_comb(idx, kk, nr));
		}
	}
	st.overrides += (int64_t)(out->size() - first);
	return CNB_OK;
}

/* test hook (CPU): the model on a sequence of node orders.  p: the search (num_nodes, max_parent_set, max_complexity,
 * parent_sizes, pools, edge_prob; order ignored); cats [N] label space; never_kept [N] or null; orders [norders][N]
 * 1-based labels; winners [norders][N][P][P]: the listed labels (0-padded, all 0 = none) of the winner of (order, child
 * label - 1, parent count - 1) on the equal-categories path, as the device would report it.  Every HIT (all_hits) or
 * every hit that changes what the device computed (an override: what the search itself acts on) comes back as a
 * row of 4 + CNB_MAXP ints -- 1 equal-categories path / 2 otherwise, order index, child label, parent count, listed
 * parent labels -- with the entry's prior beside it.  Returns the number of hits (rows beyond cap are dropped), < 0 on
 * an error. */
extern "C" int64_t cnb_cache_trace(const cnb_params *p, const int32_t *cats, const uint8_t *never_kept, int32_t norders,
		const int32_t *orders, const int32_t *winners, int32_t all_hits, int32_t cap, int32_t *rows, double *priors) {
	if (!p || !cats || !orders || norders < 0) return -1;
	const int N = p->num_nodes, P = p->max_parent_set;
	CnbScoreCache cache(N, P);
	std::vector<int32_t> eff;
	int64_t nrows = 0;
	for (int o = 0; o < norders; o++) {
		cnb_params q = *p;
		q.order = orders + (size_t)o * N;
		if (q.num_samples < 1) q.num_samples = 1;
		if (never_kept) {                                      /* as cnb_ctx_plan: no candidates for nodes without samples */
			eff.resize(N);
			for (int i = 0; i < N; i++) eff[i] = never_kept[i] ? 0 : (p->parent_sizes ? p->parent_sizes[i] : P);
			q.parent_sizes = eff.data();
		}
		CnbPlan pl;
		if (cnb_build_plan(&q, cats, 0, 1, &pl, nullptr) != CNB_OK) return -2;
		std::vector<int64_t> win((size_t)std::max<int64_t>(pl.num_options, 1), -1);
		std::vector<int32_t> pos_of(N);
		for (int i = 0; i < N; i++) pos_of[pl.order[i]] = i;
		for (const CnbBlock &b : pl.blocks) {
			if (!pl.nodes[b.child].equal_branch || !winners) continue;
			const int32_t *w = winners + (((size_t)o * N + pl.order[b.child]) * P + (b.d - 1)) * P;
			if (w[0] <= 0) continue;
			int idx[CNB_MAXP], k = 0;
			const int32_t *fre = pl.lists.data() + b.free_off, *fix = pl.lists.data() + b.fix_off;
			for (int t = 0; t < b.d; t++) {
				const int par = pos_of[w[t] - 1];
				bool fixed = false;
				for (int u = 0; u < b.nfix; u++) fixed = fixed || fix[u] == par;
				if (fixed) continue;
				const int32_t *at = std::lower_bound(fre, fre + b.nfree, par);
				if (at == fre + b.nfree || *at != par || k >= CNB_MAXP) return -3;
				idx[k++] = (int)(at - fre);
			}
			std::sort(idx, idx + k);
			int64_t rank = 0;
			int prev = -1;
			for (int i = 0; i < k; i++) {
				for (int j = prev + 1; j < idx[i]; j++) rank += cnb_binom(b.nfree - 1 - j, k - 1 - i);
				prev = idx[i];
			}
			win[b.opt_off] = b.start + rank;
		}
		std::vector<CnbOverride> ov;
		cache.emit_all_hits = all_hits != 0;
		if (cache.walk(pl, p, never_kept, win.data(), &ov) != CNB_OK) return -4;
		for (const CnbOverride &v : ov) {
			if (nrows < cap && rows) {
				int32_t *r = rows + (size_t)nrows * (4 + CNB_MAXP);
				memset(r, 0, (4 + CNB_MAXP) * sizeof(int32_t));
				r[0] = v.fid >= 0 ? 1 : 2;
				r[1] = o;
				r[2] = pl.order[v.child] + 1;
				r[3] = v.d;
				for (int t = 0; t < v.d; t++) r[4 + t] = pl.order[v.pars[t]] + 1;
				if (priors) priors[nrows] = v.prior;
			}
			nrows++;
		}
	}
	return nrows;
}
