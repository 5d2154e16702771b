/* cnb_cache.h -- host-side model of the reference's score cache (/root/reference/src/cache.cpp, cache.h).
 *
 * R runs cnSearchSA and cnSearchHist with the cache ON (R/catnet.search.R:293-294, R/catnet.histo.R) after releasing
 * it, and the cache is visible in the results: a hit returns the parent set in the sequence it was LISTED in when it was
 * first scored (under another node order), the table laid out in that sequence, the log-likelihood summed in that table
 * order and the prior multiplied up in the node order of that other order (cache.cpp:142-216).  So cache-ON networks list
 * parents differently, carry log-likelihoods that differ in the last bits, and near-ties can resolve to other parent
 * sets than with the cache off.
 *
 * The device re-scores every order anyway; what has to be reproduced is WHICH listing the reference would have used for
 * every candidate, and that is a pure function of the sequence of evaluated orders (plus, on the equal-categories path,
 * of the winners the device reports): the table is a direct-mapped array of g_ncache slots addressed by a product of
 * primes over the sorted pool (cache.cpp:174-184), a store replaces whatever is in its slot (:307-311), nothing else is
 * ever evicted.  CnbScoreCache keeps that table WITHOUT the scores -- key, listed sequence, prior -- and walks a planned
 * order in the reference's sequence of lookups and stores (catnet_search2.h:509-524, 642-650 equal categories: one
 * entry per (node, pool, parent count) holding the winner; :721-733, 789-797 otherwise: one entry per candidate).
 * Every hit whose listing or prior differs from what the device computed for the current order becomes an override: the
 * family is re-scored with its parents in the cached sequence and patched into the option table before the DP
 * (cnb_api.cu: apply_cache), and the exported network lists them that way.  No CUDA in here (CPU test-suite).
 */
#ifndef CNB_CACHE_H
#define CNB_CACHE_H
#include <stdint.h>
#include <unordered_map>
#include <vector>
#include "cnb_internal.h"

/* one candidate whose score / listing the cache changes, in the order space of the plan that was walked */
struct CnbOverride {
	int64_t opt;               /* option slot (CnbBlock.opt_off [+ rank on the unequal-categories path]) */
	int64_t fid;               /* family id inside its group (the equal-categories path replaces the block's winner by it) */
	int32_t child, d;
	int32_t cx;                /* complexity of the family */
	int32_t pars[CNB_MAXP];    /* order positions, in the cached listing sequence */
	double prior;              /* the prior stored with the entry (0 without edge priors) */
};

struct CnbCacheStats {
	int64_t lookups = 0, hits = 0, stores = 0, evictions = 0, overrides = 0;
};

class CnbScoreCache {
public:
	CnbScoreCache(int32_t num_nodes, int32_t max_parent_set) : N(num_nodes), P(max_parent_set) {}
	/* Walk the planned order pl in the reference's sequence.  params: the search's parameters (parent_sizes for nodes the
	 * plan gives no candidates); never_kept [N] label space or null (nodes perturbed in every sample: the unequal path
	 * still stores their candidates); winners [pl.num_options] = the device's option table family ids (only the slots of
	 * equal-categories blocks are read: the winner of the block, -1 = no candidate with a finite score).  Appends the
	 * overrides of this order (ascending option slot per node; sorted by opt on return). */
	int walk(const CnbPlan &pl, const cnb_params *params, const uint8_t *never_kept, const int64_t *winners,
			std::vector<CnbOverride> *out);
	const CnbCacheStats &stats() const { return st; }
	uint32_t slots() const { return ncache; }
	uint32_t bits() const { return nbits; }
	bool emit_all_hits = false;          /* test hook: every hit becomes an override, also those that change nothing */
private:
	struct Entry {
		int32_t node, npars, npool;      /* labels 1-based */
		int32_t pool_off;                /* sorted pool labels: a block of N ints in arena, or -1 = pool_small (up to CNB_MAXP labels) */
		int32_t pool_small[CNB_MAXP];
		int32_t pars[CNB_MAXP];          /* labels 1-based, listing sequence */
		double prior;
	};
	const int32_t *pool_of(const Entry &e) const { return e.pool_off < 0 ? e.pool_small : &arena[e.pool_off]; }
	int32_t N, P;
	uint32_t ncache = 0, nbits = 0;
	std::vector<int32_t> flat;           /* slot -> entry index + 1 (tables up to 2^24 slots) */
	std::unordered_map<uint32_t, int32_t> sparse;   /* larger tables: the reference allocates them in full, we do not */
	std::vector<Entry> entries;
	std::vector<int32_t> arena, free_blocks;   /* pools beyond CNB_MAXP labels: blocks of N ints, recycled when their entry is evicted */
	/* entries whose pool IS their parent set (every entry of the unequal-categories path) with one or two parents, by key:
	 * dense1[(node-1)*N + a-1], dense2[((node-1)*N + a-1)*N + b-1] (a < b) = entry index + 1, 0 = not in the table.  Kept in
	 * step with the slots (a store enters its key and removes the key it evicts), so a lookup of such a key needs no hash:
	 * it hits exactly when the slot of that key still holds it.  Up to 160 nodes (16 MB). */
	bool dense_on = false;
	std::vector<int32_t> dense1, dense2;
	int32_t *dense_cell(const int32_t *pool_sorted, int npool, int node, int npars);
	CnbCacheStats st;
	void init();
	uint32_t slot_of(const int32_t *pool_sorted, int npool, int node, int npars) const;
	Entry *lookup(const int32_t *pool_sorted, int npool, int node, int npars);
	Entry *store(const int32_t *pool_sorted, int npool, int node, int npars);
};

/* the Bernoulli edge prior of (child | pars) under the plan's order, on the host, operation for operation as dev_prior
 * (cnb_device.cuh) and catnet_search2.h:584-611 */
bool cnb_cache_wanted(int use_cache, int num_threads);
double cnb_host_prior(const CnbPlan &pl, int child, const int32_t *pars, int d);

#endif
