/* oracle/catnet_oracle.h -- TEST INFRASTRUCTURE: plain-C restatement of catnet's structure-search scoring path.
 * See catnet_oracle.c.  Label-space conventions are those of the .Call arguments (and of include/catnet_b200.h):
 * samples [M][N] 1-based categories (NA = INT_MIN or < 1), order = 1-based labels per position, pools = [N][N]
 * rows of 1-based labels 0-padded (empty row = no parents allowed), edge = R's column-major edgeProb[child,parent]. */
#ifndef CATNET_ORACLE_H
#define CATNET_ORACLE_H
#define ORC_MAXP 16
typedef struct orc_result orc_result;

int orc_family_score(const int *samples0, int N, int M, const int *cats, int child, const int *pars, int d,
		const unsigned char *keep, double *counts, double *loglik_out, int *ncount_out);
orc_result *orc_search_order(int N, int M, const int *samples_lbl, const int *pert_lbl, int maxParentSet,
		const int *parentSizes_lbl, int maxComplexity, const int *order, const int *numCats_lbl,
		const int *parentsPool_lbl, const int *fixedPool_lbl, const double *edge_lbl);
orc_result *orc_search_sa(int N, int M, const int *samples_lbl, const int *pert_lbl, int maxParentSet,
		const int *parentSizes_lbl, int maxComplexity, const int *numCats_lbl, const int *parentsPool_lbl,
		const int *fixedPool_lbl, const double *edge_lbl, int selectMode, const int *startOrder, double tempStart,
		double tempCoolFact, int tempCheckOrders, int maxIter, double orderShuffles, double stopDiff, int numThreads);
int orc_search_hist(int N, int M, const int *samples_lbl, const int *pert_lbl, int maxParentSet,
		const int *parentSizes_lbl, int maxComplexity, const int *numCats_lbl, const int *parentsPool_lbl,
		const int *fixedPool_lbl, int scoreSel, int weight, int maxIter, int numThreads, double *hist);
/* catnet_entropy.cpp pairwise metrics: kind 0 entropy, 1 KL, 2 Pearson -> out[n2*N + n1]; 3 entropy order -> order_out */
int orc_pairwise(int kind, int N, int M, const int *samples_lbl, const int *pert_lbl, double *out, int *order_out);
/* fixed network (catnet_class.h:551-879): mode 0 sampleLoglikVector, 1 bySampleLoglikVector, 2 sampleNodeLoglik of
 * every node, 3 sampleLoglik; samples0 zero-based, parents0 [N][maxParents] 0-based in listed order */
int orc_net_loglik(int N, int M, const int *samples0, const int *pert, const int *cats, int maxParents,
		const int *numParents, const int *parents0, const double *probs, const long long *offsets, int mode,
		double *out);
int orc_net_set_prob(int N, int M, const int *samples0, const int *pert, const int *cats, int maxParents,
		const int *numParents, const int *parents0, const long long *offsets, double *probs_out, double *loglik_out,
		int *sizes_out);
/* cnSamples (rcatnet.cpp:524-650) under the splitmix64 stream seeded by `seed`; out [M][N]; returns uniforms consumed */
long orc_net_samples(int N, const int *cats, int maxParents, const int *numParents, const int *parents0,
		const double *probs, const long long *offsets, int M, const int *pert, double naRate, unsigned long long seed,
		int *out);
/* the reference's score cache (cache.cpp): orc_set_cache(1) makes orc_search_sa / orc_search_hist (and orc_search_order)
 * look candidates up and store them as CATNET_SEARCH2::estimate does; orc_cache_release = ReleaseCache */
void orc_set_cache(int on);
void orc_cache_release(void);
/* test hook: record what the cache does (rows of 4 + ORC_MAXP ints: kind 1 equal-categories hit / 2 other hit / 3 store of an
 * equal-categories winner, search serial, child label, parent count, listed parent labels; the hit's prior; every
 * search's node order [serial][N]) -- checks the library's host-side model of the cache on the CPU */
void orc_cache_trace(int *rows, double *priors, int cap_rows, int *orders, int cap_orders);
void orc_cache_trace_counts(int *nrows, int *nsearches);
void orc_set_seed(unsigned long long seed);
void orc_set_unif(double (*fn)(void *), void *ctx);   /* caller-owned generator; cleared by orc_set_seed */
int orc_result_nslots(const orc_result *r);
int orc_result_exists(const orc_result *r, int k);
int orc_result_net(orc_result *r, int k, int *complexity, double *loglik);
int orc_result_node(const orc_result *r, int k, int node, int *npars, int *pars, double *loglik, double *prior,
		int *sampleSize, int *probsize, double *probs, int cap);
int orc_result_counts(const orc_result *r, int k, int node, double *counts, int cap);
int orc_result_sa_info(const orc_result *r, int *order_out, int *niter, int *naccept);
int orc_result_sa_trace(const orc_result *r, double *trace, int *accept, int cap);
void orc_result_free(orc_result *r);
#endif
