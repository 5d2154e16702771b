/*
 * catnet_b200.h -- C ABI of the B200-native structure-search scoring path of catnet.
 *
 * This is the drop-in boundary: exactly what the reference's .Call glue for cnSearchOrder / cnSearchSA binds.
 * Plain pointers and sizes only (no torch, no SEXP); all arrays are HOST memory unless a name ends in _dev.
 *
 * What each entry point replaces in the reference (/root/reference):
 *   cnb_search_order      catnetOptimalNetsForOrder (src/catnet_rexport.h:36-40, registered src/ccnInit.c:30)
 *                         -> RCatnetSearch::estimateCatnets (src/rcatnet_search.cpp:57-312)
 *                         -> CATNET_SEARCH2::estimate       (src/catnet_search2.h:152-897)
 *   cnb_search_sa         catnetOptimalNetsSA (src/catnet_rexport.h:42-48, src/ccnInit.c:31)
 *                         -> RCatnetSearchSA::search        (src/rcatnet_sa.cpp:253-935)
 *   cnb_search_hist       catnetParHistogram (src/catnet_rexport.h:50-54, src/ccnInit.c:32)
 *                         -> RCatnetSearchHist::search      (src/rcatnet_hist.cpp:164-573)
 *   cnb_release_cache     catnetReleaseCache (src/catnet_rexport.h:62, src/cache.cpp:62-76)
 *   cnb_set_seed          catnetSetSeed      (src/catnet_rexport.h:64, src/ccnInit.c:26)
 *   cnb_params            SEARCH_PARAMETERS  (src/search_params.h:39-58) as filled from the .Call arguments
 *   cnb_result_*          RCatnet::genRcatnet / genProbList (src/rcatnet.cpp:174-345): the S4 catNetwork slots
 *   cnb_family_scores     CATNET::setParents + setNodeSampleProb (src/catnet_class.h:393-433, 818-879) -- test hook
 *
 * Conventions kept from the reference:
 *   - samples arrive as R hands them to .Call: INTSXP matrix numnodes x numsamples, column-major, i.e. element
 *     [sample j][node i] at j*numnodes + i; values are 1-based category indices; NA_INTEGER (INT_MIN) or any
 *     value < 1 is missing (rcatnet_search.cpp:151-160).
 *   - node order, pools and fixed parents are 1-based node labels, as in R.
 *   - return codes are the reference's CATNET_ERR_* (src/utils.h:57-60) plus CNB_ERR_CUDA; nothing longjmps
 *     across this boundary (the R shim turns codes into Rf_error on the calling thread).
 *   - "-infinity" log-likelihood is -FLT_MAX, as in the reference.
 *
 * There is no CPU fallback: every scoring call needs the sm_100a kernels and fails with CNB_ERR_CUDA without a GPU.
 */
#ifndef CATNET_B200_H
#define CATNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CNB_OK          0
#define CNB_ERR_MEM   (-10)   /* CATNET_ERR_MEM   */
#define CNB_ERR_PARAM (-20)   /* CATNET_ERR_PARAM */
#define CNB_ERR_PROC  (-30)   /* CATNET_ERR_PROC  */
#define CNB_ERR_CUDA  (-40)   /* CUDA runtime / no device / kernels missing */

#define CNB_NA_INTEGER INT32_MIN

typedef struct cnb_ctx cnb_ctx;        /* device context: packed sample matrix resident in HBM */
typedef struct cnb_result cnb_result;  /* list of networks indexed by complexity */

/* ---- arguments of ccnOptimalNetsForOrder (label space, exactly as R passes them) ---- */
typedef struct cnb_params {
	int32_t num_nodes;             /* N */
	int32_t num_samples;           /* M */
	const int32_t *samples;        /* [M][N], 1-based categories, NA = CNB_NA_INTEGER or < 1 */
	const int32_t *perturbations;  /* [M][N] or NULL; non-zero = sample excluded for that node as a child */
	int32_t max_parent_set;        /* rMaxParents */
	const int32_t *parent_sizes;   /* [N] or NULL (rParentSizes) */
	int32_t max_complexity;        /* rMaxComplexity (clamped up to N, catnet_search2.h:182) */
	const int32_t *order;          /* [N] 1-based labels, position -> label; NULL = 1..N (rOrder) */
	const int32_t *num_cats;       /* [N] categories per node (rNodeCats given as 1..C), or NULL = infer
	                                  max-min+1 from the data (catnet_search2.h:208-235) */
	const int32_t *parents_pool;   /* [N][N] rows of 1-based labels, 0-terminated/padded; an empty row means
	                                  "no parents allowed" (rcatnet_search.cpp:216-222); NULL = unrestricted */
	const int32_t *fixed_parents;  /* [N][N] same encoding; NULL = none (rFixedParentsPool) */
	const double  *edge_prob;      /* [N*N] the R matrix edgeProb[child,parent] in R's column-major layout, i.e.
	                                  element at parent*N + child; NULL = no prior (rMatEdgeLiks) */
	int32_t echo;                  /* rEcho: progress lines on stderr */
	/* --- extensions (zero = reference behaviour on one GPU) --- */
	int32_t device;                /* CUDA device ordinal for the one-shot calls */
	int32_t num_devices;           /* >1: use devices [device, device+num_devices) of this process (clamped to the
	                                  devices present): cnb_search_order shards the candidate parent sets over them,
	                                  cnb_search_sa runs one chain per device, cnb_search_hist deals the orders out
	                                  (cnb_mctx below).  0: the CNB_NUM_DEVICES environment variable (a number or "all")
	                                  decides, so that an unchanged R session can use the whole box; absent = 1 */
} cnb_params;

/* ---- extra arguments of ccnOptimalNetsSA ---- */
typedef struct cnb_sa_params {
	int32_t select_mode;           /* rModel: 0 = highest complexity, 1 = "AIC", 2 = "BIC" (rcatnet_sa.cpp:323-327) */
	const int32_t *start_order;    /* [N] 1-based labels (rStartOrder) */
	double temp_start, temp_cool_fact;
	int32_t temp_check_orders, max_iter;
	double order_shuffles;         /* <0: "jump" moves; fractional part = Bernoulli extra shuffle (:371-377) */
	double stop_diff;
	int32_t num_threads;           /* rThreads: proposals evaluated per SA step (first accepted wins, :646-657) */
	int32_t use_cache;             /* rUseCache (R passes TRUE, R/catnet.search.R:293-294): reproduce what the reference's score
	                                  cache does to the result -- parents listed, tables laid out and log-likelihoods summed in
	                                  the sequence of the order a family was FIRST scored under (src/cache.cpp:142-216; ignored
	                                  beyond 8 threads, rcatnet_sa.cpp:315-319).  The device still re-scores every order; a host
	                                  model of the cache's slots decides the listing (csrc/cnb_cache.h).  0 = the cache-off result */
	const double *dir_prob;        /* [N*N] R layout or NULL (rDirProbs) */
	uint64_t seed;                 /* uniform stream (splitmix64) when unif is NULL; 0 = the cnb_set_seed value */
	/* the caller's own uniform generator (the R shim passes R's unif_rand, so that set.seed() reproduces stock catnet's
	 * proposals: rcatnet_sa.cpp:146,151,225,745, utils.h:156,190).  Called on the calling thread only, draw for draw in
	 * the reference's order; must return values in (0,1).  NULL = the library's splitmix64 stream from `seed`. */
	double (*unif)(void *unif_ctx);
	void *unif_ctx;
} cnb_sa_params;

/* ---- extra arguments of ccnParHistogram (cnSearchHist, R/catnet.histo.R:45-138) ---- */
typedef struct cnb_hist_params {
	int32_t score;                 /* rScore: 0 = highest complexity, 1 = "AIC", 2 = "BIC" (rcatnet_hist.cpp:212-216) */
	int32_t weight;                /* rWeight: 0 = 1, 1 = exp(loglik/(M N)) ("likelihood"), 2 = exp(score/(M N)) */
	int32_t max_iter;              /* rMaxIter: orders to evaluate (rounded up to whole batches of num_threads) */
	int32_t num_threads;           /* rThreads: orders per batch (duplicates inside a batch are skipped) */
	int32_t use_cache;             /* rUseCache: as for cnb_sa_params (the histogram's weights follow the cached log-likelihoods) */
	uint64_t seed;                 /* uniform stream (splitmix64), as for cnb_sa_params */
	double (*unif)(void *unif_ctx);   /* the caller's generator, as for cnb_sa_params (rcatnet_hist.cpp:124,129) */
	void *unif_ctx;
} cnb_hist_params;

/* ---- one-shot entry points (what the R shim calls) ---- */
int cnb_search_order(const cnb_params *p, cnb_result **out);
int cnb_search_sa(const cnb_params *p, const cnb_sa_params *sa, cnb_result **out);
/* hist_out: double[N*N], element [parent*N + child] in label space -- the vector the R code reshapes with
 * matrix(vhisto, nrow = numnodes) (R/catnet.histo.R:131); iterations_out may be NULL */
int cnb_search_hist(const cnb_params *p, const cnb_hist_params *hp, double *hist_out, int32_t *iterations_out);
/* frees the idle per-device context (streams, events, grow-only device buffers) that one-shot calls park for the next one */
void cnb_release_cache(void);
void cnb_set_seed(int32_t seed);
const char *cnb_last_error(void);
const char *cnb_version(void);

/* ---- pairwise metrics: catnetEntropyPairwise / catnetKLpairwise / catnetPearsonPairwise /
 * catnetEntropyOrder, src/catnet_entropy.cpp:38-911, prototypes src/catnet_rexport.h, .Call table src/ccnInit.c:29-45
 * (ccnEntropyPairwise/2, ccnKLPairwise/2, ccnPearsonPairwise/2, ccnEntropyOrder/2) ----
 * samples / perturbations: int32 [M][N] as in cnb_params (R's N x M column-major matrices), categories >= 1, no
 * missing value (CNB_ERR_PARAM otherwise; the reference is undefined there); categories of a node = its value range.
 * out: double[N*N], out[n2*N + n1] = the reference's klmat -- R reshapes it with matrix(mat, N, N) to [n1, n2]:
 *   ENTROPY  H(n1 | n2) over the samples where n1 is not perturbed, diagonal H(n1);
 *   KL       sum_i2,i1 p(i1|i2) log(p(i1|i2) / q(i1|i2)), p from the samples where n1 is not perturbed, q from all;
 *   PEARSON  sum_i2,i1 (n(i2,i1) - n(i2) q(i1|i2))^2 / (n(i2) q(i1|i2)), n from the unperturbed samples of n1.
 * Without perturbations KL and PEARSON are identically 0, as in the reference (its unperturbed sub-sample is empty). */
#define CNB_PAIRWISE_ENTROPY 0
#define CNB_PAIRWISE_KL 1
#define CNB_PAIRWISE_PEARSON 2
int cnb_pairwise(int32_t kind, int32_t num_nodes, int32_t num_samples, const int32_t *samples,
		const int32_t *perturbations, int32_t device, double *out);
/* order_out: int32[N], 1-based nodes by decreasing sum of mutual informations, ties handled like the reference's
 * _order (src/utils.h:97-133): a run of equal ranks yields one node and zeros */
int cnb_entropy_order(int32_t num_nodes, int32_t num_samples, const int32_t *samples, const int32_t *perturbations,
		int32_t device, int32_t *order_out);

/* ---- a given network against data: cnLoglik, cnNodeLoglik, cnSetProb ----
 * R: .Call("ccnLoglik", object, data, perturbations, bysample) R/catnet.loglik.R:98-100, .Call("ccnNodeLoglik", object,
 * node, data, perturbations) :354-356, .Call("ccnSetProb", object, data, perturbations) R/catnet.samples.R:423-425;
 * registered in src/ccnInit.c:32-34, prototypes src/catnet_rexport.h.  The bodies of that glue are absent from the
 * reference snapshot; the semantics are those of the class methods it forwards to: CATNET::sampleLoglikVector,
 * bySampleLoglikVector, sampleNodeLoglik (src/catnet_class.h:616-815) and setNodeSampleProb (:818-879).
 * cnb_network is what RCatnet::RCatnet(SEXP) reads off the S4 catNetwork (src/rcatnet.cpp:38-172). */
typedef struct cnb_network {
	int32_t num_nodes;
	int32_t max_parents;          /* row stride of parents[] (>= every num_parents[i], <= 8) */
	const int32_t *num_cats;      /* [N] categories per node */
	const int32_t *num_parents;   /* [N] */
	const int32_t *parents;       /* [N][max_parents] 1-based nodes in the object's listed order: the first listed parent is
	                                 the most significant index of the node's table (src/problist.h:252-263) */
	const double *probs;          /* conditional tables, node after node, child category fastest; NULL for cnb_set_prob */
	const int64_t *prob_offsets;  /* [N + 1] start of every node's table in probs[] */
} cnb_network;
#define CNB_LOGLIK_NODES 0        /* out[N]: per node, mean over its complete unperturbed samples of log P(x | parents) */
#define CNB_LOGLIK_BYSAMPLE 1     /* out[M]: per sample, sum over the nodes */
/* samples / perturbations as in cnb_params; values outside 1..num_cats[i] (NA_INTEGER, < 1) are missing and skip the
 * sample for the families they touch; a value above num_cats[i] is CNB_ERR_PARAM ("Data has more categories than the
 * object", R/catnet.loglik.R:91-92).  A zero probability met by a counted sample gives -FLT_MAX. */
int cnb_loglik(const cnb_network *net, int32_t num_samples, const int32_t *samples, const int32_t *perturbations,
		int32_t mode, int32_t device, double *out);
/* the per-node values of CNB_LOGLIK_NODES for the listed nodes only (1-based) */
int cnb_node_loglik(const cnb_network *net, int32_t num_sel, const int32_t *nodes, int32_t num_samples,
		const int32_t *samples, const int32_t *perturbations, int32_t device, double *out);
/* cnSetProb: tables of the network's families estimated from the data (counts normalised per parent configuration,
 * unobserved configurations uniform, src/problist.h:193-222) into probs_out (laid out by net->prob_offsets), the
 * families' sample-averaged log-likelihoods into loglik_out[N], their sample sizes into sizes_out[N] */
int cnb_set_prob(const cnb_network *net, int32_t num_samples, const int32_t *samples, const int32_t *perturbations,
		int32_t device, double *probs_out, double *loglik_out, int32_t *sizes_out);

/* ---- cnSamples: .Call("ccnSamples", object, numsamples, perturbations, naRate)
 * R/catnet.samples.R:60-62, src/ccnInit.c:42 -> RCatnet::genSamples, src/rcatnet.cpp:524-650 ----
 * out: int32 [M][N] (R's N x M matrix), 1-based categories, NA_INTEGER where naRate blanked a value.  perturbations
 * [M][N] or NULL: a value in 1..num_cats[i] fixes node i in that sample.  Nodes are drawn parents first
 * (CATNET::getOrder), one uniform per unfixed (node, sample), category = first i with u <= p_0 + ... + p_i.
 * R's unif_rand cannot cross a C ABI: the uniforms come from the library's splitmix64 stream (draw t is a function of
 * seed and t), consumed in exactly the reference's order -- so with the same stream injected into the reference the
 * samples are identical. */
int cnb_samples(const cnb_network *net, int32_t num_samples, const int32_t *perturbations, double na_rate, uint64_t seed,
		int32_t device, int32_t *out);
/* the same with perturbations and output in device memory (the matrix can go straight into cnb_ctx_set_samples_dev) */
int cnb_samples_dev(const cnb_network *net, int32_t num_samples, const int32_t *perturbations_dev, double na_rate,
		uint64_t seed, int32_t device, int32_t *out_dev);

/* ---- resident-context API (SA, bench, multi-process sharding) ---- */
int cnb_ctx_create(int32_t device, cnb_ctx **out);
void cnb_ctx_destroy(cnb_ctx *ctx);
/* copy the sample matrix to the device and pack it (K1): uint8 column-major codes + one-hot bit planes */
int cnb_ctx_set_samples(cnb_ctx *ctx, int32_t num_nodes, int32_t num_samples, const int32_t *samples,
		const int32_t *perturbations, const int32_t *num_cats);
/* same, but the [M][N] int32 matrix (and optional perturbations) already live in device memory */
int cnb_ctx_set_samples_dev(cnb_ctx *ctx, int32_t num_nodes, int32_t num_samples, const int32_t *samples_dev,
		const int32_t *perturbations_dev, const int32_t *num_cats);
/* same with one BYTE per value: 0 = missing, 1..254 = category, 255 = not a category (the form in which the library
 * itself moves large host matrices over PCIe; multi-process callers all-gather this instead of the int32 matrix) */
int cnb_ctx_set_samples_dev8(cnb_ctx *ctx, int32_t num_nodes, int32_t num_samples, const uint8_t *samples_dev,
		const int32_t *perturbations_dev, const int32_t *num_cats);
/* first step of a multi-process upload: rows [row0, row1) of the host matrix (pageable memory is fine) are narrowed to that
 * byte form by host threads and copied to dst_dev + row0 * num_nodes on the context's stream; the ranks all-gather the byte
 * matrix over NVLink and pass it to cnb_ctx_set_samples_dev8.  CNB_ERR_PARAM if a label is above 255. */
int cnb_ctx_upload_rows8(cnb_ctx *ctx, int32_t num_nodes, const int32_t *samples, int64_t row0, int64_t row1,
		uint8_t *dst_dev, int32_t max_threads);
/* the same on a stream of the caller's (NULL = the context's), staging the bytes at stage_offset of a pinned buffer of
 * stage_capacity values (0 = a buffer for just this call): successive calls with disjoint staging ranges let the host
 * threads narrow one slice while the copy engine still moves the previous one (the streamed search below) */
int cnb_ctx_upload_rows8_on(cnb_ctx *ctx, int32_t num_nodes, const int32_t *samples, int64_t row0, int64_t row1,
		uint8_t *dst_dev, int32_t max_threads, void *copy_stream, int64_t stage_offset, int64_t stage_capacity);
/* categories per node as used on the device (given or inferred), label space */
int cnb_ctx_num_cats(const cnb_ctx *ctx, int32_t *num_cats_out);
/* search for one order on the resident data; samples/perturbations/num_cats/device fields of p are ignored */
int cnb_ctx_search_order(cnb_ctx *ctx, const cnb_params *p, cnb_result **out);
/* simulated annealing over orders on the resident data (per-order evaluation on the device) */
int cnb_ctx_search_sa(cnb_ctx *ctx, const cnb_params *p, const cnb_sa_params *sa, cnb_result **out);
/* cnSearchHist on the resident data */
int cnb_ctx_search_hist(cnb_ctx *ctx, const cnb_params *p, const cnb_hist_params *hp, double *hist_out,
		int32_t *iterations_out);

/* pairwise metrics on the resident data (whatever categories the context was given: empty categories change nothing) */
int cnb_ctx_pairwise(cnb_ctx *ctx, int32_t kind, double *out);
int cnb_ctx_entropy_order(cnb_ctx *ctx, int32_t *order_out);

/* a given network on the resident data (the context must have been given net->num_cats as its categories) */
int cnb_ctx_loglik(cnb_ctx *ctx, const cnb_network *net, int32_t mode, int32_t num_sel, const int32_t *nodes, double *out);
int cnb_ctx_set_prob(cnb_ctx *ctx, const cnb_network *net, double *probs_out, double *loglik_out, int32_t *sizes_out);

/* ---- one process, several GPUs (the shape an R session has; reference: the worker pool inside one process,
 * src/rcatnet_sa.cpp:475-652, src/thread.h).  A cnb_mctx owns one context per device and one host thread per device.
 *   set_samples   every device copies only its 1/G share of the host matrix over its own PCIe link (compacted to bytes
 *                 by host threads, so pageable memory -- R's -- is fine) and completes the others' copies over NVLink
 *                 (peer copies); then each packs the full matrix
 *   search_order  every device plans, scores its shard of the candidate parent sets and writes the scores straight into
 *                 the first device's score buffer over NVLink (peer stores from the scoring kernels; a peer copy of the
 *                 segment where peer access is not available); selection + DP + export on the first device only
 *   search_sa     one chain per device (chain 0 = the single-device chain, draw for draw; chain g > 0 draws from
 *                 splitmix64(seed + g)), merged per complexity
 *   search_hist   the orders of a batch are evaluated side by side, one per device, and accumulated in the reference's
 *                 order: the histogram is bit-identical to the single-device one
 * Small problems stay on the first device (the exchange costs more than it saves): CNB_MULTI_MIN_WORK (family-samples,
 * default 2e10) sets the threshold, 0 shards always. */
typedef struct cnb_mctx cnb_mctx;
/* num_devices < -1 (also accepted by cnb_params.num_devices) is a test hook: |num_devices| contexts on the ONE device
 * first_device, so that the whole orchestration (shares, peer stores, chains, dealt orders) runs on a single-GPU box */
int cnb_mctx_create(int32_t first_device, int32_t num_devices, cnb_mctx **out);
void cnb_mctx_destroy(cnb_mctx *m);
int32_t cnb_mctx_num_devices(const cnb_mctx *m);
cnb_ctx *cnb_mctx_ctx(cnb_mctx *m, int32_t i);           /* the context of the i-th device (timing, test hooks) */
int cnb_mctx_set_samples(cnb_mctx *m, int32_t num_nodes, int32_t num_samples, const int32_t *samples,
		const int32_t *perturbations, const int32_t *num_cats);
int cnb_mctx_search_order(cnb_mctx *m, const cnb_params *p, cnb_result **out);
/* the two halves of cnb_mctx_search_order: plan + score + exchange + selection + DP (result stays on the first device),
 * then network export */
int cnb_mctx_search_light(cnb_mctx *m, const cnb_params *p);
int cnb_mctx_collect(cnb_mctx *m, cnb_result **out);
int cnb_mctx_search_sa(cnb_mctx *m, const cnb_params *p, const cnb_sa_params *sa, cnb_result **out);
int cnb_mctx_search_hist(cnb_mctx *m, const cnb_params *p, const cnb_hist_params *hp, double *hist_out,
		int32_t *iterations_out);
/* device timing across the devices: mark(0) / mark(1) record an event on the first device's stream (every other
 * device's work of a search is ordered behind the first device's stream position at the start of that search and
 * before its selection); elapsed = ms between the two marks, after synchronising every device */
int cnb_mctx_mark(cnb_mctx *m, int32_t which);
int cnb_mctx_elapsed_ms(cnb_mctx *m, float *ms);
/* devices the last search really used (1 when the problem was below the sharding threshold) */
int32_t cnb_mctx_active(const cnb_mctx *m);

/* sharded search, one process per GPU: plan -> score own shard -> (caller all-gathers) -> finish.
 * The score buffer is one fp64 per candidate family, laid out rank-major: segment r (seg_len doubles) holds rank r's
 * families.  cnb_ctx_score_shard writes this rank's segment at scores_dev + rank*seg_len; after an all-gather of
 * the segments every rank passes the full buffer to cnb_ctx_finish, which runs arg-max + DP on the device. */
int cnb_ctx_plan(cnb_ctx *ctx, const cnb_params *p, int32_t rank, int32_t world, int64_t *seg_len, int64_t *num_families);
int cnb_ctx_score_shard(cnb_ctx *ctx, double *scores_dev);
int cnb_ctx_finish(cnb_ctx *ctx, const double *scores_dev, cnb_result **out);
/* the two halves of cnb_ctx_finish: selection + DP on the device (result stays there), then network export */
int cnb_ctx_select_dp(cnb_ctx *ctx, const double *scores_dev);
int cnb_ctx_collect(cnb_ctx *ctx, cnb_result **out);
/* ---- streamed search: the samples are still on their way to the device ----
 * For a large complete matrix (categories given, <= 4 of them, no perturbations / pools / fixed parents, >= 8192 samples)
 * the host -> device copy, the packing and the tensor-core contraction overlap: cnb_ctx_stream_begin plans the search
 * (rank / world as in cnb_ctx_plan) and cuts the samples into num_slices slices of whole sample groups, rows
 * [slice_rows[k], slice_rows[k+1]) (num_slices = 0: the problem does not qualify, use the resident calls);
 * cnb_ctx_stream_slice(k) is told that slice k of the BYTE matrix [M][N] (cnb_ctx_set_samples_dev8's form, base pointer
 * bytes_dev) is complete once ready_event (a cudaEvent_t; NULL = once the work already queued on the context's stream is
 * done) has happened: it packs the slice on a side stream and runs one pass of the table kernel over this rank's tiles for
 * the slice's samples, adding to the passes before it -- the caller meanwhile uploads / all-gathers slice k + 1
 * (cnb_ctx_upload_rows8_on a stream of its own); cnb_ctx_stream_end scores this rank's shard into scores_dev exactly as
 * cnb_ctx_score_shard would (then: gather, cnb_ctx_finish).  ok = 0: a missing value or a category outside its node's
 * range turned up -- nothing was scored, the caller continues with cnb_ctx_set_samples_dev8 + cnb_ctx_plan +
 * cnb_ctx_score_shard on the byte matrix it has assembled.  cnb_search_order does all this by itself for host matrices
 * of 256 MB and more (CNB_STREAM_MIN, bytes or "off"; CNB_STREAM_SLICES, relative slice sizes, default "1,5,10"). */
int cnb_ctx_stream_begin(cnb_ctx *ctx, const cnb_params *p, int32_t rank, int32_t world, int64_t *seg_len,
		int64_t *num_families, int32_t *num_slices, int64_t *slice_rows, int32_t slice_rows_cap);
int cnb_ctx_stream_slice(cnb_ctx *ctx, int32_t k, const uint8_t *bytes_dev, void *ready_event);
int cnb_ctx_stream_end(cnb_ctx *ctx, double *scores_dev, int32_t *ok);
/* Sharded searches whose three-parent group runs on the tensor cores deal the T tiles of every column tile to the ranks:
 * a rank then scores a range of parent triples for ALL children of the column tile, i.e. writes into every rank's segment
 * of the score buffer.  The caller opts in by saying how the buffers are combined: mode 1 = every rank's scores_dev is
 * the SAME buffer (peer memory; what cnb_search_order with num_devices does), mode 2 = every rank has its own buffer,
 * cnb_ctx_score_shard starts it at -inf and the caller combines the buffers with an element-wise MAX (all-reduce) instead
 * of the all-gather whenever cnb_ctx_scattered() says the last cnb_ctx_score_shard scattered; mode 0 (default) = never:
 * such groups then take the bit-plane kernels on several ranks. */
int cnb_ctx_allow_scatter(cnb_ctx *ctx, int32_t mode);
int32_t cnb_ctx_scattered(const cnb_ctx *ctx);
/* launch on the caller's CUDA stream (e.g. the stream its NCCL collectives use) instead of the context's own
 * non-blocking stream; NULL selects the context's own stream again -- to name the legacy default stream pass
 * cudaStreamLegacy ((cudaStream_t)0x1).  Work the caller orders against this library (copies into the sample
 * matrix, the all-gather of the score segments) must be on that stream. */
int cnb_ctx_set_stream(cnb_ctx *ctx, void *cuda_stream);

/* per-phase device times of the last search on this context, CUDA events on the library's stream (ms) */
typedef struct cnb_timing {
	float pack_ms, plan_ms, score_ms, select_ms, dp_ms, collect_ms, total_ms;
	float score_ms_by_parents[8];   /* [d] scoring kernel time for candidate sets with d parents */
	int64_t families_by_parents[8];
	int64_t num_families;           /* candidate families scored by this rank */
	int64_t algorithmic_bytes;      /* sum over scored families of (d+1)*M + 8 */
	int32_t kernel_launches;
	int32_t fast_path;              /* 1 = bit-plane kernels, 0 = generic histogram kernel */
	int64_t h2d_bytes, d2h_bytes;   /* host<->device traffic since the last cnb_ctx_set_samples* call */
	int64_t tensor_tiles;           /* tcgen05 tiles (2 x 128 rows x up to 192 columns) run by this rank (0 = tensor-core
	                                   path not used) */
	float tensor_ms;                /* device time of the tcgen05 table kernel alone (CUDA events) */
	int32_t tensor_kind;            /* operand type of that kernel: 4 = E2M1 (kind::mxf4), 8 = u8 (kind::i8), 0 = none */
	int64_t tensor_macs;            /* multiply-accumulates it executed: sum over tiles of 256 rows * its columns * padded samples */
	int64_t tensor3_tiles;          /* tcgen05 tiles of the three-parent group (0 = that group ran on the bit-plane kernels) */
} cnb_timing;
int cnb_ctx_timing(const cnb_ctx *ctx, cnb_timing *out);
/* the same for the last one-shot cnb_search_order of the calling thread (its context is parked, not reachable);
 * fast_path = 2 there means the streamed form ran (upload overlapped with the table kernel) */
int cnb_last_call_timing(cnb_timing *out);

/* ---- test hooks ---- */
/* counts + log-lik of explicit families; children/parents are 0-based node LABELS; counts_out is
 * [nfam][table_stride] int32 in the reference's table layout (first listed parent most significant, child
 * fastest; problist.h:143-155, 252-263); loglik_out is the sample-averaged log-lik (catnet_class.h:867-871). */
int cnb_family_scores(cnb_ctx *ctx, int32_t nfam, int32_t num_parents, const int32_t *children,
		const int32_t *parents, int32_t table_stride, int32_t *counts_out, double *loglik_out,
		int32_t *ncount_out, int32_t force_generic);
/* the scores the last cnb_ctx_score_shard wrote for explicit candidate families (whichever kernel scored their group, the
 * tcgen05 tables included): children / parents are order POSITIONS (parents ascending, [nfam][num_parents]); scores_dev is
 * the full rank-major buffer.  NaN for a family the plan does not offer.  bench.py's parity_spot reads the benchmarked
 * kernel's own output with this. */
int cnb_ctx_search_scores(cnb_ctx *ctx, const double *scores_dev, int32_t nfam, int32_t num_parents,
		const int32_t *children, const int32_t *parents, double *out);
/* kernel selection of this context for the parity tests (0 = the library chooses).  Bits of `on`: 1 generic histogram
 * kernel everywhere; 2 direct bit-plane scorer (no inclusion-exclusion); 4 tensor-core scorer whenever applicable, 8
 * never; 16 u8 operands (kind::i8) instead of E2M1; 32 one DP launch per node; 64 grid-barrier DP; 128 pair tables by
 * the POPC kernel (set before the samples); 512 no inclusion-exclusion for three parents; 1024 three parents never on
 * the tensor cores; 2048 wavefront DP with one node per exchange; 4096 generic tensor-core epilogue also for
 * 4-category data; 8192 full-table tensor tiles (the form that takes missing values and perturbation masks) also on
 * complete data; 16384 trapezoid DP in which every slot re-sums its own candidates (no CTA-wide work list); 32768 the
 * histogram scorer with one CTA per family, 65536 with a warp per family and __match_any_sync aggregation (default: a warp
 * per family, shared-memory atomics); 131072 tensor-core tiles always in the 28 x 64 geometry of four categories.  Every
 * combination gives the same networks.  The one-shot calls take the same bits from CNB_FORCE_BITS (fuzzing). */
int cnb_ctx_force_generic(cnb_ctx *ctx, int32_t on);
/* evaluate the device's log on n inputs (must be bit-identical to the host libm the reference links) */
int cnb_device_log(int32_t device, const double *x, int64_t n, double *out);
/* the same log evaluated on the host by this library (no GPU needed) */
void cnb_host_log(const double *x, int64_t n, double *out);
/* the host tail of cnb_entropy_order (ranks = sums of mutual informations, the reference's _order with its tie behaviour)
 * on a given entropy matrix mat[n2 * N + n1]; no GPU needed */
void cnb_host_entropy_order(const double *mat, int32_t num_nodes, int32_t *order_out);
/* lexicographic unrank used by the device enumerator (host mirror; catnet_search2.h:102-147 order) */
int cnb_unrank_combination(int32_t n, int32_t k, int64_t rank, int32_t *out);
int64_t cnb_num_families(const cnb_params *p);
/* consistency of the tensor-core tiling of the unrestricted two-parent group of N nodes (host only): number of
 * tiles, or < 0 if family <-> tile slot is not a bijection */
int64_t cnb_tile_selftest(int32_t num_nodes);
int64_t cnb_tile_selftest_full(int32_t num_nodes);   /* the full-table geometry (16 pairs x 48 column nodes per tile) */
/* the geometries of data with free_cats + 1 categories per node: 28 x 64 slots (free_cats = 3), 64 x 96 (2), 256 x 192 (1) */
int64_t cnb_tile_selftest_geo(int32_t num_nodes, int32_t free_cats);
/* the same for the tiling of the unrestricted three-parent group (free_cats = max categories - 1, in the geometry of that
 * many free categories; negative: -free_cats free values in the default 28 x 64 tiles) */
int64_t cnb_tile3_selftest(int32_t num_nodes, int32_t free_cats);
/* the host-side model of the reference's score cache (src/cache.cpp:142-317; csrc/cnb_cache.h) on a sequence of node orders,
 * no GPU needed.  p: the search (num_nodes, max_parent_set, parent_sizes, pools, edge_prob; order ignored); cats [N] per
 * label; never_kept [N] bytes or NULL (nodes perturbed in every sample); orders [norders][N] 1-based labels; winners
 * [norders][N][P][P] or NULL: the listed labels (0-padded) of the winner of (order, child label - 1, parent count - 1) where
 * all pool members have equal category counts, as the device would report it.  Every cache HIT (all_hits != 0), or only
 * the hits that change what the device computed for the order (the overrides the search acts on), comes back as a row of
 * 4 + 8 ints -- 1 equal-categories path / 2 otherwise, order index, child label, parent count, the parent labels in the
 * cached listing sequence -- with the entry's prior beside it.  Returns the number of hits (rows beyond cap are dropped),
 * < 0 on an error. */
int64_t cnb_cache_trace(const cnb_params *p, const int32_t *cats, const uint8_t *never_kept, int32_t norders,
		const int32_t *orders, const int32_t *winners, int32_t all_hits, int32_t cap, int32_t *rows, double *priors);

/* ---- result accessors (the catNetwork slots) ---- */
int32_t cnb_result_num_networks(const cnb_result *r);
int32_t cnb_result_num_nodes(const cnb_result *r);
/* network i (ascending complexity): complexity, log-likelihood */
int cnb_result_network(const cnb_result *r, int32_t i, int32_t *complexity, double *loglik);
/* parents of every node of network i, order space (node k = k-th of the order; 0-based positions):
 * num_parents[N]; parents [N][max_parents_cap] */
int cnb_result_parents(const cnb_result *r, int32_t i, int32_t *num_parents, int32_t *parents, int32_t max_parents_cap);
/* the order network i is expressed in: position -> 1-based node label (differs per network after a merge) */
int cnb_result_order(const cnb_result *r, int32_t i, int32_t *order_out);
/* categories per node of network i, in its order space */
int cnb_result_cats(const cnb_result *r, int32_t i, int32_t *cats_out);
/* raw sample counts of a node's table (before normalisation); returns the table size */
int cnb_result_counts(const cnb_result *r, int32_t i, int32_t node, int32_t *counts, int32_t cap);
/* node-level view: log-lik, prior, sample size, CPT (normalised like problist.h:193-222); returns table size */
int cnb_result_node(const cnb_result *r, int32_t i, int32_t node, double *loglik, double *prior, int32_t *sample_size,
		double *probs, int32_t probs_cap);
/* SA only: accepted order (1-based labels), iterations, accepted steps */
int cnb_result_sa_info(const cnb_result *r, int32_t *order_out, int32_t *niter, int32_t *naccept);
/* SA only: log-lik of every evaluated proposal and whether it was accepted; returns the trace length */
int cnb_result_sa_trace(const cnb_result *r, double *loglik, int32_t *accepted, int32_t cap);
/* merge b into a per complexity keeping the higher likelihood (R/catnet.search.R:317-332) */
int cnb_result_merge(cnb_result *a, const cnb_result *b);
void cnb_result_free(cnb_result *r);

#ifdef __cplusplus
}
#endif
#endif /* CATNET_B200_H */
