/* cnb_ctx.h -- the device context behind the opaque cnb_ctx of the C ABI. */
#ifndef CNB_CTX_H
#define CNB_CTX_H
#include <cuda_runtime.h>
#include <vector>
#include "cnb_internal.h"
#include "cnb_result.h"
#include "cnb_cache.h"

#define CNB_HIST_MAX_CELLS 16384   /* largest conditional table of the generic scorer (shared memory) */
#define CNB_NEV_GROUPS 8
#define CNB_NEV_TEN 8        /* timed batches of the tensor-core kernel */

enum { EV_PACK0, EV_PACK1, EV_PLAN0, EV_PLAN1, EV_SCORE0, EV_SCORE1, EV_SEL0, EV_SEL1, EV_DP1, EV_COL0, EV_COL1,
	EV_GROUP0, EV_TEN0 = EV_GROUP0 + 2 * CNB_NEV_GROUPS, CNB_NEV = EV_TEN0 + 2 * CNB_NEV_TEN };

struct DevBuf {
	void *ptr = nullptr;
	size_t cap = 0;
	bool view = false;          /* ptr points into another buffer (the packed plan, cnb_ctx.b_plan): never freed from here */
	int ensure(size_t bytes);   /* grow-only; contents are not preserved */
	void release();
	void alias(void *p, size_t bytes);   /* become a view of bytes at p (an own allocation is freed first) */
};

/* grow-only page-locked host staging buffer (device -> host results) */
struct PinBuf {
	void *ptr = nullptr;
	size_t cap = 0;
	int ensure(size_t bytes);
	void release();
};

/* what a plan depends on besides the order, when it depends on the order only through the order itself (cnb_ctx_plan) */
struct CnbPlanKey {
	int32_t N, M, P, K, rank, world, tile_mode, cats, order_free;
};

struct cnb_ctx {
	CnbPlanKey plan_key = {};
	bool plan_key_valid = false;
	int device = 0;
	cudaStream_t stream = nullptr, own_stream = nullptr;
	cudaStream_t copy_stream = nullptr;   /* host -> device copies of the streamed one-shot search (created on first use) */
	cudaStream_t prep_stream = nullptr;   /* ... and the packing of a slice, under the table pass over the previous one */
	std::vector<cudaEvent_t> ev_slices;   /* "slice k is on the device" */
	bool tables_resident = false;         /* b_counts still holds the tile tables of the whole two-parent group of the current plan
	                                         (one batch, one rank): the export reads the returned networks' tables from there */
	int allow_scatter = 0;                /* a sharded search may write scores outside this rank's segment (three-parent T tiles):
	                                         1 = every rank stores into ONE buffer (peer memory), 2 = the caller max-reduces the buffers */
	bool scattered = false;               /* the last cnb_ctx_score_shard did */
	bool no_reduced_tiles = false;        /* tensor-core tiles always contract 3 free categories per node (test / A-B switch) */
	int hist_mode = 1, hist_mode_default = 1;   /* generic scorer: 0 one CTA per family, 1 a warp per family + shared atomics, 2 + __match_any_sync (CNB_HIST_MODE) */
	bool stream_active = false;           /* between cnb_ctx_stream_begin and cnb_ctx_stream_end */
	std::vector<int> stream_gend;         /* last sample group (exclusive) of every slice */
	int stream_next = 0, stream_nq = 0;
	long long stream_tiles = 0;           /* this rank's tiles of the two-parent group */
	bool streaming = false;               /* the streamed one-shot search is driving this context: cnb_ctx_plan leaves the data
	                                         kernels to it, cnb_ctx_score_shard finds the two-parent tables contracted already */
	cudaEvent_t ev[CNB_NEV];
	int N = 0, M = 0, W = 0, Mpad = 0, max_cats = 0;
	bool has_pert = false, have_samples = false, planned = false, dp_done = false, force_generic = false, last_fast = false, pairs_valid = false, no_ie = false, clean = false;
	int tensor_mode = 0;               /* 0 auto (large M), 1 always when applicable, -1 never */
	int tensor_i8 = 0;                 /* 1: kind::i8 operands instead of kind::mxf4 (test hook) */
	bool dp_per_node = false;          /* one DP launch per node instead of the fused cooperative kernel (test hook) */
	bool in_host_set = false;          /* cnb_ctx_set_samples is calling cnb_ctx_set_samples_dev (keeps its H2D count) */
	bool pairs_popc = false;           /* pair tables by the POPC kernel even for large M (test hook) */
	bool dp_no_wave = false;           /* grid-barrier DP kernel instead of the wavefront one (test hook) */
	bool triples_valid = false;        /* b_triples holds the three-way tables of this data set (built on first use) */
	bool no_tensor3 = false;           /* three-parent families never on the tensor cores (test hook) */
	int tensor3_tiles = 0;             /* T tiles of the last search (0 = three-parent group not on the tensor cores) */
	bool no_ie3 = false;               /* three-parent families by the direct bit-plane / histogram scorers (test hook) */
	std::vector<uint8_t> never_kept;   /* [N] label space: node perturbed in every sample (no candidate parent sets) */
	bool any_never_kept = false;
	bool generic_epilogue = false;     /* k_umma_epilogue<false> also when every node has 4 categories (test hook) */
	bool pairs_ord = false;            /* b_pairs_ord holds the tables of all ORDERED pairs under the child's keep mask, missing values
	                                      skipped: the 0- and 1-parent families of masked data need no pass over the samples */
	bool dp_chains = false;            /* trapezoid DP: every slot re-sums its own candidates instead of the CTA-wide work list (test hook) */
	bool force_full = false;           /* full-table tensor tiles also on complete data (test hook) */
	bool dp_no_trap = false;           /* k_dp_wave instead of the trapezoid kernel (test hook) */
	bool has_na = false;               /* the data set has missing values */
	bool quiet = false;                /* no per-phase timing events (the SA / histogram drivers: fewer driver calls per evaluated order) */
	bool pairs_full = false;           /* b_pairs holds the two-way tables over ALL samples (perturbation masks ignored) */
	bool ignore_pert = false;          /* score as if there were no perturbation masks (full tables of the pairwise metrics) */
	/* concurrent evaluation of several orders on this device (the numThreads proposals of a cnSearchSA step, the orders of a
	 * cnSearchHist batch): sibling contexts holding a copy of the packed data, one host thread each (cnb_ctx_batch) */
	std::vector<cnb_ctx *> siblings;
	struct Crew *crew = nullptr;
	unsigned long long data_epoch = 0, cloned_epoch = 0;   /* set_samples count of this context / of the data a sibling holds */
	/* the reference's score cache (cnSearchSA / cnSearchHist with use_cache, cnb_cache.h): the driver's model of it; while it
	 * is set, cnb_ctx_select_dp walks every planned order through it and patches the option table before the DP, and
	 * cnb_ctx_collect lists the parents of the returned families in the cached sequence */
	CnbScoreCache *cache = nullptr;
	std::vector<CnbOverride> cache_ov;     /* overrides of the order evaluated last, ascending option slot */
	std::vector<int32_t> cache_psizes;     /* the search's parentSizes (label space; empty = none) and maxParentSet, kept by cnb_ctx_plan */
	int32_t cache_P = 0;
	std::vector<int64_t> cache_h_opt, cache_h_fid;
	std::vector<int32_t> cache_h_cx;
	std::vector<double> cache_h_prior;
	bool defer_ovf = false;                /* a quiet cnb_ctx_score_shard left its overflow flag on the device for cnb_ctx_dp_launch */
	int cache_ovf = 0;                     /* overflow flag of the overrides' scorer, copied back behind the DP */
	int dp_final = 0, tensor_batches = 0;
	double host_ms[4] = {0, 0, 0, 0};   /* CNB_TRACE: host wall time of plan / score / select+dp, searches */
	std::vector<int32_t> cats_lbl;
	std::vector<int32_t> blk_equal;
	std::vector<int32_t> sel_chunks, sel_chunk0;   /* (block, chunk) per CTA of k_select; first chunk of every block */
	std::vector<CnbDpNode> dp_nodes;
	std::vector<unsigned char> h_exists;
	std::vector<double> h_L;
	CnbPlan plan;
	cnb_timing timing = {};
	PinBuf pin_sum;                    /* DP summary on its way back (overflow flag of the cache overrides, exists[K+1], L[K+1]): page-locked so
	                                      that the copy is queued behind the DP instead of waiting for it (cnb_ctx_dp_launch / _wait) */
	PinBuf pin_sel, pin_stage;         /* pin_stage: host-compacted copy of a large sample matrix on its way to the device */
	DevBuf b_counts_p, b_pairs_ord, b_sel_slots, b_stage_s, b_stage_p, b_pw_kept, b_pw_out, b_net_int, b_net_off, b_net_probs, b_net_logp, b_net_sel, b_net_fz, b_sel_chunks, b_sel_chunk0, b_sel_part_s, b_sel_part_f, b_triples, b_rows_a3, b_tile_base3;
	/* packed samples */
	DevBuf b_vmin, b_vmax, b_cats_lbl, b_flag, b_codes, b_planes_lbl, b_keep_lbl, b_planes, b_keep, b_pairs, b_rows_a, b_rows_b;
	/* plan: one allocation and one host -> device copy per planned order (b_plan, staged in pin_plan); the buffers below are views into it */
	DevBuf b_plan;
	PinBuf pin_plan;
	DevBuf b_order, b_cats, b_lists, b_blocks, b_group_blocks, b_groups, b_edge, b_blk_equal, b_tile_base, b_dtile_off, b_fix1;
	bool plan_has_fix1 = false;
	/* scoring */
	DevBuf b_scores, b_counts, b_ex_child, b_ex_pars, b_ex_nd, b_ex_ll, b_ex_prior, b_ex_score, b_ex_ncount, b_ex_counts;
	/* selection + DP */
	DevBuf b_ov_opt, b_ov_fid, b_ov_cx, b_ov_prior;
	DevBuf b_base_v, b_opt_score, b_opt_cx, b_opt_fid, b_dp_exists, b_dp_L, b_dp_pfx, b_bp_opt, b_bp_src, b_sel, b_dp_nodes, b_dp_flags;
	std::vector<DevBuf *> all_bufs() {
		return {&b_counts_p, &b_pairs_ord, &b_vmin, &b_vmax, &b_cats_lbl, &b_flag, &b_codes, &b_planes_lbl, &b_keep_lbl, &b_planes, &b_keep,
			&b_order, &b_cats, &b_lists, &b_blocks, &b_group_blocks, &b_groups, &b_edge, &b_blk_equal,
			&b_scores, &b_counts, &b_ex_child, &b_ex_pars, &b_ex_nd, &b_ex_ll, &b_ex_prior, &b_ex_score, &b_ex_ncount,
			&b_ex_counts, &b_base_v, &b_opt_score, &b_opt_cx, &b_opt_fid, &b_dp_exists, &b_dp_L, &b_dp_pfx, &b_bp_opt,
			&b_bp_src, &b_sel, &b_pairs, &b_tile_base, &b_dtile_off, &b_rows_a, &b_rows_b, &b_dp_nodes, &b_sel_slots, &b_dp_flags, &b_stage_s, &b_stage_p, &b_pw_kept, &b_pw_out, &b_net_int, &b_net_off, &b_net_probs, &b_net_logp, &b_net_sel, &b_net_fz, &b_sel_chunks, &b_sel_chunk0, &b_sel_part_s, &b_sel_part_f, &b_triples, &b_rows_a3, &b_tile_base3, &b_fix1,
			&b_ov_opt, &b_ov_fid, &b_ov_cx, &b_ov_prior, &b_plan};
	}
};

uint64_t cnb_default_seed(void);   /* what cnb_set_seed stored (cnb_sa.cpp) */
/* one-shot entry points: take over / park the idle context of a device (cnb_api.cu) */
int cnb_ctx_acquire(int32_t device, cnb_ctx **out);
void cnb_ctx_park(cnb_ctx *c);
void cnb_pool_release(void);
/* contexts able to evaluate `want` orders at once on c's device: c itself and up to want - 1 siblings holding a copy of its
 * packed data (small data sets only; 1 = evaluate one after the other).  cnb_ctx_batch_run evaluates orders[t] on ctxs[t]
 * (plan + score + selection + DP, no export) concurrently, one host thread each. */
int cnb_ctx_batch(cnb_ctx *c, int want, std::vector<cnb_ctx *> *ctxs);
int cnb_ctx_batch_run(cnb_ctx *c, const std::vector<cnb_ctx *> &ctxs, const cnb_params *p, int n, const int32_t *const *orders,
		bool to_options = false);   /* to_options: stop after the option table (cnb_ctx_search_to_options) */
/* internal phases shared with the SA driver (cnb_sa.cpp) */
int cnb_ctx_search_light(cnb_ctx *c, const cnb_params *p);           /* plan + score + DP, no network export */
int cnb_ctx_search_to_options(cnb_ctx *c, const cnb_params *p);      /* plan + score + option table; then: */
int cnb_ctx_dp_launch(cnb_ctx *c);                                   /* [score-cache overrides,] DP, summary copy enqueued */
int cnb_ctx_dp_wait(cnb_ctx *c);                                     /* ... drained: the evaluation is complete */
int cnb_ctx_dp_summary(const cnb_ctx *c, std::vector<int32_t> *complexity, std::vector<double> *loglik);
/* parents (order positions, [N][CNB_MAXP]) of the surviving network with the given complexity, after select_dp */
int cnb_ctx_network_parents(cnb_ctx *c, int32_t complexity, std::vector<int32_t> *num_parents, std::vector<int32_t> *parents);

/* upload paths shared with the multi-device driver (cnb_multi.cu) */
int cnb_ctx_set_samples_impl(cnb_ctx *c, int32_t N, int32_t M, const void *samples_dev, int elem_bytes, const int32_t *pert_dev,
		const int32_t *num_cats);
int cnb_ctx_upload_bytes(cnb_ctx *c, const int32_t *samples, size_t v0, size_t v1, uint8_t *dst_dev, int max_threads, bool *overflow,
		cudaStream_t copy_stream = nullptr, size_t pin_offset = 0, size_t pin_capacity = 0);
void cnb_set_last_call_timing(const cnb_timing &t);   /* what cnb_last_call_timing reports for the calling thread */
void cnb_ctx_stream_cancel(cnb_ctx *c);   /* drop a streamed search begun with cnb_ctx_stream_begin (drains the context's streams) */
/* cnSearchHist, one order (cnb_hist.cpp) */
struct CnbHistEval {
	bool found = false;
	double w = 0;
	std::vector<int32_t> npar, pars;
};
int cnb_hist_eval_order(cnb_ctx *c, const cnb_params *q, const cnb_hist_params *hp, CnbHistEval *ev);
int cnb_hist_pick(cnb_ctx *c, const cnb_hist_params *hp, CnbHistEval *ev);   /* after cnb_ctx_search_light on c */
void cnb_hist_add(double *hist, int N, const std::vector<int> &order, const CnbHistEval &ev);
void cnb_gen_permutation(std::vector<int> &out, int n, struct CnbRng &rng);
/* one process, several devices (cnb_multi.cu): devices a one-shot call should use (cnb_params.num_devices, else the
 * CNB_NUM_DEVICES environment variable -- a number or "all" --, clamped to the devices present from p->device on) */
int cnb_multi_devices(const cnb_params *p);
int cnb_multi_search_order(const cnb_params *p, cnb_result **out);
int cnb_multi_search_sa(const cnb_params *p, const cnb_sa_params *sa, cnb_result **out);
int cnb_multi_search_hist(const cnb_params *p, const cnb_hist_params *hp, double *hist, int32_t *iterations);
void cnb_multi_release(void);
/* union of two evaluations: per complexity the higher likelihood, a's on ties; complexities only b has are added */
int cnb_result_union(cnb_result *a, const cnb_result *b);

#endif
