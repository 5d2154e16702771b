/* cnb_kernels.h -- host-callable launchers of the kernels in cnb_kernels.cu */
#ifndef CNB_KERNELS_H
#define CNB_KERNELS_H
#include <cuda_runtime.h>
#include "cnb_internal.h"

struct DevData;
struct DevFamilies;

/* where a scoring launch puts its results (any pointer may be NULL) */
struct ScoreOut {
	double *scores;            /* rank-major score buffer, see dev_score_index */
	long long chunk, seg_off, seg_len;
	double *loglik;            /* [nfam] sample-averaged log-lik without prior */
	double *prior;             /* [nfam] */
	int *ncount;               /* [nfam] */
	int *counts;               /* [nfam][counts_stride] reference table layout; must be zeroed by the caller */
	int counts_stride;
};

struct DpState {
	int K;
	unsigned char *exists;     /* [K+1] */
	double *L;                 /* [K+1] network log-lik (re-summed, catnet_class.h:475-485) */
	double *pfx;               /* [K+1] in-order partial sum over the nodes already decided */
};

/* samples: [M][N] of int32 (elem_bytes 4) or of uint8 with 0 = missing (elem_bytes 1, the host-compacted copy) */
int cnbk_minmax(cudaStream_t st, const void *samples, int elem_bytes, int N, int M, int *vmin, int *vmax);
/* [w_begin, w_end): the 32-sample words this launch covers (w_end < 0 = all W) -- a slice of the samples on its way in */
int cnbk_pack(cudaStream_t st, const void *samples, int elem_bytes, const int32_t *pert, int N, int M, int W, const int *vmin,
		const int *cats, uint8_t *codes, uint4 *planes, uint32_t *keep, int *bad, int w_begin = 0, int w_end = -1);
int cnbk_permute(cudaStream_t st, const uint4 *planes_lbl, const uint32_t *keep_lbl, const int *order, int N, int W,
		uint4 *planes, uint32_t *keep, int w_begin = 0, int w_end = -1);
bool cnbk_planes_supported(int d, int max_cats);
int cnbk_score_planes(cudaStream_t st, const DevData &Dt, const DevFamilies &F, int max_cats, long long fid_begin,
		long long fid_end, int slices, const ScoreOut &O);
int cnbk_pair_tables(cudaStream_t st, const uint4 *planes_lbl, int N, int W, int *out);
/* ordered: pair_tab = one table per ORDERED (parent, child) pair, counted under the child's keep mask and without the
 * samples where either value is missing (cnbk_pair_tables_umma_full) */
int cnbk_score_pairs1(cudaStream_t st, const DevData &Dt, const DevFamilies &F, int max_cats, long long b, long long e,
		const ScoreOut &O, const int *pair_tab, bool ordered = false);
int cnbk_score_planes_ie2(cudaStream_t st, const DevData &Dt, const DevFamilies &F, int max_cats, long long b,
		long long e, const ScoreOut &O, const int *pair_tab);
/* three-way tables of all node triples (label space, [C(N,3)][64]) and the three-parent inclusion-exclusion scorer */
int cnbk_triple_tables(cudaStream_t st, const uint4 *planes_lbl, int N, int W, int max_cats, const int *pair_tab, int *out);
int cnbk_score_planes_ie3(cudaStream_t st, const DevData &Dt, const DevFamilies &F, int max_cats, long long b,
		long long e, const ScoreOut &O, const int *triple_tab);
/* tensor-core scorer (cnb_umma.cu): operand bit rows, the tcgen05 table kernel (fp4 = 1: kind::mxf4, 0: kind::i8) */
int cnbk_umma_row_quads(int W, int fp4, int *ngroups);
/* [q_begin, q_end): the quads (4 sample words) this launch writes (q_end < 0 = all of cnbk_umma_row_quads) */
int cnbk_umma_rows(cudaStream_t st, const uint4 *planes, int N, int W, int fp4, uint4 *rows_a, uint4 *rows_b, int q_begin = 0,
		int q_end = -1);
/* operand rows of the full-table tiles: 4 rows per node, or 8 with keep masks (4 plain + 4 under the node's own mask) */
int cnbk_umma_rows_full(cudaStream_t st, const uint4 *planes, const uint32_t *keep, int N, int W, uint4 *rows_a, uint4 *rows_b);
/* local tiles [local_begin, local_end) of rank T.rank (global tile = rank + local * world) */
/* [g_begin, g_end): the sample groups (3 stages x 256 samples with E2M1 operands) this launch contracts (g_end < 0 = all);
 * accumulate: add to the counts already there (a later slice of the samples) instead of storing */
int cnbk_umma_tables(cudaStream_t st, const uint4 *rows_a, const uint4 *rows_b, int N, int W, int fp4,
		const CnbTiles &T, long long local_begin, long long local_end, int *counts, int g_begin = 0, int g_end = -1,
		int accumulate = 0);
/* samples per group of the table kernel */
int cnbk_umma_group_samples(int fp4);
/* pair tables (label space) from P tiles that ran on ORDER-space operand rows: order[pos] = label */
int cnbk_pairs_from_tiles_order(cudaStream_t st, const int *counts, const int *order, int N, int M, int *pair_tab);
long long cnbk_pair_tile_run(cudaStream_t st, const uint4 *rows_a, const uint4 *rows_b, int N, int W, int *counts, int g_begin,
		int g_end, int accumulate);
long long cnbk_umma_column_macs(int W, int fp4);
/* all two-way tables on the tensor cores (no missing values, M < 2^24); scratch sized by the caller */
long long cnbk_pair_tiles(int N);
int cnbk_pair_tables_umma(cudaStream_t st, const uint4 *planes_lbl, int N, int M, int W, uint4 *rows_a, uint4 *rows_b,
		int *counts, int *pair_tab);   /* multiply-accumulates per MMA column of one tile, padding included */
/* the same for data with missing values / perturbation masks: tables of all ORDERED pairs (parent u, child v), label space,
 * [u N + v][16]: samples where both values are present and v is not perturbed (keep_lbl may be NULL); full-table P tiles */
long long cnbk_pair_tiles_full(int N);
int cnbk_pair_tables_umma_full(cudaStream_t st, const uint4 *planes_lbl, const uint32_t *keep_lbl, int N, int W, uint4 *rows_a,
		uint4 *rows_b, int *counts, int *pair_ord);
int cnbk_umma_epilogue(cudaStream_t st, const DevData &Dt, long long local_begin, long long local_end, const int *counts,
		const int *pair_tab, const ScoreOut &O, bool all_four_cats = false);   /* all_four_cats: every node has 4 categories */
/* tables, log-likelihoods and sample sizes of the explicit two-parent families [f_begin, f_end) (positions; ex_pars
 * [..][CNB_MAXP]) read off the resident tile tables of the whole tiled group (world 1, one batch) */
int cnbk_umma_export(cudaStream_t st, const DevData &Dt, long long f_begin, long long f_end, const int *ex_child, const int *ex_pars,
		const int *counts, const int *pair_tab, const ScoreOut &O);
#define CNBK_UMMA_SLOT_INTS 36
#define CNBK_UMMA_FULL_SLOT_INTS 64
/* three-parent families on the tensor cores: pair-side rows incl. the virtual pair nodes, then cnbk_umma_tables with a
 * CnbTiles of kind 3 and cnbk_umma_epilogue3 per column tile */
int cnbk_umma_rows3(cudaStream_t st, const uint4 *planes, int N, int W, int nvirt, int f1, uint4 *rows_a3);
int cnbk_umma_epilogue3(cudaStream_t st, const DevData &Dt, const DevFamilies &F, int max_cats, long long b, long long e,
		int c0, const int *counts, const int *triple_tab, const ScoreOut &O);
/* the same for triples [t_begin, t_end) of one column tile [c0, ce) (batches of its tiles; tiles dealt to several devices) */
int cnbk_umma_epilogue3_range(cudaStream_t st, const DevData &Dt, const DevFamilies &F, int max_cats, long long t_begin, long long t_end,
		long long q_base, int c0, int ce, const int *counts, const int *triple_tab, const ScoreOut &O);
/* pairwise metrics (cnb_pairwise.cu): kind 0 entropy, 1 KL, 2 Pearson from the tables of the families (child n1,
 * parent n2), family index n1 * (N - 1) + rank of n2 among the others; out[n2 * N + n1] */
int cnbk_pairwise_metric(cudaStream_t st, int kind, int N, const int *cats, const int *kept, const int *full,
		const int *diag_tab, int stride, double *out);
void cnb_entropy_order_from_matrix(const double *mat, int N, int32_t *order_out);
/* a fixed network on the device (cnb_network.cu): parents 0-based labels in listed order, tables concatenated by off */
struct CnbDevNet {
	int N, maxp;
	const int *cats, *npar, *pars;   /* [N], [N], [N][maxp] */
	const long long *off;            /* [N + 1] */
	const double *logp;              /* log of every table cell, +inf where the probability is not > 0 */
};
int cnbk_net_logp(cudaStream_t st, const double *probs, long long n, double *logp);
int cnbk_net_nodes(cudaStream_t st, const CnbDevNet &net, const uint8_t *codes, const uint32_t *keep_lbl, int Mpad,
		const int *sel, int nsel, double *out);
int cnbk_net_bysample(cudaStream_t st, const CnbDevNet &net, const uint8_t *codes, const uint32_t *keep_lbl, int M,
		int Mpad, int *first_zero, bool any_zero, double *out);
int cnbk_keep_count(cudaStream_t st, const uint32_t *keep_lbl, int N, int W, int *out);
/* cnSamples: order = sampling order (parents first); net.logp points at the PROBABILITY tables here; pert [M][N] or
 * NULL; cnt / base are [N][chunks] (chunks of cnbk_sample_chunk() samples), only used with perturbations */
int cnbk_sample_chunk(void);
int cnbk_sample_count(cudaStream_t st, const CnbDevNet &net, const int *order, const int *pert, int M, int *cnt);
int cnbk_sample(cudaStream_t st, const CnbDevNet &net, const int *order, const int *pert, const long long *base, int M,
		unsigned long long seed, int *out);
int cnbk_sample_na(cudaStream_t st, int N, int M, int k, unsigned long long seed, unsigned long long base, int *out);
int cnbk_epilogue_counts(cudaStream_t st, const DevData &Dt, const DevFamilies &F, long long b, long long e,
		const ScoreOut &O);
int cnbk_fill_f64(cudaStream_t st, double *p, long long n, double v);
int cnbk_score_hist(cudaStream_t st, const DevData &Dt, const DevFamilies &F, long long b, long long e, int max_cells,
		const ScoreOut &O, int *overflow, int mode);
/* arg-max per (child, parent count) block: `chunks` = (block, chunk) per CTA, cnbk_select_chunk() candidates each;
 * blk_chunk0[nblocks + 1] = first chunk of every block; part_* = nchunks partial results */
int cnbk_select_chunk(void);
int cnbk_select(cudaStream_t st, const DevData &Dt, int nblocks, const CnbGroup *groups, const int *blk_equal,
		const int2 *chunks, int nchunks, const int *blk_chunk0, const double *scores, long long seg_len, double *opt_score,
		int *opt_cx, long long *opt_fid, double *part_score, long long *part_fid);
/* overrides of the score-cache model into the option table (cnb_cache.h) */
int cnbk_patch_options(cudaStream_t st, int n, const long long *opt, const long long *fid, const int *cx, const double *prior,
		const double *ex_ll, const double *ex_score, double *opt_score, int *opt_cx, long long *opt_fid);
int cnbk_dp_init(cudaStream_t st, const DpState &S, int base_complexity, const double *base_v, int N);
int cnbk_dp_node(cudaStream_t st, const DpState &cur, const DpState &nxt, int c, int N, int base_cx_c,
		const double *base_v, long long opt_begin, long long opt_end, const double *opt_score, const int *opt_cx,
		const long long *opt_fid, long long *bp_opt, int *bp_src);
int cnbk_dp_all(cudaStream_t st, const DpState &s0, const DpState &s1, int N, int base_complexity, const CnbDpNode *nodes,
		const double *base_v, const double *opt_score, const int *opt_cx, const long long *opt_fid, long long *bp_opt,
		int *bp_src);
int cnbk_dp_mixed(cudaStream_t st, const DpState &s0, const DpState &s1, int N, int base_complexity, const CnbDpNode *nodes,
		const double *base_v, const double *opt_score, const int *opt_cx, const long long *opt_fid, long long *bp_opt,
		int *bp_src);
/* trapezoid wavefront (T nodes per neighbour exchange); final state in buffer (*steps_out) % 3 */
int cnbk_dp_trap(cudaStream_t st, const DpState &s0, const DpState &s1, const DpState &s2, int N, int base_complexity,
		int halo, const CnbDpNode *nodes, const double *base_v, const double *opt_score, const int *opt_cx,
		const long long *opt_fid, long long *bp_opt, int *bp_src, unsigned *flags, int flags_cap, int *steps_out,
		int compact = 1);   /* compact: the re-summation chains of a CTA's viable candidates run from a dense work list */
int cnbk_dp_wave(cudaStream_t st, const DpState &s0, const DpState &s1, const DpState &s2, int N, int base_complexity,
		int halo, const CnbDpNode *nodes, const double *base_v, const double *opt_score, const int *opt_cx,
		const long long *opt_fid, long long *bp_opt, int *bp_src, unsigned *flags, int flags_cap);
int cnbk_dp_collect(cudaStream_t st, const DpState &fin, int N, const long long *bp_opt, const int *bp_src,
		long long *sel);
int cnbk_dp_collect_list(cudaStream_t st, const DpState &fin, int N, const long long *bp_opt, const int *bp_src,
		const int *slots, int nslots, long long *sel);
int cnbk_device_log(cudaStream_t st, const double *x, long long n, double *out);
#endif
