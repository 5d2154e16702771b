/* cnb_device.cuh -- device-side views of the plan and the family enumerator (K3). */
#ifndef CNB_DEVICE_CUH
#define CNB_DEVICE_CUH

#include <cuda_runtime.h>
#include "cnb_internal.h"
#include "cnb_kernels.h"      /* ScoreOut */
#include "cnb_log.h"

/* packed samples + plan tables, all device pointers */
struct DevData {
	int32_t N, M, W;            /* W = ceil(M/32) words per node */
	int32_t Mpad;               /* W*32 */
	const uint4 *planes;        /* [W][N] order space: .x/.y/.z/.w = one-hot bit planes of categories 0..3 */
	const uint32_t *keep;       /* [W][N] order space: bit = sample kept for that child (perturbations), or NULL */
	const uint8_t *codes;       /* [N_label][Mpad] zero-based codes, 255 = NA (label space) */
	const uint32_t *keep_lbl;   /* [W][N_label] or NULL */
	const int32_t *order;       /* [N] position -> label */
	const int32_t *cats;        /* [N] order space */
	const int32_t *lists;
	const CnbBlock *blocks;
	const int32_t *group_blocks;
	const double *edge;         /* [N*N] order space [parent*N+child] or NULL */
	const int32_t *fix1;        /* [N] order space: the node's only fixed parent or -1; NULL when no node has fixed parents */
	CnbTiles tiles;             /* tensor-core tiling of the two-parent group (tile_base NULL if not tiled) */
};

__device__ __forceinline__ unsigned plane_of(const uint4 &v, int i) {
	return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
}

/* one launch class: either a plan group (implicit enumeration) or an explicit family list */
struct DevFamilies {
	int32_t d;                  /* parents per family (explicit lists may override per family) */
	int32_t nblk;
	int32_t blk_off;
	int32_t explicit_list;
	const int32_t *ex_child;    /* [nfam] positions */
	const int32_t *ex_pars;     /* [nfam][CNB_MAXP] positions */
	const int32_t *ex_nd;       /* [nfam] or NULL */
	int64_t nfam;
};

__device__ __forceinline__ long long dev_binom(int n, int k) {
	if (k < 0 || k > n) return 0;
	long long r = 1;
	for (int j = 1; j <= k; j++) r = r * (n - k + j) / j;
	return r;
}

/* lexicographic unrank: the candidate order of combinationSets() (catnet_search2.h:102-147) */
__device__ __forceinline__ void dev_unrank(int n, int k, long long r, int *out) {
	int base = 0;
	for (int t = 0; t < k; t++) {
		int kk = k - t;
		int i;
		if (kk == 1) {
			i = (int)r;
			r = 0;
		} else {
			long long total = dev_binom(n, kk);
			int lo = 0, hi = n - kk;
			while (lo < hi) {
				int mid = (lo + hi + 1) >> 1;
				if (total - dev_binom(n - mid, kk) <= r) lo = mid;
				else hi = mid - 1;
			}
			r -= total - dev_binom(n - lo, kk);
			i = lo;
		}
		out[t] = base + i;
		base += i + 1;
		n -= i + 1;
	}
}

/* family id -> child position and parent positions in table order (fixed parents first, then free ascending;
 * catnet_search2.h:532-553).  Returns the number of parents. */
__device__ __forceinline__ int dev_family(const DevData &D, const DevFamilies &F, long long fid, int &child, int *pars) {
	if (F.explicit_list) {
		child = F.ex_child[fid];
		int d = F.ex_nd ? F.ex_nd[fid] : F.d;
		for (int i = 0; i < d; i++) pars[i] = F.ex_pars[fid * CNB_MAXP + i];
		return d;
	}
	const int32_t *gb = D.group_blocks + F.blk_off;
	int lo = 0, hi = F.nblk - 1;
	while (lo < hi) {
		int mid = (lo + hi + 1) >> 1;
		if (D.blocks[gb[mid]].start <= fid) lo = mid;
		else hi = mid - 1;
	}
	const CnbBlock &b = D.blocks[gb[lo]];
	child = b.child;
	for (int i = 0; i < b.nfix; i++) pars[i] = D.lists[b.fix_off + i];
	int idx[CNB_MAXP];
	int k = b.d - b.nfix;
	dev_unrank(b.nfree, k, fid - b.start, idx);
	for (int i = 0; i < k; i++) pars[b.nfix + i] = D.lists[b.free_off + idx[i]];
	return b.d;
}

/* Bernoulli edge prior of one candidate (catnet_search2.h:584-611): log prod e[p->n] + log prod (1 - e[i->n]) over
 * the non-parents, products taken in exactly the reference's order; -FLT_MAX if either product is 0. */
__device__ __forceinline__ double dev_prior(const DevData &D, int child, const int *pars, int d) {
	const int N = D.N;
	double t = 1.0;
	for (int i = 0; i < d; i++) t = __dmul_rn(t, D.edge[(size_t)pars[i] * N + child]);
	double prior = (t > 0) ? cnb_log(t) : CNB_NEG_INF;
	double t2 = 1.0;
	for (int i = 0; i < N; i++) {
		if (i == child) continue;
		bool isp = false;
		for (int j = 0; j < d; j++) isp |= (pars[j] == i);
		if (!isp) t2 = __dmul_rn(t2, __dsub_rn(1.0, D.edge[(size_t)i * N + child]));
	}
	t2 = (t2 > 0) ? cnb_log(t2) : CNB_NEG_INF;
	if (prior == CNB_NEG_INF || t2 == CNB_NEG_INF) return CNB_NEG_INF;
	return __dadd_rn(prior, t2);
}

/* log-lik of one count table in the reference's exact operation order (problist.h:224-250 then
 * catnet_class.h:867-871): for each parent configuration ascending, pp = 1/N_k, loglik += n*log(n*pp) for n > 0
 * ascending; finally divide by ncount if ncount > 1.  cnt(cfg, k) fetches a cell. */
template <class Fetch>
__device__ __forceinline__ double dev_loglik(int ncfg, int cc, Fetch cnt, int *ncount_out) {
	double ll = 0.0;
	long long ncount = 0;
	for (int cfg = 0; cfg < ncfg; cfg++) {
		long long nk = 0;
		for (int k = 0; k < cc; k++) nk += cnt(cfg, k);
		if (nk <= 0) continue;
		ncount += nk;
		double pp = __ddiv_rn(1.0, (double)nk);
		for (int k = 0; k < cc; k++) {
			int n = cnt(cfg, k);
			if (n > 0) {
				double dn = (double)n;
				ll = __dadd_rn(ll, __dmul_rn(dn, cnb_log(__dmul_rn(dn, pp))));
			}
		}
	}
	if (ncount > 1) ll = __ddiv_rn(ll, (double)ncount);
	*ncount_out = (int)ncount;
	return ll;
}

/* position of family fid of group (chunk, seg_off) in the rank-major score buffer */
__device__ __forceinline__ long long dev_score_index(long long fid, long long chunk, long long seg_off, long long seg_len) {
	long long r = fid / chunk;
	return r * seg_len + seg_off + (fid - r * chunk);
}


/* score (+ prior), log-lik, prior and sample size of family fid into whatever outputs the launch asked for */
__device__ __forceinline__ void dev_store_outputs(const DevData &D, const ScoreOut &O, long long fid, int child,
		const int *pars, int d, double ll, int ncount) {
	double prior = 0.0;
	double score = ll;
	if (D.edge) {
		prior = dev_prior(D, child, pars, d);
		score = __dadd_rn(ll, prior);                     /* fLogLik += priorLogLik (catnet_search2.h:610) */
	}
	if (O.scores) O.scores[dev_score_index(fid, O.chunk, O.seg_off, O.seg_len)] = score;
	if (O.loglik) O.loglik[fid] = ll;
	if (O.prior) O.prior[fid] = prior;
	if (O.ncount) O.ncount[fid] = ncount;
}

/* ---- three-way tables of the data set (label space): triple (u < v < w) at rank C(w,3) + C(v,2) + u, 64 ints,
 * cell [i_u][i_v][i_w] at (i_u * 4 + i_v) * 4 + i_w ---- */
__device__ __forceinline__ long long triple_index(int u, int v, int w) {
	return (long long)w * (w - 1) * (w - 2) / 6 + (long long)v * (v - 1) / 2 + u;
}

/* n(a = i, b = j, c = k) for node labels a, b, c in any order */
__device__ __forceinline__ int triple_count(const int *__restrict__ tt, int a, int i, int b, int j, int c, int k) {
	int t;
	if (a > b) { t = a; a = b; b = t; t = i; i = j; j = t; }
	if (b > c) { t = b; b = c; c = t; t = j; j = k; k = t; }
	if (a > b) { t = a; a = b; b = t; t = i; i = j; j = t; }
	return tt[triple_index(a, b, c) * 64 + (i * 4 + j) * 4 + k];
}


/* Complete the four-way table of family (p0, p1, p2 | child) whose (CM-1)^4 free cells sit in `mine`
 * (cell (i,j,k,l) at ((((i*CM+j)*CM+k)*CM+l) * BS), strided over the threads of the CTA in shared memory): the slices with a
 * last category follow, axis after axis, from the three-way tables of the four triples inside the family (no value is
 * missing, so every margin is complete).  Returns the reference-order log-likelihood (problist.h:224-250). */
template <int CM, int BS>
__device__ __forceinline__ double dev_ie3_finish(int *mine, const DevData &Dt, const int *__restrict__ tt, const int *pars,
		int child, int *ncount) {
	constexpr int F1 = CM - 1;
#define CELL(i, j, k, l) mine[((((i) * CM + (j)) * CM + (k)) * CM + (l)) * BS]
	const int l0 = Dt.order[pars[0]], l1 = Dt.order[pars[1]], l2 = Dt.order[pars[2]], lc = Dt.order[child];
	/* axis of p0: known afterwards for every i, and j, k, l free */
	for (int j = 0; j < F1; j++)
		for (int k = 0; k < F1; k++)
			for (int l = 0; l < F1; l++) {
				int s = 0;
				for (int i = 0; i < F1; i++) s += CELL(i, j, k, l);
				CELL(F1, j, k, l) = triple_count(tt, l1, j, l2, k, lc, l) - s;
			}
	/* axis of p1 */
	for (int i = 0; i < CM; i++)
		for (int k = 0; k < F1; k++)
			for (int l = 0; l < F1; l++) {
				int s = 0;
				for (int j = 0; j < F1; j++) s += CELL(i, j, k, l);
				CELL(i, F1, k, l) = triple_count(tt, l0, i, l2, k, lc, l) - s;
			}
	/* axis of p2 */
	for (int i = 0; i < CM; i++)
		for (int j = 0; j < CM; j++)
			for (int l = 0; l < F1; l++) {
				int s = 0;
				for (int k = 0; k < F1; k++) s += CELL(i, j, k, l);
				CELL(i, j, F1, l) = triple_count(tt, l0, i, l1, j, lc, l) - s;
			}
	/* axis of the child */
	for (int i = 0; i < CM; i++)
		for (int j = 0; j < CM; j++)
			for (int k = 0; k < CM; k++) {
				int s = 0;
				for (int l = 0; l < F1; l++) s += CELL(i, j, k, l);
				CELL(i, j, k, F1) = triple_count(tt, l0, i, l1, j, l2, k) - s;
			}
#undef CELL
	const int pc0 = Dt.cats[pars[0]], pc1 = Dt.cats[pars[1]], pc2 = Dt.cats[pars[2]], cc = Dt.cats[child];
	return dev_loglik(pc0 * pc1 * pc2, cc, [&](int cfg, int l) {
		const int k = cfg % pc2, ij = cfg / pc2, j = ij % pc1, i = ij / pc1;
		return mine[(((i * CM + j) * CM + k) * CM + l) * BS];
	}, ncount);
}

#endif
