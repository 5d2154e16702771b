/* cnb_kernels.cu -- hand-written sm_100a kernels of the structure-search scoring path.
 *
 *  K1  k_minmax / k_pack / k_permute   int32 [sample][node] -> uint8 codes + one-hot bit planes (+ keep masks)
 *                                      replaces the per-call re-marshalling of rcatnet_search.cpp:151-174 and the
 *                                      recode of catnet_search2.h:201-257
 *  K2a k_score_planes<D,CM>            family contingency table by AND + POPC over bit planes, one thread per
 *                                      candidate parent set, fused fp64 log-lik epilogue
 *                                      replaces CATNET::setNodeSampleProb + PROB_LIST::loglikelihood
 *                                      (catnet_class.h:818-879, problist.h:224-263)
 *  K2b k_score_hist                    generic path (any category count / table size): one CTA per family,
 *                                      shared-memory histogram
 *  K3  dev_family (cnb_device.cuh)     combinadic unrank of the family id, replaces combinationSets (:102-147)
 *      k_score_planes_ie2 / _ie3       the same on complete data: only the free cells are counted, the rest follows from the
 *                                      two- / three-way tables of the data set (k_pair_tables, k_triple_tables)
 *      k_score_pairs0 / k_score_pairs1 zero- and one-parent families read straight off those tables (ordered variants for
 *                                      data with missing values / perturbation masks)
 *  K4  k_select / k_select_combine     per (node, d) first-maximum / option table, replaces :613-621
 *  K5  k_dp_trap (default)             knapsack over complexity with the reference's exact fp64 comparison semantics,
 *      k_dp_wave / k_dp_all / k_dp_mixed / k_dp_init + k_dp_node, k_dp_collect_list
 *                                      replaces :653-677, :800-827, :847-865: one cooperative launch for all nodes --
 *                                      trapezoid wavefront (several nodes per neighbour exchange, candidates of a CTA
 *                                      re-summed from a dense work list), one-node wavefront,
 *                                      grid barrier per node, a warp per slot for mixed-category nodes -- or a launch per node
 *  The two-parent tables on the tensor cores (tcgen05) are in cnb_umma.cu.
 */
#include <atomic>
#include <stdio.h>
#include <cooperative_groups.h>
#include "cnb_device.cuh"
#include "cnb_kernels.h"
#include "cnb_tiles.h"

#define SCORE_BS 128

/* ------------------------------------------------------------------ K1: pack ---- */
/* T = int32_t (what R hands over) or uint8_t (the host-compacted copy: 0 = missing, 255 = not a category) */
template <class T>
__global__ void k_minmax(const T *__restrict__ samples, int N, int M, int *__restrict__ vmin, int *__restrict__ vmax) {
	int v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= N) return;
	int lo = INT32_MAX, hi = -INT32_MAX;
	for (int j = blockIdx.y; j < M; j += gridDim.y) {
		int x = samples[(size_t)j * N + v];
		if (x == INT32_MIN || x < 1) continue;            /* NA (rcatnet_search.cpp:156) */
		lo = min(lo, x);
		hi = max(hi, x);
	}
	if (lo <= hi) {
		atomicMin(&vmin[v], lo);
		atomicMax(&vmax[v], hi);
	}
}

/* one thread per (32-sample word, node): reads are coalesced across nodes, one uint4 of planes per thread */
template <class T>
__global__ void k_pack(const T *__restrict__ samples, const int32_t *__restrict__ pert, int N, int M, int W,
		const int *__restrict__ vmin, const int *__restrict__ cats, uint8_t *__restrict__ codes,
		uint4 *__restrict__ planes, uint32_t *__restrict__ keep, int *__restrict__ bad, int w_begin) {
	/* bad[0]: a value outside the node's categories; bad[2]: at least one missing value in the matrix */
	/* one-dimensional grid (no 65,535 limit on the sample words), node chunks of a word adjacent */
	const int nchunk = (N + blockDim.x - 1) / blockDim.x;
	int v = (blockIdx.x % nchunk) * blockDim.x + threadIdx.x;
	int w = w_begin + blockIdx.x / nchunk;                 /* the launch covers the sample words from w_begin on */
	if (v >= N) return;
	const int base = vmin[v], C = cats[v];
	bool any_na = false;
	uint32_t pl[4] = {0, 0, 0, 0}, kp = 0;
	uint32_t bytes[8];
#pragma unroll
	for (int q = 0; q < 8; q++) bytes[q] = 0;
#pragma unroll 8
	for (int s = 0; s < 32; s++) {
		int j = w * 32 + s;
		int code = 255;
		if (j < M) {
			int x = samples[(size_t)j * N + v];
			if (!(x == INT32_MIN || x < 1)) {
				code = x - base;
				if (code < 0 || code >= C) {              /* "Incorrect categories" (catnet_search2.h:247-250) */
					*bad = 1;
					code = 255;
				}
			}
			if (!pert || pert[(size_t)j * N + v] == 0) kp |= 1u << s;
			any_na |= (code == 255);
		}
		if (code < 4) pl[code] |= 1u << s;
		bytes[s >> 2] |= (uint32_t)code << ((s & 3) * 8);
	}
	planes[(size_t)w * N + v] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
	if (any_na) bad[2] = 1;
	if (keep) keep[(size_t)w * N + v] = kp;
	uint4 *dst = reinterpret_cast<uint4 *>(codes + (size_t)v * W * 32 + (size_t)w * 32);
	dst[0] = make_uint4(bytes[0], bytes[1], bytes[2], bytes[3]);
	dst[1] = make_uint4(bytes[4], bytes[5], bytes[6], bytes[7]);
}

/* label space -> order space for one search (what the reference does by re-copying the whole int matrix) */
__global__ void k_permute(const uint4 *__restrict__ planes_lbl, const uint32_t *__restrict__ keep_lbl,
		const int *__restrict__ order, int N, int W, uint4 *__restrict__ planes, uint32_t *__restrict__ keep, int w_begin) {
	const int nchunk = (N + blockDim.x - 1) / blockDim.x;
	int p = (blockIdx.x % nchunk) * blockDim.x + threadIdx.x;
	int w = w_begin + blockIdx.x / nchunk;
	if (p >= N) return;
	int l = order[p];
	uint4 v = planes_lbl[(size_t)w * N + l];
	planes[(size_t)w * N + p] = v;
	if (keep) keep[(size_t)w * N + p] = keep_lbl[(size_t)w * N + l];
}

/* ------------------------------------------------------------------ epilogue ---- */

/* ------------------------------------------------------------------ K2a: bit-plane scorer ---- */
template <int D, int CM> struct CellCount { static constexpr int value = CM * CellCount<D - 1, CM>::value; };
template <int CM> struct CellCount<0, CM> { static constexpr int value = CM; };

template <int D>
__device__ __forceinline__ void load_word(const DevData &Dt, int w, int child, const int *pars, uint4 (&v)[D + 1]) {
	const uint4 *row = Dt.planes + (size_t)w * Dt.N;
#pragma unroll
	for (int i = 0; i < D; i++) v[i] = __ldg(row + pars[i]);
	uint4 c = __ldg(row + child);
	if (Dt.keep) {
		uint32_t m = __ldg(Dt.keep + (size_t)w * Dt.N + child);
		c.x &= m; c.y &= m; c.z &= m; c.w &= m;
	}
	v[D] = c;
}

/* counts[((i0*CM+i1)*CM+..)*CM + k] += popc(P0[i0] & P1[i1] & .. & C[k]) : every index is a compile-time constant */
template <int D, int CM>
__device__ __forceinline__ void accumulate(unsigned (&cnt)[CellCount<D, CM>::value], const uint4 (&v)[D + 1]) {
	if constexpr (D == 1) {
#pragma unroll
		for (int i = 0; i < CM; i++)
#pragma unroll
			for (int k = 0; k < CM; k++)
				cnt[i * CM + k] += __popc(plane_of(v[0], i) & plane_of(v[1], k));
	} else if constexpr (D == 2) {
#pragma unroll
		for (int i = 0; i < CM; i++)
#pragma unroll
			for (int j = 0; j < CM; j++) {
				unsigned ab = plane_of(v[0], i) & plane_of(v[1], j);
#pragma unroll
				for (int k = 0; k < CM; k++)
					cnt[(i * CM + j) * CM + k] += __popc(ab & plane_of(v[2], k));
			}
	} else {
		static_assert(D == 3, "bit-plane scorer is instantiated for 1..3 parents");
#pragma unroll
		for (int i = 0; i < CM; i++)
#pragma unroll
			for (int j = 0; j < CM; j++) {
				unsigned ab = plane_of(v[0], i) & plane_of(v[1], j);
#pragma unroll
				for (int l = 0; l < CM; l++) {
					unsigned abc = ab & plane_of(v[2], l);
#pragma unroll
					for (int k = 0; k < CM; k++)
						cnt[((i * CM + j) * CM + l) * CM + k] += __popc(abc & plane_of(v[3], k));
				}
			}
	}
}

/* One thread per candidate family; the word loop is software-pipelined one word ahead.
 * gridDim.y > 1 splits the sample words (few families, many samples): partial tables are then added into
 * O.counts with atomics and k_epilogue_counts finishes.  Otherwise the epilogue runs here, through a
 * conflict-free shared-memory transpose of the per-thread counters. */
template <int D, int CM>
__global__ void __launch_bounds__(SCORE_BS) k_score_planes(DevData Dt, DevFamilies F, long long fid_begin,
		long long fid_end, int words_per_slice, ScoreOut O) {
	constexpr int CELLS = CellCount<D, CM>::value;
	extern __shared__ __align__(16) int s_cnt[];                           /* [CELLS][SCORE_BS] */
	const long long fid = fid_begin + (long long)blockIdx.x * SCORE_BS + threadIdx.x;
	if (fid >= fid_end) return;
	int child, pars[CNB_MAXP];
	dev_family(Dt, F, fid, child, pars);

	unsigned cnt[CELLS];
#pragma unroll
	for (int i = 0; i < CELLS; i++) cnt[i] = 0;

	const int w0 = blockIdx.y * words_per_slice;
	const int w1 = min(Dt.W, w0 + words_per_slice);
	uint4 nxt[D + 1], cur[D + 1];
	if (w0 < w1) load_word<D>(Dt, w0, child, pars, nxt);
	for (int w = w0; w < w1; w++) {
#pragma unroll
		for (int i = 0; i <= D; i++) cur[i] = nxt[i];
		if (w + 1 < w1) load_word<D>(Dt, w + 1, child, pars, nxt);
		accumulate<D, CM>(cnt, cur);
	}

	/* actual category counts: parents (table order) then child */
	int pc[D], cc = Dt.cats[child];
	int ncfg = 1;
#pragma unroll
	for (int i = 0; i < D; i++) {
		pc[i] = Dt.cats[pars[i]];
		ncfg *= pc[i];
	}

	if (O.counts) {
		/* reference layout: ((p0*C1 + p1)*C2 + ..)*Cc + k  (problist.h:143-155, 252-263) */
		int *dst = O.counts + (size_t)fid * O.counts_stride;
#pragma unroll
		for (int s = 0; s < CELLS; s++) {
			int k = s % CM, rest = s / CM, ref = 0, mul = 1;
			bool ok = k < cc;
#pragma unroll
			for (int i = D - 1; i >= 0; i--) {
				int ci = rest % CM;
				rest /= CM;
				ok = ok && ci < pc[i];
				ref += ci * mul;
				mul *= pc[i];
			}
			if (ok && cnt[s]) atomicAdd(dst + ref * cc + k, (int)cnt[s]);
		}
		if (gridDim.y > 1) return;                          /* k_epilogue_counts finishes */
	}
#pragma unroll
	for (int s = 0; s < CELLS; s++) s_cnt[s * SCORE_BS + threadIdx.x] = (int)cnt[s];
	const int *mine = s_cnt + threadIdx.x;
	int ncount;
	double ll = dev_loglik(ncfg, cc, [&](int cfg, int k) {
		/* cfg is the reference's mixed-radix configuration index; map it to the CM-radix register index */
		int rest = cfg, sidx = 0, mul = 1;
#pragma unroll
		for (int i = D - 1; i >= 0; i--) {
			sidx += (rest % pc[i]) * mul;
			rest /= pc[i];
			mul *= CM;
		}
		return mine[(sidx * CM + k) * SCORE_BS];
	}, &ncount);
	dev_store_outputs(Dt, O, fid, child, pars, D, ll, ncount);
}


/* ---- pair tables: n(x = i, y = j) for every unordered pair of nodes (label space, x < y), 16 ints per pair.
 * They are the two-way margins the inclusion-exclusion scorer below completes its tables with; computed once
 * per sample matrix (cnb_ctx_set_samples), valid while no value is missing. */
__device__ __forceinline__ long long pair_index(int x, int y) { return (long long)y * (y - 1) / 2 + x; }   /* x < y */

__global__ void __launch_bounds__(128) k_pair_tables(const uint4 *__restrict__ planes_lbl, int N, int W,
		long long npairs, int *__restrict__ out) {
	long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= npairs) return;
	int y = (int)((1.0 + sqrt(1.0 + 8.0 * (double)p)) * 0.5);
	while ((long long)y * (y - 1) / 2 > p) y--;
	while ((long long)(y + 1) * y / 2 <= p) y++;
	int x = (int)(p - (long long)y * (y - 1) / 2);
	unsigned cnt[16];
#pragma unroll
	for (int i = 0; i < 16; i++) cnt[i] = 0;
	for (int w = 0; w < W; w++) {
		uint4 a = __ldg(planes_lbl + (size_t)w * N + x), b = __ldg(planes_lbl + (size_t)w * N + y);
#pragma unroll
		for (int i = 0; i < 4; i++)
#pragma unroll
			for (int j = 0; j < 4; j++) cnt[i * 4 + j] += __popc(plane_of(a, i) & plane_of(b, j));
	}
#pragma unroll
	for (int i = 0; i < 16; i++) out[p * 16 + i] = (int)cnt[i];
}

/* n(u = i, v = j) from the pair tables, u and v node labels in any order */
__device__ __forceinline__ int pair_count(const int *__restrict__ pt, int u, int v, int i, int j) {
	return u < v ? pt[pair_index(u, v) * 16 + i * 4 + j] : pt[pair_index(v, u) * 16 + j * 4 + i];
}

/* K2a': two-parent scorer with inclusion-exclusion.  With no missing values the (CM-1)^3 "free" cells
 * n(i,j,k), i,j,k < CM-1, determine the table: every cell that involves a last category follows from the
 * two-way margins n(p0,p1), n(p0,c), n(p1,c).  27 AND+POPC per word instead of 64 at CM = 4 (8 vs 27 at CM = 3,
 * 1 vs 8 at CM = 2): the kernel is bound by the POPC issue rate, so this is the speed-up. */
template <int CM>
__global__ void __launch_bounds__(SCORE_BS) k_score_planes_ie2(DevData Dt, DevFamilies F, long long fid_begin,
		long long fid_end, ScoreOut O, const int *__restrict__ pair_tab) {
	constexpr int F1 = CM - 1, FREE = F1 * F1 * F1;
	extern __shared__ __align__(16) int s_cnt[];          /* [CM^3][SCORE_BS] */
	const long long fid = fid_begin + (long long)blockIdx.x * SCORE_BS + threadIdx.x;
	if (fid >= fid_end) return;
	int child, pars[CNB_MAXP];
	dev_family(Dt, F, fid, child, pars);
	unsigned cnt[FREE > 0 ? FREE : 1];
#pragma unroll
	for (int i = 0; i < FREE; i++) cnt[i] = 0;
	uint4 nxt[3], cur[3];
	const int W = Dt.W;
	load_word<2>(Dt, 0, child, pars, nxt);
	for (int w = 0; w < W; w++) {
#pragma unroll
		for (int i = 0; i < 3; i++) cur[i] = nxt[i];
		if (w + 1 < W) load_word<2>(Dt, w + 1, child, pars, nxt);
#pragma unroll
		for (int i = 0; i < F1; i++)
#pragma unroll
			for (int j = 0; j < F1; j++) {
				unsigned ab = plane_of(cur[0], i) & plane_of(cur[1], j);
#pragma unroll
				for (int k = 0; k < F1; k++) cnt[(i * F1 + j) * F1 + k] += __popc(ab & plane_of(cur[2], k));
			}
	}
	/* complete the table in shared memory: cell (i,j,k) at [((i*CM+j)*CM+k)*SCORE_BS + tid] */
	int *mine = s_cnt + threadIdx.x;
#define CELL(i, j, k) mine[(((i) * CM + (j)) * CM + (k)) * SCORE_BS]
#pragma unroll
	for (int i = 0; i < F1; i++)
#pragma unroll
		for (int j = 0; j < F1; j++)
#pragma unroll
			for (int k = 0; k < F1; k++) CELL(i, j, k) = (int)cnt[(i * F1 + j) * F1 + k];
	const int l0 = Dt.order[pars[0]], l1 = Dt.order[pars[1]], lc = Dt.order[child];
	for (int i = 0; i < F1; i++)                          /* last child category */
		for (int j = 0; j < F1; j++) {
			int s = 0;
			for (int k = 0; k < F1; k++) s += CELL(i, j, k);
			CELL(i, j, F1) = pair_count(pair_tab, l0, l1, i, j) - s;
		}
	for (int i = 0; i < F1; i++)                          /* last category of the second parent */
		for (int k = 0; k < CM; k++) {
			int s = 0;
			for (int j = 0; j < F1; j++) s += CELL(i, j, k);
			CELL(i, F1, k) = pair_count(pair_tab, l0, lc, i, k) - s;
		}
	for (int j = 0; j < CM; j++)                          /* last category of the first parent */
		for (int k = 0; k < CM; k++) {
			int s = 0;
			for (int i = 0; i < F1; i++) s += CELL(i, j, k);
			CELL(F1, j, k) = pair_count(pair_tab, l1, lc, j, k) - s;
		}
#undef CELL
	const int pc0 = Dt.cats[pars[0]], pc1 = Dt.cats[pars[1]], cc = Dt.cats[child];
	int ncount;
	double ll = dev_loglik(pc0 * pc1, cc, [&](int cfg, int k) {
		int i = cfg / pc1, j = cfg - i * pc1;
		return mine[((i * CM + j) * CM + k) * SCORE_BS];
	}, &ncount);
	dev_store_outputs(Dt, O, fid, child, pars, 2, ll, ncount);
}

/* ---- three-way tables of all node triples (label space, once per data set) and the three-parent scorer built on them ----
 * triple (u < v < w) lives at rank C(w,3) + C(v,2) + u, 64 ints: cell [i_u][i_v][i_w] at (i_u * 4 + i_v) * 4 + i_w. */
/* one thread per triple: (CM-1)^3 free cells by POPC, the rest from the pair tables (as k_score_planes_ie2) */
template <int CM>
__global__ void __launch_bounds__(128) k_triple_tables(const uint4 *__restrict__ planes_lbl, int N, int W, long long ntriples,
		const int *__restrict__ pair_tab, int *__restrict__ out) {
	constexpr int F1 = CM - 1;
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= ntriples) return;
	/* unrank: largest w with C(w,3) <= t, then v with C(v,2) <= rest */
	int w = (int)cbrt(6.0 * (double)t) + 2;
	while ((long long)w * (w - 1) * (w - 2) / 6 > t) w--;
	while ((long long)(w + 1) * w * (w - 1) / 6 <= t) w++;
	long long r = t - (long long)w * (w - 1) * (w - 2) / 6;
	int v = (int)((1.0 + sqrt(1.0 + 8.0 * (double)r)) / 2.0);
	while ((long long)v * (v - 1) / 2 > r) v--;
	while ((long long)(v + 1) * v / 2 <= r) v++;
	const int u = (int)(r - (long long)v * (v - 1) / 2);
	unsigned cnt[F1 * F1 * F1 > 0 ? F1 * F1 * F1 : 1];
#pragma unroll
	for (int i = 0; i < F1 * F1 * F1; i++) cnt[i] = 0;
	for (int q = 0; q < W; q++) {
		const uint4 a = __ldg(planes_lbl + (size_t)q * N + u), b = __ldg(planes_lbl + (size_t)q * N + v),
				c = __ldg(planes_lbl + (size_t)q * N + w);
#pragma unroll
		for (int i = 0; i < F1; i++)
#pragma unroll
			for (int j = 0; j < F1; j++) {
				const unsigned ab = plane_of(a, i) & plane_of(b, j);
#pragma unroll
				for (int k = 0; k < F1; k++) cnt[(i * F1 + j) * F1 + k] += __popc(ab & plane_of(c, k));
			}
	}
	int tab[64];
#pragma unroll
	for (int i = 0; i < 64; i++) tab[i] = 0;
#define CELL(i, j, k) tab[((i) * 4 + (j)) * 4 + (k)]
#pragma unroll
	for (int i = 0; i < F1; i++)
#pragma unroll
		for (int j = 0; j < F1; j++)
#pragma unroll
			for (int k = 0; k < F1; k++) CELL(i, j, k) = (int)cnt[(i * F1 + j) * F1 + k];
#pragma unroll
	for (int i = 0; i < F1; i++)
#pragma unroll
		for (int j = 0; j < F1; j++) {
			int s = 0;
#pragma unroll
			for (int k = 0; k < F1; k++) s += CELL(i, j, k);
			CELL(i, j, F1) = pair_count(pair_tab, u, v, i, j) - s;
		}
#pragma unroll
	for (int i = 0; i < F1; i++)
#pragma unroll
		for (int k = 0; k < CM; k++) {
			int s = 0;
#pragma unroll
			for (int j = 0; j < F1; j++) s += CELL(i, j, k);
			CELL(i, F1, k) = pair_count(pair_tab, u, w, i, k) - s;
		}
#pragma unroll
	for (int j = 0; j < CM; j++)
#pragma unroll
		for (int k = 0; k < CM; k++) {
			int s = 0;
#pragma unroll
			for (int i = 0; i < F1; i++) s += CELL(i, j, k);
			CELL(F1, j, k) = pair_count(pair_tab, v, w, j, k) - s;
		}
#undef CELL
	int4 *dst = reinterpret_cast<int4 *>(out + t * 64);
#pragma unroll
	for (int i = 0; i < 16; i++) dst[i] = make_int4(tab[4 * i], tab[4 * i + 1], tab[4 * i + 2], tab[4 * i + 3]);
}

/* K2a'': three-parent scorer with inclusion-exclusion.  Only the (CM-1)^4 free cells are counted over the samples
 * (16 AND+POPC per word instead of 81 at CM = 3, 81 instead of 256 at CM = 4); the slices with a last category follow,
 * axis after axis, from the complete three-way tables of the four triples inside {p0, p1, p2, child}. */
template <int CM, int BS>
__global__ void __launch_bounds__(BS) k_score_planes_ie3(DevData Dt, DevFamilies F, long long fid_begin, long long fid_end,
		ScoreOut O, const int *__restrict__ tt) {
	constexpr int F1 = CM - 1, FREE = F1 * F1 * F1 * F1;
	extern __shared__ __align__(16) int s_cnt[];          /* [CM^4][BS] */
	const long long fid = fid_begin + (long long)blockIdx.x * BS + threadIdx.x;
	if (fid >= fid_end) return;
	int child, pars[CNB_MAXP];
	dev_family(Dt, F, fid, child, pars);
	unsigned cnt[FREE > 0 ? FREE : 1];
#pragma unroll
	for (int i = 0; i < FREE; i++) cnt[i] = 0;
	uint4 nxt[4], cur[4];
	const int W = Dt.W;
	load_word<3>(Dt, 0, child, pars, nxt);
	for (int w = 0; w < W; w++) {
#pragma unroll
		for (int i = 0; i < 4; i++) cur[i] = nxt[i];
		if (w + 1 < W) load_word<3>(Dt, w + 1, child, pars, nxt);
#pragma unroll
		for (int i = 0; i < F1; i++)
#pragma unroll
			for (int j = 0; j < F1; j++) {
				const unsigned ab = plane_of(cur[0], i) & plane_of(cur[1], j);
#pragma unroll
				for (int k = 0; k < F1; k++) {
					const unsigned abc = ab & plane_of(cur[2], k);
#pragma unroll
					for (int l = 0; l < F1; l++) cnt[((i * F1 + j) * F1 + k) * F1 + l] += __popc(abc & plane_of(cur[3], l));
				}
			}
	}
	int *mine = s_cnt + threadIdx.x;
#define CELL(i, j, k, l) mine[((((i) * CM + (j)) * CM + (k)) * CM + (l)) * BS]
#pragma unroll
	for (int i = 0; i < F1; i++)
#pragma unroll
		for (int j = 0; j < F1; j++)
#pragma unroll
			for (int k = 0; k < F1; k++)
#pragma unroll
				for (int l = 0; l < F1; l++) CELL(i, j, k, l) = (int)cnt[((i * F1 + j) * F1 + k) * F1 + l];
#undef CELL
	int ncount;
	const double ll = dev_ie3_finish<CM, BS>(mine, Dt, tt, pars, child, &ncount);
	dev_store_outputs(Dt, O, fid, child, pars, 3, ll, ncount);
}

/* Single-parent families when no value is missing: the family table IS the two-way table of (parent, child) that
 * k_pair_tables counted once per data set -- no pass over the samples at all, only the log-lik epilogue
 * (problist.h:224-250) and, when asked for, the table in the reference's layout. */
/* ORDERED: pair_tab holds one table per ORDERED pair (parent u, child v) at (u N + v) 16 + i 4 + k -- the counts over the
 * samples where both values are present and v is not perturbed (data with missing values / perturbation masks; tables
 * from the full-table P tiles, cnbk_pair_tables_umma_full) */
template <bool ORDERED>
__global__ void __launch_bounds__(128) k_score_pairs1(DevData Dt, DevFamilies F, long long fid_begin, long long fid_end,
		ScoreOut O, const int *__restrict__ pair_tab) {
	const long long fid = fid_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (fid >= fid_end) return;
	int child, pars[CNB_MAXP];
	dev_family(Dt, F, fid, child, pars);
	const int lp = Dt.order[pars[0]], lc = Dt.order[child];
	const int pc = Dt.cats[pars[0]], cc = Dt.cats[child];
	const int *ot = pair_tab + ((long long)lp * Dt.N + lc) * 16;
	auto count = [&](int i, int k) { return ORDERED ? ot[i * 4 + k] : pair_count(pair_tab, lp, lc, i, k); };
	int ncount;
	double ll = dev_loglik(pc, cc, count, &ncount);
	if (O.counts) {
		int *tab = O.counts + (size_t)fid * O.counts_stride;
		for (int i = 0; i < pc; i++)
			for (int k = 0; k < cc; k++) tab[i * cc + k] = count(i, k);
	}
	dev_store_outputs(Dt, O, fid, child, pars, 1, ll, ncount);
}

/* Families without parents on clean data: the marginal counts are row sums of any two-way table of the node. */
template <bool ORDERED>
__global__ void __launch_bounds__(128) k_score_pairs0(DevData Dt, DevFamilies F, long long fid_begin, long long fid_end,
		ScoreOut O, const int *__restrict__ pair_tab) {
	const long long fid = fid_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (fid >= fid_end) return;
	int child, pars[CNB_MAXP];
	dev_family(Dt, F, fid, child, pars);
	const int lc = Dt.order[child], lo = lc == 0 ? 1 : 0, cc = Dt.cats[child];
	int marg[CNB_PLANE_CATS];
	for (int k = 0; k < CNB_PLANE_CATS; k++) {
		int s = 0;
		if (ORDERED) s = pair_tab[((long long)lc * Dt.N + lc) * 16 + k * 4 + k];   /* the node against itself, under its own mask */
		else for (int i = 0; i < CNB_PLANE_CATS; i++) s += pair_count(pair_tab, lo, lc, i, k);
		marg[k] = s;
	}
	int ncount;
	double ll = dev_loglik(1, cc, [&](int, int k) { return marg[k]; }, &ncount);
	if (O.counts) {
		int *tab = O.counts + (size_t)fid * O.counts_stride;
		for (int k = 0; k < cc; k++) tab[k] = marg[k];
	}
	dev_store_outputs(Dt, O, fid, child, pars, 0, ll, ncount);
}

/* epilogue over tables already in global memory (split-sample mode): one thread per family */
__global__ void k_epilogue_counts(DevData Dt, DevFamilies F, long long fid_begin, long long fid_end, ScoreOut O) {
	const long long fid = fid_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (fid >= fid_end) return;
	int child, pars[CNB_MAXP];
	int d = dev_family(Dt, F, fid, child, pars);
	int cc = Dt.cats[child], ncfg = 1;
	for (int i = 0; i < d; i++) ncfg *= Dt.cats[pars[i]];
	const int *tab = O.counts + (size_t)fid * O.counts_stride;
	int ncount;
	double ll = dev_loglik(ncfg, cc, [&](int cfg, int k) { return tab[cfg * cc + k]; }, &ncount);
	dev_store_outputs(Dt, O, fid, child, pars, d, ll, ncount);
}

/* ------------------------------------------------------------------ K2b: generic histogram scorer ---- */
/* One CTA per family, table in shared memory (ints) followed by the per-cell fp64 terms.  Handles any number of
 * categories (<= 254) and parents (<= CNB_MAXP) as long as the table fits (tsize <= hist_max_cells). */
__global__ void __launch_bounds__(256) k_score_hist(DevData Dt, DevFamilies F, long long fid_begin, long long fid_end,
		int max_cells, ScoreOut O, int *__restrict__ overflow) {
	extern __shared__ __align__(16) int s_tab[];
	__shared__ int sh_child, sh_d, sh_pars[CNB_MAXP], sh_pc[CNB_MAXP], sh_cc, sh_tsize;
	double *terms = reinterpret_cast<double *>(s_tab + ((max_cells + 1) & ~1));
	for (long long fid = fid_begin + blockIdx.x; fid < fid_end; fid += gridDim.x) {
		__syncthreads();
		if (threadIdx.x == 0) {
			int child, pars[CNB_MAXP];
			int d = dev_family(Dt, F, fid, child, pars);
			long long ts = Dt.cats[child];
			for (int i = 0; i < d; i++) {
				sh_pars[i] = pars[i];
				sh_pc[i] = Dt.cats[pars[i]];
				ts *= sh_pc[i];
				if (ts > max_cells) ts = (long long)max_cells + 1;
			}
			sh_child = child;
			sh_d = d;
			sh_cc = Dt.cats[child];
			sh_tsize = (int)ts;
		}
		__syncthreads();
		const int d = sh_d, cc = sh_cc, tsize = sh_tsize, child = sh_child;
		if (tsize > max_cells) {
			if (threadIdx.x == 0) *overflow = 1;
			continue;
		}
		for (int i = threadIdx.x; i < tsize; i += blockDim.x) s_tab[i] = 0;
		__syncthreads();
		const uint8_t *ccol = Dt.codes + (size_t)Dt.order[child] * Dt.Mpad;
		const uint8_t *pcol[CNB_MAXP];
		for (int i = 0; i < d; i++) pcol[i] = Dt.codes + (size_t)Dt.order[sh_pars[i]] * Dt.Mpad;
		const uint32_t *kcol = Dt.keep_lbl ? Dt.keep_lbl + Dt.order[child] : nullptr;
		/* four samples per thread per step: 32-bit loads of the uint8 columns */
		for (int s4 = threadIdx.x; s4 < Dt.Mpad / 4; s4 += blockDim.x) {
			uint32_t cv = reinterpret_cast<const uint32_t *>(ccol)[s4];
			uint32_t pv[CNB_MAXP];
			for (int i = 0; i < d; i++) pv[i] = reinterpret_cast<const uint32_t *>(pcol[i])[s4];
			uint32_t km = 0xF;
			if (kcol) km = (kcol[(size_t)(s4 >> 3) * Dt.N] >> ((s4 & 7) * 4)) & 0xF;
#pragma unroll
			for (int q = 0; q < 4; q++) {
				int x = (cv >> (8 * q)) & 255;
				bool ok = x < cc && ((km >> q) & 1);
				int idx = 0;
				for (int i = 0; i < d; i++) {
					int pvq = (pv[i] >> (8 * q)) & 255;
					ok = ok && pvq < sh_pc[i];
					idx = idx * sh_pc[i] + pvq;
				}
				if (ok) atomicAdd(&s_tab[idx * cc + x], 1);
			}
		}
		__syncthreads();
		/* per-cell terms in parallel, then one thread adds them in the reference's order */
		const int ncfg = tsize / cc;
		for (int cell = threadIdx.x; cell < tsize; cell += blockDim.x) {
			int cfg = cell / cc;
			long long nk = 0;
			for (int k = 0; k < cc; k++) nk += s_tab[cfg * cc + k];
			int n = s_tab[cell];
			double t = 0.0;
			if (n > 0) {
				double pp = __ddiv_rn(1.0, (double)nk);
				t = __dmul_rn((double)n, cnb_log(__dmul_rn((double)n, pp)));
			}
			terms[cell] = t;
		}
		if (O.counts)
			for (int cell = threadIdx.x; cell < tsize; cell += blockDim.x)
				O.counts[(size_t)fid * O.counts_stride + cell] = s_tab[cell];
		__syncthreads();
		if (threadIdx.x == 0) {
			double ll = 0.0;
			long long ncount = 0;
			for (int cfg = 0; cfg < ncfg; cfg++)
				for (int k = 0; k < cc; k++) {
					int n = s_tab[cfg * cc + k];
					ncount += n;
					if (n > 0) ll = __dadd_rn(ll, terms[cfg * cc + k]);
				}
			if (ncount > 1) ll = __ddiv_rn(ll, (double)ncount);
			int pars[CNB_MAXP];
			for (int i = 0; i < d; i++) pars[i] = sh_pars[i];
			dev_store_outputs(Dt, O, fid, child, pars, d, ll, (int)ncount);
		}
	}
}

/* ------------------------------------------------------------------ K2c: warp-privatised histogram scorer ---- */
/* The scorer for nodes with more than four categories (no bit planes, no tensor tiles): one WARP per family, its
 * conditional table in the warp's own part of shared memory, so that no two warps ever add to the same cell.  Every lane
 * takes 16 consecutive samples per step -- one 128-bit load of the child's and of every parent's uint8 column (the
 * column-major code matrix stays in L2: N x M bytes) -- and turns each into the mixed-radix cell index
 * (catnet_class.h:842-857: parent configuration x child category; a missing value or a sample perturbed for the child is
 * skipped, :849).  MATCH: lanes that hit the same cell in a step are found with __match_any_sync and the first of
 * them adds their number with a plain read-modify-write -- no atomics at all, which is what pays on skewed data where
 * most samples fall into a few cells; otherwise one shared-memory atomicAdd per sample (the atomics of a warp on one
 * address are serialised by the hardware).  The fp64 epilogue (problist.h:224-250) stays in the warp: the 32 lanes
 * compute the terms n log(n / N_k) of 32 consecutive cells side by side, then the terms are added in the reference's
 * order -- cell after cell, empty cells skipped -- by a chain of shuffles. */
/* 0xFF bytes of w -> 0x80 in that byte, else 0 (a code of 255 = missing value or padding behind the last sample) */
__device__ __forceinline__ uint32_t hist_ff_bytes(uint32_t w) { return ((w & 0x7F7F7F7Fu) + 0x01010101u) & w & 0x80808080u; }

/* 16 samples of one lane: the cell indices of TWO samples at a time in the halves of a 32-bit word (tables have fewer
 * than 65,536 cells, so (p0 * C1 + p1) * Cc + x never carries from the low half into the high one), one ATOMS.POPC.INC per
 * sample.  CHECK: samples with a missing value in the family (code 255) or perturbed for the child (keep bit clear) are
 * skipped -- their bytes are cleared first, so that they cannot carry into the neighbouring sample's half. */
template <int D, bool CHECK>
__device__ __forceinline__ void hist_chunk16(int *tab, const uint4 cv, const uint4 *pv, const int *pc, int cc, uint32_t km) {
	const uint32_t cwv[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
	for (int w = 0; w < 4; w++) {
		uint32_t cw = cwv[w], pw[D > 0 ? D : 1], inv = 0;
#pragma unroll
		for (int i = 0; i < D; i++) pw[i] = w == 0 ? pv[i].x : (w == 1 ? pv[i].y : (w == 2 ? pv[i].z : pv[i].w));
		if (CHECK) {
			inv = hist_ff_bytes(cw);
#pragma unroll
			for (int i = 0; i < D; i++) inv |= hist_ff_bytes(pw[i]);
			const uint32_t drop = ~(km >> (4 * w)) & 0xFu;             /* keep bit j of the word -> bit 7 of byte j */
			inv |= ((drop * 0x00204081u) & 0x01010101u) << 7;
			const uint32_t m = (inv >> 7) * 0xFFu;
			cw &= ~m;
#pragma unroll
			for (int i = 0; i < D; i++) pw[i] &= ~m;
		}
		uint32_t lo, hi;                                              /* samples (0, 1) and (2, 3) of the word */
		if (D == 0) {
			lo = __byte_perm(cw, 0, 0x4140);
			hi = __byte_perm(cw, 0, 0x4342);
		} else {
			lo = __byte_perm(pw[0], 0, 0x4140);
			hi = __byte_perm(pw[0], 0, 0x4342);
#pragma unroll
			for (int i = 1; i < D; i++) {
				lo = lo * (uint32_t)pc[i] + __byte_perm(pw[i], 0, 0x4140);
				hi = hi * (uint32_t)pc[i] + __byte_perm(pw[i], 0, 0x4342);
			}
			lo = lo * (uint32_t)cc + __byte_perm(cw, 0, 0x4140);
			hi = hi * (uint32_t)cc + __byte_perm(cw, 0, 0x4342);
		}
		if (!CHECK || !(inv & 0x00000080u)) atomicAdd(&tab[lo & 0xFFFFu], 1);
		if (!CHECK || !(inv & 0x00008000u)) atomicAdd(&tab[lo >> 16], 1);
		if (!CHECK || !(inv & 0x00800000u)) atomicAdd(&tab[hi & 0xFFFFu], 1);
		if (!CHECK || !(inv & 0x80000000u)) atomicAdd(&tab[hi >> 16], 1);
	}
}

/* MODE 0: shared atomics, every sample checked; 1: shared atomics on data without missing values or masks (only the chunk
 * that holds the padding behind the last sample is checked); 2: __match_any_sync aggregation */
template <int D, int MODE>
__global__ void __launch_bounds__(256) k_score_hist_warp(DevData Dt, DevFamilies F, long long fid_begin, long long fid_end,
		int cells_per_warp, ScoreOut O, int *__restrict__ overflow) {
	constexpr bool MATCH = MODE == 2;
	extern __shared__ __align__(16) int s_tab[];
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
	int *tab = s_tab + (size_t)wib * cells_per_warp;
	const unsigned FULLM = 0xFFFFFFFFu;
	const int chunks16 = Dt.Mpad >> 4;                  /* Mpad is a multiple of 32 */
	const int whole16 = Dt.M >> 4;                      /* chunks without padding */
	for (long long fid = fid_begin + (long long)blockIdx.x * wpb + wib; fid < fid_end; fid += (long long)gridDim.x * wpb) {
		int child, pars[CNB_MAXP];
		dev_family(Dt, F, fid, child, pars);            /* every lane: uniform, no divergence */
		const int cc = Dt.cats[child];
		int pc[D > 0 ? D : 1];
		int tsize = cc;
#pragma unroll
		for (int i = 0; i < D; i++) { pc[i] = Dt.cats[pars[i]]; tsize = tsize > cells_per_warp ? tsize : tsize * pc[i]; }
		if (tsize > cells_per_warp) {                   /* uniform over the warp */
			if (lane == 0) *overflow = 1;
			continue;
		}
		for (int i = lane; i < tsize; i += 32) tab[i] = 0;
		__syncwarp();
		const int clbl = Dt.order[child];
		const uint4 *ccol = reinterpret_cast<const uint4 *>(Dt.codes + (size_t)clbl * Dt.Mpad);
		const uint4 *pcol[D > 0 ? D : 1];
#pragma unroll
		for (int i = 0; i < D; i++) pcol[i] = reinterpret_cast<const uint4 *>(Dt.codes + (size_t)Dt.order[pars[i]] * Dt.Mpad);
		const uint32_t *kcol = Dt.keep_lbl ? Dt.keep_lbl + clbl : nullptr;
		for (int t0 = 0; t0 < chunks16; t0 += 32) {
			const int t = t0 + lane;
			const bool in = t < chunks16;
			uint4 cv = make_uint4(~0u, ~0u, ~0u, ~0u), pv[D > 0 ? D : 1];
			uint32_t km = 0xFFFFu;
			if (in) {
				cv = __ldg(ccol + t);
#pragma unroll
				for (int i = 0; i < D; i++) pv[i] = __ldg(pcol[i] + t);
				if (kcol) km = (__ldg(kcol + (size_t)(t >> 1) * Dt.N) >> ((t & 1) * 16)) & 0xFFFFu;
			}
			if constexpr (!MATCH) {
				if (in) {
					if (MODE == 1 && t < whole16) hist_chunk16<D, false>(tab, cv, pv, pc, cc, km);
					else hist_chunk16<D, true>(tab, cv, pv, pc, cc, km);
				}
			} else {
#pragma unroll
			for (int q = 0; q < 16; q++) {
				const uint32_t cw = q < 4 ? cv.x : (q < 8 ? cv.y : (q < 12 ? cv.z : cv.w));
				const int x = (cw >> (8 * (q & 3))) & 255;
				bool ok = in && x < cc && ((km >> q) & 1);
				int idx = 0;
#pragma unroll
				for (int i = 0; i < D; i++) {
					const uint32_t pw = q < 4 ? pv[i].x : (q < 8 ? pv[i].y : (q < 12 ? pv[i].z : pv[i].w));
					const int v = (pw >> (8 * (q & 3))) & 255;
					ok = ok && v < pc[i];
					idx = idx * pc[i] + v;
				}
				idx = idx * cc + x;
				/* lanes without a sample get keys of their own, so that they meet nobody */
				const unsigned peers = __match_any_sync(FULLM, ok ? idx : (int)(0x80000000u | lane));
				if (ok && lane == __ffs(peers) - 1) tab[idx] += __popc(peers);
				__syncwarp();
			}
			}
		}
		__syncwarp();
		double ll = 0.0;
		long long ncount = 0;
		for (int c0 = 0; c0 < tsize; c0 += 32) {
			const int cell = c0 + lane;
			int n = 0;
			double term = 0.0;
			if (cell < tsize) {
				n = tab[cell];
				if (O.counts) O.counts[(size_t)fid * O.counts_stride + cell] = n;
				if (n > 0) {
					const int cfg = cell / cc;
					long long nk = 0;
					for (int k = 0; k < cc; k++) nk += tab[cfg * cc + k];
					const double pp = __ddiv_rn(1.0, (double)nk), dn = (double)n;
					term = __dmul_rn(dn, cnb_log(__dmul_rn(dn, pp)));
				}
			}
			int tot = n;
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(FULLM, tot, o);
			ncount += tot;
			unsigned m = __ballot_sync(FULLM, n > 0);
			while (m) {                                    /* the reference's order: cells ascending, empty ones skipped */
				const int l = __ffs(m) - 1;
				m &= m - 1;
				ll = __dadd_rn(ll, __shfl_sync(FULLM, term, l));
			}
		}
		if (ncount > 1) ll = __ddiv_rn(ll, (double)ncount);
		if (lane == 0) dev_store_outputs(Dt, O, fid, child, pars, D, ll, (int)ncount);
		__syncwarp();
	}
}

/* ------------------------------------------------------------------ K4: selection / option table ---- */
/* One CTA per candidate block.  Equal-categories path: first maximum in enumeration order (strict '<' running
 * max from -FLT_MAX, catnet_search2.h:559,613) -> one DP option.  Otherwise every candidate becomes an option. */
#define SEL_CHUNK 4096        /* candidates per CTA of k_select: large blocks are cut so that the machine is filled evenly */
__global__ void __launch_bounds__(256) k_select(DevData Dt, const CnbGroup *__restrict__ groups,
		const int *__restrict__ blk_equal, const int2 *__restrict__ chunks, const double *__restrict__ scores,
		long long seg_len, double *__restrict__ opt_score, int *__restrict__ opt_cx, long long *__restrict__ opt_fid,
		double *__restrict__ part_score, long long *__restrict__ part_fid) {
	const int blk = chunks[blockIdx.x].x;
	const CnbBlock &b = Dt.blocks[blk];
	const CnbGroup &g = groups[b.group];
	const long long r_begin = (long long)chunks[blockIdx.x].y * SEL_CHUNK;
	const long long r_end = r_begin + SEL_CHUNK < b.ncomb ? r_begin + SEL_CHUNK : b.ncomb;
	DevFamilies F;
	F.d = b.d; F.nblk = g.nblk; F.blk_off = g.blk_off; F.explicit_list = 0; F.nfam = g.nfam;
	F.ex_child = nullptr; F.ex_pars = nullptr; F.ex_nd = nullptr;
	auto complexity_of = [&](long long fid) {
		int child, pars[CNB_MAXP];
		int d = dev_family(Dt, F, fid, child, pars);
		int cx = Dt.cats[child] - 1;
		for (int i = 0; i < d; i++) cx *= Dt.cats[pars[i]];
		return cx;
	};
	/* where the score of candidate `fid` lives: by family id, or by tile slot for the tensor-core group */
	auto score_of = [&](long long fid) {
		long long slot = fid;
		if (g.tiled) {
			int ab[2];
			if (g.tiled == 1) dev_unrank(b.nfree, 2, fid - b.start, ab);      /* unrestricted block: positions themselves */
			else {
				/* a subset of the unrestricted group (pools, fixed parents): the candidate's two parents, ascending */
				int ch, pars[CNB_MAXP];
				dev_family(Dt, F, fid, ch, pars);
				ab[0] = min(pars[0], pars[1]);
				ab[1] = max(pars[0], pars[1]);
			}
			long long tile;
			int within;
			cnb_family_slot(Dt.tiles, Dt.N, ab[0], ab[1], b.child, tile, within);
			/* tile t is scored by rank t % world as its local tile t / world */
			return scores[(tile % g.tile_world) * seg_len + g.seg_off + (tile / g.tile_world) * (Dt.tiles.tp * Dt.tiles.tc)
					+ within];
		}
		return scores[dev_score_index(slot, g.chunk, g.seg_off, seg_len)];
	};
	if (!blk_equal[blk]) {
		for (long long r = r_begin + threadIdx.x; r < r_end; r += blockDim.x) {
			long long fid = b.start + r;
			opt_score[b.opt_off + r] = score_of(fid);
			opt_cx[b.opt_off + r] = complexity_of(fid);
			opt_fid[b.opt_off + r] = fid;
		}
		return;
	}
	double best = CNB_NEG_INF;
	long long best_fid = -1;
	for (long long r = r_begin + threadIdx.x; r < r_end; r += blockDim.x) {
		long long fid = b.start + r;
		double s = score_of(fid);
		if (s > best) { best = s; best_fid = fid; }       /* ascending r per thread: first max kept */
	}
	__shared__ double sh_s[256];
	__shared__ long long sh_f[256];
	sh_s[threadIdx.x] = best;
	sh_f[threadIdx.x] = best_fid;
	__syncthreads();
	for (int off = 128; off > 0; off >>= 1) {
		if (threadIdx.x < off) {
			double s2 = sh_s[threadIdx.x + off];
			long long f2 = sh_f[threadIdx.x + off];
			double s1 = sh_s[threadIdx.x];
			long long f1 = sh_f[threadIdx.x];
			if (f2 >= 0 && (f1 < 0 || s2 > s1 || (s2 == s1 && f2 < f1))) {
				sh_s[threadIdx.x] = s2;
				sh_f[threadIdx.x] = f2;
			}
		}
		__syncthreads();
	}
	if (threadIdx.x == 0) {
		part_score[blockIdx.x] = sh_s[0];
		part_fid[blockIdx.x] = sh_f[0];
	}
}

/* one thread per equal-category block: its chunks in enumeration order, strict '>' keeps the first maximum */
__global__ void k_select_combine(DevData Dt, const CnbGroup *__restrict__ groups, const int *__restrict__ blk_equal,
		const int *__restrict__ blk_chunk0, int nblocks, const double *__restrict__ part_score,
		const long long *__restrict__ part_fid, double *__restrict__ opt_score, int *__restrict__ opt_cx,
		long long *__restrict__ opt_fid) {
	const int blk = blockIdx.x * blockDim.x + threadIdx.x;
	if (blk >= nblocks || !blk_equal[blk]) return;
	const CnbBlock &b = Dt.blocks[blk];
	const CnbGroup &g = groups[b.group];
	double best = CNB_NEG_INF;
	long long f = -1;
	for (int ch = blk_chunk0[blk]; ch < blk_chunk0[blk + 1]; ch++)
		if (part_fid[ch] >= 0 && (f < 0 || part_score[ch] > best)) { best = part_score[ch]; f = part_fid[ch]; }
	int cx = 0;
	if (f >= 0) {
		DevFamilies F;
		F.d = b.d; F.nblk = g.nblk; F.blk_off = g.blk_off; F.explicit_list = 0; F.nfam = g.nfam;
		F.ex_child = nullptr; F.ex_pars = nullptr; F.ex_nd = nullptr;
		int child, pars[CNB_MAXP];
		int d = dev_family(Dt, F, f, child, pars);
		cx = Dt.cats[child] - 1;
		for (int i = 0; i < d; i++) cx *= Dt.cats[pars[i]];
	}
	opt_score[b.opt_off] = best;
	opt_fid[b.opt_off] = f;
	opt_cx[b.opt_off] = cx;
}

/* The reference's score cache as seen by the option table (cnb_cache.h): candidate i sits in option slot opt[i]; its
 * log-likelihood ex_ll[i] was re-scored with the parents in the cached listing sequence; its term is that plus the prior
 * the cache stored with the entry (prior == null: no edge priors, the explicit scorer's score as it is).  fid[i] >= 0
 * (equal-categories path): the slot's winner becomes the cached parent set. */
__global__ void k_patch_options(int n, const long long *__restrict__ opt, const long long *__restrict__ fid,
		const int *__restrict__ cx, const double *__restrict__ prior, const double *__restrict__ ex_ll,
		const double *__restrict__ ex_score, double *__restrict__ opt_score, int *__restrict__ opt_cx,
		long long *__restrict__ opt_fid) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const long long o = opt[i];
	opt_score[o] = prior ? __dadd_rn(ex_ll[i], prior[i]) : ex_score[i];      /* fLogLik += priorLogLik (catnet_search2.h:610) */
	opt_cx[o] = cx[i];
	if (fid[i] >= 0) opt_fid[o] = fid[i];
}

/* ------------------------------------------------------------------ K5: DP over complexity ---- */
/* loglik() of a network = sum over nodes in index order of (node loglik + prior), re-summed from zero
 * (catnet_class.h:469-485).  pfx = that sum over nodes < c, f = the candidate's term for node c. */
__device__ __forceinline__ double dev_resum(double pfx, double f, int c, int N, const double *__restrict__ base_v) {
	double acc = __dadd_rn(pfx, f);
	for (int i = c + 1; i < N; i++) acc = __dadd_rn(acc, base_v[i]);
	return acc;
}

__global__ void k_dp_init(DpState S, int base_complexity, const double *__restrict__ base_v, int N) {
	int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j > S.K) return;
	bool ex = (j == base_complexity);
	double l = 0.0;
	if (ex) for (int i = 0; i < N; i++) l = __dadd_rn(l, base_v[i]);
	S.exists[j] = ex;
	S.L[j] = l;
	S.pfx[j] = ex ? __dadd_rn(0.0, base_v[0]) : 0.0;
}

/* one launch per node c, one thread per target complexity slot j (the reference loops options outermost and
 * source slots innermost; per target slot that is a scan over options in order, each with ONE source slot). */
__global__ void __launch_bounds__(128) k_dp_node(DpState cur, DpState nxt, int c, int N, int base_cx_c,
		const double *__restrict__ base_v, long long opt_begin, long long opt_end,
		const double *__restrict__ opt_score, const int *__restrict__ opt_cx, const long long *__restrict__ opt_fid,
		long long *__restrict__ bp_opt, int *__restrict__ bp_src) {
	int j = blockIdx.x * blockDim.x + threadIdx.x;
	const int K = cur.K;
	if (j > K) return;
	const double bv = base_v[c];
	bool has = false;
	double R = CNB_NEG_INF;
	long long bo = -1;
	int bs = -1;
	for (long long o = opt_begin; o < opt_end; o++) {
		if (opt_fid[o] < 0) continue;                       /* no winner for that d (:637-640) */
		int k = j + base_cx_c - opt_cx[o];
		if (k < 0 || k > K || !cur.exists[k]) continue;
		double f = opt_score[o];
		double t = __dadd_rn(__dsub_rn(cur.L[k], bv), f);  /* :665-666 */
		if (!has && t > CNB_NEG_INF) { has = true; R = CNB_NEG_INF; }   /* empty net: loglik() = -FLT_MAX */
		if (has && R < t) {                                  /* :670-671 strict */
			bo = o;
			bs = k;
			R = dev_resum(cur.pfx[k], f, c, N, base_v);
		}
	}
	bool take = false;
	if (cur.exists[j]) take = has && bo >= 0 && cur.L[j] < R;   /* :847-859 */
	else take = has && bo >= 0;
	size_t bpi = (size_t)c * (K + 1) + j;
	if (take) {
		nxt.exists[j] = 1;
		nxt.L[j] = R;
		nxt.pfx[j] = __dadd_rn(cur.pfx[bs], opt_score[bo]);
		bp_opt[bpi] = bo;
		bp_src[bpi] = bs;
	} else {
		nxt.exists[j] = cur.exists[j];
		nxt.L[j] = cur.L[j];
		nxt.pfx[j] = __dadd_rn(cur.pfx[j], bv);
		bp_opt[bpi] = -1;
		bp_src[bpi] = j;
	}
}

/* One target slot j of node c (shared by the two fused DP kernels).  CG = loads of the previous state bypass L1
 * (another CTA wrote them during this launch). */
template <bool CG> __device__ __forceinline__ double dp_ld(const double *p) { return CG ? __ldcg(p) : *p; }
template <bool CG> __device__ __forceinline__ unsigned char dp_ld(const unsigned char *p) { return CG ? __ldcg(p) : *p; }

template <bool CG>
__device__ __forceinline__ void dp_slot(const DpState &cur, const DpState &nxt, int j, int c, int N, int K,
		const CnbDpNode &nd, int nopt, const long long (&o_fid)[4], const int (&o_cx)[4], const double (&o_sc)[4],
		const double *sbase, double bv, const double *__restrict__ opt_score, const int *__restrict__ opt_cx,
		const long long *__restrict__ opt_fid, long long *__restrict__ bp_opt, int *__restrict__ bp_src, bool write_bp = true) {
	bool has = false;
	double R = CNB_NEG_INF;
	long long bo = -1;
	int bs = -1;
	double bf = 0.0;
	if (nopt <= 4) {
		/* few options (one per parent count on the equal-categories path): the re-summations, chains of
		 * N - c dependent fp64 adds, run side by side for the viable ones (compacted to the front, as many
		 * chains as the busiest lane of the warp needs) before the sequential decision */
		double acc[4] = {0.0, 0.0, 0.0, 0.0}, t[4] = {0.0, 0.0, 0.0, 0.0}, fs[4] = {0.0, 0.0, 0.0, 0.0};
		int ks[4] = {0, 0, 0, 0}, qs[4] = {0, 0, 0, 0}, nq = 0;
#pragma unroll
		for (int q = 0; q < 4; q++) {
			if (q >= nopt || o_fid[q] < 0) continue;            /* no winner for that d (:637-640) */
			const int k = j + nd.base_cx - o_cx[q];
			if (k < 0 || k > K || !dp_ld<CG>(cur.exists + k)) continue;
			const double f = o_sc[q];
			ks[nq] = k;
			qs[nq] = q;
			fs[nq] = f;
			t[nq] = __dadd_rn(__dsub_rn(dp_ld<CG>(cur.L + k), bv), f);      /* :665-666 */
			acc[nq] = __dadd_rn(dp_ld<CG>(cur.pfx + k), f);
			nq++;
		}
		const int nqmax = __reduce_max_sync(__activemask(), nq);
		/* the terms are fetched 8 ahead of the dependent adds (shared-memory latency off the chain) */
#define DP_CHAIN(NQ) { \
			int i = c + 1; \
			for (; i + 8 <= N; i += 8) { \
				double b[8]; \
				_Pragma("unroll") for (int u = 0; u < 8; u++) b[u] = sbase[i + u]; \
				_Pragma("unroll") for (int u = 0; u < 8; u++) \
					_Pragma("unroll") for (int q = 0; q < NQ; q++) acc[q] = __dadd_rn(acc[q], b[u]); \
			} \
			for (; i < N; i++) { \
				const double b = sbase[i]; \
				_Pragma("unroll") for (int q = 0; q < NQ; q++) acc[q] = __dadd_rn(acc[q], b); \
			} }
		if (nqmax == 1) DP_CHAIN(1)
		else if (nqmax == 2) DP_CHAIN(2)
		else if (nqmax > 2) DP_CHAIN(4)
#undef DP_CHAIN
#pragma unroll
		for (int q = 0; q < 4; q++) {
			if (q >= nq) continue;
			if (!has && t[q] > CNB_NEG_INF) { has = true; R = CNB_NEG_INF; }   /* empty net: loglik() = -FLT_MAX */
			if (has && R < t[q]) { bo = nd.opt_begin + qs[q]; bs = ks[q]; R = acc[q]; bf = fs[q]; }   /* :670-671 strict */
		}
	} else
		for (long long o = nd.opt_begin; o < nd.opt_end; o++) {
			if (opt_fid[o] < 0) continue;                       /* no winner for that d (:637-640) */
			int k = j + nd.base_cx - opt_cx[o];
			if (k < 0 || k > K || !dp_ld<CG>(cur.exists + k)) continue;
			double f = opt_score[o];
			double t = __dadd_rn(__dsub_rn(dp_ld<CG>(cur.L + k), bv), f);  /* :665-666 */
			if (!has && t > CNB_NEG_INF) { has = true; R = CNB_NEG_INF; }   /* empty net: loglik() = -FLT_MAX */
			if (has && R < t) {                                  /* :670-671 strict */
				bo = o;
				bs = k;
				R = dev_resum(dp_ld<CG>(cur.pfx + k), f, c, N, sbase);
				bf = f;
			}
		}
	const bool cur_ex = dp_ld<CG>(cur.exists + j) != 0;
	bool take = false;
	if (cur_ex) take = has && bo >= 0 && dp_ld<CG>(cur.L + j) < R;   /* :847-859 */
	else take = has && bo >= 0;
	size_t bpi = (size_t)c * (K + 1) + j;
	if (take) {
		nxt.exists[j] = 1;
		nxt.L[j] = R;
		nxt.pfx[j] = __dadd_rn(dp_ld<CG>(cur.pfx + bs), bf);
		if (write_bp) { bp_opt[bpi] = bo; bp_src[bpi] = bs; }
	} else {
		nxt.exists[j] = cur_ex;
		nxt.L[j] = dp_ld<CG>(cur.L + j);
		nxt.pfx[j] = __dadd_rn(dp_ld<CG>(cur.pfx + j), bv);
		if (write_bp) { bp_opt[bpi] = -1; bp_src[bpi] = j; }
	}
}

/* the node's arguments and its first four options do not depend on the DP state: both fused kernels fetch them one
 * node ahead, behind the barrier */
#define DP_PREFETCH(node) { \
		_Pragma("unroll") for (int q = 0; q < 4; q++) { \
			const long long o = (node).opt_begin + q; \
			const bool in = o < (node).opt_end; \
			pf_fid[q] = in ? opt_fid[o] : -1; \
			pf_cx[q] = in ? opt_cx[o] : 0; \
			pf_sc[q] = in ? opt_score[o] : 0.0; } }

/* The same recurrence for every node in ONE cooperative launch (grid barrier between nodes) instead of N - 1
 * launches: at C5 the 499 launches cost 7 ms, which is a quarter of an 8-GPU step. */
__global__ void __launch_bounds__(256) k_dp_all(DpState s0, DpState s1, int N, int base_complexity,
		const CnbDpNode *__restrict__ nodes, const double *__restrict__ base_v, const double *__restrict__ opt_score,
		const int *__restrict__ opt_cx, const long long *__restrict__ opt_fid, long long *__restrict__ bp_opt,
		int *__restrict__ bp_src) {
	cooperative_groups::grid_group grid = cooperative_groups::this_grid();
	extern __shared__ double sbase[];                    /* base_v: every re-summation walks it */
	const int K = s0.K, nthr = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
	for (int i = threadIdx.x; i < N; i += blockDim.x) sbase[i] = base_v[i];
	__syncthreads();
	for (int j = t0; j <= K; j += nthr) {
		bool ex = (j == base_complexity);
		double l = 0.0;
		if (ex) for (int i = 0; i < N; i++) l = __dadd_rn(l, base_v[i]);
		s0.exists[j] = ex;
		s0.L[j] = l;
		s0.pfx[j] = ex ? __dadd_rn(0.0, base_v[0]) : 0.0;
	}
	CnbDpNode nd_n = nodes[N > 1 ? 1 : 0];
	long long pf_fid[4];
	int pf_cx[4];
	double pf_sc[4];
	DP_PREFETCH(nd_n)
	grid.sync();
	for (int c = 1; c < N; c++) {
		const DpState &cur = (c & 1) ? s0 : s1, &nxt = (c & 1) ? s1 : s0;
		const CnbDpNode nd = nd_n;
		long long o_fid[4];
		int o_cx[4];
		double o_sc[4];
#pragma unroll
		for (int q = 0; q < 4; q++) { o_fid[q] = pf_fid[q]; o_cx[q] = pf_cx[q]; o_sc[q] = pf_sc[q]; }
		if (c + 1 < N) {
			nd_n = nodes[c + 1];
			DP_PREFETCH(nd_n)
		}
		const int nopt = (int)min((long long)5, (long long)(nd.opt_end - nd.opt_begin));
		for (int j = t0; j <= K; j += nthr)
			dp_slot<false>(cur, nxt, j, c, N, K, nd, nopt, o_fid, o_cx, o_sc, sbase, sbase[c], opt_score, opt_cx, opt_fid, bp_opt, bp_src);
		grid.sync();
	}
}

/* Nodes whose pool has unequal category counts hand EVERY candidate to the DP (catnet_search2.h:491-496, :800-827):
 * hundreds of options per node on ALARM.  The per-slot scan "R < t ? take" is sequential in the option order, but
 * its state only changes at record-breaking options, so a WARP takes a slot: the lanes evaluate 32 options at a time
 * (t = delta-updated log-lik), a ballot finds the first option that beats the running R, only that one is re-summed
 * (the reference's loglik() re-summation, a chain of N - c fp64 adds), and the ballot is repeated for the lanes behind
 * it.  One cooperative launch, grid barrier per node.  Identical decisions to the sequential scan. */
#define DPM_OPTS 2048          /* options of a node staged in shared memory (score + complexity) */
#define DPM_SLOTS 4096         /* DP slots staged in shared memory (log-lik + exists) */
/* ONE: the whole DP in one CTA of BS threads (a plain launch, block barriers); an experiment that did not pay, see
 * cnbk_dp_mixed */
template <int BS, bool ONE>
__global__ void __launch_bounds__(BS) k_dp_mixed(DpState s0, DpState s1, int N, int base_complexity,
		const CnbDpNode *__restrict__ nodes, const double *__restrict__ base_v, const double *__restrict__ opt_score,
		const int *__restrict__ opt_cx, const long long *__restrict__ opt_fid, long long *__restrict__ bp_opt,
		int *__restrict__ bp_src, int stage_state) {
	auto barrier = [] { if constexpr (ONE) __syncthreads(); else cooperative_groups::this_grid().sync(); };
	/* sbase[N] | s_sc[2][DPM_OPTS] | s_L[slots] | s_pfx[slots] | s_cx[2][DPM_OPTS] | s_ex[slots]: the scan of a slot reads a
	 * node's options and the previous state 32 at a time, and every record-breaking candidate starts its re-summation from
	 * the source slot's prefix sum; out of global memory every such read is an L2 round trip on the critical path of the
	 * node step.  The options of node c + 1 do not depend on the DP state: they are staged into the other buffer BEFORE the
	 * barrier that ends node c, so that after it only the state is fetched. */
	extern __shared__ double sbase[];
	const int K = s0.K, nthr = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
	const int lane = threadIdx.x & 31, gw = t0 >> 5, nwarps = nthr >> 5;
	const int nslots = stage_state ? K + 1 : 0;
	double *s_sc2 = sbase + N, *s_L = s_sc2 + 2 * DPM_OPTS, *s_pfx = s_L + nslots;
	int *s_cx2 = reinterpret_cast<int *>(s_pfx + nslots);
	unsigned char *s_ex = reinterpret_cast<unsigned char *>(s_cx2 + 2 * DPM_OPTS);
	/* a node's options into buffer (node & 1): complexity INT_MIN = no winner for that d */
	auto stage_options = [&](int node) {
		const CnbDpNode q = nodes[node];
		const long long n = q.opt_end - q.opt_begin;
		if (n > DPM_OPTS) return;
		double *sc = s_sc2 + (node & 1) * DPM_OPTS;
		int *cx = s_cx2 + (node & 1) * DPM_OPTS;
		for (int o = threadIdx.x; o < (int)n; o += blockDim.x) {
			sc[o] = opt_score[q.opt_begin + o];
			cx[o] = opt_fid[q.opt_begin + o] >= 0 ? opt_cx[q.opt_begin + o] : INT_MIN;
		}
	};
	for (int i = threadIdx.x; i < N; i += blockDim.x) sbase[i] = base_v[i];
	__syncthreads();
	for (int j = t0; j <= K; j += nthr) {
		bool ex = (j == base_complexity);
		double l = 0.0;
		if (ex) for (int i = 0; i < N; i++) l = __dadd_rn(l, sbase[i]);
		s0.exists[j] = ex;
		s0.L[j] = l;
		s0.pfx[j] = ex ? __dadd_rn(0.0, sbase[0]) : 0.0;
	}
	if (N > 1) stage_options(1);
	barrier();
	for (int c = 1; c < N; c++) {
		const DpState &cur = (c & 1) ? s0 : s1, &nxt = (c & 1) ? s1 : s0;
		const CnbDpNode nd = nodes[c];
		const double bv = sbase[c];
		const long long nopt = nd.opt_end - nd.opt_begin;
		const bool opts_staged = nopt <= DPM_OPTS;
		const double *s_sc = s_sc2 + (c & 1) * DPM_OPTS;
		const int *s_cx = s_cx2 + (c & 1) * DPM_OPTS;
		/* the previous state */
		for (int k = threadIdx.x; k < nslots; k += blockDim.x) {
			s_L[k] = cur.L[k];
			s_pfx[k] = cur.pfx[k];
			s_ex[k] = cur.exists[k];
		}
		__syncthreads();
		const double *Lp = stage_state ? s_L : cur.L;
		const double *Pp = stage_state ? s_pfx : cur.pfx;
		const unsigned char *Ep = stage_state ? s_ex : cur.exists;
		for (int j = gw; j <= K; j += nwarps) {
			bool has = false;
			double R = CNB_NEG_INF, bf = 0.0;
			long long bo = -1;
			int bs = -1;
			for (long long o0 = 0; o0 < nopt; o0 += 32) {
				const long long o = nd.opt_begin + o0 + lane;
				bool viable = false;
				double t = 0.0, f = 0.0;
				int k = 0;
				if (o0 + lane < nopt) {
					const int cx = opts_staged ? s_cx[o0 + lane] : (opt_fid[o] >= 0 ? opt_cx[o] : INT_MIN);   /* :637-640 */
					if (cx != INT_MIN) {
						k = j + nd.base_cx - cx;
						if (k >= 0 && k <= K && Ep[k]) {
							viable = true;
							f = opts_staged ? s_sc[o0 + lane] : opt_score[o];
							t = __dadd_rn(__dsub_rn(Lp[k], bv), f);      /* :665-666 */
						}
					}
				}
				unsigned behind = 0xffffffffu;                         /* lanes not yet passed by the scan */
				for (;;) {
					/* empty net: loglik() = -FLT_MAX, the first finite t replaces it; then strict '<' (:670-671) */
					const bool cand = viable && ((behind >> lane) & 1u) && (has ? R < t : t > CNB_NEG_INF);
					const unsigned m = __ballot_sync(0xffffffffu, cand);
					if (!m) break;
					const int l = __ffs(m) - 1;
					double r = 0.0;
					if (lane == l) r = dev_resum(Pp[k], f, c, N, sbase);
					R = __shfl_sync(0xffffffffu, r, l);
					bf = __shfl_sync(0xffffffffu, f, l);
					bs = __shfl_sync(0xffffffffu, k, l);
					bo = nd.opt_begin + o0 + l;
					has = true;
					behind = l == 31 ? 0u : (0xffffffffu << (l + 1));
				}
			}
			if (lane == 0) {
				const bool cur_ex = Ep[j] != 0;
				bool take = false;
				if (cur_ex) take = has && bo >= 0 && Lp[j] < R;   /* :847-859 */
				else take = has && bo >= 0;
				size_t bpi = (size_t)c * (K + 1) + j;
				if (take) {
					nxt.exists[j] = 1;
					nxt.L[j] = R;
					nxt.pfx[j] = __dadd_rn(Pp[bs], bf);
					bp_opt[bpi] = bo;
					bp_src[bpi] = bs;
				} else {
					nxt.exists[j] = cur_ex;
					nxt.L[j] = Lp[j];
					nxt.pfx[j] = __dadd_rn(Pp[j], bv);
					bp_opt[bpi] = -1;
					bp_src[bpi] = j;
				}
			}
		}
		if (c + 1 < N) stage_options(c + 1);                       /* behind this node's scan, ahead of the barrier */
		barrier();
	}
}

/* Wavefront version.  An option only ever looks LEFT, at slot j - (option complexity - base complexity), and by at
 * most `halo` <= 256 slots: a CTA that owns 256 consecutive slots needs, for node c, its own slots and its left
 * neighbour's of node c - 1.  So instead of a grid barrier per node every CTA waits for its left neighbour's progress
 * counter (flags[b] = steps finished; step 0 = initialisation, step c = node c) and the CTAs run as a diagonal
 * wavefront.  Three state buffers: step c writes S[c % 3], which the RIGHT neighbour last read during its step
 * c - 2, hence the second, rarely blocking wait.  All CTAs are co-resident (cooperative launch). */
__device__ __forceinline__ unsigned dp_flag_load(const unsigned *p) {
	unsigned v;
	asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__global__ void __launch_bounds__(256) k_dp_wave(DpState s0, DpState s1, DpState s2, int N, int base_complexity,
		const CnbDpNode *__restrict__ nodes, const double *__restrict__ base_v, const double *__restrict__ opt_score,
		const int *__restrict__ opt_cx, const long long *__restrict__ opt_fid, long long *__restrict__ bp_opt,
		int *__restrict__ bp_src, unsigned *flags) {
	extern __shared__ double sbase[];
	const int K = s0.K, b = blockIdx.x, nb = gridDim.x, j = b * blockDim.x + threadIdx.x;
	const bool mine = j <= K;
	for (int i = threadIdx.x; i < N; i += blockDim.x) sbase[i] = base_v[i];
	__syncthreads();
	if (mine) {
		bool ex = (j == base_complexity);
		double l = 0.0;
		if (ex) for (int i = 0; i < N; i++) l = __dadd_rn(l, sbase[i]);
		s0.exists[j] = ex;
		s0.L[j] = l;
		s0.pfx[j] = ex ? __dadd_rn(0.0, sbase[0]) : 0.0;
	}
	CnbDpNode nd_n = nodes[N > 1 ? 1 : 0];
	long long pf_fid[4];
	int pf_cx[4];
	double pf_sc[4];
	DP_PREFETCH(nd_n)
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();
		atomicExch(&flags[b], 1u);
	}
	for (int c = 1; c < N; c++) {
		const DpState &cur = c % 3 == 1 ? s0 : (c % 3 == 2 ? s1 : s2), &nxt = c % 3 == 1 ? s1 : (c % 3 == 2 ? s2 : s0);
		const CnbDpNode nd = nd_n;
		long long o_fid[4];
		int o_cx[4];
		double o_sc[4];
#pragma unroll
		for (int q = 0; q < 4; q++) { o_fid[q] = pf_fid[q]; o_cx[q] = pf_cx[q]; o_sc[q] = pf_sc[q]; }
		if (c + 1 < N) {
			nd_n = nodes[c + 1];
			DP_PREFETCH(nd_n)
		}
		if (threadIdx.x == 0) {
			if (b > 0) while (dp_flag_load(&flags[b - 1]) < (unsigned)c) { }
			if (b + 1 < nb) while (dp_flag_load(&flags[b + 1]) + 1 < (unsigned)c) { }
		}
		__syncthreads();
		const int nopt = (int)min((long long)5, (long long)(nd.opt_end - nd.opt_begin));
		if (mine)
			dp_slot<true>(cur, nxt, j, c, N, K, nd, nopt, o_fid, o_cx, o_sc, sbase, sbase[c], opt_score, opt_cx, opt_fid, bp_opt, bp_src);
		__syncthreads();
		if (threadIdx.x == 0) {
			__threadfence();
			atomicExch(&flags[b], (unsigned)c + 1u);
		}
	}
}
/* Trapezoid version of the wavefront: a CTA owns 256 consecutive slots and ALSO recomputes the 256 slots to their left,
 * in shared memory, so that it can take T nodes per exchange with its neighbours instead of one: after t nodes the
 * recomputed slots are still right from the (t * halo)-th on, and T * halo <= 256 keeps the owned ones right.  The
 * flag / state round trips through L2 (what bounds k_dp_wave: 499 + 94 steps of ~6 us at C5) are paid once per T
 * nodes; the state lives in shared memory in between.  Back-pointers are written by the owner only.  Same three
 * global state buffers and neighbour protocol as k_dp_wave, per macro-step. */
#define TRAP_TMAX 16          /* nodes per neighbour exchange at most */
/* blockDim.x = owned slots + TRAP_S recomputed slots to their left (the launcher uses 256 + 512: with the work list the
 * recomputed slots cost a gather and a decision each, not a re-summation chain, so a wide left margin -- more nodes per
 * exchange -- is cheap) */
__global__ void __launch_bounds__(1024) k_dp_trap(DpState s0, DpState s1, DpState s2, int N, int base_complexity,
		const CnbDpNode *__restrict__ nodes, const double *__restrict__ base_v, const double *__restrict__ opt_score,
		const int *__restrict__ opt_cx, const long long *__restrict__ opt_fid, long long *__restrict__ bp_opt,
		int *__restrict__ bp_src, unsigned *flags, int T, int halo, int compact, int TRAP_S) {
	extern __shared__ double sbase[];                    /* [N], then the two state buffers, then the work list */
	__shared__ int s_nwork;
	__shared__ CnbDpNode s_nd[TRAP_TMAX];
	__shared__ long long s_fid[TRAP_TMAX][4];
	__shared__ int s_cx[TRAP_TMAX][4];
	__shared__ double s_sc[TRAP_TMAX][4];
	const int BW = blockDim.x, SO = BW - TRAP_S;            /* buffer width, owned slots */
	double *s_L = sbase + N, *s_pfx = s_L + 2 * BW, *s_work = s_pfx + 2 * BW;   /* work list: 4 candidates per thread at most */
	unsigned char *s_ex = reinterpret_cast<unsigned char *>(s_work + 4 * BW);
	const int K = s0.K, b = blockIdx.x, nb = gridDim.x, tid = threadIdx.x;
	const int lo = b * SO - TRAP_S, j = lo + tid;            /* threads [0, 256) recompute the left neighbour's slots */
	const bool owned = tid >= TRAP_S, in_range = j >= 0 && j <= K;
	for (int i = tid; i < N; i += blockDim.x) sbase[i] = base_v[i];
	__syncthreads();
	if (owned && in_range) {
		bool ex = (j == base_complexity);
		double l = 0.0;
		if (ex) for (int i = 0; i < N; i++) l = __dadd_rn(l, sbase[i]);
		s0.exists[j] = ex;
		s0.L[j] = l;
		s0.pfx[j] = ex ? __dadd_rn(0.0, sbase[0]) : 0.0;
	}
	__syncthreads();
	if (tid == 0) {
		__threadfence();
		atomicExch(&flags[b], 1u);
	}
	const int nsteps = (N - 1 + T - 1) / T;
	/* A node raises a network's complexity by at most `halo`, so after macro-step m no slot above base_complexity + m T halo
	 * exists: a CTA whose owned slots all lie above that has nothing to do in steps 1..m_skip -- it announces them as done and
	 * starts at m_skip + 1 (its slots read "no network" from the zeroed state buffers).  Without this the CTAs run in lockstep
	 * down a diagonal, nsteps + CTAs macro-steps long; with it the DP ends after about nsteps of them (C5: 194 -> ~100). */
	int m_skip = 0;
	if (b * SO > base_complexity) m_skip = min(nsteps, (b * SO - base_complexity - 1) / (T * halo));
	if (m_skip > 0) {
		__syncthreads();
		if (tid == 0) {
			__threadfence();
			atomicExch(&flags[b], (unsigned)m_skip + 1u);
		}
	}
	for (int m = m_skip + 1; m <= nsteps; m++) {
		const DpState &gin = (m - 1) % 3 == 0 ? s0 : ((m - 1) % 3 == 1 ? s1 : s2), &gout = m % 3 == 0 ? s0 : (m % 3 == 1 ? s1 : s2);
		const int c_lo = (m - 1) * T + 1, nT = min(T, N - c_lo);
		if (tid == 0) {
			if (b > 0) while (dp_flag_load(&flags[b - 1]) < (unsigned)m) { }
			if (b + 1 < nb) while (dp_flag_load(&flags[b + 1]) + 1 < (unsigned)m) { }
		}
		if (tid >= 32 && tid < 32 + nT) {                    /* the nodes of this macro-step and their options */
			const int t = tid - 32;
			const CnbDpNode nd = nodes[c_lo + t];
			s_nd[t] = nd;
#pragma unroll
			for (int q = 0; q < 4; q++) {
				const long long o = nd.opt_begin + q;
				const bool in = o < nd.opt_end;
				s_fid[t][q] = in ? opt_fid[o] : -1;
				s_cx[t][q] = in ? opt_cx[o] : 0;
				s_sc[t][q] = in ? opt_score[o] : 0.0;
			}
		}
		__syncthreads();
		s_ex[tid] = in_range ? __ldcg(gin.exists + j) : 0;
		s_L[tid] = in_range ? __ldcg(gin.L + j) : 0.0;
		s_pfx[tid] = in_range ? __ldcg(gin.pfx + j) : 0.0;
		__syncthreads();
		for (int t = 0; t < nT; t++) {
			const int c = c_lo + t, rb = (t & 1) * BW, wb = ((t + 1) & 1) * BW;
			DpState cur, nxt;
			cur.K = nxt.K = K;
			cur.exists = s_ex + rb - lo; cur.L = s_L + rb - lo; cur.pfx = s_pfx + rb - lo;
			nxt.exists = s_ex + wb - lo; nxt.L = s_L + wb - lo; nxt.pfx = s_pfx + wb - lo;
			const bool act = in_range && j >= lo + (t + 1) * halo;
			if (!compact) {
				if (act) {
					const CnbDpNode nd = s_nd[t];
					long long o_fid[4];
					int o_cx[4];
					double o_sc[4];
#pragma unroll
					for (int q = 0; q < 4; q++) { o_fid[q] = s_fid[t][q]; o_cx[q] = s_cx[t][q]; o_sc[q] = s_sc[t][q]; }
					const int nopt = (int)min((long long)5, (long long)(nd.opt_end - nd.opt_begin));
					dp_slot<false>(cur, nxt, j, c, N, K, nd, nopt, o_fid, o_cx, o_sc, sbase, sbase[c], opt_score, opt_cx, opt_fid,
							bp_opt, bp_src, owned);
				}
				__syncthreads();
				continue;
			}
			/* Only one slot in ten holds a network (C5: 2,489 of 24,001), so a warp's 32 slots have a handful of viable
			 * candidates -- and every one of them needs the reference's re-summation, a chain of N - c dependent fp64
			 * additions, which a warp pays for in full however few of its lanes take part (the DP was bound by the fp64 pipe:
			 * 512 threads x up to 4 chains per node and CTA).  So the candidates of the whole CTA are first compacted into a work
			 * list, the chains run on as few full warps as that takes, and each slot then decides on the re-summed values. */
			if (tid == 0) s_nwork = 0;
			__syncthreads();
			const CnbDpNode nd = s_nd[t];
			const double bv = sbase[c];
			int nq = 0, ks[4], qs[4], w0 = 0;
			double tq[4], fs[4];
			if (act) {
				const int nopt = (int)min((long long)4, (long long)(nd.opt_end - nd.opt_begin));
				double xs[4];
#pragma unroll
				for (int q = 0; q < 4; q++) {
					if (q >= nopt || s_fid[t][q] < 0) continue;           /* no winner for that d (:637-640) */
					const int k = j + nd.base_cx - s_cx[t][q];
					if (k < 0 || k > K || !cur.exists[k]) continue;
					const double f = s_sc[t][q];
					ks[nq] = k;
					qs[nq] = q;
					fs[nq] = f;
					tq[nq] = __dadd_rn(__dsub_rn(cur.L[k], bv), f);          /* :665-666 */
					xs[nq] = __dadd_rn(cur.pfx[k], f);
					nq++;
				}
				if (nq) {
					w0 = atomicAdd(&s_nwork, nq);
#pragma unroll
					for (int q = 0; q < 4; q++)
						if (q < nq) s_work[w0 + q] = xs[q];
				}
			}
			__syncthreads();
			for (int w = tid; w < s_nwork; w += BW) {
				double acc = s_work[w];
				int i = c + 1;
				for (; i + 8 <= N; i += 8) {                           /* terms fetched 8 ahead of the dependent additions */
					double bb[8];
#pragma unroll
					for (int u = 0; u < 8; u++) bb[u] = sbase[i + u];
#pragma unroll
					for (int u = 0; u < 8; u++) acc = __dadd_rn(acc, bb[u]);
				}
				for (; i < N; i++) acc = __dadd_rn(acc, sbase[i]);
				s_work[w] = acc;
			}
			__syncthreads();
			if (act) {
				bool has = false;
				double R = CNB_NEG_INF, bf = 0.0;
				long long bo = -1;
				int bs = -1;
#pragma unroll
				for (int q = 0; q < 4; q++) {
					if (q >= nq) continue;
					if (!has && tq[q] > CNB_NEG_INF) { has = true; R = CNB_NEG_INF; }   /* empty net: loglik() = -FLT_MAX */
					if (has && R < tq[q]) { bo = nd.opt_begin + qs[q]; bs = ks[q]; R = s_work[w0 + q]; bf = fs[q]; }   /* :670-671 strict */
				}
				const bool cur_ex = cur.exists[j] != 0;
				const bool take = cur_ex ? (has && bo >= 0 && cur.L[j] < R) : (has && bo >= 0);   /* :847-859 */
				const size_t bpi = (size_t)c * (K + 1) + j;
				if (take) {
					nxt.exists[j] = 1;
					nxt.L[j] = R;
					nxt.pfx[j] = __dadd_rn(cur.pfx[bs], bf);
					if (owned) { bp_opt[bpi] = bo; bp_src[bpi] = bs; }
				} else {
					nxt.exists[j] = cur_ex;
					nxt.L[j] = cur.L[j];
					nxt.pfx[j] = __dadd_rn(cur.pfx[j], bv);
					if (owned) { bp_opt[bpi] = -1; bp_src[bpi] = j; }
				}
			}
			__syncthreads();
		}
		if (owned && in_range) {
			const int fb = (nT & 1) * BW + tid;
			gout.exists[j] = s_ex[fb];
			gout.L[j] = s_L[fb];
			gout.pfx[j] = s_pfx[fb];
		}
		__syncthreads();
		if (tid == 0) {
			__threadfence();
			atomicExch(&flags[b], (unsigned)m + 1u);
		}
	}
}
#undef DP_PREFETCH

/* walk the back-pointers of every surviving slot: sel[j][c] = option chosen for node c (-1 = base family) */
__global__ void k_dp_collect(DpState fin, int N, const long long *__restrict__ bp_opt, const int *__restrict__ bp_src,
		long long *__restrict__ sel) {
	int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j > fin.K || !fin.exists[j]) return;
	int jj = j;
	sel[(size_t)j * N] = -1;
	for (int c = N - 1; c >= 1; c--) {
		size_t bpi = (size_t)c * (fin.K + 1) + jj;
		sel[(size_t)j * N + c] = bp_opt[bpi];
		jj = bp_src[bpi];
	}
}

/* same walk for a list of slots only (the surviving ones): sel[e][c] for e-th listed slot */
__global__ void k_dp_collect_list(DpState fin, int N, const long long *__restrict__ bp_opt, const int *__restrict__ bp_src,
		const int *__restrict__ slots, int nslots, long long *__restrict__ sel) {
	int e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= nslots) return;
	int jj = slots[e];
	sel[(size_t)e * N] = -1;
	for (int c = N - 1; c >= 1; c--) {
		size_t bpi = (size_t)c * (fin.K + 1) + jj;
		sel[(size_t)e * N + c] = bp_opt[bpi];
		jj = bp_src[bpi];
	}
}

__global__ void k_device_log(const double *__restrict__ x, long long n, double *__restrict__ out) {
	long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = cnb_log(x[i]);
}

/* ------------------------------------------------------------------ launchers ---- */
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cnb_set_error("CUDA: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); return CNB_ERR_CUDA; } } while (0)

int cnbk_minmax(cudaStream_t st, const void *samples, int elem_bytes, int N, int M, int *vmin, int *vmax) {
	dim3 grid((N + 127) / 128, min(M, 1024));
	if (elem_bytes == 1) k_minmax<uint8_t><<<grid, 128, 0, st>>>((const uint8_t *)samples, N, M, vmin, vmax);
	else k_minmax<int32_t><<<grid, 128, 0, st>>>((const int32_t *)samples, N, M, vmin, vmax);
	CK(cudaGetLastError());
	return CNB_OK;
}

int cnbk_pack(cudaStream_t st, const void *samples, int elem_bytes, const int32_t *pert, int N, int M, int W, const int *vmin,
		const int *cats, uint8_t *codes, uint4 *planes, uint32_t *keep, int *bad, int w_begin, int w_end) {
	if (w_end < 0) w_end = W;
	if (w_end <= w_begin) return CNB_OK;
	dim3 grid((unsigned)((long long)(w_end - w_begin) * ((N + 127) / 128)));
	if (elem_bytes == 1) k_pack<uint8_t><<<grid, 128, 0, st>>>((const uint8_t *)samples, pert, N, M, W, vmin, cats, codes, planes, keep, bad, w_begin);
	else k_pack<int32_t><<<grid, 128, 0, st>>>((const int32_t *)samples, pert, N, M, W, vmin, cats, codes, planes, keep, bad, w_begin);
	CK(cudaGetLastError());
	return CNB_OK;
}

int cnbk_permute(cudaStream_t st, const uint4 *planes_lbl, const uint32_t *keep_lbl, const int *order, int N, int W,
		uint4 *planes, uint32_t *keep, int w_begin, int w_end) {
	if (w_end < 0) w_end = W;
	if (w_end <= w_begin) return CNB_OK;
	dim3 grid((unsigned)((long long)(w_end - w_begin) * ((N + 127) / 128)));
	k_permute<<<grid, 128, 0, st>>>(planes_lbl, keep_lbl, order, N, W, planes, keep, w_begin);
	CK(cudaGetLastError());
	return CNB_OK;
}

bool cnbk_planes_supported(int d, int max_cats) {
	if (max_cats > CNB_PLANE_CATS || d < 1 || d > 3) return false;
	if (d == 3 && max_cats > 3) return false;            /* 256 counters per thread: use the histogram kernel */
	return true;
}

template <int D, int CM>
static int launch_planes(cudaStream_t st, const DevData &Dt, const DevFamilies &F, long long b, long long e, int slices,
		const ScoreOut &O) {
	constexpr int CELLS = CellCount<D, CM>::value;
	size_t smem = (size_t)CELLS * SCORE_BS * sizeof(int);
	if (smem > 48 * 1024)     /* per device, so set on every launch (cheap) */
		CK(cudaFuncSetAttribute(k_score_planes<D, CM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	long long n = e - b;
	int wps = (Dt.W + slices - 1) / slices;
	dim3 grid((unsigned)((n + SCORE_BS - 1) / SCORE_BS), slices);
	k_score_planes<D, CM><<<grid, SCORE_BS, smem, st>>>(Dt, F, b, e, wps, O);
	CK(cudaGetLastError());
	return CNB_OK;
}

int cnbk_score_planes(cudaStream_t st, const DevData &Dt, const DevFamilies &F, int max_cats, long long fid_begin,
		long long fid_end, int slices, const ScoreOut &O) {
	if (fid_end <= fid_begin) return CNB_OK;
	if (slices > 1 && !O.counts) { cnb_set_error("sample slicing needs a counts buffer"); return CNB_ERR_PROC; }
	const int cm = max_cats <= 2 ? 2 : (max_cats == 3 ? 3 : 4);
#define CASE(D_, C_) if (F.d == D_ && cm == C_) return launch_planes<D_, C_>(st, Dt, F, fid_begin, fid_end, slices, O)
	CASE(1, 2); CASE(1, 3); CASE(1, 4);
	CASE(2, 2); CASE(2, 3); CASE(2, 4);
	CASE(3, 2); CASE(3, 3);
#undef CASE
	cnb_set_error("bit-plane scorer: unsupported (d=%d, cats=%d)", F.d, max_cats);
	return CNB_ERR_PROC;
}

int cnbk_pair_tables(cudaStream_t st, const uint4 *planes_lbl, int N, int W, int *out) {
	long long npairs = (long long)N * (N - 1) / 2;
	if (npairs <= 0) return CNB_OK;
	k_pair_tables<<<(unsigned)((npairs + 127) / 128), 128, 0, st>>>(planes_lbl, N, W, npairs, out);
	CK(cudaGetLastError());
	return CNB_OK;
}

template <int CM>
static int launch_ie2(cudaStream_t st, const DevData &Dt, const DevFamilies &F, long long b, long long e,
		const ScoreOut &O, const int *pair_tab) {
	size_t smem = (size_t)CM * CM * CM * SCORE_BS * sizeof(int);
	if (smem > 48 * 1024)
		CK(cudaFuncSetAttribute(k_score_planes_ie2<CM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	k_score_planes_ie2<CM><<<(unsigned)((e - b + SCORE_BS - 1) / SCORE_BS), SCORE_BS, smem, st>>>(Dt, F, b, e, O, pair_tab);
	CK(cudaGetLastError());
	return CNB_OK;
}

/* two-parent families, no missing values, no perturbation mask: inclusion-exclusion scorer */
int cnbk_score_planes_ie2(cudaStream_t st, const DevData &Dt, const DevFamilies &F, int max_cats, long long b,
		long long e, const ScoreOut &O, const int *pair_tab) {
	if (e <= b) return CNB_OK;
	if (F.d != 2 || max_cats > 4 || !pair_tab) { cnb_set_error("ie2 scorer: unsupported launch"); return CNB_ERR_PROC; }
	if (max_cats <= 2) return launch_ie2<2>(st, Dt, F, b, e, O, pair_tab);
	if (max_cats == 3) return launch_ie2<3>(st, Dt, F, b, e, O, pair_tab);
	return launch_ie2<4>(st, Dt, F, b, e, O, pair_tab);
}

/* three-way tables of all node triples, label space: out [C(N,3)][64] */
int cnbk_triple_tables(cudaStream_t st, const uint4 *planes_lbl, int N, int W, int max_cats, const int *pair_tab, int *out) {
	const long long nt = (long long)N * (N - 1) * (N - 2) / 6;
	if (nt <= 0) return CNB_OK;
	if (max_cats > 4 || !pair_tab) { cnb_set_error("triple tables: unsupported data"); return CNB_ERR_PROC; }
	const unsigned grid = (unsigned)((nt + 127) / 128);
	if (max_cats <= 2) k_triple_tables<2><<<grid, 128, 0, st>>>(planes_lbl, N, W, nt, pair_tab, out);
	else if (max_cats == 3) k_triple_tables<3><<<grid, 128, 0, st>>>(planes_lbl, N, W, nt, pair_tab, out);
	else k_triple_tables<4><<<grid, 128, 0, st>>>(planes_lbl, N, W, nt, pair_tab, out);
	CK(cudaGetLastError());
	return CNB_OK;
}

template <int CM, int BS>
static int launch_ie3(cudaStream_t st, const DevData &Dt, const DevFamilies &F, long long b, long long e,
		const ScoreOut &O, const int *triple_tab) {
	size_t smem = (size_t)CM * CM * CM * CM * BS * sizeof(int);
	if (smem > 48 * 1024)
		CK(cudaFuncSetAttribute(k_score_planes_ie3<CM, BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	k_score_planes_ie3<CM, BS><<<(unsigned)((e - b + BS - 1) / BS), BS, smem, st>>>(Dt, F, b, e, O, triple_tab);
	CK(cudaGetLastError());
	return CNB_OK;
}

/* three-parent families, no missing values, no perturbation mask: inclusion-exclusion over the triple tables */
int cnbk_score_planes_ie3(cudaStream_t st, const DevData &Dt, const DevFamilies &F, int max_cats, long long b,
		long long e, const ScoreOut &O, const int *triple_tab) {
	if (e <= b) return CNB_OK;
	if (F.d != 3 || max_cats > 4 || !triple_tab) { cnb_set_error("ie3 scorer: unsupported launch"); return CNB_ERR_PROC; }
	if (max_cats <= 2) return launch_ie3<2, 128>(st, Dt, F, b, e, O, triple_tab);
	if (max_cats == 3) return launch_ie3<3, 128>(st, Dt, F, b, e, O, triple_tab);
	return launch_ie3<4, 64>(st, Dt, F, b, e, O, triple_tab);
}

int cnbk_score_pairs1(cudaStream_t st, const DevData &Dt, const DevFamilies &F, int max_cats, long long b, long long e,
		const ScoreOut &O, const int *pair_tab, bool ordered) {
	if (e <= b) return CNB_OK;
	if (F.d > 1 || max_cats > 4 || !pair_tab) { cnb_set_error("pair-table scorer: unsupported launch"); return CNB_ERR_PROC; }
	const unsigned grid = (unsigned)((e - b + 127) / 128);
	if (F.d == 0 && ordered) k_score_pairs0<true><<<grid, 128, 0, st>>>(Dt, F, b, e, O, pair_tab);
	else if (F.d == 0) k_score_pairs0<false><<<grid, 128, 0, st>>>(Dt, F, b, e, O, pair_tab);
	else if (ordered) k_score_pairs1<true><<<grid, 128, 0, st>>>(Dt, F, b, e, O, pair_tab);
	else k_score_pairs1<false><<<grid, 128, 0, st>>>(Dt, F, b, e, O, pair_tab);
	CK(cudaGetLastError());
	return CNB_OK;
}

int cnbk_epilogue_counts(cudaStream_t st, const DevData &Dt, const DevFamilies &F, long long b, long long e,
		const ScoreOut &O) {
	if (e <= b) return CNB_OK;
	k_epilogue_counts<<<(unsigned)((e - b + 127) / 128), 128, 0, st>>>(Dt, F, b, e, O);
	CK(cudaGetLastError());
	return CNB_OK;
}

template <int D, int MODE>
static int launch_hist_warp_mode(cudaStream_t st, const DevData &Dt, const DevFamilies &F, long long b, long long e, int cpw, int wpb,
		const ScoreOut &O, int *overflow) {
	static int sms = 0;
	if (!sms) {
		int dev = 0;
		CK(cudaGetDevice(&dev));
		CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
	}
	const size_t smem = (size_t)cpw * wpb * sizeof(int);
	const long long nblk = (e - b + wpb - 1) / wpb;
	int per_sm = 1;
	if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_score_hist_warp<D, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_score_hist_warp<D, MODE>, wpb * 32, smem));
	const unsigned grid = (unsigned)std::min<long long>(nblk, (long long)sms * std::max(per_sm, 1));
	k_score_hist_warp<D, MODE><<<grid, wpb * 32, smem, st>>>(Dt, F, b, e, cpw, O, overflow);
	CK(cudaGetLastError());
	return CNB_OK;
}

template <int D>
static int launch_hist_warp(cudaStream_t st, const DevData &Dt, const DevFamilies &F, long long b, long long e, int cpw, int wpb, int kmode,
		const ScoreOut &O, int *overflow) {
	if (kmode == 2) return launch_hist_warp_mode<D, 2>(st, Dt, F, b, e, cpw, wpb, O, overflow);
	if (kmode == 1) return launch_hist_warp_mode<D, 1>(st, Dt, F, b, e, cpw, wpb, O, overflow);
	return launch_hist_warp_mode<D, 0>(st, Dt, F, b, e, cpw, wpb, O, overflow);
}

int cnbk_score_hist(cudaStream_t st, const DevData &Dt, const DevFamilies &F, long long b, long long e, int max_cells,
		const ScoreOut &O, int *overflow, int mode) {
	if (e <= b) return CNB_OK;
	/* mode 0: one CTA per family (k_score_hist: any parent count, tables up to max_cells); 1 / 2: a warp per family with
	 * its own table, atomics / __match_any_sync -- for up to three parents of one count and tables that leave room for at
	 * least one warp's table in 96 KB */
	const int cpw = (max_cells + 31) & ~31;
	if ((mode & 3) != 0 && F.d >= 0 && F.d <= 3 && !F.ex_nd && (size_t)cpw * sizeof(int) <= 96 * 1024 && cpw < 65536) {
		int wpb = 8;
		while (wpb > 1 && (size_t)cpw * wpb * sizeof(int) > 96 * 1024) wpb >>= 1;
		/* kernel mode: __match_any_sync, or atomics with / without per-sample checks (bit 2 of mode: the data has neither
		 * missing values nor a keep mask) */
		const int kmode = (mode & 3) == 2 ? 2 : (((mode & 4) && !Dt.keep_lbl) ? 1 : 0);
		switch (F.d) {
		case 0: return launch_hist_warp<0>(st, Dt, F, b, e, cpw, wpb, kmode, O, overflow);
		case 1: return launch_hist_warp<1>(st, Dt, F, b, e, cpw, wpb, kmode, O, overflow);
		case 2: return launch_hist_warp<2>(st, Dt, F, b, e, cpw, wpb, kmode, O, overflow);
		default: return launch_hist_warp<3>(st, Dt, F, b, e, cpw, wpb, kmode, O, overflow);
		}
	}
	size_t smem = (size_t)((max_cells + 1) & ~1) * sizeof(int) + (size_t)max_cells * sizeof(double);
	if (smem > 40 * 1024)                                 /* the kernel's static shared variables count towards the 48 KB default too */
		CK(cudaFuncSetAttribute(k_score_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	long long n = e - b;
	unsigned grid = (unsigned)(n < (1ll << 30) ? n : (1ll << 30));
	k_score_hist<<<grid, 256, smem, st>>>(Dt, F, b, e, max_cells, O, overflow);
	CK(cudaGetLastError());
	return CNB_OK;
}

__global__ void k_fill_f64(double *__restrict__ p, long long n, double v) {
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

int cnbk_fill_f64(cudaStream_t st, double *p, long long n, double v) {
	if (n <= 0) return CNB_OK;
	k_fill_f64<<<(unsigned)std::min<long long>((n + 255) / 256, 148 * 16), 256, 0, st>>>(p, n, v);
	CK(cudaGetLastError());
	return CNB_OK;
}

int cnbk_select_chunk(void) { return SEL_CHUNK; }

int cnbk_select(cudaStream_t st, const DevData &Dt, int nblocks, const CnbGroup *groups, const int *blk_equal,
		const int2 *chunks, int nchunks, const int *blk_chunk0, const double *scores, long long seg_len, double *opt_score,
		int *opt_cx, long long *opt_fid, double *part_score, long long *part_fid) {
	if (nblocks <= 0 || nchunks <= 0) return CNB_OK;
	k_select<<<nchunks, 256, 0, st>>>(Dt, groups, blk_equal, chunks, scores, seg_len, opt_score, opt_cx, opt_fid, part_score, part_fid);
	CK(cudaGetLastError());
	k_select_combine<<<(nblocks + 127) / 128, 128, 0, st>>>(Dt, groups, blk_equal, blk_chunk0, nblocks, part_score, part_fid,
			opt_score, opt_cx, opt_fid);
	CK(cudaGetLastError());
	return CNB_OK;
}

int cnbk_patch_options(cudaStream_t st, int n, const long long *opt, const long long *fid, const int *cx, const double *prior,
		const double *ex_ll, const double *ex_score, double *opt_score, int *opt_cx, long long *opt_fid) {
	if (n <= 0) return CNB_OK;
	k_patch_options<<<(n + 127) / 128, 128, 0, st>>>(n, opt, fid, cx, prior, ex_ll, ex_score, opt_score, opt_cx, opt_fid);
	CK(cudaGetLastError());
	return CNB_OK;
}

int cnbk_dp_init(cudaStream_t st, const DpState &S, int base_complexity, const double *base_v, int N) {
	k_dp_init<<<(S.K + 1 + 127) / 128, 128, 0, st>>>(S, base_complexity, base_v, N);
	CK(cudaGetLastError());
	return CNB_OK;
}

int cnbk_dp_node(cudaStream_t st, const DpState &cur, const DpState &nxt, int c, int N, int base_cx_c,
		const double *base_v, long long opt_begin, long long opt_end, const double *opt_score, const int *opt_cx,
		const long long *opt_fid, long long *bp_opt, int *bp_src) {
	k_dp_node<<<(cur.K + 1 + 127) / 128, 128, 0, st>>>(cur, nxt, c, N, base_cx_c, base_v, opt_begin, opt_end,
			opt_score, opt_cx, opt_fid, bp_opt, bp_src);
	CK(cudaGetLastError());
	return CNB_OK;
}

/* all nodes in one cooperative launch; returns CNB_ERR_PROC (and launches nothing) when the device cannot co-schedule
 * the grid, so that the caller falls back to one launch per node */
int cnbk_dp_all(cudaStream_t st, const DpState &s0, const DpState &s1, int N, int base_complexity, const CnbDpNode *nodes,
		const double *base_v, const double *opt_score, const int *opt_cx, const long long *opt_fid, long long *bp_opt,
		int *bp_src) {
	int dev = 0, coop = 0, sms = 0, per_sm = 0;
	CK(cudaGetDevice(&dev));
	CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
	CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
	const size_t smem = (size_t)N * sizeof(double);
	if (smem > 48 * 1024) return CNB_ERR_PROC;
	CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_dp_all, 256, smem));
	if (!coop || per_sm < 1) return CNB_ERR_PROC;
	const int blocks = std::min(sms * per_sm, std::max(1, (s0.K + 1 + 255) / 256));   /* one slot per thread if it fits */
	void *args[] = {(void *)&s0, (void *)&s1, (void *)&N, (void *)&base_complexity, (void *)&nodes, (void *)&base_v,
		(void *)&opt_score, (void *)&opt_cx, (void *)&opt_fid, (void *)&bp_opt, (void *)&bp_src};
	CK(cudaLaunchCooperativeKernel((const void *)k_dp_all, dim3(blocks), dim3(256), args, smem, st));
	return CNB_OK;
}

/* warp-per-slot DP for searches with many options per node; CNB_ERR_PROC (nothing launched) when refused */
int cnbk_dp_mixed(cudaStream_t st, const DpState &s0, const DpState &s1, int N, int base_complexity, const CnbDpNode *nodes,
		const double *base_v, const double *opt_score, const int *opt_cx, const long long *opt_fid, long long *bp_opt,
		int *bp_src) {
	int dev = 0, coop = 0, sms = 0, per_sm = 0;
	CK(cudaGetDevice(&dev));
	CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
	CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
	if ((size_t)N * sizeof(double) > 16 * 1024) return CNB_ERR_PROC;
	/* the previous state is staged in shared memory when it fits (K + 1 <= DPM_SLOTS), a node's options when there are
	 * at most DPM_OPTS of them */
	int stage_state = s0.K + 1 <= DPM_SLOTS;
	const size_t smem = (size_t)N * sizeof(double) + (size_t)2 * DPM_OPTS * (sizeof(double) + sizeof(int))
			+ (stage_state ? (size_t)2 * (s0.K + 1) * sizeof(double) + (size_t)((s0.K + 1 + 15) & ~15) : 0);
	/* one CTA of 1024 threads, no grid barrier (CNB_DP_ONE_CTA: largest K + 1 that takes this form).  Off by default: on
	 * ALARM (601 slots, hundreds of options per node) it is 3.9 x SLOWER than the cooperative grid -- cnSearchSA 1.32 s
	 * against 0.34 s -- the scan of a slot's options is real work for 600 warps, not barrier latency */
	static int one_max = -1;
	if (one_max < 0) { const char *e = getenv("CNB_DP_ONE_CTA"); one_max = e ? atoi(e) : 0; }
	if (s0.K + 1 <= one_max) {
		if (smem > 48 * 1024)
			CK(cudaFuncSetAttribute(k_dp_mixed<1024, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		k_dp_mixed<1024, true><<<1, 1024, smem, st>>>(s0, s1, N, base_complexity, nodes, base_v, opt_score, opt_cx, opt_fid, bp_opt, bp_src,
				stage_state);
		CK(cudaGetLastError());
		return CNB_OK;
	}
	/* the opt-in above 48 KB is a property of (function, device): asked for once per size, not once per evaluated order */
	static std::atomic<size_t> opted[64];                /* zero-initialised; several host threads evaluate orders side by side */
	if (smem > 48 * 1024 && (dev < 0 || dev >= 64 || smem > opted[dev].load())) {
		CK(cudaFuncSetAttribute(k_dp_mixed<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		if (dev >= 0 && dev < 64) opted[dev].store(smem);
	}
	CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_dp_mixed<256, false>, 256, smem));
	if (!coop || per_sm < 1) return CNB_ERR_PROC;
	const int blocks = std::min(sms * per_sm, std::max(1, (s0.K + 1 + 7) / 8));      /* one warp per slot if it fits */
	void *args[] = {(void *)&s0, (void *)&s1, (void *)&N, (void *)&base_complexity, (void *)&nodes, (void *)&base_v,
		(void *)&opt_score, (void *)&opt_cx, (void *)&opt_fid, (void *)&bp_opt, (void *)&bp_src, (void *)&stage_state};
	CK(cudaLaunchCooperativeKernel((const void *)k_dp_mixed<256, false>, dim3(blocks), dim3(256), args, smem, st));
	return CNB_OK;
}

/* wavefront DP: needs one slot per thread in co-resident CTAs and options that look at most 256 slots left;
 * CNB_ERR_PROC (nothing launched) otherwise.  flags: gridDim.x counters, zeroed here. */
int cnbk_dp_wave(cudaStream_t st, const DpState &s0, const DpState &s1, const DpState &s2, int N, int base_complexity,
		int halo, const CnbDpNode *nodes, const double *base_v, const double *opt_score, const int *opt_cx,
		const long long *opt_fid, long long *bp_opt, int *bp_src, unsigned *flags, int flags_cap) {
	int dev = 0, coop = 0, sms = 0, per_sm = 0;
	CK(cudaGetDevice(&dev));
	CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
	CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
	const size_t smem = (size_t)N * sizeof(double);
	if (smem > 48 * 1024 || halo > 256 || halo < 0) return CNB_ERR_PROC;
	CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_dp_wave, 256, smem));
	const int blocks = (s0.K + 1 + 255) / 256;
	if (!coop || per_sm < 1 || blocks > sms * per_sm || blocks > flags_cap) return CNB_ERR_PROC;
	CK(cudaMemsetAsync(flags, 0, (size_t)blocks * sizeof(unsigned), st));
	void *args[] = {(void *)&s0, (void *)&s1, (void *)&s2, (void *)&N, (void *)&base_complexity, (void *)&nodes,
		(void *)&base_v, (void *)&opt_score, (void *)&opt_cx, (void *)&opt_fid, (void *)&bp_opt, (void *)&bp_src, (void *)&flags};
	CK(cudaLaunchCooperativeKernel((const void *)k_dp_wave, dim3(blocks), dim3(256), args, smem, st));
	return CNB_OK;
}

/* trapezoid DP: at most 4 options per node, halo * T <= 256; *steps_out = macro-steps (final state in buffer steps % 3);
 * CNB_ERR_PROC (nothing launched) when it does not apply */
int cnbk_dp_trap(cudaStream_t st, const DpState &s0, const DpState &s1, const DpState &s2, int N, int base_complexity,
		int halo, const CnbDpNode *nodes, const double *base_v, const double *opt_score, const int *opt_cx,
		const long long *opt_fid, long long *bp_opt, int *bp_src, unsigned *flags, int flags_cap, int *steps_out,
		int compact) {
	int dev = 0, coop = 0, sms = 0, per_sm = 0;
	CK(cudaGetDevice(&dev));
	CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
	CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
	if (halo < 1 || N < 3) return CNB_ERR_PROC;
	/* the recomputed margin must not reach past the left neighbour's own slots: the neighbour protocol waits for one CTA on
	 * either side (wider margins measured no faster anyway: the per-node latency dominates, not the exchanges) */
	int TRAP_S = 256;
	int T = std::min(TRAP_TMAX, TRAP_S / halo);
	if (T < 2) return CNB_ERR_PROC;
	/* 256 owned slots per CTA: with 768 (fewer CTAs, a shorter pipeline fill) the DP of C5 took 4.8 instead of 2.3 ms --
	 * the re-summation chains of 1024 slots per SM cost more than the fill saves */
	const int owned = 256, threads = owned + TRAP_S;
	const size_t smem = (size_t)N * sizeof(double) + (size_t)2 * threads * (2 * sizeof(double) + 1) + (size_t)4 * threads * sizeof(double);
	if (smem > 160 * 1024) return CNB_ERR_PROC;
	if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_dp_trap, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_dp_trap, threads, smem));
	const int blocks = (s0.K + 1 + owned - 1) / owned;
	if (!coop || per_sm < 1 || blocks > sms * per_sm || blocks > flags_cap) return CNB_ERR_PROC;
	CK(cudaMemsetAsync(flags, 0, (size_t)blocks * sizeof(unsigned), st));
	/* CTAs skip the macro-steps in which none of their slots can exist yet: every state buffer starts as "no network" */
	CK(cudaMemsetAsync(s0.exists, 0, (size_t)s0.K + 1, st));
	CK(cudaMemsetAsync(s1.exists, 0, (size_t)s0.K + 1, st));
	CK(cudaMemsetAsync(s2.exists, 0, (size_t)s0.K + 1, st));
	void *args[] = {(void *)&s0, (void *)&s1, (void *)&s2, (void *)&N, (void *)&base_complexity, (void *)&nodes,
		(void *)&base_v, (void *)&opt_score, (void *)&opt_cx, (void *)&opt_fid, (void *)&bp_opt, (void *)&bp_src, (void *)&flags,
		(void *)&T, (void *)&halo, (void *)&compact, (void *)&TRAP_S};
	CK(cudaLaunchCooperativeKernel((const void *)k_dp_trap, dim3(blocks), dim3(threads), args, smem, st));
	*steps_out = (N - 1 + T - 1) / T;
	return CNB_OK;
}

int cnbk_dp_collect(cudaStream_t st, const DpState &fin, int N, const long long *bp_opt, const int *bp_src,
		long long *sel) {
	k_dp_collect<<<(fin.K + 1 + 127) / 128, 128, 0, st>>>(fin, N, bp_opt, bp_src, sel);
	CK(cudaGetLastError());
	return CNB_OK;
}

int cnbk_dp_collect_list(cudaStream_t st, const DpState &fin, int N, const long long *bp_opt, const int *bp_src,
		const int *slots, int nslots, long long *sel) {
	if (nslots <= 0) return CNB_OK;
	k_dp_collect_list<<<(nslots + 63) / 64, 64, 0, st>>>(fin, N, bp_opt, bp_src, slots, nslots, sel);
	CK(cudaGetLastError());
	return CNB_OK;
}

int cnbk_device_log(cudaStream_t st, const double *x, long long n, double *out) {
	if (n <= 0) return CNB_OK;
	k_device_log<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, n, out);
	CK(cudaGetLastError());
	return CNB_OK;
}
