/* cnb_plan.cpp -- host-side planner: turns the .Call arguments into the candidate-block table the kernels walk.
 *
 * Restates, for one node order, what CATNET_SEARCH2::estimate derives before it scores anything
 * (/root/reference/src/catnet_search2.h):
 *   :182-183  maxComplexity clamped up to numnodes
 *   :456-489  per child: predecessors allowed by parentsPool, split into fixed / free parents
 *   :491-496  "equal categories" test over the child's pool, which picks one of the two DP code paths
 *   :498-503  maxpars = min(maxParentSet, parentSizes[child], pool size)
 *   :509,:530 one candidate block per parent count d: fixed parents + every increasing subset of the free pool
 * and the order-space marshalling of RCatnetSearch::estimateCatnets (src/rcatnet_search.cpp:141-268).
 * Pure host C++ (no CUDA) so that it is covered by the CPU test-suite.
 */
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include "cnb_internal.h"
#include "cnb_tiles.h"

static thread_local char g_err[512] = "";

void cnb_set_error(const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
}

extern "C" const char *cnb_last_error(void) { return g_err; }

int64_t cnb_binom(int64_t n, int64_t k) {
	if (k < 0 || k > n) return 0;
	if (k > n - k) k = n - k;
	int64_t r = 1;
	for (int64_t j = 1; j <= k; j++) {
		/* r*(n-k+j)/j stays integral; guard overflow */
		if (r > INT64_MAX / (n - k + j)) return INT64_MAX;
		r = r * (n - k + j) / j;
	}
	return r;
}

/* rank -> k-subset of {0..n-1} in lexicographic order: the order combinationSets() emits (catnet_search2.h:102-147) */
int cnb_unrank_lex(int32_t n, int32_t k, int64_t r, int32_t *out) {
	if (k < 0 || k > n || r < 0 || r >= cnb_binom(n, k)) return CNB_ERR_PARAM;
	int32_t base = 0;
	for (int32_t t = 0; t < k; t++) {
		int32_t kk = k - t;
		int64_t total = cnb_binom(n, kk);
		/* largest i with  total - C(n-i, kk) <= r  */
		int32_t lo = 0, hi = n - kk;
		while (lo < hi) {
			int32_t mid = (lo + hi + 1) >> 1;
			if (total - cnb_binom(n - mid, kk) <= r) lo = mid;
			else hi = mid - 1;
		}
		r -= total - cnb_binom(n - lo, kk);
		out[t] = base + lo;
		base += lo + 1;
		n -= lo + 1;
	}
	return CNB_OK;
}

extern "C" int cnb_unrank_combination(int32_t n, int32_t k, int64_t rank, int32_t *out) {
	return cnb_unrank_lex(n, k, rank, out);
}

int cnb_build_plan(const cnb_params *p, const int32_t *cats_lbl, int32_t rank, int32_t world, CnbPlan *plan,
		std::string *err, int tile_mode) {
	char buf[256];
#define FAIL(code, ...) do { snprintf(buf, sizeof(buf), __VA_ARGS__); if (err) *err = buf; cnb_set_error("%s", buf); return code; } while (0)
	if (!p || !plan) FAIL(CNB_ERR_PARAM, "null parameters");
	const int N = p->num_nodes, M = p->num_samples;
	if (N < 1 || M < 1) FAIL(CNB_ERR_PARAM, "Invalid sample");   /* catnet_search2.h:180 */
	if (world < 1 || rank < 0 || rank >= world) FAIL(CNB_ERR_PARAM, "invalid shard %d/%d", rank, world);
	CnbPlan &pl = *plan;
	pl = CnbPlan();
	pl.N = N;
	pl.M = M;
	pl.P = p->max_parent_set;
	pl.K = std::max(p->max_complexity, N);
	pl.rank = rank;
	pl.world = world;

	/* order: rcatnet_search.cpp:111-128 */
	pl.order.resize(N);
	std::vector<int32_t> pos_of(N, -1);
	for (int i = 0; i < N; i++) {
		int lbl = p->order ? p->order[i] : i + 1;
		if (lbl <= 0 || lbl > N) FAIL(CNB_ERR_PARAM, "Invalid nodeOrder parameter");
		if (pos_of[lbl - 1] >= 0) FAIL(CNB_ERR_PARAM, "Invalid nodeOrder parameter");
		pos_of[lbl - 1] = i;
		pl.order[i] = lbl - 1;
	}
	pl.cats.resize(N);
	for (int i = 0; i < N; i++) {
		pl.cats[i] = cats_lbl[pl.order[i]];
		if (pl.cats[i] < 1) FAIL(CNB_ERR_PARAM, "Incorrect categories");
		pl.max_cats = std::max(pl.max_cats, pl.cats[i]);
	}

	/* pools -> membership in order space (rcatnet_search.cpp:191-256) */
	auto load_pool = [&](const int32_t *pool, std::vector<uint8_t> &memb, std::vector<uint8_t> &has_row) {
		memb.assign((size_t)N * N, 0);
		has_row.assign(N, 0);
		for (int i = 0; i < N; i++) {
			const int32_t *row = pool + (size_t)pl.order[i] * N;
			for (int j = 0; j < N && row[j] > 0; j++) {
				has_row[i] = 1;
				if (row[j] <= N) memb[(size_t)i * N + pos_of[row[j] - 1]] = 1;
			}
		}
	};
	std::vector<uint8_t> pool_m, pool_has, fix_m, fix_has;
	if (p->parents_pool) load_pool(p->parents_pool, pool_m, pool_has);
	if (p->fixed_parents) load_pool(p->fixed_parents, fix_m, fix_has);

	if (p->edge_prob) {
		pl.has_prior = true;
		pl.edge.resize((size_t)N * N);
		for (int j = 0; j < N; j++)
			for (int i = 0; i < N; i++)
				pl.edge[(size_t)j * N + i] = p->edge_prob[(size_t)pl.order[j] * N + pl.order[i]];
	}

	pl.nodes.resize(N);
	std::vector<int32_t> fix, fre;
	pl.fix1.assign(N, -1);
	for (int c = 0; c < N; c++) {
		CnbNodePlan &nd = pl.nodes[c];
		fix.clear();
		fre.clear();
		for (int j = 0; j < c; j++) {
			/* an empty parentsPool row is encoded by the reference as a self reference: nothing allowed */
			bool allow = !p->parents_pool || pool_m[(size_t)c * N + j];
			if (!allow) continue;
			bool fixd = p->fixed_parents && fix_m[(size_t)c * N + j];
			(fixd ? fix : fre).push_back(j);
		}
		nd.nfix = (int32_t)fix.size();
		pl.fix1[c] = fix.size() == 1 ? fix[0] : -1;
		nd.fix_off = (int32_t)pl.lists.size();
		pl.lists.insert(pl.lists.end(), fix.begin(), fix.end());
		int32_t free_off = (int32_t)pl.lists.size();
		pl.lists.insert(pl.lists.end(), fre.begin(), fre.end());
		nd.nfree = (int32_t)fre.size();
		nd.free_off = free_off;
		int64_t cx = pl.cats[c] - 1;
		for (int v : fix) cx *= pl.cats[v];
		if (cx > INT32_MAX) FAIL(CNB_ERR_PARAM, "fixed parent set of node %d too complex", c + 1);
		nd.base_cx = (int32_t)cx;
		pl.base_complexity += nd.base_cx;
		/* equal categories over parset = free ++ fixed (catnet_search2.h:488-496) */
		nd.equal_branch = 1;
		int first_cat = !fre.empty() ? pl.cats[fre[0]] : (!fix.empty() ? pl.cats[fix[0]] : 0);
		for (int v : fre) if (pl.cats[v] != first_cat) nd.equal_branch = 0;
		for (int v : fix) if (pl.cats[v] != first_cat) nd.equal_branch = 0;
		int maxpars = p->max_parent_set;
		if (p->parent_sizes && p->parent_sizes[pl.order[c]] < maxpars) maxpars = p->parent_sizes[pl.order[c]];
		if (maxpars > (int)(fre.size() + fix.size())) maxpars = (int)(fre.size() + fix.size());
		nd.first_blk = (int32_t)pl.blocks.size();
		nd.nblk = 0;
		if (nd.nfix > CNB_MAXP) FAIL(CNB_ERR_PARAM, "more than %d fixed parents", CNB_MAXP);
		if (c == 0) continue;                                   /* main loop starts at node 1 (:444) */
		for (int d = nd.nfix + 1; d <= maxpars; d++) {
			if (d > CNB_MAXP) FAIL(CNB_ERR_PARAM, "maxParentSet > %d is not supported", CNB_MAXP);
			CnbBlock b;
			memset(&b, 0, sizeof(b));
			b.child = c;
			b.d = d;
			b.nfix = nd.nfix;
			b.nfree = (int32_t)fre.size();
			b.fix_off = nd.fix_off;
			b.free_off = free_off;
			b.ncomb = cnb_binom(b.nfree, d - nd.nfix);
			if (b.ncomb <= 0) continue;
			if (b.ncomb == INT64_MAX) FAIL(CNB_ERR_PARAM, "too many candidate parent sets");
			pl.blocks.push_back(b);
			nd.nblk++;
			pl.max_d = std::max(pl.max_d, d);
		}
	}
	if (pl.base_complexity > pl.K)
		FAIL(CNB_ERR_PARAM, "maxComplexity %d below the complexity %d of the base network", pl.K, pl.base_complexity);

	/* groups by parent count; family ids inside a group follow block order (child ascending) */
	for (int d = 1; d <= pl.max_d; d++) {
		CnbGroup g;
		memset(&g, 0, sizeof(g));
		g.d = d;
		g.blk_off = (int32_t)pl.group_blocks.size();
		for (size_t b = 0; b < pl.blocks.size(); b++) {
			if (pl.blocks[b].d != d) continue;
			pl.blocks[b].group = (int32_t)pl.groups.size();
			pl.blocks[b].start = g.nfam;
			g.nfam += pl.blocks[b].ncomb;
			pl.group_blocks.push_back((int32_t)b);
			g.nblk++;
		}
		if (!g.nblk) continue;
		g.chunk = (g.nfam + world - 1) / world;
		g.full = g.nblk == N - d;
		for (int k = 0; k < g.nblk && g.full; k++) {
			const CnbBlock &b = pl.blocks[pl.group_blocks[g.blk_off + k]];
			g.full = b.nfix == 0 && b.nfree == b.child;
		}
		/* Tensor-core tiling of the two-parent group.  The tiles always cover ALL families (a < b | c) of the order: tile (child
		 * tile, pair tile) slots map one-to-one onto the unrestricted candidate families.  tiled = 1: every node c >= 2 offers all
		 * pairs of its predecessors (no pools, no fixed parents, no size limit), a candidate's rank gives its slot directly.
		 * tiled = 2: the candidates are a subset (parent pools, fixed parents, per-node parent-set sizes; every candidate is still
		 * two predecessors of its child): the selection looks each candidate's slot up from its parents -- worth it while the
		 * subset is at least 1/16 of the full group (a tensor-core slot costs 1/27 .. 1/60 of a bit-plane family).  Tiles cost
		 * the same, so ranks get equal shares, dealt round-robin. */
		if (tile_mode && d == 2 && N >= 3) {
			bool full = g.nblk == N - 2;
			for (int k = 0; k < g.nblk && full; k++) {
				const CnbBlock &b = pl.blocks[pl.group_blocks[g.blk_off + k]];
				full = full && b.nfix == 0 && b.nfree == b.child;
			}
			const bool subset = !full && g.nfam * 16 >= cnb_binom(N, 3);
			if (subset) full = true;
			const int tiled_kind = subset ? 2 : 1;
			if (full) {
				/* tile_mode 3: the inclusion-exclusion tiles contract only as many free categories as the data has
				 * (cnb_tiles.h: cnb_geo_*) -- more families per MMA for data with two or three categories per node */
				const int fr = tile_mode == 3 ? cnb_geo_fr(pl.max_cats) : 3;
				const int tp = tile_mode == 2 ? CNB_FULL_PAIRS : cnb_geo_pairs(fr), tc = tile_mode == 2 ? CNB_FULL_CHILD : cnb_geo_child(fr);
				pl.tile_pairs = tp;
				pl.tile_cols = tc;
				pl.tile_full = tile_mode == 2;
				pl.tile_fr = fr;
				int nct = (N + tc - 1) / tc;
				int cpt = (N + nct - 1) / nct;                      /* equal child tiles (500 nodes: 8 x 63, not 7 x 64 + 52) */
				pl.tile_child = cpt;
				pl.tile_base.assign(nct + 1, 0);
				pl.dtile_off.assign((size_t)nct * CNB_DT_STRIDE, 0);
				for (int ct = 0; ct < nct; ct++) {
					const int c0 = ct * cpt, ce = std::min(c0 + cpt, N);
					int64_t *off = &pl.dtile_off[(size_t)ct * CNB_DT_STRIDE];
					const int ndp = cnb_d_pairs(c0, ce), ndt = (ndp + tp - 1) / tp;
					for (int p = 0; p < ndt; p++) off[p + 1] = off[p] + cnb_d_chunks(c0, ce, p, tp, tc);
					for (int p = ndt + 1; p < CNB_DT_STRIDE; p++) off[p] = off[ndt];
					pl.tile_base[ct + 1] = pl.tile_base[ct] + cnb_f_tiles(c0, tp) + off[ndt];
				}
				g.tiled = tiled_kind;
				g.ntiles = pl.tile_base[nct];
				g.tiles_per_rank = (g.ntiles + world - 1) / world;
				g.tile_world = world;
				g.tile_child = cpt;
				g.chunk = g.tiles_per_rank * tp * tc;
			}
		}
		g.seg_off = pl.seg_len;
		pl.seg_len += g.chunk;
		pl.num_families += g.nfam;
		pl.groups.push_back(g);
	}
	/* DP option table: one slot per block on the equal-categories path, one per candidate otherwise */
	for (int c = 0; c < N; c++) {
		CnbNodePlan &nd = pl.nodes[c];
		nd.opt_begin = pl.num_options;
		for (int b = nd.first_blk; b < nd.first_blk + nd.nblk; b++) {
			pl.blocks[b].opt_off = pl.num_options;
			pl.num_options += nd.equal_branch ? 1 : pl.blocks[b].ncomb;
		}
		nd.opt_end = pl.num_options;
	}
	return CNB_OK;
#undef FAIL
}

/* host-side view of the plan's tiling (tile_base / dtile_off point into the plan's vectors) */
void cnb_plan_tiles(const CnbPlan &pl, CnbTiles *T) {
	memset(T, 0, sizeof(*T));
	T->tile_base = (const long long *)pl.tile_base.data();
	T->dtile_off = (const long long *)pl.dtile_off.data();
	T->nct = pl.tile_base.empty() ? 0 : (int)pl.tile_base.size() - 1;
	T->child = pl.tile_child;
	T->world = 1;
	T->tp = pl.tile_pairs;
	T->tc = pl.tile_cols;
	T->fr = pl.tile_fr;
	T->full = pl.tile_full;
	T->row_stride = 4;
}

int cnb_tile_columns(const CnbPlan &pl, int64_t tile) {
	CnbTiles T;
	cnb_plan_tiles(pl, &T);
	CnbTileInfo ti;
	cnb_tile_decode(T, pl.N, tile, ti);
	return ((pl.tile_full ? 4 : pl.tile_fr) * (ti.ze - ti.zb) + 15) / 16 * 16;
}

/* test hook: the tiling of the unrestricted two-parent group of N nodes is a bijection between families and tile
 * slots (family -> slot -> family), every tile has at least one column, and no two families share a slot.
 * Returns the number of tiles, or a negative value at the first inconsistency. */
static int64_t tile_selftest(int32_t N, int tile_mode, int ncats = 2) {
	if (N < 3) return 0;
	cnb_params q;
	memset(&q, 0, sizeof(q));
	q.num_nodes = N;
	q.num_samples = 1;
	q.max_parent_set = 2;
	q.max_complexity = 1 << 30;
	std::vector<int32_t> cats(N, ncats);
	CnbPlan pl;
	if (cnb_build_plan(&q, cats.data(), 0, 1, &pl, nullptr, tile_mode) != CNB_OK || pl.tile_base.empty()) return -1;
	CnbTiles T;
	cnb_plan_tiles(pl, &T);
	const int tp = T.tp, tc = T.tc;
	if (tile_mode == 3 && (tp != cnb_geo_pairs(ncats - 1) || tc != cnb_geo_child(ncats - 1) || T.fr != ncats - 1)) return -8;
	const int64_t ntiles = pl.tile_base[T.nct];
	std::vector<unsigned char> used((size_t)ntiles * tp * tc, 0);
	int64_t nfam = 0;
	for (int c = 2; c < N; c++)
		for (int b = 1; b < c; b++)
			for (int a = 0; a < b; a++) {
				long long tile;
				int within;
				cnb_family_slot(T, N, a, b, c, tile, within);
				if (tile < 0 || tile >= ntiles || within < 0 || within >= tp * tc) return -2;
				unsigned char &u = used[(size_t)tile * tp * tc + within];
				if (u) return -3;
				u = 1;
				CnbTileInfo ti;
				cnb_tile_decode(T, N, tile, ti);
				int x, y;
				const int p = within / tc, j = within % tc;
				if (j >= ti.ze - ti.zb || !cnb_tile_pair(ti, p, x, y)) return -4;
				const bool dk = ti.kind != 0;
				const int a2 = dk ? ti.zb + j : x, b2 = dk ? x : y, c2 = dk ? y : ti.zb + j;
				if (a2 != a || b2 != b || c2 != c) return -5;
				nfam++;
			}
	if (nfam != (int64_t)N * (N - 1) * (N - 2) / 6) return -6;
	for (int64_t t = 0; t < ntiles; t++) {
		CnbTileInfo ti;
		cnb_tile_decode(T, N, t, ti);
		if (ti.ze <= ti.zb || ti.ze - ti.zb > tc || cnb_tile_columns(pl, t) < 16) return -7;
	}
	return ntiles;
}

extern "C" int64_t cnb_tile_selftest(int32_t N) { return tile_selftest(N, 1); }
/* the same for the full-table geometry (16 pairs x 48 column nodes) */
extern "C" int64_t cnb_tile_selftest_full(int32_t N) { return tile_selftest(N, 2); }
/* ... and for the reduced inclusion-exclusion geometries of data with `free_cats` + 1 categories per node (free_cats = 1, 2, 3) */
extern "C" int64_t cnb_tile_selftest_geo(int32_t N, int32_t free_cats) {
	if (free_cats < 1 || free_cats > 3) return -9;
	return tile_selftest(N, 3, free_cats + 1);
}

extern "C" int64_t cnb_num_families(const cnb_params *p) {
	if (!p || p->num_nodes < 1) return -1;
	std::vector<int32_t> cats(p->num_nodes, 2);
	if (p->num_cats) cats.assign(p->num_cats, p->num_cats + p->num_nodes);
	CnbPlan pl;
	cnb_params q = *p;
	if (q.num_samples < 1) q.num_samples = 1;
	if (cnb_build_plan(&q, cats.data(), 0, 1, &pl, nullptr) != CNB_OK) return -1;
	return pl.num_families;
}

/* test hook: the T tiling of the unrestricted three-parent group (cnb_tiles.h, kind 3) is consistent between the side that
 * builds the tiles (cnb_tile_decode / cnb_tile_pair, as the table kernel uses them) and the side that reads a family's
 * slots (the arithmetic of k_umma_epilogue3): every family (a < b < c | x) and value i < f1 of a finds, in the column tile
 * of x, the slot whose "pair" is (virtual node of (a, b, i), c) and whose column is x, no two share a slot.  Returns the
 * number of tiles, negative at the first inconsistency. */
extern "C" int64_t cnb_tile3_selftest(int32_t N, int32_t f1) {
	/* f1 > 0: the geometry of f1 free categories (what the library runs); f1 < 0: -f1 free values in the default 28 x 64 tiles */
	const int geo = f1 < 0 ? 3 : f1;
	if (f1 < 0) f1 = -f1;
	if (N < 4 || f1 < 1 || f1 > 3) return 0;
	const int TP = cnb_geo_pairs(geo), TCH = cnb_geo_child(geo);
	const int nct = (N + TCH - 1) / TCH;
	std::vector<long long> base3(nct + 1, 0);
	for (int ct = 0; ct < nct; ct++) base3[ct + 1] = base3[ct] + cnb_t_tiles(cnb_t_ce(N, nct, ct, TCH), f1, TP);
	CnbTiles T;
	memset(&T, 0, sizeof(T));
	T.tp = TP;
	T.tc = TCH;
	T.fr = geo;
	T.tile_base = base3.data();
	T.nct = nct;
	T.child = TCH;
	T.world = 1;
	T.self_pairs = 2;
	T.f1 = f1;
	T.nvirt = N * (N - 1) / 2 * f1;
	std::vector<unsigned char> used((size_t)base3[nct] * TP * TCH, 0);
	for (int x = 3; x < N; x++) {
		int ct = 0;
		while (cnb_t_ce(N, nct, ct, TCH) <= x) ct++;
		const int c0 = cnb_t_c0(N, nct, ct, TCH), ce = cnb_t_ce(N, nct, ct, TCH);
		if (x < c0 || x >= ce) return -1;
		for (int c = 2; c < x; c++)
			for (int b = 1; b < c; b++)
				for (int a = 0; a < b; a++)
					for (int i = 0; i < f1; i++) {
						const long long t = cnb_binom3(c) + (long long)b * (b - 1) / 2 + a, q = t * f1 + i;
						const long long tile = base3[ct] + q / TP;
						if (tile >= base3[ct + 1]) return -2;
						unsigned char &u = used[((size_t)tile * TP + q % TP) * TCH + (x - c0)];
						if (u) return -3;
						u = 1;
						CnbTileInfo ti;
						cnb_tile_decode(T, N, tile, ti);
						int px, py;
						if (ti.kind != 3 || ti.zb != c0 || ti.ze != ce || !cnb_tile_pair(ti, (int)(q % TP), px, py)) return -4;
						if (py != c || px != N + (int)(((long long)b * (b - 1) / 2 + a) * f1 + i)) return -5;
					}
	}
	return base3[nct];
}
