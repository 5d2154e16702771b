/* cnb_tiles.h -- geometry of the tensor-core tiling of the two-parent group, shared by the host planner
 * (cnb_plan.cpp), the table kernel and its epilogue (cnb_umma.cu) and the selection kernel (cnb_kernels.cu).
 *
 * Families (a < b | c) with a, b < c in order space; the children are cut into child tiles of T.child <= 64
 * consecutive positions [c0, ce).  Inside child tile ct there are two kinds of MMA tiles, both 28 "pair" slots
 * (2 blocks x 14 pairs x 9 configurations = the A rows) by up to 64 "column" nodes (x 3 categories = the B rows):
 *
 *   F tiles  pairs (a, b) with b < c0, ranked q = b(b-1)/2 + a; columns = the children c0..ce-1.  Every slot is a
 *            family.
 *   D tiles  the "diagonal" families whose second parent lies inside the child tile: pairs (b, c) with
 *            c0 <= b < c < ce, ranked by b then c; columns = the FIRST parent a, in chunks of 64 positions,
 *            as far as the largest b of the tile's pairs.  (With (a, b) as the pair and c > b as the columns, as in
 *            the F tiles, a diagonal tile would use half its columns on average and every MMA has a fixed cost.)
 *
 *   P tiles  (T.self_pairs, used once per data set for the two-way tables of ALL node pairs, label space): "pair"
 *            slot q is the single node q (both members of the pair are q, so only the 3 configurations (i, i) of
 *            its 9 rows are non-zero and row (i, i) is the plain one-hot row of category i); columns = all nodes in
 *            chunks of 64.  Tile t = (pair tile t / nchunks, chunk t % nchunks).
 *
 *   T tiles  (T.self_pairs == 2, three-parent families (a < b < c | x)): columns = the children x of column tile ct; the
 *            column tiles are aligned to the END of the order ([N - 64, N), [N - 128, N - 64), ... the first one takes
 *            the rest), because a column tile [c0, ce) costs C(ce - 1, 3) f1 / 28 tiles whatever its width; "pair" slot q
 *            of that column tile = (triple t = q / f1, value i = q % f1 of a) with the triples (a < b < c), c < ce - 1, ranked t = C(c,3) + C(b,2) + a; its two "pair" nodes are the VIRTUAL node
 *            N + (b(b-1)/2 + a) f1 + i -- whose three operand rows are [x_a = i][x_b = j], j < 3 -- and the real node c,
 *            so the slot's 9 x 3 cells are n(a = i, b = j, c = k, x = l).  tile_base[ct] = first tile of column tile ct.
 *            The slots of children x <= c hold no family (the diagonal is not transposed here).
 *
 * Tile ids: tile_base[ct] + [0, nF) are the F tiles of ct, then the D tiles pair tile after pair tile, chunk after
 * chunk; dtile_off[ct * CNB_DT_STRIDE + p] = D tiles of ct before pair tile p.
 */
#ifndef CNB_TILES_H
#define CNB_TILES_H
#include "cnb_internal.h"
#include <math.h>

#ifdef __CUDACC__
#define CNB_HD __host__ __device__ __forceinline__
#else
#define CNB_HD inline
#endif

#define CNB_DT_STRIDE 74      /* >= ceil(64 * 63 / 2 / 28) + 1 (default geometry) and >= ceil(48 * 47 / 2 / 16) + 1 = 72 (full-table
                                 geometry) pair tiles of D pairs per child tile, plus the total */
/* default geometry of a CnbTiles (the inclusion-exclusion tiles, the P and the T tiles) */
#define CNB_TILES_DEFAULT_GEOMETRY(T) do { (T).tp = CNB_TILE_PAIRS; (T).tc = CNB_TILE_CHILD; (T).fr = 3; } while (0)
/* Geometry of the inclusion-exclusion tiles by the number fr of free categories per node (largest category count - 1):
 * a pair takes fr^2 of the 128 rows of a block, a column node fr of the 192 columns, a slot holds fr^2 configurations x fr
 * counts (padded to int4 at fr = 3).  With fewer categories the same MMA covers more families: 28 x 64 = 1,792 slots at
 * fr = 3, 64 x 96 = 6,144 at fr = 2 (3.4 x fewer MACs per family), 256 x 192 = 49,152 at fr = 1. */
CNB_HD int cnb_geo_fr(int max_cats) { return max_cats >= 4 ? 3 : (max_cats == 3 ? 2 : 1); }
CNB_HD int cnb_geo_pairs(int fr) { return fr >= 3 ? CNB_TILE_PAIRS : (fr == 2 ? 64 : 256); }
CNB_HD int cnb_geo_child(int fr) { return fr >= 3 ? CNB_TILE_CHILD : (fr == 2 ? 96 : 192); }
CNB_HD int cnb_geo_cst(int fr) { return fr >= 3 ? 4 : fr; }                 /* ints stored per (configuration, column node) */
CNB_HD int cnb_geo_slot_ints(int fr) { return fr * fr * cnb_geo_cst(fr); }  /* 36, 8, 1 */

struct CnbTileInfo {
	int kind;                 /* 0 = F, 1 = D, 2 = P, 3 = T */
	int tp, tc;               /* pair slots and column slots of a tile (CnbTiles) */
	int n, f1;                /* kind 3: real nodes, free categories */
	int ct, c0, ce;           /* child tile and its children [c0, ce) */
	long long pair0;          /* rank of the tile's first pair inside its region (F: q, D: rank among the D pairs) */
	long long npairs;         /* pairs of the region */
	int zb, ze;               /* column nodes [zb, ze): children (F) or first parents (D); slot column = node - zb */
};

CNB_HD void cnb_pair_of(long long q, int &a, int &b) {      /* q = b(b-1)/2 + a, a < b */
	b = (int)((1.0 + sqrt(1.0 + 8.0 * (double)q)) * 0.5);
	while ((long long)b * (b - 1) / 2 > q) b--;
	while ((long long)(b + 1) * b / 2 <= q) b++;
	a = (int)(q - (long long)b * (b - 1) / 2);
}
CNB_HD long long cnb_f_pairs(int c0) { return (long long)c0 * (c0 - 1) / 2; }
CNB_HD long long cnb_f_tiles(int c0, int tp) { return (cnb_f_pairs(c0) + tp - 1) / tp; }
CNB_HD int cnb_d_pairs(int c0, int ce) { return (ce - c0) * (ce - c0 - 1) / 2; }
/* rank of D pair (b, c) of child tile [c0, ce), and back */
CNB_HD int cnb_d_rank(int c0, int ce, int b, int c) {
	const int k = b - c0;                                  /* pairs with a smaller b: sum_{t<k} (w - 1 - t) */
	return k * (ce - c0 - 1) - k * (k - 1) / 2 + (c - b - 1);
}
CNB_HD void cnb_d_pair(int c0, int ce, int q, int &b, int &c) {
	b = c0;
	while (q >= ce - 1 - b) { q -= ce - 1 - b; b++; }
	c = b + 1 + q;
}
/* chunks of tc first parents a D pair tile (tp pairs) needs: up to the b of its last pair */
CNB_HD int cnb_d_chunks(int c0, int ce, int ptile, int tp, int tc) {
	const int np = cnb_d_pairs(c0, ce);
	int last = ptile * tp + tp - 1;
	if (last > np - 1) last = np - 1;
	int b, c;
	cnb_d_pair(c0, ce, last, b, c);
	return (b + tc - 1) / tc;
}

CNB_HD long long cnb_binom3(long long n) { return n < 3 ? 0 : n * (n - 1) * (n - 2) / 6; }
/* kind 3: children [c0, ce) of column tile ct of nct = ceil(N / tc), aligned to the end of the order */
CNB_HD int cnb_t_ce(int N, int nct, int ct, int tc) { return N - tc * (nct - 1 - ct); }
CNB_HD int cnb_t_c0(int N, int nct, int ct, int tc) { return ct == 0 ? 0 : N - tc * (nct - ct); }
/* t = C(c,3) + C(b,2) + a, a < b < c */
CNB_HD void cnb_triple_of(long long t, int &a, int &b, int &c) {
	c = (int)cbrt(6.0 * (double)t) + 2;
	while (cnb_binom3(c) > t) c--;
	while (cnb_binom3(c + 1) <= t) c++;
	cnb_pair_of(t - cnb_binom3(c), a, b);
}

CNB_HD void cnb_tile_decode(const CnbTiles &T, int N, long long tile, CnbTileInfo &ti) {
	ti.n = N;
	ti.f1 = T.f1;
	ti.tp = T.tp;
	ti.tc = T.tc;
	if (T.self_pairs == 2) {
		int ct = 0;
		while (ct + 1 < T.nct && T.tile_base[ct + 1] <= tile) ct++;
		ti.kind = 3;
		ti.ct = ct;
		ti.c0 = cnb_t_c0(N, T.nct, ct, T.tc);
		ti.ce = cnb_t_ce(N, T.nct, ct, T.tc);
		ti.pair0 = (tile - T.tile_base[ct]) * T.tp;
		ti.npairs = cnb_binom3(ti.ce - 1) * T.f1;
		ti.zb = ti.c0;
		ti.ze = ti.ce;
		return;
	}
	if (T.self_pairs) {
		const int nch = (N + T.tc - 1) / T.tc;
		ti.kind = 2;
		ti.ct = 0;
		ti.c0 = 0;
		ti.ce = N;
		ti.pair0 = (tile / nch) * T.tp;
		ti.npairs = N;
		ti.zb = (int)(tile % nch) * T.tc;
		ti.ze = ti.zb + T.tc < N ? ti.zb + T.tc : N;
		return;
	}
	int lo = 0, hi = T.nct - 1;
	while (lo < hi) {
		int mid = (lo + hi + 1) >> 1;
		if (T.tile_base[mid] <= tile) lo = mid;
		else hi = mid - 1;
	}
	ti.ct = lo;
	ti.c0 = lo * T.child;
	ti.ce = ti.c0 + T.child < N ? ti.c0 + T.child : N;
	const long long pt = tile - T.tile_base[lo], nf = cnb_f_tiles(ti.c0, T.tp);
	if (pt < nf) {
		ti.kind = 0;
		ti.pair0 = pt * T.tp;
		ti.npairs = cnb_f_pairs(ti.c0);
		ti.zb = ti.c0;
		ti.ze = ti.ce;
		return;
	}
	const long long *off = T.dtile_off + (size_t)lo * CNB_DT_STRIDE;
	const int d = (int)(pt - nf);
	int p = 0;
	while (off[p + 1] <= d) p++;
	ti.kind = 1;
	ti.pair0 = (long long)p * T.tp;
	ti.npairs = cnb_d_pairs(ti.c0, ti.ce);
	int last = p * T.tp + T.tp - 1;
	if (last > ti.npairs - 1) last = (int)ti.npairs - 1;
	int b, c;
	cnb_d_pair(ti.c0, ti.ce, last, b, c);
	ti.zb = (d - (int)off[p]) * T.tc;
	ti.ze = ti.zb + T.tc < b ? ti.zb + T.tc : b;
}

/* the two "pair" nodes (x, y) of pair slot ps of a tile; false if the slot is past the end of the region */
CNB_HD bool cnb_tile_pair(const CnbTileInfo &ti, int ps, int &x, int &y) {
	const long long q = ti.pair0 + ps;
	if (q >= ti.npairs) return false;
	if (ti.kind == 0) cnb_pair_of(q, x, y);
	else if (ti.kind == 1) cnb_d_pair(ti.c0, ti.ce, (int)q, x, y);
	else if (ti.kind == 2) x = y = (int)q;
	else {
		int a, b, c;
		cnb_triple_of(q / ti.f1, a, b, c);
		x = ti.n + (int)(((long long)b * (b - 1) / 2 + a) * ti.f1 + q % ti.f1);
		y = c;
	}
	return true;
}

/* family (a < b | c) -> tile and slot inside the tile (pair slot * T.tc + column slot) */
CNB_HD void cnb_family_slot(const CnbTiles &T, int N, int a, int b, int c, long long &tile, int &within) {
	const int ct = c / T.child, c0 = ct * T.child, ce = c0 + T.child < N ? c0 + T.child : N;
	if (b < c0) {
		const long long q = (long long)b * (b - 1) / 2 + a;
		tile = T.tile_base[ct] + q / T.tp;
		within = (int)(q % T.tp) * T.tc + (c - c0);
	} else {
		const int q = cnb_d_rank(c0, ce, b, c), p = q / T.tp;
		tile = T.tile_base[ct] + cnb_f_tiles(c0, T.tp) + T.dtile_off[(size_t)ct * CNB_DT_STRIDE + p] + a / T.tc;
		within = (q % T.tp) * T.tc + a % T.tc;
	}
}

/* kind 3: tiles of column tile [c0, ce) and the slot of (triple rank t, value i of a, child x) */
CNB_HD long long cnb_t_tiles(int ce, int f1, int tp) { return (cnb_binom3(ce - 1) * f1 + tp - 1) / tp; }

#endif
