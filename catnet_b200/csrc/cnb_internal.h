/* cnb_internal.h -- structures shared by the host planner, the CUDA kernels and the C ABI. */
#ifndef CNB_INTERNAL_H
#define CNB_INTERNAL_H

#include <stdint.h>
#include <float.h>
#include <string>
#include <vector>
#include "../../include/catnet_b200.h"

#define CNB_MAXP 8            /* max parents of one family handled by the kernels */
#define CNB_NEG_INF (-(double)FLT_MAX)   /* the reference's "-infinity" (catnet_search2.h:559) */
#define CNB_PLANE_CATS 4      /* bit-plane layout holds up to 4 categories per node (one uint4 per 32 samples) */

/* One candidate block = all parent sets of one size for one child (HOT LOOP 2/3, catnet_search2.h:509-553):
 * the nfix fixed parents first, then every increasing (d - nfix)-subset of the free pool, lexicographic. */
struct CnbBlock {
	int32_t child;      /* order position */
	int32_t d;          /* total parents */
	int32_t nfix;
	int32_t nfree;
	int32_t fix_off;    /* into lists[]: fixed parents (positions, ascending) */
	int32_t free_off;   /* into lists[]: free pool (positions, ascending) */
	int32_t group;      /* index of the group (same d) this block belongs to */
	int32_t pad;
	int64_t ncomb;      /* C(nfree, d - nfix) */
	int64_t start;      /* first family id of the block inside its group */
	int64_t opt_off;    /* first option slot of this block (DP option table) */
};

/* All blocks with the same number of parents: one kernel launch class. */
struct CnbGroup {
	int32_t d;
	int32_t nblk;
	int32_t blk_off;    /* into group_blocks[] (block ids sorted by start) */
	int32_t tiled;      /* 1, 2: scored by the tensor-core kernel; scores are addressed by tile slot, not family id (1: the group
	                       is the unrestricted one, slot from the candidate's rank; 2: a subset of it, slot from its parents) */
	int64_t nfam;       /* families in the group */
	int64_t chunk;      /* score slots per rank: ceil(nfam / world), or tiles_per_rank * slots per tile when tiled */
	int64_t seg_off;    /* offset of the group inside one rank segment of the score buffer */
	int64_t ntiles;     /* tiled only: tiles of the group; tile t belongs to rank t % world (local index t / world) */
	int64_t tiles_per_rank;   /* ceil(ntiles / world): local tile slots of every rank segment */
	int32_t tile_world;
	int32_t tile_child; /* children per child tile (<= CNB_TILE_CHILD, the slot stride) */
	int32_t full;       /* every block offers all d-subsets of the child's predecessors (no pools, fixed parents, limits) */
	int32_t pad2;
};

/* per-node arguments of the DP recurrence (k_dp_all) */
struct CnbDpNode {
	int64_t opt_begin, opt_end;
	int32_t base_cx, pad;
};

/* tensor-core tiling handed to the kernels of cnb_umma.cu and to k_select */
struct CnbTiles {
	const long long *tile_base;   /* [nct + 1] first tile of every child tile */
	const long long *dtile_off;   /* [nct][CNB_DT_STRIDE] D tiles of the child tile before each of its D pair tiles */
	int32_t nct, child;           /* child tiles, children per child tile */
	int32_t rank, world;
	int32_t self_pairs;           /* 1: the pair-table tiling instead (cnb_tiles.h, kind 2); tile_base/dtile_off unused;
	                                 2: the three-parent tiling (kind 3): tile_base = first tile of every column tile */
	int32_t pad;                  /* table kernel: L1 prefetch distance of the generators */
	int32_t f1;                   /* kind 3: free categories per node (max categories - 1) */
	int32_t nvirt;                /* kind 3: virtual pair nodes behind the N real ones in the A operand rows */
	int32_t tp, tc;               /* pair slots x column slots of a tile: 28 x 64 (inclusion-exclusion tiles, P and T tiles) or
	                                 16 x 48 (full-table tiles); slot stride inside a tile = tc */
	int32_t full;                 /* 1: full-table tiles -- 16 rows per pair (all (i, j)), 4 columns per column node, the child's
	                                 keep mask ANDed into its rows, a missing value sets no bit: no inclusion-exclusion, so
	                                 missing values and perturbation masks count correctly */
	int32_t row_stride;           /* full: operand rows per node, 4 (no masks) or 8 (4 plain, then 4 under the node's own keep mask) */
	int32_t fr;                   /* inclusion-exclusion tiles: free categories contracted per node -- 3 (data with up to four
	                                 categories per node: 9 rows per pair, 3 columns per column node, 28 x 64 slots), 2 (up to
	                                 three categories: 4 rows, 2 columns, 64 x 96 slots) or 1 (binary data: 1 row, 1 column,
	                                 256 x 192 slots); cnb_geo_* in cnb_tiles.h */
};
#define CNB_FULL_PAIRS 16     /* full-table tile: 2 blocks of 8 parent pairs x 16 rows ... */
#define CNB_FULL_CHILD 48     /* ... x up to 48 column nodes x 4 columns (192): 768 family slots per tile */

#define CNB_TILE_PAIRS 28     /* tensor-core tile: 2 blocks of 14 parent pairs x 9 rows (2 x 126 of 128) ... */
#define CNB_TILE_CHILD 64     /* ... x up to 64 children x 3 columns (192): 1792 family slots per tile */

struct CnbNodePlan {
	int32_t equal_branch;   /* catnet_search2.h:491-496 */
	int32_t nfix, fix_off;
	int32_t nfree, free_off;/* the free pool in lists[] (also when the node has no candidate block) */
	int32_t base_cx;        /* (C-1) * prod C_fixed */
	int32_t first_blk, nblk;/* blocks of this node, d ascending */
	int64_t opt_begin, opt_end;  /* DP options of the node */
};

struct CnbPlan {
	int32_t N = 0, M = 0, P = 0, K = 0;        /* K = max complexity (slots 0..K) */
	int32_t rank = 0, world = 1;
	bool has_prior = false;
	std::vector<int32_t> order;                /* position -> 0-based label */
	std::vector<int32_t> fix1;                 /* [N] the node's only fixed parent (position), or -1: a two-parent candidate (fixed f, free x)
	                                              is LISTED [f, x] even when x precedes f, and its table and log-likelihood follow the listing */
	std::vector<int32_t> cats;                 /* order space */
	std::vector<int32_t> lists;
	std::vector<CnbBlock> blocks;
	std::vector<CnbGroup> groups;
	std::vector<int32_t> group_blocks;
	std::vector<CnbNodePlan> nodes;
	std::vector<double> edge;                  /* order space [parent*N + child], empty if no prior */
	std::vector<int64_t> tile_base;            /* tiled group: first tile of every child tile (+ total at the end) */
	std::vector<int64_t> dtile_off;            /* see cnb_tiles.h */
	int32_t tile_child = 0;                    /* children per child tile */
	int32_t tile_pairs = CNB_TILE_PAIRS, tile_cols = CNB_TILE_CHILD;   /* geometry of the tiled group (CnbTiles.tp / .tc) */
	int32_t tile_full = 0;                     /* the tiled group runs as full-table tiles */
	int32_t tile_fr = 3;                       /* free categories per node the inclusion-exclusion tiles contract (CnbTiles.fr) */
	int64_t seg_len = 0;                       /* doubles per rank segment */
	int64_t num_families = 0;
	int64_t num_options = 0;
	int32_t base_complexity = 0;
	int32_t max_cats = 0;
	int32_t max_d = 0;
};

/* host planner (no CUDA): cnb_plan.cpp */
/* tile_mode: 0 = the two-parent group is not tiled, 1 = inclusion-exclusion tiles (complete data), 2 = full-table tiles */
int cnb_build_plan(const cnb_params *p, const int32_t *cats_lbl, int32_t rank, int32_t world, CnbPlan *plan,
		std::string *err, int tile_mode = 0);
int64_t cnb_binom(int64_t n, int64_t k);
/* MMA columns (children x 3, rounded up to 16) tile `tile` of the tiled group really contracts */
int cnb_tile_columns(const CnbPlan &pl, int64_t tile);
/* host-side view of the plan's tiling: geometry, mode, tile_base / dtile_off pointing into the plan's vectors */
void cnb_plan_tiles(const CnbPlan &pl, CnbTiles *T);
int cnb_unrank_lex(int32_t n, int32_t k, int64_t r, int32_t *out);

void cnb_set_error(const char *fmt, ...);

#endif
