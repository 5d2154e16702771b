/* cnb_umma.cu -- two-parent family tables on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a only.
 *
 * The contingency table of family (a, b | c) is a contraction over the sample axis:
 *     n(i, j, k) = sum_s  [x_a(s) = i][x_b(s) = j] * [x_c(s) = k]
 * i.e. D = A * B^T with A = one-hot rows of the pair configurations (pair (a,b) x configurations (i,j)) and
 * B = one-hot rows of the child categories (child c x categories k), K = samples.  Both operands are 0/1 values
 * generated ON THE FLY from one-hot bit rows (HBM/L2 traffic stays at one bit per sample and category), staged in
 * shared memory in the K-major SWIZZLE_128B canonical layout and multiplied with
 *     tcgen05.mma.kind::mxf4.block_scale   (E2M1 x E2M1 -> f32, unit E8M0 scales; 256 samples per 128-byte row), or
 *     tcgen05.mma.kind::i8                 (u8 x u8 -> s32; 128 samples per 128-byte row; test/fallback kind).
 * Every product is exactly 1 (mxf4) or 128 (i8), so the TMEM accumulators hold exact counts (M < 2^24).
 *
 * One CTA per SM (all 512 TMEM columns): 16 operand-generator warps (one thread per operand row), one MMA-issuing
 * warp, a 3-stage shared-memory ring handed over with mbarriers (generators -> full[s] -> MMA -> tcgen05.commit ->
 * empty[s]).  A tile is 2 blocks of 14 parent pairs (2 x 126 of 128 rows, two accumulators) x 80 children
 * (240 columns): the child rows of a stage are generated once and consumed by both pair blocks.
 *
 * Replaces, for unrestricted two-parent candidate sets without missing values, the counting loop of
 * CATNET::setNodeSampleProb (/root/reference/src/catnet_class.h:842-857); the log-lik epilogue
 * (problist.h:224-250) runs in k_umma_epilogue over the integer tables, in the reference's operation order.
 */
#include <stdio.h>
#include "cnb_device.cuh"
#include "cnb_kernels.h"

/* Inclusion-exclusion (see k_score_planes_ie2): only the 3 x 3 x 3 "free" cells of every table are contracted;
 * cells with a last category follow from the pair tables.  So a pair needs 9 rows and a child 3 columns. */
#define UM_NA 2                           /* pair blocks (accumulators) per tile */
#define UM_BLK_PAIRS 14                   /* 14 pairs x 9 rows = 126 of the 128 rows of one MMA */
#define UM_PAIRS CNB_TILE_PAIRS
#define UM_CHILD CNB_TILE_CHILD
#define UM_SLOTS (UM_PAIRS * UM_CHILD)
#define UM_SLOT_INTS 36                   /* 9 configurations x (3 counts + 1 pad): int4 stores */
#define UM_N (UM_CHILD * 3)               /* 240 = MMA N */
#define UM_A_BYTES (128 * 128)            /* one pair block of a stage: 128 rows x 128 B */
#define UM_B_OFF (UM_NA * UM_A_BYTES)
#define UM_STAGE (UM_B_OFF + UM_N * 128)  /* 63,488 B */
#define UM_STAGES 3
#define UM_SMEM (UM_STAGES * UM_STAGE + 1024)
#define UM_GEN_WARPS 16                   /* warps 0-7: pair rows (2 x 128), warps 8-15: child rows (240 of 256) */
#define UM_THREADS ((UM_GEN_WARPS + 1) * 32)
#define UM_TMEM_COLS 512
#define UM_TMEM_SF 480                    /* mxf4: unit scale factors in columns 480..511 */
static_assert(UM_PAIRS == UM_NA * UM_BLK_PAIRS, "tile geometry");
static_assert(UM_N % 16 == 0 && UM_N <= 256 && UM_NA * UM_N <= UM_TMEM_SF, "tile geometry");
static_assert(UM_STAGE % 1024 == 0, "stages keep the 1024-byte swizzle alignment");

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
	uint32_t ok = 0, addr = smem_u32(bar);
	while (!ok)
		asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
				: "=r"(ok) : "r"(addr), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
	asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
/* D[tmem] (+)= A[smem desc] * B[smem desc]^T, u8 x u8 -> s32, M = 128, K = 32 */
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
	uint32_t z = 0;
	asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
			"tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n"
			::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(z) : "memory");
}
/* D[tmem] (+)= (A * sfa) * (B * sfb)^T, E2M1 x E2M1 -> f32, one E8M0 scale per 32 elements, M = 128, K = 64 */
__device__ __forceinline__ void umma_mxf4(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate,
		uint32_t tmem_sfa, uint32_t tmem_sfb) {
	asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
			"tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n}\n"
			::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb) : "memory");
}
/* K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 in [0,14),
 * LBO = 1 in [16,30), SBO = 1024 B >> 4 in [32,46) (stride between 8-row groups), version 1 in [46,48),
 * layout SWIZZLE_128B = 2 in [61,64) */
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
	return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void pair_of(long long q, int &a, int &b) {      /* q = b(b-1)/2 + a, a < b */
	b = (int)((1.0 + sqrt(1.0 + 8.0 * (double)q)) * 0.5);
	while ((long long)b * (b - 1) / 2 > q) b--;
	while ((long long)(b + 1) * b / 2 <= q) b++;
	a = (int)(q - (long long)b * (b - 1) / 2);
}

/* tile id -> (child tile, pair tile); tile_base[ct] = first tile of child tile ct */
__device__ __forceinline__ void tile_of(const long long *__restrict__ tile_base, int nct, long long tile, int &ct, long long &pt) {
	int lo = 0, hi = nct - 1;
	while (lo < hi) {
		int mid = (lo + hi + 1) >> 1;
		if (tile_base[mid] <= tile) lo = mid;
		else hi = mid - 1;
	}
	ct = lo;
	pt = tile - tile_base[lo];
}

/* Children a tile really contracts.  Pairs are ranked by their larger member b first, so the 28 pairs of a tile
 * share (almost) one b; only children behind the tile's smallest b can form a family, which trims the "diagonal"
 * tiles (b inside the child tile) to the columns that count.  [cb, ce) = child positions, c0 = first of the
 * child tile (slot origin).  Host mirror: cnb_tile_columns (cnb_plan.cpp). */
__device__ __forceinline__ void tile_children(const CnbTiles &T, int N, int ct, long long pt, int &c0, int &cb, int &ce) {
	int a, b;
	pair_of(pt * UM_PAIRS, a, b);
	c0 = ct * T.child;
	ce = min(c0 + T.child, N);
	cb = max(c0, b + 1);
}

/* ---- operand rows --------------------------------------------------------------------------------------------
 * k_umma_rows turns the order-space bit planes [word][node] (uint4 = the 4 categories of one 32-sample word) into
 * operand bit rows [quad][node*3 + cat] (uint4 = 4 consecutive 32-sample words = 128 samples of one (node, cat < 3)
 * row): the generator threads of a warp own consecutive rows, so their 16-byte loads of a stage coalesce into a
 * few 128-byte lines.  rows_a holds the plain bits (bit s of word w = sample 32w+s), rows_b the same samples with
 * the bits permuted so that both operands are produced by one AND per output word:
 *   mxf4: K position (word j < 4, nibble i < 8) of a 32-sample source word holds sample 4i+j.
 *         A nibble = x & (1<<j) for j < 3 (E2M1 0.5, 1, 2) and (x>>1) & 4 for j = 3 (2);
 *         B nibble from y' (nibble bits 0 and 2 swapped) = y' & (4>>j) for j < 3 (2, 1, 0.5) and (y'>>3) & 1 (0.5):
 *         every product is bit * bit * 1.
 *   i8:   K position (word j < 8, byte i < 4) holds sample 8i+j; A byte = x & (1<<j) = bit * 2^j, B byte from y'
 *         (bits reversed inside every byte) = y' & (0x80>>j) = bit * 2^(7-j): every product is bit * bit * 128.
 * Quads past the last sample word are zero (whole stages + one stage of prefetch slack). */
__device__ __forceinline__ uint32_t rows_b_word(uint32_t y, int fp4) {
	if (fp4) return ((y & 0x11111111u) << 2) | (y & 0x22222222u) | ((y & 0x44444444u) >> 2) | (y & 0x88888888u);
	return __byte_perm(__brev(y), 0, 0x0123);
}

__global__ void __launch_bounds__(128) k_umma_rows(const uint4 *__restrict__ planes, int N, int W, int nquads,
		uint4 *__restrict__ rows_a, uint4 *__restrict__ rows_b, int fp4) {
	const int n = blockIdx.x * blockDim.x + threadIdx.x, qd = blockIdx.y;
	if (n >= N) return;
	uint4 v[4];
#pragma unroll
	for (int k = 0; k < 4; k++) {
		const int w = qd * 4 + k;
		v[k] = w < W ? __ldg(planes + (size_t)w * N + n) : make_uint4(0, 0, 0, 0);
	}
	const size_t o = (size_t)qd * (3 * N) + 3 * n;
	rows_a[o + 0] = make_uint4(v[0].x, v[1].x, v[2].x, v[3].x);
	rows_a[o + 1] = make_uint4(v[0].y, v[1].y, v[2].y, v[3].y);
	rows_a[o + 2] = make_uint4(v[0].z, v[1].z, v[2].z, v[3].z);
	rows_b[o + 0] = make_uint4(rows_b_word(v[0].x, fp4), rows_b_word(v[1].x, fp4), rows_b_word(v[2].x, fp4), rows_b_word(v[3].x, fp4));
	rows_b[o + 1] = make_uint4(rows_b_word(v[0].y, fp4), rows_b_word(v[1].y, fp4), rows_b_word(v[2].y, fp4), rows_b_word(v[3].y, fp4));
	rows_b[o + 2] = make_uint4(rows_b_word(v[0].z, fp4), rows_b_word(v[1].z, fp4), rows_b_word(v[2].z, fp4), rows_b_word(v[3].z, fp4));
}

/* ---- operand generation: one 32-sample source word -> output words ------------------------------------------- */
#define STS128(addr, imm, a, b, c, d) asm volatile("st.shared.v4.b32 [%0+%1], {%2, %3, %4, %5};" \
		::"r"(addr), "n"(imm), "r"(a), "r"(b), "r"(c), "r"(d) : "memory")

/* mxf4: source word -> one 16-byte chunk */
template <bool B_SIDE, int IMM>
__device__ __forceinline__ void gen4(uint32_t addr, uint32_t x) {
	if (!B_SIDE) STS128(addr, IMM, x & 0x11111111u, x & 0x22222222u, x & 0x44444444u, (x >> 1) & 0x44444444u);
	else STS128(addr, IMM, x & 0x44444444u, x & 0x22222222u, x & 0x11111111u, (x >> 3) & 0x11111111u);
}
/* i8: source word -> two 16-byte chunks */
template <bool B_SIDE, int IMM>
__device__ __forceinline__ void gen8(uint32_t addr_lo, uint32_t addr_hi, uint32_t x) {
	if (!B_SIDE) {
		STS128(addr_lo, IMM, x & 0x01010101u, x & 0x02020202u, x & 0x04040404u, x & 0x08080808u);
		STS128(addr_hi, IMM, x & 0x10101010u, x & 0x20202020u, x & 0x40404040u, x & 0x80808080u);
	} else {
		STS128(addr_lo, IMM, x & 0x80808080u, x & 0x40404040u, x & 0x20202020u, x & 0x10101010u);
		STS128(addr_hi, IMM, x & 0x08080808u, x & 0x04040404u, x & 0x02020202u, x & 0x01010101u);
	}
}
/* one stage of one operand row: FP4 consumes two uint4 (256 samples), I8 one (128 samples) */
template <bool FP4, bool B_SIDE, int IMM>
__device__ __forceinline__ void gen_row(const uint32_t (&off)[8], uint4 v0, uint4 v1) {
	if (FP4) {
		gen4<B_SIDE, IMM>(off[0], v0.x); gen4<B_SIDE, IMM>(off[1], v0.y); gen4<B_SIDE, IMM>(off[2], v0.z); gen4<B_SIDE, IMM>(off[3], v0.w);
		gen4<B_SIDE, IMM>(off[4], v1.x); gen4<B_SIDE, IMM>(off[5], v1.y); gen4<B_SIDE, IMM>(off[6], v1.z); gen4<B_SIDE, IMM>(off[7], v1.w);
	} else {
		gen8<B_SIDE, IMM>(off[0], off[1], v0.x); gen8<B_SIDE, IMM>(off[2], off[3], v0.y);
		gen8<B_SIDE, IMM>(off[4], off[5], v0.z); gen8<B_SIDE, IMM>(off[6], off[7], v0.w);
	}
}
__device__ __forceinline__ uint4 and4(uint4 a, uint4 b) { return make_uint4(a.x & b.x, a.y & b.y, a.z & b.z, a.w & b.w); }

template <bool FP4>
__global__ void __launch_bounds__(UM_THREADS, 1) k_umma_tables(const uint4 *__restrict__ rows_a, const uint4 *__restrict__ rows_b,
		int N, int ngroups, CnbTiles T, long long local_begin, int *__restrict__ counts) {
	constexpr int SRCQ = FP4 ? 2 : 1;       /* quads (uint4) of source bits per row and stage */
	const int NR = 3 * N;                   /* operand rows per quad */
	extern __shared__ uint8_t smem_raw[];
	__shared__ __align__(8) uint64_t bar_full[UM_STAGES], bar_empty[UM_STAGES], bar_done;
	__shared__ uint32_t tmem_base_sh;
	const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const long long tile = T.rank + (local_begin + blockIdx.x) * T.world;
	int ct, c0, cb, ce;
	long long pt;
	tile_of(T.tile_base, T.nct, tile, ct, pt);
	tile_children(T, N, ct, pt, c0, cb, ce);
	const int nchild = ce - cb, ncols = (3 * nchild + 15) & ~15;        /* MMA N of this tile */

	if (warp == UM_GEN_WARPS) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)), "n"(UM_TMEM_COLS));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
	}
	if (tid == 0) {
		for (int s = 0; s < UM_STAGES; s++) {
			mbar_init(&bar_full[s], UM_GEN_WARPS);
			mbar_init(&bar_empty[s], 1);
		}
		mbar_init(&bar_done, 1);
		asm volatile("fence.mbarrier_init.release.cluster;");
	}
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;");
	const uint32_t tmem = tmem_base_sh;
	if (FP4) {
		/* unit scale factors (E8M0 127 = 2^0) in every byte of columns 480..511 of all 128 lanes */
		if (warp < 4) {
			const uint32_t one = 0x7F7F7F7Fu;
#pragma unroll
			for (int h = 0; h < 2; h++)
				asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};"
						::"r"(tmem + ((uint32_t)(warp * 32) << 16) + UM_TMEM_SF + h * 16), "r"(one) : "memory");
			asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
		}
		asm volatile("tcgen05.fence::before_thread_sync;");
		__syncthreads();
		asm volatile("tcgen05.fence::after_thread_sync;");
	}

	if (warp < UM_GEN_WARPS) {
		/* ---- operand generators: thread = one operand row for the whole K loop ---- */
		const bool a_side = warp < 8;
		bool active = true;
		const uint4 *s0, *s1;
		uint32_t dst;
		if (a_side) {
			/* row r = pair*9 + i*3 + j of pair block h; rows 126/127 and pairs past the end repeat valid data whose
			 * accumulators nobody reads */
			const int h = warp >> 2, r = tid & 127;
			int pa, pb;
			pair_of(pt * UM_PAIRS + h * UM_BLK_PAIRS + min(r / 9, UM_BLK_PAIRS - 1), pa, pb);
			if (pb >= N) { pa = 0; pb = 1; }
			s0 = rows_a + (pa * 3 + (r % 9) / 3);
			s1 = rows_a + (pb * 3 + r % 3);
			dst = h * UM_A_BYTES + (r >> 3) * 1024u + (r & 7u) * 128u;
		} else {
			/* row r = child*3 + k of the tile's children [cb, ce); rows past the MMA's columns are not generated */
			const int r = min(tid - 256, UM_N - 1);
			active = tid - 256 < ncols;
			const int c = min(cb + r / 3, ce - 1);
			s0 = s1 = rows_b + (c * 3 + r % 3);
			dst = UM_B_OFF + (r >> 3) * 1024u + (r & 7u) * 128u;
		}
		uint32_t off[8];
		{
			const uint32_t r7 = (dst >> 7) & 7u;
#pragma unroll
			for (int k = 0; k < 8; k++) off[k] = smem + dst + ((k ^ r7) << 4);
		}
		uint4 n0, n1, m0, m1;          /* next stage's source bits (row 0 / row 1 of an A thread) */
		n0 = __ldg(s0);
		n1 = FP4 ? __ldg(s0 + NR) : n0;
		m0 = a_side ? __ldg(s1) : n0;
		m1 = (a_side && FP4) ? __ldg(s1 + NR) : n0;
		if (!active) {
			/* idle child-row threads only keep their warp's arrivals going */
			for (int g = 0; g < ngroups; g++) {
#pragma unroll
				for (int u = 0; u < UM_STAGES; u++) {
					if (g > 0) mbar_wait(&bar_empty[u], (g - 1) & 1);
					asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
					__syncwarp();
					if (lane == 0) mbar_arrive(&bar_full[u]);
				}
			}
		} else
		for (int g = 0; g < ngroups; g++) {
#define UM_STAGE_BODY(u) { \
			uint4 v0 = n0, v1 = n1; \
			if (a_side) { v0 = and4(n0, m0); v1 = and4(n1, m1); } \
			s0 += SRCQ * NR; s1 += SRCQ * NR; \
			n0 = __ldg(s0); \
			if (FP4) n1 = __ldg(s0 + NR); \
			if (a_side) { m0 = __ldg(s1); if (FP4) m1 = __ldg(s1 + NR); } \
			if (g > 0) mbar_wait(&bar_empty[u], (g - 1) & 1); \
			if (a_side) gen_row<FP4, false, (u) * UM_STAGE>(off, v0, v1); \
			else gen_row<FP4, true, (u) * UM_STAGE>(off, v0, v1); \
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); \
			__syncwarp(); \
			if (lane == 0) mbar_arrive(&bar_full[u]); }
			UM_STAGE_BODY(0)
			UM_STAGE_BODY(1)
			UM_STAGE_BODY(2)
#undef UM_STAGE_BODY
		}

		/* ---- epilogue: TMEM -> registers -> global tables.  Row = pair*9 + (i*3+j); column = child*3 + k.
		 * A warp may only touch its own lane quarter (warp & 3); the four warps of a quarter split the two
		 * accumulators in halves of 120 columns (40 children), read in chunks of 24 columns (8 children). */
		mbar_wait(&bar_done, 0);
		asm volatile("tcgen05.fence::after_thread_sync;");
		const int lq = warp & 3, h = warp >> 3, half = (warp >> 2) & 1;
		const int row = lq * 32 + lane, cfg = row % 9;
		const bool row_ok = row < UM_BLK_PAIRS * 9;
		/* accumulator column 3j+k = child cb + j, whose slot inside the tile is (cb - c0) + j */
		int *out = counts + ((size_t)blockIdx.x * UM_SLOTS + (size_t)(h * UM_BLK_PAIRS + row / 9) * UM_CHILD + (cb - c0) + half * 40) * UM_SLOT_INTS + cfg * 4;
#pragma unroll 1
		for (int q = 0; q < 5 && half * 40 + q * 8 < nchild; q++) {
			uint32_t v[24];
			const uint32_t taddr = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(h * UM_N + half * 120 + q * 24);
#define LD8(o, a) asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" \
			: "=r"(v[o]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7]) \
			: "r"(a))
			LD8(0, taddr);
			LD8(8, taddr + 8);
			LD8(16, taddr + 16);
#undef LD8
			asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
			if (row_ok) {
#pragma unroll
				for (int j = 0; j < 8; j++) {
					if (half * 40 + q * 8 + j >= nchild) break;
					int x0, x1, x2;
					if (FP4) {
						x0 = __float2int_rn(__uint_as_float(v[j * 3])); x1 = __float2int_rn(__uint_as_float(v[j * 3 + 1]));
						x2 = __float2int_rn(__uint_as_float(v[j * 3 + 2]));
					} else {
						x0 = (int)(v[j * 3] >> 7); x1 = (int)(v[j * 3 + 1] >> 7); x2 = (int)(v[j * 3 + 2] >> 7);
					}
					*reinterpret_cast<int4 *>(out + (size_t)(q * 8 + j) * UM_SLOT_INTS) = make_int4(x0, x1, x2, 0);
				}
			}
		}
	} else {
		/* ---- MMA issuer: one elected lane, four K = 32-byte instructions per pair block and stage ---- */
		const uint32_t idesc = FP4 ? ((1u << 7) | (1u << 10) | ((uint32_t)(ncols >> 3) << 17) | (1u << 23) | (8u << 24))   /* E2M1 x E2M1, E8M0 scales */
		                           : ((2u << 4) | ((uint32_t)(ncols >> 3) << 17) | (8u << 24));                           /* s32 acc, u8 x u8 */
		for (int g = 0; g < ngroups; g++) {
#pragma unroll
			for (int u = 0; u < UM_STAGES; u++) {
				mbar_wait(&bar_full[u], g & 1);
				asm volatile("tcgen05.fence::after_thread_sync;");
				if (lane == 0) {
					const uint32_t buf = smem + u * UM_STAGE;
					const uint64_t db = umma_desc(buf + UM_B_OFF);
#pragma unroll
					for (int k = 0; k < 4; k++) {                  /* +2: 32 bytes in 16-byte units */
#pragma unroll
						for (int h = 0; h < UM_NA; h++) {
							const uint64_t da = umma_desc(buf + h * UM_A_BYTES);
							if (FP4) umma_mxf4(tmem + h * UM_N, da + 2 * k, db + 2 * k, idesc, (g | u | k) != 0, tmem + UM_TMEM_SF, tmem + UM_TMEM_SF + 16);
							else umma_i8(tmem + h * UM_N, da + 2 * k, db + 2 * k, idesc, (g | u | k) != 0);
						}
					}
					umma_commit(&bar_empty[u]);
				}
				__syncwarp();
			}
		}
		if (lane == 0) umma_commit(&bar_done);
		__syncwarp();
	}
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	if (warp == UM_GEN_WARPS)
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(UM_TMEM_COLS));
}

/* log-lik epilogue over the tables of a tile range: one thread per (tile, pair, child) slot.  Completes the table
 * from the 27 free cells and the pair tables (inclusion-exclusion), then the reference's fp64 sum.  The score goes
 * to the slot's position in the rank-major buffer (slot id = tile*UM_SLOTS + pair*UM_CHILD + child plays the
 * family id). */
__device__ __forceinline__ int um_pair_count(const int *__restrict__ pt, int u, int v, int i, int j) {
	return u < v ? pt[((long long)v * (v - 1) / 2 + u) * 16 + i * 4 + j] : pt[((long long)u * (u - 1) / 2 + v) * 16 + j * 4 + i];
}

__global__ void __launch_bounds__(128) k_umma_epilogue(DevData Dt, long long local_begin, long long local_end,
		const int *__restrict__ counts, const int *__restrict__ pair_tab, ScoreOut O) {
	const CnbTiles &T = Dt.tiles;
	const long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const long long local = local_begin + slot / UM_SLOTS;
	if (local >= local_end) return;
	const long long tile = T.rank + local * T.world;
	const int within = (int)(slot % UM_SLOTS), p = within / UM_CHILD, j = within % UM_CHILD;
	int ct;
	long long pt;
	tile_of(T.tile_base, T.nct, tile, ct, pt);
	int a, b;
	pair_of(pt * UM_PAIRS + p, a, b);
	const int c = ct * T.child + j;
	if (j >= T.child || c >= Dt.N || b >= c) return;      /* not a family: parents must precede the child */
	const int *tab = counts + (size_t)slot * UM_SLOT_INTS;
	int full[64];
#define CELL(i, jj, k) full[((i) * 4 + (jj)) * 4 + (k)]
	for (int i = 0; i < 3; i++)
		for (int jj = 0; jj < 3; jj++)
			for (int k = 0; k < 3; k++) CELL(i, jj, k) = tab[(i * 3 + jj) * 4 + k];
	const int l0 = Dt.order[a], l1 = Dt.order[b], lc = Dt.order[c];
	for (int i = 0; i < 3; i++)
		for (int jj = 0; jj < 3; jj++)
			CELL(i, jj, 3) = um_pair_count(pair_tab, l0, l1, i, jj) - (CELL(i, jj, 0) + CELL(i, jj, 1) + CELL(i, jj, 2));
	for (int i = 0; i < 3; i++)
		for (int k = 0; k < 4; k++)
			CELL(i, 3, k) = um_pair_count(pair_tab, l0, lc, i, k) - (CELL(i, 0, k) + CELL(i, 1, k) + CELL(i, 2, k));
	for (int jj = 0; jj < 4; jj++)
		for (int k = 0; k < 4; k++)
			CELL(3, jj, k) = um_pair_count(pair_tab, l1, lc, jj, k) - (CELL(0, jj, k) + CELL(1, jj, k) + CELL(2, jj, k));
	const int pc0 = Dt.cats[a], pc1 = Dt.cats[b], cc = Dt.cats[c];
	int ncount;
	double ll = dev_loglik(pc0 * pc1, cc, [&](int cfg, int k) {
		int i = cfg / pc1, jj = cfg - i * pc1;
		return CELL(i, jj, k);
	}, &ncount);
#undef CELL
	double score = ll;
	if (Dt.edge) {
		int pars[2] = {a, b};
		score = __dadd_rn(ll, dev_prior(Dt, c, pars, 2));
	}
	O.scores[T.rank * O.seg_len + O.seg_off + local * UM_SLOTS + within] = score;
}

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cnb_set_error("CUDA: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); return CNB_ERR_CUDA; } } while (0)

/* quads (4 sample words) per operand row for W sample words: whole groups of UM_STAGES stages plus one stage of
 * prefetch slack */
int cnbk_umma_row_quads(int W, int fp4, int *ngroups) {
	const int qps = fp4 ? 2 : 1;                          /* quads per stage */
	const int nst = (W + 4 * qps - 1) / (4 * qps), ng = (nst + UM_STAGES - 1) / UM_STAGES;
	if (ngroups) *ngroups = ng;
	return (ng * UM_STAGES + 1) * qps;
}

int cnbk_umma_rows(cudaStream_t st, const uint4 *planes, int N, int W, int fp4, uint4 *rows_a, uint4 *rows_b) {
	const int nq = cnbk_umma_row_quads(W, fp4, nullptr);
	dim3 grid((N + 127) / 128, nq);
	k_umma_rows<<<grid, 128, 0, st>>>(planes, N, W, nq, rows_a, rows_b, fp4);
	CK(cudaGetLastError());
	return CNB_OK;
}

int cnbk_umma_tables(cudaStream_t st, const uint4 *rows_a, const uint4 *rows_b, int N, int W, int fp4,
		const CnbTiles &T, long long tile_begin, long long tile_end, int *counts) {
	if (tile_end <= tile_begin) return CNB_OK;
	int ng;
	cnbk_umma_row_quads(W, fp4, &ng);
	if (fp4) {
		CK(cudaFuncSetAttribute(k_umma_tables<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, UM_SMEM));
		k_umma_tables<true><<<(unsigned)(tile_end - tile_begin), UM_THREADS, UM_SMEM, st>>>(rows_a, rows_b, N, ng, T, tile_begin, counts);
	} else {
		CK(cudaFuncSetAttribute(k_umma_tables<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, UM_SMEM));
		k_umma_tables<false><<<(unsigned)(tile_end - tile_begin), UM_THREADS, UM_SMEM, st>>>(rows_a, rows_b, N, ng, T, tile_begin, counts);
	}
	CK(cudaGetLastError());
	return CNB_OK;
}

long long cnbk_umma_column_macs(int W, int fp4) {
	int ng;
	cnbk_umma_row_quads(W, fp4, &ng);
	return (long long)UM_NA * 128 * ((long long)ng * UM_STAGES * (fp4 ? 256 : 128));
}

int cnbk_umma_epilogue(cudaStream_t st, const DevData &Dt, long long tile_begin, long long tile_end, const int *counts,
		const int *pair_tab, const ScoreOut &O) {
	long long slots = (tile_end - tile_begin) * UM_SLOTS;
	if (slots <= 0) return CNB_OK;
	k_umma_epilogue<<<(unsigned)((slots + 127) / 128), 128, 0, st>>>(Dt, tile_begin, tile_end, counts, pair_tab, O);
	CK(cudaGetLastError());
	return CNB_OK;
}
