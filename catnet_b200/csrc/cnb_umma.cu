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
 * One CTA per SM (all 512 TMEM columns), warp-specialised:
 *   warps 0-7   pair rows: thread = one A row (2 blocks x 128); A goes to TENSOR MEMORY with tcgen05.st (per block a ring
 *               of 3 half-stage slots, UM_ASTAGES)
 *   warps 8-13  column rows: thread = one B row (192); B goes to shared memory, SWIZZLE_128B K-major (3-stage ring,
 *               UM_BSTAGES)
 *   warps 14-15 one elected lane each issues the tcgen05.mma of one pair block: A from TMEM, B from shared memory
 * handed over with mbarriers (generators -> full -> MMA -> tcgen05.commit -> empty).  The tile geometry is a template
 * parameter (UmGeo<GEO>), the pipeline is the same code for all of them:
 *   inclusion-exclusion tiles (complete data): only the free cells of a table are contracted, the rest follows from the
 *               pair tables.  fr = 3 free categories per node (data with four categories): 2 blocks of 14 parent pairs
 *               x 9 rows (2 x 126 of 128 rows, two accumulators) x up to 64 column nodes x 3 columns (192);
 *               fr = 2 (three categories): 2 x 32 pairs x 4 rows, 96 column nodes x 2 columns; fr = 1 (binary data):
 *               2 x 128 pairs x 1 row, 192 column nodes x 1 column -- an MMA costs the same for any N <= 192, so fewer
 *               rows and columns per family mean more families per MMA
 *   full-table tiles (missing values, perturbation masks): 2 blocks of 8 pairs x 16 rows x up to 48 column nodes x 4
 *               columns (192): all 64 cells, the child's keep mask in its rows, a missing value sets no bit
 * The column rows of a stage are generated once and consumed by both pair blocks, and the pair rows never touch shared
 * memory -- the shared memory pipe (generator stores + tensor-core operand reads) was the measured limiter of the
 * all-shared-memory version (profiles/r01_umma_v3_ss_summary.txt).
 *
 * Replaces, for two-parent candidate sets (all of them, or a pooled / fixed-parent subset looked up by slot), the
 * counting loop of CATNET::setNodeSampleProb (/root/reference/src/catnet_class.h:842-857); the log-lik epilogue
 * (problist.h:224-250) runs in k_umma_epilogue over the integer tables, in the reference's operation order.
 */
#include <stdio.h>
#include "cnb_device.cuh"
#include "cnb_kernels.h"
#include "cnb_tiles.h"

/* Inclusion-exclusion (see k_score_planes_ie2): only the 3 x 3 x 3 "free" cells of every table are contracted;
 * cells with a last category follow from the pair tables.  So a pair needs 9 rows and a child 3 columns. */
#define UM_NA 2                           /* pair blocks (accumulators) per tile */
#define UM_BLK_PAIRS 14                   /* 14 pairs x 9 rows = 126 of the 128 rows of one MMA */
#define UM_PAIRS CNB_TILE_PAIRS
#define UM_CHILD CNB_TILE_CHILD
#define UM_SLOTS (UM_PAIRS * UM_CHILD)
#define UM_SLOT_INTS 36                   /* 9 configurations x (3 counts + 1 pad): int4 stores */
#define UM_N ((UM_CHILD * 3 + 15) / 16 * 16)   /* 192 = largest MMA N */
#define UM_STAGE (UM_N * 128)             /* child rows of one stage: 24,576 B */
#define UM_BSTAGES 3                      /* shared-memory ring of the child rows (stages) */
#define UM_ASTAGES 3                      /* tensor-memory ring of the pair rows, per pair block: 3 slots of HALF a stage
                                             (2 k-steps = 16 columns), i.e. 1.5 stages deep */
#define UM_SMEM (UM_BSTAGES * UM_STAGE + 1024)
#define UM_A_WARPS 8
#define UM_B_WARPS 6
#define UM_MMA_WARP (UM_A_WARPS + UM_B_WARPS)   /* first of the UM_NA MMA-issuing warps (one per pair block) */
#define UM_THREADS ((UM_MMA_WARP + UM_NA) * 32)
#define UM_TMEM_COLS 512
#define UM_TMEM_A (UM_NA * UM_N)          /* 384: pair rows, UM_NA blocks x UM_ASTAGES slots x 2 k-steps x 8 columns = 96 columns */
#define UM_TMEM_SF 480                    /* mxf4: unit scale factors in columns 480..511 */
/* Full-table tiles (CnbTiles.full): every cell is contracted -- 16 rows per pair (all (i, j)), 4 columns per column node;
 * a missing value sets no bit in any of its node's rows and a child's rows carry its perturbation keep mask, so the
 * 64 counts of a slot ARE the family's table whatever the data (no inclusion-exclusion).  64 / 27 = 2.4 x the MACs. */
#define UMF_BLK_PAIRS 8                   /* 8 pairs x 16 rows = 128 rows of one MMA */
#define UMF_PAIRS CNB_FULL_PAIRS
#define UMF_CHILD CNB_FULL_CHILD
#define UMF_SLOTS (UMF_PAIRS * UMF_CHILD)
#define UMF_SLOT_INTS 64                  /* 16 configurations x 4 counts: int4 stores */
static_assert(UMF_PAIRS == UM_NA * UMF_BLK_PAIRS && UMF_CHILD * 4 <= UM_N, "full-table tile geometry");
static_assert(UM_PAIRS == UM_NA * UM_BLK_PAIRS, "tile geometry");
static_assert(UM_N <= 256 && UM_TMEM_A + UM_NA * UM_ASTAGES * 16 <= UM_TMEM_SF, "tensor memory budget");
static_assert(UM_STAGE % 1024 == 0, "stages keep the 1024-byte swizzle alignment");
static_assert(UM_B_WARPS * 32 >= UM_N && UM_NA == 2 && UM_BSTAGES == 3 && UM_ASTAGES == 3, "warp roles / ring geometry");
/* half-stage 2u + p (stage u of a group, half p) of a pair block lives in slot (2u + p) % 3 of that block's ring; every
 * slot is used twice per group of 3 stages */
#define UM_ASLOT(u, p) ((2 * (u) + (p)) % 3)
#define UM_AUSE(u, p) ((2 * (u) + (p)) / 3)

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
/* The MMA-side instructions are issued by ONE lane, but from warp-convergent code: the whole warp executes the asm
 * and the instruction itself is predicated on `issue` (non-zero in the elected lane only).  Inside a divergent
 * `if (lane == 0)` the compiler has to move every operand to uniform registers with an ELECT / R2UR.BROADCAST /
 * BRA.U.ANY loop per instruction, which made the issuing warp the bottleneck of the whole kernel. */
__device__ __forceinline__ void umma_commit(uint32_t bar_addr, uint32_t issue) {
	asm volatile("{\n.reg .pred e;\nsetp.ne.b32 e, %1, 0;\n"
			"@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(bar_addr), "r"(issue) : "memory");
}
/* D[tmem] (+)= A[tmem] * B[smem desc]^T, u8 x u8 -> s32, M = 128, K = 32 */
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate,
		uint32_t issue) {
	uint32_t z = 0;
	asm volatile("{\n.reg .pred p, e;\nsetp.ne.b32 p, %4, 0;\nsetp.ne.b32 e, %6, 0;\n"
			"@e tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n}\n"
			::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(z), "r"(issue) : "memory");
}
/* D[tmem] (+)= (A[tmem] * sfa) * (B[smem desc] * sfb)^T, E2M1 x E2M1 -> f32, one E8M0 scale per 32 elements,
 * M = 128, K = 64 */
__device__ __forceinline__ void umma_mxf4(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate,
		uint32_t tmem_sfa, uint32_t tmem_sfb, uint32_t issue) {
	asm volatile("{\n.reg .pred p, e;\nsetp.ne.b32 p, %4, 0;\nsetp.ne.b32 e, %7, 0;\n"
			"@e tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], [%1], %2, %3, [%5], [%6], p;\n}\n"
			::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb), "r"(issue) : "memory");
}
/* K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 in [0,14),
 * LBO = 1 in [16,30), SBO = 1024 B >> 4 in [32,46) (stride between 8-row groups), version 1 in [46,48),
 * layout SWIZZLE_128B = 2 in [61,64) */
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
	return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
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

__global__ void __launch_bounds__(128) k_umma_rows(const uint4 *__restrict__ planes, int N, int W, int q_begin,
		uint4 *__restrict__ rows_a, uint4 *__restrict__ rows_b, int fp4, int stride_a) {
	const int nchunk = (N + blockDim.x - 1) / blockDim.x;   /* one-dimensional grid: no 65,535 limit on the quads */
	const int n = (blockIdx.x % nchunk) * blockDim.x + threadIdx.x, qd = q_begin + blockIdx.x / nchunk;
	if (n >= N) return;
	uint4 v[4];
#pragma unroll
	for (int k = 0; k < 4; k++) {
		const int w = qd * 4 + k;
		v[k] = w < W ? __ldg(planes + (size_t)w * N + n) : make_uint4(0, 0, 0, 0);
	}
	const size_t o = (size_t)qd * (3 * N) + 3 * n, oa = (size_t)qd * stride_a + 3 * n;
	rows_a[oa + 0] = make_uint4(v[0].x, v[1].x, v[2].x, v[3].x);
	rows_a[oa + 1] = make_uint4(v[0].y, v[1].y, v[2].y, v[3].y);
	rows_a[oa + 2] = make_uint4(v[0].z, v[1].z, v[2].z, v[3].z);
	if (!rows_b) return;
	rows_b[o + 0] = make_uint4(rows_b_word(v[0].x, fp4), rows_b_word(v[1].x, fp4), rows_b_word(v[2].x, fp4), rows_b_word(v[3].x, fp4));
	rows_b[o + 1] = make_uint4(rows_b_word(v[0].y, fp4), rows_b_word(v[1].y, fp4), rows_b_word(v[2].y, fp4), rows_b_word(v[3].y, fp4));
	rows_b[o + 2] = make_uint4(rows_b_word(v[0].z, fp4), rows_b_word(v[1].z, fp4), rows_b_word(v[2].z, fp4), rows_b_word(v[3].z, fp4));
}

/* operand rows of the full-table tiles: [quad][node * RS + v * 4 + cat], cat < 4; v = 0 the node's plain one-hot rows
 * (the node as a parent), v = 1 (RS = 8, perturbation masks present) the same rows under the node's own keep mask (the
 * node as the child).  A missing value has no bit in any plane, so it drops out of every product by itself. */
__global__ void __launch_bounds__(128) k_umma_rows_full(const uint4 *__restrict__ planes, const uint32_t *__restrict__ keep, int N, int W,
		uint4 *__restrict__ rows_a, uint4 *__restrict__ rows_b, int RS) {
	const int nchunk = (N + blockDim.x - 1) / blockDim.x;
	const int n = (blockIdx.x % nchunk) * blockDim.x + threadIdx.x, qd = blockIdx.x / nchunk;
	if (n >= N) return;
	uint32_t v[4][4], m[4];                  /* [cat][word of the quad] */
#pragma unroll
	for (int k = 0; k < 4; k++) {
		const int w = qd * 4 + k;
		const uint4 x = w < W ? __ldg(planes + (size_t)w * N + n) : make_uint4(0, 0, 0, 0);
		v[0][k] = x.x; v[1][k] = x.y; v[2][k] = x.z; v[3][k] = x.w;
		m[k] = (keep && w < W) ? __ldg(keep + (size_t)w * N + n) : 0xFFFFFFFFu;
	}
	const size_t o = (size_t)qd * ((size_t)RS * N) + (size_t)RS * n;
#pragma unroll
	for (int c = 0; c < 4; c++) {
		rows_a[o + c] = make_uint4(v[c][0], v[c][1], v[c][2], v[c][3]);
		rows_b[o + c] = make_uint4(rows_b_word(v[c][0], 1), rows_b_word(v[c][1], 1), rows_b_word(v[c][2], 1), rows_b_word(v[c][3], 1));
		if (RS == 8) {
			rows_a[o + 4 + c] = make_uint4(v[c][0] & m[0], v[c][1] & m[1], v[c][2] & m[2], v[c][3] & m[3]);
			rows_b[o + 4 + c] = make_uint4(rows_b_word(v[c][0] & m[0], 1), rows_b_word(v[c][1] & m[1], 1), rows_b_word(v[c][2] & m[2], 1),
					rows_b_word(v[c][3] & m[3], 1));
		}
	}
}

/* ---- operand generation: one 32-sample source word -> output words ------------------------------------------- */
#define STS128(addr, imm, a, b, c, d) asm volatile("st.shared.v4.b32 [%0+%1], {%2, %3, %4, %5};" \
		::"r"(addr), "n"(imm), "r"(a), "r"(b), "r"(c), "r"(d) : "memory")
#define STTM8(addr, a, b, c, d, e, f, g, h) asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" \
		::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e), "r"(f), "r"(g), "r"(h) : "memory")

/* child row (B operand), one stage -> 8 chunks of 16 bytes in shared memory.  mxf4: chunk c = source word c;
 * i8: chunks 2c, 2c+1 = source word c */
template <bool FP4, int IMM>
__device__ __forceinline__ void gen_row_b(const uint32_t (&off)[8], uint4 v0, uint4 v1) {
	if (FP4) {
		const uint32_t x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
		for (int c = 0; c < 8; c++)
			STS128(off[c], IMM, x[c] & 0x44444444u, x[c] & 0x22222222u, x[c] & 0x11111111u, (x[c] >> 3) & 0x11111111u);
	} else {
		const uint32_t x[4] = {v0.x, v0.y, v0.z, v0.w};
#pragma unroll
		for (int c = 0; c < 4; c++) {
			STS128(off[2 * c], IMM, x[c] & 0x80808080u, x[c] & 0x40404040u, x[c] & 0x20202020u, x[c] & 0x10101010u);
			STS128(off[2 * c + 1], IMM, x[c] & 0x08080808u, x[c] & 0x04040404u, x[c] & 0x02020202u, x[c] & 0x01010101u);
		}
	}
}
/* pair row (A operand), one stage -> 4 k-steps of 8 tensor-memory columns (32 bytes of K each) in this thread's lane;
 * column j of a k-step = bytes 4j..4j+3 of the 32-byte K slice, the same order as the shared-memory chunks above */
/* HALF selects k-steps 2*HALF, 2*HALF+1 of the stage (mxf4: the uint4 v holds their 4 source words; i8: v holds the
 * stage's 4 source words, one per k-step) and writes them to the 16 columns at taddr */
template <bool FP4, int HALF>
__device__ __forceinline__ void gen_half_a(uint32_t taddr, uint4 v) {
	if (FP4) {
		const uint32_t x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
		for (int k = 0; k < 2; k++) {
			const uint32_t p = x[2 * k], q = x[2 * k + 1];
			STTM8(taddr + k * 8, p & 0x11111111u, p & 0x22222222u, p & 0x44444444u, (p >> 1) & 0x44444444u,
					q & 0x11111111u, q & 0x22222222u, q & 0x44444444u, (q >> 1) & 0x44444444u);
		}
	} else {
		const uint32_t x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
		for (int k = 0; k < 2; k++) {
			const uint32_t p = x[2 * HALF + k];
			STTM8(taddr + k * 8, p & 0x01010101u, p & 0x02020202u, p & 0x04040404u, p & 0x08080808u,
					p & 0x10101010u, p & 0x20202020u, p & 0x40404040u, p & 0x80808080u);
		}
	}
}
__device__ __forceinline__ uint4 and4(uint4 a, uint4 b) { return make_uint4(a.x & b.x, a.y & b.y, a.z & b.z, a.w & b.w); }
/* the generators load their source bits one stage ahead; this pulls the stage after that into L1, so that the load
 * proper finds it there (ncu: 6.9 % of the kernel's stall samples sat on the first use of the loaded rows) */
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

/* issue loop of pair block H (one thread): everything but the barrier phase and the instruction descriptor is a
 * compile-time constant */
template <bool FP4, int H>
__device__ __forceinline__ void mma_issue_loop(uint32_t smem, uint32_t idesc, int ngroups, uint64_t *bar_full_a,
		uint64_t *bar_empty_a, uint64_t *bar_full_b, uint64_t *bar_empty_b, uint64_t *bar_done) {
	const uint64_t db0 = umma_desc(smem);
	const uint32_t b_ea = smem_u32(&bar_empty_a[H * UM_ASTAGES]), b_eb = smem_u32(&bar_empty_b[0]);
	constexpr uint32_t td = H * UM_N, ta0 = UM_TMEM_A + H * (UM_ASTAGES * 16);
	for (int g = 0; g < ngroups; g++) {
#pragma unroll
		for (int u = 0; u < UM_BSTAGES; u++) {
			mbar_wait(&bar_full_b[u], g & 1);
#pragma unroll
			for (int p = 0; p < 2; p++) {
				mbar_wait(&bar_full_a[H * UM_ASTAGES + UM_ASLOT(u, p)], UM_AUSE(u, p) & 1);
				asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
				for (int kk = 0; kk < 2; kk++) {               /* +2: 32 bytes of K in 16-byte units */
					const int k = 2 * p + kk;
					const uint32_t ta = ta0 + UM_ASLOT(u, p) * 16 + kk * 8;
					const uint64_t db = db0 + (uint64_t)(u * (UM_STAGE >> 4) + 2 * k);
					if (FP4) umma_mxf4(td, ta, db, idesc, (g | u | k) != 0, UM_TMEM_SF, UM_TMEM_SF + 16, 1);
					else umma_i8(td, ta, db, idesc, (g | u | k) != 0, 1);
				}
				umma_commit(b_ea + UM_ASLOT(u, p) * 8, 1);
			}
			umma_commit(b_eb + u * 8, 1);
		}
	}
	umma_commit(smem_u32(bar_done), 1);
}

/* tile geometry as compile-time constants.  GEO 0: inclusion-exclusion tiles, 3 free categories per node; 1: full-table
 * tiles; 2 / 3: inclusion-exclusion tiles with 2 / 1 free categories (data with at most three / two categories per node) */
template <int GEO> struct UmGeo {
	static constexpr bool FULL = GEO == 1;
	static constexpr int FR = GEO == 0 ? 3 : (GEO == 2 ? 2 : (GEO == 3 ? 1 : 4));
	static constexpr int CPC = FR;                    /* columns per column node */
	static constexpr int RPP = FR * FR;               /* rows per pair */
	static constexpr int BLKP = 128 / RPP;            /* pairs per block of 128 rows: 14, 8, 32, 128 */
	static constexpr int TC = FULL ? UMF_CHILD : (GEO == 0 ? UM_CHILD : UM_N / FR);   /* column nodes per tile: 64, 48, 96, 192 */
	static constexpr int CST = CPC >= 3 ? 4 : CPC;    /* ints stored per (configuration, column node) */
	static constexpr int SLOT_INTS = RPP * CST;       /* 36, 64, 8, 1 */
	static constexpr int SLOTS = UM_NA * BLKP * TC;
};
static_assert(UmGeo<0>::BLKP == UM_BLK_PAIRS && UmGeo<0>::SLOT_INTS == UM_SLOT_INTS && UmGeo<0>::SLOTS == UM_SLOTS, "default geometry");
static_assert(UmGeo<1>::BLKP == UMF_BLK_PAIRS && UmGeo<1>::SLOT_INTS == UMF_SLOT_INTS && UmGeo<1>::SLOTS == UMF_SLOTS, "full-table geometry");

template <bool FP4, int GEO>
__global__ void __launch_bounds__(UM_THREADS, 1) k_umma_tables(const uint4 *__restrict__ rows_a, const uint4 *__restrict__ rows_b,
		int N, int g_begin, int g_end, CnbTiles T, long long local_begin, int *__restrict__ counts, int whole, int split, int accumulate) {
	constexpr int SRCQ = FP4 ? 2 : 1;       /* quads (uint4) of source bits per row and stage */
	const int ngroups_all = g_end - g_begin;   /* the sample groups of this launch (all of them, or a slice of the samples) */
	using G = UmGeo<GEO>;
	constexpr bool FULL = G::FULL;
	constexpr int CPC = G::CPC, RPP = G::RPP, BLKP = G::BLKP, CST = G::CST;
	constexpr int SLOTS = G::SLOTS, SLOT_INTS = G::SLOT_INTS, TC = G::TC;
	const int RS = FULL ? T.row_stride : 3; /* operand rows per node */
	const int NRB = RS * N, NRA = RS * (N + T.nvirt);   /* operand rows per quad: column side, pair side (real + virtual nodes) */
	/* CTAs [0, whole) contract all the samples of one tile each; the tiles behind them -- the launch's last, partial
	 * wave -- are cut into `split` sample slices, one CTA each, whose counts are added up in global memory */
	const bool sliced = (int)blockIdx.x >= whole;
	const int tile_local = sliced ? whole + ((int)blockIdx.x - whole) / split : (int)blockIdx.x;
	const int slice = sliced ? ((int)blockIdx.x - whole) % split : 0;
	const int g_first = g_begin + (sliced ? (int)((long long)slice * ngroups_all / split) : 0);
	const int ngroups = g_begin + (sliced ? (int)((long long)(slice + 1) * ngroups_all / split) : ngroups_all) - g_first;
	const size_t src_quads = (size_t)g_first * UM_BSTAGES * SRCQ;   /* quads before this CTA's first stage */
	const int pf = T.pad;                   /* L1 prefetch distance in stages beyond the one being loaded */
	const size_t pf_quads = (size_t)pf * SRCQ;
	extern __shared__ uint8_t smem_raw[];
	__shared__ __align__(8) uint64_t bar_full_b[UM_BSTAGES], bar_empty_b[UM_BSTAGES], bar_full_a[UM_NA * UM_ASTAGES], bar_empty_a[UM_NA * UM_ASTAGES], bar_done;
	__shared__ uint32_t tmem_base_sh;
	const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const long long tile = T.rank + (local_begin + tile_local) * T.world;
	CnbTileInfo ti;                          /* cnb_tiles.h: pair slots (A rows) and column nodes (B rows) of this tile */
	cnb_tile_decode(T, N, tile, ti);
	const int nchild = ti.ze - ti.zb, ncols = (CPC * nchild + 15) & ~15;   /* column nodes and MMA N of this tile */

	if (warp == UM_MMA_WARP) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)), "n"(UM_TMEM_COLS));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
	}
	if (tid == 0) {
		for (int s = 0; s < UM_BSTAGES; s++) {
			mbar_init(&bar_full_b[s], UM_B_WARPS);
			mbar_init(&bar_empty_b[s], UM_NA);           /* one tcgen05.commit per MMA warp */
		}
		for (int s = 0; s < UM_NA * UM_ASTAGES; s++) {
			mbar_init(&bar_full_a[s], UM_A_WARPS / UM_NA);   /* the 4 warps of one pair block */
			mbar_init(&bar_empty_a[s], 1);                   /* the commit of that block's MMA warp */
		}
		mbar_init(&bar_done, UM_NA);
		asm volatile("fence.mbarrier_init.release.cluster;");
	}
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;");
	const uint32_t tmem = tmem_base_sh;
	if (tmem != 0) __trap();                 /* all 512 columns are ours: the MMA issuer uses literal column addresses */
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

	if (warp < UM_A_WARPS) {
		/* ---- pair rows -> tensor memory.  Thread = row r = pair*9 + i*3 + j of pair block h, TMEM lane r; rows
		 * 126/127 and pair slots past the end repeat valid data whose accumulators nobody reads ---- */
		const int h = warp >> 2, r = tid & 127;
		int pa, pb;
		if (!cnb_tile_pair(ti, h * BLKP + min(r / RPP, BLKP - 1), pa, pb)) { pa = 0; pb = 1; }
		/* full tables: in a D tile the pair is (second parent, CHILD): the child's rows are the ones under its keep mask */
		const int vb = (FULL && ti.kind == 1 && RS == 8) ? 4 : 0;
		const uint4 *s0 = rows_a + src_quads * NRA + (pa * RS + (r % RPP) / CPC), *s1 = rows_a + src_quads * NRA + (pb * RS + vb + r % CPC);
		const uint32_t ta = tmem + ((uint32_t)((warp & 3) * 32) << 16) + UM_TMEM_A + h * (UM_ASTAGES * 16);
		uint64_t *const full_a = &bar_full_a[h * UM_ASTAGES], *const empty_a = &bar_empty_a[h * UM_ASTAGES];
		uint4 n0, n1, m0, m1;          /* next stage's source bits of the two nodes */
		n0 = __ldg(s0);
		m0 = __ldg(s1);
		n1 = FP4 ? __ldg(s0 + NRA) : n0;
		m1 = FP4 ? __ldg(s1 + NRA) : m0;
		for (int g = 0; g < ngroups; g++) {
#pragma unroll
			for (int u = 0; u < UM_BSTAGES; u++) {
				const uint4 v0 = and4(n0, m0), v1 = and4(n1, m1);
				s0 += SRCQ * NRA; s1 += SRCQ * NRA;
				n0 = __ldg(s0);
				m0 = __ldg(s1);
				if (FP4) { n1 = __ldg(s0 + NRA); m1 = __ldg(s1 + NRA); }
				if (g * UM_BSTAGES + u + 1 + pf <= ngroups * UM_BSTAGES) {   /* inside the rows (one stage of slack) */
					prefetch_l1(s0 + pf_quads * NRA);
					prefetch_l1(s1 + pf_quads * NRA);
					if (FP4) { prefetch_l1(s0 + pf_quads * NRA + NRA); prefetch_l1(s1 + pf_quads * NRA + NRA); }
				}

				/* each half of the stage goes to its own ring slot (use number 2g + UM_AUSE of that slot) */
#define UM_HALF_BODY(p, v) { \
				if (g > 0 || UM_AUSE(u, p) > 0) mbar_wait(&empty_a[UM_ASLOT(u, p)], (UM_AUSE(u, p) + 1) & 1); \
				asm volatile("tcgen05.fence::after_thread_sync;"); \
				gen_half_a<FP4, p>(ta + UM_ASLOT(u, p) * 16, v); \
				asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); \
				asm volatile("tcgen05.fence::before_thread_sync;"); \
				__syncwarp(); \
				if (lane == 0) mbar_arrive(&full_a[UM_ASLOT(u, p)]); }
				UM_HALF_BODY(0, v0)
				UM_HALF_BODY(1, (FP4 ? v1 : v0))
#undef UM_HALF_BODY
			}
		}

		/* ---- epilogue: TMEM -> registers -> global tables.  Row = pair slot*RPP + configuration; accumulator column
		 * CPC*j+k = column node zb + j (column slot j of the tile).  A warp may only touch its own lane quarter
		 * (warp & 3); each of the 8 warps reads its rows of one accumulator in chunks of 8 nodes. */
		mbar_wait(&bar_done, 0);
		asm volatile("tcgen05.fence::after_thread_sync;");
		const int lq = warp & 3;
		const int row = lq * 32 + lane, cfg = row % RPP;
		const bool row_ok = row < BLKP * RPP;
		int *out = counts + ((size_t)tile_local * SLOTS + (size_t)(h * BLKP + row / RPP) * TC) * SLOT_INTS + cfg * CST;
#pragma unroll 1
		for (int q = 0; q * 8 < nchild; q++) {
			uint32_t v[8 * CPC];
			const uint32_t taddr = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(h * UM_N + q * 8 * CPC);
#define LD8(o, a) asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" \
			: "=r"(v[o]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7]) \
			: "r"(a))
			LD8(0, taddr);
			if constexpr (CPC >= 2) LD8(8, taddr + 8);
			if constexpr (CPC >= 3) LD8(16, taddr + 16);
			if constexpr (CPC >= 4) LD8(24, taddr + 24);
#undef LD8
			asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
			if (row_ok) {
#pragma unroll
				for (int j = 0; j < 8; j++) {
					if (q * 8 + j >= nchild) break;
					int x[4] = {0, 0, 0, 0};
#pragma unroll
					for (int k = 0; k < CPC; k++)
						x[k] = FP4 ? __float2int_rn(__uint_as_float(v[j * CPC + k])) : (int)(v[j * CPC + k] >> 7);
					int *dst = out + (size_t)(q * 8 + j) * SLOT_INTS;
					if (sliced || accumulate == 2 || (accumulate && CST < 4)) {
						/* accumulate == 2: a later slice of the samples added with fire-and-forget reductions (RED): the 16-byte
						 * read-modify-write below leaves the SM -- tensor memory held, no other CTA can start -- waiting for
						 * the reads */
#pragma unroll
						for (int k = 0; k < CPC; k++) atomicAdd(dst + k, x[k]);
					} else if (accumulate) {
						/* a later slice of the samples: this CTA alone owns the tile in this launch */
						const int4 o = *reinterpret_cast<const int4 *>(dst);
						*reinterpret_cast<int4 *>(dst) = make_int4(o.x + x[0], o.y + x[1], o.z + x[2], o.w + x[3]);
					} else if constexpr (CST == 4) *reinterpret_cast<int4 *>(dst) = make_int4(x[0], x[1], x[2], x[3]);
					else if constexpr (CST == 2) *reinterpret_cast<int2 *>(dst) = make_int2(x[0], x[1]);
					else *dst = x[0];
				}
			}
		}
	} else if (warp < UM_MMA_WARP) {
		/* ---- column rows -> shared memory.  Thread = row r = node*3 + k of the tile's column nodes [zb, ze); rows
		 * past the MMA's columns are not generated, their threads only keep the warp's arrivals going ---- */
		const int rr = tid - UM_A_WARPS * 32, r = min(rr, UM_N - 1);
		const bool active = rr < ncols;
		/* full tables: in an F tile the columns are the CHILDREN: their rows under the keep mask */
		const int vc = (FULL && (ti.kind == 0 || ti.kind == 2) && RS == 8) ? 4 : 0;   /* P tiles: the columns are the children too */
		const uint4 *s0 = rows_b + src_quads * NRB + (min(ti.zb + r / CPC, ti.ze - 1) * RS + vc + r % CPC);
		uint32_t off[8];
#pragma unroll
		for (int k = 0; k < 8; k++) off[k] = smem + (r >> 3) * 1024u + (r & 7u) * 128u + ((k ^ (r & 7u)) << 4);
		uint4 n0, n1;
		n0 = __ldg(s0);
		n1 = FP4 ? __ldg(s0 + NRB) : n0;
		for (int g = 0; g < ngroups; g++) {
#define UM_STAGE_BODY(u) { \
			const uint4 v0 = n0, v1 = n1; \
			s0 += SRCQ * NRB; \
			if (active) { n0 = __ldg(s0); if (FP4) n1 = __ldg(s0 + NRB); } \
			if (active && g * UM_BSTAGES + (u) + 1 + pf <= ngroups * UM_BSTAGES) { prefetch_l1(s0 + pf_quads * NRB); if (FP4) prefetch_l1(s0 + pf_quads * NRB + NRB); } \
			if (g > 0) mbar_wait(&bar_empty_b[u], (g - 1) & 1); \
			if (active) gen_row_b<FP4, (u) * UM_STAGE>(off, v0, v1); \
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); \
			__syncwarp(); \
			if (lane == 0) mbar_arrive(&bar_full_b[u]); }
			UM_STAGE_BODY(0)
			UM_STAGE_BODY(1)
			UM_STAGE_BODY(2)
#undef UM_STAGE_BODY
		}
	} else {
		/* ---- MMA issuers: warp UM_MMA_WARP + h owns pair block h (its own accumulator, so the two instruction
		 * streams never touch the same tensor-memory columns); ONE lane per warp runs the issue loop, four
		 * K = 32-byte MMAs per stage.  A single issuing warp was the bottleneck (profiles/r01_umma_v4_ts_summary.txt:
		 * 92 % of its samples inside the issue sequence, 7 cycles per instruction).  Operands are compile-time
		 * constants or warp-uniform values: the CTA owns all 512 tensor-memory columns, so its allocation starts at
		 * column 0 (checked above). ---- */
		const uint32_t idesc = FP4 ? ((1u << 7) | (1u << 10) | ((uint32_t)(ncols >> 3) << 17) | (1u << 23) | (8u << 24))   /* E2M1 x E2M1, E8M0 scales */
		                           : ((2u << 4) | ((uint32_t)(ncols >> 3) << 17) | (8u << 24));                           /* s32 acc, u8 x u8 */
		if (lane == 0) {
			if (warp == UM_MMA_WARP) mma_issue_loop<FP4, 0>(smem, idesc, ngroups, bar_full_a, bar_empty_a, bar_full_b, bar_empty_b, &bar_done);
			else mma_issue_loop<FP4, 1>(smem, idesc, ngroups, bar_full_a, bar_empty_a, bar_full_b, bar_empty_b, &bar_done);
		}
	}
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	if (warp == UM_MMA_WARP)
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(UM_TMEM_COLS));
}

/* log-lik epilogue over the tables of a tile range: one thread per (tile, pair, child) slot.  Completes the table
 * from the 27 free cells and the pair tables (inclusion-exclusion), then the reference's fp64 sum.  The score goes
 * to the slot's position in the rank-major buffer (slot id = tile*UM_SLOTS + pair*UM_CHILD + child plays the
 * family id). */
__device__ __forceinline__ int um_pair_count(const int *__restrict__ pt, int u, int v, int i, int j) {
	return u < v ? pt[((long long)v * (v - 1) / 2 + u) * 16 + i * 4 + j] : pt[((long long)u * (u - 1) / 2 + v) * 16 + j * 4 + i];
}

/* the three 4 x 4 tables of the family's node pairs, oriented [first][second], by 16-byte loads (UNI4 path) */
__device__ __forceinline__ void um_pair_table(const int *__restrict__ pt, int u, int v, int (&t)[16]) {
	const bool sw = u > v;
	const int lo = sw ? v : u, hi = sw ? u : v;
	const int4 *p = reinterpret_cast<const int4 *>(pt + ((long long)hi * (hi - 1) / 2 + lo) * 16);
	int r[16];
#pragma unroll
	for (int q = 0; q < 4; q++) {
		const int4 x = __ldg(p + q);
		r[4 * q] = x.x; r[4 * q + 1] = x.y; r[4 * q + 2] = x.z; r[4 * q + 3] = x.w;
	}
#pragma unroll
	for (int i = 0; i < 4; i++)
#pragma unroll
		for (int j = 0; j < 4; j++) t[i * 4 + j] = sw ? r[j * 4 + i] : r[i * 4 + j];
}

/* the terms n * log(n * pp), pp = 1 / N_k, of one parent configuration added to ll in the reference's order (child
 * categories ascending, empty cells skipped; problist.h:236-247).  A call, not 16 inlined copies (150 KB of code with
 * the logs).  The four logs are independent -- only the additions are ordered -- so when all of them take the
 * branch-free table path of glibc's algorithm (cnb_log_far: every ratio outside [0.9375, 1.065)) they are evaluated
 * side by side, four dependency chains deep instead of one. */
__device__ __noinline__ double um_config_terms(double ll, int n0, int n1, int n2, int n3, long long nk) {
	const double pp = __ddiv_rn(1.0, (double)nk);
	const int nn[4] = {n0, n1, n2, n3};
	double dn[4], x[4];
	bool far = true;
#pragma unroll
	for (int k = 0; k < 4; k++) {
		dn[k] = (double)nn[k];
		x[k] = __dmul_rn(dn[k], pp);
		far = far && (nn[k] <= 0 || cnb_log_is_far((uint64_t)__double_as_longlong(x[k])));
	}
	if (far) {
		double t[4];
#pragma unroll
		for (int k = 0; k < 4; k++) t[k] = __dmul_rn(dn[k], cnb_log_far((uint64_t)__double_as_longlong(x[k])));   /* empty cell: unused */
#pragma unroll
		for (int k = 0; k < 4; k++)
			if (nn[k] > 0) ll = __dadd_rn(ll, t[k]);
	} else {
#pragma unroll
		for (int k = 0; k < 4; k++)
			if (nn[k] > 0) ll = __dadd_rn(ll, __dmul_rn(dn[k], cnb_log(x[k])));
	}
	return ll;
}

/* UNI4: every node has exactly 4 categories (C5): all loop bounds and table indices are compile-time constants, the
 * 64-cell table stays in registers (the generic path indexes it dynamically through local memory: 7,185 instructions
 * per family, 17 % of them fp64 arithmetic, profiles/r01_umma_epilogue_summary.txt) */
template <bool UNI4, bool FULL>
__global__ void __launch_bounds__(128) k_umma_epilogue(DevData Dt, long long local_begin, long long local_end,
		const int *__restrict__ counts, const int *__restrict__ pair_tab, ScoreOut O) {
	const CnbTiles &T = Dt.tiles;
	/* geometry: constants for full-table tiles and for the 4-category specialisation (3 free categories), else the plan's */
	const int fr = FULL ? 4 : (UNI4 ? 3 : T.fr);
	const int SLOTS = FULL ? UMF_SLOTS : (UNI4 ? UM_SLOTS : T.tp * T.tc), TC = FULL ? UMF_CHILD : (UNI4 ? UM_CHILD : T.tc);
	const int SLOT_INTS = FULL ? UMF_SLOT_INTS : (UNI4 ? UM_SLOT_INTS : cnb_geo_slot_ints(T.fr));
	const long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const long long local = local_begin + slot / SLOTS;
	if (local >= local_end) return;
	const long long tile = T.rank + local * T.world;
	const int within = (int)(slot % SLOTS), p = within / TC, j = within % TC;
	CnbTileInfo ti;
	cnb_tile_decode(T, Dt.N, tile, ti);
	int x, y;
	if (j >= ti.ze - ti.zb || !cnb_tile_pair(ti, p, x, y)) return;
	/* F tile: pair (a, b), column = child c.  D tile: pair (b, c), column = first parent a, which must precede b. */
	const bool dk = ti.kind != 0;
	const int a = dk ? ti.zb + j : x, b = dk ? x : y, c = dk ? y : ti.zb + j;
	if (a >= b) return;
	const int *tab = counts + (size_t)slot * SLOT_INTS;   /* the tables of a launch start at its first tile */
	int full[64];
#define CELL(i, jj, k) full[((i) * 4 + (jj)) * 4 + (k)]
	if (FULL) {
		/* slot layout: 16 pair configurations x 4 column categories: the table itself, sixteen 16-byte loads */
#pragma unroll
		for (int r = 0; r < 16; r++) {
			const int4 v = __ldg(reinterpret_cast<const int4 *>(tab) + r);
			const int x4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
			for (int q = 0; q < 4; q++) {
				if (dk) CELL(q, r / 4, r % 4) = x4[q];
				else CELL(r / 4, r % 4, q) = x4[q];
			}
		}
	} else if (UNI4 || fr == 3) {
	/* slot layout: 9 pair configurations x (3 column categories + pad): nine 16-byte loads */
#pragma unroll
	for (int r = 0; r < 9; r++) {
		const int4 v = __ldg(reinterpret_cast<const int4 *>(tab) + r);
		const int x3[3] = {v.x, v.y, v.z};
#pragma unroll
		for (int q = 0; q < 3; q++) {
			/* F tile: row r = (i, jj), column category q = k.  D tile: row r = (jj, k) of the pair (b, c), column q = i */
			if (dk) CELL(q, r / 3, r % 3) = x3[q];
			else CELL(r / 3, r % 3, q) = x3[q];
		}
	}
	} else {
	/* fewer free categories: fr^2 configurations x fr counts, unpadded */
	for (int r = 0; r < fr * fr; r++)
		for (int q = 0; q < fr; q++) {
			const int v = __ldg(tab + r * fr + q);
			if (dk) CELL(q, r / fr, r % fr) = v;
			else CELL(r / fr, r % fr, q) = v;
		}
	}
	const int l0 = Dt.order[a], l1 = Dt.order[b], lc = Dt.order[c];
	/* the reference sums a table in the order of the family's LISTED parents (fixed parents first, then the free ones
	 * ascending; catnet_search2.h:532-553): the only two-parent listing that is not ascending is (fixed b, free a < b) */
	const bool listed_ba = Dt.fix1 && Dt.fix1[c] == b;
	double ll;
	if (UNI4) {
		if (!FULL) {
		int p01[16], p0c[16], p1c[16];
		um_pair_table(pair_tab, l0, l1, p01);
		um_pair_table(pair_tab, l0, lc, p0c);
		um_pair_table(pair_tab, l1, lc, p1c);
#pragma unroll
		for (int i = 0; i < 3; i++)
#pragma unroll
			for (int jj = 0; jj < 3; jj++)
				CELL(i, jj, 3) = p01[i * 4 + jj] - (CELL(i, jj, 0) + CELL(i, jj, 1) + CELL(i, jj, 2));
#pragma unroll
		for (int i = 0; i < 3; i++)
#pragma unroll
			for (int k = 0; k < 4; k++)
				CELL(i, 3, k) = p0c[i * 4 + k] - (CELL(i, 0, k) + CELL(i, 1, k) + CELL(i, 2, k));
#pragma unroll
		for (int jj = 0; jj < 4; jj++)
#pragma unroll
			for (int k = 0; k < 4; k++)
				CELL(3, jj, k) = p1c[jj * 4 + k] - (CELL(0, jj, k) + CELL(1, jj, k) + CELL(2, jj, k));
		}
		/* dev_loglik with constant bounds, same operation order: configurations ascending, pp = 1 / N_k,
		 * ll += n * log(n * pp) for n > 0, division by the sample count last */
		ll = 0.0;
		long long ncount = 0;
#pragma unroll
		for (int cfg0 = 0; cfg0 < 16; cfg0++) {
			/* a candidate (fixed parent b, free parent a < b) is listed [b, a]: its configurations run b-major */
			const int cfg = listed_ba ? (cfg0 & 3) * 4 + (cfg0 >> 2) : cfg0;
			const int n0 = full[cfg * 4], n1 = full[cfg * 4 + 1], n2 = full[cfg * 4 + 2], n3 = full[cfg * 4 + 3];
			const long long nk = (long long)n0 + n1 + n2 + n3;
			if (nk > 0) {
				ncount += nk;
				ll = um_config_terms(ll, n0, n1, n2, n3, nk);
			}
		}
		if (ncount > 1) ll = __ddiv_rn(ll, (double)ncount);
	} else {
	if (!FULL) {
	/* the cells with a last category (index fr), axis after axis, from the pair tables */
	for (int i = 0; i < fr; i++)
		for (int jj = 0; jj < fr; jj++) {
			int sum = 0;
			for (int k = 0; k < fr; k++) sum += CELL(i, jj, k);
			CELL(i, jj, fr) = um_pair_count(pair_tab, l0, l1, i, jj) - sum;
		}
	for (int i = 0; i < fr; i++)
		for (int k = 0; k <= fr; k++) {
			int sum = 0;
			for (int jj = 0; jj < fr; jj++) sum += CELL(i, jj, k);
			CELL(i, fr, k) = um_pair_count(pair_tab, l0, lc, i, k) - sum;
		}
	for (int jj = 0; jj <= fr; jj++)
		for (int k = 0; k <= fr; k++) {
			int sum = 0;
			for (int i = 0; i < fr; i++) sum += CELL(i, jj, k);
			CELL(fr, jj, k) = um_pair_count(pair_tab, l1, lc, jj, k) - sum;
		}
	}
	const int pc0 = Dt.cats[a], pc1 = Dt.cats[b], cc = Dt.cats[c];
	int ncount;
	if (listed_ba)
		ll = dev_loglik(pc0 * pc1, cc, [&](int cfg, int k) {
			int jj = cfg / pc0, i = cfg - jj * pc0;
			return CELL(i, jj, k);
		}, &ncount);
	else
		ll = dev_loglik(pc0 * pc1, cc, [&](int cfg, int k) {
			int i = cfg / pc1, jj = cfg - i * pc1;
			return CELL(i, jj, k);
		}, &ncount);
	}
#undef CELL
	double score = ll;
	if (Dt.edge) {
		int pars[2] = {a, b};
		score = __dadd_rn(ll, dev_prior(Dt, c, pars, 2));
	}
	O.scores[T.rank * O.seg_len + O.seg_off + local * SLOTS + within] = score;
}

/* The tables of explicit two-parent families (the networks a search returns) read back from the tile tables that are still
 * resident after the search, instead of counting them again over the samples: slot -> 64 cells (completed from the pair
 * tables for inclusion-exclusion tiles) -> the reference's table layout (first listed parent most significant, child
 * fastest), log-likelihood, sample size. */
template <bool FULL>
__global__ void __launch_bounds__(128) k_umma_export(DevData Dt, long long f_begin, long long f_end, const int *__restrict__ ex_child,
		const int *__restrict__ ex_pars, const int *__restrict__ counts, const int *__restrict__ pair_tab, ScoreOut O) {
	const long long f = f_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= f_end) return;
	const CnbTiles &T = Dt.tiles;
	const int fr = FULL ? 4 : T.fr, cst = cnb_geo_cst(fr);
	const int SLOTS = FULL ? UMF_SLOTS : T.tp * T.tc, SLOT_INTS = FULL ? UMF_SLOT_INTS : cnb_geo_slot_ints(T.fr);
	const int c = ex_child[f], p0 = ex_pars[f * CNB_MAXP], p1 = ex_pars[f * CNB_MAXP + 1];
	const int a = min(p0, p1), b = max(p0, p1);
	long long tile;
	int within;
	cnb_family_slot(T, Dt.N, a, b, c, tile, within);
	const bool dk = b >= (c / T.child) * T.child;           /* D tile: pair (b, c), column a */
	const int *tab = counts + ((size_t)tile * SLOTS + within) * SLOT_INTS;
	int full[64];
#define CELL(i, jj, k) full[((i) * 4 + (jj)) * 4 + (k)]
	if (FULL) {
		for (int r = 0; r < 16; r++)
			for (int q = 0; q < 4; q++) {
				const int v = tab[r * 4 + q];
				if (dk) CELL(q, r / 4, r % 4) = v;
				else CELL(r / 4, r % 4, q) = v;
			}
	} else {
		for (int r = 0; r < fr * fr; r++)
			for (int q = 0; q < fr; q++) {
				const int v = tab[r * cst + q];
				if (dk) CELL(q, r / fr, r % fr) = v;
				else CELL(r / fr, r % fr, q) = v;
			}
		const int l0 = Dt.order[a], l1 = Dt.order[b], lc = Dt.order[c];
		for (int i = 0; i < fr; i++)
			for (int jj = 0; jj < fr; jj++) {
				int sum = 0;
				for (int k = 0; k < fr; k++) sum += CELL(i, jj, k);
				CELL(i, jj, fr) = um_pair_count(pair_tab, l0, l1, i, jj) - sum;
			}
		for (int i = 0; i < fr; i++)
			for (int k = 0; k <= fr; k++) {
				int sum = 0;
				for (int jj = 0; jj < fr; jj++) sum += CELL(i, jj, k);
				CELL(i, fr, k) = um_pair_count(pair_tab, l0, lc, i, k) - sum;
			}
		for (int jj = 0; jj <= fr; jj++)
			for (int k = 0; k <= fr; k++) {
				int sum = 0;
				for (int i = 0; i < fr; i++) sum += CELL(i, jj, k);
				CELL(fr, jj, k) = um_pair_count(pair_tab, l1, lc, jj, k) - sum;
			}
	}
	/* table order = the family's LISTED parents (p0 most significant) */
	const bool sw = p0 > p1;
	const int pc0 = Dt.cats[p0], pc1 = Dt.cats[p1], cc = Dt.cats[c];
	auto cell = [&](int i0, int i1, int k) { return sw ? CELL(i1, i0, k) : CELL(i0, i1, k); };
	int ncount;
	const double ll = dev_loglik(pc0 * pc1, cc, [&](int cfg, int k) { return cell(cfg / pc1, cfg % pc1, k); }, &ncount);
	if (O.counts) {
		int *dst = O.counts + (size_t)f * O.counts_stride;
		for (int i0 = 0; i0 < pc0; i0++)
			for (int i1 = 0; i1 < pc1; i1++)
				for (int k = 0; k < cc; k++) dst[(i0 * pc1 + i1) * cc + k] = cell(i0, i1, k);
	}
#undef CELL
	int pars[CNB_MAXP];
	pars[0] = p0;
	pars[1] = p1;
	dev_store_outputs(Dt, O, f, c, pars, 2, ll, ncount);
}

/* pair tables from the P tiles: one thread per unordered node pair (u < v).  The 3 x 3 free cells come from the tile
 * slot (pair slot u, column v), rows (i, i); the margins of u and v from the diagonal slots (u, u) and (v, v); cells
 * with a last category by inclusion-exclusion (no value is missing, so every margin sums to M). */
__global__ void __launch_bounds__(128) k_pairs_from_tiles(const int *__restrict__ counts, int N, int M, int *__restrict__ pair_tab) {
	const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= (long long)N * (N - 1) / 2) return;
	int u, v;
	cnb_pair_of(q, u, v);
	const int nch = (N + UM_CHILD - 1) / UM_CHILD;
	auto slot = [&](int x, int z) {
		const long long tile = (long long)(x / UM_PAIRS) * nch + z / UM_CHILD;
		return counts + ((size_t)tile * UM_SLOTS + (size_t)(x % UM_PAIRS) * UM_CHILD + z % UM_CHILD) * UM_SLOT_INTS;
	};
	const int *uv = slot(u, v), *uu = slot(u, u), *vv = slot(v, v);
	int n[4][4], mu[4], mv[4], su = 0, sv = 0;
	for (int i = 0; i < 3; i++) {
		mu[i] = uu[(i * 3 + i) * 4 + i];
		mv[i] = vv[(i * 3 + i) * 4 + i];
		su += mu[i];
		sv += mv[i];
	}
	mu[3] = M - su;
	mv[3] = M - sv;
	for (int i = 0; i < 3; i++) {
		int s = 0;
		for (int k = 0; k < 3; k++) { n[i][k] = uv[(i * 3 + i) * 4 + k]; s += n[i][k]; }
		n[i][3] = mu[i] - s;
	}
	for (int k = 0; k < 4; k++) n[3][k] = mv[k] - (n[0][k] + n[1][k] + n[2][k]);
	int *out = pair_tab + q * 16;
	for (int i = 0; i < 4; i++)
		for (int k = 0; k < 4; k++) out[i * 4 + k] = n[i][k];
}

/* the same from P tiles that ran on ORDER-space operand rows (the streamed one-shot search builds only those): pair slot /
 * column = positions, the table goes to its label pair, oriented [lower label][higher label] */
__global__ void __launch_bounds__(128) k_pairs_from_tiles_order(const int *__restrict__ counts, const int *__restrict__ order, int N, int M,
		int *__restrict__ pair_tab) {
	const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= (long long)N * (N - 1) / 2) return;
	int u, v;
	cnb_pair_of(q, u, v);                                   /* positions u < v */
	const int nch = (N + UM_CHILD - 1) / UM_CHILD;
	auto slot = [&](int x, int z) {
		const long long tile = (long long)(x / UM_PAIRS) * nch + z / UM_CHILD;
		return counts + ((size_t)tile * UM_SLOTS + (size_t)(x % UM_PAIRS) * UM_CHILD + z % UM_CHILD) * UM_SLOT_INTS;
	};
	const int *uv = slot(u, v), *uu = slot(u, u), *vv = slot(v, v);
	int n[4][4], mu[4], mv[4], su = 0, sv = 0;
	for (int i = 0; i < 3; i++) {
		mu[i] = uu[(i * 3 + i) * 4 + i];
		mv[i] = vv[(i * 3 + i) * 4 + i];
		su += mu[i];
		sv += mv[i];
	}
	mu[3] = M - su;
	mv[3] = M - sv;
	for (int i = 0; i < 3; i++) {
		int s = 0;
		for (int k = 0; k < 3; k++) { n[i][k] = uv[(i * 3 + i) * 4 + k]; s += n[i][k]; }
		n[i][3] = mu[i] - s;
	}
	for (int k = 0; k < 4; k++) n[3][k] = mv[k] - (n[0][k] + n[1][k] + n[2][k]);
	const int lu = order[u], lv = order[v];
	const bool sw = lu > lv;
	const int lo = sw ? lv : lu, hi = sw ? lu : lv;
	int *out = pair_tab + ((long long)hi * (hi - 1) / 2 + lo) * 16;
	for (int i = 0; i < 4; i++)
		for (int k = 0; k < 4; k++) out[sw ? k * 4 + i : i * 4 + k] = n[i][k];
}

/* three-parent families of one column tile from its T tiles: one thread per family (a < b < c | x).  The f1 slots
 * (triple, i) hold the cells n(a = i, b = j, c = k, x = l) at (j * 3 + k) * 4 + l; the free ones (all indices < CM - 1) go
 * to shared memory, dev_ie3_finish completes the table from the three-way tables and sums the log-likelihood. */
template <int CM, int BS>
__global__ void __launch_bounds__(BS) k_umma_epilogue3(DevData Dt, DevFamilies F, long long fid_begin, long long fid_end,
		int c0, const int *__restrict__ counts, const int *__restrict__ tt, ScoreOut O) {
	constexpr int F1 = CM - 1;
	extern __shared__ __align__(16) int s_cnt[];          /* [CM^4][BS] */
	const long long fid = fid_begin + (long long)blockIdx.x * BS + threadIdx.x;
	if (fid >= fid_end) return;
	int child, pars[CNB_MAXP];
	dev_family(Dt, F, fid, child, pars);
	const int a = pars[0], b = pars[1], c = pars[2];
	const long long t = cnb_binom3(c) + (long long)b * (b - 1) / 2 + a;
	int *mine = s_cnt + threadIdx.x;
	/* the T tiles of CM categories have the geometry of F1 = CM - 1 free categories (cnb_tiles.h: cnb_geo_*) */
	constexpr int TP = F1 >= 3 ? CNB_TILE_PAIRS : (F1 == 2 ? 64 : 256), TCH = F1 >= 3 ? CNB_TILE_CHILD : (F1 == 2 ? 96 : 192);
	constexpr int CST = F1 >= 3 ? 4 : F1, SINTS = F1 * F1 * CST;
#pragma unroll
	for (int i = 0; i < F1; i++) {
		const long long q = t * F1 + i;
		const int *tab = counts + ((size_t)(q / TP) * (TP * TCH) + (size_t)(q % TP) * TCH + (child - c0)) * SINTS;
#pragma unroll
		for (int j = 0; j < F1; j++)
#pragma unroll
			for (int k = 0; k < F1; k++) {
				int x[3] = {0, 0, 0};
				if constexpr (CST == 4) {
					const int4 v = *reinterpret_cast<const int4 *>(tab + (j * 3 + k) * 4);
					x[0] = v.x; x[1] = v.y; x[2] = v.z;
				} else if constexpr (CST == 2) {
					const int2 v = *reinterpret_cast<const int2 *>(tab + (j * 2 + k) * 2);
					x[0] = v.x; x[1] = v.y;
				} else x[0] = tab[0];
#pragma unroll
				for (int l = 0; l < F1; l++) mine[((((i * CM + j) * CM + k) * CM + l)) * BS] = x[l];
			}
	}
	int ncount;
	const double ll = dev_ie3_finish<CM, BS>(mine, Dt, tt, pars, child, &ncount);
	dev_store_outputs(Dt, O, fid, child, pars, 3, ll, ncount);
}

/* The same for a RANGE of triples of one column tile: one thread per (triple t in [t_begin, t_end), child x of the column
 * tile [c0, ce)) with t < C(x, 3).  counts holds the tiles that contain the slots q = t * F1 + i from q_base on (q_base is a
 * multiple of the tile's pair slots), so a column tile can be contracted in batches that fit the table buffer, and its
 * tiles can be dealt to several devices.  The family id is the child's block start plus the LEXICOGRAPHIC rank of (a, b, c)
 * among the 3-subsets of the child's predecessors (the candidate order of combinationSets, catnet_search2.h:102-147);
 * the group is the unrestricted one, block k = child 3 + k. */
template <int CM, int BS>
__global__ void __launch_bounds__(BS) k_umma_epilogue3_range(DevData Dt, DevFamilies F, long long t_begin, long long t_end, long long q_base,
		int c0, int ce, const int *__restrict__ counts, const int *__restrict__ tt, ScoreOut O) {
	constexpr int F1 = CM - 1;
	extern __shared__ __align__(16) int s_cnt[];          /* [CM^4][BS] */
	const long long t = t_begin + (long long)blockIdx.x * BS + threadIdx.x;
	const int child = (c0 > 3 ? c0 : 3) + (int)blockIdx.y;
	if (t >= t_end || child >= ce || t >= cnb_binom3(child)) return;
	int a, b, c;
	cnb_triple_of(t, a, b, c);
	int pars[CNB_MAXP];
	pars[0] = a; pars[1] = b; pars[2] = c;
	const long long x = child;
	const long long lex = cnb_binom3(x) - cnb_binom3(x - a) + (x - a - 1) * (x - a - 2) / 2 - (x - b) * (x - b - 1) / 2 + (c - b - 1);
	const long long fid = Dt.blocks[Dt.group_blocks[F.blk_off + (child - 3)]].start + lex;
	int *mine = s_cnt + threadIdx.x;
	constexpr int TP = F1 >= 3 ? CNB_TILE_PAIRS : (F1 == 2 ? 64 : 256), TCH = F1 >= 3 ? CNB_TILE_CHILD : (F1 == 2 ? 96 : 192);
	constexpr int CST = F1 >= 3 ? 4 : F1, SINTS = F1 * F1 * CST;
#pragma unroll
	for (int i = 0; i < F1; i++) {
		const long long q = t * F1 + i - q_base;
		const int *tab = counts + ((size_t)(q / TP) * (TP * TCH) + (size_t)(q % TP) * TCH + (child - c0)) * SINTS;
#pragma unroll
		for (int j = 0; j < F1; j++)
#pragma unroll
			for (int k = 0; k < F1; k++) {
				int v[3] = {0, 0, 0};
				if constexpr (CST == 4) {
					const int4 w = *reinterpret_cast<const int4 *>(tab + (j * 3 + k) * 4);
					v[0] = w.x; v[1] = w.y; v[2] = w.z;
				} else if constexpr (CST == 2) {
					const int2 w = *reinterpret_cast<const int2 *>(tab + (j * 2 + k) * 2);
					v[0] = w.x; v[1] = w.y;
				} else v[0] = tab[0];
#pragma unroll
				for (int l = 0; l < F1; l++) mine[((((i * CM + j) * CM + k) * CM + l)) * BS] = v[l];
			}
	}
	int ncount;
	const double ll = dev_ie3_finish<CM, BS>(mine, Dt, tt, pars, child, &ncount);
	dev_store_outputs(Dt, O, fid, child, pars, 3, ll, ncount);
}

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cnb_set_error("CUDA: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); return CNB_ERR_CUDA; } } while (0)

/* quads (4 sample words) per operand row for W sample words: whole groups of UM_BSTAGES stages plus one stage of
 * prefetch slack */
int cnbk_umma_row_quads(int W, int fp4, int *ngroups) {
	const int qps = fp4 ? 2 : 1;                          /* quads per stage */
	const int nst = (W + 4 * qps - 1) / (4 * qps), ng = (nst + UM_BSTAGES - 1) / UM_BSTAGES;
	if (ngroups) *ngroups = ng;
	return (ng * UM_BSTAGES + 1) * qps;
}

int cnbk_umma_rows_full(cudaStream_t st, const uint4 *planes, const uint32_t *keep, int N, int W, uint4 *rows_a, uint4 *rows_b) {
	const int nq = cnbk_umma_row_quads(W, 1, nullptr);
	dim3 grid((unsigned)((long long)nq * ((N + 127) / 128)));
	k_umma_rows_full<<<grid, 128, 0, st>>>(planes, keep, N, W, rows_a, rows_b, keep ? 8 : 4);
	CK(cudaGetLastError());
	return CNB_OK;
}

int cnbk_umma_rows(cudaStream_t st, const uint4 *planes, int N, int W, int fp4, uint4 *rows_a, uint4 *rows_b, int q_begin, int q_end) {
	const int nq = cnbk_umma_row_quads(W, fp4, nullptr);
	if (q_end < 0 || q_end > nq) q_end = nq;
	if (q_end <= q_begin) return CNB_OK;
	dim3 grid((unsigned)((long long)(q_end - q_begin) * ((N + 127) / 128)));
	k_umma_rows<<<grid, 128, 0, st>>>(planes, N, W, q_begin, rows_a, rows_b, fp4, 3 * N);
	CK(cudaGetLastError());
	return CNB_OK;
}

/* ---- three-parent families on the tensor cores (T tiles, cnb_tiles.h) ---------------------------------------------
 * pair-side rows of the real nodes (stride 3 (N + nvirt)) followed by those of the virtual pair nodes
 * v = (b(b-1)/2 + a) f1 + i: row j = [x_a = i][x_b = j], j < 3 (mxf4 only) */
__global__ void __launch_bounds__(128) k_umma_rows_virtual(const uint4 *__restrict__ planes, int N, int W, int nvirt, int f1,
		uint4 *__restrict__ rows_a) {
	const int v = blockIdx.x * blockDim.x + threadIdx.x, qd = blockIdx.y;
	if (v >= nvirt) return;
	int a, b;
	cnb_pair_of(v / f1, a, b);
	const int i = v % f1;
	uint32_t r[3][4];
#pragma unroll
	for (int k = 0; k < 4; k++) {
		const int w = qd * 4 + k;
		uint4 pa = make_uint4(0, 0, 0, 0), pb = pa;
		if (w < W) { pa = __ldg(planes + (size_t)w * N + a); pb = __ldg(planes + (size_t)w * N + b); }
		const uint32_t m = i == 0 ? pa.x : (i == 1 ? pa.y : pa.z);
		r[0][k] = m & pb.x;
		r[1][k] = m & pb.y;
		r[2][k] = m & pb.z;
	}
	const size_t o = (size_t)qd * (3 * (size_t)(N + nvirt)) + 3 * (size_t)(N + v);
#pragma unroll
	for (int j = 0; j < 3; j++) rows_a[o + j] = make_uint4(r[j][0], r[j][1], r[j][2], r[j][3]);
}

int cnbk_umma_rows3(cudaStream_t st, const uint4 *planes, int N, int W, int nvirt, int f1, uint4 *rows_a3) {
	const int nq = cnbk_umma_row_quads(W, 1, nullptr);
	if (nq > 65535) { cnb_set_error("three-parent tensor path: too many samples"); return CNB_ERR_PROC; }
	dim3 grid((unsigned)((long long)nq * ((N + 127) / 128)));
	k_umma_rows<<<grid, 128, 0, st>>>(planes, N, W, 0, rows_a3, nullptr, 1, 3 * (N + nvirt));
	CK(cudaGetLastError());
	dim3 gv((nvirt + 127) / 128, nq);
	k_umma_rows_virtual<<<gv, 128, 0, st>>>(planes, N, W, nvirt, f1, rows_a3);
	CK(cudaGetLastError());
	return CNB_OK;
}

/* how the last, partial wave of `tiles` equal tiles on `sms` CTA slots is cut: slices per tile (1 = not cut) */
static int tail_split(long long tiles, int sms, int ngroups) {
	const long long rest = tiles % sms;
	if (rest == 0 || tiles < sms || ngroups < 8) return 1;
	int best = 1;
	double best_t = 1.0;                              /* duration of the tail in tile units */
	for (int s = 2; s <= 4; s++) {
		const double t = (double)((rest * s + sms - 1) / sms) / s + 0.04 * s;   /* + prologue / drain per extra CTA */
		if (t < best_t - 1e-9) { best_t = t; best = s; }
	}
	return best;
}

int cnbk_umma_group_samples(int fp4) { return UM_BSTAGES * (fp4 ? 256 : 128); }

int cnbk_umma_tables(cudaStream_t st, const uint4 *rows_a, const uint4 *rows_b, int N, int W, int fp4,
		const CnbTiles &T, long long tile_begin, long long tile_end, int *counts, int g_begin, int g_end, int accumulate) {
	if (tile_end <= tile_begin) return CNB_OK;
	int ng_all;
	cnbk_umma_row_quads(W, fp4, &ng_all);
	if (g_end < 0 || g_end > ng_all) g_end = ng_all;
	if (g_end <= g_begin) return CNB_OK;
	const int ng = g_end - g_begin;
	static int sms = 0;
	if (!sms) {
		int dev = 0;
		CK(cudaGetDevice(&dev));
		CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
	}
	/* one CTA per SM (all of tensor memory): with T tiles the last wave holds T mod SMs of them; those are cut into
	 * sample slices so that the wave is full and shorter (1535 tiles per rank at C5 on 8 GPUs: 10.5 instead of 11 waves) */
	const long long tiles = tile_end - tile_begin;
	static int pf_dist = -1;
	if (pf_dist < 0) { const char *e = getenv("CNB_PF_DIST"); pf_dist = e ? atoi(e) : 1; }
	CnbTiles Tp = T;
	Tp.pad = pf_dist;
	const int split = T.self_pairs ? 1 : tail_split(tiles, sms, ng);
	const long long whole = split > 1 ? tiles - tiles % sms : tiles;
	const long long grid = whole + (tiles - whole) * split;
	const size_t tile_ints = T.full ? (size_t)UMF_SLOTS * UMF_SLOT_INTS : (size_t)T.tp * T.tc * cnb_geo_slot_ints(T.fr);
	if (split > 1 && !accumulate)
		CK(cudaMemsetAsync(counts + (size_t)whole * tile_ints, 0, (size_t)(tiles - whole) * tile_ints * sizeof(int), st));
#define UM_LAUNCH(FP4_, GEO_) do { \
		CK(cudaFuncSetAttribute(k_umma_tables<FP4_, GEO_>, cudaFuncAttributeMaxDynamicSharedMemorySize, UM_SMEM)); \
		k_umma_tables<FP4_, GEO_><<<(unsigned)grid, UM_THREADS, UM_SMEM, st>>>(rows_a, rows_b, N, g_begin, g_end, Tp, tile_begin, counts, \
				(int)whole, split, accumulate); } while (0)
	if (!T.full && (T.tp != cnb_geo_pairs(T.fr) || T.tc != cnb_geo_child(T.fr) || T.fr < 1 || T.fr > 3)) {
		cnb_set_error("table kernel: tile geometry %d x %d does not belong to %d free categories", T.tp, T.tc, T.fr);
		return CNB_ERR_PROC;
	}
	if (!fp4 && (T.full || T.fr != 3)) { cnb_set_error("full-table and reduced tiles need E2M1 operands"); return CNB_ERR_PROC; }
	if (T.full) UM_LAUNCH(true, 1);
	else if (!fp4) UM_LAUNCH(false, 0);
	else if (T.fr == 3) UM_LAUNCH(true, 0);
	else if (T.fr == 2) UM_LAUNCH(true, 2);
	else UM_LAUNCH(true, 3);
#undef UM_LAUNCH
	CK(cudaGetLastError());
	return CNB_OK;
}

long long cnbk_umma_column_macs(int W, int fp4) {
	int ng;
	cnbk_umma_row_quads(W, fp4, &ng);
	return (long long)UM_NA * 128 * ((long long)ng * UM_BSTAGES * (fp4 ? 256 : 128));
}

int cnbk_umma_epilogue(cudaStream_t st, const DevData &Dt, long long tile_begin, long long tile_end, const int *counts,
		const int *pair_tab, const ScoreOut &O, bool all_four_cats) {
	const bool full = Dt.tiles.full != 0;
	long long slots = (tile_end - tile_begin) * (full ? UMF_SLOTS : (long long)Dt.tiles.tp * Dt.tiles.tc);
	if (slots <= 0) return CNB_OK;
	const unsigned grid = (unsigned)((slots + 127) / 128);
	if (full && all_four_cats) k_umma_epilogue<true, true><<<grid, 128, 0, st>>>(Dt, tile_begin, tile_end, counts, pair_tab, O);
	else if (full) k_umma_epilogue<false, true><<<grid, 128, 0, st>>>(Dt, tile_begin, tile_end, counts, pair_tab, O);
	else if (all_four_cats) k_umma_epilogue<true, false><<<grid, 128, 0, st>>>(Dt, tile_begin, tile_end, counts, pair_tab, O);
	else k_umma_epilogue<false, false><<<grid, 128, 0, st>>>(Dt, tile_begin, tile_end, counts, pair_tab, O);
	CK(cudaGetLastError());
	return CNB_OK;
}

int cnbk_umma_export(cudaStream_t st, const DevData &Dt, long long f_begin, long long f_end, const int *ex_child, const int *ex_pars,
		const int *counts, const int *pair_tab, const ScoreOut &O) {
	if (f_end <= f_begin) return CNB_OK;
	const unsigned grid = (unsigned)((f_end - f_begin + 127) / 128);
	if (Dt.tiles.full) k_umma_export<true><<<grid, 128, 0, st>>>(Dt, f_begin, f_end, ex_child, ex_pars, counts, pair_tab, O);
	else k_umma_export<false><<<grid, 128, 0, st>>>(Dt, f_begin, f_end, ex_child, ex_pars, counts, pair_tab, O);
	CK(cudaGetLastError());
	return CNB_OK;
}

/* two-way tables of all node pairs on the tensor cores (label space, once per data set): operand rows of the label-space
 * planes, P tiles, conversion.  rows_a/rows_b/counts are scratch of the caller's size (cnbk_umma_row_quads,
 * cnbk_pair_tiles * slots). */
long long cnbk_pair_tiles(int N) {
	return (long long)((N + UM_PAIRS - 1) / UM_PAIRS) * ((N + UM_CHILD - 1) / UM_CHILD);
}

int cnbk_pair_tables_umma(cudaStream_t st, const uint4 *planes_lbl, int N, int M, int W, uint4 *rows_a, uint4 *rows_b,
		int *counts, int *pair_tab) {
	int rc = cnbk_umma_rows(st, planes_lbl, N, W, 1, rows_a, rows_b);
	if (rc != CNB_OK) return rc;
	CnbTiles T;
	memset(&T, 0, sizeof(T));
	CNB_TILES_DEFAULT_GEOMETRY(T);
	T.self_pairs = 1;
	T.world = 1;
	rc = cnbk_umma_tables(st, rows_a, rows_b, N, W, 1, T, 0, cnbk_pair_tiles(N), counts);
	if (rc != CNB_OK) return rc;
	const long long np = (long long)N * (N - 1) / 2;
	k_pairs_from_tiles<<<(unsigned)((np + 127) / 128), 128, 0, st>>>(counts, N, M, pair_tab);
	CK(cudaGetLastError());
	return CNB_OK;
}

/* ordered pair tables from the full-table P tiles: "pair" slot u is the single node u, whose rows (i, i) are its plain
 * one-hot rows; column node v carries its keep mask: cell ((i, i), k) = n(x_u = i, x_v = k | v kept, both present) */
__global__ void __launch_bounds__(128) k_ordered_pairs_from_tiles(const int *__restrict__ counts, int N, int *__restrict__ pair_ord) {
	const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= (long long)N * N) return;
	const int u = (int)(q / N), v = (int)(q % N);
	const int nch = (N + UMF_CHILD - 1) / UMF_CHILD;
	const long long tile = (long long)(u / UMF_PAIRS) * nch + v / UMF_CHILD;
	const int *src = counts + ((size_t)tile * UMF_SLOTS + (size_t)(u % UMF_PAIRS) * UMF_CHILD + v % UMF_CHILD) * UMF_SLOT_INTS;
	int4 *out = reinterpret_cast<int4 *>(pair_ord + q * 16);
#pragma unroll
	for (int i = 0; i < 4; i++) out[i] = *reinterpret_cast<const int4 *>(src + (i * 4 + i) * 4);
}

long long cnbk_pair_tiles_full(int N) {
	return (long long)((N + UMF_PAIRS - 1) / UMF_PAIRS) * ((N + UMF_CHILD - 1) / UMF_CHILD);
}

int cnbk_pair_tables_umma_full(cudaStream_t st, const uint4 *planes_lbl, const uint32_t *keep_lbl, int N, int W, uint4 *rows_a,
		uint4 *rows_b, int *counts, int *pair_ord) {
	int rc = cnbk_umma_rows_full(st, planes_lbl, keep_lbl, N, W, rows_a, rows_b);
	if (rc != CNB_OK) return rc;
	CnbTiles T;
	memset(&T, 0, sizeof(T));
	T.tp = UMF_PAIRS;
	T.tc = UMF_CHILD;
	T.full = 1;
	T.row_stride = keep_lbl ? 8 : 4;
	T.self_pairs = 1;
	T.world = 1;
	rc = cnbk_umma_tables(st, rows_a, rows_b, N, W, 1, T, 0, cnbk_pair_tiles_full(N), counts);
	if (rc != CNB_OK) return rc;
	const long long np = (long long)N * N;
	k_ordered_pairs_from_tiles<<<(unsigned)((np + 127) / 128), 128, 0, st>>>(counts, N, pair_ord);
	CK(cudaGetLastError());
	return CNB_OK;
}

/* P tiles over one slice of the samples, on operand rows that already exist (order or label space) */
long long cnbk_pair_tile_run(cudaStream_t st, const uint4 *rows_a, const uint4 *rows_b, int N, int W, int *counts, int g_begin,
		int g_end, int accumulate) {
	CnbTiles T;
	memset(&T, 0, sizeof(T));
	CNB_TILES_DEFAULT_GEOMETRY(T);
	T.self_pairs = 1;
	T.world = 1;
	return cnbk_umma_tables(st, rows_a, rows_b, N, W, 1, T, 0, cnbk_pair_tiles(N), counts, g_begin, g_end, accumulate);
}

int cnbk_pairs_from_tiles_order(cudaStream_t st, const int *counts, const int *order, int N, int M, int *pair_tab) {
	const long long np = (long long)N * (N - 1) / 2;
	if (np < 1) return CNB_OK;
	k_pairs_from_tiles_order<<<(unsigned)((np + 127) / 128), 128, 0, st>>>(counts, order, N, M, pair_tab);
	CK(cudaGetLastError());
	return CNB_OK;
}

/* T tiles [tile_begin, tile_end) of one column tile (T.self_pairs == 2, T.tile_base on the device) */
template <int CM, int BS>
static int launch_epilogue3(cudaStream_t st, const DevData &Dt, const DevFamilies &F, long long b, long long e, int c0,
		const int *counts, const int *tt, const ScoreOut &O) {
	const size_t smem = (size_t)CM * CM * CM * CM * BS * sizeof(int);
	if (smem > 48 * 1024)
		CK(cudaFuncSetAttribute(k_umma_epilogue3<CM, BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	k_umma_epilogue3<CM, BS><<<(unsigned)((e - b + BS - 1) / BS), BS, smem, st>>>(Dt, F, b, e, c0, counts, tt, O);
	CK(cudaGetLastError());
	return CNB_OK;
}

template <int CM, int BS>
static int launch_epilogue3_range(cudaStream_t st, const DevData &Dt, const DevFamilies &F, long long t_begin, long long t_end,
		long long q_base, int c0, int ce, const int *counts, const int *tt, const ScoreOut &O) {
	const size_t smem = (size_t)CM * CM * CM * CM * BS * sizeof(int);
	if (smem > 48 * 1024)
		CK(cudaFuncSetAttribute(k_umma_epilogue3_range<CM, BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	const int first = c0 > 3 ? c0 : 3;
	if (ce <= first) return CNB_OK;
	dim3 grid((unsigned)((t_end - t_begin + BS - 1) / BS), (unsigned)(ce - first));
	k_umma_epilogue3_range<CM, BS><<<grid, BS, smem, st>>>(Dt, F, t_begin, t_end, q_base, c0, ce, counts, tt, O);
	CK(cudaGetLastError());
	return CNB_OK;
}

/* triples [t_begin, t_end) x the children of column tile [c0, ce); counts starts at the tile that holds slot q_base */
int cnbk_umma_epilogue3_range(cudaStream_t st, const DevData &Dt, const DevFamilies &F, int max_cats, long long t_begin, long long t_end,
		long long q_base, int c0, int ce, const int *counts, const int *triple_tab, const ScoreOut &O) {
	if (t_end <= t_begin) return CNB_OK;
	if (max_cats <= 2) return launch_epilogue3_range<2, 128>(st, Dt, F, t_begin, t_end, q_base, c0, ce, counts, triple_tab, O);
	if (max_cats == 3) return launch_epilogue3_range<3, 128>(st, Dt, F, t_begin, t_end, q_base, c0, ce, counts, triple_tab, O);
	return launch_epilogue3_range<4, 64>(st, Dt, F, t_begin, t_end, q_base, c0, ce, counts, triple_tab, O);
}

/* families [b, e) = those whose child lies in the column tile that starts at child c0; counts holds the tables of that
 * column tile's tiles from its first one on */
int cnbk_umma_epilogue3(cudaStream_t st, const DevData &Dt, const DevFamilies &F, int max_cats, long long b, long long e,
		int c0, const int *counts, const int *triple_tab, const ScoreOut &O) {
	if (e <= b) return CNB_OK;
	if (max_cats <= 2) return launch_epilogue3<2, 128>(st, Dt, F, b, e, c0, counts, triple_tab, O);
	if (max_cats == 3) return launch_epilogue3<3, 128>(st, Dt, F, b, e, c0, counts, triple_tab, O);
	return launch_epilogue3<4, 64>(st, Dt, F, b, e, c0, counts, triple_tab, O);
}
