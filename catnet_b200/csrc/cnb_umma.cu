/* cnb_umma.cu -- two-parent family tables on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a only.
 *
 * The contingency table of family (a, b | c) is a contraction over the sample axis:
 *     n(i, j, k) = sum_s  [x_a(s) = i][x_b(s) = j] * [x_c(s) = k]
 * i.e. D = A * B^T with A = one-hot rows of the pair configurations (pair (a,b) x 16 configurations (i,j)) and
 * B = one-hot rows of the child categories (child c x 4 categories k), K = samples.  Both operands are 0/1 bytes,
 * generated ON THE FLY from the bit planes (HBM/L2 traffic stays at one bit per sample and category), staged in
 * shared memory in the K-major SWIZZLE_128B canonical layout and multiplied with tcgen05.mma.kind::i8 (u8 x u8 ->
 * s32, exact), accumulators in TMEM.  One CTA = one 128 x 256 tile over all samples; two CTAs per SM so that one
 * CTA's operand generation overlaps the other's MMAs.
 *
 * Replaces, for unrestricted two-parent candidate sets without missing values, the counting loop of
 * CATNET::setNodeSampleProb (/root/reference/src/catnet_class.h:842-857); the log-lik epilogue
 * (problist.h:224-250) runs in k_umma_epilogue over the integer tables, in the reference's operation order.
 */
#include <stdio.h>
#include "cnb_device.cuh"
#include "cnb_kernels.h"

/* Inclusion-exclusion (see k_score_planes_ie2): only the 3 x 3 x 3 "free" cells of every table are contracted;
 * cells with a last category follow from the pair tables.  So a pair needs 9 rows and a child 3 columns:
 * 14 pairs (126 of 128 rows) x 85 children (255 of 256 columns) = 1190 families per 128 x 256 tile. */
#define UM_PAIRS CNB_TILE_PAIRS
#define UM_CHILD CNB_TILE_CHILD
#define UM_SLOTS (UM_PAIRS * UM_CHILD)
#define UM_SLOT_INTS 36                   /* 9 configurations x (3 counts + 1 pad): int4 stores */
#define UM_ROWS_A 128
#define UM_ROWS_B 256
#define UM_BK 128                         /* bytes (= samples) per stage: one 128B swizzle row, 4 words */
#define UM_THREADS 256
#define UM_STAGE_A (UM_ROWS_A * UM_BK)    /* 16 KB */
#define UM_STAGE_B (UM_ROWS_B * UM_BK)    /* 32 KB */
#define UM_STAGE (UM_STAGE_A + UM_STAGE_B)
#define UM_SMEM (2 * UM_STAGE + 1024)     /* two stages + alignment slack */
#define UM_TMEM_COLS 256

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
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
/* D[tmem] (+)= A[smem desc] * B[smem desc]^T, u8 x u8 -> s32, M = 128, N = 256, K = 32 */
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
	uint32_t z = 0;
	asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
			"tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n"
			::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(z) : "memory");
}
/* K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 in [0,14),
 * LBO = 1 in [16,30), SBO = 1024 B >> 4 in [32,46) (stride between 8-row groups), version 1 in [46,48),
 * layout SWIZZLE_128B = 2 in [61,64) */
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
	return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

/* byte offset of (row r, 16-byte chunk ch) in the K-major SWIZZLE_128B layout: 8-row groups of 1024 B, the chunk
 * index XOR-ed with the row inside the group (Swizzle<3,4,3>) */
__device__ __forceinline__ uint32_t sw128(uint32_t r, uint32_t ch) {
	return (r >> 3) * 1024u + (r & 7u) * 128u + ((ch ^ (r & 7u)) << 4);
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

__global__ void __launch_bounds__(UM_THREADS, 2) k_umma_tables(const uint4 *__restrict__ planes,
		const uint4 *__restrict__ planes_rev, int N, int W, const long long *__restrict__ tile_base, int nct,
		long long tile_begin, int *__restrict__ counts) {
	extern __shared__ uint8_t smem_raw[];
	__shared__ __align__(8) uint64_t bar_empty[2], bar_done;
	__shared__ uint32_t tmem_base_sh;
	uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const long long tile = tile_begin + blockIdx.x;
	int ct;
	long long pt;
	tile_of(tile_base, nct, tile, ct, pt);
	const int c0 = ct * UM_CHILD;

	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)), "n"(UM_TMEM_COLS));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
	}
	if (tid == 32) {
		mbar_init(&bar_empty[0], 1);
		mbar_init(&bar_empty[1], 1);
		mbar_init(&bar_done, 1);
		asm volatile("fence.mbarrier_init.release.cluster;");
	}
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;");
	const uint32_t tmem = tmem_base_sh;

	/* ---- per-thread generation roles (fixed for the whole K loop) ----
	 * task = (row, 16-byte chunk); 256 threads -> chunk = tid & 7 (word tid>>1 & 3, half tid & 1), rows tid>>3 + 32k.
	 * A row r = pair*9 + i*3 + j (i, j < 3); B row r = child*3 + k (k < 3).  Unused rows are zeroed once. */
	const int ch = tid & 7, wi = ch >> 1, u0 = (ch & 1) * 4, r0 = tid >> 3;
	int a_na[4], a_nb[4], a_ci[4], a_cj[4];          /* node positions and categories of the 4 A rows; -1 = no row */
	uint32_t a_off[4];
#pragma unroll
	for (int k = 0; k < 4; k++) {
		int r = r0 + 32 * k;
		a_off[k] = sw128(r, ch);
		a_na[k] = -1;
		if (r < UM_PAIRS * 9) {
			int pa, pb;
			pair_of(pt * UM_PAIRS + r / 9, pa, pb);
			if (pb < N) { a_na[k] = pa; a_nb[k] = pb; a_ci[k] = (r % 9) / 3; a_cj[k] = r % 3; }
		}
	}
	int b_n[8], b_ck[8];
	uint32_t b_off[8];
#pragma unroll
	for (int k = 0; k < 8; k++) {
		int r = r0 + 32 * k;
		b_off[k] = UM_STAGE_A + sw128(r, ch);
		int c = c0 + r / 3;
		b_n[k] = (r < UM_CHILD * 3 && c < N) ? c : -1;
		b_ck[k] = r % 3;
	}
	/* rows that never receive data must read as zero in both stages */
	for (int st = 0; st < 2; st++) {
#pragma unroll
		for (int k = 0; k < 4; k++)
			if (a_na[k] < 0) *reinterpret_cast<uint4 *>(smem + st * UM_STAGE + a_off[k]) = make_uint4(0, 0, 0, 0);
#pragma unroll
		for (int k = 0; k < 8; k++)
			if (b_n[k] < 0) *reinterpret_cast<uint4 *>(smem + st * UM_STAGE + b_off[k]) = make_uint4(0, 0, 0, 0);
	}

	const int nst = (W + 3) / 4;
	const uint32_t idesc = (2u << 4) | ((UM_ROWS_B >> 3) << 17) | ((UM_ROWS_A >> 4) << 24);   /* s32 acc, u8 x u8, K-major */

	/* prefetch stage 0 operands (one 32-bit plane word per task) */
	uint32_t fa[4], fb[4], fc[8];
	auto fetch = [&](int st) {
		const int w = st * 4 + wi;
		const bool in = w < W;
		const uint32_t *row = reinterpret_cast<const uint32_t *>(planes + (size_t)(in ? w : 0) * N);
		const uint32_t *rrow = reinterpret_cast<const uint32_t *>(planes_rev + (size_t)(in ? w : 0) * N);
#pragma unroll
		for (int k = 0; k < 4; k++) {
			fa[k] = fb[k] = 0;
			if (in && a_na[k] >= 0) {
				fa[k] = __ldg(row + 4 * a_na[k] + a_ci[k]);
				fb[k] = __ldg(row + 4 * a_nb[k] + a_cj[k]);
			}
		}
#pragma unroll
		for (int k = 0; k < 8; k++) fc[k] = (in && b_n[k] >= 0) ? __ldg(rrow + 4 * b_n[k] + b_ck[k]) : 0;
	};
	fetch(0);

	for (int it = 0; it < nst; it++) {
		const int s = it & 1;
		uint8_t *buf = smem + s * UM_STAGE;
		/* Operand bytes without shifts or multiplies.  Inside a 32-sample word the K position 4u+v holds sample
		 * u + 8v: the A byte is  x & (0x01 << u)  = bit * 2^u, the B byte comes from the plane word with the bits of
		 * every byte reversed,  y' & (0x01 << (7-u))  = bit * 2^(7-u); every product is bit*bit*128, so the
		 * accumulators hold 128 x count (exact in s32 for M < 2^24; the epilogue shifts right by 7). */
		uint32_t ga[4], gc[8];
#pragma unroll
		for (int k = 0; k < 4; k++) ga[k] = fa[k] & fb[k];
#pragma unroll
		for (int k = 0; k < 8; k++) gc[k] = fc[k];
		if (it + 1 < nst) fetch(it + 1);                  /* loads for the next stage fly during this one */
		if (it >= 2) mbar_wait(&bar_empty[s], ((it >> 1) - 1) & 1);   /* MMAs of stage it-2 have read this buffer */
#pragma unroll
		for (int k = 0; k < 4; k++)
			if (a_na[k] >= 0)
				*reinterpret_cast<uint4 *>(buf + a_off[k]) = make_uint4(ga[k] & (0x01010101u << u0), ga[k] & (0x02020202u << u0),
						ga[k] & (0x04040404u << u0), ga[k] & (0x08080808u << u0));
#pragma unroll
		for (int k = 0; k < 8; k++)
			if (b_n[k] >= 0)
				*reinterpret_cast<uint4 *>(buf + b_off[k]) = make_uint4(gc[k] & (0x80808080u >> u0), gc[k] & (0x40404040u >> u0),
						gc[k] & (0x20202020u >> u0), gc[k] & (0x10101010u >> u0));
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   /* generic-proxy stores -> async proxy (MMA) */
		__syncthreads();
		if (warp == 0) {
			if (lane == 0) {
				asm volatile("tcgen05.fence::after_thread_sync;");
				const uint64_t da = umma_desc(smem_u32(buf)), db = umma_desc(smem_u32(buf + UM_STAGE_A));
#pragma unroll
				for (int k = 0; k < UM_BK / 32; k++)          /* K = 32 bytes per instruction: +2 in 16-byte units */
					umma_i8(tmem, da + 2 * k, db + 2 * k, idesc, (it | k) != 0);
				umma_commit(&bar_empty[s]);
				if (it == nst - 1) umma_commit(&bar_done);
			}
			__syncwarp();
		}
	}
	mbar_wait(&bar_done, 0);
	asm volatile("tcgen05.fence::after_thread_sync;");

	/* ---- epilogue: TMEM -> registers -> global tables.  Row = pair*9 + (i*3+j); column = child*3 + k.
	 * A warp may only touch its own lane quarter (warp & 3); the two warp halves split the columns in 48-column
	 * chunks (16 children): chunks 0-2 / 3-5 (the last one holds 5 children + the unused column). */
	const int lq = warp & 3, hh = warp >> 2;
	const int row = lq * 32 + lane, p = row / 9, cfg = row % 9;
	const bool row_ok = row < UM_PAIRS * 9;
	int *out = counts + ((size_t)blockIdx.x * UM_SLOTS + (size_t)p * UM_CHILD) * UM_SLOT_INTS + cfg * 4;
#pragma unroll 1
	for (int q = hh * 3; q < hh * 3 + 3; q++) {
		uint32_t v[48];
		const uint32_t taddr = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(q * 48);
#define LD16(o, a) asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
		: "=r"(v[o]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7]), \
		  "=r"(v[o + 8]), "=r"(v[o + 9]), "=r"(v[o + 10]), "=r"(v[o + 11]), "=r"(v[o + 12]), "=r"(v[o + 13]), "=r"(v[o + 14]), "=r"(v[o + 15]) \
		: "r"(a))
		LD16(0, taddr);
		if (q < 5) {
			LD16(16, taddr + 16);
			LD16(32, taddr + 32);
		}
#undef LD16
		asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
		const int nchild = q < 5 ? 16 : UM_CHILD - 80;
		if (row_ok) {
#pragma unroll
			for (int j16 = 0; j16 < 16; j16++)
				if (j16 < nchild)
					*reinterpret_cast<int4 *>(out + (size_t)(q * 16 + j16) * UM_SLOT_INTS) =
							make_int4((int)(v[j16 * 3] >> 7), (int)(v[j16 * 3 + 1] >> 7), (int)(v[j16 * 3 + 2] >> 7), 0);
		}
	}
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	if (warp == 0)
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(UM_TMEM_COLS));
}

/* log-lik epilogue over the tables of a tile range: one thread per (tile, pair, child) slot.  Completes the table
 * from the 27 free cells and the pair tables (inclusion-exclusion), then the reference's fp64 sum.  The score goes
 * to the slot's position in the rank-major buffer (slot id = tile*1190 + pair*85 + child plays the family id). */
__device__ __forceinline__ int um_pair_count(const int *__restrict__ pt, int u, int v, int i, int j) {
	return u < v ? pt[((long long)v * (v - 1) / 2 + u) * 16 + i * 4 + j] : pt[((long long)u * (u - 1) / 2 + v) * 16 + j * 4 + i];
}

__global__ void __launch_bounds__(128) k_umma_epilogue(DevData Dt, const long long *__restrict__ tile_base, int nct,
		long long tile_begin, long long tile_end, const int *__restrict__ counts, const int *__restrict__ pair_tab,
		ScoreOut O) {
	const long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const long long tile = tile_begin + slot / UM_SLOTS;
	if (tile >= tile_end) return;
	const int within = (int)(slot % UM_SLOTS), p = within / UM_CHILD, j = within % UM_CHILD;
	int ct;
	long long pt;
	tile_of(tile_base, nct, tile, ct, pt);
	int a, b;
	pair_of(pt * UM_PAIRS + p, a, b);
	const int c = ct * UM_CHILD + j;
	if (c >= Dt.N || b >= c) return;                      /* not a family: parents must precede the child */
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
	O.scores[dev_score_index(tile * UM_SLOTS + within, O.chunk, O.seg_off, O.seg_len)] = score;
}

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cnb_set_error("CUDA: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); return CNB_ERR_CUDA; } } while (0)

int cnbk_umma_tables(cudaStream_t st, const uint4 *planes, const uint4 *planes_rev, int N, int W,
		const long long *tile_base, int nct, long long tile_begin, long long tile_end, int *counts) {
	if (tile_end <= tile_begin) return CNB_OK;
	CK(cudaFuncSetAttribute(k_umma_tables, cudaFuncAttributeMaxDynamicSharedMemorySize, UM_SMEM));
	k_umma_tables<<<(unsigned)(tile_end - tile_begin), UM_THREADS, UM_SMEM, st>>>(planes, planes_rev, N, W, tile_base, nct, tile_begin, counts);
	CK(cudaGetLastError());
	return CNB_OK;
}

int cnbk_umma_epilogue(cudaStream_t st, const DevData &Dt, const long long *tile_base, int nct, long long tile_begin,
		long long tile_end, const int *counts, const int *pair_tab, const ScoreOut &O) {
	long long slots = (tile_end - tile_begin) * UM_SLOTS;
	if (slots <= 0) return CNB_OK;
	k_umma_epilogue<<<(unsigned)((slots + 127) / 128), 128, 0, st>>>(Dt, tile_base, nct, tile_begin, tile_end, counts, pair_tab, O);
	CK(cudaGetLastError());
	return CNB_OK;
}
