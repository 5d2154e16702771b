/* cnb_api.cu -- the C ABI (include/catnet_b200.h): device context, orchestration of K1..K5, result export.
 *
 * Call structure mirrors the reference: cnb_search_order = RCatnetSearch::estimateCatnets
 * (src/rcatnet_search.cpp:57-312) around CATNET_SEARCH2::estimate (src/catnet_search2.h:152-897), with the scoring,
 * selection and DP on the device.  No CPU fallback exists: without a CUDA device every scoring call returns
 * CNB_ERR_CUDA.
 */
#include <stdio.h>
#include <math.h>
#include <string.h>
#include <algorithm>
#include <functional>
#include <chrono>
#include <map>
#include <memory>
#include "cnb_device.cuh"
#include "cnb_kernels.h"
#include "cnb_result.h"
#include "cnb_ctx.h"
#include "cnb_crew.h"
#include "cnb_tiles.h"

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cnb_set_error("CUDA: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); return CNB_ERR_CUDA; } } while (0)
#define RC(call) do { int rc_ = (call); if (rc_ != CNB_OK) return rc_; } while (0)
#define H2D(c, dst, src, n) do { CK(cudaMemcpyAsync((dst), (src), (n), cudaMemcpyHostToDevice, (c)->stream)); (c)->timing.h2d_bytes += (int64_t)(n); } while (0)
#define D2H(c, dst, src, n) do { CK(cudaMemcpyAsync((dst), (src), (n), cudaMemcpyDeviceToHost, (c)->stream)); (c)->timing.d2h_bytes += (int64_t)(n); } while (0)

extern "C" const char *cnb_version(void) { return "catnet_b200 0.1 (sm_100a)"; }

/* ------------------------------------------------------------------ context ---- */
int DevBuf::ensure(size_t bytes) {
	if (bytes <= cap) return CNB_OK;
	if (view) { ptr = nullptr; view = false; }            /* outgrown its place in the packed plan: its own allocation again */
	if (ptr) cudaFree(ptr);
	ptr = nullptr;
	cap = 0;
	size_t want = bytes + bytes / 8 + 256;
	cudaError_t e = cudaMalloc(&ptr, want);
	if (e != cudaSuccess) {
		cnb_set_error("Insufficient device memory (%zu bytes): %s", want, cudaGetErrorString(e));
		return CNB_ERR_MEM;
	}
	cap = want;
	return CNB_OK;
}
void DevBuf::release() {
	if (ptr && !view) cudaFree(ptr);
	ptr = nullptr;
	cap = 0;
	view = false;
}
void DevBuf::alias(void *p, size_t bytes) {
	if (ptr && !view) cudaFree(ptr);
	ptr = p;
	cap = bytes;
	view = true;
}

extern "C" int cnb_ctx_create(int32_t device, cnb_ctx **out) {
	if (!out) return CNB_ERR_PARAM;
	*out = nullptr;
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev < 1) {
		cnb_set_error("no CUDA device: the scoring path has no CPU fallback (%s)", cudaGetErrorString(e));
		return CNB_ERR_CUDA;
	}
	if (device < 0 || device >= ndev) {
		cnb_set_error("invalid device %d (have %d)", device, ndev);
		return CNB_ERR_PARAM;
	}
	CK(cudaSetDevice(device));
	cnb_ctx *c = new cnb_ctx();
	c->device = device;
	CK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
	c->stream = c->own_stream;
	for (int i = 0; i < CNB_NEV; i++) CK(cudaEventCreate(&c->ev[i]));
	if (const char *e = getenv("CNB_HIST_MODE")) c->hist_mode = c->hist_mode_default = std::max(0, std::min(2, atoi(e)));
	*out = c;
	return CNB_OK;
}

int PinBuf::ensure(size_t bytes) {
	if (bytes <= cap) return CNB_OK;
	if (ptr) cudaFreeHost(ptr);
	ptr = nullptr;
	cap = 0;
	if (cudaHostAlloc(&ptr, bytes, cudaHostAllocDefault) != cudaSuccess) {
		ptr = nullptr;
		cnb_set_error("Insufficient memory (pinned host, %zu bytes)", bytes);
		return CNB_ERR_MEM;
	}
	cap = bytes;
	return CNB_OK;
}
void PinBuf::release() {
	if (ptr) cudaFreeHost(ptr);
	ptr = nullptr;
	cap = 0;
}

extern "C" void cnb_ctx_destroy(cnb_ctx *c) {
	if (!c) return;
	if (c->crew) {
		c->crew->stop();
		delete c->crew;
		c->crew = nullptr;
	}
	for (cnb_ctx *s : c->siblings) cnb_ctx_destroy(s);
	c->siblings.clear();
	if (c->host_ms[3] > 0)
		fprintf(stderr, "[cnb trace] %.0f searches: plan %.3f score %.3f select+dp %.3f ms per search (host wall)\n", c->host_ms[3],
				c->host_ms[0] / c->host_ms[3], c->host_ms[1] / c->host_ms[3], c->host_ms[2] / c->host_ms[3]);
	cudaSetDevice(c->device);
	cudaStreamSynchronize(c->stream);
	c->pin_sel.release();
	c->pin_stage.release();
	c->pin_sum.release();
	c->pin_plan.release();
	for (DevBuf *b : c->all_bufs()) b->release();
	for (int i = 0; i < CNB_NEV; i++) cudaEventDestroy(c->ev[i]);
	for (cudaEvent_t ev : c->ev_slices) cudaEventDestroy(ev);
	if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
	if (c->prep_stream) cudaStreamDestroy(c->prep_stream);
	cudaStreamDestroy(c->own_stream);
	delete c;
}

/* ---- idle context kept for the one-shot entry points ----
 * Every .Call of the R shim is a one-shot: create a context, copy and pack the samples, search, destroy.  For small
 * problems creating the streams / events and allocating (then freeing) some forty device buffers costs more than the
 * search (C1: 5-16 ms against 1.1 ms).  So a released one-shot context is parked, one per device, with its grow-only
 * buffers, and the next one-shot on that device takes it over.  cnb_release_cache() frees the parked contexts. */
#include <mutex>
static std::mutex g_pool_mutex;
static cnb_ctx *g_pool[64];

int cnb_ctx_acquire(int32_t device, cnb_ctx **out) {
	{
		std::lock_guard<std::mutex> lock(g_pool_mutex);
		if (device >= 0 && device < 64 && g_pool[device]) {
			*out = g_pool[device];
			g_pool[device] = nullptr;
			(*out)->have_samples = (*out)->planned = (*out)->dp_done = false;
			return CNB_OK;
		}
	}
	return cnb_ctx_create(device, out);
}

void cnb_ctx_park(cnb_ctx *c) {
	if (!c) return;
	cudaSetDevice(c->device);
	/* a context whose last call left a CUDA error behind is not worth keeping */
	if (cudaStreamSynchronize(c->stream) != cudaSuccess || cudaPeekAtLastError() != cudaSuccess) {
		cnb_ctx_destroy(c);
		return;
	}
	c->stream = c->own_stream;
	{
		std::lock_guard<std::mutex> lock(g_pool_mutex);
		if (c->device >= 0 && c->device < 64 && !g_pool[c->device]) {
			g_pool[c->device] = c;
			return;
		}
	}
	cnb_ctx_destroy(c);
}

void cnb_pool_release(void) {
	for (int d = 0; d < 64; d++) {
		cnb_ctx *c = nullptr;
		{
			std::lock_guard<std::mutex> lock(g_pool_mutex);
			c = g_pool[d];
			g_pool[d] = nullptr;
		}
		if (c) cnb_ctx_destroy(c);
	}
}

static float ev_ms(cnb_ctx *c, int a, int b) {
	float ms = 0;
	if (c->quiet) return 0;                               /* no per-phase events were recorded (EVREC) */
	cudaEventElapsedTime(&ms, c->ev[a], c->ev[b]);
	return ms;
}
/* per-phase timing event on the context's stream; the SA / histogram drivers evaluate thousands of small orders whose cost
 * is the number of driver calls per order, so they switch these off (cnb_ctx.quiet: cnb_timing's phase times read 0) */
#define EVREC(c, id) do { if (!(c)->quiet) CK(cudaEventRecord((c)->ev[(id)], (c)->stream)); } while (0)

/* samples_dev: [M][N] int32 (elem_bytes 4) or uint8 with 0 = missing (elem_bytes 1) in device memory */
int cnb_ctx_set_samples_impl(cnb_ctx *c, int32_t N, int32_t M, const void *samples_dev, int elem_bytes,
		const int32_t *pert_dev, const int32_t *num_cats) {
	if (!c || N < 1 || M < 1 || !samples_dev) { cnb_set_error("Invalid sample"); return CNB_ERR_PARAM; }
	CK(cudaSetDevice(c->device));
	/* a failure part-way must not leave the previous data set looking valid next to the new sizes */
	c->have_samples = c->planned = c->dp_done = c->tables_resident = false;
	c->N = N;
	c->M = M;
	c->W = (M + 31) / 32;
	c->Mpad = c->W * 32;
	c->has_pert = pert_dev != nullptr;
	if (!c->in_host_set) c->timing.h2d_bytes = c->timing.d2h_bytes = 0;   /* traffic is counted per data set */
	cudaStream_t st = c->stream;
	CK(cudaEventRecord(c->ev[EV_PACK0], st));
	RC(c->b_vmin.ensure(N * sizeof(int)));
	RC(c->b_vmax.ensure(N * sizeof(int)));
	RC(c->b_cats_lbl.ensure(N * sizeof(int)));
	RC(c->b_flag.ensure(4 * sizeof(int)));
	std::vector<int> vmin(N, 1), cats(N);
	if (num_cats) {
		for (int i = 0; i < N; i++) {
			cats[i] = num_cats[i];
			if (cats[i] < 1 || cats[i] > 254) { cnb_set_error("Incorrect categories"); return CNB_ERR_PARAM; }
		}
	} else {
		/* categories inferred as max - min + 1 over the non-NA values (catnet_search2.h:208-235) */
		std::vector<int> hi(N, -INT32_MAX), lo(N, INT32_MAX);
		H2D(c, c->b_vmin.ptr, lo.data(), N * sizeof(int));
		H2D(c, c->b_vmax.ptr, hi.data(), N * sizeof(int));
		RC(cnbk_minmax(st, samples_dev, elem_bytes, N, M, (int *)c->b_vmin.ptr, (int *)c->b_vmax.ptr));
		D2H(c, lo.data(), c->b_vmin.ptr, N * sizeof(int));
		D2H(c, hi.data(), c->b_vmax.ptr, N * sizeof(int));
		CK(cudaStreamSynchronize(st));
		for (int i = 0; i < N; i++) {
			if (lo[i] > hi[i]) { cnb_set_error("node %d has no observed value", i + 1); return CNB_ERR_PARAM; }
			vmin[i] = lo[i];
			cats[i] = hi[i] - lo[i] + 1;
			if (cats[i] > 254) { cnb_set_error("more than 254 categories"); return CNB_ERR_PARAM; }
		}
	}
	c->cats_lbl = cats;
	c->max_cats = *std::max_element(cats.begin(), cats.end());
	H2D(c, c->b_vmin.ptr, vmin.data(), N * sizeof(int));
	H2D(c, c->b_cats_lbl.ptr, cats.data(), N * sizeof(int));
	CK(cudaMemsetAsync(c->b_flag.ptr, 0, 4 * sizeof(int), st));
	size_t words = (size_t)c->W * N;
	RC(c->b_codes.ensure((size_t)N * c->Mpad));
	RC(c->b_planes_lbl.ensure(words * sizeof(uint4)));
	RC(c->b_planes.ensure(words * sizeof(uint4)));
	if (c->has_pert) {
		RC(c->b_keep_lbl.ensure(words * sizeof(uint32_t)));
		RC(c->b_keep.ensure(words * sizeof(uint32_t)));
	}
	RC(cnbk_pack(st, samples_dev, elem_bytes, pert_dev, N, M, c->W, (int *)c->b_vmin.ptr, (int *)c->b_cats_lbl.ptr,
			(uint8_t *)c->b_codes.ptr, (uint4 *)c->b_planes_lbl.ptr,
			c->has_pert ? (uint32_t *)c->b_keep_lbl.ptr : nullptr, (int *)c->b_flag.ptr));
	int flags[4] = {0, 0, 0, 0};
	D2H(c, flags, c->b_flag.ptr, sizeof(flags));
	CK(cudaStreamSynchronize(st));
	if (flags[0]) { cnb_set_error("Incorrect categories"); return CNB_ERR_PARAM; }   /* catnet_search2.h:248 */
	/* nodes perturbed in EVERY sample: the reference scores their candidates on zero samples, setNodeSampleProb returns
	 * -FLT_MAX (catnet_class.h:824-826) and no candidate is ever taken (catnet_search2.h:637-640, 813-826): the planner
	 * gives such a node no candidate parent sets (cnb_ctx_plan) */
	c->never_kept.assign(N, 0);
	c->any_never_kept = false;
	if (c->has_pert) {
		std::vector<uint32_t> kp(words), any(N, 0);
		D2H(c, kp.data(), c->b_keep_lbl.ptr, words * sizeof(uint32_t));
		CK(cudaStreamSynchronize(st));
		for (size_t w = 0; w < (size_t)c->W; w++)
			for (int v = 0; v < N; v++) any[v] |= kp[w * N + v];
		for (int v = 0; v < N; v++)
			if (!any[v]) { c->never_kept[v] = 1; c->any_never_kept = true; }
	}
	/* two-way margins for the inclusion-exclusion scorer: only meaningful without missing values / masks */
	c->pairs_valid = false;
	c->pairs_full = false;
	c->triples_valid = false;
	c->has_na = flags[2] != 0;
	c->clean = !flags[2] && !c->has_pert && c->max_cats <= CNB_PLANE_CATS;   /* no NA, no masks, bit planes valid */
	/* the tables count ALL samples: with perturbation masks they serve only the "whole sample" side of the pairwise
	 * metrics (pairs_full), never the search */
	if (!flags[2] && c->max_cats <= CNB_PLANE_CATS && N >= 3 && !c->no_ie) {
		RC(c->b_pairs.ensure((size_t)N * (N - 1) / 2 * 16 * sizeof(int)));
		if (M >= 8192 && M < (1 << 24) && !c->pairs_popc && c->tensor_mode >= 0) {
			/* enough samples to fill the MMA pipeline: one-hot GEMM of all nodes against all nodes (18 -> 2 ms at C5) */
			const size_t row_bytes = (size_t)cnbk_umma_row_quads(c->W, 1, nullptr) * sizeof(uint4) * 3 * N;
			RC(c->b_rows_a.ensure(row_bytes));
			RC(c->b_rows_b.ensure(row_bytes));
			RC(c->b_counts.ensure((size_t)cnbk_pair_tiles(N) * CNB_TILE_PAIRS * CNB_TILE_CHILD * CNBK_UMMA_SLOT_INTS * sizeof(int)));
			RC(cnbk_pair_tables_umma(st, (const uint4 *)c->b_planes_lbl.ptr, N, M, c->W, (uint4 *)c->b_rows_a.ptr,
					(uint4 *)c->b_rows_b.ptr, (int *)c->b_counts.ptr, (int *)c->b_pairs.ptr));
		} else
			RC(cnbk_pair_tables(st, (const uint4 *)c->b_planes_lbl.ptr, N, c->W, (int *)c->b_pairs.ptr));
		c->pairs_valid = !c->has_pert;
		c->pairs_full = true;
	}
	/* missing values or perturbation masks: the tables of all ORDERED pairs (parent, child) -- under the child's keep mask,
	 * samples with a missing value skipped -- on the tensor cores (full-table P tiles), so that the 0- and 1-parent families
	 * need no pass over the samples either (124,750 one-parent families at C5: 17 ms of POPC per order otherwise) */
	c->pairs_ord = false;
	if ((c->has_na || c->has_pert) && c->max_cats <= CNB_PLANE_CATS && M < (1 << 24) && c->tensor_mode >= 0 && !c->pairs_popc
			&& !c->no_ie && !c->tensor_i8) {
		const size_t row_bytes = (size_t)cnbk_umma_row_quads(c->W, 1, nullptr) * sizeof(uint4) * (c->has_pert ? 8 : 4) * N;
		RC(c->b_rows_a.ensure(row_bytes));
		RC(c->b_rows_b.ensure(row_bytes));
		RC(c->b_counts.ensure((size_t)cnbk_pair_tiles_full(N) * CNB_FULL_PAIRS * CNB_FULL_CHILD * CNBK_UMMA_FULL_SLOT_INTS * sizeof(int)));
		RC(c->b_pairs_ord.ensure((size_t)N * N * 16 * sizeof(int)));
		RC(cnbk_pair_tables_umma_full(st, (const uint4 *)c->b_planes_lbl.ptr, c->has_pert ? (const uint32_t *)c->b_keep_lbl.ptr : nullptr, N,
				c->W, (uint4 *)c->b_rows_a.ptr, (uint4 *)c->b_rows_b.ptr, (int *)c->b_counts.ptr, (int *)c->b_pairs_ord.ptr));
		c->pairs_ord = true;
	}
	CK(cudaEventRecord(c->ev[EV_PACK1], st));
	CK(cudaStreamSynchronize(st));
	c->timing.pack_ms = ev_ms(c, EV_PACK0, EV_PACK1);
	c->have_samples = true;
	c->data_epoch++;
	return CNB_OK;
}

/* ------------------------------------------------------------------ sibling contexts ---- */
/* the packed data set of src (label space: codes, bit planes, keep masks, pair tables) copied device to device into dst */
static int clone_data(cnb_ctx *src, cnb_ctx *dst) {
	CK(cudaSetDevice(src->device));
	const int N = src->N;
	dst->N = N; dst->M = src->M; dst->W = src->W; dst->Mpad = src->Mpad; dst->max_cats = src->max_cats;
	dst->has_pert = src->has_pert; dst->clean = src->clean; dst->has_na = src->has_na;
	dst->pairs_valid = src->pairs_valid; dst->pairs_full = src->pairs_full; dst->pairs_ord = src->pairs_ord;
	dst->triples_valid = false;
	dst->cats_lbl = src->cats_lbl; dst->never_kept = src->never_kept; dst->any_never_kept = src->any_never_kept;
	dst->force_generic = src->force_generic; dst->no_ie = src->no_ie; dst->tensor_mode = src->tensor_mode; dst->tensor_i8 = src->tensor_i8;
	dst->dp_per_node = src->dp_per_node; dst->dp_no_wave = src->dp_no_wave; dst->pairs_popc = src->pairs_popc; dst->no_ie3 = src->no_ie3;
	dst->no_tensor3 = src->no_tensor3; dst->dp_no_trap = src->dp_no_trap; dst->generic_epilogue = src->generic_epilogue;
	dst->force_full = src->force_full;
	dst->planned = dst->dp_done = false;
	const size_t words = (size_t)src->W * N;
	struct { DevBuf *d; DevBuf *s; size_t n; bool on; } cp[] = {
		{&dst->b_vmin, &src->b_vmin, N * sizeof(int), true},
		{&dst->b_cats_lbl, &src->b_cats_lbl, N * sizeof(int), true},
		{&dst->b_codes, &src->b_codes, (size_t)N * src->Mpad, true},
		{&dst->b_planes_lbl, &src->b_planes_lbl, words * sizeof(uint4), true},
		{&dst->b_keep_lbl, &src->b_keep_lbl, words * sizeof(uint32_t), src->has_pert},
		{&dst->b_pairs, &src->b_pairs, (size_t)N * (N - 1) / 2 * 16 * sizeof(int), src->pairs_full || src->pairs_valid},
		{&dst->b_pairs_ord, &src->b_pairs_ord, (size_t)N * N * 16 * sizeof(int), src->pairs_ord},
	};
	for (auto &e : cp) {
		if (!e.on || !e.n) continue;
		RC(e.d->ensure(e.n));
		CK(cudaMemcpyAsync(e.d->ptr, e.s->ptr, e.n, cudaMemcpyDeviceToDevice, src->stream));
	}
	RC(dst->b_flag.ensure(4 * sizeof(int)));
	CK(cudaMemsetAsync(dst->b_flag.ptr, 0, 4 * sizeof(int), src->stream));
	RC(dst->b_planes.ensure(words * sizeof(uint4)));
	if (src->has_pert) RC(dst->b_keep.ensure(words * sizeof(uint32_t)));
	CK(cudaStreamSynchronize(src->stream));
	dst->have_samples = true;
	dst->cloned_epoch = src->data_epoch;
	return CNB_OK;
}

int cnb_ctx_batch(cnb_ctx *c, int want, std::vector<cnb_ctx *> *ctxs) {
	ctxs->assign(1, c);
	if (!c || !c->have_samples || want <= 1) return CNB_OK;
	/* a copy of the data per concurrent order: worth it for the small matrices SA runs on (ALARM: 3 MB), not for C5 */
	static const char *env = getenv("CNB_BATCH_MAX_BYTES");
	const size_t cap = env ? (size_t)strtoull(env, nullptr, 10) : ((size_t)64 << 20);
	const size_t bytes = (size_t)c->N * c->Mpad + (size_t)c->W * c->N * 20;
	if (bytes > cap) return CNB_OK;
	want = std::min(want, 16);
	while ((int)c->siblings.size() < want - 1) {
		cnb_ctx *s = nullptr;
		RC(cnb_ctx_create(c->device, &s));
		c->siblings.push_back(s);
	}
	for (int k = 0; k < want - 1; k++) {
		cnb_ctx *s = c->siblings[k];
		if (s->cloned_epoch != c->data_epoch || !s->have_samples) RC(clone_data(c, s));
		ctxs->push_back(s);
	}
	if (!c->crew || c->crew->G < want) {
		if (c->crew) { c->crew->stop(); delete c->crew; }
		c->crew = new Crew();
		c->crew->start(want);
	}
	return CNB_OK;
}

int cnb_ctx_batch_run(cnb_ctx *c, const std::vector<cnb_ctx *> &ctxs, const cnb_params *p, int n, const int32_t *const *orders,
		bool to_options) {
	if (n < 1) return CNB_OK;
	if (n > (int)ctxs.size() || (n > 1 && (!c->crew || c->crew->G < n))) { cnb_set_error("cnb_ctx_batch_run: %d orders, %d contexts", n, (int)ctxs.size()); return CNB_ERR_PROC; }
	auto job = [&](int t) -> int {
		cnb_params q = *p;
		q.order = orders[t];
		return to_options ? cnb_ctx_search_to_options(ctxs[t], &q) : cnb_ctx_search_light(ctxs[t], &q);
	};
	if (n == 1) return job(0);
	return c->crew->run(n, job);
}

extern "C" int cnb_ctx_set_samples_dev(cnb_ctx *c, int32_t N, int32_t M, const int32_t *samples_dev,
		const int32_t *pert_dev, const int32_t *num_cats) {
	return cnb_ctx_set_samples_impl(c, N, M, samples_dev, 4, pert_dev, num_cats);
}

extern "C" int cnb_ctx_set_samples_dev8(cnb_ctx *c, int32_t N, int32_t M, const uint8_t *samples_dev,
		const int32_t *pert_dev, const int32_t *num_cats) {
	return cnb_ctx_set_samples_impl(c, N, M, samples_dev, 1, pert_dev, num_cats);
}

/* int32 -> uint8 with saturation (x < 1, NA_INTEGER included -> 0 = missing; x > 255 -> 255): 16 values per step, SSE2.
 * Returns true when some value was above 255: the byte form cannot carry it (labels such as 300..302 are legal when the
 * categories are inferred as max - min + 1), and the caller falls back to the int32 upload. */
#include <emmintrin.h>
#include <atomic>
#include <thread>
static bool compact_rows(const int32_t *src, uint8_t *dst, size_t n) {
	size_t i = 0;
	const bool stream = ((uintptr_t)dst & 15) == 0;              /* non-temporal stores: the bytes are read next by the copy engine */
	const __m128i lim = _mm_set1_epi32(255);
	__m128i over = _mm_setzero_si128();
	for (; i + 16 <= n; i += 16) {
		const __m128i a = _mm_loadu_si128((const __m128i *)(src + i)), b = _mm_loadu_si128((const __m128i *)(src + i + 4));
		const __m128i c = _mm_loadu_si128((const __m128i *)(src + i + 8)), d = _mm_loadu_si128((const __m128i *)(src + i + 12));
		over = _mm_or_si128(over, _mm_or_si128(_mm_or_si128(_mm_cmpgt_epi32(a, lim), _mm_cmpgt_epi32(b, lim)),
				_mm_or_si128(_mm_cmpgt_epi32(c, lim), _mm_cmpgt_epi32(d, lim))));
		const __m128i v = _mm_packus_epi16(_mm_packs_epi32(a, b), _mm_packs_epi32(c, d));
		if (stream) _mm_stream_si128((__m128i *)(dst + i), v);
		else _mm_storeu_si128((__m128i *)(dst + i), v);
	}
	bool big = _mm_movemask_epi8(over) != 0;
	for (; i < n; i++) {
		big = big || src[i] > 255;
		dst[i] = (uint8_t)(src[i] < 1 ? 0 : (src[i] > 255 ? 255 : src[i]));
	}
	_mm_sfence();
	return big;
}

/* Large host matrices cross PCIe as bytes: the categories fit 8 bits, the int32 matrix R holds is 4x that.  Host threads
 * compact chunk after chunk of values [v0, v1) into the context's pinned buffer while the copy engine moves the chunks
 * already done to dst_dev + v0 (C5: 2 GB in 40 ms as int32 against 0.5 GB behind the compaction).  The source may be
 * pageable memory (R's): the threads read it, only the pinned bytes are DMA'd.  *overflow = a value above 255 was met. */
int cnb_ctx_upload_bytes(cnb_ctx *c, const int32_t *samples, size_t v0, size_t v1, uint8_t *dst_dev, int max_threads,
		bool *overflow, cudaStream_t copy_stream, size_t pin_offset, size_t pin_capacity) {
	/* copy_stream: the stream the chunks are copied on (NULL = the context's); the values are staged at pin_offset of a
	 * pinned buffer of pin_capacity values (default: a buffer for just this call): a caller that uploads slice after slice
	 * gives every slice its own part of the buffer, so that a slice is compacted while the previous one is still being copied */
	const size_t total = v1 - v0;
	*overflow = false;
	if (!total) return CNB_OK;
	if (!copy_stream) copy_stream = c->stream;
	if (!pin_capacity) { pin_offset = 0; pin_capacity = total; }
	if (pin_offset + total > pin_capacity) { cnb_set_error("upload: staging range %zu + %zu exceeds %zu", pin_offset, total, pin_capacity); return CNB_ERR_PARAM; }
	RC(c->pin_stage.ensure(pin_capacity));
	uint8_t *pin = (uint8_t *)c->pin_stage.ptr + pin_offset;
	/* values per chunk: 16 MB of int32 -> 4 MB for whole matrices; a rank's share of a small slice still gives every thread
	 * a few chunks (one 4 M chunk = one thread = 2.4 ms for the 3.9 M values of C5's first slice on 8 ranks) */
	unsigned hw = std::thread::hardware_concurrency();
	const unsigned want_threads = std::min(hw ? hw : 4u, (unsigned)std::max(max_threads, 1));
	size_t chunk = (size_t)4 << 20;
	while (chunk > ((size_t)128 << 10) && total < chunk * want_threads * 3) chunk >>= 1;
	const int nchunks = (int)((total + chunk - 1) / chunk);
	const int nthreads = (int)std::max(1u, std::min(want_threads, (unsigned)nchunks));
	/* every thread (the caller is one of them: nobody spins waiting) narrows a chunk and queues its copy itself; copies of
	 * one stream may be issued from several threads, and the chunks go to disjoint destinations, so their order is free */
	std::atomic<int> next(0), big(0), failed(0);
	const int device = c->device;
	auto work = [&](bool set_device) {
		if (set_device) cudaSetDevice(device);
		for (;;) {
			const int k = next.fetch_add(1);
			if (k >= nchunks) return;
			const size_t b = (size_t)k * chunk, e = std::min(total, b + chunk);
			if (compact_rows(samples + v0 + b, pin + b, e - b)) big.store(1, std::memory_order_relaxed);
			if (cudaMemcpyAsync(dst_dev + v0 + b, pin + b, e - b, cudaMemcpyHostToDevice, copy_stream) != cudaSuccess)
				failed.store(1, std::memory_order_relaxed);
		}
	};
	std::vector<std::thread> pool;
	for (int t = 1; t < nthreads; t++) pool.emplace_back(work, true);
	work(false);
	for (auto &th : pool) th.join();
	if (failed.load()) { cnb_set_error("CUDA: %s", cudaGetErrorString(cudaGetLastError())); return CNB_ERR_CUDA; }
	c->timing.h2d_bytes += (int64_t)total;
	*overflow = big.load() != 0;
	return CNB_OK;
}

/* Not with perturbations (a second int32 matrix: plain path).  *fallback = the byte form cannot carry this matrix. */
static int set_samples_compacted(cnb_ctx *c, int32_t N, int32_t M, const int32_t *samples, const int32_t *num_cats, bool *fallback) {
	const size_t total = (size_t)N * M;
	RC(c->b_stage_s.ensure(total));
	RC(cnb_ctx_upload_bytes(c, samples, 0, total, (uint8_t *)c->b_stage_s.ptr, 16, fallback));
	if (*fallback) { cudaStreamSynchronize(c->stream); return CNB_OK; }
	c->in_host_set = true;
	int rc = cnb_ctx_set_samples_impl(c, N, M, c->b_stage_s.ptr, 1, nullptr, num_cats);
	c->in_host_set = false;
	return rc;
}

extern "C" int cnb_ctx_set_samples(cnb_ctx *c, int32_t N, int32_t M, const int32_t *samples,
		const int32_t *pert, const int32_t *num_cats) {
	if (!c || N < 1 || M < 1 || !samples) { cnb_set_error("Invalid sample"); return CNB_ERR_PARAM; }
	CK(cudaSetDevice(c->device));
	size_t bytes = (size_t)N * M * sizeof(int32_t);
	c->timing.h2d_bytes = c->timing.d2h_bytes = 0;
	/* CNB_HOST_COMPACT_MIN: smallest matrix (bytes) that takes the compacting route; "off" disables it (tests, tuning) */
	size_t compact_min = (size_t)256 << 20;
	bool no_compact = false;
	if (const char *e = getenv("CNB_HOST_COMPACT_MIN")) {
		if (!strcmp(e, "off")) no_compact = true;
		else compact_min = (size_t)strtoull(e, nullptr, 10);
	}
	if (!pert && bytes >= compact_min && !no_compact) {
		bool fallback = false;
		int rc = set_samples_compacted(c, N, M, samples, num_cats, &fallback);
		cudaStreamSynchronize(c->stream);
		if (rc != CNB_OK || !fallback) return rc;
		c->timing.h2d_bytes = 0;                                  /* a label above 255: the int32 route carries it */
	}
	/* staging buffers stay with the context (grow-only): allocating and freeing 2 GB per call costs more than the copy */
	DevBuf &tmp_s = c->b_stage_s, &tmp_p = c->b_stage_p;
	int rc = tmp_s.ensure(bytes);
	if (rc == CNB_OK && pert) rc = tmp_p.ensure(bytes);
	if (rc == CNB_OK) {
		cudaError_t e = cudaMemcpyAsync(tmp_s.ptr, samples, bytes, cudaMemcpyHostToDevice, c->stream);
		if (e == cudaSuccess && pert) e = cudaMemcpyAsync(tmp_p.ptr, pert, bytes, cudaMemcpyHostToDevice, c->stream);
		if (e != cudaSuccess) { cnb_set_error("CUDA: %s", cudaGetErrorString(e)); rc = CNB_ERR_CUDA; }
		c->timing.h2d_bytes += (int64_t)bytes * (pert ? 2 : 1);
	}
	if (rc == CNB_OK) {
		c->in_host_set = true;
		rc = cnb_ctx_set_samples_dev(c, N, M, (const int32_t *)tmp_s.ptr, pert ? (const int32_t *)tmp_p.ptr : nullptr, num_cats);
		c->in_host_set = false;
	}
	cudaStreamSynchronize(c->stream);
	return rc;
}

/* multi-process callers: rows [row0, row1) of the host int32 matrix (pageable is fine: host threads read it) become bytes
 * at dst_dev + row0 * N on the context's stream; the ranks then all-gather the byte matrix and hand it to
 * cnb_ctx_set_samples_dev8 */
extern "C" int cnb_ctx_upload_rows8_on(cnb_ctx *c, int32_t N, const int32_t *samples, int64_t row0, int64_t row1, uint8_t *dst_dev,
		int32_t max_threads, void *copy_stream, int64_t stage_offset, int64_t stage_capacity) {
	if (!c || N < 1 || !samples || row0 < 0 || row1 < row0 || !dst_dev || stage_offset < 0 || stage_capacity < 0) {
		cnb_set_error("cnb_ctx_upload_rows8: bad arguments");
		return CNB_ERR_PARAM;
	}
	CK(cudaSetDevice(c->device));
	bool ovf = false;
	RC(cnb_ctx_upload_bytes(c, samples, (size_t)row0 * N, (size_t)row1 * N, dst_dev, max_threads > 0 ? max_threads : 16, &ovf,
			(cudaStream_t)copy_stream, (size_t)stage_offset, (size_t)stage_capacity));
	if (ovf) { cnb_set_error("a category label above 255 cannot travel as a byte"); return CNB_ERR_PARAM; }
	return CNB_OK;
}

extern "C" int cnb_ctx_upload_rows8(cnb_ctx *c, int32_t N, const int32_t *samples, int64_t row0, int64_t row1, uint8_t *dst_dev,
		int32_t max_threads) {
	if (c) c->timing.h2d_bytes = c->timing.d2h_bytes = 0;
	return cnb_ctx_upload_rows8_on(c, N, samples, row0, row1, dst_dev, max_threads, nullptr, 0, 0);
}

extern "C" int cnb_ctx_num_cats(const cnb_ctx *c, int32_t *out) {
	if (!c || !c->have_samples || !out) return CNB_ERR_PARAM;
	for (int i = 0; i < c->N; i++) out[i] = c->cats_lbl[i];
	return CNB_OK;
}

extern "C" int cnb_ctx_set_stream(cnb_ctx *c, void *cuda_stream) {
	if (!c) return CNB_ERR_PARAM;
	cudaStreamSynchronize(c->stream);
	c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
	return CNB_OK;
}

extern "C" int cnb_ctx_allow_scatter(cnb_ctx *c, int32_t mode) {
	if (!c || mode < 0 || mode > 2) return CNB_ERR_PARAM;
	c->allow_scatter = mode;
	return CNB_OK;
}

extern "C" int32_t cnb_ctx_scattered(const cnb_ctx *c) { return c && c->scattered ? 1 : 0; }

extern "C" int cnb_ctx_force_generic(cnb_ctx *c, int32_t on) {
	if (!c) return CNB_ERR_PARAM;
	c->force_generic = (on & 1) != 0;     /* bit 0: generic histogram kernel everywhere */
	c->no_ie = (on & 2) != 0;             /* bit 1: keep the direct bit-plane scorer (no inclusion-exclusion) */
	c->tensor_mode = (on & 4) ? 1 : ((on & 8) ? -1 : 0);   /* bit 2: tensor-core path whenever applicable; bit 3: never */
	c->tensor_i8 = (on & 16) ? 1 : 0;                      /* bit 4: u8 operands (kind::i8) instead of E2M1 (kind::mxf4) */
	c->dp_per_node = (on & 32) != 0;                       /* bit 5: one DP launch per node (no cooperative kernel) */
	c->dp_no_wave = (on & 64) != 0;                        /* bit 6: grid-barrier DP kernel instead of the wavefront one */
	c->pairs_popc = (on & 128) != 0;                       /* bit 7: pair tables by the POPC kernel (set BEFORE set_samples) */
	c->no_ie3 = (on & 512) != 0;                           /* bit 9: no inclusion-exclusion for three-parent families */
	c->no_tensor3 = (on & 1024) != 0;                      /* bit 10: three-parent families never on the tensor cores */
	c->dp_no_trap = (on & 2048) != 0;                      /* bit 11: wavefront DP one node per exchange (k_dp_wave) */
	c->generic_epilogue = (on & 4096) != 0;                /* bit 12: generic epilogue of the tensor-core scorer also for 4-category data */
	c->force_full = (on & 8192) != 0;
	c->hist_mode = (on & 32768) ? 0 : ((on & 65536) ? 2 : c->hist_mode_default);   /* bit 15: generic scorer one CTA per family; bit 16: a warp per family with __match_any_sync */
	c->no_reduced_tiles = (on & 131072) != 0;               /* bit 17: 28 x 64 tiles (3 free categories) whatever the data's categories */
	c->dp_chains = (on & 16384) != 0;                      /* bit 14: trapezoid DP, every slot re-sums its own candidates (no work list) */                      /* bit 13: full-table tiles also on complete data (no inclusion-exclusion) */
	if (c->no_ie) c->pairs_valid = false;
	return CNB_OK;
}

extern "C" int cnb_ctx_timing(const cnb_ctx *c, cnb_timing *out) {
	if (!c || !out) return CNB_ERR_PARAM;
	*out = c->timing;
	return CNB_OK;
}

static int search_order_streamed(cnb_ctx *c, const cnb_params *p, cnb_result **out, bool *done);
static thread_local cnb_timing g_last_timing = {};
void cnb_set_last_call_timing(const cnb_timing &t) { g_last_timing = t; }
extern "C" int cnb_last_call_timing(cnb_timing *out) {
	if (!out) return CNB_ERR_PARAM;
	*out = g_last_timing;
	return CNB_OK;
}

/* ------------------------------------------------------------------ plan upload ---- */
static DevData make_devdata(cnb_ctx *c) {
	DevData D;
	D.N = c->N; D.M = c->M; D.W = c->W; D.Mpad = c->Mpad;
	D.planes = (const uint4 *)c->b_planes.ptr;
	const bool masks = c->has_pert && !c->ignore_pert;
	D.keep = masks ? (const uint32_t *)c->b_keep.ptr : nullptr;
	D.codes = (const uint8_t *)c->b_codes.ptr;
	D.keep_lbl = masks ? (const uint32_t *)c->b_keep_lbl.ptr : nullptr;
	D.order = (const int32_t *)c->b_order.ptr;
	D.cats = (const int32_t *)c->b_cats.ptr;
	D.lists = (const int32_t *)c->b_lists.ptr;
	D.blocks = (const CnbBlock *)c->b_blocks.ptr;
	D.group_blocks = (const int32_t *)c->b_group_blocks.ptr;
	D.edge = c->plan.has_prior ? (const double *)c->b_edge.ptr : nullptr;
	D.fix1 = c->plan_has_fix1 ? (const int32_t *)c->b_fix1.ptr : nullptr;
	cnb_plan_tiles(c->plan, &D.tiles);
	D.tiles.tile_base = c->plan.tile_base.empty() ? nullptr : (const long long *)c->b_tile_base.ptr;
	D.tiles.dtile_off = c->plan.tile_base.empty() ? nullptr : (const long long *)c->b_dtile_off.ptr;
	D.tiles.rank = c->plan.rank;
	D.tiles.world = c->plan.world;
	D.tiles.row_stride = masks ? 8 : 4;
	return D;
}

template <class T>
static int upload(cnb_ctx *c, DevBuf &b, const std::vector<T> &v) {
	RC(b.ensure(std::max<size_t>(v.size(), 1) * sizeof(T)));
	if (!v.empty()) H2D(c, b.ptr, v.data(), v.size() * sizeof(T));
	return CNB_OK;
}

extern "C" int cnb_ctx_plan(cnb_ctx *c, const cnb_params *p, int32_t rank, int32_t world, int64_t *seg_len,
		int64_t *num_families) {
	if (!c || !c->have_samples || !p) { cnb_set_error("context has no samples"); return CNB_ERR_PARAM; }
	CK(cudaSetDevice(c->device));
	cnb_params q = *p;
	q.num_nodes = c->N;
	q.num_samples = c->M;
	EVREC(c, EV_PLAN0);
	c->cache_P = p->max_parent_set;
	if (p->parent_sizes) c->cache_psizes.assign(p->parent_sizes, p->parent_sizes + c->N);
	else c->cache_psizes.clear();
	/* the previous plan's host vectors may still be the source of an async copy */
	CK(cudaStreamSynchronize(c->stream));
	std::string err;
	std::vector<int32_t> eff_sizes;
	if (c->any_never_kept) {                       /* label space, like parent_sizes: no candidates for those nodes */
		eff_sizes.resize(c->N);
		for (int i = 0; i < c->N; i++)
			eff_sizes[i] = c->never_kept[i] ? 0 : (p->parent_sizes ? p->parent_sizes[i] : p->max_parent_set);
		q.parent_sizes = eff_sizes.data();
	}
	/* tensor-core tiling of the two-parent group: clean data (no NA, no masks, <= 4 categories) and enough samples
	 * per family to amortise a 128 x 256 tile (or forced on by the test hook) */
	const bool tensor_ok = !c->force_generic && c->tensor_mode >= 0 && (c->tensor_mode > 0 || c->M >= 8192)
			&& c->M < (1 << 24);          /* accumulators hold counts in f32 (mxf4) or 128 x count in s32 (i8) */
	/* 1: inclusion-exclusion tiles (27 cells contracted, the rest from the pair tables: complete data only);
	 * 2: full-table tiles (all 64 cells: missing values and perturbation masks count correctly) */
	int tile_mode = 0;
	if (tensor_ok && c->clean && c->pairs_valid && !c->force_full) tile_mode = (c->no_reduced_tiles || c->tensor_i8) ? 1 : 3;
	else if (tensor_ok && c->max_cats <= CNB_PLANE_CATS && !c->tensor_i8) tile_mode = 2;
	c->tables_resident = false;
	/* The plan is a pure function of its inputs, and without pools, fixed parents, parent-set sizes, priors and with one
	 * category count for all nodes the ORDER enters it only as the order itself (every position offers all subsets of its
	 * predecessors, whoever they are): such a plan is kept and a later call with the same shape -- the next order of an SA
	 * chain or a histogram batch on such data, the next step of a benchmark loop -- replaces just the order. */
	bool uniform = true;
	for (int i = 1; i < c->N; i++) uniform = uniform && c->cats_lbl[i] == c->cats_lbl[0];
	const bool order_free = uniform && !p->parents_pool && !p->fixed_parents && !q.parent_sizes && !p->edge_prob;
	const CnbPlanKey key = {c->N, c->M, q.max_parent_set, q.max_complexity, rank, world, tile_mode, c->cats_lbl[0], order_free ? 1 : 0};
	const bool reuse = order_free && c->plan_key_valid && !memcmp(&key, &c->plan_key, sizeof(key)) && !getenv("CNB_NO_PLAN_REUSE");
	c->plan_key_valid = false;
	CnbPlan &pl = c->plan;
	if (reuse) {
		std::vector<char> seen(c->N, 0);
		for (int i = 0; i < c->N; i++) {                       /* rcatnet_search.cpp:111-128, as in cnb_build_plan */
			const int lbl = p->order ? p->order[i] : i + 1;
			if (lbl <= 0 || lbl > c->N || seen[lbl - 1]) { cnb_set_error("Invalid nodeOrder parameter"); return CNB_ERR_PARAM; }
			seen[lbl - 1] = 1;
			pl.order[i] = lbl - 1;
		}
		RC(upload(c, c->b_order, pl.order));
	} else {
	RC(cnb_build_plan(&q, c->cats_lbl.data(), rank, world, &c->plan, &err, tile_mode));
	const bool masks = c->has_pert && !c->ignore_pert;
	/* the plan's tables travel as ONE block: staged in page-locked memory, one copy, the per-table buffers are views into it
	 * (an SA chain plans thousands of small orders, and an order costs about as much as the driver calls it makes) */
	struct PlanItem { DevBuf *b; const void *src; size_t bytes; size_t off; };
	std::vector<PlanItem> items;
	auto add = [&](DevBuf &b, const void *src, size_t bytes) { items.push_back({&b, src, bytes, 0}); };
	if (!pl.tile_base.empty()) {
		add(c->b_tile_base, pl.tile_base.data(), pl.tile_base.size() * sizeof(int64_t));
		add(c->b_dtile_off, pl.dtile_off.data(), pl.dtile_off.size() * sizeof(int64_t));
		const size_t rows_per_node = pl.tile_full ? (masks ? 8 : 4) : 3;
		size_t row_bytes = (size_t)cnbk_umma_row_quads(c->W, !c->tensor_i8, nullptr) * sizeof(uint4) * rows_per_node * c->N;
		RC(c->b_rows_a.ensure(row_bytes));
		RC(c->b_rows_b.ensure(row_bytes));
	}
	add(c->b_order, pl.order.data(), pl.order.size() * sizeof(int32_t));
	add(c->b_cats, pl.cats.data(), pl.cats.size() * sizeof(int32_t));
	c->plan_has_fix1 = false;
	for (int v : pl.fix1) c->plan_has_fix1 = c->plan_has_fix1 || v >= 0;
	if (c->plan_has_fix1) add(c->b_fix1, pl.fix1.data(), pl.fix1.size() * sizeof(int32_t));
	add(c->b_lists, pl.lists.data(), pl.lists.size() * sizeof(int32_t));
	add(c->b_blocks, pl.blocks.data(), pl.blocks.size() * sizeof(CnbBlock));
	add(c->b_group_blocks, pl.group_blocks.data(), pl.group_blocks.size() * sizeof(int32_t));
	add(c->b_groups, pl.groups.data(), pl.groups.size() * sizeof(CnbGroup));
	if (pl.has_prior) add(c->b_edge, pl.edge.data(), pl.edge.size() * sizeof(double));
	c->dp_nodes.resize(pl.nodes.size());
	for (size_t i = 0; i < pl.nodes.size(); i++) {
		c->dp_nodes[i].opt_begin = pl.nodes[i].opt_begin;
		c->dp_nodes[i].opt_end = pl.nodes[i].opt_end;
		c->dp_nodes[i].base_cx = pl.nodes[i].base_cx;
		c->dp_nodes[i].pad = 0;
	}
	add(c->b_dp_nodes, c->dp_nodes.data(), c->dp_nodes.size() * sizeof(CnbDpNode));
	c->blk_equal.assign(pl.blocks.size(), 0);
	for (size_t b = 0; b < pl.blocks.size(); b++) c->blk_equal[b] = pl.nodes[pl.blocks[b].child].equal_branch;
	add(c->b_blk_equal, c->blk_equal.data(), c->blk_equal.size() * sizeof(int32_t));
	/* selection: large candidate blocks are cut into chunks, one CTA each */
	c->sel_chunks.clear();
	c->sel_chunk0.assign(pl.blocks.size() + 1, 0);
	for (size_t b = 0; b < pl.blocks.size(); b++) {
		const int64_t nch = std::max<int64_t>(1, (pl.blocks[b].ncomb + cnbk_select_chunk() - 1) / cnbk_select_chunk());
		for (int64_t k = 0; k < nch; k++) { c->sel_chunks.push_back((int32_t)b); c->sel_chunks.push_back((int32_t)k); }
		c->sel_chunk0[b + 1] = c->sel_chunk0[b] + (int32_t)nch;
	}
	add(c->b_sel_chunks, c->sel_chunks.data(), c->sel_chunks.size() * sizeof(int32_t));
	add(c->b_sel_chunk0, c->sel_chunk0.data(), c->sel_chunk0.size() * sizeof(int32_t));
	size_t total = 0;
	for (PlanItem &it : items) {
		it.off = total;
		total += (std::max<size_t>(it.bytes, 16) + 255) / 256 * 256;
	}
	const void *old_plan = c->b_plan.ptr;
	RC(c->b_plan.ensure(total));
	RC(c->pin_plan.ensure(total));
	if (c->b_plan.ptr != old_plan)                        /* the block moved: no table of an earlier plan may keep pointing into it */
		for (DevBuf *b : {&c->b_tile_base, &c->b_dtile_off, &c->b_order, &c->b_cats, &c->b_fix1, &c->b_lists, &c->b_blocks, &c->b_group_blocks,
				&c->b_groups, &c->b_edge, &c->b_dp_nodes, &c->b_blk_equal, &c->b_sel_chunks, &c->b_sel_chunk0})
			if (b->view) b->release();
	for (const PlanItem &it : items) {
		if (it.bytes) memcpy((char *)c->pin_plan.ptr + it.off, it.src, it.bytes);
		it.b->alias((char *)c->b_plan.ptr + it.off, std::max<size_t>(it.bytes, 16));
	}
	H2D(c, c->b_plan.ptr, c->pin_plan.ptr, total);
	}
	c->plan_key = key;
	c->plan_key_valid = true;
	const bool masks = c->has_pert && !c->ignore_pert;
	c->timing.kernel_launches = 0;
	if (c->streaming) {                                /* the streamed one-shot search moves the samples in slice by slice itself */
		EVREC(c, EV_PLAN1);
		c->planned = true;
		c->dp_done = false;
		if (seg_len) *seg_len = pl.seg_len;
		if (num_families) *num_families = pl.num_families;
		return CNB_OK;
	}
	/* samples into order space (the reference re-copies the int matrix per order, rcatnet_search.cpp:151-160) */
	RC(cnbk_permute(c->stream, (const uint4 *)c->b_planes_lbl.ptr,
			c->has_pert ? (const uint32_t *)c->b_keep_lbl.ptr : nullptr, (const int *)c->b_order.ptr, c->N, c->W,
			(uint4 *)c->b_planes.ptr, c->has_pert ? (uint32_t *)c->b_keep.ptr : nullptr));
	c->timing.kernel_launches = 1;                    /* k_permute; the later phases add theirs */
	if (!pl.tile_base.empty()) {
		/* operand bit rows of the tensor-core scorer, order space */
		if (pl.tile_full)
			RC(cnbk_umma_rows_full(c->stream, (const uint4 *)c->b_planes.ptr, masks ? (const uint32_t *)c->b_keep.ptr : nullptr, c->N, c->W,
					(uint4 *)c->b_rows_a.ptr, (uint4 *)c->b_rows_b.ptr));
		else
			RC(cnbk_umma_rows(c->stream, (const uint4 *)c->b_planes.ptr, c->N, c->W, !c->tensor_i8, (uint4 *)c->b_rows_a.ptr,
					(uint4 *)c->b_rows_b.ptr));
		c->timing.kernel_launches++;
	}
	EVREC(c, EV_PLAN1);
	c->planned = true;
	c->dp_done = false;
	if (seg_len) *seg_len = pl.seg_len;
	if (num_families) *num_families = pl.num_families;
	return CNB_OK;
}

/* ------------------------------------------------------------------ scoring ---- */
static int table_stride(int max_cats, int d) {
	long long s = max_cats;
	for (int i = 0; i < d; i++) {
		s *= max_cats;
		if (s > (1 << 24)) return 1 << 24;
	}
	return (int)s;
}

/* three-way tables of the data set, built on first use (C(N,3) x 256 B); false when they would not fit comfortably */
static int ensure_triples(cnb_ctx *c, bool *ok) {
	const long long nt = (long long)c->N * (c->N - 1) * (c->N - 2) / 6;
	*ok = nt * 256 <= (8ll << 30) && c->pairs_valid;
	if (!*ok || c->triples_valid) return CNB_OK;
	RC(c->b_triples.ensure((size_t)std::max<long long>(nt, 1) * 64 * sizeof(int)));
	RC(cnbk_triple_tables(c->stream, (const uint4 *)c->b_planes_lbl.ptr, c->N, c->W, c->max_cats, (const int *)c->b_pairs.ptr,
			(int *)c->b_triples.ptr));
	c->timing.kernel_launches++;
	c->triples_valid = true;
	return CNB_OK;
}

/* can the unrestricted three-parent group g run as T tiles?  (sizes and switches only: every rank decides alike) */
static bool tensor3_eligible(const cnb_ctx *c, const CnbGroup &g) {
	const CnbPlan &pl = c->plan;
	const int N = c->N;
	if (g.d != 3 || !g.full || (pl.world != 1 && !c->allow_scatter) || pl.tile_base.empty() || pl.tile_full || !c->clean || !c->pairs_valid
			|| c->force_generic || c->no_ie || c->no_ie3 || c->no_tensor3 || c->tensor_i8 || c->tensor_mode < 0
			|| (c->tensor_mode == 0 && c->M < 8192) || c->M >= (1 << 24) || N < 4 || N > 64 * 32
			|| cnbk_umma_row_quads(c->W, 1, nullptr) > 65535) return false;
	const int f1 = (c->max_cats <= 2 ? 2 : (c->max_cats == 3 ? 3 : 4)) - 1;
	const long long nvirt = (long long)N * (N - 1) / 2 * f1, nt = (long long)N * (N - 1) * (N - 2) / 6;
	const size_t row_bytes = (size_t)cnbk_umma_row_quads(c->W, 1, nullptr) * sizeof(uint4) * 3 * (size_t)(N + nvirt);
	return row_bytes <= (16ull << 30) && nvirt <= (1ll << 28) && nt * 256 <= (24ll << 30);
}

/* the whole unrestricted three-parent group on the tensor cores (T tiles, cnb_tiles.h): per column tile of children the
 * table kernel contracts (virtual pair node (a, b, i), c) x children, the epilogue completes the four-way tables from the
 * three-way tables of the data set.  A column tile is contracted in BATCHES of tiles that fit the table buffer (so the number
 * of nodes is no longer bounded by one column tile's tables), and with several ranks every rank takes a contiguous share
 * of each column tile's tiles -- i.e. a range of triples (a, b, c) for ALL children of the column tile, so its scores land
 * in every rank's segment of the score buffer ("scattered": the exchange is then a max-reduction over a buffer that
 * starts at -inf, or peer stores into one buffer; cnb_ctx_allow_scatter).  *done = false (nothing launched) when the group
 * does not qualify. */
static int score_tensor3(cnb_ctx *c, const DevData &D, const DevFamilies &F, const CnbGroup &g, const ScoreOut &O, bool *done) {
	*done = false;
	const CnbPlan &pl = c->plan;
	const int N = c->N;
	if (!tensor3_eligible(c, g)) return CNB_OK;
	bool ok = false;
	RC(ensure_triples(c, &ok));
	if (!ok) {
		if (pl.world != 1) { cnb_set_error("three-parent tensor path: the three-way tables do not fit on this rank"); return CNB_ERR_MEM; }
		return CNB_OK;
	}
	const int cm = c->max_cats <= 2 ? 2 : (c->max_cats == 3 ? 3 : 4), f1 = cm - 1;
	const long long nvirt = (long long)N * (N - 1) / 2 * f1;
	/* tiles of f1 free categories: f1^2 rows per (virtual pair node, c) slot, f1 columns per child (cnb_tiles.h) */
	const int tp3 = cnb_geo_pairs(f1), tc3 = cnb_geo_child(f1);
	const int nct = (N + tc3 - 1) / tc3;
	std::vector<long long> base3(nct + 1, 0);
	for (int ct = 0; ct < nct; ct++) base3[ct + 1] = base3[ct] + cnb_t_tiles(cnb_t_ce(N, nct, ct, tc3), f1, tp3);
	const long long slot_bytes = (long long)cnb_geo_slot_ints(f1) * sizeof(int) * tp3 * tc3;
	const size_t row_bytes = (size_t)cnbk_umma_row_quads(c->W, 1, nullptr) * sizeof(uint4) * 3 * (size_t)(N + nvirt);
	/* a batch / a rank's share starts on a slot that begins a triple: tile boundaries where tile * tp3 is a multiple of f1 */
	int align = 1;
	while ((long long)align * tp3 % f1) align++;
	long long budget = 4ll << 30;                            /* table buffer of a batch (CNB_T3_BATCH_BYTES: tests) */
	if (const char *e = getenv("CNB_T3_BATCH_BYTES")) budget = std::max<long long>(strtoll(e, nullptr, 10), slot_bytes);
	long long batch = std::max<long long>(budget / slot_bytes / align * align, align);
	cudaStream_t st = c->stream;
	RC(upload(c, c->b_tile_base3, base3));
	RC(c->b_rows_a3.ensure(row_bytes));
	RC(cnbk_umma_rows3(st, (const uint4 *)c->b_planes.ptr, N, c->W, (int)nvirt, f1, (uint4 *)c->b_rows_a3.ptr));
	c->timing.kernel_launches += 2;
	c->tables_resident = false;                        /* the T tiles take the table buffer */
	long long widest = 1;
	for (int ct = 0; ct < nct; ct++) widest = std::max(widest, std::min(batch, base3[ct + 1] - base3[ct]));
	RC(c->b_counts.ensure((size_t)widest * slot_bytes));
	CnbTiles T3;
	memset(&T3, 0, sizeof(T3));
	T3.tp = tp3;
	T3.tc = tc3;
	T3.fr = f1;
	T3.tile_base = (const long long *)c->b_tile_base3.ptr;
	T3.nct = nct;
	T3.child = tc3;
	T3.world = 1;                                      /* contiguous tile ranges: the dealing is done here, not in the kernel */
	T3.self_pairs = 2;
	T3.f1 = f1;
	T3.nvirt = (int)nvirt;
	long long mine = 0;
	for (int ct = 0; ct < nct; ct++) {
		const int c0 = cnb_t_c0(N, nct, ct, tc3), ce = cnb_t_ce(N, nct, ct, tc3);
		const long long ntile = base3[ct + 1] - base3[ct], ntrip = cnb_binom3(ce - 1);
		if (ntile < 1 || ce <= 3) continue;
		const long long per = ((ntile + pl.world - 1) / pl.world + align - 1) / align * align;
		const long long lo = std::min(ntile, (long long)pl.rank * per), hi = std::min(ntile, lo + per);
		for (long long tb = lo; tb < hi; tb += batch) {
			const long long te = std::min(hi, tb + batch);
			const long long t_begin = tb * tp3 / f1, t_end = te == ntile ? ntrip : std::min(ntrip, te * tp3 / f1);
			RC(cnbk_umma_tables(st, (const uint4 *)c->b_rows_a3.ptr, (const uint4 *)c->b_rows_b.ptr, N, c->W, 1, T3, base3[ct] + tb,
					base3[ct] + te, (int *)c->b_counts.ptr));
			RC(cnbk_umma_epilogue3_range(st, D, F, c->max_cats, t_begin, t_end, tb * tp3, c0, ce, (const int *)c->b_counts.ptr,
					(const int *)c->b_triples.ptr, O));
			c->timing.kernel_launches += 2;
			mine += te - tb;
		}
	}
	c->tensor3_tiles = (int)mine;
	c->timing.tensor3_tiles = mine;
	*done = true;
	return CNB_OK;
}

/* the 0- and 1-parent families can be read off two-way tables of the data set (unordered ones of complete data, ordered
 * ones under the child's keep mask otherwise) */
static bool have_pair_tables(const cnb_ctx *c) {
	if (c->ignore_pert) return c->pairs_full;
	return c->pairs_valid || (c->pairs_ord && !c->no_ie);
}

/* score families [b, e) of one launch class, choosing the kernel */
static int score_range(cnb_ctx *c, const DevData &D, const DevFamilies &F, int d_max, long long b, long long e,
		ScoreOut O, bool force_generic, bool allow_slicing) {
	if (e <= b) return CNB_OK;
	cudaStream_t st = c->stream;
	/* two-way tables that hold the 0- and 1-parent families outright: the unordered ones of complete data, or the ordered
	 * ones counted under the child's keep mask / without missing values */
	const bool pairs_unord = c->ignore_pert ? c->pairs_full : c->pairs_valid;
	const bool ordered = !pairs_unord && c->pairs_ord && !c->ignore_pert && !c->no_ie;
	const bool pairs_ok = pairs_unord || ordered;
	const int *pair_tab = ordered ? (const int *)c->b_pairs_ord.ptr : (const int *)c->b_pairs.ptr;
	if (!force_generic && !F.ex_nd && F.d == 0 && pairs_ok) {
		/* no parents: marginal counts from the pair tables */
		c->last_fast = true;
		RC(cnbk_score_pairs1(c->stream, D, F, c->max_cats, b, e, O, pair_tab, ordered));
		c->timing.kernel_launches++;
		return CNB_OK;
	}
	if (!force_generic && !F.ex_nd && F.d == 3 && c->pairs_valid && !c->ignore_pert && !O.counts && !c->no_ie && !c->no_ie3
			&& c->max_cats <= CNB_PLANE_CATS) {
		/* three parents, clean data: only the free cells are counted; the rest from the three-way tables of the data set */
		bool ok = false;
		RC(ensure_triples(c, &ok));
		if (ok) {
			c->last_fast = true;
			RC(cnbk_score_planes_ie3(st, D, F, c->max_cats, b, e, O, (const int *)c->b_triples.ptr));
			c->timing.kernel_launches++;
			return CNB_OK;
		}
	}
	bool fast = !force_generic && !F.ex_nd && cnbk_planes_supported(F.d, c->max_cats);
	c->last_fast = fast;
	if (fast) {
		int slices = 1;
		if (allow_slicing) {
			/* few families x many samples: slice the sample words so that the machine is filled */
			long long n = e - b, target = 148ll * 1024;
			if (n < target) slices = (int)std::min<long long>(std::min<long long>((target + n - 1) / n, std::max(1, c->W / 8)), 65535);
		}
		if (F.d == 1 && pairs_ok && !F.ex_nd) {
			/* one parent: the table was counted with the pair tables */
			RC(cnbk_score_pairs1(st, D, F, c->max_cats, b, e, O, pair_tab, ordered));
			c->timing.kernel_launches++;
			return CNB_OK;
		}
		if (slices == 1 && !O.counts && F.d == 2 && pairs_unord && !F.explicit_list) {
			RC(cnbk_score_planes_ie2(st, D, F, c->max_cats, b, e, O, (const int *)c->b_pairs.ptr));
			c->timing.kernel_launches++;
			return CNB_OK;
		}
		if (slices > 1 && !O.counts) {
			int stride = table_stride(c->max_cats, F.d);
			size_t bytes = (size_t)(e - b) * stride * sizeof(int);
			c->tables_resident = false;
			RC(c->b_counts.ensure(bytes));
			CK(cudaMemsetAsync(c->b_counts.ptr, 0, bytes, st));
			O.counts = (int *)c->b_counts.ptr - (size_t)b * stride;
			O.counts_stride = stride;
		}
		RC(cnbk_score_planes(st, D, F, c->max_cats, b, e, slices, O));
		c->timing.kernel_launches++;
		if (slices > 1) {
			RC(cnbk_epilogue_counts(st, D, F, b, e, O));
			c->timing.kernel_launches++;
		}
		return CNB_OK;
	}
	int cells = table_stride(c->max_cats, d_max);
	if (cells > CNB_HIST_MAX_CELLS) cells = CNB_HIST_MAX_CELLS;
	RC(cnbk_score_hist(st, D, F, b, e, cells, O, (int *)c->b_flag.ptr + 1, c->hist_mode | (c->has_na ? 0 : 4)));
	c->timing.kernel_launches++;
	return CNB_OK;
}

extern "C" int cnb_ctx_score_shard(cnb_ctx *c, double *scores_dev) {
	if (!c || !c->planned || !scores_dev) { cnb_set_error("cnb_ctx_plan must run first"); return CNB_ERR_PARAM; }
	CK(cudaSetDevice(c->device));
	CnbPlan &pl = c->plan;
	DevData D = make_devdata(c);
	cudaStream_t st = c->stream;
	c->timing.num_families = 0;
	c->timing.algorithmic_bytes = 0;
	memset(c->timing.score_ms_by_parents, 0, sizeof(c->timing.score_ms_by_parents));
	memset(c->timing.families_by_parents, 0, sizeof(c->timing.families_by_parents));
	CK(cudaMemsetAsync((int *)c->b_flag.ptr + 1, 0, sizeof(int), st));
	EVREC(c, EV_SCORE0);
	c->timing.fast_path = 1;
	c->timing.tensor_tiles = 0;
	c->timing.tensor_ms = 0;
	c->timing.tensor_macs = 0;
	c->timing.tensor_kind = 0;
	c->tensor_batches = 0;
	c->tensor3_tiles = 0;
	c->timing.tensor3_tiles = 0;
	/* a sharded three-parent group on T tiles writes scores into every rank's segment: with the max-reduction exchange the
	 * buffer starts at -inf (cnb_ctx_allow_scatter) */
	c->scattered = false;
	if (pl.world > 1 && c->allow_scatter)
		for (const CnbGroup &g : pl.groups) c->scattered = c->scattered || (!g.tiled && tensor3_eligible(c, g));
	if (c->scattered && c->allow_scatter == 2)
		RC(cnbk_fill_f64(st, scores_dev, (long long)pl.seg_len * pl.world, -INFINITY));
	int gi = 0;
	for (const CnbGroup &g : pl.groups) {
		DevFamilies F;
		memset(&F, 0, sizeof(F));
		F.d = g.d; F.nblk = g.nblk; F.blk_off = g.blk_off; F.nfam = g.nfam;
		long long b = (long long)pl.rank * g.chunk, e = std::min<long long>(g.nfam, b + g.chunk);
		ScoreOut O;
		memset(&O, 0, sizeof(O));
		O.scores = scores_dev; O.chunk = g.chunk; O.seg_off = g.seg_off; O.seg_len = pl.seg_len;
		if (gi < CNB_NEV_GROUPS) EVREC(c, EV_GROUP0 + 2 * gi);
		if (g.tiled) {
			/* tcgen05 path: integer tables by one-hot GEMM, tile batches bounded by the table buffer */
			/* tiles are dealt round-robin (tile t -> rank t % world): every rank gets the same mix of full and
			 * diagonal tiles; t0..t1 are LOCAL tile indices */
			long long t0 = 0, t1 = (g.ntiles - pl.rank + pl.world - 1) / pl.world;
			bool uni4 = true;                                   /* every node with exactly 4 categories: specialised epilogue */
			for (int v : pl.cats) uni4 = uni4 && v == 4;
			const long long batch = 16384, slot_bytes = (pl.tile_full ? CNBK_UMMA_FULL_SLOT_INTS : cnb_geo_slot_ints(pl.tile_fr)) * sizeof(int)
					* (long long)pl.tile_pairs * pl.tile_cols;
			RC(c->b_counts.ensure((size_t)std::min<long long>(batch, std::max<long long>(t1 - t0, 1)) * slot_bytes));
			int nb = 0;
			for (long long tb = t0; tb < t1; tb += batch, nb++) {
				long long te = std::min(t1, tb + batch);
				if (nb < CNB_NEV_TEN) EVREC(c, EV_TEN0 + 2 * nb);
				if (!c->streaming)                            /* streamed search: the tables were contracted slice by slice */
					RC(cnbk_umma_tables(st, (const uint4 *)c->b_rows_a.ptr, (const uint4 *)c->b_rows_b.ptr, c->N, c->W,
							!c->tensor_i8, D.tiles, tb, te, (int *)c->b_counts.ptr));
				if (nb < CNB_NEV_TEN) EVREC(c, EV_TEN0 + 2 * nb + 1);
				RC(cnbk_umma_epilogue(st, D, tb, te, (const int *)c->b_counts.ptr,
						(const int *)c->b_pairs.ptr, O, uni4 && !c->generic_epilogue));
				c->timing.kernel_launches += 2;
			}
			c->timing.tensor_tiles += t1 - t0;
			c->tensor_batches = nb;
			c->tables_resident = nb == 1 && pl.world == 1;
			long long cols = 0;
			for (long long t = t0; t < t1; t++) cols += cnb_tile_columns(pl, pl.rank + t * pl.world);
			c->timing.tensor_macs = cols * cnbk_umma_column_macs(c->W, !c->tensor_i8);
			c->timing.tensor_kind = c->tensor_i8 ? 8 : 4;
			b = 0;
			e = t1 > t0 ? (long long)((double)g.nfam * (double)(t1 - t0) / (double)g.ntiles) : 0;
		} else {
			bool done = false;
			RC(score_tensor3(c, D, F, g, O, &done));
			if (!done) {
				RC(score_range(c, D, F, g.d, b, e, O, c->force_generic, true));
				if (e > b && !c->last_fast) c->timing.fast_path = 0;
			}
		}
		if (gi < CNB_NEV_GROUPS) EVREC(c, EV_GROUP0 + 2 * gi + 1);
		if (e > b) {
			c->timing.num_families += e - b;
			c->timing.algorithmic_bytes += (e - b) * ((long long)(g.d + 1) * c->M + 8);
			if (g.d < 8) c->timing.families_by_parents[g.d] += e - b;
		}
		gi++;
	}
	EVREC(c, EV_SCORE1);
	if (c->quiet) {
		/* an evaluation of the SA / histogram drivers: nobody reads the phase times, and the overflow flag of the generic
		 * scorer travels with the DP summary (cnb_ctx_dp_launch / _wait) -- the selection and the DP are queued behind the
		 * scorers without the host waiting in between */
		c->defer_ovf = true;
		return CNB_OK;
	}
	int ovf = 0;
	D2H(c, &ovf, (int *)c->b_flag.ptr + 1, sizeof(int));
	CK(cudaStreamSynchronize(st));
	if (ovf) { cnb_set_error("conditional table exceeds %d cells", CNB_HIST_MAX_CELLS); return CNB_ERR_MEM; }
	c->timing.plan_ms = ev_ms(c, EV_PLAN0, EV_PLAN1);
	c->timing.score_ms = ev_ms(c, EV_SCORE0, EV_SCORE1);
	for (int i = 0; i < c->tensor_batches && i < CNB_NEV_TEN; i++) c->timing.tensor_ms += ev_ms(c, EV_TEN0 + 2 * i, EV_TEN0 + 2 * i + 1);
	if (c->tensor_batches > CNB_NEV_TEN) c->timing.tensor_ms *= (float)c->tensor_batches / CNB_NEV_TEN;
	gi = 0;
	for (const CnbGroup &g : pl.groups) {
		if (gi < CNB_NEV_GROUPS && g.d < 8)
			c->timing.score_ms_by_parents[g.d] += ev_ms(c, EV_GROUP0 + 2 * gi, EV_GROUP0 + 2 * gi + 1);
		gi++;
	}
	return CNB_OK;
}

/* explicit family list (positions) -> per-family outputs in device buffers of the context */
static int score_explicit(cnb_ctx *c, const std::vector<int32_t> &child, const std::vector<int32_t> &pars,
		const std::vector<int32_t> &nd, int fixed_d, int stride, bool want_counts, bool force_generic) {
	size_t n = child.size();
	RC(upload(c, c->b_ex_child, child));
	RC(upload(c, c->b_ex_pars, pars));
	bool var_d = fixed_d < 0;
	if (var_d) RC(upload(c, c->b_ex_nd, nd));
	RC(c->b_ex_ll.ensure(n * sizeof(double)));
	RC(c->b_ex_prior.ensure(n * sizeof(double)));
	RC(c->b_ex_score.ensure(n * sizeof(double)));
	RC(c->b_ex_ncount.ensure(n * sizeof(int)));
	ScoreOut O;
	memset(&O, 0, sizeof(O));
	O.scores = (double *)c->b_ex_score.ptr; O.chunk = std::max<long long>(n, 1); O.seg_off = 0; O.seg_len = n;
	O.loglik = (double *)c->b_ex_ll.ptr;
	O.prior = (double *)c->b_ex_prior.ptr;
	O.ncount = (int *)c->b_ex_ncount.ptr;
	if (want_counts) {
		RC(c->b_ex_counts.ensure(n * stride * sizeof(int)));
		CK(cudaMemsetAsync(c->b_ex_counts.ptr, 0, n * stride * sizeof(int), c->stream));
		O.counts = (int *)c->b_ex_counts.ptr;
		O.counts_stride = stride;
	}
	DevFamilies F;
	memset(&F, 0, sizeof(F));
	F.explicit_list = 1;
	F.d = var_d ? 0 : fixed_d;
	F.nfam = n;
	F.ex_child = (const int32_t *)c->b_ex_child.ptr;
	F.ex_pars = (const int32_t *)c->b_ex_pars.ptr;
	F.ex_nd = var_d ? (const int32_t *)c->b_ex_nd.ptr : nullptr;
	int dmax = fixed_d;
	if (var_d) { dmax = 0; for (int v : nd) dmax = std::max(dmax, v); }
	DevData D = make_devdata(c);
	RC(score_range(c, D, F, dmax, 0, (long long)n, O, force_generic || var_d || (fixed_d == 0 && !have_pair_tables(c)), false));
	return CNB_OK;
}

/* ------------------------------------------------------------------ selection + DP ---- */
/* explicit families sorted by parent count: one launch class per count (fixed d, so the bit-plane / pair-table
 * kernels apply), tables + log-lik + prior + sample size into the b_ex_* buffers */
static int score_explicit_groups(cnb_ctx *c, const std::vector<int32_t> &child, const std::vector<int32_t> &pars,
		const std::vector<int32_t> &nd, int stride, bool from_search = false) {
	size_t n = child.size();
	RC(upload(c, c->b_ex_child, child));
	RC(upload(c, c->b_ex_pars, pars));
	RC(c->b_ex_ll.ensure(n * sizeof(double)));
	RC(c->b_ex_prior.ensure(n * sizeof(double)));
	RC(c->b_ex_score.ensure(n * sizeof(double)));
	RC(c->b_ex_ncount.ensure(n * sizeof(int)));
	RC(c->b_ex_counts.ensure(n * stride * sizeof(int)));
	CK(cudaMemsetAsync(c->b_ex_counts.ptr, 0, n * stride * sizeof(int), c->stream));
	ScoreOut O;
	memset(&O, 0, sizeof(O));
	O.scores = (double *)c->b_ex_score.ptr; O.chunk = std::max<long long>(n, 1); O.seg_off = 0; O.seg_len = n;
	O.loglik = (double *)c->b_ex_ll.ptr;
	O.prior = (double *)c->b_ex_prior.ptr;
	O.ncount = (int *)c->b_ex_ncount.ptr;
	O.counts = (int *)c->b_ex_counts.ptr;
	O.counts_stride = stride;
	DevData D = make_devdata(c);
	for (size_t b = 0; b < n;) {
		size_t e = b;
		while (e < n && nd[e] == nd[b]) e++;
		DevFamilies F;
		memset(&F, 0, sizeof(F));
		F.explicit_list = 1;
		F.d = nd[b];
		F.nfam = n;
		F.ex_child = (const int32_t *)c->b_ex_child.ptr;
		F.ex_pars = (const int32_t *)c->b_ex_pars.ptr;
		if (nd[b] == 2 && c->tables_resident && from_search && !c->force_generic) {
			/* the two-parent tables of the search are still on the device: read the families' tables from their tile slots */
			RC(cnbk_umma_export(c->stream, D, (long long)b, (long long)e, F.ex_child, F.ex_pars, (const int *)c->b_counts.ptr,
					(const int *)c->b_pairs.ptr, O));
			c->timing.kernel_launches++;
		} else
			RC(score_range(c, D, F, nd[b], (long long)b, (long long)e, O, c->force_generic || (nd[b] == 0 && !have_pair_tables(c)), true));
		b = e;
	}
	return CNB_OK;
}

/* The reference's score cache (cnb_cache.h): the option table of the planned order is in place; walk the order through the
 * model of the cache (the winners of the equal-categories blocks come back from the device first), re-score every
 * candidate the cache answers differently -- parents in the cached listing sequence -- and patch its term, and on the
 * equal-categories path the winner itself, into the option table.  The DP then runs on what the reference's DP saw. */
static int apply_cache(cnb_ctx *c) {
	CnbPlan &pl = c->plan;
	cudaStream_t st = c->stream;
	if (pl.world != 1) { cnb_set_error("the score cache model needs the whole search on one rank"); return CNB_ERR_PARAM; }
	const size_t nopt = (size_t)pl.num_options;
	c->cache_ov.clear();
	std::vector<int64_t> win(std::max<size_t>(nopt, 1), -1);
	bool any_equal = false;
	for (size_t b = 0; b < pl.blocks.size(); b++) any_equal = any_equal || c->blk_equal[b];
	if (any_equal && nopt) {
		if (nopt <= ((size_t)1 << 17)) D2H(c, win.data(), c->b_opt_fid.ptr, nopt * sizeof(long long));
		else
			for (size_t b = 0; b < pl.blocks.size(); b++)
				if (c->blk_equal[b]) D2H(c, &win[pl.blocks[b].opt_off], (const long long *)c->b_opt_fid.ptr + pl.blocks[b].opt_off, sizeof(long long));
		CK(cudaStreamSynchronize(st));
	}
	cnb_params cp;
	memset(&cp, 0, sizeof(cp));
	cp.max_parent_set = c->cache_P;
	cp.parent_sizes = c->cache_psizes.empty() ? nullptr : c->cache_psizes.data();
	RC(c->cache->walk(pl, &cp, c->any_never_kept ? c->never_kept.data() : nullptr, win.data(), &c->cache_ov));
	const size_t n = c->cache_ov.size();
	if (!n) return CNB_OK;
	/* grouped by parent count, so that each group runs on its fast kernel (as the export does) */
	std::vector<int32_t> perm(n);
	for (size_t i = 0; i < n; i++) perm[i] = (int32_t)i;
	std::stable_sort(perm.begin(), perm.end(), [&](int32_t u, int32_t v) { return c->cache_ov[u].d < c->cache_ov[v].d; });
	std::vector<int32_t> s_child(n), s_pars(n * CNB_MAXP, 0), s_nd(n);
	c->cache_h_opt.resize(n);
	c->cache_h_fid.resize(n);
	c->cache_h_cx.resize(n);
	c->cache_h_prior.resize(n);
	int dmax = 0;
	for (size_t k = 0; k < n; k++) {
		const CnbOverride &ov = c->cache_ov[perm[k]];
		s_child[k] = ov.child;
		s_nd[k] = ov.d;
		dmax = std::max(dmax, ov.d);
		for (int q = 0; q < ov.d; q++) s_pars[k * CNB_MAXP + q] = ov.pars[q];
		c->cache_h_opt[k] = ov.opt;
		c->cache_h_fid[k] = ov.fid;
		c->cache_h_cx[k] = ov.cx;
		c->cache_h_prior[k] = ov.prior;
	}
	RC(score_explicit_groups(c, s_child, s_pars, s_nd, table_stride(c->max_cats, dmax), true));
	RC(upload(c, c->b_ov_opt, c->cache_h_opt));
	RC(upload(c, c->b_ov_fid, c->cache_h_fid));
	RC(upload(c, c->b_ov_cx, c->cache_h_cx));
	if (pl.has_prior) RC(upload(c, c->b_ov_prior, c->cache_h_prior));
	RC(cnbk_patch_options(st, (int)n, (const long long *)c->b_ov_opt.ptr, (const long long *)c->b_ov_fid.ptr, (const int *)c->b_ov_cx.ptr,
			pl.has_prior ? (const double *)c->b_ov_prior.ptr : nullptr, (const double *)c->b_ex_ll.ptr, (const double *)c->b_ex_score.ptr,
			(double *)c->b_opt_score.ptr, (int *)c->b_opt_cx.ptr, (long long *)c->b_opt_fid.ptr));
	c->timing.kernel_launches++;
	c->cache_ovf = 1;                                                   /* the flag travels with the DP summary (cnb_ctx_dp_launch) */
	return CNB_OK;
}

/* selection + DP in three steps, so that the SA driver can run the first for several proposals side by side, then hand them
 * to the score-cache model one after the other while their DPs run behind one another's host work (cnb_sa.cpp):
 *   cnb_ctx_select      base network, option table (k_select)            -- enqueued, no host wait
 *   cnb_ctx_dp_launch   [score-cache overrides,] DP kernels, summary copy -- enqueued (the cache model waits for the winners)
 *   cnb_ctx_dp_wait     stream drained, timings, overflow check
 * cnb_ctx_select_dp is the three in a row. */
int cnb_ctx_select(cnb_ctx *c, const double *scores_dev) {
	if (!c || !c->planned) { cnb_set_error("cnb_ctx_plan must run first"); return CNB_ERR_PARAM; }
	CK(cudaSetDevice(c->device));
	CnbPlan &pl = c->plan;
	cudaStream_t st = c->stream;
	const int N = pl.N;
	c->dp_done = false;
	EVREC(c, EV_SEL0);
	/* base network: every node with its fixed parents only (catnet_search2.h:359-441) */
	{
		std::vector<int32_t> child(N), pars((size_t)N * CNB_MAXP, 0), nd(N);
		for (int i = 0; i < N; i++) {
			child[i] = i;
			nd[i] = pl.nodes[i].nfix;
			for (int k = 0; k < nd[i]; k++) pars[(size_t)i * CNB_MAXP + k] = pl.lists[pl.nodes[i].fix_off + k];
		}
		bool no_fixed = true;
		for (int i = 0; i < N; i++) no_fixed = no_fixed && nd[i] == 0;
		if (no_fixed && have_pair_tables(c) && !c->force_generic) RC(score_explicit(c, child, pars, nd, 0, 0, false, false));
		else RC(score_explicit(c, child, pars, nd, -1, 0, false, true));
		RC(c->b_base_v.ensure(N * sizeof(double)));
		CK(cudaMemcpyAsync(c->b_base_v.ptr, c->b_ex_score.ptr, N * sizeof(double), cudaMemcpyDeviceToDevice, st));
	}
	DevData D = make_devdata(c);
	size_t nopt = (size_t)std::max<int64_t>(pl.num_options, 1);
	RC(c->b_opt_score.ensure(nopt * sizeof(double)));
	RC(c->b_opt_cx.ensure(nopt * sizeof(int)));
	RC(c->b_opt_fid.ensure(nopt * sizeof(long long)));
	const int nchunks = (int)(c->sel_chunks.size() / 2);
	RC(c->b_sel_part_s.ensure(std::max(nchunks, 1) * sizeof(double)));
	RC(c->b_sel_part_f.ensure(std::max(nchunks, 1) * sizeof(long long)));
	RC(cnbk_select(st, D, (int)pl.blocks.size(), (const CnbGroup *)c->b_groups.ptr, (const int *)c->b_blk_equal.ptr,
			(const int2 *)c->b_sel_chunks.ptr, nchunks, (const int *)c->b_sel_chunk0.ptr, scores_dev, pl.seg_len,
			(double *)c->b_opt_score.ptr, (int *)c->b_opt_cx.ptr, (long long *)c->b_opt_fid.ptr,
			(double *)c->b_sel_part_s.ptr, (long long *)c->b_sel_part_f.ptr));
	c->timing.kernel_launches += 2;
	return CNB_OK;
}

int cnb_ctx_dp_launch(cnb_ctx *c) {
	if (!c || !c->planned) { cnb_set_error("cnb_ctx_plan must run first"); return CNB_ERR_PARAM; }
	CK(cudaSetDevice(c->device));
	CnbPlan &pl = c->plan;
	cudaStream_t st = c->stream;
	const int N = pl.N, K = pl.K;
	c->cache_ovf = 0;
	if (c->cache) RC(apply_cache(c));
	else c->cache_ov.clear();
	EVREC(c, EV_SEL1);
	/* DP state, double buffered */
	size_t slots = (size_t)K + 1;
	RC(c->b_dp_exists.ensure(3 * slots));
	RC(c->b_dp_L.ensure(3 * slots * sizeof(double)));
	RC(c->b_dp_pfx.ensure(3 * slots * sizeof(double)));
	RC(c->b_bp_opt.ensure((size_t)N * slots * sizeof(long long)));
	RC(c->b_bp_src.ensure((size_t)N * slots * sizeof(int)));
	DpState S[3];
	for (int i = 0; i < 3; i++) {
		S[i].K = K;
		S[i].exists = (unsigned char *)c->b_dp_exists.ptr + i * slots;
		S[i].L = (double *)c->b_dp_L.ptr + i * slots;
		S[i].pfx = (double *)c->b_dp_pfx.ptr + i * slots;
	}
	const double *base_v = (const double *)c->b_base_v.ptr;
	int cur = 0;
	/* All nodes in one cooperative launch: as a wavefront over CTAs of 256 slots when no option reaches further left
	 * than that (an option moves a network by its extra complexity, at most (C - 1) * maxC^d slots), else with a grid
	 * barrier per node; one launch per node if cooperative launches are refused. */
	int halo = 0;
	for (const CnbBlock &bk : pl.blocks) {
		long long reach = pl.cats[bk.child] - 1;
		for (int k = 0; k < bk.d && reach <= 4096; k++) reach *= c->max_cats;
		halo = (int)std::max<long long>(halo, std::min<long long>(reach, 1 << 20));
	}
	const int flag_cap = 4096;
	RC(c->b_dp_flags.ensure(flag_cap * sizeof(unsigned)));
	int fused = CNB_ERR_PROC;
	long long max_opts = 0;
	for (const CnbNodePlan &nd : pl.nodes) max_opts = std::max<long long>(max_opts, nd.opt_end - nd.opt_begin);
	if (!c->dp_per_node && !c->dp_no_wave && max_opts > 4) {
		/* some node hands every candidate to the DP (unequal categories): a warp per slot */
		fused = cnbk_dp_mixed(st, S[0], S[1], N, pl.base_complexity, (const CnbDpNode *)c->b_dp_nodes.ptr, base_v,
				(const double *)c->b_opt_score.ptr, (const int *)c->b_opt_cx.ptr, (const long long *)c->b_opt_fid.ptr,
				(long long *)c->b_bp_opt.ptr, (int *)c->b_bp_src.ptr);
		if (fused == CNB_OK) {
			cur = (N - 1) & 1;
			c->timing.kernel_launches += 1;
		}
	} else if (!c->dp_per_node && !c->dp_no_wave) {
		/* few options per node: the trapezoid wavefront (T nodes per neighbour exchange), else one node per exchange */
		int steps = 0;
		if (!c->dp_no_trap)
			fused = cnbk_dp_trap(st, S[0], S[1], S[2], N, pl.base_complexity, halo, (const CnbDpNode *)c->b_dp_nodes.ptr, base_v,
					(const double *)c->b_opt_score.ptr, (const int *)c->b_opt_cx.ptr, (const long long *)c->b_opt_fid.ptr,
					(long long *)c->b_bp_opt.ptr, (int *)c->b_bp_src.ptr, (unsigned *)c->b_dp_flags.ptr, flag_cap, &steps,
					c->dp_chains ? 0 : 1);
		if (fused == CNB_OK) {
			cur = steps % 3;                  /* macro-step m reads S[(m - 1) % 3] and writes S[m % 3] */
			c->timing.kernel_launches += 1;
		} else if (fused == CNB_ERR_PROC) {
			fused = cnbk_dp_wave(st, S[0], S[1], S[2], N, pl.base_complexity, halo, (const CnbDpNode *)c->b_dp_nodes.ptr, base_v,
					(const double *)c->b_opt_score.ptr, (const int *)c->b_opt_cx.ptr, (const long long *)c->b_opt_fid.ptr,
					(long long *)c->b_bp_opt.ptr, (int *)c->b_bp_src.ptr, (unsigned *)c->b_dp_flags.ptr, flag_cap);
			if (fused == CNB_OK) {
				cur = (N - 1) % 3;            /* step c reads S[(c - 1) % 3] and writes S[c % 3] */
				c->timing.kernel_launches += 1;
			}
		}
	}
	if (fused == CNB_ERR_PROC && !c->dp_per_node) {
		fused = cnbk_dp_all(st, S[0], S[1], N, pl.base_complexity, (const CnbDpNode *)c->b_dp_nodes.ptr, base_v,
				(const double *)c->b_opt_score.ptr, (const int *)c->b_opt_cx.ptr, (const long long *)c->b_opt_fid.ptr,
				(long long *)c->b_bp_opt.ptr, (int *)c->b_bp_src.ptr);
		if (fused == CNB_OK) {
			cur = (N - 1) & 1;                /* node c reads S[(c & 1) ^ 1] and writes S[c & 1] */
			c->timing.kernel_launches += 1;
		}
	}
	if (fused == CNB_OK) {
	} else if (fused == CNB_ERR_PROC) {
		RC(cnbk_dp_init(st, S[0], pl.base_complexity, base_v, N));
		for (int n = 1; n < N; n++) {
			const CnbNodePlan &nd = pl.nodes[n];
			RC(cnbk_dp_node(st, S[cur], S[cur ^ 1], n, N, nd.base_cx, base_v, nd.opt_begin, nd.opt_end,
					(const double *)c->b_opt_score.ptr, (const int *)c->b_opt_cx.ptr, (const long long *)c->b_opt_fid.ptr,
					(long long *)c->b_bp_opt.ptr, (int *)c->b_bp_src.ptr));
			cur ^= 1;
		}
		c->timing.kernel_launches += N;
	} else
		return fused;
	c->dp_final = cur;
	EVREC(c, EV_DP1);
	/* summary to the host: which slots exist and their log-lik */
	c->h_exists.resize(slots);
	c->h_L.resize(slots);
	const size_t off_e = 16, off_l = 16 + (slots + 7) / 8 * 8;
	RC(c->pin_sum.ensure(off_l + slots * sizeof(double)));
	char *pin = (char *)c->pin_sum.ptr;
	*(int *)pin = 0;
	if (c->cache_ovf || c->defer_ovf) D2H(c, pin, (int *)c->b_flag.ptr + 1, sizeof(int));   /* overflow flag of the scorers (deferred) and of the overrides' */
	c->defer_ovf = false;
	D2H(c, pin + off_e, S[cur].exists, slots);
	D2H(c, pin + off_l, S[cur].L, slots * sizeof(double));
	return CNB_OK;
}

int cnb_ctx_dp_wait(cnb_ctx *c) {
	if (!c || !c->planned) return CNB_ERR_PARAM;
	CK(cudaSetDevice(c->device));
	CK(cudaStreamSynchronize(c->stream));
	const size_t slots = c->h_exists.size(), off_e = 16, off_l = 16 + (slots + 7) / 8 * 8;
	const char *pin = (const char *)c->pin_sum.ptr;
	if (!pin || c->pin_sum.cap < off_l + slots * sizeof(double)) { cnb_set_error("cnb_ctx_dp_wait without cnb_ctx_dp_launch"); return CNB_ERR_PROC; }
	memcpy(c->h_exists.data(), pin + off_e, slots);
	memcpy(c->h_L.data(), pin + off_l, slots * sizeof(double));
	if (*(const int *)pin) { cnb_set_error("conditional table exceeds %d cells", CNB_HIST_MAX_CELLS); return CNB_ERR_MEM; }
	c->timing.select_ms = ev_ms(c, EV_SEL0, EV_SEL1);
	c->timing.dp_ms = ev_ms(c, EV_SEL1, EV_DP1);
	c->dp_done = true;
	return CNB_OK;
}

extern "C" int cnb_ctx_select_dp(cnb_ctx *c, const double *scores_dev) {
	RC(cnb_ctx_select(c, scores_dev));
	RC(cnb_ctx_dp_launch(c));
	return cnb_ctx_dp_wait(c);
}

/* host view of the DP outcome: (complexity, log-lik) of every surviving network */
int cnb_ctx_dp_summary(const cnb_ctx *c, std::vector<int32_t> *complexity, std::vector<double> *loglik) {
	if (!c || !c->dp_done) return CNB_ERR_PROC;
	complexity->clear();
	loglik->clear();
	for (size_t j = 0; j < c->h_exists.size(); j++)
		if (c->h_exists[j]) {
			complexity->push_back((int32_t)j);
			loglik->push_back(c->h_L[j]);
		}
	return CNB_OK;
}

/* parents of ONE surviving network (cnSearchHist needs nothing else): back-pointer walk for its slot, option ->
 * candidate parent set on the host */
int cnb_ctx_network_parents(cnb_ctx *c, int32_t complexity, std::vector<int32_t> *num_parents, std::vector<int32_t> *parents) {
	if (!c || !c->dp_done || !num_parents || !parents) return CNB_ERR_PROC;
	CK(cudaSetDevice(c->device));
	CnbPlan &pl = c->plan;
	cudaStream_t st = c->stream;
	const int N = pl.N, K = pl.K;
	if (complexity < 0 || complexity > K || !c->h_exists[complexity]) return CNB_ERR_PARAM;
	size_t slots = (size_t)K + 1;
	DpState fin;
	fin.K = K;
	fin.exists = (unsigned char *)c->b_dp_exists.ptr + c->dp_final * slots;
	fin.L = (double *)c->b_dp_L.ptr + c->dp_final * slots;
	fin.pfx = (double *)c->b_dp_pfx.ptr + c->dp_final * slots;
	std::vector<int32_t> one(1, complexity);
	RC(c->b_sel.ensure((size_t)N * sizeof(long long)));
	RC(upload(c, c->b_sel_slots, one));
	RC(cnbk_dp_collect_list(st, fin, N, (const long long *)c->b_bp_opt.ptr, (const int *)c->b_bp_src.ptr,
			(const int *)c->b_sel_slots.ptr, 1, (long long *)c->b_sel.ptr));
	c->timing.kernel_launches++;
	/* a small option table comes back whole with the selection (one wait); a large one entry by entry after it */
	const size_t nopt = (size_t)pl.num_options;
	const bool whole = nopt > 0 && nopt <= ((size_t)1 << 17);
	RC(c->pin_sel.ensure(std::max<size_t>((2 * (size_t)N + (whole ? nopt : 0)) * sizeof(long long), 16)));
	long long *sel = (long long *)c->pin_sel.ptr, *fid = sel + N, *all = fid + N;
	D2H(c, sel, c->b_sel.ptr, (size_t)N * sizeof(long long));
	if (whole) D2H(c, all, c->b_opt_fid.ptr, nopt * sizeof(long long));
	CK(cudaStreamSynchronize(st));
	if (whole) {
		for (int n = 0; n < N; n++)
			if (sel[n] >= 0) fid[n] = all[sel[n]];
	} else {
		for (int n = 0; n < N; n++)
			if (sel[n] >= 0) D2H(c, &fid[n], (const long long *)c->b_opt_fid.ptr + sel[n], sizeof(long long));
		CK(cudaStreamSynchronize(st));
	}
	num_parents->assign(N, 0);
	parents->assign((size_t)N * CNB_MAXP, 0);
	for (int n = 0; n < N; n++) {
		const CnbNodePlan &nd = pl.nodes[n];
		int32_t *out = parents->data() + (size_t)n * CNB_MAXP;
		if (sel[n] < 0) {                         /* base family: fixed parents only */
			(*num_parents)[n] = nd.nfix;
			for (int k = 0; k < nd.nfix; k++) out[k] = pl.lists[nd.fix_off + k];
			continue;
		}
		int bi = nd.first_blk;
		while (bi + 1 < nd.first_blk + nd.nblk && pl.blocks[bi + 1].opt_off <= sel[n]) bi++;
		const CnbBlock &b = pl.blocks[bi];
		int idx[CNB_MAXP];
		cnb_unrank_lex(b.nfree, b.d - b.nfix, fid[n] - b.start, idx);
		for (int k = 0; k < b.nfix; k++) out[k] = pl.lists[b.fix_off + k];
		for (int k = 0; k < b.d - b.nfix; k++) out[b.nfix + k] = pl.lists[b.free_off + idx[k]];
		(*num_parents)[n] = b.d;
	}
	return CNB_OK;
}

/* materialise the networks: back-pointer walk on the device, then tables of the distinct selected families */
extern "C" int cnb_ctx_collect(cnb_ctx *c, cnb_result **out) {
	if (!c || !c->dp_done || !out) return CNB_ERR_PROC;
	CK(cudaSetDevice(c->device));
	CnbPlan &pl = c->plan;
	cudaStream_t st = c->stream;
	const int N = pl.N, K = pl.K;
	size_t slots = (size_t)K + 1;
	CK(cudaEventRecord(c->ev[EV_COL0], st));
	static const bool trace = getenv("CNB_TRACE") != nullptr;
	auto now = [] { return std::chrono::steady_clock::now(); };
	auto ms_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(now() - t).count(); };
	auto t_begin = now();
	double t_sel = 0, t_walk = 0, t_tables = 0;
	DpState fin;
	fin.K = K;
	fin.exists = (unsigned char *)c->b_dp_exists.ptr + c->dp_final * slots;
	fin.L = (double *)c->b_dp_L.ptr + c->dp_final * slots;
	fin.pfx = (double *)c->b_dp_pfx.ptr + c->dp_final * slots;
	/* back-pointer walk for the surviving slots only; the selection table comes back through pinned memory */
	std::vector<int32_t> live;
	for (size_t j = 0; j < slots; j++)
		if (c->h_exists[j]) live.push_back((int32_t)j);
	const size_t nlive = live.size();
	RC(c->b_sel.ensure(std::max<size_t>(nlive, 1) * N * sizeof(long long)));
	RC(upload(c, c->b_sel_slots, live));
	RC(cnbk_dp_collect_list(st, fin, N, (const long long *)c->b_bp_opt.ptr, (const int *)c->b_bp_src.ptr,
			(const int *)c->b_sel_slots.ptr, (int)nlive, (long long *)c->b_sel.ptr));
	c->timing.kernel_launches++;
	const size_t nopt = (size_t)pl.num_options;
	RC(c->pin_sel.ensure(std::max<size_t>((nlive * N + nopt) * sizeof(long long), 16)));
	const long long *sel = (const long long *)c->pin_sel.ptr, *opt_fid = sel + nlive * N;
	if (nlive) D2H(c, c->pin_sel.ptr, c->b_sel.ptr, nlive * N * sizeof(long long));
	if (nopt) D2H(c, (void *)opt_fid, c->b_opt_fid.ptr, nopt * sizeof(long long));
	CK(cudaStreamSynchronize(st));
	t_sel = ms_since(t_begin);

	std::unique_ptr<cnb_result> res(new cnb_result());
	res->N = N;
	std::vector<int32_t> ord1(N);
	for (int i = 0; i < N; i++) ord1[i] = pl.order[i] + 1;
	res->orders.push_back(ord1);
	res->order_cats.push_back(pl.cats);
	/* distinct families: one per option index, one per base family of a node */
	std::vector<int32_t> fam_of_opt(nopt, -1), fam_of_base(N, -1);
	std::vector<int32_t> ex_child, ex_pars, ex_nd, ov_of_fam;
	auto add_family = [&](long long o, int node) -> int {
		int32_t &slot = o >= 0 ? fam_of_opt[(size_t)o] : fam_of_base[node];
		if (slot >= 0) return slot;
		int id = (int)ex_child.size();
		slot = id;
		ex_child.push_back(node);
		size_t base = ex_pars.size();
		ex_pars.resize(base + CNB_MAXP, 0);
		const CnbNodePlan &nd = pl.nodes[node];
		if (o < 0) {
			ex_nd.push_back(nd.nfix);
			for (int k = 0; k < nd.nfix; k++) ex_pars[base + k] = pl.lists[nd.fix_off + k];
		} else {
			/* option -> block (options of a node are laid out block after block) */
			int bi = nd.first_blk;
			while (bi + 1 < nd.first_blk + nd.nblk && pl.blocks[bi + 1].opt_off <= o) bi++;
			const CnbBlock &b = pl.blocks[bi];
			long long r = opt_fid[o] - b.start;
			int idx[CNB_MAXP];
			cnb_unrank_lex(b.nfree, b.d - b.nfix, r, idx);
			for (int k = 0; k < b.nfix; k++) ex_pars[base + k] = pl.lists[b.fix_off + k];
			for (int k = 0; k < b.d - b.nfix; k++) ex_pars[base + b.nfix + k] = pl.lists[b.free_off + idx[k]];
			ex_nd.push_back(b.d);
			/* the score-cache model answered this candidate: its parents in the cached listing sequence (cnb_cache.h) */
			auto ov = std::lower_bound(c->cache_ov.begin(), c->cache_ov.end(), o, [](const CnbOverride &a, long long v) { return a.opt < v; });
			if (ov != c->cache_ov.end() && ov->opt == o) {
				for (int k = 0; k < b.d; k++) ex_pars[base + k] = ov->pars[k];
				ov_of_fam.resize(id + 1, -1);
				ov_of_fam[id] = (int32_t)(ov - c->cache_ov.begin());
			}
		}
		return id;
	};
	/* the selection table is networks x nodes (C5: 2,489 x 500 option indices): a few host threads mark the options in
	 * use, the distinct families are numbered (base families, then options ascending), the threads fill in the networks */
	unsigned hw = std::thread::hardware_concurrency();
	const int nthr = nlive * (size_t)N >= ((size_t)1 << 18) ? (int)std::max(1u, std::min(8u, hw ? hw : 1u)) : 1;
	auto over_networks = [&](const std::function<void(size_t, size_t)> &fn) {
		if (nthr == 1) { fn(0, nlive); return; }
		std::vector<std::thread> pool;
		for (int t = 0; t < nthr; t++) pool.emplace_back(fn, nlive * t / nthr, nlive * (t + 1) / nthr);
		for (auto &th : pool) th.join();
	};
	std::vector<uint8_t> used_opt(nopt, 0), used_base(N, 0);
	over_networks([&](size_t e0, size_t e1) {
		for (size_t e = e0; e < e1; e++)
			for (int n = 0; n < N; n++) {
				const long long o = sel[e * N + n];
				if (o >= 0) used_opt[(size_t)o] = 1;               /* idempotent byte stores: threads may repeat each other */
				else used_base[n] = 1;
			}
	});
	for (int n = 0; n < N; n++)
		if (used_base[n]) add_family(-1, n);
	for (int n = 0; n < N; n++)
		for (long long o = pl.nodes[n].opt_begin; o < pl.nodes[n].opt_end; o++)
			if (used_opt[(size_t)o]) add_family(o, n);
	res->nets.resize(nlive);
	over_networks([&](size_t e0, size_t e1) {
		for (size_t e = e0; e < e1; e++) {
			const size_t j = (size_t)live[e];
			CnbNet &net = res->nets[e];
			net.complexity = (int32_t)j;
			net.loglik = c->h_L[j];
			net.fam.resize(N);
			for (int n = 0; n < N; n++) {
				const long long o = sel[e * N + n];
				net.fam[n] = o >= 0 ? fam_of_opt[(size_t)o] : fam_of_base[n];
			}
		}
	});
	size_t nf = ex_child.size();
	t_walk = ms_since(t_begin);
	int dmax = 0;
	for (int v : ex_nd) dmax = std::max(dmax, v);
	int stride = table_stride(c->max_cats, dmax);
	if (nf) {
		/* tables of the distinct families, grouped by parent count so that each group runs on its fast kernel */
		std::vector<int32_t> perm(nf), pos(nf);
		for (size_t f = 0; f < nf; f++) perm[f] = (int32_t)f;
		std::stable_sort(perm.begin(), perm.end(), [&](int32_t u, int32_t v) { return ex_nd[u] < ex_nd[v]; });
		std::vector<int32_t> s_child(nf), s_pars(nf * CNB_MAXP), s_nd(nf);
		for (size_t k = 0; k < nf; k++) {
			const int32_t f = perm[k];
			pos[f] = (int32_t)k;
			s_child[k] = ex_child[f];
			s_nd[k] = ex_nd[f];
			for (int q = 0; q < CNB_MAXP; q++) s_pars[k * CNB_MAXP + q] = ex_pars[(size_t)f * CNB_MAXP + q];
		}
		RC(score_explicit_groups(c, s_child, s_pars, s_nd, stride, true));     /* families in the order space of the plan */
		std::vector<int32_t> counts(nf * stride), ncount(nf);
		std::vector<double> ll(nf), prior(nf);
		D2H(c, counts.data(), c->b_ex_counts.ptr, counts.size() * sizeof(int));
		D2H(c, ncount.data(), c->b_ex_ncount.ptr, nf * sizeof(int));
		D2H(c, ll.data(), c->b_ex_ll.ptr, nf * sizeof(double));
		D2H(c, prior.data(), c->b_ex_prior.ptr, nf * sizeof(double));
		CK(cudaEventRecord(c->ev[EV_COL1], st));
		CK(cudaStreamSynchronize(st));
		t_tables = ms_since(t_begin);
		res->fams.resize(nf);
		for (size_t f = 0; f < nf; f++) {
			CnbFamily &fm = res->fams[f];
			fm.child = ex_child[f];
			fm.d = ex_nd[f];
			fm.cc = pl.cats[fm.child];
			int ts = fm.cc;
			for (int k = 0; k < fm.d; k++) {
				fm.pars[k] = ex_pars[f * CNB_MAXP + k];
				fm.pc[k] = pl.cats[fm.pars[k]];
				ts *= fm.pc[k];
			}
			const size_t k = (size_t)pos[f];              /* where the grouped launch put this family */
			fm.counts.assign(counts.begin() + k * stride, counts.begin() + k * stride + ts);
			fm.loglik = ll[k];
			fm.prior = prior[k];
			if (pl.has_prior && f < ov_of_fam.size() && ov_of_fam[f] >= 0) fm.prior = c->cache_ov[ov_of_fam[f]].prior;   /* the prior the cache stored */
			fm.ncount = ncount[k];
		}
		c->timing.collect_ms = ev_ms(c, EV_COL0, EV_COL1);
	}
	if (trace)
		fprintf(stderr, "[cnb] collect: %zu networks, %zu distinct families; selection on the host after %.2f ms, families listed %.2f, tables back %.2f, "
				"result built %.2f ms\n", nlive, nf, t_sel, t_walk, t_tables, ms_since(t_begin));
	*out = res.release();
	return CNB_OK;
}

extern "C" int cnb_ctx_finish(cnb_ctx *c, const double *scores_dev, cnb_result **out) {
	RC(cnb_ctx_select_dp(c, scores_dev));
	RC(cnb_ctx_collect(c, out));
	c->timing.total_ms = c->timing.plan_ms + c->timing.score_ms + c->timing.select_ms + c->timing.dp_ms + c->timing.collect_ms;
	return CNB_OK;
}

/* plan + scores + option table of one order: everything the score-cache model does not touch (cnb_sa.cpp runs it for the
 * proposals of a step side by side; cnb_ctx_dp_launch / cnb_ctx_dp_wait finish the evaluation) */
int cnb_ctx_search_to_options(cnb_ctx *c, const cnb_params *p) {
	int64_t seg = 0, nf = 0;
	RC(cnb_ctx_plan(c, p, 0, 1, &seg, &nf));
	RC(c->b_scores.ensure(std::max<int64_t>(seg, 1) * sizeof(double)));
	RC(cnb_ctx_score_shard(c, (double *)c->b_scores.ptr));
	return cnb_ctx_select(c, (const double *)c->b_scores.ptr);
}

int cnb_ctx_search_light(cnb_ctx *c, const cnb_params *p) {
	int64_t seg = 0, nf = 0;
	/* CNB_TRACE=1: host wall time per phase, accumulated per context and printed when it is destroyed */
	static const bool trace = getenv("CNB_TRACE") != nullptr;
	auto now = [] { return std::chrono::steady_clock::now(); };
	auto t0 = now();
	RC(cnb_ctx_plan(c, p, 0, 1, &seg, &nf));
	auto t1 = now();
	RC(c->b_scores.ensure(std::max<int64_t>(seg, 1) * sizeof(double)));
	RC(cnb_ctx_score_shard(c, (double *)c->b_scores.ptr));
	auto t2 = now();
	RC(cnb_ctx_select_dp(c, (const double *)c->b_scores.ptr));
	auto t3 = now();
	if (trace) {
		c->host_ms[0] += std::chrono::duration<double, std::milli>(t1 - t0).count();
		c->host_ms[1] += std::chrono::duration<double, std::milli>(t2 - t1).count();
		c->host_ms[2] += std::chrono::duration<double, std::milli>(t3 - t2).count();
		c->host_ms[3] += 1;
	}
	return CNB_OK;
}

extern "C" int cnb_ctx_search_order(cnb_ctx *c, const cnb_params *p, cnb_result **out) {
	if (!out) return CNB_ERR_PARAM;
	*out = nullptr;
	RC(cnb_ctx_search_light(c, p));
	RC(cnb_ctx_collect(c, out));
	c->timing.total_ms = c->timing.plan_ms + c->timing.score_ms + c->timing.select_ms + c->timing.dp_ms + c->timing.collect_ms;
	return CNB_OK;
}

/* ------------------------------------------------------------------ streamed search ---- */
/* A search on a large complete matrix that is still on its way to the device: host -> device copy, packing and the
 * tensor-core contraction overlap.  The planner needs no samples (the categories are given), so the search is planned
 * first (cnb_ctx_stream_begin); the samples then come in SLICES of whole sample groups (cnb_ctx_stream_slice): while the
 * caller moves slice k + 1 (host threads narrow it to bytes, the copy engine / NVLink carries it), the packing stream
 * turns slice k into bit planes and operand rows (k_pack / k_permute / k_umma_rows on the slice's words -- memory-bound
 * kernels that share the SMs with the table kernel's one CTA per SM) and the context's stream runs one pass of the table
 * kernel over ALL of this rank's tiles for the slice's sample groups -- the first pass stores, the later ones add (each tile
 * is owned by one CTA per pass: plain 16-byte read-modify-writes).  The pair tables are contracted the same way (P tiles on
 * the same order-space operand rows).  cnb_ctx_stream_end completes the pair tables and scores this rank's shard as the
 * resident path does (the tiled group finds its tables in place).  A problem that does not qualify gets no slices; a
 * missing value or a bad category found while packing makes stream_end report ok = 0: the caller then takes the resident
 * path (the byte matrix is on the device by then). */
static void stream_abort(cnb_ctx *c) {
	c->streaming = false;
	c->stream_active = false;
	c->have_samples = c->planned = false;
	c->clean = c->pairs_valid = c->pairs_full = false;
}

void cnb_ctx_stream_cancel(cnb_ctx *c) {
	if (!c) return;
	cudaSetDevice(c->device);
	if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
	if (c->prep_stream) cudaStreamSynchronize(c->prep_stream);
	cudaStreamSynchronize(c->stream);
	stream_abort(c);
}

extern "C" int cnb_ctx_stream_begin(cnb_ctx *c, const cnb_params *p, int32_t rank, int32_t world, int64_t *seg_len,
		int64_t *num_families, int32_t *num_slices, int64_t *slice_rows, int32_t slice_cap) {
	if (!c || !p || !num_slices) return CNB_ERR_PARAM;
	*num_slices = 0;
	c->stream_active = false;
	const int N = p->num_nodes, M = p->num_samples;
	if (!p->num_cats || p->perturbations || p->parents_pool || p->fixed_parents || p->parent_sizes || N < 3
			|| M < 8192 || M >= (1 << 24) || p->max_parent_set < 2
			|| c->force_generic || c->tensor_mode < 0 || c->tensor_i8 || c->no_ie || c->force_full || c->pairs_popc) return CNB_OK;
	int max_cats = 0;
	for (int i = 0; i < N; i++) {
		if (p->num_cats[i] < 1 || p->num_cats[i] > CNB_PLANE_CATS) return CNB_OK;
		max_cats = std::max(max_cats, p->num_cats[i]);
	}
	CK(cudaSetDevice(c->device));
	cudaStream_t st = c->stream;
	if (!c->prep_stream) {
		/* highest priority: the packing kernels of slice k + 1 must be dispatched WHILE the table kernel's pass over slice k
		 * still has thousands of CTAs waiting (measured without it: the block scheduler starts them only when the pass has
		 * issued its last CTA, i.e. every pass was followed by a gap of one packing time) */
		int lo = 0, hi = 0;
		CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
		CK(cudaStreamCreateWithPriority(&c->prep_stream, cudaStreamNonBlocking, hi));
	}
	/* ---- what cnb_ctx_set_samples would set up, minus the data ---- */
	c->have_samples = c->planned = c->dp_done = c->tables_resident = false;
	c->N = N; c->M = M; c->W = (M + 31) / 32; c->Mpad = c->W * 32;
	c->has_pert = false;
	c->timing.h2d_bytes = c->timing.d2h_bytes = 0;
	c->cats_lbl.assign(p->num_cats, p->num_cats + N);
	c->max_cats = max_cats;
	c->never_kept.assign(N, 0);
	c->any_never_kept = false;
	c->pairs_ord = c->triples_valid = false;
	c->has_na = false;
	const size_t words = (size_t)c->W * N;
	std::vector<int> vmin(N, 1);
	CK(cudaEventRecord(c->ev[EV_PACK0], st));
	RC(c->b_vmin.ensure(N * sizeof(int)));
	RC(c->b_cats_lbl.ensure(N * sizeof(int)));
	RC(c->b_flag.ensure(4 * sizeof(int)));
	RC(c->b_codes.ensure((size_t)N * c->Mpad));
	RC(c->b_planes_lbl.ensure(words * sizeof(uint4)));
	RC(c->b_planes.ensure(words * sizeof(uint4)));
	RC(c->b_pairs.ensure((size_t)N * (N - 1) / 2 * 16 * sizeof(int)));
	H2D(c, c->b_vmin.ptr, vmin.data(), N * sizeof(int));
	H2D(c, c->b_cats_lbl.ptr, c->cats_lbl.data(), N * sizeof(int));
	CK(cudaMemsetAsync(c->b_flag.ptr, 0, 4 * sizeof(int), st));
	/* ---- plan (as for complete data: inclusion-exclusion tiles) ---- */
	c->clean = c->pairs_valid = c->pairs_full = true;          /* provisional: checked against the pack flags in stream_end */
	c->have_samples = true;
	c->streaming = true;
	int64_t seg = 0, nf = 0;
	int rc = cnb_ctx_plan(c, p, rank, world, &seg, &nf);
	const CnbPlan &pl = c->plan;
	const CnbGroup *g2 = nullptr;
	for (const CnbGroup &g : pl.groups) if (g.tiled) g2 = &g;
	if (rc != CNB_OK || !g2 || g2->tiled != 1 || pl.tile_full || g2->tiles_per_rank > 16384) {
		stream_abort(c);
		return rc;                                                 /* not a tiled search after all: the resident path */
	}
	c->stream_tiles = (g2->ntiles - pl.rank + pl.world - 1) / pl.world;   /* this rank's tiles (dealt round-robin) */
	const size_t slot_bytes = (size_t)cnb_geo_slot_ints(pl.tile_fr) * sizeof(int) * pl.tile_pairs * pl.tile_cols;
	RC(c->b_counts.ensure((size_t)std::max<long long>(c->stream_tiles, 1) * slot_bytes));
	RC(c->b_counts_p.ensure((size_t)cnbk_pair_tiles(N) * CNB_TILE_PAIRS * CNB_TILE_CHILD * CNBK_UMMA_SLOT_INTS * sizeof(int)));
	/* ---- slices: whole sample groups, a small first one ---- */
	int ng_all = 0;
	c->stream_nq = cnbk_umma_row_quads(c->W, 1, &ng_all);
	const int gs = cnbk_umma_group_samples(1);
	const char *e_sl = getenv("CNB_STREAM_SLICES");
	std::vector<int> parts;
	/* one device: the table kernel outlasts the upload, so a small first slice starts it early (measured at C5: 174 / 170 /
	 * 178 ms for 5 / 3 / 2 slices); four ranks and more: the upload outlasts each rank's share of the contraction, so the
	 * last slice is small too -- its pass is all that is left when the upload ends */
	for (const char *q = e_sl ? e_sl : (world >= 4 ? "1,3,4,4,3,1" : "1,5,10"); *q;) {
		parts.push_back(std::max(1, atoi(q)));
		while (*q && *q != ',') q++;
		if (*q == ',') q++;
	}
	int psum = 0;
	for (int v : parts) psum += v;
	c->stream_gend.clear();                                        /* last group (exclusive) of every slice */
	for (size_t k = 0, acc = 0; k < parts.size(); k++) {
		acc += parts[k];
		int ge = (int)((long long)ng_all * acc / psum);
		if (k + 1 == parts.size()) ge = ng_all;
		if (ge > (c->stream_gend.empty() ? 0 : c->stream_gend.back())) c->stream_gend.push_back(ge);
	}
	const int ns = (int)c->stream_gend.size();
	if (slice_rows && slice_cap < ns + 1) { stream_abort(c); cnb_set_error("cnb_ctx_stream_begin: %d slices", ns); return CNB_ERR_PARAM; }
	while ((int)c->ev_slices.size() < 4 * ns) {                    /* per slice: "copied" (the one-shot form), "here", "packed", "contracted" */
		cudaEvent_t ev;
		CK(cudaEventCreate(&ev));                                  /* with timing: CNB_TRACE prints the slices' time line */
		c->ev_slices.push_back(ev);
	}
	if (slice_rows) {
		slice_rows[0] = 0;
		for (int k = 0; k < ns; k++) slice_rows[k + 1] = k + 1 == ns ? M : std::min<int64_t>(M, (int64_t)c->stream_gend[k] * gs);
	}
	/* the plan's uploads (order, categories) precede the packing */
	CK(cudaEventRecord(c->ev[EV_PLAN1], st));
	CK(cudaStreamWaitEvent(c->prep_stream, c->ev[EV_PLAN1], 0));
	CK(cudaEventRecord(c->ev[EV_SCORE0], st));
	c->stream_active = true;
	c->stream_next = 0;
	*num_slices = ns;
	if (seg_len) *seg_len = seg;
	if (num_families) *num_families = nf;
	return CNB_OK;
}

/* rows [slice_rows[k], slice_rows[k + 1]) of the byte matrix [M][N] (base bytes_dev) are complete once ready_event (a
 * cudaEvent_t) has happened -- NULL: once the work already queued on the context's stream has */
extern "C" int cnb_ctx_stream_slice(cnb_ctx *c, int32_t k, const uint8_t *bytes_dev, void *ready_event) {
	if (!c || !c->stream_active || k != c->stream_next || k >= (int)c->stream_gend.size() || !bytes_dev) {
		cnb_set_error("cnb_ctx_stream_slice: slices come in order after cnb_ctx_stream_begin");
		return CNB_ERR_PARAM;
	}
	CK(cudaSetDevice(c->device));
	cudaStream_t st = c->stream, ps = c->prep_stream;
	const int N = c->N, M = c->M, gs = cnbk_umma_group_samples(1), ns = (int)c->stream_gend.size();
	const int g0 = k ? c->stream_gend[k - 1] : 0, g1 = c->stream_gend[k];
	const bool last = k + 1 == ns;
	const size_t r0 = std::min<size_t>(M, (size_t)g0 * gs), r1 = last ? (size_t)M : std::min<size_t>(M, (size_t)g1 * gs);
	if (ready_event) CK(cudaStreamWaitEvent(ps, (cudaEvent_t)ready_event, 0));
	else {
		CK(cudaEventRecord(c->ev_slices[4 * k], st));
		CK(cudaStreamWaitEvent(ps, c->ev_slices[4 * k], 0));
	}
	CK(cudaEventRecord(c->ev_slices[4 * k + 1], ps));
	const int w0 = (int)(r0 / 32), w1 = last ? c->W : (int)(r1 / 32);
	RC(cnbk_pack(ps, bytes_dev, 1, nullptr, N, M, c->W, (int *)c->b_vmin.ptr, (int *)c->b_cats_lbl.ptr, (uint8_t *)c->b_codes.ptr,
			(uint4 *)c->b_planes_lbl.ptr, nullptr, (int *)c->b_flag.ptr, w0, w1));
	RC(cnbk_permute(ps, (const uint4 *)c->b_planes_lbl.ptr, nullptr, (const int *)c->b_order.ptr, N, c->W, (uint4 *)c->b_planes.ptr, nullptr, w0, w1));
	RC(cnbk_umma_rows(ps, (const uint4 *)c->b_planes.ptr, N, c->W, 1, (uint4 *)c->b_rows_a.ptr, (uint4 *)c->b_rows_b.ptr, w0 / 4,
			last ? c->stream_nq : w1 / 4));
	CK(cudaEventRecord(c->ev_slices[4 * k + 2], ps));
	CK(cudaStreamWaitEvent(st, c->ev_slices[4 * k + 2], 0));
	DevData D = make_devdata(c);
	static int acc_mode = -1;                                      /* how later slices add to the tables: 1 = 16-byte read-modify-write, 2 = RED */
	if (acc_mode < 0) { const char *e = getenv("CNB_STREAM_ACC"); acc_mode = e ? std::max(1, std::min(2, atoi(e))) : 2; }
	RC((int)cnbk_pair_tile_run(st, (const uint4 *)c->b_rows_a.ptr, (const uint4 *)c->b_rows_b.ptr, N, c->W, (int *)c->b_counts_p.ptr, g0, g1, k > 0 ? acc_mode : 0));
	RC(cnbk_umma_tables(st, (const uint4 *)c->b_rows_a.ptr, (const uint4 *)c->b_rows_b.ptr, N, c->W, 1, D.tiles, 0, c->stream_tiles,
			(int *)c->b_counts.ptr, g0, g1, k > 0 ? acc_mode : 0));
	CK(cudaEventRecord(c->ev_slices[4 * k + 3], st));
	c->timing.kernel_launches += 5;
	c->stream_next = k + 1;
	return CNB_OK;
}

extern "C" int cnb_ctx_stream_end(cnb_ctx *c, double *scores_dev, int32_t *ok) {
	if (!c || !ok || !c->stream_active || !scores_dev) { cnb_set_error("cnb_ctx_stream_end: no streamed search"); return CNB_ERR_PARAM; }
	*ok = 0;
	CK(cudaSetDevice(c->device));
	cudaStream_t st = c->stream;
	const bool complete = c->stream_next == (int)c->stream_gend.size();
	int flags[4] = {0, 0, 0, 0};
	D2H(c, flags, c->b_flag.ptr, sizeof(flags));
	cudaError_t e1 = cudaStreamSynchronize(c->prep_stream), e2 = cudaStreamSynchronize(st);
	if (e1 != cudaSuccess || e2 != cudaSuccess) {
		stream_abort(c);
		cnb_set_error("CUDA: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
		return CNB_ERR_CUDA;
	}
	if (!complete || flags[0] || flags[2]) {
		/* a value outside its node's categories or a missing value: the resident path decides (it reports the error, or
		 * scores the data with the kernels that take missing values) */
		stream_abort(c);
		return CNB_OK;
	}
	c->stream_active = false;
	if (getenv("CNB_TRACE")) {                                     /* the slices' time line on the device, ms after the plan */
		for (int k = 0; k < (int)c->stream_gend.size(); k++) {
			float a = 0, b = 0, d = 0;
			cudaEventElapsedTime(&a, c->ev[EV_SCORE0], c->ev_slices[4 * k + 1]);
			cudaEventElapsedTime(&b, c->ev[EV_SCORE0], c->ev_slices[4 * k + 2]);
			cudaEventElapsedTime(&d, c->ev[EV_SCORE0], c->ev_slices[4 * k + 3]);
			fprintf(stderr, "[cnb] dev %d slice %d (groups < %d): here %.2f  packed %.2f  contracted %.2f ms\n", c->device, k, c->stream_gend[k], a, b, d);
		}
	}
	int rc = cnbk_pairs_from_tiles_order(st, (const int *)c->b_counts_p.ptr, (const int *)c->b_order.ptr, c->N, c->M, (int *)c->b_pairs.ptr);
	CK(cudaEventRecord(c->ev[EV_PACK1], st));
	c->data_epoch++;
	if (rc == CNB_OK) rc = cnb_ctx_score_shard(c, scores_dev);   /* the tiled group finds its tables in place */
	c->streaming = false;
	if (rc != CNB_OK) return rc;
	c->timing.pack_ms = ev_ms(c, EV_PACK0, EV_PACK1);
	*ok = 1;
	return CNB_OK;
}

/* the one-shot form: this process owns the host matrix and one device */
static int search_order_streamed(cnb_ctx *c, const cnb_params *p, cnb_result **out, bool *done) {
	*done = false;
	const int N = p->num_nodes, M = p->num_samples;
	const char *e_min = getenv("CNB_STREAM_MIN");     /* read per call: tests switch it */
	const size_t stream_min = e_min ? (!strcmp(e_min, "off") ? (size_t)-1 : (size_t)strtoull(e_min, nullptr, 10)) : ((size_t)256 << 20);
	if (!p->samples || N < 1 || M < 1 || (size_t)N * M * sizeof(int32_t) < stream_min) return CNB_OK;
	int64_t seg = 0, nf = 0, rows[65];
	int32_t ns = 0;
	RC(cnb_ctx_stream_begin(c, p, 0, 1, &seg, &nf, &ns, rows, 65));
	if (ns < 1) return CNB_OK;
	CK(cudaSetDevice(c->device));
	if (!c->copy_stream) CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
	const size_t total = (size_t)N * M;
	int rc = c->b_stage_s.ensure(total);
	if (rc == CNB_OK) rc = c->b_scores.ensure(std::max<int64_t>(seg, 1) * sizeof(double));
	bool overflow = false;
	for (int k = 0; k < ns && rc == CNB_OK && !overflow; k++) {
		/* host threads narrow slice k to bytes while the copy engine moves the chunks already done and the device works on
		 * the slices before it */
		rc = cnb_ctx_upload_bytes(c, p->samples, (size_t)rows[k] * N, (size_t)rows[k + 1] * N, (uint8_t *)c->b_stage_s.ptr, 16, &overflow,
				c->copy_stream, (size_t)rows[k] * N, total);
		if (rc != CNB_OK || overflow) break;
		cudaError_t e = cudaEventRecord(c->ev_slices[4 * k], c->copy_stream);
		if (e != cudaSuccess) { cnb_set_error("CUDA: %s", cudaGetErrorString(e)); rc = CNB_ERR_CUDA; break; }
		rc = cnb_ctx_stream_slice(c, k, (const uint8_t *)c->b_stage_s.ptr, c->ev_slices[4 * k]);
	}
	cudaStreamSynchronize(c->copy_stream);
	int32_t ok = 0;
	if (rc == CNB_OK) rc = cnb_ctx_stream_end(c, (double *)c->b_scores.ptr, &ok);   /* a label above 255 left slices out: ok = 0 */
	else { cudaStreamSynchronize(c->prep_stream); cudaStreamSynchronize(c->stream); stream_abort(c); }
	if (rc != CNB_OK || !ok) return rc;
	RC(cnb_ctx_select_dp(c, (const double *)c->b_scores.ptr));
	RC(cnb_ctx_collect(c, out));
	c->timing.total_ms = c->timing.pack_ms + c->timing.score_ms + c->timing.select_ms + c->timing.dp_ms + c->timing.collect_ms;
	*done = true;
	return CNB_OK;
}

/* ------------------------------------------------------------------ one-shot entry (the .Call body) ---- */
extern "C" int cnb_search_order(const cnb_params *p, cnb_result **out) {
	if (!p || !out) return CNB_ERR_PARAM;
	*out = nullptr;
	if (cnb_multi_devices(p) > 1 || cnb_multi_devices(p) < -1) return cnb_multi_search_order(p, out);   /* one process, several GPUs: cnb_multi.cu */
	cnb_ctx *c = nullptr;
	RC(cnb_ctx_acquire(p->device, &c));
	/* CNB_FORCE_BITS: the kernel-selection switches of cnb_ctx_force_generic for the one-shot calls (fuzzing the tensor-core
	 * and histogram paths on small problems: tools/fuzz_gpu.py) */
	const char *fb = getenv("CNB_FORCE_BITS");
	cnb_ctx_force_generic(c, fb ? atoi(fb) : 0);
	bool done = false;
	int rc = search_order_streamed(c, p, out, &done);  /* large complete matrices: upload overlapped with the table kernel */
	if (rc == CNB_OK && !done) {
		rc = cnb_ctx_set_samples(c, p->num_nodes, p->num_samples, p->samples, p->perturbations, p->num_cats);
		if (rc == CNB_OK) rc = cnb_ctx_search_order(c, p, out);
	}
	g_last_timing = c->timing;
	g_last_timing.fast_path = done ? 2 : g_last_timing.fast_path;   /* 2 = the streamed form ran */
	cnb_ctx_park(c);
	return rc;
}

/* ------------------------------------------------------------------ test hooks ---- */
extern "C" int cnb_family_scores(cnb_ctx *c, int32_t nfam, int32_t d, const int32_t *children, const int32_t *parents,
		int32_t stride, int32_t *counts_out, double *loglik_out, int32_t *ncount_out, int32_t force_generic) {
	if (!c || !c->have_samples || nfam < 0 || d < 0 || d > CNB_MAXP) return CNB_ERR_PARAM;
	CK(cudaSetDevice(c->device));
	/* identity order, no pools, no prior: families are addressed by label */
	cnb_params q;
	memset(&q, 0, sizeof(q));
	q.num_nodes = c->N;
	q.num_samples = c->M;
	q.max_parent_set = 0;
	q.max_complexity = INT32_MAX / 2;
	RC(cnb_ctx_plan(c, &q, 0, 1, nullptr, nullptr));
	std::vector<int32_t> ch(children, children + nfam), pa((size_t)nfam * CNB_MAXP, 0), nd;
	for (int f = 0; f < nfam; f++)
		for (int k = 0; k < d; k++) pa[(size_t)f * CNB_MAXP + k] = parents[(size_t)f * d + k];
	int need = 1;
	if (counts_out) {
		need = table_stride(c->max_cats, d);
		if (stride < need) { cnb_set_error("table_stride %d < %d", stride, need); return CNB_ERR_PARAM; }
	}
	RC(score_explicit(c, ch, pa, nd, d, stride, counts_out != nullptr, force_generic != 0));
	cudaStream_t st = c->stream;
	if (counts_out) D2H(c, counts_out, c->b_ex_counts.ptr, (size_t)nfam * stride * sizeof(int));
	if (loglik_out) D2H(c, loglik_out, c->b_ex_ll.ptr, nfam * sizeof(double));
	if (ncount_out) D2H(c, ncount_out, c->b_ex_ncount.ptr, nfam * sizeof(int));
	int ovf = 0;
	D2H(c, &ovf, (int *)c->b_flag.ptr + 1, sizeof(int));
	CK(cudaStreamSynchronize(st));
	if (ovf) { cnb_set_error("conditional table exceeds %d cells", CNB_HIST_MAX_CELLS); return CNB_ERR_MEM; }
	return CNB_OK;
}

/* The scores the SEARCH itself produced (cnb_ctx_score_shard, whichever kernel scored the group -- the tcgen05 tables
 * included) for explicit candidate families, read back from the score buffer: children / parents are order POSITIONS,
 * parents ascending, every family with d parents.  bench.py compares them with the reference's scores of the same families
 * on the same matrix ("parity_spot").  A family the plan does not offer (pool, fixed parents, parent-set size) gives NaN. */
extern "C" int cnb_ctx_search_scores(cnb_ctx *c, const double *scores_dev, int32_t nfam, int32_t d, const int32_t *children,
		const int32_t *parents, double *out) {
	if (!c || !c->planned || !scores_dev || nfam < 0 || d < 0 || d > CNB_MAXP || !out) { cnb_set_error("cnb_ctx_search_scores: bad arguments"); return CNB_ERR_PARAM; }
	CK(cudaSetDevice(c->device));
	const CnbPlan &pl = c->plan;
	std::vector<long long> index(nfam, -1);
	CnbTiles T;
	cnb_plan_tiles(pl, &T);
	for (int f = 0; f < nfam; f++) {
		const int child = children[f];
		if (child < 0 || child >= pl.N) continue;
		const CnbNodePlan &nd = pl.nodes[child];
		for (int bi = nd.first_blk; bi < nd.first_blk + nd.nblk; bi++) {
			const CnbBlock &b = pl.blocks[bi];
			if (b.d != d) continue;
			/* the family's parents = the block's fixed parents + a subset of its free pool: lexicographic rank of the subset */
			int idx[CNB_MAXP], k = 0;
			bool ok = true;
			for (int q = 0; q < d && ok; q++) {
				const int par = parents[(size_t)f * d + q];
				bool fixed = false;
				for (int t = 0; t < b.nfix; t++) fixed = fixed || pl.lists[b.fix_off + t] == par;
				if (fixed) continue;
				int pos = -1;
				for (int t = 0; t < b.nfree; t++) if (pl.lists[b.free_off + t] == par) pos = t;
				if (pos < 0 || k >= d - b.nfix) ok = false;
				else idx[k++] = pos;
			}
			if (!ok || k != d - b.nfix) break;
			for (int i = 1; i < k; i++)                      /* k <= 8: insertion sort */
				for (int j = i; j > 0 && idx[j - 1] > idx[j]; j--) std::swap(idx[j - 1], idx[j]);
			long long rank = 0;
			int prev = -1;
			for (int i = 0; i < k; i++) {
				for (int j = prev + 1; j < idx[i]; j++) rank += cnb_binom(b.nfree - 1 - j, k - 1 - i);
				prev = idx[i];
			}
			const CnbGroup &g = pl.groups[b.group];
			const long long fid = b.start + rank;
			if (g.tiled) {
				long long tile;
				int within;
				const int p0 = parents[(size_t)f * d], p1 = parents[(size_t)f * d + 1];
				cnb_family_slot(T, pl.N, std::min(p0, p1), std::max(p0, p1), child, tile, within);
				index[f] = (tile % g.tile_world) * pl.seg_len + g.seg_off + (tile / g.tile_world) * ((long long)T.tp * T.tc) + within;
			} else {
				const long long r = fid / g.chunk;
				index[f] = r * pl.seg_len + g.seg_off + (fid - r * g.chunk);
			}
			break;
		}
	}
	for (int f = 0; f < nfam; f++) {
		if (index[f] < 0) { out[f] = nan(""); continue; }
		CK(cudaMemcpyAsync(&out[f], scores_dev + index[f], sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	}
	CK(cudaStreamSynchronize(c->stream));
	return CNB_OK;
}

/* ------------------------------------------------------------------ pairwise metrics ---- */
/* catnet_entropy.cpp:38-911 on the resident data.  The tables of all ordered pairs are those of the explicit
 * one-parent families (child n1, parent n2) -- counted by the search path's scorers: pair tables from the tensor
 * cores / POPC on complete data, bit planes under the child's perturbation mask, shared-memory histograms beyond 4
 * categories -- then reduced by cnbk_pairwise_metric.  out: double[N*N], out[n2*N + n1]. */
extern "C" int cnb_ctx_pairwise(cnb_ctx *c, int32_t kind, double *out) {
	if (!c || !c->have_samples || !out || kind < CNB_PAIRWISE_ENTROPY || kind > CNB_PAIRWISE_PEARSON) {
		cnb_set_error("cnb_ctx_pairwise: bad arguments");
		return CNB_ERR_PARAM;
	}
	/* the reference decrements every value and takes the range (catnet_entropy.cpp:61-95): a missing value is
	 * undefined behaviour there; here it is an error */
	if (c->has_na) { cnb_set_error("pairwise metrics need complete data (no NA)"); return CNB_ERR_PARAM; }
	CK(cudaSetDevice(c->device));
	const int N = c->N;
	if (kind != CNB_PAIRWISE_ENTROPY && !c->has_pert) {
		/* the "perturbed sub-sample" of KL / Pearson is empty without perturbations (numsamplesPert = 0, :594-602):
		 * every distance is +0 */
		memset(out, 0, (size_t)N * N * sizeof(double));
		return CNB_OK;
	}
	cnb_params q;
	memset(&q, 0, sizeof(q));
	q.num_nodes = N;
	q.num_samples = c->M;
	q.max_parent_set = 0;
	q.max_complexity = INT32_MAX / 2;
	RC(cnb_ctx_plan(c, &q, 0, 1, nullptr, nullptr));           /* identity order */
	cudaStream_t st = c->stream;
	const int stride = table_stride(c->max_cats, 1);
	const size_t nfam = (size_t)N * (N - 1);
	std::vector<int32_t> ch(nfam), pa(nfam * CNB_MAXP, 0), nd;
	for (int n1 = 0, f = 0; n1 < N; n1++)
		for (int n2 = 0; n2 < N; n2++)
			if (n2 != n1) { ch[f] = n1; pa[(size_t)f * CNB_MAXP] = n2; f++; }
	const int *kept = nullptr, *full = nullptr, *diag = nullptr;
	if (N == 1) {
		std::vector<int32_t> c0(1, 0), p0(CNB_MAXP, 0);
		RC(score_explicit(c, c0, p0, nd, 0, stride, true, c->force_generic));
		diag = (const int *)c->b_ex_counts.ptr;
	} else if (kind == CNB_PAIRWISE_ENTROPY) {
		RC(score_explicit(c, ch, pa, nd, 1, stride, true, c->force_generic));
		kept = (const int *)c->b_ex_counts.ptr;
	} else {
		RC(score_explicit(c, ch, pa, nd, 1, stride, true, c->force_generic));
		RC(c->b_pw_kept.ensure(nfam * stride * sizeof(int)));
		CK(cudaMemcpyAsync(c->b_pw_kept.ptr, c->b_ex_counts.ptr, nfam * stride * sizeof(int), cudaMemcpyDeviceToDevice, st));
		kept = (const int *)c->b_pw_kept.ptr;
		c->ignore_pert = true;                                 /* whole-sample tables (:609-611, :832-834) */
		int rc = score_explicit(c, ch, pa, nd, 1, stride, true, c->force_generic);
		c->ignore_pert = false;
		RC(rc);
		full = (const int *)c->b_ex_counts.ptr;
	}
	RC(c->b_pw_out.ensure((size_t)N * N * sizeof(double)));
	RC(cnbk_pairwise_metric(st, kind, N, (const int *)c->b_cats_lbl.ptr, kept, full, diag, stride, (double *)c->b_pw_out.ptr));
	c->timing.kernel_launches++;
	D2H(c, out, c->b_pw_out.ptr, (size_t)N * N * sizeof(double));
	int ovf = 0;
	D2H(c, &ovf, (int *)c->b_flag.ptr + 1, sizeof(int));
	CK(cudaStreamSynchronize(st));
	if (ovf) { cnb_set_error("conditional table exceeds %d cells", CNB_HIST_MAX_CELLS); return CNB_ERR_MEM; }
	return CNB_OK;
}

extern "C" int cnb_ctx_entropy_order(cnb_ctx *c, int32_t *order_out) {
	if (!c || !c->have_samples || !order_out) { cnb_set_error("cnb_ctx_entropy_order: bad arguments"); return CNB_ERR_PARAM; }
	std::vector<double> mat((size_t)c->N * c->N);
	RC(cnb_ctx_pairwise(c, CNB_PAIRWISE_ENTROPY, mat.data()));
	cnb_entropy_order_from_matrix(mat.data(), c->N, order_out);
	return CNB_OK;
}

static int pairwise_one_shot(int32_t kind, int32_t N, int32_t M, const int32_t *samples, const int32_t *pert,
		int32_t device, double *out, int32_t *order_out) {
	if (N < 1 || M < 1 || !samples) { cnb_set_error("Data should be a matrix"); return CNB_ERR_PARAM; }
	cnb_ctx *c = nullptr;
	RC(cnb_ctx_acquire(device, &c));
	int rc = cnb_ctx_set_samples(c, N, M, samples, pert, nullptr);   /* categories = value range, as the reference */
	if (rc == CNB_OK) rc = order_out ? cnb_ctx_entropy_order(c, order_out) : cnb_ctx_pairwise(c, kind, out);
	cnb_ctx_park(c);
	return rc;
}

extern "C" int cnb_pairwise(int32_t kind, int32_t N, int32_t M, const int32_t *samples, const int32_t *pert,
		int32_t device, double *out) {
	if (!out) { cnb_set_error("cnb_pairwise: no output"); return CNB_ERR_PARAM; }
	return pairwise_one_shot(kind, N, M, samples, pert, device, out, nullptr);
}

extern "C" int cnb_entropy_order(int32_t N, int32_t M, const int32_t *samples, const int32_t *pert, int32_t device,
		int32_t *order_out) {
	if (!order_out) { cnb_set_error("cnb_entropy_order: no output"); return CNB_ERR_PARAM; }
	return pairwise_one_shot(CNB_PAIRWISE_ENTROPY, N, M, samples, pert, device, nullptr, order_out);
}

/* ------------------------------------------------------------------ a given network ---- */
/* validate a cnb_network against the resident data and flatten it: 0-based parents, table sizes */
static int check_network(cnb_ctx *c, const cnb_network *net, bool need_probs, std::vector<int32_t> *ints,
		std::vector<long long> *off, bool against_data = true) {
	if (!c || (against_data && !c->have_samples) || !net || !net->num_cats || !net->num_parents || !net->prob_offsets
			|| (need_probs && !net->probs) || net->max_parents < 0 || (net->max_parents > 0 && !net->parents) || net->num_nodes < 1) {
		cnb_set_error("network: bad arguments");
		return CNB_ERR_PARAM;
	}
	const int N = net->num_nodes, maxp = std::max(net->max_parents, 1);
	if (against_data && N != c->N) { cnb_set_error("The number of nodes in the object and data should be equal"); return CNB_ERR_PARAM; }
	ints->assign((size_t)2 * N + (size_t)N * maxp, 0);
	off->assign(N + 1, 0);
	for (int i = 0; i < N; i++) {
		if (net->num_cats[i] < 1 || net->num_cats[i] > 254) { cnb_set_error("node %d: %d categories", i + 1, net->num_cats[i]); return CNB_ERR_PARAM; }
		if (against_data && net->num_cats[i] != c->cats_lbl[i]) {
			cnb_set_error("node %d: the object has %d categories, the resident data %d", i + 1, net->num_cats[i], c->cats_lbl[i]);
			return CNB_ERR_PARAM;
		}
		const int d = net->num_parents[i];
		if (d < 0 || d > net->max_parents || d > CNB_MAXP) { cnb_set_error("node %d: %d parents (at most %d)", i + 1, d, CNB_MAXP); return CNB_ERR_PARAM; }
		(*ints)[i] = net->num_cats[i];
		(*ints)[N + i] = d;
		long long size = net->num_cats[i];
		for (int k = 0; k < d; k++) {
			const int q = net->parents[(size_t)i * net->max_parents + k] - 1;
			if (q < 0 || q >= N || q == i) { cnb_set_error("node %d: invalid parent", i + 1); return CNB_ERR_PARAM; }
			(*ints)[(size_t)2 * N + (size_t)i * maxp + k] = q;
			size *= net->num_cats[q];
		}
		if (net->prob_offsets[i + 1] - net->prob_offsets[i] != size || net->prob_offsets[i] < 0) {
			cnb_set_error("node %d: table of %lld cells expected", i + 1, size);
			return CNB_ERR_PARAM;
		}
		(*off)[i] = net->prob_offsets[i];
		(*off)[i + 1] = net->prob_offsets[i + 1];
	}
	return CNB_OK;
}

extern "C" int cnb_ctx_loglik(cnb_ctx *c, const cnb_network *net, int32_t mode, int32_t num_sel, const int32_t *nodes,
		double *out) {
	std::vector<int32_t> ints;
	std::vector<long long> off;
	RC(check_network(c, net, true, &ints, &off));
	if (!out || (mode != CNB_LOGLIK_NODES && mode != CNB_LOGLIK_BYSAMPLE) || num_sel < 0 || (num_sel > 0 && !nodes)) {
		cnb_set_error("cnb_ctx_loglik: bad arguments");
		return CNB_ERR_PARAM;
	}
	CK(cudaSetDevice(c->device));
	cudaStream_t st = c->stream;
	const int N = c->N, maxp = std::max(net->max_parents, 1);
	const long long base = off[0], total = off[N] - base;
	bool any_zero = false;
	for (long long k = 0; k < total && !any_zero; k++) any_zero = !(net->probs[base + k] > 0);
	for (auto &o : off) o -= base;
	RC(upload(c, c->b_net_int, ints));
	RC(upload(c, c->b_net_off, off));
	RC(c->b_net_probs.ensure(std::max<long long>(total, 1) * sizeof(double)));
	RC(c->b_net_logp.ensure(std::max<long long>(total, 1) * sizeof(double)));
	H2D(c, c->b_net_probs.ptr, net->probs + base, total * sizeof(double));
	RC(cnbk_net_logp(st, (const double *)c->b_net_probs.ptr, total, (double *)c->b_net_logp.ptr));
	CnbDevNet dn;
	dn.N = N;
	dn.maxp = maxp;
	dn.cats = (const int *)c->b_net_int.ptr;
	dn.npar = dn.cats + N;
	dn.pars = dn.cats + 2 * N;
	dn.off = (const long long *)c->b_net_off.ptr;
	dn.logp = (const double *)c->b_net_logp.ptr;
	const uint32_t *keep = c->has_pert ? (const uint32_t *)c->b_keep_lbl.ptr : nullptr;
	size_t nout;
	if (mode == CNB_LOGLIK_NODES) {
		std::vector<int32_t> sel(num_sel);
		for (int k = 0; k < num_sel; k++) {
			if (nodes[k] < 1 || nodes[k] > N) { cnb_set_error("Wrong node"); return CNB_ERR_PARAM; }   /* R/catnet.loglik.R:341 */
			sel[k] = nodes[k] - 1;
		}
		nout = num_sel ? num_sel : N;
		if (num_sel) RC(upload(c, c->b_net_sel, sel));
		RC(c->b_pw_out.ensure(nout * sizeof(double)));
		RC(cnbk_net_nodes(st, dn, (const uint8_t *)c->b_codes.ptr, keep, c->Mpad, num_sel ? (const int *)c->b_net_sel.ptr : nullptr,
				(int)nout, (double *)c->b_pw_out.ptr));
		c->timing.kernel_launches = 2;
	} else {
		nout = c->M;
		RC(c->b_pw_out.ensure(nout * sizeof(double)));
		RC(c->b_net_fz.ensure((size_t)N * sizeof(int)));
		RC(cnbk_net_bysample(st, dn, (const uint8_t *)c->b_codes.ptr, keep, c->M, c->Mpad, (int *)c->b_net_fz.ptr, any_zero,
				(double *)c->b_pw_out.ptr));
		c->timing.kernel_launches = any_zero ? 3 : 2;
	}
	D2H(c, out, c->b_pw_out.ptr, nout * sizeof(double));
	CK(cudaStreamSynchronize(st));
	return CNB_OK;
}

extern "C" int cnb_ctx_set_prob(cnb_ctx *c, const cnb_network *net, double *probs_out, double *loglik_out,
		int32_t *sizes_out) {
	std::vector<int32_t> ints;
	std::vector<long long> off;
	RC(check_network(c, net, false, &ints, &off));
	if (!probs_out) { cnb_set_error("cnb_ctx_set_prob: no output"); return CNB_ERR_PARAM; }
	CK(cudaSetDevice(c->device));
	const int N = c->N, maxp = std::max(net->max_parents, 1);
	cnb_params q;
	memset(&q, 0, sizeof(q));
	q.num_nodes = N;
	q.num_samples = c->M;
	q.max_complexity = INT32_MAX / 2;
	RC(cnb_ctx_plan(c, &q, 0, 1, nullptr, nullptr));           /* identity order: families by label */
	/* the families of the network, grouped by parent count (one launch class each) */
	std::vector<int32_t> perm(N), ch(N), pa((size_t)N * CNB_MAXP, 0), nd(N);
	for (int i = 0; i < N; i++) perm[i] = i;
	std::stable_sort(perm.begin(), perm.end(), [&](int u, int v) { return ints[N + u] < ints[N + v]; });
	int dmax = 0;
	for (int k = 0; k < N; k++) {
		const int i = perm[k];
		ch[k] = i;
		nd[k] = ints[N + i];
		dmax = std::max(dmax, nd[k]);
		for (int p = 0; p < nd[k]; p++) pa[(size_t)k * CNB_MAXP + p] = ints[(size_t)2 * N + (size_t)i * maxp + p];
	}
	const int stride = table_stride(c->max_cats, dmax);
	if (stride >= (1 << 24)) { cnb_set_error("conditional table too large"); return CNB_ERR_MEM; }
	RC(score_explicit_groups(c, ch, pa, nd, stride));
	cudaStream_t st = c->stream;
	std::vector<int32_t> counts((size_t)N * stride), ncount(N), kept(N, c->M);
	std::vector<double> ll(N);
	D2H(c, counts.data(), c->b_ex_counts.ptr, counts.size() * sizeof(int));
	D2H(c, ncount.data(), c->b_ex_ncount.ptr, N * sizeof(int));
	D2H(c, ll.data(), c->b_ex_ll.ptr, N * sizeof(double));
	if (c->has_pert) {
		RC(c->b_net_fz.ensure((size_t)N * sizeof(int)));
		RC(cnbk_keep_count(st, (const uint32_t *)c->b_keep_lbl.ptr, N, c->W, (int *)c->b_net_fz.ptr));
		D2H(c, kept.data(), c->b_net_fz.ptr, N * sizeof(int));
	}
	int ovf = 0;
	D2H(c, &ovf, (int *)c->b_flag.ptr + 1, sizeof(int));
	CK(cudaStreamSynchronize(st));
	if (ovf) { cnb_set_error("conditional table exceeds %d cells", CNB_HIST_MAX_CELLS); return CNB_ERR_MEM; }
	for (int k = 0; k < N; k++) {
		const int i = perm[k], cc = c->cats_lbl[i];
		const long long size = off[i + 1] - off[i];
		const int32_t *tab = counts.data() + (size_t)k * stride;
		double *dst = probs_out + off[i];
		for (long long b = 0; b < size; b += cc) {                /* problist.h:193-222 */
			double pp = 0;
			for (int x = 0; x < cc; x++) pp += (double)tab[b + x];
			if (pp > 0) {
				pp = 1 / pp;
				for (int x = 0; x < cc; x++) dst[b + x] = (double)tab[b + x] * pp;
			} else {
				const double favg = (double)1 / (double)cc;
				double acc = 0;
				for (int x = 0; x < cc - 1; x++) { dst[b + x] = favg; acc += favg; }
				dst[b + cc - 1] = 1 - acc;
			}
		}
		/* setNodeSampleProb on an empty sample returns -FLT_MAX before touching the table (catnet_class.h:824-826) */
		if (loglik_out) loglik_out[i] = kept[i] < 1 ? CNB_NEG_INF : ll[k];
		if (sizes_out) sizes_out[i] = ncount[k];
	}
	return CNB_OK;
}

/* one-shot forms: what ccnLoglik / ccnNodeLoglik / ccnSetProb call */
template <class Fn>
static int with_network_ctx(const cnb_network *net, int32_t M, const int32_t *samples, const int32_t *pert, int32_t device, Fn fn) {
	if (!net || !net->num_cats || net->num_nodes < 1 || M < 1 || !samples) { cnb_set_error("'data' should be a matrix of categories"); return CNB_ERR_PARAM; }
	cnb_ctx *c = nullptr;
	RC(cnb_ctx_acquire(device, &c));
	int rc = cnb_ctx_set_samples(c, net->num_nodes, M, samples, pert, net->num_cats);
	if (rc == CNB_OK) rc = fn(c);
	cnb_ctx_park(c);
	return rc;
}

extern "C" int cnb_loglik(const cnb_network *net, int32_t M, const int32_t *samples, const int32_t *pert, int32_t mode,
		int32_t device, double *out) {
	return with_network_ctx(net, M, samples, pert, device, [&](cnb_ctx *c) { return cnb_ctx_loglik(c, net, mode, 0, nullptr, out); });
}

extern "C" int cnb_node_loglik(const cnb_network *net, int32_t num_sel, const int32_t *nodes, int32_t M,
		const int32_t *samples, const int32_t *pert, int32_t device, double *out) {
	if (num_sel < 1 || !nodes) { cnb_set_error("Wrong node"); return CNB_ERR_PARAM; }
	return with_network_ctx(net, M, samples, pert, device,
			[&](cnb_ctx *c) { return cnb_ctx_loglik(c, net, CNB_LOGLIK_NODES, num_sel, nodes, out); });
}

extern "C" int cnb_set_prob(const cnb_network *net, int32_t M, const int32_t *samples, const int32_t *pert,
		int32_t device, double *probs_out, double *loglik_out, int32_t *sizes_out) {
	return with_network_ctx(net, M, samples, pert, device,
			[&](cnb_ctx *c) { return cnb_ctx_set_prob(c, net, probs_out, loglik_out, sizes_out); });
}

/* ------------------------------------------------------------------ cnSamples ---- */
/* CATNET::getOrder (catnet_class.h:1119-1157): repeatedly the lowest unplaced node all of whose parents are placed */
static int sampling_order(int N, int maxp, const std::vector<int32_t> &ints, std::vector<int32_t> *order) {
	std::vector<char> found(N, 0);
	order->clear();
	for (int k = 0; k < N; k++) {
		int j = 0;
		for (; j < N; j++) {
			if (found[j]) continue;
			bool ready = true;
			for (int p = 0; p < ints[N + j] && ready; p++) ready = found[ints[(size_t)2 * N + (size_t)j * maxp + p]];
			if (ready) break;
		}
		if (j >= N) { cnb_set_error("the network has a directed cycle"); return CNB_ERR_PARAM; }
		found[j] = 1;
		order->push_back(j);
	}
	return CNB_OK;
}

/* out_dev: int32 [M][N] in device memory; pert_dev: int32 [M][N] in device memory or NULL */
static int samples_on_device(cnb_ctx *c, const cnb_network *net, int32_t M, const int32_t *pert_dev, double na_rate,
		uint64_t seed, int32_t *out_dev) {
	std::vector<int32_t> ints, order;
	std::vector<long long> off;
	RC(check_network(c, net, true, &ints, &off, false));
	if (M < 1 || !out_dev) { cnb_set_error("The number of samples should be integer"); return CNB_ERR_PARAM; }
	const int N = net->num_nodes, maxp = std::max(net->max_parents, 1);
	RC(sampling_order(N, maxp, ints, &order));
	cudaStream_t st = c->stream;
	const long long base0 = off[0], total = off[N] - base0;
	for (auto &o : off) o -= base0;
	RC(upload(c, c->b_net_int, ints));
	RC(upload(c, c->b_net_off, off));
	RC(upload(c, c->b_net_sel, order));
	RC(c->b_net_probs.ensure(std::max<long long>(total, 1) * sizeof(double)));
	H2D(c, c->b_net_probs.ptr, net->probs + base0, total * sizeof(double));
	CnbDevNet dn;
	dn.N = N;
	dn.maxp = maxp;
	dn.cats = (const int *)c->b_net_int.ptr;
	dn.npar = dn.cats + N;
	dn.pars = dn.cats + 2 * N;
	dn.off = (const long long *)c->b_net_off.ptr;
	dn.logp = (const double *)c->b_net_probs.ptr;                /* the sampler reads probabilities, not logs */
	unsigned long long drawn = (unsigned long long)N * (unsigned long long)M;
	const long long *base = nullptr;
	if (pert_dev) {
		/* uniforms are consumed by the unfixed entries only, node after node in sampling order: where each chunk of
		 * samples starts in the stream */
		const int nchunks = (M + cnbk_sample_chunk() - 1) / cnbk_sample_chunk();
		const size_t n = (size_t)N * nchunks;
		RC(c->b_net_fz.ensure(n * sizeof(int)));
		RC(cnbk_sample_count(st, dn, (const int *)c->b_net_sel.ptr, pert_dev, M, (int *)c->b_net_fz.ptr));
		std::vector<int32_t> cnt(n);
		std::vector<long long> start(n);
		D2H(c, cnt.data(), c->b_net_fz.ptr, n * sizeof(int));
		CK(cudaStreamSynchronize(st));
		long long acc = 0;
		for (size_t k = 0; k < n; k++) { start[k] = acc; acc += cnt[k]; }
		drawn = (unsigned long long)acc;
		RC(upload(c, c->b_net_logp, start));
		base = (const long long *)c->b_net_logp.ptr;
	}
	RC(cnbk_sample(st, dn, (const int *)c->b_net_sel.ptr, pert_dev, base, M, seed, out_dev));
	if (na_rate < 0) na_rate = 0;                                /* R/catnet.samples.R:54-57 */
	if (na_rate > 1) na_rate = 1;
	const int k = (int)(na_rate * N);                            /* rcatnet.cpp:601 */
	if (k > 0 && k < N) RC(cnbk_sample_na(st, N, M, k, seed, drawn, out_dev));
	CK(cudaStreamSynchronize(st));
	return CNB_OK;
}

extern "C" int cnb_samples_dev(const cnb_network *net, int32_t M, const int32_t *pert_dev, double na_rate, uint64_t seed,
		int32_t device, int32_t *out_dev) {
	cnb_ctx *c = nullptr;
	RC(cnb_ctx_acquire(device, &c));
	int rc = samples_on_device(c, net, M, pert_dev, na_rate, seed, out_dev);
	cnb_ctx_park(c);
	return rc;
}

extern "C" int cnb_samples(const cnb_network *net, int32_t M, const int32_t *pert, double na_rate, uint64_t seed,
		int32_t device, int32_t *out) {
	if (!net || M < 1 || !out) { cnb_set_error("cnb_samples: bad arguments"); return CNB_ERR_PARAM; }
	cnb_ctx *c = nullptr;
	RC(cnb_ctx_acquire(device, &c));
	const size_t bytes = (size_t)net->num_nodes * M * sizeof(int32_t);
	int rc = c->b_stage_s.ensure(bytes);
	if (rc == CNB_OK && pert) rc = c->b_stage_p.ensure(bytes);
	if (rc == CNB_OK && pert && cudaMemcpyAsync(c->b_stage_p.ptr, pert, bytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) {
		cnb_set_error("CUDA: copy of the perturbations failed");
		rc = CNB_ERR_CUDA;
	}
	if (rc == CNB_OK)
		rc = samples_on_device(c, net, M, pert ? (const int32_t *)c->b_stage_p.ptr : nullptr, na_rate, seed, (int32_t *)c->b_stage_s.ptr);
	if (rc == CNB_OK && (cudaMemcpyAsync(out, c->b_stage_s.ptr, bytes, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess
			|| cudaStreamSynchronize(c->stream) != cudaSuccess)) {
		cnb_set_error("CUDA: copy of the samples failed");
		rc = CNB_ERR_CUDA;
	}
	cnb_ctx_park(c);
	return rc;
}

extern "C" int cnb_device_log(int32_t device, const double *x, int64_t n, double *out) {
	cnb_ctx *c = nullptr;
	RC(cnb_ctx_acquire(device, &c));
	DevBuf a, b;
	int rc = a.ensure(n * sizeof(double));
	if (rc == CNB_OK) rc = b.ensure(n * sizeof(double));
	if (rc == CNB_OK) {
		cudaMemcpyAsync(a.ptr, x, n * sizeof(double), cudaMemcpyHostToDevice, c->stream);
		rc = cnbk_device_log(c->stream, (const double *)a.ptr, n, (double *)b.ptr);
		cudaMemcpyAsync(out, b.ptr, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
		if (cudaStreamSynchronize(c->stream) != cudaSuccess) { cnb_set_error("CUDA: device log failed"); rc = CNB_ERR_CUDA; }
	}
	a.release();
	b.release();
	cnb_ctx_park(c);
	return rc;
}

extern "C" void cnb_host_log(const double *x, int64_t n, double *out) {
	for (int64_t i = 0; i < n; i++) out[i] = cnb_log(x[i]);
}
