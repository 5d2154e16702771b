/* cnb_multi.cu -- one process, several GPUs: the multi-device form of the drop-in entry points.
 *
 * The reference's concurrency lives inside one process: RCatnetSearchSA::search starts a pool of worker threads and hands
 * each an order to evaluate (/root/reference/src/rcatnet_sa.cpp:475-652, src/thread.h, src/pthread.cpp); an R session
 * cannot launch one process per GPU.  So the multi-GPU path an R user takes is this one: a cnb_mctx owns one device
 * context and one host thread per GPU (the calling thread drives the first device), and
 *   - every device copies only its 1/G share of the host sample matrix over its own PCIe link and completes its peers'
 *     copies over NVLink (cudaMemcpyPeerAsync), then packs the full matrix;
 *   - cnSearchOrder: every device plans and scores its shard of the candidate parent sets; the scoring kernels store the
 *     scores straight into the first device's score buffer through NVLink peer memory (no separate exchange step; a peer
 *     copy of the segment where peer access is unavailable); selection, DP and export run on the first device only;
 *   - cnSearchSA: one chain per device, merged per complexity (north_star; R/catnet.search.R:317-332 for the rule);
 *   - cnSearchHist: the orders of a batch side by side, accumulated in the reference's order (bit-identical histogram).
 */
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <limits.h>
#include <algorithm>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "cnb_ctx.h"
#include "cnb_crew.h"
#include "cnb_rng.h"

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cnb_set_error("CUDA: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); return CNB_ERR_CUDA; } } while (0)
#define RC(call) do { int rc_ = (call); if (rc_ != CNB_OK) return rc_; } while (0)

struct cnb_mctx {
	int first = 0, G = 0, active = 1, loaded = 0;
	bool alias = false;                    /* test hook: all G contexts on the ONE device `first` */
	std::vector<cnb_ctx *> cs;
	std::vector<cudaEvent_t> ev_done, ev_share;
	std::vector<std::vector<cudaEvent_t>> ev_slice;   /* streamed search: [device][slice] "this device's share of the slice has reached every peer" */
	cudaEvent_t ev_start = nullptr, ev_mark[2] = {nullptr, nullptr};
	bool peer = false;
	int N = 0, M = 0;
	Crew crew;
};

static double env_double(const char *name, double dflt) {
	const char *e = getenv(name);
	if (!e || !*e) return dflt;
	return strtod(e, nullptr);
}

int cnb_multi_devices(const cnb_params *p) {
	long n = p->num_devices;
	if (n == 0) {
		const char *e = getenv("CNB_NUM_DEVICES");
		if (e && *e) n = !strcmp(e, "all") ? INT_MAX : strtol(e, nullptr, 10);
	}
	if (n < -1) return (int)std::max<long>(n, -16);     /* test hook: |n| contexts on the one device (cnb_mctx_create) */
	if (n <= 1) return 1;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess) return 1;
	return (int)std::max<long>(1, std::min<long>(n, ndev - p->device));
}

extern "C" int cnb_mctx_create(int32_t first, int32_t num_devices, cnb_mctx **out) {
	if (!out) return CNB_ERR_PARAM;
	*out = nullptr;
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev < 1) {
		cnb_set_error("no CUDA device: the scoring path has no CPU fallback (%s)", cudaGetErrorString(e));
		return CNB_ERR_CUDA;
	}
	if (first < 0 || first >= ndev || num_devices == 0 || num_devices < -16) { cnb_set_error("invalid device range %d+%d (have %d)", first, num_devices, ndev); return CNB_ERR_PARAM; }
	std::unique_ptr<cnb_mctx> m(new cnb_mctx());
	m->first = first;
	m->alias = num_devices < 0;
	m->G = m->alias ? -num_devices : std::min(num_devices, ndev - first);
	const int G = m->G;
	const int step = m->alias ? 0 : 1;
	m->cs.assign(G, nullptr);
	m->ev_done.assign(G, nullptr);
	m->ev_share.assign(G, nullptr);
	int rc = CNB_OK;
	for (int g = 0; g < G && rc == CNB_OK; g++) rc = cnb_ctx_acquire(first + g * step, &m->cs[g]);
	/* NVLink peer access between every pair: the scoring kernels of device g store into device 0's score buffer, and the
	 * sample shares travel device to device.  CNB_NO_PEER=1 keeps everything on staged peer copies (test hook). */
	bool peer = G > 1 && !getenv("CNB_NO_PEER");
	for (int g = 0; g < G && rc == CNB_OK && peer && !m->alias; g++)
		for (int h = 0; h < G && peer; h++) {
			if (g == h) continue;
			int can = 0;
			if (cudaDeviceCanAccessPeer(&can, first + g, first + h) != cudaSuccess || !can) { peer = false; break; }
			cudaSetDevice(first + g);
			cudaError_t pe = cudaDeviceEnablePeerAccess(first + h, 0);
			if (pe == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
			else if (pe != cudaSuccess) { cudaGetLastError(); peer = false; }
		}
	m->peer = peer;
	/* with peer access every device stores its scores into the first device's buffer: the three-parent T tiles may then be
	 * dealt to the devices by triple range (cnb_ctx_allow_scatter mode 1) */
	for (int g = 0; g < G && rc == CNB_OK; g++) m->cs[g]->allow_scatter = (peer || m->alias) ? 1 : 0;
	for (int g = 0; g < G && rc == CNB_OK; g++) {
		if (cudaSetDevice(first + g * step) != cudaSuccess || cudaEventCreateWithFlags(&m->ev_done[g], cudaEventDisableTiming) != cudaSuccess
				|| cudaEventCreateWithFlags(&m->ev_share[g], cudaEventDisableTiming) != cudaSuccess) {
			cnb_set_error("CUDA: event creation failed on device %d", first + g);
			rc = CNB_ERR_CUDA;
		}
	}
	if (rc == CNB_OK) {
		cudaSetDevice(first);
		if (cudaEventCreateWithFlags(&m->ev_start, cudaEventDisableTiming) != cudaSuccess || cudaEventCreate(&m->ev_mark[0]) != cudaSuccess
				|| cudaEventCreate(&m->ev_mark[1]) != cudaSuccess) {
			cnb_set_error("CUDA: event creation failed on device %d", first);
			rc = CNB_ERR_CUDA;
		}
	}
	if (rc != CNB_OK) {
		for (int g = 0; g < G; g++) if (m->cs[g]) cnb_ctx_park(m->cs[g]);
		return rc;
	}
	m->crew.start(G);
	*out = m.release();
	return CNB_OK;
}

extern "C" void cnb_mctx_destroy(cnb_mctx *m) {
	if (!m) return;
	m->crew.stop();
	for (int g = 0; g < m->G; g++) {
		if (m->cs[g]) cudaSetDevice(m->cs[g]->device);
		if (m->cs[g]) cudaStreamSynchronize(m->cs[g]->stream);
		if (m->ev_done[g]) cudaEventDestroy(m->ev_done[g]);
		if (m->ev_share[g]) cudaEventDestroy(m->ev_share[g]);
		if (g < (int)m->ev_slice.size()) for (cudaEvent_t ev : m->ev_slice[g]) cudaEventDestroy(ev);
	}
	cudaSetDevice(m->first);
	if (m->ev_start) cudaEventDestroy(m->ev_start);
	for (int i = 0; i < 2; i++) if (m->ev_mark[i]) cudaEventDestroy(m->ev_mark[i]);
	for (int g = 0; g < m->G; g++) if (m->cs[g]) cnb_ctx_park(m->cs[g]);       /* the per-device idle pool keeps the buffers */
	delete m;
}

extern "C" int32_t cnb_mctx_num_devices(const cnb_mctx *m) { return m ? m->G : 0; }
extern "C" int32_t cnb_mctx_active(const cnb_mctx *m) { return m ? m->active : 0; }
extern "C" cnb_ctx *cnb_mctx_ctx(cnb_mctx *m, int32_t i) { return (m && i >= 0 && i < m->G) ? m->cs[i] : nullptr; }

/* ------------------------------------------------------------------ samples ---- */
/* samples on the first `ndev` devices.  Large matrices: sample rows are dealt out, each device uploads its rows (as bytes,
 * compacted by host threads from the possibly pageable source) and sends them to its peers; small ones: every device
 * copies the whole matrix over its own PCIe link. */
static int mctx_load(cnb_mctx *m, int ndev, int32_t N, int32_t M, const int32_t *samples, const int32_t *pert, const int32_t *num_cats) {
	if (!m || N < 1 || M < 1 || !samples) { cnb_set_error("Invalid sample"); return CNB_ERR_PARAM; }
	ndev = std::max(1, std::min(ndev, m->G));
	m->loaded = 0;
	m->N = N;
	m->M = M;
	const size_t total = (size_t)N * M;
	const double share_min = env_double("CNB_MULTI_SHARE_MIN", 32.0 * (1 << 20));      /* bytes of int32 matrix */
	bool shared = ndev > 1 && (double)total * 4 >= share_min;
	if (shared) {
		const int G = ndev;
		const size_t rows = ((size_t)M + G - 1) / G;
		std::vector<char> big(G, 0);
		unsigned hw = std::thread::hardware_concurrency();
		const int per_dev_threads = (int)std::max(1u, std::min(16u, (hw ? hw : 8u) / (unsigned)G));
		/* A: staging buffers of every device exist before any peer writes into them */
		RC(m->crew.run(G, [&](int g) -> int {
			cnb_ctx *c = m->cs[g];
			CK(cudaSetDevice(c->device));
			c->timing.h2d_bytes = c->timing.d2h_bytes = 0;
			RC(c->b_stage_s.ensure(total));
			if (pert) RC(c->b_stage_p.ensure(total * sizeof(int32_t)));
			return CNB_OK;
		}));
		/* B: own share host -> device, then device -> every peer */
		RC(m->crew.run(G, [&](int g) -> int {
			cnb_ctx *c = m->cs[g];
			CK(cudaSetDevice(c->device));
			const size_t v0 = std::min(total, (size_t)g * rows * N), v1 = std::min(total, (size_t)(g + 1) * rows * N);
			bool ovf = false;
			RC(cnb_ctx_upload_bytes(c, samples, v0, v1, (uint8_t *)c->b_stage_s.ptr, per_dev_threads, &ovf));
			big[g] = ovf;
			if (pert && v1 > v0) {
				CK(cudaMemcpyAsync((int32_t *)c->b_stage_p.ptr + v0, pert + v0, (v1 - v0) * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
				c->timing.h2d_bytes += (int64_t)((v1 - v0) * sizeof(int32_t));
			}
			for (int k = 1; k < G && v1 > v0; k++) {
				cnb_ctx *d = m->cs[(g + k) % G];                       /* staggered: no two devices target the same peer first */
				CK(cudaMemcpyPeerAsync((uint8_t *)d->b_stage_s.ptr + v0, d->device, (uint8_t *)c->b_stage_s.ptr + v0, c->device, v1 - v0, c->stream));
				if (pert)
					CK(cudaMemcpyPeerAsync((int32_t *)d->b_stage_p.ptr + v0, d->device, (int32_t *)c->b_stage_p.ptr + v0, c->device,
							(v1 - v0) * sizeof(int32_t), c->stream));
			}
			CK(cudaEventRecord(m->ev_share[g], c->stream));
			return CNB_OK;
		}));
		bool any_big = false;
		for (int g = 0; g < G; g++) any_big = any_big || big[g];
		if (!any_big) {
			/* C: every share has arrived -> pack */
			RC(m->crew.run(G, [&](int g) -> int {
				cnb_ctx *c = m->cs[g];
				CK(cudaSetDevice(c->device));
				for (int h = 0; h < G; h++)
					if (h != g) CK(cudaStreamWaitEvent(c->stream, m->ev_share[h], 0));
				c->in_host_set = true;
				int rc = cnb_ctx_set_samples_impl(c, N, M, c->b_stage_s.ptr, 1, pert ? (const int32_t *)c->b_stage_p.ptr : nullptr, num_cats);
				c->in_host_set = false;
				return rc;
			}));
			m->loaded = G;
			return CNB_OK;
		}
		/* a label above 255 (legal when the categories are inferred): the int32 route below carries it */
		for (int g = 0; g < G; g++) { cudaSetDevice(m->cs[g]->device); cudaStreamSynchronize(m->cs[g]->stream); }
	}
	RC(m->crew.run(ndev, [&](int g) -> int { return cnb_ctx_set_samples(m->cs[g], N, M, samples, pert, num_cats); }));
	m->loaded = ndev;
	return CNB_OK;
}

extern "C" int cnb_mctx_set_samples(cnb_mctx *m, int32_t N, int32_t M, const int32_t *samples, const int32_t *pert,
		const int32_t *num_cats) {
	if (!m) return CNB_ERR_PARAM;
	return mctx_load(m, m->G, N, M, samples, pert, num_cats);
}

/* ------------------------------------------------------------------ cnSearchOrder ---- */
/* devices worth using for this search: the exchange and the extra host hand-offs cost ~0.1 ms, so small problems stay on
 * the first device.  Work = candidate families x samples; CNB_MULTI_MIN_WORK = 0 shards always. */
static int devices_for(const cnb_mctx *m, const cnb_params *p, int M, int avail) {
	if (avail <= 1) return 1;
	const double min_work = env_double("CNB_MULTI_MIN_WORK", 2e10);
	/* unrestricted count sum_d C(N, d + 1): an upper bound under pools / parentSizes, which is all a threshold needs */
	double fam = 0;
	for (int d = 0; d <= p->max_parent_set && d < p->num_nodes; d++) fam += (double)cnb_binom(p->num_nodes, d + 1);
	return fam * (double)M >= min_work ? avail : 1;
}

extern "C" int cnb_mctx_search_light(cnb_mctx *m, const cnb_params *p) {
	if (!m || !p || m->loaded < 1) { cnb_set_error("cnb_mctx: no samples"); return CNB_ERR_PARAM; }
	cnb_params q = *p;
	q.num_nodes = m->N;
	const int A = devices_for(m, &q, m->M, m->loaded);
	m->active = A;
	cnb_ctx *c0 = m->cs[0];
	if (A == 1) return cnb_ctx_search_light(c0, p);
	CK(cudaSetDevice(c0->device));
	CK(cudaEventRecord(m->ev_start, c0->stream));
	std::vector<int64_t> seg(A, 0);
	RC(m->crew.run(A, [&](int g) -> int {
		cnb_ctx *c = m->cs[g];
		CK(cudaSetDevice(c->device));
		if (g) CK(cudaStreamWaitEvent(c->stream, m->ev_start, 0));   /* device time of the search starts on the first device */
		int64_t nf = 0;
		return cnb_ctx_plan(c, p, g, A, &seg[g], &nf);
	}));
	const size_t score_bytes = (size_t)std::max<int64_t>(seg[0] * A, 1) * sizeof(double);
	CK(cudaSetDevice(c0->device));
	RC(c0->b_scores.ensure(score_bytes));
	RC(m->crew.run(A, [&](int g) -> int {
		cnb_ctx *c = m->cs[g];
		CK(cudaSetDevice(c->device));
		double *target = (double *)c0->b_scores.ptr;
		if (g && !m->peer) {
			RC(c->b_scores.ensure(score_bytes));
			target = (double *)c->b_scores.ptr;
		}
		/* with peer access the scoring kernels store each score over NVLink into the first device's buffer: segment g
		 * is complete there when this device's stream drains (cnb_ctx_score_shard returns synchronised) */
		RC(cnb_ctx_score_shard(c, target));
		if (g && !m->peer)
			CK(cudaMemcpyPeerAsync((double *)c0->b_scores.ptr + (size_t)g * seg[0], c0->device, target + (size_t)g * seg[0], c->device,
					(size_t)seg[0] * sizeof(double), c->stream));
		if (g) CK(cudaEventRecord(m->ev_done[g], c->stream));
		return CNB_OK;
	}));
	CK(cudaSetDevice(c0->device));
	for (int g = 1; g < A; g++) CK(cudaStreamWaitEvent(c0->stream, m->ev_done[g], 0));
	return cnb_ctx_select_dp(c0, (const double *)c0->b_scores.ptr);
}

extern "C" int cnb_mctx_collect(cnb_mctx *m, cnb_result **out) {
	if (!m || !out) return CNB_ERR_PARAM;
	return cnb_ctx_collect(m->cs[0], out);
}

extern "C" int cnb_mctx_search_order(cnb_mctx *m, const cnb_params *p, cnb_result **out) {
	if (!out) return CNB_ERR_PARAM;
	*out = nullptr;
	RC(cnb_mctx_search_light(m, p));
	RC(cnb_mctx_collect(m, out));
	cnb_timing &t = m->cs[0]->timing;
	t.total_ms = t.plan_ms + t.score_ms + t.select_ms + t.dp_ms + t.collect_ms;
	return CNB_OK;
}

extern "C" int cnb_mctx_mark(cnb_mctx *m, int32_t which) {
	if (!m || which < 0 || which > 1) return CNB_ERR_PARAM;
	CK(cudaSetDevice(m->cs[0]->device));
	CK(cudaEventRecord(m->ev_mark[which], m->cs[0]->stream));
	return CNB_OK;
}

extern "C" int cnb_mctx_elapsed_ms(cnb_mctx *m, float *ms) {
	if (!m || !ms) return CNB_ERR_PARAM;
	for (int g = m->G - 1; g >= 0; g--) {
		CK(cudaSetDevice(m->cs[g]->device));
		CK(cudaStreamSynchronize(m->cs[g]->stream));
	}
	CK(cudaEventElapsedTime(ms, m->ev_mark[0], m->ev_mark[1]));
	return CNB_OK;
}

/* ------------------------------------------------------------------ cnSearchSA: one chain per device ---- */
extern "C" int cnb_mctx_search_sa(cnb_mctx *m, const cnb_params *p, const cnb_sa_params *sa, cnb_result **out) {
	if (!m || !p || !sa || !out || m->loaded < 1) { cnb_set_error("cnb_mctx: no samples"); return CNB_ERR_PARAM; }
	*out = nullptr;
	const int A = m->loaded;
	m->active = A;
	std::vector<cnb_result *> res(A, nullptr);
	std::vector<int> rcs(A, CNB_OK);
	const uint64_t base = sa->seed ? sa->seed : cnb_default_seed();
	/* chain 0 is the single-device chain (the caller's generator, on the calling thread); the others own a splitmix64
	 * stream: a failing extra chain only costs its networks */
	m->crew.run(A, [&](int g) -> int {
		cnb_sa_params s = *sa;
		if (g) {
			s.unif = nullptr;
			s.unif_ctx = nullptr;
			s.seed = base + 0x632BE59BD9B4E019ull * (uint64_t)g;
			if (!s.seed) s.seed = (uint64_t)g;
		}
		rcs[g] = cnb_ctx_search_sa(m->cs[g], p, &s, &res[g]);
		return g ? CNB_OK : rcs[0];
	});
	if (rcs[0] != CNB_OK) {
		for (int g = 0; g < A; g++) cnb_result_free(res[g]);
		return rcs[0];
	}
	for (int g = 1; g < A; g++) {
		if (rcs[g] == CNB_OK && res[g]) cnb_result_union(res[0], res[g]);
		cnb_result_free(res[g]);
	}
	*out = res[0];
	return CNB_OK;
}

/* ------------------------------------------------------------------ cnSearchHist: orders side by side ---- */
extern "C" int cnb_mctx_search_hist(cnb_mctx *m, const cnb_params *p, const cnb_hist_params *hp, double *hist, int32_t *iterations) {
	if (!m || !p || !hp || !hist || m->loaded < 1) { cnb_set_error("cnb_mctx: no samples"); return CNB_ERR_PARAM; }
	const int N = m->N, A = m->loaded;
	m->active = A;
	const int nDrives = hp->num_threads < 1 ? 1 : hp->num_threads;
	if (cnb_cache_wanted(hp->use_cache, nDrives)) {
		/* the reference's score cache is one table that the orders pass through one after the other (cnb_cache.h): one device */
		m->active = 1;
		return cnb_ctx_search_hist(m->cs[0], p, hp, hist, iterations);
	}
	memset(hist, 0, sizeof(double) * (size_t)N * N);
	CnbRng rng(hp->seed ? hp->seed : cnb_default_seed(), hp->unif, hp->unif_ctx);
	cnb_params q0 = *p;
	q0.edge_prob = nullptr;                      /* rcatnet_hist.cpp:268-277 */
	int niter = 0;
	std::vector<std::vector<int>> orders;
	std::vector<char> skip;
	std::vector<CnbHistEval> evs;
	for (;;) {
		if (niter >= hp->max_iter) break;
		/* the loop (rcatnet_hist.cpp:282-548) draws whole batches of nDrives orders until niter >= maxIter, and a batch adds
		 * at most nDrives: at least ceil(remaining / nDrives) more batches WILL be drawn, so that many can be drawn ahead
		 * without consuming a uniform the reference would not -- enough of them to give every device an order */
		const int remaining = hp->max_iter - niter;
		const int sure = (remaining + nDrives - 1) / nDrives;
		const int nb = std::max(1, std::min(sure, (A + nDrives - 1) / nDrives));
		orders.assign((size_t)nb * nDrives, std::vector<int>(N));
		skip.assign((size_t)nb * nDrives, 0);
		for (int b = 0; b < nb; b++)
			for (int n = 0; n < nDrives; n++) {
				const size_t k = (size_t)b * nDrives + n;
				cnb_gen_permutation(orders[k], N, rng);
				for (int nn = 0; nn < n && !skip[k]; nn++) skip[k] = (orders[(size_t)b * nDrives + nn] == orders[k]);
			}
		std::vector<int> todo;
		for (size_t k = 0; k < orders.size(); k++) if (!skip[k]) todo.push_back((int)k);
		evs.assign(orders.size(), CnbHistEval());
		RC(m->crew.run(std::min<int>(A, std::max<int>((int)todo.size(), 1)), [&](int g) -> int {
			cnb_params q = q0;
			for (size_t t = g; t < todo.size(); t += A) {
				q.order = orders[todo[t]].data();
				RC(cnb_hist_eval_order(m->cs[g], &q, hp, &evs[todo[t]]));
			}
			return CNB_OK;
		}));
		bool all_stop = false;
		for (int b = 0; b < nb; b++) {
			int nstop = 1;
			for (int n = 0; n < nDrives; n++) {
				const size_t k = (size_t)b * nDrives + n;
				if (skip[k]) continue;
				nstop = 0;
				if (!evs[k].found) continue;
				cnb_hist_add(hist, N, orders[k], evs[k]);
				niter++;
				if (p->echo) fprintf(stderr, "Iteration %d\\%d\n", niter, hp->max_iter);
			}
			if (nstop) { all_stop = true; break; }
		}
		if (all_stop) break;
	}
	if (iterations) *iterations = niter;
	return CNB_OK;
}

/* ------------------------------------------------------------------ one-shot forms ---- */
/* the idle multi-device context of the one-shot calls (as cnb_ctx_park for single devices): keeps the worker threads,
 * the peer mappings and the events between .Call's */
static std::mutex g_mmu;
static cnb_mctx *g_mpool = nullptr;

static int mctx_acquire(int first, int G, cnb_mctx **out) {
	{
		std::lock_guard<std::mutex> lk(g_mmu);
		if (g_mpool && g_mpool->first == first && g_mpool->G == std::abs(G) && g_mpool->alias == (G < 0)) {
			*out = g_mpool;
			g_mpool = nullptr;
			(*out)->loaded = 0;
			for (cnb_ctx *c : (*out)->cs) c->have_samples = c->planned = c->dp_done = false;
			return CNB_OK;
		}
	}
	return cnb_mctx_create(first, G, out);
}

static void mctx_park(cnb_mctx *m) {
	if (!m) return;
	bool healthy = true;
	for (int g = 0; g < m->G; g++) {
		cudaSetDevice(m->cs[g]->device);
		if (cudaStreamSynchronize(m->cs[g]->stream) != cudaSuccess || cudaPeekAtLastError() != cudaSuccess) healthy = false;
	}
	if (healthy) {
		std::lock_guard<std::mutex> lk(g_mmu);
		if (!g_mpool) { g_mpool = m; return; }
	}
	cnb_mctx_destroy(m);
}

void cnb_multi_release(void) {
	cnb_mctx *m = nullptr;
	{
		std::lock_guard<std::mutex> lk(g_mmu);
		m = g_mpool;
		g_mpool = nullptr;
	}
	if (m) cnb_mctx_destroy(m);
}

/* The one-shot search on a large complete host matrix, streamed over A devices (the multi-device form of
 * search_order_streamed in cnb_api.cu): every device plans its shard of the candidate parent sets first, then the samples
 * come in slices -- device g narrows its 1/A share of a slice's rows to bytes and copies it over its own PCIe link, sends
 * it to its peers over NVLink (copy engines, on the copy stream), and every device packs the completed slice and runs one
 * pass of the table kernel over its own tiles while the host threads are already narrowing the next slice.  *done = false:
 * the problem does not qualify or the data has missing values -- the caller takes the resident path (from the byte
 * matrices on the devices when they are complete: m->loaded says so). */
static int mctx_search_streamed(cnb_mctx *m, int A, const cnb_params *p, cnb_result **out, bool *done) {
	*done = false;
	const int N = p->num_nodes, M = p->num_samples;
	const char *e_min = getenv("CNB_STREAM_MIN");
	const size_t stream_min = e_min ? (!strcmp(e_min, "off") ? (size_t)-1 : (size_t)strtoull(e_min, nullptr, 10)) : ((size_t)256 << 20);
	if (A < 2 || !p->samples || p->perturbations || N < 1 || M < 1 || (size_t)N * M * sizeof(int32_t) < stream_min) return CNB_OK;
	const size_t total = (size_t)N * M;
	m->loaded = 0;
	m->N = N;
	m->M = M;
	m->active = A;
	std::vector<int64_t> seg(A, 0);
	std::vector<int32_t> ns(A, 0);
	std::vector<std::vector<int64_t>> rows(A, std::vector<int64_t>(65, 0));
	cnb_ctx *c0 = m->cs[0];
	CK(cudaSetDevice(c0->device));
	CK(cudaEventRecord(m->ev_start, c0->stream));
	int rc = m->crew.run(A, [&](int g) -> int {
		cnb_ctx *c = m->cs[g];
		CK(cudaSetDevice(c->device));
		if (g) CK(cudaStreamWaitEvent(c->stream, m->ev_start, 0));
		int64_t nf = 0;
		RC(cnb_ctx_stream_begin(c, p, g, A, &seg[g], &nf, &ns[g], rows[g].data(), 65));
		if (ns[g] < 1) return CNB_OK;
		if (!c->copy_stream) CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
		RC(c->b_stage_s.ensure(total));
		return CNB_OK;
	});
	bool all = rc == CNB_OK;
	for (int g = 0; g < A; g++) all = all && ns[g] == ns[0] && ns[g] > 0 && seg[g] == seg[0];
	auto cancel = [&]() { for (int g = 0; g < A; g++) cnb_ctx_stream_cancel(m->cs[g]); };
	if (!all) { cancel(); return rc; }
	const int S = ns[0];
	if ((int)m->ev_slice.size() < m->G) m->ev_slice.resize(m->G);
	for (int g = 0; g < A; g++) {
		CK(cudaSetDevice(m->cs[g]->device));
		while ((int)m->ev_slice[g].size() < S) {
			cudaEvent_t ev;
			CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
			m->ev_slice[g].push_back(ev);
		}
	}
	const size_t score_bytes = (size_t)std::max<int64_t>(seg[0] * A, 1) * sizeof(double);
	CK(cudaSetDevice(c0->device));
	rc = c0->b_scores.ensure(score_bytes);
	if (rc != CNB_OK) { cancel(); return rc; }
	unsigned hw = std::thread::hardware_concurrency();
	const int per_dev_threads = (int)std::max(1u, std::min(16u, (hw ? hw : 8u) / (unsigned)A));
	std::vector<char> big(A, 0);
	size_t stage_cap = 0;                                          /* values a device stages in all: its share of every slice */
	for (int k = 0; k < S; k++) stage_cap += (size_t)((rows[0][k + 1] - rows[0][k] + A - 1) / A) * N;
	std::vector<size_t> stage_off(A, 0);
	for (int k = 0; k < S && rc == CNB_OK; k++) {
		const int64_t r0 = rows[0][k], r1 = rows[0][k + 1], sh = (r1 - r0 + A - 1) / A;
		/* own share of the slice: host -> device, device -> every peer (both on the copy stream) */
		rc = m->crew.run(A, [&](int g) -> int {
			cnb_ctx *c = m->cs[g];
			CK(cudaSetDevice(c->device));
			const size_t v0 = (size_t)std::min(r1, r0 + g * sh) * N, v1 = (size_t)std::min(r1, r0 + (g + 1) * sh) * N;
			bool ovf = false;
			RC(cnb_ctx_upload_bytes(c, p->samples, v0, v1, (uint8_t *)c->b_stage_s.ptr, per_dev_threads, &ovf, c->copy_stream, stage_off[g], stage_cap));
			stage_off[g] += v1 - v0;
			big[g] = big[g] || ovf;
			for (int j = 1; j < A && v1 > v0; j++) {
				cnb_ctx *d = m->cs[(g + j) % A];
				CK(cudaMemcpyPeerAsync((uint8_t *)d->b_stage_s.ptr + v0, d->device, (uint8_t *)c->b_stage_s.ptr + v0, c->device, v1 - v0, c->copy_stream));
			}
			CK(cudaEventRecord(m->ev_slice[g][k], c->copy_stream));
			return CNB_OK;
		});
		bool any_big = false;
		for (int g = 0; g < A; g++) any_big = any_big || big[g];
		if (rc != CNB_OK || any_big) break;
		/* every share's event is recorded: each device waits for all of them, packs the slice, contracts it */
		rc = m->crew.run(A, [&](int g) -> int {
			cnb_ctx *c = m->cs[g];
			CK(cudaSetDevice(c->device));
			for (int h = 0; h < A; h++)                            /* the packing stream waits, the copy stream goes on with slice k + 1 */
				if (h != g) CK(cudaStreamWaitEvent(c->prep_stream, m->ev_slice[h][k], 0));
			return cnb_ctx_stream_slice(c, k, (const uint8_t *)c->b_stage_s.ptr, m->ev_slice[g][k]);
		});
	}
	bool any_big = false;
	for (int g = 0; g < A; g++) any_big = any_big || big[g];
	if (rc != CNB_OK || any_big) { cancel(); return rc; }       /* a label above 255: the int32 route carries it */
	std::vector<int32_t> ok(A, 0);
	rc = m->crew.run(A, [&](int g) -> int {
		cnb_ctx *c = m->cs[g];
		CK(cudaSetDevice(c->device));
		CK(cudaStreamSynchronize(c->copy_stream));
		double *target = (double *)c0->b_scores.ptr;
		if (g && !m->peer) {
			RC(c->b_scores.ensure(score_bytes));
			target = (double *)c->b_scores.ptr;
		}
		RC(cnb_ctx_stream_end(c, target, &ok[g]));
		if (!ok[g]) return CNB_OK;
		if (g && !m->peer)
			CK(cudaMemcpyPeerAsync((double *)c0->b_scores.ptr + (size_t)g * seg[0], c0->device, target + (size_t)g * seg[0], c->device,
					(size_t)seg[0] * sizeof(double), c->stream));
		if (g) CK(cudaEventRecord(m->ev_done[g], c->stream));
		return CNB_OK;
	});
	bool all_ok = rc == CNB_OK;
	for (int g = 0; g < A; g++) all_ok = all_ok && ok[g];
	if (!all_ok) {
		cancel();
		if (rc != CNB_OK) return rc;
		/* missing values (or a value outside its node's categories: the resident path reports it): the byte matrix is
		 * complete on every device, pack it there */
		RC(m->crew.run(A, [&](int g) -> int {
			cnb_ctx *c = m->cs[g];
			CK(cudaSetDevice(c->device));
			c->in_host_set = true;
			int r = cnb_ctx_set_samples_impl(c, N, M, c->b_stage_s.ptr, 1, nullptr, p->num_cats);
			c->in_host_set = false;
			return r;
		}));
		m->loaded = A;
		return CNB_OK;
	}
	m->loaded = A;
	CK(cudaSetDevice(c0->device));
	for (int g = 1; g < A; g++) CK(cudaStreamWaitEvent(c0->stream, m->ev_done[g], 0));
	RC(cnb_ctx_select_dp(c0, (const double *)c0->b_scores.ptr));
	RC(cnb_ctx_collect(c0, out));
	cnb_timing &t = c0->timing;
	t.total_ms = t.pack_ms + t.score_ms + t.select_ms + t.dp_ms + t.collect_ms;
	*done = true;
	return CNB_OK;
}

int cnb_multi_search_order(const cnb_params *p, cnb_result **out) {
	const int G = cnb_multi_devices(p);
	cnb_mctx *m = nullptr;
	RC(mctx_acquire(p->device, G, &m));
	/* only the devices this search will use receive the samples */
	const int A = devices_for(m, p, p->num_samples, m->G);
	const char *fb = getenv("CNB_FORCE_BITS");                   /* kernel-selection switches for fuzzing (cnb_api.cu) */
	for (int g = 0; g < m->G; g++) cnb_ctx_force_generic(m->cs[g], fb ? atoi(fb) : 0);
	bool done = false;
	int rc = mctx_search_streamed(m, A, p, out, &done);
	if (rc == CNB_OK && !done) {
		if (m->loaded != A) rc = mctx_load(m, A, p->num_nodes, p->num_samples, p->samples, p->perturbations, p->num_cats);
		if (rc == CNB_OK) rc = cnb_mctx_search_order(m, p, out);
	}
	cnb_timing t = m->cs[0]->timing;                              /* the first device's phases, the copies of all devices */
	for (int g = 1; g < m->G; g++) { t.h2d_bytes += m->cs[g]->timing.h2d_bytes; t.d2h_bytes += m->cs[g]->timing.d2h_bytes; }
	if (done) t.fast_path = 2;                                    /* 2 = the streamed form ran */
	cnb_set_last_call_timing(t);
	mctx_park(m);
	return rc;
}

int cnb_multi_search_sa(const cnb_params *p, const cnb_sa_params *sa, cnb_result **out) {
	const int G = cnb_multi_devices(p);
	cnb_mctx *m = nullptr;
	RC(mctx_acquire(p->device, G, &m));
	int rc = mctx_load(m, m->G, p->num_nodes, p->num_samples, p->samples, p->perturbations, p->num_cats);
	if (rc == CNB_OK) rc = cnb_mctx_search_sa(m, p, sa, out);
	mctx_park(m);
	return rc;
}

int cnb_multi_search_hist(const cnb_params *p, const cnb_hist_params *hp, double *hist, int32_t *iterations) {
	const int G = cnb_multi_devices(p);
	cnb_mctx *m = nullptr;
	RC(mctx_acquire(p->device, G, &m));
	int rc = mctx_load(m, m->G, p->num_nodes, p->num_samples, p->samples, p->perturbations, p->num_cats);
	if (rc == CNB_OK) rc = cnb_mctx_search_hist(m, p, hp, hist, iterations);
	mctx_park(m);
	return rc;
}
