/* cnb_network.cu -- likelihood of the data under a GIVEN network (cnLoglik, cnNodeLoglik).
 *
 * CATNET::sampleLoglikVector / sampleNodeLoglik (catnet_class.h:616-687, 767-815) add log P(x_i | parents) over the
 * samples IN SAMPLE ORDER, one fp64 accumulator per node; CATNET::bySampleLoglikVector (:689-765) adds over the nodes
 * in node order, one accumulator per sample.  fp64 addition is not associative, so both orders are kept:
 *   k_net_nodes     one CTA per node: three producer warps turn 1536 samples at a time into log-probabilities
 *                   (uint8 codes, 16 samples per 128-bit load; table of logs evaluated once per cell by k_net_logp with
 *                   glibc's log, cnb_log.h), one consumer lane adds them in order from shared memory while the
 *                   producers fill the other buffer.  The chain of dependent DADDs bounds it (M x add latency);
 *   k_net_bysample  one thread per sample, nodes in order; coalesced byte loads.
 * A zero probability gives -FLT_MAX; in the by-sample pass it also ends the node's pass over the samples (the
 * reference's `break` leaves the sample loop), which k_net_first_zero reproduces with the first such sample per node. */
#include <cuda_runtime.h>
#include <float.h>
#include <limits.h>
#include <math_constants.h>
#include "cnb_kernels.h"
#include "cnb_log.h"

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cnb_set_error("CUDA: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); return CNB_ERR_CUDA; } } while (0)

#define NET_PRODUCERS 96                       /* warps 1..3 */
#define NET_CHUNK (NET_PRODUCERS * 16)         /* samples per buffer */
#define NET_PAD(k) ((k) + ((k) >> 4))          /* one spare double per 16: producer stores spread over the banks */
#define NET_BUF (NET_CHUNK + NET_CHUNK / 16)

/* log of every table cell; +inf marks a cell without probability (the reference tests `> 0`) */
__global__ void k_net_logp(const double *__restrict__ probs, long long n, double *__restrict__ logp) {
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const double p = probs[i];
	logp[i] = p > 0 ? cnb_log(p) : CUDART_INF;
}

__device__ __forceinline__ int net_byte(const uint4 &v, int t) {
	const uint32_t w = t < 8 ? (t < 4 ? v.x : v.y) : (t < 12 ? v.z : v.w);
	return (w >> ((t & 3) * 8)) & 255;
}

__global__ void __launch_bounds__(128) k_net_nodes(CnbDevNet net, const uint8_t *__restrict__ codes,
		const uint32_t *__restrict__ keep_lbl, int Mpad, const int *__restrict__ sel, double *__restrict__ out) {
	__shared__ double vals[2][NET_BUF];
	__shared__ int s_count[4], s_zero[4];
	const int i = sel ? sel[blockIdx.x] : blockIdx.x;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int cc = net.cats[i], d = net.npar[i];
	const double *__restrict__ logp = net.logp + net.off[i];
	int par[CNB_MAXP], pc[CNB_MAXP];
#pragma unroll
	for (int p = 0; p < CNB_MAXP; p++) {
		par[p] = p < d ? net.pars[(size_t)i * net.maxp + p] : 0;
		pc[p] = p < d ? net.cats[par[p]] : 1;
	}
	const int nchunks = (Mpad + NET_CHUNK - 1) / NET_CHUNK;
	int ncount = 0, zero = 0;
	double acc = 0.0;
	auto produce = [&](int chunk) {
		const int k0 = (threadIdx.x - 32) * 16, j = chunk * NET_CHUNK + k0;
		double *dst = vals[chunk & 1];
		if (j >= Mpad) {
#pragma unroll
			for (int t = 0; t < 16; t++) dst[NET_PAD(k0 + t)] = 0.0;
			return;
		}
		const uint4 cv = *reinterpret_cast<const uint4 *>(codes + (size_t)i * Mpad + j);
		uint4 pv[CNB_MAXP];
#pragma unroll
		for (int p = 0; p < CNB_MAXP; p++)
			if (p < d) pv[p] = *reinterpret_cast<const uint4 *>(codes + (size_t)par[p] * Mpad + j);
		uint32_t kp = 0xFFFFu;
		if (keep_lbl) kp = (keep_lbl[(size_t)(j >> 5) * net.N + i] >> (j & 31)) & 0xFFFFu;
#pragma unroll
		for (int t = 0; t < 16; t++) {
			const int x = net_byte(cv, t);
			bool ok = x < cc && ((kp >> t) & 1);
			int idx = 0;
#pragma unroll
			for (int p = 0; p < CNB_MAXP; p++)
				if (p < d) {
					const int v = net_byte(pv[p], t);
					ok = ok && v < pc[p];
					idx = idx * pc[p] + v;
				}
			double val = 0.0;                     /* adding +0 equals skipping the sample */
			if (ok) {
				val = logp[(size_t)idx * cc + x];
				if (val == CUDART_INF) { zero = 1; val = 0.0; }
				else ncount++;
			}
			dst[NET_PAD(k0 + t)] = val;
		}
	};
	if (warp > 0) produce(0);
	__syncthreads();
	for (int c = 0; c < nchunks; c++) {
		if (warp > 0) {
			if (c + 1 < nchunks) produce(c + 1);
		} else if (lane == 0) {
			const double *src = vals[c & 1];
#pragma unroll 16
			for (int k = 0; k < NET_CHUNK; k++) acc = __dadd_rn(acc, src[NET_PAD(k)]);
		}
		__syncthreads();
	}
	/* samples counted and zero-probability hits of the producers */
	for (int o = 16; o; o >>= 1) {
		ncount += __shfl_xor_sync(0xffffffffu, ncount, o);
		zero |= __shfl_xor_sync(0xffffffffu, zero, o);
	}
	if (lane == 0) { s_count[warp] = ncount; s_zero[warp] = zero; }
	__syncthreads();
	if (threadIdx.x == 0) {
		const int n = s_count[1] + s_count[2] + s_count[3];
		double r = acc;
		if (s_zero[1] | s_zero[2] | s_zero[3]) r = -(double)FLT_MAX;
		else if (n > 1) r = __ddiv_rn(r, (double)n);
		out[blockIdx.x] = r;
	}
}

/* cell of sample j in node i's table, or -1 when the sample does not count for the node */
__device__ __forceinline__ long long net_cell(const CnbDevNet &net, const uint8_t *__restrict__ codes,
		const uint32_t *__restrict__ keep_lbl, int Mpad, int i, int j) {
	if (keep_lbl && !((keep_lbl[(size_t)(j >> 5) * net.N + i] >> (j & 31)) & 1)) return -1;
	const int cc = net.cats[i], x = codes[(size_t)i * Mpad + j];
	if (x >= cc) return -1;
	const int d = net.npar[i];
	long long idx = 0;
	for (int p = 0; p < d; p++) {
		const int q = net.pars[(size_t)i * net.maxp + p], qc = net.cats[q], v = codes[(size_t)q * Mpad + j];
		if (v >= qc) return -1;
		idx = idx * qc + v;
	}
	return net.off[i] + idx * cc + x;
}

__global__ void k_net_first_zero(CnbDevNet net, const uint8_t *__restrict__ codes, const uint32_t *__restrict__ keep_lbl,
		int M, int Mpad, int *__restrict__ first_zero) {
	const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
	if (j >= M) return;
	const long long cell = net_cell(net, codes, keep_lbl, Mpad, i, j);
	if (cell >= 0 && net.logp[cell] == CUDART_INF) atomicMin(&first_zero[i], j);
}

__global__ void k_net_bysample(CnbDevNet net, const uint8_t *__restrict__ codes, const uint32_t *__restrict__ keep_lbl,
		int M, int Mpad, const int *__restrict__ first_zero, double *__restrict__ out) {
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= M) return;
	double acc = 0.0;
	for (int i = 0; i < net.N; i++) {
		if (first_zero && j > first_zero[i]) continue;    /* the node's pass ended before this sample */
		const long long cell = net_cell(net, codes, keep_lbl, Mpad, i, j);
		if (cell < 0) continue;
		const double v = net.logp[cell];
		acc = v == CUDART_INF ? -(double)FLT_MAX : __dadd_rn(acc, v);
	}
	out[j] = acc;
}

/* samples in which node i is not perturbed (all of them without masks) */
__global__ void k_keep_count(const uint32_t *__restrict__ keep_lbl, int N, int W, int *__restrict__ out) {
	const int i = blockIdx.x;
	int n = 0;
	for (int w = threadIdx.x; w < W; w += blockDim.x) n += __popc(keep_lbl[(size_t)w * N + i]);
	for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
	if (threadIdx.x == 0) out[i] = n;
}

int cnbk_net_logp(cudaStream_t st, const double *probs, long long n, double *logp) {
	if (n <= 0) return CNB_OK;
	k_net_logp<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(probs, n, logp);
	CK(cudaGetLastError());
	return CNB_OK;
}

int cnbk_net_nodes(cudaStream_t st, const CnbDevNet &net, const uint8_t *codes, const uint32_t *keep_lbl, int Mpad,
		const int *sel, int nsel, double *out) {
	if (nsel <= 0) return CNB_OK;
	k_net_nodes<<<nsel, 128, 0, st>>>(net, codes, keep_lbl, Mpad, sel, out);
	CK(cudaGetLastError());
	return CNB_OK;
}

int cnbk_net_bysample(cudaStream_t st, const CnbDevNet &net, const uint8_t *codes, const uint32_t *keep_lbl, int M,
		int Mpad, int *first_zero, bool any_zero, double *out) {
	if (any_zero) {
		CK(cudaMemsetAsync(first_zero, 0x7f, (size_t)net.N * sizeof(int), st));     /* 0x7f7f7f7f > any sample index */
		dim3 grid((unsigned)((M + 255) / 256), (unsigned)net.N);
		k_net_first_zero<<<grid, 256, 0, st>>>(net, codes, keep_lbl, M, Mpad, first_zero);
		CK(cudaGetLastError());
	}
	k_net_bysample<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(net, codes, keep_lbl, M, Mpad, any_zero ? first_zero : nullptr, out);
	CK(cudaGetLastError());
	return CNB_OK;
}

int cnbk_keep_count(cudaStream_t st, const uint32_t *keep_lbl, int N, int W, int *out) {
	k_keep_count<<<N, 32, 0, st>>>(keep_lbl, N, W, out);
	CK(cudaGetLastError());
	return CNB_OK;
}

/* ------------------------------------------------------------------ cnSamples ----
 * RCatnet::genSamples (rcatnet.cpp:524-650): nodes in the order CATNET::getOrder gives (parents first), per node the
 * samples in order, ONE uniform per (node, sample) unless a perturbation fixes the value; the category is the first
 * i with u <= p_0 + ... + p_i (fp64, in order; i = C if none).  Then, for naRate, N uniforms per sample pick the
 * k = (int)(naRate * N) nodes that become NA (largest (int)(N u) first, lowest node on ties).
 * The library's uniform stream is splitmix64 (cnb_rng.h), which is counter based: draw t is a function of seed and t
 * alone, so every (node, sample) computes its own uniform: t = draws of the earlier nodes + unfixed samples before j. */
__device__ __forceinline__ double net_uniform(unsigned long long seed, unsigned long long t) {
	unsigned long long z = seed + (t + 1ull) * 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	z = z ^ (z >> 31);
	return __dmul_rn(__dadd_rn((double)(z >> 11), 0.5), 1.0 / 9007199254740992.0);
}

#define SMP_CHUNK 256
/* unfixed (node, sample) entries per (order position, chunk of 256 samples) */
__global__ void __launch_bounds__(SMP_CHUNK) k_sample_count(CnbDevNet net, const int *__restrict__ order,
		const int *__restrict__ pert, int M, int nchunks, int *__restrict__ cnt) {
	const int k = blockIdx.y, i = order[k], j = blockIdx.x * SMP_CHUNK + threadIdx.x;
	int free_ = 0;
	if (j < M) {
		const int v = pert[(size_t)j * net.N + i];
		free_ = !(v >= 1 && v <= net.cats[i]);
	}
	const int n = __syncthreads_count(free_);
	if (threadIdx.x == 0) cnt[(size_t)k * nchunks + blockIdx.x] = n;
}

/* one thread per sample, nodes in sampling order; probs = the network's tables (net.logp points at them here).
 * base[k * nchunks + chunk] = uniforms consumed before that chunk of order position k (only with perturbations) */
__global__ void __launch_bounds__(SMP_CHUNK) k_sample(CnbDevNet net, const int *__restrict__ order,
		const int *__restrict__ pert, const long long *__restrict__ base, int M, int nchunks, unsigned long long seed,
		int *__restrict__ out) {
	__shared__ int warp_tot[SMP_CHUNK / 32];
	const int j = blockIdx.x * SMP_CHUNK + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const bool live = j < M;
	int *row = out + (size_t)(live ? j : 0) * net.N;
	for (int k = 0; k < net.N; k++) {
		const int i = order[k], cc = net.cats[i];
		int fixed = 0;
		unsigned long long t = (unsigned long long)k * (unsigned long long)M + (unsigned long long)j;
		if (pert) {
			if (live) {
				const int v = pert[(size_t)j * net.N + i];
				if (v >= 1 && v <= cc) fixed = v;                  /* rcatnet.cpp:578-584 */
			}
			/* rank of this sample among the chunk's unfixed ones */
			const unsigned m = __ballot_sync(0xffffffffu, live && !fixed);
			if (lane == 0) warp_tot[warp] = __popc(m);
			__syncthreads();
			int before = __popc(m & ((1u << lane) - 1));
			for (int w = 0; w < warp; w++) before += warp_tot[w];
			__syncthreads();
			t = (unsigned long long)(base[(size_t)k * nchunks + blockIdx.x] + before);
		}
		if (!live) continue;
		if (fixed) { row[i] = fixed; continue; }
		const int d = net.npar[i];
		long long idx = 0;
		for (int p = 0; p < d; p++) {
			const int q = net.pars[(size_t)i * net.maxp + p], qc = net.cats[q];
			int v = row[q] - 1;                                    /* drawn earlier by this thread */
			v = v < 0 ? 0 : (v >= qc ? qc - 1 : v);                /* the reference dereferences NULL there */
			idx = idx * qc + v;
		}
		const double *p = net.logp + net.off[i] + idx * cc;
		const double u = net_uniform(seed, t);
		double v = 0.0;
		int c = 0;
		for (; c < cc; c++) {
			v = __dadd_rn(v, p[c]);
			if (u <= v) break;
		}
		row[i] = c + 1;
	}
}

/* naRate: k nodes per sample become NA (rcatnet.cpp:601-626).  paux[ii] = (int)(N u) from N uniforms per sample; the
 * reference takes the first maximum k times, i.e. the k first elements by (value descending, node ascending). */
__global__ void k_sample_na(int N, int M, int k, unsigned long long seed, unsigned long long base, int *__restrict__ out) {
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= M) return;
	int pv = N + 1, pi = -1;                                       /* the previous pick (value, node) */
	for (int r = 0; r < k; r++) {
		int bv = -1, bi = -1;
		for (int ii = 0; ii < N; ii++) {
			const int a = (int)__dmul_rn((double)N, net_uniform(seed, base + (unsigned long long)j * N + ii));
			const bool after = a < pv || (a == pv && ii > pi);      /* not picked yet */
			if (after && a > bv) { bv = a; bi = ii; }
		}
		if (bi < 0) break;
		out[(size_t)j * N + bi] = INT_MIN;                         /* R_NaInt */
		pv = bv;
		pi = bi;
	}
}

int cnbk_sample_chunk(void) { return SMP_CHUNK; }

int cnbk_sample_count(cudaStream_t st, const CnbDevNet &net, const int *order, const int *pert, int M, int *cnt) {
	const int nchunks = (M + SMP_CHUNK - 1) / SMP_CHUNK;
	dim3 grid(nchunks, net.N);
	k_sample_count<<<grid, SMP_CHUNK, 0, st>>>(net, order, pert, M, nchunks, cnt);
	CK(cudaGetLastError());
	return CNB_OK;
}

int cnbk_sample(cudaStream_t st, const CnbDevNet &net, const int *order, const int *pert, const long long *base, int M,
		unsigned long long seed, int *out) {
	const int nchunks = (M + SMP_CHUNK - 1) / SMP_CHUNK;
	k_sample<<<nchunks, SMP_CHUNK, 0, st>>>(net, order, pert, base, M, nchunks, seed, out);
	CK(cudaGetLastError());
	return CNB_OK;
}

int cnbk_sample_na(cudaStream_t st, int N, int M, int k, unsigned long long seed, unsigned long long base, int *out) {
	k_sample_na<<<(M + 127) / 128, 128, 0, st>>>(N, M, k, seed, base, out);
	CK(cudaGetLastError());
	return CNB_OK;
}
