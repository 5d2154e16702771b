/* cnb_pairwise.cu -- the pairwise metrics of catnet_entropy.cpp from two-way count tables.
 *
 * catnetEntropyPairwise (:38-242), catnetEntropyOrder (:244-468), catnetKLpairwise (:470-691) and
 * catnetPearsonPairwise (:693-911) each loop over ordered node pairs (n1, n2), count a two-way table over the samples
 * and reduce it in fp64.  Here the tables come from the family scorers of the search path (family "child n1, parent
 * n2": table[i2 * C_n1 + i1]) -- `kept` over the samples where n1 is not perturbed, `full` over all samples -- and
 * this file holds the reductions: one thread per ordered pair, the reference's operation order, glibc's log
 * (cnb_log.h), no contraction (every fp64 operation is an explicit round-to-nearest intrinsic). */
#include <cuda_runtime.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "cnb_kernels.h"
#include "cnb_log.h"

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cnb_set_error("CUDA: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); return CNB_ERR_CUDA; } } while (0)

/* table of the explicit family (child n1, parent n2) in the list the API builds: n1 * (N - 1) + rank of n2 among the
 * other nodes */
__device__ __forceinline__ size_t pw_family(int N, int n1, int n2) {
	return (size_t)n1 * (N - 1) + (n2 < n1 ? n2 : n2 - 1);
}

/* conditional entropy H(n1 | n2), catnet_entropy.cpp:183-203; cnt(i2, i1) fetches a cell */
template <class Fetch>
__device__ __forceinline__ double pw_cond_entropy(int c2, int c1, Fetch cnt) {
	double floglik = 0.0, fsum = 0.0;
	for (int i = 0; i < c2; i++) {
		double faux = 0.0;
		for (int j = 0; j < c1; j++) {
			const int n = cnt(i, j);
			faux = __dadd_rn(faux, (double)n);
			if (n > 0) floglik = __dadd_rn(floglik, __dmul_rn((double)n, cnb_log((double)n)));
		}
		fsum = __dadd_rn(fsum, faux);
		if (faux > 0) floglik = __dsub_rn(floglik, __dmul_rn(faux, cnb_log(faux)));
	}
	if (fsum > 0) floglik = __ddiv_rn(floglik, fsum);
	return -floglik;
}

/* kind 0: entropy.  diag_tab (only for N == 1) holds the no-parent tables [N][stride]; otherwise the marginal of n1
 * over its kept samples is a column sum of any of its kept tables (no value is missing). */
__global__ void __launch_bounds__(128) k_pairwise_metric(int kind, int N, const int *__restrict__ cats,
		const int *__restrict__ kept, const int *__restrict__ full, const int *__restrict__ diag_tab, int stride,
		double *__restrict__ out) {
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= (long long)N * N) return;
	/* consecutive threads = consecutive n2 for one n1, like the family list */
	const int n1 = (int)(t / N), n2 = (int)(t % N);
	const int c1 = cats[n1], c2 = cats[n2];
	double *dst = out + (size_t)n2 * N + n1;            /* klmat[nnode2*numnodes + nnode1] */
	if (n1 == n2) {
		if (kind != 0) { *dst = 0.0; return; }
		/* marginal entropy, :163-181 */
		const int other = n1 == 0 ? 1 : 0;
		const int co = N > 1 ? cats[other] : 1;
		const int *tab = N > 1 ? kept + pw_family(N, n1, other) * stride : diag_tab + (size_t)n1 * stride;
		double floglik = 0.0, fsum = 0.0;
		for (int j = 0; j < c1; j++) {
			int n = 0;
			for (int i = 0; i < co; i++) n += tab[i * c1 + j];
			fsum = __dadd_rn(fsum, (double)n);
			if (n > 0) floglik = __dadd_rn(floglik, __dmul_rn((double)n, cnb_log((double)n)));
		}
		if (fsum > 0) {
			floglik = __dsub_rn(floglik, __dmul_rn(fsum, cnb_log(fsum)));
			floglik = __ddiv_rn(floglik, fsum);
		}
		*dst = -floglik;
		return;
	}
	const int *kt = kept + pw_family(N, n1, n2) * stride;
	if (kind == 0) {
		*dst = pw_cond_entropy(c2, c1, [&](int i, int j) { return kt[i * c1 + j]; });
		return;
	}
	const int *ft = full + pw_family(N, n1, n2) * stride;
	double floglik = 0.0;
	for (int i = 0; i < c2; i++) {
		/* row sums are sums of exact integers: any order gives the reference's doubles (:613-634, :836-846) */
		long long s1 = 0, s2 = 0;
		for (int j = 0; j < c1; j++) { s1 += kt[i * c1 + j]; s2 += ft[i * c1 + j]; }
		const double f2 = s2 > 0 ? __ddiv_rn(1.0, (double)s2) : 1.0;
		if (kind == 1) {                                 /* :636-647 */
			const double f1 = s1 > 0 ? __ddiv_rn(1.0, (double)s1) : 1.0;
			for (int j = 0; j < c1; j++) {
				const double a = s1 > 0 ? __dmul_rn((double)kt[i * c1 + j], f1) : (double)kt[i * c1 + j];
				const double b = s2 > 0 ? __dmul_rn((double)ft[i * c1 + j], f2) : (double)ft[i * c1 + j];
				if (a > 0 && b > 0) floglik = __dadd_rn(floglik, __dmul_rn(a, cnb_log(__ddiv_rn(a, b))));
				else if (a != 0 && b == 0) floglik = (double)FLT_MAX;
			}
		} else {                                         /* :848-864 */
			if (s1 <= 0) continue;
			const double fsum = (double)s1;
			for (int j = 0; j < c1; j++) {
				const double b = s2 > 0 ? __dmul_rn((double)ft[i * c1 + j], f2) : (double)ft[i * c1 + j];
				const double faux = __dsub_rn((double)kt[i * c1 + j], __dmul_rn(fsum, b));
				if (b > 0) floglik = __dadd_rn(floglik, __ddiv_rn(__dmul_rn(faux, faux), __dmul_rn(fsum, b)));
				else if (faux != 0 && b == 0) floglik = (double)FLT_MAX;
			}
		}
	}
	*dst = __dadd_rn(0.0, floglik);                      /* klmat[...] += floglik on a zeroed matrix */
}

int cnbk_pairwise_metric(cudaStream_t st, int kind, int N, const int *cats, const int *kept, const int *full,
		const int *diag_tab, int stride, double *out) {
	const long long n = (long long)N * N;
	k_pairwise_metric<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(kind, N, cats, kept, full, diag_tab, stride, out);
	CK(cudaGetLastError());
	return CNB_OK;
}

/* catnetEntropyOrder's tail on the host (:428-440 and _order, utils.h:97-133, decreasing): mat is the entropy matrix
 * klmat[n2*N + n1]; rank of n1 = sum over n2 of H(n1) - H(n1 | n2); nodes by decreasing rank, 1-based.  The
 * reference's _order marks every element equal to the current value as used and keeps the LAST of them, so a run of t
 * tied ranks fills one position and leaves t - 1 positions at their initial -1 (0 after the +1): reproduced as is,
 * with position 0 (never initialised there, always written for N >= 2) starting from 0. */
void cnb_entropy_order_from_matrix(const double *mat, int N, int32_t *order_out) {
	std::vector<double> ranks(N), sorted, aux;
	for (int n1 = 0; n1 < N; n1++) {
		const double faux = mat[(size_t)n1 * N + n1];
		double s = 0;
		for (int n2 = 0; n2 < N; n2++) s += faux - mat[(size_t)n2 * N + n1];
		ranks[n1] = s;
	}
	for (int i = 0; i < N; i++) order_out[i] = 0;
	if (N >= 2) {
		sorted = ranks;
		aux = ranks;
		std::sort(sorted.begin(), sorted.end());
		for (int j = 1; j < N; j++) order_out[j] = -1;
		for (int j = N - 1; j >= 0; j--)
			for (int i = 0; i < N; i++)
				if (aux[i] == sorted[j]) {
					order_out[N - 1 - j] = i;
					aux[i] = (double)FLT_MAX;
				}
	}
	for (int i = 0; i < N; i++) order_out[i]++;
}

/* test hook (no GPU): the host tail of cnb_entropy_order on a given entropy matrix mat[n2 * N + n1] */
extern "C" void cnb_host_entropy_order(const double *mat, int32_t N, int32_t *order_out) {
	cnb_entropy_order_from_matrix(mat, N, order_out);
}
