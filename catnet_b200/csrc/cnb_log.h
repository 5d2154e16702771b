/* cnb_log.h -- fp64 natural log that is BIT-IDENTICAL to glibc (>= 2.28) log() on FMA-capable x86-64.
 *
 * The reference scores a family with  loglik += n * log(n * (1/N))  using the host libm (problist.h:244), and both
 * the arg-max over candidate parent sets (catnet_search2.h:613, strict '<', first wins) and the DP over complexity
 * (:665-671) compare those fp64 sums exactly.  A device log that is merely "1 ulp" flips mathematically tied
 * candidates; so the device (and the host side of this library) evaluates the very algorithm glibc runs:
 * sysdeps/ieee754/dbl-64/e_log.c (ARM optimized-routines log, N = 128 table, order-5 polynomial, order-11 near 1),
 * in the operation order and with the FMA contractions of the ifunc-selected FMA build, using glibc's own table
 * (cnb_log_data.h, extracted by tools/extract_libm_log.py).  Every operation below is a single IEEE-754 fp64
 * add/mul/fma, so CUDA and x86 agree bit for bit.  tests/test_abi.py::test_host_log_is_libm_log and tests/test_gpu_parity.py::test_device_log_is_glibc_log check it against libm.
 *
 * Origin and licence: the ALGORITHM is that of glibc's sysdeps/ieee754/dbl-64/e_log.c, which is ARM's optimized-routines
 * log (Szabolcs Nagy; optimized-routines is MIT / Apache-2.0 WITH LLVM-exception, glibc ships it under the LGPL-2.1+); this
 * file is an independent restatement of that published algorithm, not a copy of either source file, and the 128-entry table
 * and the polynomial coefficients in cnb_log_data.h are numeric constants read out of this image's libm.so.6 by
 * tools/extract_libm_log.py.  A redistributor who treats the table as LGPL-covered data can regenerate cnb_log_data.h from
 * the libm of the target system with that script (it is the target's libm the results must match anyway).
 *
 * Domain: finite x > 0 (normal or subnormal).  The hot path only ever passes n/N in (0, 1] and positive priors.
 */
#ifndef CNB_LOG_H
#define CNB_LOG_H
#include <stdint.h>
#include <string.h>
#include "cnb_log_data.h"

#ifdef __CUDACC__
#define CNB_HD __host__ __device__ __forceinline__
#else
#define CNB_HD static inline
#endif
#include <math.h>

static const double cnb_log_A[5] = CNB_LOG_A;
static const double cnb_log_B[11] = CNB_LOG_B;
static const double cnb_log_T[256] = CNB_LOG_TAB;
#ifdef __CUDACC__
/* device copies: declared in both compilation passes (never guard a __device__ variable by __CUDA_ARCH__) */
__device__ static const double cnb_dlog_A[5] = CNB_LOG_A;
__device__ static const double cnb_dlog_B[11] = CNB_LOG_B;
__device__ static const double cnb_dlog_T[256] = CNB_LOG_TAB;
#endif
static inline uint64_t cnb_asu64_(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double cnb_asf64_(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }

#ifdef __CUDA_ARCH__
#define CNB_TA cnb_dlog_A
#define CNB_TB cnb_dlog_B
#define CNB_TT cnb_dlog_T
#define CNB_FMA(a, b, c) __fma_rn((a), (b), (c))
#define CNB_MUL(a, b) __dmul_rn((a), (b))
#define CNB_ADD(a, b) __dadd_rn((a), (b))
#define CNB_SUB(a, b) __dsub_rn((a), (b))
#define CNB_ASU64(x) ((uint64_t)__double_as_longlong(x))
#define CNB_ASF64(u) __longlong_as_double((long long)(u))
#else
#define CNB_TA cnb_log_A
#define CNB_TB cnb_log_B
#define CNB_TT cnb_log_T
#define CNB_FMA(a, b, c) fma((a), (b), (c))
/* host: plain operators; host code is built with -ffp-contract=off */
#define CNB_MUL(a, b) ((a) * (b))
#define CNB_ADD(a, b) ((a) + (b))
#define CNB_SUB(a, b) ((a) - (b))
#define CNB_ASU64(x) cnb_asu64_(x)
#define CNB_ASF64(u) cnb_asf64_(u)
#endif

/* the table path of the algorithm for the bits ix of a NORMAL x outside [1 - 2^-4, 1 + 0x1.09p-4): branch-free, so
 * several of them can be interleaved (cnb_umma.cu evaluates the four terms of a parent configuration side by side) */
CNB_HD double cnb_log_far(uint64_t ix) {
	const double *A = CNB_TA;
	uint64_t tmp = ix - 0x3fe6000000000000ull;
	int i = (int)((tmp >> 45) & 127);
	int k = (int)((int64_t)tmp >> 52);
	uint64_t iz = ix - (tmp & (0xfffull << 52));
	double invc = CNB_TT[2 * i], logc = CNB_TT[2 * i + 1];
	double z = CNB_ASF64(iz);
	double kd = (double)k;
	double w = CNB_FMA(kd, CNB_LN2HI, logc);
	double r = CNB_FMA(z, invc, -1.0);
	double p1 = CNB_FMA(r, A[2], A[1]);
	double hi = CNB_ADD(r, w);
	double r2 = CNB_MUL(r, r);
	double lo = CNB_FMA(kd, CNB_LN2LO, CNB_ADD(CNB_SUB(w, hi), r));
	double r3 = CNB_MUL(r, r2);
	double q = CNB_FMA(r, A[4], A[3]);
	lo = CNB_FMA(r2, A[0], lo);
	q = CNB_FMA(q, r2, p1);
	return CNB_ADD(CNB_FMA(r3, q, lo), hi);
}
/* true if cnb_log(x) takes the table path without the subnormal rescaling, i.e. equals cnb_log_far(bits of x) */
CNB_HD bool cnb_log_is_far(uint64_t ix) {
	return !(ix - 0x3fee000000000000ull < 0x0003090000000000ull) && !((uint32_t)(ix >> 48) - 0x0010u >= 0x7ff0u - 0x0010u);
}

CNB_HD double cnb_log(double x) {
	const double *B = CNB_TB;
	uint64_t ix = CNB_ASU64(x);
	/* |x - 1| small: 1 - 2^-4 <= x < 1 + 0x1.09p-4 */
	if (ix - 0x3fee000000000000ull < 0x0003090000000000ull) {
		if (ix == 0x3ff0000000000000ull) return 0.0;
		double r = CNB_SUB(x, 1.0);
		double r2 = CNB_MUL(r, r);
		double r3 = CNB_MUL(r, r2);
		double p0 = CNB_FMA(r2, B[3], CNB_FMA(r, B[2], B[1]));
		double p1 = CNB_FMA(r2, B[6], CNB_FMA(r, B[5], B[4]));
		double p2 = CNB_FMA(r3, B[10], CNB_FMA(r2, B[9], CNB_FMA(r, B[8], B[7])));
		double p = CNB_FMA(CNB_FMA(p2, r3, p1), r3, p0);
		double t = CNB_FMA(r, 0x1p27, r);
		double rhi = CNB_FMA(-0x1p27, r, t);
		double rlo = CNB_SUB(r, rhi);
		double rhi2 = CNB_MUL(rhi, rhi);
		double hi = CNB_FMA(rhi2, B[0], r);
		double lo = CNB_FMA(rhi2, B[0], CNB_SUB(r, hi));
		lo = CNB_FMA(CNB_MUL(B[0], rlo), CNB_ADD(r, rhi), lo);
		return CNB_ADD(hi, CNB_FMA(p, r3, lo));
	}
	if ((uint32_t)(ix >> 48) - 0x0010u >= 0x7ff0u - 0x0010u) {
		/* subnormal (x > 0 assumed): scale by 2^52 */
		ix = CNB_ASU64(CNB_MUL(x, 0x1p52));
		ix -= 52ull << 52;
	}
	return cnb_log_far(ix);
}
#endif
