/* umma_peak.cu -- issue-rate floor of the tcgen05.mma shapes the table kernel uses (measurement tool, not product).
 *
 *   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/umma_peak tools/umma_peak.cu && tools/umma_peak
 *
 * One CTA per SM; one thread issues ITER back-to-back MMAs (M = 128, N swept, K = 32 bytes) on whatever bytes lie in
 * shared / tensor memory, commits once, and the CTA reports cycles per MMA.  Variants: kind::mxf4.block_scale with A
 * from shared memory (SS) or tensor memory (TS), kind::i8 SS/TS.  Two accumulators are alternated like the table
 * kernel does.  Prints MAC/clk/SM and the chip-wide TFLOP/s at the measured clock.
 */
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t a) {
	return (uint64_t)((a >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
	uint32_t ok = 0, addr = smem_u32(bar);
	while (!ok)
		asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
				: "=r"(ok) : "r"(addr), "r"(parity) : "memory");
}

/* variant: 0 mxf4 SS, 1 mxf4 TS, 2 i8 SS, 3 i8 TS */
__global__ void __launch_bounds__(128, 1) k_peak(int variant, int N, int iters, int nacc, long long *cycles) {
	extern __shared__ uint8_t raw[];
	__shared__ __align__(8) uint64_t bar;
	__shared__ uint32_t tbase;
	const uint32_t smem = (smem_u32(raw) + 1023u) & ~1023u;
	const int warp = threadIdx.x >> 5;
	for (int i = threadIdx.x; i < (128 + 256) * 128 / 4; i += blockDim.x) ((uint32_t *)(raw + (smem - smem_u32(raw))))[i] = 0x22222222u;
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "n"(512));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
	}
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
		asm volatile("fence.mbarrier_init.release.cluster;");
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;");
	const uint32_t tmem = tbase;
	{
		const uint32_t one = 0x7F7F7F7Fu;       /* unit scales + some A data in tensor memory */
		for (int c = 384; c < 512; c += 16)
			asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};"
					::"r"(tmem + ((uint32_t)(warp * 32) << 16) + c), "r"(one) : "memory");
		asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;");
	if (threadIdx.x == 0) {
		const bool fp4 = variant < 2, ts = variant & 1;
		const uint32_t idesc = fp4 ? ((1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (1u << 23) | (8u << 24))
		                           : ((2u << 4) | ((uint32_t)(N >> 3) << 17) | (8u << 24));
		const uint64_t da = umma_desc(smem), db = umma_desc(smem + 128 * 128);
		const uint32_t ta = tmem + 384, sfa = tmem + 480, sfb = tmem + 496;
		long long t0 = clock64();
		for (int i = 0; i < iters; i++) {
			const uint32_t d = tmem + (nacc > 1 ? (i & 1) * 176 : 0);
			const uint32_t acc = i >= nacc;
			const uint64_t a = da + 2 * (i & 3), b = db + 2 * (i & 3);
			const uint32_t at = ta + 8 * (i & 3);
			uint32_t z = 0;
			if (fp4 && !ts)
				asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n}\n"
						::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(sfa), "r"(sfb) : "memory");
			else if (fp4)
				asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], [%1], %2, %3, [%5], [%6], p;\n}\n"
						::"r"(d), "r"(at), "l"(b), "r"(idesc), "r"(acc), "r"(sfa), "r"(sfb) : "memory");
			else if (!ts)
				asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n"
						::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(z) : "memory");
			else
				asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n}\n"
						::"r"(d), "r"(at), "l"(b), "r"(idesc), "r"(acc), "r"(z) : "memory");
		}
		asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
		mbar_wait(&bar, 0);
		long long t1 = clock64();
		cycles[blockIdx.x] = t1 - t0;
	}
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

/* --quick: only the two shapes bench.py quotes, as one JSON line */
static double run_one(int sms, int smem, long long *cyc, int v, int N, int nacc, int iters, double *clk_per_mma) {
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	k_peak<<<sms, 128, smem>>>(v, N, 64, nacc, cyc);
	cudaEventRecord(e0);
	k_peak<<<sms, 128, smem>>>(v, N, iters, nacc, cyc);
	cudaEventRecord(e1);
	if (cudaDeviceSynchronize() != cudaSuccess) return -1.0;
	float ms;
	cudaEventElapsedTime(&ms, e0, e1);
	long long h[1024];
	cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
	double avg = 0;
	for (int i = 0; i < sms; i++) avg += (double)h[i];
	*clk_per_mma = avg / sms / iters;
	const double kel = v < 2 ? 64.0 : 32.0;
	return 2.0 * 128.0 * N * kel * iters * sms / (ms * 1e-3) / 1e12;
}

int main(int argc, char **argv) {
	if (argc > 1 && argv[1][0] == '-' && argv[1][1] == '-' && argv[1][2] == 'q') {
		int sms = 0;
		cudaSetDevice(0);
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
		long long *cyc;
		cudaMalloc(&cyc, sms * sizeof(long long));
		const int smem = (128 + 256) * 128 + 1024;
		cudaFuncSetAttribute(k_peak, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
		double c_ss, c_ts;
		const double ss = run_one(sms, smem, cyc, 0, 256, 1, 16384, &c_ss), ts = run_one(sms, smem, cyc, 1, 192, 2, 16384, &c_ts);
		if (ss < 0 || ts < 0) { printf("{\"error\": \"cuda\"}\n"); return 1; }
		printf("{\"mxf4_ss_n256_tflops\": %.1f, \"mxf4_ss_n256_clk_per_mma\": %.2f, \"mxf4_ts_n192_tflops\": %.1f, "
				"\"mxf4_ts_n192_clk_per_mma\": %.2f, \"sms\": %d}\n", ss, c_ss, ts, c_ts, sms);
		return 0;
	}
	int dev = 0, sms = 0, khz = 0;
	cudaSetDevice(dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
	long long *cyc;
	cudaMalloc(&cyc, sms * sizeof(long long));
	const int smem = (128 + 256) * 128 + 1024;
	cudaFuncSetAttribute(k_peak, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	const char *names[4] = {"mxf4 SS", "mxf4 TS", "i8 SS", "i8 TS"};
	const int ns[] = {16, 32, 48, 64, 96, 112, 128, 144, 160, 176, 192, 208, 224, 240, 256};
	const int iters = 8192;
	printf("SMs %d, clock attr %d kHz; cycles per MMA (M=128, K=32 B), all SMs busy\n", sms, khz);
	for (int v = 0; v < 4; v++)
		for (int nacc = 1; nacc <= 2; nacc++)
			for (int N : ns) {
				if (nacc == 2 && N > 176) continue;
				cudaEvent_t e0, e1;
				cudaEventCreate(&e0); cudaEventCreate(&e1);
				k_peak<<<sms, 128, smem>>>(v, N, 64, nacc, cyc);       /* warm-up */
				cudaEventRecord(e0);
				k_peak<<<sms, 128, smem>>>(v, N, iters, nacc, cyc);
				cudaEventRecord(e1);
				cudaError_t err = cudaDeviceSynchronize();
				if (err != cudaSuccess) { printf("%s N=%d: %s\n", names[v], N, cudaGetErrorString(err)); return 1; }
				float ms;
				cudaEventElapsedTime(&ms, e0, e1);
				long long h[1024];
				cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
				double avg = 0;
				for (int i = 0; i < sms; i++) avg += (double)h[i];
				avg /= sms;
				const double kel = v < 2 ? 64.0 : 32.0;
				const double cpm = avg / iters, macs = 128.0 * N * kel;
				printf("%-8s acc=%d N=%3d  %7.1f clk/MMA  %8.0f MAC/clk/SM  %7.1f TFLOP/s (event time %.3f ms)\n", names[v], nacc, N, cpm,
						macs / cpm, 2.0 * macs * iters * sms / (ms * 1e-3) / 1e12, ms);
			}
	return 0;
}
