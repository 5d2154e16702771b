/* cnb_rng.h -- the uniform stream of the SA / histogram drivers.
 * The reference draws every proposal from R's unif_rand (rcatnet_sa.cpp:146,151,225,745, rcatnet_hist.cpp:124,129,
 * utils.h:156,190).  A caller that owns such a generator hands it over as a function pointer (cnb_sa_params.unif /
 * cnb_hist_params.unif; the R shim passes unif_rand, so set.seed() reproduces stock catnet's proposals): the drivers run
 * on the calling thread and consume it draw for draw in the reference's order.  Without one the library owns a
 * documented generator: splitmix64 mapped to (0,1) as ((x>>11)+0.5)*2^-53; oracle/ref_driver.cpp injects the identical
 * stream into the reference, so SA trajectories compare draw for draw. */
#ifndef CNB_RNG_H
#define CNB_RNG_H
#include <stdint.h>
struct CnbRng {
	uint64_t state;
	long calls;
	double (*fn)(void *);
	void *fn_ctx;
	explicit CnbRng(uint64_t seed, double (*f)(void *) = nullptr, void *ctx = nullptr) : state(seed), calls(0), fn(f), fn_ctx(ctx) {}
	double unif() {
		calls++;
		if (fn) return fn(fn_ctx);
		uint64_t z = (state += 0x9E3779B97F4A7C15ull);
		z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
		z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
		z = z ^ (z >> 31);
		return ((double)(z >> 11) + 0.5) * (1.0 / 9007199254740992.0);
	}
};
#endif
