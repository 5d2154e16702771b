/* cnb_rng.h -- the uniform stream of the SA driver.
 * R's unif_rand (used by rcatnet_sa.cpp:146,151,225,745 and utils.h:156,190) is not reachable through a C ABI,
 * so the library owns a documented generator: splitmix64 mapped to (0,1) as ((x>>11)+0.5)*2^-53.
 * oracle/ref_driver.cpp injects the identical stream into the reference, so SA trajectories compare draw for draw. */
#ifndef CNB_RNG_H
#define CNB_RNG_H
#include <stdint.h>
struct CnbRng {
	uint64_t state;
	long calls;
	explicit CnbRng(uint64_t seed) : state(seed), calls(0) {}
	double unif() {
		uint64_t z = (state += 0x9E3779B97F4A7C15ull);
		z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
		z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
		z = z ^ (z >> 31);
		calls++;
		return ((double)(z >> 11) + 0.5) * (1.0 / 9007199254740992.0);
	}
};
#endif
