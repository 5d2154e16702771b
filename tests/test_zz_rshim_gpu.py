"""GPU suite, last file: the R shim (catnet_b200/csrc/rshim.cpp) linked against the REAL libcatnet_b200.so and driven
through the miniature R API of tests/rstub/minir.cpp -- `.Call("ccnOptimalNetsForOrder", ...)` with R-style arguments
in, S4 catNetwork objects out, the pairwise metrics and the given-network routines, compared with the oracle.  (tests/test_rshim.py runs the same shim on the CPU with the
C ABI answered by the oracle.)"""
import numpy as np
import pytest

from helpers import constrained_problem, random_problem
from test_rshim import check_given_network_calls, check_networks, check_sa_and_histogram

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def r():
    from rshim_harness import R
    return R("real")


@pytest.mark.parametrize("seed", range(700, 712))
def test_dot_call_optimal_nets_for_order(r, seed):
    from oracle import port
    from rshim_harness import optimal_nets_for_order
    s, P, K, kw = constrained_problem(seed)
    nets = optimal_nets_for_order(r, s, P, K, real_args=bool(seed % 2), **kw)
    theirs = port.search_order(s, P, K, **kw)
    check_networks(nets, theirs, kw["order"], kw["num_cats"])


def test_dot_call_pairwise(r):
    from oracle import port
    s, c = random_problem(11, 7, 120, None)
    pert = (np.random.default_rng(1).random(s.shape) < 0.2).astype(np.int32)
    for kind, routine in [("entropy", "ccnEntropyPairwise"), ("kl", "ccnKLPairwise"), ("pearson", "ccnPearsonPairwise")]:
        for p in (None, pert):
            out = r.call(routine, r.data_matrix(s, real=True), r.NULL if p is None else r.data_matrix(p))
            assert np.array_equal(r.reals(out).reshape(7, 7).T, port.pairwise(kind, s, p))
    assert r.ints(r.call("ccnEntropyOrder", r.data_matrix(s), r.NULL)) == [int(v) for v in port.pairwise("order", s)]


@pytest.mark.parametrize("seed", range(800, 808))
def test_dot_call_given_network(r, seed):
    """ccnSetProb -> ccnLoglik -> ccnNodeLoglik -> ccnSamples over the library, against the oracle"""
    from oracle import port
    check_given_network_calls(r, port, seed)


@pytest.mark.parametrize("seed", range(900, 903))
def test_dot_call_sa_and_histogram(r, seed):
    """ccnOptimalNetsSA and ccnParHistogram over the library: the chain and the histogram of the oracle under the seed
    the shim draws from (the miniature) R's RNG"""
    from oracle import port
    check_sa_and_histogram(r, port, seed)
