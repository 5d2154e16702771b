"""GPU suite, last file (tests written after the round's GPU budget was spent, so their first run is the round-end
suite; they sort last so that they cannot mask the others under -x): the R shim (catnet_b200/csrc/rshim.cpp) linked against the REAL libcatnet_b200.so and driven
through the miniature R API of tests/rstub/minir.cpp -- `.Call("ccnOptimalNetsForOrder", ...)` with R-style arguments
in, S4 catNetwork objects out, the pairwise metrics and the given-network routines, compared with the oracle.  (tests/test_rshim.py runs the same shim on the CPU with the
C ABI answered by the oracle.)"""
import numpy as np
import pytest

from helpers import compare_search, constrained_problem, fully_perturbed_problem, random_problem
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


@pytest.mark.parametrize("seed", range(1100, 1112))
def test_node_perturbed_in_every_sample(seed):
    """a node without a single unperturbed sample never gets parents (tests/test_oracle.py pins the restatement on
    this against the compiled reference); one-shot call and resident context"""
    from catnet_b200 import _abi
    from oracle import port
    s, P, K, kw = fully_perturbed_problem(seed)
    theirs = port.search_order(s, P, K, **kw)
    ours = _abi.search_order(s, max_parent_set=P, max_complexity=K, **kw)
    assert len(ours) == len(theirs) and compare_search(ours, theirs) == len(theirs)
    with _abi.Context(0) as ctx:
        ctx.set_samples(s, perturbations=kw["perturbations"], num_cats=kw["num_cats"])
        kw2 = {k: v for k, v in kw.items() if k not in ("perturbations", "num_cats")}
        again = ctx.search_order(max_parent_set=P, max_complexity=K, **kw2)
    assert compare_search(again, theirs) == len(theirs)
