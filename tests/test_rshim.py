"""CPU suite, part 4: the R-facing shim (catnet_b200/csrc/rshim.cpp).  The image has no R, so the shim is
(1) syntax-checked against stand-ins of R's headers (tests/rstub), (2) its .Call table is compared with the reference's
registration (names and arity, src/ccnInit.c:29-45), and (3) it is RUN against a miniature implementation of the R API
(tests/rstub/minir.cpp) with the C ABI answered by the oracle (tests/rstub/mock_abi.cpp): the S4 catNetwork objects it
builds from R-style arguments must be the oracle's networks.  tests/test_zz_rshim_gpu.py runs the same shim over the
real library on a B200."""
import os
import re
import subprocess

import numpy as np
import pytest

from helpers import constrained_problem, random_problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "catnet_b200", "csrc", "rshim.cpp")


def test_rshim_syntax():
    r = subprocess.run(["g++", "-std=c++14", "-fsyntax-only", "-Wall", "-I" + os.path.join(ROOT, "tests", "rstub"),
                        "-DCNB_STANDALONE_RSHIM", SHIM], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_call_table_matches_reference_registration():
    src = open(SHIM).read()
    table = dict((m.group(1), int(m.group(2))) for m in re.finditer(r'\{"(ccn\w+)",\s*\(DL_FUNC\)&\w+,\s*(\d+)\}', src))
    # names and arities the reference registers for this path (src/ccnInit.c:30-32, 36-39, 41, 43)
    assert table == {"ccnOptimalNetsForOrder": 12, "ccnOptimalNetsSA": 23, "ccnParHistogram": 14, "ccnReleaseCache": 0,
                     "ccnSetSeed": 1, "ccnEntropyPairwise": 2, "ccnEntropyOrder": 2, "ccnKLPairwise": 2,
                     "ccnPearsonPairwise": 2, "ccnLoglik": 4, "ccnNodeLoglik": 4, "ccnSetProb": 3, "ccnSamples": 4}
    for name, arity in table.items():
        sig = re.search(r"SEXP %s\(([^)]*)\)" % name, src).group(1)
        nargs = 0 if sig.strip() == "void" else sig.count("SEXP")
        assert nargs == arity, (name, nargs, arity)


def check_networks(nets, theirs, order, cats, names=None):
    """nets: catNetwork dicts read back from the shim; theirs: oracle.port.search_order on the same problem"""
    assert [n["complexity"] for n in nets] == [t["complexity"] for t in theirs]
    cats_pos = [int(cats[o - 1]) for o in order]
    for a, b in zip(nets, theirs):
        assert a["parents"] == b["parents"] and a["loglik"] == b["loglik"] and a["sample_sizes"] == b["sample_sizes"]
        assert a["nodes"] == [names[o - 1] if names else "N%d" % o for o in order]
        assert [len(c) for c in a["categories"]] == cats_pos and a["maxCategories"] == max(cats_pos)
        assert a["maxParents"] == max(len(p) for p in a["parents"])
        for x, y in zip(a["probs"], b["probs"]):
            assert np.array_equal(x, y)


@pytest.mark.parametrize("seed", range(700, 724))
def test_shim_runs_against_mini_r(seed):
    from oracle import port
    from rshim_harness import R, optimal_nets_for_order
    s, P, K, kw = constrained_problem(seed)
    r = R("mock")
    nets = optimal_nets_for_order(r, s, P, K, real_args=bool(seed % 2), **kw)
    theirs = port.search_order(s, P, K, **kw)
    assert len(theirs) > 0
    check_networks(nets, theirs, kw["order"], kw["num_cats"])
    assert r.warnings() == ""


def test_shim_argument_handling():
    from oracle import port
    from rshim_harness import R, RError, optimal_nets_for_order
    s, c = random_problem(3, 6, 80, 3)
    r = R("mock")
    # data that is not a matrix -> R error with the reference's message (rcatnet_search.cpp:69-70)
    with pytest.raises(RError, match="Data is not a matrix"):
        r.call("ccnOptimalNetsForOrder", r.integer(s), *([r.NULL] * 11))
    # a short nodeOrder -> warning + default order (rcatnet_search.cpp:112-116); no catIndices -> inferred categories
    nets = optimal_nets_for_order(r, s, 2, 60, order=[1, 2, 3])
    assert "Invalid nodeOrder" in r.warnings()
    check_networks(nets, port.search_order(s, 2, 60), list(range(1, 7)), c)
    # parent pools as R hands them over: NULL entries and empty vectors are nodes without candidates
    pool = [None, [1], [], [1, 2], [3], [2, 4]]
    out = r.call("ccnOptimalNetsForOrder", r.data_matrix(s), r.NULL, r.numeric([2]), r.NULL, r.numeric([100]), r.NULL,
                 r.list([r.integer(np.arange(1, v + 1)) for v in c]),
                 r.list([r.NULL if p is None else r.integer(p) for p in pool]), r.NULL, r.NULL, r.logical(False), r.logical(False))
    check_networks([r.catnetwork(e) for e in r.elts(out)],
                   port.search_order(s, 2, 100, num_cats=c, parents_pool=[p or [] for p in pool]), list(range(1, 7)), c)
    # wrong arity and unknown routines never reach the shim
    with pytest.raises(RError, match="wrong number of arguments"):
        r.call("ccnOptimalNetsForOrder", r.NULL)
    with pytest.raises(RError, match="no such routine"):
        r.call("ccnMarginalProb", r.NULL)
    # an ABI error becomes an R error carrying the library's message
    with pytest.raises(RError, match="not in the mock"):
        r.call("ccnOptimalNetsSA", r.NULL, r.data_matrix(s), r.NULL, r.numeric([2]), r.NULL, r.numeric([60]), r.NULL, r.NULL,
               r.NULL, r.numeric([0]), r.NULL, r.matrix(r.numeric(np.full(36, 0.5)), 6, 6), r.string("BIC"),
               r.integer(np.arange(1, 7)), r.numeric([1]), r.numeric([0.9]), r.numeric([2]), r.numeric([4]), r.numeric([1]),
               r.numeric([1e-10]), r.numeric([1]), r.logical(True), r.logical(False))
    with pytest.raises(RError, match="cnet should be a catNetwork"):
        r.call("ccnSamples", r.NULL, r.integer([5]), r.NULL, r.numeric([0.0]))
    # ccnSetSeed / ccnReleaseCache
    r.call("ccnSetSeed", r.integer([77]))
    assert r.L.mock_seed() == 77
    assert r.is_null(r.call("ccnReleaseCache"))


def test_shim_releases_the_result_when_r_leaves_by_an_error():
    """an R error while the networks are being built (here: the allocator fails at the 40th vector of the export) leaves
    the shim by a non-local exit; the library's result then hangs on an external pointer and R's collector frees it
    (export_result leaked it).  In the normal course the shim releases it itself and the finalizer finds nothing"""
    from rshim_harness import R, RError, optimal_nets_for_order
    s, c = random_problem(5, 6, 90, 3)
    r = R("mock")
    r.L.mock_live_results.restype = int
    live0 = r.L.mock_live_results()
    nets = optimal_nets_for_order(r, s, 2, 80, num_cats=c)
    assert len(nets) > 1 and r.L.mock_live_results() == live0
    assert r.L.mr_run_finalizers() == 0                     # nothing was left to the collector
    args = [r.data_matrix(s), r.NULL, r.numeric([2]), r.NULL, r.numeric([80]), r.NULL,
            r.list([r.integer(np.arange(1, v + 1)) for v in c]), r.NULL, r.NULL, r.NULL, r.logical(False), r.logical(False)]
    r.L.mr_fail_alloc_after(40)
    with pytest.raises(RError, match="cannot allocate"):
        r.call("ccnOptimalNetsForOrder", *args)
    r.L.mr_fail_alloc_after(-1)
    assert r.L.mock_live_results() == live0 + 1             # the shim was left before it could free the result ...
    assert r.L.mr_run_finalizers() == 1                     # ... the finalizer of its guard does
    assert r.L.mock_live_results() == live0


@pytest.mark.parametrize("kind,routine", [("entropy", "ccnEntropyPairwise"), ("kl", "ccnKLPairwise"),
                                          ("pearson", "ccnPearsonPairwise")])
def test_shim_pairwise(kind, routine):
    from oracle import port
    from rshim_harness import R, RError
    s, c = random_problem(11, 7, 120, None)
    pert = (np.random.default_rng(1).random(s.shape) < 0.2).astype(np.int32)
    r = R("mock")
    for p in (None, pert):
        out = r.call(routine, r.data_matrix(s, real=True), r.NULL if p is None else r.data_matrix(p))
        mat = r.reals(out).reshape(7, 7).T          # R: matrix(mat, numnodes, numnodes), column-major
        assert np.array_equal(mat, port.pairwise(kind, s, p))
    order = r.ints(r.call("ccnEntropyOrder", r.data_matrix(s), r.NULL))
    assert order == [int(v) for v in port.pairwise("order", s)]
    with pytest.raises(RError, match="Data should be a matrix"):
        r.call(routine, r.integer(s), r.NULL)


def network_problem(seed):
    rng = np.random.default_rng(seed)
    n, m = int(rng.integers(2, 9)), int(rng.integers(5, 200))
    s, c = random_problem(seed, n, m, [None, 2, 3, 4][seed % 4], na_frac=[0.0, 0.1][seed % 2])
    parents = [sorted(rng.choice(i, size=min(i, int(rng.integers(0, 4))), replace=False).tolist(), reverse=seed % 3 == 0)
               for i in range(n)]
    pert = (rng.random(s.shape) < 0.3).astype(np.int32) if seed % 3 == 1 else None
    return s, c, parents, pert


def check_given_network_calls(r, port, seed):
    """ccnSetProb -> ccnLoglik (per node / by sample) -> ccnNodeLoglik -> ccnSamples on one seeded network"""
    from rshim_harness import mini_r_uniform
    s, c, parents, pert = network_problem(seed)
    n = len(c)
    s0 = np.where(s < 1, 99, s - 1).astype(np.int32)
    data = r.data_matrix(s)
    rp = r.NULL if pert is None else r.data_matrix(pert)
    fitted = r.call("ccnSetProb", r.new_catnetwork(c, parents), data, rp)
    net = r.catnetwork(fitted)
    probs, ll, sizes = port.net_set_prob(s0, c, parents, pert)
    assert net["parents"] == [list(p) for p in parents] and net["sample_sizes"] == [int(v) for v in sizes]
    assert all(np.array_equal(a, b) for a, b in zip(net["probs"], probs))
    total = 0.0
    for v in ll:
        total += float(v)
    assert net["loglik"] == (total if np.all(ll > -3.4e38) else -np.inf)
    nodes = port.net_loglik("nodes", s0, c, parents, probs, pert)
    expect = np.where(nodes > -3.4e38, nodes, -np.inf)
    assert np.array_equal(r.reals(r.call("ccnLoglik", fitted, data, rp, r.logical(False))), expect)
    by = port.net_loglik("bysample", s0, c, parents, probs, pert)
    assert np.array_equal(r.reals(r.call("ccnLoglik", fitted, data, rp, r.logical(True))), np.where(by > -3.4e38, by, -np.inf))
    sel = [n, 1] if n > 1 else [1]
    assert np.array_equal(r.reals(r.call("ccnNodeLoglik", fitted, r.numeric(sel), data, rp)), expect[np.array(sel) - 1])
    r.L.mr_set_rng(1000 + seed)
    drawn = r.ints(r.call("ccnSamples", fitted, r.numeric([57]), r.NULL, r.numeric([0.1 * (seed % 2)])))
    lib_seed = int(mini_r_uniform(1000 + seed) * 9007199254740992.0)
    theirs, _ = port.net_samples(c, parents, probs, 57, None, 0.1 * (seed % 2), lib_seed)
    assert np.array_equal(np.array(drawn, dtype=np.int64).reshape(57, n), theirs.astype(np.int64))


@pytest.mark.parametrize("seed", range(800, 816))
def test_shim_given_network_calls(seed):
    from oracle import port
    from rshim_harness import R
    check_given_network_calls(R("mock"), port, seed)


def test_shim_network_argument_errors():
    from rshim_harness import R, RError
    s, c, parents, _ = network_problem(801)
    r = R("mock")
    obj = r.new_catnetwork(c, parents)
    with pytest.raises(RError, match="cnet should be a catNetwork"):
        r.call("ccnLoglik", r.list([]), r.data_matrix(s), r.NULL, r.logical(False))
    with pytest.raises(RError, match="should be a matrix"):
        r.call("ccnLoglik", obj, r.integer(s), r.NULL, r.logical(False))
    with pytest.raises(RError, match="number of nodes"):
        r.call("ccnSetProb", obj, r.data_matrix(s[:, :-1]), r.NULL)
    with pytest.raises(RError, match="number of samples"):
        r.call("ccnSamples", obj, r.numeric([0]), r.NULL, r.numeric([0.0]))
    # nodes x samples beyond what an int-indexed buffer addresses: refused before anything is allocated (the product used to
    # be formed in int), and the error is raised with none of the call's C++ objects alive (rshim.cpp: guarded)
    with pytest.raises(RError, match="too large"):
        r.call("ccnSamples", obj, r.numeric([2 ** 30]), r.NULL, r.numeric([0.0]))
    # the shim is still usable afterwards
    assert len(r.ints(r.call("ccnSamples", r.call("ccnSetProb", obj, r.data_matrix(s), r.NULL), r.numeric([5]), r.NULL,
                             r.numeric([0.0])))) == 5 * len(c)


SA_KW = dict(temp_start=0.05, temp_cool_fact=0.8, temp_check_orders=3, max_iter=24, order_shuffles=1.5, stop_diff=1e-10,
             num_threads=3)


def check_sa_and_histogram(r, port, seed):
    """.Call("ccnOptimalNetsSA", ...23) and .Call("ccnParHistogram", ...14) with the arguments R/catnet.search.R:280-296
    and R/catnet.histo.R:116-129 pass; the chain draws from R's own generator through cnb_sa_params.unif, so the oracle
    fed the same generator (the miniature R's unif_rand, re-seeded) must propose and accept the same orders"""
    import ctypes
    r_unif = ctypes.cast(r.L.unif_rand, ctypes.c_void_p)
    s, c = random_problem(seed, 6 + seed % 3, 90, [None, 3][seed % 2])
    n = len(c)
    names = ["G%d" % (i + 1) for i in range(n)]
    start = (np.random.default_rng(seed).permutation(n) + 1).astype(np.int32)
    model = ["BIC", "AIC", "none"][seed % 3]
    cat_indices = r.list([r.integer(np.arange(1, v + 1)) for v in c])
    r.L.mr_set_rng(50 + seed)
    out = r.call("ccnOptimalNetsSA", _names(r, names),
                 r.data_matrix(s), r.NULL, r.numeric([2]), r.NULL, r.numeric([n * 20]), cat_indices, r.NULL, r.NULL,
                 r.numeric([0]), r.NULL, r.NULL, r.string(model), r.numeric(start), r.numeric([SA_KW["temp_start"]]),
                 r.numeric([SA_KW["temp_cool_fact"]]), r.numeric([SA_KW["temp_check_orders"]]), r.numeric([SA_KW["max_iter"]]),
                 r.numeric([SA_KW["order_shuffles"]]), r.numeric([SA_KW["stop_diff"]]), r.numeric([SA_KW["num_threads"]]),
                 r.logical(True), r.logical(False))
    r.L.mr_set_rng(50 + seed)
    theirs = port.search_sa(s, 2, n * 20, start, select_mode={"AIC": 1, "BIC": 2}.get(model, 0), unif=r_unif, num_cats=c,
                            want_probs=True, use_cache=True, **SA_KW)   # R passes TRUE (R/catnet.search.R:293-294), so does the call above
    nets = [r.catnetwork(e) for e in r.elts(out)]
    check_networks(nets, theirs["nets"], [int(v) for v in theirs["order"]], c, names)
    # the parent histogram: numeric vector N*N that R reshapes with matrix(vhisto, nrow = numnodes)
    r.L.mr_set_rng(90 + seed)
    vh = r.reals(r.call("ccnParHistogram", r.data_matrix(s), r.NULL, r.numeric([2]), r.NULL, r.numeric([n * 20]), cat_indices,
                        r.NULL, r.NULL, r.string(model), r.numeric([seed % 3]), r.numeric([8]), r.numeric([2]),
                        r.logical(True), r.logical(False)))
    r.L.mr_set_rng(90 + seed)
    hist, _ = port.search_hist(s, 2, n * 20, score={"AIC": 1, "BIC": 2}.get(model, 0), weight=seed % 3, max_iter=8,
                               num_threads=2, unif=r_unif, num_cats=c, use_cache=True)
    assert np.array_equal(vh.reshape(n, n), hist)


def _names(r, names):
    """character vector"""
    v = r.L.mr_category_names(len(names))
    for i, nm in enumerate(names):
        r.L.mr_set_string(v, i, nm.encode())
    return v


@pytest.mark.parametrize("seed", range(900, 906))
def test_shim_sa_and_histogram(seed):
    from oracle import port
    from rshim_harness import R
    check_sa_and_histogram(R("mock"), port, seed)
