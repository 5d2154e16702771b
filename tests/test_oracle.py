"""CPU suite, part 1: pin the oracle restatement (oracle/catnet_oracle.c).

  * against the committed golden fixtures, which are outputs of the reference's own C++ (tools/make_golden.py);
  * against the reference core directly (oracle/_ref/libcatnet_ref.so) where it was built (needs /root/reference
    at build time; the .so travels to the GPU box), on further seeded problems.
All fp64 comparisons are bit-exact."""
import numpy as np
import pytest

from cases import (CASES, HIST_CASES, NETWORK_CASES, PAIRWISE_CASES, SA_CASES, SAMPLES_CASES, build_case, build_hist_case,
                   build_network_case, build_pairwise_case, build_samples_case, ref_args)
from helpers import (check_network_case, compare_golden, compare_search, constrained_problem, fully_perturbed_problem,
                     load_golden, random_problem)
from oracle import port


def test_family_scores_match_golden():
    for g in load_golden("ref_family_scores"):
        s, c = random_problem(g["seed"], 8, 150, g["cats"], g["na"])
        s0 = np.where(s < 1, 99, s - 1).astype(np.int32)
        counts, ll, nc = port.family_score(s0, c, g["child"], g["parents"])
        assert [int(v) for v in counts] == g["counts"]
        assert ll == float.fromhex(g["loglik"]) and nc == g["ncount"]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "seed%d" % c["seed"])
def test_search_order_matches_golden(case):
    s, kw = build_case(case)
    nets = port.search_order(s, kw["max_parent_set"], kw["max_complexity"], **ref_args(kw))
    compare_golden(nets, load_golden("ref_search_order")["seed%d" % case["seed"]])


def test_config_c1_matches_golden():
    from catnet_b200 import synth
    c = synth.config("C1")
    nets = port.search_order(c["samples"], 2, c["max_complexity"], order=c["order"] + 1, num_cats=c["cats"],
                             want_probs=False)
    assert compare_golden(nets, load_golden("ref_configs")["C1"]) == 819


@pytest.mark.parametrize("case", SA_CASES, ids=lambda c: "seed%d" % c["seed"])
def test_sa_matches_golden(case):
    s, cats = random_problem(case["seed"], case["n"], case["m"], case.get("cats"))
    r = port.search_sa(s, case["P"], case["K"], np.arange(1, case["n"] + 1), num_cats=cats, **case["sa"])
    g = load_golden("ref_search_sa")["seed%d" % case["seed"]]
    assert [int(v) for v in r["order"]] == g["order"]
    assert (r["niter"], r["naccept"]) == (g["niter"], g["naccept"])
    assert [float(v) for v in r["trace"]] == [float.fromhex(v) for v in g["trace"]]
    assert [int(v) for v in r["trace_accept"]] == g["trace_accept"]
    compare_golden(r["nets"], g["nets"])


@pytest.mark.parametrize("case", HIST_CASES, ids=lambda c: "seed%d" % c["seed"])
def test_hist_matches_golden(case):
    """cnSearchHist: the restatement reproduces the reference core's histogram bit for bit"""
    s, kw = build_hist_case(case)
    h, it = port.search_hist(s, case["P"], case["K"], **kw, **case["hist"])
    g = load_golden("ref_search_hist")["seed%d" % case["seed"]]
    assert it == g["iterations"]
    assert h.tolist() == [[float.fromhex(v) for v in row] for row in g["hist"]]


def test_port_hist_matches_reference_core(ref):
    s, cats = random_problem(78, 7, 110, 3)
    kw = dict(score=2, weight=1, max_iter=10, num_threads=3, seed=99, num_cats=cats)
    a, ia = ref.search_hist(s, 2, 150, use_cache=True, **kw)
    b, ib = port.search_hist(s, 2, 150, **kw)
    assert ia == ib and np.array_equal(a, b)


@pytest.mark.parametrize("seed", range(100, 112))
def test_port_matches_reference_core(ref, seed):
    rng = np.random.default_rng(seed)
    n, m = int(rng.integers(2, 10)), int(rng.integers(1, 200))
    cats = [None, 2, 3, 4][seed % 4]
    s, c = random_problem(seed, n, m, cats, na_frac=[0.0, 0.1][seed % 2])
    kw = dict(num_cats=c)
    if seed % 3 == 0:
        kw["perturbations"] = (rng.random(s.shape) < 0.2).astype(np.int32)
    if seed % 5 == 0:
        kw["edge_prob"] = 0.1 + 0.8 * rng.random((n, n))
    order = (rng.permutation(n) + 1).astype(np.int32)
    K = int(n * int(c.max()) ** 2 * 3)
    a = ref.search_order(s, 2, K, order=order, **kw)
    b = port.search_order(s, 2, K, order=order, **kw)
    assert compare_search(b, a) == len(a)


@pytest.mark.parametrize("seed", range(700, 740))
def test_port_matches_reference_core_constraints(ref, seed):
    """the restatement against the compiled reference with everything the R front-end can pass: parent pools, fixed
    parents, per-node parent-set sizes, 1..3 parents, tight and loose complexity limits, the reference's cache on/off"""
    s, P, K, kw = constrained_problem(seed)
    a = ref.search_order(s, P, K, use_cache=bool(seed % 2), **kw)
    b = port.search_order(s, P, K, **kw)
    assert len(a) == len(b) and compare_search(b, a) == len(a)


@pytest.mark.parametrize("seed", range(1100, 1130))
def test_port_node_perturbed_in_every_sample(ref, seed):
    """found by fuzzing the restatement against the compiled reference (constrained_problem(3042)): a node without a
    single unperturbed sample never gets parents"""
    s, P, K, kw = fully_perturbed_problem(seed)
    a = ref.search_order(s, P, K, use_cache=bool(seed % 2), **kw)
    b = port.search_order(s, P, K, **kw)
    assert len(a) == len(b) and compare_search(b, a) == len(a)
    gone = [int(np.where(kw["order"] == v + 1)[0][0]) for v in np.where((kw["perturbations"] != 0).all(axis=0))[0]]
    fixed = kw.get("fixed_parents")
    for net in b:
        for pos in gone:
            assert len(net["parents"][pos]) == (len(fixed[kw["order"][pos] - 1]) if fixed else 0) or fixed


def test_port_sa_matches_reference_core(ref):
    s, cats = random_problem(77, 7, 90, None)
    kw = dict(select_mode=2, temp_start=0.05, temp_cool_fact=0.8, temp_check_orders=3, max_iter=30,
              order_shuffles=1.5, stop_diff=1e-10, num_threads=3, seed=4242, num_cats=cats)
    a = ref.search_sa(s, 2, 200, np.arange(1, 8), use_cache=False, **kw)
    b = port.search_sa(s, 2, 200, np.arange(1, 8), **kw)
    assert np.array_equal(a["order"], b["order"]) and a["niter"] == b["niter"] and a["naccept"] == b["naccept"]
    assert np.array_equal(a["trace"], b["trace"]) and np.array_equal(a["trace_accept"], b["trace_accept"])
    compare_search(b["nets"], a["nets"], check_probs=False)


@pytest.mark.parametrize("case", PAIRWISE_CASES, ids=lambda c: "seed%d" % c["seed"])
def test_pairwise_matches_golden(case):
    """catnet_entropy.cpp: entropy / KL / Pearson matrices and the entropy order, bit for bit"""
    s, p = build_pairwise_case(case)
    g = load_golden("ref_pairwise")["seed%d" % case["seed"]]
    for kind in ("entropy", "kl", "pearson"):
        assert port.pairwise(kind, s, p).tolist() == [[float.fromhex(v) for v in row] for row in g[kind]], kind
    if g["order"] is not None:
        assert port.pairwise("order", s, p).tolist() == g["order"]


@pytest.mark.parametrize("seed", range(300, 330))
def test_port_pairwise_matches_reference(ref, seed):
    from oracle import ref as r
    if not r.pairwise_available():
        pytest.skip("oracle/_ref/libcatnet_ref_pairwise.so not built")
    rng = np.random.default_rng(seed)
    m, n, c = int(rng.integers(1, 300)), int(rng.integers(2, 9)), int(rng.integers(1, 7))
    s = rng.integers(1, c + 1, size=(m, n)).astype(np.int32)
    if n > 2:
        s[:, 2] = np.where(rng.random(m) < 0.7, s[:, 1], s[:, 2])
    if seed % 5 == 0:
        s[:, 0] = s[:, 1]
    p = (rng.random((m, n)) < rng.random()).astype(np.int32) if seed % 2 else None
    for kind in ("entropy", "kl", "pearson", "order"):
        assert np.array_equal(r.pairwise(kind, s, p), port.pairwise(kind, s, p)), kind


def test_port_pairwise_rejects_missing_values():
    s, _ = random_problem(1, 4, 20, 3)
    s[3, 1] = -2147483648
    with pytest.raises(ValueError):
        port.pairwise("entropy", s)


@pytest.mark.parametrize("case", NETWORK_CASES, ids=lambda c: "seed%d" % c["seed"])
def test_network_matches_golden(case):
    """cnSetProb / cnLoglik / cnNodeLoglik: the restatement against the reference's own CATNET methods, bit for bit"""
    check_network_case(port, case, build_network_case, load_golden("ref_network")["seed%d" % case["seed"]])


@pytest.mark.parametrize("seed", range(500, 520))
def test_port_network_matches_reference(ref, seed):
    rng = np.random.default_rng(seed)
    n, m = int(rng.integers(1, 9)), int(rng.integers(6, 300))
    s, c = random_problem(seed, n, m, [None, 2, 3, 5][seed % 4], na_frac=[0, 0.08][seed % 2])
    s0 = np.where(s < 1, 99, s - 1).astype(np.int32)
    parents = [sorted(rng.choice(i, size=min(i, int(rng.integers(0, 4))), replace=False).tolist(), reverse=seed % 3 == 0)
               for i in range(n)]
    pert = (rng.random(s.shape) < 0.3).astype(np.int32) if seed % 3 == 1 else None
    h = max(1, m // 2)
    pa, la, sa = ref.net_set_prob(s0[:h], c, parents, None if pert is None else pert[:h])
    pb, lb, sb = port.net_set_prob(s0[:h], c, parents, None if pert is None else pert[:h])
    assert all(np.array_equal(x, y) for x, y in zip(pa, pb)) and np.array_equal(la, lb) and np.array_equal(sa, sb)
    if seed % 5 == 0:
        pa = [np.where(t < 0.2, 0.0, t) for t in pa]
    for mode in ("nodes", "bysample", "node", "total"):
        if mode == "bysample" and max(len(p) for p in parents) == 0 and pert is None:
            continue
        assert np.array_equal(ref.net_loglik(mode, s0, c, parents, pa, pert), port.net_loglik(mode, s0, c, parents, pa, pert))


@pytest.mark.parametrize("case", SAMPLES_CASES, ids=lambda c: "seed%d" % c["seed"])
def test_samples_match_golden(case):
    """cnSamples: the restatement draws the reference's own samples under the injected uniform stream"""
    cats, parents, probs, pert, na = build_samples_case(case)
    g = load_golden("ref_samples")["seed%d" % case["seed"]]
    s, calls = port.net_samples(cats, parents, probs, case["m"], pert, na, seed=case["seed"])
    assert calls == g["calls"] and s.ravel().tolist() == g["samples"]


def test_port_samples_match_reference(ref):
    from oracle import ref as r
    if not r.samples_available():
        pytest.skip("oracle/_ref/libcatnet_ref_samples.so not built")
    for seed in range(700, 712):
        cats, parents, probs, pert, na = build_samples_case(dict(seed=seed, n=1 + seed % 9, m=10 + 37 * (seed % 7), maxcat=5,
                                                                 maxpar=3, pert=[0, 0.4][seed % 2], na=[0, 0.2, 0.7][seed % 3]))
        m = 10 + 37 * (seed % 7)
        a, ca = r.net_samples(cats, parents, probs, m, pert, na, seed)
        b, cb = port.net_samples(cats, parents, probs, m, pert, na, seed)
        assert ca == cb and np.array_equal(a, b)

@pytest.mark.parametrize("seed", __import__("cases").SA_CACHE_SEEDS)
def test_sa_and_hist_with_the_score_cache_match_golden(seed):
    """the restated score cache (cache_get / cache_set in oracle/catnet_oracle.c) against the compiled reference run with
    its cache ON and one thread: trajectory, parents in their LISTED sequence, every node's log-likelihood, histogram"""
    from cases import sa_cache_problem
    s, c, P, K, start, sa, kw, _ = sa_cache_problem(seed)
    g = load_golden("ref_search_sa_cache")["seed%d" % seed]
    r = port.search_sa(s, P, K, start, num_cats=c, use_cache=True, **dict(sa, num_threads=1), **kw)
    assert [int(v) for v in r["order"]] == g["order"]
    assert (r["niter"], r["naccept"]) == (g["niter"], g["naccept"])
    assert [float(v) for v in r["trace"]] == [float.fromhex(v) for v in g["trace"]]
    assert [int(v) for v in r["trace_accept"]] == g["trace_accept"]
    compare_golden(r["nets"], g["nets"])
    hk = dict(score=seed % 3, weight=1 + seed % 2, max_iter=10, num_threads=1, seed=sa["seed"])
    kh = {k: v for k, v in kw.items() if k != "edge_prob"}
    h, it = port.search_hist(s, 2, K, use_cache=True, num_cats=c, **hk, **kh)
    assert it == g["hist_iterations"]
    assert h.tolist() == [[float.fromhex(v) for v in row] for row in g["hist"]]
    if g["differs_from_cache_off"]:                              # the fixture is not the cache-off result in disguise
        off = port.search_sa(s, P, K, start, num_cats=c, use_cache=False, **dict(sa, num_threads=1), **kw)
        assert ([float(v) for v in off["trace"]] != [float.fromhex(v) for v in g["trace"]]
                or [n["parents"] for n in off["nets"]] != [n["parents"] for n in g["nets"]]
                or [n["loglik"] for n in off["nets"]] != [float.fromhex(n["loglik"]) for n in g["nets"]])
