"""Shared helpers for the parity tests: random problems and network comparison."""
import numpy as np

NA = -2147483648


def random_problem(seed, n, m, cats=None, na_frac=0.0, structured=True):
    """[M][N] int32 1-based samples with some dependence between nodes, optional NAs. cats: int, list, or None=mixed 2..4"""
    rng = np.random.default_rng(seed)
    if cats is None:
        c = rng.integers(2, 5, size=n)
    elif np.isscalar(cats):
        c = np.full(n, cats)
    else:
        c = np.asarray(cats)
    s = np.zeros((m, n), dtype=np.int32)
    for i in range(n):
        base = rng.integers(0, c[i], size=m)
        if structured and i > 0:
            p = rng.integers(0, i)
            copy = rng.random(m) < 0.5
            base = np.where(copy, (s[:, p] - 1) % c[i], base)
        s[:, i] = base + 1
    # make sure every category occurs (so that inferred and given category counts agree)
    for i in range(n):
        s[:c[i], i] = np.arange(1, c[i] + 1)
    if na_frac > 0:
        mask = rng.random(s.shape) < na_frac
        mask[:int(c.max()), :] = False
        s[mask] = NA
    return s, c.astype(np.int32)


def nets_by_complexity(nets):
    return {n["complexity"]: n for n in nets}


def compare_search(ours, theirs, rel=1e-9, check_probs=True):
    """ours: list from catnet_b200._abi.read_result; theirs: list from oracle.ref.search_order.
    Returns the number of networks whose log-lik is bit-identical."""
    a, b = nets_by_complexity(ours), nets_by_complexity(theirs)
    assert sorted(a) == sorted(b), "different sets of complexity slots: only ours %s, only ref %s" % (
        sorted(set(a) - set(b))[:10], sorted(set(b) - set(a))[:10])
    exact = 0
    for cx in sorted(a):
        x, y = a[cx], b[cx]
        assert x["parents"] == y["parents"], "complexity %d: parents differ\n ours %s\n ref  %s" % (
            cx, x["parents"], y["parents"])
        assert abs(x["loglik"] - y["loglik"]) <= rel * abs(y["loglik"]) + 1e-300, (cx, x["loglik"], y["loglik"])
        exact += x["loglik"] == y["loglik"]
        assert x["sample_sizes"] == y["sample_sizes"], cx
        for k in range(len(x["parents"])):
            assert abs(x["node_loglik"][k] - y["node_loglik"][k]) <= rel * abs(y["node_loglik"][k]) + 1e-300
            if check_probs and len(x["probs"]):
                np.testing.assert_allclose(x["probs"][k], y["probs"][k], rtol=1e-12, atol=0)
    return exact


def load_golden(name):
    import gzip
    import json
    import os
    with gzip.open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".json.gz"), "rt") as f:
        return json.load(f)


def compare_golden(ours, golden):
    """ours: list of network dicts; golden: list packed by tools/make_golden.py (hex floats). Bit-exact."""
    a = nets_by_complexity(ours)
    b = {g["complexity"]: g for g in golden}
    assert sorted(a) == sorted(b), "different complexity slots: only ours %s, only reference %s" % (
        sorted(set(a) - set(b))[:10], sorted(set(b) - set(a))[:10])
    for cx in sorted(a):
        x, y = a[cx], b[cx]
        assert x["parents"] == y["parents"], "complexity %d: parents differ\n ours %s\n ref  %s" % (
            cx, x["parents"], y["parents"])
        assert x["loglik"] == float.fromhex(y["loglik"]), (cx, x["loglik"], float.fromhex(y["loglik"]))
        if "node_loglik" in y:
            assert [float(v) for v in x["node_loglik"]] == [float.fromhex(v) for v in y["node_loglik"]], cx
            assert x["sample_sizes"] == y["sample_sizes"], cx
    return len(a)


def unhex(v):
    return np.array([float.fromhex(x) for x in v], dtype=np.float64)


def check_network_case(impl, case, build, g, zero_based=True):
    """impl: object with net_set_prob(samples0, cats, parents, pert) and net_loglik(mode, samples0, cats, parents, probs,
    pert) (oracle.port / oracle.ref signatures); compares with the fixture g bit for bit"""
    s, cats, parents, pert = build(case)
    s0 = np.where(s < 1, 99, s - 1).astype(np.int32)
    h = max(1, len(s) // 2)
    probs, ll, sizes = impl.net_set_prob(s0[:h], cats, parents, None if pert is None else pert[:h])
    assert all(np.array_equal(a, unhex(b)) for a, b in zip(probs, g["set_prob"]))
    assert np.array_equal(ll, unhex(g["set_loglik"])) and [int(v) for v in sizes] == g["set_sizes"]
    used = [unhex(t) for t in g["probs_used"]]
    assert np.array_equal(impl.net_loglik("nodes", s0, cats, parents, used, pert), unhex(g["nodes"]))
    assert np.array_equal(impl.net_loglik("node", s0, cats, parents, used), unhex(g["node"]))
    if "bysample" in g:
        assert np.array_equal(impl.net_loglik("bysample", s0, cats, parents, used, pert), unhex(g["bysample"]))


def constrained_problem(seed):
    """a small seeded search problem with everything the R front-end can pass (parent pools, fixed parents, per-node
    parent-set sizes, perturbations, edge priors, 1..3 parents, a tight or loose complexity limit):
    -> samples, max_parent_set, max_complexity, keyword arguments of oracle.ref / oracle.port search_order"""
    rng = np.random.default_rng(seed)
    n, m = int(rng.integers(3, 9)), int(rng.integers(2, 150))
    s, c = random_problem(seed, n, m, [None, 2, 3, 4, None][seed % 5], na_frac=[0.0, 0.15, 0.0][seed % 3])
    P = 1 + seed % 3
    kw = dict(num_cats=c, order=(rng.permutation(n) + 1).astype(np.int32))
    if seed % 2 == 0:
        kw["parents_pool"] = [[int(v) for v in rng.choice(np.arange(1, n + 1), size=rng.integers(0, n), replace=False)
                               if v != i + 1] for i in range(n)]
    if seed % 3 == 1:
        kw["fixed_parents"] = [[int(v) for v in rng.choice(np.arange(1, n + 1), size=rng.integers(0, 2), replace=False)
                                if v != i + 1] for i in range(n)]
    if seed % 4 == 1:
        kw["parent_sizes"] = rng.integers(0, P + 1, size=n).astype(np.int32)
    if seed % 4 == 2:
        kw["perturbations"] = (rng.random(s.shape) < 0.25).astype(np.int32)
    if seed % 5 == 3:
        kw["edge_prob"] = 0.05 + 0.9 * rng.random((n, n))
    lo = int(np.sum(c - 1))
    K = int(rng.integers(lo, lo + n * int(c.max()) ** P * 2 + 1))
    return s, P, K, kw


def fully_perturbed_problem(seed):
    """a search problem in which one or two nodes are perturbed in EVERY sample (the reference then scores their
    candidates on zero samples and never gives them parents: catnet_class.h:824-826, catnet_search2.h:637-640)"""
    rng = np.random.default_rng(seed)
    n, m = int(rng.integers(4, 9)), int(rng.integers(5, 120))
    s, c = random_problem(seed, n, m, [3, None, 2, None][seed % 4], na_frac=[0.0, 0.1][seed % 2])
    pert = (rng.random(s.shape) < 0.2).astype(np.int32)
    for v in rng.choice(n, size=1 + seed % 2, replace=False):
        pert[:, v] = 1
    kw = dict(num_cats=c, order=(rng.permutation(n) + 1).astype(np.int32), perturbations=pert)
    if seed % 3 == 0:
        kw["edge_prob"] = 0.1 + 0.8 * rng.random((n, n))
    if seed % 5 == 1:
        kw["fixed_parents"] = [[int(v) for v in rng.choice(np.arange(1, n + 1), size=rng.integers(0, 2), replace=False)
                                if v != i + 1] for i in range(n)]
    P = 1 + seed % 3
    return s, P, int(np.sum(c - 1)) + n * int(c.max()) ** P, kw
