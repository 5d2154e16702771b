"""Seeded search problems shared by the GPU parity tests, the oracle tests and tools/make_golden.py.
They cover what the reference's estimate() branches on (catnet_search2.h): equal vs mixed categories, NA,
perturbations, parent pools, fixed parents, edge priors, parentSizes, tiny/ragged sample counts, one node,
category inference, > 4 categories, maxComplexity cut-off, non-identity order."""
import numpy as np

from helpers import random_problem

CASES = [
    dict(seed=1, n=8, m=200, cats=None, P=2),
    dict(seed=2, n=9, m=120, cats=3, P=3),
    dict(seed=3, n=10, m=64, cats=2, P=3),
    dict(seed=4, n=8, m=300, cats=None, P=2, na=0.07),
    dict(seed=5, n=7, m=90, cats=4, P=2, na=0.05, pert=True),
    dict(seed=6, n=12, m=40, cats=2, P=2, K=30),
    dict(seed=7, n=9, m=500, cats=None, P=3, order="random"),
    dict(seed=8, n=9, m=150, cats=3, P=2, pools=True),
    dict(seed=9, n=9, m=150, cats=None, P=3, fixed=True),
    dict(seed=10, n=8, m=150, cats=3, P=2, prior=True),
    dict(seed=11, n=8, m=150, cats=None, P=2, prior=True, fixed=True, pools=True, psizes=True),
    dict(seed=12, n=6, m=10, cats=2, P=2),
    dict(seed=13, n=5, m=2, cats=2, P=2),
    dict(seed=14, n=1, m=20, cats=3, P=2),
    dict(seed=15, n=9, m=100, cats=None, P=2, infer=True, na=0.05),
    dict(seed=16, n=8, m=100, cats=5, P=2),            # > 4 categories: generic histogram kernel
    dict(seed=17, n=7, m=100, cats=[2, 6, 3, 2, 7, 3, 2], P=2),
    dict(seed=18, n=8, m=100, cats=4, P=3),            # d=3 with 4 categories: generic kernel for that group
    dict(seed=19, n=9, m=33, cats=3, P=4),             # 4 parents
    dict(seed=20, n=8, m=200, cats=2, P=2, pert=True, prior=True, order="random"),
    dict(seed=21, n=8, m=150, cats=None, P=2, prior="soft"),
    dict(seed=22, n=9, m=120, cats=3, P=3, prior="soft", fixed=True, order="random"),
]

SA_CASES = [
    dict(seed=31, n=7, m=80, cats=None, P=2, K=200,
         sa=dict(select_mode=0, temp_start=0.01, temp_cool_fact=0.9, temp_check_orders=5, max_iter=40,
                 order_shuffles=-1.5, stop_diff=1e-10, num_threads=2, seed=99)),
    dict(seed=32, n=8, m=120, cats=3, P=2, K=150,
         sa=dict(select_mode=2, temp_start=0.0, temp_cool_fact=1.0, temp_check_orders=1000, max_iter=30,
                 order_shuffles=0, stop_diff=1e-10, num_threads=4, seed=7)),
    dict(seed=33, n=6, m=60, cats=2, P=2, K=60,
         sa=dict(select_mode=1, temp_start=1.0, temp_cool_fact=0.5, temp_check_orders=4, max_iter=50,
                 order_shuffles=2.3, stop_diff=1e-3, num_threads=1, seed=12345)),
]


# cnSearchHist (rcatnet_hist.cpp): score 0 highest complexity / 1 AIC / 2 BIC; weight 0 one / 1 likelihood / 2 score
HIST_CASES = [
    dict(seed=41, n=7, m=90, cats=3, P=2, K=200, na=0.0, hist=dict(score=2, weight=1, max_iter=12, num_threads=2, seed=5)),
    dict(seed=42, n=6, m=70, cats=None, P=2, K=150, na=0.0, hist=dict(score=1, weight=2, max_iter=7, num_threads=3, seed=17)),
    dict(seed=43, n=8, m=120, cats=2, P=3, K=120, na=0.08, hist=dict(score=0, weight=0, max_iter=6, num_threads=1, seed=3)),
    dict(seed=44, n=9, m=200, cats=4, P=2, K=400, na=0.0, pert=True, hist=dict(score=2, weight=1, max_iter=8, num_threads=4, seed=11)),
    dict(seed=45, n=3, m=40, cats=2, P=2, K=30, na=0.0, hist=dict(score=2, weight=1, max_iter=9, num_threads=4, seed=2)),
]


# pairwise metrics (catnet_entropy.cpp): pert = fraction of perturbed (node, sample) entries; dup = a duplicated column
# (ties in the entropy order); shift = categories not starting at 1; allpert = one node perturbed in every sample
PAIRWISE_CASES = [
    dict(seed=51, n=7, m=200, cats=None),
    dict(seed=52, n=6, m=150, cats=3, pert=0.3),
    dict(seed=53, n=8, m=100, cats=2, pert=0.5, dup=True),
    dict(seed=54, n=5, m=90, cats=6, pert=0.2),                    # > 4 categories: histogram kernel
    dict(seed=55, n=1, m=30, cats=3),
    dict(seed=56, n=2, m=5, cats=2, pert=0.4),
    dict(seed=57, n=9, m=257, cats=4, pert=0.25, allpert=True, dup=True),
    dict(seed=58, n=40, m=9000, cats=None, shift=True),            # M >= 8192: pair tables from the tensor cores
    dict(seed=59, n=12, m=9000, cats=4, pert=0.35),
    dict(seed=60, n=6, m=33, cats=[2, 7, 3, 2, 5, 3], pert=0.3, shift=True),
]


def build_pairwise_case(c):
    """-> (samples [M][N] 1-based without NA, perturbations [M][N] or None)"""
    from helpers import random_problem
    s, _ = random_problem(c["seed"], c["n"], c["m"], c.get("cats"))
    rng = np.random.default_rng(2000 + c["seed"])
    if c.get("dup") and c["n"] > 2:
        s[:, 2] = s[:, 1]
    if c.get("shift"):
        s[:, ::2] += 2
    p = None
    if c.get("pert"):
        p = (rng.random(s.shape) < c["pert"]).astype(np.int32)
        if c.get("allpert"):
            p[:, 0] = 1
        if c.get("dup") and c["n"] > 2:
            p[:, 2] = p[:, 1]
    return s, p


# a given network against data (cnSetProb / cnLoglik / cnNodeLoglik): maxpar = parents per node at most, rev = parents
# listed in descending order (the listed order is the table's index order), zero = probabilities below 0.15 replaced
# by 0 before the likelihood is taken (zero-probability cells hit by samples)
NETWORK_CASES = [
    dict(seed=61, n=6, m=200, cats=None, maxpar=2),
    dict(seed=62, n=8, m=150, cats=3, maxpar=2, na=0.08),
    dict(seed=63, n=7, m=120, cats=2, maxpar=3, pert=0.3),
    dict(seed=64, n=5, m=90, cats=6, maxpar=2, na=0.05, pert=0.2),
    dict(seed=65, n=1, m=30, cats=3, maxpar=0),
    dict(seed=66, n=8, m=300, cats=4, maxpar=2, pert=0.25, zero=True),
    dict(seed=67, n=9, m=3000, cats=None, maxpar=3, na=0.04, zero=True, rev=True),
    dict(seed=68, n=6, m=40, cats=3, maxpar=3, rev=True),
    dict(seed=69, n=4, m=64, cats=2, maxpar=1, pert=0.5, allpert=True),
]


# cnSamples: random DAG whose topological order is a random permutation of the nodes
SAMPLES_CASES = [
    dict(seed=71, n=6, m=200, maxcat=4, maxpar=2),
    dict(seed=72, n=9, m=1000, maxcat=3, maxpar=3, na=0.3),
    dict(seed=73, n=5, m=257, maxcat=5, maxpar=2, pert=0.3),
    dict(seed=74, n=12, m=700, maxcat=4, maxpar=3, pert=0.5, na=0.5),
    dict(seed=75, n=1, m=50, maxcat=3, maxpar=0, na=0.9),
    dict(seed=76, n=7, m=3, maxcat=2, maxpar=2, pert=0.2, na=0.99),
]


def build_samples_case(c):
    """-> (cats, parents (0-based, listed order), probs (flat tables), perturbations [M][N] or None, na_rate)"""
    rng = np.random.default_rng(4000 + c["seed"])
    n, m = c["n"], c["m"]
    cats = rng.integers(1 if c["seed"] % 2 else 2, c["maxcat"] + 1, size=n).astype(np.int32)
    pos = np.argsort(rng.permutation(n))
    parents = []
    for i in range(n):
        earlier = [int(q) for q in range(n) if pos[q] < pos[i]]
        k = min(len(earlier), int(rng.integers(0, c["maxpar"] + 1)))
        parents.append([int(v) for v in rng.choice(earlier, size=k, replace=False)] if k else [])
    probs = []
    for i, p in enumerate(parents):
        rows = int(np.prod([cats[q] for q in p])) if p else 1
        t = rng.random((rows, cats[i])) ** 3
        t /= t.sum(axis=1, keepdims=True)
        probs.append(t.ravel())
    pert = None
    if c.get("pert"):
        pert = (rng.integers(0, c["maxcat"] + 2, size=(m, n)) * (rng.random((m, n)) < c["pert"])).astype(np.int32)
    return cats, parents, probs, pert, c.get("na", 0.0)


def build_network_case(c):
    """-> (samples [M][N] 1-based with NA, cats [N], parents (lists of 0-based nodes, listed order), perturbations);
    the tables are estimated on the first half of the samples"""
    from helpers import random_problem
    s, cats = random_problem(c["seed"], c["n"], c["m"], c.get("cats"), c.get("na", 0.0))
    rng = np.random.default_rng(3000 + c["seed"])
    parents = []
    for i in range(c["n"]):
        k = min(i, int(rng.integers(0, c["maxpar"] + 1)))
        p = sorted(int(v) for v in rng.choice(i, size=k, replace=False)) if k else []
        parents.append(p[::-1] if c.get("rev") else p)
    pert = None
    if c.get("pert"):
        pert = (rng.random(s.shape) < c["pert"]).astype(np.int32)
        if c.get("allpert"):
            pert[:, 1] = 1
    return s, cats, parents, pert


def build_hist_case(c):
    """-> (samples [M][N], kwargs of the checkers' search_hist)"""
    from helpers import random_problem
    s, cats = random_problem(c["seed"], c["n"], c["m"], c.get("cats"), c.get("na", 0.0))
    kw = dict(num_cats=cats)
    if c.get("pert"):
        kw["perturbations"] = (np.random.default_rng(c["seed"]).random(s.shape) < 0.15).astype(np.int32)
    return s, kw


def build_case(c):
    s, cats = random_problem(c["seed"], c["n"], c["m"], c.get("cats"), c.get("na", 0.0))
    n = c["n"]
    rng = np.random.default_rng(1000 + c["seed"])
    kw = dict(max_parent_set=c["P"])
    kw["max_complexity"] = c.get("K", int(n * (int(cats.max()) ** c["P"]) * (int(cats.max()) - 1)))
    kw["order"] = (rng.permutation(n) + 1).astype(np.int32) if c.get("order") == "random" else None
    if not c.get("infer"):
        kw["num_cats"] = cats
    if c.get("pert"):
        kw["perturbations"] = (rng.random(s.shape) < 0.15).astype(np.int32)
    if c.get("pools"):
        kw["parents_pool"] = [[int(v) for v in rng.choice(np.arange(1, n + 1), size=rng.integers(0, n), replace=False)]
                              for _ in range(n)]
    if c.get("fixed"):
        kw["fixed_parents"] = [[int(v) for v in rng.choice(np.arange(1, n + 1), size=rng.integers(0, 3), replace=False)]
                               for _ in range(n)]
    if c.get("prior"):
        e = rng.random((n, n))
        if c["prior"] == "soft":
            e = 0.05 + 0.9 * e
        else:
            e[rng.random((n, n)) < 0.1] = 0.0
            e[rng.random((n, n)) < 0.05] = 1.0
        kw["edge_prob"] = e
    if c.get("psizes"):
        kw["parent_sizes"] = rng.integers(0, c["P"] + 1, size=n).astype(np.int32)
    return s, kw


def ref_args(kw):
    """keyword arguments of oracle.ref.search_order / oracle.port.search_order for a case built by build_case"""
    return dict(order=kw["order"], perturbations=kw.get("perturbations"), parent_sizes=kw.get("parent_sizes"),
                num_cats=kw.get("num_cats"), parents_pool=kw.get("parents_pool"),
                fixed_parents=kw.get("fixed_parents"), edge_prob=kw.get("edge_prob"))

# cnSearchSA / cnSearchHist with the reference's score cache ON (what R passes, R/catnet.search.R:293-294): fixtures from the
# compiled reference with ONE thread (with more its stores interleave by timing).  big = enough complete samples for the
# tensor-core tiles (the families the cache answers are re-scored in their cached listing from the resident tile tables).
SA_CACHE_SEEDS = list(range(24)) + [100, 101, 102, 103]


def sa_cache_problem(seed):
    """-> (samples, cats, P, K, start_order, sa kwargs (num_threads = 1 + seed % 3), search kwargs, never_kept or None)"""
    from helpers import random_problem
    rng = np.random.default_rng(seed)
    big = seed >= 100
    n, m = int(rng.integers(3, 11)), int(rng.integers(8, 300))
    if big:
        n, m = 9 + seed % 3, 8200 + 13 * (seed % 4)
    s, c = random_problem(seed, n, m, [None, 2, 3, None][seed % 4] if not big else [None, 4, 3, None][seed % 4],
                          na_frac=0.0 if big else [0, 0.1][seed % 2])
    kw, never = {}, None
    if seed % 3 == 0 and not big:
        pert = (rng.random(s.shape) < 0.2).astype(np.int32)
        if seed % 6 == 0:
            pert[:, int(rng.integers(0, n))] = 1              # a node perturbed in every sample
        kw["perturbations"] = pert
        never = (pert.min(axis=0) > 0).astype(np.uint8)
    if seed % 7 == 0:
        kw["parents_pool"] = [sorted((rng.choice(n, size=int(rng.integers(0, n)), replace=False) + 1).tolist()) for _ in range(n)]
    if seed % 11 == 0:
        kw["fixed_parents"] = [sorted((rng.choice(n, size=int(rng.integers(0, 2)), replace=False) + 1).tolist()) for _ in range(n)]
    if seed % 5 == 0:
        kw["edge_prob"] = 0.1 + 0.8 * rng.random((n, n))
    if seed % 13 == 0:
        kw["parent_sizes"] = rng.integers(0, 4, size=n).astype(np.int32)
    sa = dict(select_mode=seed % 3, temp_start=[1.0, 0.05, 10.0][seed % 3], temp_cool_fact=0.9,
              temp_check_orders=int(rng.integers(2, 8)), max_iter=int(rng.integers(30, 90)) if not big else 24,
              order_shuffles=[1.0, 1.5, -1.0, -2.5, 0.0, 3.0][seed % 6], stop_diff=1e-10,
              num_threads=1 + seed % 3, seed=int(rng.integers(1, 1 << 40)))
    start = (rng.permutation(n) + 1).astype(np.int32)
    P = 2 + (seed % 4 == 0)
    K = int(np.sum(c - 1)) + n * int(c.max()) ** P
    return s, c, P, K, start, sa, kw, never
