"""Synthetic networks and samples for the parity tests and bench.py (fixtures either side of the hot path).

Numpy restatements of the reference's generators, driven by numpy's PCG64 instead of R's RNG (R streams are not
reproducible here and need not be: oracle and GPU consume the same arrays):

  random_parents        R/catnet.dags.R:5-35      genRandomParents
  order_nodes_descend   R/catnet.dags.R:91-124    orderNodesDescend
  random_cpts           R/catnet.probs.R:118-144  setRandomProb (delta1 = delta2 = 0.01, floor to 1e-3)
  sample_network        src/rcatnet.cpp:568-602   RCatnet::genSamples (topological order, first k with u <= cumsum)
  discretize_uniform    R/catnet.quant.R:87-102   cnDiscretize(mode = "uniform")
  default_max_complexity R/catnet.search.R:190-197

Networks are dicts {"cats": int[N], "parents": list of 0-based int lists, "cpts": list of float arrays
[prod(parent cats), C_node] with the FIRST listed parent most significant (problist.h:143-155)}.
"""
import json
import math
import os

import numpy as np

_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def order_nodes_descend(parents):
    n = len(parents)
    done = [False] * n
    path = []
    again = True
    while again:
        again = False
        for i in range(n):
            if not done[i] and all(done[p] for p in parents[i]):
                path.append(i)
                done[i] = True
                again = True
    return path


def random_parents(n, maxparents, rng):
    idx = rng.permutation(n)
    pars = [[] for _ in range(n)]
    for i in range(1, n):
        npars = int(math.floor(2 * maxparents * rng.random() + 0.5))
        if npars <= maxparents:
            continue
        npars = min(npars - maxparents, i)
        pick = np.sort(rng.choice(i, size=npars, replace=False))
        pars[idx[i]] = [int(idx[p]) for p in pick]
    norder = order_nodes_descend(pars)
    pos = {v: k for k, v in enumerate(norder)}
    return [sorted(p, key=lambda v: pos[v]) for p in pars]


def _random_dist(c, rng, d1=0.01, d2=0.01):
    limit = 10 * (1 + math.floor(1 / (0.5 - d1 - d2))) ** c
    p = None
    for _ in range(int(limit) + 1):
        p = rng.random(c)
        p = p / p.sum()
        if not (np.any((p < d1) | (p > 1 - d1)) or np.any((p > 0.5 - d2) & (p < 0.5 + d2))):
            break
    p = np.floor(1000 * p) / 1000
    p[0] = 1 - p[1:].sum()
    return p


def random_cpts(parents, cats, rng):
    cpts = []
    for i, ps in enumerate(parents):
        rows = int(np.prod([cats[p] for p in ps], dtype=np.int64)) if ps else 1
        cpts.append(np.stack([_random_dist(int(cats[i]), rng) for _ in range(rows)]))
    return cpts


def random_network(n, cats, maxparents, rng):
    cats = np.full(n, cats, dtype=np.int32) if np.isscalar(cats) else np.asarray(cats, dtype=np.int32)
    parents = random_parents(n, maxparents, rng)
    return {"cats": cats, "parents": parents, "cpts": random_cpts(parents, cats, rng)}


def sample_network(net, m, rng, dtype=np.int32):
    """Ancestral sampling; returns [M][N] 1-based categories (the layout R hands to .Call, node fastest)."""
    cats, parents, cpts = net["cats"], net["parents"], net["cpts"]
    n = len(parents)
    out = np.zeros((m, n), dtype=dtype)
    for i in order_nodes_descend(parents):
        cfg = np.zeros(m, dtype=np.int64)
        for p in parents[i]:
            cfg = cfg * int(cats[p]) + (out[:, p].astype(np.int64) - 1)
        cum = np.cumsum(cpts[i], axis=1)[cfg]          # [M][C]
        u = rng.random(m)
        k = (u[:, None] > cum).sum(axis=1)              # first k with u <= cum[k]
        out[:, i] = np.minimum(k, int(cats[i]) - 1) + 1
    return out


def discretize_uniform(x, ncats):
    """x: [nodes][samples] float; returns 1-based int32 categories, per-node uniform bins (qq_i > x rule)."""
    x = np.asarray(x, dtype=np.float64)
    lo = x.min(axis=1, keepdims=True)
    hi = x.max(axis=1, keepdims=True)
    out = np.full(x.shape, ncats, dtype=np.int32)
    for i in range(ncats, 0, -1):
        qq = lo + (hi - lo) * (i / ncats)
        out[qq > x] = i
    return out


def default_max_complexity(n, max_categories, max_parent_set):
    """R/catnet.search.R:191 -- evaluated in floating point and truncated (1775 for ALARM P=2, not 1776)."""
    return int(n * math.exp(math.log(max_categories) * max_parent_set) * (max_categories - 1))


def num_families(n, max_parent_set):
    return sum(math.comb(n, d + 1) for d in range(1, max_parent_set + 1))


# ---- committed fixtures derived from the reference's data/*.rda (tools/make_fixtures.py) ----
def load_alarm():
    with open(os.path.join(_GOLDEN, "alarmnet.json")) as f:
        j = json.load(f)
    cats = np.asarray(j["cats"], dtype=np.int32)
    return {"cats": cats, "parents": j["parents"], "nodes": j["nodes"],
            "cpts": [np.asarray(c, dtype=np.float64).reshape(-1, int(cats[i])) for i, c in enumerate(j["cpts"])]}


def load_breast():
    """100 genes x 98 samples, 2 categories (demo/breast.r:6-16 pipeline). Returns [M=98][N=100] int32, 1-based."""
    a = np.load(os.path.join(_GOLDEN, "breast_100x98.npy"))
    return np.ascontiguousarray(a.T.astype(np.int32))


def config(name, m=None, seed=None):
    """The BASELINE.json configs.  Returns dict(samples [M][N] 1-based, cats, order (0-based
    positions -> node), max_parent_set, max_complexity)."""
    if name in ("C1", "C4"):
        net = load_alarm()
        m = m or (1000 if name == "C1" else 20000)
        seed = seed if seed is not None else (1001 if name == "C1" else 4004)
        s = sample_network(net, m, np.random.Generator(np.random.PCG64(seed)))
        mc = default_max_complexity(37, 4, 2) if name == "C1" else 600
        return {"samples": s, "cats": net["cats"], "order": np.asarray(order_nodes_descend(net["parents"])),
                "max_parent_set": 2, "max_complexity": mc, "net": net}
    if name == "C2":
        s = load_breast()
        return {"samples": s, "cats": np.full(100, 2, dtype=np.int32), "order": np.arange(100),
                "max_parent_set": 3, "max_complexity": default_max_complexity(100, 2, 3)}
    if name in ("C3", "C5"):
        n, c, p = (100, 3, 3) if name == "C3" else (500, 4, 2)
        m = m or (100000 if name == "C3" else 1000000)
        seed = seed if seed is not None else (3003 if name == "C3" else 5005)
        rng = np.random.Generator(np.random.PCG64(seed))
        net = random_network(n, c, p, rng)
        s = sample_network(net, m, rng, dtype=np.int8 if name == "C5" else np.int32)
        return {"samples": s, "cats": net["cats"], "order": np.asarray(order_nodes_descend(net["parents"])),
                "max_parent_set": p, "max_complexity": default_max_complexity(n, c, p), "net": net}
    raise KeyError(name)
