"""Host-side mirror of the reference's R front-ends for the search path (R is not available in this image, so the
argument handling that sits above the .Call boundary is restated in Python, same names and meaning):

    cnSearchOrder  R/catnet.search.R:46-261     cnSearchSA  R/catnet.search.R:340-586 (+ .searchSA :263-338)
    .categorizeSample  R/catnet.categor.R:5-166  (integer / numeric matrices; data frames of factors are R-only)

Everything below validates and marshals; all scoring goes through the C ABI (catnet_b200._abi), which has no CPU
fallback.  Data layout follows R: `data` is a nodes x samples matrix (rows = nodes), values 1..C or NaN/None/<1 = NA.
"""
import math
import time
import warnings

import numpy as np

from . import _abi

NA = _abi.NA_INTEGER


class CatNetwork(dict):
    """slots of the S4 catNetwork the reference returns (R/catnet.class.R:4-20); nodes are in DATA order"""
    __getattr__ = dict.__getitem__


class CatNetworkEvaluate:
    """R/catnet.class.R:231-256"""

    def __init__(self, numnodes, numsamples):
        self.numnodes, self.numsamples = numnodes, numsamples
        self.nets, self.complexity, self.loglik, self.time = [], [], [], 0.0

    def find(self, complexity):
        """cnFind: the network with the given complexity, else the closest below it (R/catnet.find.R)"""
        best = None
        for net in self.nets:
            if net["complexity"] <= complexity and (best is None or net["complexity"] > best["complexity"]):
                best = net
        return best or (self.nets[0] if self.nets else None)

    def find_bic(self):
        m = self.numsamples
        return max(self.nets, key=lambda n: m * n["likelihood"] - 0.5 * math.log(m) * n["complexity"])

    def find_aic(self):
        m = self.numsamples
        return max(self.nets, key=lambda n: m * n["likelihood"] - n["complexity"])


def categorize_sample(data, perturbations=None, node_cats=None):
    """.categorizeSample for numeric input: per node the sorted distinct non-NA values become categories 1..C
    (R/catnet.categor.R:20-68); with node_cats the given value lists define the indices."""
    x = np.asarray(data, dtype=np.float64)
    if x.ndim != 2:
        raise ValueError("data should be a matrix or data frame")
    n, m = x.shape
    out = np.full((n, m), NA, dtype=np.int32)
    cats = []
    for i in range(n):
        ok = ~np.isnan(x[i])
        levels = sorted(set(x[i][ok].tolist())) if node_cats is None else list(node_cats[i])
        cats.append(levels)
        lut = {v: k + 1 for k, v in enumerate(levels)}
        for j in np.nonzero(ok)[0]:
            v = x[i, j]
            if v not in lut:
                raise ValueError("Incorrect categories")
            out[i, j] = lut[v]
    pert = None if perturbations is None else np.asarray(perturbations, dtype=np.int32)
    return out, pert, cats, max(len(c) for c in cats)


def _names_to_ids(lst, nodenames, what):
    if lst is None:
        return None
    if len(lst) != len(nodenames):
        raise ValueError("Wrong %s parameter" % what)
    out = []
    for row in lst:
        ids = []
        for v in (row if row is not None else []):
            if isinstance(v, str):
                if v not in nodenames:
                    raise ValueError("Invalid %s" % what)
                ids.append(nodenames.index(v) + 1)
            else:
                ids.append(int(v))
        out.append(ids)
    return out


def _reorder_to_data(net, order, categories, nodenames):
    """cnReorderNodes (R/catnet.def.R:691-741): express a network returned in order space in data order"""
    n = len(order)
    lbl = [int(v) - 1 for v in order]                 # position -> data index
    parents = [None] * n
    probs = [None] * n
    sizes = [0] * n
    for pos in range(n):
        parents[lbl[pos]] = [lbl[p] + 1 for p in net["parents"][pos]]      # 1-based data indices, table order
        probs[lbl[pos]] = net["probs"][pos] if net["probs"] else None
        sizes[lbl[pos]] = net["sample_sizes"][pos]
    return CatNetwork(numnodes=n, nodes=list(nodenames), parents=parents, categories=categories,
                      probabilities=probs, complexity=net["complexity"], likelihood=net["loglik"],
                      nodeSampleSizes=sizes, maxParents=max((len(p) for p in parents), default=0),
                      maxCategories=max(len(c) for c in categories))


def _common(data, perturbations, maxParentSet, parentSizes, maxComplexity, nodeCats, parentsPool, fixedParents,
            nodenames):
    data, perturbations, categories, max_categories = categorize_sample(data, perturbations, nodeCats)
    n, m = data.shape
    if n < 1 or m < 1:
        raise ValueError("Invalid sample")
    nodenames = [str(v) for v in (nodenames if nodenames is not None else range(1, n + 1))]
    if len(set(nodenames)) != n:
        raise ValueError("Repeated node names")
    maxParentSet = int(maxParentSet)
    if maxParentSet < 1:
        if parentSizes is not None:
            maxParentSet = int(max(parentSizes))
        maxParentSet = max(maxParentSet, 1)
    if parentSizes is not None and len(parentSizes) == n:
        parentSizes = np.clip(np.asarray(parentSizes, dtype=np.int32), 0, maxParentSet)
    else:
        parentSizes = None
    parentsPool = _names_to_ids(parentsPool, nodenames, "parentsPool")
    fixedParents = _names_to_ids(fixedParents, nodenames, "fixedParents")
    if maxComplexity <= 0:       # evaluated in floating point and truncated, like R (catnet.search.R:191)
        maxComplexity = int(n * math.exp(math.log(max_categories) * maxParentSet) * (max_categories - 1))
    min_complexity = sum(len(c) - 1 for c in categories)
    maxComplexity = max(int(maxComplexity), min_complexity)
    return (data, perturbations, categories, max_categories, n, m, nodenames, maxParentSet, parentSizes,
            parentsPool, fixedParents, maxComplexity)


def _truncate_fixed(fixedParents, maxParentSet):
    if fixedParents is None:
        return None
    out = []
    for i, row in enumerate(fixedParents):
        if len(row) > maxParentSet:
            warnings.warn("fixedParents for node %d exceeds the maximum parent size. It's reduced to %d" % (i + 1, maxParentSet))
            row = row[:maxParentSet]
        out.append(row)
    return out


def cnSearchOrder(data, perturbations=None, maxParentSet=0, parentSizes=None, maxComplexity=0, nodeOrder=None,
                  nodeCats=None, parentsPool=None, fixedParents=None, edgeProb=None, echo=False, nodeNames=None,
                  device=0, numDevices=0):
    """R/catnet.search.R:46-261.  Returns a CatNetworkEvaluate whose nets are in data order."""
    t1 = time.time()
    (data, perturbations, categories, max_categories, n, m, nodenames, maxParentSet, parentSizes, parentsPool,
     fixedParents, maxComplexity) = _common(data, perturbations, maxParentSet, parentSizes, maxComplexity, nodeCats,
                                            parentsPool, fixedParents, nodeNames)
    if edgeProb is not None:
        edgeProb = np.array(edgeProb, dtype=np.float64)
        if edgeProb.shape != (n, n):
            raise ValueError("edgeProb should be square matrix of length the number of nodes")
        edgeProb = np.clip(edgeProb, 0, 1)
        edgeProb[np.isnan(edgeProb)] = 0.5
        if (edgeProb == 1).any():                                    # certain edges become fixed parents (:122-135)
            fixedParents = [list(r) for r in (fixedParents or [[] for _ in range(n)])]
            for i in range(n):
                for j in range(n):
                    if edgeProb[i, j] == 1 and len(fixedParents[i]) < maxParentSet:
                        fixedParents[i].append(j + 1)
    fixedParents = _truncate_fixed(fixedParents, maxParentSet)
    if nodeOrder is None:
        nodeOrder = list(range(1, n + 1))
    if len(nodeOrder) != n:
        raise ValueError("Invalid nodeOrder")
    nodeOrder = [nodenames.index(v) + 1 if isinstance(v, str) else int(v) for v in nodeOrder]
    num_cats = None if nodeCats is None else [len(c) for c in categories]
    nets = _abi.search_order(np.ascontiguousarray(data.T), perturbations=None if perturbations is None else
                             np.ascontiguousarray(perturbations.T), max_parent_set=maxParentSet,
                             parent_sizes=parentSizes, max_complexity=maxComplexity, order=nodeOrder,
                             num_cats=num_cats, parents_pool=parentsPool, fixed_parents=fixedParents,
                             edge_prob=edgeProb, echo=echo, device=device, num_devices=numDevices)
    ev = CatNetworkEvaluate(n, m)
    for net in nets:
        ev.nets.append(_reorder_to_data(net, net["order"], categories, nodenames))
    ev.complexity = [t["complexity"] for t in ev.nets]
    ev.loglik = [t["likelihood"] for t in ev.nets]
    ev.time = time.time() - t1
    return ev


def cnSearchSA(data, perturbations=None, maxParentSet=0, parentSizes=None, maxComplexity=0, nodeCats=None,
               parentsPool=None, fixedParents=None, edgeProb=None, dirProb=None, selectMode="BIC", tempStart=1,
               tempCoolFact=0.9, tempCheckOrders=10, maxIter=100, orderShuffles=1, stopDiff=0, numThreads=2,
               priorSearch=None, echo=False, nodeNames=None, seed=None, device=0):
    """R/catnet.search.R:340-586.  `seed` feeds the library's uniform stream (R seeds its own RNG instead)."""
    t1 = time.time()
    (data, perturbations, categories, max_categories, n, m, nodenames, maxParentSet, parentSizes, parentsPool,
     fixedParents, maxComplexity) = _common(data, perturbations, maxParentSet, parentSizes, maxComplexity, nodeCats,
                                            parentsPool, fixedParents, nodeNames)
    fixedParents = _truncate_fixed(fixedParents, maxParentSet)
    if priorSearch is not None and (priorSearch.numnodes != n or priorSearch.numsamples != m):
        raise ValueError("priorSearch's number of nodes and sample size are not compatible with those of the data")
    if tempStart < 0:
        warnings.warn("tempStart is set to 0")
        tempStart = 0
    tempCoolFact = min(max(tempCoolFact, 0), 1)
    tempCheckOrders = max(int(tempCheckOrders), 1)
    maxIter = max(int(maxIter), tempCheckOrders)
    if dirProb is not None:
        d = np.array(dirProb, dtype=np.float64)
        with np.errstate(invalid="ignore", divide="ignore"):
            d = d / (d + d.T)
        d[np.isnan(d) | (d < 0) | (d > 1)] = 0.5
        dirProb = np.clip(d, 1e-8, 1 - 1e-8)
    rng = np.random.default_rng(seed)
    if priorSearch is None or not priorSearch.nets:
        start = (rng.permutation(n) + 1).astype(np.int32)               # sample(1:numnodes)
        prior = None
    else:
        prior = priorSearch
        opt = {"AIC": prior.find_aic, "BIC": prior.find_bic}.get(selectMode, lambda: prior.find(maxComplexity))()
        from .synth import order_nodes_descend
        start = np.asarray(order_nodes_descend([[p - 1 for p in ps] for ps in opt["parents"]]), dtype=np.int32) + 1
    sa = dict(select_mode={"AIC": 1, "BIC": 2}.get(selectMode, 0), start_order=start, temp_start=tempStart,
              temp_cool_fact=tempCoolFact, temp_check_orders=tempCheckOrders, max_iter=maxIter,
              order_shuffles=orderShuffles, stop_diff=stopDiff, num_threads=numThreads, use_cache=True,
              dir_prob=dirProb, seed=int(rng.integers(1, 2 ** 62)))
    num_cats = None if nodeCats is None else [len(c) for c in categories]
    res = _abi.search_sa(np.ascontiguousarray(data.T), sa, want_probs=True,
                         perturbations=None if perturbations is None else np.ascontiguousarray(perturbations.T),
                         max_parent_set=maxParentSet, parent_sizes=parentSizes, max_complexity=maxComplexity,
                         num_cats=num_cats, parents_pool=parentsPool, fixed_parents=fixedParents,
                         edge_prob=edgeProb, echo=echo, device=device)
    ev = CatNetworkEvaluate(n, m)
    for net in res["nets"]:
        new = _reorder_to_data(net, net["order"], categories, nodenames)
        if prior is not None:                                           # .searchSA :317-332
            old = prior.find(new["complexity"])
            if old is not None and old["complexity"] == new["complexity"] and old["likelihood"] > new["likelihood"]:
                new = old
        ev.nets.append(new)
    ev.complexity = [t["complexity"] for t in ev.nets]
    ev.loglik = [t["likelihood"] for t in ev.nets]
    ev.time = (prior.time if prior else 0.0) + time.time() - t1
    ev.order = res["order"]
    return ev


def cnSearchHist(data, perturbations=None, maxParentSet=1, parentSizes=None, maxComplexity=0, nodeCats=None,
                 parentsPool=None, fixedParents=None, score="BIC", weight="likelihood", maxIter=32, numThreads=2,
                 echo=False, nodeNames=None, seed=None, device=0):
    """R/catnet.histo.R:45-138.  Returns the numnodes x numnodes matrix R builds with
    matrix(vhisto, nrow=numnodes): element [child, parent] = summed weight of the edge parent -> child over the
    networks selected for `maxIter` random node orders.  `seed` feeds the library's uniform stream."""
    (data, perturbations, categories, max_categories, n, m, nodenames, maxParentSet, parentSizes, parentsPool,
     fixedParents, maxComplexity) = _common(data, perturbations, maxParentSet, parentSizes, maxComplexity, nodeCats,
                                            parentsPool, fixedParents, nodeNames)
    numThreads = max(int(numThreads), 1)
    maxIter = max(int(maxIter), numThreads)
    nweight = {"likelihood": 1, "score": 2}.get(weight, 0)
    rng = np.random.default_rng(seed)
    hist = dict(score={"AIC": 1, "BIC": 2}.get(score, 0), weight=nweight, max_iter=maxIter, num_threads=numThreads,
                use_cache=True, seed=int(rng.integers(1, 2 ** 62)))
    num_cats = None if nodeCats is None else [len(c) for c in categories]
    v, _ = _abi.search_hist(np.ascontiguousarray(data.T), hist,
                            perturbations=None if perturbations is None else np.ascontiguousarray(perturbations.T),
                            max_parent_set=maxParentSet, parent_sizes=parentSizes, max_complexity=maxComplexity,
                            num_cats=num_cats, parents_pool=parentsPool, fixed_parents=fixedParents, echo=echo,
                            device=device)
    # the C vector is [parent*N + child]; R's matrix(v, nrow = N) is column-major, i.e. element [child, parent]
    return np.ascontiguousarray(v.T)


# ------------------------------------------------------------------ pairwise metrics ----
def _pairwise_front(data, perturbations):
    """argument handling shared by cnEntropy / cnEdgeDistanceKL / cnEdgeDistancePearson / cnEntropyOrder
    (R/catnet.entropy.R:39-58): matrix of categories (nodes x samples) -> categorised int32 [M][N] for the C ABI"""
    x = np.asarray(data)
    if x.ndim != 2:
        raise ValueError("'data' should be a matrix or data frame of categories")
    cat, pert, _, _ = categorize_sample(x, perturbations)
    if (cat == NA).any():
        raise ValueError("pairwise metrics need complete data")
    return np.ascontiguousarray(cat.T), None if pert is None else np.ascontiguousarray(pert.T)


def cnEntropy(data, perturbations=None, device=0):
    """R/catnet.entropy.R:39-67: matrix of conditional entropies, [n1, n2] = H(n1 | n2), diagonal H(n1)"""
    s, p = _pairwise_front(data, perturbations)
    return _abi.pairwise("entropy", s, p, device)


def cnEdgeDistanceKL(data, perturbations, device=0):
    """R/catnet.entropy.R:69-101"""
    if perturbations is None:
        warnings.warn("Perturbations are essential for estimating the pairwise causality")
    s, p = _pairwise_front(data, perturbations)
    return _abi.pairwise("kl", s, p, device)


def cnEdgeDistancePearson(data, perturbations, device=0):
    """R/catnet.chisq.R:2-34"""
    if perturbations is None:
        warnings.warn("Perturbations are essential for estimating the pairwise causality")
    s, p = _pairwise_front(data, perturbations)
    return _abi.pairwise("pearson", s, p, device)


def cnEntropyOrder(data, perturbations=None, device=0):
    """R/catnet.entropy.R:103-132: node order (1-based) by decreasing total mutual information"""
    s, p = _pairwise_front(data, perturbations)
    return _abi.entropy_order(s, p, device)


# ------------------------------------------------------------------ a given network against data ----
def _network_front(obj, data, perturbations):
    """argument handling of cnLoglik / cnNodeLoglik (R/catnet.loglik.R:54-96, 300-350): data is nodes x samples with the
    object's category values; returns (samples [M][N] int32, perturbations [M][N] or None, cats, parents 0-based)"""
    x = np.asarray(data)
    if x.ndim != 2:
        raise ValueError("'data' should be a matrix or data frame of categories")
    if x.shape[0] != obj["numnodes"]:
        raise ValueError("The number of nodes in  the object and data should be equal")
    cat, pert, _, _ = categorize_sample(x, perturbations, obj["categories"])
    cats = np.array([len(c) for c in obj["categories"]], dtype=np.int32)
    parents = [[int(q) - 1 for q in (p or [])] for p in obj["parents"]]
    return (np.ascontiguousarray(cat.T), None if pert is None else np.ascontiguousarray(pert.T), cats, parents)


def _neg_inf(v):
    v = np.asarray(v, dtype=np.float64).copy()
    v[v <= -3.4028234663852886e38] = -np.inf          # -FLT_MAX is the library's minus infinity (rcatnet.cpp:295-300)
    return v


def cnLoglik(obj, data, perturbations=None, bysample=False, device=0):
    """R/catnet.loglik.R:54-106: log-likelihood of the data under the network; by sample, or summed over the nodes"""
    s, p, cats, parents = _network_front(obj, data, perturbations)
    v = _neg_inf(_abi.loglik(s, cats, parents, obj["probabilities"], p, "bysample" if bysample else "nodes", device))
    return v if bysample else float(v.sum())


def cnNodeLoglik(obj, node, data, perturbations=None, device=0):
    """R/catnet.loglik.R:300-364: per-node log-likelihood of the listed nodes (1-based indices or names)"""
    s, p, cats, parents = _network_front(obj, data, perturbations)
    nodes = [obj["nodes"].index(v) + 1 if isinstance(v, str) else int(v) for v in np.atleast_1d(node)]
    if any(v < 1 or v > obj["numnodes"] for v in nodes):
        raise ValueError("Wrong node")
    return _neg_inf(_abi.node_loglik(s, cats, parents, obj["probabilities"], nodes, p, device))


def cnSetProb(obj, data, perturbations=None, nodeCats=None, device=0):
    """R/catnet.samples.R:368-441: the network with its tables re-estimated from the data (categories from the data or
    from nodeCats)"""
    x = np.asarray(data)
    if x.ndim != 2:
        raise ValueError("'data' should be a matrix or data frame of categories")
    if x.shape[0] != obj["numnodes"]:
        raise ValueError("The number of nodes in the object and data should be equal.")
    cat, pert, categories, max_categories = categorize_sample(x, perturbations, nodeCats)
    cats = np.array([len(c) for c in categories], dtype=np.int32)
    parents = [[int(q) - 1 for q in (p or [])] for p in obj["parents"]]
    probs, ll, sizes = _abi.set_prob(np.ascontiguousarray(cat.T), cats, parents,
                                     None if pert is None else np.ascontiguousarray(pert.T), device)
    ll = _neg_inf(ll)
    complexity = int(sum((cats[i] - 1) * np.prod([cats[q] for q in parents[i]], dtype=np.int64) for i in range(len(cats))))
    new = CatNetwork(obj)
    new.update(categories=categories, maxCategories=max_categories, probabilities=probs, likelihood=float(ll.sum()),
               nodeSampleSizes=[int(v) for v in sizes], complexity=complexity)
    return new


def cnSamples(obj, numsamples=1, perturbations=None, as_index=True, naRate=0.0, seed=1, device=0):
    """R/catnet.samples.R:27-95 (output="matrix"): nodes x samples matrix drawn from the network; as_index=False maps the
    indices to the object's category values.  perturbations: nodes x samples (or one vector per node), 0 = free."""
    n = obj["numnodes"]
    numsamples = int(numsamples)
    pert = None
    if perturbations is not None:
        pert = np.asarray(perturbations, dtype=np.int32)
        if pert.ndim == 1:
            pert = np.repeat(pert[:, None], numsamples, axis=1)
        if pert.shape != (n, numsamples):
            raise ValueError("The perturbation size should be %d by %d" % (n, numsamples))
        pert = np.ascontiguousarray(pert.T)
    naRate = min(max(float(naRate), 0.0), 1.0)
    cats = np.array([len(c) for c in obj["categories"]], dtype=np.int32)
    parents = [[int(q) - 1 for q in (p or [])] for p in obj["parents"]]
    s = _abi.samples(cats, parents, obj["probabilities"], numsamples, pert, naRate, seed, device).T
    if as_index:
        return s
    out = np.full(s.shape, np.nan, dtype=np.float64)
    for i in range(n):
        ok = s[i] != NA
        out[i, ok] = np.asarray(obj["categories"][i], dtype=np.float64)[s[i, ok] - 1]
    return out
