"""ctypes view of oracle/_build/libcatnet_oracle.so -- the plain-C restatement of the path (TEST INFRASTRUCTURE).
Same call shapes as oracle/ref.py so that tests can swap one checker for the other."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_build", "libcatnet_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            subprocess.check_call(["make", "-C", _HERE, "oracle"])
        L = C.CDLL(_PATH)
        vp, ip, dp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)
        L.orc_family_score.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, vp, C.c_int, vp, vp, dp, ip]
        L.orc_search_order.argtypes = [C.c_int, C.c_int, vp, vp, C.c_int, vp, C.c_int, vp, vp, vp, vp, vp]
        L.orc_search_order.restype = vp
        L.orc_search_sa.argtypes = [C.c_int, C.c_int, vp, vp, C.c_int, vp, C.c_int, vp, vp, vp, vp, C.c_int, vp,
                                    C.c_double, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]
        L.orc_search_sa.restype = vp
        L.orc_search_hist.argtypes = [C.c_int, C.c_int, vp, vp, C.c_int, vp, C.c_int, vp, vp, vp, C.c_int, C.c_int,
                                      C.c_int, C.c_int, vp]
        L.orc_search_hist.restype = C.c_int
        L.orc_set_seed.argtypes = [C.c_ulonglong]
        L.orc_result_nslots.argtypes = [vp]
        L.orc_result_exists.argtypes = [vp, C.c_int]
        L.orc_result_net.argtypes = [vp, C.c_int, ip, dp]
        L.orc_result_node.argtypes = [vp, C.c_int, C.c_int, ip, vp, dp, dp, ip, ip, vp, C.c_int]
        L.orc_result_counts.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int]
        L.orc_result_sa_info.argtypes = [vp, vp, ip, ip]
        L.orc_result_sa_trace.argtypes = [vp, vp, vp, C.c_int]
        L.orc_result_free.argtypes = [vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def _pool(pool, n):
    if pool is None:
        return None
    m = np.zeros((n, n), dtype=np.int32)
    for i, row in enumerate(pool):
        if row is not None:
            row = list(row)
            m[i, :len(row)] = row
    return m


def family_score(samples0, num_cats, child, parents, keep=None):
    L = lib()
    s = _i32(samples0)
    M, N = s.shape
    cats, pars = _i32(num_cats), _i32(parents)
    size = int(cats[child]) * int(np.prod([cats[p] for p in pars], dtype=np.int64)) if len(pars) else int(cats[child])
    counts = np.zeros(size, dtype=np.float64)
    ll, nc = C.c_double(), C.c_int()
    k = None if keep is None else np.ascontiguousarray(keep, dtype=np.uint8)
    L.orc_family_score(_p(s), N, M, _p(cats), int(child), _p(pars), len(pars), _p(k), _p(counts), C.byref(ll), C.byref(nc))
    return counts, ll.value, nc.value


def _collect(L, h, N, want_probs=True):
    nets = []
    pars = np.zeros(64, dtype=np.int32)
    probs = np.zeros(1 << 16, dtype=np.float64)
    for k in range(L.orc_result_nslots(h)):
        if not L.orc_result_exists(h, k):
            continue
        cx, ll = C.c_int(), C.c_double()
        L.orc_result_net(h, k, C.byref(cx), C.byref(ll))
        net = {"slot": k, "complexity": cx.value, "loglik": ll.value, "parents": [], "node_loglik": [],
               "node_prior": [], "sample_sizes": [], "probs": []}
        for i in range(N):
            npar, nl, pr, ss, ps = C.c_int(), C.c_double(), C.c_double(), C.c_int(), C.c_int()
            L.orc_result_node(h, k, i, C.byref(npar), _p(pars), C.byref(nl), C.byref(pr), C.byref(ss), C.byref(ps),
                              _p(probs) if want_probs else None, len(probs))
            net["parents"].append([int(x) for x in pars[:npar.value]])
            net["node_loglik"].append(nl.value)
            net["node_prior"].append(pr.value)
            net["sample_sizes"].append(ss.value)
            if want_probs:
                net["probs"].append(probs[:ps.value].copy())
        nets.append(net)
    return nets


def search_order(samples, max_parent_set, max_complexity, order=None, perturbations=None, parent_sizes=None,
                 num_cats=None, parents_pool=None, fixed_parents=None, edge_prob=None, use_cache=False,
                 want_probs=True):
    L = lib()
    s = _i32(samples)
    M, N = s.shape
    e = None if edge_prob is None else np.ascontiguousarray(np.asarray(edge_prob, dtype=np.float64).T)
    h = L.orc_search_order(N, M, _p(s), _p(_i32(perturbations)), int(max_parent_set), _p(_i32(parent_sizes)),
                           int(max_complexity), _p(_i32(order)), _p(_i32(num_cats)), _p(_pool(parents_pool, N)),
                           _p(_pool(fixed_parents, N)), _p(e))
    if not h:
        return []
    try:
        return _collect(L, h, N, want_probs)
    finally:
        L.orc_result_free(h)


def _cache(L, on):
    """the reference's score cache (cache.cpp) restated: released (as R does before every call,
    R/catnet.search.R:279) and switched on or off for the next search"""
    L.orc_set_cache.argtypes = [C.c_int]
    L.orc_set_cache.restype = None
    L.orc_cache_release.argtypes = []
    L.orc_cache_release.restype = None
    L.orc_cache_release()
    L.orc_set_cache(int(bool(on)))


def _set_unif(L, unif):
    if unif is not None:
        L.orc_set_unif.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_set_unif.restype = None
        L.orc_set_unif(unif, None)


def search_sa(samples, max_parent_set, max_complexity, start_order, select_mode=0, temp_start=1.0,
              temp_cool_fact=0.9, temp_check_orders=10, max_iter=100, order_shuffles=1.0, stop_diff=1e-10,
              num_threads=1, use_cache=False, seed=1, perturbations=None, parent_sizes=None, num_cats=None,
              parents_pool=None, fixed_parents=None, edge_prob=None, want_probs=False, unif=None):
    """unif: address of a `double f(void *)` generator consumed instead of the seeded splitmix64 stream (what R's
    unif_rand is to the reference)"""
    L = lib()
    s = _i32(samples)
    M, N = s.shape
    L.orc_set_seed(int(seed))
    _set_unif(L, unif)
    _cache(L, use_cache)
    e = None if edge_prob is None else np.ascontiguousarray(np.asarray(edge_prob, dtype=np.float64).T)
    h = L.orc_search_sa(N, M, _p(s), _p(_i32(perturbations)), int(max_parent_set), _p(_i32(parent_sizes)),
                        int(max_complexity), _p(_i32(num_cats)), _p(_pool(parents_pool, N)),
                        _p(_pool(fixed_parents, N)), _p(e), int(select_mode), _p(_i32(start_order)),
                        float(temp_start), float(temp_cool_fact), int(temp_check_orders), int(max_iter),
                        float(order_shuffles), float(stop_diff), int(num_threads))
    _cache(L, False)
    if not h:
        return None
    try:
        nets = _collect(L, h, N, want_probs)
        order = np.zeros(N, dtype=np.int32)
        niter, nacc = C.c_int(), C.c_int()
        nt = L.orc_result_sa_info(h, _p(order), C.byref(niter), C.byref(nacc))
        trace = np.zeros(max(nt, 1), dtype=np.float64)
        acc = np.zeros(max(nt, 1), dtype=np.int32)
        L.orc_result_sa_trace(h, _p(trace), _p(acc), nt)
        return {"nets": nets, "order": order, "niter": niter.value, "naccept": nacc.value, "trace": trace[:nt],
                "trace_accept": acc[:nt]}
    finally:
        L.orc_result_free(h)


def search_hist(samples, max_parent_set, max_complexity, score=2, weight=1, max_iter=32, num_threads=2, seed=1,
                perturbations=None, parent_sizes=None, num_cats=None, parents_pool=None, fixed_parents=None, unif=None,
                use_cache=False):
    """cnSearchHist: returns (hist [N][N] with hist[parent][child], iterations).  score: 0 highest complexity, 1 AIC,
    2 BIC; weight: 0 one, 1 exp(loglik/(M N)), 2 exp(score/(M N))."""
    L = lib()
    s = _i32(samples)
    M, N = s.shape
    L.orc_set_seed(int(seed))
    _set_unif(L, unif)
    _cache(L, use_cache)
    hist = np.zeros((N, N), dtype=np.float64)
    it = L.orc_search_hist(N, M, _p(s), _p(_i32(perturbations)), int(max_parent_set), _p(_i32(parent_sizes)),
                           int(max_complexity), _p(_i32(num_cats)), _p(_pool(parents_pool, N)),
                           _p(_pool(fixed_parents, N)), int(score), int(weight), int(max_iter), int(num_threads), _p(hist))
    _cache(L, False)
    return hist, it


PAIRWISE_KINDS = {"entropy": 0, "kl": 1, "pearson": 2, "order": 3}


def pairwise(kind, samples, perturbations=None):
    """catnet_entropy.cpp restated (orc_pairwise).  samples [M][N] 1-based, no NA.  Returns out[n1][n2] like
    oracle.ref.pairwise (R's matrix(mat, N, N)); for "order" the 1-based order vector."""
    L = lib()
    L.orc_pairwise.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    s = _i32(samples)
    M, N = s.shape
    out = np.zeros((N, N), dtype=np.float64)
    order = np.zeros(N, dtype=np.int32)
    rc = L.orc_pairwise(PAIRWISE_KINDS[kind], N, M, _p(s), _p(_i32(perturbations)), _p(out), _p(order))
    if rc != 0:
        raise ValueError("orc_pairwise: missing or non-positive category")
    return order if kind == "order" else out.T.copy()


NET_MODES = {"nodes": 0, "bysample": 1, "node": 2, "total": 3}


def flatten_network(cats, parents, probs=None):
    """see oracle.ref.flatten_network (same layout)"""
    cats = _i32(cats)
    N = len(cats)
    maxp = max([len(p) for p in parents] + [1])
    npar = np.array([len(p) for p in parents], dtype=np.int32)
    pars = np.zeros((N, maxp), dtype=np.int32)
    off = np.zeros(N + 1, dtype=np.int64)
    for i, p in enumerate(parents):
        pars[i, :len(p)] = p
        size = int(cats[i])
        for q in p:
            size *= int(cats[q])
        off[i + 1] = off[i] + size
    flat = None if probs is None else np.concatenate([np.asarray(t, dtype=np.float64).ravel() for t in probs])
    return cats, maxp, npar, pars, flat, off


def net_loglik(mode, samples0, cats, parents, probs, perturbations=None):
    L = lib()
    L.orc_net_loglik.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_void_p]
    s = _i32(samples0)
    M, N = s.shape
    cats, maxp, npar, pars, flat, off = flatten_network(cats, parents, probs)
    out = np.zeros(M if mode == "bysample" else N, dtype=np.float64)
    L.orc_net_loglik(N, M, _p(s), _p(_i32(perturbations)), _p(cats), maxp, _p(npar), _p(pars), _p(flat), _p(off),
                     NET_MODES[mode], _p(out))
    return out[0] if mode == "total" else out


def net_set_prob(samples0, cats, parents, perturbations=None):
    L = lib()
    L.orc_net_set_prob.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 6
    s = _i32(samples0)
    M, N = s.shape
    cats, maxp, npar, pars, _, off = flatten_network(cats, parents)
    probs = np.zeros(off[-1], dtype=np.float64)
    ll = np.zeros(N, dtype=np.float64)
    sizes = np.zeros(N, dtype=np.int32)
    L.orc_net_set_prob(N, M, _p(s), _p(_i32(perturbations)), _p(cats), maxp, _p(npar), _p(pars), _p(off), _p(probs),
                       _p(ll), _p(sizes))
    return [probs[off[i]:off[i + 1]].copy() for i in range(N)], ll, sizes


def net_samples(cats, parents, probs, num_samples, perturbations=None, na_rate=0.0, seed=1):
    """cnSamples restated (orc_net_samples): ([M][N] int32, uniforms consumed)"""
    L = lib()
    L.orc_net_samples.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                  C.c_void_p, C.c_double, C.c_ulonglong, C.c_void_p]
    L.orc_net_samples.restype = C.c_long
    cats, maxp, npar, pars, flat, off = flatten_network(cats, parents, probs)
    N, M = len(cats), int(num_samples)
    out = np.zeros((M, N), dtype=np.int32)
    calls = L.orc_net_samples(N, _p(cats), maxp, _p(npar), _p(pars), _p(flat), _p(off), M, _p(_i32(perturbations)),
                              float(na_rate), int(seed), _p(out))
    if calls < 0:
        raise ValueError("cyclic network")
    return out, calls
