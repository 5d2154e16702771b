"""ctypes view of oracle/_ref/libcatnet_ref.so -- the reference's own C++ search core (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The library is built by oracle/Makefile from the sources under /root/reference/src (see ref_driver.cpp).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libcatnet_ref.so")
_lib = None

NA_INT = -2147483648  # R's NA_INTEGER


def available():
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libcatnet_ref.so missing: run `make -C oracle ref` where /root/reference exists")
        L = C.CDLL(_PATH)
        ip, dp, vp = C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_void_p
        L.ref_family_score.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, vp, C.c_int, vp, dp, ip]
        L.ref_family_score.restype = C.c_int
        L.ref_search_order.argtypes = [C.c_int, C.c_int, vp, vp, C.c_int, vp, C.c_int, vp, vp, vp, vp, vp, C.c_int]
        L.ref_search_order.restype = vp
        L.ref_search_sa.argtypes = [C.c_int, C.c_int, vp, vp, C.c_int, vp, C.c_int, vp, vp, vp, vp, vp,
                                    C.c_int, vp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double,
                                    C.c_int, C.c_int]
        L.ref_search_sa.restype = vp
        L.ref_search_hist.argtypes = [C.c_int, C.c_int, vp, vp, C.c_int, vp, C.c_int, vp, vp, vp, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_int, vp]
        L.ref_search_hist.restype = C.c_int
        L.ref_result_nslots.argtypes = [vp]
        L.ref_result_exists.argtypes = [vp, C.c_int]
        L.ref_result_net.argtypes = [vp, C.c_int, ip, dp]
        L.ref_result_node.argtypes = [vp, C.c_int, C.c_int, ip, vp, dp, dp, ip, ip, vp, C.c_int]
        L.ref_result_free.argtypes = [vp]
        L.ref_result_sa_info.argtypes = [vp, vp, ip, ip]
        L.ref_result_sa_trace.argtypes = [vp, vp, vp, C.c_int]
        L.ref_set_seed.argtypes = [C.c_uint64]
        L.ref_rng_calls.restype = C.c_long
        L.ref_score_family_list.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, C.c_int, C.c_int, vp]
        L.ref_score_family_list.restype = C.c_double
        L.ref_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _pool_matrix(pool, N):
    """list-of-lists of 1-based labels (None/[] = none) -> [N][N] int32, 0 padded."""
    if pool is None:
        return None
    m = np.zeros((N, N), dtype=np.int32)
    for i, row in enumerate(pool):
        if row is None:
            continue
        row = list(row)
        m[i, :len(row)] = row
    return m


def family_score(samples0, num_cats, child, parents):
    """samples0: [M][N] int32 zero-based codes.  Returns (counts float64[table], loglik, ncount)."""
    L = lib()
    s = _i32(samples0)
    M, N = s.shape
    cats = _i32(num_cats)
    pars = _i32(parents)
    size = int(cats[child]) * int(np.prod([cats[p] for p in pars], dtype=np.int64)) if len(pars) else int(cats[child])
    counts = np.zeros(size, dtype=np.float64)
    ll = C.c_double()
    nc = C.c_int()
    n = L.ref_family_score(_p(s), N, M, _p(cats), int(child), _p(pars), len(pars), _p(counts), C.byref(ll), C.byref(nc))
    assert n == size, (n, size)
    return counts, ll.value, nc.value


def score_family_list(samples0, num_cats, children, parents):
    L = lib()
    s = _i32(samples0)
    M, N = s.shape
    ch = _i32(children)
    pa = _i32(parents).reshape(len(ch), -1)
    out = np.zeros(len(ch), dtype=np.float64)
    L.ref_score_family_list(_p(s), N, M, _p(_i32(num_cats)), _p(ch), _p(pa), pa.shape[1], len(ch), _p(out))
    return out


def _collect(L, h, N, want_probs=True):
    nslots = L.ref_result_nslots(h)
    nets = []
    pars = np.zeros(N, dtype=np.int32)
    probs = np.zeros(1 << 16, dtype=np.float64)
    for k in range(nslots):
        if not L.ref_result_exists(h, k):
            continue
        cx, ll = C.c_int(), C.c_double()
        L.ref_result_net(h, k, C.byref(cx), C.byref(ll))
        net = {"slot": k, "complexity": cx.value, "loglik": ll.value, "parents": [], "node_loglik": [],
               "node_prior": [], "sample_sizes": [], "probs": []}
        for i in range(N):
            npar, nl, pr, ss, ps = C.c_int(), C.c_double(), C.c_double(), C.c_int(), C.c_int()
            rc = L.ref_result_node(h, k, i, C.byref(npar), _p(pars), C.byref(nl), C.byref(pr), C.byref(ss),
                                   C.byref(ps), _p(probs), len(probs))
            assert rc == 0
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
    """samples: [M][N] int32, 1-based categories in LABEL space (NA = NA_INT or <1).  order: 1-based labels.
    edge_prob: the R matrix as numpy [N][N] with edge_prob[child][parent] (R passes it column-major, so the flat
    buffer handed to C is edge_prob.T.ravel(); rcatnet_search.cpp:263-266)."""
    L = lib()
    s = _i32(samples)
    M, N = s.shape
    order = _i32(np.arange(1, N + 1) if order is None else order)
    e = None if edge_prob is None else _f64(np.asarray(edge_prob, dtype=np.float64).T)
    h = L.ref_search_order(N, M, _p(s), _p(_i32(perturbations)), int(max_parent_set), _p(_i32(parent_sizes)),
                           int(max_complexity), _p(order), _p(_i32(num_cats)), _p(_pool_matrix(parents_pool, N)),
                           _p(_pool_matrix(fixed_parents, N)), _p(e), int(bool(use_cache)))
    try:
        return _collect(L, h, N, want_probs)
    finally:
        L.ref_result_free(h)
        L.ref_release_cache()


def search_sa(samples, max_parent_set, max_complexity, start_order, select_mode=0, temp_start=1.0,
              temp_cool_fact=0.9, temp_check_orders=10, max_iter=100, order_shuffles=1.0, stop_diff=1e-10,
              num_threads=1, use_cache=True, seed=1, perturbations=None, parent_sizes=None, num_cats=None,
              parents_pool=None, fixed_parents=None, edge_prob=None, dir_prob=None, want_probs=False):
    L = lib()
    s = _i32(samples)
    M, N = s.shape
    L.ref_release_cache()
    L.ref_set_seed(int(seed))
    e = None if edge_prob is None else _f64(np.asarray(edge_prob, dtype=np.float64).T)
    dpm = None if dir_prob is None else _f64(np.asarray(dir_prob, dtype=np.float64).T)
    h = L.ref_search_sa(N, M, _p(s), _p(_i32(perturbations)), int(max_parent_set), _p(_i32(parent_sizes)),
                        int(max_complexity), _p(_i32(num_cats)), _p(_pool_matrix(parents_pool, N)),
                        _p(_pool_matrix(fixed_parents, N)), _p(e), _p(dpm), int(select_mode),
                        _p(_i32(start_order)), float(temp_start), float(temp_cool_fact), int(temp_check_orders),
                        int(max_iter), float(order_shuffles), float(stop_diff), int(num_threads),
                        int(bool(use_cache)))
    try:
        nets = _collect(L, h, N, want_probs)
        order = np.zeros(N, dtype=np.int32)
        niter, nacc = C.c_int(), C.c_int()
        nt = L.ref_result_sa_info(h, _p(order), C.byref(niter), C.byref(nacc))
        trace = np.zeros(max(nt, 1), dtype=np.float64)
        acc = np.zeros(max(nt, 1), dtype=np.int32)
        L.ref_result_sa_trace(h, _p(trace), _p(acc), nt)
        return {"nets": nets, "order": order, "niter": niter.value, "naccept": nacc.value,
                "trace": trace[:nt], "trace_accept": acc[:nt], "rng_calls": L.ref_rng_calls()}
    finally:
        L.ref_result_free(h)
        L.ref_release_cache()


def search_hist(samples, max_parent_set, max_complexity, score=2, weight=1, max_iter=32, num_threads=2, use_cache=True,
                seed=1, perturbations=None, parent_sizes=None, num_cats=None, parents_pool=None, fixed_parents=None):
    """cnSearchHist through the reference's own estimate()/threads/cache (ref_driver.cpp:ref_search_hist).
    Returns (hist [N][N] with hist[parent][child], iterations)."""
    L = lib()
    s = _i32(samples)
    M, N = s.shape
    L.ref_release_cache()
    L.ref_set_seed(int(seed))
    hist = np.zeros((N, N), dtype=np.float64)
    it = L.ref_search_hist(N, M, _p(s), _p(_i32(perturbations)), int(max_parent_set), _p(_i32(parent_sizes)),
                           int(max_complexity), _p(_i32(num_cats)), _p(_pool_matrix(parents_pool, N)),
                           _p(_pool_matrix(fixed_parents, N)), int(score), int(weight), int(max_iter),
                           int(num_threads), int(bool(use_cache)), _p(hist))
    L.ref_release_cache()
    return hist, it


# ---------------------------------------------------------------- pairwise metrics ----
_PW_PATH = os.path.join(_HERE, "_ref", "libcatnet_ref_pairwise.so")
_pw = None
PAIRWISE_KINDS = {"entropy": 0, "kl": 1, "pearson": 2, "order": 3}


def pairwise_available():
    return os.path.exists(_PW_PATH)


def pairwise(kind, samples, perturbations=None):
    """The reference's own catnetEntropyPairwise / catnetKLpairwise / catnetPearsonPairwise / catnetEntropyOrder
    (catnet_entropy.cpp, compiled unmodified; ref_pairwise.cpp).  samples: [M][N] int32 1-based categories, no NA.
    Returns the R matrix `matrix(mat, N, N)` as numpy [N][N] with out[n1][n2] (= mat[n2*N + n1]); for "order" the
    1-based node order (0 where the reference's tie handling leaves a hole, utils.h:97-133)."""
    global _pw
    if _pw is None:
        if not pairwise_available():
            raise RuntimeError("oracle/_ref/libcatnet_ref_pairwise.so missing: run `make -C oracle ref` where /root/reference exists")
        _pw = C.CDLL(_PW_PATH)
        _pw.ref_pairwise.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _pw.ref_pairwise.restype = C.c_int
        _pw.ref_pairwise_last_error.restype = C.c_char_p
    s = _i32(samples)
    M, N = s.shape
    p = _i32(perturbations)
    out = np.zeros((N, N), dtype=np.float64)
    order = np.zeros(N, dtype=np.int32)
    rc = _pw.ref_pairwise(PAIRWISE_KINDS[kind], N, M, _p(s), _p(p), _p(out), _p(order))
    if rc != 0:
        raise RuntimeError("reference pairwise: " + _pw.ref_pairwise_last_error().decode())
    return order if kind == "order" else out.T.copy()


# ---------------------------------------------------------------- fixed network ----
def flatten_network(cats, parents, probs=None):
    """cats [N]; parents: list of lists of 0-based nodes in listed order (first = most significant table index);
    probs: list of flat tables (child fastest) or None -> the arrays the C entry points take"""
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
    flat = None
    if probs is not None:
        flat = np.concatenate([np.asarray(t, dtype=np.float64).ravel() for t in probs])
        assert len(flat) == off[-1]
    return cats, maxp, npar, pars, flat, off


NET_MODES = {"nodes": 0, "bysample": 1, "node": 2, "total": 3}


def net_loglik(mode, samples0, cats, parents, probs, perturbations=None):
    """CATNET::sampleLoglikVector ("nodes") / bySampleLoglikVector ("bysample") / sampleNodeLoglik per node ("node") /
    sampleLoglik ("total") of the reference, on a network given by (cats, parents, probs)"""
    L = lib()
    L.ref_net_loglik.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_void_p]
    s = _i32(samples0)
    M, N = s.shape
    cats, maxp, npar, pars, flat, off = flatten_network(cats, parents, probs)
    out = np.zeros(M if mode == "bysample" else N, dtype=np.float64)
    rc = L.ref_net_loglik(N, M, _p(s), _p(_i32(perturbations)), _p(cats), maxp, _p(npar), _p(pars), _p(flat), _p(off),
                          NET_MODES[mode], _p(out))
    assert rc == 0
    return out[0] if mode == "total" else out


def net_set_prob(samples0, cats, parents, perturbations=None):
    """cnSetProb through CATNET::setNodeSampleProb(.., bNormalize=1): (list of tables, loglik [N], sample sizes [N])"""
    L = lib()
    L.ref_net_set_prob.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 6
    s = _i32(samples0)
    M, N = s.shape
    cats, maxp, npar, pars, _, off = flatten_network(cats, parents)
    probs = np.zeros(off[-1], dtype=np.float64)
    ll = np.zeros(N, dtype=np.float64)
    sizes = np.zeros(N, dtype=np.int32)
    rc = L.ref_net_set_prob(N, M, _p(s), _p(_i32(perturbations)), _p(cats), maxp, _p(npar), _p(pars), _p(off), _p(probs),
                            _p(ll), _p(sizes))
    assert rc == 0
    return [probs[off[i]:off[i + 1]].copy() for i in range(N)], ll, sizes


# ---------------------------------------------------------------- cnSamples ----
_SM_PATH = os.path.join(_HERE, "_ref", "libcatnet_ref_samples.so")
_sm = None


def samples_available():
    return os.path.exists(_SM_PATH)


def net_samples(cats, parents, probs, num_samples, perturbations=None, na_rate=0.0, seed=1):
    """RCatnet::genSamples of the reference (rcatnet.cpp:524-650, compiled unmodified; ref_samples.cpp) under the injected
    splitmix64 uniform stream.  Returns ([M][N] int32 1-based, NA = NA_INT; uniforms consumed)."""
    global _sm
    if _sm is None:
        if not samples_available():
            raise RuntimeError("oracle/_ref/libcatnet_ref_samples.so missing: run `make -C oracle ref` where /root/reference exists")
        _sm = C.CDLL(_SM_PATH)
        _sm.ref_samples.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                    C.c_void_p, C.c_double, C.c_ulonglong, C.c_void_p]
        _sm.ref_samples.restype = C.c_long
    cats, maxp, npar, pars, flat, off = flatten_network(cats, parents, probs)
    N, M = len(cats), int(num_samples)
    out = np.zeros((M, N), dtype=np.int32)
    calls = _sm.ref_samples(N, _p(cats), maxp, _p(npar), _p(pars), _p(flat), _p(off), M, _p(_i32(perturbations)),
                            float(na_rate), int(seed), _p(out))
    assert calls >= 0
    return out, calls
