"""Runs catnet_b200/csrc/rshim.cpp without R: tests/rstub/minir.cpp implements the part of R's C API the shim uses,
and this module builds the SEXP arguments of a `.Call`, invokes the routine through the table the shim registers, and
turns the S4 catNetwork objects it returns into the dictionaries the other tests compare.

backend "mock": the C ABI is answered by the oracle (tests/rstub/mock_abi.cpp) -- runs anywhere;
backend "real": the shim is linked against catnet_b200/libcatnet_b200.so -- needs a B200."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rstub")
INTSXP, REALSXP, STRSXP, VECSXP, S4SXP, NILSXP = 13, 14, 16, 19, 25, 0
_libs = {}


def lib(backend):
    if backend not in _libs:
        path = os.path.join(HERE, "_build", "librshim_%s.so" % backend)
        # pytest-xdist workers share the tree: one of them (re)builds at a time, the others wait and find it up to date
        import fcntl
        os.makedirs(os.path.join(HERE, "_build"), exist_ok=True)
        with open(os.path.join(HERE, "_build", ".lock"), "w") as lock:
            fcntl.flock(lock, fcntl.LOCK_EX)
            rc = subprocess.call(["make", "-s", "-C", HERE, "_build/librshim_%s.so" % backend])
        if rc != 0 and not os.path.exists(path):     # a prebuilt library (from __graft_entry__.build()) is good enough
            raise RuntimeError("cannot build %s" % path)
        L = C.CDLL(path)
        vp = C.c_void_p
        for name, res, args in [("mr_nil", vp, []), ("mr_int_vector", vp, [vp, C.c_int]), ("mr_real_vector", vp, [vp, C.c_int]),
                                ("mr_logical", vp, [C.c_int]), ("mr_string", vp, [C.c_char_p]), ("mr_list", vp, [C.c_int]),
                                ("mr_list_set", None, [vp, C.c_int, vp]), ("mr_set_dim", vp, [vp, C.c_int, C.c_int]),
                                ("mr_type", C.c_int, [vp]), ("mr_length", C.c_int, [vp]), ("mr_elt", vp, [vp, C.c_int]),
                                ("mr_chars", C.c_char_p, [vp]), ("mr_ints", C.POINTER(C.c_int), [vp]),
                                ("mr_reals", C.POINTER(C.c_double), [vp]), ("mr_slot", vp, [vp, C.c_char_p]),
                                ("mr_last_error", C.c_char_p, []), ("mr_warnings", C.c_char_p, []), ("mr_clear", None, []),
                                ("mr_set_rng", None, [C.c_ulonglong]), ("mr_new_s4", vp, [C.c_char_p]),
                                ("mr_set_slot", None, [vp, C.c_char_p, vp]), ("mr_category_names", vp, [C.c_int]),
                                ("mr_set_string", None, [vp, C.c_int, C.c_char_p]), ("mr_call", vp, [C.c_char_p, C.POINTER(vp), C.c_int]),
                                ("mr_fail_alloc_after", None, [C.c_long]), ("mr_run_finalizers", C.c_int, [])]:
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _libs[backend] = L
    return _libs[backend]


class RError(Exception):
    pass


class R:
    """a tiny R: value constructors and .Call"""

    def __init__(self, backend):
        self.L = lib(backend)
        self.NULL = self.L.mr_nil()

    def integer(self, v):
        a = np.ascontiguousarray(v, dtype=np.int32).ravel()
        return self.L.mr_int_vector(a.ctypes.data, len(a))

    def numeric(self, v):
        a = np.ascontiguousarray(v, dtype=np.float64).ravel()
        return self.L.mr_real_vector(a.ctypes.data, len(a))

    def logical(self, v):
        return self.L.mr_logical(int(bool(v)))

    def string(self, v):
        return self.L.mr_string(v.encode())

    def list(self, items):
        l = self.L.mr_list(len(items))
        for i, it in enumerate(items):
            self.L.mr_list_set(l, i, it)
        return l

    def matrix(self, sexp, nrow, ncol):
        return self.L.mr_set_dim(sexp, nrow, ncol)

    def data_matrix(self, samples, real=False):
        """samples [M][N] (numpy, sample-major) -> R's nodes x samples matrix: the same bytes, column-major"""
        m, n = samples.shape
        return self.matrix(self.numeric(samples) if real else self.integer(samples), n, m)

    def int_lists(self, lists):
        return self.NULL if lists is None else self.list([self.integer(v) for v in lists])

    def nested_probs(self, flat, parent_cats, cc):
        """flat table (first listed parent most significant, child fastest) -> R's nested lists of numeric vectors"""
        flat = np.asarray(flat, dtype=np.float64)
        if not parent_cats:
            return self.numeric(flat[:cc])
        step = len(flat) // parent_cats[0]
        return self.list([self.nested_probs(flat[j * step:(j + 1) * step], parent_cats[1:], cc) for j in range(parent_cats[0])])

    def new_catnetwork(self, cats, parents, probs=None):
        """S4 catNetwork with the slots RCatnet::RCatnet(SEXP) reads; parents: lists of 0-based nodes in listed order"""
        n = len(cats)
        obj = self.L.mr_new_s4(b"catNetwork")
        self.L.mr_set_slot(obj, b"numnodes", self.integer([n]))
        self.L.mr_set_slot(obj, b"nodes", self.L.mr_category_names(n))
        self.L.mr_set_slot(obj, b"parents", self.list([self.integer([q + 1 for q in p]) if len(p) else self.NULL for p in parents]))
        self.L.mr_set_slot(obj, b"categories", self.list([self.L.mr_category_names(int(c)) for c in cats]))
        self.L.mr_set_slot(obj, b"maxParents", self.integer([max([len(p) for p in parents] + [0])]))
        self.L.mr_set_slot(obj, b"maxCategories", self.integer([int(max(cats))]))
        self.L.mr_set_slot(obj, b"complexity", self.integer([int(sum(
            (int(cats[i]) - 1) * int(np.prod([cats[q] for q in parents[i]], dtype=np.int64)) for i in range(n)))]))
        self.L.mr_set_slot(obj, b"likelihood", self.numeric([0.0]))
        self.L.mr_set_slot(obj, b"nodeSampleSizes", self.integer([0] * n))
        if probs is None:
            probs = [np.full(int(cats[i]) * int(np.prod([cats[q] for q in parents[i]], dtype=np.int64)), 1.0 / cats[i])
                     for i in range(n)]
        self.L.mr_set_slot(obj, b"probabilities", self.list([
            self.nested_probs(probs[i], [int(cats[q]) for q in parents[i]], int(cats[i])) for i in range(n)]))
        return obj

    def call(self, name, *args):
        self.L.mr_clear()
        arr = (C.c_void_p * max(len(args), 1))(*args)
        out = self.L.mr_call(name.encode(), arr, len(args))
        if not out:
            raise RError(self.L.mr_last_error().decode())
        return out

    def warnings(self):
        return self.L.mr_warnings().decode()

    # ---- reading values back ----
    def is_null(self, s):
        return s == self.NULL or self.L.mr_type(s) == NILSXP

    def ints(self, s):
        return [int(self.L.mr_ints(s)[i]) for i in range(self.L.mr_length(s))]

    def reals(self, s):
        return np.array([self.L.mr_reals(s)[i] for i in range(self.L.mr_length(s))], dtype=np.float64)

    def strings(self, s):
        return [self.L.mr_chars(self.L.mr_elt(s, i)).decode() for i in range(self.L.mr_length(s))]

    def elts(self, s):
        return [self.L.mr_elt(s, i) for i in range(self.L.mr_length(s))]

    def slot(self, s, name):
        v = self.L.mr_slot(s, name.encode())
        assert v, "no slot %s" % name
        return v

    def flatten_probs(self, s):
        """nested probability lists (first listed parent outermost) -> flat table, leaves = numeric vectors"""
        if self.L.mr_type(s) == REALSXP:
            return self.reals(s)
        return np.concatenate([self.flatten_probs(e) for e in self.elts(s)])

    def catnetwork(self, s):
        """an S4 catNetwork -> dict with the fields helpers.compare_search reads (parents as 0-based positions)"""
        assert self.L.mr_type(s) == S4SXP and self.L.mr_chars(s).decode() == "catNetwork"
        n = self.ints(self.slot(s, "numnodes"))[0]
        parents = [[] if self.is_null(p) else [v - 1 for v in self.ints(p)] for p in self.elts(self.slot(s, "parents"))]
        cats = [self.strings(c) for c in self.elts(self.slot(s, "categories"))]
        probs = [self.flatten_probs(p) for p in self.elts(self.slot(s, "probabilities"))]
        assert len(parents) == n and len(cats) == n and len(probs) == n
        return {"numnodes": n, "nodes": self.strings(self.slot(s, "nodes")), "parents": parents, "categories": cats,
                "probs": probs, "complexity": self.ints(self.slot(s, "complexity"))[0],
                "loglik": float(self.reals(self.slot(s, "likelihood"))[0]),
                "sample_sizes": self.ints(self.slot(s, "nodeSampleSizes")),
                "maxParents": self.ints(self.slot(s, "maxParents"))[0],
                "maxCategories": self.ints(self.slot(s, "maxCategories"))[0]}


def optimal_nets_for_order(r, samples, max_parent_set, max_complexity, order=None, perturbations=None,
                           parent_sizes=None, num_cats=None, parents_pool=None, fixed_parents=None, edge_prob=None,
                           real_args=True):
    """.Call("ccnOptimalNetsForOrder", ...) with the arguments R/catnet.search.R:205-214 passes; numbers go in as
    doubles (what R hands over unless the caller wrote 2L) when real_args"""
    s = np.ascontiguousarray(samples, dtype=np.int32)
    m, n = s.shape
    num = r.numeric if real_args else r.integer
    cat_indices = r.NULL if num_cats is None else r.list([r.integer(np.arange(1, c + 1)) for c in num_cats])
    edge = r.NULL if edge_prob is None else r.matrix(r.numeric(np.asarray(edge_prob, dtype=np.float64).T), n, n)
    out = r.call("ccnOptimalNetsForOrder", r.data_matrix(s),
                 r.NULL if perturbations is None else r.data_matrix(np.asarray(perturbations)),
                 num([max_parent_set]), r.NULL if parent_sizes is None else num(parent_sizes), num([max_complexity]),
                 r.NULL if order is None else num(order), cat_indices, r.int_lists(parents_pool),
                 r.int_lists(fixed_parents), edge, r.logical(False), r.logical(False))
    if r.is_null(out):
        return []
    return [r.catnetwork(e) for e in r.elts(out)]


def mini_r_uniform(state):
    """the miniature R's unif_rand (splitmix64, tests/rstub/minir.cpp) after mr_set_rng(state): first draw"""
    mask = (1 << 64) - 1
    z = (state + 0x9e3779b97f4a7c15) & mask
    z = ((z ^ (z >> 30)) * 0xbf58476d1ce4e5b9) & mask
    z = ((z ^ (z >> 27)) * 0x94d049bb133111eb) & mask
    z ^= z >> 31
    return (z >> 11) * (1.0 / 9007199254740992.0)
