"""ctypes binding of catnet_b200/libcatnet_b200.so (the C ABI of include/catnet_b200.h).

The library is the product; this module only marshals numpy arrays into it.  It fails loudly when the library is
missing or was built without the sm_100a kernels -- there is no Python or CPU fallback for the scoring path.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcatnet_b200.so")

CNB_OK, CNB_ERR_MEM, CNB_ERR_PARAM, CNB_ERR_PROC, CNB_ERR_CUDA = 0, -10, -20, -30, -40
NA_INTEGER = -2147483648
MAXP = 8


class CnbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("catnet_b200 error %d: %s" % (code, msg))
        self.code = code


class Params(C.Structure):
    _fields_ = [("num_nodes", C.c_int32), ("num_samples", C.c_int32), ("samples", C.c_void_p),
                ("perturbations", C.c_void_p), ("max_parent_set", C.c_int32), ("parent_sizes", C.c_void_p),
                ("max_complexity", C.c_int32), ("order", C.c_void_p), ("num_cats", C.c_void_p),
                ("parents_pool", C.c_void_p), ("fixed_parents", C.c_void_p), ("edge_prob", C.c_void_p),
                ("echo", C.c_int32), ("device", C.c_int32), ("num_devices", C.c_int32)]


class SaParams(C.Structure):
    _fields_ = [("select_mode", C.c_int32), ("start_order", C.c_void_p), ("temp_start", C.c_double),
                ("temp_cool_fact", C.c_double), ("temp_check_orders", C.c_int32), ("max_iter", C.c_int32),
                ("order_shuffles", C.c_double), ("stop_diff", C.c_double), ("num_threads", C.c_int32),
                ("use_cache", C.c_int32), ("dir_prob", C.c_void_p), ("seed", C.c_uint64), ("unif", C.c_void_p),
                ("unif_ctx", C.c_void_p)]


class HistParams(C.Structure):
    _fields_ = [("score", C.c_int32), ("weight", C.c_int32), ("max_iter", C.c_int32), ("num_threads", C.c_int32),
                ("use_cache", C.c_int32), ("seed", C.c_uint64), ("unif", C.c_void_p), ("unif_ctx", C.c_void_p)]


UNIF_FN = C.CFUNCTYPE(C.c_double, C.c_void_p)     # double (*unif)(void *): cnb_sa_params.unif / cnb_hist_params.unif


class Network(C.Structure):
    _fields_ = [("num_nodes", C.c_int32), ("max_parents", C.c_int32), ("num_cats", C.c_void_p),
                ("num_parents", C.c_void_p), ("parents", C.c_void_p), ("probs", C.c_void_p),
                ("prob_offsets", C.c_void_p)]


class Timing(C.Structure):
    _fields_ = [("pack_ms", C.c_float), ("plan_ms", C.c_float), ("score_ms", C.c_float), ("select_ms", C.c_float),
                ("dp_ms", C.c_float), ("collect_ms", C.c_float), ("total_ms", C.c_float),
                ("score_ms_by_parents", C.c_float * 8), ("families_by_parents", C.c_int64 * 8),
                ("num_families", C.c_int64), ("algorithmic_bytes", C.c_int64), ("kernel_launches", C.c_int32),
                ("fast_path", C.c_int32), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("tensor_tiles", C.c_int64), ("tensor_ms", C.c_float), ("tensor_kind", C.c_int32),
                ("tensor_macs", C.c_int64), ("tensor3_tiles", C.c_int64)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k not in ("score_ms_by_parents", "families_by_parents")}
        d["score_ms_by_parents"] = list(self.score_ms_by_parents)
        d["families_by_parents"] = list(self.families_by_parents)
        return d


# every symbol include/catnet_b200.h declares (tests/test_abi.py checks the library exports them all)
SYMBOLS = [
    "cnb_search_order", "cnb_search_sa", "cnb_search_hist", "cnb_ctx_search_hist", "cnb_release_cache", "cnb_set_seed", "cnb_last_error", "cnb_version",
    "cnb_ctx_create", "cnb_ctx_destroy", "cnb_ctx_set_samples", "cnb_ctx_set_samples_dev", "cnb_ctx_num_cats",
    "cnb_ctx_search_order", "cnb_ctx_search_sa", "cnb_ctx_plan", "cnb_ctx_score_shard", "cnb_ctx_finish",
    "cnb_ctx_select_dp", "cnb_ctx_collect", "cnb_ctx_set_stream", "cnb_ctx_timing", "cnb_ctx_force_generic", "cnb_family_scores", "cnb_tile_selftest", "cnb_device_log", "cnb_host_log",
    "cnb_unrank_combination", "cnb_num_families", "cnb_result_num_networks", "cnb_result_num_nodes",
    "cnb_result_network", "cnb_result_parents", "cnb_result_order", "cnb_result_cats", "cnb_result_counts", "cnb_result_node",
    "cnb_result_sa_info", "cnb_result_sa_trace", "cnb_result_merge", "cnb_result_free",
    "cnb_pairwise", "cnb_entropy_order", "cnb_ctx_pairwise", "cnb_ctx_entropy_order",
    "cnb_loglik", "cnb_node_loglik", "cnb_set_prob", "cnb_ctx_loglik", "cnb_ctx_set_prob",
    "cnb_samples", "cnb_samples_dev", "cnb_ctx_set_samples_dev8", "cnb_tile3_selftest", "cnb_host_entropy_order",
    "cnb_mctx_create", "cnb_mctx_destroy", "cnb_mctx_num_devices", "cnb_mctx_ctx", "cnb_mctx_set_samples",
    "cnb_mctx_search_order", "cnb_mctx_search_light", "cnb_mctx_collect", "cnb_mctx_search_sa", "cnb_mctx_search_hist",
    "cnb_mctx_mark", "cnb_mctx_elapsed_ms", "cnb_mctx_active", "cnb_ctx_search_scores",
    "cnb_ctx_upload_rows8", "cnb_tile_selftest_full", "cnb_last_call_timing",
    "cnb_tile_selftest_geo", "cnb_ctx_allow_scatter", "cnb_ctx_scattered", "cnb_ctx_upload_rows8_on", "cnb_ctx_stream_begin", "cnb_ctx_stream_slice", "cnb_ctx_stream_end",
    "cnb_cache_trace",
]

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("catnet_b200: %s not built (run `python -c 'import __graft_entry__ as g; g.build()'` or "
                          "`make -C catnet_b200/csrc`); there is no fallback implementation" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    P = C.POINTER
    sig = {
        "cnb_search_order": (i32, [P(Params), P(vp)]),
        "cnb_search_sa": (i32, [P(Params), P(SaParams), P(vp)]),
        "cnb_release_cache": (None, []),
        "cnb_set_seed": (None, [i32]),
        "cnb_last_error": (C.c_char_p, []),
        "cnb_version": (C.c_char_p, []),
        "cnb_ctx_create": (i32, [i32, P(vp)]),
        "cnb_ctx_destroy": (None, [vp]),
        "cnb_ctx_set_samples": (i32, [vp, i32, i32, vp, vp, vp]),
        "cnb_ctx_set_samples_dev": (i32, [vp, i32, i32, vp, vp, vp]),
        "cnb_ctx_set_samples_dev8": (i32, [vp, i32, i32, vp, vp, vp]),
        "cnb_ctx_num_cats": (i32, [vp, vp]),
        "cnb_ctx_search_order": (i32, [vp, P(Params), P(vp)]),
        "cnb_ctx_search_sa": (i32, [vp, P(Params), P(SaParams), P(vp)]),
        "cnb_ctx_search_hist": (i32, [vp, P(Params), P(HistParams), vp, P(i32)]),
        "cnb_search_hist": (i32, [P(Params), P(HistParams), vp, P(i32)]),
        "cnb_ctx_plan": (i32, [vp, P(Params), i32, i32, P(i64), P(i64)]),
        "cnb_ctx_score_shard": (i32, [vp, vp]),
        "cnb_ctx_finish": (i32, [vp, vp, P(vp)]),
        "cnb_ctx_select_dp": (i32, [vp, vp]),
        "cnb_ctx_collect": (i32, [vp, P(vp)]),
        "cnb_ctx_set_stream": (i32, [vp, vp]),
        "cnb_ctx_timing": (i32, [vp, P(Timing)]),
        "cnb_last_call_timing": (i32, [P(Timing)]),
        "cnb_ctx_force_generic": (i32, [vp, i32]),
        "cnb_tile_selftest": (C.c_int64, [i32]),
        "cnb_cache_trace": (C.c_int64, [P(Params), vp, vp, i32, vp, vp, i32, i32, vp, vp]),
        "cnb_tile_selftest_full": (C.c_int64, [i32]),
        "cnb_tile3_selftest": (C.c_int64, [i32, i32]),
        "cnb_tile_selftest_geo": (C.c_int64, [i32, i32]),
        "cnb_ctx_allow_scatter": (i32, [vp, i32]),
        "cnb_ctx_scattered": (i32, [vp]),
        "cnb_host_entropy_order": (None, [vp, i32, vp]),
        "cnb_family_scores": (i32, [vp, i32, i32, vp, vp, i32, vp, vp, vp, i32]),
        "cnb_device_log": (i32, [i32, vp, i64, vp]),
        "cnb_host_log": (None, [vp, i64, vp]),
        "cnb_unrank_combination": (i32, [i32, i32, i64, vp]),
        "cnb_num_families": (i64, [P(Params)]),
        "cnb_result_num_networks": (i32, [vp]),
        "cnb_result_num_nodes": (i32, [vp]),
        "cnb_result_network": (i32, [vp, i32, P(i32), P(dbl)]),
        "cnb_result_parents": (i32, [vp, i32, vp, vp, i32]),
        "cnb_result_order": (i32, [vp, i32, vp]),
        "cnb_result_cats": (i32, [vp, i32, vp]),
        "cnb_result_counts": (i32, [vp, i32, i32, vp, i32]),
        "cnb_result_node": (i32, [vp, i32, i32, P(dbl), P(dbl), P(i32), vp, i32]),
        "cnb_result_sa_info": (i32, [vp, vp, P(i32), P(i32)]),
        "cnb_result_sa_trace": (i32, [vp, vp, vp, i32]),
        "cnb_result_merge": (i32, [vp, vp]),
        "cnb_result_free": (None, [vp]),
        "cnb_pairwise": (i32, [i32, i32, i32, vp, vp, i32, vp]),
        "cnb_entropy_order": (i32, [i32, i32, vp, vp, i32, vp]),
        "cnb_ctx_pairwise": (i32, [vp, i32, vp]),
        "cnb_ctx_entropy_order": (i32, [vp, vp]),
        "cnb_loglik": (i32, [P(Network), i32, vp, vp, i32, i32, vp]),
        "cnb_node_loglik": (i32, [P(Network), i32, vp, i32, vp, vp, i32, vp]),
        "cnb_set_prob": (i32, [P(Network), i32, vp, vp, i32, vp, vp, vp]),
        "cnb_ctx_loglik": (i32, [vp, P(Network), i32, i32, vp, vp]),
        "cnb_ctx_set_prob": (i32, [vp, P(Network), vp, vp, vp]),
        "cnb_samples": (i32, [P(Network), i32, vp, dbl, C.c_uint64, i32, vp]),
        "cnb_samples_dev": (i32, [P(Network), i32, vp, dbl, C.c_uint64, i32, vp]),
        "cnb_ctx_search_scores": (i32, [vp, vp, i32, i32, vp, vp, vp]),
        "cnb_ctx_upload_rows8": (i32, [vp, i32, vp, i64, i64, vp, i32]),
        "cnb_ctx_upload_rows8_on": (i32, [vp, i32, vp, i64, i64, vp, i32, vp, i64, i64]),
        "cnb_ctx_stream_begin": (i32, [vp, P(Params), i32, i32, P(i64), P(i64), P(i32), vp, i32]),
        "cnb_ctx_stream_slice": (i32, [vp, i32, vp, vp]),
        "cnb_ctx_stream_end": (i32, [vp, vp, P(i32)]),
        "cnb_mctx_create": (i32, [i32, i32, P(vp)]),
        "cnb_mctx_destroy": (None, [vp]),
        "cnb_mctx_num_devices": (i32, [vp]),
        "cnb_mctx_active": (i32, [vp]),
        "cnb_mctx_ctx": (vp, [vp, i32]),
        "cnb_mctx_set_samples": (i32, [vp, i32, i32, vp, vp, vp]),
        "cnb_mctx_search_order": (i32, [vp, P(Params), P(vp)]),
        "cnb_mctx_search_light": (i32, [vp, P(Params)]),
        "cnb_mctx_collect": (i32, [vp, P(vp)]),
        "cnb_mctx_search_sa": (i32, [vp, P(Params), P(SaParams), P(vp)]),
        "cnb_mctx_search_hist": (i32, [vp, P(Params), P(HistParams), vp, P(i32)]),
        "cnb_mctx_mark": (i32, [vp, i32]),
        "cnb_mctx_elapsed_ms": (i32, [vp, P(C.c_float)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(rc):
    if rc != CNB_OK:
        raise CnbError(rc, lib().cnb_last_error().decode("utf-8", "replace"))


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def pool_matrix(pool, n):
    """list (per node, label order) of lists of 1-based labels -> [N][N] int32 rows, 0-padded; None passes through."""
    if pool is None:
        return None
    m = np.zeros((n, n), dtype=np.int32)
    for i, row in enumerate(pool):
        if row is None:
            continue
        row = [int(v) for v in np.atleast_1d(row)]
        if len(row) > n:
            raise ValueError("pool row longer than the number of nodes")
        m[i, :len(row)] = row
    return m


class Keep:
    """keeps numpy buffers alive while a Params struct points at them"""

    def __init__(self):
        self.refs = []

    def __call__(self, a):
        if a is not None:
            self.refs.append(a)
        return ptr(a)


def make_params(keep, n, m, samples=None, perturbations=None, max_parent_set=1, parent_sizes=None, max_complexity=0,
                order=None, num_cats=None, parents_pool=None, fixed_parents=None, edge_prob=None, echo=False,
                device=0, num_devices=0):
    p = Params()
    p.num_nodes, p.num_samples = n, m
    p.samples = keep(i32(samples))
    p.perturbations = keep(i32(perturbations))
    p.max_parent_set = int(max_parent_set)
    p.parent_sizes = keep(i32(parent_sizes))
    p.max_complexity = int(max_complexity)
    p.order = keep(i32(order))
    p.num_cats = keep(i32(num_cats))
    p.parents_pool = keep(pool_matrix(parents_pool, n))
    p.fixed_parents = keep(pool_matrix(fixed_parents, n))
    # R hands the matrix edgeProb[child, parent] column-major: flat index parent*N + child
    p.edge_prob = keep(None if edge_prob is None else f64(np.asarray(edge_prob, dtype=np.float64).T))
    p.echo = int(bool(echo))
    p.device, p.num_devices = int(device), int(num_devices)
    return p


def read_result(h, want_probs=True, want_counts=False):
    """cnb_result* -> list of dicts (ascending complexity), the slots of the reference's catNetwork objects."""
    L = lib()
    n = L.cnb_result_num_nodes(h)
    out = []
    npar = np.zeros(n, dtype=np.int32)
    pars = np.zeros((n, MAXP), dtype=np.int32)
    order = np.zeros(n, dtype=np.int32)
    probs = np.zeros(1 << 16, dtype=np.float64)
    counts = np.zeros(1 << 16, dtype=np.int32)
    for i in range(L.cnb_result_num_networks(h)):
        cx, ll = C.c_int32(), C.c_double()
        check(L.cnb_result_network(h, i, C.byref(cx), C.byref(ll)))
        check(L.cnb_result_parents(h, i, ptr(npar), ptr(pars), MAXP))
        check(L.cnb_result_order(h, i, ptr(order)))
        net = {"complexity": cx.value, "loglik": ll.value, "order": order.copy(),
               "parents": [[int(v) for v in pars[k, :npar[k]]] for k in range(n)],
               "node_loglik": [], "node_prior": [], "sample_sizes": [], "probs": [], "counts": []}
        for k in range(n):
            nl, pr, ss = C.c_double(), C.c_double(), C.c_int32()
            ts = L.cnb_result_node(h, i, k, C.byref(nl), C.byref(pr), C.byref(ss), ptr(probs) if want_probs else None,
                                   len(probs))
            if ts < 0:
                check(ts)
            net["node_loglik"].append(nl.value)
            net["node_prior"].append(pr.value)
            net["sample_sizes"].append(ss.value)
            if want_probs:
                net["probs"].append(probs[:ts].copy())
            if want_counts:
                tc = L.cnb_result_counts(h, i, k, ptr(counts), len(counts))
                net["counts"].append(counts[:tc].copy())
        out.append(net)
    return out


class Context:
    """cnb_ctx: a sample matrix packed and resident on one GPU."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        check(lib().cnb_ctx_create(int(device), C.byref(self.h)))
        self.n = self.m = 0

    def close(self):
        if self.h:
            lib().cnb_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_samples(self, samples, perturbations=None, num_cats=None):
        s = i32(samples)
        self.m, self.n = s.shape
        check(lib().cnb_ctx_set_samples(self.h, self.n, self.m, ptr(s), ptr(i32(perturbations)), ptr(i32(num_cats))))

    def set_samples_dev(self, samples_dev_ptr, n, m, perturbations_dev_ptr=None, num_cats=None):
        self.n, self.m = n, m
        check(lib().cnb_ctx_set_samples_dev(self.h, n, m, samples_dev_ptr, perturbations_dev_ptr, ptr(i32(num_cats))))

    def set_samples_dev8(self, samples_dev_ptr, n, m, perturbations_dev_ptr=None, num_cats=None):
        """device matrix [M][N] of bytes: 0 = missing, 1..254 = category"""
        self.n, self.m = n, m
        check(lib().cnb_ctx_set_samples_dev8(self.h, n, m, samples_dev_ptr, perturbations_dev_ptr, ptr(i32(num_cats))))

    def upload_rows8(self, samples, row0, row1, dst_dev_ptr, max_threads=0):
        """rows [row0, row1) of the host int32 matrix -> bytes at dst_dev_ptr + row0 * N (first step of a sharded upload)"""
        s = i32(samples)
        check(lib().cnb_ctx_upload_rows8(self.h, s.shape[1], ptr(s), int(row0), int(row1), dst_dev_ptr, int(max_threads)))

    def upload_rows8_on(self, samples, row0, row1, dst_dev_ptr, max_threads, cuda_stream, stage_offset, stage_capacity):
        """the same on the caller's stream, staged at stage_offset (values) of a pinned buffer of stage_capacity values"""
        s = i32(samples)
        check(lib().cnb_ctx_upload_rows8_on(self.h, s.shape[1], ptr(s), int(row0), int(row1), dst_dev_ptr, int(max_threads),
                                            int(cuda_stream) if cuda_stream else None, int(stage_offset), int(stage_capacity)))

    def stream_begin(self, rank, world, n, m, **kw):
        """plan a streamed search (samples still on the host) -> (seg_len, num_families, slice row bounds); no bounds = the
        problem does not qualify, use the resident calls.  kw as for plan() plus num_cats (required)."""
        self.n, self.m = n, m
        keep = Keep()
        p = make_params(keep, n, m, **kw)
        seg, nf, ns = C.c_int64(), C.c_int64(), C.c_int32()
        rows = np.zeros(65, dtype=np.int64)
        check(lib().cnb_ctx_stream_begin(self.h, C.byref(p), int(rank), int(world), C.byref(seg), C.byref(nf), C.byref(ns), ptr(rows), 65))
        return seg.value, nf.value, [int(r) for r in rows[:ns.value + 1]] if ns.value else []

    def stream_slice(self, k, bytes_dev_ptr, ready_event=None):
        check(lib().cnb_ctx_stream_slice(self.h, int(k), bytes_dev_ptr, int(ready_event) if ready_event else None))

    def stream_end(self, scores_dev_ptr):
        ok = C.c_int32()
        check(lib().cnb_ctx_stream_end(self.h, scores_dev_ptr, C.byref(ok)))
        return bool(ok.value)

    def num_cats(self):
        out = np.zeros(self.n, dtype=np.int32)
        check(lib().cnb_ctx_num_cats(self.h, ptr(out)))
        return out

    def force_generic(self, on=True):
        check(lib().cnb_ctx_force_generic(self.h, int(on)))

    def timing(self):
        t = Timing()
        check(lib().cnb_ctx_timing(self.h, C.byref(t)))
        return t.as_dict()

    def search_order_raw(self, **kw):
        keep = Keep()
        p = make_params(keep, self.n, self.m, **kw)
        h = C.c_void_p()
        check(lib().cnb_ctx_search_order(self.h, C.byref(p), C.byref(h)))
        return h

    def search_order(self, want_probs=True, want_counts=False, **kw):
        h = self.search_order_raw(**kw)
        try:
            return read_result(h, want_probs, want_counts)
        finally:
            lib().cnb_result_free(h)

    def plan(self, rank, world, **kw):
        keep = Keep()
        p = make_params(keep, self.n, self.m, **kw)
        seg, nf = C.c_int64(), C.c_int64()
        check(lib().cnb_ctx_plan(self.h, C.byref(p), int(rank), int(world), C.byref(seg), C.byref(nf)))
        return seg.value, nf.value

    def score_shard(self, scores_dev_ptr):
        check(lib().cnb_ctx_score_shard(self.h, scores_dev_ptr))

    def allow_scatter(self, mode):
        """0 never, 1 all ranks share one score buffer, 2 own buffers combined by an element-wise MAX (dist.gather_scores)"""
        check(lib().cnb_ctx_allow_scatter(self.h, int(mode)))

    def scattered(self):
        return bool(lib().cnb_ctx_scattered(self.h))

    def select_dp(self, scores_dev_ptr):
        check(lib().cnb_ctx_select_dp(self.h, scores_dev_ptr))

    def collect_raw(self):
        h = C.c_void_p()
        check(lib().cnb_ctx_collect(self.h, C.byref(h)))
        return h

    def set_stream(self, cuda_stream):
        """launch on the caller's stream.  torch reports its default stream as handle 0, which the C ABI reads as
        "the context's own (non-blocking) stream": pass CUDA's explicit legacy-default-stream handle instead, so
        that the library's kernels and the caller's copies / NCCL collectives stay ordered."""
        CUDA_STREAM_LEGACY = 1
        check(lib().cnb_ctx_set_stream(self.h, int(cuda_stream) if cuda_stream else CUDA_STREAM_LEGACY))

    def finish_raw(self, scores_dev_ptr):
        h = C.c_void_p()
        check(lib().cnb_ctx_finish(self.h, scores_dev_ptr, C.byref(h)))
        return h

    def finish(self, scores_dev_ptr, want_probs=True):
        h = self.finish_raw(scores_dev_ptr)
        try:
            return read_result(h, want_probs)
        finally:
            lib().cnb_result_free(h)

    def family_scores(self, children, parents, want_counts=True, force_generic=False):
        """children [F], parents [F][d] (0-based labels) -> (counts int32 [F][stride] or None, loglik [F], ncount [F])"""
        ch = i32(children)
        pa = i32(parents).reshape(len(ch), -1) if len(ch) else np.zeros((0, 0), dtype=np.int32)
        d = pa.shape[1]
        cats = int(self.num_cats().max())
        stride = cats ** (d + 1)
        counts = np.zeros((len(ch), stride), dtype=np.int32) if want_counts else None
        ll = np.zeros(len(ch), dtype=np.float64)
        nc = np.zeros(len(ch), dtype=np.int32)
        check(lib().cnb_family_scores(self.h, len(ch), d, ptr(ch), ptr(pa), stride, ptr(counts), ptr(ll), ptr(nc),
                                      int(force_generic)))
        return counts, ll, nc

    def search_scores(self, scores_dev_ptr, children_pos, parents_pos):
        """scores the last score_shard wrote for explicit families (order positions, parents ascending)"""
        ch = i32(children_pos)
        pa = i32(parents_pos).reshape(len(ch), -1)
        out = np.zeros(len(ch), dtype=np.float64)
        check(lib().cnb_ctx_search_scores(self.h, scores_dev_ptr, len(ch), pa.shape[1], ptr(ch), ptr(pa), ptr(out)))
        return out

    def pairwise(self, kind):
        """kind "entropy" | "kl" | "pearson" -> numpy [N][N] laid out like R's matrix(mat, N, N): out[n1][n2]"""
        out = np.zeros((self.n, self.n), dtype=np.float64)
        check(lib().cnb_ctx_pairwise(self.h, PAIRWISE_KINDS[kind], ptr(out)))
        return out.T.copy()

    def entropy_order(self):
        out = np.zeros(self.n, dtype=np.int32)
        check(lib().cnb_ctx_entropy_order(self.h, ptr(out)))
        return out

    def loglik(self, cats, parents, probs, mode="nodes", nodes=None):
        """likelihood of the resident data under a given network (parents: lists of 0-based nodes in listed order;
        probs: list of flat tables).  mode "nodes" -> [N] (or the listed 1-based `nodes`), "bysample" -> [M]"""
        keep = Keep()
        net, off = make_network(keep, cats, parents, probs)
        sel = i32(nodes)
        nsel = 0 if sel is None else len(sel)
        out = np.zeros(self.m if mode == "bysample" else (nsel or self.n), dtype=np.float64)
        check(lib().cnb_ctx_loglik(self.h, C.byref(net), LOGLIK_MODES[mode], nsel, ptr(sel), ptr(out)))
        return out

    def set_prob(self, cats, parents):
        """cnSetProb on the resident data -> (list of tables, loglik [N], sample sizes [N])"""
        keep = Keep()
        net, off = make_network(keep, cats, parents, None)
        probs = np.zeros(off[-1], dtype=np.float64)
        ll = np.zeros(self.n, dtype=np.float64)
        sizes = np.zeros(self.n, dtype=np.int32)
        check(lib().cnb_ctx_set_prob(self.h, C.byref(net), ptr(probs), ptr(ll), ptr(sizes)))
        return [probs[off[k]:off[k + 1]].copy() for k in range(self.n)], ll, sizes

    def search_hist(self, hist, **kw):
        keep = Keep()
        p = make_params(keep, self.n, self.m, **kw)
        hp = make_hist_params(**hist)
        out = np.zeros((self.n, self.n), dtype=np.float64)
        it = C.c_int32()
        check(lib().cnb_ctx_search_hist(self.h, C.byref(p), C.byref(hp), ptr(out), C.byref(it)))
        return out, it.value

    def search_sa(self, sa, want_probs=False, **kw):
        keep = Keep()
        p = make_params(keep, self.n, self.m, **kw)
        s = make_sa_params(keep, **sa)
        h = C.c_void_p()
        check(lib().cnb_ctx_search_sa(self.h, C.byref(p), C.byref(s), C.byref(h)))
        try:
            return read_sa_result(h, self.n, want_probs)
        finally:
            lib().cnb_result_free(h)


class MultiContext:
    """cnb_mctx: one process driving several GPUs (the form an R session uses): the sample matrix on every device,
    candidate parent sets sharded, SA chains / histogram orders dealt out."""

    def __init__(self, first_device=0, num_devices=2):
        self.h = C.c_void_p()
        check(lib().cnb_mctx_create(int(first_device), int(num_devices), C.byref(self.h)))
        self.g = lib().cnb_mctx_num_devices(self.h)
        self.n = self.m = 0

    def close(self):
        if self.h:
            lib().cnb_mctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_samples(self, samples, perturbations=None, num_cats=None):
        s = i32(samples)
        self.m, self.n = s.shape
        check(lib().cnb_mctx_set_samples(self.h, self.n, self.m, ptr(s), ptr(i32(perturbations)), ptr(i32(num_cats))))

    def active(self):
        return lib().cnb_mctx_active(self.h)

    def timing(self, i=0):
        t = Timing()
        check(lib().cnb_ctx_timing(lib().cnb_mctx_ctx(self.h, i), C.byref(t)))
        return t.as_dict()

    def force_generic(self, on):
        for i in range(self.g):
            check(lib().cnb_ctx_force_generic(lib().cnb_mctx_ctx(self.h, i), int(on)))

    def search_light(self, **kw):
        keep = Keep()
        p = make_params(keep, self.n, self.m, **kw)
        check(lib().cnb_mctx_search_light(self.h, C.byref(p)))

    def collect_raw(self):
        h = C.c_void_p()
        check(lib().cnb_mctx_collect(self.h, C.byref(h)))
        return h

    def search_order(self, want_probs=True, **kw):
        keep = Keep()
        p = make_params(keep, self.n, self.m, **kw)
        h = C.c_void_p()
        check(lib().cnb_mctx_search_order(self.h, C.byref(p), C.byref(h)))
        try:
            return read_result(h, want_probs)
        finally:
            lib().cnb_result_free(h)

    def search_sa(self, sa, want_probs=False, **kw):
        keep = Keep()
        p = make_params(keep, self.n, self.m, **kw)
        s = make_sa_params(keep, **sa)
        h = C.c_void_p()
        check(lib().cnb_mctx_search_sa(self.h, C.byref(p), C.byref(s), C.byref(h)))
        try:
            return read_sa_result(h, self.n, want_probs)
        finally:
            lib().cnb_result_free(h)

    def search_hist(self, hist, **kw):
        keep = Keep()
        p = make_params(keep, self.n, self.m, **kw)
        hp = make_hist_params(**hist)
        out = np.zeros((self.n, self.n), dtype=np.float64)
        it = C.c_int32()
        check(lib().cnb_mctx_search_hist(self.h, C.byref(p), C.byref(hp), ptr(out), C.byref(it)))
        return out, it.value

    def mark(self, which):
        check(lib().cnb_mctx_mark(self.h, int(which)))

    def elapsed_ms(self):
        ms = C.c_float()
        check(lib().cnb_mctx_elapsed_ms(self.h, C.byref(ms)))
        return ms.value


LOGLIK_MODES = {"nodes": 0, "bysample": 1}


def make_network(keep, cats, parents, probs=None):
    """(cats [N], parents: list of lists of 0-BASED nodes in listed order, probs: list of flat tables or None)
    -> (Network struct, offsets).  The ABI takes 1-based parents, as the S4 object holds them."""
    cats = i32(cats)
    n = len(cats)
    maxp = max([len(p) for p in parents] + [1])
    npar = np.array([len(p) for p in parents], dtype=np.int32)
    pars = np.zeros((n, maxp), dtype=np.int32)
    off = np.zeros(n + 1, dtype=np.int64)
    for k, p in enumerate(parents):
        pars[k, :len(p)] = np.asarray(p, dtype=np.int32) + 1
        size = int(cats[k])
        for q in p:
            size *= int(cats[q]) if 0 <= q < n else 1       # the library reports the invalid parent
        off[k + 1] = off[k] + size
    flat = None
    if probs is not None:
        flat = f64(np.concatenate([np.asarray(t, dtype=np.float64).ravel() for t in probs]))
        if len(flat) < off[-1]:      # a malformed network: the library reports it; never hand it a short buffer
            flat = np.concatenate([flat, np.zeros(off[-1] - len(flat))])
    net = Network()
    net.num_nodes, net.max_parents = n, maxp
    net.num_cats, net.num_parents, net.parents = keep(cats), keep(npar), keep(pars)
    net.probs, net.prob_offsets = keep(flat), keep(off)
    return net, off


def loglik(samples, cats, parents, probs, perturbations=None, mode="nodes", device=0):
    """one-shot cnb_loglik (ccnLoglik in the R shim)"""
    s = i32(samples)
    m, n = s.shape
    keep = Keep()
    net, _ = make_network(keep, cats, parents, probs)
    out = np.zeros(m if mode == "bysample" else n, dtype=np.float64)
    check(lib().cnb_loglik(C.byref(net), m, ptr(s), ptr(i32(perturbations)), LOGLIK_MODES[mode], int(device), ptr(out)))
    return out


def node_loglik(samples, cats, parents, probs, nodes, perturbations=None, device=0):
    s = i32(samples)
    m, n = s.shape
    keep = Keep()
    net, _ = make_network(keep, cats, parents, probs)
    sel = i32(nodes)
    out = np.zeros(len(sel), dtype=np.float64)
    check(lib().cnb_node_loglik(C.byref(net), len(sel), ptr(sel), m, ptr(s), ptr(i32(perturbations)), int(device), ptr(out)))
    return out


def set_prob(samples, cats, parents, perturbations=None, device=0):
    s = i32(samples)
    m, n = s.shape
    keep = Keep()
    net, off = make_network(keep, cats, parents, None)
    probs = np.zeros(off[-1], dtype=np.float64)
    ll = np.zeros(n, dtype=np.float64)
    sizes = np.zeros(n, dtype=np.int32)
    check(lib().cnb_set_prob(C.byref(net), m, ptr(s), ptr(i32(perturbations)), int(device), ptr(probs), ptr(ll), ptr(sizes)))
    return [probs[off[k]:off[k + 1]].copy() for k in range(n)], ll, sizes


def samples(cats, parents, probs, num_samples, perturbations=None, na_rate=0.0, seed=1, device=0):
    """one-shot cnb_samples (ccnSamples in the R shim): [M][N] int32, 1-based, NA_INTEGER where blanked"""
    keep = Keep()
    net, _ = make_network(keep, cats, parents, probs)
    out = np.zeros((int(num_samples), len(cats)), dtype=np.int32)
    check(lib().cnb_samples(C.byref(net), int(num_samples), ptr(i32(perturbations)), float(na_rate), int(seed), int(device),
                            ptr(out)))
    return out


def samples_dev(cats, parents, probs, num_samples, out_dev_ptr, perturbations_dev_ptr=None, na_rate=0.0, seed=1, device=0):
    """cnb_samples_dev: the [M][N] int32 matrix is written to device memory at out_dev_ptr"""
    keep = Keep()
    net, _ = make_network(keep, cats, parents, probs)
    check(lib().cnb_samples_dev(C.byref(net), int(num_samples), perturbations_dev_ptr, float(na_rate), int(seed), int(device),
                                out_dev_ptr))


PAIRWISE_KINDS = {"entropy": 0, "kl": 1, "pearson": 2}


def pairwise(kind, samples, perturbations=None, device=0):
    """one-shot cnb_pairwise (what ccnEntropyPairwise / ccnKLPairwise / ccnPearsonPairwise call in the R shim):
    samples [M][N] int32 1-based -> [N][N] with out[n1][n2] (R's matrix(mat, N, N))"""
    s = i32(samples)
    m, n = s.shape
    out = np.zeros((n, n), dtype=np.float64)
    check(lib().cnb_pairwise(PAIRWISE_KINDS[kind], n, m, ptr(s), ptr(i32(perturbations)), int(device), ptr(out)))
    return out.T.copy()


def entropy_order(samples, perturbations=None, device=0):
    s = i32(samples)
    m, n = s.shape
    out = np.zeros(n, dtype=np.int32)
    check(lib().cnb_entropy_order(n, m, ptr(s), ptr(i32(perturbations)), int(device), ptr(out)))
    return out


def _unif_ptr(unif):
    """unif: None, an address, or a ctypes UNIF_FN object (the caller keeps it alive) -> void * for the params struct"""
    if unif is None:
        return None
    return C.cast(unif, C.c_void_p) if not isinstance(unif, int) else C.c_void_p(unif)


def make_hist_params(score=2, weight=1, max_iter=32, num_threads=2, use_cache=False, seed=1, unif=None):
    h = HistParams()
    h.score, h.weight, h.max_iter, h.num_threads = int(score), int(weight), int(max_iter), int(num_threads)
    h.use_cache, h.seed = int(bool(use_cache)), int(seed)
    h.unif = _unif_ptr(unif)
    return h


def make_sa_params(keep, select_mode=0, start_order=None, temp_start=1.0, temp_cool_fact=0.9, temp_check_orders=10,
                   max_iter=100, order_shuffles=1.0, stop_diff=1e-10, num_threads=1, use_cache=False, dir_prob=None,
                   seed=1, unif=None):
    s = SaParams()
    s.select_mode = int(select_mode)
    s.start_order = keep(i32(start_order))
    s.temp_start, s.temp_cool_fact = float(temp_start), float(temp_cool_fact)
    s.temp_check_orders, s.max_iter = int(temp_check_orders), int(max_iter)
    s.order_shuffles, s.stop_diff = float(order_shuffles), float(stop_diff)
    s.num_threads, s.use_cache = int(num_threads), int(bool(use_cache))
    s.dir_prob = keep(None if dir_prob is None else f64(np.asarray(dir_prob, dtype=np.float64).T))
    s.seed = int(seed)
    s.unif = _unif_ptr(unif)
    return s


def read_sa_result(h, n, want_probs=False):
    L = lib()
    nets = read_result(h, want_probs)
    order = np.zeros(n, dtype=np.int32)
    niter, nacc = C.c_int32(), C.c_int32()
    check(L.cnb_result_sa_info(h, ptr(order), C.byref(niter), C.byref(nacc)))
    nt = L.cnb_result_sa_trace(h, None, None, 0)
    trace = np.zeros(max(nt, 1), dtype=np.float64)
    acc = np.zeros(max(nt, 1), dtype=np.int32)
    L.cnb_result_sa_trace(h, ptr(trace), ptr(acc), nt)
    return {"nets": nets, "order": order, "niter": niter.value, "naccept": nacc.value, "trace": trace[:nt],
            "trace_accept": acc[:nt]}


def last_call_timing():
    t = Timing()
    check(lib().cnb_last_call_timing(C.byref(t)))
    return t.as_dict()


def search_order(samples, want_probs=True, **kw):
    """one-shot cnb_search_order (host buffers in, networks out): the call the R shim makes"""
    s = i32(samples)
    m, n = s.shape
    keep = Keep()
    p = make_params(keep, n, m, samples=s, **kw)
    h = C.c_void_p()
    check(lib().cnb_search_order(C.byref(p), C.byref(h)))
    try:
        return read_result(h, want_probs)
    finally:
        lib().cnb_result_free(h)


def search_sa(samples, sa, want_probs=False, **kw):
    s = i32(samples)
    m, n = s.shape
    keep = Keep()
    p = make_params(keep, n, m, samples=s, **kw)
    sp = make_sa_params(keep, **sa)
    h = C.c_void_p()
    check(lib().cnb_search_sa(C.byref(p), C.byref(sp), C.byref(h)))
    try:
        return read_sa_result(h, n, want_probs)
    finally:
        lib().cnb_result_free(h)


def search_hist(samples, hist, **kw):
    """one-shot cnb_search_hist (host buffers in, histogram out): the call the R shim makes for cnSearchHist.
    hist: dict of cnb_hist_params fields.  Returns (matrix [N][N] with [parent][child], iterations)."""
    s = i32(samples)
    m, n = s.shape
    keep = Keep()
    p = make_params(keep, n, m, samples=s, **kw)
    hp = make_hist_params(**hist)
    out = np.zeros((n, n), dtype=np.float64)
    it = C.c_int32()
    check(lib().cnb_search_hist(C.byref(p), C.byref(hp), ptr(out), C.byref(it)))
    return out, it.value


def device_log(x, device=0):
    x = f64(x)
    out = np.zeros_like(x)
    check(lib().cnb_device_log(device, ptr(x), x.size, ptr(out)))
    return out


def host_log(x):
    x = f64(x)
    out = np.zeros_like(x)
    lib().cnb_host_log(ptr(x), x.size, ptr(out))
    return out


def unrank_combination(n, k, rank):
    out = np.zeros(max(k, 1), dtype=np.int32)
    check(lib().cnb_unrank_combination(n, k, int(rank), ptr(out)))
    return [int(v) for v in out[:k]]
