"""The library's host-side model of the reference's score cache (catnet_b200/csrc/cnb_cache.cpp, CPU only) against the
restated cache of the oracle (oracle/catnet_oracle.c: cache_get / cache_set, itself fuzzed against the compiled reference
with its cache ON -- tools/fuzz_oracle.py sacacheport): on the sequence of orders an SA chain evaluates, the model must
report the same hits, with the parents in the same listed sequence and the same stored prior, as the oracle's cache
produced while it ran that chain."""
import ctypes as C

import numpy as np
import pytest

from catnet_b200 import _abi
from cases import sa_cache_problem
from helpers import random_problem
from oracle import port

ROW = 4 + 8          # the library's rows (CNB_MAXP parents)
OROW = 4 + 16        # the oracle's (ORC_MAXP)


def oracle_trace(s, cats, P, K, start, sa, **kw):
    L = port.lib()
    cap, ocap = 400000, 2000
    n = s.shape[1]
    rows = np.zeros((cap, OROW), dtype=np.int32)
    priors = np.zeros(cap, dtype=np.float64)
    orders = np.zeros((ocap, n), dtype=np.int32)
    L.orc_cache_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.orc_cache_trace.restype = None
    L.orc_cache_trace_counts.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.orc_cache_trace_counts.restype = None
    L.orc_cache_trace(rows.ctypes.data, priors.ctypes.data, cap, orders.ctypes.data, ocap)
    try:
        r = port.search_sa(s, P, K, start, use_cache=True, num_cats=cats, **sa, **kw)
        nr, ns = C.c_int(), C.c_int()
        L.orc_cache_trace_counts(C.byref(nr), C.byref(ns))
    finally:
        L.orc_cache_trace(None, None, 0, None, 0)
    assert nr.value <= cap and ns.value <= ocap
    assert not rows[:, ROW:].any()
    return r, rows[:nr.value, :ROW], priors[:nr.value], orders[:ns.value]


def model_trace(n, cats, P, K, orders, winners, never_kept=None, all_hits=True, **kw):
    keep = []

    def hold(a):
        if a is not None:
            keep.append(a)
        return _abi.ptr(a)
    p = _abi.make_params(hold, n, 1, max_parent_set=P, max_complexity=K, **kw)
    cap = 400000
    rows = np.zeros((cap, ROW), dtype=np.int32)
    priors = np.zeros(cap, dtype=np.float64)
    c32 = np.ascontiguousarray(cats, dtype=np.int32)
    o32 = np.ascontiguousarray(orders, dtype=np.int32)
    w32 = np.ascontiguousarray(winners, dtype=np.int32)
    nk = None if never_kept is None else np.ascontiguousarray(never_kept, dtype=np.uint8)
    got = _abi.lib().cnb_cache_trace(C.byref(p), c32.ctypes.data, None if nk is None else nk.ctypes.data, len(o32),
                                     o32.ctypes.data, w32.ctypes.data, int(all_hits), cap, rows.ctypes.data, priors.ctypes.data)
    assert 0 <= got <= cap, got
    return rows[:got], priors[:got]


def problem(seed):
    return sa_cache_problem(seed)


def acted_on(hits, orders, kw):
    """what the search acts on: the hits that change what the device computed for the order -- every hit where the pool has
    equal category counts (the winner is replaced), otherwise the hits whose cached listing is not the one this order gives
    the family (fixed parents first, then the others by position), and all hits when priors are in use"""
    fixed = kw.get("fixed_parents")
    keep = []
    for row in hits:
        if row[0] == 1 or "edge_prob" in kw:
            keep.append(True)
            continue
        pos = {int(lbl): i for i, lbl in enumerate(orders[row[1]])}
        pars = [int(v) for v in row[4:4 + row[3]]]
        fx = set(fixed[row[2] - 1]) if fixed is not None else set()
        default = sorted([q for q in pars if q in fx], key=pos.get) + sorted([q for q in pars if q not in fx], key=pos.get)
        keep.append(pars != default)
    return np.array(keep, dtype=bool) if len(hits) else np.zeros(0, dtype=bool)


@pytest.mark.parametrize("seed", range(60))
def test_cache_model_reports_the_oracles_hits(seed):
    s, c, P, K, start, sa, kw, never = problem(seed)
    n = s.shape[1]
    r, rows, priors, orders = oracle_trace(s, c, P, K, start, sa, **kw)
    if r is None:
        pytest.skip("no network")
    # equal-categories winners the oracle stored, as the device would report them
    winners = np.zeros((len(orders), n, P, P), dtype=np.int32)
    for row in rows[rows[:, 0] == 3]:
        winners[row[1], row[2] - 1, row[3] - 1, :row[3]] = row[4:4 + row[3]]
    hits, hp = rows[rows[:, 0] != 3], priors[rows[:, 0] != 3]
    mkw = {k: v for k, v in kw.items() if k != "perturbations"}
    got, gp = model_trace(n, c, P, K, orders, winners, never_kept=never, **mkw)
    assert len(got) == len(hits), (len(got), len(hits))
    assert np.array_equal(got, hits)
    if "edge_prob" in kw:
        assert np.array_equal(gp, hp)                         # fp64 bit equality: host prior == the oracle's
    assert len(orders) > 1
    keep = acted_on(hits, orders, kw)
    got2, _ = model_trace(n, c, P, K, orders, winners, never_kept=never, all_hits=False, **mkw)
    assert np.array_equal(got2, hits[keep])


def test_cache_table_size_follows_the_reference():
    """cache.cpp:225-264: N^(P+1) below 1069 collapses to 2 and is then raised to 31 slots; otherwise the first prime >= N
    to the power P + 1"""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    # indirectly: the restated cache with these sizes reproduces the compiled reference on the fuzz seeds (sacacheport);
    # here the model's hits equal the oracle's on a problem with 37 nodes (37^3 = 50,653 slots) and mixed categories
    rng = np.random.default_rng(5)
    s, c = random_problem(5, 37, 60, None)
    sa = dict(select_mode=2, temp_start=1.0, temp_cool_fact=0.9, temp_check_orders=5, max_iter=6, order_shuffles=4.0,
              stop_diff=1e-10, num_threads=1, seed=9)
    start = (rng.permutation(37) + 1).astype(np.int32)
    K = int(np.sum(c - 1)) + 37 * 16
    r, rows, priors, orders = oracle_trace(s, c, 2, K, start, sa)
    winners = np.zeros((len(orders), 37, 2, 2), dtype=np.int32)
    for row in rows[rows[:, 0] == 3]:
        winners[row[1], row[2] - 1, row[3] - 1, :row[3]] = row[4:4 + row[3]]
    hits = rows[rows[:, 0] != 3]
    got, _ = model_trace(37, c, 2, K, orders, winners)
    assert len(hits) > 1000 and np.array_equal(got, hits)
    got2, _ = model_trace(37, c, 2, K, orders, winners, all_hits=False)
    assert 0 < len(got2) < len(got) and np.array_equal(got2, hits[acted_on(hits, orders, {})])


@pytest.mark.parametrize("n", [170, 260])
def test_cache_model_beyond_the_dense_index_and_the_flat_table(n):
    """more than 160 nodes: every lookup goes through the hash slots (no dense key index); 260 nodes with two parents:
    263^3 slots, more than the model keeps as a flat array (it keeps only the slots in use; the reference allocates them
    all) -- parent pools keep the candidate lists short"""
    rng = np.random.default_rng(n)
    s, c = random_problem(n, n, 40, None)
    pool = [sorted((rng.choice(n, size=5, replace=False) + 1).tolist()) for _ in range(n)]
    sa = dict(select_mode=2, temp_start=1.0, temp_cool_fact=0.9, temp_check_orders=5, max_iter=8, order_shuffles=6.0,
              stop_diff=1e-10, num_threads=1, seed=3)
    start = (rng.permutation(n) + 1).astype(np.int32)
    K = int(np.sum(c - 1)) + n * 16
    kw = dict(parents_pool=pool)
    r, rows, priors, orders = oracle_trace(s, c, 2, K, start, sa, **kw)
    winners = np.zeros((len(orders), n, 2, 2), dtype=np.int32)
    for row in rows[rows[:, 0] == 3]:
        winners[row[1], row[2] - 1, row[3] - 1, :row[3]] = row[4:4 + row[3]]
    hits = rows[rows[:, 0] != 3]
    got, _ = model_trace(n, c, 2, K, orders, winners, **kw)
    assert len(hits) > 100 and np.array_equal(got, hits)
    got2, _ = model_trace(n, c, 2, K, orders, winners, all_hits=False, **kw)
    assert np.array_equal(got2, hits[acted_on(hits, orders, kw)])
