#!/usr/bin/env python
"""Fuzz the plain-C restatement (oracle/catnet_oracle.c) against the reference's own C++ compiled where it lies
(oracle/_ref, needs /root/reference): every problem runs in a forked child, because the reference corrupts its heap
on some inputs (e.g. fixed parents whose base network exceeds maxComplexity).  CPU only; test infrastructure.

    python tools/fuzz_oracle.py search 1000 6000      # tests/helpers.py::constrained_problem
    python tools/fuzz_oracle.py wide 0 3000           # inferred categories, NA-heavy, 1-4 parents, priors with 0 and 1
    python tools/fuzz_oracle.py unseen 0 2000         # declared but unseen categories, limits from the base network up
    python tools/fuzz_oracle.py sa 0 1500             # cnSearchSA trajectories (cache off) + cnSearchHist
    python tools/fuzz_oracle.py rest 0 3000           # pairwise metrics, cnSetProb / cnLoglik / cnNodeLoglik, cnSamples
    python tools/fuzz_oracle.py sacache 0 1500        # the reference with its score cache ON against itself with it OFF
    python tools/fuzz_oracle.py sacacheport 0 2000    # the restated cache against the reference's (cache ON, one thread)
    python tools/fuzz_oracle.py sacachethreads 0 600  # the reference's cache-ON run with 2-4 threads against the sequential interleaving

Round 1: 10,911 + 5,955 + 2,000 + 1,500 + 3,000 problems; one mismatch (search 3042: a node perturbed in every sample, fixed
since) and five reference crashes (search 3886, 9556, 10501, 11146, 11995: fixed parents put the base
network above maxComplexity, the documented out-of-bounds write)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from helpers import compare_search, constrained_problem, random_problem  # noqa: E402
from oracle import port, ref  # noqa: E402


def check_search(seed):
    s, P, K, kw = constrained_problem(seed)
    a = ref.search_order(s, P, K, use_cache=bool(seed % 2), **kw)
    b = port.search_order(s, P, K, **kw)
    assert len(a) == len(b), (len(a), len(b))
    assert compare_search(b, a) == len(a)


def wide_problem(seed):
    """inferred categories, heavy NA, 1-4 parents, priors with exact 0 and 1, parent-set sizes"""
    rng = np.random.default_rng(seed)
    n, m = int(rng.integers(1, 11)), int(rng.integers(4, 60))
    cats = [None, 2, 3, 4, 5, [2, 3, 4, 5, 6, 7, 2, 3, 4, 5][:n]][seed % 6]
    s, c = random_problem(seed, n, m, cats, na_frac=[0, 0.3, 0.05][seed % 3])
    kw = {}
    if seed % 2 == 0:
        kw["num_cats"] = c
    elif seed % 4 == 1:                              # inferred categories with a missing bottom category
        col = int(rng.integers(0, n))
        s[s[:, col] == 1, col] = 2
    P = int(rng.integers(1, 5))
    kw["order"] = (rng.permutation(n) + 1).astype(np.int32)
    if seed % 5 == 0:
        kw["perturbations"] = (rng.random(s.shape) < [0.1, 0.6, 0.95][seed % 3]).astype(np.int32)
    if seed % 7 == 0:
        kw["edge_prob"] = np.round(rng.random((n, n)), 1)          # includes exact 0 and 1
    if seed % 11 == 0:
        kw["parent_sizes"] = rng.integers(0, P + 1, size=n).astype(np.int32)
    lo = int(np.sum(c - 1))
    K = int(rng.integers(lo, lo + n * int(c.max()) ** min(P, 3) + 2))
    return s, P, K, kw


def check_wide(seed):
    s, P, K, kw = wide_problem(seed)
    a = ref.search_order(s, P, K, use_cache=bool(seed % 2), **kw)
    b = port.search_order(s, P, K, **kw)
    assert len(a) == len(b), (len(a), len(b))
    assert compare_search(b, a) == len(a)


def check_unseen(seed):
    """declared categories that never occur; complexity limits from exactly the base network to unbounded"""
    rng = np.random.default_rng(seed)
    n, m = int(rng.integers(2, 9)), int(rng.integers(5, 80))
    s, c = random_problem(seed, n, m, [None, 2, 3][seed % 3], na_frac=[0, 0.1][seed % 2])
    c2 = (c + rng.integers(0, 3, size=n)).astype(np.int32)
    P = int(rng.integers(1, 4))
    kw = dict(num_cats=c2, order=(rng.permutation(n) + 1).astype(np.int32))
    if seed % 4 == 0:
        kw["perturbations"] = (rng.random(s.shape) < 0.3).astype(np.int32)
    lo = int(np.sum(c2 - 1))
    K = [lo, lo + 1, lo + int(c2.max()), 10 ** 6][seed % 4]
    a = ref.search_order(s, P, K, use_cache=bool(seed % 2), **kw)
    b = port.search_order(s, P, K, **kw)
    assert len(a) == len(b), (len(a), len(b))
    assert compare_search(b, a) == len(a)


def check_sa(seed):
    rng = np.random.default_rng(seed)
    n, m = int(rng.integers(3, 9)), int(rng.integers(8, 120))
    s, c = random_problem(seed, n, m, [None, 2, 3, 4][seed % 4], na_frac=[0, 0.1][seed % 2])
    kw = dict(num_cats=c)
    if seed % 3 == 0:
        kw["perturbations"] = (rng.random(s.shape) < 0.2).astype(np.int32)
    kh = dict(kw)
    if seed % 5 == 0:
        kw["edge_prob"] = 0.1 + 0.8 * rng.random((n, n))
    sa = dict(select_mode=seed % 3, temp_start=[1.0, 0.05, 10.0][seed % 3], temp_cool_fact=[0.9, 0.5][seed % 2],
              temp_check_orders=int(rng.integers(1, 5)), max_iter=int(rng.integers(5, 40)),
              order_shuffles=[1.0, 1.5, -1.0, -2.5, 0.0, 3.0][seed % 6], stop_diff=[1e-10, 0.01][seed % 2],
              num_threads=int(rng.integers(1, 5)), seed=int(rng.integers(1, 1 << 40)))
    start = (rng.permutation(n) + 1).astype(np.int32)
    K = int(np.sum(c - 1)) + n * int(c.max()) ** 2
    a = ref.search_sa(s, 2, K, start, use_cache=False, **sa, **kw)
    b = port.search_sa(s, 2, K, start, **sa, **kw)
    assert np.array_equal(a["order"], b["order"]) and a["niter"] == b["niter"] and a["naccept"] == b["naccept"]
    assert np.array_equal(a["trace"], b["trace"]) and np.array_equal(a["trace_accept"], b["trace_accept"])
    compare_search(b["nets"], a["nets"], check_probs=False)
    hk = dict(score=seed % 3, weight=seed % 3, max_iter=6, num_threads=1 + seed % 3, seed=sa["seed"])
    h1, i1 = ref.search_hist(s, 2, K, use_cache=False, **hk, **kh)
    h2, i2 = port.search_hist(s, 2, K, **hk, **kh)
    assert i1 == i2 and np.array_equal(h1, h2)


def check_sacache(seed):
    """R always runs cnSearchSA with the score cache ON (R/catnet.search.R:293-294), the library never caches: the
    reference with its cache against the reference without it, same uniform stream -- a `mismatch` here is a DIVERGENCE of
    cache-on from cache-off (the cached table was computed under an earlier order, with its parents in another order, so
    its fp64 sum can differ in the last bit and flip a comparison)"""
    rng = np.random.default_rng(seed)
    n, m = int(rng.integers(3, 11)), int(rng.integers(8, 400))
    s, c = random_problem(seed, n, m, [None, 2, 3, 4][seed % 4], na_frac=[0, 0.1][seed % 2])
    kw = dict(num_cats=c)
    if seed % 3 == 0:
        kw["perturbations"] = (rng.random(s.shape) < 0.2).astype(np.int32)
    kh = dict(kw)
    if seed % 5 == 0:
        kw["edge_prob"] = 0.1 + 0.8 * rng.random((n, n))
    sa = dict(select_mode=seed % 3, temp_start=[1.0, 0.05, 10.0, 0.0][seed % 4], temp_cool_fact=[0.9, 0.5][seed % 2],
              temp_check_orders=int(rng.integers(1, 8)), max_iter=int(rng.integers(20, 120)),
              order_shuffles=[1.0, 1.5, -1.0, -2.5, 0.0, 3.0][seed % 6], stop_diff=[1e-10, 0.01][seed % 2],
              num_threads=int(rng.integers(1, 5)), seed=int(rng.integers(1, 1 << 40)))
    start = (rng.permutation(n) + 1).astype(np.int32)
    P = 2 + (seed % 7 == 0)
    K = int(np.sum(c - 1)) + n * int(c.max()) ** P
    a = ref.search_sa(s, P, K, start, use_cache=False, **sa, **kw)
    b = ref.search_sa(s, P, K, start, use_cache=True, **sa, **kw)
    assert np.array_equal(a["order"], b["order"]) and a["niter"] == b["niter"] and a["naccept"] == b["naccept"], "trajectory"
    assert np.array_equal(a["trace_accept"], b["trace_accept"]), "acceptances"
    assert np.array_equal(a["trace"], b["trace"]), "trace values (last bit)"
    assert compare_search(b["nets"], a["nets"], check_probs=False) == len(a["nets"]), "network log-likelihoods (last bit)"
    hk = dict(score=seed % 3, weight=seed % 3, max_iter=12, num_threads=1 + seed % 3, seed=sa["seed"])
    h1, i1 = ref.search_hist(s, 2, K, use_cache=False, **hk, **kh)
    h2, i2 = ref.search_hist(s, 2, K, use_cache=True, **hk, **kh)
    assert i1 == i2 and np.array_equal(h1, h2), "histogram"


def strict_nets(a, b):
    """same complexity slots, same LISTED parent sequences, every fp64 bit-identical (network and node log-likelihoods,
    priors, probabilities)"""
    x, y = {n["complexity"]: n for n in a}, {n["complexity"]: n for n in b}
    assert sorted(x) == sorted(y), "slots"
    for cx in x:
        u, v = x[cx], y[cx]
        assert u["parents"] == v["parents"], "listed parents"
        assert u["loglik"] == v["loglik"], "network log-likelihood"
        assert u["node_loglik"] == v["node_loglik"] and u["node_prior"] == v["node_prior"], "node terms"
        assert u["sample_sizes"] == v["sample_sizes"], "sample sizes"
        assert all(np.array_equal(p, q) for p, q in zip(u["probs"], v["probs"])), "probability tables"


def check_sacacheport(seed):
    """the restatement WITH its restated score cache (oracle/catnet_oracle.c: cache_get / cache_set) against the compiled
    reference with its cache ON, one thread (with more the reference's stores interleave by timing): trajectory, listed
    parent sequences, tables and every fp64 must be identical.  Also counts how often cache-ON differs from cache-OFF in
    the restatement (exit code 5 = identical to the reference, and cache ON != cache OFF)."""
    rng = np.random.default_rng(seed)
    n, m = int(rng.integers(3, 11)), int(rng.integers(8, 400))
    s, c = random_problem(seed, n, m, [None, 2, 3, 4][seed % 4], na_frac=[0, 0.1][seed % 2])
    kw = dict(num_cats=c)
    if seed % 3 == 0:
        kw["perturbations"] = (rng.random(s.shape) < 0.2).astype(np.int32)
        if seed % 9 == 0:
            kw["perturbations"][:, int(rng.integers(0, n))] = 1      # a node perturbed in every sample
    if seed % 11 == 0:
        kw["parents_pool"] = [sorted(rng.choice(n, size=int(rng.integers(0, n)), replace=False) + 1) for _ in range(n)]
    if seed % 13 == 0:
        kw["fixed_parents"] = [sorted(rng.choice(n, size=int(rng.integers(0, 2)), replace=False) + 1) for _ in range(n)]
    kh = dict(kw)
    if seed % 5 == 0:
        kw["edge_prob"] = 0.1 + 0.8 * rng.random((n, n))
    sa = dict(select_mode=seed % 3, temp_start=[1.0, 0.05, 10.0, 0.0][seed % 4], temp_cool_fact=[0.9, 0.5][seed % 2],
              temp_check_orders=int(rng.integers(1, 8)), max_iter=int(rng.integers(20, 120)),
              order_shuffles=[1.0, 1.5, -1.0, -2.5, 0.0, 3.0][seed % 6], stop_diff=[1e-10, 0.01][seed % 2],
              num_threads=1, seed=int(rng.integers(1, 1 << 40)))
    start = (rng.permutation(n) + 1).astype(np.int32)
    P = 2 + (seed % 7 == 0)
    K = int(np.sum(c - 1)) + n * int(c.max()) ** P
    a = ref.search_sa(s, P, K, start, use_cache=True, want_probs=True, **sa, **kw)
    b = port.search_sa(s, P, K, start, use_cache=True, want_probs=True, **sa, **kw)
    assert np.array_equal(a["order"], b["order"]) and a["niter"] == b["niter"] and a["naccept"] == b["naccept"], "trajectory"
    assert np.array_equal(a["trace_accept"], b["trace_accept"]), "acceptances"
    assert np.array_equal(a["trace"], b["trace"]), "trace values"
    strict_nets(b["nets"], a["nets"])
    hk = dict(score=seed % 3, weight=seed % 3, max_iter=12, num_threads=1, seed=sa["seed"])
    h1, i1 = ref.search_hist(s, 2, K, use_cache=True, **hk, **kh)
    h2, i2 = port.search_hist(s, 2, K, use_cache=True, **hk, **kh)
    assert i1 == i2 and np.array_equal(h1, h2), "histogram"
    off = port.search_sa(s, P, K, start, use_cache=False, want_probs=True, **sa, **kw)
    try:
        assert np.array_equal(off["trace"], b["trace"]) and np.array_equal(off["order"], b["order"])
        strict_nets(off["nets"], b["nets"])
    except AssertionError:
        os._exit(5)


def check_sacachethreads(seed):
    """How the reference's OWN cache-ON result with several threads relates to the proposal-after-proposal interleaving the
    restatement (and the library) take: its worker threads fill the one global table under a mutex in whatever order they
    get there, so the outcome depends on timing.  Exit code 5 = the reference's run differs from the sequential
    interleaving (not a mismatch: counted)."""
    rng = np.random.default_rng(seed)
    n, m = int(rng.integers(4, 11)), int(rng.integers(20, 400))
    s, c = random_problem(seed, n, m, [None, 2, 3, 4][seed % 4], na_frac=[0, 0.1][seed % 2])
    sa = dict(select_mode=seed % 3, temp_start=[1.0, 0.05, 10.0, 0.0][seed % 4], temp_cool_fact=[0.9, 0.5][seed % 2],
              temp_check_orders=int(rng.integers(1, 8)), max_iter=int(rng.integers(20, 120)),
              order_shuffles=[1.0, 1.5, -1.0, -2.5, 0.0, 3.0][seed % 6], stop_diff=[1e-10, 0.01][seed % 2],
              num_threads=2 + seed % 3, seed=int(rng.integers(1, 1 << 40)))
    start = (rng.permutation(n) + 1).astype(np.int32)
    K = int(np.sum(c - 1)) + n * int(c.max()) ** 2
    a = ref.search_sa(s, 2, K, start, use_cache=True, want_probs=True, num_cats=c, **sa)
    b = port.search_sa(s, 2, K, start, use_cache=True, want_probs=True, num_cats=c, **sa)
    # the trajectory's decisions never depend on the cache beyond last bits; what may differ is listing and last bits
    try:
        assert np.array_equal(a["order"], b["order"]) and np.array_equal(a["trace"], b["trace"])
        strict_nets(b["nets"], a["nets"])
    except AssertionError:
        os._exit(5)


def check_rest(seed):
    rng = np.random.default_rng(seed)
    n, m = int(rng.integers(1, 10)), int(rng.integers(7, 150))
    cats = [None, 2, 3, 4, 7][seed % 5]
    s, c = random_problem(seed, n, m, cats)                        # the pairwise metrics take no missing values
    if seed % 4 == 0:
        s[:, 0] = s[:, -1]                                         # duplicated column: ties in the entropy order
    if seed % 6 == 1:
        s = s + 2                                                  # categories not starting at 1
    pert = None
    if seed % 3 != 0:
        pert = (rng.random(s.shape) < [0.2, 0.7][seed % 2]).astype(np.int32)
        if seed % 7 == 0:
            pert[:, int(rng.integers(0, n))] = 1
    for kind in ("entropy", "kl", "pearson", "order"):
        if kind == "order" and n == 1:
            continue                                               # the reference returns uninitialised memory there
        assert np.array_equal(ref.pairwise(kind, s, pert), port.pairwise(kind, s, pert), equal_nan=True), kind
    s, c = random_problem(seed + 7, n, m, cats, na_frac=[0, 0.2][seed % 2])
    s0 = np.where(s < 1, 99, s - 1).astype(np.int32)
    parents = [sorted(rng.choice(i, size=min(i, int(rng.integers(0, 4))), replace=False).tolist(), reverse=seed % 3 == 0)
               for i in range(n)]
    pa, la, sa = ref.net_set_prob(s0, c, parents, pert)
    pb, lb, sb = port.net_set_prob(s0, c, parents, pert)
    assert all(np.array_equal(x, y) for x, y in zip(pa, pb)) and np.array_equal(la, lb) and np.array_equal(sa, sb)
    valid = pa
    if seed % 4 == 0:
        pa = [np.where(t < 0.15, 0.0, t) for t in pa]              # zero probabilities met by counted samples
    for mode in ("nodes", "bysample", "node", "total"):
        pe = None if mode == "node" else pert
        assert np.array_equal(ref.net_loglik(mode, s0, c, parents, pa, pe), port.net_loglik(mode, s0, c, parents, pa, pe)), mode
    if ref.samples_available():
        fixed = None
        if seed % 3 == 1:
            fixed = (rng.integers(0, int(c.max()) + 2, size=(40, n)) * (rng.random((40, n)) < 0.3)).astype(np.int32)
        a, ca = ref.net_samples(c, parents, valid, 40, fixed, [0.0, 0.3][seed % 2], seed + 1)
        b, cb = port.net_samples(c, parents, valid, 40, fixed, [0.0, 0.3][seed % 2], seed + 1)
        assert ca == cb and np.array_equal(a, b)


def main():
    check = {"search": check_search, "wide": check_wide, "unseen": check_unseen, "sa": check_sa, "rest": check_rest, "sacache": check_sacache,
             "sacacheport": check_sacacheport, "sacachethreads": check_sacachethreads}[sys.argv[1]]
    lo, hi = int(sys.argv[2]), int(sys.argv[3])
    if not ref.available():
        sys.exit("oracle/_ref is not built (needs /root/reference)")
    port.lib()
    ref.lib()
    ran, mismatch, crash, differs = 0, [], [], 0
    for seed in range(lo, hi):
        pid = os.fork()
        if pid == 0:
            code = 0
            try:
                check(seed)
            except ValueError as e:                                # degenerate problem (fewer samples than categories)
                code = 4 if "broadcast" in str(e) else 3
                if code == 3:
                    sys.stderr.write("seed %d: %r\n" % (seed, str(e)[:300]))
            except BaseException as e:
                sys.stderr.write("seed %d: %r\n" % (seed, str(e)[:300]))
                code = 3
            os._exit(code)
        _, st = os.waitpid(pid, 0)
        if os.WIFSIGNALED(st):
            crash.append(seed)
        elif os.WEXITSTATUS(st) == 3:
            mismatch.append(seed)
        elif os.WEXITSTATUS(st) == 5:
            differs += 1
        if not os.WIFSIGNALED(st) and os.WEXITSTATUS(st) != 4:
            ran += 1
    print("problems %d  mismatches %s  reference crashes %s" % (ran, mismatch, crash))
    if sys.argv[1] == "sacacheport":
        print("of these, cache ON differs from cache OFF (listing order of parents or last bits) in %d" % differs)
    if sys.argv[1] == "sacachethreads":
        print("of these, the reference with 2-4 threads and its cache ON differs from the proposal-after-proposal interleaving in %d" % differs)
    sys.exit(1 if mismatch else 0)


if __name__ == "__main__":
    main()
