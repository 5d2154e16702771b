#!/usr/bin/env python
"""Fuzz the CUDA path (through the C ABI) against the plain-C restatement on the seeded generators that
tools/fuzz_oracle.py pins against the compiled reference.  Needs a B200; reads nothing under /root/reference.

    python tools/fuzz_gpu.py search 1000 3000     # tests/helpers.py::constrained_problem
    python tools/fuzz_gpu.py wide 0 3000          # tools/fuzz_oracle.py::wide_problem
    python tools/fuzz_gpu.py pert 1100 1400       # tests/helpers.py::fully_perturbed_problem
    python tools/fuzz_gpu.py many 0 1500          # 2..9 categories per node: the histogram scorers
    CNB_FORCE_BITS=4 python tools/fuzz_gpu.py search 0 1500    # the same problems through the tensor-core tiles
    python tools/fuzz_gpu.py sacache 200 1200     # cnSearchSA + cnSearchHist with use_cache (tests/cases.py::sa_cache_problem)

Totals of the runs are in DESIGN.md section 5.4."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
from helpers import compare_search, constrained_problem, fully_perturbed_problem  # noqa: E402


def many_categories_problem(seed):
    """2..9 categories per node (more than four: no bit planes, the histogram scorers), missing values, perturbations,
    pools, priors, 1..3 parents while the largest table stays below 16,384 cells"""
    import numpy as np
    from helpers import random_problem
    rng = np.random.default_rng(seed)
    n, m = int(rng.integers(3, 9)), int(rng.integers(10, 300))
    cats = rng.integers(2, 10, size=n) if seed % 3 else int(rng.integers(5, 10))
    s, c = random_problem(seed, n, m, cats, na_frac=[0.0, 0.1, 0.0, 0.03][seed % 4])
    P = 1 + seed % 3
    while int(c.max()) ** (P + 1) > 16384:
        P -= 1
    kw = dict(num_cats=c, order=(rng.permutation(n) + 1).astype(np.int32))
    if seed % 4 == 2:
        kw["perturbations"] = (rng.random(s.shape) < 0.25).astype(np.int32)
    if seed % 5 == 0:
        kw["parents_pool"] = [[int(v) for v in rng.choice(np.arange(1, n + 1), size=rng.integers(0, n), replace=False)
                               if v != i + 1] for i in range(n)]
    if seed % 7 == 3:
        kw["edge_prob"] = 0.05 + 0.9 * rng.random((n, n))
    lo = int(np.sum(c - 1))
    return s, P, int(rng.integers(lo, lo + n * int(c.max()) ** P + 1)), kw


def fuzz_sa_cache(lo, hi):
    """cnSearchSA and cnSearchHist with the reference's score cache modelled (use_cache, csrc/cnb_cache.h) against the
    restatement with its restated cache (pinned against the compiled reference by tools/fuzz_oracle.py sacacheport):
    trajectory, listed parent sequences, every fp64, tables, histogram"""
    import numpy as np
    from cases import sa_cache_problem
    from catnet_b200 import _abi
    from oracle import port
    ran, bad, differs = 0, [], 0
    for seed in range(lo, hi):
        try:
            s, c, P, K, start, sa, kw, _ = sa_cache_problem(seed)
            o = port.search_sa(s, P, K, start, num_cats=c, use_cache=True, want_probs=True, **sa, **kw)
        except ValueError:
            continue
        if o is None:
            continue
        ran += 1
        try:
            r = _abi.search_sa(s, dict(sa, start_order=start, use_cache=True), want_probs=True, max_parent_set=P,
                               max_complexity=K, num_cats=c, **kw)
            assert np.array_equal(r["order"], o["order"]) and r["niter"] == o["niter"] and r["naccept"] == o["naccept"], "trajectory"
            assert np.array_equal(r["trace"], o["trace"]) and np.array_equal(r["trace_accept"], o["trace_accept"]), "trace"
            x, y = {n["complexity"]: n for n in r["nets"]}, {n["complexity"]: n for n in o["nets"]}
            assert sorted(x) == sorted(y), "slots"
            for cx in x:
                u, v = x[cx], y[cx]
                assert u["parents"] == v["parents"], "listed parents"
                assert u["loglik"] == v["loglik"] and list(u["node_loglik"]) == list(v["node_loglik"]), "log-likelihoods"
                assert list(u["node_prior"]) == list(v["node_prior"]), "priors"
                assert all(np.array_equal(p, q) for p, q in zip(u["probs"], v["probs"])), "tables"
            hk = dict(score=seed % 3, weight=seed % 3, max_iter=9, num_threads=1 + seed % 4, seed=sa["seed"])
            kh = {k: v for k, v in kw.items() if k != "edge_prob"}
            h, it = _abi.search_hist(s, dict(hk, use_cache=True), max_parent_set=2, max_complexity=K, num_cats=c, **kh)
            h2, it2 = port.search_hist(s, 2, K, use_cache=True, num_cats=c, **hk, **kh)
            assert it == it2 and np.array_equal(h, h2), "histogram"
            off = port.search_sa(s, P, K, start, num_cats=c, use_cache=False, **sa, **kw)
            differs += (not np.array_equal(off["trace"], o["trace"]) or [n["parents"] for n in off["nets"]] != [n["parents"] for n in o["nets"]]
                        or [n["loglik"] for n in off["nets"]] != [n["loglik"] for n in o["nets"]])
        except (AssertionError, _abi.CnbError) as e:
            bad.append((seed, str(e)[:200]))
    print("sacache [%d, %d) CNB_FORCE_BITS=%s: problems %d  mismatches %d %s; cache ON differs from cache OFF in %d of them" % (
        lo, hi, os.environ.get("CNB_FORCE_BITS", "-"), ran, len(bad), bad[:20], differs))
    sys.exit(1 if bad else 0)


def main():
    if sys.argv[1] == "sacache":
        return fuzz_sa_cache(int(sys.argv[2]), int(sys.argv[3]))
    from catnet_b200 import _abi
    from fuzz_oracle import wide_problem
    from oracle import port
    gen = {"search": constrained_problem, "wide": wide_problem, "pert": fully_perturbed_problem,
           "many": many_categories_problem}[sys.argv[1]]
    ran, bad = 0, []
    for seed in range(int(sys.argv[2]), int(sys.argv[3])):
        try:
            s, P, K, kw = gen(seed)
        except ValueError:                       # fewer samples than categories
            continue
        theirs = port.search_order(s, P, K, **kw)
        try:
            ours = _abi.search_order(s, max_parent_set=P, max_complexity=K, **kw)
        except _abi.CnbError as e:               # the library refuses what the restatement answers with nothing
            ours = []
            if theirs:
                bad.append((seed, "error %s" % e))
                continue
        ran += 1
        try:
            assert len(ours) == len(theirs), (len(ours), len(theirs))
            assert compare_search(ours, theirs) == len(theirs)
        except AssertionError as e:
            bad.append((seed, str(e)[:200]))
    print("%s [%s, %s) CNB_FORCE_BITS=%s: problems %d  mismatches %d %s" % (sys.argv[1], sys.argv[2], sys.argv[3],
                                                                      os.environ.get("CNB_FORCE_BITS", "-"), ran, len(bad), bad[:20]))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
