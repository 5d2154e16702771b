"""Generate tests/golden/ref_*.json from the reference itself (oracle/_ref/libcatnet_ref.so, i.e. the unmodified C++
under /root/reference/src).  Run in the build container:  python tools/make_golden.py

The reference ships no golden vectors for this path; these fixtures are its outputs on seeded
inputs, so that the CPU suite (and the GPU box, where /root/reference does not exist) can pin the oracle restatement
and the CUDA path without re-running it.  fp64 values are stored as C99 hex floats: comparisons are bit-exact.
"""
import gzip
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import random_problem  # noqa: E402
from cases import (CASES, HIST_CASES, NETWORK_CASES, PAIRWISE_CASES, SA_CASES, SAMPLES_CASES, build_case,  # noqa: E402
                   build_hist_case, build_network_case, build_pairwise_case, build_samples_case)

from catnet_b200 import synth  # noqa: E402
from oracle import ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def dump(obj, name):
    with gzip.open(os.path.join(OUT, name + ".gz"), "wt") as f:
        json.dump(obj, f)


def pack_nets(nets, with_nodes=True):
    out = []
    for n in nets:
        d = {"complexity": n["complexity"], "loglik": float(n["loglik"]).hex(), "parents": n["parents"]}
        if with_nodes:
            d["node_loglik"] = [float(v).hex() for v in n["node_loglik"]]
            d["sample_sizes"] = n["sample_sizes"]
        out.append(d)
    return out


def make_sa_cache():
    """cnSearchSA and cnSearchHist of the compiled reference with its score cache ON and one thread (the deterministic case):
    trajectory, networks with their parents in the LISTED sequence, every node's log-likelihood, the histogram"""
    from cases import SA_CACHE_SEEDS, sa_cache_problem
    out = {}
    for seed in SA_CACHE_SEEDS:
        s, c, P, K, start, sa, kw, _ = sa_cache_problem(seed)
        sa1 = dict(sa, num_threads=1)
        r = ref.search_sa(s, P, K, start, num_cats=c, use_cache=True, want_probs=False, **sa1, **kw)
        off = ref.search_sa(s, P, K, start, num_cats=c, use_cache=False, want_probs=False, **sa1, **kw)
        differs = (not np.array_equal(r["trace"], off["trace"])
                   or [n["parents"] for n in r["nets"]] != [n["parents"] for n in off["nets"]]
                   or [n["loglik"] for n in r["nets"]] != [n["loglik"] for n in off["nets"]])
        hk = dict(score=seed % 3, weight=1 + seed % 2, max_iter=10, num_threads=1, seed=sa["seed"])
        kh = {k: v for k, v in kw.items() if k != "edge_prob"}
        h, it = ref.search_hist(s, 2, K, use_cache=True, num_cats=c, **hk, **kh)
        out["seed%d" % seed] = {"order": [int(v) for v in r["order"]], "niter": r["niter"], "naccept": r["naccept"],
                                "trace": [float(v).hex() for v in r["trace"]],
                                "trace_accept": [int(v) for v in r["trace_accept"]],
                                "nets": pack_nets(r["nets"]), "differs_from_cache_off": bool(differs),
                                "hist_iterations": it, "hist": [[float(v).hex() for v in row] for row in h]}
    dump(out, "ref_search_sa_cache.json")
    print("ref_search_sa_cache.json", len(out), "cases,", sum(v["differs_from_cache_off"] for v in out.values()),
          "of them differ from the cache-off run")


def make_pairwise():
    """catnetEntropyPairwise / KLpairwise / PearsonPairwise / EntropyOrder of the reference's own catnet_entropy.cpp
    (oracle/_ref/libcatnet_ref_pairwise.so)"""
    pw = {}
    for case in PAIRWISE_CASES:
        s, p = build_pairwise_case(case)
        g = {}
        for kind in ("entropy", "kl", "pearson"):
            g[kind] = [[float(v).hex() for v in row] for row in ref.pairwise(kind, s, p)]
        g["order"] = [int(v) for v in ref.pairwise("order", s, p)] if case["n"] > 1 else None   # N = 1: uninitialised there
        pw["seed%d" % case["seed"]] = g
    dump(pw, "ref_pairwise.json")
    print("ref_pairwise.json", len(pw))


def hexes(v):
    return [float(x).hex() for x in np.asarray(v, dtype=np.float64).ravel()]


def make_network():
    """cnSetProb / cnLoglik / cnNodeLoglik through the reference's own CATNET methods (ref_driver.cpp: ref_net_set_prob,
    ref_net_loglik): tables estimated on the first half of the samples, likelihoods taken on all of them"""
    out = {}
    for case in NETWORK_CASES:
        s, cats, parents, pert = build_network_case(case)
        s0 = np.where(s < 1, 99, s - 1).astype(np.int32)
        h = max(1, len(s) // 2)
        probs, ll, sizes = ref.net_set_prob(s0[:h], cats, parents, None if pert is None else pert[:h])
        used = [np.where(t < 0.15, 0.0, t) for t in probs] if case.get("zero") else probs
        g = {"set_prob": [hexes(t) for t in probs], "set_loglik": hexes(ll), "set_sizes": [int(v) for v in sizes],
             "probs_used": [hexes(t) for t in used],
             "nodes": hexes(ref.net_loglik("nodes", s0, cats, parents, used, pert)),
             "node": hexes(ref.net_loglik("node", s0, cats, parents, used))}
        if max(len(p) for p in parents) > 0 or pert is not None:      # the reference dereferences NULL otherwise
            g["bysample"] = hexes(ref.net_loglik("bysample", s0, cats, parents, used, pert))
        out["seed%d" % case["seed"]] = g
    dump(out, "ref_network.json")
    print("ref_network.json", len(out))


def make_samples():
    """cnSamples through the reference's own RCatnet::genSamples (oracle/_ref/libcatnet_ref_samples.so) under the injected
    splitmix64 stream; the matrices are stored whole (small) with the number of uniforms consumed"""
    out = {}
    for case in SAMPLES_CASES:
        cats, parents, probs, pert, na = build_samples_case(case)
        s, calls = ref.net_samples(cats, parents, probs, case["m"], pert, na, seed=case["seed"])
        out["seed%d" % case["seed"]] = {"calls": int(calls), "samples": s.ravel().tolist()}
    dump(out, "ref_samples.json")
    print("ref_samples.json", {k: v["calls"] for k, v in out.items()})


def main():
    if "--only-samples" in sys.argv:
        return make_samples()
    if "--only-pairwise" in sys.argv:
        return make_pairwise()
    if "--only-network" in sys.argv:
        return make_network()
    if "--only-cache" in sys.argv:
        return make_sa_cache()
    fam = []
    for seed, cats, na in [(1, None, 0.0), (2, 3, 0.1), (3, 2, 0.0), (4, 4, 0.05)]:
        s, c = random_problem(seed, 8, 150, cats, na)
        s0 = np.where(s < 1, 99, s - 1).astype(np.int32)
        rng = np.random.default_rng(seed)
        for d in (0, 1, 2, 3):
            for _ in range(6):
                nodes = rng.choice(8, size=d + 1, replace=False)
                counts, ll, nc = ref.family_score(s0, c, int(nodes[0]), [int(v) for v in nodes[1:]])
                fam.append({"seed": seed, "cats": cats, "na": na, "child": int(nodes[0]),
                            "parents": [int(v) for v in nodes[1:]], "counts": [int(v) for v in counts],
                            "loglik": float(ll).hex(), "ncount": nc})
    dump(fam, "ref_family_scores.json")
    print("ref_family_scores.json", len(fam))

    searches = {}
    for case in CASES:
        s, kw = build_case(case)
        nets = ref.search_order(s, kw["max_parent_set"], kw["max_complexity"], order=kw["order"],
                                perturbations=kw.get("perturbations"), parent_sizes=kw.get("parent_sizes"),
                                num_cats=kw.get("num_cats"), parents_pool=kw.get("parents_pool"),
                                fixed_parents=kw.get("fixed_parents"), edge_prob=kw.get("edge_prob"), want_probs=False)
        searches["seed%d" % case["seed"]] = pack_nets(nets)
    dump(searches, "ref_search_order.json")
    print("ref_search_order.json", {k: len(v) for k, v in searches.items()})

    sa = {}
    for case in SA_CASES:
        s, cats = random_problem(case["seed"], case["n"], case["m"], case.get("cats"))
        r = ref.search_sa(s, case["P"], case["K"], np.arange(1, case["n"] + 1), num_cats=cats, use_cache=False,
                          **case["sa"])
        sa["seed%d" % case["seed"]] = {"order": [int(v) for v in r["order"]], "niter": r["niter"],
                                       "naccept": r["naccept"], "trace": [float(v).hex() for v in r["trace"]],
                                       "trace_accept": [int(v) for v in r["trace_accept"]],
                                       "nets": pack_nets(r["nets"], with_nodes=False)}
    dump(sa, "ref_search_sa.json")
    print("ref_search_sa.json", {k: (v["niter"], v["naccept"]) for k, v in sa.items()})
    hist = {}
    for case in HIST_CASES:
        s, kw = build_hist_case(case)
        h, it = ref.search_hist(s, case["P"], case["K"], use_cache=False, **kw, **case["hist"])
        hc, itc = ref.search_hist(s, case["P"], case["K"], use_cache=True, **kw, **case["hist"])
        assert it == itc and (np.abs(h - hc) <= 1e-12 * np.abs(h).max()).all()     # the cache only memoises scores
        hist["seed%d" % case["seed"]] = {"iterations": it, "hist": [[float(v).hex() for v in row] for row in h]}
    dump(hist, "ref_search_hist.json")
    print("ref_search_hist.json", {k: v["iterations"] for k, v in hist.items()})
    make_sa_cache()
    make_pairwise()
    make_network()
    make_samples()
    if "--skip-configs" in sys.argv:
        return
    big = {}
    c = synth.config("C1")
    big["C1"] = pack_nets(ref.search_order(c["samples"], 2, c["max_complexity"], order=c["order"] + 1,
                                           num_cats=c["cats"], want_probs=False), with_nodes=False)
    c = synth.config("C2")
    big["C2"] = pack_nets(ref.search_order(c["samples"], 3, c["max_complexity"], num_cats=c["cats"],
                                           want_probs=False), with_nodes=False)
    c = synth.config("C3", m=300)
    big["C3_m300"] = pack_nets(ref.search_order(c["samples"], 3, c["max_complexity"], order=c["order"] + 1,
                                                num_cats=c["cats"], want_probs=False), with_nodes=False)
    dump(big, "ref_configs.json")
    print("ref_configs.json", {k: len(v) for k, v in big.items()})



if __name__ == "__main__":
    main()
