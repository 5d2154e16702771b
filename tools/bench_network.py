#!/usr/bin/env python
"""tools/bench_network.py -- wall time of cnSetProb / cnLoglik for the generating network of a
synthetic data set of the C5 shape (N nodes x 4 categories, <= 2 parents, M samples) on the resident data, next to the
reference's own CATNET methods (oracle/_ref, test infrastructure) on a reduced M (their cost is linear in M).
One JSON line on stdout.  Measurement tool, not a bench line.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, repeat):
    fn()
    ts = []
    for _ in range(repeat):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=500)
    ap.add_argument("--samples", type=int, default=1000000)
    ap.add_argument("--ref-samples", type=int, default=20000)
    ap.add_argument("--repeat", type=int, default=3)
    args = ap.parse_args()
    from catnet_b200 import _abi, synth
    from oracle import ref
    n, m = args.nodes, args.samples
    rng = np.random.Generator(np.random.PCG64(5005))
    net = synth.random_network(n, 4, 2, rng)
    s = np.ascontiguousarray(synth.sample_network(net, m, rng, dtype=np.int8), dtype=np.int32)
    cats, parents = net["cats"], [list(p) for p in net["parents"]]
    out = {"workload": "generating network of %d nodes x 4 cats (<= 2 parents) against its %d samples" % (n, m)}
    with _abi.Context(0) as ctx:
        out["set_samples_s"] = timed(lambda: ctx.set_samples(s, num_cats=cats), 1)
        res = {}
        out["set_prob_s"] = timed(lambda: res.__setitem__("p", ctx.set_prob(cats, parents)), args.repeat)
        probs = res["p"][0]
        out["loglik_nodes_s"] = timed(lambda: res.__setitem__("n", ctx.loglik(cats, parents, probs, "nodes")), args.repeat)
        out["loglik_bysample_s"] = timed(lambda: res.__setitem__("b", ctx.loglik(cats, parents, probs, "bysample")), args.repeat)
        out["loglik"] = float(res["n"].sum())
        out["bysample_mean"] = float(res["b"].mean())
    import torch
    probs_true = [np.asarray(t).ravel() for t in net["cpts"]]
    dev_out = torch.empty((m, n), dtype=torch.int32, device="cuda")

    def draw():
        _abi.samples_dev(cats, parents, probs_true, m, dev_out.data_ptr(), seed=5005)
        torch.cuda.synchronize()
    out["samples_to_device_s"] = timed(draw, args.repeat)
    out["samples_na5pct_to_device_s"] = timed(lambda: (_abi.samples_dev(cats, parents, probs_true, m, dev_out.data_ptr(),
                                                                        na_rate=0.05, seed=5005), torch.cuda.synchronize()), 1)
    if ref.samples_available():
        mr = min(args.ref_samples, m)
        t0 = time.perf_counter()
        ref.net_samples(cats, parents, probs_true, mr, None, 0.0, 5005)
        out["ref_samples_s_at_M%d" % mr] = time.perf_counter() - t0
    if ref.available():
        mr = min(args.ref_samples, m)
        s0 = (s[:mr] - 1).astype(np.int32)
        t0 = time.perf_counter()
        rp, rl, rs = ref.net_set_prob(s0, cats, parents)
        out["ref_set_prob_s_at_M%d" % mr] = time.perf_counter() - t0
        t0 = time.perf_counter()
        ref.net_loglik("nodes", s0, cats, parents, rp)
        out["ref_loglik_nodes_s_at_M%d" % mr] = time.perf_counter() - t0
        t0 = time.perf_counter()
        ref.net_loglik("bysample", s0, cats, parents, rp)
        out["ref_loglik_bysample_s_at_M%d" % mr] = time.perf_counter() - t0
        out["ref_scale_to_full_M"] = m / mr
        out["ref_threads"] = 1
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
