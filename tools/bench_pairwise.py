#!/usr/bin/env python
"""tools/bench_pairwise.py -- wall time of the pairwise metrics (cnEntropy, cnEdgeDistanceKL,
cnEdgeDistancePearson, cnEntropyOrder) on synthetic data of the C5 shape (N nodes x 4 categories x M samples), next
to the reference's own catnet_entropy.cpp (oracle/_ref/libcatnet_ref_pairwise.so, test infrastructure) on a reduced M
(its cost is linear in M).  One JSON line per case on stdout.  Measurement tool, not a bench line.
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
    ap.add_argument("--ref-samples", type=int, default=4000)
    ap.add_argument("--repeat", type=int, default=3)
    args = ap.parse_args()
    from catnet_b200 import _abi
    from oracle import ref
    n, m = args.nodes, args.samples
    rng = np.random.default_rng(5005)
    s = rng.integers(1, 5, size=(m, n), dtype=np.int32)
    s[:, 1::2] = np.where(rng.random((m, n // 2)) < 0.5, s[:, 0:n - 1:2], s[:, 1::2])
    p = (rng.random((m, n)) < 0.3).astype(np.int32)
    pairs = n * n
    for label, pert in (("no perturbations", None), ("30% perturbed entries", p)):
        out = {"workload": "pairwise metrics, %d nodes x 4 cats x %d samples, %s" % (n, m, label), "ordered_pairs": pairs}
        with _abi.Context(0) as ctx:
            out["set_samples_s"] = timed(lambda: ctx.set_samples(s, perturbations=pert), 1)
            for kind in ("entropy", "kl", "pearson"):
                out[kind + "_s"] = timed(lambda: ctx.pairwise(kind), args.repeat)
            out["entropy_order_s"] = timed(lambda: ctx.entropy_order(), args.repeat)
            out["pair_tables_per_s_entropy"] = pairs / out["entropy_s"]
        out["one_shot_entropy_host_buffers_s"] = timed(lambda: _abi.pairwise("entropy", s, pert), 1)
        if ref.pairwise_available():
            mr = min(args.ref_samples, m)
            sr, pr = s[:mr], None if pert is None else pert[:mr]
            for kind in ("entropy", "kl"):
                t0 = time.perf_counter()
                ref.pairwise(kind, sr, pr)
                t = time.perf_counter() - t0
                out["ref_%s_s_at_M%d" % (kind, mr)] = t
                out["ref_%s_s_extrapolated" % kind] = t * m / mr
            out["ref_threads"] = 1
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
