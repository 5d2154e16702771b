#!/usr/bin/env python
"""tools/bench_hist.py -- searches over nodes with MORE than four categories (no bit planes, no tensor tiles): the
histogram scorers side by side -- a warp per family with its own shared-memory table (atomics / __match_any_sync) and the
one-CTA-per-family kernel.  Default shape: 100 nodes x 8 categories x 100,000 samples, maxParentSet = 2 (161,700 + 4,950 +
100 families), the matrix drawn from a random network by the library's sampler.  One JSON line per case; the sample
matrix (10 MB of uint8 codes) stays in L2, so the bound is the shared-memory read-modify-write rate, reported as
family-samples/s.  Measurement tool, not a bench line."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=100)
    ap.add_argument("--cats", type=int, default=8)
    ap.add_argument("--samples", type=int, default=100000)
    ap.add_argument("--parents", type=int, default=2)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--modes", default="1,2,0")
    ap.add_argument("--skew", type=float, default=0.0, help="probability mass of category 1 in every column (0 = uniform-ish CPTs)")
    ap.add_argument("--na", type=float, default=0.0, help="fraction of missing values (the warp kernel then checks every sample)")
    args = ap.parse_args()
    import torch
    from catnet_b200 import _abi
    n, m, C = args.nodes, args.samples, args.cats
    rng = np.random.default_rng(5)
    s = np.zeros((m, n), dtype=np.int32)
    for i in range(n):                                       # every node copies an earlier one half of the time
        if args.skew > 0:
            p = np.full(C, (1 - args.skew) / (C - 1))
            p[0] = args.skew
            base = rng.choice(C, size=m, p=p)
        else:
            base = rng.integers(0, C, size=m)
        if i:
            par = rng.integers(0, i)
            base = np.where(rng.random(m) < 0.5, s[:, par] - 1, base)
        s[:, i] = base + 1
    if args.na > 0:
        mask = rng.random(s.shape) < args.na
        mask[:C, :] = False
        s[mask] = -2147483648
    cats = np.full(n, C, dtype=np.int32)
    K = n * (C - 1) * C ** args.parents
    kw = dict(max_parent_set=args.parents, max_complexity=K)
    names = {0: "one CTA per family (k_score_hist)", 1: "warp per family, shared atomics (k_score_hist_warp<false>)",
             2: "warp per family, __match_any_sync (k_score_hist_warp<true>)"}
    bits = {0: 32768, 1: 0, 2: 65536}
    ref = None
    with _abi.Context(0) as ctx:
        ctx.set_samples(s, num_cats=cats)
        for mode in [int(x) for x in args.modes.split(",")]:
            ctx.force_generic(bits[mode])
            seg, nf = ctx.plan(0, 1, **kw)
            scores = torch.empty(seg, dtype=torch.float64, device="cuda")

            def step():
                ctx.plan(0, 1, **kw)
                ctx.score_shard(scores.data_ptr())
                ctx.select_dp(scores.data_ptr())
            step()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                step()
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) / args.steps
            tm = ctx.timing()
            sc = scores.cpu().numpy()
            if ref is None:
                ref = sc
            print(json.dumps({"mode": mode, "kernel": names[mode], "nodes": n, "cats": C, "samples": m, "parents": args.parents, "skew": args.skew, "na": args.na,
                              "families": nf, "step_ms": 1000 * wall, "score_ms": tm["score_ms"], "score_ms_by_parents": tm["score_ms_by_parents"][:args.parents + 1],
                              "families_per_s": nf / wall, "family_samples_per_s": nf * m / (tm["score_ms"] / 1e3),
                              "fast_path": tm["fast_path"], "scores_equal_first_mode": bool(np.array_equal(sc, ref, equal_nan=True))}), flush=True)


if __name__ == "__main__":
    main()
