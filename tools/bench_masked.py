#!/usr/bin/env python
"""tools/bench_masked.py -- the C5 shape (500 nodes x 4 categories x 1M samples, maxParentSet=2) with missing values and
with perturbation masks: one order on a resident context, tensor cores (full-table tiles) against the bit-plane kernels.

The matrix is drawn on the device by the library's own sampler (cnb_samples_dev, naRate for the missing values); the
perturbation mask with torch.  One JSON line per case.  Measurement tool, not a bench line.
"""
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
    ap.add_argument("--samples", type=int, default=1000000)
    ap.add_argument("--cases", default="clean,na5,pert30,na5pert30")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--slow", action="store_true", help="also time the bit-plane kernels (about 10 s per order at full size)")
    args = ap.parse_args()
    import torch
    import bench
    from catnet_b200 import _abi
    wl = bench.build_workload("C5", args.samples)
    net, n, m = wl["net"], wl["n"], wl["m"]
    parents = [list(p) for p in net["parents"]]
    probs = [np.asarray(t, dtype=np.float64).ravel() for t in net["cpts"]]
    kw = dict(max_parent_set=2, max_complexity=wl["K"], order=wl["order"].astype(np.int32))
    dev = torch.device("cuda", 0)
    for case in args.cases.split(","):
        na = 0.05 if "na5" in case else 0.0
        s = torch.empty((m, n), dtype=torch.int32, device=dev)
        _abi.samples_dev(net["cats"], parents, probs, m, s.data_ptr(), na_rate=na, seed=77)
        pert = None
        if "pert30" in case:
            g = torch.Generator(device=dev)
            g.manual_seed(5)
            pert = (torch.rand((m, n), device=dev, generator=g) < 0.3).to(torch.int32)
        out = {"case": case, "nodes": n, "samples": m, "na_rate": na, "perturbed": 0.3 if pert is not None else 0.0}
        with _abi.Context(0) as ctx:
            ctx.set_samples_dev(s.data_ptr(), n, m, perturbations_dev_ptr=pert.data_ptr() if pert is not None else None,
                                num_cats=net["cats"])
            seg, nf = ctx.plan(0, 1, **kw)
            scores = torch.empty(seg, dtype=torch.float64, device=dev)

            def step():
                ctx.plan(0, 1, **kw)
                ctx.score_shard(scores.data_ptr())
                ctx.select_dp(scores.data_ptr())
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                step()
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) / args.steps
            tm = ctx.timing()
            h = ctx.collect_raw()
            L = _abi.lib()
            nn = L.cnb_result_num_networks(h)
            L.cnb_result_free(h)
            out.update({"step_ms": 1000 * wall, "families": nf, "families_per_s": nf / wall, "tensor_tiles": int(tm["tensor_tiles"]),
                        "tensor_ms": tm["tensor_ms"], "score_ms": tm["score_ms"], "select_ms": tm["select_ms"], "dp_ms": tm["dp_ms"],
                        "tensor_tflops": 2.0 * tm["tensor_macs"] / (tm["tensor_ms"] / 1e3) / 1e12 if tm["tensor_ms"] > 0 else None,
                        "networks": nn})
            if args.slow:
                ctx.force_generic(8)
                step()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                step()
                torch.cuda.synchronize()
                out["bit_plane_step_ms"] = 1000 * (time.perf_counter() - t0)
                h = ctx.collect_raw()
                out["bit_plane_networks"] = L.cnb_result_num_networks(h)
                L.cnb_result_free(h)
        print(json.dumps(out), flush=True)
        del s, pert


if __name__ == "__main__":
    main()
