#!/usr/bin/env python
"""tools/bench_configs.py -- wall time of the BASELINE.json configs that are NOT bench.py's headline line:
C1 (ALARM, M=1,000), C2 (breast), C3 (100 x 3 cats x 100k, three parents) through cnb_search_order with HOST buffers,
and C4 (cnSearchSA on ALARM, M=20,000, demo/alarm.r schedule) through cnb_search_sa, each next to the reference's own
C++ (oracle/_ref, test infrastructure) on this host where that finishes within --ref-budget seconds.

One JSON line per config on stdout.  Measurement tool, not a bench line.
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
    ap.add_argument("--configs", default="C1,C2,C3,C4,H4")
    ap.add_argument("--ref-budget", type=float, default=120.0, help="skip the reference when its estimate exceeds this")
    ap.add_argument("--sa-iter", type=int, default=1000)
    ap.add_argument("--repeat", type=int, default=3)
    args = ap.parse_args()
    from catnet_b200 import _abi, synth
    from oracle import ref
    have_ref = ref.available()
    for name in args.configs.split(","):
        if name == "H4":
            # cnSearchHist on the C4 data: 32 random orders, 2 per batch, BIC, likelihood weights
            cfg = synth.config("C4")
            s = np.ascontiguousarray(cfg["samples"], dtype=np.int32)
            hp = dict(score=2, weight=1, max_iter=32, num_threads=2, seed=4004)
            kw = dict(max_parent_set=2, max_complexity=600, num_cats=cfg["cats"])
            _abi.search_hist(s, dict(hp, max_iter=2), **kw)
            ts = []
            for _ in range(args.repeat):
                t0 = time.perf_counter()
                h, it = _abi.search_hist(s, hp, **kw)
                ts.append(time.perf_counter() - t0)
            out = {"config": "H4: cnSearchHist on cnSamples(alarmnet, 20000), maxParentSet=2, maxComplexity=600, BIC, "
                             "likelihood weights, maxIter=32, numThreads=2", "ours_s": min(ts), "iterations": it,
                   "hist_sum": float(h.sum())}
            if have_ref:
                t0 = time.perf_counter()
                hr, itr = ref.search_hist(s, 2, 600, score=2, weight=1, max_iter=8, num_threads=2, use_cache=True,
                                          seed=4004, num_cats=cfg["cats"])
                out.update({"ref_s_first_8_iterations": time.perf_counter() - t0, "ref_iterations": itr,
                            "ref_threads": 2, "ref_cache": True})
            print(json.dumps(out), flush=True)
            continue
        cfg = synth.config(name)
        s = np.ascontiguousarray(cfg["samples"], dtype=np.int32)
        m, n = s.shape
        order1 = (np.asarray(cfg["order"]) + 1).astype(np.int32)
        out = {"config": name, "nodes": n, "samples": m, "max_parent_set": cfg["max_parent_set"],
               "max_complexity": cfg["max_complexity"], "families": int(synth.num_families(n, cfg["max_parent_set"]))}
        if name == "C4":
            sa = dict(select_mode=2, start_order=order1[np.random.default_rng(7).permutation(n)], temp_start=0.0,
                      temp_cool_fact=1.0, temp_check_orders=1000, max_iter=args.sa_iter, order_shuffles=0.0,
                      stop_diff=1e-10, num_threads=4, use_cache=True, seed=4004)
            kw = dict(max_parent_set=2, max_complexity=600, num_cats=cfg["cats"])
            _abi.search_sa(s, dict(sa, max_iter=3), **kw)                       # warm-up (context, JIT-free)
            ts = []
            for _ in range(args.repeat):
                t0 = time.perf_counter()
                r = _abi.search_sa(s, sa, **kw)
                ts.append(time.perf_counter() - t0)
            out.update({"ours_s": min(ts), "ours_iterations": int(r["niter"]), "ours_orders_evaluated": int(len(r["trace"])),
                        "ours_best_loglik": float(max(x["loglik"] for x in r["nets"])) if r["nets"] else None})
            if have_ref:
                t0 = time.perf_counter()
                rr = ref.search_sa(s, 2, 600, sa["start_order"], select_mode=2, temp_start=0.0, temp_cool_fact=1.0,
                                   temp_check_orders=1000, max_iter=min(args.sa_iter, 40), order_shuffles=0.0,
                                   stop_diff=1e-10, num_threads=4, use_cache=True, seed=4004, num_cats=cfg["cats"])
                t40 = time.perf_counter() - t0
                out.update({"ref_s_first_40_iterations": t40, "ref_iterations": int(rr["niter"]),
                            "ref_threads": 4, "ref_cache": True})
        else:
            kw = dict(max_parent_set=cfg["max_parent_set"], max_complexity=cfg["max_complexity"], order=order1,
                      num_cats=cfg["cats"])
            import ctypes
            L = _abi.lib()
            keep = _abi.Keep()
            p = _abi.make_params(keep, n, m, samples=s, **kw)

            def one_shot():
                """the .Call body: cnb_search_order with host buffers, networks materialised on the host"""
                h = ctypes.c_void_p()
                t0 = time.perf_counter()
                _abi.check(L.cnb_search_order(ctypes.byref(p), ctypes.byref(h)))
                t = time.perf_counter() - t0
                k = L.cnb_result_num_networks(h)
                L.cnb_result_free(h)
                return t, k
            one_shot()                                                            # warm-up (module load)
            runs = [one_shot() for _ in range(args.repeat)]
            t_abi = min(t for t, _ in runs)
            t0 = time.perf_counter()
            nets = _abi.search_order(s, want_probs=False, **kw)                   # + Python marshalling of every network
            t_py = time.perf_counter() - t0
            with _abi.Context(0) as ctx:                                          # resident data: one order evaluation
                ctx.set_samples(s, num_cats=cfg["cats"])
                kw2 = {k: v for k, v in kw.items() if k != "num_cats"}
                L.cnb_result_free(ctx.search_order_raw(**kw2))
                t0 = time.perf_counter()
                L.cnb_result_free(ctx.search_order_raw(**kw2))
                t_res = time.perf_counter() - t0
                tm = ctx.timing()
            out.update({"ours_s": t_abi, "networks": len(nets), "families_per_s": out["families"] / t_abi,
                        "ours_resident_s": t_res, "families_per_s_resident": out["families"] / t_res,
                        "ours_with_python_export_s": t_py,
                        "device_ms": {k: round(tm[k], 3) for k in ("plan_ms", "score_ms", "select_ms", "dp_ms", "collect_ms")},
                        "score_ms_by_parents": [round(v, 3) for v in tm["score_ms_by_parents"][:4]]})
            if have_ref:
                # estimate the reference from its scoring primitive: ~8e7 sample visits per second per thread
                est = out["families"] * m / 8e7
                if est <= args.ref_budget:
                    t0 = time.perf_counter()
                    theirs = ref.search_order(s, cfg["max_parent_set"], cfg["max_complexity"], order=order1,
                                              num_cats=cfg["cats"], want_probs=False)
                    out.update({"ref_s": time.perf_counter() - t0, "ref_networks": len(theirs), "ref_threads": 1})
                else:
                    out.update({"ref_s": None, "ref_estimate_s": est,
                                "ref_note": "not run: estimate from 8e7 sample visits/s/thread exceeds the budget"})
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
