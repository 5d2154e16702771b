#!/usr/bin/env python
"""tools/bench_e2e.py -- the one-shot drop-in call (cnb_search_order: pageable host matrix in, networks on the host out) on
the C5 shape under different slicings of the streamed upload (CNB_STREAM_SLICES) and with streaming off.  The matrix is
drawn on the device by the library's sampler and copied to the host once.  Measurement tool, not a bench line."""
import argparse
import ctypes
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
    ap.add_argument("--slicings", default="off;1,3,4,4,4;1,5,10;1,3,12;2,6,8;1,15;1,2,5,8")
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--devices", type=int, default=0)
    args = ap.parse_args()
    import torch
    import bench
    from catnet_b200 import _abi
    L = _abi.lib()
    wl = bench.build_workload("C5", args.samples)
    net, n, m = wl["net"], wl["n"], wl["m"]
    s = torch.empty((m, n), dtype=torch.int32, device="cuda")
    _abi.samples_dev(net["cats"], [list(p) for p in net["parents"]], [np.asarray(t, dtype=np.float64).ravel() for t in net["cpts"]],
                     m, s.data_ptr(), seed=77)
    host = s.cpu().numpy()                                  # pageable
    del s
    torch.cuda.empty_cache()
    keep = _abi.Keep()
    p = _abi.make_params(keep, n, m, samples=host, num_cats=net["cats"], max_parent_set=2, max_complexity=wl["K"],
                         order=wl["order"].astype(np.int32), num_devices=args.devices)
    for sl in args.slicings.split(";"):
        if sl == "off":
            os.environ["CNB_STREAM_MIN"] = "off"
        else:
            os.environ.pop("CNB_STREAM_MIN", None)
            os.environ["CNB_STREAM_SLICES"] = sl
        ts = []
        for k in range(args.steps + 1):
            h = ctypes.c_void_p()
            t0 = time.perf_counter()
            _abi.check(L.cnb_search_order(ctypes.byref(p), ctypes.byref(h)))
            ts.append(time.perf_counter() - t0)
            nn = L.cnb_result_num_networks(h)
            L.cnb_result_free(h)
        tm = _abi.last_call_timing()
        print(json.dumps({"slices": sl, "ms": 1000 * float(np.median(ts[1:])), "min_ms": 1000 * min(ts[1:]), "networks": nn,
                          "streamed": tm["fast_path"] == 2, "stream_phase_ms": tm["pack_ms"], "score_ms": tm["score_ms"],
                          "select_ms": tm["select_ms"], "dp_ms": tm["dp_ms"], "collect_ms": tm["collect_ms"]}), flush=True)


if __name__ == "__main__":
    main()
