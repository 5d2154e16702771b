import os, sys, time, numpy as np
os.environ.setdefault("CNB_SA_TIMING", "1")   # keep the per-phase timing events in the SA evaluations (cnb_ctx.quiet)
sys.path.insert(0, '/root/repo')
from catnet_b200 import _abi, synth
sys.path.insert(0, '/root/repo')
import bench
s, cats, sa = bench.sa_workload()
for bits in (0, 8, 64, 32):
    with _abi.Context(0) as ctx:
        ctx.set_samples(s, num_cats=cats)
        if bits: ctx.force_generic(bits)
        kw = dict(max_parent_set=2, max_complexity=600)
        ctx.search_sa(dict(sa, max_iter=5, seed=1), **kw)
        t0 = time.perf_counter()
        r = ctx.search_sa(dict(sa, max_iter=300, seed=4004), **kw)
        w = time.perf_counter() - t0
        tm = ctx.timing()
        print("bits", bits, "wall/order ms %.3f" % (1000 * w / len(r["trace"])), {k: round(tm[k], 3) for k in ("plan_ms", "score_ms", "select_ms", "dp_ms")}, "launches", tm["kernel_launches"], "tiles", tm["tensor_tiles"], flush=True)
