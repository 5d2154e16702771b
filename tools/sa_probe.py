import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from catnet_b200 import _abi
import bench
s, cats, sa = bench.sa_workload()
kw = dict(max_parent_set=2, max_complexity=600)
_abi.search_sa(s, dict(sa, max_iter=3, seed=1), num_cats=cats, **kw)
for it in (100, 300, 1000):
    t0 = time.perf_counter(); r = _abi.search_sa(s, dict(sa, max_iter=it, seed=4004), num_cats=cats, **kw); w = time.perf_counter() - t0
    print("one-shot", it, "wall %.3f s" % w, "orders", len(r["trace"]), "ms/order %.3f" % (1000 * w / len(r["trace"])), flush=True)
with _abi.Context(0) as ctx:
    ctx.set_samples(s, num_cats=cats)
    ctx.search_sa(dict(sa, max_iter=5, seed=1), **kw)
    for it in (100, 300, 1000):
        t0 = time.perf_counter(); r = ctx.search_sa(dict(sa, max_iter=it, seed=4004), **kw); w = time.perf_counter() - t0
        print("ctx", it, "wall %.3f s" % w, "orders", len(r["trace"]), "ms/order %.3f" % (1000 * w / len(r["trace"])), "nets", len(r["nets"]), flush=True)
