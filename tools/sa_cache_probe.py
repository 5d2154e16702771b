"""C4 (ALARM, 20,000 samples) cnSearchSA with use_cache 0 / 1: wall time per evaluated order, whether the chains agree.
Needs a B200."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import bench  # noqa: E402
from catnet_b200 import _abi  # noqa: E402

s, cats, sa = bench.sa_workload()
kw = dict(max_parent_set=2, max_complexity=600, num_cats=cats)
_abi.search_sa(s, dict(sa, max_iter=3, seed=1), **kw)
for threads, cache, pipe in ((1, False, 1), (1, True, 1), (2, True, 1), (4, False, 1), (4, True, 1), (4, True, 0)):
    if pipe:
        os.environ.pop("CNB_CACHE_NO_PIPELINE", None)
    else:
        os.environ["CNB_CACHE_NO_PIPELINE"] = "1"
    _abi.search_sa(s, dict(sa, max_iter=8, seed=1, use_cache=cache, num_threads=threads), **kw)
    t0 = time.perf_counter()
    r = _abi.search_sa(s, dict(sa, max_iter=1000, seed=4004, use_cache=cache, num_threads=threads), **kw)
    w = time.perf_counter() - t0
    print("threads %d use_cache %d pipelined %d: wall %.3f s, %d orders, %.3f ms/order, best loglik %.12f" % (
        threads, cache, pipe, w, len(r["trace"]), 1000 * w / len(r["trace"]), max(n["loglik"] for n in r["nets"])), flush=True)
