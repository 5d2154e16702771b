"""A few cnSearchSA evaluations on C4 for an ncu launch list (kernel durations of one evaluated order with use_cache 0 and
1): ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file out.csv python tools/sa_ncu.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import bench  # noqa: E402
from catnet_b200 import _abi  # noqa: E402

s, cats, sa = bench.sa_workload()
kw = dict(max_parent_set=2, max_complexity=600, num_cats=cats)
for cache in (False, True):
    r = _abi.search_sa(s, dict(sa, max_iter=8, seed=4004, use_cache=cache, num_threads=1), **kw)
    print("use_cache", cache, len(r["trace"]), "orders", flush=True)
