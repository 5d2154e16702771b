import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from catnet_b200 import _abi
import bench
s, cats, sa = bench.sa_workload()
kw = dict(max_parent_set=2, max_complexity=600)
def run(tag):
    _abi.search_sa(s, dict(sa, max_iter=3, seed=1), num_cats=cats, **kw)
    t0 = time.perf_counter(); r = _abi.search_sa(s, dict(sa, max_iter=300, seed=4004), num_cats=cats, **kw); w = time.perf_counter() - t0
    print(tag, "ms/order %.3f" % (1000 * w / len(r["trace"])), flush=True)
run("plain")
import torch
x = torch.zeros(10, device="cuda"); torch.cuda.synchronize()
run("torch context")
y = torch.empty(20 << 30, dtype=torch.uint8, device="cuda"); torch.cuda.synchronize()
run("20 GB torch tensor live")
big = torch.randint(1, 5, (200000, 500), dtype=torch.int32, device="cuda")
ctx = _abi.Context(0)
ctx.set_samples_dev(big.data_ptr(), 500, 200000)
run("second cnb context live")
r = ctx.search_order(want_probs=False, max_parent_set=2, max_complexity=23999)
run("after a C5-like search on it")
sampler = bench.ClockSampler(0)
time.sleep(1.0)
run("clock sampler running")
sampler.stop(0, 1)
run("clock sampler stopped")
th = __import__("os").cpu_count()
print("cpus", th)
