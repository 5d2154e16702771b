import sys, time, numpy as np, os
sys.path.insert(0, '/root/repo')
from catnet_b200 import _abi
import bench
import torch
s, cats, sa = bench.sa_workload()
kw = dict(max_parent_set=2, max_complexity=600)
def run(tag):
    _abi.search_sa(s, dict(sa, max_iter=3, seed=1), num_cats=cats, **kw)
    t0 = time.perf_counter(); r = _abi.search_sa(s, dict(sa, max_iter=300, seed=4004), num_cats=cats, **kw); w = time.perf_counter() - t0
    print(tag, "ms/order %.3f" % (1000 * w / len(r["trace"])), flush=True)
dev = torch.device("cuda", 0)
x = torch.zeros(10, device=dev); torch.cuda.synchronize()
time.sleep(5)
run("torch context + 5 s")
big = torch.randint(1, 5, (1000000, 500), dtype=torch.int32, device=dev)
ctx = _abi.Context(0)
stream = torch.cuda.current_stream()
ctx.set_stream(stream.cuda_stream)
ctx.set_samples_dev(big.data_ptr(), 500, 1000000)
kw5 = dict(max_parent_set=2, max_complexity=23999)
seg, nf = ctx.plan(0, 1, **kw5)
scores = torch.empty(seg, dtype=torch.float64, device=dev)
for _ in range(3):
    ctx.plan(0, 1, **kw5); ctx.score_shard(scores.data_ptr()); ctx.select_dp(scores.data_ptr())
torch.cuda.synchronize()
run("after C5 steps on the legacy stream")
host = torch.empty((1000000, 500), dtype=torch.int32, pin_memory=True)
host.copy_(big); torch.cuda.synchronize()
run("2 GB pinned host tensor live")
ctx2 = _abi.Context(0)
ctx2.set_stream(stream.cuda_stream)
ctx2.set_samples(host.numpy())
ctx2.plan(0, 1, **kw5); ctx2.score_shard(scores.data_ptr()); h = ctx2.finish_raw(scores.data_ptr()); _abi.lib().cnb_result_free(h)
run("after e2e step")
ctx2.close()
run("ctx2 closed")
try:
    s0 = np.ascontiguousarray((big[:20000] - 1).cpu().numpy())
    r = bench.cpu_reference_rate(s0, np.full(500, 4, dtype=np.int32), np.arange(500), 2, 3.0, 16)
    print("cpu ref", r[:2])
except Exception as e:
    print("cpu ref failed", e)
run("after cpu reference threads")
print(bench.sa_leg_ours(0, 1, 300))
