import sys, numpy as np, torch, subprocess, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
if len(sys.argv) > 1:
    from catnet_b200 import _abi as abi
    from helpers import random_problem
    n, m, bits, P, rnd = [int(v) for v in sys.argv[1:6]]
    s, c = random_problem(11 * n + m, n, m, 3)
    order = (np.random.default_rng(n + m).permutation(n) + 1) if rnd else np.arange(1, n + 1)
    kw = dict(max_parent_set=P, max_complexity=n * 30, order=order)
    with abi.Context(0) as ctx:
        ctx.set_samples(s, num_cats=c)
        ctx.force_generic(bits)
        seg, nf = ctx.plan(0, 1, **kw)
        buf = torch.full((seg,), float("nan"), dtype=torch.float64, device="cuda")
        ctx.score_shard(buf.data_ptr())
        torch.cuda.synchronize()
        print("ok", sys.argv[1:], "nan", int(torch.isnan(buf).sum()))
else:
    for args in ((70, 200, 4 | 1024, 3, 1), (70, 200, 4 | 1024, 2, 1), (70, 200, 4, 3, 1), (64, 200, 4 | 1024, 3, 1), (40, 200, 4 | 1024, 3, 1), (70, 200, 4 | 1024 | 256, 3, 1)):
        r = subprocess.run([sys.executable, __file__] + [str(v) for v in args], capture_output=True, text=True)
        print(args, (r.stdout.strip().splitlines() or ["-"])[-1], "|", (r.stderr.strip().splitlines() or ["-"])[-1][-90:], flush=True)
