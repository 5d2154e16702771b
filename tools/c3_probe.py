import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from catnet_b200 import _abi, synth
cfg = synth.config("C3")
s = np.ascontiguousarray(cfg["samples"], dtype=np.int32)
order1 = (np.asarray(cfg["order"]) + 1).astype(np.int32)
kw = dict(max_parent_set=3, max_complexity=cfg["max_complexity"], order=order1)
for bits in (0, 1024, 1024 | 512):
    with _abi.Context(0) as ctx:
        ctx.set_samples(s, num_cats=cfg["cats"])
        if bits:
            ctx.force_generic(bits)
        for rep in range(2):
            t0 = time.perf_counter()
            h = ctx.search_order_raw(**kw)
            w = time.perf_counter() - t0
            _abi.lib().cnb_result_free(h)
            tm = ctx.timing()
            print("bits", bits, "rep", rep, "wall %.1f ms" % (1000 * w), {k: round(tm[k], 2) for k in ("pack_ms", "plan_ms", "score_ms", "select_ms", "dp_ms", "collect_ms")},
                  "by parents", [round(v, 2) for v in tm["score_ms_by_parents"][:4]], "launches", tm["kernel_launches"], "tiles", tm["tensor_tiles"], "tiles3", tm["tensor3_tiles"], flush=True)
