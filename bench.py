#!/usr/bin/env python
"""bench.py -- family log-lik scores/sec of the cnSearchOrder hot path on B200 (BASELINE.json metric).

A step = one per-order evaluation of the structure-search scoring path over one synthetic sample matrix:
enumerate candidate parent sets -> family tables + log-lik (sm_100a kernels) -> [exchange of the score shards]
-> per-(node, d) selection -> DP over complexity, all on the device.

Workload at every N: BASELINE.json configs[4] ("C5"): synthetic 500 nodes x 4 categories, 1,000,000 samples,
maxParentSet=2, true order; 20,833,250 candidate families, sharded over the GPUs (strong scaling).  The other configs
(C1..C3, and C4 = the cnSearchSA leg) are measured in the `configs` / `sa` blocks of the same line.

  python bench.py [--gpus N --steps K --warmup W]            our arm (N > 1: under torchrun, one rank per GPU, NCCL)
  python bench.py --single-process --gpus N ...              our arm, ONE process driving N GPUs (cnb_mctx: the form an
                                                             R session uses; host threads + NVLink peer stores, no NCCL)
  python bench.py --impl reference [...]                      the reference's own C++ scoring primitive on host cores

value  : whole-job families/s with the packed matrix already resident in HBM (timed with CUDA events on the stream
         every kernel and the collective are launched on; max over ranks)
e2e    : the same metric through the reference-facing C ABI with HOST buffers: the int32 [M][N] matrix in PAGEABLE host
         memory (what R hands over) -> H2D -> pack -> search -> networks exported to the host, all inside the timed region
Both arms draw the SAME matrix (numpy PCG64 on the host), so `config` is identical in both lines.
"""
import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------- workload
def build_workload(name, m_override=None):
    """network + search arguments of a BASELINE config; the sample matrix itself is drawn by draw_samples"""
    from catnet_b200 import synth
    if name in ("C3", "C5"):
        n, c, p = (100, 3, 3) if name == "C3" else (500, 4, 2)
        m = m_override or (100000 if name == "C3" else 1000000)
        seed = 3003 if name == "C3" else 5005
        rng = np.random.Generator(np.random.PCG64(seed))
        net = synth.random_network(n, c, p, rng)
        return {"name": name, "n": n, "m": m, "cats": net["cats"], "net": net, "seed": seed,
                "order": np.asarray(synth.order_nodes_descend(net["parents"])) + 1, "P": p,
                "K": synth.default_max_complexity(n, c, p),
                "desc": "%s: cnSearchOrder on synthetic cnRandomCatnet %d nodes x %d cats, %d samples, maxParentSet=%d, "
                        "true order" % (name, n, c, m, p)}
    if name in ("C1", "C4"):
        net = synth.load_alarm()
        m = m_override or (1000 if name == "C1" else 20000)
        return {"name": name, "n": 37, "m": m, "cats": net["cats"], "net": net, "seed": 1001 if name == "C1" else 4004,
                "order": np.asarray(synth.order_nodes_descend(net["parents"])) + 1, "P": 2,
                "K": synth.default_max_complexity(37, 4, 2) if name == "C1" else 600,
                "desc": "%s: cnSearchOrder on cnSamples(alarmnet, %d), 37 nodes, true order, maxParentSet=2" % (name, m)}
    if name == "C2":
        s = synth.load_breast()
        return {"name": name, "n": 100, "m": 98, "cats": np.full(100, 2, dtype=np.int32), "samples": s, "seed": 0,
                "order": np.arange(1, 101), "P": 3, "K": synth.default_max_complexity(100, 2, 3),
                "desc": "C2: cnSearchOrder on breast (100 genes x 98 samples, 2 cats), maxParentSet=3"}
    raise SystemExit("unknown config " + name)


def draw_samples(wl):
    """the sample matrix of a workload, [M][N] int8, 1-based: ancestral sampling (src/rcatnet.cpp:568-602 semantics) with
    numpy PCG64(seed) on the host -- the SAME generator in both arms, so both score the same matrix"""
    from catnet_b200 import synth
    if "samples" in wl:
        return np.ascontiguousarray(wl["samples"]).astype(np.int8)
    return synth.sample_network(wl["net"], wl["m"], np.random.Generator(np.random.PCG64(wl["seed"])), dtype=np.int8)


def config_of(wl, n_fam_total):
    """identical in both arms"""
    return {"workload": wl["desc"], "families": n_fam_total,
            "l2": "inputs larger than L2 (bit planes %.0f MB)" % (wl["n"] * ((wl["m"] + 31) // 32) * 16 / 1e6),
            "data_rng": "network numpy PCG64(%d); samples numpy PCG64(%d), drawn on the host in both arms" % (wl["seed"], wl["seed"])}


# ------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons of one GPU while the timed region runs: NVML polled every 10 ms from a thread of this
    process (a timed region of a few steps at 8 GPUs lasts ~0.1 s: nvidia-smi's 200 ms loop, the fallback when NVML cannot
    be loaded, would miss it)."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
            0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.rows, self.proc, self.nv, self.smax, self.quit = [], None, None, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = (pynvml, pynvml.nvmlDeviceGetHandleByIndex(int(index)))
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.nv[1], pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv, h = self.nv
        reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.quit:
            try:
                self.rows.append((time.time(), float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(reasons(h))))
            except Exception:
                pass
            time.sleep(0.01)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.nv:
            self.quit = True
            self.t.join(timeout=1.0)
            inside = [r for r in self.rows if t0 <= r[0] <= t1]
            how = "nvml, 10 ms"
            if not inside and self.rows:                      # a region shorter than one poll: the nearest sample
                inside = [min(self.rows, key=lambda r: abs(r[0] - 0.5 * (t0 + t1)))]
                how = "nvml, nearest sample to a region shorter than the polling period"
            mask = 0
            for r in inside:
                mask |= r[2]
            return {"sm_mhz": float(np.median([r[1] for r in inside])) if inside else None, "sm_max_mhz": self.smax,
                    "reasons": sorted(nm for bit, nm in self.BITS.items() if mask & bit), "samples": len(inside), "how": how}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 or ts > t1:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
                for nm, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "how": "nvidia-smi -lms 50"}


# ------------------------------------------------------------------------------------------------- CPU baseline
def draw_families(rng, n, d, order0, k):
    """k random candidate families with d parents: (child labels, parent labels, child positions, parent positions)"""
    ch, pa, chp, pap = [], [], [], []
    for _ in range(k):
        pos = sorted(int(x) for x in rng.choice(n, size=d + 1, replace=False))
        ch.append(order0[pos[-1]])
        pa.append([order0[x] for x in pos[:-1]])
        chp.append(pos[-1])
        pap.append(pos[:-1])
    return (np.asarray(ch, dtype=np.int32), np.asarray(pa, dtype=np.int32), np.asarray(chp, dtype=np.int32),
            np.asarray(pap, dtype=np.int32))


def cpu_reference_rate(samples0, cats, order0, p_max, budget_s, threads):
    """the reference's own scoring primitive (CATNET::setParents + setNodeSampleProb through oracle/_ref) on the
    host cores: `threads` concurrent scorers, each on its own seeded random sample of candidate families at full M.
    Returns (families/s, families, wall, the scored lists with the reference's log-likelihoods)."""
    from oracle import ref
    n = len(cats)
    rng = np.random.default_rng(12345)
    d = p_max
    ch, pa, _, _ = draw_families(rng, n, d, order0, 2)
    t = time.perf_counter()
    ref.score_family_list(samples0, cats, ch, pa)
    per_family = (time.perf_counter() - t) / 2
    k = max(2, int(budget_s / max(per_family, 1e-9)))
    lists = [draw_families(rng, n, d, order0, k) for _ in range(threads)]
    done = [0.0] * threads
    lls = [None] * threads

    def work(i):
        lls[i] = ref.score_family_list(samples0, cats, lists[i][0], lists[i][1])
        done[i] = time.perf_counter()

    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for th in ts:
        th.start()
    for th in ts:
        th.join()
    wall = max(done) - t0
    scored = {"children_pos": np.concatenate([l[2] for l in lists]), "parents_pos": np.concatenate([l[3] for l in lists]),
              "loglik": np.concatenate(lls)}
    return threads * k / wall, threads * k, wall, scored


# ------------------------------------------------------------------------------------------------- cnSearchSA (C4)
def sa_workload():
    """BASELINE config C4: ALARM, 20,000 samples, maxParentSet=2, maxComplexity=600, demo/alarm.r:23-26 schedule"""
    from catnet_b200 import synth
    cfg = synth.config("C4")
    s = np.ascontiguousarray(cfg["samples"], dtype=np.int32)
    order1 = (np.asarray(cfg["order"]) + 1).astype(np.int32)
    start = order1[np.random.default_rng(7).permutation(len(order1))]
    sa = dict(select_mode=2, start_order=start, temp_start=0.0, temp_cool_fact=1.0, temp_check_orders=1000,
              order_shuffles=0.0, stop_diff=1e-10, num_threads=4, use_cache=False)
    return s, cfg["cats"], sa


SA_DESC = ("C4: cnSearchSA on cnSamples(alarmnet, 20000), maxParentSet=2, maxComplexity=600, BIC, maxIter=%d, "
           "numThreads=4 (proposals per step), use_cache=0 (the cache-off result; `score_cache_on` = the same chain with the "
           "reference's score cache modelled, R's setting)")


def bic_best(nets, m):
    """the network cnFindBIC would pick from an evaluation (R/catnet.search.R): max of M loglik - 0.5 log(M) complexity"""
    if not nets:
        return None
    ft = 0.5 * math.log(m)
    best = max(nets, key=lambda x: m * x["loglik"] - ft * x["complexity"])
    return {"complexity": int(best["complexity"]), "loglik": float(best["loglik"]),
            "bic": float(m * best["loglik"] - ft * best["complexity"])}


def sa_leg_torchrun(rank, world, max_iter, dist_mod=None, dev=None):
    """one cnb_search_sa call per rank on ITS OWN GPU (independent chains, seed + rank), host buffers in, networks out;
    the chains are merged per complexity on rank 0 like R's .searchSA merges evaluations (R/catnet.search.R:317-332)"""
    import torch
    from catnet_b200 import _abi, dist as cdist
    s, cats, sa = sa_workload()
    kw = dict(max_parent_set=2, max_complexity=600, num_cats=cats, device=dev.index if dev is not None else 0)
    _abi.search_sa(s, dict(sa, max_iter=3, seed=1), **kw)
    t0 = time.perf_counter()
    r = _abi.search_sa(s, dict(sa, max_iter=max_iter, seed=4004 + rank), **kw)
    wall = time.perf_counter() - t0
    chains = [r["nets"]]
    if world > 1:
        t = torch.tensor([wall], dtype=torch.float64, device=dev)
        dist_mod.all_reduce(t, op=dist_mod.ReduceOp.MAX)
        wall = float(t.item())
        chains = [None] * world
        dist_mod.all_gather_object(chains, [{"complexity": x["complexity"], "loglik": x["loglik"]} for x in r["nets"]])
    merged = cdist.merge_evaluations(chains)
    m = s.shape[0]
    cache_on = None
    if world == 1:
        # R's own setting: every proposal walks the model of the reference's score cache, one after the other, and the
        # families the cache answers are re-scored in their cached listing (csrc/cnb_cache.h)
        t0 = time.perf_counter()
        rc = _abi.search_sa(s, dict(sa, max_iter=max_iter, seed=4004, use_cache=True), **kw)
        wc = time.perf_counter() - t0
        cache_on = {"wall_s": wc, "orders_evaluated": int(len(rc["trace"])), "ms_per_order": 1000.0 * wc / max(len(rc["trace"]), 1),
                    "bic_best": bic_best(rc["nets"], m),
                    "same_trajectory_as_cache_off": bool(np.array_equal(rc["trace_accept"], r["trace_accept"])
                                                         and np.array_equal(rc["order"], r["order"]))}
    return {"score_cache_on": cache_on,
            "workload": (SA_DESC % max_iter) + ", one chain per GPU, merged per complexity",
            "wall_s": wall, "chains": world, "iterations": int(r["niter"]), "orders_evaluated": int(len(r["trace"])),
            "ms_per_order": 1000.0 * wall / max(len(r["trace"]), 1),
            "chain_bic_best": [bic_best(c, m) for c in chains], "merged_bic_best": bic_best(merged, m),
            "merged_networks": len(merged),
            "best_loglik": float(max(x["loglik"] for x in merged)) if merged else None}


def sa_leg_single_process(ngpus, max_iter):
    """the same through ONE call of the drop-in entry point with cnb_params.num_devices (what CNB_NUM_DEVICES gives an R
    session): one chain per device inside the library, merged there"""
    from catnet_b200 import _abi
    s, cats, sa = sa_workload()
    kw = dict(max_parent_set=2, max_complexity=600, num_cats=cats, num_devices=ngpus)
    _abi.search_sa(s, dict(sa, max_iter=3, seed=1), **kw)
    t0 = time.perf_counter()
    r = _abi.search_sa(s, dict(sa, max_iter=max_iter, seed=4004), **kw)
    wall = time.perf_counter() - t0
    return {"workload": (SA_DESC % max_iter) + ", cnb_search_sa with num_devices=%d (one chain per device, merged)" % ngpus,
            "wall_s": wall, "chains": ngpus, "iterations": int(r["niter"]), "orders_evaluated_chain0": int(len(r["trace"])),
            "merged_bic_best": bic_best(r["nets"], s.shape[0]), "merged_networks": len(r["nets"]),
            "best_loglik": float(max(x["loglik"] for x in r["nets"])) if r["nets"] else None}


def sa_leg_reference(iters_a=30, iters_b=90):
    """the reference's SA loop around its own estimate(), threads and cache (oracle/_ref), bounded: the first orders are
    the expensive ones (empty cache), so the warm per-order time is reported separately from the difference of two runs
    of the same chain"""
    from oracle import ref
    s, cats, sa = sa_workload()

    def run(it):
        t0 = time.perf_counter()
        r = ref.search_sa(s, 2, 600, sa["start_order"], select_mode=2, temp_start=0.0, temp_cool_fact=1.0,
                          temp_check_orders=1000, max_iter=it, order_shuffles=0.0, stop_diff=1e-10, num_threads=4,
                          use_cache=True, seed=4004, num_cats=cats)
        return time.perf_counter() - t0, r
    wa, ra = run(iters_a)
    wb, rb = run(iters_b)
    na, nb = len(ra["trace"]), len(rb["trace"])
    return {"workload": (SA_DESC % iters_b) + ", cache on (reference core, bounded)",
            "wall_s": wb, "chains": 1, "iterations": int(rb["niter"]), "orders_evaluated": nb,
            "ms_per_order": 1000.0 * wb / max(nb, 1),
            "cold_ms_per_order_first_%d" % na: 1000.0 * wa / max(na, 1),
            "warm_ms_per_order_after_%d" % na: 1000.0 * (wb - wa) / max(nb - na, 1)}


# ------------------------------------------------------------------------------------------------- other configs
def configs_block(device):
    """BASELINE configs C1, C2, C3 on one GPU through the drop-in call (host buffers in, networks on the host out) and on a
    resident context; C1 / C2 are compared with the committed outputs of the compiled reference (tests/golden), C3 -- whose
    reference run takes over an hour -- with the bit-plane kernels of this library (a different kernel family)"""
    from catnet_b200 import _abi, synth
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import compare_golden, load_golden
    L = _abi.lib()
    out = {}
    golden = load_golden("ref_configs")
    for name in ("C1", "C2", "C3"):
        cfg = synth.config(name)
        s = np.ascontiguousarray(cfg["samples"], dtype=np.int32)
        m, n = s.shape
        kw = dict(max_parent_set=cfg["max_parent_set"], max_complexity=cfg["max_complexity"],
                  order=(np.asarray(cfg["order"]) + 1).astype(np.int32), num_cats=cfg["cats"], device=device)
        fam = int(synth.num_families(n, cfg["max_parent_set"]))
        keep = _abi.Keep()
        p = _abi.make_params(keep, n, m, samples=s, **kw)

        def one_shot():
            h = ctypes.c_void_p()
            t0 = time.perf_counter()
            _abi.check(L.cnb_search_order(ctypes.byref(p), ctypes.byref(h)))
            t = time.perf_counter() - t0
            k = L.cnb_result_num_networks(h)
            L.cnb_result_free(h)
            return t, k
        one_shot()
        runs = [one_shot() for _ in range(5)]
        t_one = float(np.median([t for t, _ in runs]))
        nets = _abi.search_order(s, want_probs=False, **kw)
        with _abi.Context(device) as ctx:
            ctx.set_samples(s, num_cats=cfg["cats"])
            kw2 = {k: v for k, v in kw.items() if k not in ("num_cats", "device")}
            L.cnb_result_free(ctx.search_order_raw(**kw2))
            ts = []
            for _ in range(5):
                t0 = time.perf_counter()
                L.cnb_result_free(ctx.search_order_raw(**kw2))
                ts.append(time.perf_counter() - t0)
            t_res = float(np.median(ts))
            tm = ctx.timing()
            check = None
            if name == "C3":
                ctx.force_generic(8 | 1024)                     # no tensor cores: bit-plane POPC kernels
                other = ctx.search_order(want_probs=False, **kw2)
                same = len(other) == len(nets) and all(a["parents"] == b["parents"] and a["loglik"] == b["loglik"]
                                                       for a, b in zip(nets, other))
                check = {"against": "the bit-plane kernels of this library (reference run > 1 h at this size)", "identical": bool(same)}
        if name in golden:
            try:
                k = compare_golden(nets, golden[name])
                check = {"against": "compiled reference (tests/golden/ref_configs.json.gz)", "identical": True, "networks": k}
            except AssertionError as e:
                check = {"against": "compiled reference (tests/golden/ref_configs.json.gz)", "identical": False, "error": str(e)[:200]}
        out[name] = {"nodes": n, "samples": m, "max_parent_set": cfg["max_parent_set"], "families": fam,
                     "one_shot_ms": 1000 * t_one, "resident_ms": 1000 * t_res, "families_per_s": fam / t_one,
                     "families_per_s_resident": fam / t_res, "networks": len(nets), "check": check,
                     "device_ms": {k: round(tm[k], 3) for k in ("plan_ms", "score_ms", "select_ms", "dp_ms", "collect_ms")},
                     "tensor_tiles": int(tm["tensor_tiles"]), "tensor3_tiles": int(tm["tensor3_tiles"])}
    return out


# ------------------------------------------------------------------------------------------------- roofline
def roofline_of(wl, phases, m, world, rank):
    d_dom = wl["P"]
    ph_last = phases[-1]
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    dom_ms = float(np.mean([p["score_ms_by_parents"][d_dom] for p in phases]))
    dom_fam = ph_last["families_by_parents"][d_dom]
    alg_bytes = dom_fam * ((d_dom + 1) * m + 8)
    hbm_model = {"achieved_gbs": alg_bytes / (dom_ms / 1000.0) / 1e9 if dom_ms > 0 else 0.0, "peak_gbs": hbm_peak,
                 "frac": (alg_bytes / (dom_ms / 1000.0) / 1e9 / hbm_peak) if dom_ms > 0 else 0.0,
                 "algorithmic_bytes_per_launch": alg_bytes, "group_ms": dom_ms,
                 "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback (B200_PROFILING.md)"}
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get("%s_gpus%d" % (wl["name"], world))
    except Exception:
        pass
    if ph_last["tensor_tiles"] > 0:
        # tcgen05 scorer: bound by the tensor pipe.  MEASURED_PEAKS.json only has a bf16 cuBLAS figure (nominal dense rates
        # on B200: bf16 2.25, fp8/int8 4.5, fp4 9 PFLOP/s), and 4 x that figure is EXCEEDED by this kernel, so the
        # denominator for E2M1 operands is measured here, live, by tools/umma_peak --quick: back-to-back
        # tcgen05.mma.kind::mxf4 M128 N256 on every SM.  Two builder-independent denominators are quoted beside it:
        # the nominal dense fp4 rate and the pipe arithmetic at the SM clock this run held.
        fp4 = ph_last.get("tensor_kind", 8) == 4
        t_ms = float(np.mean([p["tensor_ms"] for p in phases]))
        ops = 2.0 * ph_last["tensor_macs"]
        bf16 = float(peaks.get("bf16_tflops_sustained", 1400.0))
        ach = ops / (t_ms / 1000.0) / 1e12 if t_ms > 0 else 0.0
        extra = {}
        if fp4:
            live = None
            try:
                if rank == 0:
                    out = subprocess.run([os.path.join(ROOT, "tools", "umma_peak"), "--quick"], capture_output=True, text=True,
                                         timeout=60).stdout.strip().splitlines()[-1]
                    live = json.loads(out)
            except Exception:
                live = None
            if live and "mxf4_ss_n256_tflops" in live:
                peak = float(live["mxf4_ss_n256_tflops"])
                extra = {"ts_ceiling": float(live["mxf4_ts_n192_tflops"]), "ts_clk_per_mma": live["mxf4_ts_n192_clk_per_mma"]}
                src = ("measured live by tools/umma_peak --quick (back-to-back tcgen05.mma kind::mxf4 M128 N256, all SMs); "
                       "MEASURED_PEAKS.json has no fp4 figure (4 x its bf16_tflops_sustained = %.0f)" % (4.0 * bf16))
            else:
                peak = 8835.9
                extra = {"ts_ceiling": 8369.2}
                src = ("tools/umma_peak on this pool, recorded in profiles/r01_umma_peak_b200.txt (the binary did not run "
                       "here); MEASURED_PEAKS.json has no fp4 figure (4 x its bf16_tflops_sustained = %.0f)" % (4.0 * bf16))
            extra["frac_of_nominal_9000"] = ach / 9000.0
            extra["nominal_note"] = ("nominal dense fp4 9 PFLOP/s (B200_PROFILING.md); pipe arithmetic 16384 MAC/clk/SM x 148 SMs x "
                                     "the SM clock of this run is reported as frac_of_pipe_at_clock")
        else:
            peak = 2.0 * bf16
            src = ("2 x bf16_tflops_sustained of MEASURED_PEAKS.json (int8 nominal = 2 x bf16)" if peaks
                   else "2 x 1400 TFLOP/s fallback (B200_PROFILING.md sustained bf16)")
        useful_cells = 27.0
        roofline = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "traffic": traffic,
                    "kernel": "k_umma_tables (tcgen05.mma %s)" % ("kind::mxf4.block_scale" if fp4 else "kind::i8"),
                    "kernel_ms": t_ms, "ops_per_launch": ops, "tiles": int(ph_last["tensor_tiles"]),
                    "peak_source": src,
                    "useful_mac_fraction": dom_fam * useful_cells * m / max(ph_last["tensor_macs"], 1),
                    "hbm_model": hbm_model}
        roofline.update(extra)
    else:
        roofline = {"bound": "hbm", "achieved": hbm_model["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s",
                    "frac": hbm_model["frac"], "traffic": traffic,
                    "kernel": "k_score_planes%s<%d,%d>" % ("_ie2" if d_dom == 2 else "", d_dom, min(4, max(2, int(max(wl["cats"]))))),
                    "kernel_ms": dom_ms, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": hbm_model["peak_source"]}
    roofline["family_samples_per_s"] = dom_fam * m / (dom_ms / 1000.0) if dom_ms > 0 else 0.0
    return roofline


def add_pipe_fraction(roofline, clocks):
    """the tensor pipe's own arithmetic at the clock this run held: 16384 E2M1 MAC/clk/SM x 148 SMs"""
    if roofline.get("bound") == "tensor" and clocks.get("sm_mhz") and "frac_of_nominal_9000" in roofline:
        pipe = 2.0 * 16384 * 148 * clocks["sm_mhz"] * 1e6 / 1e12
        roofline["pipe_at_clock_tflops"] = pipe
        roofline["frac_of_pipe_at_clock"] = roofline["achieved"] / pipe


# ------------------------------------------------------------------------------------------------- reference arm
def reference_arm(args, wl, n_fam_total, threads, rank):
    if rank != 0:
        return 0
    s = draw_samples(wl)
    s0 = np.ascontiguousarray(s.astype(np.int32) - 1)
    order0 = wl["order"] - 1
    for _ in range(max(args.warmup, 0)):
        cpu_reference_rate(s0, wl["cats"], order0, wl["P"], 0.5, threads)
    rates, t_all = [], 0.0
    per_step = max(2.0, min(20.0, 120.0 / max(args.steps, 1)))
    nfam = 0
    for _ in range(args.steps):
        r, k, wall, _ = cpu_reference_rate(s0, wl["cats"], order0, wl["P"], per_step, threads)
        rates.append(r)
        t_all += wall
        nfam = k
    v = float(np.mean(rates))
    sample = "%d random candidate families with %d parents at full M=%d per step, %d concurrent reference scorers" % (
        nfam, wl["P"], wl["m"], threads)
    sa_ref = None
    if not args.no_sa:
        try:
            sa_ref = sa_leg_reference()
        except Exception as e:
            sa_ref = {"unavailable": str(e)}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "families/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * t_all / max(args.steps, 1),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(wl, n_fam_total),
        "cpu_baseline": {"value": v, "unit": "families/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": "families/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "sa": sa_ref}))
    return 0


METRIC = "family log-lik scores/sec (cnSearchOrder)"


# ------------------------------------------------------------------------------------------------- one process, N GPUs
def single_process_block(ngpus, wl, host_i32, steps, warmup, n_fam_total, want_e2e=True):
    """cnb_mctx: ONE process drives ngpus devices (host threads, NVLink peer stores into the first device's score buffer).
    Device time between two marks on the first device's stream, every device synchronised."""
    from catnet_b200 import _abi
    L = _abi.lib()
    kw = dict(max_parent_set=wl["P"], max_complexity=wl["K"], order=wl["order"].astype(np.int32))
    out = {"n_gpus": ngpus, "how": "one process, cnb_mctx: a host thread per device, scoring kernels store scores into device 0 "
                                  "over NVLink peer memory, selection + DP + export on device 0 only"}
    with _abi.MultiContext(0, ngpus) as mc:
        t0 = time.perf_counter()
        mc.set_samples(host_i32, num_cats=wl["cats"])
        out["set_samples_first_ms"] = 1000 * (time.perf_counter() - t0)
        for _ in range(max(warmup, 3)):
            mc.search_light(**kw)
        mc.mark(0)
        for _ in range(steps):
            mc.search_light(**kw)
        mc.mark(1)
        ms = mc.elapsed_ms()
        out.update({"value": n_fam_total * steps / (ms / 1000.0), "unit": "families/s", "ms_per_step": ms / steps,
                    "steps": steps, "active_devices": mc.active(),
                    "tensor_ms_per_device": [round(mc.timing(i)["tensor_ms"], 3) for i in range(mc.g)],
                    "phases_ms_device0": {k: mc.timing(0)[k] for k in ("plan_ms", "score_ms", "select_ms", "dp_ms")}})
        h = mc.collect_raw()
        nn = L.cnb_result_num_networks(h)
        cx, ll = ctypes.c_int32(), ctypes.c_double()
        L.cnb_result_network(h, nn - 1, ctypes.byref(cx), ctypes.byref(ll))
        L.cnb_result_free(h)
        out["result"] = {"networks": nn, "max_complexity": cx.value, "loglik": ll.value}
    if want_e2e:
        # THE drop-in call an R session makes: cnb_search_order with the pageable host matrix in, networks on the host out,
        # cnb_params.num_devices = ngpus (or CNB_NUM_DEVICES in the environment).  Inside: slices of the samples dealt over
        # the devices, peer copies, one table-kernel pass per slice on every device (cnb_multi.cu::mctx_search_streamed).
        keep = _abi.Keep()
        p1 = _abi.make_params(keep, host_i32.shape[1], host_i32.shape[0], samples=host_i32, num_cats=wl["cats"], device=0,
                              num_devices=ngpus, **kw)

        def e2e_step():
            h = ctypes.c_void_p()
            _abi.check(L.cnb_search_order(ctypes.byref(p1), ctypes.byref(h)))
            k = L.cnb_result_num_networks(h)
            L.cnb_result_free(h)
            return k
        e2e_step()
        e_steps = max(1, min(steps, 3))
        t0 = time.perf_counter()
        for _ in range(e_steps):
            k = e2e_step()
        ems = 1000 * (time.perf_counter() - t0)                # the call returns synchronised: host wall time is the span
        tm = _abi.last_call_timing()
        out["e2e"] = {"value": n_fam_total * e_steps / (ems / 1000.0), "unit": "families/s", "ms_per_step": ems / e_steps,
                      "steps": e_steps, "host_memory": "pageable", "timed": "host wall clock around the synchronous calls",
                      "call": "cnb_search_order(num_devices=%d), %s" % (ngpus, "streamed" if tm["fast_path"] == 2 else "resident path"),
                      "h2d_bytes_per_step": int(tm["h2d_bytes"]), "d2h_bytes_per_step": int(tm["d2h_bytes"]),
                      "device0_ms": {k_: float(tm[k_]) for k_ in ("pack_ms", "score_ms", "select_ms", "dp_ms", "collect_ms")},
                      "networks": k}
    return out


# ------------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C5")
    ap.add_argument("--samples", type=int, default=None, help="override M (debug only; the line is then not a bench line)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--kernel-bits", type=int, default=0,
                    help="debug: cnb_ctx_force_generic bits (8 = no tensor cores, 16 = u8 operands); not a bench line")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-sa", action="store_true", help="skip the cnSearchSA wall-time leg (BASELINE config C4)")
    ap.add_argument("--sa-iter", type=int, default=1000)
    ap.add_argument("--no-configs", action="store_true", help="skip the C1..C3 block")
    ap.add_argument("--single-process", action="store_true",
                    help="ONE process drives --gpus N devices through cnb_mctx (no torchrun, no NCCL): the R-session form")
    ap.add_argument("--no-single-process-block", action="store_true",
                    help="under torchrun: skip the extra single-process measurement rank 0 makes after the main one")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != max(args.gpus, 1) and world > 1:
        log("warning: WORLD_SIZE %d != --gpus %d" % (world, args.gpus))
    wl = build_workload(args.config, args.samples)
    n_fam_total = sum(math.comb(wl["n"], d + 1) for d in range(1, wl["P"] + 1))
    threads = os.cpu_count() or 1

    if args.impl == "reference":
        return reference_arm(args, wl, n_fam_total, threads, rank)

    import torch
    import torch.distributed as dist
    from catnet_b200 import _abi
    _abi.lib()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the scoring path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cpu_group = None
    if world > 1:
        from catnet_b200 import dist as cdist
        dist.init_process_group("nccl", device_id=dev, pg_options=cdist.nccl_options())
        cpu_group = dist.new_group(backend="gloo")       # host-side barrier while rank 0 measures the single-process form

    # ---- synthetic data: rank 0 draws on the host (numpy, the same generator as the reference arm), everybody receives it
    t0 = time.time()
    m, n = wl["m"], wl["n"]
    host8 = None
    if rank == 0:
        host8 = draw_samples(wl)
        samples_dev8 = torch.as_tensor(host8.view(np.uint8), device=dev)
    else:
        samples_dev8 = torch.empty((m, n), dtype=torch.uint8, device=dev)
    if world > 1:
        dist.broadcast(samples_dev8, 0)
    torch.cuda.synchronize()
    log("rank %d: data %s ready in %.1fs" % (rank, tuple(samples_dev8.shape), time.time() - t0))
    kw = dict(max_parent_set=wl["P"], max_complexity=wl["K"], order=wl["order"].astype(np.int32))

    if args.single_process:
        if world > 1:
            raise SystemExit("--single-process is launched WITHOUT torchrun")
        host_i32 = np.ascontiguousarray(host8.astype(np.int32))          # pageable, as R's matrix
        del samples_dev8
        sampler = ClockSampler(0)
        tw0 = time.time()
        sp = single_process_block(args.gpus, wl, host_i32, args.steps, args.warmup, n_fam_total, not args.no_e2e)
        tw1 = time.time()
        clocks = sampler.stop(tw0, tw1)
        sa_res = None
        if not args.no_sa and args.config == "C5" and args.samples is None:
            try:
                sa_res = sa_leg_single_process(args.gpus, args.sa_iter)
            except Exception as e:
                sa_res = {"unavailable": str(e)}
        line = {"metric": METRIC, "value": sp["value"], "unit": "families/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": sp["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "e2m1 one-hot operands, f32 accumulators holding exact counts, f64 log-lik",
                "data": "synthetic", "config": config_of(wl, n_fam_total), "clocks": clocks, "mode": "single-process",
                "e2e": sp.get("e2e"), "single_process": sp, "sa": sa_res}
        print(json.dumps(line))
        return 0

    ctx = _abi.Context(local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.set_samples_dev8(samples_dev8.data_ptr(), n, m, num_cats=wl["cats"])
    if args.kernel_bits:
        ctx.force_generic(args.kernel_bits)
    seg, nfam = ctx.plan(rank, world, **kw)
    assert nfam == n_fam_total, (nfam, n_fam_total)
    scores = torch.empty(max(seg * world, 1), dtype=torch.float64, device=dev)
    my_seg = scores[rank * seg:(rank + 1) * seg]

    def gather():
        # in place: rank r's segment already sits at offset r of the output (ncclAllGather's in-place form)
        if world > 1:
            dist.all_gather_into_tensor(scores, my_seg)

    def step():
        ctx.plan(rank, world, **kw)
        ctx.score_shard(scores.data_ptr())
        gather()
        ctx.select_dp(scores.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local_rank)
    time.sleep(0.25)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phases = []
    tw0 = time.time()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
        phases.append(ctx.timing())
    ev1.record(stream)
    barrier()
    tw1 = time.time()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop(tw0, tw1)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    value = n_fam_total * args.steps / (ms / 1000.0)

    roofline = roofline_of(wl, phases, m, world, rank)
    add_pipe_fraction(roofline, clocks)

    # ---- the reference's CPU path beside it (rank 0, N=1 only) + parity of the benchmarked kernel's own scores
    cpu, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            s0 = np.ascontiguousarray(host8.astype(np.int32) - 1)
            r, k, wall, scored = cpu_reference_rate(s0, wl["cats"], wl["order"] - 1, wl["P"], args.cpu_budget, threads)
            del s0
            cpu = {"value": r, "unit": "families/s", "cores": threads, "kind": "reference",
                   "sample": "%d random candidate families with %d parents at full M=%d, %d concurrent reference "
                             "scorers (CATNET::setParents+setNodeSampleProb), %.1fs wall" % (k, wl["P"], m, threads, wall)}
            # the SAME families as the timed search scored them (tile slots of the tcgen05 tables), read from the score
            # buffer of the last timed step, against the reference's values on the same matrix: bit equality
            ours = ctx.search_scores(scores.data_ptr(), scored["children_pos"], scored["parents_pos"])
            bad = int(np.sum(ours.view(np.int64) != scored["loglik"].view(np.int64)))
            parity = {"families": int(len(ours)), "mismatches": bad,
                      "what": "scores of the timed search (k_umma_tables + k_umma_epilogue) vs the compiled reference's "
                              "setNodeSampleProb on the same matrix, fp64 bit equality (equal log-lik <=> equal counts summed "
                              "in the same order)"}
            if bad:
                worst = int(np.argmax(np.abs(ours - scored["loglik"])))
                parity["first_mismatch"] = {"child_pos": int(scored["children_pos"][worst]),
                                            "parents_pos": [int(v) for v in scored["parents_pos"][worst]],
                                            "ours": float(ours[worst]), "reference": float(scored["loglik"][worst])}
        except Exception as e:  # the checker library is optional for the product path
            cpu = {"value": None, "unit": "families/s", "cores": threads, "kind": "reference", "sample": "unavailable: %s" % e}

    # ---- e2e through the C ABI with host buffers (pageable, as R's)
    e2e = None
    if not args.no_e2e:
        hnp = np.empty((m, n), dtype=np.int32)
        if rank == 0:
            hnp[:] = host8
        else:
            hnp[:] = samples_dev8.cpu().numpy().view(np.int8)
        host = torch.from_numpy(hnp)
        ctx2 = _abi.Context(local_rank)
        ctx2.set_stream(stream.cuda_stream)
        L = _abi.lib()
        wall = {"set_samples_ms": 0.0, "plan_score_ms": 0.0, "finish_ms": 0.0, "h2d_bytes": 0}

        keep1 = _abi.Keep()
        p1 = _abi.make_params(keep1, n, m, samples=hnp, num_cats=wl["cats"], device=local_rank, **kw) if world == 1 else None

        def e2e_step():
            t_a = time.perf_counter()
            if world == 1:
                # THE drop-in call: cnb_search_order with the host matrix in, networks on the host out (inside the library the
                # upload is cut into slices and overlapped with the table kernel)
                h = ctypes.c_void_p()
                _abi.check(L.cnb_search_order(ctypes.byref(p1), ctypes.byref(h)))
                nn = L.cnb_result_num_networks(h)
                cx, ll = ctypes.c_int32(), ctypes.c_double()
                L.cnb_result_network(h, nn - 1, ctypes.byref(cx), ctypes.byref(ll))
                L.cnb_result_free(h)
                wall["plan_score_ms"] += (time.perf_counter() - t_a) * 1e3
                return (nn, cx.value, ll.value)
            # every rank narrows and copies 1/world of every slice of the host matrix, an NVLink all-gather completes the slice,
            # the table kernel contracts it while the next slice is on its way (catnet_b200/dist.py::search_order_streamed);
            # selection + DP + network export (D2H) on rank 0 only
            from catnet_b200 import dist as cdist
            h, info = cdist.search_order_streamed(ctx2, host, rank, world, dev, wl["cats"], finish_rank=0, raw=True, **kw)
            wall["h2d_bytes"] = info["h2d_bytes"]
            wall["streamed"] = info["streamed"]
            res = None
            if rank == 0:
                nn = L.cnb_result_num_networks(h)
                cx, ll = ctypes.c_int32(), ctypes.c_double()
                L.cnb_result_network(h, nn - 1, ctypes.byref(cx), ctypes.byref(ll))
                L.cnb_result_free(h)
                res = (nn, cx.value, ll.value)
            wall["plan_score_ms"] += (time.perf_counter() - t_a) * 1e3
            return res

        e2e_step()
        barrier()
        for k_ in ("set_samples_ms", "plan_score_ms", "finish_ms"):
            wall[k_] = 0.0
        e_steps = max(1, min(args.steps, 3))
        ev0.record(stream)
        for _ in range(e_steps):
            res = e2e_step()
        ev1.record(stream)
        barrier()
        ems = ev0.elapsed_time(ev1)
        if world == 1:
            ems = max(ems, sum(wall[k_] for k_ in ("set_samples_ms", "plan_score_ms", "finish_ms")))   # the call returns synchronised
        t_ms = torch.tensor([ems], dtype=torch.float64, device=dev)
        tm = ctx2.timing() if world > 1 else _abi.last_call_timing()
        # streamed: the library counted the uploaded slices itself; resident fall-back: set_samples_dev8 restarts the count
        traffic = torch.tensor([int(tm["h2d_bytes"]) + (0 if wall.get("streamed") or world == 1 else int(wall["h2d_bytes"])),
                                int(tm["d2h_bytes"])], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(traffic, op=dist.ReduceOp.SUM)
        ems = float(t_ms.item())
        if rank == 0:
            e2e = {"value": n_fam_total * e_steps / (ems / 1000.0), "unit": "families/s",
                   "h2d_bytes_per_step": int(traffic[0].item()), "d2h_bytes_per_step": int(traffic[1].item()),
                   "ms_per_step": ems / e_steps, "steps": e_steps, "host_memory": "pageable",
                   "call": ("cnb_search_order (one-shot C ABI call, %s)" % ("streamed: upload overlapped with the table kernel"
                                                                          if tm.get("fast_path") == 2 else "resident path")
                            if world == 1 else ("dist.search_order_streamed: per slice upload_rows8_on + all-gather + stream_slice (pack, table "
                                                "kernel pass); stream_end, score all-gather, finish on rank 0" if wall.get("streamed") else
                                                "upload_rows8 + all-gather + set_samples_dev8 + plan + score_shard + all-gather + finish on rank 0")),
                   "host_wall_ms_per_step": {k_: wall[k_] / e_steps for k_ in ("set_samples_ms", "plan_score_ms", "finish_ms")},
                   "device_ms": {k_: float(tm[k_]) for k_ in ("pack_ms", "score_ms", "select_ms", "dp_ms", "collect_ms")},
                   "result": {"networks": res[0], "max_complexity": res[1], "loglik": res[2]}}
        ctx2.close()
        del host, hnp

    # ---- cnSearchSA wall time (second half of the BASELINE metric): one chain per GPU, merged
    sa_res = None
    if not args.no_sa and args.config == "C5" and args.samples is None:
        try:
            sa_res = sa_leg_torchrun(rank, world, args.sa_iter, dist if world > 1 else None, dev)
        except Exception as e:
            sa_res = {"unavailable": str(e)}

    ph = phases[-1]
    ctx.close()
    del scores, my_seg, samples_dev8
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs (rank 0 at N = 1) and the single-process form of this very workload (rank 0, N > 1)
    configs = None
    if rank == 0 and world == 1 and not args.no_configs and args.config == "C5" and args.samples is None:
        try:
            configs = configs_block(local_rank)
        except Exception as e:
            configs = {"unavailable": str(e)}
    single = None
    if world > 1 and not args.no_single_process_block:
        dist.barrier(group=cpu_group)                       # every rank has released its GPU; the others now wait on the host
        if rank == 0:
            try:
                host_i32 = np.ascontiguousarray(host8.astype(np.int32))
                single = single_process_block(world, wl, host_i32, args.steps, args.warmup, n_fam_total, not args.no_e2e)
                del host_i32
                if not args.no_sa and args.config == "C5" and args.samples is None:
                    single["sa"] = sa_leg_single_process(world, args.sa_iter)
            except Exception as e:
                single = {"unavailable": str(e)}
        dist.barrier(group=cpu_group)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "families/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": ("e2m1 one-hot operands, f32 accumulators holding exact counts, f64 log-lik"
                      if ph.get("tensor_kind", 0) == 4 else "u32 counts + f64 log-lik"), "data": "synthetic",
            "config": config_of(wl, n_fam_total), "families_per_rank": int(ph["num_families"]),
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "parity_spot": parity, "e2e": e2e,
            "gpu_launches": int(ph["kernel_launches"]) * args.steps,
            "phases_ms": {k: ph[k] for k in ("plan_ms", "score_ms", "select_ms", "dp_ms")},
            "sa": sa_res, "configs": configs, "single_process": single,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if parity and parity["mismatches"]:
        log("PARITY FAILURE: %d of %d spot-checked families differ from the reference" % (parity["mismatches"], parity["families"]))
        return 3
    return 0


if __name__ == "__main__":
    sys.exit(main())
