#!/usr/bin/env python
"""bench.py -- family log-lik scores/sec of the cnSearchOrder hot path on B200 (BASELINE.json metric).

A step = one per-order evaluation of the structure-search scoring path over one synthetic sample matrix:
enumerate candidate parent sets -> family tables + log-lik (sm_100a kernels) -> [all-gather of the score shards]
-> per-(node, d) selection -> DP over complexity, all on the device.

Workload at every N: BASELINE.json configs[4] ("C5"): synthetic 500 nodes x 4 categories, 1,000,000 samples,
maxParentSet=2, true order; 20,833,250 candidate families, sharded over the ranks (strong scaling).
Other configs (--config C1..C4) are parity-test shapes, not bench lines.

  python bench.py [--gpus N --steps K --warmup W]            our arm
  python bench.py --impl reference [...]                      the reference's own C++ scoring primitive on host cores
  torchrun ... bench.py --gpus N ...                          one rank per GPU, NCCL all-gather of the score shards

value  : whole-job families/s with the packed matrix already resident in HBM (timed with CUDA events on the stream
         every kernel and the collective are launched on; max over ranks)
e2e    : the same metric through the reference-facing C ABI with HOST buffers: int32 [M][N] matrix in pinned host
         memory -> H2D -> pack -> search -> networks exported to the host, all inside the timed region
"""
import argparse
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
    """network + search arguments of a BASELINE config; the sample matrix itself is drawn later (GPU or numpy)"""
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


def sample_on_device(wl, device):
    """ancestral sampling (src/rcatnet.cpp:568-602 semantics) with torch on the GPU: [M][N] int32, 1-based"""
    import torch
    from catnet_b200 import synth
    net, m, n = wl["net"], wl["m"], wl["n"]
    g = torch.Generator(device=device)
    g.manual_seed(wl["seed"])
    out = torch.zeros((m, n), dtype=torch.int32, device=device)
    for i in synth.order_nodes_descend(net["parents"]):
        cfg = torch.zeros(m, dtype=torch.int64, device=device)
        for p in net["parents"][i]:
            cfg = cfg * int(net["cats"][p]) + (out[:, p].to(torch.int64) - 1)
        cum = torch.as_tensor(np.cumsum(net["cpts"][i], axis=1), device=device)[cfg]
        u = torch.rand(m, generator=g, device=device, dtype=torch.float64)
        k = (u[:, None] > cum).sum(dim=1).clamp_(max=int(net["cats"][i]) - 1)
        out[:, i] = (k + 1).to(torch.int32)
    return out


# ------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
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
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------- CPU baseline
def cpu_reference_rate(samples0, cats, order0, p_max, budget_s, threads):
    """the reference's own scoring primitive (CATNET::setParents + setNodeSampleProb through oracle/_ref) on the
    host cores: `threads` concurrent scorers, each on its own seeded random sample of candidate families at full M."""
    from oracle import ref
    n, m = len(cats), samples0.shape[0]
    rng = np.random.default_rng(12345)
    # calibrate: 1 thread, a few families
    d = p_max
    def draw(k):
        ch, pa = [], []
        for _ in range(k):
            pos = sorted(rng.choice(n, size=d + 1, replace=False))
            ch.append(order0[pos[-1]])
            pa.append([order0[x] for x in pos[:-1]])
        return np.asarray(ch, dtype=np.int32), np.asarray(pa, dtype=np.int32)
    ch, pa = draw(2)
    t = time.perf_counter()
    ref.score_family_list(samples0, cats, ch, pa)
    per_family = (time.perf_counter() - t) / 2
    k = max(2, int(budget_s / max(per_family, 1e-9)))
    lists = [draw(k) for _ in range(threads)]
    done = [0.0] * threads

    def work(i):
        ref.score_family_list(samples0, cats, lists[i][0], lists[i][1])
        done[i] = time.perf_counter()

    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for th in ts:
        th.start()
    for th in ts:
        th.join()
    wall = max(done) - t0
    return threads * k / wall, threads * k, wall


# ------------------------------------------------------------------------------------------------- cnSearchSA (C4)
def sa_workload():
    """BASELINE config C4: ALARM, 20,000 samples, maxParentSet=2, maxComplexity=600, demo/alarm.r:23-26 schedule"""
    from catnet_b200 import synth
    cfg = synth.config("C4")
    s = np.ascontiguousarray(cfg["samples"], dtype=np.int32)
    order1 = (np.asarray(cfg["order"]) + 1).astype(np.int32)
    start = order1[np.random.default_rng(7).permutation(len(order1))]
    sa = dict(select_mode=2, start_order=start, temp_start=0.0, temp_cool_fact=1.0, temp_check_orders=1000,
              order_shuffles=0.0, stop_diff=1e-10, num_threads=4, use_cache=True)
    return s, cfg["cats"], sa


def sa_leg_ours(rank, world, max_iter, dist_mod=None, dev=None):
    """wall time of one cnb_search_sa call per GPU (independent chains, seed + rank), host buffers in, networks out"""
    import torch
    from catnet_b200 import _abi
    s, cats, sa = sa_workload()
    # every rank's chain on ITS OWN GPU (cnb_params.device): on one shared device the processes time-slice
    kw = dict(max_parent_set=2, max_complexity=600, num_cats=cats, device=dev.index if dev is not None else 0)
    _abi.search_sa(s, dict(sa, max_iter=3, seed=1), **kw)
    t0 = time.perf_counter()
    r = _abi.search_sa(s, dict(sa, max_iter=max_iter, seed=4004 + rank), **kw)
    wall = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([wall], dtype=torch.float64, device=dev)
        dist_mod.all_reduce(t, op=dist_mod.ReduceOp.MAX)
        wall = float(t.item())
    return {"workload": "C4: cnSearchSA on cnSamples(alarmnet, 20000), maxParentSet=2, maxComplexity=600, BIC, maxIter=%d, "
                        "numThreads=4 (proposals per step), one chain per GPU" % max_iter,
            "wall_s": wall, "chains": world, "iterations": int(r["niter"]), "orders_evaluated": int(len(r["trace"])),
            "ms_per_order": 1000.0 * wall / max(len(r["trace"]), 1),
            "best_loglik": float(max(x["loglik"] for x in r["nets"])) if r["nets"] else None}


def sa_leg_reference(budget_iter):
    """the reference's SA loop around its own estimate(), threads and cache (oracle/_ref), bounded to budget_iter
    iterations; the first iterations are the expensive ones (empty cache)"""
    from oracle import ref
    s, cats, sa = sa_workload()
    t0 = time.perf_counter()
    r = ref.search_sa(s, 2, 600, sa["start_order"], select_mode=2, temp_start=0.0, temp_cool_fact=1.0,
                      temp_check_orders=1000, max_iter=budget_iter, order_shuffles=0.0, stop_diff=1e-10, num_threads=4,
                      use_cache=True, seed=4004, num_cats=cats)
    wall = time.perf_counter() - t0
    return {"workload": "C4: cnSearchSA on cnSamples(alarmnet, 20000), maxParentSet=2, maxComplexity=600, BIC, maxIter=%d, "
                        "numThreads=4, cache on (reference core, bounded)" % budget_iter,
            "wall_s": wall, "chains": 1, "iterations": int(r["niter"]), "orders_evaluated": int(len(r["trace"])),
            "ms_per_order": 1000.0 * wall / max(len(r["trace"]), 1)}


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
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != max(args.gpus, 1) and world > 1:
        log("warning: WORLD_SIZE %d != --gpus %d" % (world, args.gpus))
    wl = build_workload(args.config, args.samples)
    metric = "family log-lik scores/sec (cnSearchOrder)"
    n_fam_total = sum(math.comb(wl["n"], d + 1) for d in range(1, wl["P"] + 1))
    threads = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return 0
        from catnet_b200 import synth
        if "samples" in wl:
            s = wl["samples"]
        else:
            s = synth.sample_network(wl["net"], wl["m"], np.random.Generator(np.random.PCG64(wl["seed"])), dtype=np.int8)
        s0 = np.ascontiguousarray(s.astype(np.int32) - 1)
        order0 = wl["order"] - 1
        for _ in range(max(args.warmup, 0)):
            cpu_reference_rate(s0, wl["cats"], order0, wl["P"], 0.5, threads)
        rates, t_all = [], 0.0
        per_step = max(2.0, min(20.0, 120.0 / max(args.steps, 1)))
        nfam = 0
        for _ in range(args.steps):
            r, k, wall = cpu_reference_rate(s0, wl["cats"], order0, wl["P"], per_step, threads)
            rates.append(r)
            t_all += wall
            nfam = k
        v = float(np.mean(rates))
        sample = "%d random candidate families with %d parents at full M=%d per step, %d concurrent reference scorers" % (
            nfam, wl["P"], wl["m"], threads)
        sa_ref = None
        if not args.no_sa:
            try:
                sa_ref = sa_leg_reference(30)
            except Exception as e:
                sa_ref = {"unavailable": str(e)}
        print(json.dumps({
            "impl": "reference", "metric": metric, "value": v, "unit": "families/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * t_all / max(args.steps, 1),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["desc"], "families": n_fam_total},
            "cpu_baseline": {"value": v, "unit": "families/s", "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": v, "unit": "families/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "sa": sa_ref}))
        return 0

    import torch
    import torch.distributed as dist
    from catnet_b200 import _abi
    _abi.lib()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the scoring path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- synthetic data: rank 0 draws, everybody receives the same matrix
    t0 = time.time()
    if "samples" in wl:
        samples_dev = torch.as_tensor(wl["samples"], dtype=torch.int32, device=dev)
    elif rank == 0:
        samples_dev = sample_on_device(wl, dev)
    else:
        samples_dev = torch.empty((wl["m"], wl["n"]), dtype=torch.int32, device=dev)
    if world > 1:
        dist.broadcast(samples_dev, 0)
    torch.cuda.synchronize()
    log("rank %d: data %s ready in %.1fs" % (rank, tuple(samples_dev.shape), time.time() - t0))
    m, n = samples_dev.shape
    kw = dict(max_parent_set=wl["P"], max_complexity=wl["K"], order=wl["order"].astype(np.int32))

    ctx = _abi.Context(local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.set_samples_dev(samples_dev.data_ptr(), n, m, num_cats=wl["cats"])
    if args.kernel_bits:
        ctx.force_generic(args.kernel_bits)
    seg, nfam = ctx.plan(rank, world, **kw)
    assert nfam == n_fam_total, (nfam, n_fam_total)
    scores = torch.empty(max(seg * world, 1), dtype=torch.float64, device=dev)

    def step():
        ctx.plan(rank, world, **kw)
        ctx.score_shard(scores.data_ptr())
        if world > 1:
            dist.all_gather_into_tensor(scores, scores[rank * seg:(rank + 1) * seg].clone())
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

    # ---- roofline of the dominant kernel, measured live (the library's CUDA events on the launching stream)
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
        # tcgen05.mma.kind::mxf4 M128 N256 on every SM (the best rate the instruction reaches; the table kernel issues
        # the A-from-tensor-memory form at N <= 192, whose own ceiling is reported as ts_ceiling).  u8 operands
        # (kind::i8, test path): 2 x the measured sustained bf16 figure.
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
        else:
            peak = 2.0 * bf16
            src = ("2 x bf16_tflops_sustained of MEASURED_PEAKS.json (int8 nominal = 2 x bf16)" if peaks
                   else "2 x 1400 TFLOP/s fallback (B200_PROFILING.md sustained bf16)")
        roofline = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "traffic": traffic,
                    "kernel": "k_umma_tables (tcgen05.mma %s)" % ("kind::mxf4.block_scale" if fp4 else "kind::i8"),
                    "kernel_ms": t_ms, "ops_per_launch": ops, "tiles": int(ph_last["tensor_tiles"]),
                    "peak_source": src,
                    "useful_mac_fraction": dom_fam * 27.0 * m / max(ph_last["tensor_macs"], 1),
                    "hbm_model": hbm_model}
        roofline.update(extra)
    else:
        roofline = {"bound": "hbm", "achieved": hbm_model["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s",
                    "frac": hbm_model["frac"], "traffic": traffic,
                    "kernel": "k_score_planes%s<%d,%d>" % ("_ie2" if d_dom == 2 else "", d_dom, min(4, max(2, int(max(wl["cats"]))))),
                    "kernel_ms": dom_ms, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": hbm_model["peak_source"]}
    roofline["family_samples_per_s"] = dom_fam * m / (dom_ms / 1000.0) if dom_ms > 0 else 0.0

    # ---- e2e through the C ABI with host buffers
    e2e = None
    if not args.no_e2e:
        host = torch.empty((m, n), dtype=torch.int32, pin_memory=True)
        host.copy_(samples_dev)
        torch.cuda.synchronize()
        hnp = host.numpy()
        ctx2 = _abi.Context(local_rank)
        ctx2.set_stream(stream.cuda_stream)
        L = _abi.lib()

        wall = {"set_samples_ms": 0.0, "plan_score_ms": 0.0, "finish_ms": 0.0, "h2d_bytes": 0}

        def e2e_step():
            t_a = time.perf_counter()
            if world == 1:
                ctx2.set_samples(hnp, num_cats=wl["cats"])       # H2D + pack + pair tables; returns synchronised
            else:
                # every rank copies 1/world of the host matrix, NVLink all-gather, pack (catnet_b200/dist.py)
                from catnet_b200 import dist as cdist
                _, nb = cdist.set_samples_sharded(ctx2, host, rank, world, dev, num_cats=wl["cats"])
                wall["h2d_bytes"] = nb
            t_b = time.perf_counter()
            ctx2.plan(rank, world, **kw)
            ctx2.score_shard(scores.data_ptr())
            if world > 1:
                dist.all_gather_into_tensor(scores, scores[rank * seg:(rank + 1) * seg].clone())
            t_c = time.perf_counter()
            h = ctx2.finish_raw(scores.data_ptr())                # selection + DP + network export (D2H)
            t_d = time.perf_counter()
            wall["set_samples_ms"] += (t_b - t_a) * 1e3
            wall["plan_score_ms"] += (t_c - t_b) * 1e3
            wall["finish_ms"] += (t_d - t_c) * 1e3
            nn = L.cnb_result_num_networks(h)
            import ctypes
            cx, ll = ctypes.c_int32(), ctypes.c_double()
            L.cnb_result_network(h, nn - 1, ctypes.byref(cx), ctypes.byref(ll))
            L.cnb_result_free(h)
            return nn, cx.value, ll.value

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
        t_ms = torch.tensor([ems], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        ems = float(t_ms.item())
        tm = ctx2.timing()
        e2e = {"value": n_fam_total * e_steps / (ems / 1000.0), "unit": "families/s",
               "h2d_bytes_per_step": (int(tm["h2d_bytes"]) + int(wall["h2d_bytes"])) * world,
               "d2h_bytes_per_step": int(tm["d2h_bytes"]) * world,
               "ms_per_step": ems / e_steps, "steps": e_steps,
               "host_wall_ms_per_step": {k_: wall[k_] / e_steps for k_ in ("set_samples_ms", "plan_score_ms", "finish_ms")},
               "device_ms": {k_: float(tm[k_]) for k_ in ("pack_ms", "score_ms", "select_ms", "dp_ms", "collect_ms")},
               "result": {"networks": res[0], "max_complexity": res[1], "loglik": res[2]}}
        ctx2.close()

    # ---- the reference's CPU path beside it (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            s0 = np.ascontiguousarray((samples_dev - 1).cpu().numpy())
            r, k, wall = cpu_reference_rate(s0, wl["cats"], wl["order"] - 1, wl["P"], args.cpu_budget, threads)
            cpu = {"value": r, "unit": "families/s", "cores": threads, "kind": "reference",
                   "sample": "%d random candidate families with %d parents at full M=%d, %d concurrent reference "
                             "scorers (CATNET::setParents+setNodeSampleProb), %.1fs wall" % (k, wl["P"], m, threads, wall)}
        except Exception as e:  # the checker library is optional for the product path
            cpu = {"value": None, "unit": "families/s", "cores": threads, "kind": "reference", "sample": "unavailable: %s" % e}

    # ---- cnSearchSA wall time (second half of the BASELINE metric): one chain per GPU
    sa_res = None
    if not args.no_sa and args.config == "C5" and args.samples is None:
        try:
            sa_res = sa_leg_ours(rank, world, args.sa_iter, dist if world > 1 else None, dev)
        except Exception as e:
            sa_res = {"unavailable": str(e)}

    if rank == 0:
        ph = phases[-1]
        line = {
            "metric": metric, "value": value, "unit": "families/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": ("e2m1 one-hot operands, f32 accumulators holding exact counts, f64 log-lik"
                      if ph.get("tensor_kind", 0) == 4 else "u32 counts + f64 log-lik"), "data": "synthetic",
            "config": {"workload": wl["desc"], "families": n_fam_total, "families_per_rank": int(ph["num_families"]),
                       "l2": "inputs larger than L2 (bit planes %.0f MB)" % (n * ((m + 31) // 32) * 16 / 1e6),
                       "data_rng": "network numpy PCG64(%d); samples torch CUDA Philox(%d)" % (wl["seed"], wl["seed"])},
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(ph["kernel_launches"]) * args.steps,
            "phases_ms": {k: ph[k] for k in ("plan_ms", "score_ms", "select_ms", "dp_ms")},
            "sa": sa_res,
        }
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
