"""One process, several GPUs (cnb_mctx; catnet_b200/csrc/cnb_multi.cu) -- the multi-GPU path an R session takes.

Every test runs in two forms: "alias" = G device contexts on the ONE GPU of the box (the library's test hook
num_devices = -G: the whole orchestration -- dealt sample rows, peer copies, sharded scoring into the first device's
score buffer, chains, dealt orders -- runs on a single-GPU box, so nothing here is skipped by the driver), and "real" =
G distinct GPUs when the box has them (gpurun --gpus 2/8).  The single-device result is the reference for both; it is
itself compared with the oracle in test_gpu_parity.py.
"""
import ctypes
import os

import numpy as np
import pytest

from helpers import NA, compare_search, constrained_problem, random_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def abi():
    from catnet_b200 import _abi
    _abi.lib()
    return _abi


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _forms(g):
    """num_devices arguments to try for G devices"""
    out = [-g]
    if _ngpu() >= g:
        out.append(g)
    return out


@pytest.fixture(autouse=True)
def shard_everything(monkeypatch):
    monkeypatch.setenv("CNB_MULTI_MIN_WORK", "0")          # the test problems are far below the sharding threshold
    monkeypatch.setenv("CNB_MULTI_SHARE_MIN", "0")         # ... and below the dealt-upload threshold
    yield


def same_nets(a, b):
    assert compare_search(a, b) == len(b)
    for x, y in zip(a, b):
        assert x["loglik"] == y["loglik"] and x["node_loglik"] == y["node_loglik"]
        assert all(np.array_equal(p, q) for p, q in zip(x["probs"], y["probs"]))


@pytest.mark.parametrize("g", [2, 3, 8])
@pytest.mark.parametrize("seed,n,m,cats,na,P", [(6, 12, 500, None, 0.02, 2), (7, 9, 257, 3, 0.0, 3), (8, 14, 1000, 4, 0.0, 2),
                                                (9, 7, 33, 2, 0.1, 1), (10, 10, 300, 6, 0.0, 2)])
def test_one_shot_search_order(abi, g, seed, n, m, cats, na, P):
    """cnb_search_order with cnb_params.num_devices: identical networks, tables and log-likelihoods"""
    s, c = random_problem(seed, n, m, cats, na)
    K = int(np.sum(c - 1)) + n * int(c.max()) ** P
    a = abi.search_order(s, max_parent_set=P, max_complexity=K, num_cats=c)
    for nd in _forms(g):
        b = abi.search_order(s, max_parent_set=P, max_complexity=K, num_cats=c, num_devices=nd)
        same_nets(b, a)


@pytest.mark.parametrize("seed", range(700, 720))
def test_constrained_problems(abi, seed):
    """pools, fixed parents, parent sizes, perturbations, priors, random orders -- sharded over 3 devices"""
    s, P, K, kw = constrained_problem(seed)
    try:
        a = abi.search_order(s, max_parent_set=P, max_complexity=K, **kw)
    except abi.CnbError as e:
        with pytest.raises(abi.CnbError) as e2:
            abi.search_order(s, max_parent_set=P, max_complexity=K, num_devices=-3, **kw)
        assert e2.value.code == e.code
        return
    for nd in _forms(3):
        same_nets(abi.search_order(s, max_parent_set=P, max_complexity=K, num_devices=nd, **kw), a)


def test_environment_variable_selects_devices(abi, monkeypatch):
    """what an unchanged R session does: CNB_NUM_DEVICES in the environment, cnb_params.num_devices = 0"""
    s, c = random_problem(11, 10, 400, None, 0.0)
    a = abi.search_order(s, max_parent_set=2, max_complexity=500, num_cats=c)
    monkeypatch.setenv("CNB_NUM_DEVICES", "-2")
    same_nets(abi.search_order(s, max_parent_set=2, max_complexity=500, num_cats=c), a)
    monkeypatch.setenv("CNB_NUM_DEVICES", "all")           # clamped to the devices present: any box
    same_nets(abi.search_order(s, max_parent_set=2, max_complexity=500, num_cats=c), a)
    abi.lib().cnb_release_cache()


@pytest.mark.parametrize("nopeer", [False, True])
@pytest.mark.parametrize("g", [2, 4])
def test_tensor_path_sharded(abi, monkeypatch, g, nopeer):
    """the tcgen05 table kernel on every device, tiles dealt round-robin, scores stored into the first device's buffer
    (peer stores) or copied there (CNB_NO_PEER)"""
    if nopeer:
        monkeypatch.setenv("CNB_NO_PEER", "1")
    s, c = random_problem(21, 70, 9000, 4, 0.0)
    order = (np.random.default_rng(3).permutation(70) + 1).astype(np.int32)
    kw = dict(max_parent_set=2, max_complexity=70 * 48, order=order)
    with abi.Context(0) as ctx:
        ctx.set_samples(s, num_cats=c)
        a = ctx.search_order(**kw)
        assert ctx.timing()["tensor_tiles"] > 0
    for nd in _forms(g):
        with abi.MultiContext(0, nd) as mc:
            mc.set_samples(s, num_cats=c)
            b = mc.search_order(**kw)
            assert mc.active() == g
            tiles = [mc.timing(i)["tensor_tiles"] for i in range(g)]
            assert min(tiles) > 0 and sum(tiles) == ctx_tiles(abi, s, c, kw)
            same_nets(b, a)
            # a second order on the resident data (the SA / bench shape), then back
            kw2 = dict(kw, order=order[::-1].copy())
            with abi.Context(0) as ctx:
                ctx.set_samples(s, num_cats=c)
                a2 = ctx.search_order(**kw2)
            same_nets(mc.search_order(**kw2), a2)
            same_nets(mc.search_order(**kw), a)


def ctx_tiles(abi, s, c, kw):
    with abi.Context(0) as ctx:
        ctx.set_samples(s, num_cats=c)
        ctx.search_order(**kw)
        return ctx.timing()["tensor_tiles"]


@pytest.mark.parametrize("what", ["plain", "na", "pert", "inferred", "label300"])
def test_dealt_upload(abi, what):
    """every device uploads its rows as bytes and sends them to its peers; perturbations travel the same way; a label
    above 255 (legal with inferred categories) falls back to the int32 upload"""
    s, c = random_problem(31, 9, 1003, 3, 0.05 if what == "na" else 0.0)
    pert = (np.random.default_rng(5).random(s.shape) < 0.2).astype(np.int32) if what == "pert" else None
    cats = None if what in ("inferred", "label300") else c
    if what == "label300":
        s = np.where(s >= 1, s + 299, s).astype(np.int32)   # labels 300..302: three categories, min 300
    kw = dict(max_parent_set=2, max_complexity=400, num_cats=cats, perturbations=pert)
    a = abi.search_order(s, **kw)
    for nd in _forms(3) + _forms(2):
        same_nets(abi.search_order(s, num_devices=nd, **kw), a)


def test_label_above_255_single_device_compacted(abi, monkeypatch):
    """the host-compacted upload saturated labels above 255; it now hands such matrices to the int32 route"""
    monkeypatch.setenv("CNB_HOST_COMPACT_MIN", "0")
    s, c = random_problem(32, 6, 500, 3, 0.0)
    a = abi.search_order(s, max_parent_set=2, max_complexity=200)
    b = abi.search_order(np.where(s >= 1, s + 299, s).astype(np.int32), max_parent_set=2, max_complexity=200)
    same_nets(b, a)


SA = dict(select_mode=2, temp_start=0.05, temp_cool_fact=0.8, temp_check_orders=4, max_iter=40, order_shuffles=1.5,
          stop_diff=1e-10, num_threads=2, use_cache=True)


@pytest.mark.parametrize("g", [2, 3])
def test_sa_chains(abi, g):
    """one chain per device: chain 0 is the single-device chain draw for draw; the merge keeps, per complexity, the
    best network over the chains (so it is never worse than the single-device result)"""
    s, c = random_problem(41, 8, 300, None, 0.0)
    start = (np.random.default_rng(2).permutation(8) + 1).astype(np.int32)
    kw = dict(max_parent_set=2, max_complexity=300, num_cats=c)
    one = abi.search_sa(s, dict(SA, start_order=start, seed=17), want_probs=True, **kw)
    for nd in _forms(g):
        many = abi.search_sa(s, dict(SA, start_order=start, seed=17), want_probs=True, num_devices=nd, **kw)
        assert np.array_equal(many["trace"], one["trace"]) and np.array_equal(many["order"], one["order"])
        assert many["niter"] == one["niter"] and many["naccept"] == one["naccept"]
        a = {n["complexity"]: n for n in one["nets"]}
        b = {n["complexity"]: n for n in many["nets"]}
        assert set(a) <= set(b)
        better = 0
        for cx, net in a.items():
            assert b[cx]["loglik"] >= net["loglik"]
            better += b[cx]["loglik"] > net["loglik"]
        # every merged network is a valid evaluation: re-scoring its parents under its own order reproduces it
        for cx, net in b.items():
            assert len(net["parents"]) == 8 and len(net["order"]) == 8


def test_sa_follows_the_callers_generator(abi):
    """cnb_sa_params.unif: a generator handed over by the caller (R's unif_rand in the shim) is consumed draw for draw
    -- a ctypes splitmix64 with the library's own mapping must reproduce the seeded chain exactly"""
    s, c = random_problem(42, 7, 200, 3, 0.0)
    start = (np.random.default_rng(4).permutation(7) + 1).astype(np.int32)
    kw = dict(max_parent_set=2, max_complexity=200, num_cats=c)
    seeded = abi.search_sa(s, dict(SA, start_order=start, seed=99), **kw)
    state = [99]
    mask = (1 << 64) - 1

    def splitmix(_):
        state[0] = (state[0] + 0x9E3779B97F4A7C15) & mask
        z = state[0]
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & mask
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & mask
        z ^= z >> 31
        return ((z >> 11) + 0.5) * (1.0 / 9007199254740992.0)

    fn = abi.UNIF_FN(splitmix)
    called = abi.search_sa(s, dict(SA, start_order=start, seed=0, unif=fn), **kw)
    assert np.array_equal(called["trace"], seeded["trace"]) and np.array_equal(called["order"], seeded["order"])
    state[0] = 99
    hist_a, it_a = abi.search_hist(s, dict(score=2, weight=1, max_iter=9, num_threads=2, seed=0, unif=fn), **kw)
    hist_b, it_b = abi.search_hist(s, dict(score=2, weight=1, max_iter=9, num_threads=2, seed=99), **kw)
    assert it_a == it_b and np.array_equal(hist_a, hist_b)


@pytest.mark.parametrize("threads,iters", [(1, 7), (2, 9), (3, 10), (5, 12)])
def test_histogram_orders_dealt_out(abi, threads, iters):
    """cnSearchHist: orders evaluated side by side, accumulated in the reference's order -> bit-identical histogram and
    the same number of uniforms consumed"""
    s, c = random_problem(51, 8, 250, None, 0.03)
    kw = dict(max_parent_set=2, max_complexity=300, num_cats=c)
    for weight in (0, 1, 2):
        hp = dict(score=2, weight=weight, max_iter=iters, num_threads=threads, seed=5)
        a, ia = abi.search_hist(s, hp, **kw)
        for nd in _forms(2) + _forms(4):
            b, ib = abi.search_hist(s, hp, num_devices=nd, **kw)
            assert ia == ib and np.array_equal(a, b)


def test_below_threshold_stays_on_one_device(abi, monkeypatch):
    monkeypatch.delenv("CNB_MULTI_MIN_WORK")
    s, c = random_problem(61, 8, 200, None, 0.0)
    with abi.MultiContext(0, -2) as mc:
        mc.set_samples(s, num_cats=c)
        nets = mc.search_order(max_parent_set=2, max_complexity=300)
        assert mc.active() == 1
    same_nets(nets, abi.search_order(s, max_parent_set=2, max_complexity=300, num_cats=c))


def test_device_time_marks(abi):
    s, c = random_problem(62, 40, 9000, 4, 0.0)
    with abi.MultiContext(0, -2) as mc:
        mc.set_samples(s, num_cats=c)
        mc.search_light(max_parent_set=2, max_complexity=2000)
        mc.mark(0)
        for _ in range(3):
            mc.search_light(max_parent_set=2, max_complexity=2000)
        mc.mark(1)
        ms = mc.elapsed_ms()
        assert 0 < ms < 5000
        h = mc.collect_raw()
        assert abi.lib().cnb_result_num_networks(h) > 0
        abi.lib().cnb_result_free(h)


@pytest.mark.parametrize("n,m,cats,g", [(30, 200, 3, 2), (70, 300, 2, 3), (40, 8300, 4, 2), (101, 120, 3, 4)])
def test_three_parent_tiles_dealt_over_devices(abi, monkeypatch, n, m, cats, g):
    """three parents on the tensor cores over several devices: every device contracts a range of parent triples for all
    children of a column tile and stores its scores straight into the first device's buffer (peer memory)"""
    s, c = random_problem(23 * n + m, n, m, cats)
    kw = dict(max_parent_set=3, max_complexity=n * 30, num_cats=c, order=(np.random.default_rng(n + m).permutation(n) + 1))
    monkeypatch.setenv("CNB_FORCE_BITS", "1024")               # bit planes
    a = abi.search_order(s, want_probs=False, **kw)
    assert abi.last_call_timing()["tensor3_tiles"] == 0
    monkeypatch.setenv("CNB_FORCE_BITS", "4")                  # tensor cores whatever the sample count
    one = abi.search_order(s, want_probs=False, **kw)
    whole = abi.last_call_timing()["tensor3_tiles"]
    assert whole > 0 and compare_search(one, a, check_probs=False) == len(a)
    for nd in _forms(g):
        b = abi.search_order(s, want_probs=False, num_devices=nd, **kw)
        t = abi.last_call_timing()["tensor3_tiles"]
        assert 0 < t < whole                                   # the first device's share
        assert compare_search(b, a, check_probs=False) == len(a)
    abi.lib().cnb_release_cache()


# ------------------------------------------------------------------ streamed search over several devices
@pytest.mark.parametrize("n,m,cats,slices,g", [(70, 9000, 4, "1,3,4,4,3,1", 2), (40, 8192, 3, "1", 3), (130, 20001, 4, "1,1", 2),
                                               (65, 12345, 2, "1,2,3", 4), (30, 100000, 4, "1,5,10", 8)])
def test_streamed_search_over_devices(abi, monkeypatch, n, m, cats, slices, g):
    """cnb_search_order(num_devices = G) on a large complete matrix: every slice of the samples is dealt over the devices
    (1/G of its rows over each PCIe link, peer copies complete it everywhere) and contracted on every device -- over that
    device's tiles -- while the next slice is on its way.  Same networks, tables and log-likelihoods as one device."""
    s, c = random_problem(17 * n + m, n, m, cats)
    kw = dict(max_parent_set=2, max_complexity=n * 16 * 3, num_cats=c, order=(np.random.default_rng(m).permutation(n) + 1))
    monkeypatch.setenv("CNB_STREAM_MIN", "off")
    a = abi.search_order(s, **kw)
    assert abi.last_call_timing()["fast_path"] != 2
    monkeypatch.setenv("CNB_STREAM_SLICES", slices)
    for nd in _forms(g):
        monkeypatch.setenv("CNB_STREAM_MIN", "off")
        same_nets(abi.search_order(s, num_devices=nd, **kw), a)
        assert abi.last_call_timing()["fast_path"] != 2
        monkeypatch.setenv("CNB_STREAM_MIN", "0")
        b = abi.search_order(s, num_devices=nd, **kw)
        t = abi.last_call_timing()
        assert t["fast_path"] == 2 and t["tensor_tiles"] > 0 and t["h2d_bytes"] >= s.size
        same_nets(b, a)
    abi.lib().cnb_release_cache()


def test_streamed_search_over_devices_steps_aside(abi, monkeypatch):
    """a missing value found while packing the last slice: every device packs the byte matrix it has assembled and the
    resident kernels score it; a value outside its node's categories is reported; a label above 255 goes the int32 way"""
    monkeypatch.setenv("CNB_STREAM_MIN", "0")
    s, c = random_problem(78, 40, 9000, 4)
    kw = dict(max_parent_set=2, max_complexity=40 * 48, num_cats=c)
    s_na = s.copy()
    s_na[8000, 7] = NA
    monkeypatch.setenv("CNB_STREAM_MIN", "off")
    want = abi.search_order(s_na, **kw)
    ok = abi.search_order(s, **kw)
    monkeypatch.setenv("CNB_STREAM_MIN", "0")
    for nd in _forms(2) + _forms(3):
        got = abi.search_order(s_na, num_devices=nd, **kw)
        assert abi.last_call_timing()["fast_path"] != 2
        same_nets(got, want)
        bad = s.copy()
        bad[8100, 3] = 9
        with pytest.raises(abi.CnbError) as e:
            abi.search_order(bad, num_devices=nd, **kw)
        assert "Incorrect categories" in str(e.value)
        big = s.copy()
        big[5, 2] = 300
        with pytest.raises(abi.CnbError) as e:
            abi.search_order(big, num_devices=nd, **kw)
        assert "Incorrect categories" in str(e.value)
        again = abi.search_order(s, num_devices=nd, **kw)     # the parked contexts stream again afterwards
        assert abi.last_call_timing()["fast_path"] == 2
        same_nets(again, ok)
    abi.lib().cnb_release_cache()


def test_streamed_search_one_rank_python_driver(abi, monkeypatch):
    """catnet_b200/dist.py::search_order_streamed with a world of one: the slice loop of the multi-process driver
    (upload_rows8_on on a side stream, event per slice, stream_slice, stream_end) against the resident path"""
    import torch
    from catnet_b200 import dist as cdist
    dev = torch.device("cuda:0")
    s, c = random_problem(79, 60, 20000, 4)
    kw = dict(max_parent_set=2, max_complexity=60 * 48, order=(np.random.default_rng(2).permutation(60) + 1))
    with abi.Context(0) as ctx:
        ctx.set_samples(s, num_cats=c)
        want = ctx.search_order(**kw)
    monkeypatch.setenv("CNB_STREAM_SLICES", "1,2,2,1")
    with abi.Context(0) as ctx:
        got, info = cdist.search_order_streamed(ctx, s, 0, 1, dev, c, **kw)
        assert info["streamed"] and info["slices"] == 4 and info["h2d_bytes"] == s.size
        same_nets(got, want)
        s_na = s.copy()
        s_na[19990, 3] = NA
        got, info = cdist.search_order_streamed(ctx, s_na, 0, 1, dev, c, **kw)
        assert not info["streamed"]
    with abi.Context(0) as ctx:
        ctx.set_samples(s_na, num_cats=c)
        same_nets(got, ctx.search_order(**kw))


def _nccl_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["CNB_STREAM_SLICES"] = "1,3,4,4,3,1"
    torch.cuda.set_device(rank)
    from catnet_b200 import dist as cdist0
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank), pg_options=cdist0.nccl_options())
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from helpers import random_problem as rp
        from catnet_b200 import _abi, dist as cdist
        dev = torch.device("cuda", rank)
        res = []
        for seed, n, m, na in ((5, 70, 33000, False), (6, 45, 9001, False), (7, 50, 12000, True)):
            s, c = rp(seed, n, m, 4)
            if na:
                s[m - 5, 3] = -2147483648
            kw = dict(max_parent_set=2, max_complexity=n * 48, order=(np.random.default_rng(seed).permutation(n) + 1))
            with _abi.Context(rank) as ctx:
                ctx.set_samples(s, num_cats=c)
                want = ctx.search_order(**kw)
            with _abi.Context(rank) as ctx:
                got, info = cdist.search_order_streamed(ctx, s, rank, world, dev, c, **kw)
            same = len(got) == len(want) and all(
                x["loglik"] == y["loglik"] and x["complexity"] == y["complexity"]
                and all(np.array_equal(p, q_) for p, q_ in zip(x["parents"], y["parents"]))
                and all(np.array_equal(p, q_) for p, q_ in zip(x["probs"], y["probs"])) for x, y in zip(got, want))
            res.append((bool(same), bool(info["streamed"]), int(info["h2d_bytes"]), s.size))
        # three parents on the tensor cores, the T tiles dealt to the ranks: scores scattered over the segments, max-reduced
        s, c = rp(9, 40, 400, 3)
        kw = dict(max_parent_set=3, max_complexity=40 * 30, order=(np.random.default_rng(9).permutation(40) + 1))
        with _abi.Context(rank) as ctx:
            ctx.set_samples(s, num_cats=c)
            ctx.force_generic(1024)
            want = ctx.search_order(want_probs=False, **kw)
            ctx.force_generic(4)
            got = cdist.search_order_sharded(ctx, rank, world, dev, want_probs=False, **kw)
            scat, t3 = ctx.scattered(), ctx.timing()["tensor3_tiles"]
        same = len(got) == len(want) and all(x["loglik"] == y["loglik"] and all(np.array_equal(p, q_) for p, q_ in zip(x["parents"], y["parents"]))
                                             for x, y in zip(got, want))
        res.append((bool(same), bool(scat), int(t3), 0))
        q.put((rank, res))
    except Exception as e:                                     # the parent fails at once instead of waiting for the queue
        q.put((rank, "error: %r" % (e,)))
        raise
    finally:
        dist.destroy_process_group()


def test_streamed_search_two_ranks_nccl():
    """the multi-process driver on two real GPUs: slices all-gathered in place over NCCL while the table kernel runs"""
    if _ngpu() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import socket
    import torch.multiprocessing as mp
    sk = socket.socket()
    sk.bind(("127.0.0.1", 0))
    port = sk.getsockname()[1]
    sk.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank in (0, 1):
        assert not isinstance(out[rank], str), out[rank]
        (s1, st1, h1, sz1), (s2, st2, h2, sz2), (s3, st3, _, _), (s4, scat, t3, _) = out[rank]
        assert s1 and s2 and s3 and st1 and st2 and not st3
        assert s4 and scat and t3 > 0
        assert abs(h1 - sz1 / 2) <= 70 * 8 and abs(h2 - sz2 / 2) <= 45 * 8
